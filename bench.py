"""Benchmark of the B200-native ORT captioning hot path (driver contract: one JSON line on stdout).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--images B]

Workload (BASELINE.json configs[2]): ORT 6x512, 95 % randomly pruned binarized-mask weights, beam-3 incremental
decoding, max length 16, 36 regions x 2048-d synthetic features + boxes, 512 images per GPU per step.
A "step" = one batch through encoder + 16 decode steps + beam bookkeeping.  N > 1: one process per GPU (torchrun),
images sharded by rank, no collective on the data path (weak scaling).

  value : captions/s (one caption = the best beam of one image), inputs already resident in HBM.
  e2e   : same through OrtEngine.sample() with pinned HOST fp32 features/boxes (H2D inside the timed region) and a
          device->host read of the decoded tokens + log-probs every step.
  roofline     : dominant kernel (the tcgen05 GEMM), algorithmic FLOPs / CUDA-event time of its launches in one
                 instrumented (graph-less) step, against MEASURED_PEAKS.json.
  cpu_baseline : the CPU oracle (a port of the reference's PyTorch path) on this box's host cores, bounded sample.
  --impl reference : the same CPU arm alone (the reference is pure Python/PyTorch; its hot path restated in
                 oracle/ort_oracle.py is what runs — /root/reference does not exist on the GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(d_model=512, dim_feedforward=2048, num_layers=6, num_heads=8, max_seq_length=16, att_feat_size=2048,
           vocab_size=10000)
SPARSITY = 0.95
BEAM = 3
N_BOX = 36
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant GEMM shape (M=1536, N=512, K=512: 384 of the
# 623 GEMM launches of a step), ncu --set full, profiles/r01b_ncu_full_summary.txt (inf_gemm64: 5.29 MB read, 0 written:
# the 3 MB output tile stays in the 126 MB L2 for the next kernel)
NCU_TRAFFIC_DOMINANT_GEMM = 5293312
METRIC = "ort95_beam3_captions_per_sec"
UNIT = "captions/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons}


def cpu_arm(steps, warmup, images):
    """Reference arm / cpu_baseline: the oracle port of the reference's PyTorch path on the host cores."""
    from oracle import ort_oracle as O
    from sparse_caption_b200.engine import ModelCfg
    from sparse_caption_b200 import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = ModelCfg(CFG)
    sd = synthetic.random_state_dict(cfg, seed=1234, sparsity=SPARSITY)
    ocfg = O.Cfg(**CFG)
    att, boxes = synthetic.synthetic_inputs(images, N_BOX, CFG["att_feat_size"], seed=8888)
    opt = {"beam_size": BEAM}
    with torch.no_grad():
        for _ in range(warmup):
            O.sample(sd, ocfg, att[:8], boxes[:8], None, opt)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.sample(sd, ocfg, att, boxes, None, opt)
        dt = time.perf_counter() - t0
    return images * steps / dt, dt / steps, cores


def time_train_gemms(prof, dev):
    """Every distinct GEMM launch of one training step (forward with its dropout / residual epilogue, dX, dX with the
    activation-mask epilogue, weight gradient incl. its split-K reduction) re-issued through the same entry point, 20 launches
    inside one CUDA graph, CUDA events around two replays: microseconds without the host's launch gaps.  Returns
    (total us per step, total flop per step)."""
    from sparse_caption_b200 import kernels as KK
    groups = {}
    for name, meta, a, b in prof:
        if meta and meta[0] == "gemm_bf16":
            key = (name,) + tuple(meta[1:])
            groups[key] = groups.get(key, 0) + 1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot_us, tot_fl = 0.0, 0.0
    for key, cnt in groups.items():
        name, d1, d2, d3 = key[:4]
        if name == "sc_linear_wgrad_rowmajor":
            N, Kd, M = d1, d2, d3
            dy = torch.randn(M, N, device=dev).bfloat16(); x = torch.randn(M, Kd, device=dev).bfloat16()
            W = torch.randn(N, Kd, device=dev); S = torch.randn(N, Kd, device=dev)
            gW, gS = torch.empty_like(W), torch.empty_like(S)
            wsb = torch.empty(4 * N * Kd, device=dev)
            run = lambda i: KK.linear_wgrad_rowmajor(dy, x, W, S, KK.MASK_BERNOULLI, gW, gS, workspace=wsb, seed=3, stream_id=5)
            fl = 2.0 * M * N * Kd
        else:
            M, N, Kd = d1, d2, d3
            ys = key[6]
            x = torch.randn(M, Kd, device=dev).bfloat16(); w = torch.randn(N, Kd, device=dev).bfloat16()
            outs = [torch.empty(M, N, device=dev, dtype=torch.bfloat16 if ys == 2 else torch.float32) for _ in range(2)]
            fl = 2.0 * M * N * Kd
            if name == "sc_linear_hmask":
                h = torch.randn(M, N, device=dev).clamp_min(0).bfloat16(); cs = torch.zeros(N, device=dev)
                run = lambda i: KK.linear_hmask(x, w, h, outs[i % 2], scale=1.1, colsum=cs)
            elif name == "sc_linear_dropout":
                has_res, relu, drop, tile = key[8], key[9], key[10], key[11]
                bias = torch.randn(N, device=dev); res = torch.randn(M, N, device=dev) if has_res else None
                run = lambda i: KK.linear_dropout(x, w, bias, residual=res, relu=relu, out=outs[i % 2], p=0.1 if drop else 0.0,
                                                  drop_seed=1, drop_stream=2, tile_n=tile)
            else:
                has_res = key[8] if len(key) > 8 else False
                res = torch.randn(M, N, device=dev) if has_res else None
                run = lambda i: KK.linear(x, w, None, residual=res, out=outs[i % 2])
        run(0)
        torch.cuda.synchronize(dev)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for i in range(20):
                run(i)
        gr.replay()
        torch.cuda.synchronize(dev)
        e0.record(); gr.replay(); gr.replay(); e1.record()
        torch.cuda.synchronize(dev)
        tot_us += e0.elapsed_time(e1) * 1e3 / 40 * cnt
        tot_fl += fl * cnt
    return tot_us, tot_fl


def train_arm(args, dev, world, rank, dist_mod=None):
    """SMP training arm (BASELINE.json configs[1]): ORT supermask training, bf16 tensor-core GEMMs with fp32 master
    weights + fp32 mask logits, 5 captions/image with the encoder run once, Bernoulli masks + dropout + sparsity
    loss + clip + Adam inside the timed step.  Single-GPU here; the N-GPU variant adds the NCCL all-reduce of the
    flat weight+logit gradient buffers (tests/test_ddp_cpu.py covers the sharding logic)."""
    from sparse_caption_b200 import lib, synthetic
    from sparse_caption_b200.engine import ModelCfg
    from sparse_caption_b200.trainer import OrtTrainer
    cfg = ModelCfg(dict(CFG, max_seq_length=17))
    from sparse_caption_b200 import distributed as D
    sd = synthetic.random_state_dict(cfg, seed=1234, sparsity=0.0, device=dev)
    tr = OrtTrainer(sd, cfg, mask_type="supermask", precision="bf16", device=dev, seed=8888,  # same mask seed on every rank
                    use_graph=not args.no_train_graph, fused_st=not (world > 1 and args.exchange == "sharded"))
    B, S, T = args.train_images, 5, 17
    g = torch.Generator().manual_seed(8888 + rank)
    att, boxes = synthetic.synthetic_inputs(B, N_BOX, CFG["att_feat_size"], seed=8888 + rank, pin=True)
    R = B * S
    seqs = torch.zeros(R, T + 1, dtype=torch.long)
    masks = torch.zeros(R, T + 1)
    lens = torch.randint(6, T - 1, (R,), generator=g)
    for r in range(R):
        n = int(lens[r])
        seqs[r, 0] = 2
        seqs[r, 1:1 + n] = torch.randint(4, CFG["vocab_size"], (n,), generator=g)
        seqs[r, 1 + n] = 3
        masks[r, :n + 2] = 1
    seqs, masks = seqs.pin_memory(), masks.pin_memory()
    opt = dict(lr=3e-4, sparsity_target=0.95, sparsity_weight=30.0, current_step=100, max_step=1000)
    all_reduce = exchange = None
    if world > 1 and args.exchange == "sharded":
        exchange = D.ShardedExchange(device=dev)  # reduce-scatter -> Adam on the owned shard -> all-gather, per bucket
    else:
        all_reduce = D.make_all_reduce(async_op=True)  # NCCL SUM of the gradient buckets each backward phase finishes
    gtok = D.global_token_count(masks.to(dev), T)

    def step():
        return tr.train_step(att, boxes, seqs, masks, seq_per_img=S, all_reduce=all_reduce, exchange=exchange, global_tokens=gtok, **opt)

    def barrier():
        if dist_mod is not None:
            dist_mod.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(3):
        step()
    barrier()
    before = lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.train_steps):
        loss = step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if dist_mod is not None:
        dist_mod.all_reduce(t, op=dist_mod.ReduceOp.MAX)
    ms = float(t[0]) / args.train_steps
    launches = (lib.launch_count - before) // args.train_steps
    if rank != 0:
        return None
    # GEMM share / tensor roofline from one instrumented step
    lib.profile = []
    tr.use_graph = False  # per-launch event timing needs the eager launch sequence
    tr.train_step(att, boxes, seqs, masks, seq_per_img=S, all_reduce=None, global_tokens=gtok, **opt)  # rank-local
    torch.cuda.synchronize(dev)
    prof, lib.profile = lib.profile, None
    gemm_ms = sum(a.elapsed_time(b) for n, m, a, b in prof if m and m[0] == "gemm_bf16")
    gemm_fl = sum(2.0 * m[1] * m[2] * m[3] for n, m, a, b in prof if m and m[0] == "gemm_bf16")
    tot_ms = sum(a.elapsed_time(b) for n, m, a, b in prof)
    peaks = load_peaks()
    tf_eager = gemm_fl / (gemm_ms / 1e3) / 1e12 if gemm_ms else 0.0
    try:
        g_us, g_fl = time_train_gemms(prof, dev)
        tf = g_fl / g_us / 1e6
    except Exception as ex:  # diagnostics only: fall back to the event timing of the eager step
        print(f"in-graph training GEMM timing skipped: {ex}", file=sys.stderr)
        g_us, tf = None, tf_eager
    return {"metric": "smp_train_images_per_sec", "value": world * B / (ms / 1e3), "unit": "images/s", "n_gpus": world, "ms_per_step": ms,
            "collective": "none (1 GPU)" if world == 1 else (
                "per gradient bucket (4 backward phases): NCCL reduce-scatter -> Adam on the owned 1/N shard -> all-gather of the updated "
                "weights + mask logits, on a side stream under the next phase" if exchange is not None else
                "NCCL all-reduce(sum) of fp32 weight+logit gradient buckets, started after each of the 4 backward phases (overlaps the next phase)"),
            "images_per_gpu_per_step": B, "captions_per_image": S, "positions": T, "dtype": "bf16 GEMM / fp32 master+logits",
            "loss": float(loss), "gpu_launches_per_step": launches, "h2d_bytes_per_step": att.numel() * 4 + boxes.numel() * 4 + seqs.numel() * 8 + masks.numel() * 4,
            "includes": "H2D of the batch, Bernoulli masks, dropout, sparsity loss, clip + Adam (2 groups)",
            "algorithmic_gflop_per_step": gemm_fl / 1e9,
            "roofline": {"kernel": "sc_gemm_bf16_kernel (fwd + dgrad + wgrad launches of one step, each through its own entry point)",
                         "bound": "tensor", "achieved": tf, "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": tf / peaks["tf_burst"],
                         "gemm_us_per_step_in_graph": g_us, "achieved_eager_events": tf_eager,
                         "note": "in-graph timing per shape (20 launches, CUDA events); wgrad figures include the split-K reduction kernel",
                         "share_of_step": gemm_ms / tot_ms if tot_ms else None}}


def _claim_stdout():
    """Route everything that libraries print on fd 1 (e.g. NCCL's version banner) to stderr; the returned file object
    is the real stdout, used for the ONE JSON line of the driver contract."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    return real


def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--images", type=int, default=512, help="images per GPU per step")
    ap.add_argument("--backend", default="dense", choices=["dense", "csr", "sell", "auto"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the SMP training arm (BASELINE.json configs[1])")
    ap.add_argument("--train-images", type=int, default=50, help="images per GPU per training step (5 captions each)")
    ap.add_argument("--train-steps", type=int, default=20)
    ap.add_argument("--exchange", default="allreduce", choices=["sharded", "allreduce"],
                    help="data-parallel gradient exchange of the training arm (2 GPUs measured: allreduce 6.05, sharded 6.18 ms/step)")
    ap.add_argument("--no-train-graph", action="store_true", help="diagnostic: eager launches instead of the captured training graph")
    ap.add_argument("--cpu-images", type=int, default=64)
    ap.add_argument("--ln-fold", action="store_true", help="LayerNorm folded into the consuming GEMMs (sc_linear_ln) instead of separate LayerNorm kernels")
    ap.add_argument("--prefetch", action="store_true", help="e2e arm: H2D of the next batch on a copy stream underneath the previous decode of the same slot")
    ap.add_argument("--no-fuse-topk", action="store_true", help="diagnostic: materialise the logits (sc_linear + sc_beam_step) instead of the fused generator + beam row pass")
    ap.add_argument("--no-pdl", action="store_true", help="diagnostic: disable programmatic dependent launch")
    ap.add_argument("--slots", type=int, default=0, help="batches in flight (pipeline slots: stream + workspaces + graphs each); 0 = 8 when the timed region is long enough "
                         "to amortise the pipeline's fill and drain (>= 24 steps), else 4")
    args = ap.parse_args()
    if args.slots <= 0:
        # the K timed steps go round-robin over the slots; a slot count that divides K keeps every round of the pipeline
        # full (K = 10 -> 5 slots: two full rounds instead of 4 + 4 + 2), long runs amortise fill / drain anyway
        args.slots = 8 if args.steps >= 24 else next((sl for sl in (8, 7, 6, 5, 4) if args.steps % sl == 0), 4)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": "ORT 6x512 95%-sparse binarized-mask beam-3 inference, L=16, V=10000, 36x2048 features + boxes "
                          f"({args.images} images/GPU/step) [BASELINE.json configs[2]]",
              "images_per_gpu_per_step": args.images, "beam": BEAM, "max_len": 16, "sparsity": SPARSITY,
              "sharding": f"images by rank x{world}, no data-path collective", "decoder_gemm_backend": args.backend,
              "batches_in_flight": args.slots,
              "l2_policy": "inputs_larger_than_L2 (151 MB fresh fp32 features + ~0.7 GB of activations/KV per step vs 126 MB L2)"}

    if args.impl == "reference":
        if rank != 0:
            return
        # K timed steps and W warm-up steps as asked; one step = a bounded sample (--cpu-images, default 64 images of the
        # 512-image workload: ~0.6 s on 16 cores), so the default K=10 / W=3 run takes ~10 s
        warm = max(1, args.warmup)
        steps = max(1, args.steps)
        v, spp, cores = cpu_arm(steps, warm, args.cpu_images)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
                          "warmup": warm, "ms_per_step": spp * 1e3, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                           "sample": f"{steps} x {args.cpu_images} images, same model/beam/length, torch fp32 on host cores"},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), file=out, flush=True)
        return

    assert torch.cuda.is_available(), "bench.py --impl b200 needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from sparse_caption_b200 import lib, synthetic
    from sparse_caption_b200.engine import ModelCfg, OrtEngine
    lib.load()
    if args.no_pdl:
        from sparse_caption_b200 import kernels as _K
        _K.set_pdl(False)
    cfg = ModelCfg(CFG)
    sd = synthetic.random_state_dict(cfg, seed=1234, sparsity=SPARSITY, device=dev)
    eng = OrtEngine(sd, cfg, precision="bf16", sparse_backend=args.backend, device=dev, ln_fold=args.ln_fold,
                    fuse_topk=not args.no_fuse_topk,
                    dec_tiles={"o": 3256, "co": 3256, "cq": 3128, "ff2": 3256} if args.slots >= 8 else None)
    B = args.images
    # two distinct pinned host batches, alternated
    host = [synthetic.synthetic_inputs(B, N_BOX, CFG["att_feat_size"], seed=8888 + rank + 100 * i, pin=True) for i in range(2)]
    opt = {"beam_size": BEAM}

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- device-resident arm: inputs already in HBM ----------------
    # `slots` batches are in flight at once: slot s = its own stream, workspaces and CUDA graphs (engine.submit);
    # every timed step is still one full batch (encoder + 16 decode steps + beam bookkeeping).
    S = max(1, args.slots)
    slots = list(range(1, S + 1))
    cur = torch.cuda.current_stream(dev)
    encs = {}
    for s in slots:
        # inputs resident in HBM (fp32, as the data loader delivers them) -> the slot's workspaces + graphs
        eng.submit(host[s & 1][0].to(dev), host[s & 1][1].to(dev), None, opt, slot=s)
        encs[s] = eng._get_enc_ws(B, N_BOX, False, s)
    eng.wait(host=True)
    torch.cuda.synchronize(dev)
    dws = eng._get_dec_ws(B, BEAM, N_BOX, False, slots[0])
    launches_per_step = encs[slots[0]].launches + dws.launches

    def dev_step(i):
        s = slots[i % S]
        with torch.cuda.stream(eng.stream(s)):
            eng.run_encoder(encs[s])
            eng.decode(encs[s], opt)

    def fork():
        for s in slots:
            eng.stream(s).wait_stream(cur)

    def join():
        for s in slots:
            cur.wait_stream(eng.stream(s))

    fork()
    for i in range(args.warmup):
        dev_step(i)
    join()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:  # one nvidia-smi poller per job, not per rank
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fork()
    for i in range(args.steps):
        dev_step(i)
    join()
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1)

    # ---------------- end-to-end arm: pinned host inputs -> tokens on the host ----------------
    out_seq = [torch.empty(B, BEAM, 16, dtype=torch.int32).pin_memory() for _ in slots]
    out_lp = [torch.empty(B, BEAM, 16, dtype=torch.float32).pin_memory() for _ in slots]

    def e2e_step(i):
        att, boxes = host[i & 1]
        k = i % S
        eng.submit(att, boxes, None, opt, slot=slots[k], out=(out_seq[k], out_lp[k]), prefetch=args.prefetch)

    for i in range(max(args.warmup, S)):  # every slot's host-input path (its own workspaces / graphs with the fused ingest) is warm
        e2e_step(i)
    eng.wait()
    barrier()
    e0.record()
    for i in range(args.steps):
        e2e_step(i)
    eng.wait()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    sampler.stop_flag = True
    if rank == 0:
        sampler.join(timeout=2)
    h2d = host[0][0].numel() * 4 + host[0][1].numel() * 4
    d2h = out_seq[0].numel() * 4 + out_lp[0].numel() * 4

    t = torch.tensor([ms_dev, ms_e2e], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(t[0]), float(t[1])
    value = world * B * args.steps / (ms_dev / 1e3)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)

    train = None
    if not args.no_train:
        train = train_arm(args, dev, world, rank, dist_mod=dist)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---------------- roofline leg: one instrumented step without graphs ----------------
    peaks = load_peaks()
    eng2 = OrtEngine(sd, cfg, precision="bf16", sparse_backend=args.backend, device=dev, use_graphs=False,
                     ln_fold=args.ln_fold, fuse_topk=not args.no_fuse_topk)
    enc2 = eng2.encode(host[0][0], host[0][1])
    eng2.decode(enc2, opt)
    torch.cuda.synchronize(dev)
    lib.profile = []
    eng2.run_encoder(enc2)
    eng2.decode(enc2, opt)
    torch.cuda.synchronize(dev)
    prof, lib.profile = lib.profile, None
    by = {}
    for name, meta, a, b in prof:
        key = meta[0] if meta else name
        d = by.setdefault(key, {"ms": 0.0, "n": 0, "flops": 0.0, "bytes": 0.0})
        d["ms"] += a.elapsed_time(b)
        d["n"] += 1
        if meta and meta[0].startswith("gemm"):
            _, M, N, Kd, xs, wsz, ys = meta[:7]
            d["flops"] += 2.0 * M * N * Kd
            d["bytes"] += M * Kd * xs + N * Kd * wsz + M * N * ys
        elif meta and meta[0] in ("csr_spmm", "sell_spmm"):
            _, M, N, Kd, xs, nnz, ys = meta
            d["flops"] += 2.0 * M * nnz
            d["bytes"] += M * Kd * xs + nnz * (xs + 2) + M * N * ys
    total_ms = sum(d["ms"] for d in by.values())
    top = max(by, key=lambda k: by[k]["ms"])
    g = by.get("gemm_bf16", by[top])
    # The per-launch events above include the host's launch gaps (eager ctypes launches): they give the kernel's SHARE of
    # the step.  The duration behind `achieved` is measured without them: every GEMM shape of the step, 40 launches inside
    # one CUDA graph on the current stream, CUDA events around two replays.
    shapes = {}
    for name, meta, a, b in prof:
        if meta and meta[0] == "gemm_bf16":
            key = tuple(meta[1:7]) + tuple(meta[8:10])
            shapes[key] = shapes.get(key, 0) + 1
    from sparse_caption_b200 import kernels as KK
    gemm_us, gemm_fl, dom = 0.0, 0.0, None
    for (M, N, Kd, xs, wsz, ys, has_res, relu), cnt in shapes.items():
        x = torch.randn(M, Kd, device=dev).bfloat16()
        w = torch.randn(N, Kd, device=dev).bfloat16()
        bias = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev) if has_res else None
        if ys == 0:  # generator fused with the beam row pass: no output tile, 12-float records per (row, tile half)
            part = torch.empty(M, KK.linear_topk_parts(N), 12, device=dev)
            run = lambda i: KK.linear_topk(x, w, bias, part, candidates=BEAM)
        else:
            outs = [torch.empty(M, N, device=dev, dtype=torch.bfloat16 if ys == 2 else torch.float32) for _ in range(4)]
            run = lambda i: KK.linear(x, w, bias, residual=res, relu=relu, out=outs[i % 4])
        run(0)
        torch.cuda.synchronize(dev)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for i in range(40):
                run(i)
        gr.replay()
        torch.cuda.synchronize(dev)
        e0.record(); gr.replay(); gr.replay(); e1.record()
        torch.cuda.synchronize(dev)
        us = e0.elapsed_time(e1) * 1e3 / 80
        gemm_us += us * cnt
        gemm_fl += 2.0 * M * N * Kd * cnt
        if dom is None or us * cnt > dom[0]:
            dom = (us * cnt, M, N, Kd, us, cnt, xs * M * Kd + wsz * N * Kd + ys * M * N + (4 * M * N if has_res else 0))
    # the same shapes with `slots` streams running them concurrently - the regime of the timed region (S batches in flight):
    # microseconds of wall time per GEMM = elapsed / (streams * launches)
    conc_us, conc_fl = 0.0, 0.0
    try:
        streams = [torch.cuda.Stream(dev) for _ in range(S)]
        cur_s = torch.cuda.current_stream(dev)
        for (M, N, Kd, xs, wsz, ys, has_res, relu), cnt in shapes.items():
            graphs = []
            for st in streams:
                x = torch.randn(M, Kd, device=dev).bfloat16()
                w = torch.randn(N, Kd, device=dev).bfloat16()
                bias = torch.randn(N, device=dev)
                res = torch.randn(M, N, device=dev) if has_res else None
                if ys == 0:
                    part = torch.empty(M, KK.linear_topk_parts(N), 12, device=dev)
                    run = (lambda x, w, bias, part: (lambda i: KK.linear_topk(x, w, bias, part, candidates=BEAM)))(x, w, bias, part)
                else:
                    outs = [torch.empty(M, N, device=dev, dtype=torch.bfloat16 if ys == 2 else torch.float32) for _ in range(2)]
                    run = (lambda x, w, bias, res, outs: (lambda i: KK.linear(x, w, bias, residual=res, relu=relu, out=outs[i % 2])))(x, w, bias, res, outs)
                run(0)
                torch.cuda.synchronize(dev)
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    for i in range(20):
                        run(i)
                graphs.append((gr, run))

            def go():
                for st, (gr, _) in zip(streams, graphs):
                    st.wait_stream(cur_s)
                    with torch.cuda.stream(st):
                        gr.replay()
                        gr.replay()
                for st in streams:
                    cur_s.wait_stream(st)
            go()
            torch.cuda.synchronize(dev)
            e0.record(); go(); e1.record()
            torch.cuda.synchronize(dev)
            us = e0.elapsed_time(e1) * 1e3 / (S * 40)
            conc_us += us * cnt
            conc_fl += 2.0 * M * N * Kd * cnt
            del graphs
    except Exception as ex:  # diagnostics only
        print(f"concurrent GEMM timing skipped: {ex}", file=sys.stderr)
    tf_conc = conc_fl / conc_us / 1e6 if conc_us else None
    tf = gemm_fl / gemm_us / 1e6 if gemm_us else 0.0
    tf_dom = 2.0 * dom[1] * dom[2] * dom[3] / dom[4] / 1e6 if dom else 0.0
    roofline = {"kernel": "sc_gemm_bf16_kernel (tcgen05/TMEM/TMA; every GEMM launch of one step)", "bound": "tensor", "achieved": tf,
                "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": tf / peaks["tf_burst"],
                # dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant shape, ncu --set full
                # (profiles/r01b_ncu_full_summary.txt); cold-cache capture, one launch
                "traffic": NCU_TRAFFIC_DOMINANT_GEMM,
                "peak_source": f"{peaks['src']} MEASURED_PEAKS.json bf16_tflops (burst: shapes timed alone, in-graph)",
                "achieved_in_flight": tf_conc, "frac_in_flight": (tf_conc / peaks["tf_sus"]) if tf_conc else None,
                "in_flight_note": f"same shapes, {S} streams concurrently (the timed region's regime), wall time per GEMM; vs sustained peak",
                "launches": g["n"], "avg_launch_us": gemm_us / max(1, g["n"]),
                "algorithmic_gflop_per_step": gemm_fl / 1e9,
                "dominant_shape": None if dom is None else {"M": dom[1], "N": dom[2], "K": dom[3], "launches": dom[5], "us_per_launch": dom[4],
                                                            "tflops": tf_dom, "algorithmic_bytes": dom[6],
                                                            "hbm_floor_us": dom[6] / (peaks["hbm"] * 1e3)},
                "share_of_step": g["ms"] / total_ms if total_ms else None,
                "share_source": "per-launch CUDA events of one un-graphed step (host launch gaps included); the ncu launch list "
                                "profiles/r01b_infer_launches_summary.txt gives the same share",
                "kernel_time_breakdown_ms": {k: round(v["ms"], 3) for k, v in sorted(by.items(), key=lambda kv: -kv[1]["ms"])}}

    cpu = None
    if not args.no_cpu_baseline and world == 1:  # (rank 0 at N = 1 only; at N > 1 the driver reads it from the N = 1 line)
        # bounded sample of ~10-20 s of CPU work: rate from one 64-image pass, then one pass sized from it
        v0, _, cores = cpu_arm(1, 1, args.cpu_images)
        n_img = int(min(1024, max(args.cpu_images, 32 * round(v0 * 12 / 32))))
        v, spp, cores = cpu_arm(1, 0, n_img)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "seconds": spp,
               "sample": f"1 x {n_img} images (same model, beam 3, L=16; sized for ~12 s), torch fp32 oracle port of the reference path on host cores"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches_per_step * args.steps, "clocks": sampler.summary(), "roofline": roofline,
            "cpu_baseline": cpu, "train": train}
    print(json.dumps(line), file=out, flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
