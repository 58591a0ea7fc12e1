"""BASELINE.json configs[3] and [4] at full size on the B200 path (sanity + agreement with the fp32 verification mode):
  [3] ACORT-base (2 unique layers applied 3x, share_att='kv', radix vocab 771, L=26) at 99.1 % sparsity, beam 5, 512 images
  [4] SCST rollout for ORT: one beam-5 rollout + one greedy rollout, 1024 images (CIDEr scoring excluded)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sparse_caption_b200 import synthetic
from sparse_caption_b200.engine import ModelCfg, OrtEngine
from sparse_caption_b200.detok import radix_to_word_ids
dev = torch.device("cuda")

def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = fn()
    e1.record(); torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / n

# ---- configs[3]: ACORT ----
acfg = ModelCfg(dict(bench.CFG, vocab_size=771, max_seq_length=26, share_att_encoder="kv", share_att_decoder="kv",
                     share_layer_encoder=(0, 0, 0, 1, 1, 1), share_layer_decoder=(0, 0, 0, 1, 1, 1), eos_token_id=770, bos_token_id=769))
sd = synthetic.random_state_dict(acfg, seed=5, sparsity=0.991, device=dev)
att, boxes = synthetic.synthetic_inputs(512, 36, 2048, seed=7, pin=True)
eng = OrtEngine(sd, acfg, precision="bf16", device=dev)
(seq, lp), ms = timed(lambda: eng.sample(att, boxes, None, {"beam_size": 5}))
assert tuple(seq.shape) == (512, 5, 26) and torch.isfinite(lp).all()
words, n = radix_to_word_ids(seq[:, 0], 768, 2, eos_id=770)
print(f"configs[3] ACORT 99.1% beam-5 L=26, 512 images: {ms:.2f} ms/batch ({512 / ms * 1e3:.0f} captions/s incl. H2D), "
      f"radix -> {words.shape[1]} word slots")
ref = OrtEngine(sd, acfg, precision="fp32", device=dev)
rseq, rlp = ref.sample(att[:32], boxes[:32], None, {"beam_size": 5})
agree = float((seq[:32, 0, 0] == rseq[:, 0, 0]).float().mean())
same = seq[:32, 0, 0] == rseq[:, 0, 0]
print(f"  bf16 vs fp32 mode, first token of the best beam identical on {agree * 100:.0f}% of 32 images, "
      f"max |dlogp| on those {float((lp[:32, 0, 0][same] - rlp[:, 0, 0][same]).abs().max()):.4f}")
del eng, ref
# ---- configs[4]: SCST rollout ----
ocfg = ModelCfg(bench.CFG)
sd = synthetic.random_state_dict(ocfg, seed=6, sparsity=0.95, device=dev)
att, boxes = synthetic.synthetic_inputs(1024, 36, 2048, seed=8, pin=True)
eng = OrtEngine(sd, ocfg, precision="bf16", device=dev)
def rollout():
    enc = eng.encode(att, boxes)
    s5, l5 = eng.decode(enc, {"beam_size": 5})
    s1, l1 = eng.decode(enc, {"beam_size": 1})
    return s5, l5, s1, l1
(s5, l5, s1, l1), ms = timed(rollout)
assert tuple(s5.shape) == (1024, 5, 16) and tuple(s1.shape) == (1024, 1, 16) and torch.isfinite(l5).all() and torch.isfinite(l1).all()
print(f"configs[4] SCST rollout (beam-5 + greedy, shared encoder pass), 1024 images: {ms:.2f} ms ({1024 / ms * 1e3:.0f} images/s incl. H2D)")
sr, lr_ = eng.sample(att[:64], boxes[:64], None, {"beam_size": 0, "num_random_sample": 5, "sample_seed": 1})
print(f"  multinomial rollout (5 samples/image, 64 images): shape {tuple(sr.shape)}, mean log-prob {float(lr_.mean()):.3f}")
