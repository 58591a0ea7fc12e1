"""Data-parallel plumbing (one process per GPU, torch.distributed over NCCL / NVLink; gloo on CPU for tests).

The path shards by image: inference and SCST rollouts need no collective; SMP training needs one
exchange per step — a SUM all-reduce of the flat gradient buffers (weights AND mask logits), with every rank's loss
normalised by the GLOBAL token count so that the summed gradient equals the single-process gradient of the whole
batch (utils/losses.py:42 normalises by sum(mask)).  Value clipping happens after the exchange, inside the fused
Adam kernel, as in utils/optim.py:187-191.  All ranks must share the mask seed (sampler.set_mask_seed / OrtTrainer
seed): the reference draws ONE Bernoulli mask per layer per step for the whole batch (masked_layer.py:97).
"""
import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous block partition of n_items (images) over ranks; the first n_items % world ranks get one extra."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def global_token_count(masks: torch.Tensor, T: int, group=None) -> torch.Tensor:
    """Sum over all ranks of the loss-mask entries masks[:, 1:T+1] (the denominator of LanguageModelCriterion)."""
    cnt = masks[:, 1: T + 1].float().sum().reshape(1).clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    return cnt


def make_all_reduce(group=None, async_op=False):
    """Callable for OrtTrainer.train_step(all_reduce=...): in-place SUM over ranks of one (slice of a) flat gradient
    buffer.  ``async_op``: return the work handle instead of waiting, so that the trainer can overlap the exchange of
    the buckets one backward phase finished with the next phase (OrtTrainer.grad_buckets); ``handle.wait()`` orders
    the current CUDA stream after the collective (NCCL) or blocks the host (gloo)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None

    def _ar(flat):
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return work if async_op else None

    return _ar
