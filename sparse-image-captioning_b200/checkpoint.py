"""On-disk formats of the pruned checkpoints <-> the layouts the kernels consume.

The reference writes, after supermask training (scripts/train_n_prune_transformer.py:251-291, pruning/prune.py:176-226):
  model_best_pruned_sparse.pth : every pruned weight as a COO tensor (``.to_sparse()``), masks dropped     (14.5 MB for ORT @95 %)
  model_best_pruned.pth        : the same weights dense, zeros in place
  model_best_bin_mask.pth      : weights + 0/1 masks (binarized supermasks), for ``mask_freeze`` fine-tuning / SCST
and evaluates the dense class on the densified sparse file (eval_model.py:64-88, ``densify_state_dict``).  These helpers read
any of the three (and plain ``*_prune`` checkpoints with logits) into an ``OrtEngine`` - COO entries go to the engine as they are,
which packs them into the dense bf16 / CSR / sliced-ELL operand it runs - and write the same three files from a pruned model,
so reference checkpoints run unmodified and checkpoints written here load in the reference.
"""
import os
from typing import Dict, Optional, Tuple

import torch

from . import prune
from .engine import ModelCfg, OrtEngine

_SUFFIX = "_pruning_mask"


def classify(state_dict: Dict[str, torch.Tensor]) -> str:
    """"sparse" (COO weights), "bin_mask" (weights + 0/1 masks), "supermask" (weights + logits) or "dense"."""
    if any(torch.is_tensor(v) and v.is_sparse for v in state_dict.values()):
        return "sparse"
    masks = [v for k, v in state_dict.items() if k.endswith(_SUFFIX)]
    if not masks:
        return "dense"
    binary = all(bool(((m == 0) | (m == 1)).all()) for m in masks)
    return "bin_mask" if binary else "supermask"


def to_dense_class(state_dict: Dict[str, torch.Tensor], mask_type: Optional[str] = None) -> Tuple[Dict[str, torch.Tensor], str]:
    """Any of the checkpoint flavours -> dense-class (`relation_transformer`) state dict (COO tensors are kept sparse: the
    engine densifies / packs them on the device) and the flavour found."""
    kind = classify(state_dict)
    if kind in ("sparse", "dense"):
        return dict(state_dict), kind
    if kind == "bin_mask":
        return prune.fold_masks(state_dict, prune.MASK_FREEZE), kind          # masks are used raw (0 / 1)
    return prune.fold_masks(state_dict, mask_type or prune.REGULAR), kind     # logits: rint(sigmoid(S)) (eval semantics)


def load_checkpoint(path: str, map_location="cpu") -> Dict[str, torch.Tensor]:
    return torch.load(path, map_location=map_location, weights_only=False)


def engine_from_checkpoint(path: str, config, *, device="cuda", precision="bf16", sparse_backend="dense", **kw) -> OrtEngine:
    """`eval_model.py` in one call: checkpoint file -> inference engine (dense tensor-core GEMMs by default; ``sparse_backend``
    "sell" / "csr" / "auto" run the pruned decoder linears from the packed sparse formats)."""
    sd, _ = to_dense_class(load_checkpoint(path))
    cfg = config if isinstance(config, ModelCfg) else ModelCfg(config)
    return OrtEngine(sd, cfg, precision=precision, sparse_backend=sparse_backend, device=device, **kw)


@torch.no_grad()
def save_pruned_checkpoints(model, log_dir: str, stem: str = "model_best") -> Dict[str, str]:
    """`maybe_prune_best_model` (train_n_prune_transformer.py:251-291): prune the weights with the (binarized) masks, then write
    ``<stem>_pruned_sparse.pth``, ``<stem>_pruned.pth`` and - for supermasks - ``<stem>_bin_mask.pth``; also ``sparsities.csv``."""
    os.makedirs(log_dir, exist_ok=True)
    model.prune_weights()
    overall, nnz, tensor_sps, names = model.all_mask_sparsities
    out = {"sparse": os.path.join(log_dir, f"{stem}_pruned_sparse.pth"), "dense": os.path.join(log_dir, f"{stem}_pruned.pth")}
    cpu = lambda sd: {k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in sd.items()}
    torch.save(cpu(model.state_dict_sparse(discard_pruning_mask=True, prune_weights=False)), out["sparse"])
    torch.save(cpu(model.state_dict_dense(discard_pruning_mask=True, prune_weights=False)), out["dense"])
    if model.mask_type == prune.REGULAR:
        out["bin_mask"] = os.path.join(log_dir, f"{stem}_bin_mask.pth")
        torch.save(cpu(model.state_dict_dense(discard_pruning_mask=False, prune_weights=False, binarize_supermasks=True)), out["bin_mask"])
    with open(os.path.join(log_dir, "sparsities.csv"), "w") as f:
        f.write(f"sparsity,nnz,{','.join(names)}\n")
        f.write(f"{float(overall):.5f},{int(nnz)},{','.join(f'{float(s):.5f}' for s in tensor_sps)}")
    return out
