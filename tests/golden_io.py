"""Loader for the committed golden fixtures (tests/golden/*.npz, made by make_golden.py)."""
import ast
import os

import numpy as np
import torch

from oracle import ort_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    out = {"w": {}, "g": {}, "u": {}}
    for k in z.files:
        if k.startswith("w::"):
            out["w"][k[3:]] = torch.from_numpy(z[k])
        elif k.startswith("g::"):
            out["g"][k[3:]] = torch.from_numpy(z[k])
        elif k.startswith("u::"):
            out["u"][k[3:]] = torch.from_numpy(z[k])
        elif k in ("cfg_keys", "cfg_vals"):
            continue
        else:
            out[k] = torch.from_numpy(z[k])
    if "cfg_keys" in z.files:
        cfg = {str(k): ast.literal_eval(str(v)) for k, v in zip(z["cfg_keys"], z["cfg_vals"])}
        out["cfg_dict"] = cfg
        out["cfg"] = O.Cfg(**cfg)
        out["w"]["model.tgt_embed.1.pe"] = O.positional_encoding(cfg["d_model"], 5000).unsqueeze(0)
    return out
