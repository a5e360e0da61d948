"""Golden vectors for SURVEY 8 f2 by EXECUTING THE UNMODIFIED reference class
layers/win_attention.py::WinBasedAttention (timm stubbed as in oracle/ref_harness.py).  Build container only:

    python tests/golden/make_golden_winattn.py
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402


def main():
    for name, kw in (("timm", {}), ("timm.models", {}),
                     ("timm.models.layers", dict(DropPath=nn.Identity, to_2tuple=lambda x: (x, x),
                                                 trunc_normal_=nn.init.trunc_normal_))):
        m = types.ModuleType(name); m.__dict__.update(kw); sys.modules[name] = m
    wa = ref_harness.load_leaf("layers/win_attention.py", "ref_win_attention")
    out = {}
    for tag, (dim, heads, ws, shift, B, H, W) in {"noshift": (192, 8, 8, 0, 2, 16, 24), "shift": (192, 8, 8, 4, 1, 16, 16),
                                                  "small": (64, 4, 4, 2, 2, 8, 12)}.items():
        torch.manual_seed(11)
        blk = wa.WinBasedAttention(dim=dim, num_heads=heads, window_size=ws, shift_size=shift).eval()
        with torch.no_grad():
            blk.attn.relative_position_bias_table.normal_(0, 0.5)       # default init (std .02) would hide bias errors
            blk.attn.qkv.weight.mul_(2.0)
            x = torch.randn(B, dim, H, W)
            y = blk(x)
        out[f"{tag}_cfg"] = np.array([dim, heads, ws, shift, B, H, W])
        out[f"{tag}_x"] = x.numpy(); out[f"{tag}_y"] = y.numpy()
        for k, v in blk.state_dict().items():
            out[f"{tag}_sd_{k}"] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "win_attention.npz"), **out)
    print("wrote win_attention.npz", {k: v.shape for k, v in out.items() if k.endswith(("_x", "_y"))})


if __name__ == "__main__":
    main()
