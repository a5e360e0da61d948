"""Window attention core kernel at the bench shape, for an ncu capture."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import ldic_b200
from ldic_b200 import ops
dev = torch.device("cuda", 0)
B, H, W, C, heads, ws = 4, 288, 480, 192, 8, 8
qkv = [torch.randn(B, H, W, C, device=dev).to(torch.bfloat16) for _ in range(3)]
bias = torch.randn((2 * ws - 1) ** 2, heads, device=dev) * 0.1     # relative_position_bias_table, indexed inside the kernel
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    ops.window_attention_core(qkv[0], qkv[1], qkv[2], bias, heads, ws, 4)
torch.cuda.synchronize()
print("done")
