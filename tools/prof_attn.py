"""Fused attention kernels alone at the decoder shape of the bench (B=64, T=1000, H=2, dk=128)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs2b200 import sub  # noqa: E402

ops = sub("ops")
B, T, H, dk = (int(sys.argv[1]) if len(sys.argv) > 1 else 64), 1000, 2, 128
qkv = torch.randn(B, T, 3 * H * dk, device="cuda").to(torch.bfloat16)
lens = torch.randint(300, T + 1, (B,), device="cuda")
d_out = torch.randn(B, T, H * dk, device="cuda").to(torch.bfloat16)
for _ in range(3):
    out, lse2 = ops.attn_fwd(qkv, lens, H, dk)
    dqkv = ops.attn_bwd(qkv, out, d_out, lse2, lens, H, dk)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record(); out, lse2 = ops.attn_fwd(qkv, lens, H, dk); e[1].record()
dqkv = ops.attn_bwd(qkv, out, d_out, lse2, lens, H, dk); e[2].record(); torch.cuda.synchronize()
print("fwd %.3f ms  bwd %.3f ms" % (e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
