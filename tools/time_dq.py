import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs2b200 import sub
ops = sub("ops")
B, T, H, dk = 64, 1000, 2, 128
qkv = torch.randn(B, T, 3 * H * dk, device="cuda").to(torch.bfloat16)
lens = torch.full((B,), T, device="cuda")
d_out = torch.randn(B, T, H * dk, device="cuda").to(torch.bfloat16)
out, lse2 = ops.attn_fwd(qkv, lens, H, dk)
for _ in range(2): ops.attn_bwd(qkv, out, d_out, lse2, lens, H, dk)
torch.cuda.synchronize()
import torch.cuda.profiler
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): ops.attn_bwd(qkv, out, d_out, lse2, lens, H, dk)
    torch.cuda.synchronize()
for e in prof.key_averages():
    if "attn" in e.key: print(os.environ.get("FS2_ATTN_DBG","0"), e.key[:40], "%.1f us" % (e.device_time_total / e.count))
