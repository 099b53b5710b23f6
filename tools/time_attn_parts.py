"""Times the three attention-backward kernels separately (CUDA events around the whole call minus variants)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs2b200 import sub
ops, synth = sub("ops"), sub("synthetic")
B, T, H, dk = 64, 1000, 2, 128
batch = synth.make_batch(**synth.CONFIGS["C2"])
for tag, lens in (("dense", torch.full((B,), T, dtype=torch.int64, device="cuda")), ("ragged", batch[7].clamp(max=T).cuda())):
    qkv = torch.randn(B, T, 3 * H * dk, device="cuda").to(torch.bfloat16)
    d_o = torch.randn(B, T, H * dk, device="cuda").to(torch.bfloat16)
    sched = None if os.environ.get("FS2_NO_ATTN_SCHED") else ops.attn_schedule(lens, T, H)
    out, lse2 = ops.attn_fwd(qkv, lens, H, dk, sched)
    from torch.profiler import profile, ProfilerActivity
    for _ in range(3): ops.attn_bwd(qkv, out, d_o, lse2, lens, H, dk, sched)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            ops.attn_fwd(qkv, lens, H, dk, sched)
            ops.attn_bwd(qkv, out, d_o, lse2, lens, H, dk, sched)
        torch.cuda.synchronize()
    print(tag, "FS2_ATTN_DBG=%s" % os.environ.get("FS2_ATTN_DBG", "0"))
    for e in prof.key_averages():
        if "attn" in e.key:
            print("   %-40s %8.1f us" % (e.key[:40], e.device_time_total / e.count))
