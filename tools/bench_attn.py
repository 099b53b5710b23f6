"""Fused attention kernels alone on the C2 decoder shape (B=64, T=1000, H=2, dk=128), dense and with the C2
batch's ragged lengths: CUDA events, L2 flushed, GPU kept busy until the launch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs2b200 import sub  # noqa: E402

ops = sub("ops")
synth = sub("synthetic")


def timeit(fn, iters=9):
    flush = torch.zeros(128 << 20, dtype=torch.float32, device="cuda")
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.sum()
        torch.cuda._sleep(400000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    B, T, H, dk = 64, 1000, 2, 128
    batch = synth.make_batch(**synth.CONFIGS["C2"])
    ragged = batch[7].clamp(max=T).cuda()
    dense = torch.full((B,), T, dtype=torch.int64, device="cuda")
    qkv = torch.randn(B, T, 3 * H * dk, device="cuda").to(torch.bfloat16)
    d_o = torch.randn(B, T, H * dk, device="cuda").to(torch.bfloat16)
    for tag, lens in (("dense", dense), ("ragged C2", ragged)):
        fl = float((lens.double() ** 2).sum()) * 4 * dk * H  # QK^T + PV, algorithmic
        sched = None if os.environ.get("FS2_NO_ATTN_SCHED") else ops.attn_schedule(lens, T, H)
        out, lse2 = ops.attn_fwd(qkv, lens, H, dk, sched)
        us_f = timeit(lambda: ops.attn_fwd(qkv, lens, H, dk, sched))
        us_b = timeit(lambda: ops.attn_bwd(qkv, out, d_o, lse2, lens, H, dk, sched))
        print("%-10s fwd %7.1f us (%6.1f TFLOP/s algorithmic)   bwd %7.1f us (%6.1f TFLOP/s, 2.5x fwd flops)"
              % (tag, us_f, fl / us_f / 1e6, us_b, 2.5 * fl / us_b / 1e6))


if __name__ == "__main__":
    main()
