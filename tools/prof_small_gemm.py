"""The epilogue-bound small-K GEMM of the FFT block, alone: fused Q|K|V projection (N=768, K=256) on a ragged
C2 batch (B=64, T=1000).  Used for `ncu --set full --import-source on` captures of gemm_tc2_kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs2b200 import sub  # noqa: E402

ops = sub("ops")
synth = sub("synthetic")
B, T = 64, 1000
batch = synth.make_batch(**synth.CONFIGS["C2"])
lens = batch[7].clamp(max=T).cuda()
x = torch.randn(B * T, 256, device="cuda").to(torch.bfloat16)
w = torch.randn(768, 256, device="cuda").to(torch.bfloat16)
bias = torch.zeros(768, device="cuda")
for _ in range(4):
    y = ops.linear_fwd(x, w, bias, lens=lens, T=T, tail=ops.NO_TAIL)
torch.cuda.synchronize()
print("ok")
