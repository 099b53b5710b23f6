"""The dominant kernel of the step, alone: gemm_tc_kernel<256,4> on the decoder's k=9 Conv1d forward
(B=64, T=1000, 256 -> 1024).  Used for the `ncu --set full` capture under profiles/."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs2b200 import sub  # noqa: E402

ops = sub("ops")
B, T = 64, 1000
x = torch.randn(B, T, 256, device="cuda").to(torch.bfloat16)
wp = torch.randn(1024, 9, 256, device="cuda").to(torch.bfloat16)
bias = torch.zeros(1024, device="cuda")
for _ in range(4):
    y = ops.conv_fwd(x, wp, bias, relu=True)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
