"""PostNet BatchNorm kernels at the C2 shape ([64000 x 512]) for `ncu --set full` (kernel-name regex:bn_)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs2b200 import sub  # noqa: E402

ops = sub("ops")
L = ops._L()
M, C = 64000, 512
yb = torch.randn(M, C, device="cuda").to(torch.bfloat16)
dob = torch.randn(M, C, device="cuda").to(torch.bfloat16)
stats, dst = torch.zeros(2, C, device="cuda"), torch.zeros(2, C, device="cuda")
ws = torch.empty(L.fs2_bn_workspace_floats(M, C), device="cuda")
gb, bb = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
ob, dyb = torch.empty_like(yb), torch.empty_like(yb)
keep = torch.empty(M, C // 8, dtype=torch.uint8, device="cuda")
seed = torch.zeros(1, dtype=torch.int64, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    L.fs2_bn_stats_bf16(yb.data_ptr(), M, C, ws.data_ptr(), stats.data_ptr(), 0.1, None, None, None, ops._st())
    L.fs2_bn_apply_fwd(yb.data_ptr(), stats.data_ptr(), gb.data_ptr(), bb.data_ptr(), M, C, 1, 0.5, 7,
                       seed.data_ptr(), ob.data_ptr(), None, None, keep.data_ptr(), ops._st())
    L.fs2_bn_bwd(dob.data_ptr(), 0, yb.data_ptr(), stats.data_ptr(), gb.data_ptr(), bb.data_ptr(), M, C, 1, 0.5, 7,
                 seed.data_ptr(), keep.data_ptr(), ws.data_ptr(), dst.data_ptr(), None, None, dyb.data_ptr(), ops._st())
    torch.cuda.synchronize()
print("ok")
