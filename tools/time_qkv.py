import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs2b200 import sub
G = sub("gemm")
M, N, K = 64000, 768, 256
x = torch.randn(M, K, device="cuda").to(torch.bfloat16); w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
bias = torch.randn(N, device="cuda"); y = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
f = lambda: G.gemm(G.operand(x, K, M), G.operand(w, K, N), y, M, N, K, bias=bias)
for _ in range(3): f()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(os.environ.get("FS2_GEMM_DBG", "0"), "qkv gemm %.1f us" % (1e3 * sorted(ts)[2]))
