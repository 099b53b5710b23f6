import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs2b200 import sub
G = sub("gemm")
M = 64000
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return 1e3*sorted(ts)[2]
for (N, K) in [(256, 256), (768, 256), (256, 1024)]:
    dy = torch.randn(M, N, device="cuda").to(torch.bfloat16); x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    dw = torch.zeros(N, K, device="cuda")
    for splits in (4, 8, 16, 31, 62, 74, 148):
        us = t(lambda: G.wgrad(G.operand(dy, N, M, mn_major=True), G.operand(x, K, M, mn_major=True), dw, N, K, splits=splits))
        print("wgrad N=%d K=%d splits=%d: %.1f us" % (N, K, splits, us))
