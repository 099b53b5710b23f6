import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs2b200 import sub
ops = sub("ops"); L = ops._L()
M, C = 64000, 512
y = torch.randn(M, C, device="cuda").to(torch.bfloat16)
stats = torch.zeros(2, C, device="cuda"); L.fs2_bn_stats_bf16(y.data_ptr(), M, C, stats.data_ptr(), ops._st())
g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda"); out = torch.empty_like(y)
seed = torch.zeros(1, dtype=torch.int64, device="cuda")
dout = torch.randn(M, C, device="cuda").to(torch.bfloat16); dst = torch.zeros(2, C, device="cuda"); dy = torch.empty_like(y)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn):
    for _ in range(2): fn()
    ts = []
    for _ in range(5):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return 1e3 * sorted(ts)[2]
for act in (0, 1):
    for p in (0.0, 0.5):
        us = t(lambda: L.fs2_bn_apply_fwd(y.data_ptr(), stats.data_ptr(), g.data_ptr(), b.data_ptr(), M, C, act, p, 7, seed.data_ptr(), out.data_ptr(), None, None, ops._st()))
        us2 = t(lambda: L.fs2_bn_bwd(dout.data_ptr(), 0, y.data_ptr(), stats.data_ptr(), g.data_ptr(), b.data_ptr(), M, C, act, p, 7, seed.data_ptr(), dst.data_ptr(), dy.data_ptr(), ops._st()))
        print("act=%d p=%.1f: apply %.1f us (%.2f TB/s)  bwd(reduce+apply) %.1f us" % (act, p, us, 2 * M * C * 2 / us / 1e6, us2))
us = t(lambda: L.fs2_bn_stats_bf16(y.data_ptr(), M, C, stats.data_ptr(), ops._st())); print("stats %.1f us" % us)
x = torch.randn(250, 256, 256, device="cuda").to(torch.bfloat16)
for p in (0.0, 0.2):
    us = t(lambda: ops.ln_fwd(x, x, g[:256], b[:256], None, p, 1, 5)); print("ln_fwd p=%.1f %.1f us" % (p, us))
c = torch.zeros(1024, device="cuda"); xx = torch.randn(64000, 1024, device="cuda").to(torch.bfloat16)
us = t(lambda: ops.colsum(xx, c)); print("colsum 64000x1024 %.1f us (%.2f TB/s)" % (us, 64000 * 1024 * 2 / us / 1e6))
