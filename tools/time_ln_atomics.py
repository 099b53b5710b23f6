"""LayerNorm backward on the ragged C2 batch with and without its final dgamma / dbeta / dbias atomics."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs2b200 import sub
ops, synth = sub("ops"), sub("synthetic")
B, T, C = 64, 1000, 256
lens = synth.make_batch(**synth.CONFIGS["C2"])[7].clamp(max=T).cuda()
bf = torch.bfloat16
x, res, dy = (torch.randn(B, T, C, device="cuda").to(bf) for _ in range(3))
g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
dg, db, dbias = (torch.zeros(C, device="cuda") for _ in range(3))
flush = torch.zeros(96 << 20, device="cuda")
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(9):
        flush.sum(); torch.cuda._sleep(300000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[4]
y, mean, rstd, keep = ops.ln_fwd(x, res, g, b, lens, 0.2, 1, 5)
L = ops._L()
def run(dg_, db_, dbias_):
    dx = torch.empty_like(x); dres = torch.empty_like(x)
    def f():
        ops._ck(L.fs2_ln_bwd_bf16(dy.data_ptr(), x.data_ptr(), res.data_ptr(), g.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                  lens.data_ptr(), B, T, C, 0.2, 1, 0, keep.data_ptr(), dx.data_ptr(), dres.data_ptr(),
                                  None if dg_ is None else dg_.data_ptr(), None if db_ is None else db_.data_ptr(),
                                  None if dbias_ is None else dbias_.data_ptr(), ops._st()), "ln_bwd")
    return t(f)
print("ln_bwd with atomics   : %.1f us" % run(dg, db, dbias))
print("ln_bwd without atomics: %.1f us" % run(None, None, None))
