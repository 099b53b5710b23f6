"""k = 1 FFN convolution + LayerNorm on the ragged C2 batch: separate kernels vs the fused GEMM epilogue."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs2b200 import sub
ops, synth, G = sub("ops"), sub("synthetic"), sub("gemm")
B, T, D, Dh = 64, 1000, 256, 1024
batch = synth.make_batch(**synth.CONFIGS["C2"])
lens = batch[7].clamp(max=T).cuda()
bf = torch.bfloat16
h = torch.randn(B, T, Dh, device="cuda").to(bf)
x = torch.randn(B, T, D, device="cuda").to(bf)
w2p = (torch.randn(D, 1, Dh, device="cuda") * Dh ** -0.5).to(bf)
b2 = torch.zeros(D, device="cuda")
gamma, beta = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
flush = torch.zeros(96 << 20, device="cuda")
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(9):
        flush.sum(); torch.cuda._sleep(300000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[4]
y = torch.empty(B, T, D, dtype=bf, device="cuda"); v = torch.empty_like(y)
mean = torch.empty(B * T, device="cuda"); rstd = torch.empty(B * T, device="cuda")
keep = torch.empty(B * T, D // 8, dtype=torch.uint8, device="cuda")
seed = ops._Rng.tensor(torch.device("cuda"))
def fused(p):
    G.gemm(G.operand(h, Dh, T, B), G.operand(w2p, Dh, D), y, T, D, Dh, Z=B, bias=b2, d_zdiv=1, d_zdiv_stride=T * D,
           row_lens=lens, tail_rows=0, ln=dict(gamma=gamma, beta=beta, res=x, p=p, salt=5, seed_dev=seed if p > 0 else None,
                                               v=v, mean=mean, rstd=rstd, keep=keep if p > 0 else None))
f = ops.conv_fwd(h, w2p, b2, lens=lens, tail=ops.NO_TAIL)
for p in (0.0, 0.2):
    a = t(lambda: ops.conv_fwd(h, w2p, b2, lens=lens, tail=ops.NO_TAIL))
    b = t(lambda: ops.ln_fwd(f, x, gamma, beta, lens, p, 1, 5))
    c = t(lambda: fused(p))
    print("p=%.1f  conv %.1f us + LN %.1f us = %.1f   fused %.1f us (%s)" % (p, a, b, a + b, c, ops._L().fs2_last_kernel().decode()))
