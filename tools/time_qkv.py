"""Ablation timing of the small-K GEMMs of one decoder layer on the ragged C2 batch (FS2_GEMM_DBG bits: 1 no global
stores, 2 no epilogue, 4 no TMA loads, 8 no MMA)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs2b200 import sub
ops, synth, G = sub("ops"), sub("synthetic"), sub("gemm")
B, T = 64, 1000
batch = synth.make_batch(**synth.CONFIGS["C2"])
lens = batch[7].clamp(max=T).cuda()
V = int(lens.sum())
bf = torch.bfloat16
x = torch.randn(B * T, 256, device="cuda").to(bf)
x1k = torch.randn(B, T, 1024, device="cuda").to(bf)
wqkv = torch.randn(768, 256, device="cuda").to(bf); bq = torch.zeros(768, device="cuda")
wo = torch.randn(256, 256, device="cuda").to(bf); bo = torch.zeros(256, device="cuda")
w2p = torch.randn(256, 1, 1024, device="cuda").to(bf); b2 = torch.zeros(256, device="cuda")
dqkv = torch.randn(B * T, 768, device="cuda").to(bf)
flush = torch.zeros(96 << 20, device="cuda")
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(7):
        if not os.environ.get('FS2_NOFLUSH'): flush.sum()
        torch.cuda._sleep(300000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[3]
x3 = x.view(B, T, 256)
cases = [
 ("qkv fwd   K=256  N=768 ", lambda: ops.linear_fwd(x, wqkv, bq, lens=lens, T=T, tail=ops.NO_TAIL), 2.0 * V * 768 * 256, (V * 256 + V * 768) * 2),
 ("oproj fwd K=256  N=256 ", lambda: ops.linear_fwd(x, wo, bo, lens=lens, T=T, tail=ops.NO_TAIL), 2.0 * V * 256 * 256, (V * 512) * 2),
 ("k1 fwd    K=1024 N=256 ", lambda: ops.conv_fwd(x1k, w2p, b2, lens=lens, tail=ops.NO_TAIL), 2.0 * V * 1024 * 256, (V * 1280) * 2),
 ("k1 dgrad  K=256  N=1024", lambda: ops.conv_dgrad(x3, w2p, 1024, lens=lens, tail=4), 2.0 * V * 1024 * 256, (V * 1280) * 2),
 ("qkv dgrad K=768  N=256 ", lambda: ops.linear_dgrad(dqkv, wqkv, epilogue=G.EPI_ADD_AUX, aux=x, lens=lens, T=T), 2.0 * V * 768 * 256, (V * 1280) * 2),
]
print("FS2_GEMM_DBG=%s  valid rows %d" % (os.environ.get("FS2_GEMM_DBG", "0"), V))
for name, fn, fl, by in cases:
    us = t(fn)
    print("  %s %7.1f us  %6.0f TFLOP/s  %5.0f GB/s (algorithmic)" % (name, us, fl / us / 1e6, by / us / 1e3))
