"""Split-count sweep of the three small weight gradients of a decoder layer on the ragged C2 batch
(output projection 256x256, QKV 768x256, k = 1 FFN conv 256x1024; reduction = 35,945 valid frames)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs2b200 import sub
ops, synth, G = sub("ops"), sub("synthetic"), sub("gemm")
B, T = 64, 1000
batch = synth.make_batch(**synth.CONFIGS["C2"])
lens = batch[7].clamp(max=T).cuda()
valid = (torch.arange(T, device="cuda")[None, :] < lens[:, None])[..., None]
bf = torch.bfloat16
mk = lambda c: (torch.randn(B, T, c, device="cuda") * valid).to(bf)
d256, x256, d768, x1024 = mk(256), mk(256), mk(768), mk(1024)
flush = torch.zeros(96 << 20, device="cuda")
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(7):
        flush.sum(); torch.cuda._sleep(300000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[3]
cases = [("oproj 256x256 ", d256, x256, 256, 256), ("qkv   768x256 ", d768, x256, 768, 256), ("k1    256x1024", d256, x1024, 256, 1024)]
for name, dy, x, Mo, No in cases:
    dw = torch.zeros(Mo, No, device="cuda")
    a, b = ops._wgrad_operands(dy.view(B * T, Mo), x.view(B * T, No), lens, T)
    cur = ops._splits(Mo, No, 1, (B * T + 63) // 64)
    line = "%s now %3d:" % (name, cur)
    for s in sorted(set([cur, 4, 9, 18, 24, 37, 74])):
        us = t(lambda: G.wgrad(a, b, dw, Mo, No, splits=s, row_lens=lens))
        line += "  s=%d %.1f" % (s, us)
    print(line + "  (%s)" % ops._L().fs2_last_kernel().decode())
