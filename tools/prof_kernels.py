"""One launch sequence of the heavy kernels of a C2 decoder layer on the ragged batch (for `ncu --set full`):
  python tools/prof_kernels.py [reps]
Kernels appear in this order, `reps` times each (default 2; profile the last): attention fwd, attention bwd (prep,
dK/dV, dQ), QKV fwd, out-proj fwd, k=1 conv fwd, k=9 conv fwd, k=1 conv dgrad, QKV dgrad, k=9 conv dgrad / wgrad,
LayerNorm fwd / bwd."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs2b200 import sub  # noqa: E402

ops, synth, G = sub("ops"), sub("synthetic"), sub("gemm")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B, T, D, Dh, H = 64, 1000, 256, 1024, 2
batch = synth.make_batch(**synth.CONFIGS["C2"])
lens = batch[7].clamp(max=T).cuda()
bf = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s: (torch.randn(*s, device="cuda", generator=g) * 0.5).to(bf)
x, dy = rnd(B, T, D), rnd(B, T, D)
h, dh = rnd(B, T, Dh), rnd(B, T, Dh)
qkv, dqkv = rnd(B, T, 3 * D), rnd(B, T, 3 * D)
w1p, w2p, wqkv, wo = rnd(Dh, 9, D), rnd(D, 1, Dh), rnd(3 * D, D), rnd(D, D)
b1, b2, bq, bo = (torch.zeros(n, device="cuda") for n in (Dh, D, 3 * D, D))
gw1 = torch.zeros(Dh, 9, D, device="cuda").permute(0, 2, 1)
gb1 = torch.zeros(Dh, device="cuda")
hmask = torch.empty(B * T, Dh // 64, dtype=torch.int64, device="cuda")
x2, dy2 = x.view(B * T, D), dy.view(B * T, D)
NT = ops.NO_TAIL
gamma, beta = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
dg, db, dbias = (torch.zeros(D, device="cuda") for _ in range(3))
sched = ops.attn_schedule(lens, T, H)
o3, lse = ops.attn_fwd(qkv, lens, H, D // H, sched)
y, mean, rstd, keep = ops.ln_fwd(x, dy, gamma, beta, lens, 0.2, 1, 5)
torch.cuda.synchronize()
seq = [
    lambda: ops.attn_fwd(qkv, lens, H, D // H, sched),
    lambda: ops.attn_bwd(qkv, o3, dy, lse, lens, H, D // H, sched),
    lambda: ops.linear_fwd(x2, wqkv, bq, lens=lens, T=T, tail=NT),
    lambda: ops.linear_fwd(x2, wo, bo, lens=lens, T=T, tail=NT),
    lambda: ops.conv_fwd(h, w2p, b2, lens=lens, tail=NT),
    lambda: ops.conv_fwd(x, w1p, b1, relu=True, lens=lens, tail=NT, relu_mask=hmask),
    lambda: ops.conv_dgrad(dy, w2p, Dh, epilogue=G.EPI_RELU_BWD, lens=lens, tail=4, relu_mask=hmask),
    lambda: ops.linear_dgrad(dqkv.view(B * T, 3 * D), wqkv, epilogue=G.EPI_ADD_AUX, aux=x2, lens=lens, T=T),
    lambda: ops.conv_dgrad(dh, w1p, D, epilogue=G.EPI_ADD_AUX, aux=x, lens=lens),
    lambda: ops.conv_wgrad(dh, x, gw1, lens=lens, dbias=gb1),  # + the fused bias gradient, as in the step
    lambda: ops.ln_fwd(x, dy, gamma, beta, lens, 0.2, 1, 5),
    lambda: ops.ln_bwd(dy, x, dy, gamma, mean, rstd, lens, 0.2, 1, keep, dg, db, True, dbias=dbias),
]
for fn in seq:
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
print("ok")
