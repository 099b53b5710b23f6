"""k = 9 FFN Conv1d weight gradient on the ragged C2 batch: alone, with the fused bias gradient (fs2_gemm::a_colsum),
and the separate column-sum launch it replaces."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs2b200 import sub
ops, synth, G = sub("ops"), sub("synthetic"), sub("gemm")
B, T, Co, Ci, k = 64, 1000, 1024, 256, 9
batch = synth.make_batch(**synth.CONFIGS["C2"])
lens = batch[7].clamp(max=T).cuda()
valid = (torch.arange(T, device="cuda")[None, :] < lens[:, None])[..., None]
bf = torch.bfloat16
dy = (torch.randn(B, T, Co, device="cuda") * valid).to(bf)
x = (torch.randn(B, T, Ci, device="cuda") * valid).to(bf)
dw = torch.zeros(Co, k, Ci, device="cuda").permute(0, 2, 1)
db = torch.zeros(Co, device="cuda")
flush = torch.zeros(96 << 20, device="cuda")
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(9):
        flush.sum()
        torch.cuda._sleep(300000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[4]
print("wgrad k=9 alone        %7.1f us" % t(lambda: ops.conv_wgrad(dy, x, dw, lens=lens)))
print("wgrad k=9 + fused bias %7.1f us" % t(lambda: ops.conv_wgrad(dy, x, dw, lens=lens, dbias=db)))
print("column sums alone      %7.1f us" % t(lambda: ops.colsum(dy.view(B * T, Co), db, lens=lens, T=T)))
db.zero_(); ops.conv_wgrad(dy, x, dw, lens=lens, dbias=db)
ref = dy.float().sum((0, 1))
print("bias gradient rel err  %.2e" % ((db - ref).norm() / ref.norm()).item())
