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

# fused Q|K|V projection: weight gradient (768 x 256, three segments) with / without the Q and V bias gradients
HD, D = 256, 256
dqkv = (torch.randn(B, T, 3 * HD, device="cuda") * valid).to(bf)
xq = (torch.randn(B, T, D, device="cuda") * valid).to(bf)
gb = [(torch.zeros(HD, D, device="cuda"), None) if i % 2 == 0 else (torch.zeros(HD, device="cuda"), None) for i in range(6)]
d2, x2 = dqkv.view(B * T, 3 * HD), xq.view(B * T, D)
print("QKV wgrad alone (bias_done) %7.1f us" % t(lambda: ops.qkv_param_grads(d2, x2, gb, HD, lens=lens, T=T, bias_done=True)))
print("QKV wgrad + fused Q/V bias  %7.1f us" % t(lambda: ops.qkv_param_grads(d2, x2, gb, HD, lens=lens, T=T)))
print("Q/V column sums alone       %7.1f us" % t(lambda: ops._ck(ops._L().fs2_colsum_ragged2_bf16(
    d2.data_ptr(), 3 * HD, B, T, HD, lens.data_ptr(), 2 * HD, gb[1][0].data_ptr(), gb[5][0].data_ptr(), ops._st()), "colsum")))
for g_ in gb: g_[0].zero_()
ops.qkv_param_grads(d2, x2, gb, HD, lens=lens, T=T)
refb = dqkv.float().sum((0, 1))
print("Q / V bias rel err %.2e %.2e, K bias untouched: %s" % (
    ((gb[1][0] - refb[:HD]).norm() / refb[:HD].norm()).item(),
    ((gb[5][0] - refb[2 * HD:]).norm() / refb[2 * HD:].norm()).item(), bool((gb[3][0] == 0).all())))
