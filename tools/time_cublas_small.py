"""cuBLAS (torch.matmul, bf16) on the small-K / narrow-N GEMM shapes of one C2 decoder layer, L2 flushed before every
launch: the yardstick for the hand-written kernels on shapes that are bound by operand streaming, not by the MMA."""
import torch
V = 35945
bf = torch.bfloat16
flush = torch.zeros(96 << 20, device="cuda")
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(9):
        flush.sum(); torch.cuda._sleep(300000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[4]
for name, M, N, K in [("qkv fwd", V, 768, 256), ("oproj fwd", V, 256, 256), ("k1 fwd", V, 256, 1024),
                      ("k1 dgrad", V, 1024, 256), ("qkv dgrad", V, 256, 768), ("k9-as-gemm", V, 1024, 2304)]:
    a = torch.randn(M, K, device="cuda").to(bf)
    w = torch.randn(N, K, device="cuda").to(bf)
    out = torch.empty(M, N, device="cuda", dtype=bf)
    us = t(lambda: torch.matmul(a, w.t(), out=out))
    print("  cuBLAS %-10s M=%d N=%4d K=%4d %7.1f us  %6.0f TFLOP/s  %5.0f GB/s" % (name, M, N, K, us, 2.0 * M * N * K / us / 1e6, (M * K + M * N) * 2 / us / 1e3))
src = torch.empty(36 << 20, device="cuda", dtype=bf); dst = torch.empty_like(src)
print("  copy 72 MB -> 72 MB: %.1f us" % t(lambda: dst.copy_(src)))
