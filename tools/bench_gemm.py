"""Micro-benchmark of the tcgen05 GEMM engine on the FastSpeech2 shapes (CUDA events, L2 flushed)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs2b200 import sub  # noqa: E402

G = sub("gemm")


def timeit(fn, iters=20):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda._sleep(400000)  # no launch gap inside the timed window
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def rnd(*s):
    return torch.randn(*s, device="cuda").to(torch.bfloat16)


def main():
    B, T = (int(sys.argv[1]) if len(sys.argv) > 1 else 16), (int(sys.argv[2]) if len(sys.argv) > 2 else 850)
    M = B * T
    rows = []
    # linear shapes
    for name, (m, n, k) in {"qkv": (M, 768, 256), "oproj": (M, 256, 256), "conv_k1": (M, 256, 1024),
                            "big": (8192, 8192, 8192)}.items():
        x, w = rnd(m, k), rnd(n, k)
        y = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: G.gemm(G.operand(x, k, m), G.operand(w, k, n), y, m, n, k))
        rows.append((name, ms, 2.0 * m * n * k / ms / 1e9))
    # conv k9 fwd 256 -> 1024
    x = rnd(B, T, 256)
    wp = rnd(1024, 9 * 256)
    y = torch.empty(B, T, 1024, device="cuda", dtype=torch.bfloat16)
    ms = timeit(lambda: G.gemm(G.operand(x, 256, T, B), G.operand(wp, 9 * 256, 1024), y, T, 1024, 256,
                               Z=B, taps=9, tap_shift0=-4, b_tap_kstride=256, d_zdiv=1,
                               d_zdiv_stride=T * 1024))
    rows.append(("conv_k9_fwd", ms, 2.0 * M * 1024 * 2304 / ms / 1e9))
    # conv k9 wgrad
    dy = rnd(B, T, 1024)
    dw = torch.zeros(1024, 9, 256, device="cuda")
    for splits in (1, 2, 4):
        ms = timeit(lambda: G.wgrad(G.operand(dy, 1024, T, B, mn_major=True),
                                    G.operand(x, 256, T, B, mn_major=True), dw, 1024, 256, taps=9,
                                    tap_shift0=-4, ldd=2304, d_col_stride=1, d_tap_stride=256,
                                    splits=splits))
        rows.append(("conv_k9_wgrad_s%d" % splits, ms, 2.0 * M * 1024 * 2304 / ms / 1e9))
    # attention bmm QK^T
    qkv = rnd(B, T, 768)
    Tp = ((T + 63) // 64) * 64
    s = torch.empty(B * 2, T, Tp, device="cuda")
    ms = timeit(lambda: G.gemm(G.operand(qkv, 768, T, B, zdiv=2, zmod_stride=128),
                               G.operand(qkv, 768, T, B, inner_base=256, zdiv=2, zmod_stride=128), s, T,
                               T, 128, Z=B * 2, ldd=Tp, d_zdiv=1, d_zdiv_stride=T * Tp))
    rows.append(("attn_qk_f32out", ms, 2.0 * B * 2 * T * T * 128 / ms / 1e9))
    for name, ms, tf in rows:
        print("%-22s %8.3f ms  %8.1f TFLOP/s" % (name, ms, tf))


if __name__ == "__main__":
    main()
