"""Micro-benchmark of the HBM-bound kernels on the C2 shapes (B=64, Ts=200, Tm=1000): CUDA events on the
launching stream, L2 flushed (512 MiB read sweep) before every launch, median of 9.  Prints achieved GB/s on the ALGORITHMIC
bytes of each kernel against the measured copy bandwidth of MEASURED_PEAKS.json.

  python tools/bench_membound.py [--json profiles/membound_rNN.json]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fs2b200 import sub  # noqa: E402

ops = sub("ops")
synth = sub("synthetic")
L = ops._L()
BF16 = torch.bfloat16


def timeit(fn, iters=9):
    flush = timeit.flush
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.sum()  # read-only sweep of 512 MiB: L2 ends up full of CLEAN lines (a write flush would make the
        #              timed kernel pay for evicting ~126 MB of dirty lines, which is not its own traffic)
        torch.cuda._sleep(400000)  # keep the GPU busy while the host enqueues e0 / kernel / e1: no launch gap is timed
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    timeit.flush = torch.zeros(128 << 20, dtype=torch.float32, device="cuda")
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    rows = []

    def rec(name, us, nbytes):
        gbs = nbytes / us / 1e3
        rows.append({"kernel": name, "us": round(us, 2), "algorithmic_MB": round(nbytes / 1e6, 2),
                     "GBps": round(gbs, 1), "frac_of_peak": round(gbs / peak, 3)})
        print("%-58s %8.1f us %8.1f MB %8.0f GB/s  %4.1f %%" % (name, us, nbytes / 1e6, gbs, 100 * gbs / peak))

    # size-matched yardstick: a plain device copy moving the same ~100 MB / ~330 MB as the kernels below (the
    # MEASURED_PEAKS figure is a 4 GiB copy; launch ramp and tail are a visible share of a 20-50 us kernel)
    for mb in (98, 328):
        src = torch.empty(mb * 500_000 // 2, dtype=torch.bfloat16, device="cuda")
        dst = torch.empty_like(src)
        us = timeit(lambda: dst.copy_(src))
        rec("yardstick: torch copy_ moving %d MB in total" % mb, us, 2 * src.numel() * 2)
    del src, dst

    B, T, C = 64, 1000, 256
    batch = synth.make_batch(**synth.CONFIGS["C2"])
    mel_lens = batch[7].clamp(max=T).cuda()
    valid = int(mel_lens.sum())
    full = torch.full((B,), T, dtype=torch.int64, device="cuda")
    x = torch.randn(B, T, C, device="cuda").to(BF16)
    res = torch.randn(B, T, C, device="cuda").to(BF16)
    dy = torch.randn(B, T, C, device="cuda").to(BF16)
    g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    dg, db, dbias = (torch.zeros(C, device="cuda") for _ in range(3))
    for tag, lens, nrows in (("dense", full, B * T), ("ragged C2 lens", mel_lens, valid)):
        for p in (0.0, 0.2):
            y, mean, rstd, keep = ops.ln_fwd(x, res, g, b, lens, p, 1, 5)
            us = timeit(lambda: ops.ln_fwd(x, res, g, b, lens, p, 1, 5))
            # reads x, res on valid rows; writes y on all rows
            rec("ln_fwd dropout+res+LN+padzero p=%.1f %s" % (p, tag), us, (2 * nrows + B * T) * C * 2)
            us = timeit(lambda: ops.ln_bwd(dy, x, res, g, mean, rstd, lens, p, 1, keep, dg, db, True, dbias=dbias))
            nout = 2 if p > 0 else 1
            rec("ln_bwd (+dgamma,dbeta,dbias) p=%.1f %s" % (p, tag), us, (3 * nrows + nout * B * T) * C * 2)

    # PostNet BatchNorm (M = B*T rows, C = 512), padded frames included by definition
    M, Cb = B * T, 512
    yb = torch.randn(M, Cb, device="cuda").to(BF16)
    stats = torch.zeros(2, Cb, device="cuda")
    ws = torch.empty(L.fs2_bn_workspace_floats(M, Cb), device="cuda")

    def bn_stats():
        L.fs2_bn_stats_bf16(yb.data_ptr(), M, Cb, ws.data_ptr(), stats.data_ptr(), 0.1, None, None, None, ops._st())

    bn_stats()
    gb, bb = torch.ones(Cb, device="cuda"), torch.zeros(Cb, device="cuda")
    ob = torch.empty_like(yb)
    seed = torch.zeros(1, dtype=torch.int64, device="cuda")
    dob = torch.randn(M, Cb, device="cuda").to(BF16)
    dst = torch.zeros(2, Cb, device="cuda")
    dyb = torch.empty_like(yb)
    keepb = torch.empty(M, Cb // 8, dtype=torch.uint8, device="cuda")
    us = timeit(bn_stats)
    rec("bn_stats [64000 x 512] (fixed-order partials + finalize)", us, M * Cb * 2)
    us = timeit(lambda: L.fs2_bn_apply_fwd(yb.data_ptr(), stats.data_ptr(), gb.data_ptr(), bb.data_ptr(), M, Cb, 1,
                                           0.5, 7, seed.data_ptr(), ob.data_ptr(), None, None, keepb.data_ptr(),
                                           ops._st()))
    rec("bn_apply + tanh + dropout 0.5", us, 2 * M * Cb * 2)
    us = timeit(lambda: L.fs2_bn_bwd(dob.data_ptr(), 0, yb.data_ptr(), stats.data_ptr(), gb.data_ptr(),
                                     bb.data_ptr(), M, Cb, 1, 0.5, 7, seed.data_ptr(), keepb.data_ptr(), ws.data_ptr(),
                                     dst.data_ptr(), None, None, dyb.data_ptr(), ops._st()))
    rec("bn_bwd (reduce + finalize + apply, 3 kernels)", us, 5 * M * Cb * 2)

    # bias-gradient column sums of the FFN hidden gradient [64000 x 1024]
    xx = torch.randn(B * T, 1024, device="cuda").to(BF16)
    out = torch.zeros(1024, device="cuda")
    us = timeit(lambda: ops.colsum(xx, out))
    rec("colsum [64000 x 1024] dense", us, B * T * 1024 * 2)
    us = timeit(lambda: ops.colsum(xx, out, lens=mel_lens, T=T))
    rec("colsum [64000 x 1024] ragged C2 lens", us, valid * 1024 * 2)

    # LengthRegulator: index + fused gather (+speaker row + sinusoid), backward segment sum
    Ts = 200
    dur = batch[11].cuda()
    xs = torch.randn(B, Ts, C, device="cuda").to(BF16)
    cum, idx, mel_len = ops.lr_index(dur, T)
    spk = torch.randn(B, C, device="cuda")
    pe = torch.randn(T + 1, C, device="cuda")
    outlr = torch.empty(B, T, C, device="cuda", dtype=BF16)
    us = timeit(lambda: ops.lr_index(dur, T))
    rec("lr_index (int64 cumsum + per-frame search)", us, B * Ts * 8 * 2 + B * T * 4)
    us = timeit(lambda: L.fs2_lr_gather_fused_bf16(xs.data_ptr(), idx.data_ptr(), spk.data_ptr(), pe.data_ptr(), B,
                                                   Ts, T, T, C, outlr.data_ptr(), ops._st()))
    rec("lr_gather_fused (+spk +PE)", us, (B * Ts + B * T) * C * 2 + B * T * 4)
    dxs = torch.empty_like(xs)
    us = timeit(lambda: L.fs2_lr_bwd_bf16(dy.data_ptr(), cum.data_ptr(), B, Ts, T, C, dxs.data_ptr(), ops._st()))
    rec("lr_bwd (segment sum)", us, (valid + B * Ts) * C * 2 + B * Ts * 8)

    # loss
    n_mel = 80
    mel = torch.randn(B, T, n_mel, device="cuda", requires_grad=True)
    post = torch.randn(B, T, n_mel, device="cuda", requires_grad=True)
    pp, ep, dp = (torch.randn(B, Ts, device="cuda", requires_grad=True) for _ in range(3))
    mel_t = batch[6][:, :T].cuda()
    src_lens = batch[4].cuda()

    pt, et = batch[9].cuda(), batch[10].cuda()

    def loss_fwd2():
        return ops.FastSpeech2LossFn.apply(mel, post, pp, ep, dp, mel_t, pt, et, dur, src_lens, mel_lens)

    us = timeit(loss_fwd2)
    rec("loss fwd (1 kernel + torch glue)", us, 3 * B * T * n_mel * 4 + 5 * B * Ts * 4)
    out6 = loss_fwd2()

    def loss_bwd():
        torch.autograd.grad(out6[0], (mel, post, pp, ep, dp), retain_graph=True)

    us = timeit(loss_bwd)
    rec("loss bwd", us, 5 * B * T * n_mel * 4)
    if "--json" in sys.argv:
        path = sys.argv[sys.argv.index("--json") + 1]
        json.dump({"peak_hbm_GBps": peak, "shapes": "C2: B=64, Ts=200, Tm=1000, valid mel frames %d" % valid,
                   "method": "CUDA events, 512 MiB read-sweep L2 flush before every launch, median of 9", "kernels": rows},
                  open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
