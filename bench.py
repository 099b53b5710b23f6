#!/usr/bin/env python
"""bench.py -- mel frames/s of one FastSpeech2 fwd + loss + bwd (+ gradient all-reduce) step.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
  python bench.py --impl reference ...                     (the reference algorithm on the host CPU)
  python bench.py --config C3|C5|C1|C4 ...                 (the other BASELINE.json configurations)
  python bench.py --balanced ...                           (N > 1: frame-balanced partition of the global batch)

Workload (BASELINE.json configs[1], SURVEY.md 8d "C2"): LibriTTS-shaped multi-speaker batch, 64
utterances PER GPU (weak scaling), src_len ~ U{20..200}, durations ~ U{1..11}, mel clipped to 1000
frames, bf16 operands / fp32 accumulation, dropout ON, train-mode BatchNorm, synthetic data and
seeded random-init weights.  `value` times graph replays with inputs resident in HBM; `e2e` times the
public TrainStep calls a training loop makes every step: run() on the batch prefetched during the previous
step, prefetch_batch() of the next one (pinned host -> device on a copy stream, overlapping the step), and
read_losses() (device -> host read of the six losses, synchronising).  One JSON line on stdout (rank 0).

Besides the contract keys the line carries: `roofline` (the kernel with the largest share of the step, timed on the
RAGGED call the step makes) + `roofline_kernels` (every tensor-bound kernel family of a decoder layer) +
`roofline_hbm` (LayerNorm, LengthRegulator, loss, BatchNorm against the measured HBM peak), `sustained` (>= 3 s of
back-to-back steps with clocks), `stock_gpu_baseline` (the reference algorithm in eager PyTorch on the same B200:
cuBLAS / cuDNN kernels, fp32 and bf16 autocast), `cpu_baseline`, and at N > 1 `dp_check_rel_err` (averaged flat
gradient vs the all_gather mean of the local gradients), `per_rank_frames`, `per_rank_solo_ms`, `exposed_comm_ms`.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mel_frames_per_sec_fastspeech2_fwd_bwd"
UNIT = "mel frames/s"
WORKLOADS = {
    "C1": "C1: LJSpeech-shaped single-speaker FastSpeech2 training step, batch 16 per GPU",
    "C2": "C2: LibriTTS-shaped multi-speaker FastSpeech2 training step, batch 64 per GPU",
    "C2_8": "C2/8: LibriTTS-shaped multi-speaker FastSpeech2 training step, batch 8 per GPU (global 64 on 8 GPUs)",
    "C3": "C3: few-shot FSCL query-batch step (AISHELL-3/KSS-shaped task, batch 8, averaged speaker), one task per GPU",
    "C4": "C4: long-utterance stress (220 phonemes, skewed durations, ~2000 frames truncated to 1500), batch 4",
    "C5": "C5: cross-lingual few-shot fine-tune step (CSS10-shaped), batch 4 per GPU, postnet on",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--balanced", action="store_true",
                    help="N > 1: deal the utterances of the global batch to the ranks so that every rank gets the "
                         "same number of mel frames (serpentine over the sorted lengths) instead of its own draw")
    ap.add_argument("--sustained-seconds", type=float, default=3.0)
    ap.add_argument("--quick", action="store_true", help="skip roofline / sustained / baseline extras")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


# --------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# --------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        if sm:
            out["sm_mhz"] = sm[len(sm) // 2]
            out["sm_min_mhz"] = sm[0]
            out["sm_max_mhz"] = max(int(r[1]) for r in rows if r[1].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i, n in enumerate(names):
            if any(len(r) > 2 + i and r[2 + i].strip().lower().startswith("active") for r in rows):
                out["reasons"].append(n)
        pw = []
        for r in rows:
            try:
                pw.append(float(r[6]))
            except Exception:
                pass
        if pw:
            out["power_w_max"] = max(pw)
        out["samples"] = len(sm)
        return out


# --------------------------------------------------------------------------------------------------
# the reference algorithm on the host CPU (oracle port; /root/reference does not exist on the box)
# --------------------------------------------------------------------------------------------------
def sub_batch(batch, n_utt):
    """the first n_utt utterances of a batch, re-padded to their own maximum lengths"""
    n_utt = min(n_utt, batch[3].shape[0])
    Ts = int(batch[4][:n_utt].max())
    Tm = int(batch[7][:n_utt].max())
    return (batch[0][:n_utt], batch[1][:n_utt], batch[2][:n_utt], batch[3][:n_utt, :Ts], batch[4][:n_utt], Ts,
            batch[6][:n_utt, :Tm], batch[7][:n_utt], Tm, batch[9][:n_utt, :Ts], batch[10][:n_utt, :Ts],
            batch[11][:n_utt, :Ts], batch[12][:n_utt])


def oracle_step_fn(config, batch, n_utt, device="cpu", autocast=False):
    """One fwd + loss + bwd of the reference algorithm (oracle/fs2_oracle.py: plain PyTorch restatement of the
    reference's modules, dropout off) on `device`.  Only the baseline legs of bench.py call this."""
    import torch

    from fs2b200 import sub
    from oracle import fs2_oracle  # the ONLY place bench.py runs oracle code: the baseline legs

    synth = sub("synthetic")
    M = sub("lightning.model")
    cfg, model, _, call_kw = synth.build_config(config, M)
    sd = {k: v.detach().to(device) for k, v in model.state_dict().items()}
    sb = synth.to_device(sub_batch(batch, n_utt), device)
    frames = int(torch.clamp(sb[7], max=cfg["max_seq_len"]).sum())
    params = {k: v for k, v in sd.items() if v.is_floating_point() and "position_enc" not in k
              and not k.endswith("_bins") and "running_" not in k}
    for v in params.values():
        v.requires_grad_(True)

    def step():
        for v in params.values():
            v.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = fs2_oracle.forward(sd, cfg, sb[2], sb[3], *sb[4:12], lang_args=sb[12], **call_kw)
            losses = fs2_oracle.loss(sb[:12], out)
        losses[0].backward()
        return losses[0]

    sample = "%d of the %d utterances of the same seeded batch (Ts=%d, Tm=%d)" % (
        sb[3].shape[0], batch[3].shape[0], int(sb[5]), int(sb[8]))
    return step, frames, sample


def run_reference(args):
    import torch

    from fs2b200 import sub

    synth = sub("synthetic")
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = synth.make_batch(**synth.CONFIGS[args.config])
    step, frames, sample = oracle_step_fn(args.config, batch, n_utt=16)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    dt = time.time() - t0
    val = frames * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.config],
                   "note": "reference algorithm (oracle port of the reference's PyTorch path, fp32, dropout off) on "
                           "the host CPU with all host threads; rank 0 only; /root/reference itself needs "
                           "pytorch_lightning / dlhlp_lib and does not exist on the GPU box (DESIGN.md section 7)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# algorithmic work
# --------------------------------------------------------------------------------------------------
def flops_fwd(B, Ts, Tm):
    """algorithmic forward FLOPs of one step on padded shapes."""
    return B * Ts * (4 * (5767168 + 1024 * Ts) + 2360832) + B * Tm * (6 * (5767168 + 1024 * Tm) + 40960 + 8683520)


def flops_fwd_ragged(src_lens, mel_lens, Ts, Tm):
    """Forward FLOPs needed for the reference's OUTPUTS: the FFT blocks zero every padded row after each
    sub-layer (transformer/Layers.py:25,28) and mask padded keys, so only the valid rows of every utterance
    do useful work there; the variance adaptor and the PostNet (BatchNorm statistics include padded frames)
    keep their padded shapes.  This is what the ragged tile schedule executes, up to tile rounding."""
    f = 0
    for ls in src_lens:
        ls = min(int(ls), Ts)
        f += ls * 4 * (5767168 + 1024 * ls)
    for lm in mel_lens:
        lm = min(int(lm), Tm)
        f += lm * 6 * (5767168 + 1024 * lm)
    B = len(src_lens)
    return f + B * Ts * 2360832 + B * Tm * (40960 + 8683520)


# --------------------------------------------------------------------------------------------------
# kernel-level rooflines, measured live on the calls the step makes (ragged C2 decoder shapes)
# --------------------------------------------------------------------------------------------------
class KTimer:
    """CUDA events on the launching stream (torch's current stream = the stream the ops launch on), L2 flushed with
    a read-only sweep before every launch, the GPU kept busy while the host enqueues, median of `iters`."""

    def __init__(self, iters=7):
        import torch

        self.torch = torch
        self.iters = iters
        self.flush = torch.zeros(96 << 20, dtype=torch.float32, device="cuda")  # 384 MiB > 126 MB L2

    def __call__(self, fn):
        torch = self.torch
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(self.iters):
            self.flush.sum()
            torch.cuda._sleep(300000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]


def kernel_rooflines(base, cfg, bf16_peak, hbm_peak, step_ms, traffic):
    """Times, alone, every heavy kernel of ONE decoder layer (and the HBM-bound kernels north_star names) on the
    ragged shapes of this rank's batch.  Returns (roofline, roofline_kernels, roofline_hbm)."""
    import torch

    from fs2b200 import sub

    ops, G = sub("ops"), sub("gemm")
    BF16 = torch.bfloat16
    dev = torch.device("cuda")
    timer = KTimer()
    t = cfg["transformer"]
    D, Dh, H = t["decoder_hidden"], t["conv_filter_size"], t["decoder_head"]
    k1, k2 = t["conv_kernel_size"]
    n_dec, n_enc = t["decoder_layer"], t["encoder_layer"]
    B = base[3].shape[0]
    Tm = min(int(base[8]), cfg["max_seq_len"])
    lens = torch.clamp(base[7], max=Tm).to(dev)
    V = int(lens.sum())  # valid decoder rows
    sq = float((lens.double() ** 2).sum())  # sum of len^2 (attention work)
    g = torch.Generator(device="cuda").manual_seed(3)
    rnd = lambda *s: (torch.randn(*s, device=dev, generator=g) * 0.5).to(BF16)
    x, dy = rnd(B, Tm, D), rnd(B, Tm, D)
    h, dh = rnd(B, Tm, Dh), rnd(B, Tm, Dh)
    qkv, dqkv = rnd(B, Tm, 3 * D), rnd(B, Tm, 3 * D)
    w1p, w2p = rnd(Dh, k1, D), rnd(D, k2, Dh)
    wqkv, wo = rnd(3 * D, D), rnd(D, D)
    b1, b2, bq, bo = (torch.zeros(n, device=dev) for n in (Dh, D, 3 * D, D))
    gw1 = torch.zeros(Dh, k1, D, device=dev).permute(0, 2, 1)  # [Co][k][Ci] storage, as in the gradient bucket
    gw2 = torch.zeros(D, Dh, 1, device=dev)
    gqkv = [[torch.zeros(D, D, device=dev)], [torch.zeros(D, device=dev)]] * 3
    gwo = torch.zeros(D, D, device=dev)
    hmask = torch.empty(B * Tm, Dh // 64, dtype=torch.int64, device=dev)
    x2, dy2, attn2 = x.view(B * Tm, D), dy.view(B * Tm, D), rnd(B * Tm, D)
    NT = ops.NO_TAIL

    fams = []  # (label, kernel, per-step launches in decoder layers, flops, callable)

    def fam(label, kernel, flops, fn, per_layer=1):
        ms = timer(fn)
        if "+" not in kernel and not kernel.startswith("attn"):  # the engine the descriptor was routed to
            kernel = ops._L().fs2_last_kernel().decode() or kernel
        fams.append({"label": label, "kernel": kernel, "avg_launch_ms": ms, "flops_per_launch": flops,
                     "achieved": flops / ms / 1e9, "launches_per_step": per_layer * n_dec,
                     "share_of_step": per_layer * n_dec * ms / step_ms})

    fam("FFN Conv1d k=%d %d->%d fwd (+bias+ReLU+mask)" % (k1, D, Dh), "conv_tc2_kernel", 2.0 * V * Dh * D * k1,
        lambda: ops.conv_fwd(x, w1p, b1, relu=True, lens=lens, tail=NT, relu_mask=hmask))
    fam("FFN Conv1d k=%d input-gradient (+residual)" % k1, "conv_tc2_kernel", 2.0 * V * Dh * D * k1,
        lambda: ops.conv_dgrad(dh, w1p, D, epilogue=G.EPI_ADD_AUX, aux=x, lens=lens))
    fam("FFN Conv1d k=%d weight-gradient" % k1, "gemm_tc2_kernel", 2.0 * V * Dh * D * k1,
        lambda: ops.conv_wgrad(dh, x, gw1, lens=lens))
    fam("FFN Conv1d k=%d %d->%d fwd" % (k2, Dh, D), "gemm_tc2_kernel", 2.0 * V * Dh * D * k2,
        lambda: ops.conv_fwd(h, w2p, b2, lens=lens, tail=NT))
    fam("FFN Conv1d k=%d input-gradient (+ReLU mask)" % k2, "gemm_tc2_kernel", 2.0 * V * Dh * D * k2,
        lambda: ops.conv_dgrad(dy, w2p, Dh, epilogue=G.EPI_RELU_BWD, lens=lens, tail=(k1 - 1) // 2, relu_mask=hmask))
    fam("FFN Conv1d k=%d weight-gradient" % k2, "gemm_tc2_kernel", 2.0 * V * Dh * D * k2,
        lambda: ops.conv_wgrad(dy, h, gw2, lens=lens))
    fam("QKV projection fwd (N=%d, K=%d)" % (3 * D, D), "gemm_tc2_kernel", 2.0 * V * 3 * D * D,
        lambda: ops.linear_fwd(x2, wqkv, bq, lens=lens, T=Tm, tail=NT))
    fam("QKV projection input-gradient (+residual)", "gemm_tc2_kernel", 2.0 * V * 3 * D * D,
        lambda: ops.linear_dgrad(dqkv.view(B * Tm, 3 * D), wqkv, epilogue=G.EPI_ADD_AUX, aux=x2, lens=lens, T=Tm))
    fam("QKV projection weight-gradient (+ Q / V bias column sums)", "gemm_tc2_kernel + colsum_kernel", 2.0 * V * 3 * D * D,
        lambda: ops.qkv_param_grads(dqkv.view(B * Tm, 3 * D), x2, gqkv, D, lens=lens, T=Tm))
    fam("output projection fwd", "gemm_tc2_kernel", 2.0 * V * D * D,
        lambda: ops.linear_fwd(attn2, wo, bo, lens=lens, T=Tm, tail=NT))
    fam("output projection input-gradient", "gemm_tc2_kernel", 2.0 * V * D * D,
        lambda: ops.linear_dgrad(dy2, wo, lens=lens, T=Tm, tail=NT))
    fam("output projection weight-gradient", "gemm_tc2_kernel", 2.0 * V * D * D,
        lambda: ops.linear_wgrad(dy2, attn2, gwo, lens=lens, T=Tm))
    dk = D // H
    sched = ops.attn_schedule(lens, Tm, H)  # longest-first work order, as transformer/Models.py::_run_layers builds it
    o3, lse = ops.attn_fwd(qkv, lens, H, dk, sched)
    fam("fused attention fwd (QK^T, softmax, PV)", "attn_fwd2_kernel", 4.0 * sq * D,
        lambda: ops.attn_fwd(qkv, lens, H, dk, sched))
    fam("fused attention bwd (dK/dV + dQ kernels)", "attn_bwd_prep + attn_bwd_dkv2 + attn_bwd_dq2", 8.0 * sq * D,
        lambda: ops.attn_bwd(qkv, o3, dy, lse, lens, H, dk, sched))
    for f in fams:
        f["bound"], f["peak"], f["unit"] = "tensor", bf16_peak, "TFLOP/s"
        f["frac"] = f["achieved"] / bf16_peak
        f["traffic"] = traffic.get(f["label"])
    # the dominant KERNEL: entries that time several kernels in one call (attention backward = row-dot prologue + dK/dV
    # + dQ; QKV weight gradient + column sums) explain the step below but are not one kernel's roofline
    top = max((f for f in fams if "+" not in f["kernel"]), key=lambda f: f["share_of_step"])
    roof = dict(top)
    roof["note"] = ("largest share of the step among the single-kernel launches of a decoder layer (share = launches per step "
                    "x this time / step time); timed ALONE on the ragged call the step makes (valid rows %d of %d "
                    "padded), L2 flushed, CUDA events on the launching stream; flops = valid rows only" % (V, B * Tm))

    # ---- HBM-bound kernels north_star names: LayerNorm (+dropout+residual), LengthRegulator, losses; + BatchNorm
    hb = []

    def rec(label, nbytes, fn, launches):
        ms = timer(fn)
        gbs = nbytes / ms / 1e6
        hb.append({"label": label, "bound": "hbm", "avg_launch_ms": ms, "bytes_per_launch": nbytes, "achieved": gbs,
                   "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "launches_per_step": launches,
                   "share_of_step": launches * ms / step_ms, "traffic": traffic.get(label)})

    p = t["decoder_dropout"]
    gamma, beta = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    dg, db, dbias = (torch.zeros(D, device=dev) for _ in range(3))
    y, mean, rstd, keep = ops.ln_fwd(x, dy, gamma, beta, lens, p, 1, 5)
    rec("LayerNorm fwd (dropout + residual + LN + pad-zero)", (2 * V + B * Tm) * D * 2,
        lambda: ops.ln_fwd(x, dy, gamma, beta, lens, p, 1, 5), 2 * n_dec)
    rec("LayerNorm bwd (+dgamma/dbeta/dbias, stored keep bits)", (3 * V + 2 * B * Tm) * D * 2,
        lambda: ops.ln_bwd(dy, x, dy, gamma, mean, rstd, lens, p, 1, keep, dg, db, True, dbias=dbias), 2 * n_dec)
    Ts = int(base[5])
    dur = base[11].to(dev)
    Tm_full = int(base[8])
    xs = rnd(B, Ts, D)
    spk = torch.randn(B, D, device=dev)
    pe = torch.randn(Tm + 1, D, device=dev)
    cum, idx, mel_len = ops.lr_index(dur, Tm_full)
    rec("LengthRegulator index (cumsum + search)", 8 * B * Ts + 4 * B * Tm_full + 8 * B * Ts,
        lambda: ops.lr_index(dur, Tm_full), 1)
    rec("LengthRegulator gather (+speaker row +sinusoid)", (B * Ts + B * Tm) * D * 2 + 4 * B * Tm,
        lambda: ops.LengthRegulate.apply(xs, cum, idx, Tm_full, Tm, spk, pe), 1)
    dout = rnd(B, Tm, D)
    dx = torch.empty(B, Ts, D, dtype=BF16, device=dev)
    rec("LengthRegulator bwd (segment sum)", (B * Ts + B * Tm) * D * 2 + 8 * B * Ts,
        lambda: ops._ck(ops._L().fs2_lr_bwd_bf16(dout.data_ptr(), cum.data_ptr(), B, Ts, Tm, D, dx.data_ptr(),
                                                 ops._st()), "lr_bwd"), 1)
    n_mel = 80
    mel = torch.randn(B, Tm, n_mel, device=dev)
    post = torch.randn(B, Tm, n_mel, device=dev)
    mel_t = base[6].to(dev)
    pp, ep, dp = (torch.randn(B, Ts, device=dev) for _ in range(3))
    src_lens = base[4].to(dev)
    pt, et = base[9].to(dev), base[10].to(dev)
    lf = lambda: ops.FastSpeech2LossFn.apply(mel, post, pp, ep, dp, mel_t, pt, et, dur, src_lens, lens)
    rec("masked loss fwd (5 terms, one kernel)", 3 * V * n_mel * 4 + 5 * B * Ts * 4, lf, 1)
    melg, postg = mel.clone().requires_grad_(), post.clone().requires_grad_()
    out = ops.FastSpeech2LossFn.apply(melg, postg, pp, ep, dp, mel_t, pt, et, dur, src_lens, lens)
    rec("masked loss bwd", 3 * V * n_mel * 4 + 2 * B * Tm * n_mel * 4 + 5 * B * Ts * 4,
        lambda: torch.autograd.grad(out[0], (melg, postg), retain_graph=True), 1)
    C = 512
    M = B * Tm
    yb, dob = rnd(B, Tm, C), rnd(B, Tm, C)
    L = ops._L()
    ws = torch.empty(L.fs2_bn_workspace_floats(M, C), device=dev)
    stats, dstats = torch.empty(2, C, device=dev), torch.empty(2, C, device=dev)
    gb, bb = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    nxt, dyb = torch.empty_like(yb), torch.empty_like(yb)
    keepb = torch.empty(M, C // 8, dtype=torch.uint8, device=dev)
    seed = ops._Rng.tensor(dev)
    rec("BatchNorm statistics (fixed-order, + finalize)", M * C * 2,
        lambda: L.fs2_bn_stats_bf16(yb.data_ptr(), M, C, ws.data_ptr(), stats.data_ptr(), 0.1, None, None, None,
                                    ops._st()), 3)
    rec("BatchNorm apply + tanh + dropout", 2 * M * C * 2,
        lambda: L.fs2_bn_apply_fwd(yb.data_ptr(), stats.data_ptr(), gb.data_ptr(), bb.data_ptr(), M, C, 1, 0.5, 7,
                                   seed.data_ptr(), nxt.data_ptr(), None, None, keepb.data_ptr(), ops._st()), 3)
    rec("BatchNorm bwd (reduce + finalize + apply)", 5 * M * C * 2,
        lambda: L.fs2_bn_bwd(dob.data_ptr(), 0, yb.data_ptr(), stats.data_ptr(), gb.data_ptr(), bb.data_ptr(), M, C,
                             1, 0.5, 7, seed.data_ptr(), keepb.data_ptr(), ws.data_ptr(), dstats.data_ptr(), None, None,
                             dyb.data_ptr(), ops._st()), 3)
    return roof, fams, hb


def time_optimizer(rt, step, hbm_peak):
    """sumsq + Adam + advance on the 34.5 M-parameter flat buffers; call only after the last graph replay
    (FusedAdam re-homes the parameters into its flat buffer)."""
    import torch

    cfg = {"scheduler_type": "sqrt", "optimizer": {"betas": [0.9, 0.98], "eps": 1e-9, "weight_decay": 0.0,
                                                  "grad_clip_thresh": 1.0, "warm_up_step": 4000,
                                                  "anneal_steps": [30000, 40000, 50000], "anneal_rate": 0.3}}
    opt = rt.FusedAdam(step.buckets, train_config=cfg)  # config/train/baseline.yaml
    n = step.buckets.flat.numel()
    for _ in range(3):
        opt.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = 10
    torch.cuda._sleep(2000000)
    e0.record()
    for _ in range(k):
        opt.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / k
    nbytes = 8 * 4 * n  # g (norm) + p, g, m, v reads + p, m, v writes
    return {"ms_per_step": ms, "algorithmic_MB": nbytes / 1e6, "GBps": nbytes / ms / 1e6,
            "frac_of_hbm_peak": nbytes / ms / 1e6 / hbm_peak, "kernels": "sumsq + adam_step + advance",
            "note": "not included in `value` (the metric is fwd+bwd); pass optimizer= to TrainStep to capture it"}


def stock_gpu_baseline(config, base, n_steps=3):
    """The reference algorithm in eager PyTorch on THIS GPU (cuBLAS / cuDNN / ATen kernels; none of ours): fp32 with
    TF32 off, and under bf16 autocast.  Dropout off (favours this leg).  A baseline, never part of `value`."""
    import torch

    out = {"note": "oracle port of the reference's PyTorch modules run eagerly on the same B200 (library kernels "
                   "only, dropout off); CUDA-event timed, 1 warm-up + %d steps" % n_steps}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        for key, ac in (("fp32", False), ("bf16_autocast", True)):
            try:
                step, frames, sample = oracle_step_fn(config, base, n_utt=base[3].shape[0], device="cuda", autocast=ac)
                step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(n_steps):
                    step()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / n_steps
                out[key + "_ms"] = ms
                out[key + "_frames_per_s"] = frames / ms * 1e3
                out["sample"] = sample
                del step
                torch.cuda.empty_cache()
            except Exception as e:  # e.g. out of memory for the fp32 attention scores
                out[key + "_error"] = repr(e)[:200]
                torch.cuda.empty_cache()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    return out


def dbg(msg):
    if os.environ.get("FS2_DEBUG"):
        print("[rank %s %.1f] %s" % (os.environ.get("RANK", "0"), time.time() % 1000, msg), file=sys.stderr, flush=True)


def load_traffic():
    """DRAM bytes per launch from the committed ncu --set full captures (profiles/r2_kernel_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "r2_kernel_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def make_rank_batch(synth, config, rank, world, balanced):
    """This rank's batch.  Default: every rank draws its own utterances (seed + 1000 * rank).  --balanced: the ranks'
    draws are pooled into the global batch and dealt back serpentine over the sorted mel lengths, so that every
    rank holds the same number of utterances and (nearly) the same number of mel frames -- the reference's own split
    is an arbitrary `batch_size // device_count` (FastSpeech2DataModule.py:102)."""
    import torch

    kw = dict(synth.CONFIGS[config])
    if not balanced or world == 1:
        kw["seed"] = kw["seed"] + 1000 * rank
        return synth.make_batch(**kw)
    pool = []
    for r in range(world):
        k = dict(kw)
        k["seed"] = kw["seed"] + 1000 * r
        pool.append(synth.make_batch(**k))
    items = [(int(b[7][i]), r, i) for r, b in enumerate(pool) for i in range(b[3].shape[0])]
    items.sort(reverse=True)
    mine = []
    for j, it in enumerate(items):
        rnd, pos = divmod(j, world)
        owner = pos if rnd % 2 == 0 else world - 1 - pos
        if owner == rank:
            mine.append(it)
    Ts = max(int(pool[r][4][i]) for _, r, i in mine)
    Tm = max(int(pool[r][7][i]) for _, r, i in mine)

    def gather(slot, T=None):
        rows = []
        for _, r, i in mine:
            t = pool[r][slot][i]
            if T is not None:
                pad = T - min(t.shape[0], T)
                t = t[:T]
                if pad:
                    t = torch.cat([t, t.new_zeros((pad,) + tuple(t.shape[1:]))])
            rows.append(t)
        return torch.stack(rows)

    ids = ["utt%04d" % j for j in range(len(mine))]
    return (ids, ["" for _ in ids], gather(2), gather(3, Ts), gather(4), Ts, gather(6, Tm), gather(7), Tm,
            gather(9, Ts), gather(10, Ts), gather(11, Ts), gather(12))


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from fs2b200 import sub
    synth = sub("synthetic")

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION) off it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    hbm, bf16_burst, bf16_sus, peak_src = peaks()
    cabi, ops, rt = sub("_cabi"), sub("ops"), sub("runtime")
    M = sub("lightning.model")

    cfg, model, loss_fn, call_kw = synth.build_config(args.config, M, device=dev)
    base = make_rank_batch(synth, args.config, rank, world, args.balanced)
    B, Ts, Tm = base[3].shape[0], int(base[5]), int(base[8])
    frames = synth.count_real_frames(base, cfg["max_seq_len"])
    dbg("model built, B=%d Ts=%d Tm=%d" % (B, Ts, Tm))

    # ---- N > 1: a second captured step WITHOUT the reduction (same dropout salts): local gradients for the
    # gradient-average check and this rank's solo step time (what the step costs with no collective at all)
    solo = None
    salt0, seed0 = 1234 + rank, None
    if world > 1 and not args.no_graph:
        ops.manual_seed(salt0, dev)
        lb = rt.GradBuckets(model.parameters(), device=dev)
        lb.world = 1
        solo = rt.TrainStep(model, loss_fn, base, use_graph=not args.no_graph, buckets=lb, device=dev,
                            model_kwargs=call_kw)
    ops.manual_seed(salt0, dev)  # identical per-call-site salts in both captures
    n0 = cabi.launch_count()
    step = rt.TrainStep(model, loss_fn, base, use_graph=not args.no_graph, device=dev, model_kwargs=call_kw)
    launches_total = cabi.launch_count() - n0
    dbg("TrainStep ready")
    launches_per_step = launches_total // 3 if not args.no_graph else None  # 2 warm-up bodies + 1 captured

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dp_check = None
    if solo is not None:
        # averaged flat gradient (the timed graph, NCCL buckets overlapped with the backward) against the mean of
        # the ranks' LOCAL gradients exchanged with a plain all_gather; same dropout counter on both replays
        seed = ops._Rng.tensor(dev)
        c0 = seed.clone()
        step.run()
        torch.cuda.synchronize()
        avg = step.buckets.flat.clone()
        seed.copy_(c0)
        solo.run()
        torch.cuda.synchronize()
        loc = solo.buckets.flat.clone()
        gathered = [torch.empty_like(loc) for _ in range(world)]
        dist.all_gather(gathered, loc)
        mean = torch.stack(gathered).mean(0)
        err = ((avg - mean).norm() / mean.norm()).reshape(1)
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        dp_check = float(err[0])
        del gathered, mean, avg, loc
        if not (dp_check <= 1e-3):
            if rank == 0:
                print(json.dumps({"error": "data-parallel gradient check failed", "dp_check_rel_err": dp_check}),
                      flush=True)
            barrier()
            step.close()
            solo.close()
            dist.destroy_process_group()
            sys.exit(3)

    # three host variants of the batch with the same padded shape (utterance order permuted)
    variants = []
    g = torch.Generator().manual_seed(7 + rank)
    for v in range(3):
        perm = torch.randperm(B, generator=g) if v else torch.arange(B)
        hv = {}
        for i in rt.step._TENSOR_SLOTS:
            t = base[i][perm]
            hv[i] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            hv[i].copy_(t)
        variants.append(hv)

    def timed(fn, k, sync_ranks=True):
        if sync_ranks:
            barrier()
        else:
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        if sync_ranks:
            barrier()
        else:
            torch.cuda.synchronize()
        wall = (time.time() - t0) * 1e3
        ms = torch.tensor([e0.elapsed_time(e1), wall], device=dev)
        if world > 1 and sync_ranks:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0]), float(ms[1])

    W = max(args.warmup, 3)
    for i in range(W):
        step.run()
    dbg("warm-up done")
    clocks = Clocks(local) if rank == 0 else None
    n_before = cabi.launch_count()
    dev_ms, wall_ms = timed(lambda i: step.run(), args.steps)
    eager_launches = cabi.launch_count() - n_before
    dbg("timed loop done %.2f ms" % dev_ms)

    def e2e_iter_sync(i):
        # public API, one training step: run on the batch prefetched during the previous step, start the
        # pinned-host -> device copy of the next batch (overlaps this step), read this step's six losses back
        # and WAIT for them (the host enqueues the next step only after this one has finished)
        step.run()
        step.prefetch_batch(pinned=variants[(i + 1) % 3])
        step.read_losses()

    def e2e_iter(i):
        # the same three calls with asynchronous logging: the 24-byte read of this step's losses is enqueued and
        # the host picks up the PREVIOUS step's values, so step i+1 is enqueued while step i runs
        step.run()
        step.prefetch_batch(pinned=variants[(i + 1) % 3])
        step.read_losses_lagged()

    step.prefetch_batch(pinned=variants[0])
    for i in range(2):
        e2e_iter_sync(i)
    e2e_ms, e2e_wall = timed(e2e_iter_sync, args.steps)
    for i in range(2):
        e2e_iter(i)
    e2e_lag_ms, e2e_lag_wall = timed(e2e_iter, args.steps)
    clk = clocks.stop() if clocks else None
    losses = step.read_losses().tolist()

    # ---- sustained: >= N seconds of back-to-back replays (clocks under sustained load, not the burst of 20 steps)
    sustained = None
    if not args.quick and args.sustained_seconds > 0:
        n_sus = max(int(args.sustained_seconds * 1e3 / (dev_ms / args.steps)) + 1, args.steps)
        if world > 1:
            t = torch.tensor([n_sus], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            n_sus = int(t[0])
        clocks2 = Clocks(local) if rank == 0 else None
        sus_ms, _ = timed(lambda i: step.run(), n_sus)
        clk2 = clocks2.stop() if clocks2 else None
        sustained = {"steps": n_sus, "seconds": sus_ms / 1e3, "ms_per_step": sus_ms / n_sus, "clocks": clk2}

    # ---- per-rank solo time (no collective in the graph) and frames
    per_rank = None
    if solo is not None:
        for i in range(3):
            solo.run()
        solo_ms, _ = timed(lambda i: solo.run(), args.steps, sync_ranks=False)
        t = torch.tensor([float(frames), solo_ms / args.steps], device=dev)
        allr = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
        per_rank = {"frames": [int(a[0]) for a in allr], "solo_ms": [float(a[1]) for a in allr]}

    tot = torch.tensor([float(frames)], device=dev)
    if world > 1:
        dist.all_reduce(tot)
    total_frames = float(tot[0])
    value = total_frames * args.steps / (dev_ms / 1e3)
    e2e_val = total_frames * args.steps / (max(e2e_ms, e2e_wall) / 1e3)
    launches = (launches_per_step * args.steps) if launches_per_step else eager_launches

    if rank == 0:
        Tm_eff = min(Tm, cfg["max_seq_len"])
        fl_padded = 3.0 * flops_fwd(B, Ts, Tm_eff)
        fl = 3.0 * flops_fwd_ragged(base[4].tolist(), base[7].tolist(), Ts, Tm_eff)
        step_ms = dev_ms / args.steps
        step_tf = fl / step_ms / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.config], "per_gpu_batch": B, "global_batch": B * world, "Ts_pad": Ts,
                       "Tm_pad": Tm, "real_mel_frames_per_gpu": frames, "parallelism": "dp%d" % world,
                       "dropout": "on", "cuda_graph": not args.no_graph, "padded_frames": "skipped (ragged tiles)",
                       "partition": "frame-balanced (serpentine over sorted mel lengths)" if args.balanced
                       else "independent draw per rank",
                       "l2": "step working set (~GBs of activations) exceeds the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": step.h2d_bytes,
                    "d2h_bytes_per_step": step.d2h_bytes, "ms_per_step": max(e2e_ms, e2e_wall) / args.steps,
                    "loss_readback": "blocking: the host waits for the six losses of every step before it enqueues "
                                     "the next one",
                    # the same loop with TrainStep.read_losses_lagged (losses picked up one step late, the host
                    # never waits for the step it has just launched): no idle gaps between steps -- and therefore
                    # the clocks of a sustained run rather than of a burst
                    "ms_per_step_lagged_readback": max(e2e_lag_ms, e2e_lag_wall) / args.steps},
            "gpu_launches": launches, "launches_per_step": launches_per_step,
            "step_tensor": {"algorithmic_tflop_per_step": fl / 1e12, "padded_shape_tflop_per_step": fl_padded / 1e12,
                            "achieved_tflops": step_tf, "frac_of_sustained_peak": step_tf / bf16_sus,
                            "peak_sustained": bf16_sus,
                            "note": "algorithmic = valid rows only in the FFT blocks (padded frames are skipped "
                                    "by the ragged tile schedule); padded_shape = SURVEY.md 8d formula"},
            "clocks": clk, "losses": losses, "wall_ms_per_step": wall_ms / args.steps,
        }
        if sustained:
            line["sustained"] = sustained
            line["sustained"]["frames_per_s"] = total_frames / (sustained["ms_per_step"] / 1e3)
        if world > 1:
            line["dp_check_rel_err"] = dp_check
        if per_rank is not None:
            line["per_rank_frames"] = per_rank["frames"]
            line["per_rank_solo_ms"] = per_rank["solo_ms"]
            line["exposed_comm_ms"] = step_ms - max(per_rank["solo_ms"])
            line["collective"] = "NCCL all_reduce(AVG) per 32 MiB fp32 bucket on a side stream, captured in the graph"
        if not args.quick:
            try:
                roof, fams, hb = kernel_rooflines(base, cfg, bf16_burst, hbm, step_ms, load_traffic())
                src = "MEASURED_PEAKS.json (bf16_tflops burst: kernels timed alone; hbm_gbs)" if peak_src == "measured" \
                    else "fallback 1590 TFLOP/s / 6650 GB/s"
                roof["peak_source"] = src
                line["roofline"] = roof
                line["roofline_kernels"] = fams
                line["roofline_hbm"] = hb
            except Exception as e:
                line["roofline"] = {"error": repr(e)[:300]}
            try:  # row 8f-1: fused clip + Adam + LR schedule on the flat buffers (NOT part of `value`)
                line["optimizer_step"] = time_optimizer(rt, step, hbm)
            except Exception as e:
                line["optimizer_step"] = {"error": repr(e)}
        if world == 1 and not args.quick:
            try:
                line["stock_gpu_baseline"] = stock_gpu_baseline(args.config, base)
            except Exception as e:
                line["stock_gpu_baseline"] = {"error": repr(e)[:300]}
        if world == 1 and not args.no_cpu_baseline and not args.quick:
            try:
                torch.set_num_threads(os.cpu_count() or 1)
                cstep, cframes, sample = oracle_step_fn(args.config, base, n_utt=16)
                cstep()
                t0 = time.time()
                n = 4
                for _ in range(n):
                    cstep()
                dt = time.time() - t0
                line["cpu_baseline"] = {"value": cframes * n / dt, "unit": UNIT, "cores": os.cpu_count(),
                                        "kind": "port", "sample": sample + ", fp32, dropout off, 1 warm-up + 4 timed "
                                        "steps (the unmodified reference needs pytorch_lightning / dlhlp_lib and is "
                                        "not on the GPU box; the port is pinned to it by tests/golden)"}
            except Exception as e:  # the baseline must never take the GPU number down with it
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                        "sample": "failed: %r" % (e,)}
        print(json.dumps(line), flush=True)
    if world > 1:
        # drop the captured graphs (and their NCCL kernels) before the process group goes away: a clean shutdown
        barrier()
        step.close()
        if solo is not None:
            solo.close()
        del step, solo
        import gc
        gc.collect()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
