#!/usr/bin/env python
"""bench.py -- mel frames/s of one FastSpeech2 fwd + loss + bwd (+ gradient all-reduce) step.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
  python bench.py --impl reference ...                     (the reference algorithm on the host CPU)

Workload (BASELINE.json configs[1], SURVEY.md 8d "C2"): LibriTTS-shaped multi-speaker batch, 64
utterances PER GPU (weak scaling), src_len ~ U{20..200}, durations ~ U{1..11}, mel clipped to 1000
frames, bf16 operands / fp32 accumulation, dropout ON, train-mode BatchNorm, synthetic data and
seeded random-init weights.  `value` times graph replays with inputs resident in HBM; `e2e` times the
public TrainStep calls a training loop makes every step: run() on the batch prefetched during the previous
step, prefetch_batch() of the next one (pinned host -> device on a copy stream, overlapping the step), and
read_losses() (device -> host read of the six losses, synchronising).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mel_frames_per_sec_fastspeech2_fwd_bwd"
UNIT = "mel frames/s"
WORKLOAD = "C2: LibriTTS-shaped multi-speaker FastSpeech2 training step, batch 64 per GPU"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

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
            out["sm_max_mhz"] = max(int(r[1]) for r in rows if r[1].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i, n in enumerate(names):
            if any(len(r) > 2 + i and r[2 + i].strip().lower().startswith("active") for r in rows):
                out["reasons"].append(n)
        out["samples"] = len(sm)
        return out


# --------------------------------------------------------------------------------------------------
# the reference algorithm on the host CPU (oracle port; /root/reference does not exist on the box)
# --------------------------------------------------------------------------------------------------
def cpu_reference_step_fn(cfg, batch, n_utt):
    import torch

    from fs2b200 import sub
    from oracle import fs2_oracle  # the ONLY place bench.py runs oracle code: the CPU baseline legs

    synth = sub("synthetic")
    M = sub("lightning.model")
    spk = {"emb_type": "table", "speakers": list(range(247))} if cfg.get("multi_speaker") else None
    tmpl = (M.FastSpeech2(cfg, spk_config=spk) if spk else M.FastSpeech2(cfg)).state_dict()
    sd = synth.init_state_dict(tmpl, 0)
    # bounded sample: the first n_utt utterances of the same batch (re-padded to their own max)
    Ts = int(batch[4][:n_utt].max())
    Tm = int(batch[7][:n_utt].max())
    sub_b = (batch[0][:n_utt], batch[1][:n_utt], batch[2][:n_utt], batch[3][:n_utt, :Ts], batch[4][:n_utt], Ts,
             batch[6][:n_utt, :Tm], batch[7][:n_utt], Tm, batch[9][:n_utt, :Ts], batch[10][:n_utt, :Ts],
             batch[11][:n_utt, :Ts], batch[12][:n_utt])
    frames = int(torch.clamp(sub_b[7], max=cfg["max_seq_len"]).sum())

    def step():
        _, losses, grads = fs2_oracle.step(sd, cfg, sub_b)
        return float(losses[0])

    return step, frames, "first %d of the %d utterances of the same seeded batch (Ts=%d, Tm=%d)" % (
        n_utt, batch[3].shape[0], Ts, Tm)


def run_reference(args):
    import torch

    from fs2b200 import sub

    synth = sub("synthetic")
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = synth.model_cfg(multi_speaker=True)
    batch = synth.make_batch(**synth.CONFIGS[args.config])
    step, frames, sample = cpu_reference_step_fn(cfg, batch, n_utt=16)
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
        "config": {"workload": WORKLOAD, "note": "reference algorithm (oracle port of the reference's PyTorch "
                   "path, fp32, dropout off) on the host CPU; rank 0 only"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
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


def dominant_kernel_roofline(B, Tm, bf16_peak):
    """The step's dominant kernel on its largest instance: the decoder's k=9 Conv1d (256 -> 1024) forward,
    dense (all 64 x 1000 rows), i.e. conv_tc2_kernel (2-CTA tiles + activation-halo reuse)."""
    import torch

    from fs2b200 import sub

    ops = sub("ops")
    x = torch.randn(B, Tm, 256, device="cuda").to(torch.bfloat16)
    wp = torch.randn(1024, 9, 256, device="cuda").to(torch.bfloat16)
    bias = torch.zeros(1024, device="cuda")
    flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ops.conv_fwd(x, wp, bias, relu=True)
    ts = []
    for _ in range(10):
        flush.zero_()
        torch.cuda._sleep(400000)  # the GPU spins while the host enqueues e0 / kernel / e1 (no launch gap timed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv_fwd(x, wp, bias, relu=True)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sum(ts) / len(ts)
    fl = 2.0 * B * Tm * 1024 * 2304
    ach = fl / ms / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            pass
    return {"bound": "tensor", "achieved": ach, "peak": bf16_peak, "unit": "TFLOP/s", "frac": ach / bf16_peak,
            "traffic": traffic, "kernel": "conv_tc2_kernel decoder Conv1d k=9 256->1024 fwd, dense (M=%d)" % (B * Tm),
            "avg_launch_ms": ms, "flops_per_launch": fl}


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


def dbg(msg):
    if os.environ.get("FS2_DEBUG"):
        print("[rank %s %.1f] %s" % (os.environ.get("RANK", "0"), time.time() % 1000, msg), file=sys.stderr, flush=True)


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
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
        dist.init_process_group("nccl", device_id=dev)
    hbm, bf16_burst, bf16_sus, peak_src = peaks()
    cabi, ops, rt = sub("_cabi"), sub("ops"), sub("runtime")
    M = sub("lightning.model")

    cfg = synth.model_cfg(multi_speaker=True)
    spk = {"emb_type": "table", "speakers": list(range(247))}
    torch.manual_seed(0)
    model = M.FastSpeech2(cfg, spk_config=spk)
    model.load_state_dict(synth.init_state_dict(model.state_dict(), 0))
    model = model.to(dev).train()
    loss_fn = M.FastSpeech2Loss(cfg)
    ops.manual_seed(1234 + rank, dev)

    kw = dict(synth.CONFIGS[args.config])
    kw["seed"] = kw["seed"] + 1000 * rank  # every rank draws its own utterances (weak scaling)
    base = synth.make_batch(**kw)
    if world > 1:  # a common padded shape is not required across ranks; each rank owns its graph
        pass
    B, Ts, Tm = base[3].shape[0], int(base[5]), int(base[8])
    frames = synth.count_real_frames(base, cfg["max_seq_len"])

    dbg("model built, B=%d Ts=%d Tm=%d" % (B, Ts, Tm))
    n0 = cabi.launch_count()
    step = rt.TrainStep(model, loss_fn, base, use_graph=not args.no_graph, device=dev)
    launches_total = cabi.launch_count() - n0
    dbg("TrainStep ready")
    launches_per_step = launches_total // 3 if not args.no_graph else None  # 2 warm-up bodies + 1 captured

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

    def load(v):
        for i, t in variants[v % 3].items():
            step.static[i].copy_(t, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        barrier()
        wall = (time.time() - t0) * 1e3
        ms = torch.tensor([e0.elapsed_time(e1), wall], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0]), float(ms[1])

    load(0)
    for i in range(max(args.warmup, 3)):
        step.run()
    dbg("warm-up done")
    clocks = Clocks(local) if rank == 0 else None
    n_before = cabi.launch_count()
    dev_ms, wall_ms = timed(lambda i: step.run(), args.steps)
    eager_launches = cabi.launch_count() - n_before
    dbg("timed loop done %.2f ms" % dev_ms)

    def e2e_iter(i):
        # public API, one training step: run on the batch prefetched during the previous step, start the
        # pinned-host -> device copy of the next batch (overlaps this step), read this step's six losses back
        step.run()
        step.prefetch_batch(pinned=variants[(i + 1) % 3])
        step.read_losses()

    step.prefetch_batch(pinned=variants[0])
    for i in range(2):
        e2e_iter(i)
    e2e_ms, e2e_wall = timed(e2e_iter, args.steps)
    clk = clocks.stop() if clocks else None
    losses = step.read_losses().tolist()

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
        step_tf = fl / (dev_ms / args.steps) / 1e9
        roof = dominant_kernel_roofline(B, min(Tm, cfg["max_seq_len"]), bf16_burst)
        roof["peak_source"] = "MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone)" if peak_src == "measured" \
            else "fallback 1590 TFLOP/s"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": B, "global_batch": B * world, "Ts_pad": Ts,
                       "Tm_pad": Tm, "real_mel_frames_per_gpu": frames, "parallelism": "dp%d" % world,
                       "dropout": "on", "cuda_graph": not args.no_graph, "padded_frames": "skipped (ragged tiles)",
                       "l2": "step working set (~GBs of activations) exceeds the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": step.h2d_bytes,
                    "d2h_bytes_per_step": step.d2h_bytes, "ms_per_step": max(e2e_ms, e2e_wall) / args.steps},
            "gpu_launches": launches, "launches_per_step": launches_per_step,
            "roofline": roof,
            "step_tensor": {"algorithmic_tflop_per_step": fl / 1e12, "padded_shape_tflop_per_step": fl_padded / 1e12,
                            "achieved_tflops": step_tf, "frac_of_sustained_peak": step_tf / bf16_sus,
                            "peak_sustained": bf16_sus,
                            "note": "algorithmic = valid rows only in the FFT blocks (padded frames are skipped "
                                    "by the ragged tile schedule); padded_shape = SURVEY.md 8d formula"},
            "clocks": clk, "losses": losses, "wall_ms_per_step": wall_ms / args.steps,
        }
        try:  # row 8f-1: fused clip + Adam + LR schedule on the flat buffers, timed on its own (NOT part of `value`)
            line["optimizer_step"] = time_optimizer(rt, step, hbm)
        except Exception as e:
            line["optimizer_step"] = {"error": repr(e)}
        if world == 1 and not args.no_cpu_baseline:
            try:
                torch.set_num_threads(os.cpu_count() or 1)
                cstep, cframes, sample = cpu_reference_step_fn(cfg, base, n_utt=16)
                cstep()
                t0 = time.time()
                n = 4
                for _ in range(n):
                    cstep()
                dt = time.time() - t0
                line["cpu_baseline"] = {"value": cframes * n / dt, "unit": UNIT, "cores": os.cpu_count(),
                                        "kind": "port", "sample": sample + ", 1 warm-up + 4 timed steps"}
            except Exception as e:  # the baseline must never take the GPU number down with it
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                        "sample": "failed: %r" % (e,)}
        print(json.dumps(line), flush=True)
    if world > 1:
        # Captured NCCL kernels keep the communicator busy; tearing the process group down after graph
        # capture can block (observed on 2xB200, torch 2.11 / NCCL 2.28).  Everything is flushed and
        # synchronised, so leave without the destructor.
        barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
