"""N-rank NCCL check of the data-parallel step (run under torchrun): every rank steps on ITS OWN batch; the averaged
flat gradient must equal the mean of the per-rank local gradients (gathered with a separate all_gather), with the
weight gradients produced on the side stream and the NCCL buckets launched while the backward is still running."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs2b200 import sub  # noqa: E402
synth = sub("synthetic")  # the package's own generators (oracle/ is for tests only)
from tests.util_parity import disable_dropout  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    M, rt = sub("lightning.model"), sub("runtime")
    cfg = synth.model_cfg(encoder_layer=2, decoder_layer=2)

    def build():
        m = M.FastSpeech2(cfg)
        m.load_state_dict(synth.init_state_dict(m.state_dict(), 0))
        return disable_dropout(m.to(dev).train()), M.FastSpeech2Loss(cfg)

    batch = synth.make_batch(B=4, src_len=(10, 40), dur=synth.uniform_dur(1, 8), seed=50 + (0 if os.environ.get("FS2_DP_SAME") else rank))
    # local gradient (no reduction): same code path with world forced to 1
    m1, l1 = build()
    b1 = rt.GradBuckets(m1.parameters(), device=dev)
    b1.world = 1
    s1 = rt.TrainStep(m1, l1, batch, use_graph=False, buckets=b1, device=dev)
    s1.run()
    local_flat = b1.flat.clone()
    gathered = [torch.empty_like(local_flat) for _ in range(world)]
    # ranks have different padded shapes but the same parameters -> same flat layout
    dist.all_gather(gathered, local_flat)
    mean = torch.stack(gathered).mean(0)
    for use_graph in (False, True):
        m2, l2 = build()
        b2 = rt.GradBuckets(m2.parameters(), bucket_bytes=4 << 20, device=dev)  # several buckets -> overlap path
        if os.environ.get("FS2_DP_NO_OVERLAP"):
            b2.overlap = False
        s2 = rt.TrainStep(m2, l2, batch, use_graph=use_graph, buckets=b2, device=dev)
        s2.run()
        torch.cuda.synchronize()
        err = ((b2.flat - mean).norm() / mean.norm()).item()
        if err > 5e-2 and rank == 0:
            tot = torch.stack(gathered).sum(0)
            for bi, (s0, e0) in enumerate(b2.buckets):
                f = b2.flat[s0:e0]
                rel = lambda ref: ((f - ref[s0:e0]).norm() / (ref[s0:e0].norm() + 1e-30)).item()
                print("  bucket %d [%d:%d] vs mean %.2e  vs sum %.2e  vs local %.2e  vs local/2 %.2e |f|=%.3e |mean|=%.3e"
                      % (bi, s0, e0, rel(mean), rel(tot), rel(local_flat), rel(local_flat / 2), f.norm().item(),
                         mean[s0:e0].norm().item()),
                      flush=True)
        print("rank %d graph=%s buckets=%d rel-err of averaged gradient vs mean of local gradients: %.3e"
              % (rank, use_graph, len(b2.buckets), err), flush=True)
        if not os.environ.get("FS2_DP_DIAG"):
            assert err < 5e-2, err  # run-to-run noise floor of the bf16 path (tools/debug_determinism.py)
    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        print("dp check ok", flush=True)
    os._exit(0)


if __name__ == "__main__":
    main()
