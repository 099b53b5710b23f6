"""Data-parallel gradient average on real hardware: 2 ranks, NCCL, one process per GPU.

Every rank steps on ITS OWN batch through the full-depth model (4 + 6 FFT layers); the flat gradient buffer after the
step -- weight gradients produced on the side stream, NCCL buckets launched from inside the backward pass, eager and
as a replayed CUDA graph -- must equal the mean of the per-rank local gradients, which are computed by the same code
with the reduction switched off and exchanged with a plain all_gather (reference semantics: Lightning `ddp` averages
gradients over ranks, main.py:34-40; the six logged losses are rank means, FastSpeech2.py:89).  Tolerance: 1e-3
norm-wise (the activation / input-gradient chain is bit-reproducible; weight gradients differ by the order of fp32
atomics, ~1e-6).  Also covers: several modules announcing gradients more than once is refused loudly, and the process
group shuts down cleanly after TrainStep.close() (no os._exit).
"""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from fs2b200 import sub
    from tests.util_parity import disable_dropout

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    synth, M, rt = sub("synthetic"), sub("lightning.model"), sub("runtime")
    cfg = synth.model_cfg(multi_speaker=True)  # full depth: 4 encoder + 6 decoder layers
    spk = {"emb_type": "table", "speakers": list(range(9))}

    def build():
        m = M.FastSpeech2(cfg, spk_config=spk)
        m.load_state_dict(synth.init_state_dict(m.state_dict(), 0))
        return disable_dropout(m.to(dev).train()), M.FastSpeech2Loss(cfg)

    # different utterances (and padded shapes) per rank
    batch = synth.make_batch(B=6, src_len=(20, 90), dur=synth.uniform_dur(1, 8), seed=50 + rank, n_speaker=9)
    m1, l1 = build()
    b1 = rt.GradBuckets(m1.parameters(), device=dev)
    b1.world = 1  # local gradient: same code path, no reduction
    s1 = rt.TrainStep(m1, l1, batch, use_graph=False, buckets=b1, device=dev)
    s1.run()
    local_flat = b1.flat.clone()
    local_losses = s1.read_losses().clone().to(dev)
    gathered = [torch.empty_like(local_flat) for _ in range(world)]
    dist.all_gather(gathered, local_flat)
    mean = torch.stack(gathered).mean(0)
    gl = [torch.empty_like(local_losses) for _ in range(world)]
    dist.all_gather(gl, local_losses)
    mean_losses = torch.stack(gl).mean(0)
    res = {}
    for use_graph in (False, True):
        m2, l2 = build()
        b2 = rt.GradBuckets(m2.parameters(), bucket_bytes=4 << 20, device=dev)  # many buckets -> overlap path
        assert len(b2.buckets) > 4 and b2._avg
        s2 = rt.TrainStep(m2, l2, batch, use_graph=use_graph, buckets=b2, device=dev)
        for _ in range(2):  # a replay must reproduce itself
            s2.run()
        torch.cuda.synchronize()
        err = ((b2.flat - mean).norm() / mean.norm()).item()
        per_bucket = max(((b2.flat[s:e] - mean[s:e]).norm() / (mean[s:e].norm() + 1e-30)).item()
                         for s, e in b2.buckets)
        lerr = ((s2.read_losses().to(dev) - mean_losses).abs() / mean_losses.abs()).max().item()
        res["graph" if use_graph else "eager"] = (err, per_bucket, lerr)
        s2.close()
    q.put((rank, res))
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()  # clean shutdown: the graphs (and their captured NCCL kernels) are gone


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_gradient_average_eager_and_graph():
    import torch.multiprocessing as mp

    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0, "rank did not shut down cleanly (exit code %r)" % (p.exitcode,)
    for rank, r in res:
        print("rank", rank, r)
        for mode, (err, per_bucket, lerr) in r.items():
            assert err <= TOL, (rank, mode, err)
            assert per_bucket <= 5 * TOL, (rank, mode, per_bucket)
            assert lerr <= 1e-5, (rank, mode, lerr)
