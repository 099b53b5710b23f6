"""Data-parallel gradient buckets on CPU: world_size 2, gloo (host logic of runtime/dp.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fs2b200 import sub


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4),
                               torch.nn.LayerNorm(4))


def _data(rank):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.randn(5 + rank, 8, generator=g), torch.randn(5 + rank, 4, generator=g)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dp = sub("runtime.dp")
    m = _model()
    buckets = dp.GradBuckets(m.parameters(), bucket_bytes=256, device=torch.device("cpu"))
    assert len(buckets.buckets) > 1
    for step in range(2):  # second step checks that zero() really resets
        buckets.zero()
        x, y = _data(rank)
        loss = ((m(x) - y) ** 2).mean()  # per-rank normalisation (mean of means, like DDP)
        loss.backward()
        # every second parameter announces itself (as the CUDA backward functions do), the rest is
        # picked up by finish()
        buckets.notify(list(m.parameters())[::2])
        buckets.finish()
        scal = buckets.reduce_scalars(torch.tensor([loss.item()]))
    # plain lists: tensors in an mp.Queue are shared-memory handles that die with the worker
    q.put((rank, buckets.flat.tolist(), [p.grad.flatten().tolist() for p in m.parameters()], float(scal)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_average_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process emulation: average of the per-rank gradients
    ref, losses = None, []
    for r in range(world):
        m = _model()
        x, y = _data(r)
        loss = ((m(x) - y) ** 2).mean()
        loss.backward()
        losses.append(loss.item())
        g = [p.grad for p in m.parameters()]
        ref = g if ref is None else [a + b for a, b in zip(ref, g)]
    ref = [g / world for g in ref]
    for rank, flat, grads, scal in res:
        for a, b in zip(grads, ref):
            assert torch.allclose(torch.tensor(a), b.flatten(), atol=1e-6)
        assert abs(scal - sum(losses) / world) < 1e-6
    assert res[0][1] == res[1][1]  # identical flat buffers on both ranks


def test_bucket_views_alias_flat_buffer_in_reverse_order():
    dp = sub("runtime.dp")
    m = _model()
    b = dp.GradBuckets(m.parameters(), bucket_bytes=128, device=torch.device("cpu"))
    params = list(m.parameters())
    assert b.flat.numel() == sum((p.numel() + 3) // 4 * 4 for p in params)
    assert all(p.main_grad.data_ptr() % 16 == 0 for p in params)
    # reverse registration order: the last parameter sits first in the flat buffer
    assert params[-1].main_grad.data_ptr() == b.flat.data_ptr()
    for p in params:
        assert p.grad.data_ptr() == p.main_grad.data_ptr()
        p.main_grad.fill_(1.0)
    assert float(b.flat.sum()) == sum(p.numel() for p in params)
    covered = sorted(b.buckets)
    assert covered[0][0] == 0 and covered[-1][1] == b.flat.numel()
    assert all(covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1))


def test_bucket_refuses_gradients_announced_after_the_reduce():
    """A parameter announced again after its bucket went on the wire (module called twice per step) must not be
    averaged silently with a partial gradient (notify() counted notifications, not parameters)."""
    dp = sub("runtime.dp")
    lin = torch.nn.Linear(8, 8)
    b = dp.GradBuckets(lin.parameters(), device=torch.device("cpu"))
    b.world = 2  # pretend: exercise the bookkeeping without a process group
    launched = []
    b._launch = lambda i: (launched.append(i), b._launched.__setitem__(i, True))
    b.zero()
    b.notify([lin.weight])
    b.notify([lin.weight])  # same parameter twice: still waiting for the bias
    assert launched == []
    b.notify([lin.bias])
    assert launched == [0]
    b.notify([lin.weight])  # late contribution
    b.world = 1
    with pytest.raises(RuntimeError):
        b.finish()
