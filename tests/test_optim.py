"""Fused clip + Adam + LR schedule.

CPU: the oracle restatement against the fixture produced by the reference's own optimizer stack
(oracle/make_golden_optim.py: reference get_optimizer / get_scheduler + clip_grad_norm_).
GPU: csrc/optim.cu through runtime.FusedAdam against the same fixture and against the oracle on a
FastSpeech2-sized flat buffer.  Tolerance: fp32, rel 2e-5 / abs 1e-7 after 12 steps (the kernel evaluates
m = b1*m + (1-b1)*g where torch uses lerp, and the bias corrections in double precision).
"""
import os

import pytest
import torch

from fs2b200 import sub
from oracle import optim_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "optim.pt")


@pytest.mark.parametrize("case", ["sqrt", "const_wd"])
def test_oracle_matches_reference_optimizer_fixture(case):
    fx = torch.load(GOLD, weights_only=False)[case]
    params = [p.clone() for p in fx["params0"]]
    lrs, norms = optim_oracle.run(params, fx["grads"], fx["cfg"])
    for a, b in zip(params, fx["expected"]):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)
    assert lrs == pytest.approx(fx["lrs"], rel=1e-7)
    assert norms == pytest.approx(fx["norms"], rel=1e-5)


def test_schedule_known_values():
    """lightning/scheduler.py:26-38 with the shipped config (warm-up 4000, anneal 0.3 at 30k/40k/50k)."""
    f = lambda k: optim_oracle.lr_factor(k, "sqrt", 4000, [30000, 40000, 50000], 0.3)
    assert f(0) == pytest.approx(1 / 4000)
    assert f(3999) == pytest.approx(1.0)
    assert f(15999) == pytest.approx(0.5)
    assert f(30000) == pytest.approx((4000 / 30001) ** 0.5 * 0.3)
    assert optim_oracle.lr_factor(9999, "const", 4000, [], 1.0) == 1.0


class _Holder(torch.nn.Module):
    def __init__(self, tensors):
        super().__init__()
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(t.clone()) for t in tensors])


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["sqrt", "const_wd"])
def test_fused_adam_matches_reference_fixture(case):
    rt = sub("runtime")
    fx = torch.load(GOLD, weights_only=False)[case]
    model = _Holder(fx["params0"]).cuda()
    buckets = rt.GradBuckets(model.parameters(), device=torch.device("cuda"))
    opt = rt.FusedAdam(buckets, train_config=fx["train_config"])
    for k, grads in enumerate(fx["grads"]):
        buckets.zero()
        for p, g in zip(model.ps, grads):
            p.main_grad.copy_(g.cuda())
        opt.step()
        assert float(opt.grad_norm) == pytest.approx(fx["norms"][k], rel=1e-5)
    assert int(opt.step_dev) == len(fx["grads"])
    for p, e in zip(model.ps, fx["expected"]):
        assert torch.allclose(p.detach().cpu(), e, rtol=2e-5, atol=1e-7), (p.detach().cpu() - e).abs().max()


@pytest.mark.gpu
def test_fused_adam_large_flat_buffer_vs_oracle_and_graph_capture():
    """34.5 M-element flat buffer (the FastSpeech2 size), 3 steps replayed from ONE captured CUDA graph: the
    device-side step counter drives lr / bias corrections, so replays must equal 3 eager oracle steps."""
    rt = sub("runtime")
    torch.manual_seed(0)
    n = 34_553_923
    model = _Holder([torch.randn(n // 2), torch.randn(n - n // 2)]).cuda()
    p0 = [p.detach().clone() for p in model.ps]
    buckets = rt.GradBuckets(model.parameters(), device=torch.device("cuda"))
    cfg = {"scheduler_type": "sqrt", "optimizer": {"betas": [0.9, 0.98], "eps": 1e-9, "weight_decay": 0.0,
                                                  "grad_clip_thresh": 1.0, "warm_up_step": 2, "anneal_steps": [],
                                                  "anneal_rate": 1.0}}
    opt = rt.FusedAdam(buckets, train_config=cfg)
    gsrc = [torch.randn_like(p) * 1e-4 for p in model.ps]  # total norm ~0.59 -> unclipped ... scaled per step below
    scale = torch.ones(1, device="cuda")

    def body():
        for p, g in zip(model.ps, gsrc):
            p.main_grad.copy_(g * scale)
        opt.step()

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        graph = torch.cuda.CUDAGraph()
        # nothing ran yet: state is pristine when the capture begins
        with torch.cuda.graph(graph, stream=s):
            body()
    grads = []
    for k, sc in enumerate((1.0, 3.0, 0.5)):  # norms ~0.59, ~1.76 (clipped), ~0.29
        scale.fill_(sc)
        graph.replay()
        grads.append([(g * sc) for g in gsrc])
    torch.cuda.synchronize()
    ref = [p.clone() for p in p0]
    ocfg = dict(lr=0.001, betas=(0.9, 0.98), eps=1e-9, weight_decay=0.0, max_norm=1.0, sched="sqrt", warmup=2,
                anneal_steps=[], anneal_rate=1.0)
    optim_oracle.run(ref, grads, ocfg)
    assert int(opt.step_dev) == 3
    for p, r in zip(model.ps, ref):
        assert torch.allclose(p.detach(), r, rtol=2e-5, atol=1e-7), (p.detach() - r).abs().max()
