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


@pytest.mark.gpu
def test_fused_adam_state_dict_round_trips_through_torch_adam():
    """FusedAdam.state_dict() is torch.optim.Adam's format (what the reference's Lightning checkpoints hold in
    `optimizer_states`, lightning/optimizer.py:5-16): a torch Adam loads it and takes the same next step, and a
    FusedAdam loads the torch optimizer's state back -- although the fused buffers keep Conv1d states in [Co][k][Ci]
    order and the buckets in reverse registration order."""
    rt = sub("runtime")
    torch.manual_seed(3)

    def make():
        torch.manual_seed(5)
        m = torch.nn.Sequential(torch.nn.Conv1d(8, 12, 5), torch.nn.Linear(7, 6), torch.nn.LayerNorm(6)).cuda()
        m[1].bias.requires_grad_(False)  # a frozen parameter keeps its index in torch's param list
        return m

    hp = dict(lr=1e-3, betas=(0.9, 0.98), eps=1e-9, weight_decay=0.0)
    m1, m2 = make(), make()
    buckets = rt.GradBuckets(m1.parameters(), device=torch.device("cuda"))
    fused = rt.FusedAdam(buckets, lr=hp["lr"], betas=hp["betas"], eps=hp["eps"], weight_decay=0.0, max_grad_norm=0.0,
                         scheduler_type="none")
    ref = torch.optim.Adam(m2.parameters(), **hp)
    gs = [[torch.randn_like(p) for p in m1.parameters()] for _ in range(4)]

    def step_fused(g):
        buckets.zero()
        for p, gi in zip(m1.parameters(), g):
            if p.requires_grad:
                p.main_grad.copy_(gi)
        fused.step()

    def step_ref(g):
        for p, gi in zip(m2.parameters(), g):
            p.grad = gi.clone() if p.requires_grad else None
        ref.step()

    for g in gs[:3]:
        step_fused(g)
        step_ref(g)
    sd = fused.state_dict()
    assert sorted(sd["state"].keys()) == [i for i, p in enumerate(m1.parameters()) if p.requires_grad]
    assert sd["state"][0]["exp_avg"].shape == (12, 8, 5)
    for i, ent in ref.state_dict()["state"].items():
        assert torch.allclose(sd["state"][i]["exp_avg"], ent["exp_avg"], rtol=1e-5, atol=1e-8)
        assert torch.allclose(sd["state"][i]["exp_avg_sq"], ent["exp_avg_sq"], rtol=1e-5, atol=1e-10)
        assert float(sd["state"][i]["step"]) == float(ent["step"]) == 3.0
    # torch Adam <- fused state, fused <- torch state: one more identical step on both
    m3 = make()
    with torch.no_grad():
        for a, b in zip(m3.parameters(), m1.parameters()):
            a.copy_(b)
    ref3 = torch.optim.Adam(m3.parameters(), **hp)
    ref3.load_state_dict(sd)
    m4 = make()
    b4 = rt.GradBuckets(m4.parameters(), device=torch.device("cuda"))
    fused4 = rt.FusedAdam(b4, lr=hp["lr"], betas=hp["betas"], eps=hp["eps"], weight_decay=0.0, max_grad_norm=0.0,
                          scheduler_type="none")
    with torch.no_grad():
        for a, b in zip(m4.parameters(), m2.parameters()):
            a.copy_(b)
    fused4.load_state_dict(ref.state_dict())
    assert int(fused4.step_dev) == 3
    for p, gi in zip(m3.parameters(), gs[3]):
        p.grad = gi.clone() if p.requires_grad else None
    ref3.step()
    step_ref(gs[3])
    b4.zero()
    for p, gi in zip(m4.parameters(), gs[3]):
        if p.requires_grad:
            p.main_grad.copy_(gi)
    fused4.step()
    for a, b, c in zip(m2.parameters(), m3.parameters(), m4.parameters()):
        assert torch.allclose(a, b, rtol=2e-5, atol=1e-7) and torch.allclose(a, c, rtol=2e-5, atol=1e-7)
    with pytest.raises(ValueError):
        fused4.load_state_dict({"step": 1})
