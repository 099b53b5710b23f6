"""Generate tests/golden/optim.pt by running the reference's optimizer stack:
torch.optim.Adam built by the reference's `get_optimizer` (lightning/optimizer.py), the reference's
`get_scheduler` (lightning/scheduler.py) and torch.nn.utils.clip_grad_norm_ (what pytorch_lightning's
`gradient_clip_val` calls, main.py:104-110) on seeded parameters / gradients.

Run in the build container:  python oracle/make_golden_optim.py
The oracle restatement (oracle/optim_oracle.py) must reproduce the result before the fixture is written.
"""
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import optim_oracle  # noqa: E402

REF = "/root/reference"


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref_opt = _load("ref_optimizer", os.path.join(REF, "lightning", "optimizer.py"))
    ref_sched = _load("ref_scheduler", os.path.join(REF, "lightning", "scheduler.py"))
    cases = {}
    for name, sched, wd in (("sqrt", "sqrt", 0.0), ("const_wd", "const", 0.01)):
        train_config = {
            "scheduler_type": sched,
            "optimizer": {"betas": [0.9, 0.98], "eps": 1e-9, "weight_decay": wd, "grad_clip_thresh": 1.0,
                          "warm_up_step": 4, "anneal_steps": [6, 9], "anneal_rate": 0.3},
        }
        g = torch.Generator().manual_seed(5)
        shapes = [(7, 5), (13,), (4, 3, 3), (1,)]
        params0 = [torch.randn(s, generator=g) for s in shapes]
        steps = 12
        # gradient scales straddle the clip threshold (norm > 1 in some steps, < 1 in others)
        grads = [[torch.randn(s, generator=g) * (0.05 if k % 3 == 0 else 0.5) for s in shapes] for k in range(steps)]

        class Holder(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.ps = torch.nn.ParameterList([torch.nn.Parameter(p.clone()) for p in params0])

        model = Holder()
        opt = ref_opt.get_optimizer(model, None, train_config)
        sch = ref_sched.get_scheduler(opt, train_config)
        lrs, norms = [], []
        for k in range(steps):
            for p, gk in zip(model.ps, grads[k]):
                p.grad = gk.clone()
            norms.append(float(torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)))
            lrs.append(opt.param_groups[0]["lr"])
            opt.step()
            sch.step()
        expected = [p.detach().clone() for p in model.ps]

        mine = [p.clone() for p in params0]
        cfg = dict(lr=0.001, betas=(0.9, 0.98), eps=1e-9, weight_decay=wd, max_norm=1.0, sched=sched, warmup=4,
                   anneal_steps=[6, 9], anneal_rate=0.3)
        o_lrs, o_norms = optim_oracle.run(mine, grads, cfg)
        for a, b in zip(mine, expected):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), (name, (a - b).abs().max())
        assert all(abs(a - b) <= 1e-12 + 1e-7 * abs(b) for a, b in zip(o_lrs, lrs)), (o_lrs, lrs)
        assert all(abs(a - b) <= 1e-5 * abs(b) for a, b in zip(o_norms, norms))
        cases[name] = {"train_config": train_config, "cfg": cfg, "params0": params0, "grads": grads,
                       "expected": expected, "lrs": lrs, "norms": norms}
        print(name, "ok; lrs", ["%.2e" % l for l in lrs])
    out = os.path.join(ROOT, "tests", "golden", "optim.pt")
    torch.save(cases, out)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
