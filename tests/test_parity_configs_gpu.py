"""Parity of the BASELINE.json configurations at their REAL sizes and depth (4 + 6 FFT layers), through the path the
benchmark times: `runtime.TrainStep` graph replay (multi-tensor weight refresh, side-stream weight gradients, deferred
joins, flat gradient buckets), against the fp32 oracle run on the host CPU on the same seeded batch and weights.

  C1  LJSpeech-shaped, B=16                     C2  the bench batch itself: B=64, Ts<=200, Tm<=1000, 247 speakers
  C3  FSCL query batch, B=8, averaged speaker   C4  B=4, 220 phonemes, skewed durations, ~2000 frames -> truncated 1500
  C5  few-shot fine-tune batch, B=4, multilingual model

Stated tolerances (bf16 operands and bf16 activation storage against an fp32 reference):
  six losses rel <= 1e-2; mel / postnet mel / predictor outputs norm-wise rel <= 3e-2;
  every parameter gradient with a non-negligible norm: cosine >= COS_MIN and norm ratio within 5 %.
The worst-cosine tensors of every case are printed and written to gpurun_out/parity_<case>.json.
"""
import json
import os

import pytest
import torch

from fs2b200 import sub
from oracle import fs2_oracle, synth
from tests.util_parity import cosine, cuda_batch, disable_dropout, rel_err

pytestmark = pytest.mark.gpu

from tests.util_parity import COS_MIN, cos_floor, ratio_tol  # noqa: E402


def _report(name, rows, extra):
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        json.dump({"case": name, "worst": rows[:25], **extra}, open(os.path.join(d, "parity_%s.json" % name), "w"),
                  indent=1)
    except OSError:
        pass


@pytest.mark.parametrize("name", ["C1", "C2", "C3", "C4", "C5"])
def test_config_through_trainstep_graph_against_cpu_oracle(name):
    M, rt = sub("lightning.model"), sub("runtime")
    cfg, model, loss_fn, call_kw = synth.build_config(name, M, device="cuda")
    model = disable_dropout(model)
    batch = synth.make_batch(**synth.CONFIGS[name])
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    # eager forward for the output tensors (the step object only exposes losses and gradients)
    b = cuda_batch(batch)
    with torch.no_grad():
        out = model(b[2], b[3], *b[4:12], lang_args=b[12], **call_kw)
    out = [o.detach().cpu() if torch.is_tensor(o) else o for o in out]
    step = rt.TrainStep(model, loss_fn, batch, use_graph=True, model_kwargs=call_kw)
    assert step.graph is not None
    for _ in range(2):
        losses = step.step_e2e(batch).clone()
    grads = {k: p.main_grad.detach().cpu().clone() for k, p in model.named_parameters() if p.requires_grad}

    # the oracle: fp32, host CPU -- the checker, never the thing measured
    params = {k: v for k, v in sd.items() if v.is_floating_point() and "position_enc" not in k
              and not k.endswith("_bins") and "running_" not in k}
    for v in params.values():
        v.requires_grad_(True)
    o_out = fs2_oracle.forward(sd, cfg, batch[2], batch[3], *batch[4:12], lang_args=batch[12], **call_kw)
    o_losses = fs2_oracle.loss(batch[:12], o_out)
    o_losses[0].backward()
    o_grads = {k: v.grad for k, v in params.items()}

    assert torch.equal(out[9], o_out[9])  # LengthRegulator lengths: bit-exact
    errs = {n: rel_err(o, r) for n, o, r in zip(("mel", "post", "pitch", "energy", "log_d"), out[:5], o_out[:5])}
    lerr = [abs(float(a) - float(r)) / abs(float(r)) for a, r in zip(losses.tolist(), o_losses)]
    gmax = max(float(g.norm()) for g in o_grads.values() if g is not None)
    rows = []
    for k, g in grads.items():
        r = o_grads.get(k)
        if r is None or float(r.norm()) < 1e-4 * gmax:
            continue
        rows.append((cosine(g, r), k, float(g.norm()) / float(r.norm()), float(r.norm())))
    rows.sort()
    print(name, "outputs rel-err", errs, "loss rel-err", max(lerr))
    for c, k, ratio, rn in rows[:8]:
        print("  cos %.5f  ratio %.4f  |g_ref| %.3e  %s" % (c, ratio, rn, k))
    _report(name, rows, {"out_rel_err": errs, "loss_rel_err": lerr, "n_tensors": len(rows),
                         "ratio_outliers": [r for r in rows if not (0.97 < r[2] < 1.03)]})
    assert all(e <= 3e-2 for e in errs.values()), errs
    assert max(lerr) <= 1e-2, lerr
    bad = [(k, round(c, 5), round(ratio, 4), rn) for c, k, ratio, rn in rows
           if c < cos_floor(k) or abs(ratio - 1.0) >= ratio_tol(k)]
    assert not bad, (name, bad)
