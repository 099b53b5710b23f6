"""End-to-end parity of the CUDA FastSpeech2 step against (a) the golden fixtures produced by the
unmodified reference and (b) the oracle restatement on larger seeded batches.

Stated tolerances for the bf16-operand / fp32-accumulate path (activations are stored in bf16 between
kernels, the reference is fp32 end to end):
  mel / postnet-mel outputs : norm-wise rel-err <= 3e-2
  predictor outputs         : norm-wise rel-err <= 3e-2
  six losses                : rel-err <= 1e-2
  parameter gradients       : cosine >= 0.99 and norm ratio within 5 % for every tensor whose reference
                              gradient norm is non-negligible (>= 1e-4 of the largest)
  LengthRegulator lengths   : torch.equal
"""
import pytest
import torch

from fs2b200 import sub
from oracle import fs2_oracle, synth
from tests.util_parity import cosine, cuda_batch, disable_dropout, load_golden, rel_err

pytestmark = pytest.mark.gpu


def build(cfg, spk_config=None):
    M = sub("lightning.model")
    model = M.FastSpeech2(cfg, spk_config=spk_config) if spk_config else M.FastSpeech2(cfg)
    model.load_state_dict(synth.init_state_dict(model.state_dict(), 0))
    model = disable_dropout(model.cuda().train())
    return model, M.FastSpeech2Loss(cfg)


def run_step(model, loss_fn, batch):
    b = cuda_batch(batch)
    out = model(b[2], b[3], *b[4:12], lang_args=b[12])
    losses = loss_fn(b[:-1], out)
    model.zero_grad(set_to_none=True)
    losses[0].backward()
    torch.cuda.synchronize()
    return out, losses


def check_grads(model, ref_digest=None, ref_full=None, ref_grads=None, cos_min=0.99):
    worst = (1.0, None)
    gmax = max((n for n, _ in ref_digest.values()), default=0.0) if ref_digest else max(
        float(g.norm()) for g in ref_grads.values() if g is not None)
    for k, p in model.named_parameters():
        if not p.requires_grad:
            continue
        if ref_grads is not None:
            rg = ref_grads.get(k)
            if rg is None:
                continue
            rn = float(rg.norm())
        else:
            if k not in ref_digest:
                continue
            rn = ref_digest[k][0]
        assert p.grad is not None, "no gradient for %s" % k
        assert torch.isfinite(p.grad).all(), k
        if rn < 1e-4 * gmax:
            continue
        ratio = float(p.grad.norm()) / rn
        assert 0.95 < ratio < 1.05, (k, ratio)
        if ref_grads is not None:
            c = cosine(p.grad, ref_grads[k])
            worst = min(worst, (c, k))
            assert c >= cos_min, (k, c)
    if ref_full:
        for k, g in ref_full.items():
            c = cosine(dict(model.named_parameters())[k].grad, g)
            worst = min(worst, (c, k))
            assert c >= cos_min, (k, c)
    return worst


@pytest.mark.parametrize("case", ["small", "spk_lang", "truncate", "f64_energy"])
def test_against_reference_golden(case):
    fx = load_golden("model_%s.pt" % case)
    model, loss_fn = build(fx["cfg"], fx["spk_config"])
    out, losses = run_step(model, loss_fn, fx["batch"])
    ref = fx["out"]
    assert out[0].shape == ref["mel"].shape and out[7].shape[1] == ref["mel_mask_len"]
    assert torch.equal(out[9].cpu(), ref["mel_len"])
    errs = {n: rel_err(o, ref[n]) for n, o in zip(("mel", "post", "pitch", "energy", "log_d"), out[:5])}
    lerr = [abs(float(l) - float(r)) / abs(float(r)) for l, r in zip(losses, fx["losses"])]
    print(case, errs, "loss rel-err", max(lerr))
    assert all(e <= 3e-2 for e in errs.values()), errs
    assert max(lerr) <= 1e-2, lerr
    worst = check_grads(model, ref_digest=fx["grad_digest"], ref_full=fx["grad_full"])
    print(case, "worst full-grad cosine", worst)


@pytest.mark.parametrize("name,over,bkw", [
    ("mid", dict(), dict(B=4, src_len=(30, 60), dur=synth.uniform_dur(1, 8), seed=31)),
    ("C4_long_skewed", dict(max_seq_len=1500, multi_speaker=True),
     dict(B=2, src_len=(220, 220), fixed_src_len=220, dur=synth.skewed_dur, seed=4, n_speaker=7)),
])
def test_against_oracle(name, over, bkw):
    cfg = synth.model_cfg(**over)
    spk = {"emb_type": "table", "speakers": list(range(7))} if over.get("multi_speaker") else None
    model, loss_fn = build(cfg, spk)
    batch = synth.make_batch(**bkw)
    out, losses = run_step(model, loss_fn, batch)
    # the oracle runs on the host CPU (fp32): it is the checker, not the thing measured
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    o_out, o_losses, o_grads = fs2_oracle.step(sd, cfg, batch)
    assert torch.equal(out[9].cpu(), o_out[9])
    errs = {n: rel_err(o, r) for n, o, r in zip(("mel", "post", "pitch", "energy", "log_d"), out[:5], o_out[:5])}
    lerr = [abs(float(l) - float(r)) / abs(float(r)) for l, r in zip(losses, o_losses)]
    print(name, errs, "loss rel-err", max(lerr))
    assert all(e <= 3e-2 for e in errs.values()), errs
    assert max(lerr) <= 1e-2, lerr
    worst = check_grads(model, ref_grads=o_grads)
    print(name, "worst grad cosine", worst)
