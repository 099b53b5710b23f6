"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference, via ref_loader).

Run in the build container:  python oracle/make_golden.py
Each model fixture holds the seeded inputs, the reference's outputs / six losses and a digest of every
parameter gradient (norm + first values; full tensors for a few small parameters).  Before a fixture
is written the oracle restatement (fs2_oracle.py) must reproduce the reference on the same inputs.
These are derived artefacts of running reference code, not copies of reference data.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fs2_oracle, ref_loader, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

MODEL_CASES = {
    # name: (cfg overrides, spk_config, make_batch kwargs)
    "small": (dict(), None, dict(B=3, src_len=(5, 20), dur=synth.uniform_dur(0, 6), seed=11)),
    "spk_lang": (dict(multi_speaker=True, multi_lingual=True, encoder_layer=2, decoder_layer=2),
                 {"emb_type": "table", "speakers": list(range(5))},
                 dict(B=2, src_len=(4, 12), dur=synth.uniform_dur(1, 5), seed=12, n_speaker=5, n_lang=3)),
    "truncate": (dict(max_seq_len=40, encoder_layer=1, decoder_layer=2), None,
                 dict(B=2, src_len=(10, 14), dur=synth.uniform_dur(2, 6), seed=13)),
    "f64_energy": (dict(encoder_layer=1, decoder_layer=1), None,
                   dict(B=2, src_len=(6, 9), dur=synth.uniform_dur(1, 4), seed=14, energy_f64=True)),
}
FULL_GRAD_KEYS = ["mel_linear.bias", "variance_adaptor.duration_predictor.linear_layer.weight",
                  "decoder.layer_stack.0.slf_attn.layer_norm.weight", "postnet.convolutions.4.1.bias",
                  "encoder.layer_stack.0.pos_ffn.w_2.bias"]


def run_model_case(name, over, spk_config, bkw):
    cfg = synth.model_cfg(**over)
    model, loss_fn = ref_loader.build_reference_model(cfg, spk_config)
    sd = synth.init_state_dict(model.state_dict(), seed=0)
    model.load_state_dict(sd)
    batch = synth.make_batch(**bkw)
    with ref_loader.no_functional_dropout():
        out = model(batch[2], batch[3], *batch[4:12], lang_args=batch[12])
        losses = loss_fn(batch[:-1], out)
        losses[0].backward()
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    # the oracle restatement must agree with the reference before the fixture is trusted
    o_out, o_losses, o_grads = fs2_oracle.step({k: v.clone() for k, v in sd.items()}, cfg, batch)
    for a, b in zip(out[:5], o_out[:5]):
        assert torch.allclose(a, b, atol=1e-5, rtol=1e-5), name
    assert torch.equal(out[9], o_out[9])
    for a, b in zip(losses, o_losses):
        assert abs(float(a) - float(b)) < 1e-5, (name, float(a), float(b))
    for k, g in grads.items():
        assert torch.allclose(g, o_grads[k], atol=1e-5, rtol=1e-4), (name, k)
    fx = {
        "cfg": cfg, "spk_config": spk_config, "weight_seed": 0,
        "batch": tuple(b for b in batch),
        "out": {"mel": out[0].detach(), "post": out[1].detach(), "pitch": out[2].detach(),
                "energy": out[3].detach(), "log_d": out[4].detach(), "mel_len": out[9].detach(),
                "mel_mask_len": out[7].shape[1]},
        "losses": torch.stack([l.detach() for l in losses]),
        "grad_digest": {k: (float(g.norm()), g.flatten()[:8].clone()) for k, g in grads.items()},
        "grad_full": {k: grads[k].clone() for k in FULL_GRAD_KEYS if k in grads},
    }
    torch.save(fx, os.path.join(OUT, "model_%s.pt" % name))
    print("wrote model_%s.pt: losses %s" % (name, [round(float(l), 5) for l in losses]))


def run_lr_cases():
    _, _, LengthRegulator = ref_loader.load()
    lr = LengthRegulator()
    g = torch.Generator().manual_seed(21)
    cases = []
    specs = [
        ("uniform", torch.randint(0, 6, (3, 9), generator=g), None),
        ("pad_to_longer", torch.randint(1, 4, (2, 7), generator=g), 40),
        ("crop", torch.randint(2, 9, (2, 11), generator=g), 17),
        ("skewed", torch.stack([synth.skewed_dur(g, 50) for _ in range(2)]), None),
        ("all_zero_row", torch.tensor([[0, 0, 0, 0], [1, 0, 2, 0]]), 5),
        ("negative", torch.tensor([[2, -3, 1, 4], [-1, -1, 5, 0]]), None),
        ("float_durations", torch.tensor([[1.9, 0.2, 3.0, -0.5], [2.5, 2.5, 0.99, 1.0]]), None),
    ]
    for name, dur, max_len in specs:
        x = torch.randn(dur.shape[0], dur.shape[1], 8, generator=g)
        out, mel_len = lr(x, dur, max_len)
        o2, l2 = fs2_oracle.length_regulate(x, dur, max_len)
        assert torch.equal(out, o2) and torch.equal(mel_len, l2), name
        cases.append({"name": name, "x": x, "dur": dur, "max_len": max_len, "out": out, "mel_len": mel_len})
    torch.save(cases, os.path.join(OUT, "length_regulator.pt"))
    print("wrote length_regulator.pt (%d cases)" % len(cases))


def run_misc():
    """bucketize semantics (appendix C.5) and the sinusoid table (Models.py:10-30)."""
    sys.path.insert(0, ref_loader.REF)
    ref_loader.load()
    mods = sys.modules["_fs2ref_transformer.Models"]
    table = mods.get_sinusoid_encoding_table(37, 256)
    bins = torch.linspace(-1, 1, 5)
    x = torch.tensor([-2, -1, -0.75, -0.5, 0, 1, 1.5])
    torch.save({"sinusoid_37x256": table, "bins": bins, "bucket_x": x, "bucket_idx": torch.bucketize(x, bins)},
               os.path.join(OUT, "misc.pt"))
    print("wrote misc.pt")


if __name__ == "__main__":
    assert ref_loader.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    run_lr_cases()
    run_misc()
    for name, (over, spk, bkw) in MODEL_CASES.items():
        run_model_case(name, over, spk, bkw)
