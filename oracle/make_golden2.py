"""Round-2 fixtures, again produced by running the UNMODIFIED reference (/root/reference via ref_loader):

  model_eval.pt         eval-mode inference (no targets: predicted durations / pitch / energy, running-statistics
                        BatchNorm, F.dropout off), one case longer than max_seq_len (sinusoid table rebuilt)
  model_frame_level.pt  training step with frame-level pitch / energy (modules.py:141-150, loss.py:50-59)
  ada_loss.pt           FastSpeech2ADALoss (loss.py:104-140) on random predictions / masks

Run in the build container:  python oracle/make_golden2.py
Before a fixture is written the oracle restatement must reproduce the reference on the same inputs.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fs2_oracle, ref_loader, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def nontrivial_running_stats(sd, seed=5):
    g = torch.Generator().manual_seed(seed)
    for k in sd:
        if k.endswith("running_mean"):
            sd[k] = 0.3 * torch.randn(sd[k].shape, generator=g)
        elif k.endswith("running_var"):
            sd[k] = 0.5 + torch.rand(sd[k].shape, generator=g)
    return sd


def run_eval_cases():
    cases = []
    for name, over, bkw in [
        ("eval_short", dict(encoder_layer=2, decoder_layer=2),
         dict(B=3, src_len=(5, 14), dur=synth.uniform_dur(1, 4), seed=61)),
        ("eval_long", dict(encoder_layer=1, decoder_layer=2, max_seq_len=24),
         dict(B=2, src_len=(8, 12), dur=synth.uniform_dur(1, 4), seed=62)),
    ]:
        cfg = synth.model_cfg(**over)
        model, _ = ref_loader.build_reference_model(cfg, None, train=False)
        sd = nontrivial_running_stats(synth.init_state_dict(model.state_dict(), seed=0))
        # a duration predictor that yields a few frames per phoneme at random init
        sd["variance_adaptor.duration_predictor.linear_layer.bias"] = torch.full((1,), 1.2)
        model.load_state_dict(sd)
        model.eval()
        batch = synth.make_batch(**bkw)
        with torch.no_grad():
            out = model(batch[2], batch[3], batch[4], batch[5], p_control=1.1, e_control=0.9, d_control=1.0)
            o = fs2_oracle.forward({k: v.clone() for k, v in sd.items()}, cfg, batch[2], batch[3], batch[4], batch[5],
                                   training=False, p_control=1.1, e_control=0.9, d_control=1.0)
        assert torch.equal(out[5], o[5]) and torch.equal(out[9], o[9]), name
        for a, b in zip(out[:5], o[:5]):
            assert torch.allclose(a, b, atol=2e-5, rtol=1e-5), name
        assert torch.equal(out[7], o[7])
        if name == "eval_long":
            assert out[0].shape[1] > cfg["max_seq_len"]
        cases.append({"name": name, "cfg": cfg, "sd_overrides": {k: sd[k] for k in sd if "running_" in k
                                                                 or k.endswith("duration_predictor.linear_layer.bias")},
                      "batch": batch, "controls": (1.1, 0.9, 1.0),
                      "out": {"mel": out[0], "post": out[1], "pitch": out[2], "energy": out[3], "log_d": out[4],
                              "d_rounded": out[5], "mel_len": out[9]}})
        print("eval case %s: mel %s" % (name, tuple(out[0].shape)))
    torch.save(cases, os.path.join(OUT, "model_eval.pt"))


def run_frame_level():
    cfg = synth.model_cfg(encoder_layer=1, decoder_layer=2)
    cfg["pitch"]["feature"] = "frame_level"
    cfg["energy"]["feature"] = "frame_level"
    model, loss_fn = ref_loader.build_reference_model(cfg, None)
    sd = synth.init_state_dict(model.state_dict(), seed=0)
    model.load_state_dict(sd)
    b = list(synth.make_batch(B=3, src_len=(5, 12), dur=synth.uniform_dur(1, 5), seed=63))
    g = torch.Generator().manual_seed(64)
    Tm = int(b[8])
    valid_m = torch.arange(Tm)[None, :] < b[7][:, None]
    b[9] = torch.randn(len(b[0]), Tm, generator=g) * valid_m   # frame-level pitch
    b[10] = torch.randn(len(b[0]), Tm, generator=g) * valid_m  # frame-level energy
    batch = tuple(b)
    with ref_loader.no_functional_dropout():
        out = model(batch[2], batch[3], *batch[4:12], lang_args=batch[12])
        losses = loss_fn(batch[:-1], out)
        losses[0].backward()
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    o_out, o_losses, o_grads = fs2_oracle.step({k: v.clone() for k, v in sd.items()}, cfg, batch)
    for a, c in zip(out[:5], o_out[:5]):
        assert torch.allclose(a, c, atol=1e-5, rtol=1e-5)
    for a, c in zip(losses, o_losses):
        assert abs(float(a) - float(c)) < 1e-5
    for k, gr in grads.items():
        assert torch.allclose(gr, o_grads[k], atol=1e-5, rtol=1e-4), k
    assert out[2].shape == (3, Tm)
    fx = {"cfg": cfg, "weight_seed": 0, "batch": batch,
          "out": {"mel": out[0].detach(), "post": out[1].detach(), "pitch": out[2].detach(),
                  "energy": out[3].detach(), "log_d": out[4].detach()},
          "losses": torch.stack([l.detach() for l in losses]),
          "grad_digest": {k: (float(gr.norm()), gr.flatten()[:8].clone()) for k, gr in grads.items()}}
    torch.save(fx, os.path.join(OUT, "model_frame_level.pt"))
    print("frame-level losses", [round(float(l), 5) for l in losses])


def run_ada_loss():
    ref_loader.load()
    ADALoss = sys.modules["_fs2ref_lightning.model.loss"].FastSpeech2ADALoss
    g = torch.Generator().manual_seed(65)
    cases = []
    for B, Tm, Tt in [(3, 37, 37), (2, 50, 64)]:  # second case: target longer than the (truncated) prediction
        lens = torch.randint(5, Tm + 1, (B,), generator=g)
        lens[0] = Tm
        masks = torch.arange(Tm)[None, :] >= lens[:, None]
        mel, post = torch.randn(B, Tm, 80, generator=g), torch.randn(B, Tm, 80, generator=g)
        tgt = torch.randn(B, Tt, 80, generator=g)
        mel.requires_grad_(True)
        post.requires_grad_(True)
        out = ADALoss()(tgt, (mel, post, masks))
        out[0].backward()
        o = fs2_oracle.ada_loss(tgt, (mel.detach(), post.detach(), masks))
        assert all(abs(float(a) - float(c)) < 1e-6 for a, c in zip(out, o))
        cases.append({"mel": mel.detach(), "post": post.detach(), "masks": masks, "target": tgt,
                      "losses": torch.stack([l.detach() for l in out]), "d_mel": mel.grad.clone(),
                      "d_post": post.grad.clone()})
    torch.save(cases, os.path.join(OUT, "ada_loss.pt"))
    print("ada losses", [c["losses"].tolist() for c in cases])


def _exec_reference_function(path, name, ns):
    """exec one top-level function of a reference source file (its imports are not executable here)"""
    src = open(os.path.join(ref_loader.REF, path)).read().split("\n")
    start = next(i for i, l in enumerate(src) if l.startswith("def %s(" % name))
    end = next((i for i in range(start + 1, len(src)) if src[i].startswith("def ")), len(src))
    exec("\n".join(src[start:end]), ns)
    return ns[name]


def run_collate():
    """The reference's own `reprocess` (lightning/collates/utils.py:8-111) with its own pad_1D / pad_2D
    (lightning/utils/tool.py:134-165), exec'd from source, on seeded ragged items, all three modes."""
    import numpy as np

    ns = {"np": np, "torch": torch}
    _exec_reference_function("lightning/utils/tool.py", "pad_1D", ns)
    _exec_reference_function("lightning/utils/tool.py", "pad_2D", ns)
    reprocess = _exec_reference_function("lightning/collates/utils.py", "reprocess", ns)
    cases = []
    for seed in (0, 1):
        rng = np.random.default_rng(seed)
        items = []
        for i in range(9):
            L = int(rng.integers(3, 40))
            dur = rng.integers(0, 9, L).astype(np.int64)
            T = max(int(dur.sum()), 1)
            items.append({"id": "utt%d" % i, "raw_text": "text %d" % i, "speaker": int(rng.integers(0, 7)),
                          "lang_id": int(rng.integers(0, 3)), "text": rng.integers(1, 80, L).astype(np.int64),
                          "mel": rng.standard_normal((T, 80)).astype(np.float32),
                          "pitch": rng.standard_normal(L).astype(np.float32),
                          "energy": rng.standard_normal(L).astype(np.float64 if seed % 2 else np.float32),
                          "duration": dur})
        idxs = [4, 0, 7, 2, 8]
        cases.append({"items": items, "idxs": idxs,
                      "out": {m: reprocess(items, idxs, mode=m) for m in ("sup", "unsup", "inference")}})
    torch.save(cases, os.path.join(OUT, "collate.pt"))
    print("collate fixture:", [len(c["out"]["sup"]) for c in cases], [len(c["out"]["inference"]) for c in cases])


if __name__ == "__main__":
    assert ref_loader.available(), "needs /root/reference"
    torch.manual_seed(0)
    run_eval_cases()
    run_frame_level()
    run_ada_loss()
    run_collate()
