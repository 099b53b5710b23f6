"""Run-to-run reproducibility of the FastSpeech2 step (same weights, same batch, dropout off): which outputs /
gradients are bit-identical across two runs, and how large the spread of the others is.

  python tools/debug_determinism.py [small|C2_8|C2] [eager|graph]

The activation / input-gradient chain holds no atomics (BatchNorm statistics and the BatchNorm backward reduction are
summed in a fixed order), so outputs, losses and every gradient that is not itself accumulated with fp32 atomics
(split-K weight gradients, LayerNorm / bias column sums, embedding scatter-adds) must be bit-identical; those
accumulate in arrival order and differ by ~1e-7 relative.  FS2_OVERLAP=0 / FS2_NO_PDL=1 / FS2_NO_DEFER=1 are the
switches to bisect a race with."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs2b200 import sub  # noqa: E402

synth = sub("synthetic")  # the package's own generators (oracle/ is for tests only)
from tests.util_parity import cuda_batch, disable_dropout, rel_err  # noqa: E402

M, ops, rt = sub("lightning.model"), sub("ops"), sub("runtime")
case = sys.argv[1] if len(sys.argv) > 1 else "small"
mode = sys.argv[2] if len(sys.argv) > 2 else "eager"


def build():
    if case == "small":
        cfg = synth.model_cfg(encoder_layer=2, decoder_layer=2)
        model = M.FastSpeech2(cfg)
        model.load_state_dict(synth.init_state_dict(model.state_dict(), 0))
        return disable_dropout(model.cuda().train()), M.FastSpeech2Loss(cfg)
    cfg, model, loss_fn, _ = synth.build_config("C2", M, device="cuda")
    return disable_dropout(model), loss_fn


batch = synth.make_batch(B=4, src_len=(10, 40), dur=synth.uniform_dur(1, 8), seed=21) if case == "small" \
    else synth.make_batch(**synth.CONFIGS[case])


def run():
    model, loss_fn = build()
    if mode == "graph":
        step = rt.TrainStep(model, loss_fn, batch, use_graph=True)
        step.step_e2e(batch)
        l = step.step_e2e(batch).clone()
        g = {k: p.main_grad.detach().clone() for k, p in model.named_parameters() if p.requires_grad}
        return [], [x for x in l], g
    b = cuda_batch(batch)
    out = model(b[2], b[3], *b[4:12], lang_args=b[12])
    l = loss_fn(b[:-1], out)
    l[0].backward()
    torch.cuda.synchronize()
    g = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    return [o.detach().clone() for o in out[:5]], [x.detach().clone() for x in l], g


o1, l1, g1 = run()
o2, l2, g2 = run()
print("case=%s mode=%s OVERLAP=%s NO_PDL=%s NO_DEFER=%s" % (case, mode, os.environ.get("FS2_OVERLAP", "1"),
                                                          os.environ.get("FS2_NO_PDL"), os.environ.get("FS2_NO_DEFER")))
for n, a, b in zip(["mel", "postnet_mel", "p_pred", "e_pred", "log_d_pred"], o1, o2):
    print("  out %-12s equal=%s rel=%.2e" % (n, torch.equal(a, b), rel_err(a, b)))
print("  loss differences", [float(a - b) for a, b in zip(l1, l2)])
rows = sorted(((rel_err(g1[k], g2[k]), k) for k in g1 if float(g2[k].norm()) > 0), reverse=True)
print("  bit-identical grads: %d of %d; worst rel %.2e" % (sum(torch.equal(g1[k], g2[k]) for k in g1), len(g1),
                                                         rows[0][0] if rows else 0.0))
for r, k in rows[:6]:
    print("    %.2e %s |g|=%.3e" % (r, k, float(g1[k].norm())))
