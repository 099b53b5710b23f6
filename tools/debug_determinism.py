"""Run-to-run reproducibility of the eager FastSpeech2 step (same weights, same batch, dropout off): which
outputs / gradients are bit-identical, and how large is the spread of those that are not."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs2b200 import sub
synth = sub("synthetic")  # the package's own generators (oracle/ is for tests only)
from tests.util_parity import cuda_batch, disable_dropout, rel_err
M = sub("lightning.model"); ops = sub("ops")
def build():
    cfg = synth.model_cfg(encoder_layer=2, decoder_layer=2)
    model = M.FastSpeech2(cfg); model.load_state_dict(synth.init_state_dict(model.state_dict(), 0))
    return disable_dropout(model.cuda().train()), M.FastSpeech2Loss(cfg)
batch = synth.make_batch(B=4, src_len=(10, 40), dur=synth.uniform_dur(1, 8), seed=21)
def eager():
    model, loss_fn = build(); b = cuda_batch(batch)
    out = model(b[2], b[3], *b[4:12], lang_args=b[12]); l = loss_fn(b[:-1], out); l[0].backward()
    torch.cuda.synchronize()
    g = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    return [o.detach().clone() for o in out[:5]], [x.detach().clone() for x in l], g
ops.OVERLAP = len(sys.argv) > 1 and sys.argv[1] == "overlap"
o1, l1, g1 = eager(); o2, l2, g2 = eager()
names = ["mel", "postnet_mel", "p_pred", "e_pred", "log_d_pred"]
for n, a, b in zip(names, o1, o2):
    print("out %-12s equal=%s rel=%.2e" % (n, torch.equal(a, b), rel_err(a, b)))
print("losses", [float(a - b) for a, b in zip(l1, l2)])
rows = sorted(((rel_err(g1[k], g2[k]), k) for k in g1 if float(g2[k].norm()) > 0), reverse=True)
print("bit-identical grads: %d of %d" % (sum(torch.equal(g1[k], g2[k]) for k in g1), len(g1)))
for r, k in rows[:12]:
    print("%.2e %s |g|=%.3e" % (r, k, float(g1[k].norm())))
