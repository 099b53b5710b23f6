import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs2b200 import sub
from oracle import synth, fs2_oracle
from tests.util_parity import disable_dropout
M, rt, ops = sub("lightning.model"), sub("runtime"), sub("ops")
cfg = synth.model_cfg(encoder_layer=2, decoder_layer=2)
dev = torch.device("cuda")
def build():
    m = M.FastSpeech2(cfg); m.load_state_dict(synth.init_state_dict(m.state_dict(), 0))
    return disable_dropout(m.to(dev).train()), M.FastSpeech2Loss(cfg)
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 50
batch = synth.make_batch(B=4, src_len=(10, 40), dur=synth.uniform_dur(1, 8), seed=seed)
flats = []
for i in range(2):
    m, l = build(); b = rt.GradBuckets(m.parameters(), device=dev)
    s = rt.TrainStep(m, l, batch, use_graph=False, buckets=b, device=dev); s.run(); torch.cuda.synchronize()
    flats.append((b.flat.clone(), {k: p.main_grad.clone() for k, p in m.named_parameters() if hasattr(p, "main_grad")}))
rel = lambda a, b: ((a - b).norm() / (b.norm() + 1e-30)).item()
print("seed", seed, "flat rel", rel(flats[0][0], flats[1][0]))
rows = sorted(((rel(flats[0][1][k], flats[1][1][k]), k, flats[1][1][k].norm().item()) for k in flats[0][1]), reverse=True)[:8]
for r in rows: print("  %.2e %-70s |g|=%.3e" % r)
# against the fp32 oracle
sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
o_out, o_losses, o_grads = fs2_oracle.step(sd, cfg, batch)
rows = sorted(((rel(flats[1][1][k].cpu(), o_grads[k]), k, o_grads[k].norm().item()) for k in o_grads if o_grads[k] is not None and k in flats[1][1]), reverse=True)[:8]
print("vs oracle:")
for r in rows: print("  %.2e %-70s |g|=%.3e" % r)
