"""Wall-clock / device timing of one FastSpeech2 fwd+loss+bwd through the boundary modules."""
import cProfile
import os
import pstats
import sys
import time

t0 = time.time()
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs2b200 import sub  # noqa: E402
synth = sub("synthetic")  # the package's own generators (oracle/ is for tests only)
from tests.util_parity import cuda_batch, disable_dropout  # noqa: E402

print("import %.1fs" % (time.time() - t0)); t0 = time.time()
M = sub("lightning.model")
cabi = sub("_cabi")
name = sys.argv[1] if len(sys.argv) > 1 else "C1"
cfg = synth.model_cfg(multi_speaker="n_speaker" in synth.CONFIGS[name])
spk = {"emb_type": "table", "speakers": list(range(247))} if cfg["multi_speaker"] else None
model = M.FastSpeech2(cfg, spk_config=spk) if spk else M.FastSpeech2(cfg)
model = model.cuda().train()
if "--nodrop" in sys.argv:
    disable_dropout(model)
loss_fn = M.FastSpeech2Loss(cfg)
batch = cuda_batch(synth.make_batch(**synth.CONFIGS[name]))
print("setup %.1fs  B=%d Ts=%d Tm=%d frames=%d" % (time.time() - t0, batch[3].shape[0], batch[5], batch[8],
                                                   synth.count_real_frames(batch, cfg["max_seq_len"])))


def step():
    out = model(batch[2], batch[3], *batch[4:12])
    losses = loss_fn(batch[:-1], out)
    model.zero_grad(set_to_none=True)
    losses[0].backward()
    return losses


for i in range(3):
    t0 = time.time(); n0 = cabi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); l = step(); e1.record(); torch.cuda.synchronize()
    print("step %d: wall %.3fs device %.2f ms launches %d loss %.4f" % (i, time.time() - t0, e0.elapsed_time(e1),
                                                                        cabi.launch_count() - n0, float(l[0])))
pr = cProfile.Profile(); pr.enable(); step(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
