"""First-order few-shot adaptation step (runtime.FirstOrderTaskStep; BASELINE.json configs[2])
against the same loop written with plain torch on the oracle restatement of the reference modules: k SGD steps on
support mini-batches, query forward / backward at the adapted weights, outer gradient = query gradient, theta
restored.  (The live reference never adapts -- fscl-orig.yaml:40-43 -- so the oracle IS the specification here.)"""
import pytest
import torch

from fs2b200 import sub
from oracle import fs2_oracle, synth
from tests.util_parity import cos_floor, cosine, disable_dropout

pytestmark = pytest.mark.gpu


def test_first_order_inner_loop_matches_torch_loop_on_the_oracle():
    M, rt = sub("lightning.model"), sub("runtime")
    cfg = synth.model_cfg(multi_speaker=True, multi_lingual=True, max_seq_len=1500, encoder_layer=2, decoder_layer=2)
    spk = {"emb_type": "table", "speakers": list(range(11))}
    model = M.FastSpeech2(cfg, spk_config=spk)
    model.load_state_dict(synth.init_state_dict(model.state_dict(), 0))
    model = disable_dropout(model.cuda().train())
    loss_fn = M.FastSpeech2Loss(cfg)
    kw = dict(average_spk_emb=True)
    k, lr = 3, 0.002  # large enough for three plain-SGD steps to matter (8.50 -> 5.21 on the oracle), small enough to be stable
    support = [synth.make_batch(B=8, src_len=(20, 70), dur=synth.uniform_dur(2, 9), seed=300 + i, n_speaker=11,
                                n_lang=8) for i in range(k)]
    # the query repeats the first support batch: three SGD steps at this rate must lower its loss visibly
    query = synth.make_batch(B=8, src_len=(20, 70), dur=synth.uniform_dur(2, 9), seed=300, n_speaker=11, n_lang=8)
    sd0 = {n: v.detach().cpu().clone() for n, v in model.state_dict().items()}

    task = rt.FirstOrderTaskStep(model, loss_fn, inner_lr=lr, model_kwargs=kw)
    theta_before = task.flat_param.clone()
    losses = task.run(support, query).cpu()
    torch.cuda.synchronize()
    assert torch.equal(task.flat_param, theta_before)  # theta is restored bit-exactly
    outer = {n: p.main_grad.detach().cpu().clone() for n, p in model.named_parameters() if p.requires_grad}

    # ---- the same loop in plain torch on the oracle (host CPU, fp32)
    def grads_of(sd, batch):
        params = {n: v for n, v in sd.items() if v.is_floating_point() and "position_enc" not in n
                  and not n.endswith("_bins") and "running_" not in n}
        for v in params.values():
            v.requires_grad_(True)
        out = fs2_oracle.forward(sd, cfg, batch[2], batch[3], *batch[4:12], lang_args=batch[12], **kw)
        ls = fs2_oracle.loss(batch[:12], out)
        g = torch.autograd.grad(ls[0], list(params.values()), allow_unused=True)
        for v in params.values():
            v.requires_grad_(False)
        return ls, dict(zip(params.keys(), g))

    sd = {n: v.clone() for n, v in sd0.items()}
    unadapted_loss = float(grads_of(sd, query)[0][0])
    for sb in support:
        _, g = grads_of(sd, sb)
        for n, gi in g.items():
            if gi is not None:
                sd[n] = sd[n] - lr * gi
    o_losses, o_grads = grads_of(sd, query)
    adapted_loss = float(o_losses[0])
    assert adapted_loss < unadapted_loss - 0.02 * abs(unadapted_loss), (adapted_loss, unadapted_loss)
    assert abs(float(losses[0]) - adapted_loss) <= 1e-2 * abs(adapted_loss), (float(losses[0]), adapted_loss)
    gmax = max(float(g.norm()) for g in o_grads.values() if g is not None)
    worst = (1.0, None)
    for n, g in outer.items():
        r = o_grads.get(n)
        if r is None or float(r.norm()) < 1e-3 * gmax:
            continue
        c = cosine(g, r)
        worst = min(worst, (c, n))
        assert c >= min(cos_floor(n), 0.99), (n, c)
    print("first-order outer gradient: worst cosine", worst, "loss", float(losses[0]), "oracle", adapted_loss,
          "unadapted", unadapted_loss)
