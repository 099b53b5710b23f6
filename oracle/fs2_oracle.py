"""ORACLE (test infrastructure, NOT product code): a plain-PyTorch fp32 restatement of the reference's
FastSpeech2 forward + loss, written functionally over the reference's state_dict keys.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this.  The product path (few-shot-cross-lingual-tts_b200/) never does.

Pinned against the UNMODIFIED reference code: oracle/make_golden.py imports /root/reference through
oracle/ref_loader.py (in the build container, where /root/reference exists), runs both on the same
seeded inputs and asserts they agree to fp32 round-off before writing tests/golden/*.pt;
tests/test_oracle.py re-checks this restatement against those fixtures everywhere (no GPU needed).
The reference itself ships no tests / golden vectors for this path.

Every function cites the reference lines it restates.  Dropout is the identity here (parity mode,
SURVEY.md appendix C.4); BatchNorm runs in training mode (batch statistics over all B*T positions).
"""
import math

import torch
import torch.nn.functional as F


def mask_from_lengths(lengths, max_len):
    """True = padding.  lightning/utils/tool.py:63-74 (in-repo copy of dlhlp_lib's function)."""
    return torch.arange(int(max_len), device=lengths.device)[None, :] >= lengths[:, None]


def length_regulate(x, duration, max_len):
    """lightning/model/modules.py:169-196 + lightning/utils/tool.py:168-186, loop form."""
    outs, lens = [], []
    for xb, db in zip(x, duration):
        reps = [max(int(d), 0) for d in db.tolist()]
        rows = [xb[i:i + 1].expand(r, -1) for i, r in enumerate(reps)]
        e = torch.cat(rows, 0) if rows else xb[:0]
        lens.append(e.shape[0])
        outs.append(e)
    if max_len is None:
        max_len = max(lens)
    padded = [F.pad(e, (0, 0, 0, max_len - e.shape[0])) for e in outs]  # negative pad crops, like F.pad
    return torch.stack(padded), torch.tensor(lens, dtype=torch.int64, device=x.device)


def _mha(sd, pre, x, key_mask, n_head):
    """transformer/SubLayers.py:29-57 + transformer/Modules.py:14-25 (dropout = identity)."""
    B, T, D = x.shape
    dk = sd[pre + "w_qs.weight"].shape[0] // n_head
    q = F.linear(x, sd[pre + "w_qs.weight"], sd[pre + "w_qs.bias"]).view(B, T, n_head, dk)
    k = F.linear(x, sd[pre + "w_ks.weight"], sd[pre + "w_ks.bias"]).view(B, T, n_head, dk)
    v = F.linear(x, sd[pre + "w_vs.weight"], sd[pre + "w_vs.bias"]).view(B, T, n_head, dk)
    q, k, v = (t.permute(2, 0, 1, 3).reshape(n_head * B, T, dk) for t in (q, k, v))
    attn = torch.bmm(q, k.transpose(1, 2)) / math.sqrt(dk)
    m = key_mask[:, None, :].expand(-1, T, -1).repeat(n_head, 1, 1)
    attn = torch.softmax(attn.masked_fill(m, float("-inf")), dim=2)
    o = torch.bmm(attn, v).view(n_head, B, T, dk).permute(1, 2, 0, 3).reshape(B, T, n_head * dk)
    o = F.linear(o, sd[pre + "fc.weight"], sd[pre + "fc.bias"])
    return F.layer_norm(o + x, (D,), sd[pre + "layer_norm.weight"], sd[pre + "layer_norm.bias"])


def _ffn(sd, pre, x):
    """transformer/SubLayers.py:85-93."""
    D = x.shape[-1]
    w1, w2 = sd[pre + "w_1.weight"], sd[pre + "w_2.weight"]
    h = F.conv1d(x.transpose(1, 2), w1, sd[pre + "w_1.bias"], padding=(w1.shape[2] - 1) // 2)
    h = F.conv1d(F.relu(h), w2, sd[pre + "w_2.bias"], padding=(w2.shape[2] - 1) // 2).transpose(1, 2)
    return F.layer_norm(h + x, (D,), sd[pre + "layer_norm.weight"], sd[pre + "layer_norm.bias"])


def fft_stack(sd, pre, x, mask, n_layers, n_head):
    """transformer/Layers.py:21-30 applied n_layers times (Models.py:159-166 / 230-237)."""
    for i in range(n_layers):
        p = "%slayer_stack.%d." % (pre, i)
        x = _mha(sd, p + "slf_attn.", x, mask, n_head).masked_fill(mask[..., None], 0)
        x = _ffn(sd, p + "pos_ffn.", x).masked_fill(mask[..., None], 0)
    return x


def variance_predictor(sd, pre, x, mask):
    """lightning/model/modules.py:244-252 (conv1d_2 has literal padding=1, :232)."""
    c = pre + "conv_layer."
    w1 = sd[c + "conv1d_1.conv.weight"]
    h = F.conv1d(x.transpose(1, 2), w1, sd[c + "conv1d_1.conv.bias"], padding=(w1.shape[2] - 1) // 2)
    h = F.relu(h).transpose(1, 2)
    h = F.layer_norm(h, (h.shape[-1],), sd[c + "layer_norm_1.weight"], sd[c + "layer_norm_1.bias"])
    h = F.conv1d(h.transpose(1, 2), sd[c + "conv1d_2.conv.weight"], sd[c + "conv1d_2.conv.bias"], padding=1)
    h = F.relu(h).transpose(1, 2)
    h = F.layer_norm(h, (h.shape[-1],), sd[c + "layer_norm_2.weight"], sd[c + "layer_norm_2.bias"])
    out = F.linear(h, sd[pre + "linear_layer.weight"], sd[pre + "linear_layer.bias"]).squeeze(-1)
    return out.masked_fill(mask, 0.0) if mask is not None else out


def postnet(sd, pre, x, n_layers=5, running=None, training=True):
    """transformer/Layers.py:129-137; train-mode BatchNorm over all B*T positions (padding included), or -- with
    training=False -- eval-mode BatchNorm on the running statistics stored in `sd` (F.dropout(..., self.training) is
    the identity there)."""
    x = x.transpose(1, 2)
    for i in range(n_layers):
        c = "%sconvolutions.%d." % (pre, i)
        w = sd[c + "0.conv.weight"]
        x = F.conv1d(x, w, sd[c + "0.conv.bias"], padding=(w.shape[2] - 1) // 2)
        rm = rv = None
        if not training:
            rm, rv = sd[c + "1.running_mean"], sd[c + "1.running_var"]
        elif running is not None:
            rm, rv = running[c + "1.running_mean"], running[c + "1.running_var"]
        x = F.batch_norm(x, rm, rv, sd[c + "1.weight"], sd[c + "1.bias"], training=training, momentum=0.1,
                         eps=1e-5)
        if i < n_layers - 1:
            x = torch.tanh(x)
    return x.transpose(1, 2)


def sinusoid_table(n_position, d_hid):
    """transformer/Models.py:10-30 with padding_idx=None (the eval-mode long-sequence branch, :148-153, :211-218)."""
    import numpy as np

    pos = np.arange(n_position)[:, None].astype(np.float64)
    hid = np.arange(d_hid)[None, :]
    ang = pos / np.power(10000, 2 * (hid // 2) / d_hid)
    ang[:, 0::2] = np.sin(ang[:, 0::2])
    ang[:, 1::2] = np.cos(ang[:, 1::2])
    return torch.FloatTensor(ang)


def forward(sd, cfg, speaker_args, texts, src_lens, max_src_len, mels=None, mel_lens=None, max_mel_len=None,
            p_targets=None, e_targets=None, d_targets=None, lang_args=None, average_spk_emb=False,
            running=None, training=True, p_control=1.0, e_control=1.0, d_control=1.0, inject=None):
    """lightning/model/fastspeech2m.py:48-163.  training=True: train-mode semantics (decoder truncation to
    max_seq_len, Models.py:220-228, batch-statistics BatchNorm); training=False: eval mode (no truncation, sinusoid
    table rebuilt for sequences longer than max_seq_len, running-statistics BatchNorm).  Without targets the
    predictions drive the embeddings / durations (modules.py:84-91,133-139); `inject` = {"pitch", "energy",
    "duration"} optionally replaces those predictions by externally supplied values (the parity tests feed the CUDA
    path's own predictions so that bucket / rounding decisions are identical on both sides).  Pitch / energy are
    phoneme- or frame-level per cfg (modules.py:117-128,141-150).  Returns the reference's 10-tuple."""
    t = cfg["transformer"]
    inject = inject or {}
    max_src_len = int(max_src_len)
    src_masks = mask_from_lengths(src_lens, max_src_len)
    mel_masks = mask_from_lengths(mel_lens, int(max_mel_len)) if mel_lens is not None else None
    # Encoder2 (Models.py:139-166)
    if not training and max_src_len > cfg["max_seq_len"]:
        x = texts + sinusoid_table(max_src_len, t["encoder_hidden"])[None].to(texts.device)
    else:
        x = texts + sd["encoder.position_enc"][:, :max_src_len]
    x = fft_stack(sd, "encoder.", x, src_masks, t["encoder_layer"], t["encoder_head"])
    spk = None
    if "speaker_emb.model.weight" in sd:  # fastspeech2m.py:84-89 (table embedding)
        spk = sd["speaker_emb.model.weight"][speaker_args]
        if average_spk_emb:
            spk = spk.mean(0, keepdim=True).expand(x.shape[0], -1)
        x = x + spk[:, None, :]
    if "language_emb.model.weight" in sd and lang_args is not None:  # :98-101
        x = x + sd["language_emb.model.weight"][lang_args][:, None, :]
    # VarianceAdaptor (modules.py:104-160)
    va = "variance_adaptor."

    def variance(name, x, target, mask, control):  # modules.py:82-102
        pred = variance_predictor(sd, va + name + "_predictor.", x, mask)
        if target is None:
            pred = pred * control
            target = inject.get(name, pred.detach())
        emb = sd[va + name + "_embedding.weight"][torch.bucketize(target, sd[va + name + "_bins"])]
        return pred, x + emb

    log_d = variance_predictor(sd, va + "duration_predictor.", x, src_masks)
    p_pred = e_pred = None
    if cfg["pitch"]["feature"] == "phoneme_level":
        p_pred, x = variance("pitch", x, p_targets, src_masks, p_control)
    if cfg["energy"]["feature"] == "phoneme_level":
        e_pred, x = variance("energy", x, e_targets, src_masks, e_control)
    if d_targets is not None:
        d_rounded = d_targets
        x, mel_len = length_regulate(x, d_targets, max_mel_len)
    else:
        d_rounded = torch.clamp(torch.round(torch.exp(log_d.detach()) - 1) * d_control, min=0)
        d_rounded = inject.get("duration", d_rounded)
        x, mel_len = length_regulate(x, d_rounded, max_mel_len)
        mel_masks = mask_from_lengths(mel_len, x.shape[1])
    if cfg["pitch"]["feature"] == "frame_level":
        p_pred, x = variance("pitch", x, p_targets, mel_masks, p_control)
    if cfg["energy"]["feature"] == "frame_level":
        e_pred, x = variance("energy", x, e_targets, mel_masks, e_control)
    if spk is not None:  # :132-136 (max(mel_lens) == x.shape[1])
        x = x + spk[:, None, :]
    # Decoder (Models.py:205-237)
    if not training and x.shape[1] > cfg["max_seq_len"]:
        x = x + sinusoid_table(x.shape[1], t["decoder_hidden"])[None].to(x.device)
    else:
        T = min(x.shape[1], cfg["max_seq_len"])
        x = x[:, :T] + sd["decoder.position_enc"][:, :T]
        mel_masks = mel_masks[:, :T]
    x = fft_stack(sd, "decoder.", x, mel_masks, t["decoder_layer"], t["decoder_head"])
    mel = F.linear(x, sd["mel_linear.weight"], sd["mel_linear.bias"])
    post = postnet(sd, "postnet.", mel, running=running, training=training) + mel
    return (mel, post, p_pred, e_pred, log_d, d_rounded, src_masks, mel_masks, src_lens, mel_len)


def loss(inputs, predictions, pitch_level="phoneme_level", energy_level="phoneme_level"):
    """lightning/model/loss.py:15-89 (pitch / energy masked by the source or the mel mask, :47-60)."""
    mel_t, _, _, p_t, e_t, d_t = inputs[6:12]
    mel, post, p_pred, e_pred, log_d, _, src_masks, mel_masks, _, _ = predictions
    sm, mm = ~src_masks, ~mel_masks
    log_d_t = torch.log(d_t.float() + 1)
    mel_t = mel_t[:, : mm.shape[1], :]
    l1 = lambda a, b: (a.masked_select(mm[..., None]) - b.masked_select(mm[..., None])).abs().mean()
    mse = lambda a, b, m=sm: ((a.masked_select(m) - b.masked_select(m).float()) ** 2).mean()
    mel_loss, post_loss = l1(mel, mel_t), l1(post, mel_t)
    pitch_loss = mse(p_pred, p_t, sm if pitch_level == "phoneme_level" else mm)
    energy_loss = mse(e_pred, e_t, sm if energy_level == "phoneme_level" else mm)
    dur_loss = mse(log_d, log_d_t)
    total = mel_loss + post_loss + dur_loss + pitch_loss + energy_loss
    return total, mel_loss, post_loss, pitch_loss, energy_loss, dur_loss


def step(sd, cfg, batch, grad_keys=None):
    """One fwd + loss + bwd on a 13-tuple batch (collates/utils.py:70-85) whose texts slot already holds
    the embedded phonemes [B,Ts,d].  Returns (predictions, losses, {key: grad})."""
    params = {k: v for k, v in sd.items() if v.is_floating_point() and "position_enc" not in k
              and not k.endswith("_bins") and "running_" not in k}
    for v in params.values():
        v.requires_grad_(True)
    out = forward(sd, cfg, batch[2], batch[3], *batch[4:12], lang_args=batch[12])
    losses = loss(batch[:12], out, cfg["pitch"]["feature"], cfg["energy"]["feature"])
    keys = list(params) if grad_keys is None else list(grad_keys)
    grads = torch.autograd.grad(losses[0], [params[k] for k in keys], allow_unused=True)
    for v in params.values():
        v.requires_grad_(False)
    return out, losses, dict(zip(keys, grads))


def ada_loss(mel_targets, predictions):
    """FastSpeech2ADALoss, lightning/model/loss.py:104-140: L1(mel) + L1(postnet mel) over the valid mel elements."""
    mel, post, mel_masks = predictions
    mm = ~mel_masks
    mel_t = mel_targets[:, : mm.shape[1], :]
    l1 = lambda a: (a.masked_select(mm[..., None]) - mel_t.masked_select(mm[..., None])).abs().mean()
    mel_loss, post_loss = l1(mel), l1(post)
    return mel_loss + post_loss, mel_loss, post_loss
