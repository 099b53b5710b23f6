"""Deterministic synthetic batches and seeded weights for benchmarks and tests.
There is no dataset in the build / bench environment; these generators produce LJSpeech / LibriTTS /
AISHELL-3-shaped batches of the exact tuple layout the reference's collate emits.

Batches have the reference's 13-tuple layout (lightning/collates/utils.py:70-85) except that slot 3
holds the already-embedded phonemes [B, Ts, d] (the live model is "headless", fastspeech2m.py:19-20).
"""
import hashlib
import math

import torch

MODEL_CFG = {  # config/model/fastspeech2.yaml:1-38 (values only; multi_speaker switched per test)
    "transformer": {"encoder_layer": 4, "encoder_head": 2, "encoder_hidden": 256, "decoder_layer": 6,
                    "decoder_head": 2, "decoder_hidden": 256, "conv_filter_size": 1024,
                    "conv_kernel_size": [9, 1], "encoder_dropout": 0.2, "decoder_dropout": 0.2},
    "variance_predictor": {"filter_size": 256, "kernel_size": 3, "dropout": 0.5},
    "variance_embedding": {"pitch_quantization": "linear", "energy_quantization": "linear", "n_bins": 256},
    "pitch": {"feature": "phoneme_level", "normalization": True},
    "energy": {"feature": "phoneme_level", "normalization": True},
    "multi_speaker": False, "max_seq_len": 1000, "speaker_emb": "table",
}


def model_cfg(**over):
    import copy

    cfg = copy.deepcopy(MODEL_CFG)
    for k, v in over.items():
        if k in cfg["transformer"]:
            cfg["transformer"][k] = v
        else:
            cfg[k] = v
    return cfg


def _gen_for(key, seed):
    h = hashlib.sha256(("%s|%d" % (key, seed)).encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:7], "little"))
    return g


def init_state_dict(template, seed=0):
    """Deterministic values for every learnable tensor of a FastSpeech2 state_dict (any implementation
    with the reference's keys).  Frozen tables (position_enc, *_bins) and BN bookkeeping are kept."""
    out = {}
    for k, v in template.items():
        if "position_enc" in k or k.endswith("_bins") or k.endswith("num_batches_tracked"):
            out[k] = v.clone()
        elif k.endswith("running_mean"):
            out[k] = torch.zeros_like(v)
        elif k.endswith("running_var"):
            out[k] = torch.ones_like(v)
        else:
            g = _gen_for(k, seed)
            norm_like = ("layer_norm" in k) or (".1." in k and "postnet" in k)
            if v.dim() == 1 and norm_like and k.endswith("weight"):
                t = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
            elif v.dim() == 1:
                t = 0.1 * torch.randn(v.shape, generator=g)
            elif "embedding" in k or "_emb." in k:
                t = 0.5 * torch.randn(v.shape, generator=g)
            else:
                fan_in = v[0].numel()
                bound = 1.0 / math.sqrt(fan_in)
                t = (torch.rand(v.shape, generator=g) * 2 - 1) * bound
            out[k] = t.to(v.dtype)
    return out


def make_batch(B, src_len, dur, seed, d_model=256, n_mel=80, max_mel=None, n_speaker=None, n_lang=None,
               fixed_src_len=None, device="cpu", energy_f64=False):
    """src_len=(lo,hi) inclusive; dur: callable(gen, n) -> int64 durations for n phonemes."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    if fixed_src_len is not None:
        src_lens = torch.full((B,), fixed_src_len, dtype=torch.int64)
    else:
        src_lens = torch.randint(src_len[0], src_len[1] + 1, (B,), generator=g, dtype=torch.int64)
    Ts = int(src_lens.max())
    durations = torch.zeros(B, Ts, dtype=torch.int64)
    for b in range(B):
        n = int(src_lens[b])
        d = dur(g, n)
        if max_mel is not None:  # clip the utterance to max_mel frames (cleaning step of the corpus)
            c = torch.cumsum(d, 0)
            d = torch.where(c <= max_mel, d, torch.clamp(d - (c - max_mel), min=0))
        durations[b, :n] = d
    mel_lens = durations.sum(1)
    Tm = int(mel_lens.max())
    valid_s = torch.arange(Ts)[None, :] < src_lens[:, None]
    valid_m = torch.arange(Tm)[None, :] < mel_lens[:, None]
    emb = torch.randn(B, Ts, d_model, generator=g) * valid_s[..., None]
    mels = torch.randn(B, Tm, n_mel, generator=g) * valid_m[..., None]
    pitch = torch.randn(B, Ts, generator=g) * valid_s
    energy = torch.randn(B, Ts, generator=g) * valid_s
    if energy_f64:
        energy = energy.double()
    spk = torch.randint(0, n_speaker, (B,), generator=g) if n_speaker else torch.zeros(B, dtype=torch.int64)
    lang = torch.randint(0, n_lang, (B,), generator=g) if n_lang else torch.zeros(B, dtype=torch.int64)
    ids = ["utt%04d" % i for i in range(B)]
    batch = (ids, ["" for _ in ids], spk, emb, src_lens, Ts, mels, mel_lens, Tm, pitch, energy, durations, lang)
    return to_device(batch, device)


def to_device(batch, device):
    return tuple(x.to(device) if torch.is_tensor(x) else x for x in batch)


def uniform_dur(lo, hi):
    return lambda g, n: torch.randint(lo, hi + 1, (n,), generator=g, dtype=torch.int64)


def skewed_dur(g, n):
    """C4: 90 % U{0..3} (zeros included), 10 % U{40..120}."""
    small = torch.randint(0, 4, (n,), generator=g, dtype=torch.int64)
    big = torch.randint(40, 121, (n,), generator=g, dtype=torch.int64)
    pick = torch.rand(n, generator=g) < 0.1
    return torch.where(pick, big, small)


CONFIGS = {
    # name: kwargs for make_batch  
    "C1": dict(B=16, src_len=(60, 140), dur=uniform_dur(1, 11), seed=1),
    "C2": dict(B=64, src_len=(20, 200), dur=uniform_dur(1, 11), seed=2, max_mel=1000, n_speaker=247),
    "C2_8": dict(B=8, src_len=(20, 200), dur=uniform_dur(1, 11), seed=2, max_mel=1000, n_speaker=247),
    "C3": dict(B=8, src_len=(20, 70), dur=uniform_dur(2, 9), seed=3, n_speaker=247, n_lang=8),
    "C4": dict(B=4, src_len=(220, 220), dur=skewed_dur, seed=4, fixed_src_len=220, n_speaker=247, n_lang=8),
    "C5": dict(B=4, src_len=(60, 160), dur=uniform_dur(1, 9), seed=5, max_mel=1000, n_speaker=247, n_lang=8),
}


# model settings per BASELINE.json config (SURVEY.md 8d "Model config"): C1 single-speaker fastspeech2.yaml; C2
# multi-speaker (247 LibriTTS speakers); C3-C5 multilingual-fastspeech2.yaml (multi_lingual, max_seq_len 1500)
CONFIG_MODEL = {
    "C1": dict(),
    "C2": dict(multi_speaker=True),
    "C2_8": dict(multi_speaker=True),
    "C3": dict(multi_speaker=True, multi_lingual=True, max_seq_len=1500),
    "C4": dict(multi_speaker=True, multi_lingual=True, max_seq_len=1500),
    "C5": dict(multi_speaker=True, multi_lingual=True, max_seq_len=1500),
}
CONFIG_CALL = {"C3": dict(average_spk_emb=True)}  # FSCL task step (TransEmbOrig.py:93-126)
N_SPEAKER = 247


def build_config(name, M, device=None, weight_seed=0):
    """(cfg, model, loss_fn, model_kwargs) of a BASELINE.json configuration; M = the lightning.model package."""
    cfg = model_cfg(**CONFIG_MODEL[name])
    spk = {"emb_type": "table", "speakers": list(range(N_SPEAKER))} if cfg.get("multi_speaker") else None
    model = M.FastSpeech2(cfg, spk_config=spk) if spk else M.FastSpeech2(cfg)
    model.load_state_dict(init_state_dict(model.state_dict(), weight_seed))
    if device is not None:
        model = model.to(device)
    return cfg, model.train(), M.FastSpeech2Loss(cfg), dict(CONFIG_CALL.get(name, {}))


def count_real_frames(batch, max_seq_len):
    """mel frames that reach the loss: sum_b min(mel_len_b, max_seq_len)."""
    return int(torch.clamp(batch[7], max=max_seq_len).sum())
