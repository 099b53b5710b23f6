"""ORACLE-side test infrastructure: import the UNMODIFIED reference hot path from /root/reference.

Only usable where /root/reference exists (the build container); used by make_golden.py and by
tests/test_oracle.py::test_oracle_matches_reference (skipped elsewhere).  Recipe: SURVEY.md appendix B.
Nothing from the reference is copied: missing third-party modules are stubbed, `pad` is exec'd from
the reference file at import time.
"""
import json
import os
import sys
import types

import torch
import torch.nn as nn

REF = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF, "lightning", "model"))


def load():
    """Returns (FastSpeech2, FastSpeech2Loss, LengthRegulator) classes of the reference."""
    if "fs2_ref_loaded" in sys.modules:
        m = sys.modules["fs2_ref_loaded"]
        return m.FastSpeech2, m.FastSpeech2Loss, m.LengthRegulator
    saved = {k: sys.modules.get(k) for k in ("lightning", "lightning.utils", "lightning.utils.tool", "Define",
                                              "text", "text.symbols", "transformer")}

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class LightningModule(nn.Module):
        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def freeze(self):
            for p in self.parameters():
                p.requires_grad = False

    mod("pytorch_lightning", LightningModule=LightningModule)

    def get_mask_from_lengths(lengths, max_len=None):
        if max_len is None:
            max_len = int(torch.max(lengths).item())
        return torch.arange(0, int(max_len), device=lengths.device)[None, :] >= lengths[:, None]

    mod("dlhlp_lib")
    mod("dlhlp_lib.utils")
    mod("dlhlp_lib.utils.tool", get_mask_from_lengths=get_mask_from_lengths)
    mod("dlhlp_lib.utils.numeric", torch_exist_nan=lambda x: bool((x != x).any()))
    mod("dlhlp_lib.audio", AUDIO_CONFIG={"mel": {"n_mel_channels": 80}, "audio": {"sampling_rate": 22050},
                                         "stft": {"hop_length": 256}})
    mod("resemblyzer", VoiceEncoder=type("VoiceEncoder", (nn.Module,), {}))
    t = mod("text")
    t.__path__ = []
    mod("text.symbols", symbols=["s%d" % i for i in range(300)])
    stats = json.load(open(os.path.join(REF, "stats.json")))
    mod("Define", ALLSTATS={"global": stats["pitch"] + stats["energy"]}, DEVICE=torch.device("cpu"),
        NOLID=False, DEBUG=False)
    lt = mod("lightning")
    lt.__path__ = [os.path.join(REF, "lightning")]
    lu = mod("lightning.utils")
    lu.__path__ = [os.path.join(REF, "lightning", "utils")]
    src = open(os.path.join(REF, "lightning", "utils", "tool.py")).read().split("\n")
    start = next(i for i, l in enumerate(src) if l.startswith("def pad("))
    end = next(i for i in range(start + 1, len(src)) if src[i].startswith("def "))
    ns = {"torch": torch, "F": torch.nn.functional}
    exec("\n".join(src[start:end]), ns)
    mod("lightning.utils.tool", pad=ns["pad"])
    sys.path.insert(0, REF)
    for k in [k for k in sys.modules if k == "transformer" or k.startswith("transformer.")]:
        del sys.modules[k]
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        from lightning.model.fastspeech2m import FastSpeech2  # noqa: E402
        from lightning.model.loss import FastSpeech2Loss  # noqa: E402
        from lightning.model.modules import LengthRegulator  # noqa: E402
    sys.path.remove(REF)
    holder = types.ModuleType("fs2_ref_loaded")
    holder.FastSpeech2, holder.FastSpeech2Loss, holder.LengthRegulator = FastSpeech2, FastSpeech2Loss, LengthRegulator
    sys.modules["fs2_ref_loaded"] = holder
    # keep the reference's `lightning` / `transformer` / `Define` private to the loaded classes
    for k in list(sys.modules):
        if k.split(".")[0] in ("lightning", "transformer", "Define", "text"):
            sys.modules["_fs2ref_" + k] = sys.modules.pop(k)
    for k, v in saved.items():
        if v is not None:
            sys.modules[k] = v
    return FastSpeech2, FastSpeech2Loss, LengthRegulator


def build_reference_model(cfg, spk_config=None, train=True, no_dropout=True):
    import contextlib
    import io

    FastSpeech2, FastSpeech2Loss, _ = load()
    with contextlib.redirect_stdout(io.StringIO()):
        m = FastSpeech2(cfg, spk_config=spk_config) if spk_config else FastSpeech2(cfg)
    m.train(train)
    if no_dropout:  # parity mode: dropout = identity while BatchNorm stays in batch-stat mode
        for mm in m.modules():
            if isinstance(mm, nn.Dropout):
                mm.p = 0.0
    return m, FastSpeech2Loss(cfg)


class no_functional_dropout:
    """PostNet calls F.dropout(x, 0.5, self.training) directly (transformer/Layers.py:133-134)."""

    def __enter__(self):
        self._orig = torch.nn.functional.dropout
        torch.nn.functional.dropout = lambda x, p=0.5, training=True, inplace=False: x
        return self

    def __exit__(self, *a):
        torch.nn.functional.dropout = self._orig
