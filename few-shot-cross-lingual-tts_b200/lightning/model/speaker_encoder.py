"""Speaker / language embedding tables (reference: lightning/model/speaker_encoder.py:102-167).

Only the `table` and `shared` speaker embeddings and the language table are on the FastSpeech2 hot
path; the d-vector / GE2E encoders are frozen third-party LSTMs (resemblyzer) and are out of scope.
The lookups are [B]-sized, so they stay plain nn.Embedding calls; the broadcast add over time and its
backward are CUDA kernels (ops.AddRowVec / the fused LengthRegulator epilogue).
"""
import torch
import torch.nn as nn


class SpeakerEncoder(nn.Module):
    def __init__(self, model_config, spk_config):
        super().__init__()
        self.emb_type = spk_config["emb_type"]
        d = model_config["transformer"]["encoder_hidden"]
        if self.emb_type == "table":
            self.model = nn.Embedding(len(spk_config["speakers"]), d)
        elif self.emb_type == "shared":
            self.model = nn.Embedding(1, d)
        else:
            raise NotImplementedError(
                "speaker emb_type %r needs the resemblyzer / GE2E encoders, which are outside the "
                "FastSpeech2 hot path (use 'table' or 'shared')" % self.emb_type)

    def forward(self, args):
        if self.emb_type == "table":
            return self.model(args)
        return self.model(torch.zeros_like(args))


class LanguageEncoder(nn.Module):
    def __init__(self, model_config, lang_config):
        super().__init__()
        self.emb_type = lang_config["emb_type"]
        if self.emb_type != "table":
            raise NotImplementedError
        self.model = nn.Embedding(100, model_config["transformer"]["encoder_hidden"])  # up to 100 languages

    def forward(self, args):
        return self.model(args)
