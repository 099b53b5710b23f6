"""Headless FastSpeech2 (takes already-embedded phonemes) behind the reference's constructor, forward
signature, 10-tuple result and state_dict keys (reference: lightning/model/fastspeech2m.py:19-163).
"""
import torch
import torch.nn as nn

from ... import Define, ops
from ..._act import to_act
from ...transformer import Decoder, Encoder2, PostNet
from ..utils.tool import get_mask_from_lengths
from .modules import VarianceAdaptor
from .speaker_encoder import LanguageEncoder, SpeakerEncoder

N_MEL_CHANNELS = 80  # dlhlp_lib AUDIO_CONFIG["mel"]["n_mel_channels"] (config/preprocess/*.yaml)


class FastSpeech2(nn.Module):
    def __init__(self, model_config, **kwargs):
        super().__init__()
        self.model_config = model_config
        self.encoder = Encoder2(model_config)
        self.variance_adaptor = VarianceAdaptor(model_config)
        self.decoder = Decoder(model_config)
        self.mel_linear = nn.Linear(model_config["transformer"]["decoder_hidden"],
                                    kwargs.get("n_mel_channels", N_MEL_CHANNELS))
        self.postnet = PostNet(n_mel_channels=kwargs.get("n_mel_channels", N_MEL_CHANNELS))
        self.speaker_emb = None
        if model_config.get("multi_speaker", False):
            self.speaker_emb = SpeakerEncoder(model_config, kwargs["spk_config"])
        self.language_emb = None
        if model_config.get("multi_lingual", False):
            self.language_emb = LanguageEncoder(model_config, {"emb_type": "table"})

    @property
    def device(self):
        return next(self.parameters()).device

    def _speaker_rows(self, speaker_args, B, average_spk_emb):
        spk = self.speaker_emb(speaker_args)
        if average_spk_emb:
            spk = spk.mean(dim=0, keepdim=True).expand(B, -1)
        return spk

    def forward(self, speaker_args, texts, src_lens, max_src_len, mels=None, mel_lens=None, max_mel_len=None,
                p_targets=None, e_targets=None, d_targets=None, lang_args=None, p_control=1.0, e_control=1.0,
                d_control=1.0, average_spk_emb=False):
        dev = self.device
        if not texts.is_cuda:
            to_act(texts)  # raises: there is no CPU path
        B = texts.shape[0]
        max_src_len = int(max_src_len)
        src_lens = src_lens.to(dev, torch.int64)
        src_masks = get_mask_from_lengths(src_lens, max_src_len)
        mel_masks = None
        if mel_lens is not None:
            mel_lens = mel_lens.to(dev, torch.int64)
            mel_masks = get_mask_from_lengths(mel_lens, int(max_mel_len))

        x = self.encoder(texts, src_masks, lens=src_lens, out_dtype=torch.bfloat16)
        if x.dtype != torch.bfloat16:
            x = x.to(torch.bfloat16)
        spk = None
        if self.speaker_emb is not None:
            spk = self._speaker_rows(speaker_args, B, average_spk_emb)
            x = ops.AddRowVec.apply(x, spk)
        if not Define.NOLID and self.language_emb is not None and lang_args is not None:
            x = ops.AddRowVec.apply(x, self.language_emb(lang_args))

        # variance adaptor; the `+ speaker` after it, the decoder's truncation and `+ position_enc`
        # (fastspeech2m.py:132-141, Models.py:220-226) are folded into the LengthRegulator gather.
        # (frame-level pitch / energy run their predictors on the LengthRegulator output, modules.py:141-150, so the
        # adds cannot be folded into the gather there)
        va = self.variance_adaptor
        frame_level = "frame_level" in (va.pitch_feature_level, va.energy_feature_level)
        teacher = d_targets is not None and max_mel_len is not None and not frame_level
        if teacher:
            T_lr = int(max_mel_len)
            t_dec = self.decoder.out_len(T_lr)
            if not self.decoder.training and T_lr > self.decoder.max_seq_len:
                pos_table = self.decoder._table_for(T_lr, dev)[0]
            else:
                pos_table = self.decoder.position_enc
            spk2 = spk  # the same rows as above (one gather, one backward)
            fused = (t_dec, spk2, pos_table)
        else:
            fused = None
        (x, p_pred, e_pred, log_d_pred, d_rounded, mel_lens_out, mel_masks) = self.variance_adaptor(
            x, src_masks, mel_masks, max_mel_len, p_targets, e_targets, d_targets, p_control, e_control,
            d_control, src_lens=src_lens, _fused=fused)

        # decoder key mask / frame validity: the INPUT mel_lens when they are given (the reference builds mel_masks
        # from them, fastspeech2m.py:70-74), the LengthRegulator's lengths otherwise (inference)
        len_src = mel_lens if mel_lens is not None else mel_lens_out
        if fused is not None:
            dec_lens = torch.clamp(len_src, max=t_dec)
            x = self.decoder.forward_prepared(x, dec_lens)
            mel_masks = mel_masks[:, :t_dec] if mel_masks is not None else None
        else:
            if self.speaker_emb is not None:
                x = ops.AddRowVec.apply(x, self._speaker_rows(speaker_args, B, average_spk_emb))
            dec_lens = torch.clamp(len_src, max=x.shape[1])
            x, mel_masks = self.decoder(x, mel_masks, lens=dec_lens)
            if x.dtype != torch.bfloat16:
                x = x.to(torch.bfloat16)
        mel = ops.LinearF32Out.apply(x, self.mel_linear.weight, self.mel_linear.bias)
        postnet_mel = self.postnet.forward_residual(mel)
        ops.join_branches()  # the variance predictors ran next to the decoder / PostNet (ops.branch)
        return (mel, postnet_mel, p_pred, e_pred, log_d_pred, d_rounded, src_masks, mel_masks, src_lens,
                mel_lens_out)
