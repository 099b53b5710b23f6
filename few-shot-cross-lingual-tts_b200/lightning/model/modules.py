"""VarianceAdaptor, LengthRegulator, VariancePredictor, Conv with the reference's signatures and
state_dict keys (reference: lightning/model/modules.py:18-298).  Math: ops.VariancePredictorFn,
ops.BucketEmbedAdd, ops.LengthRegulate (sm_100a kernels)."""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from ... import Define, ops
from ..._act import from_act, lens_from_mask, to_act
from ..utils.tool import get_mask_from_lengths


class Conv(nn.Module):
    """Channels-last Conv1d wrapper (modules.py:255-298); parameter container + inference helper."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=0, dilation=1, bias=True,
                 w_init="linear"):
        super().__init__()
        if stride != 1 or dilation != 1:
            raise NotImplementedError("stride / dilation 1 only")
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                              padding=padding, dilation=dilation, bias=bias)

    def forward(self, x):
        k = self.conv.kernel_size[0]
        if self.conv.padding[0] != (k - 1) // 2:
            raise NotImplementedError("'same' padding only")
        xb, dt = to_act(x.contiguous())
        return from_act(ops.conv_fwd(xb, ops.pack_conv(self.conv.weight), self.conv.bias.detach()), dt)


class VariancePredictor(nn.Module):
    """Duration / pitch / energy predictor (modules.py:199-252)."""

    def __init__(self, model_config):
        super().__init__()
        self.input_size = model_config["transformer"]["encoder_hidden"]
        self.filter_size = model_config["variance_predictor"]["filter_size"]
        self.kernel = model_config["variance_predictor"]["kernel_size"]
        self.conv_output_size = model_config["variance_predictor"]["filter_size"]
        self.dropout = model_config["variance_predictor"]["dropout"]
        self.conv_layer = nn.Sequential(OrderedDict([
            ("conv1d_1", Conv(self.input_size, self.filter_size, kernel_size=self.kernel,
                              padding=(self.kernel - 1) // 2)),
            ("relu_1", nn.ReLU()),
            ("layer_norm_1", nn.LayerNorm(self.filter_size)),
            ("dropout_1", nn.Dropout(self.dropout)),
            ("conv1d_2", Conv(self.filter_size, self.filter_size, kernel_size=self.kernel, padding=1)),
            ("relu_2", nn.ReLU()),
            ("layer_norm_2", nn.LayerNorm(self.filter_size)),
            ("dropout_2", nn.Dropout(self.dropout)),
        ]))
        self.linear_layer = nn.Linear(self.conv_output_size, 1)

    def forward(self, encoder_output, mask, *, lens=None):
        x, _ = to_act(encoder_output)
        B, T, _ = x.shape
        use_mask = mask is not None or lens is not None
        if lens is None:
            lens = lens_from_mask(mask, T, B, x.device)
        c = self.conv_layer
        p = self.dropout if self.training else 0.0
        return ops.VariancePredictorFn.apply(
            x, lens, c.conv1d_1.conv.weight, c.conv1d_1.conv.bias, c.layer_norm_1.weight, c.layer_norm_1.bias,
            c.conv1d_2.conv.weight, c.conv1d_2.conv.bias, c.layer_norm_2.weight, c.layer_norm_2.bias,
            self.linear_layer.weight, self.linear_layer.bias, p, use_mask)


class LengthRegulator(nn.Module):
    """Repeat phoneme i of each utterance max(int(d_i), 0) times, pad / crop to max_len
    (modules.py:163-196).  Integer cumsum + gather on the device: no `.item()` per phoneme."""

    def forward(self, x, duration, max_len):
        if not x.is_cuda:
            to_act(x)
        if max_len is None:  # reference pads to the longest expansion: needs the lengths on the host
            d = duration.detach()
            d = d.to(torch.int64) if not d.is_floating_point() else d.to(torch.float32).trunc().to(torch.int64)
            max_len = int(d.clamp_min(0).sum(1).max().item())
        max_len = int(max_len)
        cum, idx, mel_len = ops.lr_index(duration.detach(), max_len)
        if x.dtype not in (torch.bfloat16, torch.float32):
            x = x.float()
        out = ops.LengthRegulate.apply(x, cum, idx, max_len, max_len, None, None)
        return out, mel_len

    # aliases kept for callers of the reference's helper methods
    def LR(self, x, duration, max_len):
        return self.forward(x, duration, max_len)


class VarianceAdaptor(nn.Module):
    """duration / pitch / energy prediction, bucketised embeddings, length regulation
    (modules.py:18-160)."""

    def __init__(self, model_config):
        super().__init__()
        self.duration_predictor = VariancePredictor(model_config)
        self.length_regulator = LengthRegulator()
        self.pitch_predictor = VariancePredictor(model_config)
        self.energy_predictor = VariancePredictor(model_config)
        self.pitch_feature_level = model_config["pitch"]["feature"]
        self.energy_feature_level = model_config["energy"]["feature"]
        assert self.pitch_feature_level in ["phoneme_level", "frame_level"]
        assert self.energy_feature_level in ["phoneme_level", "frame_level"]
        ve = model_config["variance_embedding"]
        n_bins = ve["n_bins"]
        assert ve["pitch_quantization"] in ["linear", "log"] and ve["energy_quantization"] in ["linear", "log"]
        (p_min, p_max, p_mean, p_std, e_min, e_max, e_mean, e_std) = Define.ALLSTATS["global"]
        if model_config["pitch"]["normalization"]:
            p_min, p_max = (p_min - p_mean) / p_std, (p_max - p_mean) / p_std
        if model_config["energy"]["normalization"]:
            e_min, e_max = (e_min - e_mean) / e_std, (e_max - e_mean) / e_std

        def edges(lo, hi, kind):
            if kind == "log":
                return torch.exp(torch.linspace(np.log(lo), np.log(hi), n_bins - 1))
            return torch.linspace(lo, hi, n_bins - 1)

        self.pitch_bins = nn.Parameter(edges(p_min, p_max, ve["pitch_quantization"]), requires_grad=False)
        self.energy_bins = nn.Parameter(edges(e_min, e_max, ve["energy_quantization"]), requires_grad=False)
        d = model_config["transformer"]["encoder_hidden"]
        self.pitch_embedding = nn.Embedding(n_bins, d)
        self.energy_embedding = nn.Embedding(n_bins, d)

    # -- (prediction, x + embedding): the reference returns the embedding; adding it is fused here --------
    def _variance(self, predictor, table, bins, x, target, mask, lens, control, branch=None):
        if target is not None and branch is not None:
            # teacher forcing: the prediction only feeds the loss -- its launches go on a branch stream, the
            # embedding of the TARGET is added on the current one (ops.branch)
            with ops.branch(branch):
                prediction = predictor(x, mask, lens=lens)
            return prediction, ops.BucketEmbedAdd.apply(x, target, bins, table.weight)
        prediction = predictor(x, mask, lens=lens)
        if target is None:
            prediction = prediction * control
            target = prediction.detach()
        return prediction, ops.BucketEmbedAdd.apply(x, target, bins, table.weight)

    def get_pitch_embedding(self, x, target, mask, control):
        xb, dt = to_act(x)
        pred, y = self._variance(self.pitch_predictor, self.pitch_embedding, self.pitch_bins, xb, target, mask,
                                 None, control)
        return pred, from_act(y - xb, dt)

    def get_energy_embedding(self, x, target, mask, control):
        xb, dt = to_act(x)
        pred, y = self._variance(self.energy_predictor, self.energy_embedding, self.energy_bins, xb, target, mask,
                                 None, control)
        return pred, from_act(y - xb, dt)

    def forward(self, x, src_mask, mel_mask=None, max_len=None, pitch_target=None, energy_target=None,
                duration_target=None, p_control=1.0, e_control=1.0, d_control=1.0, *, src_lens=None,
                _fused=None):
        """Same 7-tuple as the reference.  `_fused=(out_len, spk_rows, pos_table)` is used by FastSpeech2
        to fold `+ speaker`, decoder truncation and `+ position_enc` into the gather."""
        xb, dt = to_act(x)
        B, Ts, _ = xb.shape
        if src_lens is None:
            src_lens = lens_from_mask(src_mask, Ts, B, xb.device)
        if duration_target is not None:  # the durations come from the batch: log_d only feeds the loss
            with ops.branch(0):
                log_d = self.duration_predictor(xb, src_mask, lens=src_lens)
        else:
            log_d = self.duration_predictor(xb, src_mask, lens=src_lens)
        pitch_pred = energy_pred = None
        if self.pitch_feature_level == "phoneme_level":
            pitch_pred, xb = self._variance(self.pitch_predictor, self.pitch_embedding, self.pitch_bins, xb,
                                            pitch_target, src_mask, src_lens, p_control, branch=1)
        if self.energy_feature_level == "phoneme_level":
            energy_pred, xb = self._variance(self.energy_predictor, self.energy_embedding, self.energy_bins,
                                             xb, energy_target, src_mask, src_lens, e_control, branch=2)
        if duration_target is not None:
            durations = duration_target
            duration_rounded = duration_target
        else:
            duration_rounded = torch.clamp(torch.round(torch.exp(log_d.detach()) - 1) * d_control, min=0)
            durations = duration_rounded
        if max_len is None:
            d = durations.detach()
            d = d.trunc().to(torch.int64) if d.is_floating_point() else d
            max_len = int(d.clamp_min(0).sum(1).max().item())  # inference: one host sync, as the reference
        max_len = int(max_len)
        cum, idx, mel_len = ops.lr_index(durations.detach(), max_len)
        if duration_target is None:
            mel_mask = get_mask_from_lengths(mel_len, max_len)
        frame_level = "frame_level" in (self.pitch_feature_level, self.energy_feature_level)
        if _fused is not None and not frame_level:
            out_len, spk, pe = _fused
            xo = ops.LengthRegulate.apply(xb, cum, idx, max_len, out_len, spk, pe)
        else:
            xo = ops.LengthRegulate.apply(xb, cum, idx, max_len, max_len, None, None)
            mel_lens_c = torch.clamp(mel_len, max=max_len)
            if self.pitch_feature_level == "frame_level":
                pitch_pred, xo = self._variance(self.pitch_predictor, self.pitch_embedding, self.pitch_bins, xo,
                                                pitch_target, mel_mask, mel_lens_c, p_control)
            if self.energy_feature_level == "frame_level":
                energy_pred, xo = self._variance(self.energy_predictor, self.energy_embedding,
                                                 self.energy_bins, xo, energy_target, mel_mask, mel_lens_c,
                                                 e_control)
        if _fused is None:
            xo = from_act(xo, dt)
            ops.join_branches()  # stand-alone use: every returned tensor is ready on the caller's stream
        return xo, pitch_pred, energy_pred, log_d, duration_rounded, mel_len, mel_mask
