"""FastSpeech2Loss / FastSpeech2ADALoss with the reference's call signatures
(reference: lightning/model/loss.py:5-140); the arithmetic is ops.FastSpeech2LossFn (CUDA)."""
import torch
import torch.nn as nn

from ... import ops


def _lens_from_mask(mask):
    """valid lengths of a bool [B, T] padding mask (True = padding, prefix-valid as get_mask_from_lengths builds it)"""
    return (mask.shape[1] - mask.sum(1)).to(torch.int64)


class FastSpeech2Loss(nn.Module):
    """masked L1(mel) + L1(postnet mel) + MSE(log-duration) + MSE(pitch) + MSE(energy), pitch / energy
    phoneme-level or frame-level (loss.py:47-60)."""

    def __init__(self, model_config, **kwargs):
        super().__init__()
        self.pitch_feature_level = model_config["pitch"]["feature"]
        self.energy_feature_level = model_config["energy"]["feature"]
        for lvl in (self.pitch_feature_level, self.energy_feature_level):
            assert lvl in ("phoneme_level", "frame_level"), lvl

    def forward(self, inputs, predictions):
        mel_targets, _, _, pitch_targets, energy_targets, duration_targets = inputs[6:12]
        (mel_pred, post_pred, pitch_pred, energy_pred, log_d_pred, _, src_masks, mel_masks, src_lens,
         mel_lens) = predictions
        # validity of a mel frame is the model's mel mask (loss.py:36-39: built from the INPUT mel_lens in teacher-forced
        # mode and cut to the decoder's output length), not the sum of the durations
        if mel_masks is not None:
            mel_valid = _lens_from_mask(mel_masks)
        else:
            mel_valid = torch.clamp(mel_lens, max=mel_pred.shape[1])
        return ops.FastSpeech2LossFn.apply(mel_pred, post_pred, pitch_pred, energy_pred, log_d_pred,
                                           mel_targets, pitch_targets, energy_targets, duration_targets,
                                           src_lens, mel_valid, self.pitch_feature_level == "frame_level",
                                           self.energy_feature_level == "frame_level")


class FastSpeech2ADALoss(nn.Module):
    """Mel-only variant (loss.py:104-140): total = L1(mel) + L1(postnet mel), each a mean over the valid mel elements
    of `mel_masks` (True = padding); the target is cut to the mask length."""

    def forward(self, inputs, predictions):
        mel_targets = inputs
        mel_pred, post_pred, mel_masks = predictions
        B = mel_masks.shape[0]
        mel_lens = _lens_from_mask(mel_masks)
        dev = mel_pred.device
        # the variance terms are fed one always-valid zero element each: they come out as exactly 0
        z = torch.zeros(B, 1, dtype=torch.float32, device=dev)
        zl = torch.zeros(B, 1, dtype=torch.int64, device=dev)
        out = ops.FastSpeech2LossFn.apply(mel_pred, post_pred, z, z, z, mel_targets, z, z, zl,
                                          torch.ones(B, dtype=torch.int64, device=dev), mel_lens)
        return out[1] + out[2], out[1], out[2]
