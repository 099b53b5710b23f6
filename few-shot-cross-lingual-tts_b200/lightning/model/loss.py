"""FastSpeech2Loss / FastSpeech2ADALoss with the reference's call signatures
(reference: lightning/model/loss.py:5-140); the arithmetic is ops.FastSpeech2LossFn (CUDA)."""
import torch
import torch.nn as nn

from ... import ops


class FastSpeech2Loss(nn.Module):
    """masked L1(mel) + L1(postnet mel) + MSE(log-duration) + MSE(pitch) + MSE(energy)."""

    def __init__(self, model_config, **kwargs):
        super().__init__()
        self.pitch_feature_level = model_config["pitch"]["feature"]
        self.energy_feature_level = model_config["energy"]["feature"]
        if self.pitch_feature_level != "phoneme_level" or self.energy_feature_level != "phoneme_level":
            raise NotImplementedError("the fused loss kernel covers phoneme-level pitch / energy "
                                      "(every shipped model config, config/model/*.yaml)")

    def forward(self, inputs, predictions):
        mel_targets, _, _, pitch_targets, energy_targets, duration_targets = inputs[6:12]
        (mel_pred, post_pred, pitch_pred, energy_pred, log_d_pred, _, src_masks, mel_masks, src_lens,
         mel_lens) = predictions
        return ops.FastSpeech2LossFn.apply(mel_pred, post_pred, pitch_pred, energy_pred, log_d_pred,
                                           mel_targets, pitch_targets, energy_targets, duration_targets,
                                           src_lens, mel_lens)


class FastSpeech2ADALoss(nn.Module):
    """Mel-only variant (loss.py:104-140): total = L1(mel) + L1(postnet mel)."""

    def forward(self, inputs, predictions):
        mel_targets = inputs
        mel_pred, post_pred, mel_masks = predictions
        B, Tm = mel_masks.shape
        mel_lens = (Tm - mel_masks.sum(1)).to(torch.int64)
        z = torch.zeros(B, 1, dtype=torch.float32, device=mel_pred.device)
        zl = torch.zeros(B, 1, dtype=torch.int64, device=mel_pred.device)
        out = ops.FastSpeech2LossFn.apply(mel_pred, post_pred, z, z, z, mel_targets, z, z, zl,
                                          torch.ones(B, dtype=torch.int64, device=mel_pred.device), mel_lens)
        return out[1] + out[2], out[1], out[2]
