"""PhonemeQueryExtractor with the reference's constructor / forward signature (reference:
lightning/model/reduction.py:42-110): phoneme-level average of the SSL frames of every segment, then the class-wise
average over all segments of the same phoneme in the support set.  Both stages are one CUDA pass per utterance
(`fs2_segment_class_accum_f32`: the same duration-cumsum segmentation as the LengthRegulator backward)."""
import torch.nn as nn

from ... import ops


class PhonemeQueryExtractor(nn.Module):
    def __init__(self, mode: str = "average", two_stage: bool = True):
        super().__init__()
        if mode != "average":
            raise NotImplementedError('only mode="average" (every shipped algorithm config) runs on the CUDA path')
        self.two_stage = two_stage

    def forward(self, representations, avg_frames, n_symbols, phonemes):
        """-> fp32 [1, n_symbols, *dims]; classes that never occur are zero rows (reduction.py:104-107)."""
        return ops.phoneme_class_mean(representations, avg_frames, n_symbols, phonemes, self.two_stage).unsqueeze(0)
