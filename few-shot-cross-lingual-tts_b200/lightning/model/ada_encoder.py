"""ADAEncoder with the reference's constructor / forward signature and state_dict keys
(reference: lightning/model/ada_encoder.py:11-25): `embedding` = Linear(d_in -> encoder_hidden) on the mel
frames, then the headless FFT-block encoder.  Both run on the sm_100a kernels (ops.LinearFn, Encoder2)."""
import torch.nn as nn

from ... import ops
from ..._act import from_act, to_act
from ...transformer import Encoder2
from ..utils.tool import get_mask_from_lengths


class ADAEncoder(nn.Module):
    def __init__(self, d_in, config):
        super().__init__()
        encoder_dim = config["transformer"]["encoder_hidden"]
        if d_in % 8:
            raise NotImplementedError("the GEMM engine needs a multiple of 8 input features (mel: 80)")
        self.embedding = nn.Linear(d_in, encoder_dim)
        self.encoder = Encoder2(config)

    def forward(self, x, lengths, embed=True):
        dt = x.dtype
        if embed:
            xb, _ = to_act(x)
            x = ops.LinearFn.apply(xb, self.embedding.weight, self.embedding.bias)
        mask = get_mask_from_lengths(lengths).to(x.device)
        return from_act(self.encoder(x, mask), dt)
