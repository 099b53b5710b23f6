"""FFT-block sub-layers with the reference's constructors, forward signatures and state_dict keys
(reference: transformer/SubLayers.py:8-93).  The nn.Linear / nn.Conv1d / nn.LayerNorm children are
parameter containers only: the math runs in ops.MHASublayer / ops.FFNSublayer (sm_100a kernels).
"""
import torch.nn as nn

from .. import ops
from .._act import from_act, lens_from_mask, to_act
from .Modules import ScaledDotProductAttention


class MultiHeadAttention(nn.Module):
    """LN(dropout(fc(softmax(QK^T/sqrt(dk) + keymask) V)) + q)   (SubLayers.py:29-57)."""

    def __init__(self, n_head, d_model, d_k, d_v, dropout=0.1):
        super().__init__()
        if d_k != d_v:
            raise NotImplementedError("d_k == d_v is assumed (as in every reference config)")
        self.n_head, self.d_k, self.d_v = n_head, d_k, d_v
        self.w_qs = nn.Linear(d_model, n_head * d_k)
        self.w_ks = nn.Linear(d_model, n_head * d_k)
        self.w_vs = nn.Linear(d_model, n_head * d_v)
        self.attention = ScaledDotProductAttention(temperature=d_k ** 0.5)
        self.layer_norm = nn.LayerNorm(d_model)
        self.fc = nn.Linear(n_head * d_v, d_model)
        self.dropout = nn.Dropout(dropout)

    def forward(self, q, k, v, mask=None, *, lens=None, zero_pad=False):
        if not (q is k and k is v):
            raise NotImplementedError("only self-attention (q is k is v) is on the FastSpeech2 path")
        x, dt = to_act(q)
        B, T, _ = x.shape
        if lens is None:
            lens = lens_from_mask(mask, T, B, x.device)
        p = self.dropout.p if self.training else 0.0
        y = ops.MHASublayer.apply(x, lens, self.w_qs.weight, self.w_qs.bias, self.w_ks.weight,
                                  self.w_ks.bias, self.w_vs.weight, self.w_vs.bias, self.fc.weight,
                                  self.fc.bias, self.layer_norm.weight, self.layer_norm.bias, self.n_head,
                                  p, zero_pad)
        # attention probabilities are never consumed on the training path (transformer/Models.py:163-166
        # drops them); they are not materialised in the caller's layout.
        return from_act(y, dt), None


class PositionwiseFeedForward(nn.Module):
    """LN(dropout(conv_k2(relu(conv_k1(x)))) + x)   (SubLayers.py:60-93), channels-last implicit GEMMs."""

    def __init__(self, d_in, d_hid, kernel_size, dropout=0.1):
        super().__init__()
        self.w_1 = nn.Conv1d(d_in, d_hid, kernel_size=kernel_size[0], padding=(kernel_size[0] - 1) // 2)
        self.w_2 = nn.Conv1d(d_hid, d_in, kernel_size=kernel_size[1], padding=(kernel_size[1] - 1) // 2)
        self.layer_norm = nn.LayerNorm(d_in)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x, *, lens=None, zero_pad=False):
        xb, dt = to_act(x)
        B, T, _ = xb.shape
        if zero_pad and lens is None:
            raise ValueError("zero_pad needs lens")
        if lens is None:
            lens = lens_from_mask(None, T, B, xb.device)
        p = self.dropout.p if self.training else 0.0
        y = ops.FFNSublayer.apply(xb, lens, self.w_1.weight, self.w_1.bias, self.w_2.weight, self.w_2.bias,
                                  self.layer_norm.weight, self.layer_norm.bias, p, zero_pad)
        return from_act(y, dt)
