"""FFTBlock, ConvNorm and PostNet with the reference's signatures / state_dict keys
(reference: transformer/Layers.py:11-137)."""
import torch.nn as nn

from .. import ops
from .._act import from_act, lens_from_mask, to_act
from .SubLayers import MultiHeadAttention, PositionwiseFeedForward


class FFTBlock(nn.Module):
    """attention sub-layer -> zero pads -> conv-FFN sub-layer -> zero pads   (Layers.py:21-30).
    The two masked_fill passes are fused into the LayerNorm kernels (zero_pad)."""

    def __init__(self, d_model, n_head, d_k, d_v, d_inner, kernel_size, dropout=0.1):
        super().__init__()
        self.slf_attn = MultiHeadAttention(n_head, d_model, d_k, d_v, dropout=dropout)
        self.pos_ffn = PositionwiseFeedForward(d_model, d_inner, kernel_size, dropout=dropout)

    def forward(self, enc_input, mask=None, slf_attn_mask=None, *, lens=None):
        x, dt = to_act(enc_input)
        B, T, _ = x.shape
        if lens is None:
            lens = lens_from_mask(mask if mask is not None else slf_attn_mask, T, B, x.device)
        y, attn = self.slf_attn(x, x, x, mask=slf_attn_mask, lens=lens, zero_pad=True)
        y = self.pos_ffn(y, lens=lens, zero_pad=True)
        return from_act(y, dt), attn


class ConvNorm(nn.Module):
    """Parameter container with the reference's key layout (`<name>.conv.weight`, Layers.py:33-64)."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=None, dilation=1,
                 bias=True, w_init_gain="linear"):
        super().__init__()
        if padding is None:
            assert kernel_size % 2 == 1
            padding = int(dilation * (kernel_size - 1) / 2)
        if stride != 1 or dilation != 1 or padding != (kernel_size - 1) // 2:
            raise NotImplementedError("the sm_100a Conv1d path covers stride 1, dilation 1, 'same' padding")
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                              padding=padding, dilation=dilation, bias=bias)

    def forward(self, signal):
        """[B, C, T] in / out like the reference; runs the channels-last implicit GEMM (inference helper)."""
        x, dt = to_act(signal.transpose(1, 2).contiguous())
        y = ops.conv_fwd(x, ops.pack_conv(self.conv.weight), self.conv.bias.detach())
        return from_act(y, dt).transpose(1, 2)


class PostNet(nn.Module):
    """Five Conv1d(k=5) + BatchNorm1d blocks, tanh on the first four, dropout 0.5 on all
    (Layers.py:67-137).  forward returns the PostNet output WITHOUT the residual, like the reference;
    `forward_residual` is the fused `postnet(x) + x` used by FastSpeech2."""

    def __init__(self, n_mel_channels=80, postnet_embedding_dim=512, postnet_kernel_size=5,
                 postnet_n_convolutions=5):
        super().__init__()
        self.convolutions = nn.ModuleList()
        dims = [n_mel_channels] + [postnet_embedding_dim] * (postnet_n_convolutions - 1) + [n_mel_channels]
        for i in range(postnet_n_convolutions):
            gain = "linear" if i == postnet_n_convolutions - 1 else "tanh"
            self.convolutions.append(nn.Sequential(
                ConvNorm(dims[i], dims[i + 1], kernel_size=postnet_kernel_size, stride=1,
                         padding=int((postnet_kernel_size - 1) / 2), dilation=1, w_init_gain=gain),
                nn.BatchNorm1d(dims[i + 1])))
        self.p_dropout = 0.5  # hard-coded in the reference (Layers.py:133-134)

    def _flat_tensors(self):
        ts = []
        for blk in self.convolutions:
            conv, bn = blk[0].conv, blk[1]
            ts += [conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                   bn.num_batches_tracked]
        return ts

    def forward_residual(self, x):
        """postnet(x) + x in fp32 (fastspeech2m.py:145); x: [B, T, n_mel] fp32."""
        return ops.PostNetFn.apply(x.float(), self.training, self.p_dropout, len(self.convolutions),
                                   *self._flat_tensors())

    def forward(self, x):
        xf = x.float()
        return (self.forward_residual(xf) - xf).to(x.dtype)
