"""Encoder / Encoder2 / Decoder stacks (reference: transformer/Models.py:10-237)."""
import numpy as np
import torch
import torch.nn as nn

from .. import ops
from .._act import from_act, lens_from_mask, to_act
from . import Constants
from .Layers import FFTBlock


def get_sinusoid_encoding_table(n_position, d_hid, padding_idx=None):
    """PE[p, j] = sin / cos (p / 10000^(2*(j//2)/d_hid)) computed in float64 then cast to fp32, exactly
    the numbers of the reference's python-list version (Models.py:10-30), vectorised."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    table = pos / np.power(10000, 2 * (j // 2) / d_hid)
    table[:, 0::2] = np.sin(table[:, 0::2])
    table[:, 1::2] = np.cos(table[:, 1::2])
    if padding_idx is not None:
        table[padding_idx] = 0.0
    return torch.FloatTensor(table)


class _FFTStack(nn.Module):
    """Shared body: frozen sinusoid table + N FFT blocks."""

    def _build(self, config, prefix):
        tcfg = config["transformer"]
        d_model = tcfg[prefix + "_hidden"]
        n_head = tcfg[prefix + "_head"]
        self.max_seq_len = config["max_seq_len"]
        self.d_model = d_model
        self.position_enc = nn.Parameter(
            get_sinusoid_encoding_table(self.max_seq_len + 1, d_model).unsqueeze(0), requires_grad=False)
        self.layer_stack = nn.ModuleList([
            FFTBlock(d_model, n_head, d_model // n_head, d_model // n_head, tcfg["conv_filter_size"],
                     tcfg["conv_kernel_size"], dropout=tcfg[prefix + "_dropout"])
            for _ in range(tcfg[prefix + "_layer"])])

    def _table_for(self, T, device):
        if not self.training and T > self.max_seq_len:  # eval long-sequence branch (Models.py:148-153)
            return get_sinusoid_encoding_table(T, self.d_model).to(device), T
        return self.position_enc, T

    def _run_layers(self, x, lens):
        if not self.layer_stack:
            return x
        # one longest-first work order for the attention kernels of all layers (same lengths in every layer)
        with ops.attn_schedule_scope(lens, x.shape[1], self.layer_stack[0].slf_attn.n_head):
            for layer in self.layer_stack:
                x, _ = layer(x, lens=lens)
        return x


class Encoder2(_FFTStack):
    """Encoder without the embedding layer: emb + PE -> 4 FFT blocks (Models.py:103-166)."""

    def __init__(self, config):
        super().__init__()
        self._build(config, "encoder")

    def forward(self, emb_src_seq, mask, return_attns=False, *, lens=None, out_dtype=None):
        """out_dtype (keyword, not in the reference signature): FastSpeech2.forward asks for the bf16 activations
        directly instead of a caller-dtype (fp32) copy that it would cast back."""
        B, T = emb_src_seq.shape[0], emb_src_seq.shape[1]
        dt = out_dtype if out_dtype is not None else emb_src_seq.dtype
        if not emb_src_seq.is_cuda:
            to_act(emb_src_seq)  # raises: no CPU path
        if lens is None:
            lens = lens_from_mask(mask, T, B, emb_src_seq.device)
        table, _ = self._table_for(T, emb_src_seq.device)
        x = ops.PosEncAdd.apply(emb_src_seq, table, T)
        return from_act(self._run_layers(x, lens), dt if dt != torch.float64 else torch.float32)


class Encoder(_FFTStack):
    """Encoder with its own nn.Embedding (Models.py:33-100); same stack behind an embedding gather."""

    def __init__(self, config, n_src_vocab=None):
        super().__init__()
        if n_src_vocab is None:
            try:
                from text.symbols import symbols  # the reference's symbol table, when it is importable
                n_src_vocab = len(symbols) + 1
            except Exception:
                n_src_vocab = config.get("n_src_vocab", 361)
        self.src_word_emb = nn.Embedding(n_src_vocab, config["transformer"]["encoder_hidden"],
                                         padding_idx=Constants.PAD)
        self._build(config, "encoder")

    def forward(self, src_seq, mask, return_attns=False, *, lens=None, out_dtype=None):
        B, T = src_seq.shape
        if lens is None:
            lens = lens_from_mask(mask, T, B, src_seq.device)
        emb = ops.EmbeddingFn.apply(src_seq, self.src_word_emb.weight, Constants.PAD)
        table, _ = self._table_for(T, src_seq.device)
        x = ops.PosEncAdd.apply(emb, table, T)
        return from_act(self._run_layers(x, lens), out_dtype if out_dtype is not None else torch.float32)


class Decoder(_FFTStack):
    """(truncate to max_seq_len in training) + PE -> 6 FFT blocks; returns (out, mask) (Models.py:169-237)."""

    def __init__(self, config):
        super().__init__()
        self._build(config, "decoder")

    def out_len(self, T):
        return T if (not self.training and T > self.max_seq_len) else min(T, self.max_seq_len)

    def forward(self, enc_seq, mask, return_attns=False, *, lens=None):
        B, T = enc_seq.shape[0], enc_seq.shape[1]
        dt = enc_seq.dtype
        if not enc_seq.is_cuda:
            to_act(enc_seq)
        t_out = self.out_len(T)
        if lens is None:
            lens = lens_from_mask(mask, T, B, enc_seq.device)
        table, _ = self._table_for(T, enc_seq.device)
        x = ops.PosEncAdd.apply(enc_seq, table, t_out)
        out = self.forward_prepared(x, lens)
        return from_act(out, dt), (mask[:, :t_out] if mask is not None else None)

    def forward_prepared(self, x, lens):
        """x already holds `enc_seq[:, :t_out] + PE` in bf16 (the fused LengthRegulator epilogue)."""
        return self._run_layers(x, lens)
