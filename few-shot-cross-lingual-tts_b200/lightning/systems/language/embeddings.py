"""MultilingualEmbedding: per-language phoneme tables, looked up in the (virtual) concatenation of all tables
(reference: lightning/systems/language/embeddings.py:8-31).  The gather and its scatter-add backward are CUDA
kernels (ops.MultiEmbeddingFn / ops.EmbeddingFn) that resolve an id against the cumulative row counts, so the
concatenated table is never built; the per-language ParameterDict (state_dict keys `tables.table-<id>`) is the
reference's."""
from math import sqrt
from typing import Optional

import torch
import torch.nn as nn

from .... import ops


class MultilingualEmbedding(nn.Module):
    def __init__(self, id2symbols, dim: int, padding_idx: int = 0):
        super().__init__()
        self.id2symbols = id2symbols
        self.dim = dim
        self.padding_idx = padding_idx
        self.tables = nn.ParameterDict()
        for symbol_id, v in id2symbols.items():
            if len(v) > 0:
                val = sqrt(3.0) * sqrt(2.0 / (len(v) + dim))  # uniform bound for a xavier-like std
                w_init = torch.empty(len(v), dim).uniform_(-val, val)
                w_init[padding_idx].fill_(0)
                self.tables[f"table-{symbol_id}"] = nn.Parameter(w_init)

    def forward(self, x, symbol_id: Optional[str] = None):
        """x: int64 ids [B, T] -> fp32 [B, T, dim] (bf16 internally; the model re-casts on entry)."""
        if symbol_id is None:
            # ids index the concatenation of all tables; the reference's per-call torch.cat is replaced by a lookup
            # against the cumulative row counts inside the kernel (gradients go straight to each table)
            return ops.MultiEmbeddingFn.apply(x, self.padding_idx, *self.tables.values())
        return ops.EmbeddingFn.apply(x, self.tables[f"table-{symbol_id}"], self.padding_idx)


class SoftMultiAttCodebook2(nn.Module):
    """Codebook attention that turns phoneme-level SSL queries into phoneme embeddings (reference:
    lightning/systems/language/embeddings.py:77-142): softmax(weight_raw)-weighted sum over the upstream layers
    (NaN -> 0), `q_linear`, then `num_heads`-head attention (temperature sqrt(head dim)) of every query over the
    `att_banks` keys with the `emb_banks` values.  Same constructor, forward signature and state_dict keys
    (`emb_banks`, `att_banks`, `weight_raw`, `q_linear.{weight,bias}`); the reference reads the upstream geometry from
    its global `Define` module, here the same names are read from this package's `Define` unless passed explicitly.
    Unlike the reference, `ref` is not modified in place (NaNs are zeroed on the fly)."""

    def __init__(self, codebook_size, embed_dim, num_heads, upstream_dim=None, upstream=None, layer_idx=None,
                 n_layers=25):
        super().__init__()
        from .... import Define

        self.codebook_size, self.d_word_vec, self.num_heads = codebook_size, embed_dim, num_heads
        assert embed_dim % num_heads == 0
        upstream = Define.UPSTREAM if upstream is None else upstream
        layer_idx = Define.LAYER_IDX if layer_idx is None else layer_idx
        upstream_dim = Define.UPSTREAM_DIM if upstream_dim is None else upstream_dim
        self.emb_banks = nn.Parameter(torch.randn(codebook_size, embed_dim))
        self.att_banks = nn.Parameter(torch.randn(codebook_size, embed_dim))
        self.temperature = (embed_dim // num_heads) ** 0.5
        self.layered = upstream != "mel" and upstream is not None
        if self.layered and layer_idx is not None:
            w = torch.ones(1, n_layers, 1) * float("-inf")  # one-hot layer selection (embeddings.py:96-100)
            w[0][layer_idx][0] = 10.0
            self.weight_raw = nn.Parameter(w, requires_grad=False)
        self.q_linear = nn.Linear(upstream_dim, embed_dim)

    def forward(self, ref, need_weights=False):
        """ref: [B, L, n_layer, dim] (or [B, L, dim] for the "mel" upstream) -> ([B, L, embed_dim] fp32, None)."""
        if need_weights:
            raise NotImplementedError("attention weights are not materialised on the CUDA path")
        B, Lq = ref.shape[0], ref.shape[1]
        if self.layered:
            flat = ref.reshape(B * Lq, ref.shape[2], ref.shape[3])
            w_raw = self.weight_raw
        else:
            flat, w_raw = ref.reshape(B * Lq, ref.shape[-1]), None
        out = ops.CodebookAttnFn.apply(flat, w_raw, self.q_linear.weight, self.q_linear.bias, self.att_banks,
                                       self.emb_banks, self.num_heads, self.temperature)
        return out.view(B, Lq, self.d_word_vec), None
