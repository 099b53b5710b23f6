"""MultilingualEmbedding: per-language phoneme tables, looked up in the concatenation of all tables
(reference: lightning/systems/language/embeddings.py:8-31).  The gather and its scatter-add backward
are CUDA kernels (ops.EmbeddingFn); the concatenation keeps the per-language ParameterDict (and so the
state_dict keys `tables.table-<id>`) of the reference."""
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
            table = torch.cat([p for p in self.tables.values()], dim=0)
        else:
            table = self.tables[f"table-{symbol_id}"]
        return ops.EmbeddingFn.apply(x, table, self.padding_idx)
