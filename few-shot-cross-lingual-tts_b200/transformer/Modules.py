"""Scaled dot-product attention (reference: transformer/Modules.py:6-25).

In the B200 path the attention core is part of ops.MHASublayer (QK^T and PV are batched tcgen05 GEMMs,
the masked softmax is one kernel in between); this class keeps the reference's constructor and, when
called on its own, runs the same kernels on separate q / k / v tensors.
"""
import math

import torch
import torch.nn as nn

from .. import gemm as G
from .. import ops
from .._act import from_act, lens_from_mask, to_act


class ScaledDotProductAttention(nn.Module):
    def __init__(self, temperature):
        super().__init__()
        self.temperature = temperature
        self.softmax = nn.Softmax(dim=2)  # kept for attribute parity; the CUDA softmax is used

    def forward(self, q, k, v, mask=None):
        """q,k,v: [Z, T, d]; mask: bool [Z, Tq, Tk] (True = masked key). Forward-only helper."""
        if torch.is_grad_enabled() and any(t.requires_grad for t in (q, k, v)):
            raise NotImplementedError(
                "stand-alone ScaledDotProductAttention is inference-only; training goes through "
                "MultiHeadAttention (fused sub-layer with a hand-written backward)")
        (qb, dt), (kb, _), (vb, _) = to_act(q.contiguous()), to_act(k.contiguous()), to_act(v.contiguous())
        Z, Tq, d = qb.shape
        Tk = kb.shape[1]
        assert Tq == Tk, "self-attention shapes only"
        lens = lens_from_mask(mask, Tk, Z, qb.device)
        Tp = (Tk + 127) // 128 * 128
        S = torch.empty(Z, Tq, Tp, dtype=torch.float32, device=qb.device)
        G.gemm(G.operand(qb, d, Tq, Z), G.operand(kb, d, Tk, Z), S, Tq, Tk, d, Z=Z, ldd=Tp,
               alpha=1.0 / float(self.temperature), d_zdiv=1, d_zdiv_stride=Tq * Tp)
        P = torch.empty(Z, Tq, Tp, dtype=torch.bfloat16, device=qb.device)
        ops._ck(ops._L().fs2_softmax_fwd(S.data_ptr(), lens.data_ptr(), Z, 1, Tq, Tp, P.data_ptr(), ops._st()),
                "softmax_fwd")
        out = torch.empty(Z, Tq, d, dtype=torch.bfloat16, device=qb.device)
        G.gemm(G.operand(P, Tp, Tq, Z), G.operand(vb, d, Tk, Z, mn_major=True), out, Tq, d, Tk, Z=Z, ldd=d,
               d_zdiv=1, d_zdiv_stride=Tq * d)
        return from_act(out, dt), from_act(P[:, :, :Tk], dt)
