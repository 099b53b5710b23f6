"""Host-side descriptors for the tcgen05 GEMM engine (fs2_gemm_bf16, include/fs2b200.h).

Only shape bookkeeping lives here; all arithmetic happens in csrc/gemm_tc.cu.
"""
import os

import torch

from . import _cabi
from ._cabi import (EPI_ADD_AUX, EPI_NONE, EPI_RELU, EPI_RELU_BWD, GEMM_NORMAL, GEMM_WGRAD, Gemm,
                    Operand)

__all__ = ["operand", "run", "gemm", "wgrad", "EPI_NONE", "EPI_RELU", "EPI_RELU_BWD", "EPI_ADD_AUX"]


def _impl():
    # FS2_GEMM_IMPL=simt selects the CUDA-core cross-check kernel (debug only).
    return 1 if os.environ.get("FS2_GEMM_IMPL", "tc") == "simt" else 0


def _stream():
    return torch.cuda.current_stream().cuda_stream


_WS = {}


def _workspace():
    """Split-K scratch of the Conv1d kernel (fs2_gemm::workspace): one buffer per stream (launches that may run
    concurrently must not share it) and per graph capture; the kernel leaves its counters zero again."""
    st = torch.cuda.current_stream()
    ws, fresh = scratch(_WS, st, (), int(_cabi.lib().fs2_gemm_workspace_bytes()), torch.uint8)
    if fresh:
        ws[:4096].zero_()  # the arrival counters; the partial-sum slabs are written before they are read
    return ws


def scratch(cache, st, extra_key, numel, dtype):
    """Self-resetting kernel scratch: one buffer per (device, stream, CUDA-graph capture).  Inside a capture the
    buffer is new and the caller's zeroing becomes a node of THAT graph (it runs on every replay), so no graph relies
    on another one -- or on a kernel that was aborted -- having left the counters at zero; buffers of finished
    captures are dropped here (their memory stays with the graph that uses it).  -> (buffer, needs_zeroing)"""
    cap = int(_cabi.lib().fs2_stream_capture_id(st.cuda_stream))
    key = (st.device_index, st.cuda_stream, cap) + tuple(extra_key)
    ws = cache.get(key)
    if ws is not None:
        return ws, False
    if cap:
        for k in [k for k in cache if k[:2] == key[:2] and k[2] not in (0, cap)]:
            del cache[k]
    ws = cache[key] = torch.empty(int(numel), dtype=dtype, device=st.device)
    return ws, True


def operand(t, inner, rows, batches=1, ld=None, batch_stride=None, mn_major=False, inner_base=0,
            zdiv=1, zmod_stride=0):
    """Describe bf16 tensor `t` as [batches][rows][inner] with explicit element strides."""
    assert t.dtype == torch.bfloat16 and t.is_cuda, "GEMM operands are bf16 CUDA tensors"
    ld = int(ld if ld is not None else inner)
    batch_stride = int(batch_stride if batch_stride is not None else ld * rows)
    o = Operand()
    o.ptr = t.data_ptr()
    o.ld, o.batch_stride = ld, batch_stride
    o.inner, o.rows, o.batches = int(inner), int(rows), int(batches)
    o.mn_major = 1 if mn_major else 0
    o.inner_base, o.zdiv, o.zmod_stride = int(inner_base), int(zdiv), int(zmod_stride)
    return o


def run(g, impl=None):
    rc = _cabi.lib().fs2_gemm_bf16(g, _impl() if impl is None else impl, _stream())
    _cabi.check(rc, "fs2_gemm_bf16")


def gemm(a, b, d, M, N, K, Z=1, taps=1, tap_shift0=0, b_tap_kstride=0, ldd=None, bias=None,
         epilogue=EPI_NONE, aux=None, ld_aux=0, aux_batch_stride=0, alpha=1.0, d_zdiv=1,
         d_zdiv_stride=0, d_zmod_stride=0, impl=None, row_lens=None, lens_zdiv=1, tail_rows=0, relu_mask=None,
         ln=None):
    """D[z] = epi(alpha * sum_tap A[z][m+shift+tap] . B[z][n][tap*kstride+k] + bias).

    row_lens (int64 [Z / lens_zdiv]): rows m >= row_lens of D[z] are written as zero and row tiles that
    hold only such rows are skipped (padded frames of an utterance batch).  tail_rows: rows zeroed behind the
    last scheduled row tile (0 = all: every element of D is defined; n > 0: an n-row halo; < 0: none)."""
    g = Gemm()
    g.a, g.b = a, b
    g.mode = GEMM_NORMAL
    g.M, g.N, g.K, g.Z = int(M), int(N), int(K), int(Z)
    g.taps, g.tap_shift0, g.b_tap_kstride = int(taps), int(tap_shift0), int(b_tap_kstride)
    g.splits = 1
    g.epilogue = epilogue
    assert d.dtype in (torch.bfloat16, torch.float32)
    g.d_f32 = 1 if d.dtype == torch.float32 else 0
    g.d_atomic = 0
    g.d_zdiv, g.d_zdiv_stride, g.d_zmod_stride = int(d_zdiv), int(d_zdiv_stride), int(d_zmod_stride)
    g.alpha = float(alpha)
    g.d = d.data_ptr()
    g.ldd = int(ldd if ldd is not None else N)
    g.d_col_stride, g.d_tap_stride = 1, 0
    g.bias = bias.data_ptr() if bias is not None else None
    if bias is not None:
        assert bias.dtype == torch.float32
    g.aux = aux.data_ptr() if aux is not None else None
    g.ld_aux, g.aux_batch_stride = int(ld_aux), int(aux_batch_stride)
    _set_lens(g, row_lens, lens_zdiv)
    g.tail_zero_rows = int(tail_rows)
    if relu_mask is not None:  # int64 [Z*M, N/64]: written by EPI_RELU, read by EPI_RELU_BWD instead of `aux`
        assert relu_mask.dtype == torch.int64 and relu_mask.is_contiguous() and relu_mask.numel() == Z * M * (N // 64)
        g.relu_mask = relu_mask.data_ptr()
    if ln is not None:  # fused dropout + residual + LayerNorm + pad-zero epilogue (fs2_gemm::ln_*)
        g.ln_gamma, g.ln_beta = ln["gamma"].data_ptr(), ln["beta"].data_ptr()
        res = ln["res"]
        assert res.dtype == torch.bfloat16 and res.is_contiguous()
        g.ln_res, g.ld_res, g.res_batch_stride = res.data_ptr(), int(N), int(M * N)
        g.ln_p_drop, g.ln_seed = float(ln["p"]), int(ln["salt"])
        g.ln_seed_dev = ln["seed_dev"].data_ptr() if ln.get("seed_dev") is not None else None
        g.ln_v, g.ln_mean, g.ln_rstd = ln["v"].data_ptr(), ln["mean"].data_ptr(), ln["rstd"].data_ptr()
        g.ln_keep = ln["keep"].data_ptr() if ln.get("keep") is not None else None
    if g.taps > 1 and not g.d_f32:  # last partial wave of the Conv1d schedule: split over the reduction
        ws = _workspace()
        g.workspace, g.workspace_bytes = ws.data_ptr(), ws.numel()
    run(g, impl)


def _set_lens(g, row_lens, lens_zdiv):
    if row_lens is not None:
        assert row_lens.dtype == torch.int64 and row_lens.is_cuda and row_lens.is_contiguous()
        g.row_lens = row_lens.data_ptr()
        g.lens_zdiv = int(lens_zdiv)


def wgrad(a, b, d, M, N, taps=1, tap_shift0=0, ldd=None, d_col_stride=1, d_tap_stride=0, splits=1,
          accumulate=True, impl=None, segments=None, row_lens=None, lens_zdiv=1, a_colsum=None, a_colsum_seg=None):
    """D[m][tap][n] (+)= sum_{z,r} A[z][r][m] * B[z][r+shift+tap][n]; D is fp32.

    row_lens: promise that rows r >= row_lens[z] of A[z] are zero -> their 64-row blocks are skipped.
    a_colsum: f32 [M] that receives += sum_{z,r} A[z][r][m] (the bias gradient; fs2_gemm::a_colsum);
    a_colsum_seg: with `segments`, one f32 [rows_per_segment] tensor (or None) per segment instead."""
    if segments is not None:  # (rows_per_segment, [tensor, ...]): row block i of D lives in tensors[i]
        d = segments[1][0]
    assert d.dtype == torch.float32
    g = Gemm()
    g.a, g.b = a, b
    g.mode = GEMM_WGRAD
    g.M, g.N, g.K, g.Z = int(M), int(N), 0, 1
    g.taps, g.tap_shift0, g.b_tap_kstride = int(taps), int(tap_shift0), 0
    g.splits = int(splits)
    g.epilogue = EPI_NONE
    g.d_f32 = 1
    g.d_atomic = 1 if (accumulate or splits > 1) else 0
    g.d_zdiv, g.d_zdiv_stride, g.d_zmod_stride = 1, 0, 0
    g.alpha = 1.0
    g.d = d.data_ptr()
    g.ldd = int(ldd if ldd is not None else N * taps)
    g.d_col_stride, g.d_tap_stride = int(d_col_stride), int(d_tap_stride)
    g.bias = None
    g.aux = None
    g.ld_aux = g.aux_batch_stride = 0
    if segments is not None:
        g.d_seg_rows = int(segments[0])
        for i, t in enumerate(segments[1]):
            assert t.dtype == torch.float32 and t.is_contiguous()
            g.d_seg[i] = t.data_ptr()
    _set_lens(g, row_lens, lens_zdiv)
    if a_colsum is not None:
        assert a_colsum.dtype == torch.float32 and a_colsum.is_contiguous() and a_colsum.numel() == M
        g.a_colsum = a_colsum.data_ptr()
    if a_colsum_seg is not None:
        assert segments is not None and len(a_colsum_seg) == len(segments[1])
        for i, t in enumerate(a_colsum_seg):
            if t is not None:
                assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == segments[0]
                g.a_colsum_seg[i] = t.data_ptr()
    run(g, impl)
