"""Fused attention kernels (tcgen05, S/P never in HBM) vs a plain PyTorch fp32 reference."""
import math

import pytest
import torch

from fs2b200 import sub
from tests.util_parity import rel_err

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16


def reference(qkv, lens, H, dk):
    B, T, _ = qkv.shape
    HD = H * dk
    q, k, v = (qkv[..., i * HD:(i + 1) * HD].float().view(B, T, H, dk).permute(0, 2, 1, 3) for i in range(3))
    s = (q @ k.transpose(-1, -2)) / math.sqrt(dk)  # [B,H,T,T]
    kmask = torch.arange(T, device=qkv.device)[None, :] >= lens[:, None]
    s = s.masked_fill(kmask[:, None, None, :], float("-inf"))
    p = torch.softmax(s, -1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B, T, HD)
    lse2 = torch.logsumexp(s, -1) * math.log2(math.e)  # [B,H,T]
    return o, lse2, p


@pytest.mark.parametrize("B,T,lens", [(2, 128, [128, 77]), (3, 200, [200, 1, 130]), (2, 70, [70, 33]),
                                      (1, 1000, [913]), (4, 384, [384, 383, 129, 128]),
                                      (3, 512, [512, 100, 0]), (2, 1000, [300, 1000])])
def test_attn_fwd(B, T, lens):
    torch.manual_seed(T)
    H, dk = 2, 128
    ops = sub("ops")
    qkv = (torch.randn(B, T, 3 * H * dk, device="cuda") * 1.5).to(BF16)
    lens_t = torch.tensor(lens, device="cuda")
    out, lse2 = ops.attn_fwd(qkv, lens_t, H, dk)
    ref, ref_lse2, _ = reference(qkv, lens_t, H, dk)
    qvalid = torch.arange(T, device="cuda")[None, :] < lens_t[:, None]
    assert torch.isfinite(out.float()).all()
    assert rel_err(out[qvalid], ref[qvalid]) < 1e-2, rel_err(out[qvalid], ref[qvalid])
    assert out[~qvalid].abs().sum() == 0  # padded query rows produce exact zeros
    l2 = lse2.view(B, H, T)
    m = qvalid[:, None, :].expand(-1, H, -1)
    assert torch.allclose(l2[m], ref_lse2[m], atol=2e-3, rtol=1e-4)
    assert torch.isinf(l2[~m]).all()


@pytest.mark.parametrize("B,T,lens", [(2, 128, [128, 77]), (3, 200, [200, 1, 130]), (2, 70, [70, 33]),
                                      (1, 1000, [913]), (2, 333, [333, 64]), (3, 512, [512, 100, 0]),
                                      (2, 1000, [300, 1000])])
def test_attn_bwd(B, T, lens):
    torch.manual_seed(T + 1)
    H, dk = 2, 128
    HD = H * dk
    ops = sub("ops")
    qkv = (torch.randn(B, T, 3 * HD, device="cuda")).to(BF16)
    lens_t = torch.tensor(lens, device="cuda")
    qvalid = torch.arange(T, device="cuda")[None, :] < lens_t[:, None]
    d_out = torch.randn(B, T, HD, device="cuda").to(BF16) * qvalid[..., None]  # padded rows carry no gradient
    out, lse2 = ops.attn_fwd(qkv, lens_t, H, dk)
    dbq, dbv = torch.full((HD,), 0.5, device="cuda"), torch.full((HD,), -0.25, device="cuda")
    dqkv = ops.attn_bwd(qkv, out, d_out, lse2, lens_t, H, dk, dbias_q=dbq, dbias_v=dbv)
    # Q / V bias gradients: column sums of the stored (bf16) dQ / dV, accumulated on top of the buffers' contents
    assert torch.allclose(dbq, dqkv[..., :HD].float().sum((0, 1)) + 0.5, rtol=2e-3, atol=2e-3)
    assert torch.allclose(dbv, dqkv[..., 2 * HD:].float().sum((0, 1)) - 0.25, rtol=2e-3, atol=2e-3)
    qr = qkv.float().requires_grad_()
    ref, _, _ = reference(qr, lens_t, H, dk)
    (ref * qvalid[..., None]).backward(d_out.float())
    g = qr.grad
    if min(lens) == 0:  # torch's softmax over an all-masked row is NaN; the gradient of an empty utterance is 0
        g = torch.where(lens_t[:, None, None] > 0, g, torch.zeros_like(g))
    assert torch.isfinite(dqkv.float()).all()
    for name, sl in (("dq", slice(0, HD)), ("dk", slice(HD, 2 * HD)), ("dv", slice(2 * HD, 3 * HD))):
        e = rel_err(dqkv[..., sl], g[..., sl])
        assert e < 2e-2, (name, e)
    # keys / queries beyond the length get exactly zero gradient
    assert dqkv[~qvalid].abs().sum() == 0


@pytest.mark.parametrize("B,T,lens", [(5, 700, [700, 13, 400, 0, 129]), (3, 128, [128, 128, 1])])
def test_attention_work_order_is_a_permutation_and_changes_no_bit(B, T, lens):
    """fs2_attn_schedule: longest utterance first, padded-only tiles last, every (z, tile) exactly once; the kernels
    produce bit-identical results in scheduled and natural order (same CTAs, different launch order)."""
    torch.manual_seed(5)
    H, dk = 2, 128
    ops = sub("ops")
    lens_t = torch.tensor(lens, device="cuda")
    sched = ops.attn_schedule(lens_t, T, H)
    nq = (T + 127) // 128
    e = sched.cpu().tolist()
    assert sorted(e) == sorted((z << 8) | t for z in range(B * H) for t in range(nq))
    work = [lens[(x >> 8) // H] if (x & 255) * 128 < lens[(x >> 8) // H] else 0 for x in e]
    assert work == sorted(work, reverse=True)
    qkv = torch.randn(B, T, 3 * H * dk, device="cuda").to(BF16)
    d_out = torch.randn(B, T, H * dk, device="cuda").to(BF16)
    o1, l1 = ops.attn_fwd(qkv, lens_t, H, dk)
    o2, l2 = ops.attn_fwd(qkv, lens_t, H, dk, sched)
    assert torch.equal(o1, o2) and torch.equal(l1, l2)
    g1 = ops.attn_bwd(qkv, o1, d_out, l1, lens_t, H, dk)
    g2 = ops.attn_bwd(qkv, o1, d_out, l1, lens_t, H, dk, sched)
    assert torch.equal(g1, g2)


def test_fused_and_unfused_sublayer_agree(monkeypatch):
    """The fused attention path against the GEMM -> softmax -> GEMM composition, whole sub-layer fwd+bwd."""
    S = sub("transformer.SubLayers")
    torch.manual_seed(3)
    mha = S.MultiHeadAttention(2, 256, 128, 128, dropout=0.0).cuda().train()
    x = torch.randn(3, 150, 256, device="cuda")
    lens = torch.tensor([150, 40, 97], device="cuda")
    w = torch.randn(3, 150, 256, device="cuda")
    res = {}
    for mode in ("fused", "unfused"):
        monkeypatch.setenv("FS2_ATTN", mode)
        xi = x.clone().requires_grad_()
        mha.zero_grad(set_to_none=True)
        y, _ = mha(xi, xi, xi, lens=lens, zero_pad=True)
        (y * w).sum().backward()
        res[mode] = (y.detach(), xi.grad, mha.w_ks.weight.grad.clone(), mha.fc.bias.grad.clone())
    for a, b in zip(res["fused"], res["unfused"]):
        assert rel_err(a, b) < 2e-2
