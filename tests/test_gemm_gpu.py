"""tcgen05 GEMM engine vs a plain PyTorch fp32 reference (and vs the CUDA-core cross-check kernel).

Every case feeds bf16 operands; the reference multiplies the same bf16 values in fp32.
Tolerances: fp32 outputs rel-err <= 2e-3 (accumulation order), bf16 outputs <= 1e-2 (rounding).
"""
import pytest
import torch

from fs2b200 import sub

pytestmark = pytest.mark.gpu

G = None


def setup_module(module):
    global G
    G = sub("gemm")


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).to(torch.bfloat16)


IMPLS = [0, 1]  # 0 = tcgen05, 1 = CUDA-core cross-check


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("M,N,K", [(300, 768, 256), (128, 256, 64), (1000, 256, 1024), (77, 80, 256),
                                   (4096, 1024, 256), (513, 200, 192)])
@pytest.mark.parametrize("relu", [False, True])
def test_linear(impl, M, N, K, relu):
    torch.manual_seed(M + N + K)
    x, w = rnd(M, K), rnd(N, K, scale=K ** -0.5)
    bias = torch.randn(N, device="cuda")
    for out_dtype, tol in ((torch.bfloat16, 1e-2), (torch.float32, 2e-3)):
        y = torch.full((M, N), float("nan"), device="cuda", dtype=out_dtype)
        G.gemm(G.operand(x, K, M), G.operand(w, K, N), y, M, N, K, bias=bias,
               epilogue=G.EPI_RELU if relu else G.EPI_NONE, impl=impl)
        ref = x.float() @ w.float().t() + bias
        if relu:
            ref = ref.relu()
        assert torch.isfinite(y.float()).all()
        assert rel_err(y, ref) < tol, (impl, out_dtype, rel_err(y, ref))


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("B,T,Cin,Cout,k", [(3, 200, 256, 1024, 9), (2, 77, 256, 256, 3),
                                            (2, 130, 512, 512, 5), (2, 130, 80, 512, 5),
                                            (2, 130, 512, 80, 5), (5, 40, 1024, 256, 1)])
def test_conv1d_channels_last(impl, B, T, Cin, Cout, k):
    torch.manual_seed(B * T + Cin + k)
    x = rnd(B, T, Cin)
    w = (torch.randn(Cout, Cin, k, device="cuda") * (Cin * k) ** -0.5).to(torch.bfloat16)
    bias = torch.randn(Cout, device="cuda")
    cpad = ((Cin + 63) // 64) * 64
    wp = torch.zeros(Cout, k, cpad, device="cuda", dtype=torch.bfloat16)
    wp[:, :, :Cin] = w.permute(0, 2, 1)
    y = torch.full((B, T, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    G.gemm(G.operand(x, Cin, T, B), G.operand(wp, k * cpad, Cout), y, T, Cout, Cin, Z=B, taps=k,
           tap_shift0=-((k - 1) // 2), b_tap_kstride=cpad, bias=bias, d_zdiv=1,
           d_zdiv_stride=T * Cout, impl=impl)
    ref = torch.nn.functional.conv1d(x.float().transpose(1, 2), w.float(), bias,
                                     padding=(k - 1) // 2).transpose(1, 2)
    assert torch.isfinite(y.float()).all()
    assert rel_err(y, ref) < 1e-2, rel_err(y, ref)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("B,T,H", [(2, 200, 2), (3, 77, 2), (1, 850, 2)])
def test_attention_bmms(impl, B, T, H):
    """QK^T from the fused [B,T,3*H*dk] buffer, then P.V with V as an MN-major operand."""
    dk = 128
    torch.manual_seed(T)
    C = 3 * H * dk
    qkv = rnd(B, T, C)
    Tp = ((T + 63) // 64) * 64
    s = torch.full((B * H, T, Tp), float("nan"), device="cuda")
    a = G.operand(qkv, C, T, B, inner_base=0, zdiv=H, zmod_stride=dk)
    b = G.operand(qkv, C, T, B, inner_base=H * dk, zdiv=H, zmod_stride=dk)
    G.gemm(a, b, s, T, T, dk, Z=B * H, ldd=Tp, alpha=dk ** -0.5, d_zdiv=1, d_zdiv_stride=T * Tp,
           impl=impl)
    q = qkv[..., : H * dk].float().view(B, T, H, dk).permute(0, 2, 1, 3)
    kk = qkv[..., H * dk: 2 * H * dk].float().view(B, T, H, dk).permute(0, 2, 1, 3)
    v = qkv[..., 2 * H * dk:].float().view(B, T, H, dk).permute(0, 2, 1, 3)
    sref = (q @ kk.transpose(-1, -2)) * dk ** -0.5  # [B,H,T,T]; z = b*H + h
    assert rel_err(s[:, :, :T].reshape(B, H, T, T), sref) < 2e-3
    p = torch.zeros(B * H, T, Tp, device="cuda", dtype=torch.bfloat16)
    p[:, :, :T] = torch.softmax(sref, -1).reshape(B * H, T, T).to(torch.bfloat16)
    o = torch.full((B, T, H * dk), float("nan"), device="cuda", dtype=torch.bfloat16)
    pa = G.operand(p, Tp, T, B * H)
    vb = G.operand(qkv, C, T, B, mn_major=True, inner_base=2 * H * dk, zdiv=H, zmod_stride=dk)
    G.gemm(pa, vb, o, T, dk, T, Z=B * H, ldd=H * dk, d_zdiv=H, d_zdiv_stride=T * H * dk,
           d_zmod_stride=dk, impl=impl)
    oref = (p[:, :, :T].float().reshape(B, H, T, T) @ v).permute(0, 2, 1, 3).reshape(B, T, H * dk)
    assert torch.isfinite(o.float()).all()
    assert rel_err(o, oref) < 1e-2, rel_err(o, oref)


@pytest.mark.parametrize("impl", IMPLS)
def test_mn_major_both(impl):
    """dK = dS^T Q : A = dS^T (MN-major), B = Q^T (MN-major), batched."""
    torch.manual_seed(5)
    Z, Tq, Tk, dk = 4, 150, 150, 128
    Tp = 192
    ds = torch.zeros(Z, Tq, Tp, device="cuda", dtype=torch.bfloat16)
    ds[:, :, :Tk] = rnd(Z, Tq, Tk)
    q = rnd(Z, Tq, dk)
    out = torch.full((Z, Tk, dk), float("nan"), device="cuda", dtype=torch.bfloat16)
    a = G.operand(ds, Tk, Tq, Z, ld=Tp, batch_stride=Tq * Tp, mn_major=True)
    b = G.operand(q, dk, Tq, Z, mn_major=True)
    G.gemm(a, b, out, Tk, dk, Tq, Z=Z, ldd=dk, d_zdiv=1, d_zdiv_stride=Tk * dk, impl=impl)
    ref = ds[:, :, :Tk].float().transpose(1, 2) @ q.float()
    assert rel_err(out, ref) < 1e-2, rel_err(out, ref)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("epi", ["relu_bwd", "add_aux"])
def test_aux_epilogues(impl, epi):
    torch.manual_seed(9)
    M, N, K = 333, 1024, 256
    x, w = rnd(M, K), rnd(N, K, scale=K ** -0.5)
    aux = rnd(M, N)
    y = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    G.gemm(G.operand(x, K, M), G.operand(w, K, N), y, M, N, K, aux=aux, ld_aux=N,
           epilogue=G.EPI_RELU_BWD if epi == "relu_bwd" else G.EPI_ADD_AUX, impl=impl)
    ref = x.float() @ w.float().t()
    ref = ref * (aux.float() > 0) if epi == "relu_bwd" else ref + aux.float()
    assert rel_err(y, ref) < 1e-2, rel_err(y, ref)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("M,N,K,splits", [(1000, 768, 256, 1), (1000, 768, 256, 5), (3000, 256, 1024, 7),
                                          (500, 80, 256, 3)])
def test_wgrad_linear(impl, M, N, K, splits):
    """dW[N,K] = dY^T X, accumulated on top of existing contents."""
    torch.manual_seed(M)
    dy, x = rnd(M, N), rnd(M, K)
    dw = torch.ones(N, K, device="cuda")
    G.wgrad(G.operand(dy, N, M, mn_major=True), G.operand(x, K, M, mn_major=True), dw, N, K,
            splits=splits, impl=impl)
    ref = dy.float().t() @ x.float() + 1.0
    assert rel_err(dw, ref) < 2e-3, rel_err(dw, ref)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("B,T,Cin,Cout,k,splits", [(3, 200, 256, 1024, 9, 2), (2, 77, 256, 256, 3, 1),
                                                   (2, 130, 80, 512, 5, 3), (2, 130, 512, 80, 5, 2)])
def test_wgrad_conv(impl, B, T, Cin, Cout, k, splits):
    """dW accumulated in the GEMM-friendly [co][tap][ci] layout (unit column stride, vector atomics)."""
    torch.manual_seed(T)
    x = rnd(B, T, Cin).float().requires_grad_()
    w = torch.zeros(Cout, Cin, k, device="cuda", requires_grad=True)
    dy = rnd(B, T, Cout)
    y = torch.nn.functional.conv1d(x.transpose(1, 2), w, None, padding=(k - 1) // 2).transpose(1, 2)
    y.backward(dy.float())
    dw = torch.zeros(Cout, k, Cin, device="cuda")
    xb = x.detach().to(torch.bfloat16)
    G.wgrad(G.operand(dy, Cout, T, B, mn_major=True), G.operand(xb, Cin, T, B, mn_major=True), dw,
            Cout, Cin, taps=k, tap_shift0=-((k - 1) // 2), ldd=Cin * k, d_col_stride=1,
            d_tap_stride=Cin, splits=splits, impl=impl)
    assert rel_err(dw, w.grad.permute(0, 2, 1)) < 2e-3, rel_err(dw, w.grad.permute(0, 2, 1))


@pytest.mark.parametrize("impl", IMPLS)
def test_wgrad_row_segments(impl):
    """Fused Q|K|V weight gradient: row blocks of one [3*HD, D] result land in three tensors."""
    torch.manual_seed(11)
    M, HD, D = 1500, 256, 256
    dy, x = rnd(M, 3 * HD), rnd(M, D)
    outs = [torch.full((HD, D), float(i), device="cuda") for i in range(3)]
    G.wgrad(G.operand(dy, 3 * HD, M, mn_major=True), G.operand(x, D, M, mn_major=True), None, 3 * HD, D, splits=3,
            segments=(HD, outs), impl=impl)
    ref = dy.float().t() @ x.float()
    for i in range(3):
        assert rel_err(outs[i], ref[i * HD:(i + 1) * HD] + float(i)) < 2e-3


@pytest.mark.parametrize("B,T,Cin,Cout,k", [(3, 200, 256, 1024, 9), (2, 130, 80, 512, 5), (2, 77, 1024, 256, 1)])
def test_ops_conv_wgrad_reference_layout(B, T, Cin, Cout, k):
    ops = sub("ops")
    torch.manual_seed(T)
    x = rnd(B, T, Cin)
    dy = rnd(B, T, Cout)
    w = torch.zeros(Cout, Cin, k, device="cuda", requires_grad=True)
    y = torch.nn.functional.conv1d(x.float().transpose(1, 2), w, None, padding=(k - 1) // 2).transpose(1, 2)
    y.backward(dy.float())
    dw = torch.ones(Cout, Cin, k, device="cuda")
    ops.conv_wgrad(dy, x, dw)
    assert rel_err(dw, w.grad + 1.0) < 2e-3


# --------------------------------------------------------------------------------------------------
# ragged rows (fs2_gemm::row_lens): padded frames are skipped, their output rows are zero
# --------------------------------------------------------------------------------------------------
def _lens(vals):
    return torch.tensor(vals, device="cuda", dtype=torch.int64)


RAGGED_LENS = [[200, 1, 129, 0, 128, 77], [0, 0, 0], [300, 300], [5]]


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("lens", RAGGED_LENS)
@pytest.mark.parametrize("T,N,K,f32", [(300, 768, 256, False), (300, 256, 1024, False), (257, 80, 256, True),
                                       (300, 1024, 256, False)])
def test_ragged_linear(impl, lens, T, N, K, f32):
    """Batched per-utterance linear layer: valid rows equal the dense result, padded rows are exactly 0."""
    B = len(lens)
    torch.manual_seed(T + N + K + B)
    x, w = rnd(B, T, K), rnd(N, K, scale=K ** -0.5)
    bias = torch.randn(N, device="cuda")
    aux = None if f32 else rnd(B, T, N)
    y = torch.full((B, T, N), float("nan"), device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
    G.gemm(G.operand(x, K, T, B), G.operand(w, K, N), y, T, N, K, Z=B, bias=bias, d_zdiv=1, d_zdiv_stride=T * N,
           aux=aux, ld_aux=N, aux_batch_stride=T * N, epilogue=G.EPI_NONE if f32 else G.EPI_ADD_AUX,
           row_lens=_lens(lens), impl=impl)
    ref = x.float() @ w.float().t() + bias
    if aux is not None:
        ref = ref + aux.float()
    valid = torch.arange(T, device="cuda")[None, :] < _lens(lens).clamp(max=T)[:, None]
    ref = ref * valid[..., None]
    assert torch.isfinite(y.float()).all()
    assert (y[~valid] == 0).all()
    if valid.any():
        assert rel_err(y, ref) < (2e-3 if f32 else 1e-2), rel_err(y, ref)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("lens", [[200, 1, 129, 0, 128, 77], [0, 0], [260, 255, 256, 257], [129, 1, 1], [7]])
@pytest.mark.parametrize("Cin,Cout,k", [(256, 1024, 9), (1024, 256, 1), (256, 256, 3), (80, 512, 5)])
def test_ragged_conv(impl, lens, Cin, Cout, k):
    """Implicit-GEMM Conv1d with row_lens: same values as the dense conv on valid rows (the halo may read
    padded input rows, which the caller keeps at zero), zeros on padded rows."""
    B, T = len(lens), 260
    torch.manual_seed(Cin + k + B)
    ln = _lens(lens).clamp(max=T)
    valid = torch.arange(T, device="cuda")[None, :] < ln[:, None]
    x = rnd(B, T, Cin) * valid[..., None]
    w = (torch.randn(Cout, Cin, k, device="cuda") * (Cin * k) ** -0.5).to(torch.bfloat16)
    bias = torch.randn(Cout, device="cuda")
    cpad = ((Cin + 63) // 64) * 64  # packed weights [Co][k][Cpad] (zero-padded channels), as ops.pack_conv
    wp = torch.zeros(Cout, k, cpad, device="cuda", dtype=torch.bfloat16)
    wp[:, :, :Cin] = w.permute(0, 2, 1)
    y = torch.full((B, T, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    G.gemm(G.operand(x, Cin, T, B), G.operand(wp, k * cpad, Cout), y, T, Cout, Cin, Z=B, taps=k,
           tap_shift0=-((k - 1) // 2), b_tap_kstride=cpad, bias=bias, epilogue=G.EPI_RELU, d_zdiv=1,
           d_zdiv_stride=T * Cout, row_lens=_lens(lens), impl=impl)
    ref = torch.nn.functional.conv1d(x.float().transpose(1, 2), w.float(), bias,
                                     padding=(k - 1) // 2).transpose(1, 2).relu() * valid[..., None]
    assert (y[~valid] == 0).all()
    if valid.any():
        assert rel_err(y, ref) < 1e-2, rel_err(y, ref)


def _split_lens(n_pair_tiles, T, seed):
    """Utterance lengths whose 128-row tiles add up to 2 * n_pair_tiles (- 1 when `seed` is odd: filler half)."""
    g = torch.Generator().manual_seed(seed)
    want = 2 * n_pair_tiles - (seed & 1)
    lens, tiles = [], 0
    while tiles < want:
        t = min(int(torch.randint(1, T // 128 + 1, (1,), generator=g)), want - tiles)
        lens.append(min(T, t * 128 - int(torch.randint(0, 100, (1,), generator=g))))
        tiles += t
    return lens


@pytest.mark.parametrize("Cin,Cout,k,epi,n_pair_tiles", [
    (1024, 256, 9, "aux", 157),      # FFN k=9 input gradient at C2: rounds of 74, 74, 9 -> 4 slices of 4 channel blocks
    (1024, 256, 9, "aux", 104),      # remainder 30 -> 2 slices
    (1024, 512, 5, "relu_mask", 80),  # two column tiles: 160 pair tiles, remainder 12 -> 4 slices; mask written
    (1024, 256, 5, "relu_bwd_mask", 85),  # mask read, odd tile count (filler half of the last pair)
    (1024, 512, 5, "bias", 41),      # 82 pair tiles, remainder 8
    (256, 1024, 9, "relu_mask", 40),  # short reduction (36 k-blocks): scheduled without a split
])
def test_conv_split_tail(Cin, Cout, k, epi, n_pair_tiles):
    """Last partial round of the persistent Conv1d schedule split over the channel blocks (fs2_gemm::workspace):
    same values as the CUDA-core kernel, bit-identical from launch to launch (fixed-order reduction; the counters in
    the workspace return to zero)."""
    T = 1000
    seed = n_pair_tiles + k
    lens = _split_lens(n_pair_tiles, T, seed)
    B = len(lens)
    torch.manual_seed(seed)
    ln = _lens(lens)
    valid = torch.arange(T, device="cuda")[None, :] < ln[:, None]
    x = rnd(B, T, Cin) * valid[..., None]
    wp = (torch.randn(Cout, k, Cin, device="cuda") * (Cin * k) ** -0.5).to(torch.bfloat16)
    bias = torch.randn(Cout, device="cuda") if epi in ("bias", "relu_mask") else None
    aux = rnd(B, T, Cout) if epi == "aux" else None
    mask = None
    if epi in ("relu_mask", "relu_bwd_mask"):
        mask = torch.randint(-2 ** 62, 2 ** 62, (B * T, Cout // 64), device="cuda", dtype=torch.int64)
    epilogue = {"aux": G.EPI_ADD_AUX, "bias": G.EPI_NONE, "relu_mask": G.EPI_RELU, "relu_bwd_mask": G.EPI_RELU_BWD}[epi]

    def run(impl, m):
        guard = 8192  # the split-K epilogue computes its own addresses: nothing may land outside D
        buf = torch.full((guard + B * T * Cout + guard,), float("nan"), device="cuda", dtype=torch.bfloat16)
        buf[:guard] = 7.0
        buf[-guard:] = 7.0
        y = buf[guard:guard + B * T * Cout].view(B, T, Cout)
        G.gemm(G.operand(x, Cin, T, B), G.operand(wp, k * Cin, Cout), y, T, Cout, Cin, Z=B, taps=k,
               tap_shift0=-((k - 1) // 2), b_tap_kstride=Cin, bias=bias, epilogue=epilogue, aux=aux, ld_aux=Cout,
               aux_batch_stride=T * Cout, d_zdiv=1, d_zdiv_stride=T * Cout, row_lens=ln, relu_mask=m, impl=impl)
        assert (buf[:guard] == 7.0).all() and (buf[-guard:] == 7.0).all()
        return y

    m0 = mask.clone() if mask is not None else None
    m1 = mask.clone() if mask is not None else None
    if epi == "relu_mask":
        m1.zero_()  # the CUDA-core cross-check ORs its bits in
    ref = run(1, m1)
    y = run(0, m0)
    assert torch.isfinite(y.float()).all() and (y[~valid] == 0).all()
    assert rel_err(y, ref) < 4e-3, rel_err(y, ref)
    if epi == "relu_mask":  # the mask words of valid rows agree except where the pre-activation rounds across zero
        vm = valid.reshape(-1)
        diff = (m0[vm] ^ m1[vm])
        flipped = sum(bin(int(v) & (2 ** 64 - 1)).count("1") for v in diff[diff != 0].tolist())
        assert flipped <= 1e-4 * vm.sum().item() * Cout
    for _ in range(3):  # counters reset, fixed summation order
        m2 = mask.clone() if mask is not None else None
        assert torch.equal(run(0, m2), y)
        if epi == "relu_mask":
            assert torch.equal(m2[valid.reshape(-1)], m0[valid.reshape(-1)])


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("lens", [[200, 1, 129, 0, 128, 77], [0, 0], [300, 64, 65]])
@pytest.mark.parametrize("N,K,k,splits", [(768, 256, 1, 1), (768, 256, 1, 4), (256, 1024, 9, 3), (80, 256, 1, 2)])
def test_ragged_wgrad(impl, lens, N, K, k, splits):
    """Weight gradient with row_lens: reduction blocks of padded frames are skipped (dY is zero there)."""
    B, T = len(lens), 300
    torch.manual_seed(N + K + k + B)
    ln = _lens(lens).clamp(max=T)
    valid = torch.arange(T, device="cuda")[None, :] < ln[:, None]
    dy = rnd(B, T, N) * valid[..., None]
    x = rnd(B, T, K)  # padded rows of x are arbitrary (finite): they only ever meet zero rows of dY
    if k > 1:
        x = x * valid[..., None]
    dw = torch.ones(N, k, K, device="cuda")
    G.wgrad(G.operand(dy, N, T, B, mn_major=True), G.operand(x, K, T, B, mn_major=True), dw, N, K, taps=k,
            tap_shift0=-((k - 1) // 2), ldd=K * k, d_col_stride=1, d_tap_stride=K, splits=splits,
            row_lens=_lens(lens), impl=impl)
    xf = x.float().requires_grad_(False)
    w = torch.zeros(N, K, k, device="cuda", requires_grad=True)
    yy = torch.nn.functional.conv1d(xf.transpose(1, 2), w, None, padding=(k - 1) // 2).transpose(1, 2)
    yy.backward(dy.float())
    ref = w.grad.permute(0, 2, 1) + 1.0
    assert rel_err(dw, ref) < 2e-3, rel_err(dw, ref)


@pytest.mark.parametrize("Co,Ci,k,lens,T", [
    (1024, 256, 9, [300, 64, 65, 1, 0, 257], 300),    # FFN conv: tap groups 3 + 3 + 3
    (512, 512, 5, [130] * 7, 130),                      # PostNet conv (dense): groups 3 + 2 with their own split counts
    (256, 256, 3, [200, 17, 90], 200),                  # predictor conv: one group
    (256, 384, 2, [100, 100], 100),                     # even kernel size
    (768, 128, 7, [500, 333], 500),                     # 1.5 co pair tiles, one ci tile, groups 4 + 3
])
def test_conv_wgrad_tap_groups(Co, Ci, k, lens, T):
    """csrc/wgrad_taps.cu: one X tile with halo serves all taps of a group (row-shifted MN-major descriptors)."""
    B = len(lens)
    torch.manual_seed(Co + Ci + k)
    ln = _lens(lens).clamp(max=T)
    valid = torch.arange(T, device="cuda")[None, :] < ln[:, None]
    dy = rnd(B, T, Co) * valid[..., None]
    x = rnd(B, T, Ci) * valid[..., None]
    dw = torch.full((Co, k, Ci), 0.5, device="cuda")
    shift = -((k - 1) // 2)
    G.wgrad(G.operand(dy, Co, T, B, mn_major=True), G.operand(x, Ci, T, B, mn_major=True), dw, Co, Ci, taps=k,
            tap_shift0=shift, ldd=Ci * k, d_col_stride=1, d_tap_stride=Ci, splits=4, row_lens=_lens(lens))
    # reference: dw[co][tap][ci] = sum_t dy[t][co] * x[t + tap + shift][ci]
    xf, dyf = x.float(), dy.float()
    ref = torch.empty(Co, k, Ci, device="cuda")
    for tap in range(k):
        s_ = tap + shift
        xs = torch.zeros_like(xf)
        if s_ >= 0:
            xs[:, :T - s_] = xf[:, s_:]
        else:
            xs[:, -s_:] = xf[:, :T + s_]
        ref[:, tap] = torch.einsum("btc,bti->ci", dyf, xs)
    assert rel_err(dw, ref + 0.5) < 2e-3, rel_err(dw, ref + 0.5)


def test_split_tail_scratch_is_per_graph_capture():
    """fs2_gemm::workspace handed out by gemm.py: every CUDA-graph capture gets its own buffer whose counters are
    zeroed by a node of THAT graph -- a graph that is replayed before any other graph (or eager call) has run on its
    stream, and after the scratch of an earlier capture was dirtied, still reduces its split-K tail correctly."""
    Cin, Cout, k, T = 1024, 256, 9, 1000
    lens = _split_lens(157, T, 3)
    B = len(lens)
    torch.manual_seed(11)
    ln = _lens(lens)
    valid = torch.arange(T, device="cuda")[None, :] < ln[:, None]
    x = rnd(B, T, Cin) * valid[..., None]
    wp = (torch.randn(Cout, k, Cin, device="cuda") * (Cin * k) ** -0.5).to(torch.bfloat16)

    def conv(y):
        G.gemm(G.operand(x, Cin, T, B), G.operand(wp, k * Cin, Cout), y, T, Cout, Cin, Z=B, taps=k,
               tap_shift0=-((k - 1) // 2), b_tap_kstride=Cin, d_zdiv=1, d_zdiv_stride=T * Cout, row_lens=ln)

    ref = torch.empty(B, T, Cout, device="cuda", dtype=torch.bfloat16)
    conv(ref)
    torch.cuda.synchronize()
    outs, graphs = [], []
    for _ in range(2):
        y = torch.full((B, T, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            conv(y)
        outs.append(y)
        graphs.append(g)
    # dirty every cached scratch buffer (as a kernel aborted half-way would), then replay the SECOND graph first
    for ws in list(G._WS.values()):
        ws.fill_(0x5A)
    graphs[1].replay()
    graphs[0].replay()
    torch.cuda.synchronize()
    for ws in list(G._WS.values()):
        ws[:4096].zero_()  # leave the eager buffers usable for the tests that follow
    assert torch.equal(outs[1], ref) and torch.equal(outs[0], ref)


@pytest.mark.parametrize("Co,Ci,k,lens,T,ragged", [
    (1024, 256, 9, [300, 64, 65, 1, 0, 257], 300, True),   # FFN conv: fused into the tap-group kernel, 2 ci tiles
    (1024, 256, 9, [1000, 613, 127, 899] * 4, 1000, True), # many frame blocks per split, both CTAs of every pair
    (512, 512, 5, [130] * 7, 130, False),                  # dense, 4 ci tiles, groups of different size
    (768, 128, 7, [500, 333], 500, True),                  # 1.5 co pair tiles (rows >= M of the last tile are skipped)
    (384, 256, 3, [200, 17, 90], 200, True),               # 1.5 co pair tiles in the fused kernel: the peer CTA of the last pair has no rows
    (256, 256, 1, [200, 17, 90], 200, True),               # k = 1: not the tap-group kernel -> separate column-sum launch
    (256, 80, 5, [64, 200], 200, False),                   # Ci % 128 != 0: fallback, dense
])
def test_conv_wgrad_bias_gradient(Co, Ci, k, lens, T, ragged):
    """fs2_gemm::a_colsum: db[co] += sum_{z,t} dY[z][t][co] next to the weight gradient -- read from the dY tiles of
    csrc/wgrad_taps.cu (the peer CTA's through distributed shared memory), else one column-sum launch."""
    B = len(lens)
    torch.manual_seed(Co + Ci + k + B)
    ln = _lens(lens).clamp(max=T)
    valid = torch.arange(T, device="cuda")[None, :] < ln[:, None]
    dy = rnd(B, T, Co) * valid[..., None]
    x = rnd(B, T, Ci) * valid[..., None]
    shift = -((k - 1) // 2)
    out = []
    for with_db in (False, True):
        dw = torch.full((Co, k, Ci), 0.5, device="cuda")
        db = torch.full((Co + 8,), 0.25, device="cuda")  # 8 guard elements behind the bias gradient
        G.wgrad(G.operand(dy, Co, T, B, mn_major=True), G.operand(x, Ci, T, B, mn_major=True), dw, Co, Ci, taps=k,
                tap_shift0=shift, ldd=Ci * k, d_col_stride=1, d_tap_stride=Ci, splits=4,
                row_lens=_lens(lens) if ragged else None, a_colsum=db[:Co] if with_db else None)
        out.append((dw, db))
    (dw0, db0), (dw1, db1) = out
    assert (db0 == 0.25).all() and (db1[Co:] == 0.25).all()
    ref = dy.float().sum(dim=(0, 1)) + 0.25
    assert rel_err(db1[:Co], ref) < 1e-5, rel_err(db1[:Co], ref)
    assert rel_err(dw1, dw0) < 1e-5  # the weight gradient itself is unchanged (fp32 atomics: order only)


@pytest.mark.parametrize("HD,D,lens,T", [
    (256, 256, [1000, 613, 127, 899, 64, 1] * 3, 1000),   # the QKV projection at C2 scale: 2-CTA kernel, readers on
    (256, 256, None, 700),                                # dense single batch
    (128, 512, [300, 17], 300),                           # 384 rows = 1.5 pair tiles, two column tiles share the duty
    (64, 256, [200, 90], 200),                            # M = 192 < 256: 1-CTA kernel -> column-sum launches
])
def test_wgrad_segmented_bias_gradients(HD, D, lens, T):
    """fs2_gemm::a_colsum_seg: the Q / V bias gradients (column sums of row blocks 0 and 2 of dQKV) next to the fused
    QKV weight gradient; the K block is skipped (NULL)."""
    torch.manual_seed(HD + D + T)
    if lens is None:
        B, ln = 1, None
        valid = torch.ones(1, T, 1, device="cuda", dtype=torch.bool)
    else:
        B, ln = len(lens), _lens(lens)
        valid = (torch.arange(T, device="cuda")[None, :] < ln[:, None])[..., None]
    dqkv = rnd(B, T, 3 * HD) * valid
    x = rnd(B, T, D) * valid
    ws = [torch.full((HD, D), 0.5, device="cuda") for _ in range(3)]
    bs = [torch.full((HD + 8,), 0.25, device="cuda") for _ in range(3)]
    a = G.operand(dqkv, 3 * HD, T, B, mn_major=True)
    b = G.operand(x, D, T, B, mn_major=True)
    G.wgrad(a, b, None, 3 * HD, D, splits=5, segments=(HD, ws), row_lens=ln,
            a_colsum_seg=[bs[0][:HD], None, bs[2][:HD]])
    ref_w = torch.einsum("btc,btd->cd", dqkv.float(), x.float()) + 0.5
    ref_b = dqkv.float().sum((0, 1)) + 0.25
    for i in range(3):
        assert rel_err(ws[i], ref_w[i * HD:(i + 1) * HD]) < 2e-3
        assert (bs[i][HD:] == 0.25).all()
    assert rel_err(bs[0][:HD], ref_b[:HD]) < 1e-5 and rel_err(bs[2][:HD], ref_b[2 * HD:]) < 1e-5
    assert (bs[1] == 0.25).all()


def test_ragged_many_batches_falls_back_to_dense_schedule():
    """More utterances than the on-chip schedule table holds: every tile is computed, rows still zeroed."""
    B, T, N, K = 300, 40, 256, 256
    torch.manual_seed(3)
    lens = torch.randint(0, T + 1, (B,), device="cuda", dtype=torch.int64)
    x, w = rnd(B, T, K), rnd(N, K, scale=K ** -0.5)
    y = torch.full((B, T, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    G.gemm(G.operand(x, K, T, B), G.operand(w, K, N), y, T, N, K, Z=B, d_zdiv=1, d_zdiv_stride=T * N,
           row_lens=lens)
    valid = torch.arange(T, device="cuda")[None, :] < lens[:, None]
    ref = (x.float() @ w.float().t()) * valid[..., None]
    assert (y[~valid] == 0).all() and rel_err(y, ref) < 1e-2


@pytest.mark.parametrize("tail", [-1, 4, 0])
@pytest.mark.parametrize("T,N", [(700, 1024), (200, 256)])
def test_ragged_tail_rows(tail, T, N):
    """fs2_gemm::tail_zero_rows: rows behind the last scheduled row tile are zeroed completely (0), for an
    n-row halo only (n > 0) or not at all (< 0); rows inside scheduled tiles are always defined."""
    lens = [T, 1, 130, 0, 300][:5]
    B, K = len(lens), 256
    torch.manual_seed(tail + T)
    x, w = rnd(B, T, K), rnd(N, K, scale=K ** -0.5)
    y = torch.full((B, T, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    G.gemm(G.operand(x, K, T, B), G.operand(w, K, N), y, T, N, K, Z=B, d_zdiv=1, d_zdiv_stride=T * N,
           row_lens=_lens(lens), tail_rows=tail)
    ref = x.float() @ w.float().t()
    unit = 128  # scheduling unit of the ragged schedule (2-CTA kernels pair up any two 128-row tiles)
    for b, ln in enumerate(lens):
        ln = min(ln, T)
        covered = min(-(-ln // unit) * unit, T)
        assert rel_err(y[b, :ln], ref[b, :ln]) < 1e-2 if ln else True
        assert (y[b, ln:covered] == 0).all()
        z_end = T if tail == 0 else min(T, covered + max(tail, 0))
        assert (y[b, covered:z_end] == 0).all()
        assert torch.isnan(y[b, z_end:].float()).all()  # untouched


@pytest.mark.parametrize("case", ["ragged_conv", "ragged_linear", "dense_conv", "wgrad"])
def test_outputs_stay_inside_their_buffer(case):
    """Guard bands around the output: no kernel of the engine may write a byte outside D (compute-sanitizer is not
    available on the test pool, so the bounds are checked with sentinels)."""
    torch.manual_seed(17)
    B, T, Cin, Cout, k = 5, 333, 256, 512, 5
    lens = _lens([333, 0, 130, 257, 64])
    x = rnd(B, T, Cin)
    guard = 4096
    if case == "wgrad":
        dy = rnd(B, T, Cout)
        buf = torch.full((guard + Cout * k * Cin + guard,), 7.0, device="cuda")
        dw = buf[guard:guard + Cout * k * Cin].view(Cout, k, Cin)
        dw.zero_()
        G.wgrad(G.operand(dy, Cout, T, B, mn_major=True), G.operand(x, Cin, T, B, mn_major=True), dw, Cout, Cin,
                taps=k, tap_shift0=-2, ldd=Cin * k, d_col_stride=1, d_tap_stride=Cin, splits=3, row_lens=lens)
        assert torch.isfinite(dw).all()
    else:
        n = B * T * Cout
        buf = torch.full((guard + n + guard,), 7.0, device="cuda", dtype=torch.bfloat16)
        y = buf[guard:guard + n].view(B, T, Cout)
        if case == "ragged_linear":
            w = rnd(Cout, Cin)
            G.gemm(G.operand(x, Cin, T, B), G.operand(w, Cin, Cout), y, T, Cout, Cin, Z=B, d_zdiv=1,
                   d_zdiv_stride=T * Cout, row_lens=lens, tail_rows=3)
        else:
            wp = rnd(Cout, k * Cin)
            G.gemm(G.operand(x, Cin, T, B), G.operand(wp, k * Cin, Cout), y, T, Cout, Cin, Z=B, taps=k, tap_shift0=-2,
                   b_tap_kstride=Cin, d_zdiv=1, d_zdiv_stride=T * Cout,
                   row_lens=lens if case == "ragged_conv" else None)
    torch.cuda.synchronize()
    assert (buf[:guard] == 7.0).all() and (buf[-guard:] == 7.0).all()


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("lens", [None, [300, 1, 129, 0, 128]])
def test_relu_bitmask_forward_and_backward(impl, lens):
    """fs2_gemm::relu_mask: the ReLU forward writes a 1-bit mask of its positive outputs, the ReLU-backward
    epilogue of a later GEMM reads it instead of the bf16 activation (same result as the `aux` path)."""
    torch.manual_seed(29)
    B, T, K, N = 5, 300, 256, 1024
    x, w = rnd(B, T, K), rnd(N, K, scale=K ** -0.5)
    bias = torch.randn(N, device="cuda") * 0.1
    rl = None if lens is None else _lens(lens)
    h = torch.zeros(B, T, N, device="cuda", dtype=torch.bfloat16)
    mask = torch.zeros(B * T, N // 64, device="cuda", dtype=torch.int64)
    G.gemm(G.operand(x, K, T, B), G.operand(w, K, N), h, T, N, K, Z=B, bias=bias, epilogue=G.EPI_RELU, d_zdiv=1,
           d_zdiv_stride=T * N, row_lens=rl, relu_mask=mask, impl=impl)
    valid = torch.ones(B, T, dtype=torch.bool, device="cuda") if lens is None else \
        torch.arange(T, device="cuda")[None, :] < rl[:, None]
    bits = ((mask.view(B, T, N // 64, 1) >> torch.arange(64, device="cuda")) & 1).bool().view(B, T, N)
    assert torch.equal(bits[valid], (h > 0)[valid])
    # backward of the ReLU through a second GEMM: dh = (dy @ W2) * (h > 0), once with the mask, once with aux = h
    dy, w2 = rnd(B, T, 256), rnd(256, N, scale=256 ** -0.5)  # W2 [256, N] read as an MN-major B operand
    outs = []
    for kw in (dict(relu_mask=mask), dict(aux=h, ld_aux=N, aux_batch_stride=T * N)):
        dh = torch.full((B, T, N), float("nan"), device="cuda", dtype=torch.bfloat16)
        G.gemm(G.operand(dy, 256, T, B), G.operand(w2, N, 256, mn_major=True), dh, T, N, 256, Z=B,
               epilogue=G.EPI_RELU_BWD, d_zdiv=1, d_zdiv_stride=T * N, row_lens=rl, impl=impl, **kw)
        outs.append(dh)
    assert torch.equal(outs[0][valid], outs[1][valid])
    ref = (dy.float() @ w2.float()) * (h > 0)
    assert rel_err(outs[0][valid], ref[valid]) < 1e-2


# ---- gemm_sk.cu: B-stationary small-K kernel (resident weights, TMA-store epilogue) ----------------------------------
@pytest.mark.parametrize("lens", [None, [300, 1, 129, 0, 128, 257], [128, 128, 128], [5], [0, 0, 300]])
@pytest.mark.parametrize("N,K,b_mn", [(768, 256, False), (256, 256, False), (1024, 256, True), (256, 128, True),
                                      (320, 192, False), (512, 80, False)])
def test_small_k_resident_weight_kernel(lens, N, K, b_mn):
    """Shapes the dispatcher sends to gemm_sk_kernel (NORMAL, K <= 256, N >= 256, bf16 output): K- and MN-major
    weights, partial column blocks (N = 320), K that is not a multiple of 64, odd numbers of row tiles (filler
    half of the last CTA pair), empty utterances; the output is a column slice of a wider buffer (ldd > N) whose
    other columns and whose rows past each utterance's T must stay untouched (TMA store clipping)."""
    torch.manual_seed(N + K + (0 if lens is None else len(lens)))
    B, T = (3, 300) if lens is None else (len(lens), 300)
    x = rnd(B, T, K)
    w = rnd(K, N, scale=K ** -0.5) if b_mn else rnd(N, K, scale=K ** -0.5)
    bias = torch.randn(N, device="cuda")
    ldd = N + 64
    buf = torch.full((B, T + 3, ldd), 7.0, device="cuda", dtype=torch.bfloat16)  # 3 guard rows per utterance
    d = buf[:, :T, 32:32 + N]
    rl = None if lens is None else _lens(lens)
    bop = G.operand(w, N, K, mn_major=True) if b_mn else G.operand(w, K, N)
    G.gemm(G.operand(x, K, T, B), bop, d, T, N, K, Z=B, ldd=ldd, bias=bias, d_zdiv=1, d_zdiv_stride=(T + 3) * ldd,
           row_lens=rl)
    ref = x.float() @ (w.float() if b_mn else w.float().t()) + bias
    if lens is not None:
        valid = torch.arange(T, device="cuda")[None, :] < rl[:, None]
        ref = ref * valid[..., None]
    assert rel_err(d, ref) < 1e-2 or float(ref.norm()) == 0.0
    if lens is not None:
        assert float(d.float()[~valid].abs().sum()) == 0.0  # padded frames are written as zeros (tail_rows = 0)
    assert (buf[:, T:, :] == 7.0).all() and (buf[:, :, :32] == 7.0).all() and (buf[:, :, 32 + N:] == 7.0).all()
