"""tcgen05 GEMM vs torch fp32 matmul on bf16-rounded operands (GPU)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

# bf16 output rounding is 2^-9 relative; accumulation is fp32.
RTOL, ATOL = 1.0e-2, 1.0e-2


def _mk(shape, dev, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev).to(torch.bfloat16)


SHAPES = [
    (128, 128, 64), (128, 128, 128), (256, 256, 256), (200, 384, 128), (1000, 512, 128),
    (4096, 128, 512), (2560, 768, 768), (2560, 3072, 768), (2560, 768, 3072), (777, 1024, 1024),
    (9216, 48 + 16, 128), (36864, 1536, 512), (64, 768, 768), (100, 8, 64), (333, 264, 200),
]


@pytest.mark.parametrize("m,n,k", SHAPES)
def test_gemm_kmajor(cuda_dev, m, n, k):
    from fiber_b200 import kernels as K
    a = _mk((m, k), cuda_dev, 1)
    b = _mk((n, k), cuda_dev, 2, k ** -0.5)
    out = K.gemm(a, b)
    ref = a.float() @ b.float().t()
    torch.testing.assert_close(out.float(), ref, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("m,n,k", [(384, 128, 4096), (128, 128, 64), (512, 128, 100000), (768, 3072, 2560),
                                   (1536, 512, 36864), (264, 200, 333), (128, 48 + 16, 9216)])
def test_gemm_mnmajor_wgrad(cuda_dev, m, n, k):
    """dW[m,n] = dY[k,m]^T X[k,n] with fp32 atomic split-K accumulation."""
    from fiber_b200 import kernels as K
    dy = _mk((k, m), cuda_dev, 3, k ** -0.5)
    x = _mk((k, n), cuda_dev, 4)
    out = K.gemm(dy, x, mn_major=True, accumulate=True)
    ref = dy.float().t() @ x.float()
    torch.testing.assert_close(out, ref, rtol=2e-3, atol=2e-3)
    out1 = K.gemm(dy, x, mn_major=True, out_dtype=torch.float32)
    torch.testing.assert_close(out1, ref, rtol=2e-3, atol=2e-3)
    # fused bias gradient: db[m] = scale * sum_k dy[k, m] from the same kernel (ones-tile MMA)
    db = torch.zeros(m, device=cuda_dev)
    alpha = torch.tensor([0.5], device=cuda_dev)
    out2 = K.gemm(dy, x, mn_major=True, accumulate=True, scale=alpha, colsum=db)
    torch.testing.assert_close(out2, 0.5 * ref, rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(db, 0.5 * dy.float().sum(0), rtol=2e-3, atol=2e-3 * max(1.0, k ** 0.5 * k ** -0.5))


def test_gemm_epilogue(cuda_dev):
    from fiber_b200 import kernels as K
    m, n, k = 1000, 512, 128
    a = _mk((m, k), cuda_dev, 5)
    b = _mk((n, k), cuda_dev, 6, k ** -0.5)
    bias = torch.randn(n, device=cuda_dev)
    res = _mk((m, n), cuda_dev, 7)
    scale = torch.tensor([0.5], device=cuda_dev)
    row_scale = torch.rand(m // 100, device=cuda_dev) + 0.5
    pre = torch.empty((m, n), device=cuda_dev, dtype=torch.bfloat16)
    out = K.gemm(a, b, bias=bias, residual=res, preact=pre, scale=scale, row_scale=row_scale,
                 rows_per_scale=100, act=K.ACT_GELU)
    h = a.float() @ b.float().t() + bias
    ref = torch.nn.functional.gelu(h) * 0.5 * row_scale.repeat_interleave(100)[:, None] + res.float()
    torch.testing.assert_close(pre.float(), h, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(out.float(), ref, rtol=RTOL, atol=ATOL)
    # GELU-grad epilogue: out = (a b^T) * gelu'(aux)
    aux = _mk((m, n), cuda_dev, 8)
    out2 = K.gemm(a, b, aux=aux, act=K.ACT_GELU_GRAD)
    x = aux.float().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    ref2 = (a.float() @ b.float().t()) * x.grad
    torch.testing.assert_close(out2.float(), ref2, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("m,n,k", [(1000, 512, 128), (333, 264, 200), (4096, 128, 256), (130, 72, 64)])
def test_gemm_tma_store_epilogue(cuda_dev, m, n, k):
    """No residual / aux => thread=row epilogue with TMA stores (bias, GELU, preact, gate, DropPath)."""
    from fiber_b200 import kernels as K
    a = _mk((m, k), cuda_dev, 11)
    b = _mk((n, k), cuda_dev, 12, k ** -0.5)
    bias = torch.randn(n, device=cuda_dev)
    scale = torch.tensor([0.75], device=cuda_dev)
    rps = 50
    row_scale = torch.rand((m + rps - 1) // rps, device=cuda_dev) + 0.5
    pre = torch.zeros((m, n), device=cuda_dev, dtype=torch.bfloat16)
    out = K.gemm(a, b, bias=bias, preact=pre, scale=scale, row_scale=row_scale, rows_per_scale=rps, act=K.ACT_GELU)
    h = a.float() @ b.float().t() + bias
    ref = torch.nn.functional.gelu(h) * 0.75 * row_scale.repeat_interleave(rps)[:m, None]
    torch.testing.assert_close(pre.float(), h, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(out.float(), ref, rtol=RTOL, atol=ATOL)
    out2 = K.gemm(a, b, bias=bias)
    torch.testing.assert_close(out2.float(), h, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("combo", ["plain", "bias", "bias_gelu_pre", "gelu_grad", "res", "bias_res", "bias_scale_res",
                                   "scale", "bias_pre_scale_res"])
def test_gemm_specialised_epilogues(cuda_dev, combo):
    """Tile-aligned shapes take the compile-time specialised epilogue chunks (gemm_sm100.cu epi_chunk_full)."""
    from fiber_b200 import kernels as K
    m, n, k = 512, 768, 192
    a = _mk((m, k), cuda_dev, 21)
    b = _mk((n, k), cuda_dev, 22, k ** -0.5)
    bias = torch.randn(n, device=cuda_dev)
    res = _mk((m, n), cuda_dev, 23)
    aux = _mk((m, n), cuda_dev, 24)
    alpha = torch.tensor([0.6], device=cuda_dev)
    rs = torch.rand(m // 64, device=cuda_dev) + 0.5
    acc = a.float() @ b.float().t()
    rsf = rs.repeat_interleave(64)[:, None]
    pre = torch.zeros((m, n), device=cuda_dev, dtype=torch.bfloat16)
    if combo == "plain":
        out, ref = K.gemm(a, b), acc
    elif combo == "bias":
        out, ref = K.gemm(a, b, bias=bias), acc + bias
    elif combo == "bias_gelu_pre":
        out = K.gemm(a, b, bias=bias, act=K.ACT_GELU, preact=pre)
        ref = torch.nn.functional.gelu(acc + bias)
        torch.testing.assert_close(pre.float(), acc + bias, rtol=RTOL, atol=ATOL)
    elif combo == "gelu_grad":
        out = K.gemm(a, b, aux=aux, act=K.ACT_GELU_GRAD)
        x = aux.float().requires_grad_(True)
        torch.nn.functional.gelu(x).sum().backward()
        ref = acc * x.grad
    elif combo == "res":
        out, ref = K.gemm(a, b, residual=res), acc + res.float()
    elif combo == "bias_res":
        out, ref = K.gemm(a, b, bias=bias, residual=res), acc + bias + res.float()
    elif combo == "bias_scale_res":
        out = K.gemm(a, b, bias=bias, residual=res, row_scale=rs, rows_per_scale=64)
        ref = (acc + bias) * rsf + res.float()
    elif combo == "scale":
        out, ref = K.gemm(a, b, scale=alpha), acc * 0.6
    else:
        out = K.gemm(a, b, bias=bias, preact=pre, scale=alpha, residual=res, row_scale=rs, rows_per_scale=64)
        ref = (acc + bias) * 0.6 * rsf + res.float()
        torch.testing.assert_close(pre.float(), acc + bias, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(out.float(), ref, rtol=RTOL, atol=ATOL)


def test_vocab_decoder_padded(cuda_dev):
    """ops.VocabDecoderFn: vocabulary size that is not a multiple of 8 (MLM head), forward + all gradients."""
    from fiber_b200 import ops
    torch.manual_seed(0)
    V, Kd = 1003, 96
    x = (torch.randn(3, 40, Kd, device=cuda_dev)).requires_grad_(True)
    w = (torch.randn(V, Kd, device=cuda_dev) * Kd ** -0.5).requires_grad_(True)
    b = torch.randn(V, device=cuda_dev).requires_grad_(True)
    y = ops.VocabDecoderFn.apply(x, w, b)
    assert y.shape == (3, 40, V)
    labels = torch.randint(0, V, (120,), device=cuda_dev)
    labels[::3] = -100
    loss = torch.nn.functional.cross_entropy(y.view(-1, V).float(), labels, ignore_index=-100)
    loss.backward()
    xr, wr, br = (t.detach().clone().requires_grad_(True) for t in (x, w, b))
    yr = xr.to(torch.bfloat16).float() @ wr.to(torch.bfloat16).float().t() + br
    lr = torch.nn.functional.cross_entropy(yr.view(-1, V), labels, ignore_index=-100)
    lr.backward()
    torch.testing.assert_close(y, yr, rtol=RTOL, atol=ATOL)
    assert abs(loss.item() - lr.item()) < 2e-3
    for name, a, r in (("dx", x.grad, xr.grad), ("dw", w.grad, wr.grad), ("db", b.grad, br.grad)):
        err = (a - r).abs().max().item()
        assert err <= 2e-2 * max(r.abs().max().item(), 1e-6), "%s: %g vs %g" % (name, err, r.abs().max().item())


# Opt-in epilogues (kernel template parameter EPI = 1): not yet validated on hardware, run with
# FIBER_B200_EXPERIMENTAL=1 (tools/gpu_round2a.sh).
@pytest.mark.parametrize("m,n,k,with_bias", [(128, 128, 64, True), (256, 512, 128, True), (1024, 2048, 512, True),
                                             (2560, 3072, 768, True), (147456 // 8, 512, 128, False),
                                             (384, 160, 96, True)])
def test_gemm_gelu_cache_epilogues(cuda_dev, m, n, k, with_bias):
    from fiber_b200 import kernels as K
    a = _mk((m, k), cuda_dev, 11)
    b = _mk((n, k), cuda_dev, 12, k ** -0.5)
    bias = torch.randn(n, device=cuda_dev) if with_bias else None
    # act 3: out = GELU(v), second output = GELU'(v); GELU must be bit-identical to the default two-pass epilogue
    gp = torch.empty((m, n), device=cuda_dev, dtype=torch.bfloat16)
    out = K.gemm(a, b, bias=bias, act=K.ACT_GELU_CACHE, preact=gp)
    h = torch.empty((m, n), device=cuda_dev, dtype=torch.bfloat16)
    out_ref = K.gemm(a, b, bias=bias, act=K.ACT_GELU, preact=h)
    assert torch.equal(out, out_ref)
    # act 5: both outputs of the default two-pass epilogue, bit for bit, from one pass
    h1 = torch.empty_like(h)
    out1 = K.gemm(a, b, bias=bias, act=K.ACT_GELU_ONEPASS, preact=h1)
    assert torch.equal(out1, out_ref) and torch.equal(h1, h)
    v = (a.float() @ b.float().t() + (bias if with_bias else 0.0)).requires_grad_(True)
    torch.nn.functional.gelu(v).sum().backward()
    torch.testing.assert_close(gp.float(), v.grad, rtol=1e-2, atol=1e-2)
    # act 4: dgrad through the GELU = (dz W) * GELU'(h); compare with the default act-2 path on the same operands
    dz = _mk((m, k), cuda_dev, 13)
    w_t = _mk((n, k), cuda_dev, 14, k ** -0.5)
    dh = K.gemm(dz, w_t, aux=gp, act=K.ACT_MUL_AUX)
    dh_ref = K.gemm(dz, w_t, aux=h, act=K.ACT_GELU_GRAD)
    assert torch.equal(K.gemm(dz, w_t, aux=h, act=K.ACT_GELU_GRAD_PF), dh_ref)  # act 7 == act 2, bit for bit
    torch.testing.assert_close(dh.float(), (dz.float() @ w_t.float().t()) * gp.float(), rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(dh.float(), dh_ref.float(), rtol=3e-2, atol=3e-2)


def test_gemm_gelu_cache_rejects_unsupported(cuda_dev):
    from fiber_b200 import kernels as K
    a = _mk((100, 64), cuda_dev, 1)
    b = _mk((128, 64), cuda_dev, 2)
    with pytest.raises(RuntimeError):  # M % 128 != 0
        K.gemm(a, b, act=K.ACT_GELU_CACHE, preact=torch.empty((100, 128), device=cuda_dev, dtype=torch.bfloat16))


@pytest.mark.parametrize("m,n,k", [(256, 512, 512), (1152, 512, 2048), (9216, 128, 128), (2560, 768, 3072), (384, 160, 96)])
@pytest.mark.parametrize("with_bias,with_scale,with_rows", [(True, False, False), (True, True, True), (False, False, True),
                                                            (False, True, False)])
def test_gemm_residual_prefetch_is_bit_identical(cuda_dev, m, n, k, with_bias, with_scale, with_rows):
    """act 6 (ACT_RES_PF, the default at K <= 512) == the plain residual epilogue, bit for bit."""
    from fiber_b200 import kernels as K
    a = _mk((m, k), cuda_dev, 21)
    b = _mk((n, k), cuda_dev, 22, k ** -0.5)
    res = _mk((m, n), cuda_dev, 23)
    kw = dict(residual=res)
    if with_bias:
        kw["bias"] = torch.randn(n, device=cuda_dev)
    if with_scale:
        kw["scale"] = torch.tensor([0.7], device=cuda_dev)
    if with_rows:
        rps = 128
        kw["row_scale"] = torch.rand(m // rps, device=cuda_dev) + 0.5
        kw["rows_per_scale"] = rps
    old = K.RES_PREFETCH
    try:
        K.set_res_prefetch(0)
        ref = K.gemm(a, b, **kw)
        K.set_res_prefetch(2)
        out = K.gemm(a, b, **kw)
    finally:
        K.set_res_prefetch(old)
    assert torch.equal(out, ref)


@pytest.mark.parametrize("count", [0, 1, 100, 128, 300, 1000])
def test_gemm_device_row_count(cuda_dev, count):
    """fiber_gemm_args.row_count: a device scalar says how many leading activation rows carry work.  K-major launches leave
    the output row tiles past it untouched; MN-major (wgrad) launches reduce over the leading rows only (whole 64-row
    k-blocks: the caller zero-fills up to the block edge, as the fused MLM cross-entropy backward does)."""
    from fiber_b200 import kernels as K
    m, n, k = 1000, 384, 256
    a = _mk((m, k), cuda_dev, 11)
    b = _mk((n, k), cuda_dev, 12, k ** -0.5)
    cnt = torch.tensor([count], device=cuda_dev, dtype=torch.int32)
    out = torch.full((m, n), 7.0, device=cuda_dev, dtype=torch.bfloat16)
    K.gemm(a, b, out=out, row_count=cnt)
    edge = min(m, (count + 127) // 128 * 128)
    ref = a.float() @ b.float().t()
    torch.testing.assert_close(out[:edge].float(), ref[:edge], rtol=RTOL, atol=ATOL)
    assert bool((out[edge:] == 7.0).all())
    # wgrad: dW = dY^T X over the first ceil(count / 64) * 64 rows
    dy = _mk((m, n), cuda_dev, 13)
    x = _mk((m, k), cuda_dev, 14)
    kedge = min(m, (count + 63) // 64 * 64)
    db = torch.zeros(n, device=cuda_dev)
    dw = K.gemm(dy, x, mn_major=True, accumulate=True, colsum=db, row_count=cnt)
    ref_w = dy[:kedge].float().t() @ x[:kedge].float()
    torch.testing.assert_close(dw, ref_w, rtol=2e-2, atol=2e-2 * max(1.0, kedge ** 0.5))
    torch.testing.assert_close(db, dy[:kedge].float().sum(0), rtol=2e-2, atol=2e-2 * max(1.0, kedge ** 0.5))


PAIR_SHAPES = [(256, 256, 256), (512, 384, 512), (2560, 768, 768), (2560, 3072, 768), (36864, 1536, 512),
               (36864, 512, 2048), (9216, 4096, 1024), (147456, 512, 512), (768, 264, 320)]


@pytest.mark.parametrize("m,n,k", PAIR_SHAPES)
def test_gemm_cta_pairs(cuda_dev, m, n, k):
    """fiber_set_option("gemm_cta2", 7): the same GEMMs as CTA pairs (2-CTA clusters, tcgen05.mma.cta_group::2, 256 x 256 pair
    tiles, half of the B tile per CTA) — plain, bias + residual + scale, and the two-box GELU epilogues — against the
    single-CTA kernel (bit for bit: same per-element accumulation order) and the fp32 reference."""
    from fiber_b200 import kernels as K, lib
    a = _mk((m, k), cuda_dev, 21)
    b = _mk((n, k), cuda_dev, 22, k ** -0.5)
    bias = torch.randn(n, device=cuda_dev)
    res = _mk((m, n), cuda_dev, 23)
    aux = _mk((m, n), cuda_dev, 24)
    sc = torch.tensor([0.75], device=cuda_dev)

    def run_all():
        outs = [K.gemm(a, b), K.gemm(a, b, bias=bias, residual=res, scale=sc)]
        if n % 32 == 0 and m % 128 == 0:
            pre = torch.empty((m, n), device=cuda_dev, dtype=torch.bfloat16)
            outs.append(K.gemm(a, b, bias=bias, preact=pre, act=K.ACT_GELU_CACHE))
            outs.append(pre)
            outs.append(K.gemm(a, b, aux=aux, act=K.ACT_MUL_AUX))
            outs.append(K.gemm(a, b, bias=bias, residual=res, act=K.ACT_RES_PF))
        return outs

    lib.set_option("gemm_cta2", 0)
    base = run_all()
    try:
        lib.set_option("gemm_cta2", 7)  # bit 2: also below the default K >= 1024 threshold
        n0 = lib.get_option("gemm_cta2_launches")
        pair = run_all()
        torch.cuda.synchronize()
        assert lib.get_option("gemm_cta2_launches") - n0 == (len(pair) - 1 if len(pair) > 2 else 2)
    finally:
        lib.set_option("gemm_cta2", -1)
    ref = a.float() @ b.float().t()
    torch.testing.assert_close(pair[0].float(), ref, rtol=RTOL, atol=ATOL)
    for x, y in zip(base, pair):
        assert torch.equal(x, y)


def test_programmatic_dependent_launch_option(cuda_dev):
    """fiber_set_option("pdl", 1): every launch carries the programmatic-stream-serialization attribute and every kernel
    waits (griddepcontrol.wait) before its first global access — a chain of dependent launches (GEMM -> LayerNorm ->
    GEMM with residual -> wgrad) must give the results of plain stream order."""
    from fiber_b200 import kernels as K, lib
    a = _mk((1024, 512), cuda_dev, 31)
    w1 = _mk((512, 512), cuda_dev, 32, 512 ** -0.5)
    w2 = _mk((384, 512), cuda_dev, 33, 512 ** -0.5)
    g, b = torch.ones(512, device=cuda_dev), torch.zeros(512, device=cuda_dev)

    def chain():
        h = K.gemm(a, w1)
        y = K.layernorm_fwd(h, g, b, 1e-5)[0]
        o = K.gemm(y, w2, residual=_mk((1024, 384), cuda_dev, 34))
        dw = K.gemm(o, y, mn_major=True, accumulate=True)
        return o, dw

    lib.set_option("pdl", 0)
    ref = chain()
    try:
        lib.set_option("pdl", 1)
        assert lib.get_option("pdl") == 1
        for _ in range(3):
            out = chain()
        torch.cuda.synchronize()
    finally:
        lib.set_option("pdl", -1)
    assert torch.equal(out[0], ref[0])
    torch.testing.assert_close(out[1], ref[1], rtol=1e-4, atol=1e-2)  # split-K atomics: summation order varies


@pytest.mark.parametrize("rows,m,n", [(4096, 512, 384), (36864, 2048, 512), (9000, 256, 1024), (147456, 512, 2048)])
def test_gemm_cta_pairs_wgrad(cuda_dev, rows, m, n):
    """fiber_set_option("gemm_cta2", 8 | 4): wgrad launches (MN-major operands, split-K atomics, fused column sums) as CTA
    pairs against the single-CTA kernel and the fp32 reference."""
    from fiber_b200 import kernels as K, lib
    dy = _mk((rows, m), cuda_dev, 41, 0.5)
    x = _mk((rows, n), cuda_dev, 42, 0.5)

    def run():
        db = torch.zeros(m, device=cuda_dev)
        dw = K.gemm(dy, x, mn_major=True, accumulate=True, colsum=db)
        return dw, db

    lib.set_option("gemm_cta2", 0)
    base = run()
    try:
        lib.set_option("gemm_cta2", 12)
        n0 = lib.get_option("gemm_cta2_launches")
        pair = run()
        torch.cuda.synchronize()
        assert lib.get_option("gemm_cta2_launches") - n0 == 1
    finally:
        lib.set_option("gemm_cta2", -1)
    ref_w = dy.float().t() @ x.float()
    ref_b = dy.float().sum(0)
    tol = 2e-2 * max(1.0, rows ** 0.5 * 0.25)
    torch.testing.assert_close(pair[0], ref_w, rtol=2e-2, atol=tol)
    torch.testing.assert_close(pair[1], ref_b, rtol=2e-2, atol=tol)
    torch.testing.assert_close(pair[0], base[0], rtol=1e-3, atol=tol * 0.1)  # split-K atomics: order varies
    torch.testing.assert_close(pair[1], base[1], rtol=1e-3, atol=tol * 0.1)
