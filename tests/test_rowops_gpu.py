"""Row-wise kernels vs torch fp32 on the same bf16 inputs (GPU)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import fiber_oracle as O

pytestmark = pytest.mark.gpu


def _rand(shape, dev, seed, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev).to(dtype)


def _close(a, b, tol, name):
    a, b = a.float(), b.float()
    err = (a - b).abs().max().item()
    ref = b.abs().max().item()
    assert err <= tol * max(ref, 1e-3), "%s: max err %g vs ref max %g" % (name, err, ref)


@pytest.mark.parametrize("rows,C,add", [(1000, 128, False), (333, 256, True), (77, 512, False), (2560, 768, True),
                                        (500, 1024, False), (144, 2048, False), (9, 8, False),
                                        (30001, 256, False), (40999, 512, False), (2560, 768, False),
                                        (70003, 128, False)])
def test_layernorm_fwd_bwd(cuda_dev, rows, C, add):
    from fiber_b200 import kernels as K
    x = _rand((rows, C), cuda_dev, 1)
    x2 = _rand((rows, C), cuda_dev, 2) if add else None
    g = _rand((C,), cuda_dev, 3, 0.1, torch.float32) + 1.0
    b = _rand((C,), cuda_dev, 4, 0.1, torch.float32)
    dy = _rand((rows, C), cuda_dev, 5)
    dres = _rand((rows, C), cuda_dev, 6)
    y, mean, rstd, s = K.layernorm_fwd(x, g, b, 1e-5, add=x2, want_sum=add)
    xf = (x.float() + (x2.float() if add else 0)).requires_grad_(True)
    gf, bf = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.layer_norm(xf, (C,), gf, bf, 1e-5)
    _close(y, ref, 1e-2, "y")
    if add:
        _close(s, xf, 1e-2, "sum")
    ref.backward(dy.float())
    dg, db = torch.zeros_like(g), torch.zeros_like(b)
    dx = K.layernorm_bwd(dy, x, mean, rstd, g, add=x2, dres=dres, dgamma=dg, dbeta=db)
    _close(dx, xf.grad + dres.float(), 1.5e-2, "dx")
    _close(dg, gf.grad, 1e-2, "dgamma")
    _close(db, bf.grad, 1e-2, "dbeta")


@pytest.mark.parametrize("B,H,C", [(3, 12, 64), (2, 24, 128), (3, 12, 256), (2, 8, 512)])
def test_layernorm_patch_merging(cuda_dev, B, H, C):
    from fiber_b200 import kernels as K
    x = _rand((B * H * H, C), cuda_dev, 1)
    g = _rand((4 * C,), cuda_dev, 3, 0.1, torch.float32) + 1.0
    b = _rand((4 * C,), cuda_dev, 4, 0.1, torch.float32)
    y, mean, rstd, _ = K.layernorm_fwd(x, g, b, 1e-5, merge=(B, H, H))
    xf = x.float().view(B, H * H, C).requires_grad_(True)
    gf, bf = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    xm = xf.view(B, H // 2, 2, H // 2, 2, C).permute(0, 1, 3, 4, 2, 5).reshape(B * (H // 2) ** 2, 4 * C)
    ref = F.layer_norm(xm, (4 * C,), gf, bf, 1e-5)
    _close(y, ref, 1e-2, "y")
    dy = _rand(tuple(ref.shape), cuda_dev, 5)
    ref.backward(dy.float())
    dg, db = torch.zeros_like(g), torch.zeros_like(b)
    dx = K.layernorm_bwd(dy, x, mean, rstd, g, merge=(B, H, H), dgamma=dg, dbeta=db)
    _close(dx.view(B, H * H, C), xf.grad, 1.5e-2, "dx")
    _close(dg, gf.grad, 1e-2, "dgamma")
    _close(db, bf.grad, 1e-2, "dbeta")


def test_colsum_dot_scale_dropout_cast(cuda_dev):
    from fiber_b200 import kernels as K
    x = _rand((5000, 384), cuda_dev, 1)
    rs = torch.rand(50, device=cuda_dev) + 0.5
    sc = torch.tensor([0.25], device=cuda_dev)
    out = K.colsum(x, scale=sc, row_scale=rs, rows_per_scale=100)
    ref = (x.float() * rs.repeat_interleave(100)[:, None]).sum(0) * 0.25
    _close(out, ref, 2e-3, "colsum")
    out = K.colsum(x[:, 128:256])
    _close(out, x[:, 128:256].float().sum(0), 2e-3, "colsum view")
    y = _rand((5000, 384), cuda_dev, 2)
    d = K.dot(x, y)
    _close(d, (x.float() * y.float()).sum().view(1), 2e-3, "dot")
    z = K.scale_rows(x, rs, 100)
    _close(z, x.float() * rs.repeat_interleave(100)[:, None], 1e-2, "scale_rows")
    dr = K.dropout(x, 0.1, 7)
    dr2 = K.dropout(x, 0.1, 7)
    assert torch.equal(dr, dr2)
    kept = (dr != 0).float().mean().item()
    assert abs(kept - 0.9) < 0.01
    m = dr != 0
    _close(dr[m], x.float()[m] / 0.9, 1e-2, "dropout scale")
    f = torch.randn(1001, device=cuda_dev)
    assert torch.equal(K.cast_bf16(f), f.to(torch.bfloat16))


def test_cast_transpose_and_patch_gather(cuda_dev):
    from fiber_b200 import kernels as K
    w = torch.randn(100, 48, device=cuda_dev)
    wo = torch.zeros(100, 64, device=cuda_dev, dtype=torch.bfloat16)
    wt = torch.zeros(64, 104, device=cuda_dev, dtype=torch.bfloat16)
    K.cast_transpose(w, wo, wt)
    assert torch.equal(wo[:, :48], w.to(torch.bfloat16)) and float(wo[:, 48:].abs().sum()) == 0
    assert torch.equal(wt[:48, :100], w.t().to(torch.bfloat16))
    img = torch.randn(2, 3, 32, 32, device=cuda_dev)
    p = K.patch_gather(img)
    ref = img.view(2, 3, 8, 4, 8, 4).permute(0, 2, 4, 1, 3, 5).reshape(2 * 64, 48)
    assert torch.equal(p[:, :48], ref.to(torch.bfloat16)) and float(p[:, 48:].abs().sum()) == 0


def test_embeddings_gather_scatter(cuda_dev):
    from fiber_b200 import kernels as K
    V, C, L = 200, 768, 40
    word = torch.randn(V, C, device=cuda_dev) * 0.1
    pos = torch.randn(L + 2, C, device=cuda_dev) * 0.1
    typ = torch.randn(1, C, device=cuda_dev) * 0.1
    ids = torch.randint(3, V, (5, L), device=cuda_dev)
    ids[1, 20:] = 1
    ids[3, 33:] = 1
    out = K.embed_gather(ids, word, pos, typ)
    pid = O.roberta_position_ids(ids)
    ref = word[ids] + pos[pid] + typ[0]
    _close(out.view(5, L, C), ref, 1e-2, "embed")
    d = _rand((5 * L, C), cuda_dev, 3)
    dw, dp = torch.zeros_like(word), torch.zeros_like(pos)
    K.embed_scatter(ids, d, dw, dp)
    w2, p2 = word.clone().requires_grad_(True), pos.clone().requires_grad_(True)
    (F.embedding(ids, w2, padding_idx=1) + F.embedding(pid, p2, padding_idx=1)).backward(d.float().view(5, L, C))
    _close(dw, w2.grad, 1e-3, "dword")
    _close(dp, p2.grad, 1e-3, "dpos")


@pytest.mark.parametrize("rows,C", [(3000, 512), (1200, 128), (720, 1024)])
def test_layernorm_bwd_scaled_second_output(cuda_dev, rows, C):
    """layernorm_bwd(row_scale=...) also returns dx * s[row // rps] (the DropPath scale of the consumer branch)."""
    from fiber_b200 import kernels as K
    x = _rand((rows, C), cuda_dev, 1)
    g = _rand((C,), cuda_dev, 3, 0.1, torch.float32) + 1.0
    b = _rand((C,), cuda_dev, 4, 0.1, torch.float32)
    dy, dres = _rand((rows, C), cuda_dev, 5), _rand((rows, C), cuda_dev, 6)
    _, mean, rstd, _ = K.layernorm_fwd(x, g, b, 1e-5)
    rps = 120
    s = torch.rand(rows // rps, device=cuda_dev) + 0.5
    s[1] = 0.0
    dg, db = torch.zeros_like(g), torch.zeros_like(b)
    dx0 = K.layernorm_bwd(dy, x, mean, rstd, g, dres=dres)
    dx, dxs = K.layernorm_bwd(dy, x, mean, rstd, g, dres=dres, dgamma=dg, dbeta=db, row_scale=s, rows_per_scale=rps)
    assert torch.equal(dx, dx0)
    _close(dxs, dx0.float() * s.repeat_interleave(rps)[:, None], 1e-2, "dx_scaled")


def test_fused_adamw_matches_hf_update(cuda_dev):
    """fiber_adamw_multi against a plain restatement of transformers 4.6 AdamW.step (fiber_utils.py:247 uses it with
    betas (0.9, 0.98), eps 1e-8): three steps over tensors of awkward sizes in two groups, one with weight decay."""
    from fiber_b200.optim import FusedAdamW
    g = torch.Generator().manual_seed(3)
    sizes = [(5,), (1023,), (257, 129), (65536 + 3,), (768, 768)]
    ps = [torch.nn.Parameter(torch.randn(s, generator=g).to(cuda_dev)) for s in sizes]
    ref = [p.detach().double().clone() for p in ps]
    m = [torch.zeros_like(r) for r in ref]
    v = [torch.zeros_like(r) for r in ref]
    groups = [{"params": ps[:3], "weight_decay": 0.01, "lr": 1e-3}, {"params": ps[3:], "weight_decay": 0.0, "lr": 5e-3}]
    opt = FusedAdamW(groups, lr=1e-3, betas=(0.9, 0.98), eps=1e-8)
    b1, b2, eps = 0.9, 0.98, 1e-8
    for step in range(1, 4):
        for i, p in enumerate(ps):
            p.grad = torch.randn(p.shape, generator=g).to(cuda_dev) * (0.1 * step)
        opt.step()
        for i, p in enumerate(ps):
            lr, wd = (1e-3, 0.01) if i < 3 else (5e-3, 0.0)
            gr = p.grad.double()
            m[i] = b1 * m[i] + (1 - b1) * gr
            v[i] = b2 * v[i] + (1 - b2) * gr * gr
            step_size = lr * (1 - b2 ** step) ** 0.5 / (1 - b1 ** step)
            ref[i] = ref[i] - step_size * m[i] / (v[i].sqrt() + eps)
            ref[i] = ref[i] - lr * wd * ref[i]
    for i, p in enumerate(ps):
        torch.testing.assert_close(p.detach().double(), ref[i], rtol=2e-6, atol=2e-7)
        torch.testing.assert_close(opt.state[p]["exp_avg_sq"].double(), v[i], rtol=1e-5, atol=1e-12)
    # a parameter without gradient is skipped; LR schedulers rewrite group["lr"] and the next step sees it
    ps[0].grad = None
    before = ps[0].detach().clone()
    opt.param_groups[0]["lr"] = 0.0
    p1 = ps[1].detach().clone()
    opt.step()
    assert torch.equal(ps[0].detach(), before) and torch.equal(ps[1].detach(), p1 * (1 - 0.0 * 0.01))


@pytest.mark.gpu
@pytest.mark.parametrize("rows,V,frac", [(2560, 50265, 0.15), (80, 50265, 0.5), (384, 1000, 1.0), (200, 4096, 0.0),
                                         (2560, 50265, 1.0)])
def test_fused_mlm_decoder_cross_entropy(cuda_dev, rows, V, frac):
    """ops.MlmDecoderCEFn (fiber_mlm_ce_fwd / _bwd: decoder GEMM with cross-entropy epilogues, labelled rows first, device
    row count) against F.cross_entropy on fp32 logits of the same bf16-rounded operands (heads.py:40-43 +
    objectives.py:19-26): loss, arg-max at the labelled rows, d(hidden), d(weight), d(bias).  frac = share of labelled rows
    (0: the all-ignored batch gives nan like F.cross_entropy; 1: no row is skipped)."""
    from fiber_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(rows * 7 + V)
    Kd = 768
    h = (torch.randn(rows, Kd, generator=g) * 1.5).to(cuda_dev).to(torch.bfloat16)
    w = (torch.randn(V, Kd, generator=g) * 0.05).to(cuda_dev)
    b = (torch.randn(V, generator=g) * 0.5).to(cuda_dev)
    labels = torch.randint(0, V, (rows,), generator=g)
    labels[torch.rand(rows, generator=g) >= frac] = -100
    if frac > 0 and frac < 1:
        labels[3] = V - 1  # the last real column (next to the padded ones)
    labels = labels.to(cuda_dev)
    hp = h.clone().requires_grad_(True)
    wp, bp = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    loss, pred = ops.MlmDecoderCEFn.apply(hp, wp, bp, labels)
    # fp32 reference on the operands the kernel sees (bf16 hidden rows, bf16 weight copy, fp32 bias)
    hr = h.float().requires_grad_(True)
    wr = w.to(torch.bfloat16).float().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    logits = hr @ wr.t() + br
    ref = F.cross_entropy(logits, labels, ignore_index=-100)
    keep = labels != -100
    if frac == 0.0:
        assert torch.isnan(loss) and torch.isnan(ref)
        return
    assert abs(float(loss) - float(ref)) <= 2e-5 * abs(float(ref)), (float(loss), float(ref))
    ref_pred = logits.argmax(-1)
    agree = (pred[keep] == ref_pred[keep])
    if not bool(agree.all()):  # a disagreement must be a numerical tie of the two top logits
        bad = keep.nonzero().flatten()[~agree]
        top = logits[bad].gather(1, torch.stack([pred[bad], ref_pred[bad]], 1))
        assert (top[:, 0] - top[:, 1]).abs().max().item() < 1e-4
    assert int(pred[~keep].abs().sum()) == 0
    (loss * 3.0).backward()
    (ref * 3.0).backward()

    def l2rel(a, r):
        return ((a.float() - r.float()).norm() / r.float().norm().clamp_min(1e-30)).item()
    # d(logits) leaves the epilogue as bf16 (2^-9 relative per element), accumulated in fp32 by the dgrad / wgrad GEMMs
    assert l2rel(hp.grad, hr.grad) < 6e-3, l2rel(hp.grad, hr.grad)
    assert l2rel(wp.grad, wr.grad) < 6e-3, l2rel(wp.grad, wr.grad)
    assert l2rel(bp.grad, br.grad) < 6e-3, l2rel(bp.grad, br.grad)
    assert hp.grad[~keep].abs().max().item() == 0.0 if bool((~keep).any()) else True


@pytest.mark.parametrize("B,H,W,Hp,Wp,C,scaled", [(2, 50, 84, 60, 84, 128, True), (3, 7, 10, 12, 12, 1024, True),
                                                   (2, 24, 24, 24, 24, 256, False), (1, 25, 42, 36, 48, 512, True)])
def test_grid_copy_crop_scale_add_fwd_bwd(cuda_dev, B, H, W, Hp, Wp, C, scaled):
    """fiber_grid_copy through ops.CropScaleAddFn against the torch expression of the fine-grained block
    (fusion_swin_transformer_v2.py:336-343: crop, DropPath scale, residual) — same bf16 roundings, so bit for bit."""
    from fiber_b200 import kernels as K
    from fiber_b200 import ops
    x = _rand((B, H * W, C), cuda_dev, 1).requires_grad_(True)
    z = _rand((B * Hp * Wp, C), cuda_dev, 2).requires_grad_(True)
    s = (torch.tensor([0.0, 1.25, 1.25][:B], device=cuda_dev) if scaled else None)
    dout = _rand((B, H * W, C), cuda_dev, 3)
    out = ops.CropScaleAddFn.apply(x, z, s, (H, W), (Hp, Wp))
    out.backward(dout)
    xr, zr = x.detach().clone().requires_grad_(True), z.detach().clone().requires_grad_(True)
    zc = zr.view(B, Hp, Wp, C)[:, :H, :W].reshape(B, H * W, C)
    ref = (xr.float() + (zc.float() * s.view(B, 1, 1) if scaled else zc.float())).to(torch.bfloat16)
    ref.backward(dout)
    assert torch.equal(out, ref)
    assert torch.equal(x.grad, xr.grad)
    assert torch.equal(z.grad, zr.grad.to(torch.bfloat16))
    # zero padding alone (the F.pad of :316-321)
    src = _rand((B, H, W, C), cuda_dev, 4)
    assert torch.equal(K.grid_copy(src, (Hp, Wp)), F.pad(src, (0, 0, 0, Wp - W, 0, Hp - H)))
