"""Attention kernels (window + plain modes) vs the oracle's fp32 formulation on bf16 inputs (GPU)."""
import math
import os

import pytest
import torch

from oracle import fiber_oracle as O

pytestmark = pytest.mark.gpu


def _rand(shape, dev, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev).to(torch.bfloat16)


def _close(a, b, tol, name):
    a, b = a.float(), b.float()
    err = (a - b).abs().max().item()
    ref = b.abs().max().item()
    assert err <= tol * max(ref, 1e-3), "%s: max err %g vs ref max %g" % (name, err, ref)


WINDOW_CASES = [(2, 24, 12, 0, 2), (2, 24, 12, 6, 2), (3, 14, 7, 3, 4), (1, 12, 12, 0, 3),
                (2, 96, 12, 6, 4), (1, 36, 18, 9, 2), (2, 18, 18, 0, 3),
                (8, 48, 12, 6, 8), (4, 28, 7, 0, 4), (2, 48, 12, 5, 2)]


@pytest.mark.parametrize("B,H,ws,shift,nh", WINDOW_CASES)
def test_window_attention_fwd_bwd(cuda_dev, B, H, ws, shift, nh):
    _window_case(cuda_dev, B, H, ws, shift, nh)


# The tcgen05 generation (csrc/window_attn_tc.cu) is the default for 12x12 windows (option "winattn_tc" = 3).  Cases it
# does not cover (ws != 12, odd shifts) must fall through to the mma.sync kernels with the option set, so every case
# runs; the 12x12 ones must launch it.  Mode 0 keeps the mma.sync generation under test for every geometry.
TC_CASES = WINDOW_CASES + [(64, 12, 12, 0, 32), (5, 96, 12, 6, 4), (3, 24, 12, 6, 16), (1, 48, 12, 0, 8)]


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 5, 7, 11, 15],
                         ids=["mmasync", "tcfwd", "tcbwd", "tcboth", "tqfwd", "tqfwd_tcbwd", "tcfwd_tqbwd", "tqboth"])
@pytest.mark.parametrize("B,H,ws,shift,nh", TC_CASES)
def test_window_attention_tcgen05(cuda_dev, B, H, ws, shift, nh, mode):
    from fiber_b200 import lib
    before = lib.get_option("winattn_tc_launches")
    lib.set_option("winattn_tc", mode)
    try:
        _window_case(cuda_dev, B, H, ws, shift, nh, diag=True)
        torch.cuda.synchronize()
    finally:
        lib.set_option("winattn_tc", -1)  # back to the default
    launched = lib.get_option("winattn_tc_launches") - before
    covered = ws == 12 and shift in (0, 6)
    assert launched == ((mode & 1) + ((mode >> 1) & 1) if covered else 0)


def _diag(name, got, ref, src, N, hd):
    """Where does a window-attention tensor differ?  Error maxima by window-token class (the tcgen05 kernels treat
    tokens 0..127 and 128..143 of a window on different code paths), head-dim half and window position."""
    got, ref = got.float(), ref.float()
    B, T, C = ref.shape
    nW = src.shape[0]
    err = (got - ref).abs()[:, src.reshape(-1)].reshape(B, nW, N, C // hd, hd)  # window order
    scale = max(ref.abs().max().item(), 1e-3)
    tok = err.amax(dim=(0, 1, 3, 4)) / scale
    lines = ["%s: rel-to-max error by class (ref max %.3g)" % (name, scale)]
    lines.append("  tokens 0..127: %.3g   tokens 128..%d: %.3g" % (tok[:128].max().item() if N > 128 else tok.max().item(),
                                                               N - 1, tok[128:].max().item() if N > 128 else 0.0))
    lines.append("  head-dim 0..15: %.3g   16..31: %.3g" % ((err[..., :16].max() / scale).item(), (err[..., 16:].max() / scale).item()))
    win = err.amax(dim=(0, 2, 3, 4)) / scale
    lines.append("  first window: %.3g   last window: %.3g   worst window %d: %.3g" % (
        win[0].item(), win[-1].item(), int(win.argmax()), win.max().item()))
    lines.append("  worst tokens: %s" % [(int(i), round(float(tok[i]), 4)) for i in tok.topk(min(6, N)).indices])
    print("\n".join(lines))


def _window_case(cuda_dev, B, H, ws, shift, nh, diag=False):
    from fiber_b200 import kernels as K
    hd = 32
    C = nh * hd
    N = ws * ws
    T = H * H
    qkv = _rand((B * T, 3 * C), cuda_dev, 1)
    table = (torch.randn((2 * ws - 1) ** 2, nh, generator=torch.Generator().manual_seed(2)) * 0.5).to(cuda_dev)
    d_o = _rand((B * T, C), cuda_dev, 3)
    scale = hd ** -0.5
    win = (B, H, H, ws, shift)
    o, lse = K.attn_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], nh, hd, scale, window=win, bias_table=table)

    # reference: the oracle's index maps + explicit softmax in fp32 on the same bf16 inputs
    x = qkv.float().view(B, T, 3 * C).requires_grad_(True)
    tb = table.clone().requires_grad_(True)
    src = O.window_token_source(H, H, ws, shift).to(cuda_dev)
    nW = src.shape[0]
    xw = x[:, src.reshape(-1)].reshape(B * nW, N, 3, nh, hd).permute(2, 0, 3, 1, 4)
    s = (xw[0] * scale) @ xw[1].transpose(-2, -1)
    s = s + tb[O.relative_position_index(ws).to(cuda_dev).reshape(-1)].view(N, N, nh).permute(2, 0, 1)[None]
    m = O.shift_attn_mask(H, H, ws, shift)
    if m is not None:
        s = (s.view(B, nW, nh, N, N) + m.to(cuda_dev)[None, :, None]).view(B * nW, nh, N, N)
    ow = (torch.softmax(s, -1) @ xw[2]).transpose(1, 2).reshape(B, nW * N, C)
    ref = torch.empty(B, T, C, device=cuda_dev)
    ref = ref.index_copy(1, src.reshape(-1), ow)
    if diag:
        _diag("o", o.view(B, T, C), ref.detach(), src, N, hd)
    _close(o.view(B, T, C), ref, 2e-2, "o")
    _close(lse, torch.logsumexp(s, -1), 1e-2, "lse")
    ref.backward(d_o.float().view(B, T, C))

    dqkv = torch.empty_like(qkv)
    dtable = torch.zeros_like(table)
    K.attn_bwd(d_o, qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o, lse, nh, hd, scale,
               dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], dbias_table=dtable, window=win, bias_table=table)
    if diag:
        for i, nm in enumerate(("dq", "dk", "dv")):
            _diag(nm, dqkv.view(B, T, 3 * C)[..., i * C:(i + 1) * C], x.grad[..., i * C:(i + 1) * C], src, N, hd)
        print("dtable: max err %.3g vs ref max %.3g" % ((dtable - tb.grad).abs().max().item(), tb.grad.abs().max().item()))
    _close(dqkv.view(B, T, 3 * C), x.grad, 3e-2, "dqkv")
    _close(dtable, tb.grad, 3e-2, "dtable")


@pytest.mark.parametrize("B,nh,hd,Lq,Lk,masked", [
    (3, 12, 64, 40, 40, True),      # RoBERTa self-attention
    (2, 12, 64, 50, 50, True),      # 50-token VQA text
    (2, 12, 64, 40, 576, False),    # t2i stage 2
    (2, 12, 64, 40, 144, False),    # t2i stage 3
    (2, 16, 32, 576, 40, True),     # i2t stage 2
    (2, 32, 32, 144, 40, True),     # i2t stage 3
    (1, 4, 32, 1296, 50, True),     # i2t at 576 px
])
def test_plain_attention_fwd_bwd(cuda_dev, B, nh, hd, Lq, Lk, masked):
    _plain_case(cuda_dev, B, nh, hd, Lq, Lk, masked)


# Small-CTA configurations of the plain backward (option "attn_small", default 7) and the generic kernel (0).
@pytest.mark.parametrize("small", [7, 0], ids=["small", "generic"])
@pytest.mark.parametrize("B,nh,hd,Lq,Lk,masked", [(3, 12, 64, 40, 40, True), (2, 12, 64, 48, 48, False),
                                                  (4, 12, 64, 33, 47, True), (256, 12, 64, 40, 40, True),
                                                  (2, 12, 64, 50, 50, True),
                                                  (2, 16, 32, 576, 40, True), (2, 32, 32, 144, 40, True),
                                                  (3, 16, 32, 100, 48, False), (1, 4, 32, 1296, 50, True),
                                                  (2, 12, 64, 40, 576, False), (2, 12, 64, 40, 144, False),
                                                  (3, 12, 64, 33, 100, True)])
def test_plain_attention_small_cfg(cuda_dev, B, nh, hd, Lq, Lk, masked, small):
    from fiber_b200 import lib
    lib.set_option("attn_small", small)
    try:
        _plain_case(cuda_dev, B, nh, hd, Lq, Lk, masked)
        torch.cuda.synchronize()
    finally:
        lib.set_option("attn_small", -1)


# tcgen05 generation of plain attention for at most 64 keys per group (csrc/attention_sk.cu; option "attn_sk": bit 0
# forward, bit 1 backward).  Cases with more keys must fall through to the mma.sync kernels with the option set.
SK_CASES = [(3, 12, 64, 40, 40, True), (2, 12, 64, 50, 50, True), (4, 12, 64, 33, 47, True), (256, 12, 64, 40, 40, True),
            (2, 16, 32, 576, 40, True), (2, 32, 32, 144, 40, True), (3, 16, 32, 100, 48, False), (1, 4, 32, 1296, 50, True),
            (5, 16, 32, 576, 64, False), (3, 12, 64, 130, 64, True), (2, 12, 64, 40, 576, False), (3, 12, 64, 33, 100, True),
            # packed self-attention (Lq == Lk, L % 8 == 0: 128 / L groups per tile), incl. a partial last tile
            (5, 12, 64, 40, 40, True), (7, 4, 32, 16, 16, True), (3, 12, 64, 64, 64, False), (4, 12, 64, 48, 48, True),
            (1, 12, 64, 24, 24, True)]


@pytest.mark.parametrize("mode", [0, 1, 21, 3, 15, 31, 26],
                         ids=["mmasync", "sk_long", "sk_all", "sk_long_fb", "default", "sk_all_fb", "sk_bwd_only"])
@pytest.mark.parametrize("B,nh,hd,Lq,Lk,masked", SK_CASES)
def test_plain_attention_tcgen05(cuda_dev, B, nh, hd, Lq, Lk, masked, mode):
    from fiber_b200 import lib
    before = lib.get_option("attn_sk_launches")
    lib.set_option("attn_sk", mode)
    try:
        _plain_case(cuda_dev, B, nh, hd, Lq, Lk, masked)
        torch.cuda.synchronize()
    finally:
        lib.set_option("attn_sk", -1)
    launched = lib.get_option("attn_sk_launches") - before
    packed = Lq == Lk and 32 <= Lq <= 64 and Lq % 8 == 0
    short = Lq < 96 and (packed or mode & 16)
    fwd = Lk <= 64 and ((mode & 1 and Lq >= 96) or (mode & 4 and short))
    bwd = Lk <= 64 and ((mode & 2 and Lq >= 96) or (mode & 8 and short))
    assert launched == int(bool(fwd)) + int(bool(bwd))


@pytest.mark.parametrize("B,nh,hd,Lq,Lk", [(3, 12, 64, 40, 40), (2, 12, 64, 130, 50), (2, 16, 32, 300, 40), (5, 12, 64, 48, 48)])
def test_plain_attention_tcgen05_backward_dropout_matches_mma_sync(cuda_dev, B, nh, hd, Lq, Lk):
    """The tcgen05 backward regenerates the forward's dropout mask from the same counter hash: with the same forward
    output it must give the mma.sync backward's gradients (to bf16 rounding of P / dS)."""
    from fiber_b200 import kernels as K, lib
    C = nh * hd
    q = _rand((B * Lq, C), cuda_dev, 21)
    kv = _rand((B * Lk, 2 * C), cuda_dev, 22)
    d_o = _rand((B * Lq, C), cuda_dev, 23)
    mask = torch.zeros(B, Lk, device=cuda_dev)
    mask[1, Lk - 5:] = -10000.0
    kw = dict(groups=B, lq=Lq, lk=Lk, key_mask=mask, drop_p=0.2, seed=4242)
    scale = hd ** -0.5
    lib.set_option("attn_sk", 0)
    try:
        o, lse = K.attn_fwd(q, kv[:, :C], kv[:, C:], nh, hd, scale, **kw)
        grads = []
        for mode in (0, 26):
            lib.set_option("attn_sk", mode)
            dq, dkv = torch.empty_like(q), torch.empty_like(kv)
            K.attn_bwd(d_o, q, kv[:, :C], kv[:, C:], o, lse, nh, hd, scale, dq, dkv[:, :C], dkv[:, C:], **kw)
            grads.append((dq, dkv))
    finally:
        lib.set_option("attn_sk", -1)
    (dq0, dkv0), (dq1, dkv1) = grads
    _close(dq1, dq0.float(), 2e-2, "dq with dropout")
    _close(dkv1, dkv0.float(), 2e-2, "dk / dv with dropout")


def test_plain_attention_tcgen05_dropout_matches_mma_sync(cuda_dev):
    """Probability dropout uses the same counter-based hash in both generations: the tcgen05 forward must reproduce the
    mma.sync forward's output (same kept set, same scaling) so that either backward can follow."""
    from fiber_b200 import kernels as K, lib
    _fwd_dropout_case(cuda_dev, 3, 12, 64, 130, 50)
    _fwd_dropout_case(cuda_dev, 5, 12, 64, 40, 40)  # packed tiles


def _fwd_dropout_case(cuda_dev, B, nh, hd, Lq, Lk):
    from fiber_b200 import kernels as K, lib
    C = nh * hd
    q = _rand((B * Lq, C), cuda_dev, 11)
    kv = _rand((B * Lk, 2 * C), cuda_dev, 12)
    outs = []
    for mode in (0, 21):
        lib.set_option("attn_sk", mode)
        try:
            outs.append(K.attn_fwd(q, kv[:, :C], kv[:, C:], nh, hd, hd ** -0.5, groups=B, lq=Lq, lk=Lk, drop_p=0.25, seed=99))
        finally:
            lib.set_option("attn_sk", -1)
    (o0, l0), (o1, l1) = outs
    _close(o1, o0.float(), 2e-2, "o with dropout")
    torch.testing.assert_close(l1, l0, rtol=1e-3, atol=1e-3)
    assert (o0.float().abs() > 0).float().mean() > 0.5


def _plain_case(cuda_dev, B, nh, hd, Lq, Lk, masked):
    from fiber_b200 import kernels as K
    C = nh * hd
    q = _rand((B * Lq, C), cuda_dev, 1)
    kv = _rand((B * Lk, 2 * C), cuda_dev, 2)
    d_o = _rand((B * Lq, C), cuda_dev, 3)
    mask = None
    if masked:
        mask = torch.zeros(B, Lk, device=cuda_dev)
        mask[0, Lk - 7:] = -10000.0
    scale = 1.0 / math.sqrt(hd)
    kw = dict(groups=B, lq=Lq, lk=Lk, key_mask=mask)
    o, lse = K.attn_fwd(q, kv[:, :C], kv[:, C:], nh, hd, scale, **kw)

    qf = q.float().view(B, Lq, nh, hd).transpose(1, 2).requires_grad_(True)
    kf = kv[:, :C].float().reshape(B, Lk, nh, hd).transpose(1, 2).requires_grad_(True)
    vf = kv[:, C:].float().reshape(B, Lk, nh, hd).transpose(1, 2).requires_grad_(True)
    s = qf @ kf.transpose(-1, -2) * scale
    if mask is not None:
        s = s + mask[:, None, None, :]
    ref = (torch.softmax(s, -1) @ vf).transpose(1, 2).reshape(B * Lq, C)
    _close(o, ref, 2e-2, "o")
    ref.backward(d_o.float())
    dq = torch.empty_like(q)
    dkv = torch.empty_like(kv)
    K.attn_bwd(d_o, q, kv[:, :C], kv[:, C:], o, lse, nh, hd, scale, dq, dkv[:, :C], dkv[:, C:], **kw)
    _close(dq.view(B, Lq, nh, hd).transpose(1, 2), qf.grad, 3e-2, "dq")
    _close(dkv[:, :C].reshape(B, Lk, nh, hd).transpose(1, 2), kf.grad, 3e-2, "dk")
    _close(dkv[:, C:].reshape(B, Lk, nh, hd).transpose(1, 2), vf.grad, 3e-2, "dv")


@pytest.mark.parametrize("B,nh,hd,Lq,Lk", [(2, 12, 64, 256, 1050), (1, 12, 64, 256, 4200), (2, 16, 32, 300, 500),
                                           (2, 32, 32, 1728, 256), (1, 16, 32, 5040, 200)])
def test_plain_attention_many_queries_and_keys(cuda_dev, B, nh, hd, Lq, Lk):
    """Fine-grained t2i shapes (256 query tokens x 4200 / 1050 image keys): several query chunks AND several key chunks —
    the backward falls to the small-tile configuration that leaves room for the shared-memory dQ accumulator — and
    fine-grained i2t shapes (5040 / 1728 image queries x a 256-token text query), where the wrapper runs the backward per
    chunk of 144 keys."""
    _plain_case(cuda_dev, B, nh, hd, Lq, Lk, False)


def test_attention_dropout_statistics(cuda_dev):
    """Probability dropout: E[o] is unchanged, the mask is identical in forward and backward."""
    from fiber_b200 import kernels as K
    B, nh, hd, L = 8, 12, 64, 40
    C = nh * hd
    q = _rand((B * L, C), cuda_dev, 1, 0.2)
    kv = _rand((B * L, 2 * C), cuda_dev, 2, 0.2)
    kv[:, C:] = 1.0  # v == 1  =>  o == sum of kept probabilities / keep
    kw = dict(groups=B, lq=L, lk=L)
    o, _ = K.attn_fwd(q, kv[:, :C], kv[:, C:], nh, hd, 0.125, drop_p=0.1, seed=123, **kw)
    o2, _ = K.attn_fwd(q, kv[:, :C], kv[:, C:], nh, hd, 0.125, drop_p=0.1, seed=123, **kw)
    assert torch.equal(o, o2)
    o3, _ = K.attn_fwd(q, kv[:, :C], kv[:, C:], nh, hd, 0.125, drop_p=0.1, seed=124, **kw)
    assert not torch.equal(o, o3)
    assert abs(o.float().mean().item() - 1.0) < 0.01
    assert 0.02 < o.float().std().item() < 0.2
