"""Thin tensor-level wrappers over the C-ABI: torch is used only for device memory and streams.

Every function enqueues one (or a few) hand-written sm_100a kernels on the current CUDA stream.
Nothing here computes in torch; a missing library or a non-CUDA tensor raises.
"""
import ctypes as C

import torch

from . import lib as _lib

BF16 = torch.bfloat16
F32 = torch.float32

ACT_NONE, ACT_GELU, ACT_GELU_GRAD = 0, 1, 2
# opt-in epilogues (include/fiber_b200.h): 3 = out GELU(v), preact buffer receives GELU'(v) (one TMEM pass);
# 4 = out = accumulator * aux.  Need M % 128 == 0, N % 32 == 0, bf16 outputs.
ACT_GELU_CACHE, ACT_MUL_AUX = 3, 4
ACT_GELU_ONEPASS = 5  # out GELU(v), preact buffer receives v: the results of ACT_GELU + preact, in one TMEM pass
ACT_RES_PF = 6        # act 0 with a residual (+ bias / scale / row_scale) bit for bit, residual rows prefetched
ACT_GELU_GRAD_PF = 7  # ACT_GELU_GRAD bit for bit, aux rows prefetched into registers one chunk ahead
OUT_BF16, OUT_F32, OUT_F32_ATOMIC = 0, 1, 2
# Residual epilogues through ACT_RES_PF (bit-identical; coalesced residual loads, two store boxes per warp) when the
# reduction is short (K <= 512: the epilogue-bound proj / fc2 shapes of Swin stages 0-2, 15-30 % faster in
# gpurun_out/r2m_gemm_shapes_respf.txt; neutral at K = 2048, which keeps the default epilogue and runs as CTA pairs).
# FIBER_GEMM_RES_PREFETCH=0 / set_res_prefetch(False) turn it off, =2 applies it at every K.
RES_PREFETCH = int(__import__("os").environ.get("FIBER_GEMM_RES_PREFETCH", "1"))


def set_res_prefetch(on):
    """0 / False: never; 1 / True: K <= 512 (default); 2: every K."""
    global RES_PREFETCH
    RES_PREFETCH = int(on)

# Optional in-situ profiling (bench.py): when a list is installed here every GEMM launch is bracketed
# by CUDA events on the launching stream and (class key, flops, algorithmic bytes, start, stop) recorded.
GEMM_PROFILE = None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _req(t, dtype, name):
    if not t.is_cuda:
        raise RuntimeError("fiber_b200: %s must be a CUDA tensor (no CPU path exists)" % name)
    if t.dtype != dtype:
        raise RuntimeError("fiber_b200: %s must be %s, got %s" % (name, dtype, t.dtype))


def _rowmajor_2d(t, name):
    if t.dim() != 2 or t.stride(1) != 1:
        raise RuntimeError("fiber_b200: %s must be 2-D with unit inner stride" % name)
    return t.stride(0)


def gemm(a, b, *, mn_major=False, bias=None, residual=None, aux=None, preact=None, scale=None,
         row_scale=None, rows_per_scale=1, act=ACT_NONE, out=None, out_dtype=BF16,
         accumulate=False, splits=0, colsum=None, row_count=None):
    """out[M,N] = epilogue(A . B^T).

    mn_major=False: a is [M,K], b is [N,K]  (forward / dgrad with a transposed weight copy)
    mn_major=True : a is [K,M], b is [K,N]  (wgrad: a = dY [rows,N_out], b = X [rows,K_in]);
                    colsum (fp32 [M]) additionally accumulates scale * sum_k a[k, :] (the bias gradient)
    row_count (int32 device scalar): only the first row_count activation rows carry work — output row tiles past it are
                    not written (mn_major=False), reduction rows past it are not read (mn_major=True)
    """
    _req(a, BF16, "a"); _req(b, BF16, "b")
    lda = _rowmajor_2d(a, "a"); ldb = _rowmajor_2d(b, "b")
    if mn_major:
        k, m = a.shape; k2, n = b.shape
    else:
        m, k = a.shape; n, k2 = b.shape
    if k != k2:
        raise RuntimeError("fiber_b200.gemm: reduction dims differ (%d vs %d)" % (k, k2))
    if out is None:
        if accumulate:
            out = torch.zeros((m, n), device=a.device, dtype=F32)
        else:
            out = torch.empty((m, n), device=a.device, dtype=out_dtype)
    ldc = _rowmajor_2d(out, "out")
    if accumulate:
        _req(out, F32, "out"); out_mode = OUT_F32_ATOMIC
    else:
        out_mode = OUT_BF16 if out.dtype == BF16 else OUT_F32
    args = _lib.GemmArgs()
    args.a, args.b, args.c = a.data_ptr(), b.data_ptr(), out.data_ptr()
    args.m, args.n, args.k = m, n, k
    args.lda, args.ldb, args.ldc = lda, ldb, ldc
    args.a_major = args.b_major = 1 if mn_major else 0
    if bias is not None:
        _req(bias, F32, "bias"); args.bias = bias.data_ptr()
    if residual is not None:
        _req(residual, BF16, "residual"); args.residual = residual.data_ptr()
        args.ldr = _rowmajor_2d(residual, "residual")
    if aux is not None:
        _req(aux, BF16, "aux"); args.aux = aux.data_ptr(); args.ldaux = _rowmajor_2d(aux, "aux")
    if preact is not None:
        _req(preact, BF16, "preact"); args.preact = preact.data_ptr()
        args.ldp = _rowmajor_2d(preact, "preact")
    if scale is not None:
        _req(scale, F32, "scale"); args.scale = scale.data_ptr()
    if row_scale is not None:
        _req(row_scale, F32, "row_scale"); args.row_scale = row_scale.data_ptr()
    args.rows_per_scale = rows_per_scale
    if (RES_PREFETCH and (k <= 512 or RES_PREFETCH == 2) and act == ACT_NONE and residual is not None and preact is None and aux is None and not mn_major
            and out_mode == OUT_BF16 and m % 128 == 0 and n % 32 == 0):
        act = ACT_RES_PF  # opt-in, bit-identical to the default residual epilogue
    args.act = act
    args.out_mode = out_mode
    args.splits = splits
    if colsum is not None:
        _req(colsum, F32, "colsum")
        if not mn_major or colsum.numel() != m or not colsum.is_contiguous():
            raise RuntimeError("fiber_b200.gemm: colsum needs mn_major=True and a contiguous fp32 [M] tensor")
        args.colsum = colsum.data_ptr()
    if row_count is not None:
        _req(row_count, torch.int32, "row_count"); args.row_count = row_count.data_ptr()
    if GEMM_PROFILE is None:
        _lib.check(_lib.load().fiber_gemm(C.byref(args), _stream()), "gemm")
        return out
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(_lib.load().fiber_gemm(C.byref(args), _stream()), "gemm")
    e1.record()
    esz = 2 if out.dtype == BF16 else 4
    nbytes = 2 * (m * k + n * k) + esz * m * n + 2 * m * n * sum(t is not None for t in (residual, aux, preact))
    epi = "+".join(x for x, on in (("bias", bias is not None), ("gelu", act == ACT_GELU), ("gelu_grad", act == ACT_GELU_GRAD),
                                   ("preact", preact is not None), ("res", residual is not None),
                                   ("wgrad", mn_major)) if on) or "plain"
    GEMM_PROFILE.append(((m, n, k, epi), 2.0 * m * n * k, nbytes, e0, e1))
    return out


def _ce_args(x, w, bias, labels, row_count):
    _req(x, BF16, "x"); _req(w, BF16, "w"); _req(bias, F32, "bias"); _req(labels, torch.int32, "labels")
    m, k = x.shape
    n, k2 = w.shape
    if k != k2 or bias.numel() != n or labels.numel() != m or n % 32 != 0:
        raise RuntimeError("fiber_b200.mlm_ce: shapes x %s, w %s, bias %d, labels %d (N %% 32 == 0)"
                           % (tuple(x.shape), tuple(w.shape), bias.numel(), labels.numel()))
    args = _lib.CeArgs()
    args.x, args.w, args.bias, args.labels = x.data_ptr(), w.data_ptr(), bias.data_ptr(), labels.data_ptr()
    args.m, args.n, args.k = m, n, k
    args.ldx, args.ldw = _rowmajor_2d(x, "x"), _rowmajor_2d(w, "w")
    if row_count is not None:
        _req(row_count, torch.int32, "row_count"); args.row_count = row_count.data_ptr()
    return args, m, n


def mlm_ce_fwd(x, w, bias, labels, row_count=None):
    """Fused decoder GEMM + cross-entropy forward (include/fiber_b200.h: fiber_mlm_ce_fwd).
    Returns (lse [M] fp32, log2 domain; loss_rows [M] fp32; pred [M] int32); the logits are never written."""
    args, m, n = _ce_args(x, w, bias, labels, row_count)
    mpad = (m + 127) // 128 * 128
    part = torch.empty((4 * ((n + 255) // 256), mpad, 4), device=x.device, dtype=F32)
    xl = torch.zeros(m, device=x.device, dtype=F32)
    lse = torch.empty(m, device=x.device, dtype=F32)
    loss_rows = torch.empty(m, device=x.device, dtype=F32)
    pred = torch.empty(m, device=x.device, dtype=torch.int32)
    args.part, args.label_logit, args.lse = part.data_ptr(), xl.data_ptr(), lse.data_ptr()
    args.loss_rows, args.pred = loss_rows.data_ptr(), pred.data_ptr()
    _lib.check(_lib.load().fiber_mlm_ce_fwd(C.byref(args), _stream()), "mlm_ce_fwd")
    return lse, loss_rows, pred


def mlm_ce_bwd(x, w, bias, labels, lse, gscale, row_count=None, out=None):
    """d(logits) [M, N] bf16 = (softmax - onehot) * gscale of the fused decoder + cross-entropy (fiber_mlm_ce_bwd); row
    tiles past row_count are left untouched."""
    args, m, n = _ce_args(x, w, bias, labels, row_count)
    _req(lse, F32, "lse"); _req(gscale, F32, "gscale")
    if out is None:
        out = torch.empty((m, n), device=x.device, dtype=BF16)
    args.lse, args.gscale = lse.data_ptr(), gscale.data_ptr()
    args.dlogits, args.lddl = out.data_ptr(), _rowmajor_2d(out, "dlogits")
    _lib.check(_lib.load().fiber_mlm_ce_bwd(C.byref(args), _stream()), "mlm_ce_bwd")
    return out


def _attn_args(q, k, v, o, lse, heads, head_dim, scale, *, window=None, groups=None, lq=None, lk=None,
               key_mask=None, bias_table=None, drop_p=0.0, seed=0):
    """window = (G, H, W, ws, shift) for mode 1; otherwise plain mode with (groups, lq, lk)."""
    a = _lib.AttnArgs()
    for name, t in (("q", q), ("k", k), ("v", v), ("o", o)):
        _req(t, BF16, name)
        if t.stride(-1) != 1:
            raise RuntimeError("fiber_b200.attention: %s must have unit inner stride" % name)
    a.q, a.k, a.v, a.o = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr()
    a.ldq, a.ldk, a.ldv, a.ldo = q.stride(-2), k.stride(-2), v.stride(-2), o.stride(-2)
    _req(lse, F32, "lse"); a.lse = lse.data_ptr()
    a.heads, a.head_dim, a.scale = heads, head_dim, float(scale)
    if window is not None:
        G, H, W, ws, shift = window
        a.mode, a.groups, a.h, a.w, a.ws, a.shift = 1, G, H, W, ws, shift
        a.lq = a.lk = ws * ws
        _req(bias_table, F32, "bias_table"); a.bias_table = bias_table.data_ptr()
    else:
        a.mode, a.groups, a.lq, a.lk = 0, groups, lq, lk
        if key_mask is not None:
            _req(key_mask, F32, "key_mask"); a.key_mask = key_mask.data_ptr()
    a.drop_p, a.seed = float(drop_p), int(seed)
    return a


def attn_fwd(q, k, v, heads, head_dim, scale, **kw):
    """Returns (o, lse).  q/k/v are 2-D row-major views (rows x channels, any row stride)."""
    rows = q.shape[0]
    o = torch.empty((rows, heads * head_dim), device=q.device, dtype=BF16)
    if kw.get("window") is not None:
        G, H, W, ws, _ = kw["window"]
        lse = torch.empty((G * (H // ws) * (W // ws), heads, ws * ws), device=q.device, dtype=F32)
    else:
        lse = torch.empty((kw["groups"], heads, kw["lq"]), device=q.device, dtype=F32)
    a = _attn_args(q, k, v, o, lse, heads, head_dim, scale, **kw)
    _lib.check(_lib.load().fiber_attn_fwd(C.byref(a), _stream()), "attn_fwd")
    return o, lse


def _attn_bwd_key_chunks(d_o, q, k, v, o, lse, heads, head_dim, scale, dq, dk, dv, kw):
    """Plain attention backward with MANY queries and MANY keys per group (fine-grained i2t: 5040 / 1728 image queries x a
    256-token text query): the kernels keep the fp32 dQ accumulator of every query chunk in shared memory while they walk
    the key chunks, which does not fit there.  The softmax statistics (lse) are global, so the backward is exact per key
    chunk: run it once per chunk of <= 144 keys (contiguous copies of the few key / value rows), add the dQ parts, scatter
    the dK / dV parts.  No dropout (the keep-mask hash indexes keys by the full row length)."""
    G, Lq, Lk = kw["groups"], kw["lq"], kw["lk"]
    if kw.get("drop_p", 0.0):
        raise RuntimeError("fiber_b200.attn_bwd: dropout with > 144 queries and > 144 keys per group is not supported")
    Ck = k.shape[1]
    k3, v3 = k.reshape(G, Lk, Ck), v.reshape(G, Lk, Ck)
    km = kw.get("key_mask")
    acc = None
    for c0 in range(0, Lk, 144):
        c1 = min(Lk, c0 + 144)
        kc, vc = k3[:, c0:c1].reshape(G * (c1 - c0), Ck).contiguous(), v3[:, c0:c1].reshape(G * (c1 - c0), Ck).contiguous()
        dkc, dvc, dqc = torch.empty_like(kc), torch.empty_like(vc), torch.empty_like(dq, memory_format=torch.contiguous_format)
        kw_c = dict(kw, lk=c1 - c0, key_mask=None if km is None else km[:, c0:c1].contiguous())
        attn_bwd(d_o, q, kc, vc, o, lse, heads, head_dim, scale, dqc, dkc, dvc, **kw_c)
        acc = dqc.float() if acc is None else acc.add_(dqc)
        for dst, part in ((dk, dkc), (dv, dvc)):  # dk / dv may be column slices of a packed [G Lk, 2C] buffer
            dst.unflatten(0, (G, Lk))[:, c0:c1] = part.view(G, c1 - c0, Ck)
    dq.copy_(acc)


def attn_bwd(d_o, q, k, v, o, lse, heads, head_dim, scale, dq, dk, dv, dbias_table=None, **kw):
    """Writes dq/dk/dv (bf16 views with the layout of q/k/v); accumulates into dbias_table."""
    if (kw.get("window") is None and kw.get("lq", 0) > 144 and kw.get("lk", 0) > 144
            and kw["lq"] * head_dim * 4 > (96 << 10)):
        return _attn_bwd_key_chunks(d_o, q, k, v, o, lse, heads, head_dim, scale, dq, dk, dv, kw)
    a = _attn_args(q, k, v, o, lse, heads, head_dim, scale, **kw)
    for name, t in (("d_o", d_o), ("dq", dq), ("dk", dk), ("dv", dv)):
        _req(t, BF16, name)
    a.d_o, a.dq, a.dk, a.dv = d_o.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    a.lddo, a.lddq, a.lddk, a.lddv = d_o.stride(-2), dq.stride(-2), dk.stride(-2), dv.stride(-2)
    if dbias_table is not None:
        _req(dbias_table, F32, "dbias_table"); a.dbias_table = dbias_table.data_ptr()
    scratch = None
    if kw.get("window") is not None:
        scratch = torch.empty((q.shape[0], heads), device=q.device, dtype=F32)
        a.d_scratch = scratch.data_ptr()
    _lib.check(_lib.load().fiber_attn_bwd(C.byref(a), _stream()), "attn_bwd")


# ---------------------------------------------------------------------------------------------
# row-wise kernels
# ---------------------------------------------------------------------------------------------
def layernorm_fwd(x, gamma, beta, eps, *, add=None, want_sum=False, merge=None, save_stats=True):
    """y = LN(x [+ add]).  merge=(B, H, W): x is [B*H*W, Cin] and rows are the 2x2-merged tokens.
    Returns (y, mean, rstd, sum_or_None)."""
    _req(x, BF16, "x"); _req(gamma, F32, "gamma"); _req(beta, F32, "beta")
    a = _lib.LnArgs()
    a.in1, a.ld1 = x.data_ptr(), _rowmajor_2d(x, "x")
    if merge is not None:
        B, H, W = merge
        cin = x.shape[1]
        rows, c = B * (H // 2) * (W // 2), 4 * cin
        a.merge, a.h, a.w, a.cin = 1, H, W, cin
    else:
        rows, c = x.shape
    y = torch.empty((rows, c), device=x.device, dtype=BF16)
    a.out, a.ldo, a.rows, a.c = y.data_ptr(), c, rows, c
    a.gamma, a.beta, a.eps = gamma.data_ptr(), beta.data_ptr(), float(eps)
    mean = rstd = s = None
    if save_stats:
        mean = torch.empty(rows, device=x.device, dtype=F32)
        rstd = torch.empty(rows, device=x.device, dtype=F32)
        a.mean, a.rstd = mean.data_ptr(), rstd.data_ptr()
    if add is not None:
        _req(add, BF16, "add"); a.in2, a.ld2 = add.data_ptr(), _rowmajor_2d(add, "add")
        if want_sum:
            s = torch.empty((rows, c), device=x.device, dtype=BF16)
            a.sum_out, a.lds = s.data_ptr(), c
    _lib.check(_lib.load().fiber_layernorm_fwd(C.byref(a), _stream()), "layernorm_fwd")
    return y, mean, rstd, s


def layernorm_bwd(dy, x, mean, rstd, gamma, *, add=None, merge=None, dres=None, dgamma=None, dbeta=None,
                  row_scale=None, rows_per_scale=1):
    """Returns dx (shape of x).  dgamma/dbeta (fp32) are accumulated into when given.
    With row_scale (fp32 per-sample DropPath scales) returns (dx, dx * row_scale[row // rows_per_scale])."""
    _req(dy, BF16, "dy"); _req(x, BF16, "x")
    a = _lib.LnArgs()
    a.in1, a.ld1 = x.data_ptr(), _rowmajor_2d(x, "x")
    rows, c = dy.shape
    if merge is not None:
        B, H, W = merge
        a.merge, a.h, a.w, a.cin = 1, H, W, x.shape[1]
    if add is not None:
        a.in2, a.ld2 = add.data_ptr(), _rowmajor_2d(add, "add")
    a.rows, a.c = rows, c
    a.gamma, a.mean, a.rstd = gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr()
    a.dy, a.lddy = dy.data_ptr(), _rowmajor_2d(dy, "dy")
    dx = torch.empty_like(x)
    a.dx, a.lddx = dx.data_ptr(), dx.stride(0)
    if dres is not None:
        _req(dres, BF16, "dres"); a.dres, a.lddres = dres.data_ptr(), _rowmajor_2d(dres, "dres")
    if dgamma is not None:
        _req(dgamma, F32, "dgamma"); _req(dbeta, F32, "dbeta")
        a.dgamma, a.dbeta = dgamma.data_ptr(), dbeta.data_ptr()
    dxs = None
    if row_scale is not None:
        _req(row_scale, F32, "row_scale")
        dxs = torch.empty_like(x)
        a.row_scale, a.rows_per_scale = row_scale.data_ptr(), rows_per_scale
        a.dx_scaled, a.lddxs = dxs.data_ptr(), dxs.stride(0)
    _lib.check(_lib.load().fiber_layernorm_bwd(C.byref(a), _stream()), "layernorm_bwd")
    return dx if row_scale is None else (dx, dxs)


def colsum(x, out=None, scale=None, row_scale=None, rows_per_scale=1):
    _req(x, BF16, "x")
    m, n = x.shape
    if out is None:
        out = torch.zeros(n, device=x.device, dtype=F32)
    _lib.check(_lib.load().fiber_colsum(x.data_ptr(), _rowmajor_2d(x, "x"), m, n, out.data_ptr(), _ptr(scale),
                                        _ptr(row_scale), rows_per_scale, _stream()), "colsum")
    return out


def dot(a, b, out=None):
    _req(a, BF16, "a"); _req(b, BF16, "b")
    m, n = a.shape
    if out is None:
        out = torch.zeros(1, device=a.device, dtype=F32)
    _lib.check(_lib.load().fiber_dot(a.data_ptr(), _rowmajor_2d(a, "a"), b.data_ptr(), _rowmajor_2d(b, "b"), m, n,
                                     out.data_ptr(), _stream()), "dot")
    return out


def dropout(x, p, seed, out=None):
    _req(x, BF16, "x")
    m, n = x.shape
    if out is None:
        out = torch.empty((m, n), device=x.device, dtype=BF16)
    _lib.check(_lib.load().fiber_dropout(x.data_ptr(), _rowmajor_2d(x, "x"), out.data_ptr(), out.stride(0), m, n,
                                         float(p), int(seed), _stream()), "dropout")
    return out


def scale_rows(x, row_scale, rows_per_scale, out=None):
    _req(x, BF16, "x"); _req(row_scale, F32, "row_scale")
    m, n = x.shape
    if out is None:
        out = torch.empty((m, n), device=x.device, dtype=BF16)
    _lib.check(_lib.load().fiber_scale_rows(x.data_ptr(), _rowmajor_2d(x, "x"), out.data_ptr(), out.stride(0), m, n,
                                            row_scale.data_ptr(), rows_per_scale, _stream()), "scale_rows")
    return out


def cast_bf16(x):
    _req(x, F32, "x")
    x = x.contiguous()
    y = torch.empty(x.shape, device=x.device, dtype=BF16)
    _lib.check(_lib.load().fiber_cast_f32_bf16(x.data_ptr(), y.data_ptr(), x.numel(), _stream()), "cast")
    return y


def cast_transpose(w, w_out=None, wt_out=None):
    """fp32 [N,K] -> bf16 [N,K] view and/or bf16 [K,N] view (views may be slices of packed buffers)."""
    _req(w, F32, "w")
    n, k = w.shape
    _lib.check(_lib.load().fiber_cast_transpose(
        w.data_ptr(), _rowmajor_2d(w, "w"), n, k,
        _ptr(w_out), 0 if w_out is None else w_out.stride(0),
        _ptr(wt_out), 0 if wt_out is None else wt_out.stride(0), _stream()), "cast_transpose")


def grid_copy(src, out_hw, row_scale=None, add=None):
    """src [B, Hs, Ws, C] bf16 contiguous -> [B, Hd, Wd, C]: crop (Hd <= Hs) or zero-pad (Hd >= Hs), optionally scaled
    per sample (fp32 [B]) and added to `add` [B, Hd, Wd, C]."""
    _req(src, BF16, "src")
    if src.dim() != 4 or not src.is_contiguous():
        raise RuntimeError("fiber_b200.grid_copy expects a contiguous [B, H, W, C] tensor")
    B, Hs, Ws, Cc = src.shape
    Hd, Wd = out_hw
    out = torch.empty((B, Hd, Wd, Cc), device=src.device, dtype=BF16)
    if add is not None:
        _req(add, BF16, "add")
        if tuple(add.shape) != (B, Hd, Wd, Cc) or not add.is_contiguous():
            raise RuntimeError("fiber_b200.grid_copy: add must be contiguous [B, Hd, Wd, C]")
    if row_scale is not None:
        _req(row_scale, F32, "row_scale")
    _lib.check(_lib.load().fiber_grid_copy(src.data_ptr(), out.data_ptr(), _ptr(add), _ptr(row_scale), B, Hs, Ws, Hd, Wd, Cc,
                                           _stream()), "grid_copy")
    return out


def patch_gather(img):
    """im2row of the 4x4 / stride-4 patch embedding: fp32 [B,3,H,W] -> bf16 [B*(H/4)*(W/4), 64] (48 values + zero pad)."""
    _req(img, F32, "img")
    img = img.contiguous()
    b, ch, h, w = img.shape
    if ch != 3 or h % 4 or w % 4:
        raise RuntimeError("fiber_b200.patch_gather expects [B,3,H,W] with H and W multiples of 4")
    out = torch.empty((b * (h // 4) * (w // 4), 64), device=img.device, dtype=BF16)
    _lib.check(_lib.load().fiber_patch_gather_hw(img.data_ptr(), out.data_ptr(), b, h, w, _stream()), "patch_gather")
    return out


def embed_gather(ids, word, pos, type_, pad_id=1):
    if ids.dtype != torch.int64 or not ids.is_cuda:
        raise RuntimeError("fiber_b200.embed_gather: ids must be a CUDA int64 tensor")
    ids = ids.contiguous()
    b, l = ids.shape
    c = word.shape[1]
    out = torch.empty((b * l, c), device=ids.device, dtype=BF16)
    _lib.check(_lib.load().fiber_embed_gather(ids.data_ptr(), b, l, c, pad_id, word.data_ptr(), pos.data_ptr(),
                                              type_.data_ptr(), out.data_ptr(), c, _stream()), "embed_gather")
    return out


def embed_scatter(ids, dsum, dword, dpos, pad_id=1):
    ids = ids.contiguous()
    b, l = ids.shape
    _lib.check(_lib.load().fiber_embed_scatter(ids.data_ptr(), b, l, dsum.shape[1], pad_id, dsum.data_ptr(),
                                               _rowmajor_2d(dsum, "dsum"), dword.data_ptr(), dpos.data_ptr(),
                                               _stream()), "embed_scatter")


def axpy(x, add, alpha=None):
    """add + alpha * x (alpha: 1-element fp32 device tensor or None for 1)."""
    _req(x, BF16, "x"); _req(add, BF16, "add")
    m, n = x.shape
    out = torch.empty((m, n), device=x.device, dtype=BF16)
    _lib.check(_lib.load().fiber_axpy(x.data_ptr(), _rowmajor_2d(x, "x"), add.data_ptr(), _rowmajor_2d(add, "add"),
                                      _ptr(alpha), out.data_ptr(), n, m, n, _stream()), "axpy")
    return out
