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
OUT_BF16, OUT_F32, OUT_F32_ATOMIC = 0, 1, 2


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
         accumulate=False, splits=0):
    """out[M,N] = epilogue(A . B^T).

    mn_major=False: a is [M,K], b is [N,K]  (forward / dgrad with a transposed weight copy)
    mn_major=True : a is [K,M], b is [K,N]  (wgrad: a = dY [rows,N_out], b = X [rows,K_in])
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
    args.act = act
    args.out_mode = out_mode
    args.splits = splits
    _lib.check(_lib.load().fiber_gemm(C.byref(args), _stream()), "gemm")
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


def attn_bwd(d_o, q, k, v, o, lse, heads, head_dim, scale, dq, dk, dv, dbias_table=None, **kw):
    """Writes dq/dk/dv (bf16 views with the layout of q/k/v); accumulates into dbias_table."""
    a = _attn_args(q, k, v, o, lse, heads, head_dim, scale, **kw)
    for name, t in (("d_o", d_o), ("dq", dq), ("dk", dk), ("dv", dv)):
        _req(t, BF16, name)
    a.d_o, a.dq, a.dk, a.dv = d_o.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    a.lddo, a.lddq, a.lddk, a.lddv = d_o.stride(-2), dq.stride(-2), dk.stride(-2), dv.stride(-2)
    if dbias_table is not None:
        _req(dbias_table, F32, "dbias_table"); a.dbias_table = dbias_table.data_ptr()
    _lib.check(_lib.load().fiber_attn_bwd(C.byref(a), _stream()), "attn_bwd")
