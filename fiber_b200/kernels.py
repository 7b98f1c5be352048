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
