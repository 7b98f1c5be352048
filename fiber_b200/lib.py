"""ctypes binding of the C-ABI library (include/fiber_b200.h).

The product path has no fallback: if libfiber_b200.so is missing or a call fails, a
RuntimeError is raised.  `load()` does not need a GPU (the library only touches CUDA when an
entry point runs), which lets the CPU test-suite check that every declared symbol is exported.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FIBER_B200_LIB") or os.path.join(_HERE, "libfiber_b200.so")  # env: A/B builds (tools)

_lib = None


class GemmArgs(C.Structure):
    """Mirror of `fiber_gemm_args` (include/fiber_b200.h)."""
    _fields_ = [
        ("a", C.c_void_p), ("b", C.c_void_p), ("c", C.c_void_p),
        ("m", C.c_int32), ("n", C.c_int32), ("k", C.c_int32),
        ("lda", C.c_int64), ("ldb", C.c_int64), ("ldc", C.c_int64),
        ("a_major", C.c_int32), ("b_major", C.c_int32),
        ("bias", C.c_void_p),
        ("residual", C.c_void_p), ("ldr", C.c_int64),
        ("aux", C.c_void_p), ("ldaux", C.c_int64),
        ("preact", C.c_void_p), ("ldp", C.c_int64),
        ("scale", C.c_void_p),
        ("row_scale", C.c_void_p), ("rows_per_scale", C.c_int32),
        ("act", C.c_int32), ("out_mode", C.c_int32), ("splits", C.c_int32),
        ("colsum", C.c_void_p),
        ("row_count", C.c_void_p),
    ]


class CeArgs(C.Structure):
    """Mirror of `fiber_ce_args` (include/fiber_b200.h)."""
    _fields_ = [
        ("x", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p), ("labels", C.c_void_p), ("row_count", C.c_void_p),
        ("m", C.c_int32), ("n", C.c_int32), ("k", C.c_int32),
        ("ldx", C.c_int64), ("ldw", C.c_int64),
        ("part", C.c_void_p), ("label_logit", C.c_void_p), ("lse", C.c_void_p), ("loss_rows", C.c_void_p),
        ("pred", C.c_void_p), ("dlogits", C.c_void_p), ("lddl", C.c_int64), ("gscale", C.c_void_p),
    ]


class AttnArgs(C.Structure):
    """Mirror of `fiber_attn_args` (include/fiber_b200.h)."""
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("o", C.c_void_p), ("lse", C.c_void_p),
        ("ldq", C.c_int64), ("ldk", C.c_int64), ("ldv", C.c_int64), ("ldo", C.c_int64),
        ("mode", C.c_int32), ("groups", C.c_int32), ("heads", C.c_int32), ("lq", C.c_int32),
        ("lk", C.c_int32), ("head_dim", C.c_int32),
        ("scale", C.c_float),
        ("key_mask", C.c_void_p),
        ("h", C.c_int32), ("w", C.c_int32), ("ws", C.c_int32), ("shift", C.c_int32),
        ("bias_table", C.c_void_p),
        ("drop_p", C.c_float),
        ("seed", C.c_uint64),
        ("d_o", C.c_void_p), ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p),
        ("lddo", C.c_int64), ("lddq", C.c_int64), ("lddk", C.c_int64), ("lddv", C.c_int64),
        ("dbias_table", C.c_void_p),
        ("d_scratch", C.c_void_p),
    ]


class LnArgs(C.Structure):
    """Mirror of `fiber_ln_args` (include/fiber_b200.h)."""
    _fields_ = [
        ("in1", C.c_void_p), ("in2", C.c_void_p), ("ld1", C.c_int64), ("ld2", C.c_int64),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("eps", C.c_float),
        ("out", C.c_void_p), ("ldo", C.c_int64), ("mean", C.c_void_p), ("rstd", C.c_void_p),
        ("sum_out", C.c_void_p), ("lds", C.c_int64), ("rows", C.c_int64), ("c", C.c_int32),
        ("merge", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("cin", C.c_int32),
        ("dy", C.c_void_p), ("lddy", C.c_int64), ("dres", C.c_void_p), ("lddres", C.c_int64),
        ("dx", C.c_void_p), ("lddx", C.c_int64), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p),
        ("row_scale", C.c_void_p), ("rows_per_scale", C.c_int32), ("dx_scaled", C.c_void_p), ("lddxs", C.c_int64),
    ]


class ImageDesc(C.Structure):
    """Mirror of `fiber_image_desc` (include/fiber_b200.h)."""
    _fields_ = [
        ("src", C.c_void_p), ("stride", C.c_int64), ("h", C.c_int32), ("w", C.c_int32),
        ("box_x", C.c_int32), ("box_y", C.c_int32), ("box_w", C.c_int32), ("box_h", C.c_int32),
        ("flip", C.c_int32), ("ksize_x", C.c_int32), ("ksize_y", C.c_int32), ("planar", C.c_int32),
        ("coef_off", C.c_int64), ("tmp_off", C.c_int64), ("chan_stride", C.c_int64),
    ]


def declared_symbols():
    """Every `fiber_*` function declared in include/fiber_b200.h (parsed from the header)."""
    import re
    hdr = os.path.join(os.path.dirname(_HERE), "include", "fiber_b200.h")
    with open(hdr) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fiber_[a-z0-9_]+)\s*\(", text)))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "fiber_b200: %s not found — build it with `python -m fiber_b200.build` "
            "(there is no CPU or eager fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.fiber_last_error.restype = C.c_char_p
    lib.fiber_launch_count.restype = C.c_int64
    lib.fiber_set_option.argtypes = [C.c_char_p, C.c_int32]
    lib.fiber_get_option.argtypes = [C.c_char_p]
    lib.fiber_gemm.argtypes = [C.POINTER(GemmArgs), C.c_void_p]
    lib.fiber_mlm_ce_fwd.argtypes = [C.POINTER(CeArgs), C.c_void_p]
    lib.fiber_mlm_ce_bwd.argtypes = [C.POINTER(CeArgs), C.c_void_p]
    lib.fiber_attn_fwd.argtypes = [C.POINTER(AttnArgs), C.c_void_p]
    lib.fiber_attn_bwd.argtypes = [C.POINTER(AttnArgs), C.c_void_p]
    V, I64, I32, F = C.c_void_p, C.c_int64, C.c_int32, C.c_float
    lib.fiber_layernorm_fwd.argtypes = [C.POINTER(LnArgs), V]
    lib.fiber_layernorm_bwd.argtypes = [C.POINTER(LnArgs), V]
    lib.fiber_colsum.argtypes = [V, I64, I64, I32, V, V, V, I32, V]
    lib.fiber_dot.argtypes = [V, I64, V, I64, I64, I32, V, V]
    lib.fiber_dropout.argtypes = [V, I64, V, I64, I64, I32, F, C.c_uint64, V]
    lib.fiber_scale_rows.argtypes = [V, I64, V, I64, I64, I32, V, I32, V]
    lib.fiber_cast_f32_bf16.argtypes = [V, V, I64, V]
    lib.fiber_axpy.argtypes = [V, I64, V, I64, V, V, I64, I64, I32, V]
    lib.fiber_cast_transpose.argtypes = [V, I64, I32, I32, V, I64, V, I64, V]
    lib.fiber_grid_copy.argtypes = [V, V, V, V, I32, I32, I32, I32, I32, I32, V]
    lib.fiber_patch_gather.argtypes = [V, V, I32, I32, V]
    lib.fiber_patch_gather_hw.argtypes = [V, V, I32, I32, I32, V]
    lib.fiber_embed_gather.argtypes = [V, I32, I32, I32, I32, V, V, V, V, I64, V]
    lib.fiber_embed_scatter.argtypes = [V, I32, I32, I32, I32, V, I64, V, V, V]
    lib.fiber_adamw_multi.argtypes = [V, V, I32, I32, F, F, F, I32, V]
    lib.fiber_image_transform_plan.restype = C.c_size_t
    lib.fiber_image_transform_plan.argtypes = [C.POINTER(ImageDesc), I32, I32, I32]
    lib.fiber_image_transform.argtypes = [C.POINTER(ImageDesc), V, I32, I32, I32, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                          V, C.c_size_t, V, V]
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError("fiber_b200.%s failed (%d): %s" % (what, rc, load().fiber_last_error().decode()))


def launch_count():
    return int(load().fiber_launch_count())


def set_option(name, value):
    """Process-wide kernel selection (include/fiber_b200.h: fiber_set_option), e.g. ("winattn_tc", 3)."""
    check(load().fiber_set_option(name.encode(), int(value)), "set_option")


def get_option(name):
    v = load().fiber_get_option(name.encode())
    if v < 0:
        raise RuntimeError("fiber_b200.get_option failed: %s" % load().fiber_last_error().decode())
    return int(v)
