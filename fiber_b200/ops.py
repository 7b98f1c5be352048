"""Block-level autograd Functions of the FIBER hot path, each a straight-line sequence of C-ABI
kernel launches (fiber_b200.kernels) with a hand-written backward.

Layout: every activation is a 2-D bf16 [rows, C] tensor in IMAGE order (rows = B*H*W) or text order
(rows = B*L); Swin's roll / window_partition / window_reverse never happen as copies — the window
attention kernel gathers through the closed-form map.  Parameters stay fp32 masters (reference
state_dict layout); bf16 (and transposed) copies are cached per parameter version.

Reference semantics restated here (file:line under coarse_grained/fiber/modules/):
  SwinTransformerBlock.forward   swin_transformer.py:356-393   -> SwinBlockFn
  WindowAttention.forward        swin_transformer.py:195-261   (inside SwinBlockFn)
  PatchMerging.forward           swin_transformer.py:411-432   -> PatchMergeFn
  timm PatchEmbed                fiber_module.py:311           -> PatchEmbedFn
  RobertaEmbeddings.forward      roberta.py:169-199            -> RobertaEmbedFn
  RobertaLayer.forward           roberta.py:441-502            -> RobertaLayerFn
"""
import itertools
import math
import os
import weakref

import torch

from . import kernels as K

BF16 = torch.bfloat16
F32 = torch.float32
LN_EPS = 1e-5

_seed_counter = itertools.count(1)
_base_seed = None  # resolved lazily: torch's seed (torch.manual_seed / pl.seed_everything) and the distributed rank


def set_dropout_seed(seed):
    """Pin the dropout / DropPath-independent hash stream explicitly (benchmarks, tests)."""
    global _base_seed, _seed_counter
    _base_seed = int(seed)
    _seed_counter = itertools.count(1)


def _resolve_base_seed():
    """Dropout masks come from counter-based hashes (not torch's Philox stream).  Their base seed follows
    torch.initial_seed() — so torch.manual_seed / pl.seed_everything select the mask sequence — and differs per
    data-parallel rank (identical masks on every rank would correlate the replicas' gradients)."""
    global _base_seed
    rank = 0
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        rank = torch.distributed.get_rank()
    _base_seed = (int(torch.initial_seed()) * 2654435761 + 0x9E3779B97F4A7C15 * (rank + 1)) & 0x7FFFFFFF
    return _base_seed


def _next_seed():
    base = _base_seed if _base_seed is not None else _resolve_base_seed()
    return (base * 1000003 + next(_seed_counter)) & 0x7FFFFFFFFFFF


# ---------------------------------------------------------------------------------------------
# bf16 weight cache
# ---------------------------------------------------------------------------------------------
class _WeightCache:
    """bf16 [N,K] and transposed [K,N] copies of (packed) fp32 Linear weights; packed fp32 biases.
    Refreshed whenever any source parameter's in-place version counter changes (optimizer step,
    load_state_dict).  NOTE: writes through `param.data` (EMA, manual clipping via p.data.mul_) do not bump the version
    counter — call CACHE.clear() after such an update.  Entries whose parameters have been garbage-collected are
    evicted whenever the cache grows (models come and go in tests / sweeps)."""

    def __init__(self):
        self._w = {}
        self._b = {}
        self._sweep_at = 256

    def _sweep(self):
        if len(self._w) + len(self._b) < self._sweep_at:
            return
        for d, ri in ((self._w, 3), (self._b, 2)):
            for key in [k for k, v in d.items() if any(r() is None for r in v[ri])]:
                del d[key]
        self._sweep_at = max(256, 2 * (len(self._w) + len(self._b)))

    @staticmethod
    def _key(ps):
        return tuple(id(p) for p in ps), tuple((p._version, p.data_ptr()) for p in ps)

    @staticmethod
    def _alive(refs, ps):
        return all(r() is p for r, p in zip(refs, ps))

    def weights(self, ps, need_t=True, pad_k=0, pad_n=0):
        """ps: tuple of fp32 [N_i, K] (or conv [N, ...]) weights packed along N (zero rows up to pad_n)."""
        ids, vers = self._key(ps)
        hit = self._w.get(ids)
        if hit is not None and hit[0] == vers and self._alive(hit[3], ps) and (hit[2] is not None or not need_t):
            return hit[1], hit[2]
        with torch.no_grad():
            mats = [p.detach().reshape(p.shape[0], -1) for p in ps]
            k = mats[0].shape[1]
            kp = max(k, pad_k)
            n_rows = sum(m.shape[0] for m in mats)
            n_total = max(n_rows, pad_n)
            dev = mats[0].device
            padded = kp != k or n_total != n_rows
            w = (torch.zeros if padded else torch.empty)((n_total, kp), device=dev, dtype=BF16)
            wt = (torch.zeros if padded else torch.empty)((k, n_total), device=dev, dtype=BF16) if need_t else None
            n0 = 0
            for m in mats:
                n = m.shape[0]
                K.cast_transpose(m, w[n0:n0 + n], None if wt is None else wt[:, n0:n0 + n])
                n0 += n
        self._w[ids] = (vers, w, wt, tuple(weakref.ref(p) for p in ps))
        self._sweep()
        return w, wt

    def bias(self, ps):
        if len(ps) == 1:
            return ps[0].detach()
        ids, vers = self._key(ps)
        hit = self._b.get(ids)
        if hit is not None and hit[0] == vers and self._alive(hit[2], ps):
            return hit[1]
        with torch.no_grad():
            b = torch.cat([p.detach() for p in ps])
        self._b[ids] = (vers, b, tuple(weakref.ref(p) for p in ps))
        self._sweep()
        return b

    def clear(self):
        self._w.clear()
        self._b.clear()


CACHE = _WeightCache()


class _GradArena(dict):
    """One zero-filled fp32 buffer per backward call; gradients are named views of it (a single
    memset instead of one fill kernel per gradient tensor)."""

    def __init__(self, shapes, device):
        super().__init__()
        sizes = {}
        for n, shp in shapes.items():
            numel = 1
            for d in shp:
                numel *= d
            sizes[n] = (numel, ((numel + 3) // 4) * 4)  # keep every view 16-byte aligned
        flat = torch.zeros(sum(v[1] for v in sizes.values()), device=device, dtype=F32)
        off = 0
        for n, shp in shapes.items():
            self[n] = flat[off:off + sizes[n][0]].view(shp)
            off += sizes[n][1]


# Default (FIBER_GELU_CACHE=0 or set_gelu_cache(False) restores the h-saving epilogues): the fc1 GEMM stores GELU'(h) instead of the pre-activation h
# as its second output (one pass over TMEM, one erfc per element) and the fc2 dgrad GEMM multiplies by it instead of
# evaluating the exact-erf GELU' in its epilogue (gemm_sm100.cu, kernel template parameter EPI = 1).  GELU(h) itself is
# bit-identical to the default path; the gradient sees GELU'(h) rounded to bf16 instead of h rounded to bf16.
GELU_CACHE = os.environ.get("FIBER_GELU_CACHE", "1") == "1"
# Opt-in (FIBER_GELU_ONEPASS=1): same outputs as the default fc1 epilogue (GELU(h) and h), bit for bit, from one pass
# over TMEM instead of two; the backward is unchanged.
GELU_ONEPASS = os.environ.get("FIBER_GELU_ONEPASS", "0") == "1"


# Opt-in (FIBER_GELU_GRAD_PREFETCH=1): the default fc2-dgrad epilogue (acc * GELU'(h)), bit for bit, with the h rows
# loaded into registers one chunk ahead instead of after the accumulator wait.
GELU_GRAD_PREFETCH = os.environ.get("FIBER_GELU_GRAD_PREFETCH", "0") == "1"


def set_gelu_grad_prefetch(on):
    global GELU_GRAD_PREFETCH
    GELU_GRAD_PREFETCH = bool(on)


def set_gelu_cache(on):
    global GELU_CACHE
    GELU_CACHE = bool(on)


def set_gelu_onepass(on):
    global GELU_ONEPASS
    GELU_ONEPASS = bool(on)


def _fc1_gelu(x, w, bias, buf):
    """a = GELU(x W^T + b).  `buf` receives what the backward needs — h, or GELU'(h) when the opt-in epilogue applies
    (returned flag); _fc2_dgrad_gelu takes the same flag."""
    if (GELU_CACHE or GELU_ONEPASS) and x.shape[0] % 128 == 0 and w.shape[0] % 32 == 0:
        if GELU_CACHE:
            return K.gemm(x, w, bias=bias, act=K.ACT_GELU_CACHE, preact=buf), True
        return K.gemm(x, w, bias=bias, act=K.ACT_GELU_ONEPASS, preact=buf), False
    return K.gemm(x, w, bias=bias, act=K.ACT_GELU, preact=buf), False


def _fc2_dgrad_gelu(dz, w2_t, buf, cached):
    """dh = (dz W2) * GELU'(h)"""
    if cached:
        return K.gemm(dz, w2_t, aux=buf, act=K.ACT_MUL_AUX)
    if GELU_GRAD_PREFETCH and dz.shape[0] % 128 == 0 and w2_t.shape[0] % 32 == 0:
        return K.gemm(dz, w2_t, aux=buf, act=K.ACT_GELU_GRAD_PF)
    return K.gemm(dz, w2_t, aux=buf, act=K.ACT_GELU_GRAD)


def _wgrad(dy, x, scale=None, out=None, db=None, row_count=None):
    """dW[N,K] = dy[rows,N]^T x[rows,K]  (fp32, split-K atomics on a zeroed buffer); db[N] += colsum(dy)
    comes out of the same kernel (one extra 128x16 MMA per k-step against a tile of ones)."""
    return K.gemm(dy, x, mn_major=True, accumulate=True, scale=scale, out=out, colsum=db, row_count=row_count)


def _colsum(x, out=None, **kw):
    return K.colsum(x, out=out, **kw)


def _as2d(x):
    return x.reshape(-1, x.shape[-1])


def _to_bf16_2d(x):
    x2 = _as2d(x)
    if x2.dtype == F32:
        return K.cast_bf16(x2)
    if x2.dtype != BF16:
        raise RuntimeError("fiber_b200: activations must be bf16 or fp32, got %s" % x2.dtype)
    return x2 if x2.is_contiguous() else x2.contiguous()


# ---------------------------------------------------------------------------------------------
# generic Linear / LayerNorm (tail of infer(): cross_modal_*_transform, poolers, vit_model.norm)
# ---------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, out_fp32):
        x2 = _to_bf16_2d(x)
        w, wt = CACHE.weights((weight,))
        y = K.gemm(x2, w, bias=None if bias is None else bias.detach(),
                   out_dtype=F32 if out_fp32 else BF16)
        ctx.saved = (x2, wt, bias is not None, x.shape, x.dtype)
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, wt, has_bias, xshape, xdtype = ctx.saved
        dy2 = _to_bf16_2d(dy)
        db = torch.zeros(dy2.shape[1], device=dy2.device, dtype=F32) if has_bias else None
        dw = _wgrad(dy2, x2, db=db)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = K.gemm(dy2, wt, out_dtype=F32 if xdtype == F32 else BF16).view(xshape)
        return dx, dw, db, None


class VocabDecoderFn(torch.autograd.Function):
    """MLM decoder (heads.py:40-43): logits = x W^T + b over the 50265-word vocabulary, fp32 out.
    The vocabulary is not a multiple of 8 (TMA rows are 16 bytes), so the cached bf16 weight copy carries
    zero rows up to the next multiple of 32 (the padding MlmDecoderCEFn needs: both share the cached copy) and the
    returned logits are the [..., :V] view of the padded buffer."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        V = weight.shape[0]
        Vp = (V + 31) // 32 * 32
        x2 = _to_bf16_2d(x)
        w, wt = CACHE.weights((weight,), pad_n=Vp)
        bp = torch.zeros(Vp, device=x2.device, dtype=F32)
        bp[:V] = bias.detach()
        y = K.gemm(x2, w, bias=bp, out_dtype=F32)
        ctx.saved = (x2, wt, V, Vp, x.shape, x.dtype)
        return y.view(*x.shape[:-1], Vp)[..., :V]

    @staticmethod
    def backward(ctx, dy):
        x2, wt, V, Vp, xshape, xdtype = ctx.saved
        dyp = torch.zeros((x2.shape[0], Vp), device=x2.device, dtype=BF16)
        dyp[:, :V] = dy.reshape(-1, V)
        dbp = torch.zeros(Vp, device=x2.device, dtype=F32)
        dw = _wgrad(dyp, x2, db=dbp)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = K.gemm(dyp, wt, out_dtype=F32 if xdtype == F32 else BF16).view(xshape)
        return dx, dw[:V], dbp[:V]


class MlmDecoderCEFn(torch.autograd.Function):
    """MLM decoder + cross-entropy + arg-max in one pass (heads.py:40-43, objectives.py:19-26, my_metrics.py Accuracy):
    `F.cross_entropy(h W^T + b, labels, ignore_index=-100)` and `(h W^T + b).argmax(-1)` without the [B L, 50265] fp32
    logits ever reaching HBM.

    Only labelled rows contribute to the loss, its gradient and the accuracy, so the rows are permuted labelled-first (a
    stable device-side argsort, no host sync) and the kernels take the labelled count as a device scalar: row tiles past
    it are skipped.  Forward: tcgen05 GEMM whose epilogue reduces each accumulator tile to online-softmax partials +
    a combine kernel.  Backward: the GEMM is recomputed with an epilogue that emits bf16 d(logits) for the labelled
    row tiles; dX, dW and db are the usual dgrad / wgrad GEMMs on it.  Returns (loss, predictions [rows], 0 on ignored rows)."""

    @staticmethod
    def forward(ctx, h, weight, bias, labels):
        V = weight.shape[0]
        Vp = (V + 31) // 32 * 32
        x2 = _to_bf16_2d(h)
        lab = labels.reshape(-1)
        valid = lab >= 0
        order = torch.argsort((~valid).to(torch.uint8), stable=True)  # labelled rows first, original order within
        n_valid = valid.sum(dtype=torch.int32).reshape(1)
        xs = x2.index_select(0, order)
        ls = lab.index_select(0, order).to(torch.int32)
        w, wt = CACHE.weights((weight,), pad_n=Vp)
        bp = torch.full((Vp,), -1e30, device=x2.device, dtype=F32)  # pad columns vanish from the soft-max
        bp[:V] = bias.detach()
        lse, loss_rows, pred_s = K.mlm_ce_fwd(xs, w, bp, ls, row_count=n_valid)
        loss = (loss_rows.sum() / n_valid.to(F32)).reshape(())  # 0 / 0 = nan for an all-ignored batch, as F.cross_entropy
        pred = torch.zeros(x2.shape[0], device=x2.device, dtype=torch.int64)
        pred.index_copy_(0, order, pred_s.to(torch.int64))
        ctx.saved = (xs, ls, w, wt, bp, lse, n_valid, order, V, h.shape, h.dtype)
        ctx.mark_non_differentiable(pred)
        return loss, pred.view(labels.shape)

    @staticmethod
    def backward(ctx, dloss, _dpred):
        xs, ls, w, wt, bp, lse, n_valid, order, V, hshape, hdtype = ctx.saved
        g = (dloss.to(F32) / n_valid.to(F32)).reshape(1).contiguous()
        dl = K.mlm_ce_bwd(xs, w, bp, ls, lse, g, row_count=n_valid)
        dbp = torch.zeros(bp.shape[0], device=xs.device, dtype=F32)
        dw = _wgrad(dl, xs, db=dbp, row_count=n_valid)
        dx = None
        if ctx.needs_input_grad[0]:
            # [labelled rows, 768] over K = 50272: three row tiles x three column tiles would leave 139 SMs idle for 786
            # k-blocks (308 us in ncu) — split-K with fp32 atomics instead (9 tiles x 16 splits)
            dxs = K.gemm(dl, wt, accumulate=True, row_count=n_valid)
            dx = torch.empty(xs.shape, device=xs.device, dtype=F32 if hdtype == F32 else BF16)
            dx.index_copy_(0, order, dxs.to(dx.dtype))
            dx = dx.view(hshape)
        return dx, dw[:V], dbp[:V], None


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        x2 = _to_bf16_2d(x)
        y, mean, rstd, _ = K.layernorm_fwd(x2, weight.detach(), bias.detach(), eps)
        ctx.saved = (x2, mean, rstd, weight.detach(), x.shape)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, mean, rstd, g, xshape = ctx.saved
        dg, db = torch.zeros_like(g), torch.zeros_like(g)
        dx = K.layernorm_bwd(_to_bf16_2d(dy), x2, mean, rstd, g, dgamma=dg, dbeta=db)
        return dx.view(xshape), dg, db, None


# ---------------------------------------------------------------------------------------------
# PatchEmbed: conv 4x4/s4 as a K=48(->64) GEMM + LayerNorm
# ---------------------------------------------------------------------------------------------
class PatchEmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, proj_w, proj_b, norm_w, norm_b):
        B, _, RH, RW = img.shape
        patches = K.patch_gather(img.float())
        w, _ = CACHE.weights((proj_w,), need_t=False, pad_k=64)
        e = K.gemm(patches, w, bias=proj_b.detach())
        x, mean, rstd, _ = K.layernorm_fwd(e, norm_w.detach(), norm_b.detach(), LN_EPS)
        ctx.saved = (patches, e, mean, rstd, norm_w.detach(), proj_w.shape)
        return x.view(B, (RH // 4) * (RW // 4), proj_w.shape[0])

    @staticmethod
    def backward(ctx, dx):
        patches, e, mean, rstd, g, wshape = ctx.saved
        dg, db = torch.zeros_like(g), torch.zeros_like(g)
        de = K.layernorm_bwd(_to_bf16_2d(dx), e, mean, rstd, g, dgamma=dg, dbeta=db)
        dw = _wgrad(de, patches)[:, :48].reshape(wshape)
        return None, dw, K.colsum(de), dg, db


# ---------------------------------------------------------------------------------------------
# PatchMerging: gather+LN fused kernel, then the 4C->2C reduction GEMM (no bias)
# ---------------------------------------------------------------------------------------------
class PatchMergeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, norm_w, norm_b, red_w, H, W):
        B, T, C = x.shape
        x2 = _to_bf16_2d(x)
        y, mean, rstd, _ = K.layernorm_fwd(x2, norm_w.detach(), norm_b.detach(), LN_EPS, merge=(B, H, W))
        w, wt = CACHE.weights((red_w,))
        out = K.gemm(y, w)
        ctx.saved = (x2, y, mean, rstd, norm_w.detach(), wt, (B, H, W), x.shape)
        return out.view(B, T // 4, 2 * C)

    @staticmethod
    def backward(ctx, dout):
        x2, y, mean, rstd, g, wt, merge, xshape = ctx.saved
        d2 = _to_bf16_2d(dout)
        dw = _wgrad(d2, y)
        dy = K.gemm(d2, wt)
        dg, db = torch.zeros_like(g), torch.zeros_like(g)
        dx = K.layernorm_bwd(dy, x2, mean, rstd, g, merge=merge, dgamma=dg, dbeta=db)
        return dx.view(xshape), dg, db, dw, None, None


# ---------------------------------------------------------------------------------------------
# Swin block
# ---------------------------------------------------------------------------------------------
SWIN_PLAIN = ("norm1.weight", "norm1.bias", "attn.relative_position_bias_table", "attn.qkv.weight",
              "attn.qkv.bias", "attn.proj.weight", "attn.proj.bias", "norm2.weight", "norm2.bias",
              "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias")
SWIN_FUSED = SWIN_PLAIN + ("attn.qkv_text_i2t.weight", "attn.qkv_text_i2t.bias", "attn.qkv_i2t.weight",
                           "attn.qkv_i2t.bias", "attn.proj_i2t.weight", "attn.proj_i2t.bias", "attn.alpha_i2t",
                           "attn.norm_i2t_i.weight", "attn.norm_i2t_i.bias")


class SwinBlockFn(torch.autograd.Function):
    """forward(x [B,T,C], text [B,L,Ct] | None, text_mask [B,L] f32 | None,
               s_attn [B] | None, s_mlp [B] | None   (the two independent DropPath draws of :390-391),
               geom=(H, W, ws, shift, heads), *params in SWIN_PLAIN / SWIN_FUSED order)"""

    @staticmethod
    def forward(ctx, x, text, text_mask, s_attn, s_mlp, geom, *params):
        H, W, ws, shift, nh = geom
        B, T, C = x.shape
        hd = C // nh
        fused = text is not None
        names = SWIN_FUSED if fused else SWIN_PLAIN
        p = dict(zip(names, (t.detach() for t in params)))
        x2 = _to_bf16_2d(x)
        s = s_attn
        scale = hd ** -0.5
        win = (B, H, W, ws, shift)
        sv = {}

        ln1, sv["mean1"], sv["rstd1"], _ = K.layernorm_fwd(x2, p["norm1.weight"], p["norm1.bias"], LN_EPS)
        wqkv, sv["wqkv_t"] = CACHE.weights((params[names.index("attn.qkv.weight")],))
        qkv = K.gemm(ln1, wqkv, bias=p["attn.qkv.bias"])
        table = p["attn.relative_position_bias_table"]
        ao, lse = K.attn_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], nh, hd, scale, window=win, bias_table=table)
        wproj, sv["wproj_t"] = CACHE.weights((params[names.index("attn.proj.weight")],))
        if not fused:
            x1 = K.gemm(ao, wproj, bias=p["attn.proj.bias"], residual=x2, row_scale=s, rows_per_scale=T)
        else:
            t2 = _to_bf16_2d(text)
            L = text.shape[1]
            z = torch.empty_like(x2)
            xz = K.gemm(ao, wproj, bias=p["attn.proj.bias"], preact=z, residual=x2, row_scale=s, rows_per_scale=T)
            lnz, sv["meanz"], sv["rstdz"], _ = K.layernorm_fwd(z, p["attn.norm_i2t_i.weight"],
                                                               p["attn.norm_i2t_i.bias"], LN_EPS)
            wq2, sv["wq2_t"] = CACHE.weights((params[names.index("attn.qkv_i2t.weight")],))
            q2 = K.gemm(lnz, wq2, bias=p["attn.qkv_i2t.bias"])
            wkvt, sv["wkvt_t"] = CACHE.weights((params[names.index("attn.qkv_text_i2t.weight")],))
            kvt = K.gemm(t2, wkvt, bias=p["attn.qkv_text_i2t.bias"])
            km = None if text_mask is None else text_mask.contiguous()
            ao2, lse2 = K.attn_fwd(q2, kvt[:, :C], kvt[:, C:], nh, hd, scale, groups=B, lq=T, lk=L, key_mask=km)
            wp2, sv["wp2_t"] = CACHE.weights((params[names.index("attn.proj_i2t.weight")],))
            y = torch.empty_like(x2)
            x1 = K.gemm(ao2, wp2, bias=p["attn.proj_i2t.bias"], preact=y, scale=p["attn.alpha_i2t"], residual=xz,
                        row_scale=s, rows_per_scale=T)
            sv.update(t2=t2, z=z, lnz=lnz, q2=q2, kvt=kvt, ao2=ao2, lse2=lse2, y=y, km=km, L=L)
        ln2, sv["mean2"], sv["rstd2"], _ = K.layernorm_fwd(x1, p["norm2.weight"], p["norm2.bias"], LN_EPS)
        w1, sv["w1_t"] = CACHE.weights((params[names.index("mlp.fc1.weight")],))
        h = torch.empty((B * T, 4 * C), device=x.device, dtype=BF16)
        a, sv["gelu_cached"] = _fc1_gelu(ln2, w1, p["mlp.fc1.bias"], h)
        w2, sv["w2_t"] = CACHE.weights((params[names.index("mlp.fc2.weight")],))
        out = K.gemm(a, w2, bias=p["mlp.fc2.bias"], residual=x1, row_scale=s_mlp, rows_per_scale=T)
        sv.update(x2=x2, ln1=ln1, qkv=qkv, ao=ao, lse=lse, x1=x1, ln2=ln2, h=h, a=a, p=p, s=s, s_mlp=s_mlp, geom=geom,
                  fused=fused, shape=(B, T, C), text_shape=None if text is None else text.shape)
        ctx.sv = sv
        return out.view(B, T, C)

    @staticmethod
    def backward(ctx, dout):
        sv = ctx.sv
        p, s = sv["p"], sv["s"]
        B, T, C = sv["shape"]
        H, W, ws, shift, nh = sv["geom"]
        hd = C // nh
        scale = hd ** -0.5
        fused = sv["fused"]
        names = SWIN_FUSED if fused else SWIN_PLAIN
        g = _GradArena({n: tuple(p[n].shape) for n in names}, dout.device)
        d_out = _to_bf16_2d(dout)

        # ---- MLP branch: out = x1 + s * fc2(gelu(fc1(LN2(x1)))) ----
        dz = K.scale_rows(d_out, sv["s_mlp"], T) if sv["s_mlp"] is not None else d_out
        _wgrad(dz, sv["a"], out=g["mlp.fc2.weight"], db=g["mlp.fc2.bias"])
        dh = _fc2_dgrad_gelu(dz, sv["w2_t"], sv["h"], sv["gelu_cached"])
        _wgrad(dh, sv["ln2"], out=g["mlp.fc1.weight"], db=g["mlp.fc1.bias"])
        dln2 = K.gemm(dh, sv["w1_t"])
        # ---- attention branch: x1 = x + s * (z [+ alpha * y]); the LN backward also emits dZ = s * dx1 ----
        if s is not None:
            dx1, dZ = K.layernorm_bwd(dln2, sv["x1"], sv["mean2"], sv["rstd2"], p["norm2.weight"], dres=d_out,
                                      dgamma=g["norm2.weight"], dbeta=g["norm2.bias"], row_scale=s, rows_per_scale=T)
        else:
            dx1 = K.layernorm_bwd(dln2, sv["x1"], sv["mean2"], sv["rstd2"], p["norm2.weight"], dres=d_out,
                                  dgamma=g["norm2.weight"], dbeta=g["norm2.bias"])
            dZ = dx1
        dtext = None
        if fused:
            alpha = p["attn.alpha_i2t"]
            K.dot(dZ, sv["y"], out=g["attn.alpha_i2t"])
            _wgrad(dZ, sv["ao2"], scale=alpha, out=g["attn.proj_i2t.weight"], db=g["attn.proj_i2t.bias"])
            dao2 = K.gemm(dZ, sv["wp2_t"], scale=alpha)
            dq2 = torch.empty_like(sv["q2"])
            dkvt = torch.empty_like(sv["kvt"])
            K.attn_bwd(dao2, sv["q2"], sv["kvt"][:, :C], sv["kvt"][:, C:], sv["ao2"], sv["lse2"], nh, hd, scale,
                       dq2, dkvt[:, :C], dkvt[:, C:], groups=B, lq=T, lk=sv["L"], key_mask=sv["km"])
            _wgrad(dkvt, sv["t2"], out=g["attn.qkv_text_i2t.weight"], db=g["attn.qkv_text_i2t.bias"])
            if ctx.needs_input_grad[1]:
                dtext = K.gemm(dkvt, sv["wkvt_t"]).view(sv["text_shape"])
            _wgrad(dq2, sv["lnz"], out=g["attn.qkv_i2t.weight"], db=g["attn.qkv_i2t.bias"])
            dlnz = K.gemm(dq2, sv["wq2_t"])
            dzz = K.layernorm_bwd(dlnz, sv["z"], sv["meanz"], sv["rstdz"], p["attn.norm_i2t_i.weight"], dres=dZ,
                                  dgamma=g["attn.norm_i2t_i.weight"], dbeta=g["attn.norm_i2t_i.bias"])
        else:
            dzz = dZ
        _wgrad(dzz, sv["ao"], out=g["attn.proj.weight"], db=g["attn.proj.bias"])
        dao = K.gemm(dzz, sv["wproj_t"])
        dqkv = torch.empty_like(sv["qkv"])
        qkv = sv["qkv"]
        K.attn_bwd(dao, qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], sv["ao"], sv["lse"], nh, hd, scale,
                   dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:],
                   dbias_table=g["attn.relative_position_bias_table"], window=(B, H, W, ws, shift),
                   bias_table=p["attn.relative_position_bias_table"])
        _wgrad(dqkv, sv["ln1"], out=g["attn.qkv.weight"], db=g["attn.qkv.bias"])
        dln1 = K.gemm(dqkv, sv["wqkv_t"])
        dx = K.layernorm_bwd(dln1, sv["x2"], sv["mean1"], sv["rstd1"], p["norm1.weight"], dres=dx1,
                             dgamma=g["norm1.weight"], dbeta=g["norm1.bias"])
        ctx.sv = None
        return (dx.view(B, T, C), dtext, None, None, None, None) + tuple(g[n] for n in names)


# ---------------------------------------------------------------------------------------------
# Composable pieces (fine-grained fused backbone, modules/fusion_swin_fg.py): the same kernels as SwinBlockFn behind
# one autograd Function each, so that padding / cropping of ragged window grids can sit between them as torch ops
# ---------------------------------------------------------------------------------------------
class WindowAttnFn(torch.autograd.Function):
    """W-MSA / SW-MSA on image-ordered tokens: qkv [B*H*W, 3C] bf16 -> [B*H*W, C]; H and W multiples of the window
    (fusion_swin_transformer_v2.py:148-186 with the roll / partition / reverse of :316-338 as index math)."""

    @staticmethod
    def forward(ctx, qkv, table, geom, nh):
        B, H, W, ws, shift = geom
        C = qkv.shape[1] // 3
        hd = C // nh
        scale = hd ** -0.5
        tab = table.detach()
        q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
        o, lse = K.attn_fwd(q, k, v, nh, hd, scale, window=(B, H, W, ws, shift), bias_table=tab)
        ctx.saved = (qkv, o, lse, tab, geom, nh)
        return o

    @staticmethod
    def backward(ctx, do):
        qkv, o, lse, tab, geom, nh = ctx.saved
        C = qkv.shape[1] // 3
        hd = C // nh
        dqkv = torch.empty_like(qkv)
        dtab = torch.zeros_like(tab)
        K.attn_bwd(_to_bf16_2d(do), qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o, lse, nh, hd, hd ** -0.5,
                   dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], dbias_table=dtab, window=geom, bias_table=tab)
        ctx.saved = None  # `o` is this node's own output: drop the reference cycle node -> ctx -> o -> node now, not at the
        # next cyclic GC (the step's activations would otherwise stay allocated and every step would cudaMalloc anew)
        return dqkv, dtab, None, None


class CrossAttnFn(torch.autograd.Function):
    """Plain attention of q [G*Lq, C] against packed kv [G*Lk, 2C] with an additive key mask [G, Lk] (image -> text,
    fusion_swin_transformer_v2.py:188-222)."""

    @staticmethod
    def forward(ctx, q, kv, key_mask, G, Lq, Lk, nh):
        C = q.shape[1]
        hd = C // nh
        km = None if key_mask is None else key_mask.contiguous()
        o, lse = K.attn_fwd(q, kv[:, :C], kv[:, C:], nh, hd, hd ** -0.5, groups=G, lq=Lq, lk=Lk, key_mask=km)
        ctx.saved = (q, kv, o, lse, km, G, Lq, Lk, nh)
        return o

    @staticmethod
    def backward(ctx, do):
        q, kv, o, lse, km, G, Lq, Lk, nh = ctx.saved
        C = q.shape[1]
        hd = C // nh
        dq, dkv = torch.empty_like(q), torch.empty_like(kv)
        K.attn_bwd(_to_bf16_2d(do), q, kv[:, :C], kv[:, C:], o, lse, nh, hd, hd ** -0.5, dq, dkv[:, :C], dkv[:, C:],
                   groups=G, lq=Lq, lk=Lk, key_mask=km)
        ctx.saved = None  # see WindowAttnFn
        return dq, dkv, None, None, None, None, None


class MlpFn(torch.autograd.Function):
    """fc2(GELU(fc1(x))) with the GELU (and the GELU' cache) in the fc1 epilogue and GELU' in the fc2-dgrad epilogue."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        x2 = _to_bf16_2d(x)
        w1b, w1t = CACHE.weights((w1,))
        h = torch.empty((x2.shape[0], w1.shape[0]), device=x2.device, dtype=BF16)
        a, cached = _fc1_gelu(x2, w1b, b1.detach(), h)
        w2b, w2t = CACHE.weights((w2,))
        out = K.gemm(a, w2b, bias=b2.detach())
        ctx.saved = (x2, h, a, cached, w1t, w2t, x.shape)
        return out.view(*x.shape[:-1], w2.shape[0])

    @staticmethod
    def backward(ctx, dout):
        x2, h, a, cached, w1t, w2t, xshape = ctx.saved
        dz = _to_bf16_2d(dout)
        db2 = torch.zeros(dz.shape[1], device=dz.device, dtype=F32)
        dw2 = _wgrad(dz, a, db=db2)
        dh = _fc2_dgrad_gelu(dz, w2t, h, cached)
        db1 = torch.zeros(dh.shape[1], device=dz.device, dtype=F32)
        dw1 = _wgrad(dh, x2, db=db1)
        dx = K.gemm(dh, w1t).view(xshape)
        return dx, dw1, db1, dw2, db2


class CropScaleAddFn(torch.autograd.Function):
    """out[b, h, w] = x[b, h, w] + s[b] * z[b, h, w] for h < H, w < W with z on a (zero-padded) Hp x Wp grid: the crop,
    DropPath scale and residual add of the fine-grained Swin block (fusion_swin_transformer_v2.py:336-343) in one pass;
    backward: dx = dout, dz = zero-padded s * dout in one pass."""

    @staticmethod
    def forward(ctx, x, z, s, hw, hw_padded):
        (H, W), (Hp, Wp) = hw, hw_padded
        B, T, C = x.shape
        out = K.grid_copy(z.reshape(B, Hp, Wp, C).contiguous(), (H, W), row_scale=s, add=_to_bf16_2d(x).view(B, H, W, C))
        ctx.saved = (s, hw, hw_padded, z.shape)
        return out.view(B, T, C)

    @staticmethod
    def backward(ctx, dout):
        s, (H, W), (Hp, Wp), zshape = ctx.saved
        B, T, C = dout.shape
        d = dout.to(BF16).contiguous().view(B, H, W, C)
        dz = K.grid_copy(d, (Hp, Wp), row_scale=s)
        return dout, dz.view(zshape), None, None, None


class MlpResidualFn(torch.autograd.Function):
    """The MLP half of a Swin block as one node: out = x + s * fc2(GELU(fc1(LN(x)))) with the residual and the DropPath
    row scale in the fc2 epilogue and the residual gradient inside the LayerNorm backward — the second half of
    SwinBlockFn, for callers whose attention half needs torch ops in between (padded window grids of the fine-grained
    backbone, fusion_swin_transformer_v2.py:340-346).  x [B, T, C]; s [B] per-sample scale (0 or 1 / keep) or None."""

    @staticmethod
    def forward(ctx, x, s, n_w, n_b, eps, w1, b1, w2, b2):
        B, T, C = x.shape
        x1 = _to_bf16_2d(x)
        g = n_w.detach()
        ln, mean, rstd, _ = K.layernorm_fwd(x1, g, n_b.detach(), eps)
        w1b, w1t = CACHE.weights((w1,))
        h = torch.empty((B * T, w1.shape[0]), device=x.device, dtype=BF16)
        a, cached = _fc1_gelu(ln, w1b, b1.detach(), h)
        w2b, w2t = CACHE.weights((w2,))
        out = K.gemm(a, w2b, bias=b2.detach(), residual=x1, row_scale=s, rows_per_scale=T)
        ctx.saved = (x1, ln, mean, rstd, g, h, a, cached, w1t, w2t, s, (B, T, C))
        return out.view(B, T, C)

    @staticmethod
    def backward(ctx, dout):
        x1, ln, mean, rstd, g, h, a, cached, w1t, w2t, s, (B, T, C) = ctx.saved
        d_out = _to_bf16_2d(dout)
        dz = K.scale_rows(d_out, s, T) if s is not None else d_out
        db2 = torch.zeros(C, device=dz.device, dtype=F32)
        dw2 = _wgrad(dz, a, db=db2)
        dh = _fc2_dgrad_gelu(dz, w2t, h, cached)
        db1 = torch.zeros(dh.shape[1], device=dz.device, dtype=F32)
        dw1 = _wgrad(dh, ln, db=db1)
        dln = K.gemm(dh, w1t)
        dg, dbeta = torch.zeros_like(g), torch.zeros_like(g)
        dx = K.layernorm_bwd(dln, x1, mean, rstd, g, dres=d_out, dgamma=dg, dbeta=dbeta)
        ctx.saved = None
        return dx.view(B, T, C), None, dg, dbeta, None, dw1, db1, dw2, db2


# ---------------------------------------------------------------------------------------------
# RoBERTa
# ---------------------------------------------------------------------------------------------
class RobertaEmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, word, pos, typ, ln_w, ln_b, eps, drop_p, pad_id):
        B, L = ids.shape
        e = K.embed_gather(ids, word.detach(), pos.detach(), typ.detach(), pad_id)
        x, mean, rstd, _ = K.layernorm_fwd(e, ln_w.detach(), ln_b.detach(), eps)
        seed = None
        if drop_p > 0:
            seed = _next_seed()
            x = K.dropout(x, drop_p, seed)
        ctx.saved = (ids, e, mean, rstd, ln_w.detach(), drop_p, seed, pad_id, word.shape, pos.shape)
        return x.view(B, L, -1)

    @staticmethod
    def backward(ctx, dx):
        ids, e, mean, rstd, g, drop_p, seed, pad_id, wshape, pshape = ctx.saved
        d2 = _to_bf16_2d(dx)
        if drop_p > 0:
            d2 = K.dropout(d2, drop_p, seed)
        dg, db = torch.zeros_like(g), torch.zeros_like(g)
        de = K.layernorm_bwd(d2, e, mean, rstd, g, dgamma=dg, dbeta=db)
        dword = torch.zeros(wshape, device=de.device, dtype=F32)
        dpos = torch.zeros(pshape, device=de.device, dtype=F32)
        K.embed_scatter(ids, de, dword, dpos, pad_id)
        dtyp = K.colsum(de).view(1, -1)
        return None, dword, dpos, dtyp, dg, db, None, None, None


ROBERTA_PLAIN = ("attention.self.query.weight", "attention.self.query.bias", "attention.self.key.weight",
                 "attention.self.key.bias", "attention.self.value.weight", "attention.self.value.bias",
                 "attention.output.dense.weight", "attention.output.dense.bias",
                 "attention.output.LayerNorm.weight", "attention.output.LayerNorm.bias",
                 "intermediate.dense.weight", "intermediate.dense.bias", "output.dense.weight", "output.dense.bias",
                 "output.LayerNorm.weight", "output.LayerNorm.bias")
ROBERTA_FUSED = ROBERTA_PLAIN + ("crossattention_t2i.self.query.weight", "crossattention_t2i.self.query.bias",
                                 "crossattention_t2i.self.key.weight", "crossattention_t2i.self.key.bias",
                                 "crossattention_t2i.self.value.weight", "crossattention_t2i.self.value.bias",
                                 "crossattention_t2i.output.dense.weight", "crossattention_t2i.output.dense.bias",
                                 "alpha_t2i")


class RobertaLayerFn(torch.autograd.Function):
    """forward(h [B,L,768], mask [B,L] f32 additive, image [B,T,Cimg] | None,
               meta=(heads, last_norm, eps, hidden_drop, attn_drop), *params)"""

    @staticmethod
    def forward(ctx, h, mask, image, meta, *params):
        nh, last_norm, eps, p_h, p_a = meta
        B, L, C = h.shape
        hd = C // nh
        fused = image is not None
        names = ROBERTA_FUSED if fused else ROBERTA_PLAIN
        P = dict(zip(names, params))
        p = {k: v.detach() for k, v in P.items()}
        h2 = _to_bf16_2d(h)
        sv = {}
        scale = 1.0 / math.sqrt(hd)
        km = None if mask is None else mask.contiguous()

        wqkv, sv["wqkv_t"] = CACHE.weights((P["attention.self.query.weight"], P["attention.self.key.weight"],
                                            P["attention.self.value.weight"]))
        bqkv = CACHE.bias((P["attention.self.query.bias"], P["attention.self.key.bias"],
                           P["attention.self.value.bias"]))
        qkv = K.gemm(h2, wqkv, bias=bqkv)
        sv["seed_a"] = _next_seed() if p_a > 0 else 0
        ctxv, lse = K.attn_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], nh, hd, scale, groups=B, lq=L, lk=L,
                               key_mask=km, drop_p=p_a, seed=sv["seed_a"])
        wo, sv["wo_t"] = CACHE.weights((P["attention.output.dense.weight"],))
        a = K.gemm(ctxv, wo, bias=p["attention.output.dense.bias"])
        if p_h > 0:
            sv["seed_o"] = _next_seed()
            a = K.dropout(a, p_h, sv["seed_o"])
        a2 = a
        if fused:
            img2 = _to_bf16_2d(image)
            Tk = image.shape[1]
            wq2, sv["wq2_t"] = CACHE.weights((P["crossattention_t2i.self.query.weight"],))
            q2 = K.gemm(a, wq2, bias=p["crossattention_t2i.self.query.bias"])
            wkv2, sv["wkv2_t"] = CACHE.weights((P["crossattention_t2i.self.key.weight"],
                                                P["crossattention_t2i.self.value.weight"]))
            bkv2 = CACHE.bias((P["crossattention_t2i.self.key.bias"], P["crossattention_t2i.self.value.bias"]))
            kv2 = K.gemm(img2, wkv2, bias=bkv2)
            sv["seed_a2"] = _next_seed() if p_a > 0 else 0
            ctx2, lse2 = K.attn_fwd(q2, kv2[:, :C], kv2[:, C:], nh, hd, scale, groups=B, lq=L, lk=Tk,
                                    drop_p=p_a, seed=sv["seed_a2"])
            wo2, sv["wo2_t"] = CACHE.weights((P["crossattention_t2i.output.dense.weight"],))
            c = K.gemm(ctx2, wo2, bias=p["crossattention_t2i.output.dense.bias"])
            if p_h > 0:
                sv["seed_o2"] = _next_seed()
                c = K.dropout(c, p_h, sv["seed_o2"])
            a2 = K.axpy(c, a, p["alpha_t2i"])
            sv.update(img2=img2, Tk=Tk, q2=q2, kv2=kv2, ctx2=ctx2, lse2=lse2, c=c, image_shape=image.shape)
        ln_a, sv["mean_a"], sv["rstd_a"], _ = K.layernorm_fwd(a2, p["attention.output.LayerNorm.weight"],
                                                              p["attention.output.LayerNorm.bias"], eps, add=h2)
        wi, sv["wi_t"] = CACHE.weights((P["intermediate.dense.weight"],))
        hpre = torch.empty((B * L, wi.shape[0]), device=h.device, dtype=BF16)
        inter, sv["gelu_cached"] = _fc1_gelu(ln_a, wi, p["intermediate.dense.bias"], hpre)
        wout, sv["wout_t"] = CACHE.weights((P["output.dense.weight"],))
        f = K.gemm(inter, wout, bias=p["output.dense.bias"])
        if p_h > 0:
            sv["seed_f"] = _next_seed()
            f = K.dropout(f, p_h, sv["seed_f"])
        if last_norm:
            out, sv["mean_o"], sv["rstd_o"], _ = K.layernorm_fwd(f, p["output.LayerNorm.weight"],
                                                                 p["output.LayerNorm.bias"], eps, add=ln_a)
        else:
            out = K.axpy(f, ln_a)
        sv.update(h2=h2, qkv=qkv, ctxv=ctxv, lse=lse, a=a, a2=a2, ln_a=ln_a, hpre=hpre, inter=inter, f=f, p=p, km=km,
                  meta=meta, fused=fused, shape=(B, L, C))
        ctx.sv = sv
        return out.view(B, L, C)

    @staticmethod
    def backward(ctx, dout):
        sv = ctx.sv
        p = sv["p"]
        nh, last_norm, eps, p_h, p_a = sv["meta"]
        B, L, C = sv["shape"]
        hd = C // nh
        scale = 1.0 / math.sqrt(hd)
        fused = sv["fused"]
        names = ROBERTA_FUSED if fused else ROBERTA_PLAIN
        shapes = {n: tuple(p[n].shape) for n in names if "self." not in n}
        shapes["wqkv"], shapes["bqkv"] = (3 * C, C), (3 * C,)
        if fused:
            shapes["wkv2"], shapes["bkv2"] = (2 * C, sv["img2"].shape[1]), (2 * C,)
            shapes["crossattention_t2i.self.query.weight"] = (C, C)
            shapes["crossattention_t2i.self.query.bias"] = (C,)
        g = _GradArena(shapes, dout.device)
        d_out = _to_bf16_2d(dout)
        if last_norm:
            dsum = K.layernorm_bwd(d_out, sv["f"], sv["mean_o"], sv["rstd_o"], p["output.LayerNorm.weight"],
                                   add=sv["ln_a"], dgamma=g["output.LayerNorm.weight"],
                                   dbeta=g["output.LayerNorm.bias"])
        else:
            g["output.LayerNorm.weight"] = g["output.LayerNorm.bias"] = None
            dsum = d_out
        df = K.dropout(dsum, p_h, sv["seed_f"]) if p_h > 0 else dsum
        _wgrad(df, sv["inter"], out=g["output.dense.weight"], db=g["output.dense.bias"])
        dhpre = _fc2_dgrad_gelu(df, sv["wout_t"], sv["hpre"], sv["gelu_cached"])
        _wgrad(dhpre, sv["ln_a"], out=g["intermediate.dense.weight"], db=g["intermediate.dense.bias"])
        dln_a = K.gemm(dhpre, sv["wi_t"], residual=dsum)
        ds1 = K.layernorm_bwd(dln_a, sv["a2"], sv["mean_a"], sv["rstd_a"], p["attention.output.LayerNorm.weight"],
                              add=sv["h2"], dgamma=g["attention.output.LayerNorm.weight"],
                              dbeta=g["attention.output.LayerNorm.bias"])
        dimage = None
        if fused:
            alpha = p["alpha_t2i"]
            K.dot(ds1, sv["c"], out=g["alpha_t2i"])
            gc = K.dropout(ds1, p_h, sv["seed_o2"]) if p_h > 0 else ds1
            _wgrad(gc, sv["ctx2"], scale=alpha, out=g["crossattention_t2i.output.dense.weight"],
                   db=g["crossattention_t2i.output.dense.bias"])
            dctx2 = K.gemm(gc, sv["wo2_t"], scale=alpha)
            dq2 = torch.empty_like(sv["q2"])
            dkv2 = torch.empty_like(sv["kv2"])
            K.attn_bwd(dctx2, sv["q2"], sv["kv2"][:, :C], sv["kv2"][:, C:], sv["ctx2"], sv["lse2"], nh, hd, scale,
                       dq2, dkv2[:, :C], dkv2[:, C:], groups=B, lq=L, lk=sv["Tk"], drop_p=p_a, seed=sv["seed_a2"])
            dwkv2 = _wgrad(dkv2, sv["img2"], out=g["wkv2"], db=g["bkv2"])
            dbkv2 = g["bkv2"]
            g["crossattention_t2i.self.key.weight"], g["crossattention_t2i.self.value.weight"] = dwkv2[:C], dwkv2[C:]
            g["crossattention_t2i.self.key.bias"], g["crossattention_t2i.self.value.bias"] = dbkv2[:C], dbkv2[C:]
            if ctx.needs_input_grad[2]:
                dimage = K.gemm(dkv2, sv["wkv2_t"]).view(sv["image_shape"])
            _wgrad(dq2, sv["a"], out=g["crossattention_t2i.self.query.weight"],
                   db=g["crossattention_t2i.self.query.bias"])
            da = K.gemm(dq2, sv["wq2_t"], residual=ds1)
        else:
            da = ds1
        if p_h > 0:
            da = K.dropout(da, p_h, sv["seed_o"])
        _wgrad(da, sv["ctxv"], out=g["attention.output.dense.weight"], db=g["attention.output.dense.bias"])
        dctx = K.gemm(da, sv["wo_t"])
        qkv = sv["qkv"]
        dqkv = torch.empty_like(qkv)
        K.attn_bwd(dctx, qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], sv["ctxv"], sv["lse"], nh, hd, scale,
                   dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], groups=B, lq=L, lk=L, key_mask=sv["km"],
                   drop_p=p_a, seed=sv["seed_a"])
        dwqkv = _wgrad(dqkv, sv["h2"], out=g["wqkv"], db=g["bqkv"])
        dbqkv = g["bqkv"]
        for i, n in enumerate(("query", "key", "value")):
            g["attention.self.%s.weight" % n] = dwqkv[i * C:(i + 1) * C]
            g["attention.self.%s.bias" % n] = dbqkv[i * C:(i + 1) * C]
        dh = K.gemm(dqkv, sv["wqkv_t"], residual=ds1)
        ctx.sv = None
        return (dh.view(B, L, C), None, dimage, None) + tuple(g[n] for n in names)
