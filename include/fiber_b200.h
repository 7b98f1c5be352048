/* fiber_b200 — C-ABI of the B200-native FIBER fusion-in-the-backbone hot path.
 *
 * The reference (microsoft/FIBER, coarse_grained/) has no FFI on this path: its boundary is the
 * Python object protocol of fiber.modules.FIBERTransformerSS (fiber_module.py:26-367).  This
 * header is the native boundary our Python mirror of that protocol binds with ctypes: one entry
 * point per eager op (or fused group of ops) the reference issues on the path, each citing the
 * reference lines it replaces.  Conventions for every entry point:
 *   - plain pointers + sizes; all pointers are DEVICE pointers on the current device unless said
 *     otherwise; activations are bf16 row-major, parameters fp32 masters or bf16 copies as stated;
 *   - the caller allocates every buffer; entry points never allocate device memory and never
 *     synchronise; work is enqueued on `stream`;
 *   - returns 0 on success, negative on error; fiber_last_error() gives the message (thread-local).
 */
#ifndef FIBER_B200_H_
#define FIBER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* fiber_stream_t; /* cudaStream_t */

const char* fiber_last_error(void);
int fiber_version(void);
/* Sets per-kernel attributes (max dynamic shared memory) for the current device.  Idempotent. */
int fiber_init(void);
/* Number of kernels this library has launched since load (for bench.py's gpu_launches). */
int64_t fiber_launch_count(void);
/* Process-wide kernel selection (value -1 restores the default).
 * "winattn_tc": routes the forward (bit 0) / backward (bit 1) of 12x12-window attention (fiber_attn_fwd / fiber_attn_bwd,
 * mode 1, head_dim 32, shift 0 or 6) to the tcgen05 kernels of csrc/window_attn_tc.cu instead of the mma.sync ones; bit 2 /
 * bit 3 select their fourth generation (TMA-fed quadrant-order tiles) for the forward / backward.  Default 15, or the
 * FIBER_WINATTN_TC environment variable.
 * "attn_small": bit 0 routes fiber_attn_bwd (mode 0, head_dim 64, at most 48 queries and keys: RoBERTa self-attention,
 * roberta.py:256-326) to a 3-warp / four-CTAs-per-SM configuration of the same kernel, bit 1 the few-key case
 * (head_dim 32, at most 48 keys, more than 48 queries: i2t, swin_transformer.py:226-259) to a 4-warp / three-CTAs-per-SM
 * one, bit 2 the few-query case (head_dim 64, at most 48 queries, more than 48 keys: t2i, roberta.py:441-502) to the
 * 3-warp one; default 7 or FIBER_ATTN_SMALL.
 * "attn_sk": bit 0 routes the forward, bit 1 the backward of mode-0 attention with at most 64 keys per group and at least 96
 * queries (i2t) to the tcgen05 + TMA kernels of csrc/attention_sk.cu, bits 2 / 3 the forward / backward of self-attention
 * with Lq = Lk in {32, 40, 48, 56, 64} (RoBERTa, roberta.py:256-326: two or three samples packed into a 128-row tile), bit 4
 * every other shorter query sequence; default 15 or FIBER_ATTN_SK.
 * "gemm_cta2": bit 0 runs K-major fiber_gemm launches with N > 128, M % 256 == 0, K >= 1024 and no row_count as CTA pairs
 * (2-CTA clusters, tcgen05 cta_group::2, 256 x 256 pair tiles) for the default epilogues, bit 1 for act 3 .. 7, bit 2 lowers
 * the K threshold to 256; same results bit for bit (same accumulation order); default 3, FIBER_GEMM_CTA2.
 * "pdl": 1 launches every kernel with programmatic stream serialization (prologue overlaps the previous kernel's tail);
 * default 0, FIBER_PDL.
 * "image_variant": work per thread of fiber_image_transform's two passes — bit 0 the horizontal pass reads aligned words
 * instead of bytes, bit 1 / bit 2 eight / sixteen output columns per thread in the vertical pass instead of four, bit 3 (with
 * bit 0) eight source rows per thread instead of four, bit 4 the vertical pass per 16-row band with its plane rows staged in
 * shared memory; default 3 or FIBER_IMAGE_VARIANT; the bytes out are the same.
 * "tq_trace" (debug): 1 makes the fourth-generation window backward record an event trace of one CTA (tools/tq_trace.py).
 * Results are the same attention (swin_transformer.py:195-224) either way.  Returns 0, or -1 for an unknown name;
 * fiber_get_option returns the value ("winattn_tc_launches" / "attn_sk_launches", read-only: launches of the tcgen05
 * window / small-key kernels so far). */
int fiber_set_option(const char* name, int32_t value);
int fiber_get_option(const char* name);

/* ---- GEMM on tcgen05 tensor cores --------------------------------------------------------
 * C[M,N] = epilogue(A * B^T), bf16 operands, fp32 accumulate in TMEM.
 * Replaces every nn.Linear on the path (swin_transformer.py:202,223,240,257 + timm Mlp;
 * roberta.py:266-279,338,404,418; PatchEmbed conv as a K=48 GEMM; fiber_module.py:349-350)
 * and their autograd backward (dgrad, wgrad).
 *   a_major/b_major = 0: operand stored [rows, K] with K contiguous (forward, dgrad)
 *                   = 1: operand stored [K, rows] with rows contiguous (wgrad: dW = dY^T X)
 * Epilogue, applied in this order on v = acc:
 *   v += bias[col]; if (preact) preact[row,col] = bf16(v);
 *   act == 1: v = gelu_erf(v);  act == 2: v *= gelu_erf'(aux[row,col]);
 *   if (scale) v *= *scale;  if (row_scale) v *= row_scale[row / rows_per_scale];
 *   if (residual) v += residual[row,col];
 *   out_mode 0: c = bf16(v); 1: c = f32(v); 2: atomicAdd(f32 c, v) (split-K allowed).
 * Single-pass epilogues (K-major operands, out_mode 0, M % 128 == 0, N % 32 == 0; no scale / row_scale / residual except
 * act 6; two TMA-store boxes per warp).  act 3 / act 4 are what the FIBER path uses for fc1 / fc2 dgrad:
 *   act == 3: v += bias[col]; c = bf16(gelu_erf(v)); preact[row,col] = bf16(gelu_erf'(v))   (one pass, one exponential)
 *   act == 4: c = bf16(acc * aux[row,col])          (with aux = the act-3 second output: dgrad through the GELU)
 *   act == 5: v += bias[col]; c = bf16(gelu_erf(v)); preact[row,col] = bf16(v)    (act 1 + preact, bit for bit, one pass)
 *   act == 6: the default epilogue with residual (bias, scale, row_scale as above; no preact / aux), bit for bit, residual
 *             rows register-prefetched a chunk ahead
 *   act == 7: c = bf16(acc * gelu_erf'(aux[row,col]))   (act 2, bit for bit, aux rows register-prefetched a chunk ahead)
 */
typedef struct fiber_gemm_args {
  const void* a;
  const void* b;
  void* c;
  int32_t m, n, k;
  int64_t lda, ldb, ldc; /* leading dimensions in elements */
  int32_t a_major, b_major;
  const float* bias;
  const void* residual; /* bf16 [M, N] */
  int64_t ldr;
  const void* aux; /* bf16 [M, N] */
  int64_t ldaux;
  void* preact; /* bf16 [M, N] out */
  int64_t ldp;
  const float* scale;
  const float* row_scale;
  int32_t rows_per_scale;
  int32_t act;
  int32_t out_mode;
  int32_t splits; /* 0 = choose */
  /* MN-major (wgrad) only, out_mode 1 / 2: colsum[m] += scale * sum_k A[k, m] — the bias gradient
   * db = dY^T 1 that goes with dW = dY^T X (one extra 128 x 16 MMA per k-step against a tile of ones). */
  float* colsum;
  /* Device-side row count (int32 scalar in device memory) or NULL.  Only the first *row_count rows of the activation operand
   * carry work: K-major launches skip the output row tiles past it (those rows of c are NOT written), MN-major (wgrad) launches
   * skip the reduction k-blocks past it.  Lets a caller that compacted its valid rows to the front (the labelled MLM positions)
   * launch without reading the count back to the host. */
  const int32_t* row_count;
} fiber_gemm_args;

int fiber_gemm(const fiber_gemm_args* args, fiber_stream_t stream);

/* ---- fused MLM decoder + cross-entropy -----------------------------------------------------
 * Replaces `mlm_logits = decoder(h) + bias` (heads.py:40-43) followed by `F.cross_entropy(mlm_logits.view(-1, V), labels.view(-1),
 * ignore_index=-100)` (objectives.py:19-26) and the arg-max the accuracy metric takes of the same logits (gadgets/my_metrics.py:
 * Accuracy.update): the [rows, V] logits live only as TMEM accumulators of the tcgen05 GEMM.
 *   fiber_mlm_ce_fwd: per row r < *row_count: lse[r] = log2-domain log-sum-exp of x[r] . w^T + bias, loss_rows[r] = natural-log
 *       lse - logit[r, labels[r]] (0 when labels[r] < 0), pred[r] = arg-max column (first maximum).  Rows >= *row_count get 0.
 *   fiber_mlm_ce_bwd: dlogits[r, c] = bf16((softmax(logits[r])[c] - (c == labels[r])) * *gscale) for labelled rows, 0 for ignored
 *       rows; row tiles past *row_count are not written.  dX / dW / dbias then are plain fiber_gemm calls on dlogits
 *       (with the same row_count).
 * n % 32 == 0: pad the vocabulary with zero weight rows and a large negative bias (e.g. -1e30) so pad columns vanish. */
typedef struct fiber_ce_args {
  const void* x;            /* bf16 [M, K] */
  const void* w;            /* bf16 [N, K] */
  const float* bias;        /* [N] */
  const int32_t* labels;    /* [M], < 0 = ignored row */
  const int32_t* row_count; /* device scalar or NULL (= all M rows) */
  int32_t m, n, k;
  int64_t ldx, ldw;
  float* part;        /* fwd workspace: 4 * ceil(N / 256) * (ceil(M / 128) * 128) * 4 floats */
  float* label_logit; /* fwd workspace [M] */
  float* lse;         /* [M]: fwd out, bwd in (log2 domain) */
  float* loss_rows;   /* fwd out [M] */
  int32_t* pred;      /* fwd out [M] */
  void* dlogits;      /* bwd out bf16 [M, N] */
  int64_t lddl;
  const float* gscale; /* bwd in: device scalar */
} fiber_ce_args;

int fiber_mlm_ce_fwd(const fiber_ce_args* args, fiber_stream_t stream);
int fiber_mlm_ce_bwd(const fiber_ce_args* args, fiber_stream_t stream);

/* ---- small-sequence attention (flash-style; scores never reach HBM) ------------------------
 * mode 1 (window): Swin W-MSA/SW-MSA core, swin_transformer.py:205-222 with the roll /
 *   window_partition / window_reverse of :363-387 and the bias gather of :208-217 folded into
 *   the kernel's addressing.  q/k/v point into the IMAGE-ordered packed qkv activation
 *   [G*H*W, 3C] (q = base, k = base + C, v = base + 2C, ld = 3C); o is image-ordered [G*H*W, C].
 * mode 0 (plain): RoBERTa self-attention (roberta.py:284-322), t2i cross attention (no mask) and
 *   i2t cross attention (swin_transformer.py:245-256); rows of group g are g*Lq + i / g*Lk + j;
 *   key_mask is the additive [G, Lk] fp32 mask (0 / -10000) or NULL.
 * Scores are scale * q.k (+ bias / mask); lse [G', nH, Lq] is written for the backward pass.
 * Backward recomputes P; window mode accumulates d(bias table) into dbias_table (fp32, atomics).
 */
typedef struct fiber_attn_args {
  const void* q;
  const void* k;
  const void* v;
  void* o;
  float* lse;
  int64_t ldq, ldk, ldv, ldo;
  int32_t mode, groups, heads, lq, lk, head_dim;
  float scale;
  const float* key_mask;
  int32_t h, w, ws, shift;
  const float* bias_table; /* [(2ws-1)^2, heads] fp32 */
  float drop_p;
  uint64_t seed;
  /* backward only */
  const void* d_o;
  void* dq;
  void* dk;
  void* dv;
  int64_t lddo, lddq, lddk, lddv;
  float* dbias_table;
  float* d_scratch; /* window backward: fp32 [rows, heads] workspace for rowsum(dO*O); NULL = slow path */
} fiber_attn_args;

int fiber_attn_fwd(const fiber_attn_args* args, fiber_stream_t stream);
int fiber_attn_bwd(const fiber_attn_args* args, fiber_stream_t stream);

/* ---- LayerNorm (nn.LayerNorm: swin_transformer.py:362,391,429,240; roberta.py:197,485,422) ---
 * y = LN(in1 [+ in2]) * gamma + beta over the last dimension (width c, multiple of 8, <= 2048).
 * merge != 0 fuses PatchMerging's 2x2 neighbour gather + concat (swin_transformer.py:420-427):
 *   in1 is the image-ordered [B*H*W, cin] activation, rows = B*(H/2)*(W/2), c = 4*cin.
 * Forward writes mean/rstd (fp32 per row) for the backward pass and optionally in1+in2 (sum_out).
 * Backward: dx = LN'(dy) [+ dres]; dgamma/dbeta are ACCUMULATED (fp32 atomics; caller zeroes).
 */
typedef struct fiber_ln_args {
  const void* in1;
  const void* in2;
  int64_t ld1, ld2;
  const float* gamma;
  const float* beta;
  float eps;
  void* out;
  int64_t ldo;
  float* mean;
  float* rstd;
  void* sum_out;
  int64_t lds;
  int64_t rows;
  int32_t c;
  int32_t merge, h, w, cin;
  /* backward only */
  const void* dy;
  int64_t lddy;
  const void* dres;
  int64_t lddres;
  void* dx;
  int64_t lddx;
  float* dgamma;
  float* dbeta;
  /* optional second output of the backward: dx_scaled[row,:] = dx[row,:] * row_scale[row / rows_per_scale]
   * (the DropPath scale of the branch that consumes dx next, swin_transformer.py:390) */
  const float* row_scale;
  int32_t rows_per_scale;
  void* dx_scaled;
  int64_t lddxs;
} fiber_ln_args;

int fiber_layernorm_fwd(const fiber_ln_args* args, fiber_stream_t stream);
int fiber_layernorm_bwd(const fiber_ln_args* args, fiber_stream_t stream);

/* out[n] += (*scale) * sum_m row_scale[m / rows_per_scale] * x[m, n]   (nn.Linear bias gradients) */
int fiber_colsum(const void* x, int64_t ld, int64_t m, int32_t n, float* out, const float* scale,
                 const float* row_scale, int32_t rows_per_scale, fiber_stream_t stream);
/* *out += sum_{m,n} a[m,n] * b[m,n]   (gradients of the alpha_i2t / alpha_t2i gates) */
int fiber_dot(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t m, int32_t n, float* out,
              fiber_stream_t stream);
/* y = dropout(x; p, seed): counter-based mask, identical for forward and backward (nn.Dropout,
 * roberta.py:198,339,419) */
int fiber_dropout(const void* x, int64_t ldx, void* y, int64_t ldy, int64_t m, int32_t n, float p, uint64_t seed,
                  fiber_stream_t stream);
/* y[m,:] = x[m,:] * row_scale[m / rows_per_scale]   (timm DropPath, swin_transformer.py:390-391) */
int fiber_scale_rows(const void* x, int64_t ldx, void* y, int64_t ldy, int64_t m, int32_t n, const float* row_scale,
                     int32_t rows_per_scale, fiber_stream_t stream);
/* out = add + (*alpha) * x  (alpha NULL = 1): the alpha_t2i gate of roberta.py:483 and the
 * un-normalised residual of RobertaOutput with last_norm=False (roberta.py:420-423) */
int fiber_axpy(const void* x, int64_t ldx, const void* add, int64_t ldadd, const float* alpha, void* out, int64_t ldo,
               int64_t m, int32_t n, fiber_stream_t stream);
/* Crop / zero-pad of a bf16 token grid with the DropPath scale and the residual fused (fine-grained Swin block,
 * fine_grained/maskrcnn_benchmark/modeling/backbone/fusion_swin_transformer_v2.py:316-321 F.pad, :336-343 crop + residual):
 * dst[b,h,w,:] = (h < hs && w < ws ? src[b,h,w,:] * row_scale[b] : 0) + add[b,h,w,:] for h < hd, w < wd; src [B,hs,ws,C],
 * dst / add [B,hd,wd,C] contiguous; row_scale (fp32 [B]) and add may be NULL; C % 8 == 0. */
int fiber_grid_copy(const void* src, void* dst, const void* add, const float* row_scale, int32_t batch, int32_t hs, int32_t ws,
                    int32_t hd, int32_t wd, int32_t c, fiber_stream_t stream);
int fiber_cast_f32_bf16(const float* x, void* y, int64_t n, fiber_stream_t stream);
/* fp32 master weight [n,k] -> bf16 copy [n,k] (ld_out) and/or transposed bf16 copy [k,n] (ldt_out) */
int fiber_cast_transpose(const float* w, int64_t ldw, int32_t n, int32_t k, void* w_out, int64_t ld_out, void* wt_out,
                         int64_t ldt_out, fiber_stream_t stream);
/* timm PatchEmbed im2row: img fp32 [B,3,R,R] -> bf16 [B*(R/4)^2, 64], col = c*16+kh*4+kw, cols 48..63 zero */
int fiber_patch_gather(const float* img, void* out, int32_t batch, int32_t r, fiber_stream_t stream);
/* the same for rectangular images [B,3,H,W], H % 4 == 0 and W % 4 == 0 (fine-grained model: PatchEmbed of
 * fine_grained/maskrcnn_benchmark/modeling/backbone/fusion_swin_transformer_v2.py:548-566 after its padding) */
int fiber_patch_gather_hw(const float* img, void* out, int32_t batch, int32_t h, int32_t w, fiber_stream_t stream);
/* RobertaEmbeddings sum (roberta.py:169-196): out[t] = word[ids[t]] + pos[pos_id(t)] + type[0] */
int fiber_embed_gather(const int64_t* ids, int32_t batch, int32_t len, int32_t c, int32_t pad_id, const float* word,
                       const float* pos, const float* type, void* out, int64_t ldo, fiber_stream_t stream);
/* backward of the two gathers: dword[ids[t]] += dsum[t], dpos[pos_id(t)] += dsum[t] (pad rows skipped) */
int fiber_embed_scatter(const int64_t* ids, int32_t batch, int32_t len, int32_t c, int32_t pad_id, const void* dsum,
                        int64_t ldd, float* dword, float* dpos, fiber_stream_t stream);

/* ---- fused multi-tensor AdamW ----------------------------------------------
 * Replaces transformers.AdamW.step over the six parameter groups of fiber_utils.set_schedule
 * (coarse_grained/fiber/modules/fiber_utils.py:156-252): per element, in HF 4.6 order,
 *   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr sqrt(1-b2^t)/(1-b1^t) m / (sqrt(v) + eps);  p -= lr wd p.
 * `tensors` is a DEVICE array of fiber_adamw_tensor (one per parameter; lr / wd carry its group), `chunks` a DEVICE array
 * of n_chunks (tensor index, chunk index) int32 pairs covering every tensor in pieces of chunk_elems elements
 * (a multiple of 1024).  step counts from 1.  p_bf16 (nullable) receives a bf16 copy of the updated parameter. */
typedef struct fiber_adamw_tensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  void* p_bf16;
  int64_t n;
  float lr, wd;
} fiber_adamw_tensor;
int fiber_adamw_multi(const void* tensors, const void* chunks, int32_t n_chunks, int32_t chunk_elems, float beta1, float beta2,
                      float eps, int32_t step, fiber_stream_t stream);

/* ---- image side of the input pipeline --------------------------------------
 * Replaces, per batch instead of per image in DataLoader workers, what the reference's `albef` transform and the
 * geometric part of `albef_randaug` do to a decoded RGB image (coarse_grained/fiber/transforms/transform.py:10-17 and
 * :20-24,42-44; called from datasets/base_dataset.py:93-110 get_raw_image / get_image / get_false_image):
 *   [crop box -> ] PIL Image.resize((out_w, out_h), BICUBIC) [-> horizontal flip] -> ToTensor -> Normalize(mean, std)
 * with the libraries' own arithmetic, bit for bit: Pillow's 8-bit two-pass resampling (horizontal pass first, 22-bit
 * fixed-point coefficients evaluated in double precision, each pass rounded and clamped to a byte) and torchvision's
 * float32 (v / 255 - mean) / std.  JPEG decoding and RandomAugment's photometric / affine operations stay with the caller.
 *
 * One fiber_image_desc per image.  The caller fills src .. flip, planar and chan_stride; fiber_image_transform_plan (host only, no CUDA call)
 * fills the rest and returns the workspace size in bytes (0 on error).  Images are ragged: every descriptor has its own
 * size, row stride and crop box.  Pillow >= 11 resamples the vertical axis first when h > 100 w; such boxes are rejected. */
typedef struct fiber_image_desc {
  const uint8_t* src;                  /* DEVICE pointer: row 0, column 0 (channel 0) of the decoded image */
  int64_t stride;                      /* bytes between source rows (>= 3 w interleaved, >= w planar) */
  int32_t h, w;                        /* decoded size */
  int32_t box_x, box_y, box_w, box_h;  /* crop (torchvision resized_crop: left, top, width, height); whole image = 0,0,w,h */
  int32_t flip;                        /* 1: RandomHorizontalFlip applied after the resize */
  int32_t ksize_x, ksize_y;            /* plan: taps per output sample, horizontal / vertical */
  int32_t planar;                      /* 0: RGB bytes interleaved (PIL / numpy HWC); 1: three byte planes (CHW, what nvJPEG decoders return) */
  int64_t coef_off;                    /* plan: byte offset of this image's coefficient tables in the workspace */
  int64_t tmp_off;                     /* plan: byte offset of its horizontally resampled byte planes [3][box_h][pitch] */
  int64_t chan_stride;                 /* planar only: bytes between channel planes (>= h stride) */
} fiber_image_desc;
size_t fiber_image_transform_plan(fiber_image_desc* descs_host, int32_t n, int32_t out_h, int32_t out_w);
/* descs_host: the planned descriptors (read on the host for grid sizes); descs_dev: a DEVICE copy of the same array;
 * mean / std: HOST float[3]; out: DEVICE float32 [n, 3, out_h, out_w] (16-byte aligned, out_w % 4 == 0);
 * workspace: DEVICE, at least the planned size, 16-byte aligned.  Three launches: coefficient tables, horizontal pass
 * (bytes -> byte planes), vertical pass + normalisation. */
int fiber_image_transform(const fiber_image_desc* descs_host, const fiber_image_desc* descs_dev, int32_t n, int32_t out_h,
                          int32_t out_w, const float* mean, const float* std, void* workspace, size_t workspace_bytes,
                          float* out, fiber_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FIBER_B200_H_ */
