// fiber_b200 — extern "C" surface (see include/fiber_b200.h) and process-wide plumbing.
#include "attention.cuh"
#include "../../include/fiber_b200.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace fiber {
void set_image_variant(int v);  // image_pipeline.cu
int get_image_variant();

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// "winattn_tc": bit 0 = tcgen05 window-attention forward, bit 1 = backward (window_attn_tc.cu) for the geometries it
// covers (12x12 windows, shift 0 / 6); 0 = the mma.sync generation everywhere; bit 2 / bit 3 select the fourth
// generation (TMA-fed, quadrant token order) for the forward / backward.  Default 3 (validated on B200 in round 2:
// gpurun_out/r2a_*), overridable by FIBER_WINATTN_TC; fiber_set_option(name, -1) returns to the default.
static std::atomic<int> g_winattn_tc{-1};
// "attn_small": bit 0 routes plain attention backward with <= 48 queries and keys (head_dim 64) to the 3-warp
// configuration of attention_bwd.cu, bit 1 the <= 48-key / many-query case (head_dim 32) to the 4-warp one, bit 2 the
// <= 48-query / many-key case (head_dim 64, t2i) to the 3-warp one.
// Default 7 (validated on B200 in round 2), overridable by FIBER_ATTN_SMALL.
static std::atomic<int> g_attn_small{-1};
int option_attn_small() {
  int v = g_attn_small.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("FIBER_ATTN_SMALL");
    v = e ? (atoi(e) & 7) : 7;
    g_attn_small.store(v, std::memory_order_relaxed);
  }
  return v;
}
// "attn_sk": bit 0 routes the forward, bit 1 the backward of plain attention with at most 64 keys per group and at least 96
// queries (i2t) to the tcgen05 + TMA kernels of attention_sk.cu; bits 2 / 3 the forward / backward of self-attention shapes
// that pack two or three samples into a 128-row tile (RoBERTa, Lq = Lk in {32, 40, 48, 56, 64}); bit 4 every other short
// query sequence, unpacked (slower than the mma.sync kernels; tests).  Default 15 (validated on B200 in round 2: i2t stage 2
// forward 0.255 -> 0.158 ms, backward 0.407 -> 0.317 ms; text self-attention forward 0.042 = 0.042 ms, backward
// 0.089 -> 0.071 ms), or FIBER_ATTN_SK.
static std::atomic<int> g_attn_sk{-1};
int option_attn_sk() {
  int v = g_attn_sk.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("FIBER_ATTN_SK");
    v = e ? (atoi(e) & 31) : 15;
    g_attn_sk.store(v, std::memory_order_relaxed);
  }
  return v;
}
static std::atomic<int> g_attn_sk_launches{0};
void count_attn_sk_launch() { g_attn_sk_launches.fetch_add(1, std::memory_order_relaxed); }
static std::atomic<int> g_winattn_tc_launches{0};  // launches of the tcgen05 generation (tests check the routing)
void count_winattn_tc_launch() { g_winattn_tc_launches.fetch_add(1, std::memory_order_relaxed); }
int option_winattn_tc() {
  int v = g_winattn_tc.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("FIBER_WINATTN_TC");
    v = e ? (atoi(e) & 15) : 15;
    g_winattn_tc.store(v, std::memory_order_relaxed);
  }
  return v;
}

// "pdl": 1 = every launch carries the programmatic-stream-serialization attribute (common.cuh: launch_k), so a kernel's
// prologue overlaps the previous kernel's tail; 0 = plain stream order.  Default 0, or FIBER_PDL: measured neutral on the
// power-capped training step (profiles/r2_pdl_ab.txt: 207.3 / 210.4 ms with, 206.3 / 206.4 ms without), full GPU suite green
// either way.
static std::atomic<int> g_pdl{-1};
int option_pdl() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("FIBER_PDL");
    v = e ? (atoi(e) != 0) : 0;
    g_pdl.store(v, std::memory_order_relaxed);
  }
  return v;
}

// "gemm_cta2": bit 0 runs K-major GEMMs with 256-column tiles and M % 256 == 0 as CTA pairs (2-CTA clusters, tcgen05
// cta_group::2: 256 x 256 pair tiles, half of the B tile per CTA) for the default epilogues, bit 1 for the two-box epilogues
// (act 3 .. 7), in both cases only when K >= 1024: measured on B200 (profiles/r2_gemm_cta_pairs.txt) +3 .. +9 % at K >= 1024
// and -3 .. -20 % at K <= 512, where a tile is four to eight k-blocks and the pair's longer hand-offs (remote arrives,
// multicast commits) show.  Default 3, or FIBER_GEMM_CTA2; bit 2 lifts the K >= 1024 rule (tests, microbenchmarks); bit 3
// runs wgrad launches (MN-major operands, split-K, fused column sums) as pairs too — correct (test_gemm_cta_pairs_wgrad) but
// 2 - 4 % slower than single CTAs on every FIBER shape, so off by default.
static std::atomic<int> g_gemm_cta2{-1};
static std::atomic<int> g_gemm_cta2_launches{0};
void count_gemm_cta2_launch() { g_gemm_cta2_launches.fetch_add(1, std::memory_order_relaxed); }
int option_gemm_cta2() {
  int v = g_gemm_cta2.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("FIBER_GEMM_CTA2");
    v = e ? (atoi(e) & 15) : 3;
    g_gemm_cta2.store(v, std::memory_order_relaxed);
  }
  return v;
}

static std::atomic<int> g_tq_trace{0};  // debug: event trace of the fourth-generation window backward (tools/tq_trace.py)
int option_tq_trace() { return g_tq_trace.load(std::memory_order_relaxed); }

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

int gemm_dispatch(const fiber_gemm_args* a, cudaStream_t stream);
int mlm_ce_dispatch(const fiber_ce_args* a, int backward, cudaStream_t stream);
int attn_fwd_dispatch(const AttnParams& p, int hd, cudaStream_t stream);
int attn_bwd_dispatch(const AttnParams& p, int hd, float* d_scratch, cudaStream_t stream);

static AttnParams to_params(const fiber_attn_args* a) {
  AttnParams p;
  p.q = reinterpret_cast<const bf16*>(a->q); p.k = reinterpret_cast<const bf16*>(a->k);
  p.v = reinterpret_cast<const bf16*>(a->v); p.o = reinterpret_cast<bf16*>(a->o);
  p.lse = a->lse;
  p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv; p.ldo = a->ldo;
  p.mode = a->mode; p.G = a->groups; p.nH = a->heads; p.Lq = a->lq; p.Lk = a->lk;
  p.scale = a->scale; p.key_mask = a->key_mask;
  p.H = a->h; p.W = a->w; p.ws = a->ws; p.shift = a->shift; p.bias_table = a->bias_table;
  p.drop_p = a->drop_p; p.seed = a->seed;
  p.d_o = reinterpret_cast<const bf16*>(a->d_o); p.dq = reinterpret_cast<bf16*>(a->dq);
  p.dk = reinterpret_cast<bf16*>(a->dk); p.dv = reinterpret_cast<bf16*>(a->dv);
  p.lddo = a->lddo; p.lddq = a->lddq; p.lddk = a->lddk; p.lddv = a->lddv;
  p.dbias_table = a->dbias_table;
  return p;
}

struct LnParams {
  const bf16* in1; const bf16* in2; long long ld1, ld2; const float* gamma; const float* beta; float eps;
  bf16* out; long long ldo; float* mean; float* rstd; bf16* sum_out; long long lds; long long rows; int C;
  int merge, H, W, Cin;
  const bf16* dy; long long lddy; const bf16* dres; long long lddres; bf16* dx; long long lddx;
  float* dgamma; float* dbeta;
  const float* row_scale; int rps; bf16* dx2; long long lddx2;
};
int ln_dispatch(const LnParams& p, bool bwd, cudaStream_t stream);
int colsum_dispatch(const bf16*, long long, long long, int, float*, const float*, const float*, int, cudaStream_t);
int dot_dispatch(const bf16*, long long, const bf16*, long long, long long, int, float*, cudaStream_t);
int rowwise_scale_dispatch(const bf16*, long long, bf16*, long long, long long, int, int, float, unsigned long long,
                           const float*, int, cudaStream_t);
int cast_dispatch(const float*, bf16*, long long, cudaStream_t);
int axpy_dispatch(const bf16*, long long, const bf16*, long long, const float*, bf16*, long long, long long, int,
                  cudaStream_t);
int cast_transpose_dispatch(const float*, long long, int, int, bf16*, long long, bf16*, long long, cudaStream_t);
int patch_gather_dispatch(const float*, bf16*, int, int, int, cudaStream_t);
int grid_copy_dispatch(const bf16*, bf16*, const bf16*, const float*, int, int, int, int, int, int, cudaStream_t);
int embed_dispatch(const long long*, int, int, int, int, const float*, const float*, const float*, bf16*, long long,
                   cudaStream_t);
int embed_scatter_dispatch(const long long*, int, int, int, int, const bf16*, long long, float*, float*, cudaStream_t);

static LnParams to_ln(const fiber_ln_args* a) {
  LnParams p;
  p.in1 = reinterpret_cast<const bf16*>(a->in1); p.in2 = reinterpret_cast<const bf16*>(a->in2);
  p.ld1 = a->ld1; p.ld2 = a->ld2; p.gamma = a->gamma; p.beta = a->beta; p.eps = a->eps;
  p.out = reinterpret_cast<bf16*>(a->out); p.ldo = a->ldo; p.mean = a->mean; p.rstd = a->rstd;
  p.sum_out = reinterpret_cast<bf16*>(a->sum_out); p.lds = a->lds; p.rows = a->rows; p.C = a->c;
  p.merge = a->merge; p.H = a->h; p.W = a->w; p.Cin = a->cin;
  p.dy = reinterpret_cast<const bf16*>(a->dy); p.lddy = a->lddy;
  p.dres = reinterpret_cast<const bf16*>(a->dres); p.lddres = a->lddres;
  p.dx = reinterpret_cast<bf16*>(a->dx); p.lddx = a->lddx; p.dgamma = a->dgamma; p.dbeta = a->dbeta;
  p.row_scale = a->row_scale; p.rps = a->rows_per_scale > 0 ? a->rows_per_scale : 1;
  p.dx2 = reinterpret_cast<bf16*>(a->dx_scaled); p.lddx2 = a->lddxs;
  return p;
}

}  // namespace fiber

#define FIBER_S(s) reinterpret_cast<cudaStream_t>(s)
#define FIBER_B(p) reinterpret_cast<const fiber::bf16*>(p)
#define FIBER_BM(p) reinterpret_cast<fiber::bf16*>(p)

extern "C" {

const char* fiber_last_error(void) { return fiber::g_err; }
int fiber_version(void) { return 100; }
int64_t fiber_launch_count(void) { return fiber::g_launches.load(); }

int fiber_set_option(const char* name, int32_t value) {
  if (name && strcmp(name, "winattn_tc") == 0) {
    fiber::g_winattn_tc.store(value < 0 ? -1 : (value & 15), std::memory_order_relaxed);
    return 0;
  }
  if (name && strcmp(name, "attn_small") == 0) {
    fiber::g_attn_small.store(value < 0 ? -1 : (value & 7), std::memory_order_relaxed);
    return 0;
  }
  if (name && strcmp(name, "attn_sk") == 0) {
    fiber::g_attn_sk.store(value < 0 ? -1 : (value & 31), std::memory_order_relaxed);
    return 0;
  }
  if (name && strcmp(name, "gemm_cta2") == 0) {
    fiber::g_gemm_cta2.store(value < 0 ? -1 : (value & 15), std::memory_order_relaxed);
    return 0;
  }
  if (name && strcmp(name, "pdl") == 0) {
    fiber::g_pdl.store(value < 0 ? -1 : (value != 0), std::memory_order_relaxed);
    return 0;
  }
  if (name && strcmp(name, "tq_trace") == 0) {
    fiber::g_tq_trace.store(value > 0 ? 1 : 0, std::memory_order_relaxed);
    return 0;
  }
  if (name && strcmp(name, "image_variant") == 0) {
    fiber::set_image_variant(value);
    return 0;
  }
  fiber::set_last_error("unknown option '%s'", name ? name : "(null)");
  return -1;
}
int fiber_get_option(const char* name) {
  if (name && strcmp(name, "winattn_tc") == 0) return fiber::option_winattn_tc();
  if (name && strcmp(name, "attn_small") == 0) return fiber::option_attn_small();
  if (name && strcmp(name, "winattn_tc_launches") == 0) return fiber::g_winattn_tc_launches.load();
  if (name && strcmp(name, "pdl") == 0) return fiber::option_pdl();
  if (name && strcmp(name, "gemm_cta2") == 0) return fiber::option_gemm_cta2();
  if (name && strcmp(name, "gemm_cta2_launches") == 0) return fiber::g_gemm_cta2_launches.load();
  if (name && strcmp(name, "attn_sk") == 0) return fiber::option_attn_sk();
  if (name && strcmp(name, "attn_sk_launches") == 0) return fiber::g_attn_sk_launches.load();
  if (name && strcmp(name, "image_variant") == 0) return fiber::get_image_variant();
  fiber::set_last_error("unknown option '%s'", name ? name : "(null)");
  return -1;
}

int fiber_init(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    fiber::set_last_error("cudaGetDevice failed: %s", cudaGetErrorString(e));
    return -2;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    fiber::set_last_error("cudaGetDeviceProperties failed: %s", cudaGetErrorString(e));
    return -2;
  }
  if (prop.major != 10) {
    fiber::set_last_error("fiber_b200 requires an sm_100a device (found sm_%d%d)", prop.major,
                          prop.minor);
    return -3;
  }
  return 0;
}

int fiber_gemm(const fiber_gemm_args* args, fiber_stream_t stream) {
  return fiber::gemm_dispatch(args, reinterpret_cast<cudaStream_t>(stream));
}

int fiber_mlm_ce_fwd(const fiber_ce_args* args, fiber_stream_t stream) {
  return fiber::mlm_ce_dispatch(args, 0, reinterpret_cast<cudaStream_t>(stream));
}
int fiber_mlm_ce_bwd(const fiber_ce_args* args, fiber_stream_t stream) {
  return fiber::mlm_ce_dispatch(args, 1, reinterpret_cast<cudaStream_t>(stream));
}

int fiber_attn_fwd(const fiber_attn_args* a, fiber_stream_t stream) {
  if (!a) { fiber::set_last_error("null args"); return -1; }
  return fiber::attn_fwd_dispatch(fiber::to_params(a), a->head_dim, reinterpret_cast<cudaStream_t>(stream));
}
int fiber_attn_bwd(const fiber_attn_args* a, fiber_stream_t stream) {
  if (!a) { fiber::set_last_error("null args"); return -1; }
  return fiber::attn_bwd_dispatch(fiber::to_params(a), a->head_dim, a->d_scratch, reinterpret_cast<cudaStream_t>(stream));
}

int fiber_layernorm_fwd(const fiber_ln_args* a, fiber_stream_t s) {
  if (!a || !a->in1 || !a->out || !a->gamma || !a->beta) { fiber::set_last_error("layernorm_fwd: null argument"); return -1; }
  return fiber::ln_dispatch(fiber::to_ln(a), false, FIBER_S(s));
}
int fiber_layernorm_bwd(const fiber_ln_args* a, fiber_stream_t s) {
  if (!a || !a->in1 || !a->dy || !a->dx || !a->gamma || !a->mean || !a->rstd) {
    fiber::set_last_error("layernorm_bwd: null argument"); return -1;
  }
  if ((a->dgamma == nullptr) != (a->dbeta == nullptr)) { fiber::set_last_error("layernorm_bwd: dgamma/dbeta go together"); return -1; }
  if ((a->dx_scaled != nullptr) && (a->row_scale == nullptr || a->merge)) {
    fiber::set_last_error("layernorm_bwd: dx_scaled needs row_scale and plain (non-merging) rows"); return -1;
  }
  return fiber::ln_dispatch(fiber::to_ln(a), true, FIBER_S(s));
}
int fiber_colsum(const void* x, int64_t ld, int64_t m, int32_t n, float* out, const float* scale,
                 const float* row_scale, int32_t rps, fiber_stream_t s) {
  return fiber::colsum_dispatch(FIBER_B(x), ld, m, n, out, scale, row_scale, rps, FIBER_S(s));
}
int fiber_dot(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t m, int32_t n, float* out, fiber_stream_t s) {
  return fiber::dot_dispatch(FIBER_B(a), lda, FIBER_B(b), ldb, m, n, out, FIBER_S(s));
}
int fiber_dropout(const void* x, int64_t ldx, void* y, int64_t ldy, int64_t m, int32_t n, float p, uint64_t seed,
                  fiber_stream_t s) {
  return fiber::rowwise_scale_dispatch(FIBER_B(x), ldx, FIBER_BM(y), ldy, m, n, 0, p, seed, nullptr, 1, FIBER_S(s));
}
int fiber_scale_rows(const void* x, int64_t ldx, void* y, int64_t ldy, int64_t m, int32_t n, const float* row_scale,
                     int32_t rps, fiber_stream_t s) {
  return fiber::rowwise_scale_dispatch(FIBER_B(x), ldx, FIBER_BM(y), ldy, m, n, 1, 0.f, 0, row_scale, rps, FIBER_S(s));
}
int fiber_axpy(const void* x, int64_t ldx, const void* add, int64_t ldadd, const float* alpha, void* out, int64_t ldo,
               int64_t m, int32_t n, fiber_stream_t s) {
  return fiber::axpy_dispatch(FIBER_B(x), ldx, FIBER_B(add), ldadd, alpha, FIBER_BM(out), ldo, m, n, FIBER_S(s));
}
int fiber_cast_f32_bf16(const float* x, void* y, int64_t n, fiber_stream_t s) {
  return fiber::cast_dispatch(x, FIBER_BM(y), n, FIBER_S(s));
}
int fiber_cast_transpose(const float* w, int64_t ldw, int32_t n, int32_t k, void* w_out, int64_t ld_out, void* wt_out,
                         int64_t ldt_out, fiber_stream_t s) {
  return fiber::cast_transpose_dispatch(w, ldw, n, k, FIBER_BM(w_out), ld_out, FIBER_BM(wt_out), ldt_out, FIBER_S(s));
}
int fiber_grid_copy(const void* src, void* dst, const void* add, const float* row_scale, int32_t batch, int32_t hs, int32_t ws,
                    int32_t hd, int32_t wd, int32_t c, fiber_stream_t s) {
  return fiber::grid_copy_dispatch(FIBER_B(src), FIBER_BM(dst), FIBER_B(add), row_scale, batch, hs, ws, hd, wd, c, FIBER_S(s));
}
int fiber_patch_gather(const float* img, void* out, int32_t batch, int32_t r, fiber_stream_t s) {
  return fiber::patch_gather_dispatch(img, FIBER_BM(out), batch, r, r, FIBER_S(s));
}
int fiber_patch_gather_hw(const float* img, void* out, int32_t batch, int32_t h, int32_t w, fiber_stream_t s) {
  return fiber::patch_gather_dispatch(img, FIBER_BM(out), batch, h, w, FIBER_S(s));
}
int fiber_embed_gather(const int64_t* ids, int32_t batch, int32_t len, int32_t c, int32_t pad_id, const float* word,
                       const float* pos, const float* type, void* out, int64_t ldo, fiber_stream_t s) {
  return fiber::embed_dispatch(reinterpret_cast<const long long*>(ids), batch, len, c, pad_id, word, pos, type,
                               FIBER_BM(out), ldo, FIBER_S(s));
}
int fiber_embed_scatter(const int64_t* ids, int32_t batch, int32_t len, int32_t c, int32_t pad_id, const void* dsum,
                        int64_t ldd, float* dword, float* dpos, fiber_stream_t s) {
  return fiber::embed_scatter_dispatch(reinterpret_cast<const long long*>(ids), batch, len, c, pad_id, FIBER_B(dsum),
                                       ldd, dword, dpos, FIBER_S(s));
}

}  // extern "C"
