// fiber_b200 — plain ("mode 0") attention with at most 64 keys per group on tcgen05: the image->text cross attention of
// the fused Swin blocks (swin_transformer.py:226-259: 576 / 144 / 1296 image queries per sample against <= 50 text
// tokens, head_dim 32, additive text mask) and RoBERTa self-attention (roberta.py:256-326: <= 50 x <= 50, head_dim 64,
// -10000 padding mask, probability dropout).  The mma.sync kernels of attention_{fwd,bwd}.cu remain for everything
// else (text->image with 576 / 144 keys, 324-token windows) and as the generation to compare against ("attn_sk" option).
//
// Forward.  Work item = (group, head, tile of 128 queries).  Per item: Q [128][hd], K [64][hd], V [64][hd] arrive as
// three TMA boxes (2-D maps over the row-major activations; rows past the tensor end are zero-filled, rows of the next
// group are finite garbage that the key mask / the row predicate neutralise); S = Q K^T is ONE tcgen05.mma chain
// (M = 128, N = 64) into TMEM; thread = query row turns its 64 scores into probabilities (scale + mask in the log2 domain,
// row maximum and sum without shuffles, optional dropout) and writes them as bf16 into one SWIZZLE_128B chunk;
// O = P V is a second chain (M = 128, N = hd, K = 64) whose accumulator is drained one item later (O double-buffered).
// Six warps (4 element-wise, MMA issuer, TMA producer), 256 TMEM columns, ~80 KB of shared memory: two CTAs per SM
// overlap each other's bubbles.
#include "attention.cuh"
#include "window_tc_layout.cuh"
#include "../../include/fiber_b200.h"

#include <mutex>

namespace fiber {

void count_launch(int n = 1);

namespace {

using namespace tcl;

constexpr float SK_LOG2E = 1.4426950408889634f;
constexpr float SK_LN2 = 0.6931471805599453f;
constexpr int SK_THREADS = 192;
constexpr int SK_WARP_MMA = 4, SK_WARP_LD = 5;
constexpr int SK_KEYS = 64;

__device__ __forceinline__ void sk_wait_timeout(int tag, uint32_t parity, int it) {
  printf("fiber_b200 attention_sk: mbarrier timeout tag %d parity %u item %d block %d warp %d lane %d\n", tag, parity, it,
         blockIdx.x, threadIdx.x >> 5, threadIdx.x & 31);
  __trap();
}
__device__ __forceinline__ void sk_wait(uint64_t* bar, uint32_t parity, int tag, int it) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > 4) __nanosleep(spins > 64 ? 256 : 32);
    if (spins > (1u << 22)) sk_wait_timeout(tag, parity, it);
  }
}
__device__ __forceinline__ float sk_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 8 columns (the packed kernels read their group's columns from an 8-aligned, not 32-aligned, base)
__device__ __forceinline__ void sk_tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void sk_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// operand descriptors of a [rows][HD] bf16 tile with HD * 2-byte rows: SWIZZLE_64B for HD = 32, SWIZZLE_128B for HD = 64
template <int HD>
__device__ __forceinline__ uint64_t sk_desc_kmajor(uint32_t tile, int ks) {  // K = head dim, step = 16 elements = 32 B
  if constexpr (HD == 32) return desc_tile_kmajor(tile, ks);
  else return umma_desc_sw128(tile + ks * 32, 16, 1024);
}
template <int HD>
__device__ __forceinline__ uint64_t sk_desc_mnmajor(uint32_t tile, int kk) {  // K = rows (tokens), step = 16 rows
  if constexpr (HD == 32) return desc_tile_mnmajor(tile, kk);
  else return umma_desc_sw128(tile + kk * 2048, 8192, 1024);  // the wgrad GEMM's MN-major form (one 128-byte MN chunk)
}

template <int HD>
struct SkFwdCfg {
  static constexpr int Q_BYTES = 128 * HD * 2, KV_BYTES = SK_KEYS * HD * 2;
  static constexpr int STAGE_BYTES = Q_BYTES + 2 * KV_BYTES;
  static constexpr int STAGES = 2;
  static constexpr int OFF_P = STAGES * STAGE_BYTES;       // [128][64] bf16, 128-byte rows, SWIZZLE_128B
  static constexpr int OFF_MSK = OFF_P + 128 * 128;        // [2 stages][64] fp32: additive key mask, log2 domain
  static constexpr int OFF_BARS = OFF_MSK + 2 * SK_KEYS * 4;
  static constexpr int SMEM = 1024 + OFF_BARS + 16 * 8;
  static constexpr uint32_t S_COL0 = 0, S_COL1 = 64, O_COL0 = 128, O_COL1 = 192;
};

template <int HD, bool DROPOUT>
__global__ void __launch_bounds__(SK_THREADS, 2) attn_sk_fwd_kernel(const AttnParams p, const __grid_constant__ CUtensorMap tmQ,
                                                                   const __grid_constant__ CUtensorMap tmK,
                                                                   const __grid_constant__ CUtensorMap tmV, int ntiles,
                                                                   int n_items) {
  using Cfg = SkFwdCfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sP = smem + Cfg::OFF_P;
  float* sMsk = reinterpret_cast<float*>(smem + Cfg::OFF_MSK);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BARS);
  // timeout tags: 301 s_full, 302 o_full, 303 full, 304 s_empty, 305 p_ready, 306 stage_free, 307 o_empty
  uint64_t* full = bars;            // [2] TMA boxes of a stage have landed     (expect_tx)
  uint64_t* stage_free = bars + 2;  // [2] stage may be overwritten             (tcgen05.commit)
  uint64_t* s_full = bars + 4;      // [2] S written                            (tcgen05.commit)
  uint64_t* s_empty = bars + 6;     // [2] S read                               (4 element-wise warps)
  uint64_t* p_ready = bars + 8;     //     P in smem                            (4 element-wise warps)
  uint64_t* o_full = bars + 9;      // [2] O written, P consumed                (tcgen05.commit)
  uint64_t* o_empty = bars + 11;    // [2] O drained                            (4 element-wise warps)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 13);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_my = (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const float scale2 = p.scale * SK_LOG2E;

  pdl_trigger();
  if (warp == SK_WARP_MMA) {
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&stage_free[s], 1);
        mbar_init(&s_full[s], 1);
        mbar_init(&s_empty[s], 4);
        mbar_init(&o_full[s], 1);
        mbar_init(&o_empty[s], 4);
      }
      mbar_init(p_ready, 4);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, 256);
    tmem_relinquish();
  }
  if (warp == SK_WARP_LD && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();  // nothing above touched global memory

  if (warp < 4) {
    // ================= element-wise warps: thread = query row of the tile =================
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    uint8_t* p_row = sP + row * 128;
    const int xr = row & 7;
    const float keep_inv = DROPOUT ? 1.0f / (1.0f - p.drop_p) : 1.0f;
    float m_prev = 0.f, l_prev = 1.f;
    long long orow_prev = -1;   // global output row of the previous item (-1: row not valid)
    long long lse_prev = 0;
    int h_prev = 0;

    auto epilogue = [&](int itp) {
      const int bp = itp & 1;
      sk_wait(&o_full[bp], (itp >> 1) & 1, 302, itp);
      tc_fence_after();
      uint32_t o[HD];
      if constexpr (HD == 64) {
        tmem_ld32(lane_addr + (bp ? Cfg::O_COL1 : Cfg::O_COL0), reinterpret_cast<uint32_t(&)[32]>(o[0]));
        tmem_ld32(lane_addr + (bp ? Cfg::O_COL1 : Cfg::O_COL0) + 32, reinterpret_cast<uint32_t(&)[32]>(o[32]));
      } else {
        tmem_ld32(lane_addr + (bp ? Cfg::O_COL1 : Cfg::O_COL0), reinterpret_cast<uint32_t(&)[32]>(o[0]));
      }
      tmem_ld_wait();
      tc_fence_before();
      if (lane == 0) mbar_arrive(&o_empty[bp]);
      if (orow_prev >= 0) {
        const float inv = 1.0f / l_prev;
        bf16* dst = p.o + orow_prev * p.ldo + h_prev * HD;
#pragma unroll
        for (int c = 0; c < HD; c += 8) {
          uint4 v;
          v.x = pack_bf16(__uint_as_float(o[c]) * inv, __uint_as_float(o[c + 1]) * inv);
          v.y = pack_bf16(__uint_as_float(o[c + 2]) * inv, __uint_as_float(o[c + 3]) * inv);
          v.z = pack_bf16(__uint_as_float(o[c + 4]) * inv, __uint_as_float(o[c + 5]) * inv);
          v.w = pack_bf16(__uint_as_float(o[c + 6]) * inv, __uint_as_float(o[c + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + c) = v;
        }
        if (p.lse) p.lse[lse_prev] = (m_prev + sk_lg2(l_prev)) * SK_LN2;
      }
    };

#pragma unroll 1
    for (int it = 0; it < n_my; ++it) {
      const int item = blockIdx.x + it * gridDim.x;
      const int t = item % ntiles, gh = item / ntiles;
      const int h = gh % p.nH, g = gh / p.nH;
      const int b = it & 1;
      const int qi = t * 128 + row;
      // additive key mask of this group in the log2 domain; keys past Lk (rows of the next group) are switched off
      if (row < SK_KEYS) {
        float mk = -1e30f;
        if (row < p.Lk) mk = p.key_mask ? p.key_mask[static_cast<long long>(g) * p.Lk + row] * SK_LOG2E : 0.f;
        sMsk[b * SK_KEYS + row] = mk;
      }
      sk_bar_sync(1, 128);

      sk_wait(&s_full[b], (it >> 1) & 1, 301, it);
      tc_fence_after();
      uint32_t v[64];
      tmem_ld32(lane_addr + (b ? Cfg::S_COL1 : Cfg::S_COL0), reinterpret_cast<uint32_t(&)[32]>(v[0]));
      tmem_ld32(lane_addr + (b ? Cfg::S_COL1 : Cfg::S_COL0) + 32, reinterpret_cast<uint32_t(&)[32]>(v[32]));
      tmem_ld_wait();
      tc_fence_before();
      if (lane == 0) mbar_arrive(&s_empty[b]);
      float mx = -1e30f;
#pragma unroll
      for (int j = 0; j < 64; j += 4) {
        const float4 mk = *reinterpret_cast<const float4*>(sMsk + b * SK_KEYS + j);
        const float x0 = fmaf(__uint_as_float(v[j]), scale2, mk.x), x1 = fmaf(__uint_as_float(v[j + 1]), scale2, mk.y);
        const float x2 = fmaf(__uint_as_float(v[j + 2]), scale2, mk.z), x3 = fmaf(__uint_as_float(v[j + 3]), scale2, mk.w);
        v[j] = __float_as_uint(x0); v[j + 1] = __float_as_uint(x1);
        v[j + 2] = __float_as_uint(x2); v[j + 3] = __float_as_uint(x3);
        mx = fmaxf(mx, fmaxf(fmaxf(x0, x1), fmaxf(x2, x3)));
      }
      if (it > 0) sk_wait(&o_full[b ^ 1], ((it - 1) >> 1) & 1, 302, it);  // P V of the previous item has consumed P
      float sum = 0.f;
      const unsigned long long didx0 = DROPOUT ? ((static_cast<unsigned long long>(g) * p.nH + h) * p.Lq + qi) * p.Lk : 0ull;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        float pr[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          pr[e] = ex2_approx(__uint_as_float(v[c8 * 8 + e]) - mx);
          sum += pr[e];
          if (DROPOUT) pr[e] = dropout_keep(p.seed, didx0 + c8 * 8 + e, p.drop_p) ? pr[e] * keep_inv : 0.f;
        }
        uint4 o;
        o.x = pack_bf16(pr[0], pr[1]); o.y = pack_bf16(pr[2], pr[3]);
        o.z = pack_bf16(pr[4], pr[5]); o.w = pack_bf16(pr[6], pr[7]);
        *reinterpret_cast<uint4*>(p_row + ((c8 ^ xr) << 4)) = o;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);

      if (it > 0) epilogue(it - 1);
      m_prev = mx;
      l_prev = sum;
      h_prev = h;
      orow_prev = qi < p.Lq ? static_cast<long long>(g) * p.Lq + qi : -1;
      lse_prev = (static_cast<long long>(g) * p.nH + h) * p.Lq + qi;
      // (the mask row of stage b is rewritten two items later, behind the barrier of the item in between)
    }
    if (n_my > 0) epilogue(n_my - 1);
  } else if (warp == SK_WARP_MMA) {
    // ================= tcgen05.mma issuer =================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, SK_KEYS, 0, 0);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, HD, 0, 1);
    auto issue_s = [&](int it) {
      const int s = it & 1;
      sk_wait(&full[s], (it >> 1) & 1, 303, it);
      sk_wait(&s_empty[s], ((it >> 1) & 1) ^ 1, 304, it);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t q_addr = smem_u32(smem + s * Cfg::STAGE_BYTES), k_addr = q_addr + Cfg::Q_BYTES;
        const uint32_t d = tmem_base + (s ? Cfg::S_COL1 : Cfg::S_COL0);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          umma_f16_ss(d, sk_desc_kmajor<HD>(q_addr, ks), sk_desc_kmajor<HD>(k_addr, ks), idesc_s, ks);
        umma_commit(&s_full[s]);
      }
      __syncwarp();
    };
    if (n_my > 0) issue_s(0);
#pragma unroll 1
    for (int it = 0; it < n_my; ++it) {
      if (it + 1 < n_my) issue_s(it + 1);
      const int b = it & 1;
      sk_wait(p_ready, it & 1, 305, it);
      sk_wait(&o_empty[b], ((it >> 1) & 1) ^ 1, 307, it);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t v_addr = smem_u32(smem + b * Cfg::STAGE_BYTES + Cfg::Q_BYTES + Cfg::KV_BYTES), p_addr = smem_u32(sP);
        const uint32_t d = tmem_base + (b ? Cfg::O_COL1 : Cfg::O_COL0);
#pragma unroll
        for (int kk = 0; kk < SK_KEYS / 16; ++kk)
          umma_f16_ss(d, umma_desc_sw128(p_addr + kk * 32, 16, 1024), sk_desc_mnmajor<HD>(v_addr, kk), idesc_o, kk);
        umma_commit(&o_full[b]);
        umma_commit(&stage_free[b]);
      }
      __syncwarp();
    }
  } else {
    // ================= TMA producer =================
    if (lane == 0) {
#pragma unroll 1
      for (int it = 0; it < n_my; ++it) {
        const int item = blockIdx.x + it * gridDim.x;
        const int t = item % ntiles, gh = item / ntiles;
        const int h = gh % p.nH, g = gh / p.nH;
        const int s = it & 1;
        if (it >= 2) sk_wait(&stage_free[s], ((it >> 1) - 1) & 1, 306, it);
        uint8_t* st = smem + s * Cfg::STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], Cfg::STAGE_BYTES);
        tma_load_2d(st, &tmQ, &full[s], h * HD, g * p.Lq + t * 128);
        tma_load_2d(st + Cfg::Q_BYTES, &tmK, &full[s], h * HD, g * p.Lk);
        tma_load_2d(st + Cfg::Q_BYTES + Cfg::KV_BYTES, &tmV, &full[s], h * HD, g * p.Lk);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == SK_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// =====================================================================================================================
// Backward.  Work item = (group, head); its query tiles of 128 rows are walked in order by ONE CTA, so dK / dV accumulate
// in TMEM across the tiles and leave once per item (no atomics).  Per tile: Q, dO, O [128][hd] and K, V [64][hd] arrive as
// five TMA boxes; S = Q K^T and dP = dO V^T are two tcgen05.mma chains (M = 128, N = 64) into TMEM; thread = query row
// computes D = rowsum(dO * O) from the staged tiles, P = exp(S * scale + mask - lse) (dropout regenerated from the
// forward's counter hash), dS = P (dP - D), and writes bf16 P (after dropout) and dS into SWIZZLE_128B chunks; three more
// chains follow: dV += P^T dO, dK += dS^T Q (M = 128 key rows of which <= 64 are real: accumulator rows are independent,
// so the operand's upper 64 "keys" simply read whatever follows the chunk in shared memory and their rows are never
// drained; N = hd, K = 128 queries, MN-major operands straight from the row-major tiles) and dQ = dS K (M = 128, N = hd,
// K = 64).  Rows of the tile past the group's last query (next group's rows, finite) get P = dS = 0.
// hd = 32: 96 KB of shared memory and 256 TMEM columns, two CTAs per SM overlap each other's bubbles; hd = 64: one CTA.
// =====================================================================================================================
template <int HD>
struct SkBwdCfg {
  static constexpr int T_BYTES = 128 * HD * 2, KV_BYTES = SK_KEYS * HD * 2;
  static constexpr int STAGE_BYTES = 3 * T_BYTES + 2 * KV_BYTES;  // Q, dO, O, K, V
  static constexpr int STAGES = 2;
  static constexpr int CHUNK = 128 * 128;                          // [128 query rows][64 keys] bf16, SWIZZLE_128B
  // dS, P, then the stages: the MN-major P^T / dS^T operands span two chunks (M = 128), the second being the next
  // 16 KB of shared memory (P for dS, the first stage for P) — finite or not, those accumulator rows are discarded
  static constexpr int OFF_DS = 0, OFF_P = CHUNK, OFF_STAGES = 2 * CHUNK;
  static constexpr int OFF_MSK = OFF_STAGES + STAGES * STAGE_BYTES;  // [2][64] fp32 key mask, log2 domain
  static constexpr int OFF_BARS = OFF_MSK + 2 * SK_KEYS * 4;
  static constexpr int SMEM = 1024 + OFF_BARS + 16 * 8;
  static constexpr uint32_t S_COL = 0, DP_COL = 64, DQ_COL = 128, DV_COL = 128 + HD, DK_COL = 128 + 2 * HD;
  static constexpr uint32_t TMEM_COLS = HD == 32 ? 256 : 512;
  static constexpr int CTAS_PER_SM = HD == 32 ? 2 : 1;
};
static_assert(2 * (SkBwdCfg<32>::SMEM + 1024) <= 227 * 1024, "two hd = 32 CTAs per SM");
static_assert(SkBwdCfg<64>::SMEM <= 227 * 1024, "backward shared memory");

// D of one tile row from the staged dO / O tiles (HD * 2-byte swizzled rows)
template <int HD>
__device__ __forceinline__ float sk_row_dot(const uint8_t* sdO, const uint8_t* sO, int row) {
  float acc = 0.f;
#pragma unroll
  for (int pc = 0; pc < HD / 8; ++pc) {
    const uint32_t off = HD == 32 ? sw64_off(row, pc) : sw128_off(row, pc);
    const uint4 a = *reinterpret_cast<const uint4*>(sdO + off);
    const uint4 b = *reinterpret_cast<const uint4*>(sO + off);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fa = unpack_bf16(aw[e]), fb = unpack_bf16(bw[e]);
      acc = fmaf(fa.x, fb.x, acc);
      acc = fmaf(fa.y, fb.y, acc);
    }
  }
  return acc;
}

template <int HD, bool DROPOUT>
__global__ void __launch_bounds__(SK_THREADS, SkBwdCfg<HD>::CTAS_PER_SM) attn_sk_bwd_kernel(const AttnParams p, const __grid_constant__ CUtensorMap tmQ,
                                                                   const __grid_constant__ CUtensorMap tmdO,
                                                                   const __grid_constant__ CUtensorMap tmO,
                                                                   const __grid_constant__ CUtensorMap tmK,
                                                                   const __grid_constant__ CUtensorMap tmV, int ntiles,
                                                                   int n_items) {
  using Cfg = SkBwdCfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sP = smem + Cfg::OFF_P;
  uint8_t* sdS = smem + Cfg::OFF_DS;
  float* sMsk = reinterpret_cast<float*>(smem + Cfg::OFF_MSK);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BARS);
  // timeout tags: 401 full, 402 stage_free, 403 s_full, 404 sdp_empty, 405 pds_ready, 406 out_full, 407 acc_empty
  uint64_t* full = bars;            // [2] TMA boxes of a stage have landed                (expect_tx)
  uint64_t* stage_free = bars + 2;  // [2] stage may be overwritten                        (tcgen05.commit)
  uint64_t* s_full = bars + 4;      //     S and dP written                                (tcgen05.commit)
  uint64_t* sdp_empty = bars + 5;   //     S and dP read                                   (4 element-wise warps)
  uint64_t* pds_ready = bars + 6;   //     P and dS in smem                                (4 element-wise warps)
  uint64_t* out_full = bars + 7;    //     dQ (and dV / dK so far) written, P / dS consumed (tcgen05.commit)
  uint64_t* acc_empty = bars + 8;   //     dQ (and, after an item's last tile, dV / dK) drained (4 element-wise warps)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 9);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_my_items = (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int n_my = n_my_items * ntiles;  // units = (item, tile), tiles of an item consecutive
  const float scale2 = p.scale * SK_LOG2E;

  pdl_trigger();
  if (warp == SK_WARP_MMA) {
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&stage_free[s], 1);
      }
      mbar_init(s_full, 1);
      mbar_init(sdp_empty, 4);
      mbar_init(pds_ready, 4);
      mbar_init(out_full, 1);
      mbar_init(acc_empty, 4);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  if (warp == SK_WARP_LD && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();  // nothing above touched global memory

  if (warp < 4) {
    // ================= element-wise warps: thread = query row of the tile (and key row of the dV / dK drain) =================
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    uint8_t* p_row = sP + row * 128;
    uint8_t* ds_row = sdS + row * 128;
    const int xr = row & 7;
    const float keep_inv = DROPOUT ? 1.0f / (1.0f - p.drop_p) : 1.0f;
#pragma unroll 1
    for (int u = 0; u < n_my; ++u) {
      const int item = blockIdx.x + (u / ntiles) * gridDim.x, t = u % ntiles;
      const int h = item % p.nH, g = item / p.nH;
      const int s = u & 1;
      const int qi = t * 128 + row;
      const bool valid = qi < p.Lq;
      const uint8_t* stage = smem + Cfg::OFF_STAGES + s * Cfg::STAGE_BYTES;
      if (row < SK_KEYS) {
        float mk = -1e30f;
        if (row < p.Lk) mk = p.key_mask ? p.key_mask[static_cast<long long>(g) * p.Lk + row] * SK_LOG2E : 0.f;
        sMsk[s * SK_KEYS + row] = mk;
      }
      const float nlse2 = valid ? -p.lse[(static_cast<long long>(g) * p.nH + h) * p.Lq + qi] * SK_LOG2E : 0.f;
      sk_wait(&full[s], (u >> 1) & 1, 401, u);
      const float D = sk_row_dot<HD>(stage + Cfg::T_BYTES, stage + 2 * Cfg::T_BYTES, row);
      sk_bar_sync(1, 128);  // the mask row of this stage is complete

      sk_wait(s_full, u & 1, 403, u);
      tc_fence_after();
      const unsigned long long didx0 = DROPOUT ? ((static_cast<unsigned long long>(g) * p.nH + h) * p.Lq + qi) * p.Lk : 0ull;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {  // 32 key columns at a time
        uint32_t sv[32], dv[32];
        tmem_ld32(lane_addr + Cfg::S_COL + hf * 32, sv);
        tmem_ld32(lane_addr + Cfg::DP_COL + hf * 32, dv);
        tmem_ld_wait();
        if (hf == 1) {
          tc_fence_before();
          if (lane == 0) mbar_arrive(sdp_empty);  // S / dP may be overwritten by the next tile's MMAs
        }
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          float pd[8], ds[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int j = hf * 32 + c8 * 8 + e;
            const float x = fmaf(__uint_as_float(sv[c8 * 8 + e]), scale2, sMsk[s * SK_KEYS + j]) + nlse2;
            const float pr = valid ? ex2_approx(x) : 0.f;
            float dpv = __uint_as_float(dv[c8 * 8 + e]);
            pd[e] = pr;
            if (DROPOUT) {
              const bool keep = dropout_keep(p.seed, didx0 + j, p.drop_p);
              pd[e] = keep ? pr * keep_inv : 0.f;
              dpv = keep ? dpv * keep_inv : 0.f;
            }
            ds[e] = pr * (dpv - D);
          }
          uint4 o, d;
          o.x = pack_bf16(pd[0], pd[1]); o.y = pack_bf16(pd[2], pd[3]);
          o.z = pack_bf16(pd[4], pd[5]); o.w = pack_bf16(pd[6], pd[7]);
          d.x = pack_bf16(ds[0], ds[1]); d.y = pack_bf16(ds[2], ds[3]);
          d.z = pack_bf16(ds[4], ds[5]); d.w = pack_bf16(ds[6], ds[7]);
          const int piece = hf * 4 + c8;
          *reinterpret_cast<uint4*>(p_row + ((piece ^ xr) << 4)) = o;
          *reinterpret_cast<uint4*>(ds_row + ((piece ^ xr) << 4)) = d;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_ready);

      // drain dQ of this tile (and dV / dK of the item after its last tile)
      sk_wait(out_full, u & 1, 406, u);
      tc_fence_after();
      {
        uint32_t a[HD];
        tmem_ld32(lane_addr + Cfg::DQ_COL, reinterpret_cast<uint32_t(&)[32]>(a[0]));
        if constexpr (HD == 64) tmem_ld32(lane_addr + Cfg::DQ_COL + 32, reinterpret_cast<uint32_t(&)[32]>(a[32]));
        tmem_ld_wait();
        if (valid) {
          bf16* dst = p.dq + (static_cast<long long>(g) * p.Lq + qi) * p.lddq + h * HD;
#pragma unroll
          for (int c = 0; c < HD; c += 8) {
            uint4 v;
            v.x = pack_bf16(__uint_as_float(a[c]) * p.scale, __uint_as_float(a[c + 1]) * p.scale);
            v.y = pack_bf16(__uint_as_float(a[c + 2]) * p.scale, __uint_as_float(a[c + 3]) * p.scale);
            v.z = pack_bf16(__uint_as_float(a[c + 4]) * p.scale, __uint_as_float(a[c + 5]) * p.scale);
            v.w = pack_bf16(__uint_as_float(a[c + 6]) * p.scale, __uint_as_float(a[c + 7]) * p.scale);
            *reinterpret_cast<uint4*>(dst + c) = v;
          }
        }
      }
      if (t == ntiles - 1 && warp < 2) {  // accumulator lanes = key rows 0..63
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          uint32_t a[HD];
          tmem_ld32(lane_addr + (w ? Cfg::DK_COL : Cfg::DV_COL), reinterpret_cast<uint32_t(&)[32]>(a[0]));
          if constexpr (HD == 64) tmem_ld32(lane_addr + (w ? Cfg::DK_COL : Cfg::DV_COL) + 32, reinterpret_cast<uint32_t(&)[32]>(a[32]));
          tmem_ld_wait();
          if (row < p.Lk) {
            const float sc = w ? p.scale : 1.0f;
            bf16* dst = (w ? p.dk + (static_cast<long long>(g) * p.Lk + row) * p.lddk
                           : p.dv + (static_cast<long long>(g) * p.Lk + row) * p.lddv) + h * HD;
#pragma unroll
            for (int c = 0; c < HD; c += 8) {
              uint4 v;
              v.x = pack_bf16(__uint_as_float(a[c]) * sc, __uint_as_float(a[c + 1]) * sc);
              v.y = pack_bf16(__uint_as_float(a[c + 2]) * sc, __uint_as_float(a[c + 3]) * sc);
              v.z = pack_bf16(__uint_as_float(a[c + 4]) * sc, __uint_as_float(a[c + 5]) * sc);
              v.w = pack_bf16(__uint_as_float(a[c + 6]) * sc, __uint_as_float(a[c + 7]) * sc);
              *reinterpret_cast<uint4*>(dst + c) = v;
            }
          }
        }
      }
      tc_fence_before();
      if (lane == 0) mbar_arrive(acc_empty);
    }
  } else if (warp == SK_WARP_MMA) {
    // ================= tcgen05.mma issuer =================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, SK_KEYS, 0, 0);  // S = Q K^T, dP = dO V^T
    constexpr uint32_t idesc_kv = umma_idesc_bf16(128, HD, 1, 1);      // dV = P^T dO, dK = dS^T Q (MN x MN)
    constexpr uint32_t idesc_q = umma_idesc_bf16(128, HD, 0, 1);       // dQ = dS K (K-major x MN-major)
    const uint32_t p_addr = smem_u32(sP), ds_addr = smem_u32(sdS);
    auto issue_scores = [&](int u) {
      const int s = u & 1;
      sk_wait(&full[s], (u >> 1) & 1, 401, u);
      sk_wait(sdp_empty, (u & 1) ^ 1, 404, u);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t q_addr = smem_u32(smem + Cfg::OFF_STAGES + s * Cfg::STAGE_BYTES), do_addr = q_addr + Cfg::T_BYTES;
        const uint32_t k_addr = q_addr + 3 * Cfg::T_BYTES, v_addr = k_addr + Cfg::KV_BYTES;
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          umma_f16_ss(tmem_base + Cfg::S_COL, sk_desc_kmajor<HD>(q_addr, ks), sk_desc_kmajor<HD>(k_addr, ks), idesc_s, ks);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          umma_f16_ss(tmem_base + Cfg::DP_COL, sk_desc_kmajor<HD>(do_addr, ks), sk_desc_kmajor<HD>(v_addr, ks), idesc_s, ks);
        umma_commit(s_full);
      }
      __syncwarp();
    };
    if (n_my > 0) issue_scores(0);
#pragma unroll 1
    for (int u = 0; u < n_my; ++u) {
      const int s = u & 1, t = u % ntiles;
      sk_wait(pds_ready, u & 1, 405, u);
      sk_wait(acc_empty, (u & 1) ^ 1, 407, u);  // the previous tile's dQ (and a finished item's dV / dK) are drained
      tc_fence_after();
      if (lane == 0) {
        const uint32_t q_addr = smem_u32(smem + Cfg::OFF_STAGES + s * Cfg::STAGE_BYTES), do_addr = q_addr + Cfg::T_BYTES;
        const uint32_t k_addr = q_addr + 3 * Cfg::T_BYTES;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // 16 query rows per step
          const uint32_t acc = (t > 0 || kk > 0) ? 1u : 0u;
          umma_f16_ss(tmem_base + Cfg::DV_COL, umma_desc(p_addr + kk * 2048, Cfg::CHUNK, 1024, LAYOUT_SW128),
                      sk_desc_mnmajor<HD>(do_addr, kk), idesc_kv, acc);
          umma_f16_ss(tmem_base + Cfg::DK_COL, umma_desc(ds_addr + kk * 2048, Cfg::CHUNK, 1024, LAYOUT_SW128),
                      sk_desc_mnmajor<HD>(q_addr, kk), idesc_kv, acc);
        }
#pragma unroll
        for (int kk = 0; kk < SK_KEYS / 16; ++kk)  // 16 keys per step
          umma_f16_ss(tmem_base + Cfg::DQ_COL, umma_desc_sw128(ds_addr + kk * 32, 16, 1024), sk_desc_mnmajor<HD>(k_addr, kk),
                      idesc_q, kk);
        umma_commit(out_full);
        umma_commit(&stage_free[s]);
      }
      __syncwarp();
      if (u + 1 < n_my) issue_scores(u + 1);
    }
  } else {
    // ================= TMA producer =================
    if (lane == 0) {
#pragma unroll 1
      for (int u = 0; u < n_my; ++u) {
        const int item = blockIdx.x + (u / ntiles) * gridDim.x, t = u % ntiles;
        const int h = item % p.nH, g = item / p.nH;
        const int s = u & 1;
        if (u >= 2) sk_wait(&stage_free[s], ((u >> 1) - 1) & 1, 402, u);
        uint8_t* st = smem + Cfg::OFF_STAGES + s * Cfg::STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], Cfg::STAGE_BYTES);
        tma_load_2d(st, &tmQ, &full[s], h * HD, g * p.Lq + t * 128);
        tma_load_2d(st + Cfg::T_BYTES, &tmdO, &full[s], h * HD, g * p.Lq + t * 128);
        tma_load_2d(st + 2 * Cfg::T_BYTES, &tmO, &full[s], h * HD, g * p.Lq + t * 128);
        tma_load_2d(st + 3 * Cfg::T_BYTES, &tmK, &full[s], h * HD, g * p.Lk);
        tma_load_2d(st + 3 * Cfg::T_BYTES + Cfg::KV_BYTES, &tmV, &full[s], h * HD, g * p.Lk);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == SK_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// =====================================================================================================================
// Packed self-attention (RoBERTa self-attention, roberta.py:256-326: Lq = Lk = L, 32 <= L <= 64 tokens, L % 8 == 0).  A 128-row
// tile holds gpt = 128 / L whole groups (samples): rows [g0 L, g0 L + 128) of the row-major activations are ONE TMA box
// for Q and one each for K and V, S = Q K^T is a 128 x 128 chain of which row r only reads the L columns of its own
// group (block diagonal: gi = r / L, columns [gi L, gi L + L)), P is written block-diagonally into two SWIZZLE_128B
// chunks that were zeroed once (a row always writes the same column range), and O = P V (K = 128 keys) needs no
// masking: the off-diagonal blocks are zero.  120 of 128 tile rows carry work at L = 40 (40 of 128 unpacked).
// tcgen05.ld is warp-collective (one column address per warp) and 32 <= L means a warp's 32 rows touch at most two
// groups: the warp loads the column ranges of both and every thread selects its own.
// The backward is the same construction on the sk backward: all 128 key rows of dV / dK are real, every tile is its
// own item.
// =====================================================================================================================
template <int HD>
struct PkCfg {
  static constexpr int T_BYTES = 128 * HD * 2;
  static constexpr int CHUNK = 128 * 128;  // [128 rows][64 keys] bf16, SWIZZLE_128B
  // forward: Q, K, V per stage; backward: Q, dO, O, K, V
  static constexpr int F_STAGE = 3 * T_BYTES, B_STAGE = 5 * T_BYTES;
  static constexpr int F_OFF_P = 2 * F_STAGE, F_OFF_MSK = F_OFF_P + 2 * CHUNK, F_OFF_BARS = F_OFF_MSK + 2 * 128 * 4;
  static constexpr int F_SMEM = 1024 + F_OFF_BARS + 16 * 8;
  static constexpr int B_OFF_DS = 0, B_OFF_P = 2 * CHUNK, B_OFF_STAGES = 4 * CHUNK;
  static constexpr int B_OFF_MSK = B_OFF_STAGES + 2 * B_STAGE, B_OFF_BARS = B_OFF_MSK + 2 * 128 * 4;
  static constexpr int B_SMEM = 1024 + B_OFF_BARS + 16 * 8;
  static constexpr uint32_t FS_COL0 = 0, FS_COL1 = 128, FO_COL0 = 256, FO_COL1 = 256 + HD;
  static constexpr uint32_t BS_COL = 0, BDP_COL = 128, BDQ_COL = 256, BDV_COL = 256 + HD, BDK_COL = 256 + 2 * HD;
};
static_assert(PkCfg<64>::B_SMEM <= 227 * 1024 && PkCfg<64>::F_SMEM <= 227 * 1024, "packed attention shared memory");

struct PkGeo {
  int L, gpt, n_tiles;  // tokens per group, groups per 128-row tile, tiles per head
};

// mask of tile key column c (group g0 + c / L, key c % L) in the log2 domain; columns of no group are switched off
__device__ __forceinline__ float pk_mask(const AttnParams& p, const PkGeo& geo, int g0, int c) {
  const int gi = c / geo.L, j = c - gi * geo.L, g = g0 + gi;
  if (gi >= geo.gpt || g >= p.G) return -1e30f;
  return p.key_mask ? p.key_mask[static_cast<long long>(g) * geo.L + j] * SK_LOG2E : 0.f;
}

template <int HD, bool DROPOUT>
__global__ void __launch_bounds__(SK_THREADS, 1) attn_pk_fwd_kernel(const AttnParams p, const __grid_constant__ CUtensorMap tmQ,
                                                                   const __grid_constant__ CUtensorMap tmK,
                                                                   const __grid_constant__ CUtensorMap tmV, const PkGeo geo,
                                                                   int n_items) {
  using Cfg = PkCfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sP = smem + Cfg::F_OFF_P;
  float* sMsk = reinterpret_cast<float*>(smem + Cfg::F_OFF_MSK);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::F_OFF_BARS);
  // timeout tags: 501 s_full, 502 o_full, 503 full, 504 s_empty, 505 p_ready, 506 stage_free, 507 o_empty
  uint64_t* full = bars;
  uint64_t* stage_free = bars + 2;
  uint64_t* s_full = bars + 4;
  uint64_t* s_empty = bars + 6;
  uint64_t* p_ready = bars + 8;
  uint64_t* o_full = bars + 9;
  uint64_t* o_empty = bars + 11;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 13);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_my = (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const float scale2 = p.scale * SK_LOG2E;
  const int L = geo.L;

  pdl_trigger();
  if (warp == SK_WARP_MMA) {
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&stage_free[s], 1);
        mbar_init(&s_full[s], 1);
        mbar_init(&s_empty[s], 4);
        mbar_init(&o_full[s], 1);
        mbar_init(&o_empty[s], 4);
      }
      mbar_init(p_ready, 4);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  if (warp == SK_WARP_LD && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  for (int i = tid; i < 2 * Cfg::CHUNK / 16; i += SK_THREADS) reinterpret_cast<uint4*>(sP)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();

  if (warp < 4) {
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const int gi = row / L, qi = row - gi * L;
    const int kc0 = gi < geo.gpt ? gi * L : 0;  // first key column of this row's group
    const int gi_lo = (warp * 32) / L;          // first group of this warp's rows (warp-uniform)
    const int kcw = gi_lo < geo.gpt ? gi_lo * L : 0;
    const bool two = (warp * 32 + 31) / L != gi_lo;  // the warp's rows span two groups
    const bool second = gi != gi_lo;
    uint8_t* p_row = sP + row * 128;
    const int xr = row & 7;
    const float keep_inv = DROPOUT ? 1.0f / (1.0f - p.drop_p) : 1.0f;
    float m_prev = 0.f, l_prev = 1.f;
    long long orow_prev = -1, lse_prev = 0;
    int h_prev = 0;

    auto epilogue = [&](int itp) {
      const int bp = itp & 1;
      sk_wait(&o_full[bp], (itp >> 1) & 1, 502, itp);
      tc_fence_after();
      uint32_t o[HD];
      tmem_ld32(lane_addr + (bp ? Cfg::FO_COL1 : Cfg::FO_COL0), reinterpret_cast<uint32_t(&)[32]>(o[0]));
      if constexpr (HD == 64) tmem_ld32(lane_addr + (bp ? Cfg::FO_COL1 : Cfg::FO_COL0) + 32, reinterpret_cast<uint32_t(&)[32]>(o[32]));
      tmem_ld_wait();
      tc_fence_before();
      if (lane == 0) mbar_arrive(&o_empty[bp]);
      if (orow_prev >= 0) {
        const float inv = 1.0f / l_prev;
        bf16* dst = p.o + orow_prev * p.ldo + h_prev * HD;
#pragma unroll
        for (int c = 0; c < HD; c += 8) {
          uint4 v;
          v.x = pack_bf16(__uint_as_float(o[c]) * inv, __uint_as_float(o[c + 1]) * inv);
          v.y = pack_bf16(__uint_as_float(o[c + 2]) * inv, __uint_as_float(o[c + 3]) * inv);
          v.z = pack_bf16(__uint_as_float(o[c + 4]) * inv, __uint_as_float(o[c + 5]) * inv);
          v.w = pack_bf16(__uint_as_float(o[c + 6]) * inv, __uint_as_float(o[c + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + c) = v;
        }
        if (p.lse) p.lse[lse_prev] = (m_prev + sk_lg2(l_prev)) * SK_LN2;
      }
    };

#pragma unroll 1
    for (int it = 0; it < n_my; ++it) {
      const int item = blockIdx.x + it * gridDim.x;
      const int tg = item % geo.n_tiles, h = item / geo.n_tiles;
      const int g0 = tg * geo.gpt, g = g0 + gi;
      const bool valid = gi < geo.gpt && g < p.G;
      const int b = it & 1;
      sMsk[b * 128 + row] = pk_mask(p, geo, g0, row);
      sk_bar_sync(1, 128);

      sk_wait(&s_full[b], (it >> 1) & 1, 501, it);
      tc_fence_after();
      uint32_t v[64];
      {
        const uint32_t sbase = lane_addr + (b ? Cfg::FS_COL1 : Cfg::FS_COL0) + kcw;
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) sk_tmem_ld8(sbase + c8 * 8, v + c8 * 8);
        if (two) {
          uint32_t w[64];
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8) sk_tmem_ld8(sbase + L + c8 * 8, w + c8 * 8);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = second ? w[j] : v[j];
        }
      }
      tmem_ld_wait();
      tc_fence_before();
      if (lane == 0) mbar_arrive(&s_empty[b]);
      float mx = -1e30f;
      const float* mrow = sMsk + b * 128 + kc0;
#pragma unroll
      for (int j = 0; j < 64; j += 4) {
        float x[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          x[e] = (j + e < L) ? fmaf(__uint_as_float(v[j + e]), scale2, mrow[(j + e) & 127]) : -1e30f;
          v[j + e] = __float_as_uint(x[e]);
        }
        mx = fmaxf(mx, fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])));
      }
      if (it > 0) sk_wait(&o_full[b ^ 1], ((it - 1) >> 1) & 1, 502, it);  // P V of the previous item has consumed P
      float sum = 0.f;
      const unsigned long long didx0 = DROPOUT ? ((static_cast<unsigned long long>(g) * p.nH + h) * L + qi) * L : 0ull;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        if (c8 * 8 < L && gi < geo.gpt) {
          float pr[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            pr[e] = valid ? ex2_approx(__uint_as_float(v[c8 * 8 + e]) - mx) : 0.f;
            sum += pr[e];
            if (DROPOUT) pr[e] = dropout_keep(p.seed, didx0 + c8 * 8 + e, p.drop_p) ? pr[e] * keep_inv : 0.f;
          }
          uint4 o;
          o.x = pack_bf16(pr[0], pr[1]); o.y = pack_bf16(pr[2], pr[3]);
          o.z = pack_bf16(pr[4], pr[5]); o.w = pack_bf16(pr[6], pr[7]);
          const int piece = (kc0 >> 3) + c8;
          *reinterpret_cast<uint4*>(p_row + (piece >> 3) * Cfg::CHUNK + (((piece & 7) ^ xr) << 4)) = o;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);

      if (it > 0) epilogue(it - 1);
      m_prev = mx;
      l_prev = valid ? sum : 1.f;
      h_prev = h;
      orow_prev = valid ? static_cast<long long>(g) * L + qi : -1;
      lse_prev = (static_cast<long long>(g) * p.nH + h) * L + qi;
    }
    if (n_my > 0) epilogue(n_my - 1);
  } else if (warp == SK_WARP_MMA) {
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, HD, 0, 1);
    auto issue_s = [&](int it) {
      const int s = it & 1;
      sk_wait(&full[s], (it >> 1) & 1, 503, it);
      sk_wait(&s_empty[s], ((it >> 1) & 1) ^ 1, 504, it);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t q_addr = smem_u32(smem + s * Cfg::F_STAGE), k_addr = q_addr + Cfg::T_BYTES;
        const uint32_t d = tmem_base + (s ? Cfg::FS_COL1 : Cfg::FS_COL0);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          umma_f16_ss(d, sk_desc_kmajor<HD>(q_addr, ks), sk_desc_kmajor<HD>(k_addr, ks), idesc_s, ks);
        umma_commit(&s_full[s]);
      }
      __syncwarp();
    };
    if (n_my > 0) issue_s(0);
#pragma unroll 1
    for (int it = 0; it < n_my; ++it) {
      if (it + 1 < n_my) issue_s(it + 1);
      const int b = it & 1;
      sk_wait(p_ready, it & 1, 505, it);
      sk_wait(&o_empty[b], ((it >> 1) & 1) ^ 1, 507, it);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t v_addr = smem_u32(smem + b * Cfg::F_STAGE + 2 * Cfg::T_BYTES), p_addr = smem_u32(sP);
        const uint32_t d = tmem_base + (b ? Cfg::FO_COL1 : Cfg::FO_COL0);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)  // 16 keys per step, two 64-key chunks
          umma_f16_ss(d, umma_desc_sw128(p_addr + (kk >> 2) * Cfg::CHUNK + (kk & 3) * 32, 16, 1024), sk_desc_mnmajor<HD>(v_addr, kk),
                      idesc_o, kk);
        umma_commit(&o_full[b]);
        umma_commit(&stage_free[b]);
      }
      __syncwarp();
    }
  } else {
    if (lane == 0) {
#pragma unroll 1
      for (int it = 0; it < n_my; ++it) {
        const int item = blockIdx.x + it * gridDim.x;
        const int tg = item % geo.n_tiles, h = item / geo.n_tiles;
        const int row0 = tg * geo.gpt * L;
        const int s = it & 1;
        if (it >= 2) sk_wait(&stage_free[s], ((it >> 1) - 1) & 1, 506, it);
        uint8_t* st = smem + s * Cfg::F_STAGE;
        mbar_arrive_expect_tx(&full[s], Cfg::F_STAGE);
        tma_load_2d(st, &tmQ, &full[s], h * HD, row0);
        tma_load_2d(st + Cfg::T_BYTES, &tmK, &full[s], h * HD, row0);
        tma_load_2d(st + 2 * Cfg::T_BYTES, &tmV, &full[s], h * HD, row0);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == SK_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int HD, bool DROPOUT>
__global__ void __launch_bounds__(SK_THREADS, 1) attn_pk_bwd_kernel(const AttnParams p, const __grid_constant__ CUtensorMap tmQ,
                                                                   const __grid_constant__ CUtensorMap tmdO,
                                                                   const __grid_constant__ CUtensorMap tmO,
                                                                   const __grid_constant__ CUtensorMap tmK,
                                                                   const __grid_constant__ CUtensorMap tmV, const PkGeo geo,
                                                                   int n_items) {
  using Cfg = PkCfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sP = smem + Cfg::B_OFF_P;
  uint8_t* sdS = smem + Cfg::B_OFF_DS;
  float* sMsk = reinterpret_cast<float*>(smem + Cfg::B_OFF_MSK);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::B_OFF_BARS);
  // timeout tags: 601 full, 602 stage_free, 603 s_full, 604 sdp_empty, 605 pds_ready, 606 out_full, 607 acc_empty
  uint64_t* full = bars;
  uint64_t* stage_free = bars + 2;
  uint64_t* s_full = bars + 4;
  uint64_t* sdp_empty = bars + 5;
  uint64_t* pds_ready = bars + 6;
  uint64_t* out_full = bars + 7;
  uint64_t* acc_empty = bars + 8;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 9);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_my = (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const float scale2 = p.scale * SK_LOG2E;
  const int L = geo.L;

  pdl_trigger();
  if (warp == SK_WARP_MMA) {
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&stage_free[s], 1);
      }
      mbar_init(s_full, 1);
      mbar_init(sdp_empty, 4);
      mbar_init(pds_ready, 4);
      mbar_init(out_full, 1);
      mbar_init(acc_empty, 4);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  if (warp == SK_WARP_LD && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  for (int i = tid; i < 4 * Cfg::CHUNK / 16; i += SK_THREADS) reinterpret_cast<uint4*>(sdS)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();

  if (warp < 4) {
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const int gi = row / L, qi = row - gi * L;
    const int kc0 = gi < geo.gpt ? gi * L : 0;
    const int gi_lo = (warp * 32) / L;
    const int kcw = gi_lo < geo.gpt ? gi_lo * L : 0;
    const bool two = (warp * 32 + 31) / L != gi_lo;
    const bool second = gi != gi_lo;
    uint8_t* p_row = sP + row * 128;
    uint8_t* ds_row = sdS + row * 128;
    const int xr = row & 7;
    const float keep_inv = DROPOUT ? 1.0f / (1.0f - p.drop_p) : 1.0f;
#pragma unroll 1
    for (int u = 0; u < n_my; ++u) {
      const int item = blockIdx.x + u * gridDim.x;
      const int tg = item % geo.n_tiles, h = item / geo.n_tiles;
      const int g0 = tg * geo.gpt, g = g0 + gi;
      const bool valid = gi < geo.gpt && g < p.G;
      const int s = u & 1;
      const uint8_t* stage = smem + Cfg::B_OFF_STAGES + s * Cfg::B_STAGE;
      sMsk[s * 128 + row] = pk_mask(p, geo, g0, row);
      const float nlse2 = valid ? -p.lse[(static_cast<long long>(g) * p.nH + h) * L + qi] * SK_LOG2E : 0.f;
      sk_wait(&full[s], (u >> 1) & 1, 601, u);
      const float D = sk_row_dot<HD>(stage + Cfg::T_BYTES, stage + 2 * Cfg::T_BYTES, row);
      sk_bar_sync(1, 128);

      sk_wait(s_full, u & 1, 603, u);
      tc_fence_after();
      const unsigned long long didx0 = DROPOUT ? ((static_cast<unsigned long long>(g) * p.nH + h) * L + qi) * L : 0ull;
      const float* mrow = sMsk + s * 128 + kc0;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t sv[32], dv[32];
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          sk_tmem_ld8(lane_addr + Cfg::BS_COL + kcw + hf * 32 + c8 * 8, sv + c8 * 8);
          sk_tmem_ld8(lane_addr + Cfg::BDP_COL + kcw + hf * 32 + c8 * 8, dv + c8 * 8);
        }
        if (two) {
          uint32_t sw[32], dw[32];
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            sk_tmem_ld8(lane_addr + Cfg::BS_COL + kcw + L + hf * 32 + c8 * 8, sw + c8 * 8);
            sk_tmem_ld8(lane_addr + Cfg::BDP_COL + kcw + L + hf * 32 + c8 * 8, dw + c8 * 8);
          }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            sv[j] = second ? sw[j] : sv[j];
            dv[j] = second ? dw[j] : dv[j];
          }
        }
        tmem_ld_wait();
        if (hf == 1) {
          tc_fence_before();
          if (lane == 0) mbar_arrive(sdp_empty);
        }
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          if ((hf * 4 + c8) * 8 < L && gi < geo.gpt) {
            float pd[8], ds[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int j = hf * 32 + c8 * 8 + e;
              const float x = fmaf(__uint_as_float(sv[c8 * 8 + e]), scale2, mrow[j]) + nlse2;
              const float pr = valid ? ex2_approx(x) : 0.f;
              float dpv = __uint_as_float(dv[c8 * 8 + e]);
              pd[e] = pr;
              if (DROPOUT) {
                const bool keep = dropout_keep(p.seed, didx0 + j, p.drop_p);
                pd[e] = keep ? pr * keep_inv : 0.f;
                dpv = keep ? dpv * keep_inv : 0.f;
              }
              ds[e] = pr * (dpv - D);
            }
            uint4 o, d;
            o.x = pack_bf16(pd[0], pd[1]); o.y = pack_bf16(pd[2], pd[3]);
            o.z = pack_bf16(pd[4], pd[5]); o.w = pack_bf16(pd[6], pd[7]);
            d.x = pack_bf16(ds[0], ds[1]); d.y = pack_bf16(ds[2], ds[3]);
            d.z = pack_bf16(ds[4], ds[5]); d.w = pack_bf16(ds[6], ds[7]);
            const int piece = (kc0 >> 3) + hf * 4 + c8;
            const int off = (piece >> 3) * Cfg::CHUNK + (((piece & 7) ^ xr) << 4);
            *reinterpret_cast<uint4*>(p_row + off) = o;
            *reinterpret_cast<uint4*>(ds_row + off) = d;
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_ready);

      sk_wait(out_full, u & 1, 606, u);
      tc_fence_after();
      // thread = query row for dQ and key row for dV / dK: the same global row g0 L + row of the three outputs
      const long long grow = static_cast<long long>(g0) * L + row;
#pragma unroll
      for (int w = 0; w < 3; ++w) {
        uint32_t a[HD];
        const uint32_t col = w == 0 ? Cfg::BDQ_COL : (w == 1 ? Cfg::BDV_COL : Cfg::BDK_COL);
        tmem_ld32(lane_addr + col, reinterpret_cast<uint32_t(&)[32]>(a[0]));
        if constexpr (HD == 64) tmem_ld32(lane_addr + col + 32, reinterpret_cast<uint32_t(&)[32]>(a[32]));
        tmem_ld_wait();
        if (valid) {
          const float sc = w == 1 ? 1.0f : p.scale;
          bf16* dst = (w == 0 ? p.dq + grow * p.lddq : (w == 1 ? p.dv + grow * p.lddv : p.dk + grow * p.lddk)) + h * HD;
#pragma unroll
          for (int c = 0; c < HD; c += 8) {
            uint4 v;
            v.x = pack_bf16(__uint_as_float(a[c]) * sc, __uint_as_float(a[c + 1]) * sc);
            v.y = pack_bf16(__uint_as_float(a[c + 2]) * sc, __uint_as_float(a[c + 3]) * sc);
            v.z = pack_bf16(__uint_as_float(a[c + 4]) * sc, __uint_as_float(a[c + 5]) * sc);
            v.w = pack_bf16(__uint_as_float(a[c + 6]) * sc, __uint_as_float(a[c + 7]) * sc);
            *reinterpret_cast<uint4*>(dst + c) = v;
          }
        }
      }
      tc_fence_before();
      if (lane == 0) mbar_arrive(acc_empty);
    }
  } else if (warp == SK_WARP_MMA) {
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_kv = umma_idesc_bf16(128, HD, 1, 1);
    constexpr uint32_t idesc_q = umma_idesc_bf16(128, HD, 0, 1);
    const uint32_t p_addr = smem_u32(sP), ds_addr = smem_u32(sdS);
    auto issue_scores = [&](int u) {
      const int s = u & 1;
      sk_wait(&full[s], (u >> 1) & 1, 601, u);
      sk_wait(sdp_empty, (u & 1) ^ 1, 604, u);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t q_addr = smem_u32(smem + Cfg::B_OFF_STAGES + s * Cfg::B_STAGE), do_addr = q_addr + Cfg::T_BYTES;
        const uint32_t k_addr = q_addr + 3 * Cfg::T_BYTES, v_addr = q_addr + 4 * Cfg::T_BYTES;
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          umma_f16_ss(tmem_base + Cfg::BS_COL, sk_desc_kmajor<HD>(q_addr, ks), sk_desc_kmajor<HD>(k_addr, ks), idesc_s, ks);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          umma_f16_ss(tmem_base + Cfg::BDP_COL, sk_desc_kmajor<HD>(do_addr, ks), sk_desc_kmajor<HD>(v_addr, ks), idesc_s, ks);
        umma_commit(s_full);
      }
      __syncwarp();
    };
    if (n_my > 0) issue_scores(0);
#pragma unroll 1
    for (int u = 0; u < n_my; ++u) {
      const int s = u & 1;
      sk_wait(pds_ready, u & 1, 605, u);
      sk_wait(acc_empty, (u & 1) ^ 1, 607, u);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t q_addr = smem_u32(smem + Cfg::B_OFF_STAGES + s * Cfg::B_STAGE), do_addr = q_addr + Cfg::T_BYTES;
        const uint32_t k_addr = q_addr + 3 * Cfg::T_BYTES;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // 16 query rows per step; M = the tile's 128 key rows (two chunks)
          umma_f16_ss(tmem_base + Cfg::BDV_COL, umma_desc(p_addr + kk * 2048, Cfg::CHUNK, 1024, LAYOUT_SW128),
                      sk_desc_mnmajor<HD>(do_addr, kk), idesc_kv, kk);
          umma_f16_ss(tmem_base + Cfg::BDK_COL, umma_desc(ds_addr + kk * 2048, Cfg::CHUNK, 1024, LAYOUT_SW128),
                      sk_desc_mnmajor<HD>(q_addr, kk), idesc_kv, kk);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)  // 16 keys per step
          umma_f16_ss(tmem_base + Cfg::BDQ_COL, umma_desc_sw128(ds_addr + (kk >> 2) * Cfg::CHUNK + (kk & 3) * 32, 16, 1024),
                      sk_desc_mnmajor<HD>(k_addr, kk), idesc_q, kk);
        umma_commit(out_full);
        umma_commit(&stage_free[s]);
      }
      __syncwarp();
      if (u + 1 < n_my) issue_scores(u + 1);
    }
  } else {
    if (lane == 0) {
#pragma unroll 1
      for (int u = 0; u < n_my; ++u) {
        const int item = blockIdx.x + u * gridDim.x;
        const int tg = item % geo.n_tiles, h = item / geo.n_tiles;
        const int row0 = tg * geo.gpt * L;
        const int s = u & 1;
        if (u >= 2) sk_wait(&stage_free[s], ((u >> 1) - 1) & 1, 602, u);
        uint8_t* st = smem + Cfg::B_OFF_STAGES + s * Cfg::B_STAGE;
        mbar_arrive_expect_tx(&full[s], Cfg::B_STAGE);
        tma_load_2d(st, &tmQ, &full[s], h * HD, row0);
        tma_load_2d(st + Cfg::T_BYTES, &tmdO, &full[s], h * HD, row0);
        tma_load_2d(st + 2 * Cfg::T_BYTES, &tmO, &full[s], h * HD, row0);
        tma_load_2d(st + 3 * Cfg::T_BYTES, &tmK, &full[s], h * HD, row0);
        tma_load_2d(st + 4 * Cfg::T_BYTES, &tmV, &full[s], h * HD, row0);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == SK_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- host ---------------------------------------------------------------------------------------------------------
typedef CUresult (*SkEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
SkEncodeTiledFn sk_encode_fn() {
  static SkEncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<SkEncodeTiledFn>(sym);
  });
  return fn;
}

// 2-D map over a row-major [rows, cols] bf16 activation (row pitch ld elements): box = hd columns x box_rows rows
int sk_tmap(CUtensorMap* out, const bf16* base, long long ld, long long rows, int cols, int hd, int box_rows) {
  SkEncodeTiledFn fn = sk_encode_fn();
  FIBER_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(hd), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, hd == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FIBER_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (attention tile) failed with %d", static_cast<int>(r));
  return 0;
}

template <int HD, bool DROPOUT>
int launch_sk_fwd_t(const AttnParams& p, cudaStream_t stream) {
  using Cfg = SkFwdCfg<HD>;
  auto kern = attn_sk_fwd_kernel<HD, DROPOUT>;
  static bool attr_set = false;
  if (!attr_set) {
    FIBER_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  CUtensorMap tq, tk, tv;
  const int C = p.nH * HD;
  if (sk_tmap(&tq, p.q, p.ldq, static_cast<long long>(p.G) * p.Lq, C, HD, 128) ||
      sk_tmap(&tk, p.k, p.ldk, static_cast<long long>(p.G) * p.Lk, C, HD, SK_KEYS) ||
      sk_tmap(&tv, p.v, p.ldv, static_cast<long long>(p.G) * p.Lk, C, HD, SK_KEYS))
    return -1;
  const int ntiles = (p.Lq + 127) / 128;
  const long long n_items = static_cast<long long>(p.G) * p.nH * ntiles;
  FIBER_CHECK(n_items < (1ll << 31), "too many attention work items");
  const int grid = static_cast<int>(n_items < 2ll * num_sms() ? n_items : 2ll * num_sms());
  FIBER_CUDA(launch_k(kern, dim3(grid), dim3(SK_THREADS), Cfg::SMEM, stream, p, tq, tk, tv, ntiles, static_cast<int>(n_items)));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

template <int HD, bool DROPOUT>
int launch_sk_bwd_t(const AttnParams& p, cudaStream_t stream) {
  using Cfg = SkBwdCfg<HD>;
  auto kern = attn_sk_bwd_kernel<HD, DROPOUT>;
  static bool attr_set = false;
  if (!attr_set) {
    FIBER_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  CUtensorMap tq, tdo, to, tk, tv;
  const int C = p.nH * HD;
  const long long qrows = static_cast<long long>(p.G) * p.Lq, krows = static_cast<long long>(p.G) * p.Lk;
  if (sk_tmap(&tq, p.q, p.ldq, qrows, C, HD, 128) || sk_tmap(&tdo, p.d_o, p.lddo, qrows, C, HD, 128) ||
      sk_tmap(&to, p.o, p.ldo, qrows, C, HD, 128) || sk_tmap(&tk, p.k, p.ldk, krows, C, HD, SK_KEYS) ||
      sk_tmap(&tv, p.v, p.ldv, krows, C, HD, SK_KEYS))
    return -1;
  const int ntiles = (p.Lq + 127) / 128;
  const long long n_items = static_cast<long long>(p.G) * p.nH;
  FIBER_CHECK(n_items * ntiles < (1ll << 31), "too many attention work items");
  const long long slots = static_cast<long long>(Cfg::CTAS_PER_SM) * num_sms();
  const int grid = static_cast<int>(n_items < slots ? n_items : slots);
  FIBER_CUDA(launch_k(kern, dim3(grid), dim3(SK_THREADS), Cfg::SMEM, stream, p, tq, tdo, to, tk, tv, ntiles,
                      static_cast<int>(n_items)));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

static PkGeo pk_geo(const AttnParams& p) {
  PkGeo geo;
  geo.L = p.Lq;
  geo.gpt = 128 / p.Lq;
  geo.n_tiles = (p.G + geo.gpt - 1) / geo.gpt;
  return geo;
}

template <int HD, bool DROPOUT>
int launch_pk_fwd_t(const AttnParams& p, cudaStream_t stream) {
  using Cfg = PkCfg<HD>;
  auto kern = attn_pk_fwd_kernel<HD, DROPOUT>;
  static bool attr_set = false;
  if (!attr_set) {
    FIBER_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::F_SMEM));
    attr_set = true;
  }
  CUtensorMap tq, tk, tv;
  const int C = p.nH * HD;
  const long long rows = static_cast<long long>(p.G) * p.Lq;
  if (sk_tmap(&tq, p.q, p.ldq, rows, C, HD, 128) || sk_tmap(&tk, p.k, p.ldk, rows, C, HD, 128) ||
      sk_tmap(&tv, p.v, p.ldv, rows, C, HD, 128))
    return -1;
  const PkGeo geo = pk_geo(p);
  const long long n_items = static_cast<long long>(geo.n_tiles) * p.nH;
  FIBER_CHECK(n_items < (1ll << 31), "too many attention work items");
  const int grid = static_cast<int>(n_items < num_sms() ? n_items : num_sms());
  FIBER_CUDA(launch_k(kern, dim3(grid), dim3(SK_THREADS), Cfg::F_SMEM, stream, p, tq, tk, tv, geo, static_cast<int>(n_items)));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

template <int HD, bool DROPOUT>
int launch_pk_bwd_t(const AttnParams& p, cudaStream_t stream) {
  using Cfg = PkCfg<HD>;
  auto kern = attn_pk_bwd_kernel<HD, DROPOUT>;
  static bool attr_set = false;
  if (!attr_set) {
    FIBER_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::B_SMEM));
    attr_set = true;
  }
  CUtensorMap tq, tdo, to, tk, tv;
  const int C = p.nH * HD;
  const long long rows = static_cast<long long>(p.G) * p.Lq;
  if (sk_tmap(&tq, p.q, p.ldq, rows, C, HD, 128) || sk_tmap(&tdo, p.d_o, p.lddo, rows, C, HD, 128) ||
      sk_tmap(&to, p.o, p.ldo, rows, C, HD, 128) || sk_tmap(&tk, p.k, p.ldk, rows, C, HD, 128) ||
      sk_tmap(&tv, p.v, p.ldv, rows, C, HD, 128))
    return -1;
  const PkGeo geo = pk_geo(p);
  const long long n_items = static_cast<long long>(geo.n_tiles) * p.nH;
  FIBER_CHECK(n_items < (1ll << 31), "too many attention work items");
  const int grid = static_cast<int>(n_items < num_sms() ? n_items : num_sms());
  FIBER_CUDA(launch_k(kern, dim3(grid), dim3(SK_THREADS), Cfg::B_SMEM, stream, p, tq, tdo, to, tk, tv, geo,
                      static_cast<int>(n_items)));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace

static bool sk_aligned16(const void* ptr, long long ld) {
  return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ld % 8 == 0;
}

// plain mode, at most 64 keys per group, 16-byte addressable rows
bool attn_sk_supported(const AttnParams& p, int hd) {
  return p.mode == 0 && (hd == 32 || hd == 64) && p.Lk <= SK_KEYS && sk_aligned16(p.q, p.ldq) && sk_aligned16(p.k, p.ldk) &&
         sk_aligned16(p.v, p.ldv) && sk_aligned16(p.o, p.ldo);
}

// backward: additionally d_o / dq / dk / dv rows 16-byte addressable
bool attn_sk_bwd_supported(const AttnParams& p, int hd) {
  return attn_sk_supported(p, hd) && p.d_o && p.dq && p.dk && p.dv && p.lse && sk_aligned16(p.d_o, p.lddo) &&
         sk_aligned16(p.dq, p.lddq) && sk_aligned16(p.dk, p.lddk) && sk_aligned16(p.dv, p.lddv);
}

// packed self-attention: Lq == Lk == L, 32 <= L <= 64, L % 8 == 0 (two or three whole groups per 128-row tile)
// (L >= 32: a warp's rows then touch at most two groups, see the kernels)
bool attn_pk_shape(const AttnParams& p) { return p.Lq == p.Lk && p.Lq >= 32 && p.Lq <= 64 && p.Lq % 8 == 0; }

int launch_attn_sk_bwd(const AttnParams& p, int hd, cudaStream_t stream) {
  if (attn_pk_shape(p)) {
    if (hd == 32) return p.drop_p > 0.f ? launch_pk_bwd_t<32, true>(p, stream) : launch_pk_bwd_t<32, false>(p, stream);
    return p.drop_p > 0.f ? launch_pk_bwd_t<64, true>(p, stream) : launch_pk_bwd_t<64, false>(p, stream);
  }
  if (hd == 32) return p.drop_p > 0.f ? launch_sk_bwd_t<32, true>(p, stream) : launch_sk_bwd_t<32, false>(p, stream);
  return p.drop_p > 0.f ? launch_sk_bwd_t<64, true>(p, stream) : launch_sk_bwd_t<64, false>(p, stream);
}

int launch_attn_sk_fwd(const AttnParams& p, int hd, cudaStream_t stream) {
  if (attn_pk_shape(p)) {
    if (hd == 32) return p.drop_p > 0.f ? launch_pk_fwd_t<32, true>(p, stream) : launch_pk_fwd_t<32, false>(p, stream);
    return p.drop_p > 0.f ? launch_pk_fwd_t<64, true>(p, stream) : launch_pk_fwd_t<64, false>(p, stream);
  }
  if (hd == 32) return p.drop_p > 0.f ? launch_sk_fwd_t<32, true>(p, stream) : launch_sk_fwd_t<32, false>(p, stream);
  return p.drop_p > 0.f ? launch_sk_fwd_t<64, true>(p, stream) : launch_sk_fwd_t<64, false>(p, stream);
}

}  // namespace fiber
