// fiber_b200 — persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = epilogue( A[M,K] . B[N,K]^T )      bf16 operands, fp32 accumulation in TMEM
//
// One CTA per SM, 18 warps:
//   warp 0      TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1      MMA issuer     (one lane issues tcgen05.mma 128 x BN x 16, commits to mbarriers)
//   warps 2..17 epilogue       (thread = accumulator row: tcgen05.ld TMEM -> registers -> fused bias /
//                               GELU / GELU' / gate / DropPath / residual -> bf16 pack -> 64B-swizzled
//                               smem box -> TMA store; fp32 / atomic outputs store from registers)
// TMEM holds two BN-column fp32 accumulators so the epilogue of tile i overlaps the MMAs of
// tile i+1.  Operands may be K-major (forward, dgrad) or MN-major (wgrad: dW = dY^T X, both
// operands read straight from their row-major activations, no transposes materialised).
// Split-K work units accumulate with fp32 atomics (wgrad only).
#include "common.cuh"
#include "../../include/fiber_b200.h"

#include <cstdlib>
#include <mutex>

namespace fiber {

void count_launch(int n = 1);
void count_gemm_cta2_launch();
int option_gemm_cta2();  // capi.cu: bit 0 = CTA pairs for the default epilogues, bit 1 = for the two-box (EPI = 1) epilogues

struct GemmParams {
  int M, N, K;
  int tiles_m, tiles_n, splits, kb_total, kb_per_split;
  void* c;
  long long ldc;
  const float* bias;
  const bf16* residual;
  long long ldr;
  const bf16* aux;
  long long ldaux;
  bf16* preact;
  long long ldp;
  const float* scale;
  const float* row_scale;
  int rows_per_scale;
  int act;
  int out_mode;
  int tma_store;  // epilogue variant: thread=row math -> swizzled smem box -> TMA store (no residual/aux)
  float* colsum;  // MN-major only: colsum[m] += scale * sum_k A[k, m] (extra N=16 MMA against a tile of ones)
  // Device-side row count (or null): only the first *row_count rows of the activation operand carry work — K-major:
  // output row tiles past it are skipped; MN-major (wgrad): reduction k-blocks past it are skipped.
  const int* row_count;
  // EPI == 2 (fused MLM decoder + cross-entropy, fiber_mlm_ce_fwd / _bwd)
  const int* ce_labels;  // [M] target column, < 0 = ignored row
  float* ce_part;        // act 8 out: [tiles_n * 4][ce_mpad] float4 (max2, sum2, best2, argmax bits), log2 domain
  float* ce_xl;          // act 8 out: [M] logit at the label column (written by the one warp that owns it)
  const float* ce_lse2;  // act 9 in: [M] log2-domain log-sum-exp
  const float* ce_g;     // act 9 in: device scalar d(loss) / n_valid
  int ce_mpad;
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 576;  // TMA warp, MMA warp, 16 epilogue warps

// EPI = 1 (two-output GELU epilogues, aux / residual epilogues): these GEMMs move 1.2 - 1.4 GB per launch and were
// bound by the epilogue waiting, four times per tile and warp, for the TMA unit to finish READING the warp's only box
// before the next result could be staged (6.6 us per tile against 2.3 us of MMAs).  They trade one operand stage for a
// second box per warp: the two outputs of a chunk (and consecutive chunks) alternate boxes, and a box is rewritten
// only when the store issued two stores ago has been read (cp.async.bulk.wait_group.read 1).
// CTA2 = 1 (K-major, BN = 256): the kernel runs as CTA pairs (2-CTA clusters, tcgen05 cta_group::2).  A pair owns a
// 256 x 256 output tile: each CTA loads its own 128 rows of A and HALF of the B tile (128 of the 256 weight rows), the
// leader issues 256 x 256 x 16 MMAs that read both halves, and each CTA drains its own 128 x 256 accumulator.  A stage
// is 32 KB instead of 48 KB per CTA: a third less L2 -> shared-memory traffic per flop (what bounds the 128 x 256
// single-CTA tile at ~1.3 PFLOP/s) and two more stages in the same shared memory.
template <int BN, int EPI = 0, int CTA2 = 0>
struct GemmCfg {
  static constexpr int STAGES = CTA2 ? (EPI == 1 ? 5 : 6) : (BN == 256) ? (EPI == 1 ? 3 : 4) : (EPI == 1 ? 5 : 6);
  static constexpr int BOXES = EPI == 1 ? 2 : 1;  // EPI == 2 (cross-entropy epilogues) keeps the EPI == 0 budget
  static constexpr uint32_t A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr uint32_t B_BYTES = (CTA2 ? BN / 2 : BN) * GEMM_BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t STAGING_BYTES = 16 * 2048 * BOXES;  // 16 epilogue warps x BOXES x (32 rows x 64 B) TMA-store box
  static constexpr uint32_t SMEM_BYTES =
      1024 /*align slack*/ + STAGES * STAGE_BYTES + STAGING_BYTES + 256 /*barriers*/;
  static constexpr uint32_t TMEM_COLS = 2 * BN;
};

// One 32-row x 32-column chunk of the TMA-store epilogue with every epilogue option fixed at compile time
// (the generic runtime-flag loop below spends half of its issue slots on branches and predicates).
// Preconditions (checked by the caller, warp-uniform): all 32 rows and all 32 columns are in range.
//   ACT 0 none / 1 erf-GELU / 2 multiply by GELU'(aux);  BIAS: + bias[col];  SCALE: * sc;  RES: + residual
template <int ACT, bool BIAS, bool SCALE, bool RES>
__device__ __forceinline__ void epi_chunk_full(uint32_t taddr, uint8_t* box_row, int swz,
                                               const float* __restrict__ bias_c, const bf16* __restrict__ res_c,
                                               const bf16* __restrict__ aux_c, float sc) {
  uint4 rv[4], av[4];
  if (RES) {
#pragma unroll
    for (int j = 0; j < 4; ++j) rv[j] = *reinterpret_cast<const uint4*>(res_c + j * 8);
  }
  if (ACT == 2) {
#pragma unroll
    for (int j = 0; j < 4; ++j) av[j] = *reinterpret_cast<const uint4*>(aux_c + j * 8);
  }
  const uint64_t sc2 = pk2(sc, sc);
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    uint32_t r[16];
    tmem_ld16(taddr + hf * 16, r);
    float4 bv[4];
    if (BIAS) {
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = __ldg(reinterpret_cast<const float4*>(bias_c + hf * 16 + j * 4));
    }
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint64_t xp[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        xp[e] = pk2(__uint_as_float(r[j * 8 + 2 * e]), __uint_as_float(r[j * 8 + 2 * e + 1]));
      if (BIAS) {
        xp[0] = add2(xp[0], pk2(bv[2 * j].x, bv[2 * j].y)); xp[1] = add2(xp[1], pk2(bv[2 * j].z, bv[2 * j].w));
        xp[2] = add2(xp[2], pk2(bv[2 * j + 1].x, bv[2 * j + 1].y));
        xp[3] = add2(xp[3], pk2(bv[2 * j + 1].z, bv[2 * j + 1].w));
      }
      if (ACT == 1) gelu_erf2x4(xp);
      if (ACT == 2) {
        const uint32_t* au = reinterpret_cast<const uint32_t*>(&av[hf * 2 + j]);
        uint64_t hx[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_bf16(au[e]);
          hx[e] = pk2(f.x, f.y);
        }
        gelu_erf_grad_mul2x4(xp, hx);
      }
      if (RES) {
        const uint32_t* ru = reinterpret_cast<const uint32_t*>(&rv[hf * 2 + j]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_bf16(ru[e]);
          xp[e] = SCALE ? fma2(xp[e], sc2, pk2(f.x, f.y)) : add2(xp[e], pk2(f.x, f.y));
        }
      } else if (SCALE) {
#pragma unroll
        for (int e = 0; e < 4; ++e) xp[e] = mul2(xp[e], sc2);
      }
      float x[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) upk2(xp[e], x[2 * e], x[2 * e + 1]);
      uint4 o;
      o.x = pack_bf16(x[0], x[1]); o.y = pack_bf16(x[2], x[3]);
      o.z = pack_bf16(x[4], x[5]); o.w = pack_bf16(x[6], x[7]);
      *reinterpret_cast<uint4*>(box_row + (((hf * 2 + j) ^ swz) << 4)) = o;
    }
  }
}

// ---- opt-in epilogue (kernel template parameter EPI = 1; fiber_gemm_args.act 3 / 4) ---------------------------
// act 3 (fc1): C = GELU(acc + bias) and P = GELU'(acc + bias) from ONE pass over TMEM and one erfc per element; the
//              backward only ever needs GELU'(h), so it is stored instead of the pre-activation h.
// act 4 (fc2 dgrad): C = acc * aux — with aux = the stored GELU'(h) this replaces the 14-instruction-per-element
//              GELU' epilogue of act 2 by one multiply.
// One 32-row x 32-column chunk, thread = row, all rows / columns in range (checked on the host).
// act 5 (fc1, default backward): C = GELU(acc + bias) and P = acc + bias — what act 1 + preact computes in two passes
//              over TMEM, bit for bit, in one.
// Writes bf16 GELU into the warp's swizzled box and returns the second output (GRAD: GELU'(v), else v) packed as
// bf16 pairs, 16 bytes = 8 columns per group of four words.
template <bool GRAD>
__device__ __forceinline__ void epi_chunk_gelu_both(uint32_t taddr, uint8_t* box_row, int swz,
                                                    const float* __restrict__ bias_c, bool has_bias, uint32_t (&gp)[16]) {
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    uint32_t r[16];
    tmem_ld16(taddr + hf * 16, r);
    float4 bv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      bv[j] = has_bias ? __ldg(reinterpret_cast<const float4*>(bias_c + hf * 16 + j * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint64_t xp[4], gq[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        xp[e] = pk2(__uint_as_float(r[j * 8 + 2 * e]), __uint_as_float(r[j * 8 + 2 * e + 1]));
      xp[0] = add2(xp[0], pk2(bv[2 * j].x, bv[2 * j].y)); xp[1] = add2(xp[1], pk2(bv[2 * j].z, bv[2 * j].w));
      xp[2] = add2(xp[2], pk2(bv[2 * j + 1].x, bv[2 * j + 1].y));
      xp[3] = add2(xp[3], pk2(bv[2 * j + 1].z, bv[2 * j + 1].w));
      if (GRAD) {
        gelu_erf_both2x4(xp, gq);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) gq[e] = xp[e];  // the pre-activation itself
        gelu_erf2x4(xp);
      }
      float x[8], g[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        upk2(xp[e], x[2 * e], x[2 * e + 1]);
        upk2(gq[e], g[2 * e], g[2 * e + 1]);
      }
      uint4 o;
      o.x = pack_bf16(x[0], x[1]); o.y = pack_bf16(x[2], x[3]);
      o.z = pack_bf16(x[4], x[5]); o.w = pack_bf16(x[6], x[7]);
      *reinterpret_cast<uint4*>(box_row + (((hf * 2 + j) ^ swz) << 4)) = o;
      const int jb = hf * 2 + j;
      gp[jb * 4 + 0] = pack_bf16(g[0], g[1]); gp[jb * 4 + 1] = pack_bf16(g[2], g[3]);
      gp[jb * 4 + 2] = pack_bf16(g[4], g[5]); gp[jb * 4 + 3] = pack_bf16(g[6], g[7]);
    }
  }
}
// act 6 (proj / fc2 forward): C = (acc + bias) * sc + residual — the default residual epilogue (act 0 + residual [+ scale /
//              row_scale]), bit for bit, with the residual rows in registers one chunk ahead.
template <bool BIAS, bool SCALE>
__device__ __forceinline__ void epi_chunk_res(uint32_t taddr, uint8_t* box_row, int swz, const float* __restrict__ bias_c,
                                              const uint4 (&rv)[4], float sc) {
  const uint64_t sc2 = pk2(sc, sc);
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    uint32_t r[16];
    tmem_ld16(taddr + hf * 16, r);
    float4 bv[4];
    if (BIAS) {
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = __ldg(reinterpret_cast<const float4*>(bias_c + hf * 16 + j * 4));
    }
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint64_t xp[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        xp[e] = pk2(__uint_as_float(r[j * 8 + 2 * e]), __uint_as_float(r[j * 8 + 2 * e + 1]));
      if (BIAS) {
        xp[0] = add2(xp[0], pk2(bv[2 * j].x, bv[2 * j].y)); xp[1] = add2(xp[1], pk2(bv[2 * j].z, bv[2 * j].w));
        xp[2] = add2(xp[2], pk2(bv[2 * j + 1].x, bv[2 * j + 1].y));
        xp[3] = add2(xp[3], pk2(bv[2 * j + 1].z, bv[2 * j + 1].w));
      }
      const uint32_t* ru = reinterpret_cast<const uint32_t*>(&rv[hf * 2 + j]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16(ru[e]);
        xp[e] = SCALE ? fma2(xp[e], sc2, pk2(f.x, f.y)) : add2(xp[e], pk2(f.x, f.y));
      }
      float x[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) upk2(xp[e], x[2 * e], x[2 * e + 1]);
      uint4 o;
      o.x = pack_bf16(x[0], x[1]); o.y = pack_bf16(x[2], x[3]);
      o.z = pack_bf16(x[4], x[5]); o.w = pack_bf16(x[6], x[7]);
      *reinterpret_cast<uint4*>(box_row + (((hf * 2 + j) ^ swz) << 4)) = o;
    }
  }
}
// act 7 (fc2 dgrad, default numerics): C = acc * GELU'(aux) — what act 2 computes, bit for bit, with the aux rows
//              in registers before they are needed (below).
// One chunk of C = acc * aux (GRAD = false, act 4) or acc * GELU'(aux) (GRAD = true, act 7); av = the thread's 32 aux
// values, loaded by the caller one chunk ahead.
template <bool GRAD>
__device__ __forceinline__ void epi_chunk_mul_aux(uint32_t taddr, uint8_t* box_row, int swz, const uint4 (&av)[4]) {
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    uint32_t r[16];
    tmem_ld16(taddr + hf * 16, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint32_t* au = reinterpret_cast<const uint32_t*>(&av[hf * 2 + j]);
      float x[8];
      uint64_t xp[4], hx[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16(au[e]);
        xp[e] = pk2(__uint_as_float(r[j * 8 + 2 * e]), __uint_as_float(r[j * 8 + 2 * e + 1]));
        hx[e] = pk2(f.x, f.y);
      }
      if (GRAD) {
        gelu_erf_grad_mul2x4(xp, hx);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) xp[e] = mul2(xp[e], hx[e]);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) upk2(xp[e], x[2 * e], x[2 * e + 1]);
      uint4 o;
      o.x = pack_bf16(x[0], x[1]); o.y = pack_bf16(x[2], x[3]);
      o.z = pack_bf16(x[4], x[5]); o.w = pack_bf16(x[6], x[7]);
      *reinterpret_cast<uint4*>(box_row + (((hf * 2 + j) ^ swz) << 4)) = o;
    }
  }
}

__device__ __forceinline__ void prefetch_l2(const void* ptr) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
}

template <int BN, int MN_MAJOR, int EPI = 0, int CTA2 = 0>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmP,
                    const GemmParams p) {
  static_assert(!CTA2 || BN == 256, "CTA pairs: 256-column tiles");
  using Cfg = GemmCfg<BN, EPI, CTA2>;
  constexpr int NCTA = CTA2 ? 2 : 1;
  const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0u;  // 0 = leader (issues the MMAs, owns the full / tempty barriers)
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem;
  if constexpr (EPI == 1) {
    // alignment by pointer arithmetic on the __shared__ array: the compiler keeps the address space, so the epilogue's
    // box stores become STS.128 instead of generic ST.E.128 (to be carried over to EPI == 0 once measured)
    smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  } else {
    smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  }
  uint8_t* stage_base = smem;
  float* staging = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::STAGING_BYTES);
  uint64_t* full_bar = bars;                  // [STAGES]
  uint64_t* empty_bar = bars + STAGES;        // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;    // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  pdl_trigger();
  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tfull_bar[a], 1);
        mbar_init(&tempty_bar[a], 16 * NCTA);  // pair: the epilogue warps of both CTAs arrive on the leader's barrier
      }
      mbar_fence_init();
    }
    __syncwarp();
    if constexpr (CTA2) {
      tmem_alloc2(tmem_ptr, Cfg::TMEM_COLS);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  // Bias-gradient column sums (wgrad): a tile of bf16 ones in the (otherwise unused) TMA-store staging area is
  // the B operand of one extra 128 x 16 MMA per k-step; its accumulator sits behind the single main accumulator.
  const bool do_cs = MN_MAJOR == 1 && p.colsum != nullptr;
  if (do_cs) {
    uint32_t* ones = reinterpret_cast<uint32_t*>(staging);
    for (int i = threadIdx.x; i < 8192 / 4; i += GEMM_THREADS) ones[i] = 0x3F803F80u;
    fence_proxy_async_smem();
  }
  const int nacc = do_cs ? 1 : 2;  // accumulators in flight (the column-sum columns take the second one's place)
  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  else                __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();  // everything above touched only shared memory, TMEM and the kernel parameters

  int tiles_m = p.tiles_m, kb_total = p.kb_total;
  if (p.row_count) {
    const int cnt = __ldg(p.row_count);
    if (MN_MAJOR == 0) tiles_m = min(tiles_m, (cnt + GEMM_BM - 1) / GEMM_BM);
    else               kb_total = min(kb_total, (cnt + GEMM_BK - 1) / GEMM_BK);
  }
  if constexpr (CTA2) tiles_m /= 2;  // a pair's unit is a 256-row tile (host: M % 256 == 0, no row_count)
  const int tiles_mn = tiles_m * p.tiles_n;
  const int total_units = tiles_mn * p.splits;  // units whose k-range is empty (row_count) are skipped by every role
  const int unit0 = static_cast<int>(blockIdx.x) / NCTA, unit_step = static_cast<int>(gridDim.x) / NCTA;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = unit0; unit < total_units; unit += unit_step) {
        const int tile = unit % tiles_mn;
        const int split = unit / tiles_mn;
        const int m0 = ((tile / p.tiles_n) * NCTA + static_cast<int>(cta_rank)) * GEMM_BM;
        const int n0 = (tile % p.tiles_n) * BN;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = stage_base + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          if constexpr (CTA2) {
            // both CTAs' boxes complete on the LEADER's barrier, which expects the bytes of the whole pair
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
            const uint32_t bar = mapa_cta(smem_u32(&full_bar[stage]), 0);
            if constexpr (MN_MAJOR == 0) {
              tma_load_2d_2sm(sa, &tmA, bar, kb * GEMM_BK, m0);
              tma_load_2d_2sm(sb, &tmB, bar, kb * GEMM_BK, n0 + static_cast<int>(cta_rank) * (BN / 2));
            } else {  // wgrad: this CTA's 128 M columns and its half of the N columns, 64-wide chunks
#pragma unroll
              for (int j = 0; j < GEMM_BM / 64; ++j)
                tma_load_2d_2sm(sa + j * 8192, &tmA, bar, m0 + j * 64, kb * GEMM_BK);
#pragma unroll
              for (int j = 0; j < BN / 128; ++j)
                tma_load_2d_2sm(sb + j * 8192, &tmB, bar, n0 + static_cast<int>(cta_rank) * (BN / 2) + j * 64, kb * GEMM_BK);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if (MN_MAJOR == 0) {
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * GEMM_BK, m0);
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * GEMM_BK, n0);
          } else {
#pragma unroll
            for (int j = 0; j < GEMM_BM / 64; ++j)
              tma_load_2d(sa + j * 8192, &tmA, &full_bar[stage], m0 + j * 64, kb * GEMM_BK);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sb + j * 8192, &tmB, &full_bar[stage], n0 + j * 64, kb * GEMM_BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if constexpr (CTA2) {
        // the leader's multicast commits still arrive on this CTA's empty barriers after its last load was issued:
        // see every slot released before the CTA may leave the cluster
        for (int s = 0; s < STAGES; ++s) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (pair: the leader CTA only) =================
    constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM * NCTA, BN, MN_MAJOR, MN_MAJOR);
    constexpr uint32_t idesc_cs = umma_idesc_bf16(GEMM_BM * NCTA, 16, MN_MAJOR, MN_MAJOR);  // pair: 8 ones-columns per CTA
    const uint32_t ones_addr = smem_u32(staging);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase[2] = {0, 0};
    for (int unit = unit0; unit < total_units && cta_rank == 0; unit += unit_step) {
      const int split = unit / tiles_mn;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, kb_total);
      if (kb0 >= kb1) continue;
      const bool cs_unit = do_cs && ((unit % tiles_mn) % p.tiles_n == 0);  // first n-tile of its row
      // wait until the epilogue has drained this accumulator
      mbar_wait(&tempty_bar[acc], acc_phase[acc] ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(stage_base + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            uint64_t adesc, bdesc;
            if (MN_MAJOR == 0) {
              adesc = umma_desc_sw128(sa + k * 32, 16, 1024);
              bdesc = umma_desc_sw128(sb + k * 32, 16, 1024);
            } else {
              adesc = umma_desc_sw128(sa + k * 2048, 8192, 1024);
              bdesc = umma_desc_sw128(sb + k * 2048, 8192, 1024);
            }
            if constexpr (CTA2) umma_f16_ss2(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else                umma_f16_ss(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            if (MN_MAJOR == 1 && cs_unit) {
              if constexpr (CTA2)
                umma_f16_ss2(tmem_base + BN, adesc, umma_desc_sw128(ones_addr + k * 2048, 8192, 1024), idesc_cs,
                             (kb > kb0 || k > 0) ? 1u : 0u);
              else
                umma_f16_ss(tmem_base + BN, adesc, umma_desc_sw128(ones_addr + k * 2048, 8192, 1024), idesc_cs,
                            (kb > kb0 || k > 0) ? 1u : 0u);
            }
          }
          if constexpr (CTA2) {  // the same barrier offset in both CTAs of the pair
            umma_commit2_mc(&empty_bar[stage], 3);
            if (kb == kb1 - 1) umma_commit2_mc(&tfull_bar[acc], 3);
          } else {
            umma_commit(&empty_bar[stage]);            // frees the smem slot when the MMAs retire
            if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);  // accumulator ready for the epilogue
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      acc_phase[acc] ^= 1;
      acc = nacc == 2 ? acc ^ 1 : 0;
    }
  } else {
    // ================= epilogue (warps 2..17) =================
    // Four warps per TMEM lane quadrant, each owning a quarter of the accumulator columns.  Thread =
    // accumulator row: per 32-column chunk one tcgen05.ld, all fused math on registers (bias, erf-GELU,
    // GELU' * aux, gate, DropPath, residual; residual / aux rows are read directly, 64 contiguous bytes
    // per thread), bf16 pack, 16-byte stores into this warp's 32x32 64B-swizzled box, one TMA store per
    // box and output tensor.  fp32 / atomic outputs (wgrad, tiny fp32 tails) store straight from registers.
    const int ew = warp - 2;
    const int q = warp & 3;     // TMEM lane quadrant this warp may access
    const int cgrp = ew >> 2;   // which quarter of the BN columns
    constexpr int CPW = BN / 128;  // 32-column chunks per warp
    uint8_t* box = reinterpret_cast<uint8_t*>(staging) + ew * 2048 * Cfg::BOXES;  // BOXES x (32 rows x 64 B)
    int acc = 0;
    uint32_t acc_phase = 0;  // bit a = phase of accumulator a
    if constexpr (EPI == 1) {
      // opt-in epilogues (act 3 .. 7, see epi_chunk_gelu_both): full tiles, bf16 outputs through the TMA-store box
      const int act = p.act;
      const int swz = (lane >> 1) & 3;
      int bsel = 0;  // which of the warp's two boxes the next result goes to
      auto wait_box = [&] {  // the store issued two stores ago (the last user of box `bsel`) has been read
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
      };
      for (int unit = unit0; unit < total_units; unit += unit_step) {
        const int tile = unit % tiles_mn;
        const int m0 = ((tile / p.tiles_n) * NCTA + static_cast<int>(cta_rank)) * GEMM_BM;
        const int n0 = (tile % p.tiles_n) * BN;
        const int ncols = min(BN, p.N - n0);
        const long long row = static_cast<long long>(m0) + q * 32 + lane;
        const bool aux_mode = act == 4 || act == 6 || act == 7;  // a second bf16 [M, N] operand read per element
        float sc = 1.0f;  // act 6: gate alpha x DropPath scale of this row's sample (as the default epilogue)
        if (act == 6) {
          if (p.scale) sc = __ldg(p.scale);
          if (p.row_scale) sc *= __ldg(p.row_scale + static_cast<int>(row / p.rows_per_scale));
        }
        // The second operand (aux / residual) of a 32 x 32 chunk is loaded COALESCED — lane l fetches the 16-byte piece
        // l % 4 of rows l / 4 + 8 i, i = 0..3: eight 64-byte rows per instruction = 8 LSU wavefronts, where the
        // thread-per-row form (64 contiguous bytes per lane) costs 32 — one chunk ahead into registers, and is
        // transposed to thread-per-row through the warp's (still idle) output box.  The per-row loads were what bound
        // these epilogues: 4096 wavefronts per 128 x 256 tile, as many cycles as the tile's MMAs.
        const bf16* aux_base = nullptr;
        long long aux_ld = 0;
        if (aux_mode) {
          aux_base = act == 6 ? p.residual : p.aux;
          aux_ld = act == 6 ? p.ldr : p.ldaux;
          aux_base += (static_cast<long long>(m0) + q * 32 + (lane >> 2)) * aux_ld + n0 + (lane & 3) * 8;
        }
        uint4 av_next[4];
        if (aux_mode && cgrp * CPW * 32 < ncols) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            av_next[j] = *reinterpret_cast<const uint4*>(aux_base + static_cast<long long>(8 * j) * aux_ld + cgrp * CPW * 32);
        }
        mbar_wait(&tfull_bar[acc], (acc_phase >> acc) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int i = 0; i < CPW; ++i) {
          const int c0 = (cgrp * CPW + i) * 32;
          if (c0 >= ncols) break;
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c0;
          wait_box();
          uint8_t* bx = box + bsel * 2048;
          uint8_t* box_row = bx + lane * 64;
          if (aux_mode) {
            // registers (piece l % 4 of rows l / 4 + 8 j) -> box -> this thread's row
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int r = (lane >> 2) + 8 * j;
              *reinterpret_cast<uint4*>(bx + r * 64 + ((((lane & 3) ^ ((r >> 1) & 3))) << 4)) = av_next[j];
            }
            __syncwarp();
            uint4 av[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) av[j] = *reinterpret_cast<const uint4*>(box_row + ((j ^ swz) << 4));
            __syncwarp();  // every lane has read its row: the box may take the outputs
            if (i + 1 < CPW && c0 + 32 < ncols) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                av_next[j] = *reinterpret_cast<const uint4*>(aux_base + static_cast<long long>(8 * j) * aux_ld + c0 + 32);
            }
            if (act == 4) {
              epi_chunk_mul_aux<false>(taddr, box_row, swz, av);
            } else if (act == 7) {
              epi_chunk_mul_aux<true>(taddr, box_row, swz, av);
            } else {
              const bool scaled = p.scale != nullptr || p.row_scale != nullptr;
              const float* bias_c = p.bias + n0 + c0;
              if (p.bias) {
                if (scaled) epi_chunk_res<true, true>(taddr, box_row, swz, bias_c, av, sc);
                else        epi_chunk_res<true, false>(taddr, box_row, swz, bias_c, av, sc);
              } else {
                if (scaled) epi_chunk_res<false, true>(taddr, box_row, swz, bias_c, av, sc);
                else        epi_chunk_res<false, false>(taddr, box_row, swz, bias_c, av, sc);
              }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC, bx, n0 + c0, m0 + q * 32);
              tma_store_commit();
            }
            bsel ^= 1;
          } else {
            uint32_t gp[16];
            if (act == 3) epi_chunk_gelu_both<true>(taddr, box_row, swz, p.bias + n0 + c0, p.bias != nullptr, gp);
            else          epi_chunk_gelu_both<false>(taddr, box_row, swz, p.bias + n0 + c0, p.bias != nullptr, gp);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC, bx, n0 + c0, m0 + q * 32);
              tma_store_commit();
            }
            bsel ^= 1;
            wait_box();  // the second output goes to the other box
            uint8_t* bx2 = box + bsel * 2048;
#pragma unroll
            for (int jb = 0; jb < 4; ++jb)
              *reinterpret_cast<uint4*>(bx2 + lane * 64 + ((jb ^ swz) << 4)) =
                  make_uint4(gp[jb * 4 + 0], gp[jb * 4 + 1], gp[jb * 4 + 2], gp[jb * 4 + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmP, bx2, n0 + c0, m0 + q * 32);
              tma_store_commit();
            }
            bsel ^= 1;
          }
        }
        tc_fence_before();
        if (lane == 0) {
          if constexpr (CTA2) mbar_arrive_cluster(mapa_cta(smem_u32(&tempty_bar[acc]), 0));
          else                mbar_arrive(&tempty_bar[acc]);
        }
        acc_phase ^= (1u << acc);
        acc ^= 1;
      }
      if (lane == 0) tma_store_wait_read();  // smem must outlive the last bulk store
    } else if constexpr (EPI == 2) {
      // Fused MLM decoder + cross-entropy (heads.py:40-43 + objectives.py:19-26): the [rows, vocabulary] logits exist
      // only as TMEM accumulators.  act 8 (forward): every warp folds its 64 columns of the tile into one online
      // soft-max partial per row — running maximum, sum of exponentials (both in the log2 domain), arg-max — and the
      // warp that owns a row's label column also writes that logit.  act 9 (backward): the tile is recomputed and
      // leaves as bf16 d(logits) = (softmax - onehot) * g through the TMA-store box.
      const int act = p.act;
      const int swz = (lane >> 1) & 3;
      const int cnt = p.row_count ? __ldg(p.row_count) : p.M;
      constexpr float LOG2E = 1.4426950408889634f;
      for (int unit = unit0; unit < total_units; unit += unit_step) {
        const int tile = unit % tiles_mn;
        const int m0 = (tile / p.tiles_n) * GEMM_BM;
        const int tn = tile % p.tiles_n;
        const int n0 = tn * BN;
        const int ncols = min(BN, p.N - n0);
        const long long row = static_cast<long long>(m0) + q * 32 + lane;
        const bool row_ok = row < p.M;
        const int label = row_ok ? __ldg(p.ce_labels + row) : -1;
        float lse2 = 0.f, rs = 0.f;
        if (act == 9 && row_ok) {
          lse2 = __ldg(p.ce_lse2 + row);
          rs = (label >= 0 && row < cnt) ? __ldg(p.ce_g) : 0.f;
        }
        float m_run = -3.0e38f, s_run = 0.f, best = -3.0e38f;
        int best_i = 0;
        mbar_wait(&tfull_bar[acc], (acc_phase >> acc) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int i = 0; i < CPW; ++i) {
          const int c0 = (cgrp * CPW + i) * 32;
          if (c0 >= ncols) break;
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c0;
          uint8_t* box_row = box + lane * 64;
          if (act == 9) {
            if (lane == 0) tma_store_wait_read();  // the previous store has finished reading the box
            __syncwarp();
          }
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t r[16];
            tmem_ld16(taddr + hf * 16, r);
            float4 bv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + hf * 16 + j * 4));
            tmem_ld_wait();
            float x[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              x[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + bv[j].x;
              x[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + bv[j].y;
              x[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + bv[j].z;
              x[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + bv[j].w;
            }
            const int col0 = n0 + c0 + hf * 16;
            const int lc = label - col0;  // in [0, 16) when this half holds the label column
            if (act == 8) {
              if (static_cast<unsigned>(lc) < 16u) {
                float xl = 0.f;
#pragma unroll
                for (int e = 0; e < 16; ++e) xl = (e == lc) ? x[e] : xl;
                p.ce_xl[row] = xl;
              }
              float cm = -3.0e38f;
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                x[e] *= LOG2E;
                cm = fmaxf(cm, x[e]);
              }
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                if (x[e] > best) { best = x[e]; best_i = col0 + e; }  // strict: the first maximum wins, as torch.argmax
              }
              const float m_new = fmaxf(m_run, cm);
              float acc_s = s_run * exp2f(m_run - m_new);
#pragma unroll
              for (int e = 0; e < 16; ++e) acc_s += exp2f(x[e] - m_new);
              s_run = acc_s;
              m_run = m_new;
            } else {
              float d[16];
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const float pr = exp2f(fmaf(x[e], LOG2E, -lse2));
                d[e] = (pr - ((e == lc) ? 1.0f : 0.0f)) * rs;
              }
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                uint4 o;
                o.x = pack_bf16(d[8 * j + 0], d[8 * j + 1]); o.y = pack_bf16(d[8 * j + 2], d[8 * j + 3]);
                o.z = pack_bf16(d[8 * j + 4], d[8 * j + 5]); o.w = pack_bf16(d[8 * j + 6], d[8 * j + 7]);
                *reinterpret_cast<uint4*>(box_row + (((hf * 2 + j) ^ swz) << 4)) = o;
              }
            }
          }
          if (act == 9) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC, box, n0 + c0, m0 + q * 32);
              tma_store_commit();
            }
          }
        }
        if (act == 8)
          reinterpret_cast<float4*>(p.ce_part)[static_cast<long long>(tn * 4 + cgrp) * p.ce_mpad + row] =
              make_float4(m_run, s_run, best, __int_as_float(best_i));
        tc_fence_before();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        acc_phase ^= (1u << acc);
        acc ^= 1;
      }
      if (lane == 0) tma_store_wait_read();  // smem must outlive the last bulk store
    } else {
    const float scale = p.scale ? __ldg(p.scale) : 1.0f;
    const float* __restrict__ bias = p.bias;
    const bf16* __restrict__ residual = p.residual;
    const bf16* __restrict__ aux = p.aux;
    bf16* __restrict__ preact = p.preact;
    const float* __restrict__ row_scale = p.row_scale;
    const int act = p.act, out_mode = p.out_mode, use_tma = p.tma_store;
    for (int unit = unit0; unit < total_units; unit += unit_step) {
      const int tile = unit % tiles_mn;
      if ((unit / tiles_mn) * p.kb_per_split >= kb_total) continue;  // empty k-range (row_count)
      const int m0 = ((tile / p.tiles_n) * NCTA + static_cast<int>(cta_rank)) * GEMM_BM;
      const int n0 = (tile % p.tiles_n) * BN;
      const int ncols = min(BN, p.N - n0);
      const long long row = static_cast<long long>(m0) + q * 32 + lane;  // this thread's output row
      const bool row_ok = row < p.M;
      float sc = scale;  // gate alpha x DropPath scale of this row's sample
      if (row_scale && row_ok) sc *= __ldg(row_scale + static_cast<int>(row / p.rows_per_scale));
      const bf16* res_r = residual ? residual + row * p.ldr + n0 : nullptr;
      const bf16* aux_r = aux ? aux + row * p.ldaux + n0 : nullptr;

      // operands the epilogue reads from HBM (64 B per chunk and thread): start them towards L2 while the
      // MMAs of this tile are still running
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < CPW; ++i) {
          const int c0 = (cgrp * CPW + i) * 32;
          if (c0 < ncols) {
            if (res_r) prefetch_l2(res_r + c0);
            if (aux_r) prefetch_l2(aux_r + c0);
          }
        }
      }
      // warp-uniform fast-path key of this tile: every row in range, bf16 output through the TMA-store box
      const bool tile_fast = use_tma && (m0 + GEMM_BM <= p.M);
      const bool scaled = p.scale != nullptr || row_scale != nullptr;

      mbar_wait(&tfull_bar[acc], (acc_phase >> acc) & 1);
      tc_fence_after();
      const int passes = preact ? 2 : 1;
      for (int pass = 0; pass < passes; ++pass) {
        const bool write_pre = preact && pass == 0;
#pragma unroll 1
        for (int i = 0; i < CPW; ++i) {
          const int c0 = (cgrp * CPW + i) * 32;
          if (c0 >= ncols) break;
          if (use_tma) {
            if (lane == 0) tma_store_wait_read();  // the previous store has finished reading the box
            __syncwarp();
          }
          if (tile_fast && c0 + 32 <= ncols) {
            // compile-time specialised chunk bodies for the epilogues the FIBER path uses; anything else
            // (and every ragged edge) takes the generic runtime-flag body below
            const int eff_act = write_pre ? 0 : act;
            const bool eff_res = res_r != nullptr && !write_pre;
            const bool eff_scale = scaled && !write_pre;
            const int key = eff_act | (bias ? 4 : 0) | (eff_scale ? 8 : 0) | (eff_res ? 16 : 0);
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c0;
            uint8_t* box_row = box + lane * 64;
            const int swz = (lane >> 1) & 3;
            const float* bias_c = bias + n0 + c0;
            const bf16* res_c = res_r + c0;
            const bf16* aux_c = aux_r + c0;
            bool done = true;
            switch (key) {
              case 0:           epi_chunk_full<0, false, false, false>(taddr, box_row, swz, bias_c, res_c, aux_c, sc); break;
              case 4:           epi_chunk_full<0, true, false, false>(taddr, box_row, swz, bias_c, res_c, aux_c, sc); break;
              case 1 | 4:       epi_chunk_full<1, true, false, false>(taddr, box_row, swz, bias_c, res_c, aux_c, sc); break;
              case 2:           epi_chunk_full<2, false, false, false>(taddr, box_row, swz, bias_c, res_c, aux_c, sc); break;
              case 16:          epi_chunk_full<0, false, false, true>(taddr, box_row, swz, bias_c, res_c, aux_c, sc); break;
              case 4 | 16:      epi_chunk_full<0, true, false, true>(taddr, box_row, swz, bias_c, res_c, aux_c, sc); break;
              case 4 | 8 | 16:  epi_chunk_full<0, true, true, true>(taddr, box_row, swz, bias_c, res_c, aux_c, sc); break;
              case 8:           epi_chunk_full<0, false, true, false>(taddr, box_row, swz, bias_c, res_c, aux_c, sc); break;
              default: done = false; break;
            }
            if (done) {
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(write_pre ? &tmP : &tmC, box, n0 + c0, m0 + q * 32);
                tma_store_commit();
              }
              continue;
            }
          }
          // two 16-column halves per 32-column box: short live ranges leave registers for interleaving the
          // GELU dependency chains of four packed pairs
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int ch = c0 + hf * 16;
            uint4 rv[2], av[2];
            if (res_r && !write_pre) {
#pragma unroll
              for (int j = 0; j < 2; ++j)
                rv[j] = (row_ok && ch + j * 8 < ncols) ? *reinterpret_cast<const uint4*>(res_r + ch + j * 8)
                                                       : make_uint4(0u, 0u, 0u, 0u);
            }
            if (act == 2 && !write_pre) {
#pragma unroll
              for (int j = 0; j < 2; ++j)
                av[j] = (row_ok && ch + j * 8 < ncols) ? *reinterpret_cast<const uint4*>(aux_r + ch + j * 8)
                                                       : make_uint4(0u, 0u, 0u, 0u);
            }
            uint32_t r[16];
            tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + ch, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 2; ++j) {  // 8 columns = one 16-byte bf16 chunk; math on packed fp32 pairs
              uint64_t xp[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                xp[e] = pk2(__uint_as_float(r[j * 8 + 2 * e]), __uint_as_float(r[j * 8 + 2 * e + 1]));
              const int col = ch + j * 8;  // column inside the tile
              const bool col_ok = col < ncols;
              if (bias && col_ok) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + n0 + col));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + n0 + col + 4));
                xp[0] = add2(xp[0], pk2(b0.x, b0.y)); xp[1] = add2(xp[1], pk2(b0.z, b0.w));
                xp[2] = add2(xp[2], pk2(b1.x, b1.y)); xp[3] = add2(xp[3], pk2(b1.z, b1.w));
              }
              if (!write_pre) {
                if (act == 1) {
                  gelu_erf2x4(xp);
                } else if (act == 2) {
                  const uint32_t* au = reinterpret_cast<const uint32_t*>(&av[j]);
                  uint64_t hx[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f = unpack_bf16(au[e]);
                    hx[e] = pk2(f.x, f.y);
                  }
                  gelu_erf_grad_mul2x4(xp, hx);
                }
                const uint64_t sc2 = pk2(sc, sc);
                if (res_r) {
                  const uint32_t* ru = reinterpret_cast<const uint32_t*>(&rv[j]);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f = unpack_bf16(ru[e]);
                    xp[e] = fma2(xp[e], sc2, pk2(f.x, f.y));
                  }
                } else {
#pragma unroll
                  for (int e = 0; e < 4; ++e) xp[e] = mul2(xp[e], sc2);
                }
              }
              float x[8];
#pragma unroll
              for (int e = 0; e < 4; ++e) upk2(xp[e], x[2 * e], x[2 * e + 1]);
              const int jb = hf * 2 + j;  // 16-byte chunk inside the 64-byte box row
              if (out_mode == 0 || write_pre) {
                uint4 o;
                o.x = pack_bf16(x[0], x[1]); o.y = pack_bf16(x[2], x[3]);
                o.z = pack_bf16(x[4], x[5]); o.w = pack_bf16(x[6], x[7]);
                if (use_tma) {
                  *reinterpret_cast<uint4*>(box + lane * 64 + ((jb ^ ((lane >> 1) & 3)) << 4)) = o;
                } else if (row_ok && col_ok) {  // pitch not TMA-addressable: direct 16-byte row stores
                  bf16* dst = write_pre ? preact + row * p.ldp + n0 + col
                                        : reinterpret_cast<bf16*>(p.c) + row * p.ldc + n0 + col;
                  *reinterpret_cast<uint4*>(dst) = o;
                }
              } else if (row_ok && col_ok) {
                float* dst = reinterpret_cast<float*>(p.c) + row * p.ldc + n0 + col;
                if (out_mode == 1) {
                  *reinterpret_cast<float4*>(dst) = make_float4(x[0], x[1], x[2], x[3]);
                  *reinterpret_cast<float4*>(dst + 4) = make_float4(x[4], x[5], x[6], x[7]);
                } else {  // split-K wgrad: 16-byte vector reductions (red.global.add.v4.f32)
                  atomicAdd(reinterpret_cast<float4*>(dst), make_float4(x[0], x[1], x[2], x[3]));
                  atomicAdd(reinterpret_cast<float4*>(dst + 4), make_float4(x[4], x[5], x[6], x[7]));
                }
              }
            }
          }
          if (use_tma) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(write_pre ? &tmP : &tmC, box, n0 + c0, m0 + q * 32);
              tma_store_commit();
            }
          }
        }
      }
      if (do_cs && n0 == 0 && cgrp == 0) {  // the 16 column-sum columns are identical: take the first
        uint32_t cs[16];
        tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + BN, cs);
        tmem_ld_wait();
        if (row_ok) atomicAdd(p.colsum + row, __uint_as_float(cs[0]) * scale);
      }
      // every TMEM read of this accumulator by this warp has completed (tmem_ld_wait): hand it back
      tc_fence_before();
      if (lane == 0) {
        if constexpr (CTA2) mbar_arrive_cluster(mapa_cta(smem_u32(&tempty_bar[acc]), 0));
        else                mbar_arrive(&tempty_bar[acc]);
      }
      acc_phase ^= (1u << acc);
      acc = nacc == 2 ? acc ^ 1 : 0;
    }
    if (use_tma && lane == 0) tma_store_wait_read();  // smem must outlive the last bulk store
    }  // EPI == 0
  }

  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all();  // nothing of the peer (barriers, operand halves) is referenced past this point
  else                __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    if constexpr (CTA2) tmem_dealloc2(tmem_base, Cfg::TMEM_COLS);
    else                tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

// 2-D bf16 tensor map: dim0 = contiguous dimension (inner), dim1 = rows; 128B swizzle.
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer,
                      uint64_t row_pitch_elems, uint32_t box_inner, uint32_t box_outer,
                      CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = get_encode_fn();
  FIBER_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  FIBER_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16B aligned");
  FIBER_CHECK((row_pitch_elems * 2) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes (ld=%llu)",
              (unsigned long long)row_pitch_elems);
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {row_pitch_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FIBER_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return 0;
}

template <int BN, int MN, int EPI = 0, int CTA2 = 0>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tp,
                       const GemmParams& p, int grid, cudaStream_t stream) {
  auto kern = gemm_tcgen05_kernel<BN, MN, EPI, CTA2>;
  using Cfg = GemmCfg<BN, EPI, CTA2>;
  static bool attr_set = false;
  if (!attr_set) {
    FIBER_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  if (CTA2) FIBER_CUDA(launch_k_pair(kern, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, ta, tb, tc, tp, p));
  else      FIBER_CUDA(launch_k(kern, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, ta, tb, tc, tp, p));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int gemm_dispatch(const fiber_gemm_args* a, cudaStream_t stream) {
  FIBER_CHECK(a != nullptr, "null args");
  FIBER_CHECK(a->m > 0 && a->n > 0 && a->k > 0, "bad GEMM shape %d x %d x %d", a->m, a->n, a->k);
  FIBER_CHECK(a->a_major == a->b_major, "mixed operand majors are not supported");
  FIBER_CHECK(a->out_mode >= 0 && a->out_mode <= 2, "bad out_mode");
  FIBER_CHECK(a->act >= 0 && a->act <= 7, "bad act %d", a->act);
  FIBER_CHECK((a->act != 2 && a->act != 4 && a->act != 7) || a->aux != nullptr, "act=2 / 4 / 7 need aux");
  const bool epi1 = a->act >= 3;  // opt-in single-pass GELU + GELU' (3) / GELU + pre-activation (5) / multiply-by-aux (4)
  if (epi1) {
    FIBER_CHECK(a->a_major == 0 && a->out_mode == 0 && a->m % GEMM_BM == 0 && a->n % 32 == 0,
                "act=%d needs K-major operands, a bf16 output, M %% 128 == 0 and N %% 32 == 0", a->act);
    if (a->act == 6) {
      FIBER_CHECK(a->residual != nullptr && a->preact == nullptr && a->aux == nullptr && a->colsum == nullptr,
                  "act=6 needs residual and takes no preact / aux / colsum");
    } else {
      FIBER_CHECK(a->residual == nullptr && a->scale == nullptr && a->row_scale == nullptr && a->colsum == nullptr,
                  "act=%d does not combine with residual / scale / row_scale / colsum", a->act);
      FIBER_CHECK((a->act == 4 || a->act == 7) ? (a->preact == nullptr && a->bias == nullptr) : a->preact != nullptr,
                  "act=3 / act=5 write their second output to preact; act=4 / act=7 take no bias / preact");
    }
  }
  const int mn = a->a_major;
  const int BN = (a->n > 128) ? 256 : 128;

  GemmParams p{};
  p.M = a->m; p.N = a->n; p.K = a->k;
  p.row_count = a->row_count;
  p.tiles_m = (a->m + GEMM_BM - 1) / GEMM_BM;
  p.tiles_n = (a->n + BN - 1) / BN;
  p.kb_total = (a->k + GEMM_BK - 1) / GEMM_BK;
  const int sms = num_sms();
  int splits = a->splits;
  const int tiles = p.tiles_m * p.tiles_n;
  if (a->out_mode != 2) {
    splits = 1;
  } else if (splits <= 0) {
    // Work units = tiles x splits are dealt round-robin to one CTA per SM and all have the same depth, so the
    // launch takes ceil(units / SMs) waves of ceil(kb / splits) k-blocks (+ ~6 k-blocks of epilogue / atomics
    // per wave): pick the split count with the shortest makespan, at least 4 k-blocks per unit.
    const int max_splits = (p.kb_total + 3) / 4;
    long long best_cost = -1;
    splits = 1;
    for (int s = 1; s <= max_splits && s <= 8 * sms; ++s) {
      const long long waves = (static_cast<long long>(tiles) * s + sms - 1) / sms;
      const long long depth = (p.kb_total + s - 1) / s;
      const long long cost = waves * (depth + 6);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; splits = s; }
    }
  }
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.c = a->c; p.ldc = a->ldc;
  p.bias = a->bias;
  p.residual = reinterpret_cast<const bf16*>(a->residual); p.ldr = a->ldr;
  p.aux = reinterpret_cast<const bf16*>(a->aux); p.ldaux = a->ldaux;
  p.preact = reinterpret_cast<bf16*>(a->preact); p.ldp = a->ldp;
  p.scale = a->scale; p.row_scale = a->row_scale;
  p.rows_per_scale = a->rows_per_scale > 0 ? a->rows_per_scale : 1;
  p.act = a->act; p.out_mode = a->out_mode;
  p.colsum = a->colsum;
  FIBER_CHECK(a->colsum == nullptr || (a->a_major == 1 && a->out_mode != 0 && a->row_scale == nullptr),
              "colsum needs MN-major operands, an fp32 output and no row_scale");

  // CTA pairs (GemmCfg): 256-column tiles, whole 256-row pair tiles, no device row count; K-major launches (bit 0 / bit 1 by
  // epilogue family) without split-K, wgrad launches (bit 3) with it
  const int cta2_opt = option_gemm_cta2();
  const bool cta2 = BN == 256 && a->m % (2 * GEMM_BM) == 0 && a->row_count == nullptr &&
                    a->k >= ((cta2_opt & 4) ? 256 : 1024) &&
                    (mn == 0 ? (splits == 1 && ((cta2_opt >> (epi1 ? 1 : 0)) & 1) != 0) : (cta2_opt & 8) != 0);
  CUtensorMap ta, tb;
  if (mn == 0) {
    // A[M, K] (lda), B[N, K] (ldb): inner = K
    if (make_tmap_bf16_2d(&ta, a->a, a->k, a->m, a->lda, GEMM_BK, GEMM_BM)) return -1;
    if (make_tmap_bf16_2d(&tb, a->b, a->k, a->n, a->ldb, GEMM_BK, cta2 ? BN / 2 : BN)) return -1;
  } else {
    // A stored [K, M] (lda), B stored [K, N] (ldb): inner = M / N, 64-wide chunks
    if (make_tmap_bf16_2d(&ta, a->a, a->m, a->k, a->lda, 64, GEMM_BK)) return -1;
    if (make_tmap_bf16_2d(&tb, a->b, a->n, a->k, a->ldb, 64, GEMM_BK)) return -1;
  }
  // TMA-store epilogue (thread = accumulator row) whenever the output is bf16 with TMA-compatible pitch
  // and the fused operands are 16-byte addressable; fp32 / atomic outputs use the transposing epilogue
  CUtensorMap tc = ta, tp = ta;
  p.tma_store = 0;
  static const int tma_mode = [] {  // FIBER_TMA_STORE=0: direct row stores instead of TMA stores (debug)
    const char* e = getenv("FIBER_TMA_STORE");
    return e ? atoi(e) : 2;
  }();
  FIBER_CHECK(a->n % 8 == 0, "N must be a multiple of 8 (got %d)", a->n);
  FIBER_CHECK(a->residual == nullptr || (a->ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(a->residual) & 15) == 0),
              "residual rows must be 16-byte aligned");
  FIBER_CHECK((a->act != 2 && a->act != 4 && a->act != 7) ||
                  (a->ldaux % 8 == 0 && (reinterpret_cast<uintptr_t>(a->aux) & 15) == 0),
              "aux rows must be 16-byte aligned");
  FIBER_CHECK((a->ldc * (a->out_mode == 0 ? 2 : 4)) % 16 == 0 && (reinterpret_cast<uintptr_t>(a->c) & 15) == 0,
              "output rows must be 16-byte aligned");
  FIBER_CHECK(a->preact == nullptr || ((a->ldp * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(a->preact) & 15) == 0),
              "preact rows must be 16-byte aligned");
  if (tma_mode > 0 && a->out_mode == 0) {
    if (make_tmap_bf16_2d(&tc, a->c, a->n, a->m, a->ldc, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B)) return -1;
    if (a->preact && make_tmap_bf16_2d(&tp, a->preact, a->n, a->m, a->ldp, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B)) return -1;
    p.tma_store = 1;
  }
  const int units = tiles * p.splits;
  const int grid = units < sms ? units : sms;
  if (cta2) {
    count_gemm_cta2_launch();
    const int pair_grid = units < (sms & ~1) ? units : (sms & ~1);  // units is even (M % 256 == 0): one CTA per 128-row half
    if (epi1) {
      FIBER_CHECK(p.tma_store == 1, "act=%d needs the TMA-store epilogue", a->act);
      return launch_gemm<256, 0, 1, 1>(ta, tb, tc, tp, p, pair_grid, stream);
    }
    return mn ? launch_gemm<256, 1, 0, 1>(ta, tb, tc, tp, p, pair_grid, stream)
              : launch_gemm<256, 0, 0, 1>(ta, tb, tc, tp, p, pair_grid, stream);
  }
  if (epi1) {
    FIBER_CHECK(p.tma_store == 1, "act=%d needs the TMA-store epilogue", a->act);
    return BN == 256 ? launch_gemm<256, 0, 1>(ta, tb, tc, tp, p, grid, stream)
                     : launch_gemm<128, 0, 1>(ta, tb, tc, tp, p, grid, stream);
  }
  if (BN == 256) {
    return mn ? launch_gemm<256, 1>(ta, tb, tc, tp, p, grid, stream) : launch_gemm<256, 0>(ta, tb, tc, tp, p, grid, stream);
  }
  return mn ? launch_gemm<128, 1>(ta, tb, tc, tp, p, grid, stream) : launch_gemm<128, 0>(ta, tb, tc, tp, p, grid, stream);
}

// ------------------------------------------------------------------------------------------
// Fused MLM decoder + cross-entropy (fiber_mlm_ce_fwd / fiber_mlm_ce_bwd)
// ------------------------------------------------------------------------------------------
// One warp per row folds the row's per-warp-per-tile partials of the act-8 epilogue into the log-sum-exp, the arg-max
// (lowest column on ties) and the row's loss lse - logit[label]; rows that are ignored (label < 0) or not computed
// (row >= *row_count) get loss 0, prediction 0.
__global__ void __launch_bounds__(256) ce_combine_kernel(const float4* __restrict__ part, int nparts, int mpad, int M,
                                                         const int* __restrict__ labels, const int* __restrict__ row_count,
                                                         const float* __restrict__ xl, float* __restrict__ lse2_out,
                                                         float* __restrict__ loss_rows, int* __restrict__ pred) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int row = static_cast<int>((static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5);
  if (row >= M) return;
  const int cnt = row_count ? __ldg(row_count) : M;
  if (row >= cnt) {
    if (lane == 0) { lse2_out[row] = 0.f; loss_rows[row] = 0.f; pred[row] = 0; }
    return;
  }
  float m = -3.0e38f, s = 0.f, best = -3.0e38f;
  int bi = 0x7fffffff;
  for (int t = lane; t < nparts; t += 32) {
    const float4 v = part[static_cast<long long>(t) * mpad + row];
    const float mn = fmaxf(m, v.x);
    s = s * exp2f(m - mn) + v.y * exp2f(v.x - mn);
    m = mn;
    const int vi = __float_as_int(v.w);
    if (v.y > 0.f && (v.z > best || (v.z == best && vi < bi))) { best = v.z; bi = vi; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float b2 = __shfl_xor_sync(0xffffffffu, best, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
    const float mn = fmaxf(m, m2);
    s = s * exp2f(m - mn) + s2 * exp2f(m2 - mn);
    m = mn;
    if (b2 > best || (b2 == best && i2 < bi)) { best = b2; bi = i2; }
  }
  if (lane == 0) {
    const float lse2 = m + log2f(s);
    const int label = __ldg(labels + row);
    lse2_out[row] = lse2;
    loss_rows[row] = label >= 0 ? (lse2 - __ldg(xl + row) * 1.4426950408889634f) * 0.6931471805599453f : 0.f;
    pred[row] = bi == 0x7fffffff ? 0 : bi;
  }
}

int mlm_ce_dispatch(const fiber_ce_args* a, int backward, cudaStream_t stream) {
  FIBER_CHECK(a != nullptr, "null args");
  FIBER_CHECK(a->m > 0 && a->n > 256 && a->k > 0 && a->n % 32 == 0, "mlm_ce: bad shape %d x %d x %d (N %% 32 == 0, N > 256)",
              a->m, a->n, a->k);
  FIBER_CHECK(a->x && a->w && a->bias && a->labels && a->lse, "mlm_ce: x, w, bias, labels and lse are required");
  constexpr int BN = 256;
  GemmParams p{};
  p.M = a->m; p.N = a->n; p.K = a->k;
  p.tiles_m = (a->m + GEMM_BM - 1) / GEMM_BM;
  p.tiles_n = (a->n + BN - 1) / BN;
  p.kb_total = (a->k + GEMM_BK - 1) / GEMM_BK;
  p.splits = 1;
  p.kb_per_split = p.kb_total;
  p.bias = a->bias;
  p.row_count = a->row_count;
  p.ce_labels = a->labels;
  p.ce_mpad = p.tiles_m * GEMM_BM;
  CUtensorMap ta, tb;
  if (make_tmap_bf16_2d(&ta, a->x, a->k, a->m, a->ldx, GEMM_BK, GEMM_BM)) return -1;
  if (make_tmap_bf16_2d(&tb, a->w, a->k, a->n, a->ldw, GEMM_BK, BN)) return -1;
  CUtensorMap tc = ta;
  const int units = p.tiles_m * p.tiles_n;
  const int sms = num_sms();
  const int grid = units < sms ? units : sms;
  if (!backward) {
    FIBER_CHECK(a->part && a->label_logit && a->loss_rows && a->pred, "mlm_ce_fwd: part, label_logit, loss_rows, pred required");
    p.act = 8;
    p.ce_part = a->part;
    p.ce_xl = a->label_logit;
    if (launch_gemm<256, 0, 2>(ta, tb, tc, tc, p, grid, stream)) return -2;
    const int blocks = (a->m + 7) / 8;  // eight rows (warps) per block
    FIBER_CUDA(launch_k(ce_combine_kernel, dim3(blocks), dim3(256), 0, stream, reinterpret_cast<const float4*>(a->part),
                        p.tiles_n * 4, p.ce_mpad, a->m, a->labels, a->row_count, a->label_logit, a->lse, a->loss_rows, a->pred));
    FIBER_CUDA(cudaGetLastError());
    count_launch();
    return 0;
  }
  FIBER_CHECK(a->dlogits && a->gscale, "mlm_ce_bwd: dlogits and gscale required");
  FIBER_CHECK((a->lddl * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(a->dlogits) & 15) == 0,
              "mlm_ce_bwd: d(logits) rows must be 16-byte aligned");
  p.act = 9;
  p.ce_lse2 = a->lse;
  p.ce_g = a->gscale;
  if (make_tmap_bf16_2d(&tc, a->dlogits, a->n, a->m, a->lddl, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B)) return -1;
  return launch_gemm<256, 0, 2>(ta, tb, tc, tc, p, grid, stream);
}

}  // namespace fiber
