// fiber_b200 — Swin window attention on tcgen05 (second generation of window_attn.cu), specialised for the
// 384-px FIBER configuration: 12x12 = 144-token windows, head_dim 32 (swin_transformer.py:195-224, :363-387).
//
// STATUS: opt-in (fiber_set_option("winattn_tc", 1 | 2) or FIBER_WINATTN_TC=3); the mma.sync kernels of
// window_attn.cu stay the default until this file has been validated on a B200 (tests/test_attention_gpu.py
// runs every window case against both generations when the option is set).
//
// Why: the mma.sync generation is bound by shared-memory wavefronts (ldmatrix 49 % + staging of P / dS), not
// by the tensor pipe (ncu: profiles/r1_win{fwd,bwd}_ncu.txt).  Here the tensor core reads every operand
// straight from swizzled shared memory (tcgen05.mma, SS form) and S / dP / the output accumulators live in
// TMEM, so the LSU only sees the bias-table gather and ONE 16-byte store per 8 probabilities.
//
// Geometry.  144 = 128 + 16: query rows 0..127 of a window are one M = 128 UMMA tile; the 16 remaining rows
// (and, in the backward, the 16 remaining keys of dV / dK) are 1/9 of the work and run on two "remainder"
// warps with mma.sync on the same shared-memory tiles (ldmatrix through the swizzle).
//
// Shared-memory layouts (all UMMA-canonical, see cute/atom/mma_traits_sm100.hpp "make_umma_desc"):
//   Q / K / V / dO tile  [144 rows][32]  64-byte rows, SWIZZLE_64B (16-byte piece ^= (row >> 1) & 3).
//       K-major operand (A or B of S = Q K^T, dP = dO V^T): 8-row groups 512 B apart (SBO), k-step = +32 B.
//       MN-major B operand (V in P V, dO in P^T dO, Q in dS^T Q, K in dS K): the 32 head-dim elements are the
//       contiguous MN extent, 8 token rows = one 512-byte atom (SBO), k-step of 16 tokens = +1024 B.
//   P / dS  [rows = queries][keys] in three 64-key chunks of 128-byte rows, SWIZZLE_128B (piece ^= row & 7).
//       K-major A (P V, dS K): 8-row groups 1024 B apart, k-step = +32 B inside a chunk.
//       MN-major A (P^T dO, dS^T Q; M = keys): LBO = chunk stride, SBO = 1024 (8 query rows), k-step = +2048 B
//       — the same descriptor form the wgrad GEMM uses (gemm_sm100.cu).
//
// Roles (384 threads, one CTA per SM, persistent over the windows of one head):
//   warps 0-7   element-wise: thread = (query row, half of the 144 key columns); TMEM -> registers ->
//               bias / mask / exp2 -> bf16 -> swizzled smem; accumulator drain + global stores
//   warp  8     tcgen05.mma issuer (one lane) and TMEM owner
//   warp  9     cp.async loader of the gathered Q / K / V (/ dO) tiles (cyclic shift + partition in the address)
//   warps 10-11 remainder rows / keys on mma.sync
#include "window_common.cuh"
#include "window_tc_layout.cuh"
#include "../../include/fiber_b200.h"

#include <mutex>

namespace fiber {

void count_launch(int n = 1);
void count_winattn_tc_launch();  // capi.cu
int launch_win_bwd_prep(const AttnParams& p, float* D, cudaStream_t stream);  // window_attn.cu

namespace {

using namespace tcl;  // layouts and descriptor builders (window_tc_layout.cuh)

constexpr int TC_N = tcl::N;              // tokens per window
constexpr int TC_WS = tcl::WS;
constexpr int TC_TW2 = 2 * TC_WS - 1;     // 23
constexpr int TC_TILE = tcl::TILE;        // bytes of one [144][32] bf16 tile
constexpr int TC_THREADS = 384;
constexpr int TC_WARP_MMA = 8, TC_WARP_LD = 9, TC_WARP_R0 = 10;
constexpr int TC_TABLE_BYTES = (2 * WA_MAXTBL + 2) * 4 + 4 * WA_ROWS * 4;  // WinTables: tbl2, aq4, bj4, code, tok

__device__ __forceinline__ void tmem_ld32p(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16p(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8p(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// mbarrier wait that names itself when it times out (a protocol bug becomes a trap with a readable message):
// tag = 100 * kernel (1 forward, 2 backward) + barrier number (listed next to the barrier declarations)
__device__ __noinline__ void tc_wait_timeout(int tag, uint32_t parity, int it) {
  printf("fiber_b200 window_attn_tc: mbarrier timeout tag %d parity %u window-iteration %d block (%d,%d) warp %d lane %d\n",
         tag, parity, it, blockIdx.x, blockIdx.y, threadIdx.x >> 5, threadIdx.x & 31);
  __trap();
}
__device__ __forceinline__ void tc_wait(uint64_t* bar, uint32_t parity, int tag, int it) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
#ifndef FIBER_NO_BACKOFF
    if (++spins > 4) __nanosleep(spins > 64 ? 256 : 32);  // do not take issue slots from the working warps (common.cuh)
#else
    ++spins;
#endif
    if (spins > (1u << 22)) tc_wait_timeout(tag, parity, it);
  }
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// relative-position index pieces of key column j (compile-time after unrolling): B_j and the SW-MSA region code
__host__ __device__ constexpr int tc_bj(int j) { return (j / TC_WS) * TC_TW2 + j % TC_WS; }
__host__ __device__ constexpr int tc_code(int j) { return ((j / TC_WS) >= TC_WS / 2 ? 1 : 0) | ((j % TC_WS) >= TC_WS / 2 ? 2 : 0); }

// score2[c] = s[c] * scale2 + table2[A_i - B_j] (+ mask) for the 72 key columns j = HF * 72 + c of one query
// row; returns the row maximum over these columns.  tbl_i = table bytes + A_i.
template <int HF, bool MASKED>
__device__ __forceinline__ float tc_scores(uint32_t (&v)[72], const char* tbl_i, float scale2, const float (&madd)[4]) {
  float mx = -1e30f;
#pragma unroll
  for (int c = 0; c < 72; ++c) {
    const int j = HF * 72 + c;
    const float t = *reinterpret_cast<const float*>(tbl_i - 4 * tc_bj(j));
    float x = fmaf(__uint_as_float(v[c]), scale2, t);
    if (MASKED) x += madd[tc_code(j)];
    v[c] = __float_as_uint(x);
    mx = fmaxf(mx, x);
  }
  return mx;
}

// =================================================================================================
// Forward
// =================================================================================================
constexpr int TF_STAGES = 3;
constexpr int TF_STAGE_BYTES = 3 * TC_TILE;                 // Q, K, V
constexpr int TF_PCHUNK = tcl::F_PCHUNK;                    // P chunk: 128 query rows x 64 keys
constexpr int TF_OFF_P = TF_STAGES * TF_STAGE_BYTES;        // 82944
constexpr int TF_OFF_TBL = TF_OFF_P + 3 * TF_PCHUNK;        // 132096
constexpr int TF_OFF_RMAX = TF_OFF_TBL + ((TC_TABLE_BYTES + 15) & ~15);
constexpr int TF_OFF_RSUM = TF_OFF_RMAX + 2 * 2 * 128 * 4;
constexpr int TF_OFF_BARS = TF_OFF_RSUM + 2 * 2 * 128 * 4;
constexpr int TF_SMEM = 1024 + TF_OFF_BARS + 16 * 8;
constexpr uint32_t TF_S_COL0 = 0, TF_S_COL1 = 160, TF_O_COL = 320;

// remainder query tile (rows 128..143) of one window on mma.sync: S, online softmax over three 48-key
// sub-tiles and P V as in window_attn.cu's wf_tile, reading the SWIZZLE_64B tiles
template <bool MASKED>
__device__ __forceinline__ void tc_fwd_rem_tile(uint32_t sQ, uint32_t sK, uint32_t sV, const WinTables& T,
                                                const char* tbl_bytes, int lane, float scale2, int emask,
                                                float (&oacc)[4][4], float (&m_run)[2], float (&l_run)[2]) {
  uint32_t qf[2][4];
  {
    const int row = 128 + (lane & 7) + ((lane >> 3) & 1) * 8;
    const int pc = lane >> 4;
    ldsm_x4(sQ + sw64_off(row, pc), qf[0][0], qf[0][1], qf[0][2], qf[0][3]);
    ldsm_x4(sQ + sw64_off(row, pc + 2), qf[1][0], qf[1][1], qf[1][2], qf[1][3]);
  }
  const int rl0 = 128 + (lane >> 2);
  const int c2 = (lane & 3) * 2;
  const int aq0 = T.aq4[rl0], aq1 = T.aq4[rl0 + 8];
  int ci0 = 0, ci1 = 0;
  if (MASKED) {
    ci0 = T.code[rl0] & emask;
    ci1 = T.code[rl0 + 8] & emask;
  }
#pragma unroll 1
  for (int sub = 0; sub < 3; ++sub) {
    float s[6][4];
#pragma unroll
    for (int i = 0; i < 6; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
      for (int nt2 = 0; nt2 < 3; ++nt2) {
        const int row = sub * 48 + nt2 * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int pc = ks * 2 + ((lane >> 3) & 1);
        uint32_t b0, b1, b2, b3;
        ldsm_x4(sK + sw64_off(row, pc), b0, b1, b2, b3);
        mma16816(s[2 * nt2], qf[ks], b0, b1);
        mma16816(s[2 * nt2 + 1], qf[ks], b2, b3);
      }
    }
    float mx0 = -1e30f, mx1 = -1e30f;
#pragma unroll
    for (int nt = 0; nt < 6; ++nt) {
      const int j0 = sub * 48 + nt * 8 + c2;
      const int2 bj = *reinterpret_cast<const int2*>(T.bj4 + j0);
      int2 cj = make_int2(0, 0);
      if (MASKED) {
        cj = *reinterpret_cast<const int2*>(T.code + j0);
        cj.x &= emask;
        cj.y &= emask;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool hi = e >> 1, odd = e & 1;
        const float t = *reinterpret_cast<const float*>(tbl_bytes + ((hi ? aq1 : aq0) - (odd ? bj.y : bj.x)));
        float v = fmaf(s[nt][e], scale2, t);
        if (MASKED) {
          if ((hi ? ci1 : ci0) != (odd ? cj.y : cj.x)) v += WA_MASK2;
        }
        s[nt][e] = v;
        if (hi) mx1 = fmaxf(mx1, v); else mx0 = fmaxf(mx0, v);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m_run[0], mx0), mn1 = fmaxf(m_run[1], mx1);
    const float corr0 = ex2_approx(m_run[0] - mn0), corr1 = ex2_approx(m_run[1] - mn1);
    m_run[0] = mn0; m_run[1] = mn1;
    float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 6; ++nt) {
      s[nt][0] = ex2_approx(s[nt][0] - mn0);
      s[nt][1] = ex2_approx(s[nt][1] - mn0);
      s[nt][2] = ex2_approx(s[nt][2] - mn1);
      s[nt][3] = ex2_approx(s[nt][3] - mn1);
      ls0 += s[nt][0] + s[nt][1];
      ls1 += s[nt][2] + s[nt][3];
    }
    l_run[0] = fmaf(l_run[0], corr0, ls0);
    l_run[1] = fmaf(l_run[1], corr1, ls1);
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) {
      oacc[dt][0] *= corr0; oacc[dt][1] *= corr0;
      oacc[dt][2] *= corr1; oacc[dt][3] *= corr1;
    }
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
      uint32_t a[4];
      a[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      a[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      a[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      a[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dt2 = 0; dt2 < 2; ++dt2) {
        const int row = sub * 48 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int pc = dt2 * 2 + (lane >> 4);
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(sV + sw64_off(row, pc), b0, b1, b2, b3);
        mma16816(oacc[2 * dt2], a, b0, b1);
        mma16816(oacc[2 * dt2 + 1], a, b2, b3);
      }
    }
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1) win_attn_tc_fwd_kernel(const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ array, so that the compiler keeps the address space
  // (LDS / STS instead of generic LD / ST for the table gathers and the P / dS stores)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sP = smem + TF_OFF_P;
  WinTables T;
  T.tbl2 = reinterpret_cast<float*>(smem + TF_OFF_TBL);
  T.aq4 = reinterpret_cast<int*>(T.tbl2 + 2 * WA_MAXTBL + 2);
  T.bj4 = T.aq4 + WA_ROWS;
  T.code = T.bj4 + WA_ROWS;
  T.tok = T.code + WA_ROWS;
  const char* tbl_bytes = reinterpret_cast<const char*>(T.tbl2);
  float* rowmax = reinterpret_cast<float*>(smem + TF_OFF_RMAX);  // [window parity][column half][128]
  float* rowsum = reinterpret_cast<float*>(smem + TF_OFF_RSUM);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TF_OFF_BARS);
  // timeout tags (tc_wait): 101 s_full, 102 o_full, 103 full, 104 s_empty, 105 p_ready, 106 stage_free
  uint64_t* full = bars;            // [3] tiles of a stage have landed              (32 loader lanes)
  uint64_t* stage_free = bars + 3;  // [3] stage may be overwritten                   (tcgen05.commit + remainder warp)
  uint64_t* s_full = bars + 6;      // [2] S accumulator written                      (tcgen05.commit)
  uint64_t* s_empty = bars + 8;     // [2] S accumulator read into registers          (8 element-wise warps)
  uint64_t* p_ready = bars + 10;    //     P in smem, previous O drained              (8 element-wise warps)
  uint64_t* o_full = bars + 11;     //     O accumulator written, P consumed          (tcgen05.commit)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x;
  WinGeo geo;
  geo.H = p.H; geo.W = p.W; geo.ws = TC_WS; geo.shift = p.shift;
  geo.nWw = p.W / TC_WS; geo.nWh = p.H / TC_WS; geo.nW = geo.nWh * geo.nWw;
  const int n_groups = p.G * geo.nW;
  const int n_my = (n_groups - static_cast<int>(blockIdx.y) + static_cast<int>(gridDim.y) - 1) / static_cast<int>(gridDim.y);
  const float scale2 = p.scale * WA_LOG2E;

  pdl_trigger();
  pdl_wait();
  fill_tables(T, p.bias_table, p.nH, h, TC_WS, p.shift, TC_N, tid, TC_THREADS);
  if (warp == TC_WARP_MMA) {
    if (lane == 0) {
      for (int s = 0; s < TF_STAGES; ++s) {
        mbar_init(&full[s], 32);
        mbar_init(&stage_free[s], 2);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&s_full[b], 1);
        mbar_init(&s_empty[b], 8);
      }
      mbar_init(p_ready, 8);
      mbar_init(o_full, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < 8) {
    // ================= element-wise warps =================
    const int q = warp & 3, hf = warp >> 2, row = q * 32 + lane;
    const int tok = T.tok[row];
    const int th_i = tok & 255, tw_i = tok >> 8;
    const char* tbl_i = tbl_bytes + T.aq4[row];
    const int code_i = T.code[row];
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float m_prev = 0.f;
    long long grow_prev = 0;
    int g_prev = 0;

    // O[row, hf*16 .. +16) / l -> bf16 -> global; LSE (natural log) from the column-half-0 warp
    auto epilogue = [&](int bp) {
      const float l = rowsum[(bp * 2) * 128 + row] + rowsum[(bp * 2 + 1) * 128 + row];
      uint32_t o[16];
      tmem_ld16p(lane_addr + TF_O_COL + hf * 16, o);
      tmem_ld_wait();
      const float inv = 1.0f / l;
      uint4 v0, v1;
      v0.x = pack_bf16(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
      v0.y = pack_bf16(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
      v0.z = pack_bf16(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
      v0.w = pack_bf16(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
      v1.x = pack_bf16(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
      v1.y = pack_bf16(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
      v1.z = pack_bf16(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
      v1.w = pack_bf16(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
      bf16* dst = p.o + grow_prev * p.ldo + h * WA_HD + hf * 16;
      *reinterpret_cast<uint4*>(dst) = v0;
      *reinterpret_cast<uint4*>(dst + 8) = v1;
      if (hf == 0 && p.lse)
        p.lse[(static_cast<long long>(g_prev) * p.nH + h) * TC_N + row] = (m_prev + lg2_approx(l)) * WA_LN2;
    };

#pragma unroll 1
    for (int it = 0; it < n_my; ++it) {
      const int g = blockIdx.y + it * gridDim.y;
      const int b = it & 1;
      long long img_base; int h0, w0, emask;
      geo.decode(g, img_base, h0, w0, emask);

      tc_wait(&s_full[b], (it >> 1) & 1, 101, it);
      tc_fence_after();
      uint32_t v[72];  // fp32 bit patterns: scores, then probabilities
      {
        // every load naturally aligned to its own width (columns 0 | 32 | 64 and 72 | 80 | 96 | 128)
        const uint32_t ta = lane_addr + (b ? TF_S_COL1 : TF_S_COL0);
        if (hf == 0) {
          tmem_ld32p(ta, v);
          tmem_ld32p(ta + 32, v + 32);
          tmem_ld8p(ta + 64, v + 64);
        } else {
          tmem_ld8p(ta + 72, v);
          tmem_ld16p(ta + 80, v + 8);
          tmem_ld32p(ta + 96, v + 24);
          tmem_ld16p(ta + 128, v + 56);
        }
        tmem_ld_wait();
      }
      tc_fence_before();
      if (lane == 0) mbar_arrive(&s_empty[b]);

      float madd[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) madd[c] = ((code_i ^ c) & emask) ? WA_MASK2 : 0.f;
      float mx;
      if (hf == 0)
        mx = emask ? tc_scores<0, true>(v, tbl_i, scale2, madd) : tc_scores<0, false>(v, tbl_i, scale2, madd);
      else
        mx = emask ? tc_scores<1, true>(v, tbl_i, scale2, madd) : tc_scores<1, false>(v, tbl_i, scale2, madd);
      rowmax[(b * 2 + hf) * 128 + row] = mx;
      named_bar_sync(1 + q, 64);  // the two warps that share this TMEM lane quadrant
      const float m = fmaxf(mx, rowmax[(b * 2 + (hf ^ 1)) * 128 + row]);

      if (it > 0) {  // previous window: P V has finished (O complete, P buffer free)
        tc_wait(o_full, (it - 1) & 1, 102, it);
        tc_fence_after();
        epilogue(b ^ 1);
      }

      float sum = 0.f;
#pragma unroll
      for (int c8 = 0; c8 < 9; ++c8) {
        float pr[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          pr[e] = ex2_approx(__uint_as_float(v[c8 * 8 + e]) - m);
          sum += pr[e];
        }
        const int p8 = hf * 9 + c8;  // 16-byte piece (8 keys) of the 144-key row
        uint4 o;
        o.x = pack_bf16(pr[0], pr[1]);
        o.y = pack_bf16(pr[2], pr[3]);
        o.z = pack_bf16(pr[4], pr[5]);
        o.w = pack_bf16(pr[6], pr[7]);
        *reinterpret_cast<uint4*>(sP + pds_piece_off(row, p8, TF_PCHUNK)) = o;
      }
      rowsum[(b * 2 + hf) * 128 + row] = sum;
      fence_proxy_async_smem();  // P stores -> visible to the tensor core (async proxy)
      tc_fence_before();         // orders the TMEM reads of the epilogue before the issuer's next P V
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);

      m_prev = m;
      g_prev = g;
      grow_prev = geo.row(img_base, h0, w0, th_i, tw_i);
    }
    named_bar_sync(1 + q, 64);  // partner's row sums of the last window
    tc_wait(o_full, (n_my - 1) & 1, 102, n_my);
    tc_fence_after();
    epilogue((n_my - 1) & 1);
  } else if (warp == TC_WARP_MMA) {
    // ================= tcgen05.mma issuer =================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, TC_N, 0, 0);   // S = Q K^T   (both K-major)
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, WA_HD, 0, 1);  // O = P V     (V MN-major)
    auto issue_s = [&](int it) {
      const int s = it % TF_STAGES, b = it & 1;
      tc_wait(&full[s], (it / TF_STAGES) & 1, 103, it);
      tc_wait(&s_empty[b], ((it >> 1) & 1) ^ 1, 104, it);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t q_addr = smem_u32(smem + s * TF_STAGE_BYTES), k_addr = q_addr + TC_TILE;
        const uint32_t d = tmem_base + (b ? TF_S_COL1 : TF_S_COL0);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
          umma_f16_ss(d, desc_tile_kmajor(q_addr, ks), desc_tile_kmajor(k_addr, ks), idesc_s, ks);
        umma_commit(&s_full[b]);
      }
      __syncwarp();
    };
    issue_s(0);
#pragma unroll 1
    for (int it = 0; it < n_my; ++it) {
      if (it + 1 < n_my) issue_s(it + 1);  // next window's scores overlap this window's softmax
      tc_wait(p_ready, it & 1, 105, it);
      tc_fence_after();
      if (lane == 0) {
        const int s = it % TF_STAGES;
        const uint32_t v_addr = smem_u32(smem + s * TF_STAGE_BYTES + 2 * TC_TILE), p_addr = smem_u32(sP);
#pragma unroll
        for (int kk = 0; kk < 9; ++kk)  // 16 keys per step
          umma_f16_ss(tmem_base + TF_O_COL, desc_pds_kmajor(p_addr, kk, TF_PCHUNK), desc_tile_mnmajor(v_addr, kk),
                      idesc_o, kk);
        umma_commit(o_full);
        umma_commit(&stage_free[s]);
      }
      __syncwarp();
    }
  } else if (warp == TC_WARP_LD) {
    // ================= tile loader =================
    const int piece = lane & 3;
    const int col0 = h * WA_HD + piece * 8;
#pragma unroll 1
    for (int it = 0; it < n_my; ++it) {
      if (it > 0) {  // the previous window's copies (issued one iteration ago) have had a window's time to land
        cp_async_wait0();
        fence_proxy_async_smem();
        mbar_arrive(&full[(it - 1) % TF_STAGES]);
      }
      const int s = it % TF_STAGES;
      if (it >= TF_STAGES) tc_wait(&stage_free[s], (it / TF_STAGES - 1) & 1, 106, it);
      const int g = blockIdx.y + it * gridDim.y;
      long long img_base; int h0, w0, em;
      geo.decode(g, img_base, h0, w0, em);
      const uint32_t st = smem_u32(smem + s * TF_STAGE_BYTES);
#pragma unroll 6
      for (int k = 0; k < 18; ++k) {
        const int r = (lane >> 2) + 8 * k;
        const long long grow = geo.row(img_base, h0, w0, r / TC_WS, r % TC_WS);
        const uint32_t so = st + sw64_off(r, piece);
        cp_async16(so, p.q + grow * p.ldq + col0, true);
        cp_async16(so + TC_TILE, p.k + grow * p.ldk + col0, true);
        cp_async16(so + 2 * TC_TILE, p.v + grow * p.ldv + col0, true);
      }
      cp_async_commit();
    }
    cp_async_wait0();
    fence_proxy_async_smem();
    mbar_arrive(&full[(n_my - 1) % TF_STAGES]);
  } else {
    // ================= remainder rows 128..143 (mma.sync), windows it = k, k + 2, ... =================
    const int k = warp - TC_WARP_R0;
#pragma unroll 1
    for (int it = k; it < n_my; it += 2) {
      const int s = it % TF_STAGES;
      const int g = blockIdx.y + it * gridDim.y;
      long long img_base; int h0, w0, emask;
      geo.decode(g, img_base, h0, w0, emask);
      tc_wait(&full[s], (it / TF_STAGES) & 1, 103, it);
      const uint32_t sQ = smem_u32(smem + s * TF_STAGE_BYTES), sK = sQ + TC_TILE, sV = sK + TC_TILE;
      float m_run[2] = {-1e30f, -1e30f}, l_run[2] = {0.f, 0.f};
      float oacc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;
      if (emask)
        tc_fwd_rem_tile<true>(sQ, sK, sV, T, tbl_bytes, lane, scale2, emask, oacc, m_run, l_run);
      else
        tc_fwd_rem_tile<false>(sQ, sK, sV, T, tbl_bytes, lane, scale2, 0, oacc, m_run, l_run);
      const int rl0 = 128 + (lane >> 2);
      float inv[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float l = l_run[r];
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        inv[r] = 1.0f / l;
        const int qi = rl0 + r * 8;
        if ((lane & 3) == 0 && p.lse)
          p.lse[(static_cast<long long>(g) * p.nH + h) * TC_N + qi] = (m_run[r] + lg2_approx(l)) * WA_LN2;
      }
      // Q rows 128..143 are read by this warp only (the UMMA tile is rows 0..127): reuse them as staging
      __syncwarp();
      uint8_t* q_tile = smem + s * TF_STAGE_BYTES;
#pragma unroll
      for (int dt = 0; dt < 4; ++dt) {
        const int cb = (lane & 3) * 4;  // byte offset inside the 16-byte piece dt
        *reinterpret_cast<uint32_t*>(q_tile + sw64_off(rl0, dt) + cb) = pack_bf16(oacc[dt][0] * inv[0], oacc[dt][1] * inv[0]);
        *reinterpret_cast<uint32_t*>(q_tile + sw64_off(rl0 + 8, dt) + cb) = pack_bf16(oacc[dt][2] * inv[1], oacc[dt][3] * inv[1]);
      }
      __syncwarp();
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int c = lane + 32 * kk;
        const int i = 128 + (c >> 2), cc = c & 3;
        const long long grow = geo.row(img_base, h0, w0, i / TC_WS, i % TC_WS);
        *reinterpret_cast<uint4*>(p.o + grow * p.ldo + h * WA_HD + cc * 8) =
            *reinterpret_cast<const uint4*>(q_tile + sw64_off(i, cc));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&stage_free[s]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TC_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =================================================================================================
// Backward
//   S = Q K^T and dP = dO V^T into TMEM (rows = queries 0..127); element-wise warps turn them into
//   P = exp2(S*scale2 + bias2 - lse2) and dS = P * (dP - D) (bf16, [query][key] chunks in smem) and keep the
//   d(bias) sums of their fixed (query, key) positions in registers; then
//   dV = P^T dO, dK = dS^T Q (A MN-major: M = keys 0..127, K = the 144 queries) and dQ = dS K (A K-major).
//   Remainder warps (mma.sync): scores / dS of query rows 128..143 and dV / dK / dQ of token rows 128..143.
// =================================================================================================
constexpr int TB_STAGES = 2;
constexpr int TB_STAGE_BYTES = 4 * TC_TILE;                 // Q, dO, K, V
constexpr int TB_PCHUNK = tcl::B_PCHUNK;                    // P / dS chunk: 144 query rows x 64 keys (18432 B)
constexpr int TB_OFF_P = TB_STAGES * TB_STAGE_BYTES;        // 73728
constexpr int TB_OFF_DS = TB_OFF_P + 3 * TB_PCHUNK;         // 129024
constexpr int TB_OFF_STG = TB_OFF_DS + 3 * TB_PCHUNK;       // 184320: [2 remainder warps][16 rows][64 B]
constexpr int TB_OFF_TBL = TB_OFF_STG + 2 * 1024;
constexpr int TB_OFF_BARS = TB_OFF_TBL + ((TC_TABLE_BYTES + 15) & ~15);
constexpr int TB_SMEM = 1024 + TB_OFF_BARS + 16 * 8;
constexpr uint32_t TB_S_COL = 0, TB_DP_COL = 160, TB_DV_COL = 320, TB_DK_COL = 352, TB_DQ_COL = 384;
constexpr int TB_ACCP = 146;  // fp32 pitch of the d(bias) flush matrix aliased onto the P / dS chunks
static_assert(TC_N * TB_ACCP * 4 <= 6 * TB_PCHUNK, "flush matrix must fit in the P / dS chunks");
static_assert(TB_SMEM <= 227 * 1024, "backward shared memory");
static_assert(TF_SMEM <= 227 * 1024, "forward shared memory");

// The 72 key columns [HF * 72, +72) of one query row in nine 8-column pieces (= one 16-byte store each): P, dS,
// d(bias) sums.  The TMEM loads of piece k + 1 are in flight while piece k is computed.
template <int HF, bool MASKED>
__device__ __forceinline__ void tc_bwd_row(uint32_t lane_addr, uint8_t* sP, uint8_t* sdS, int row, const char* tbl_i,
                                           float scale2, float nlse2, float negD, const float (&madd)[4],
                                           float (&dbacc)[72]) {
  constexpr int JB = HF * 72;
  uint32_t s[2][8], dp[2][8];
  tmem_ld8p(lane_addr + TB_S_COL + JB, s[0]);
  tmem_ld8p(lane_addr + TB_DP_COL + JB, dp[0]);
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    tmem_ld_wait();
    if (k + 1 < 9) {
      tmem_ld8p(lane_addr + TB_S_COL + JB + 8 * (k + 1), s[(k + 1) & 1]);
      tmem_ld8p(lane_addr + TB_DP_COL + JB + 8 * (k + 1), dp[(k + 1) & 1]);
    }
    float pr[8], ds[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int j = JB + 8 * k + e;
      const float t = *reinterpret_cast<const float*>(tbl_i - 4 * tc_bj(j)) + nlse2;
      float x = fmaf(__uint_as_float(s[k & 1][e]), scale2, t);
      if (MASKED) x += madd[tc_code(j)];
      pr[e] = ex2_approx(x);
      ds[e] = pr[e] * (__uint_as_float(dp[k & 1][e]) + negD);
      dbacc[8 * k + e] += ds[e];
    }
    const int p8 = HF * 9 + k;  // 16-byte piece (8 keys) of the 144-key row
    const uint32_t off = pds_piece_off(row, p8, TB_PCHUNK);
    uint4 a, b;
    a.x = pack_bf16(pr[0], pr[1]); a.y = pack_bf16(pr[2], pr[3]);
    a.z = pack_bf16(pr[4], pr[5]); a.w = pack_bf16(pr[6], pr[7]);
    b.x = pack_bf16(ds[0], ds[1]); b.y = pack_bf16(ds[2], ds[3]);
    b.z = pack_bf16(ds[4], ds[5]); b.w = pack_bf16(ds[6], ds[7]);
    *reinterpret_cast<uint4*>(sP + off) = a;
    *reinterpret_cast<uint4*>(sdS + off) = b;
  }
}

// Remainder score job (window_attn.cu's w3_job_a for the row tile 128..143): 16 queries x key third `third`.
// S and dP accumulators start at -lse/scale and -D; P = exp2, dS = P*dP go to the smem chunks as bf16.
template <bool MASKED>
__device__ __forceinline__ void tc_bwd_rem_scores(uint32_t sQ, uint32_t sdO, uint32_t sK, uint32_t sV, uint8_t* sP,
                                                  uint8_t* sdS, float nl0, float nl1, float nd0, float nd1,
                                                  const WinTables& T, const char* tbl_bytes, int third, int lane,
                                                  float scale2, int emask, float (&dbacc)[6][4]) {
  uint32_t qf[2][4], dof[2][4];
  {
    const int row = 128 + (lane & 7) + ((lane >> 3) & 1) * 8;
    const int pc = lane >> 4;
    ldsm_x4(sQ + sw64_off(row, pc), qf[0][0], qf[0][1], qf[0][2], qf[0][3]);
    ldsm_x4(sQ + sw64_off(row, pc + 2), qf[1][0], qf[1][1], qf[1][2], qf[1][3]);
    ldsm_x4(sdO + sw64_off(row, pc), dof[0][0], dof[0][1], dof[0][2], dof[0][3]);
    ldsm_x4(sdO + sw64_off(row, pc + 2), dof[1][0], dof[1][1], dof[1][2], dof[1][3]);
  }
  const int rl0 = 128 + (lane >> 2);
  const int c2 = (lane & 3) * 2;
  const int aq0 = T.aq4[rl0], aq1 = T.aq4[rl0 + 8];
  int ci0 = 0, ci1 = 0;
  if (MASKED) {
    ci0 = T.code[rl0] & emask;
    ci1 = T.code[rl0 + 8] & emask;
  }
#pragma unroll
  for (int nt2 = 0; nt2 < 3; ++nt2) {  // 16 keys = two n-tiles per slice
    float s[2][4], dp[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      s[i][0] = s[i][1] = nl0; s[i][2] = s[i][3] = nl1;
      dp[i][0] = dp[i][1] = nd0; dp[i][2] = dp[i][3] = nd1;
    }
    const int krow = third * 48 + nt2 * 16 + (lane & 7) + ((lane >> 4) << 3);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int pc = ks * 2 + ((lane >> 3) & 1);
      uint32_t b0, b1, b2, b3;
      ldsm_x4(sK + sw64_off(krow, pc), b0, b1, b2, b3);
      mma16816(s[0], qf[ks], b0, b1);
      mma16816(s[1], qf[ks], b2, b3);
      ldsm_x4(sV + sw64_off(krow, pc), b0, b1, b2, b3);
      mma16816(dp[0], dof[ks], b0, b1);
      mma16816(dp[1], dof[ks], b2, b3);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int nt = nt2 * 2 + i;
      const int j0 = third * 48 + nt * 8 + c2;
      const int2 bj = *reinterpret_cast<const int2*>(T.bj4 + j0);
      int2 cj = make_int2(0, 0);
      if (MASKED) {
        cj = *reinterpret_cast<const int2*>(T.code + j0);
        cj.x &= emask;
        cj.y &= emask;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool hi = e >> 1, odd = e & 1;
        const float t = *reinterpret_cast<const float*>(tbl_bytes + ((hi ? aq1 : aq0) - (odd ? bj.y : bj.x)));
        float v = fmaf(s[i][e], scale2, t);
        if (MASKED) {
          if ((hi ? ci1 : ci0) != (odd ? cj.y : cj.x)) v += WA_MASK2;
        }
        const float pr = ex2_approx(v);
        const float ds = pr * dp[i][e];
        dbacc[nt][e] += ds;
        s[i][e] = pr;
        dp[i][e] = ds;
      }
      *reinterpret_cast<uint32_t*>(sP + pds_off(rl0, j0, TB_PCHUNK)) = pack_bf16(s[i][0], s[i][1]);
      *reinterpret_cast<uint32_t*>(sP + pds_off(rl0 + 8, j0, TB_PCHUNK)) = pack_bf16(s[i][2], s[i][3]);
      *reinterpret_cast<uint32_t*>(sdS + pds_off(rl0, j0, TB_PCHUNK)) = pack_bf16(dp[i][0], dp[i][1]);
      *reinterpret_cast<uint32_t*>(sdS + pds_off(rl0 + 8, j0, TB_PCHUNK)) = pack_bf16(dp[i][2], dp[i][3]);
    }
  }
}

// Remainder output job for token rows 128..143 (window_attn.cu's wb_phase_b with tile = 8):
//   type 0: dV = P^T dO   type 1: dK = dS^T Q   (reduction over the 144 queries; A^T read from chunk 2)
//   type 2: dQ = dS K                           (reduction over the 144 keys)
__device__ __forceinline__ void tc_bwd_rem_out(int type, uint32_t sQ, uint32_t sdO, uint32_t sK, uint32_t sP,
                                               uint32_t sdS, int lane, float (&acc)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  const uint32_t Bm = type == 0 ? sdO : (type == 1 ? sQ : sK);
  const int brow = (lane & 7) + ((lane >> 3) & 1) * 8;
  const int bpc = lane >> 4;
  if (type < 2) {
    const uint32_t Am = (type == 0 ? sP : sdS) + 2 * TB_PCHUNK;  // keys 128..143 = pieces 0, 1 of chunk 2
    const int arow = (lane & 7) + ((lane >> 4) << 3);             // reduction (query) index inside the 16-step
    const int apc = (lane >> 3) & 1;
#pragma unroll 3
    for (int kt = 0; kt < 9; ++kt) {
      uint32_t a[4], b0, b1, b2, b3;
      ldsm_x4_t(Am + sw128_off(kt * 16 + arow, apc), a[0], a[1], a[2], a[3]);
      ldsm_x4_t(Bm + sw64_off(kt * 16 + brow, bpc), b0, b1, b2, b3);
      mma16816(acc[0], a, b0, b1);
      mma16816(acc[1], a, b2, b3);
      ldsm_x4_t(Bm + sw64_off(kt * 16 + brow, bpc + 2), b0, b1, b2, b3);
      mma16816(acc[2], a, b0, b1);
      mma16816(acc[3], a, b2, b3);
    }
  } else {
    const int arow = 128 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll 3
    for (int kt = 0; kt < 9; ++kt) {
      uint32_t a[4], b0, b1, b2, b3;
      ldsm_x4(sdS + (kt >> 2) * TB_PCHUNK + sw128_off(arow, (kt & 3) * 2 + (lane >> 4)), a[0], a[1], a[2], a[3]);
      ldsm_x4_t(Bm + sw64_off(kt * 16 + brow, bpc), b0, b1, b2, b3);
      mma16816(acc[0], a, b0, b1);
      mma16816(acc[1], a, b2, b3);
      ldsm_x4_t(Bm + sw64_off(kt * 16 + brow, bpc + 2), b0, b1, b2, b3);
      mma16816(acc[2], a, b0, b1);
      mma16816(acc[3], a, b2, b3);
    }
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1) win_attn_tc_bwd_kernel(const AttnParams p, const float* __restrict__ Dg) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ array, so that the compiler keeps the address space
  // (LDS / STS instead of generic LD / ST for the table gathers and the P / dS stores)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sP = smem + TB_OFF_P;
  uint8_t* sdS = smem + TB_OFF_DS;
  WinTables T;
  T.tbl2 = reinterpret_cast<float*>(smem + TB_OFF_TBL);
  T.aq4 = reinterpret_cast<int*>(T.tbl2 + 2 * WA_MAXTBL + 2);
  T.bj4 = T.aq4 + WA_ROWS;
  T.code = T.bj4 + WA_ROWS;
  T.tok = T.code + WA_ROWS;
  const char* tbl_bytes = reinterpret_cast<const char*>(T.tbl2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TB_OFF_BARS);
  // timeout tags (tc_wait): 201 s_full, 202 rem_done, 203 acc_full, 204 full, 205 sdp_empty, 206 pds_ready,
  //                         207 acc_empty, 208 stage_free
  uint64_t* full = bars;            // [2] tiles of a stage have landed                     (32 loader lanes)
  uint64_t* stage_free = bars + 2;  // [2] stage may be overwritten                          (commit + 2 remainder warps)
  uint64_t* s_full = bars + 4;      //     S and dP accumulators written                     (commit)
  uint64_t* sdp_empty = bars + 5;   //     S and dP read into registers                      (8 element-wise warps)
  uint64_t* pds_ready = bars + 6;   //     P and dS complete in smem                         (8 + 2 warps)
  uint64_t* acc_full = bars + 7;    //     dV / dK / dQ accumulators written, P / dS consumed (commit)
  uint64_t* acc_empty = bars + 8;   //     accumulators drained                              (8 element-wise warps)
  uint64_t* rem_done = bars + 9;    //     remainder warps have read P / dS of the window    (2 remainder warps)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 10);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x;
  WinGeo geo;
  geo.H = p.H; geo.W = p.W; geo.ws = TC_WS; geo.shift = p.shift;
  geo.nWw = p.W / TC_WS; geo.nWh = p.H / TC_WS; geo.nW = geo.nWh * geo.nWw;
  const int n_groups = p.G * geo.nW;
  const int n_my = (n_groups - static_cast<int>(blockIdx.y) + static_cast<int>(gridDim.y) - 1) / static_cast<int>(gridDim.y);
  const float scale2 = p.scale * WA_LOG2E;

  pdl_trigger();
  pdl_wait();
  fill_tables(T, p.bias_table, p.nH, h, TC_WS, p.shift, TC_N, tid, TC_THREADS);
  if (warp == TC_WARP_MMA) {
    if (lane == 0) {
      for (int s = 0; s < TB_STAGES; ++s) {
        mbar_init(&full[s], 32);
        mbar_init(&stage_free[s], 3);
      }
      mbar_init(s_full, 1);
      mbar_init(sdp_empty, 8);
      mbar_init(pds_ready, 10);
      mbar_init(acc_full, 1);
      mbar_init(acc_empty, 8);
      mbar_init(rem_done, 2);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // d(bias): every thread keeps the sums of the (query, key) positions it owns in registers over all windows of the
  // CTA.  When a role has finished its windows it meets the others at named barrier 6 (all MMAs have completed: the
  // element-wise warps waited for the last acc_full) and flushes its sums into an fp32 [144][TB_ACCP] matrix aliased
  // onto the P / dS chunks — every (i, j) has exactly one owner, so no atomics.  The register arrays are scoped to
  // their role's branch so that they do not add up in the allocator.
  float* sAcc = reinterpret_cast<float*>(sP);

  if (warp < 8) {
    // ================= element-wise warps =================
    const int q = warp & 3, hf = warp >> 2, row = q * 32 + lane;
    const int tok = T.tok[row];
    const int th_i = tok & 255, tw_i = tok >> 8;
    const char* tbl_i = tbl_bytes + T.aq4[row];
    const int code_i = T.code[row];
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float dbacc[72];  // (row, hf * 72 + c)
#pragma unroll
    for (int c = 0; c < 72; ++c) dbacc[c] = 0.f;
#pragma unroll 1
    for (int it = 0; it < n_my; ++it) {
      const int g = blockIdx.y + it * gridDim.y;
      long long img_base; int h0, w0, emask;
      geo.decode(g, img_base, h0, w0, emask);
      const long long grow = geo.row(img_base, h0, w0, th_i, tw_i);
      const float nlse2 = -p.lse[(static_cast<long long>(g) * p.nH + h) * TC_N + row] * WA_LOG2E;
      const float negD = -Dg[grow * p.nH + h];
      float madd[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) madd[c] = ((code_i ^ c) & emask) ? WA_MASK2 : 0.f;

      tc_wait(s_full, it & 1, 201, it);
      tc_fence_after();
      // P / dS buffers are free: the tensor core is done with them (this thread waited for acc_full of the previous
      // window in its drain below) and so are the remainder warps' output jobs
      if (it > 0) tc_wait(rem_done, (it - 1) & 1, 202, it);
      if (hf == 0) {
        if (emask) tc_bwd_row<0, true>(lane_addr, sP, sdS, row, tbl_i, scale2, nlse2, negD, madd, dbacc);
        else       tc_bwd_row<0, false>(lane_addr, sP, sdS, row, tbl_i, scale2, nlse2, negD, madd, dbacc);
      } else {
        if (emask) tc_bwd_row<1, true>(lane_addr, sP, sdS, row, tbl_i, scale2, nlse2, negD, madd, dbacc);
        else       tc_bwd_row<1, false>(lane_addr, sP, sdS, row, tbl_i, scale2, nlse2, negD, madd, dbacc);
      }
      tc_fence_before();
      if (lane == 0) mbar_arrive(sdp_empty);  // S / dP may be overwritten by the next window's MMAs
      fence_proxy_async_smem();               // P / dS stores -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_ready);

      // drain dV / dK / dQ of token `row`, head-dim columns [hf * 16, +16)
      tc_wait(acc_full, it & 1, 203, it);
      tc_fence_after();
#pragma unroll
      for (int t = 0; t < 3; ++t) {  // one accumulator at a time: 16 live registers next to the 72 d(bias) sums
        uint32_t a[16];
        tmem_ld16p(lane_addr + (t == 0 ? TB_DV_COL : (t == 1 ? TB_DK_COL : TB_DQ_COL)) + hf * 16, a);
        tmem_ld_wait();
        const float sc = t == 0 ? 1.0f : p.scale;
        bf16* dst = (t == 0 ? p.dv + grow * p.lddv : (t == 1 ? p.dk + grow * p.lddk : p.dq + grow * p.lddq)) +
                    h * WA_HD + hf * 16;
        uint4 v0, v1;
        v0.x = pack_bf16(__uint_as_float(a[0]) * sc, __uint_as_float(a[1]) * sc);
        v0.y = pack_bf16(__uint_as_float(a[2]) * sc, __uint_as_float(a[3]) * sc);
        v0.z = pack_bf16(__uint_as_float(a[4]) * sc, __uint_as_float(a[5]) * sc);
        v0.w = pack_bf16(__uint_as_float(a[6]) * sc, __uint_as_float(a[7]) * sc);
        v1.x = pack_bf16(__uint_as_float(a[8]) * sc, __uint_as_float(a[9]) * sc);
        v1.y = pack_bf16(__uint_as_float(a[10]) * sc, __uint_as_float(a[11]) * sc);
        v1.z = pack_bf16(__uint_as_float(a[12]) * sc, __uint_as_float(a[13]) * sc);
        v1.w = pack_bf16(__uint_as_float(a[14]) * sc, __uint_as_float(a[15]) * sc);
        *reinterpret_cast<uint4*>(dst) = v0;
        *reinterpret_cast<uint4*>(dst + 8) = v1;
      }
      tc_fence_before();
      if (lane == 0) mbar_arrive(acc_empty);
    }
    named_bar_sync(6, TC_THREADS);
#pragma unroll
    for (int c = 0; c < 72; c += 2)
      *reinterpret_cast<float2*>(sAcc + row * TB_ACCP + hf * 72 + c) = make_float2(dbacc[c], dbacc[c + 1]);
  } else if (warp == TC_WARP_MMA) {
    // ================= tcgen05.mma issuer =================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, TC_N, 0, 0);    // S = Q K^T, dP = dO V^T (K-major x K-major)
    constexpr uint32_t idesc_kv = umma_idesc_bf16(128, WA_HD, 1, 1);  // dV = P^T dO, dK = dS^T Q (MN x MN)
    constexpr uint32_t idesc_q = umma_idesc_bf16(128, WA_HD, 0, 1);   // dQ = dS K (K-major x MN-major)
    const uint32_t p_addr = smem_u32(sP), ds_addr = smem_u32(sdS);
#pragma unroll 1
    for (int it = 0; it < n_my; ++it) {
      const int s = it & 1;
      const uint32_t q_addr = smem_u32(smem + s * TB_STAGE_BYTES);
      const uint32_t do_addr = q_addr + TC_TILE, k_addr = q_addr + 2 * TC_TILE, v_addr = q_addr + 3 * TC_TILE;
      tc_wait(&full[s], (it >> 1) & 1, 204, it);
      tc_wait(sdp_empty, (it & 1) ^ 1, 205, it);
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
          umma_f16_ss(tmem_base + TB_S_COL, desc_tile_kmajor(q_addr, ks), desc_tile_kmajor(k_addr, ks), idesc_s, ks);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
          umma_f16_ss(tmem_base + TB_DP_COL, desc_tile_kmajor(do_addr, ks), desc_tile_kmajor(v_addr, ks), idesc_s, ks);
        umma_commit(s_full);
      }
      __syncwarp();
      tc_wait(pds_ready, it & 1, 206, it);
      tc_wait(acc_empty, (it & 1) ^ 1, 207, it);
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int kk = 0; kk < 9; ++kk) {  // 16 queries per step
          umma_f16_ss(tmem_base + TB_DV_COL, desc_pds_mnmajor(p_addr, kk), desc_tile_mnmajor(do_addr, kk), idesc_kv, kk);
          umma_f16_ss(tmem_base + TB_DK_COL, desc_pds_mnmajor(ds_addr, kk), desc_tile_mnmajor(q_addr, kk), idesc_kv, kk);
        }
#pragma unroll
        for (int kk = 0; kk < 9; ++kk)  // 16 keys per step
          umma_f16_ss(tmem_base + TB_DQ_COL, desc_pds_kmajor(ds_addr, kk, TB_PCHUNK), desc_tile_mnmajor(k_addr, kk),
                      idesc_q, kk);
        umma_commit(acc_full);
        umma_commit(&stage_free[s]);
      }
      __syncwarp();
    }
    named_bar_sync(6, TC_THREADS);
  } else if (warp == TC_WARP_LD) {
    // ================= tile loader =================
    const int piece = lane & 3;
    const int col0 = h * WA_HD + piece * 8;
#pragma unroll 1
    for (int it = 0; it < n_my; ++it) {
      if (it > 0) {
        cp_async_wait0();
        fence_proxy_async_smem();
        mbar_arrive(&full[(it - 1) & 1]);
      }
      const int s = it & 1;
      if (it >= TB_STAGES) tc_wait(&stage_free[s], ((it >> 1) - 1) & 1, 208, it);
      const int g = blockIdx.y + it * gridDim.y;
      long long img_base; int h0, w0, em;
      geo.decode(g, img_base, h0, w0, em);
      const uint32_t st = smem_u32(smem + s * TB_STAGE_BYTES);
#pragma unroll 6
      for (int k = 0; k < 18; ++k) {
        const int r = (lane >> 2) + 8 * k;
        const long long grow = geo.row(img_base, h0, w0, r / TC_WS, r % TC_WS);
        const uint32_t so = st + sw64_off(r, piece);
        cp_async16(so, p.q + grow * p.ldq + col0, true);
        cp_async16(so + TC_TILE, p.d_o + grow * p.lddo + col0, true);
        cp_async16(so + 2 * TC_TILE, p.k + grow * p.ldk + col0, true);
        cp_async16(so + 3 * TC_TILE, p.v + grow * p.ldv + col0, true);
      }
      cp_async_commit();
    }
    cp_async_wait0();
    fence_proxy_async_smem();
    mbar_arrive(&full[(n_my - 1) & 1]);
    named_bar_sync(6, TC_THREADS);
  } else {
    // ================= remainder warps (mma.sync) =================
    // R0: score jobs of key thirds 0 and 1, then dV of token rows 128..143
    // R1: score job of key third 2, then dK and dQ of token rows 128..143
    const int k = warp - TC_WARP_R0;
    float dbr[2][6][4];  // [score job slot][n-tile][fragment element], as in window_attn.cu
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int i = 0; i < 6; ++i) dbr[a][i][0] = dbr[a][i][1] = dbr[a][i][2] = dbr[a][i][3] = 0.f;
    uint8_t* stg = smem + TB_OFF_STG + k * 1024;
    const int rl0 = 128 + (lane >> 2);
    const int tok0 = T.tok[rl0], tok1 = T.tok[rl0 + 8];
    const float inv_scale = 1.0f / p.scale;
#pragma unroll 1
    for (int it = 0; it < n_my; ++it) {
      const int s = it & 1;
      const int g = blockIdx.y + it * gridDim.y;
      long long img_base; int h0, w0, emask;
      geo.decode(g, img_base, h0, w0, emask);
      const long long lse_base = (static_cast<long long>(g) * p.nH + h) * TC_N;
      const float nl0 = -p.lse[lse_base + rl0] * inv_scale, nl1 = -p.lse[lse_base + rl0 + 8] * inv_scale;
      const float nd0 = -Dg[geo.row(img_base, h0, w0, tok0 & 255, tok0 >> 8) * p.nH + h];
      const float nd1 = -Dg[geo.row(img_base, h0, w0, tok1 & 255, tok1 >> 8) * p.nH + h];
      const uint32_t sQ = smem_u32(smem + s * TB_STAGE_BYTES);
      const uint32_t sdO = sQ + TC_TILE, sK = sQ + 2 * TC_TILE, sV = sQ + 3 * TC_TILE;

      tc_wait(&full[s], (it >> 1) & 1, 204, it);
      if (it > 0) {
        tc_wait(acc_full, (it - 1) & 1, 203, it);  // the tensor core has consumed P / dS of the previous window
        named_bar_sync(5, 64);              // ... and so has the other remainder warp
      }
#pragma unroll
      for (int jb = 0; jb < 2; ++jb) {  // static slot index: dbr stays in registers
        if (jb == 0 || k == 0) {
          const int third = k == 0 ? jb : 2;
          if (emask)
            tc_bwd_rem_scores<true>(sQ, sdO, sK, sV, sP, sdS, nl0, nl1, nd0, nd1, T, tbl_bytes, third, lane, scale2,
                                    emask, dbr[jb]);
          else
            tc_bwd_rem_scores<false>(sQ, sdO, sK, sV, sP, sdS, nl0, nl1, nd0, nd1, T, tbl_bytes, third, lane, scale2, 0,
                                     dbr[jb]);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_ready);

      tc_wait(pds_ready, it & 1, 206, it);  // every P / dS element of this window is in smem
      const int t_first = k == 0 ? 0 : 1, t_last = k == 0 ? 0 : 2;
      for (int type = t_first; type <= t_last; ++type) {
        float acc[4][4];
        tc_bwd_rem_out(type, sQ, sdO, sK, smem_u32(sP), smem_u32(sdS), lane, acc);
        const float sc = type == 0 ? 1.0f : p.scale;
        const int r_lo = lane >> 2;
#pragma unroll
        for (int dt = 0; dt < 4; ++dt) {
          const int cb = (lane & 3) * 4;
          *reinterpret_cast<uint32_t*>(stg + sw64_off(r_lo, dt) + cb) = pack_bf16(acc[dt][0] * sc, acc[dt][1] * sc);
          *reinterpret_cast<uint32_t*>(stg + sw64_off(r_lo + 8, dt) + cb) = pack_bf16(acc[dt][2] * sc, acc[dt][3] * sc);
        }
        __syncwarp();
        bf16* outp = type == 0 ? p.dv : (type == 1 ? p.dk : p.dq);
        const long long ldo = type == 0 ? p.lddv : (type == 1 ? p.lddk : p.lddq);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const int r = (lane >> 2) + 8 * kk, cc = lane & 3;
          const int tk = T.tok[128 + r];
          const long long grow = geo.row(img_base, h0, w0, tk & 255, tk >> 8);
          *reinterpret_cast<uint4*>(outp + grow * ldo + h * WA_HD + cc * 8) =
              *reinterpret_cast<const uint4*>(stg + sw64_off(r, cc));
        }
        __syncwarp();  // staging tile is free again
      }
      if (lane == 0) {
        mbar_arrive(&stage_free[s]);
        mbar_arrive(rem_done);
      }
    }
    named_bar_sync(6, TC_THREADS);
    const int qi = 128 + (lane >> 2);
#pragma unroll
    for (int jb = 0; jb < 2; ++jb) {
      if (jb == 0 || k == 0) {
        const int third = k == 0 ? jb : 2;
#pragma unroll
        for (int nt = 0; nt < 6; ++nt) {
          const int j0 = third * 48 + nt * 8 + (lane & 3) * 2;
          *reinterpret_cast<float2*>(sAcc + qi * TB_ACCP + j0) = make_float2(dbr[jb][nt][0], dbr[jb][nt][1]);
          *reinterpret_cast<float2*>(sAcc + (qi + 8) * TB_ACCP + j0) = make_float2(dbr[jb][nt][2], dbr[jb][nt][3]);
        }
      }
    }
  }

  // one global atomic per table entry (dh, dw): the sum over the <= 144 token pairs with that offset
  tc_fence_before();
  __syncthreads();
  for (int t = tid; t < TC_TW2 * TC_TW2; t += TC_THREADS) {
    const int dh = t / TC_TW2 - (TC_WS - 1), dw = t % TC_TW2 - (TC_WS - 1);
    const int ih0 = max(0, dh), ih1 = min(TC_WS, TC_WS + dh), iw0 = max(0, dw), iw1 = min(TC_WS, TC_WS + dw);
    float sum = 0.f;
    for (int ih = ih0; ih < ih1; ++ih)
      for (int iw = iw0; iw < iw1; ++iw)
        sum += sAcc[(ih * TC_WS + iw) * TB_ACCP + (ih - dh) * TC_WS + (iw - dw)];
    atomicAdd(&p.dbias_table[t * p.nH + h], sum);
  }
  if (warp == TC_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// =================================================================================================
// Fourth generation, forward ("tq" = TMA + quadrant token order)
//
// What ncu showed for the kernel above (profiles/r2_winfwd_tc_ncu.txt): 151 M warp instructions for 16384
// window-heads, 32 % of the stall samples "no instruction" (the element-wise phase was 72 columns x 4 template
// variants of straight-line code, 69 KB of SASS), a shared-memory bias gather + 2 FADD + FMNMX per score, and a loader
// warp issuing 1728 16-byte cp.async per window.  This generation:
//   * TMA: a window is four 6 x 6-token quadrants (cyclic shift 0 or 6 moves whole quadrants), each ONE 4-D box
//     (32 channels x 6 x 6 tokens) of the [B, H, W, C] activation -> 12 cp.async.bulk.tensor per window instead of
//     1728 cp.async.  The tile rows are therefore in QUADRANT ORDER (row = 36 * quadrant + 6 * (th % 6) + tw % 6);
//     attention does not care about the token order inside a window as long as bias / mask indices follow, and
//     they come from tables filled in that order (fill_tables_quad).  The swizzle of a box written at a 128-byte
//     aligned row offset follows the absolute shared-memory address (tools/tc_probe2.cu, test B), i.e. sw64_off.
//   * the 72 bias values of a thread's (query row, key half) live in REGISTERS for the whole CTA (one head per
//     CTA); the SW-MSA mask is one addend per key QUADRANT (region code = quadrant), folded into the row maximum
//     and into the exponent offset -> no per-score mask work, one code path for shifted and unshifted windows.
//   * two passes over TMEM (tcgen05.ld runs at ~800 B/clk/SM, tools/tc_probe2.cu test A, so re-reading S is free):
//     pass 1 row maximum (FFMA2 + FMNMX3 per pair), pass 2 exp2 / row sum / bf16 pack (FFMA2, FADD2, 2 MUFU,
//     FADD2, F2FP per pair): 4.1 issue slots per score instead of ~9, MUFU-bound by construction.
//   * O double-buffered in TMEM, so the drain of window i - 1 runs after P of window i has been handed to the
//     tensor core instead of in front of the exp2 phase.
// =================================================================================================
__host__ __device__ constexpr int tq_th(int r) { return 6 * ((r / 36) >> 1) + (r % 36) / 6; }
__host__ __device__ constexpr int tq_tw(int r) { return 6 * ((r / 36) & 1) + (r % 36) % 6; }

__device__ __forceinline__ void fill_tables_quad(const WinTables& t, const float* __restrict__ bias_table, int nH, int h,
                                                 int shift, int tid, int nthreads) {
  for (int i = tid; i < WA_MAXTBL; i += nthreads) {
    t.tbl2[i] = bias_table[i * nH + h] * WA_LOG2E;
    t.tbl2[WA_MAXTBL + i] = -1e30f;
  }
  for (int r = tid; r < WA_ROWS; r += nthreads) {
    const int th = tq_th(r), tw = tq_tw(r);
    const int bidx = th * TC_TW2 + tw;
    t.aq4[r] = 4 * (bidx + (TC_WS - 1) * (TC_TW2 + 1));
    t.bj4[r] = 4 * bidx;
    t.code[r] = ((th >= TC_WS - shift) ? 1 : 0) | ((tw >= TC_WS - shift) ? 2 : 0);
    t.tok[r] = th | (tw << 8);
  }
}

// One 32-byte store per lane (sm_100: STG.256).  A scattered 16-byte-per-lane store costs the LSU one wavefront per
// lane; the drains of these kernels were bound by exactly that (tools/tq_trace.py: 2100 of 10900 cycles per window).
__device__ __forceinline__ void st_global_256(void* ptr, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x),
               "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
__device__ __forceinline__ uint64_t pk2u(uint32_t a, uint32_t b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
  return r;
}
template <int X>
__device__ __forceinline__ void tq_ld(uint32_t taddr, uint32_t* v) {
  if constexpr (X == 32) tmem_ld32p(taddr, v);
  else if constexpr (X == 16) tmem_ld16p(taddr, v);
  else tmem_ld8p(taddr, v);
}

constexpr int TQF_STAGES = 3;
constexpr int TQF_STAGE_BYTES = 3 * TC_TILE;
constexpr int TQF_OFF_P = TQF_STAGES * TQF_STAGE_BYTES;
constexpr int TQF_OFF_TBL = TQF_OFF_P + 3 * TF_PCHUNK;
constexpr int TQF_OFF_RMAX = TQF_OFF_TBL + ((TC_TABLE_BYTES + 15) & ~15);
constexpr int TQF_OFF_RSUM = TQF_OFF_RMAX + 2 * 2 * 128 * 4;
constexpr int TQF_OFF_BARS = TQF_OFF_RSUM + 2 * 2 * 128 * 4;
constexpr int TQF_SMEM = 1024 + TQF_OFF_BARS + 24 * 8;
// TMEM columns.  S = Q K^T is issued as two N = 72 MMAs (key halves) whose accumulators start at 32-column aligned
// bases, so that both key halves run the SAME element-wise code (x32 | x32 | x8 loads relative to the half's base):
// all eight element-wise warps share one instruction stream — the instruction cache was the limiter (ncu: 40 % of the
// stall samples "no instruction" with one code copy per half).
constexpr uint32_t TQF_S_HALF = 96, TQF_S_BUF = 192, TQF_O_COL0 = 384, TQF_O_COL1 = 416;
static_assert(TQF_SMEM <= 227 * 1024, "forward shared memory");

// Windows of a CTA: the contiguous range [first, first + n) of the head's windows (neighbouring windows of one image
// share L2 lines; stepping to the next window needs no division).
struct TqWindowIter {
  int img, wh, ww;
  __device__ __forceinline__ void init(int g, const WinGeo& geo) {
    img = g / geo.nW;
    const int w = g - img * geo.nW;
    wh = w / geo.nWw;
    ww = w - wh * geo.nWw;
  }
  __device__ __forceinline__ void next(const WinGeo& geo) {
    if (++ww == geo.nWw) {
      ww = 0;
      if (++wh == geo.nWh) {
        wh = 0;
        ++img;
      }
    }
  }
  __device__ __forceinline__ int emask(const WinGeo& geo) const {
    return geo.shift > 0 ? ((wh == geo.nWh - 1) ? 1 : 0) | ((ww == geo.nWw - 1) ? 2 : 0) : 0;
  }
  __device__ __forceinline__ int g(const WinGeo& geo) const { return (img * geo.nWh + wh) * geo.nWw + ww; }
  // activation row of the token at window coordinates (th, tw)
  __device__ __forceinline__ long long row(const WinGeo& geo, int th, int tw) const {
    int hp = wh * geo.ws + geo.shift + th, wp = ww * geo.ws + geo.shift + tw;
    hp -= hp >= geo.H ? geo.H : 0;
    wp -= wp >= geo.W ? geo.W : 0;
    return (static_cast<long long>(img) * geo.H + hp) * geo.W + wp;
  }
};
__device__ __forceinline__ void tq_my_windows(int n_groups, int& first, int& n_my) {
  const long long y = blockIdx.y, gy = gridDim.y;
  first = static_cast<int>(y * n_groups / gy);
  n_my = static_cast<int>((y + 1) * n_groups / gy) - first;
}

// pass 1 over the key columns [OFF, OFF + X) of this thread's half: running maxima of the two key quadrants
template <int OFF, int X>
__device__ __forceinline__ void tq_pass1_chunk(uint32_t taddr, const uint64_t (&bias2)[36], uint64_t scale2, float& m0, float& m1) {
  uint32_t v[X];
  tq_ld<X>(taddr + OFF, v);
  tmem_ld_wait();
#pragma unroll
  for (int e = 0; e < X; e += 2) {
    const uint64_t x = fma2(pk2u(v[e], v[e + 1]), scale2, bias2[(OFF + e) >> 1]);
    float x0, x1;
    upk2(x, x0, x1);
    if (OFF + e < 36) m0 = fmaxf(m0, fmaxf(x0, x1));
    else m1 = fmaxf(m1, fmaxf(x0, x1));
  }
}
// pass 2: P = exp2(s * scale2 + bias + (mask addend of the key quadrant - row maximum)), row sum, bf16 -> smem
template <int OFF, int X>
__device__ __forceinline__ void tq_pass2_chunk(uint32_t taddr, const uint64_t (&bias2)[36], uint64_t scale2, uint64_t c0,
                                               uint64_t c1, uint64_t& sum2, uint8_t* p_row, uint32_t xr, int p8b) {
  uint32_t v[X];
  tq_ld<X>(taddr + OFF, v);
  tmem_ld_wait();
#pragma unroll
  for (int e8 = 0; e8 < X; e8 += 8) {
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      const int k = OFF + e8 + e;
      uint64_t x = fma2(pk2u(v[e8 + e], v[e8 + e + 1]), scale2, bias2[k >> 1]);
      x = add2(x, k < 36 ? c0 : c1);
      float x0, x1;
      upk2(x, x0, x1);
      const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
      sum2 = add2(sum2, pk2(p0, p1));
      o[e >> 1] = pack_bf16(p0, p1);
    }
    const uint32_t p8 = p8b + ((OFF + e8) >> 3);  // 16-byte piece (8 keys) of the 144-key row: chunk p8 / 8, piece p8 % 8
    static_assert(TF_PCHUNK == (1 << 14), "chunk stride");
    *reinterpret_cast<uint4*>(p_row + ((p8 >> 3) << 14) + (((p8 & 7) << 4) ^ xr)) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// element-wise role of the forward kernel: thread = (query row, key half hf), persistent over the CTA's windows
__device__ __forceinline__ void tq_fwd_elementwise(const AttnParams& p, const WinGeo& geo, const WinTables& T, const char* tbl_bytes,
                                                   uint8_t* sP, float* rowmax, float* rowsum, uint64_t* s_full, uint64_t* s_empty,
                                                   uint64_t* p_ready, uint64_t* o_full, uint64_t* o_empty, uint32_t tmem_base,
                                                   int first, int n_my, int h, int warp, int lane, float scale2f) {
  const int q = warp & 3, hf = warp >> 2, row = q * 32 + lane;
  const int tok = T.tok[row];
  const int th_i = tok & 255, tw_i = tok >> 8;
  const int code_i = T.code[row];
  const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
  const uint64_t scale2 = pk2(scale2f, scale2f);
  uint64_t bias2[36];
  {
    const char* tbl_i = tbl_bytes + T.aq4[row];
    const int* bj = T.bj4 + hf * 72;
#pragma unroll
    for (int k = 0; k < 72; k += 2) {
      const int2 b2 = *reinterpret_cast<const int2*>(bj + k);
      bias2[k >> 1] = pk2(*reinterpret_cast<const float*>(tbl_i - b2.x), *reinterpret_cast<const float*>(tbl_i - b2.y));
    }
  }
  // region codes of this half's two key quadrants (tile columns [0, 36) and [36, 72) of the half): quadrant qd has
  // code (qd >> 1) | ((qd & 1) << 1), qd = 2 * hf and 2 * hf + 1
  const int code_c0 = hf, code_c1 = hf | 2;
  // this thread's P row: pieces hf * 9 .. hf * 9 + 8 of the 18 (three 64-key chunks of 128-byte swizzled rows)
  uint8_t* p_row = sP + row * 128;
  const uint32_t xr = static_cast<uint32_t>(row & 7) << 4;
  const int p8b = hf * 9;
  float m_prev = 0.f;
  long long grow_prev = 0;
  int g_prev = 0;
  TqWindowIter wi;
  wi.init(first, geo);

  auto epilogue = [&](int itp) {  // window itp: O buffer itp & 1, row sums of parity itp & 1
    const int bp = itp & 1;
    tc_wait(&o_full[bp], (itp >> 1) & 1, 102, itp);
    tc_fence_after();
    const float l = rowsum[(bp * 2) * 128 + row] + rowsum[(bp * 2 + 1) * 128 + row];
    uint32_t o[16];
    tmem_ld16p(lane_addr + (bp ? TQF_O_COL1 : TQF_O_COL0) + hf * 16, o);
    tmem_ld_wait();
    tc_fence_before();
    if (lane == 0) mbar_arrive(&o_empty[bp]);
    const float inv = 1.0f / l;
    uint4 v0, v1;
    v0.x = pack_bf16(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
    v0.y = pack_bf16(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
    v0.z = pack_bf16(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
    v0.w = pack_bf16(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
    v1.x = pack_bf16(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
    v1.y = pack_bf16(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
    v1.z = pack_bf16(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
    v1.w = pack_bf16(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
    st_global_256(p.o + grow_prev * p.ldo + h * WA_HD + hf * 16, v0, v1);
    if (hf == 0 && p.lse)  // LSE is stored in NATURAL token order (shared with the other kernel generations)
      p.lse[(static_cast<long long>(g_prev) * p.nH + h) * TC_N + th_i * TC_WS + tw_i] = (m_prev + lg2_approx(l)) * WA_LN2;
  };

#pragma unroll 1
  for (int it = 0; it < n_my; ++it) {
    const int b = it & 1;
    const int emask = wi.emask(geo);
    const float madd0 = ((code_i ^ code_c0) & emask) ? WA_MASK2 : 0.f;
    const float madd1 = ((code_i ^ code_c1) & emask) ? WA_MASK2 : 0.f;
    const uint32_t ta = lane_addr + b * TQF_S_BUF + hf * TQF_S_HALF;

    tc_wait(&s_full[b], (it >> 1) & 1, 101, it);
    tc_fence_after();
    float m0 = -1e30f, m1 = -1e30f;
    tq_pass1_chunk<0, 32>(ta, bias2, scale2, m0, m1);
    tq_pass1_chunk<32, 32>(ta, bias2, scale2, m0, m1);
    tq_pass1_chunk<64, 8>(ta, bias2, scale2, m0, m1);
    const float mx = fmaxf(m0 + madd0, m1 + madd1);
    rowmax[(b * 2 + hf) * 128 + row] = mx;
    named_bar_sync(1 + q, 64);  // the two warps that share this TMEM lane quadrant
    const float m = fmaxf(mx, rowmax[(b * 2 + (hf ^ 1)) * 128 + row]);

    if (it > 0) tc_wait(&o_full[b ^ 1], ((it - 1) >> 1) & 1, 107, it);  // P V of the previous window has consumed P
    uint64_t sum2 = pk2(0.f, 0.f);
    const uint64_t c0 = pk2(madd0 - m, madd0 - m), c1 = pk2(madd1 - m, madd1 - m);
    tq_pass2_chunk<0, 32>(ta, bias2, scale2, c0, c1, sum2, p_row, xr, p8b);
    tq_pass2_chunk<32, 32>(ta, bias2, scale2, c0, c1, sum2, p_row, xr, p8b);
    tq_pass2_chunk<64, 8>(ta, bias2, scale2, c0, c1, sum2, p_row, xr, p8b);
    tc_fence_before();
    if (lane == 0) mbar_arrive(&s_empty[b]);  // S buffer b may be overwritten (window it + 2)
    float s_lo, s_hi;
    upk2(sum2, s_lo, s_hi);
    rowsum[(b * 2 + hf) * 128 + row] = s_lo + s_hi;
    fence_proxy_async_smem();  // P stores -> visible to the tensor core (async proxy)
    __syncwarp();
    if (lane == 0) mbar_arrive(p_ready);

    if (it > 0) epilogue(it - 1);
    m_prev = m;
    g_prev = wi.g(geo);
    grow_prev = wi.row(geo, th_i, tw_i);
    wi.next(geo);
  }
  named_bar_sync(1 + q, 64);  // partner's row sums of the last window
  epilogue(n_my - 1);
}

__global__ void __launch_bounds__(TC_THREADS, 1) win_attn_tq_fwd_kernel(const AttnParams p, const __grid_constant__ CUtensorMap tmQ,
                                                                       const __grid_constant__ CUtensorMap tmK,
                                                                       const __grid_constant__ CUtensorMap tmV) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sP = smem + TQF_OFF_P;
  WinTables T;
  T.tbl2 = reinterpret_cast<float*>(smem + TQF_OFF_TBL);
  T.aq4 = reinterpret_cast<int*>(T.tbl2 + 2 * WA_MAXTBL + 2);
  T.bj4 = T.aq4 + WA_ROWS;
  T.code = T.bj4 + WA_ROWS;
  T.tok = T.code + WA_ROWS;
  const char* tbl_bytes = reinterpret_cast<const char*>(T.tbl2);
  float* rowmax = reinterpret_cast<float*>(smem + TQF_OFF_RMAX);  // [window parity][column half][128]
  float* rowsum = reinterpret_cast<float*>(smem + TQF_OFF_RSUM);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TQF_OFF_BARS);
  // timeout tags (tc_wait): 101 s_full, 102 o_full (drain), 103 full, 104 s_empty, 105 p_ready, 106 stage_free,
  //                         107 o_full (P buffer), 108 o_empty
  uint64_t* full = bars;            // [3] TMA boxes of a stage have landed            (expect_tx)
  uint64_t* stage_free = bars + 3;  // [3] stage may be overwritten                    (tcgen05.commit + remainder warp)
  uint64_t* s_full = bars + 6;      // [2] S accumulator written                       (tcgen05.commit)
  uint64_t* s_empty = bars + 8;     // [2] S accumulator read twice                    (8 element-wise warps)
  uint64_t* p_ready = bars + 10;    //     P in smem                                   (8 element-wise warps)
  uint64_t* o_full = bars + 11;     // [2] O accumulator written, P consumed           (tcgen05.commit)
  uint64_t* o_empty = bars + 13;    // [2] O accumulator drained                       (8 element-wise warps)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 16);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x;
  WinGeo geo;
  geo.H = p.H; geo.W = p.W; geo.ws = TC_WS; geo.shift = p.shift;
  geo.nWw = p.W / TC_WS; geo.nWh = p.H / TC_WS; geo.nW = geo.nWh * geo.nWw;
  int first, n_my;
  tq_my_windows(p.G * geo.nW, first, n_my);
  const float scale2 = p.scale * WA_LOG2E;

  pdl_trigger();
  if (warp == TC_WARP_MMA) {
    if (lane == 0) {
      for (int s = 0; s < TQF_STAGES; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&stage_free[s], 2);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&s_full[b], 1);
        mbar_init(&s_empty[b], 8);
        mbar_init(&o_full[b], 1);
        mbar_init(&o_empty[b], 8);
      }
      mbar_init(p_ready, 8);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  if (warp == TC_WARP_LD && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  pdl_wait();  // first global access below (the relative-position table)
  fill_tables_quad(T, p.bias_table, p.nH, h, p.shift, tid, TC_THREADS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < 8) {
    tq_fwd_elementwise(p, geo, T, tbl_bytes, sP, rowmax, rowsum, s_full, s_empty, p_ready, o_full, o_empty, tmem_base, first,
                       n_my, h, warp, lane, scale2);
  } else if (warp == TC_WARP_MMA) {
    // ================= tcgen05.mma issuer =================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, TC_N / 2, 0, 0);  // S half = Q K_half^T (both K-major, N = 72)
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, WA_HD, 0, 1);     // O = P V     (V MN-major)
    auto issue_s = [&](int it) {
      const int s = it % TQF_STAGES, b = it & 1;
      tc_wait(&full[s], (it / TQF_STAGES) & 1, 103, it);
      tc_wait(&s_empty[b], ((it >> 1) & 1) ^ 1, 104, it);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t q_addr = smem_u32(smem + s * TQF_STAGE_BYTES), k_addr = q_addr + TC_TILE;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const uint32_t d = tmem_base + b * TQF_S_BUF + hf * TQF_S_HALF;
          const uint32_t kh_addr = k_addr + hf * 72 * 64;  // key rows 72 * hf .. : nine 8-row groups further on
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            umma_f16_ss(d, desc_tile_kmajor(q_addr, ks), desc_tile_kmajor(kh_addr, ks), idesc_s, ks);
        }
        umma_commit(&s_full[b]);
      }
      __syncwarp();
    };
    issue_s(0);
#pragma unroll 1
    for (int it = 0; it < n_my; ++it) {
      if (it + 1 < n_my) issue_s(it + 1);  // next window's scores overlap this window's softmax
      const int b = it & 1;
      tc_wait(p_ready, it & 1, 105, it);
      tc_wait(&o_empty[b], ((it >> 1) & 1) ^ 1, 108, it);
      tc_fence_after();
      if (lane == 0) {
        const int s = it % TQF_STAGES;
        const uint32_t v_addr = smem_u32(smem + s * TQF_STAGE_BYTES + 2 * TC_TILE), p_addr = smem_u32(sP);
        const uint32_t d = tmem_base + (b ? TQF_O_COL1 : TQF_O_COL0);
#pragma unroll
        for (int kk = 0; kk < 9; ++kk)  // 16 keys per step
          umma_f16_ss(d, desc_pds_kmajor(p_addr, kk, TF_PCHUNK), desc_tile_mnmajor(v_addr, kk), idesc_o, kk);
        umma_commit(&o_full[b]);
        umma_commit(&stage_free[s]);
      }
      __syncwarp();
    }
  } else if (warp == TC_WARP_LD) {
    // ================= TMA producer: 12 boxes (Q, K, V x 4 quadrants) per window =================
    if (lane == 0) {
      const int c0 = h * WA_HD;
      TqWindowIter wi;
      wi.init(first, geo);
#pragma unroll 1
      for (int it = 0; it < n_my; ++it) {
        const int s = it % TQF_STAGES;
        if (it >= TQF_STAGES) tc_wait(&stage_free[s], (it / TQF_STAGES - 1) & 1, 106, it);
        const uint32_t st = smem_u32(smem + s * TQF_STAGE_BYTES);
        mbar_arrive_expect_tx(&full[s], TQF_STAGE_BYTES);
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          int hq = wi.wh * TC_WS + p.shift + 6 * (qd >> 1), wq = wi.ww * TC_WS + p.shift + 6 * (qd & 1);
          hq -= hq >= p.H ? p.H : 0;
          wq -= wq >= p.W ? p.W : 0;
          const uint32_t dst = st + qd * 36 * 64;
          tma_load_4d(dst, &tmQ, &full[s], c0, wq, hq, wi.img);
          tma_load_4d(dst + TC_TILE, &tmK, &full[s], c0, wq, hq, wi.img);
          tma_load_4d(dst + 2 * TC_TILE, &tmV, &full[s], c0, wq, hq, wi.img);
        }
        wi.next(geo);
      }
    }
  } else {
    // ================= remainder rows 128..143 (mma.sync), windows it = k, k + 2, ... =================
    const int k = warp - TC_WARP_R0;
    TqWindowIter wi;
    wi.init(first, geo);
    if (k) wi.next(geo);
#pragma unroll 1
    for (int it = k; it < n_my; it += 2) {
      const int s = it % TQF_STAGES;
      const int g = wi.g(geo);
      const int emask = wi.emask(geo);
      tc_wait(&full[s], (it / TQF_STAGES) & 1, 103, it);
      const uint32_t sQ = smem_u32(smem + s * TQF_STAGE_BYTES), sK = sQ + TC_TILE, sV = sK + TC_TILE;
      float m_run[2] = {-1e30f, -1e30f}, l_run[2] = {0.f, 0.f};
      float oacc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;
      // one instantiation for masked and unmasked windows (emask = 0 never adds the mask): half the code
      tc_fwd_rem_tile<true>(sQ, sK, sV, T, tbl_bytes, lane, scale2, emask, oacc, m_run, l_run);
      const int rl0 = 128 + (lane >> 2);
      float inv[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float l = l_run[r];
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        inv[r] = 1.0f / l;
        const int tk = T.tok[rl0 + r * 8];
        if ((lane & 3) == 0 && p.lse)
          p.lse[(static_cast<long long>(g) * p.nH + h) * TC_N + (tk & 255) * TC_WS + (tk >> 8)] = (m_run[r] + lg2_approx(l)) * WA_LN2;
      }
      // Q rows 128..143 are read by this warp only (the UMMA tile is rows 0..127): reuse them as staging
      __syncwarp();
      uint8_t* q_tile = smem + s * TQF_STAGE_BYTES;
#pragma unroll
      for (int dt = 0; dt < 4; ++dt) {
        const int cb = (lane & 3) * 4;  // byte offset inside the 16-byte piece dt
        *reinterpret_cast<uint32_t*>(q_tile + sw64_off(rl0, dt) + cb) = pack_bf16(oacc[dt][0] * inv[0], oacc[dt][1] * inv[0]);
        *reinterpret_cast<uint32_t*>(q_tile + sw64_off(rl0 + 8, dt) + cb) = pack_bf16(oacc[dt][2] * inv[1], oacc[dt][3] * inv[1]);
      }
      __syncwarp();
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int c = lane + 32 * kk;
        const int i = 128 + (c >> 2), cc = c & 3;
        const int tk = T.tok[i];
        const long long grow = wi.row(geo, tk & 255, tk >> 8);
        *reinterpret_cast<uint4*>(p.o + grow * p.ldo + h * WA_HD + cc * 8) =
            *reinterpret_cast<const uint4*>(q_tile + sw64_off(i, cc));
      }
      fence_proxy_async_smem();  // the staging writes above precede the TMA writes of the stage's next window
      __syncwarp();
      if (lane == 0) mbar_arrive(&stage_free[s]);
      wi.next(geo);
      wi.next(geo);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TC_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =================================================================================================
// Fourth generation, backward ("tq")
//
// ncu on the third-generation backward (profiles/r2_winbwd_tc_ncu.txt): 48 % of the stall samples are long-scoreboard
// waits — the per-window global loads of LSE and D in front of the element-wise phase, and mbarrier spins while the
// tensor core works through the output MMAs.  Here:
//   * every input of a window arrives by TMA: Q, dO, K, V, O as 5 x 4 quadrant boxes plus one 576-byte bulk copy of
//     the window-head's LSE row; D = rowsum(dO * O) is computed from the staged tiles (the separate D pre-pass
//     kernel and its scratch tensor are gone: 18 KB less HBM traffic per window-head);
//   * quadrant token order as in the forward (mask addend per key quadrant, folded into the exponent offset);
//   * packed fp32x2 math: per score FFMA2/2 + FADD2/2 + MUFU + FADD2/2 + FMUL2/2 + FADD2/2 + 2 F2FP/2 + LDS;
//   * the tensor core gets S / dP of window i + 1 queued right behind the output MMAs of window i.
// =================================================================================================
// Optional event trace of CTA (0, 0) (option "tq_trace", tools/tq_trace.py): [role][window][event] = clock64()
constexpr int TQT_ROLES = 6, TQT_WINDOWS = 24, TQT_EVENTS = 8;
__device__ __forceinline__ void tq_trace(long long* tr, int role, int it, int ev) {
  if (tr != nullptr && it < TQT_WINDOWS) tr[(role * TQT_WINDOWS + it) * TQT_EVENTS + ev] = clock64();
}

__host__ __device__ constexpr int tq_bj(int j) { return tq_th(j) * TC_TW2 + tq_tw(j); }
__host__ __device__ constexpr int tq_tile_row(int th, int tw) { return 36 * (2 * (th / 6) + tw / 6) + 6 * (th % 6) + tw % 6; }

__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

constexpr int TQB_STAGES = 2;
constexpr int TQB_STAGE_BYTES = 5 * TC_TILE;                 // Q, dO, K, V, O
constexpr int TQB_OFF_P = TQB_STAGES * TQB_STAGE_BYTES;      // 92160
constexpr int TQB_OFF_DS = TQB_OFF_P + 3 * TB_PCHUNK;
constexpr int TQB_THREADS = 416;                             // 8 element-wise warps, MMA issuer, TMA producer, 3 remainder warps
constexpr int TQB_REM = 3;
constexpr int TQB_OFF_STG = TQB_OFF_DS + 3 * TB_PCHUNK;      // [3 remainder warps][16 rows][64 B]
constexpr int TQB_OFF_LSE = TQB_OFF_STG + TQB_REM * 1024;    // [2 stages][144] fp32
constexpr int TQB_OFF_TBL = TQB_OFF_LSE + TQB_STAGES * TC_N * 4;
constexpr int TQB_OFF_BARS = TQB_OFF_TBL + ((TC_TABLE_BYTES + 15) & ~15);
constexpr int TQB_SMEM = 1024 + TQB_OFF_BARS + 16 * 8;
static_assert(TQB_SMEM <= 227 * 1024, "backward shared memory");
static_assert(TQB_OFF_LSE % 16 == 0 && (TC_N * 4) % 16 == 0, "bulk copy alignment");

// D of one tile row from the staged dO / O tiles (64-byte swizzled rows), pieces [pc0, pc0 + npc)
__device__ __forceinline__ float tq_row_dot(const uint8_t* sdO, const uint8_t* sO, int row, int pc0, int npc) {
  float acc = 0.f;
  for (int pc = pc0; pc < pc0 + npc; ++pc) {
    const uint4 a = *reinterpret_cast<const uint4*>(sdO + sw64_off(row, pc));
    const uint4 b = *reinterpret_cast<const uint4*>(sO + sw64_off(row, pc));
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

// TMEM columns of the backward: S and dP are issued as two N = 72 MMAs each (key halves at 32-column aligned bases),
// so both key halves run the same element-wise code (see the forward).
constexpr uint32_t TQB_S_COL = 0, TQB_DP_COL = 192, TQB_HALF = 96, TQB_DV_COL = 384, TQB_DK_COL = 416, TQB_DQ_COL = 448;

// The 72 key columns of one (query row, key half) in nine 8-column pieces (= one 16-byte store each): P, dS, d(bias)
// sums.  The TMEM loads of piece k + 1 are in flight while piece k is computed.  tbl_h = bias-table pointer of the row,
// already moved to the half's first key (B_j of tile column 72 + c is B_j of column c plus 6 * 23).
__device__ __forceinline__ void tq_bwd_row(uint32_t ts, uint32_t tdp, uint8_t* p_row, uint8_t* ds_row, uint32_t xr, int p8b,
                                           const char* tbl_h, uint64_t scale2, uint64_t c0, uint64_t c1, uint64_t negD2,
                                           uint64_t (&dbacc2)[36]) {
  uint32_t s[2][8], dp[2][8];
  tmem_ld8p(ts, s[0]);
  tmem_ld8p(tdp, dp[0]);
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    tmem_ld_wait();
    if (k + 1 < 9) {
      tmem_ld8p(ts + 8 * (k + 1), s[(k + 1) & 1]);
      tmem_ld8p(tdp + 8 * (k + 1), dp[(k + 1) & 1]);
    }
    uint32_t po[4], dso[4];
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      const int kk = 8 * k + e;
      const uint64_t t = pk2(*reinterpret_cast<const float*>(tbl_h - 4 * tq_bj(kk)), *reinterpret_cast<const float*>(tbl_h - 4 * tq_bj(kk + 1)));
      uint64_t x = fma2(pk2u(s[k & 1][e], s[k & 1][e + 1]), scale2, t);
      x = add2(x, kk < 36 ? c0 : c1);
      float x0, x1;
      upk2(x, x0, x1);
      const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
      const uint64_t ds = mul2(pk2(p0, p1), add2(pk2u(dp[k & 1][e], dp[k & 1][e + 1]), negD2));
      dbacc2[kk >> 1] = add2(dbacc2[kk >> 1], ds);
      float d0, d1;
      upk2(ds, d0, d1);
      po[e >> 1] = pack_bf16(p0, p1);
      dso[e >> 1] = pack_bf16(d0, d1);
    }
    const uint32_t p8 = p8b + k;  // 16-byte piece (8 keys) of the 144-key row
    const uint32_t off = (p8 >> 3) * TB_PCHUNK + (((p8 & 7) << 4) ^ xr);
    *reinterpret_cast<uint4*>(p_row + off) = make_uint4(po[0], po[1], po[2], po[3]);
    *reinterpret_cast<uint4*>(ds_row + off) = make_uint4(dso[0], dso[1], dso[2], dso[3]);
  }
}

struct TqBwdBars {
  uint64_t *full, *stage_free, *s_full, *sdp_empty, *pds_ready, *acc_full, *acc_empty, *rem_done;
};

__device__ __forceinline__ void tq_bwd_elementwise(const AttnParams& p, const WinGeo& geo, const WinTables& T, const char* tbl_bytes,
                                                   uint8_t* smem, const TqBwdBars& B, uint32_t tmem_base, int first, int n_my, int h,
                                                   int warp, int lane, float scale2f, float* sAcc, long long* tr) {
  if (warp != 0 || lane != 0) tr = nullptr;  // trace: element-wise warp 0 only
  uint8_t* sP = smem + TQB_OFF_P;
  uint8_t* sdS = smem + TQB_OFF_DS;
  const float* lse_s = reinterpret_cast<const float*>(smem + TQB_OFF_LSE);
  const int q = warp & 3, hf = warp >> 2, row = q * 32 + lane;
  const int tok = T.tok[row];
  const int th_i = tok & 255, tw_i = tok >> 8;
  const int nat_i = th_i * TC_WS + tw_i;
  const char* tbl_h = tbl_bytes + T.aq4[row] - hf * (4 * 6 * TC_TW2);
  const int code_i = T.code[row];
  const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
  const uint32_t ts = lane_addr + TQB_S_COL + hf * TQB_HALF, tdp = lane_addr + TQB_DP_COL + hf * TQB_HALF;
  const uint64_t scale2 = pk2(scale2f, scale2f);
  const int code_c0 = hf, code_c1 = hf | 2;  // region codes of the half's two key quadrants (see the forward)
  uint8_t* p_row = sP + row * 128;
  uint8_t* ds_row = sdS + row * 128;
  const uint32_t xr = static_cast<uint32_t>(row & 7) << 4;
  const int p8b = hf * 9;
  uint64_t dbacc2[36];  // (row, hf * 72 + c), pairs
#pragma unroll
  for (int c = 0; c < 36; ++c) dbacc2[c] = pk2(0.f, 0.f);
  TqWindowIter wi;
  wi.init(first, geo);
#pragma unroll 1
  for (int it = 0; it < n_my; ++it) {
    const int s = it & 1;
    const int emask = wi.emask(geo);
    const long long grow = wi.row(geo, th_i, tw_i);
    const uint8_t* stage = smem + s * TQB_STAGE_BYTES;
    tq_trace(tr, 0, it, 0);
    tc_wait(&B.full[s], (it >> 1) & 1, 204, it);
    tq_trace(tr, 0, it, 1);
    const float D = tq_row_dot(stage + TC_TILE, stage + 4 * TC_TILE, row, 0, 4);
    const float nlse2 = -lse_s[s * TC_N + nat_i] * WA_LOG2E;
    const float madd0 = ((code_i ^ code_c0) & emask) ? WA_MASK2 : 0.f;
    const float madd1 = ((code_i ^ code_c1) & emask) ? WA_MASK2 : 0.f;

    tc_wait(B.s_full, it & 1, 201, it);
    tc_fence_after();
    tq_trace(tr, 0, it, 2);
    // P / dS buffers are free: the tensor core is done with them (this thread waited for acc_full of the previous
    // window in its drain below) and so are the remainder warps' output jobs
    if (it > 0) tc_wait(B.rem_done, (it - 1) & 1, 202, it);
    tq_trace(tr, 0, it, 3);
    tq_bwd_row(ts, tdp, p_row, ds_row, xr, p8b, tbl_h, scale2, pk2(madd0 + nlse2, madd0 + nlse2),
               pk2(madd1 + nlse2, madd1 + nlse2), pk2(-D, -D), dbacc2);
    tc_fence_before();
    if (lane == 0) mbar_arrive(B.sdp_empty);  // S / dP may be overwritten by the next window's MMAs
    fence_proxy_async_smem();                 // P / dS stores -> visible to the tensor core
    __syncwarp();
    if (lane == 0) mbar_arrive(B.pds_ready);
    tq_trace(tr, 0, it, 4);

    // drain dV / dK / dQ of token `row`, head-dim columns [hf * 16, +16)
    tc_wait(B.acc_full, it & 1, 203, it);
    tc_fence_after();
    tq_trace(tr, 0, it, 5);
#pragma unroll
    for (int t = 0; t < 3; ++t) {  // one accumulator at a time: 16 live registers next to the 72 d(bias) sums
      uint32_t a[16];
      tmem_ld16p(lane_addr + (t == 0 ? TQB_DV_COL : (t == 1 ? TQB_DK_COL : TQB_DQ_COL)) + hf * 16, a);
      tmem_ld_wait();
      const float sc = t == 0 ? 1.0f : p.scale;
      bf16* dst = (t == 0 ? p.dv + grow * p.lddv : (t == 1 ? p.dk + grow * p.lddk : p.dq + grow * p.lddq)) + h * WA_HD + hf * 16;
      uint4 v0, v1;
      v0.x = pack_bf16(__uint_as_float(a[0]) * sc, __uint_as_float(a[1]) * sc);
      v0.y = pack_bf16(__uint_as_float(a[2]) * sc, __uint_as_float(a[3]) * sc);
      v0.z = pack_bf16(__uint_as_float(a[4]) * sc, __uint_as_float(a[5]) * sc);
      v0.w = pack_bf16(__uint_as_float(a[6]) * sc, __uint_as_float(a[7]) * sc);
      v1.x = pack_bf16(__uint_as_float(a[8]) * sc, __uint_as_float(a[9]) * sc);
      v1.y = pack_bf16(__uint_as_float(a[10]) * sc, __uint_as_float(a[11]) * sc);
      v1.z = pack_bf16(__uint_as_float(a[12]) * sc, __uint_as_float(a[13]) * sc);
      v1.w = pack_bf16(__uint_as_float(a[14]) * sc, __uint_as_float(a[15]) * sc);
      st_global_256(dst, v0, v1);
    }
    tc_fence_before();
    if (lane == 0) mbar_arrive(B.acc_empty);
    tq_trace(tr, 0, it, 6);
    wi.next(geo);
  }
  named_bar_sync(6, TQB_THREADS);
#pragma unroll
  for (int c = 0; c < 36; ++c) {
    float a, b;
    upk2(dbacc2[c], a, b);
    *reinterpret_cast<float2*>(sAcc + row * TB_ACCP + hf * 72 + 2 * c) = make_float2(a, b);
  }
}

__global__ void __launch_bounds__(TQB_THREADS, 1) win_attn_tq_bwd_kernel(const AttnParams p, const __grid_constant__ CUtensorMap tmQ,
                                                                       const __grid_constant__ CUtensorMap tmdO,
                                                                       const __grid_constant__ CUtensorMap tmK,
                                                                       const __grid_constant__ CUtensorMap tmV,
                                                                       const __grid_constant__ CUtensorMap tmO, long long* trace) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sP = smem + TQB_OFF_P;
  uint8_t* sdS = smem + TQB_OFF_DS;
  float* lse_s = reinterpret_cast<float*>(smem + TQB_OFF_LSE);
  WinTables T;
  T.tbl2 = reinterpret_cast<float*>(smem + TQB_OFF_TBL);
  T.aq4 = reinterpret_cast<int*>(T.tbl2 + 2 * WA_MAXTBL + 2);
  T.bj4 = T.aq4 + WA_ROWS;
  T.code = T.bj4 + WA_ROWS;
  T.tok = T.code + WA_ROWS;
  const char* tbl_bytes = reinterpret_cast<const char*>(T.tbl2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TQB_OFF_BARS);
  // timeout tags (tc_wait): 201 s_full, 202 rem_done, 203 acc_full, 204 full, 205 sdp_empty, 206 pds_ready,
  //                         207 acc_empty, 208 stage_free
  TqBwdBars B;
  B.full = bars;            // [2] TMA boxes + LSE row of a stage have landed              (expect_tx)
  B.stage_free = bars + 2;  // [2] stage may be overwritten                                 (commit + 2 remainder warps)
  B.s_full = bars + 4;      //     S and dP accumulators written                            (commit)
  B.sdp_empty = bars + 5;   //     S and dP read into registers                             (8 element-wise warps)
  B.pds_ready = bars + 6;   //     P and dS complete in smem                                (8 + 2 warps)
  B.acc_full = bars + 7;    //     dV / dK / dQ accumulators written, P / dS consumed        (commit)
  B.acc_empty = bars + 8;   //     accumulators drained                                     (8 element-wise warps)
  B.rem_done = bars + 9;    //     remainder warps have read P / dS of the window           (2 remainder warps)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 10);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x;
  WinGeo geo;
  geo.H = p.H; geo.W = p.W; geo.ws = TC_WS; geo.shift = p.shift;
  geo.nWw = p.W / TC_WS; geo.nWh = p.H / TC_WS; geo.nW = geo.nWh * geo.nWw;
  int first, n_my;
  tq_my_windows(p.G * geo.nW, first, n_my);
  const float scale2 = p.scale * WA_LOG2E;

  pdl_trigger();
  if (warp == TC_WARP_MMA) {
    if (lane == 0) {
      for (int s = 0; s < TQB_STAGES; ++s) {
        mbar_init(&B.full[s], 1);
        mbar_init(&B.stage_free[s], 1 + TQB_REM);
      }
      mbar_init(B.s_full, 1);
      mbar_init(B.sdp_empty, 8);
      mbar_init(B.pds_ready, 8 + TQB_REM);
      mbar_init(B.acc_full, 1);
      mbar_init(B.acc_empty, 8);
      mbar_init(B.rem_done, TQB_REM);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  if (warp == TC_WARP_LD && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
  }
  pdl_wait();  // first global access below (the relative-position table)
  fill_tables_quad(T, p.bias_table, p.nH, h, p.shift, tid, TQB_THREADS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  float* sAcc = reinterpret_cast<float*>(sP);  // d(bias) flush matrix, aliased onto the P / dS chunks after the last window

  if (warp < 8) {
    tq_bwd_elementwise(p, geo, T, tbl_bytes, smem, B, tmem_base, first, n_my, h, warp, lane, scale2, sAcc,
                       (blockIdx.x | blockIdx.y) == 0 ? trace : nullptr);
  } else if (warp == TC_WARP_MMA) {
    // ================= tcgen05.mma issuer =================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, TC_N / 2, 0, 0);  // S, dP halves (K-major x K-major, N = 72)
    constexpr uint32_t idesc_kv = umma_idesc_bf16(128, WA_HD, 1, 1);  // dV = P^T dO, dK = dS^T Q (MN x MN)
    constexpr uint32_t idesc_q = umma_idesc_bf16(128, WA_HD, 0, 1);   // dQ = dS K (K-major x MN-major)
    const uint32_t p_addr = smem_u32(sP), ds_addr = smem_u32(sdS);
    long long* tr = ((blockIdx.x | blockIdx.y) == 0 && lane == 0) ? trace : nullptr;
    auto issue_scores = [&](int it) {
      const int s = it & 1;
      const uint32_t q_addr = smem_u32(smem + s * TQB_STAGE_BYTES);
      const uint32_t do_addr = q_addr + TC_TILE, k_addr = q_addr + 2 * TC_TILE, v_addr = q_addr + 3 * TC_TILE;
      tc_wait(&B.full[s], (it >> 1) & 1, 204, it);
      tq_trace(tr, 1, it, 0);
      tc_wait(B.sdp_empty, (it & 1) ^ 1, 205, it);
      tc_fence_after();
      tq_trace(tr, 1, it, 1);
      if (lane == 0) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {  // key rows 72 * hf ..: nine 8-row groups further on
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            umma_f16_ss(tmem_base + TQB_S_COL + hf * TQB_HALF, desc_tile_kmajor(q_addr, ks),
                        desc_tile_kmajor(k_addr + hf * 72 * 64, ks), idesc_s, ks);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            umma_f16_ss(tmem_base + TQB_DP_COL + hf * TQB_HALF, desc_tile_kmajor(do_addr, ks),
                        desc_tile_kmajor(v_addr + hf * 72 * 64, ks), idesc_s, ks);
        }
        umma_commit(B.s_full);
      }
      __syncwarp();
    };
    issue_scores(0);
#pragma unroll 1
    for (int it = 0; it < n_my; ++it) {
      const int s = it & 1;
      const uint32_t q_addr = smem_u32(smem + s * TQB_STAGE_BYTES);
      const uint32_t do_addr = q_addr + TC_TILE, k_addr = q_addr + 2 * TC_TILE;
      tc_wait(B.pds_ready, it & 1, 206, it);
      tq_trace(tr, 1, it, 2);
      tc_wait(B.acc_empty, (it & 1) ^ 1, 207, it);
      tc_fence_after();
      tq_trace(tr, 1, it, 3);
      if (lane == 0) {
#pragma unroll
        for (int kk = 0; kk < 9; ++kk) {  // 16 queries per step
          umma_f16_ss(tmem_base + TQB_DV_COL, desc_pds_mnmajor(p_addr, kk), desc_tile_mnmajor(do_addr, kk), idesc_kv, kk);
          umma_f16_ss(tmem_base + TQB_DK_COL, desc_pds_mnmajor(ds_addr, kk), desc_tile_mnmajor(q_addr, kk), idesc_kv, kk);
        }
#pragma unroll
        for (int kk = 0; kk < 9; ++kk)  // 16 keys per step
          umma_f16_ss(tmem_base + TQB_DQ_COL, desc_pds_kmajor(ds_addr, kk, TB_PCHUNK), desc_tile_mnmajor(k_addr, kk),
                      idesc_q, kk);
        umma_commit(B.acc_full);
        umma_commit(&B.stage_free[s]);
      }
      __syncwarp();
      if (it + 1 < n_my) issue_scores(it + 1);  // queued right behind the output MMAs (S / dP were released before pds_ready)
    }
    named_bar_sync(6, TQB_THREADS);
  } else if (warp == TC_WARP_LD) {
    // ================= TMA producer: 20 boxes (Q, dO, K, V, O x 4 quadrants) + the LSE row per window =================
    if (lane == 0) {
      const int c0 = h * WA_HD;
      TqWindowIter wi;
      wi.init(first, geo);
#pragma unroll 1
      for (int it = 0; it < n_my; ++it) {
        const int s = it & 1;
        tq_trace((blockIdx.x | blockIdx.y) == 0 ? trace : nullptr, 2, it, 0);
        if (it >= TQB_STAGES) tc_wait(&B.stage_free[s], ((it >> 1) - 1) & 1, 208, it);
        tq_trace((blockIdx.x | blockIdx.y) == 0 ? trace : nullptr, 2, it, 1);
        const int g = wi.g(geo), bimg = wi.img, wh = wi.wh, ww = wi.ww;
        const uint32_t st = smem_u32(smem + s * TQB_STAGE_BYTES);
        mbar_arrive_expect_tx(&B.full[s], TQB_STAGE_BYTES + TC_N * 4);
        bulk_load_1d(smem_u32(lse_s + s * TC_N), p.lse + (static_cast<long long>(g) * p.nH + h) * TC_N, TC_N * 4, &B.full[s]);
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          int hq = wh * TC_WS + p.shift + 6 * (qd >> 1), wq = ww * TC_WS + p.shift + 6 * (qd & 1);
          hq -= hq >= p.H ? p.H : 0;
          wq -= wq >= p.W ? p.W : 0;
          const uint32_t dst = st + qd * 36 * 64;
          tma_load_4d(dst, &tmQ, &B.full[s], c0, wq, hq, bimg);
          tma_load_4d(dst + 2 * TC_TILE, &tmK, &B.full[s], c0, wq, hq, bimg);
          tma_load_4d(dst + TC_TILE, &tmdO, &B.full[s], c0, wq, hq, bimg);
          tma_load_4d(dst + 3 * TC_TILE, &tmV, &B.full[s], c0, wq, hq, bimg);
          tma_load_4d(dst + 4 * TC_TILE, &tmO, &B.full[s], c0, wq, hq, bimg);
        }
        wi.next(geo);
      }
    }
    named_bar_sync(6, TQB_THREADS);
  } else {
    // ================= remainder warps (mma.sync) =================
    // warp k of three: the score job of key third k (16 queries 128..143 x 48 keys), then the output job of token rows
    // 128..143 of type k (0: dV = P^T dO, 1: dK = dS^T Q, 2: dQ = dS K).  Each job is a latency-bound chain on one warp
    // (~2200 / ~3000 cycles, tools/tq_trace.py); with two warps sharing the six jobs they were the critical path.
    const int k = warp - TC_WARP_R0;
    float dbr[6][4];  // [n-tile][fragment element], as in window_attn.cu
#pragma unroll
    for (int i = 0; i < 6; ++i) dbr[i][0] = dbr[i][1] = dbr[i][2] = dbr[i][3] = 0.f;
    uint8_t* stg = smem + TQB_OFF_STG + k * 1024;
    const int rl0 = 128 + (lane >> 2);
    const int tok0 = T.tok[rl0], tok1 = T.tok[rl0 + 8];
    const int nat0 = (tok0 & 255) * TC_WS + (tok0 >> 8), nat1 = (tok1 & 255) * TC_WS + (tok1 >> 8);
    const float inv_scale = 1.0f / p.scale;
    TqWindowIter wi;
    wi.init(first, geo);
#pragma unroll 1
    for (int it = 0; it < n_my; ++it) {
      const int s = it & 1;
      const int emask = wi.emask(geo);
      const uint8_t* stage = smem + s * TQB_STAGE_BYTES;
      const uint32_t sQ = smem_u32(stage);
      const uint32_t sdO = sQ + TC_TILE, sK = sQ + 2 * TC_TILE, sV = sQ + 3 * TC_TILE;

      long long* tr = ((blockIdx.x | blockIdx.y) == 0 && lane == 0) ? trace : nullptr;
      tq_trace(tr, 3 + k, it, 0);
      tc_wait(&B.full[s], (it >> 1) & 1, 204, it);
      tq_trace(tr, 3 + k, it, 1);
      const float nl0 = -lse_s[s * TC_N + nat0] * inv_scale, nl1 = -lse_s[s * TC_N + nat1] * inv_scale;
      // D of rows rl0 / rl0 + 8: the four lanes that share a row take one 16-byte piece each
      float d0 = tq_row_dot(stage + TC_TILE, stage + 4 * TC_TILE, rl0, lane & 3, 1);
      float d1 = tq_row_dot(stage + TC_TILE, stage + 4 * TC_TILE, rl0 + 8, lane & 3, 1);
      d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
      d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
      d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
      d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
      if (it > 0) {
        tc_wait(B.acc_full, (it - 1) & 1, 203, it);  // the tensor core has consumed P / dS of the previous window
        named_bar_sync(5, 32 * TQB_REM);             // ... and so have the other remainder warps
      }
      tq_trace(tr, 3 + k, it, 2);
      // one instantiation for masked and unmasked windows (emask = 0 never adds the mask): half the code
      tc_bwd_rem_scores<true>(sQ, sdO, sK, sV, sP, sdS, nl0, nl1, -d0, -d1, T, tbl_bytes, k, lane, scale2, emask, dbr);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(B.pds_ready);
      tq_trace(tr, 3 + k, it, 3);

      tc_wait(B.pds_ready, it & 1, 206, it);  // every P / dS element of this window is in smem
      tq_trace(tr, 3 + k, it, 4);
      {
        const int type = k;
        float acc[4][4];
        tc_bwd_rem_out(type, sQ, sdO, sK, smem_u32(sP), smem_u32(sdS), lane, acc);
        const float sc = type == 0 ? 1.0f : p.scale;
        const int r_lo = lane >> 2;
#pragma unroll
        for (int dt = 0; dt < 4; ++dt) {
          const int cb = (lane & 3) * 4;
          *reinterpret_cast<uint32_t*>(stg + sw64_off(r_lo, dt) + cb) = pack_bf16(acc[dt][0] * sc, acc[dt][1] * sc);
          *reinterpret_cast<uint32_t*>(stg + sw64_off(r_lo + 8, dt) + cb) = pack_bf16(acc[dt][2] * sc, acc[dt][3] * sc);
        }
        __syncwarp();
        bf16* outp = type == 0 ? p.dv : (type == 1 ? p.dk : p.dq);
        const long long ldo = type == 0 ? p.lddv : (type == 1 ? p.lddk : p.lddq);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const int r = (lane >> 2) + 8 * kk, cc = lane & 3;
          const int tk = T.tok[128 + r];
          const long long grow = wi.row(geo, tk & 255, tk >> 8);
          *reinterpret_cast<uint4*>(outp + grow * ldo + h * WA_HD + cc * 8) =
              *reinterpret_cast<const uint4*>(stg + sw64_off(r, cc));
        }
        __syncwarp();  // staging tile is free again
      }
      if (lane == 0) {
        mbar_arrive(&B.stage_free[s]);
        mbar_arrive(B.rem_done);
      }
      tq_trace(tr, 3 + k, it, 5);
      wi.next(geo);
    }
    named_bar_sync(6, TQB_THREADS);
    const int qi = 128 + (lane >> 2);
#pragma unroll
    for (int nt = 0; nt < 6; ++nt) {
      const int j0 = k * 48 + nt * 8 + (lane & 3) * 2;
      *reinterpret_cast<float2*>(sAcc + qi * TB_ACCP + j0) = make_float2(dbr[nt][0], dbr[nt][1]);
      *reinterpret_cast<float2*>(sAcc + (qi + 8) * TB_ACCP + j0) = make_float2(dbr[nt][2], dbr[nt][3]);
    }
  }

  // one global atomic per table entry (dh, dw): the sum over the <= 144 token pairs with that offset (sAcc is indexed by
  // TILE rows / columns, i.e. quadrant order)
  tc_fence_before();
  __syncthreads();
  for (int t = tid; t < TC_TW2 * TC_TW2; t += TQB_THREADS) {
    const int dh = t / TC_TW2 - (TC_WS - 1), dw = t % TC_TW2 - (TC_WS - 1);
    const int ih0 = max(0, dh), ih1 = min(TC_WS, TC_WS + dh), iw0 = max(0, dw), iw1 = min(TC_WS, TC_WS + dw);
    float sum = 0.f;
    for (int ih = ih0; ih < ih1; ++ih)
      for (int iw = iw0; iw < iw1; ++iw)
        sum += sAcc[tq_tile_row(ih, iw) * TB_ACCP + tq_tile_row(ih - dh, iw - dw)];
    atomicAdd(&p.dbias_table[t * p.nH + h], sum);
  }
  if (warp == TC_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static bool aligned16(const void* ptr, long long ld) {
  return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ld % 8 == 0;
}

// 12x12 windows, head_dim 32, cyclic shift 0 or 6, 16-byte addressable rows
bool win_attn_tc_supported(const AttnParams& p, int hd, bool bwd) {
  if (!(p.mode == 1 && hd == WA_HD && p.ws == TC_WS && p.Lq == TC_N && (p.shift == 0 || p.shift == TC_WS / 2) &&
        p.drop_p == 0.f && p.H % TC_WS == 0 && p.W % TC_WS == 0))
    return false;
  if (!(aligned16(p.q, p.ldq) && aligned16(p.k, p.ldk) && aligned16(p.v, p.ldv) && aligned16(p.o, p.ldo))) return false;
  if (bwd && !(aligned16(p.d_o, p.lddo) && aligned16(p.dq, p.lddq) && aligned16(p.dk, p.lddk) && aligned16(p.dv, p.lddv)))
    return false;
  return true;
}

static int tc_grid_y(const AttnParams& p) {
  const int n_groups = p.G * (p.H / TC_WS) * (p.W / TC_WS);
  int gy = num_sms() / p.nH;  // one CTA per SM (single wave), persistent over the windows of its head
  if (gy < 1) gy = 1;
  if (gy > n_groups) gy = n_groups;
  return gy;
}

// ---- TMA tensor maps of the fourth generation: [G, H, W, C] bf16 activation, box = 32 channels x 6 x 6 tokens ----
typedef CUresult (*TcEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TcEncodeTiledFn tc_encode_fn() {
  static TcEncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<TcEncodeTiledFn>(sym);
  });
  return fn;
}

struct WinMapKey {
  const void* base;
  long long ld;
  int G, H, W, C;
  bool operator==(const WinMapKey& o) const { return base == o.base && ld == o.ld && G == o.G && H == o.H && W == o.W && C == o.C; }
};
// Encoding a map costs ~1 us of host time and the same few (pointer, shape) pairs come back every step: tiny cache.
static int win_tmap(CUtensorMap* out, const bf16* base, long long ld, int G, int H, int W, int C) {
  constexpr int SLOTS = 64;
  static WinMapKey keys[SLOTS];
  static CUtensorMap maps[SLOTS];
  static int used = 0, next = 0;
  static std::mutex mu;
  const WinMapKey key{base, ld, G, H, W, C};
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < used; ++i)
    if (keys[i] == key) {
      *out = maps[i];
      return 0;
    }
  TcEncodeTiledFn fn = tc_encode_fn();
  FIBER_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                              static_cast<cuuint64_t>(G)};
  const cuuint64_t strides[3] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(W) * ld * 2,
                                 static_cast<cuuint64_t>(H) * W * ld * 2};
  const cuuint32_t box[4] = {WA_HD, 6, 6, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FIBER_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (window tile) failed with %d", static_cast<int>(r));
  const int slot = used < SLOTS ? used++ : (next = (next + 1) % SLOTS);
  keys[slot] = key;
  maps[slot] = *out;
  return 0;
}

int option_winattn_tc();  // capi.cu

static int launch_win_tq_fwd(const AttnParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    FIBER_CUDA(cudaFuncSetAttribute(win_attn_tq_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TQF_SMEM));
    attr_set = true;
  }
  CUtensorMap tq, tk, tv;
  const int C = p.nH * WA_HD;
  if (win_tmap(&tq, p.q, p.ldq, p.G, p.H, p.W, C) || win_tmap(&tk, p.k, p.ldk, p.G, p.H, p.W, C) ||
      win_tmap(&tv, p.v, p.ldv, p.G, p.H, p.W, C))
    return -1;
  FIBER_CUDA(launch_k(win_attn_tq_fwd_kernel, dim3(p.nH, tc_grid_y(p)), dim3(TC_THREADS), TQF_SMEM, stream, p, tq, tk, tv));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  count_winattn_tc_launch();
  return 0;
}

static bool aligned32(const void* ptr, long long ld) {
  return (reinterpret_cast<uintptr_t>(ptr) & 31) == 0 && ld % 16 == 0;
}

int launch_win_tc_fwd(const AttnParams& p, cudaStream_t stream) {
  if ((option_winattn_tc() & 4) && aligned32(p.o, p.ldo)) return launch_win_tq_fwd(p, stream);  // fourth generation
  static bool attr_set = false;
  if (!attr_set) {
    FIBER_CUDA(cudaFuncSetAttribute(win_attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TF_SMEM));
    attr_set = true;
  }
  FIBER_CUDA(launch_k(win_attn_tc_fwd_kernel, dim3(p.nH, tc_grid_y(p)), dim3(TC_THREADS), TF_SMEM, stream, p));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  count_winattn_tc_launch();
  return 0;
}

static long long* g_tq_trace_buf = nullptr;  // device buffer of the optional event trace (option "tq_trace")
int option_tq_trace();                       // capi.cu
extern "C" int fiber_debug_tq_trace(long long* host_dst, int n) {  // debug tool entry (tools/tq_trace.py), not part of the ABI
  const int total = TQT_ROLES * TQT_WINDOWS * TQT_EVENTS;
  if (g_tq_trace_buf == nullptr || n < total) return -1;
  return cudaMemcpy(host_dst, g_tq_trace_buf, total * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? total : -2;
}

static int launch_win_tq_bwd(const AttnParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  long long* trace = nullptr;
  if (option_tq_trace()) {
    if (g_tq_trace_buf == nullptr) FIBER_CUDA(cudaMalloc(&g_tq_trace_buf, TQT_ROLES * TQT_WINDOWS * TQT_EVENTS * sizeof(long long)));
    FIBER_CUDA(cudaMemsetAsync(g_tq_trace_buf, 0, TQT_ROLES * TQT_WINDOWS * TQT_EVENTS * sizeof(long long), stream));
    trace = g_tq_trace_buf;
  }
  if (!attr_set) {
    FIBER_CUDA(cudaFuncSetAttribute(win_attn_tq_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TQB_SMEM));
    attr_set = true;
  }
  CUtensorMap tq, tdo, tk, tv, to;
  const int C = p.nH * WA_HD;
  if (win_tmap(&tq, p.q, p.ldq, p.G, p.H, p.W, C) || win_tmap(&tdo, p.d_o, p.lddo, p.G, p.H, p.W, C) ||
      win_tmap(&tk, p.k, p.ldk, p.G, p.H, p.W, C) || win_tmap(&tv, p.v, p.ldv, p.G, p.H, p.W, C) ||
      win_tmap(&to, p.o, p.ldo, p.G, p.H, p.W, C))
    return -1;
  FIBER_CUDA(launch_k(win_attn_tq_bwd_kernel, dim3(p.nH, tc_grid_y(p)), dim3(TQB_THREADS), TQB_SMEM, stream, p, tq, tdo, tk, tv, to, trace));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  count_winattn_tc_launch();
  return 0;
}

int launch_win_tc_bwd(const AttnParams& p, float* D, cudaStream_t stream) {
  if ((option_winattn_tc() & 8) && (reinterpret_cast<uintptr_t>(p.lse) & 15) == 0 && aligned32(p.dq, p.lddq) &&
      aligned32(p.dk, p.lddk) && aligned32(p.dv, p.lddv))
    return launch_win_tq_bwd(p, stream);  // fourth generation: no D pre-pass
  if (launch_win_bwd_prep(p, D, stream)) return -2;
  static bool attr_set = false;
  if (!attr_set) {
    FIBER_CUDA(cudaFuncSetAttribute(win_attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_SMEM));
    attr_set = true;
  }
  FIBER_CUDA(launch_k(win_attn_tc_bwd_kernel, dim3(p.nH, tc_grid_y(p)), dim3(TC_THREADS), TB_SMEM, stream, p, D));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  count_winattn_tc_launch();
  return 0;
}

}  // namespace fiber
