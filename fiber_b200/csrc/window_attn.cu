// fiber_b200 — Swin window attention (W-MSA / SW-MSA, swin_transformer.py:195-224 + :363-387),
// specialised for ws*ws <= 144 tokens and head_dim 32 (every FIBER stage at 224 / 384 px).
//
// Both kernels are persistent over the windows of ONE head (blockIdx.x = head): the Q/K/V (+dO) tiles
// of window i+1 are prefetched with cp.async into the second shared-memory buffer while window i is
// computed.  Window order never exists in HBM — rows are gathered / scattered through the closed-form
// cyclic-shift + partition map.  All per-element work happens in the log2 domain:
//     score2 = (q.k) * scale*log2e + table2[A_i - B_j]  (+ -100*log2e if the pair is masked)
// with table2 = log2e * relative_position_bias_table[:, head] staged in shared memory, followed by a
// second copy of -1e30 entries that padded keys (ragged windows, e.g. 7x7 = 49 tokens) index into.
// The SW-MSA mask needs no per-window table: two tokens of a border window differ in region id iff
// their static codes  [th >= ws-shift] | [tw >= ws-shift] << 1  differ in a bit that the window
// enables (bit 0: last window row, bit 1: last window column).
//
// Tensor path: legacy mma.sync m16n8k16 bf16 (at head_dim 32 the core is bound by the per-score
// ALU/MUFU work: 128 MMA flop per exp).
#include "window_common.cuh"
#include "../../include/fiber_b200.h"

namespace fiber {

void count_launch(int n = 1);

namespace {

// =================================================================================================
// Forward: 9 warps, one 16-row query tile each, online softmax over three 48-key register tiles.
// Two CTAs per SM; ONE __syncthreads per window.
// =================================================================================================
constexpr int WF_THREADS = 288;

template <bool MASKED>
__device__ __forceinline__ void wf_tile(const bf16* sQ, const bf16* sK, const bf16* sV, const WinTables& T,
                                        const char* tbl_bytes, int warp, int lane, int n_sub, float scale2,
                                        int emask, float (&oacc)[4][4], float (&m_run)[2], float (&l_run)[2]) {
  constexpr int PITCH = WA_PITCH;
  uint32_t qf[2][4];
  {
    const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
    const int col = (lane >> 4) * 8;
    ldsm_x4(smem_u32(sQ + row * PITCH + col), qf[0][0], qf[0][1], qf[0][2], qf[0][3]);
    ldsm_x4(smem_u32(sQ + row * PITCH + col + 16), qf[1][0], qf[1][1], qf[1][2], qf[1][3]);
  }
  const int rl0 = warp * 16 + (lane >> 2);
  const int c2 = (lane & 3) * 2;
  const int aq0 = T.aq4[rl0], aq1 = T.aq4[rl0 + 8];
  int ci0 = 0, ci1 = 0;
  if (MASKED) {
    ci0 = T.code[rl0] & emask;
    ci1 = T.code[rl0 + 8] & emask;
  }
#pragma unroll 1
  for (int sub = 0; sub < n_sub; ++sub) {
    float s[6][4];
#pragma unroll
    for (int i = 0; i < 6; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
      for (int nt2 = 0; nt2 < 3; ++nt2) {
        const int row = sub * 48 + nt2 * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int col = ks * 16 + ((lane >> 3) & 1) * 8;
        uint32_t b0, b1, b2, b3;
        ldsm_x4(smem_u32(sK + row * PITCH + col), b0, b1, b2, b3);
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
        const int col = dt2 * 16 + (lane >> 4) * 8;
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(smem_u32(sV + row * PITCH + col), b0, b1, b2, b3);
        mma16816(oacc[2 * dt2], a, b0, b1);
        mma16816(oacc[2 * dt2 + 1], a, b2, b3);
      }
    }
  }
}

__global__ void __launch_bounds__(WF_THREADS, 2) win_attn_fwd_kernel(const AttnParams p) {
  pdl_trigger();
  pdl_wait();
  constexpr int PITCH = WA_PITCH;
  extern __shared__ __align__(16) uint8_t smem[];
  bf16* tiles = reinterpret_cast<bf16*>(smem);  // [2 buffers][Q, K, V][144][PITCH]
  WinTables T;
  T.tbl2 = reinterpret_cast<float*>(tiles + 2 * 3 * WA_TILE);
  T.aq4 = reinterpret_cast<int*>(T.tbl2 + 2 * WA_MAXTBL + 2);
  T.bj4 = T.aq4 + WA_ROWS;
  T.code = T.bj4 + WA_ROWS;
  T.tok = T.code + WA_ROWS;
  const char* tbl_bytes = reinterpret_cast<const char*>(T.tbl2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x;
  const int N = p.Lq;
  WinGeo geo;
  geo.H = p.H; geo.W = p.W; geo.ws = p.ws; geo.shift = p.shift;
  geo.nWw = p.W / p.ws; geo.nWh = p.H / p.ws; geo.nW = geo.nWh * geo.nWw;
  const int n_groups = p.G * geo.nW;
  const int n_tiles = (N + 15) / 16;
  const int n_sub = (N + 47) / 48;
  const float scale2 = p.scale * WA_LOG2E;

  fill_tables(T, p.bias_table, p.nH, h, p.ws, p.shift, N, tid, WF_THREADS);

  // prefetch assignment: 576 16-byte chunks per tile = exactly two per thread (rows r0 and r0 + 72)
  const int pr0 = tid >> 2, pcc = tid & 3;
  const int pth0 = pr0 / p.ws, ptw0 = pr0 % p.ws;
  const int pth1 = (pr0 + 72) / p.ws, ptw1 = (pr0 + 72) % p.ws;
  const int col0 = h * WA_HD + pcc * 8;

  auto prefetch = [&](int g, int buf) {
    long long img_base; int h0, w0, em;
    geo.decode(g, img_base, h0, w0, em);
    bf16* tb = tiles + buf * 3 * WA_TILE;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int r = pr0 + 72 * k;
      const bool valid = r < N;
      const long long grow = valid ? geo.row(img_base, h0, w0, k ? pth1 : pth0, k ? ptw1 : ptw0) : 0;
      const uint32_t so = smem_u32(tb + r * PITCH + pcc * 8);
      cp_async16(so, p.q + grow * p.ldq + col0, valid);
      cp_async16(so + WA_TILE * 2, p.k + grow * p.ldk + col0, valid);
      cp_async16(so + 2 * WA_TILE * 2, p.v + grow * p.ldv + col0, valid);
    }
    cp_async_commit();
  };

  int it = 0;
  if (static_cast<int>(blockIdx.y) < n_groups) prefetch(blockIdx.y, 0);
  for (int g = blockIdx.y; g < n_groups; g += gridDim.y, ++it) {
    const int cur = it & 1;
    cp_async_wait0();
    __syncthreads();  // tiles[cur] (and, first time, the tables) visible; everyone is done with tiles[cur^1]
    const int g_next = g + gridDim.y;
    if (g_next < n_groups) prefetch(g_next, cur ^ 1);
    if (warp >= n_tiles) continue;
    bf16* sQ = tiles + cur * 3 * WA_TILE;
    const bf16* sK = sQ + WA_TILE;
    const bf16* sV = sK + WA_TILE;
    long long img_base; int h0, w0, emask;
    geo.decode(g, img_base, h0, w0, emask);

    float m_run[2] = {-1e30f, -1e30f}, l_run[2] = {0.f, 0.f};
    float oacc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;
    if (emask)
      wf_tile<true>(sQ, sK, sV, T, tbl_bytes, warp, lane, n_sub, scale2, emask, oacc, m_run, l_run);
    else
      wf_tile<false>(sQ, sK, sV, T, tbl_bytes, warp, lane, n_sub, scale2, 0, oacc, m_run, l_run);

    const int rl0 = warp * 16 + (lane >> 2);
    float inv[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float l = l_run[r];
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      l += __shfl_xor_sync(0xffffffffu, l, 2);
      inv[r] = 1.0f / l;
      const int qi = rl0 + r * 8;
      if ((lane & 3) == 0 && qi < N && p.lse)
        p.lse[(static_cast<long long>(g) * p.nH + h) * N + qi] = (m_run[r] + lg2_approx(l)) * WA_LN2;
    }
    __syncwarp();  // this warp's Q rows are consumed (fragments in registers): reuse them as staging
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) {
      const int col = dt * 8 + (lane & 3) * 2;
      *reinterpret_cast<uint32_t*>(sQ + rl0 * PITCH + col) = pack_bf16(oacc[dt][0] * inv[0], oacc[dt][1] * inv[0]);
      *reinterpret_cast<uint32_t*>(sQ + (rl0 + 8) * PITCH + col) = pack_bf16(oacc[dt][2] * inv[1], oacc[dt][3] * inv[1]);
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int c = lane + 32 * k;
      const int r = c >> 2, cc = c & 3, i = warp * 16 + r;
      if (i < N) {
        const int tok = T.tok[i];
        const long long grow = geo.row(img_base, h0, w0, tok & 255, tok >> 8);
        *reinterpret_cast<uint4*>(p.o + grow * p.ldo + h * WA_HD + cc * 8) =
            *reinterpret_cast<const uint4*>(sQ + i * PITCH + cc * 8);
      }
    }
  }
}

constexpr int WB_ACCP = 146;  // fp32 pitch of the d(bias) flush matrix (144 x 146 x 4 B fits in the P + dS tiles)
static_assert(WA_ROWS * WB_ACCP * 4 <= 2 * WA_ROWS * WA_SP * 2, "flush matrix must fit in the P / dS tiles");

// type 0: dV[tile] = P^T dO;  type 1: dK[tile] = dS^T Q;  type 2: dQ[tile] = dS K
__device__ __forceinline__ void wb_phase_b(int type, int tile, int n_tiles, const bf16* sQ, const bf16* sdO,
                                           const bf16* sK, const bf16* sP, const bf16* sdS, int lane,
                                           float (&acc)[4][4]) {
  constexpr int PITCH = WA_PITCH, SP = WA_SP;
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  const int m0 = tile * 16;
  const bf16* Bm = type == 0 ? sdO : (type == 1 ? sQ : sK);
  const uint32_t b_base = smem_u32(Bm + ((lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + (lane >> 4) * 8);
  if (type < 2) {
    const bf16* Am = type == 0 ? sP : sdS;
    // A^T: rows of the smem tile are the reduction (query) index, columns the output (key) index
    const uint32_t a_base = smem_u32(Am + ((lane & 7) + ((lane >> 4) << 3)) * SP + m0 + ((lane >> 3) & 1) * 8);
#pragma unroll 3
    for (int kt = 0; kt < n_tiles; ++kt) {
      uint32_t a[4], b0, b1, b2, b3;
      ldsm_x4_t(a_base + kt * 16 * SP * 2, a[0], a[1], a[2], a[3]);
      ldsm_x4_t(b_base + kt * 16 * PITCH * 2, b0, b1, b2, b3);
      mma16816(acc[0], a, b0, b1);
      mma16816(acc[1], a, b2, b3);
      ldsm_x4_t(b_base + kt * 16 * PITCH * 2 + 32, b0, b1, b2, b3);
      mma16816(acc[2], a, b0, b1);
      mma16816(acc[3], a, b2, b3);
    }
  } else {
    const uint32_t a_base = smem_u32(sdS + (m0 + (lane & 7) + ((lane >> 3) & 1) * 8) * SP + (lane >> 4) * 8);
#pragma unroll 3
    for (int kt = 0; kt < n_tiles; ++kt) {
      uint32_t a[4], b0, b1, b2, b3;
      ldsm_x4(a_base + kt * 16 * 2, a[0], a[1], a[2], a[3]);
      ldsm_x4_t(b_base + kt * 16 * PITCH * 2, b0, b1, b2, b3);
      mma16816(acc[0], a, b0, b1);
      mma16816(acc[1], a, b2, b3);
      ldsm_x4_t(b_base + kt * 16 * PITCH * 2 + 32, b0, b1, b2, b3);
      mma16816(acc[2], a, b0, b1);
      mma16816(acc[3], a, b2, b3);
    }
  }
}

// =================================================================================================
// Backward: 16 warps (4 per scheduler), one CTA per SM, ONE __syncthreads per window.
// The 27 score jobs (A) and 27 output jobs (B) of a window form one dependency-ordered list
//     A(third 0) x9, A(third 1) x9, A(third 2) x9,
//     [dV, dK](key tiles of third 0), [dV, dK](third 1), [dV, dK](third 2), dQ x9
// dealt round-robin to the warps (position = warp + 16 k).  A jobs publish their 48 key columns of
// P / dS through one mbarrier per third (every lane arrives after its stores); dV/dK jobs wait only
// for the third that holds their key tile, dQ jobs for all three — so warps that finish their score
// jobs early roll straight into output jobs instead of idling at a block barrier.  A job only ever
// waits for jobs EARLIER in the list, and every warp walks its positions in order, so the earliest
// unfinished job is always runnable (no deadlock).  Score jobs work on 16-key slices (two n-tiles)
// to keep the live accumulators small; d(bias) sums stay in registers (two A slots per warp).
// =================================================================================================
constexpr int W3_WARPS = 16;
constexpr int W3_THREADS = W3_WARPS * 32;
constexpr int W3_STG = 16 * 32;  // bf16 elements of one warp's staging tile: 16 rows x 64 B, 16-byte chunks XOR-swizzled

template <bool MASKED>
__device__ __forceinline__ void w3_job_a(const bf16* sQ, const bf16* sdO, const bf16* sK, const bf16* sV, bf16* sP,
                                         bf16* sdS, const float* lse_s, const float* d_s, const WinTables& T,
                                         const char* tbl_bytes, int rt, int third, int lane, float scale2,
                                         float inv_scale, int emask, float (&dbacc)[6][4]) {
  constexpr int PITCH = WA_PITCH, SP = WA_SP;
  uint32_t qf[2][4], dof[2][4];
  {
    const int row = rt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
    const int col = (lane >> 4) * 8;
    ldsm_x4(smem_u32(sQ + row * PITCH + col), qf[0][0], qf[0][1], qf[0][2], qf[0][3]);
    ldsm_x4(smem_u32(sQ + row * PITCH + col + 16), qf[1][0], qf[1][1], qf[1][2], qf[1][3]);
    ldsm_x4(smem_u32(sdO + row * PITCH + col), dof[0][0], dof[0][1], dof[0][2], dof[0][3]);
    ldsm_x4(smem_u32(sdO + row * PITCH + col + 16), dof[1][0], dof[1][1], dof[1][2], dof[1][3]);
  }
  const int rl0 = rt * 16 + (lane >> 2);
  const int c2 = (lane & 3) * 2;
  const float nl0 = -lse_s[rl0] * inv_scale, nl1 = -lse_s[rl0 + 8] * inv_scale;
  const float nd0 = -d_s[rl0], nd1 = -d_s[rl0 + 8];
  const int aq0 = T.aq4[rl0], aq1 = T.aq4[rl0 + 8];
  int ci0 = 0, ci1 = 0;
  if (MASKED) {
    ci0 = T.code[rl0] & emask;
    ci1 = T.code[rl0 + 8] & emask;
  }
  const uint32_t kv_off = ((third * 48 + (lane & 7) + ((lane >> 4) << 3)) * PITCH + ((lane >> 3) & 1) * 8) * 2;
  const uint32_t k_base = smem_u32(sK) + kv_off, v_base = smem_u32(sV) + kv_off;
#pragma unroll
  for (int nt2 = 0; nt2 < 3; ++nt2) {  // 16 keys = two n-tiles per slice
    float s[2][4], dp[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      s[i][0] = s[i][1] = nl0; s[i][2] = s[i][3] = nl1;
      dp[i][0] = dp[i][1] = nd0; dp[i][2] = dp[i][3] = nd1;
    }
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(k_base + (nt2 * 16 * PITCH + ks * 16) * 2, b0, b1, b2, b3);
      mma16816(s[0], qf[ks], b0, b1);
      mma16816(s[1], qf[ks], b2, b3);
      ldsm_x4(v_base + (nt2 * 16 * PITCH + ks * 16) * 2, b0, b1, b2, b3);
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
      *reinterpret_cast<uint32_t*>(sP + rl0 * SP + j0) = pack_bf16(s[i][0], s[i][1]);
      *reinterpret_cast<uint32_t*>(sP + (rl0 + 8) * SP + j0) = pack_bf16(s[i][2], s[i][3]);
      *reinterpret_cast<uint32_t*>(sdS + rl0 * SP + j0) = pack_bf16(dp[i][0], dp[i][1]);
      *reinterpret_cast<uint32_t*>(sdS + (rl0 + 8) * SP + j0) = pack_bf16(dp[i][2], dp[i][3]);
    }
  }
}

__global__ void __launch_bounds__(W3_THREADS, 1) win_attn_bwd_kernel(const AttnParams p, const float* __restrict__ Dg) {
  pdl_trigger();
  pdl_wait();
  constexpr int PITCH = WA_PITCH, SP = WA_SP;
  extern __shared__ __align__(16) uint8_t smem[];
  bf16* tiles = reinterpret_cast<bf16*>(smem);  // [2 buffers][Q, dO, K, V][144][PITCH]
  bf16* sP = tiles + 2 * 4 * WA_TILE;
  bf16* sdS = sP + WA_ROWS * SP;
  bf16* stage = sdS + WA_ROWS * SP;             // [16 warps][16 rows][32]
  float* sLse = reinterpret_cast<float*>(stage + W3_WARPS * W3_STG);  // [2][144]
  float* sD = sLse + 2 * WA_ROWS;                                      // [2][144]
  WinTables T;
  T.tbl2 = sD + 2 * WA_ROWS;
  T.aq4 = reinterpret_cast<int*>(T.tbl2 + 2 * WA_MAXTBL + 2);
  T.bj4 = T.aq4 + WA_ROWS;
  T.code = T.bj4 + WA_ROWS;
  T.tok = T.code + WA_ROWS;
  int* sJob = T.tok + WA_ROWS;                                      // [64] B-job descriptors, see below
  uint64_t* bars = reinterpret_cast<uint64_t*>(sJob + 64);          // [3] one per key third
  const char* tbl_bytes = reinterpret_cast<const char*>(T.tbl2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x;
  const int N = p.Lq;
  WinGeo geo;
  geo.H = p.H; geo.W = p.W; geo.ws = p.ws; geo.shift = p.shift;
  geo.nWw = p.W / p.ws; geo.nWh = p.H / p.ws; geo.nW = geo.nWh * geo.nWw;
  const int n_groups = p.G * geo.nW;
  const int n_tiles = (N + 15) / 16;
  const int n_thirds = (N + 47) / 48;
  const int n_a = n_tiles * n_thirds;   // score jobs, list positions [0, n_a)
  const int n_b = 3 * n_tiles;          // output jobs, list positions [n_a, n_a + n_b)
  const float scale2 = p.scale * WA_LOG2E;
  const float inv_scale = 1.0f / p.scale;
  const int tw2 = 2 * p.ws - 1;

  fill_tables(T, p.bias_table, p.nH, h, p.ws, p.shift, N, tid, W3_THREADS);
  // B-job descriptor: type (0 dV, 1 dK, 2 dQ) | tile << 4 | wait mask (bit t = third t) << 12
  for (int i = tid; i < n_b; i += W3_THREADS) {
    int type, tile, dep;
    if (i < 2 * n_tiles) {  // [dV, dK] pairs in key-tile order
      tile = i >> 1; type = i & 1; dep = 1 << (tile / 3);
    } else {
      tile = i - 2 * n_tiles; type = 2; dep = (1 << n_thirds) - 1;
    }
    sJob[i] = type | (tile << 4) | (dep << 12);
  }
  if (tid == 0) {
    for (int t = 0; t < 3; ++t) mbar_init(&bars[t], n_tiles * 32);
    mbar_fence_init();
  }

  float dbacc[2][6][4];
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int i = 0; i < 6; ++i) dbacc[k][i][0] = dbacc[k][i][1] = dbacc[k][i][2] = dbacc[k][i][3] = 0.f;

  // this warp's score jobs (list positions warp and warp + 16) and first output position
  const int a_third0 = warp / n_tiles, a_rt0 = warp - a_third0 * n_tiles;
  const int a_third1 = (warp + W3_WARPS) / n_tiles, a_rt1 = (warp + W3_WARPS) - a_third1 * n_tiles;
  const bool has_a0 = warp < n_a, has_a1 = warp + W3_WARPS < n_a;
  int b_first = warp;
  while (b_first < n_a) b_first += W3_WARPS;
  b_first -= n_a;  // index into the B list

  // prefetch assignment: 576 16-byte chunks per tile; thread t takes chunk t and (t < 64) chunk t + 512
  const int pr0 = tid >> 2, pcc = tid & 3;
  const int pth0 = pr0 / p.ws, ptw0 = pr0 % p.ws;
  const int pth1 = (pr0 + 128) / p.ws, ptw1 = (pr0 + 128) % p.ws;
  const int col0 = h * WA_HD + pcc * 8;

  auto prefetch = [&](int g, int buf) {
    long long img_base; int h0, w0, em;
    geo.decode(g, img_base, h0, w0, em);
    bf16* tb = tiles + buf * 4 * WA_TILE;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (k == 1 && tid >= 64) break;
      const int r = pr0 + 128 * k;
      const bool valid = r < N;
      const long long grow = valid ? geo.row(img_base, h0, w0, k ? pth1 : pth0, k ? ptw1 : ptw0) : 0;
      const uint32_t so = smem_u32(tb + r * PITCH + pcc * 8);
      cp_async16(so, p.q + grow * p.ldq + col0, valid);
      cp_async16(so + WA_TILE * 2, p.d_o + grow * p.lddo + col0, valid);
      cp_async16(so + 2 * WA_TILE * 2, p.k + grow * p.ldk + col0, valid);
      cp_async16(so + 3 * WA_TILE * 2, p.v + grow * p.ldv + col0, valid);
      if (pcc == 0) cp_async4(smem_u32(sD + buf * WA_ROWS + r), Dg + grow * p.nH + h, valid);
      if (pcc == 1)
        cp_async4(smem_u32(sLse + buf * WA_ROWS + r),
                  p.lse + (static_cast<long long>(g) * p.nH + h) * N + (valid ? r : 0), valid);
    }
    cp_async_commit();
  };

  int it = 0;
  if (static_cast<int>(blockIdx.y) < n_groups) prefetch(blockIdx.y, 0);
  for (int g = blockIdx.y; g < n_groups; g += gridDim.y, ++it) {
    const int cur = it & 1;
    const uint32_t parity = it & 1;
    cp_async_wait0();
    __syncthreads();  // tiles[cur] visible (first time: tables, barriers); all jobs of the previous window are complete
    const int g_next = g + gridDim.y;
    if (g_next < n_groups) prefetch(g_next, cur ^ 1);
    const bf16* sQ = tiles + cur * 4 * WA_TILE;
    const bf16* sdO = sQ + WA_TILE;
    const bf16* sK = sdO + WA_TILE;
    const bf16* sV = sK + WA_TILE;
    const float* lse_s = sLse + cur * WA_ROWS;
    const float* d_s = sD + cur * WA_ROWS;
    long long img_base; int h0, w0, emask;
    geo.decode(g, img_base, h0, w0, emask);

    // ---------------- score jobs ----------------
    if (has_a0) {
      if (emask)
        w3_job_a<true>(sQ, sdO, sK, sV, sP, sdS, lse_s, d_s, T, tbl_bytes, a_rt0, a_third0, lane, scale2, inv_scale,
                       emask, dbacc[0]);
      else
        w3_job_a<false>(sQ, sdO, sK, sV, sP, sdS, lse_s, d_s, T, tbl_bytes, a_rt0, a_third0, lane, scale2, inv_scale,
                        0, dbacc[0]);
      mbar_arrive(&bars[a_third0]);  // every lane: its P / dS stores are released to the waiters
    }
    if (has_a1) {
      if (emask)
        w3_job_a<true>(sQ, sdO, sK, sV, sP, sdS, lse_s, d_s, T, tbl_bytes, a_rt1, a_third1, lane, scale2, inv_scale,
                       emask, dbacc[1]);
      else
        w3_job_a<false>(sQ, sdO, sK, sV, sP, sdS, lse_s, d_s, T, tbl_bytes, a_rt1, a_third1, lane, scale2, inv_scale,
                        0, dbacc[1]);
      mbar_arrive(&bars[a_third1]);
    }

    // ---------------- output jobs ----------------
    bf16* stg = stage + warp * W3_STG;
#pragma unroll 1
    for (int jb = b_first; jb < n_b; jb += W3_WARPS) {
      const int desc = sJob[jb];
      const int type = desc & 15, tile = (desc >> 4) & 255, dep = desc >> 12;
      if (dep & 1) mbar_wait(&bars[0], parity);
      if (dep & 2) mbar_wait(&bars[1], parity);
      if (dep & 4) mbar_wait(&bars[2], parity);
      float acc[4][4];
      wb_phase_b(type, tile, n_tiles, sQ, sdO, sK, sP, sdS, lane, acc);
      const float sc = type == 0 ? 1.0f : p.scale;
      const int r_lo = lane >> 2;
      const int swz = (r_lo >> 1) & 3;  // same for rows r_lo and r_lo + 8
#pragma unroll
      for (int dt = 0; dt < 4; ++dt) {
        const int off = ((dt ^ swz) << 3) + (lane & 3) * 2;
        *reinterpret_cast<uint32_t*>(stg + r_lo * 32 + off) = pack_bf16(acc[dt][0] * sc, acc[dt][1] * sc);
        *reinterpret_cast<uint32_t*>(stg + (r_lo + 8) * 32 + off) = pack_bf16(acc[dt][2] * sc, acc[dt][3] * sc);
      }
      __syncwarp();
      bf16* outp = type == 0 ? p.dv : (type == 1 ? p.dk : p.dq);
      const long long ldo = type == 0 ? p.lddv : (type == 1 ? p.lddk : p.lddq);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int r = (lane >> 2) + 8 * k, cc = lane & 3, i = tile * 16 + r;
        if (i < N) {
          const int tok = T.tok[i];
          const long long grow = geo.row(img_base, h0, w0, tok & 255, tok >> 8);
          *reinterpret_cast<uint4*>(outp + grow * ldo + h * WA_HD + cc * 8) =
              *reinterpret_cast<const uint4*>(stg + r * 32 + ((cc ^ ((r >> 1) & 3)) << 3));
        }
      }
      __syncwarp();  // staging tile is free again
    }
  }

  // Flush the register-resident d(bias) sums.  Every (i, j) position is owned by exactly one thread, so the
  // sums go to an fp32 [144][WB_ACCP] matrix aliased onto the P / dS tiles without atomics; each table entry
  // (dh, dw) is then the sum over the <= 144 token pairs with that offset: one global atomic per entry.
  cp_async_wait0();
  __syncthreads();
  float* sAcc = reinterpret_cast<float*>(sP);
#pragma unroll
  for (int slot = 0; slot < 2; ++slot) {
    if (slot == 0 ? has_a0 : has_a1) {
      const int third = slot == 0 ? a_third0 : a_third1, rt = slot == 0 ? a_rt0 : a_rt1;
      const int qi = rt * 16 + (lane >> 2);
#pragma unroll
      for (int nt = 0; nt < 6; ++nt) {
        const int j0 = third * 48 + nt * 8 + (lane & 3) * 2;
        *reinterpret_cast<float2*>(sAcc + qi * WB_ACCP + j0) = make_float2(dbacc[slot][nt][0], dbacc[slot][nt][1]);
        *reinterpret_cast<float2*>(sAcc + (qi + 8) * WB_ACCP + j0) = make_float2(dbacc[slot][nt][2], dbacc[slot][nt][3]);
      }
    }
  }
  __syncthreads();
  const int ws = p.ws;
  for (int t = tid; t < tw2 * tw2; t += W3_THREADS) {
    const int dh = t / tw2 - (ws - 1), dw = t % tw2 - (ws - 1);
    const int ih0 = max(0, dh), ih1 = min(ws, ws + dh), iw0 = max(0, dw), iw1 = min(ws, ws + dw);
    float sum = 0.f;
    for (int ih = ih0; ih < ih1; ++ih)
      for (int iw = iw0; iw < iw1; ++iw) sum += sAcc[(ih * ws + iw) * WB_ACCP + (ih - dh) * ws + (iw - dw)];
    atomicAdd(&p.dbias_table[t * p.nH + h], sum);
  }
}

// D[row, head] = sum_d dO[row, h*32 + d] * O[row, h*32 + d]   (pre-pass of the window backward)
__global__ void __launch_bounds__(256) win_attn_bwd_prep_kernel(const bf16* __restrict__ o, long long ldo,
                                                                const bf16* __restrict__ d_o, long long lddo,
                                                                float* __restrict__ D, long long rows, int C) {
  pdl_trigger();
  pdl_wait();
  const int vec_per_row = C / 8;
  const long long total = rows * vec_per_row;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i0 = static_cast<long long>(blockIdx.x) * blockDim.x; i0 < total; i0 += stride) {
    const long long i = i0 + threadIdx.x;
    float part = 0.f;
    long long r = 0;
    int c = 0;
    if (i < total) {
      r = i / vec_per_row;
      c = static_cast<int>(i - r * vec_per_row) * 8;
      const uint4 a = *reinterpret_cast<const uint4*>(o + r * ldo + c);
      const uint4 b = *reinterpret_cast<const uint4*>(d_o + r * lddo + c);
      const uint32_t* au = reinterpret_cast<const uint32_t*>(&a);
      const uint32_t* bu = reinterpret_cast<const uint32_t*>(&b);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 x = unpack_bf16(au[e]), y = unpack_bf16(bu[e]);
        part += x.x * y.x + x.y * y.y;
      }
    }
    part += __shfl_xor_sync(0xffffffffu, part, 1);  // 4 consecutive lanes hold one (row, head)
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    if (i < total && (threadIdx.x & 3) == 0) D[r * (C / WA_HD) + c / WA_HD] = part;
  }
}

}  // namespace

bool win_attn_supported(const AttnParams& p, int hd) {
  return p.mode == 1 && hd == WA_HD && p.Lq <= WA_ROWS && p.ws <= 12 && p.drop_p == 0.f;
}

int launch_win_fwd(const AttnParams& p, cudaStream_t stream) {
  const size_t smem = 2 * 3 * WA_TILE * 2 + (2 * WA_MAXTBL + 2) * 4 + 4 * WA_ROWS * 4;
  static bool attr_set = false;
  if (!attr_set) {
    FIBER_CUDA(cudaFuncSetAttribute(win_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int n_groups = p.G * (p.H / p.ws) * (p.W / p.ws);
  int gy = 2 * num_sms() / p.nH;  // two CTAs per SM, single wave, persistent over windows
  if (gy < 1) gy = 1;
  if (gy > n_groups) gy = n_groups;
  FIBER_CUDA(launch_k(win_attn_fwd_kernel, dim3(p.nH, gy), dim3(WF_THREADS), smem, stream, p));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// D[row, head] = rowsum(dO * O): pre-pass shared by both generations of the window backward
int launch_win_bwd_prep(const AttnParams& p, float* D, cudaStream_t stream) {
  const long long rows = static_cast<long long>(p.G) * p.H * p.W;
  const int C = p.nH * WA_HD;
  long long blocks = (rows * (C / 8) + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  FIBER_CUDA(launch_k(win_attn_bwd_prep_kernel, dim3(static_cast<int>(blocks)), dim3(256), 0, stream, p.o, p.ldo, p.d_o, p.lddo, D, rows, C));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int launch_win_bwd(const AttnParams& p, float* D, cudaStream_t stream) {
  if (launch_win_bwd_prep(p, D, stream)) return -2;
  const int n_groups = p.G * (p.H / p.ws) * (p.W / p.ws);
  int gy = num_sms() / p.nH;  // one CTA per SM (single wave), persistent over windows
  if (gy < 1) gy = 1;
  if (gy > n_groups) gy = n_groups;
  const size_t smem = 2 * 4 * WA_TILE * 2 + 2 * WA_ROWS * WA_SP * 2 + W3_WARPS * W3_STG * 2 + 4 * WA_ROWS * 4 +
                      (2 * WA_MAXTBL + 2) * 4 + 4 * WA_ROWS * 4 + 64 * 4 + 3 * 8 + 8;
  static bool attr_set = false;
  if (!attr_set) {
    FIBER_CUDA(cudaFuncSetAttribute(win_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  FIBER_CUDA(launch_k(win_attn_bwd_kernel, dim3(p.nH, gy), dim3(W3_THREADS), smem, stream, p, D));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace fiber
