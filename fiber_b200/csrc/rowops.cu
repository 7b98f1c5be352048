// fiber_b200 — HBM-bound row-wise kernels: LayerNorm fwd/bwd (optionally fused with a residual
// add or with the PatchMerging 2x2 gather), bias-gradient column sums, dot products (gate
// gradients), dropout, DropPath row scaling, fp32<->bf16 casts, weight cast+transpose,
// PatchEmbed patch gather, RoBERTa embedding gather / scatter.
// All are warp-per-row or grid-stride kernels with 128-bit accesses; fp32 math in registers.
#include "common.cuh"
#include "../../include/fiber_b200.h"

#include <cstdlib>

namespace fiber {

void count_launch(int n = 1);

// ---------------------------------------------------------------------------------------------
// LayerNorm
// ---------------------------------------------------------------------------------------------
struct LnParams {
  const bf16* in1;
  const bf16* in2;  // optional second addend (same layout as the output rows)
  long long ld1, ld2;
  const float* gamma;
  const float* beta;
  float eps;
  bf16* out;
  long long ldo;
  float* mean;
  float* rstd;
  bf16* sum_out;  // optional: in1 + in2
  long long lds;
  long long rows;
  int C;
  // PatchMerging gather (swin_transformer.py:420-427): rows = B*(H/2)*(W/2), C = 4*Cin
  int merge, H, W, Cin;
  // backward
  const bf16* dy;
  long long lddy;
  const bf16* dres;  // optional gradient added to dx (residual branch)
  long long lddres;
  bf16* dx;
  long long lddx;
  float* dgamma;
  float* dbeta;
  // optional second backward output: dx2[row,:] = dx[row,:] * row_scale[row / rps] (DropPath of the consumer)
  const float* row_scale;
  int rps;
  bf16* dx2;
  long long lddx2;
};

__device__ __forceinline__ long long ln_src_offset(const LnParams& p, long long row, int col) {
  if (!p.merge) return row * p.ld1 + col;
  const int W2 = p.W / 2, H2 = p.H / 2;
  const long long b = row / (H2 * W2);
  const int rem = static_cast<int>(row % (H2 * W2));
  const int h2 = rem / W2, w2 = rem % W2;
  const int seg = col / p.Cin, cin = col % p.Cin;
  const long long src = b * p.H * p.W + (2 * h2 + (seg & 1)) * p.W + (2 * w2 + (seg >> 1));
  return src * p.ld1 + cin;
}

__device__ __forceinline__ void load8(const bf16* ptr, float (&x)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(ptr);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = unpack_bf16(w[e]);
    x[2 * e] = f.x; x[2 * e + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(bf16* ptr, const float (&x)[8]) {
  uint4 u;
  u.x = pack_bf16(x[0], x[1]); u.y = pack_bf16(x[2], x[3]);
  u.z = pack_bf16(x[4], x[5]); u.w = pack_bf16(x[6], x[7]);
  *reinterpret_cast<uint4*>(ptr) = u;
}

// Lane mapping: LPR lanes cooperate on one row (LPR = 16 for C <= 128 so both half-warps work, else
// 32); each lane owns VPL vectors of 8 columns; ROWS row-groups are in flight per warp iteration so
// several 16-byte loads per lane are outstanding.  Statistics in one pass (sum, sum of squares) in fp32.
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int VPL, int LPR, int ROWS>
__global__ void __launch_bounds__(256, (VPL <= 2) ? 3 : 2) ln_fwd_kernel(const LnParams p) {
  pdl_trigger();
  pdl_wait();
  constexpr int RPW = 32 / LPR;  // rows per warp per group
  const int lane = threadIdx.x & 31, sl = lane % LPR, sr = lane / LPR;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const float invC = 1.0f / p.C;
  constexpr bool CACHE_GB = VPL <= 2;  // keep gamma/beta in registers only when that is cheap
  float g[CACHE_GB ? VPL : 1][8], b[CACHE_GB ? VPL : 1][8];
  if (CACHE_GB) {
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int col = (sl + LPR * v) * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        g[v][e] = col < p.C ? __ldg(p.gamma + col + e) : 0.f;
        b[v][e] = col < p.C ? __ldg(p.beta + col + e) : 0.f;
      }
    }
  }
  for (long long row0 = warp_global * (ROWS * RPW); row0 < p.rows; row0 += nwarps * (ROWS * RPW)) {
    float x[ROWS][VPL][8];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const long long row = row0 + r * RPW + sr;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int col = (sl + LPR * v) * 8;
        if (row < p.rows && col < p.C) {
          load8(p.in1 + ln_src_offset(p, row, col), x[r][v]);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) x[r][v][e] = 0.f;
        }
      }
    }
    if (p.in2) {
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const long long row = row0 + r * RPW + sr;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int col = (sl + LPR * v) * 8;
          if (row < p.rows && col < p.C) {
            float y[8];
            load8(p.in2 + row * p.ld2 + col, y);
#pragma unroll
            for (int e = 0; e < 8; ++e) x[r][v][e] += y[e];
            if (p.sum_out) store8(p.sum_out + row * p.lds + col, x[r][v]);
          }
        }
      }
    }
    float s1[ROWS], s2[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      s1[r] = s2[r] = 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          s1[r] += x[r][v][e];
          s2[r] = fmaf(x[r][v][e], x[r][v][e], s2[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      s1[r] = group_sum<LPR>(s1[r]);
      s2[r] = group_sum<LPR>(s2[r]);
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const long long row = row0 + r * RPW + sr;
      const float mean = s1[r] * invC;
      const float rstd = rsqrtf(fmaxf(s2[r] * invC - mean * mean, 0.f) + p.eps);
      if (row < p.rows) {
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int col = (sl + LPR * v) * 8;
          if (col < p.C) {
            float y[8];
            if (CACHE_GB) {
#pragma unroll
              for (int e = 0; e < 8; ++e) y[e] = fmaf((x[r][v][e] - mean) * rstd, g[v][e], b[v][e]);
            } else {
              const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + col));
              const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + col + 4));
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta + col));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.beta + col + 4));
              const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int e = 0; e < 8; ++e) y[e] = fmaf((x[r][v][e] - mean) * rstd, gg[e], bb[e]);
            }
            store8(p.out + row * p.ldo + col, y);
          }
        }
        if (sl == 0) {
          if (p.mean) p.mean[row] = mean;
          if (p.rstd) p.rstd[row] = rstd;
        }
      }
    }
  }
}

template <int VPL, int LPR, int ROWS>
__global__ void __launch_bounds__(256, (VPL <= 2) ? 2 : 1) ln_bwd_kernel(const LnParams p) {
  pdl_trigger();
  pdl_wait();
  constexpr int RPW = 32 / LPR;
  extern __shared__ float s_acc[];  // [2][C]: dgamma, dbeta block partials
  for (int i = threadIdx.x; i < 2 * p.C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, sl = lane % LPR, sr = lane / LPR;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const float invC = 1.0f / p.C;
  float dg[VPL][8], db[VPL][8];
#pragma unroll
  for (int v = 0; v < VPL; ++v)
#pragma unroll
    for (int e = 0; e < 8; ++e) dg[v][e] = db[v][e] = 0.f;

  for (long long row0 = warp_global * (ROWS * RPW); row0 < p.rows; row0 += nwarps * (ROWS * RPW)) {
    float xh[ROWS][VPL][8], gy[ROWS][VPL][8];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const long long row = row0 + r * RPW + sr;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int col = (sl + LPR * v) * 8;
        if (row < p.rows && col < p.C) {
          load8(p.in1 + ln_src_offset(p, row, col), xh[r][v]);
          load8(p.dy + row * p.lddy + col, gy[r][v]);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) xh[r][v][e] = gy[r][v][e] = 0.f;
        }
      }
    }
    if (p.in2) {
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const long long row = row0 + r * RPW + sr;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int col = (sl + LPR * v) * 8;
          if (row < p.rows && col < p.C) {
            float y[8];
            load8(p.in2 + row * p.ld2 + col, y);
#pragma unroll
            for (int e = 0; e < 8; ++e) xh[r][v][e] += y[e];
          }
        }
      }
    }
    float c1[ROWS], c2[ROWS], rs[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const long long row = row0 + r * RPW + sr;
      const bool rv = row < p.rows;
      const float mean = rv ? p.mean[row] : 0.f;
      rs[r] = rv ? p.rstd[row] : 0.f;
      c1[r] = c2[r] = 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int col = (sl + LPR * v) * 8;
        if (col < p.C) {
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + col));
          const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + col + 4));
          const float gam[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float dy = gy[r][v][e];
            const float xhat = (xh[r][v][e] - mean) * rs[r];
            xh[r][v][e] = xhat;
            gy[r][v][e] = dy * gam[e];
            c1[r] += gy[r][v][e];
            c2[r] = fmaf(gy[r][v][e], xhat, c2[r]);
            dg[v][e] = fmaf(dy, xhat, dg[v][e]);
            db[v][e] += dy;
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      c1[r] = group_sum<LPR>(c1[r]);
      c2[r] = group_sum<LPR>(c2[r]);
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const long long row = row0 + r * RPW + sr;
      if (row < p.rows) {
        const float k1 = c1[r] * invC, k2 = c2[r] * invC;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int col = (sl + LPR * v) * 8;
          if (col < p.C) {
            float dx[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) dx[e] = rs[r] * (gy[r][v][e] - k1 - xh[r][v][e] * k2);
            if (p.dres) {
              float rr[8];
              load8(p.dres + (p.merge ? ln_src_offset(p, row, col) / p.ld1 * p.lddres + (col % p.Cin)
                                      : row * p.lddres + col), rr);
#pragma unroll
              for (int e = 0; e < 8; ++e) dx[e] += rr[e];
            }
            const long long off = p.merge ? ln_src_offset(p, row, col) / p.ld1 * p.lddx + (col % p.Cin)
                                          : row * p.lddx + col;
            store8(p.dx + off, dx);
            if (p.dx2) {  // non-merging rows only (checked at the C-ABI)
              const float sc = __ldg(p.row_scale + row / p.rps);
#pragma unroll
              for (int e = 0; e < 8; ++e) dx[e] *= sc;
              store8(p.dx2 + row * p.lddx2 + col, dx);
            }
          }
        }
      }
    }
  }
  if (p.dgamma) {
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int col = (sl + LPR * v) * 8;
      if (col < p.C) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          atomicAdd(&s_acc[col + e], dg[v][e]);
          atomicAdd(&s_acc[p.C + col + e], db[v][e]);
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.C; i += blockDim.x) {
      atomicAdd(p.dgamma + i, s_acc[i]);
      atomicAdd(p.dbeta + i, s_acc[p.C + i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fast paths (plain rows: no second input, no PatchMerging gather, C == 8 * VPL * LPR exactly) — the
// Swin norm1 / norm2 / norm_i2t and final-norm LayerNorms, i.e. almost all LayerNorm bytes of a step.
// Every 16-byte load of a warp iteration (x [, dy, dres]) is issued before anything is consumed and the
// rows stay PACKED (bf16) in registers; x_hat / dy*gamma are recomputed in the second pass instead of
// being kept as fp32.  That puts ROWS*VPL*(1 or 3)*16 bytes per lane in flight with ONE exposed memory
// latency per iteration (the generic kernels expose two and hold a third of the bytes).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& u, float (&x)[8]) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = unpack_bf16(w[e]);
    x[2 * e] = f.x; x[2 * e + 1] = f.y;
  }
}
// Compiler barrier on a packed vector: stops CSE from keeping the fp32 unpacking of pass 1 alive for pass 2
// (the point of the fast kernels is that only the packed registers persist).
__device__ __forceinline__ void keep_packed(uint4& u) {
  asm volatile("" : "+r"(u.x), "+r"(u.y), "+r"(u.z), "+r"(u.w));
}
__device__ __forceinline__ void ldg8f(const float* ptr, float (&x)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(ptr));
  const float4 b = __ldg(reinterpret_cast<const float4*>(ptr + 4));
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}

// Row addressing shared by the fast kernels.  Plain rows: element offset = row * ld + col.  PatchMerging rows
// (MERGE): the 4*Cin columns of merged row (b, h2, w2) are the Cin channels of source tokens
// (2h2 + (seg & 1), 2w2 + (seg >> 1)), seg = col / Cin (swin_transformer.py:420-427); the two divisions happen
// once per row, the per-vector part (segment shift, channel offset) is a per-lane constant.
template <bool MERGE>
__device__ __forceinline__ long long ln_row_base(const LnParams& p, long long row) {
  if (!MERGE) return row;
  const int W2 = p.W / 2, H2 = p.H / 2;
  const long long b = row / (H2 * W2);
  const int rem = static_cast<int>(row - b * (H2 * W2));
  const int h2 = rem / W2, w2 = rem - h2 * W2;
  return b * p.H * p.W + static_cast<long long>(2 * h2) * p.W + 2 * w2;
}

template <int VPL, int LPR, int ROWS, bool MERGE>
__global__ void __launch_bounds__(256, 3) ln_fwd_fast_kernel(const LnParams p) {
  pdl_trigger();
  pdl_wait();
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sl = lane % LPR, sr = lane / LPR;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const float invC = 1.0f / p.C;
  int sadd[VPL], ccol[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int col = (sl + LPR * v) * 8;
    const int seg = MERGE ? col / p.Cin : 0;
    sadd[v] = MERGE ? (seg & 1) * p.W + (seg >> 1) : 0;
    ccol[v] = MERGE ? col - seg * p.Cin : col;
  }
  for (long long row0 = warp_global * (ROWS * RPW); row0 < p.rows; row0 += nwarps * (ROWS * RPW)) {
    uint4 xp[ROWS][VPL];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const long long row = row0 + r * RPW + sr;
      const long long base = ln_row_base<MERGE>(p, row < p.rows ? row : 0);
#pragma unroll
      for (int v = 0; v < VPL; ++v)
        xp[r][v] = row < p.rows ? *reinterpret_cast<const uint4*>(p.in1 + (base + sadd[v]) * p.ld1 + ccol[v])
                                : make_uint4(0u, 0u, 0u, 0u);
    }
    float s1[ROWS], s2[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      s1[r] = s2[r] = 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        float x[8];
        unpack8(xp[r][v], x);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          s1[r] += x[e];
          s2[r] = fmaf(x[e], x[e], s2[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      s1[r] = group_sum<LPR>(s1[r]);
      s2[r] = group_sum<LPR>(s2[r]);
#pragma unroll
      for (int v = 0; v < VPL; ++v) keep_packed(xp[r][v]);
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const long long row = row0 + r * RPW + sr;
      const float mean = s1[r] * invC;
      const float rstd = rsqrtf(fmaxf(s2[r] * invC - mean * mean, 0.f) + p.eps);
      if (row < p.rows) {
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int col = (sl + LPR * v) * 8;
          float x[8], g[8], b[8], y[8];
          unpack8(xp[r][v], x);
          ldg8f(p.gamma + col, g);
          ldg8f(p.beta + col, b);
#pragma unroll
          for (int e = 0; e < 8; ++e) y[e] = fmaf((x[e] - mean) * rstd, g[e], b[e]);
          store8(p.out + row * p.ldo + col, y);
        }
        if (sl == 0) {
          if (p.mean) p.mean[row] = mean;
          if (p.rstd) p.rstd[row] = rstd;
        }
      }
    }
  }
}

// SACC: the dgamma / dbeta partial sums of wide rows (VPL >= 3: 48-64 registers) live in a per-warp
// shared-memory slice instead of registers, which is what lets the packed rows stay register-resident.
template <int VPL, int LPR, int ROWS, bool MERGE, bool SACC>
__global__ void __launch_bounds__(256, 2) ln_bwd_fast_kernel(const LnParams p) {
  pdl_trigger();
  pdl_wait();
  constexpr int RPW = 32 / LPR;
  extern __shared__ __align__(16) float s_fast[];  // [2][C] block partials (+ [8 warps][2][C] when SACC)
  float* s_acc = s_fast;
  const int n_acc = SACC ? 18 * p.C : 2 * p.C;
  for (int i = threadIdx.x; i < n_acc; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, sl = lane % LPR, sr = lane / LPR;
  float* wacc = s_acc + 2 * p.C + (threadIdx.x >> 5) * 2 * p.C;  // SACC only
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const float invC = 1.0f / p.C;
  const bool has_res = p.dres != nullptr;
  int sadd[VPL], ccol[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int col = (sl + LPR * v) * 8;
    const int seg = MERGE ? col / p.Cin : 0;
    sadd[v] = MERGE ? (seg & 1) * p.W + (seg >> 1) : 0;
    ccol[v] = MERGE ? col - seg * p.Cin : col;
  }
  float dg[SACC ? 1 : VPL][8], db[SACC ? 1 : VPL][8];
#pragma unroll
  for (int v = 0; v < (SACC ? 1 : VPL); ++v)
#pragma unroll
    for (int e = 0; e < 8; ++e) dg[v][e] = db[v][e] = 0.f;

  for (long long row0 = warp_global * (ROWS * RPW); row0 < p.rows; row0 += nwarps * (ROWS * RPW)) {
    uint4 xp[ROWS][VPL], yp[ROWS][VPL], rp[ROWS][VPL];
    float mean[ROWS], rs[ROWS];
    long long base[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const long long row = row0 + r * RPW + sr;
      const bool rv = row < p.rows;
      base[r] = ln_row_base<MERGE>(p, rv ? row : 0);
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int col = (sl + LPR * v) * 8;
        const long long src = base[r] + sadd[v];
        xp[r][v] = rv ? *reinterpret_cast<const uint4*>(p.in1 + src * p.ld1 + ccol[v]) : make_uint4(0u, 0u, 0u, 0u);
        yp[r][v] = rv ? *reinterpret_cast<const uint4*>(p.dy + row * p.lddy + col) : make_uint4(0u, 0u, 0u, 0u);
        rp[r][v] = (rv && has_res) ? *reinterpret_cast<const uint4*>(p.dres + src * p.lddres + ccol[v])
                                   : make_uint4(0u, 0u, 0u, 0u);
      }
      mean[r] = rv ? p.mean[row] : 0.f;
      rs[r] = rv ? p.rstd[row] : 0.f;
    }
    float c1[ROWS], c2[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) c1[r] = c2[r] = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int col = (sl + LPR * v) * 8;
      float gam[8];
      ldg8f(p.gamma + col, gam);
      float tg[8], tb[8];  // this iteration's dgamma / dbeta contribution of vector v (SACC)
#pragma unroll
      for (int e = 0; e < 8; ++e) tg[e] = tb[e] = 0.f;
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        float x[8], dy[8];
        unpack8(xp[r][v], x);
        unpack8(yp[r][v], dy);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xhat = (x[e] - mean[r]) * rs[r];
          const float gy = dy[e] * gam[e];
          c1[r] += gy;
          c2[r] = fmaf(gy, xhat, c2[r]);
          if (SACC) {
            tg[e] = fmaf(dy[e], xhat, tg[e]);
            tb[e] += dy[e];
          } else {
            dg[v][e] = fmaf(dy[e], xhat, dg[v][e]);
            db[v][e] += dy[e];
          }
        }
      }
      if (SACC) {  // lane-private columns of a warp-private slice: plain read-modify-write, no atomics
        float4* ag = reinterpret_cast<float4*>(wacc + col);
        float4* ab = reinterpret_cast<float4*>(wacc + p.C + col);
        float4 g0 = ag[0], g1 = ag[1], b0 = ab[0], b1 = ab[1];
        g0.x += tg[0]; g0.y += tg[1]; g0.z += tg[2]; g0.w += tg[3];
        g1.x += tg[4]; g1.y += tg[5]; g1.z += tg[6]; g1.w += tg[7];
        b0.x += tb[0]; b0.y += tb[1]; b0.z += tb[2]; b0.w += tb[3];
        b1.x += tb[4]; b1.y += tb[5]; b1.z += tb[6]; b1.w += tb[7];
        ag[0] = g0; ag[1] = g1; ab[0] = b0; ab[1] = b1;
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      c1[r] = group_sum<LPR>(c1[r]);
      c2[r] = group_sum<LPR>(c2[r]);
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        keep_packed(xp[r][v]);
        keep_packed(yp[r][v]);
      }
    }
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int col = (sl + LPR * v) * 8;
      float gam[8];
      ldg8f(p.gamma + col, gam);
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const long long row = row0 + r * RPW + sr;
        if (row < p.rows) {
          const float k1 = c1[r] * invC, k2 = c2[r] * invC;
          float x[8], dy[8], rr[8], dx[8];
          unpack8(xp[r][v], x);
          unpack8(yp[r][v], dy);
          unpack8(rp[r][v], rr);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float xhat = (x[e] - mean[r]) * rs[r];
            dx[e] = fmaf(rs[r], dy[e] * gam[e] - k1 - xhat * k2, rr[e]);
          }
          store8(p.dx + (base[r] + sadd[v]) * p.lddx + ccol[v], dx);
          if (!MERGE && p.dx2) {
            const float sc = __ldg(p.row_scale + row / p.rps);
#pragma unroll
            for (int e = 0; e < 8; ++e) dx[e] *= sc;
            store8(p.dx2 + row * p.lddx2 + col, dx);
          }
        }
      }
    }
  }
  if (p.dgamma) {
    if (SACC) {
      __syncthreads();
      for (int i = threadIdx.x; i < 2 * p.C; i += blockDim.x) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_acc[2 * p.C + w * 2 * p.C + i];
        atomicAdd((i < p.C ? p.dgamma : p.dbeta - p.C) + i, t);
      }
    } else {
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int col = (sl + LPR * v) * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          atomicAdd(&s_acc[col + e], dg[v][e]);
          atomicAdd(&s_acc[p.C + col + e], db[v][e]);
        }
      }
      __syncthreads();
      for (int i = threadIdx.x; i < p.C; i += blockDim.x) {
        atomicAdd(p.dgamma + i, s_acc[i]);
        atomicAdd(p.dbeta + i, s_acc[p.C + i]);
      }
    }
  }
}

static int ln_grid(long long rows, int rows_per_warp, int blocks_per_sm) {
  const long long blocks = (rows + 8 * rows_per_warp - 1) / (8 * rows_per_warp);  // 8 warps per block
  const long long cap = static_cast<long long>(num_sms()) * blocks_per_sm;
  return static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

template <int VPL, int LPR, int ROWS>
static int ln_launch(const LnParams& p, bool bwd, cudaStream_t stream) {
  const int rpw = ROWS * (32 / LPR);
  if (!bwd) {
    FIBER_CUDA(launch_k(ln_fwd_kernel<VPL, LPR, ROWS>, dim3(ln_grid(p.rows, rpw, 6)), dim3(256), 0, stream, p));
  } else {
    // few resident blocks => few global atomics for dgamma/dbeta
    FIBER_CUDA(launch_k(ln_bwd_kernel<VPL, LPR, ROWS>, dim3(ln_grid(p.rows, rpw, 2)), dim3(256), 2 * p.C * sizeof(float), stream, p));
  }
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

template <int VPL, int LPR, int ROWS, bool MERGE>
static int ln_launch_fast(const LnParams& p, bool bwd, cudaStream_t stream) {
  const int rpw = ROWS * (32 / LPR);
  if (!bwd) {
    FIBER_CUDA(launch_k(ln_fwd_fast_kernel<VPL, LPR, ROWS, MERGE>, dim3(ln_grid(p.rows, rpw, 3)), dim3(256), 0, stream, p));
  } else {
    constexpr bool SACC = VPL >= 3;
    auto kern = ln_bwd_fast_kernel<VPL, LPR, ROWS, MERGE, SACC>;
    const size_t smem = (SACC ? 18 : 2) * p.C * sizeof(float);
    if (smem > 48 * 1024) {
      static bool attr_set = false;
      if (!attr_set) {
        FIBER_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 18 * 1024 * (int)sizeof(float)));
        attr_set = true;
      }
    }
    FIBER_CUDA(launch_k(kern, dim3(ln_grid(p.rows, rpw, 2)), dim3(256), smem, stream, p));
  }
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

static bool ln_aligned16(const void* ptr, long long ld) {
  return ptr == nullptr || ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ld % 8 == 0);
}

int ln_dispatch(const LnParams& p, bool bwd, cudaStream_t stream) {
  FIBER_CHECK(p.C % 8 == 0 && p.C >= 8 && p.C <= 2048, "LayerNorm width must be a multiple of 8 in [8, 2048] (got %d)", p.C);
  FIBER_CHECK(p.rows > 0, "LayerNorm needs rows > 0");
  if (p.merge) {
    FIBER_CHECK(p.C == 4 * p.Cin && p.Cin % 8 == 0 && p.H % 2 == 0 && p.W % 2 == 0 && !p.in2,
                "bad PatchMerging LayerNorm geometry");
  }
  const bool fast = !p.in2 && !p.sum_out && ln_aligned16(p.in1, p.ld1) &&
                    (bwd ? (ln_aligned16(p.dy, p.lddy) && ln_aligned16(p.dres, p.lddres) && ln_aligned16(p.dx, p.lddx) &&
                            ln_aligned16(p.dx2, p.lddx2))
                         : ln_aligned16(p.out, p.ldo));
  if (fast && !p.merge) {  // ROWS * VPL = up to 4 vectors per lane and tensor in flight
    if (p.C == 128) return ln_launch_fast<1, 16, 4, false>(p, bwd, stream);
    if (p.C == 256) return ln_launch_fast<1, 32, 4, false>(p, bwd, stream);
    if (p.C == 512) return bwd ? ln_launch_fast<2, 32, 1, false>(p, true, stream) : ln_launch_fast<2, 32, 2, false>(p, false, stream);
    if (p.C == 768) return ln_launch_fast<3, 32, 1, false>(p, bwd, stream);
    if (p.C == 1024) return ln_launch_fast<4, 32, 1, false>(p, bwd, stream);
  }
  if (fast && p.merge) {  // PatchMerging: C = 4 * Cin
    if (p.C == 256) return ln_launch_fast<1, 32, 4, true>(p, bwd, stream);
    if (p.C == 512) return bwd ? ln_launch_fast<2, 32, 1, true>(p, true, stream) : ln_launch_fast<2, 32, 2, true>(p, false, stream);
    if (p.C == 1024) return ln_launch_fast<4, 32, 1, true>(p, bwd, stream);
  }
  // unroll depths picked from tools/bench_ln.py on B200
  if (p.C <= 128) return ln_launch<1, 16, 2>(p, bwd, stream);
  if (p.C <= 256) return bwd ? ln_launch<1, 32, 4>(p, true, stream) : ln_launch<1, 32, 2>(p, false, stream);
  if (p.C <= 512) return ln_launch<2, 32, 1>(p, bwd, stream);
  if (p.C <= 768) return ln_launch<3, 32, 1>(p, bwd, stream);
  if (p.C <= 1024) return ln_launch<4, 32, 1>(p, bwd, stream);
  return ln_launch<8, 32, 1>(p, bwd, stream);
}

// ---------------------------------------------------------------------------------------------
// column sum (bias gradients):  out[n] (+)= scale * sum_m rs[m / rps] * x[m, n]
// ---------------------------------------------------------------------------------------------
// One block covers `cw` 16-byte vector columns (all of a row when N <= 2048) and 256 / cw rows per pass, so
// every thread issues coalesced 16-byte loads whatever N is; four independent loads are in flight per thread.
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ x, long long ld, long long M, int N,
                                                     float* out, const float* scale, const float* row_scale, int rps,
                                                     long long rows_per_block, int cw) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_part[2048];
  const int rpp = 256 / cw;                  // rows per pass
  const int slot = threadIdx.x / cw, lc = threadIdx.x - slot * cw;
  const int vcol = blockIdx.x * cw + lc;     // vector column of this thread
  const bool active = slot < rpp && vcol * 8 < N;
  const long long r0 = blockIdx.y * rows_per_block;
  const long long r1 = min(r0 + rows_per_block, M);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (active) {
    const bf16* xp = x + vcol * 8;
    long long r = r0 + slot;
    for (; r + 3LL * rpp < r1; r += 4LL * rpp) {
      float v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) load8(xp + (r + static_cast<long long>(u) * rpp) * ld, v[u]);
      if (row_scale) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float sc = row_scale[(r + static_cast<long long>(u) * rpp) / rps];
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = fmaf(v[u][e], sc, acc[e]);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += (v[0][e] + v[1][e]) + (v[2][e] + v[3][e]);
      }
    }
    for (; r < r1; r += rpp) {
      float v[8];
      load8(xp + r * ld, v);
      const float sc = row_scale ? row_scale[r / rps] : 1.0f;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(v[e], sc, acc[e]);
    }
  }
  if (slot < rpp) {
#pragma unroll
    for (int e = 0; e < 8; ++e) s_part[slot * cw * 8 + lc * 8 + e] = acc[e];
  }
  __syncthreads();
  const float sc = scale ? *scale : 1.0f;
  for (int c = threadIdx.x; c < cw * 8; c += 256) {
    float t = 0.f;
    for (int y = 0; y < rpp; ++y) t += s_part[y * cw * 8 + c];
    const int gc = blockIdx.x * cw * 8 + c;
    if (gc < N) atomicAdd(out + gc, t * sc);
  }
}

// dot(a, b) -> *out += sum a[i]*b[i]   (2-D, row-major with independent leading dimensions)
__global__ void __launch_bounds__(256) dot_kernel(const bf16* a, long long lda, const bf16* b, long long ldb,
                                                  long long M, int N, float* out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_w[8];
  const int vec_per_row = N / 8;
  const long long total = M * vec_per_row;
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / vec_per_row;
    const int c = static_cast<int>(i % vec_per_row) * 8;
    float x[8], y[8];
    load8(a + r * lda + c, x);
    load8(b + r * ldb + c, y);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc += x[e] * y[e];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = s_w[threadIdx.x];
    t += __shfl_xor_sync(0xffu, t, 4);
    t += __shfl_xor_sync(0xffu, t, 2);
    t += __shfl_xor_sync(0xffu, t, 1);
    if (threadIdx.x == 0) atomicAdd(out, t);
  }
}

// ---------------------------------------------------------------------------------------------
// elementwise
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool keep_elem(unsigned long long seed, unsigned long long idx, float p) {
  unsigned long long z = idx + seed * 0x9E3779B97F4A7C15ull + 0x2545F4914F6CDD1Dull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return static_cast<float>(static_cast<uint32_t>(z >> 40)) * (1.0f / 16777216.0f) >= p;
}

// mode 0: y = dropout(x; p, seed) (same mask for forward and backward)
// mode 1: y = x * row_scale[row / rps]
__global__ void __launch_bounds__(256) rowwise_scale_kernel(const bf16* x, long long ldx, bf16* y, long long ldy,
                                                            long long M, int N, int mode, float p,
                                                            unsigned long long seed, const float* row_scale, int rps) {
  pdl_trigger();
  pdl_wait();
  const int vec_per_row = N / 8;
  const long long total = M * vec_per_row;
  const float keep_inv = 1.0f / (1.0f - p);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / vec_per_row;
    const int c = static_cast<int>(i % vec_per_row) * 8;
    float v[8];
    load8(x + r * ldx + c, v);
    if (mode == 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        v[e] = keep_elem(seed, static_cast<unsigned long long>(r) * N + c + e, p) ? v[e] * keep_inv : 0.f;
    } else {
      const float s = row_scale[r / rps];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] *= s;
    }
    store8(y + r * ldy + c, v);
  }
}

// Crop / zero-pad of a token grid with the DropPath scale and the residual fused:
//   dst[b, h, w, :] = (h < Hs && w < Ws ? src[b, h, w, :] * row_scale[b] : 0) + add[b, h, w, :]      (h < Hd, w < Wd)
// src [B, Hs, Ws, C], dst / add [B, Hd, Wd, C], contiguous bf16; row_scale and add optional.  Hd <= Hs crops (the
// fine-grained block's `x = shortcut + drop_path(x[:, :H, :W])`, fusion_swin_transformer_v2.py:336-343), Hd > Hs pads
// with zeros (its backward, and the F.pad of :316-321).
__global__ void __launch_bounds__(256) grid_copy_kernel(const bf16* src, bf16* dst, const bf16* add, const float* row_scale,
                                                        int B, int Hs, int Ws, int Hd, int Wd, int C) {
  pdl_trigger();
  pdl_wait();
  const int vec_per_row = C / 8;
  const long long total = static_cast<long long>(B) * Hd * Wd * vec_per_row;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % vec_per_row) * 8;
    const long long tok = i / vec_per_row;
    const int w = static_cast<int>(tok % Wd);
    const int h = static_cast<int>((tok / Wd) % Hd);
    const int b = static_cast<int>(tok / (static_cast<long long>(Wd) * Hd));
    float v[8];
    if (h < Hs && w < Ws) {
      load8(src + ((static_cast<long long>(b) * Hs + h) * Ws + w) * C + c, v);
      if (row_scale) {
        const float s = row_scale[b];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= s;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
    }
    if (add) {
      float a[8];
      load8(add + tok * C + c, a);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += a[e];
    }
    store8(dst + tok * C + c, v);
  }
}

// out = add + (*alpha) * x   (gate: a + alpha_t2i * c, roberta.py:483; plain residual add when alpha == null)
__global__ void __launch_bounds__(256) axpy_kernel(const bf16* x, long long ldx, const bf16* add, long long lda,
                                                   const float* alpha, bf16* out, long long ldo, long long M, int N) {
  pdl_trigger();
  pdl_wait();
  const int vec_per_row = N / 8;
  const long long total = M * vec_per_row;
  const float al = alpha ? *alpha : 1.0f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / vec_per_row;
    const int c = static_cast<int>(i % vec_per_row) * 8;
    float v[8], w[8];
    load8(x + r * ldx + c, v);
    load8(add + r * lda + c, w);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = w[e] + al * v[e];
    store8(out + r * ldo + c, v);
  }
}

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* x, bf16* y, long long n) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x * 4) {
    if (i + 3 < n) {
      const float4 v = *reinterpret_cast<const float4*>(x + i);
      *reinterpret_cast<uint2*>(y + i) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    } else {
      for (long long j = i; j < n; ++j) y[j] = __float2bfloat16(x[j]);
    }
  }
}

// w f32 [N, K] (ldw) -> w_out bf16 [N, K] (ld_out) and wt_out bf16 [K, N] (ldt_out), either optional
__global__ void __launch_bounds__(256) cast_transpose_kernel(const float* w, long long ldw, int N, int K, bf16* w_out,
                                                             long long ld_out, bf16* wt_out, long long ldt_out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[32][33];
  const int n0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int n = n0 + i, k = k0 + tx;
    float v = 0.f;
    if (n < N && k < K) {
      v = w[n * ldw + k];
      if (w_out) w_out[n * ld_out + k] = __float2bfloat16(v);
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  if (wt_out) {
    for (int i = ty; i < 32; i += 8) {
      const int k = k0 + i, n = n0 + tx;
      if (n < N && k < K) wt_out[k * ldt_out + n] = __float2bfloat16(tile[tx][i]);
    }
  }
}

// PatchEmbed gather (timm PatchEmbed conv 4x4/s4 as a GEMM, fiber_module.py:311):
// img f32 [B,3,R,R] -> patches bf16 [B*(R/4)^2, 64]; column = c*16 + kh*4 + kw, columns 48..63 = 0.
__global__ void __launch_bounds__(256) patch_gather_kernel(const float* img, bf16* out, int B, int RH, int R) {
  pdl_trigger();
  pdl_wait();
  const int P = R / 4, PH = RH / 4;  // patches per row / per column (the image is RH x R)
  const long long total = static_cast<long long>(B) * PH * P * 16;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i & 15);
    const long long patch = i >> 4;
    uint2 o = make_uint2(0u, 0u);
    if (j < 12) {
      const int c = j >> 2, kh = j & 3;
      const int pw = static_cast<int>(patch % P), ph = static_cast<int>((patch / P) % PH);
      const long long b = patch / (static_cast<long long>(P) * PH);
      const float4 v = *reinterpret_cast<const float4*>(img + ((b * 3 + c) * RH + (ph * 4 + kh)) * R + pw * 4);
      o = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
    *reinterpret_cast<uint2*>(out + patch * 64 + j * 4) = o;
  }
}

// RoBERTa embeddings (roberta.py:169-196): out[t] = word[id] + pos[pid] + type[0]; pid from the
// running count of non-pad tokens (roberta.py:877-888).  One warp per token.
__device__ __forceinline__ int roberta_pos_id(const long long* ids_row, int l, int pad, int lane) {
  int count = 0;
  for (int base = 0; base <= l; base += 32) {
    const int j = base + lane;
    const bool nz = (j <= l) && (ids_row[j] != pad);
    count += __popc(__ballot_sync(0xffffffffu, nz));
  }
  return ids_row[l] != pad ? count + pad : pad;
}

__global__ void __launch_bounds__(256) embed_gather_kernel(const long long* ids, int B, int L, int C, int pad,
                                                           const float* word, const float* pos, const float* type,
                                                           bf16* out, long long ldo) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long t = warp_global; t < static_cast<long long>(B) * L; t += nwarps) {
    const int b = static_cast<int>(t / L), l = static_cast<int>(t % L);
    const long long id = ids[t];
    const int pid = roberta_pos_id(ids + static_cast<long long>(b) * L, l, pad, lane);
    for (int c = lane * 4; c < C; c += 128) {
      const float4 w = *reinterpret_cast<const float4*>(word + id * C + c);
      const float4 q = *reinterpret_cast<const float4*>(pos + static_cast<long long>(pid) * C + c);
      const float4 y = *reinterpret_cast<const float4*>(type + c);
      *reinterpret_cast<uint2*>(out + t * ldo + c) =
          make_uint2(pack_bf16(w.x + q.x + y.x, w.y + q.y + y.y), pack_bf16(w.z + q.z + y.z, w.w + q.w + y.w));
    }
  }
}

// scatter-add of d(embedding sum) into the word / position tables (fp32 atomics); the pad row of
// both tables receives no gradient (nn.Embedding(padding_idx=1), roberta.py:150,165).
__global__ void __launch_bounds__(256) embed_scatter_kernel(const long long* ids, int B, int L, int C, int pad,
                                                            const bf16* dsum, long long ldd, float* dword,
                                                            float* dpos) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long t = warp_global; t < static_cast<long long>(B) * L; t += nwarps) {
    const int b = static_cast<int>(t / L), l = static_cast<int>(t % L);
    const long long id = ids[t];
    const int pid = roberta_pos_id(ids + static_cast<long long>(b) * L, l, pad, lane);
    for (int c = lane * 2; c < C; c += 64) {
      const float2 g = unpack_bf16(*reinterpret_cast<const uint32_t*>(dsum + t * ldd + c));
      if (id != pad) {
        atomicAdd(dword + id * C + c, g.x);
        atomicAdd(dword + id * C + c + 1, g.y);
      }
      if (pid != pad) {
        atomicAdd(dpos + static_cast<long long>(pid) * C + c, g.x);
        atomicAdd(dpos + static_cast<long long>(pid) * C + c + 1, g.y);
      }
    }
  }
}

static int ew_grid(long long work_items) {
  long long blocks = (work_items + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  return static_cast<int>(blocks > 0 ? blocks : 1);
}

int colsum_dispatch(const bf16* x, long long ld, long long M, int N, float* out, const float* scale,
                    const float* row_scale, int rps, cudaStream_t stream) {
  FIBER_CHECK(N % 8 == 0 && M > 0, "colsum: N must be a multiple of 8");
  const int nv = N / 8;
  const int gx = (nv + 255) / 256;
  const int cw = (nv + gx - 1) / gx;  // vector columns per block (<= 256)
  const int rpp = 256 / cw;
  long long gy = (4LL * num_sms() + gx - 1) / gx;
  const long long min_rows = 8LL * rpp;  // at least two unrolled passes per block
  if (gy > (M + min_rows - 1) / min_rows) gy = (M + min_rows - 1) / min_rows;
  if (gy < 1) gy = 1;
  const long long rpb = (M + gy - 1) / gy;
  FIBER_CUDA(launch_k(colsum_kernel, dim3(gx, static_cast<unsigned>(gy)), dim3(256), 0, stream, x, ld, M, N, out, scale, row_scale,
                                                                        rps > 0 ? rps : 1, rpb, cw));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int dot_dispatch(const bf16* a, long long lda, const bf16* b, long long ldb, long long M, int N, float* out,
                 cudaStream_t stream) {
  FIBER_CHECK(N % 8 == 0 && M > 0, "dot: N must be a multiple of 8");
  int grid = ew_grid(M * (N / 8));
  if (grid > 2 * num_sms()) grid = 2 * num_sms();
  FIBER_CUDA(launch_k(dot_kernel, dim3(grid), dim3(256), 0, stream, a, lda, b, ldb, M, N, out));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int rowwise_scale_dispatch(const bf16* x, long long ldx, bf16* y, long long ldy, long long M, int N, int mode,
                           float p, unsigned long long seed, const float* row_scale, int rps, cudaStream_t stream) {
  FIBER_CHECK(N % 8 == 0 && M > 0, "rowwise op: N must be a multiple of 8");
  FIBER_CHECK(mode == 0 ? (p >= 0.f && p < 1.f) : row_scale != nullptr, "bad rowwise op arguments");
  FIBER_CUDA(launch_k(rowwise_scale_kernel, dim3(ew_grid(M * (N / 8))), dim3(256), 0, stream, x, ldx, y, ldy, M, N, mode, p, seed, row_scale,
                                                                rps > 0 ? rps : 1));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int grid_copy_dispatch(const bf16* src, bf16* dst, const bf16* add, const float* row_scale, int B, int Hs, int Ws, int Hd,
                       int Wd, int C, cudaStream_t stream) {
  FIBER_CHECK(src && dst && B > 0 && Hs > 0 && Ws > 0 && Hd > 0 && Wd > 0 && C > 0 && C % 8 == 0,
              "grid_copy: empty grid or C not a multiple of 8");
  FIBER_CUDA(launch_k(grid_copy_kernel, dim3(ew_grid(static_cast<long long>(B) * Hd * Wd * (C / 8))), dim3(256), 0, stream, src, dst,
                      add, row_scale, B, Hs, Ws, Hd, Wd, C));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int axpy_dispatch(const bf16* x, long long ldx, const bf16* add, long long lda, const float* alpha, bf16* out,
                  long long ldo, long long M, int N, cudaStream_t stream) {
  FIBER_CHECK(N % 8 == 0 && M > 0, "axpy: N must be a multiple of 8");
  FIBER_CUDA(launch_k(axpy_kernel, dim3(ew_grid(M * (N / 8))), dim3(256), 0, stream, x, ldx, add, lda, alpha, out, ldo, M, N));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int cast_dispatch(const float* x, bf16* y, long long n, cudaStream_t stream) {
  FIBER_CHECK(n > 0, "cast: empty");
  FIBER_CUDA(launch_k(cast_f32_bf16_kernel, dim3(ew_grid((n + 3) / 4)), dim3(256), 0, stream, x, y, n));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int cast_transpose_dispatch(const float* w, long long ldw, int N, int K, bf16* w_out, long long ld_out, bf16* wt_out,
                            long long ldt_out, cudaStream_t stream) {
  FIBER_CHECK(N > 0 && K > 0, "cast_transpose: empty");
  FIBER_CUDA(launch_k(cast_transpose_kernel, dim3((K + 31) / 32, (N + 31) / 32), dim3(256), 0, stream, w, ldw, N, K, w_out, ld_out, wt_out,
                                                                              ldt_out));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int patch_gather_dispatch(const float* img, bf16* out, int B, int RH, int R, cudaStream_t stream) {
  FIBER_CHECK(B > 0 && R > 0 && RH > 0 && R % 4 == 0 && RH % 4 == 0, "patch_gather: image height and width must be multiples of 4");
  FIBER_CUDA(launch_k(patch_gather_kernel, dim3(ew_grid(static_cast<long long>(B) * (RH / 4) * (R / 4) * 16)), dim3(256), 0, stream, img, out, B, RH, R));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int embed_dispatch(const long long* ids, int B, int L, int C, int pad, const float* word, const float* pos,
                   const float* type, bf16* out, long long ldo, cudaStream_t stream) {
  FIBER_CHECK(C % 4 == 0 && B > 0 && L > 0, "embed: width must be a multiple of 4");
  FIBER_CUDA(launch_k(embed_gather_kernel, dim3(ew_grid(static_cast<long long>(B) * L * 32)), dim3(256), 0, stream, ids, B, L, C, pad, word, pos,
                                                                                      type, out, ldo));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int embed_scatter_dispatch(const long long* ids, int B, int L, int C, int pad, const bf16* dsum, long long ldd,
                           float* dword, float* dpos, cudaStream_t stream) {
  FIBER_CHECK(C % 2 == 0 && B > 0 && L > 0, "embed_scatter: width must be even");
  FIBER_CUDA(launch_k(embed_scatter_kernel, dim3(ew_grid(static_cast<long long>(B) * L * 32)), dim3(256), 0, stream, ids, B, L, C, pad, dsum, ldd,
                                                                                       dword, dpos));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace fiber
