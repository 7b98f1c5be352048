// fiber_b200 — attention backward (recomputes P from Q, K and the saved log-sum-exp).
//
// One CTA (9 warps) owns all queries and keys of one (group, head) problem, so dQ/dK/dV need no
// atomics:   per 144-key chunk { per 144-query chunk {
//     phase A (warp = 16 query rows): S = QK^T, P = exp(S - lse), dP = dO V^T, dS = P*(dP - D),
//                                     dQ += dS K ; P and dS go to shared memory as bf16
//     phase B (warp = 16 key rows)  : dV += P^T dO, dK += dS^T Q  (transposed ldmatrix reads) } }
// WINDOW mode additionally reduces d(relative-position-bias table): every thread owns fixed (i,j)
// score positions, so dS is summed over all windows a CTA processes in registers and flushed once
// (shared-memory atomics -> one global atomic per table entry per CTA).
#include <algorithm>

#include "attention.cuh"
#include "../../include/fiber_b200.h"

namespace fiber {

void count_launch(int n = 1);
int attn_check(const AttnParams& p, int hd);

// NW warps = 16 * NW query rows per chunk, SK keys staged per chunk.  <9, 144> is the general configuration
// (one CTA per SM); <3, 48> serves problems with at most 48 queries and 48 keys (RoBERTa self-attention at 40
// tokens: 3072 (sample, head) problems per launch) with four CTAs per SM instead of one two-thirds-idle CTA.
// <4, 48> serves few-key problems with many queries (i2t cross attention: 576 / 144 queries x 40 keys) with three
// CTAs per SM that walk 64-query chunks.
template <int HD, bool WINDOW, int NW = 9, int SK = ATT_SKEYS>
__global__ void __launch_bounds__(NW * 32, NW == 9 ? 1 : (NW == 3 ? 4 : 3)) attn_bwd_kernel(const AttnParams p) {
  constexpr int BW_NWARPS = NW;
  constexpr int BW_QROWS = 16 * NW;
  constexpr int ATT_SKEYS = SK;        // shadows the global constant inside this kernel
  constexpr int BW_SP = SK + 8;        // pitch of the P / dS tiles (elements)
  static_assert(SK % ATT_KCHUNK == 0 && SK / 16 <= NW, "one key tile per warp in phase B");
  static_assert(!WINDOW || SK == 144, "window mode sizes its d(bias) registers for 144-key chunks");
  constexpr int PITCH = HD + 8;
  constexpr int CPR = HD / 8;
  extern __shared__ __align__(16) uint8_t smem[];
  bf16* sQ = reinterpret_cast<bf16*>(smem);
  bf16* sdO = sQ + BW_QROWS * PITCH;
  bf16* sK = sdO + BW_QROWS * PITCH;
  bf16* sV = sK + ATT_SKEYS * PITCH;
  bf16* sP = sV + ATT_SKEYS * PITCH;
  bf16* sdS = sP + BW_QROWS * BW_SP;
  float* sLse = reinterpret_cast<float*>(sdS + BW_QROWS * BW_SP);
  float* sD = sLse + BW_QROWS;
  float* sMask = sD + BW_QROWS;
  float* sTbl = sMask + ATT_SKEYS;   // window only: bias table column of head h
  float* sdTbl = sTbl + ATT_MAXTBL;  // window only: d(table) accumulator
  int* sRow = reinterpret_cast<int*>(sdTbl + ATT_MAXTBL);
  uint8_t* sTh = reinterpret_cast<uint8_t*>(sRow + ATT_MAXTOK);
  uint8_t* sTw = sTh + ATT_MAXTOK;
  uint8_t* sRid = sTw + ATT_MAXTOK;
  // fp32 dQ accumulator [nqc*144][HD], only when both the query and the key loop have several chunks
  float* sDQ = reinterpret_cast<float*>(smem + (((sRid + ATT_MAXTOK) - smem + 15) & ~static_cast<ptrdiff_t>(15)));

  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x;
  const int Lq = p.Lq, Lk = p.Lk;
  const int ws = p.ws, tw2 = 2 * ws - 1;
  const int nWw = WINDOW ? p.W / ws : 1, nW = WINDOW ? (p.H / ws) * nWw : 1;
  const int n_groups = WINDOW ? p.G * nW : p.G;
  const int nkc = (Lk + ATT_SKEYS - 1) / ATT_SKEYS;
  const int nqc = (Lq + BW_QROWS - 1) / BW_QROWS;
  const float keep_inv = p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.0f;
  const int r_lo = lane >> 2;
  const bool acc_dq_smem = nqc > 1 && nkc > 1;
  const bool db_regs = Lq <= BW_QROWS;  // fixed (i,j) ownership only with a single query/key chunk
  // Several chunks (18 x 18 windows: 324 tokens = 3 x 3 tiles of 144): d(bias) as DIAGONAL SUMS of the dS tile.  With
  // tiles that hold whole window rows (144 % ws == 0), the ws x ws block of dS between query window-row QR and key
  // window-row KR contributes the sum of its diagonal qw - kw = d to table entry (QR - KR, d).  Warp w owns the table
  // rows (QR - KR + ws - 1) % 9 == w — at most one block per query window-row and tile, 7-8 blocks per warp and full
  // tile, and nobody else ever touches those rows.  Lane L < ws walks a block with a skew, column (L + qw) mod ws of
  // row qw, so that it only ever meets the diagonals d = -L (before the wrap) and d = ws - L (after it): two running
  // sums per owned table row, at most four rows per warp, kept in registers over all windows of the CTA and written
  // once.  No atomics (the generic path's per-element shared-memory float atomics compile to compare-and-swap spin
  // loops: 105 k contended ones per window-head, two thirds of the 576-px configuration's step).
  const int dg_rows = WINDOW ? BW_QROWS / ws : 1;
  const bool db_diag = WINDOW && !db_regs && BW_NWARPS == 9 && BW_QROWS % ws == 0 && dg_rows <= 9 && ws <= 32 &&
                       2 * ws - 1 <= 36 && Lq == ws * ws && Lk == Lq;
  float dbacc[WINDOW ? 18 : 1][4];
  if (WINDOW) {
#pragma unroll
    for (int i = 0; i < 18; ++i) dbacc[i][0] = dbacc[i][1] = dbacc[i][2] = dbacc[i][3] = 0.f;
    for (int t = tid; t < tw2 * tw2; t += blockDim.x) {
      sTbl[t] = p.bias_table[t * p.nH + h];
      sdTbl[t] = 0.f;
    }
    for (int i = tid; i < ATT_MAXTOK; i += blockDim.x) {
      sTh[i] = i < Lq ? i / ws : 0;
      sTw[i] = i < Lq ? i % ws : 0;
    }
  }

  for (int g = blockIdx.y; g < n_groups; g += gridDim.y) {
    const long long qbase = static_cast<long long>(g) * Lq, kbase = static_cast<long long>(g) * Lk;
    __syncthreads();
    if (WINDOW) {
      const int b = g / nW, w = g % nW, wh = w / nWw, ww = w % nWw;
      for (int i = tid; i < ATT_MAXTOK; i += blockDim.x) {
        int row = 0, rid = 0;
        if (i < Lq) {
          const int hp = wh * ws + i / ws, wp = ww * ws + i % ws;
          row = b * p.H * p.W + ((hp + p.shift) % p.H) * p.W + (wp + p.shift) % p.W;
          rid = 3 * ((hp >= p.H - ws) + (hp >= p.H - p.shift)) + (wp >= p.W - ws) + (wp >= p.W - p.shift);
        }
        sRow[i] = row; sRid[i] = rid;
      }
      __syncthreads();
    }

    float dq[HD / 8][4];  // persists over key chunks when there is a single query chunk

    for (int kc = 0; kc < nkc; ++kc) {
      const int kc0 = kc * ATT_SKEYS;
      const int nk = min(ATT_SKEYS, Lk - kc0);
      const int nk_pad = ((nk + ATT_KCHUNK - 1) / ATT_KCHUNK) * ATT_KCHUNK;
      const int n_ktiles = (nk + 15) / 16;
      __syncthreads();
      for (int c = tid; c < nk_pad * CPR; c += blockDim.x) {
        const int r = c / CPR, cc = c % CPR, kj = kc0 + r;
        const bool valid = kj < Lk;
        const long long grow = valid ? (WINDOW ? sRow[kj] : kbase + kj) : 0;
        cp_async16(smem_u32(sK + r * PITCH + cc * 8), p.k + grow * p.ldk + h * HD + cc * 8, valid);
        cp_async16(smem_u32(sV + r * PITCH + cc * 8), p.v + grow * p.ldv + h * HD + cc * 8, valid);
      }
      if (!WINDOW) {
        for (int j = tid; j < nk_pad; j += blockDim.x)
          sMask[j] = (p.key_mask && kc0 + j < Lk) ? p.key_mask[kbase + kc0 + j] : 0.f;
      }

      float dkacc[HD / 8][4], dvacc[HD / 8][4];
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) {
        dkacc[i][0] = dkacc[i][1] = dkacc[i][2] = dkacc[i][3] = 0.f;
        dvacc[i][0] = dvacc[i][1] = dvacc[i][2] = dvacc[i][3] = 0.f;
      }

      for (int qc = 0; qc < nqc; ++qc) {
        const int q0 = qc * BW_QROWS;
        const int nq = min(BW_QROWS, Lq - q0);
        const int n_qtiles = (nq + 15) / 16;
        __syncthreads();  // previous phase B done with sQ/sdO/sP/sdS
        for (int c = tid; c < BW_QROWS * CPR; c += blockDim.x) {
          const int r = c / CPR, cc = c % CPR, qi = q0 + r;
          const bool valid = qi < Lq;
          const long long grow = valid ? (WINDOW ? sRow[qi] : qbase + qi) : 0;
          cp_async16(smem_u32(sQ + r * PITCH + cc * 8), p.q + grow * p.ldq + h * HD + cc * 8, valid);
          cp_async16(smem_u32(sdO + r * PITCH + cc * 8), p.d_o + grow * p.lddo + h * HD + cc * 8, valid);
        }
        // D_i = sum_d dO[i,d] * O[i,d]; lse_i
        for (int c = tid; c < BW_QROWS * CPR; c += blockDim.x) {
          const int r = c / CPR, cc = c % CPR, qi = q0 + r;
          float part = 0.f;
          if (qi < Lq) {
            const long long grow = WINDOW ? sRow[qi] : qbase + qi;
            const uint4 a = *reinterpret_cast<const uint4*>(p.o + grow * p.ldo + h * HD + cc * 8);
            const uint4 b = *reinterpret_cast<const uint4*>(p.d_o + grow * p.lddo + h * HD + cc * 8);
            const uint32_t* au = reinterpret_cast<const uint32_t*>(&a);
            const uint32_t* bu = reinterpret_cast<const uint32_t*>(&b);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 x = unpack_bf16(au[e]), y = unpack_bf16(bu[e]);
              part += x.x * y.x + x.y * y.y;
            }
          }
          // CPR (4 or 8) consecutive lanes hold one row
#pragma unroll
          for (int o = CPR / 2; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
          if (cc == 0) {
            sD[r] = part;
            sLse[r] = qi < Lq ? p.lse[(static_cast<long long>(g) * p.nH + h) * Lq + qi] : 0.f;
          }
        }
        cp_async_wait_all();
        __syncthreads();

        // ================= phase A =================
        if (warp < n_qtiles) {
          uint32_t qf[HD / 16][4], dof[HD / 16][4];
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks) {
            const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
            const int col = ks * 16 + (lane >> 4) * 8;
            ldsm_x4(smem_u32(sQ + row * PITCH + col), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
            ldsm_x4(smem_u32(sdO + row * PITCH + col), dof[ks][0], dof[ks][1], dof[ks][2], dof[ks][3]);
          }
          const int rl0 = warp * 16 + r_lo;  // local query row of fragment row 0
          const float lse0 = sLse[rl0], lse1 = sLse[rl0 + 8];
          const float D0 = sD[rl0], D1 = sD[rl0 + 8];
          if (kc == 0 || nqc > 1) {
#pragma unroll
            for (int i = 0; i < HD / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
          }
          if (acc_dq_smem && kc > 0) {  // continue the running sum of this query tile (thread-private slots)
#pragma unroll
            for (int i = 0; i < HD / 8; ++i) {
              const float* a0 = sDQ + (q0 + warp * 16 + r_lo) * HD + i * 8 + (lane & 3) * 2;
              dq[i][0] = a0[0]; dq[i][1] = a0[1]; dq[i][2] = a0[8 * HD]; dq[i][3] = a0[8 * HD + 1];
            }
          }
#pragma unroll
          for (int sub = 0; sub < ATT_SKEYS / ATT_KCHUNK; ++sub) {
            if (sub * ATT_KCHUNK < nk_pad) {
              float s[6][4], dp[6][4];
#pragma unroll
              for (int i = 0; i < 6; ++i) {
                s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
                dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
              }
#pragma unroll
              for (int ks = 0; ks < HD / 16; ++ks) {
#pragma unroll
                for (int nt2 = 0; nt2 < 3; ++nt2) {
                  const int row = sub * ATT_KCHUNK + nt2 * 16 + (lane & 7) + ((lane >> 4) << 3);
                  const int col = ks * 16 + ((lane >> 3) & 1) * 8;
                  uint32_t b0, b1, b2, b3;
                  ldsm_x4(smem_u32(sK + row * PITCH + col), b0, b1, b2, b3);
                  mma16816(s[2 * nt2], qf[ks], b0, b1);
                  mma16816(s[2 * nt2 + 1], qf[ks], b2, b3);
                  ldsm_x4(smem_u32(sV + row * PITCH + col), b0, b1, b2, b3);
                  mma16816(dp[2 * nt2], dof[ks], b0, b1);
                  mma16816(dp[2 * nt2 + 1], dof[ks], b2, b3);
                }
              }
#pragma unroll
              for (int nt = 0; nt < 6; ++nt) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int jl = sub * ATT_KCHUNK + nt * 8 + (lane & 3) * 2 + (e & 1);
                  const int j = kc0 + jl;
                  const int ql = rl0 + (e >> 1) * 8, qi = q0 + ql;
                  float v = s[nt][e] * p.scale;
                  if (WINDOW) {
                    const int qq = qi < Lq ? qi : 0;
                    const int jj = j < Lk ? j : 0;
                    v += sTbl[(static_cast<int>(sTh[qq]) - static_cast<int>(sTh[jj]) + ws - 1) * tw2 +
                              static_cast<int>(sTw[qq]) - static_cast<int>(sTw[jj]) + ws - 1];
                    if (p.shift > 0 && sRid[qq] != sRid[jj]) v += -100.0f;
                  } else {
                    v += sMask[jl];
                  }
                  float pr = (j < Lk) ? fast_exp(v - ((e >> 1) ? lse1 : lse0)) : 0.f;
                  float dpv = dp[nt][e];
                  float pd = pr;  // P after dropout (feeds dV)
                  if (p.drop_p > 0.f) {
                    const unsigned long long idx =
                        ((static_cast<unsigned long long>(g) * p.nH + h) * Lq + qi) * Lk + j;
                    const bool keep = dropout_keep(p.seed, idx, p.drop_p);
                    pd = keep ? pr * keep_inv : 0.f;
                    dpv = keep ? dpv * keep_inv : 0.f;
                  }
                  const float ds = pr * (dpv - ((e >> 1) ? D1 : D0));
                  s[nt][e] = pd;
                  dp[nt][e] = ds;
                  if (WINDOW) {
                    if (db_regs) {
                      dbacc[sub * 6 + nt][e] += ds;
                    } else if (!db_diag && qi < Lq && j < Lk) {
                      atomicAdd(&sdTbl[(static_cast<int>(sTh[qi]) - static_cast<int>(sTh[j]) + ws - 1) * tw2 +
                                       static_cast<int>(sTw[qi]) - static_cast<int>(sTw[j]) + ws - 1], ds);
                    }
                  }
                }
              }
              // P, dS -> smem (bf16) for phase B; dQ += dS K from registers
#pragma unroll
              for (int nt = 0; nt < 6; ++nt) {
                const int col = sub * ATT_KCHUNK + nt * 8 + (lane & 3) * 2;
                *reinterpret_cast<uint32_t*>(sP + rl0 * BW_SP + col) = pack_bf16(s[nt][0], s[nt][1]);
                *reinterpret_cast<uint32_t*>(sP + (rl0 + 8) * BW_SP + col) = pack_bf16(s[nt][2], s[nt][3]);
                *reinterpret_cast<uint32_t*>(sdS + rl0 * BW_SP + col) = pack_bf16(dp[nt][0], dp[nt][1]);
                *reinterpret_cast<uint32_t*>(sdS + (rl0 + 8) * BW_SP + col) = pack_bf16(dp[nt][2], dp[nt][3]);
              }
#pragma unroll
              for (int kk = 0; kk < 3; ++kk) {
                uint32_t a[4];
                a[0] = pack_bf16(dp[2 * kk][0], dp[2 * kk][1]);
                a[1] = pack_bf16(dp[2 * kk][2], dp[2 * kk][3]);
                a[2] = pack_bf16(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
                a[3] = pack_bf16(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
                for (int dt2 = 0; dt2 < HD / 16; ++dt2) {
                  const int row = sub * ATT_KCHUNK + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                  const int col = dt2 * 16 + (lane >> 4) * 8;
                  uint32_t b0, b1, b2, b3;
                  ldsm_x4_t(smem_u32(sK + row * PITCH + col), b0, b1, b2, b3);
                  mma16816(dq[2 * dt2], a, b0, b1);
                  mma16816(dq[2 * dt2 + 1], a, b2, b3);
                }
              }
            }
          }
        }
        __syncthreads();

        if (WINDOW && db_diag) {   // d(bias): skewed diagonal sums of this tile's dS (read-only on sdS, like phase B)
          const int rq = nq / ws, rk = nk / ws;
          const int offa = lane, offb = lane - ws, thr = ws - lane;   // column before / after the wrap, first wrapped row
          for (int qr = 0; qr < rq; ++qr) {
            // the one key window-row of this tile whose table row (QR - KR + ws - 1) this warp owns
            const int base = qc * dg_rows + qr - kc * dg_rows + ws - 1;        // QR - kc * rows + ws - 1 >= 0
            const int kr = ((base - warp) % 9 + 9) % 9;
            if (kr < rk && lane < ws) {
              const int ri = (base - kr) / 9;                                  // which of the warp's (up to four) rows
              const bf16* e0 = sdS + (qr * ws) * BW_SP + kr * ws;
              float a0 = 0.f, a1 = 0.f;
#pragma unroll 6
              for (int qw = 0; qw < ws; ++qw) {
                const bool wr = qw >= thr;
                const float v = __bfloat162float(e0[qw * BW_SP + (wr ? offb : offa) + qw]);
                a0 += wr ? 0.f : v;
                a1 += wr ? v : 0.f;
              }
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (i == ri) { dbacc[i][0] += a0; dbacc[i][1] += a1; }
            }
          }
        }

        // ================= phase B =================
        if (warp < n_ktiles) {
          const int m0 = warp * 16;  // key tile inside the smem chunk
          for (int kt = 0; kt < n_qtiles; ++kt) {
            const int k0 = kt * 16;
            uint32_t ap[4], as_[4];
            {
              const int row = k0 + (lane & 7) + ((lane >> 4) << 3);
              const int col = m0 + ((lane >> 3) & 1) * 8;
              ldsm_x4_t(smem_u32(sP + row * BW_SP + col), ap[0], ap[1], ap[2], ap[3]);
              ldsm_x4_t(smem_u32(sdS + row * BW_SP + col), as_[0], as_[1], as_[2], as_[3]);
            }
#pragma unroll
            for (int dt2 = 0; dt2 < HD / 16; ++dt2) {
              const int row = k0 + (lane & 7) + ((lane >> 3) & 1) * 8;
              const int col = dt2 * 16 + (lane >> 4) * 8;
              uint32_t b0, b1, b2, b3;
              ldsm_x4_t(smem_u32(sdO + row * PITCH + col), b0, b1, b2, b3);
              mma16816(dvacc[2 * dt2], ap, b0, b1);
              mma16816(dvacc[2 * dt2 + 1], ap, b2, b3);
              ldsm_x4_t(smem_u32(sQ + row * PITCH + col), b0, b1, b2, b3);
              mma16816(dkacc[2 * dt2], as_, b0, b1);
              mma16816(dkacc[2 * dt2 + 1], as_, b2, b3);
            }
          }
        }
        // ---- dQ: running sums in smem while more key chunks follow; store after the last ----
        if (acc_dq_smem && kc < nkc - 1) {
          if (warp < n_qtiles) {
#pragma unroll
            for (int i = 0; i < HD / 8; ++i) {
              float* a0 = sDQ + (q0 + warp * 16 + r_lo) * HD + i * 8 + (lane & 3) * 2;
              a0[0] = dq[i][0]; a0[1] = dq[i][1]; a0[8 * HD] = dq[i][2]; a0[8 * HD + 1] = dq[i][3];
            }
          }
        } else if (nqc > 1 || kc == nkc - 1) {
          __syncthreads();  // phase B finished reading sQ
          if (warp < n_qtiles) {
#pragma unroll
            for (int dt = 0; dt < HD / 8; ++dt) {
              const int col = dt * 8 + (lane & 3) * 2;
              *reinterpret_cast<uint32_t*>(sQ + (warp * 16 + r_lo) * PITCH + col) =
                  pack_bf16(dq[dt][0] * p.scale, dq[dt][1] * p.scale);
              *reinterpret_cast<uint32_t*>(sQ + (warp * 16 + r_lo + 8) * PITCH + col) =
                  pack_bf16(dq[dt][2] * p.scale, dq[dt][3] * p.scale);
            }
            __syncwarp();
            for (int c = lane; c < 16 * CPR; c += 32) {
              const int r = c / CPR, cc = c % CPR, qi = q0 + warp * 16 + r;
              if (qi < Lq) {
                const long long grow = WINDOW ? sRow[qi] : qbase + qi;
                *reinterpret_cast<uint4*>(p.dq + grow * p.lddq + h * HD + cc * 8) =
                    *reinterpret_cast<const uint4*>(sQ + (warp * 16 + r) * PITCH + cc * 8);
              }
            }
          }
        }
      }  // query chunks

      // ---- dK / dV of this key chunk: stage through this warp's rows of sK / sV ----
      __syncthreads();
      if (warp < n_ktiles) {
#pragma unroll
        for (int dt = 0; dt < HD / 8; ++dt) {
          const int col = dt * 8 + (lane & 3) * 2;
          *reinterpret_cast<uint32_t*>(sK + (warp * 16 + r_lo) * PITCH + col) =
              pack_bf16(dkacc[dt][0] * p.scale, dkacc[dt][1] * p.scale);
          *reinterpret_cast<uint32_t*>(sK + (warp * 16 + r_lo + 8) * PITCH + col) =
              pack_bf16(dkacc[dt][2] * p.scale, dkacc[dt][3] * p.scale);
          *reinterpret_cast<uint32_t*>(sV + (warp * 16 + r_lo) * PITCH + col) = pack_bf16(dvacc[dt][0], dvacc[dt][1]);
          *reinterpret_cast<uint32_t*>(sV + (warp * 16 + r_lo + 8) * PITCH + col) = pack_bf16(dvacc[dt][2], dvacc[dt][3]);
        }
        __syncwarp();
        for (int c = lane; c < 16 * CPR; c += 32) {
          const int r = c / CPR, cc = c % CPR, kj = kc0 + warp * 16 + r;
          if (kj < Lk) {
            const long long grow = WINDOW ? sRow[kj] : kbase + kj;
            *reinterpret_cast<uint4*>(p.dk + grow * p.lddk + h * HD + cc * 8) =
                *reinterpret_cast<const uint4*>(sK + (warp * 16 + r) * PITCH + cc * 8);
            *reinterpret_cast<uint4*>(p.dv + grow * p.lddv + h * HD + cc * 8) =
                *reinterpret_cast<const uint4*>(sV + (warp * 16 + r) * PITCH + cc * 8);
          }
        }
      }
    }  // key chunks
  }    // groups

  if (WINDOW) {
    // flush the register-resident d(bias) sums: smem atomics, then one global atomic per entry
    __syncthreads();
#pragma unroll
    for (int t = 0; t < (db_regs ? 18 : 0); ++t) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = t * 8 + (lane & 3) * 2 + (e & 1);
        const int qi = warp * 16 + r_lo + (e >> 1) * 8;
        if (qi < Lq && j < Lk)
          atomicAdd(&sdTbl[(static_cast<int>(sTh[qi]) - static_cast<int>(sTh[j]) + ws - 1) * tw2 +
                           static_cast<int>(sTw[qi]) - static_cast<int>(sTw[j]) + ws - 1],
                    dbacc[t][e]);
      }
    }
    if (db_diag && lane < ws) {   // rows warp, warp + 9, ... of the table: this warp is their only writer
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = warp + 9 * i;
        if (row < tw2) {
          sdTbl[row * tw2 + ws - 1 - lane] += dbacc[i][0];                       // d = -lane
          if (lane > 0) sdTbl[row * tw2 + 2 * ws - 1 - lane] += dbacc[i][1];     // d = ws - lane
        }
      }
    }
    __syncthreads();
    for (int t = tid; t < tw2 * tw2; t += blockDim.x) atomicAdd(&p.dbias_table[t * p.nH + h], sdTbl[t]);
  }
}


// shared memory the backward needs for (Lq, Lk) in a given configuration (same formula as launch_bwd)
template <int HD, int NW, int SK>
static size_t bwd_smem_bytes(const AttnParams& p) {
  constexpr int QROWS = 16 * NW, SP = SK + 8, PITCH = HD + 8;
  const int nqc = (p.Lq + QROWS - 1) / QROWS, nkc = (p.Lk + SK - 1) / SK;
  const size_t base = (2 * QROWS + 2 * SK) * PITCH * 2 + 2 * QROWS * SP * 2 + (2 * QROWS + SK) * 4 + 2 * ATT_MAXTBL * 4 +
                      ATT_MAXTOK * 4 + 3 * ATT_MAXTOK + 32;
  return base + ((nqc > 1 && nkc > 1) ? static_cast<size_t>(nqc) * QROWS * HD * 4 : 0);
}

template <int HD, bool WINDOW, int NW = 9, int SK = ATT_SKEYS>
static int launch_bwd(const AttnParams& p, cudaStream_t stream) {
  constexpr int BW_NWARPS = NW;
  constexpr int BW_QROWS = 16 * NW;
  constexpr int ATT_SKEYS = SK;
  constexpr int BW_SP = SK + 8;
  constexpr int PITCH = HD + 8;
  const int nqc = (p.Lq + BW_QROWS - 1) / BW_QROWS, nkc = (p.Lk + ATT_SKEYS - 1) / ATT_SKEYS;
  const size_t base = (2 * BW_QROWS + 2 * ATT_SKEYS) * PITCH * 2 + 2 * BW_QROWS * BW_SP * 2 +
                      (2 * BW_QROWS + ATT_SKEYS) * 4 + 2 * ATT_MAXTBL * 4 + ATT_MAXTOK * 4 +
                      3 * ATT_MAXTOK + 32;
  const size_t smem = base + ((nqc > 1 && nkc > 1) ? static_cast<size_t>(nqc) * BW_QROWS * HD * 4 : 0);
  FIBER_CHECK(smem <= 227 * 1024, "attention backward: Lq=%d with Lk=%d needs %zu bytes of shared memory", p.Lq, p.Lk, smem);
  auto kern = attn_bwd_kernel<HD, WINDOW, NW, SK>;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    FIBER_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  const int n_groups = WINDOW ? p.G * (p.H / p.ws) * (p.W / p.ws) : p.G;
  int gy = n_groups;
  if (WINDOW) {  // persistent over windows so d(bias) is flushed once per CTA
    int resident = 1;   // CTAs per SM at this shared-memory size: one wave, no tail
    FIBER_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, BW_NWARPS * 32, smem));
    const int target = std::max(1, (std::max(1, resident) * num_sms()) / p.nH);
    if (gy > target) gy = target;
  }
  dim3 grid(p.nH, gy);
  FIBER_CUDA(launch_k(kern, grid, dim3(BW_NWARPS * 32), smem, stream, p));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

bool win_attn_supported(const AttnParams& p, int hd);                        // window_attn.cu
int launch_win_bwd(const AttnParams& p, float* D, cudaStream_t stream);  // window_attn.cu
bool win_attn_tc_supported(const AttnParams& p, int hd, bool bwd);           // window_attn_tc.cu
int launch_win_tc_bwd(const AttnParams& p, float* D, cudaStream_t stream);   // window_attn_tc.cu
int option_winattn_tc();                                                     // capi.cu
int option_attn_small();                                                     // capi.cu
bool attn_sk_bwd_supported(const AttnParams& p, int hd);                     // attention_sk.cu
int launch_attn_sk_bwd(const AttnParams& p, int hd, cudaStream_t stream);    // attention_sk.cu
bool attn_pk_shape(const AttnParams& p);                                     // attention_sk.cu
int option_attn_sk();                                                        // capi.cu
void count_attn_sk_launch();

int attn_bwd_dispatch(const AttnParams& p, int hd, float* d_scratch, cudaStream_t stream) {
  if (attn_check(p, hd)) return -1;
  FIBER_CHECK(p.d_o && p.dq && p.dk && p.dv && p.lse && p.o, "attention backward needs o, lse, d_o, dq, dk, dv");
  if (p.mode == 1) {
    FIBER_CHECK(hd == 32, "window attention uses head_dim 32");
    FIBER_CHECK(p.dbias_table != nullptr, "window backward needs dbias_table");
    if ((option_winattn_tc() & 2) && d_scratch != nullptr && win_attn_tc_supported(p, hd, true))
      return launch_win_tc_bwd(p, d_scratch, stream);  // opt-in tcgen05 generation
    if (win_attn_supported(p, hd) && d_scratch != nullptr) return launch_win_bwd(p, d_scratch, stream);
    return launch_bwd<32, true>(p, stream);
  }
  // tcgen05 + TMA backward for at most 64 keys per group ("attn_sk" bit 1: >= 96 queries (i2t); bit 3: packed
  // self-attention shapes; bit 4: every other short query sequence, see attn_fwd_dispatch)
  const int sk_opt = option_attn_sk();
  if (attn_sk_bwd_supported(p, hd) &&
      (((sk_opt & 2) && p.Lq >= 96) || ((sk_opt & 8) && p.Lq < 96 && (attn_pk_shape(p) || (sk_opt & 16))))) {
    count_attn_sk_launch();
    return launch_attn_sk_bwd(p, hd, stream);
  }
  // opt-in: at most 48 queries and keys (RoBERTa self-attention at 40 tokens) on 3-warp CTAs, four per SM
  if (hd == 64 && p.Lq <= 48 && p.Lk <= 48 && (option_attn_small() & 1)) return launch_bwd<64, false, 3, 48>(p, stream);
  // opt-in (bit 2): few queries, many keys (t2i: 40 text queries x 576 / 144 image keys) on the same 3-warp CTAs,
  // walking 48-key chunks (the 48-row Q / dO tiles are re-read from L2 per chunk)
  if (hd == 64 && p.Lq <= 48 && p.Lk > 48 && (option_attn_small() & 4)) return launch_bwd<64, false, 3, 48>(p, stream);
  // opt-in (bit 1): few keys, many queries (i2t: 576 / 144 queries x 40 text tokens) on 4-warp CTAs, three per SM
  if (hd == 32 && p.Lk <= 48 && p.Lq > 48 && (option_attn_small() & 2)) return launch_bwd<32, false, 4, 48>(p, stream);
  // many queries AND many keys (fine-grained t2i: 256 query tokens x 4200 / 1050 image keys): the fp32 dQ accumulator of
  // all query chunks lives in shared memory next to the tiles; where the 9-warp / 144-key tiles leave no room for it, the
  // small-tile configurations do
  if (hd == 64 && bwd_smem_bytes<64, 9, ATT_SKEYS>(p) > 227 * 1024) {
    // 144 query rows per chunk with 48-key staging first (a third of the chunk iterations of the 48 x 48 form)
    if (bwd_smem_bytes<64, 9, 48>(p) <= 227 * 1024) return launch_bwd<64, false, 9, 48>(p, stream);
    if (bwd_smem_bytes<64, 3, 48>(p) <= 227 * 1024) return launch_bwd<64, false, 3, 48>(p, stream);
  }
  if (hd == 32 && bwd_smem_bytes<32, 9, ATT_SKEYS>(p) > 227 * 1024 && bwd_smem_bytes<32, 4, 48>(p) <= 227 * 1024)
    return launch_bwd<32, false, 4, 48>(p, stream);
  return hd == 32 ? launch_bwd<32, false>(p, stream) : launch_bwd<64, false>(p, stream);
}

}  // namespace fiber
