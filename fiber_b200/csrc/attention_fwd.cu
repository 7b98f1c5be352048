// fiber_b200 — attention forward (flash-style, never materialises the score matrix in HBM).
// See attention.cuh for the two addressing modes.  Each warp owns 16 query rows; K/V are staged
// in shared memory 144 keys at a time and consumed in 48-key register tiles with an online
// softmax; legacy mma.sync m16n8k16 bf16 tensor-core path (the attention core is MUFU/ALU-bound at
// head_dim 32).
#include "attention.cuh"
#include "../../include/fiber_b200.h"

namespace fiber {

void count_launch(int n = 1);

template <int HD, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) attn_fwd_kernel(const AttnParams p) {
  constexpr int QROWS = 16 * NWARPS;
  constexpr int PITCH = HD + 8;
  constexpr int CPR = HD / 8;  // 16-byte chunks per row
  extern __shared__ __align__(16) uint8_t smem[];
  bf16* sQ = reinterpret_cast<bf16*>(smem);
  bf16* sK = sQ + QROWS * PITCH;
  bf16* sV = sK + ATT_SKEYS * PITCH;
  float* sMask = reinterpret_cast<float*>(sV + ATT_SKEYS * PITCH);  // [ATT_SKEYS]
  float* sTbl = sMask + ATT_SKEYS;                                  // window only
  int* sRow = reinterpret_cast<int*>(sTbl + ATT_MAXTBL);
  // bias index = A_q - sB[j] with A_q = (qi/ws)*(2ws-1) + qi%ws + (ws-1)*2ws, sB[j] = (j/ws)*(2ws-1) + j%ws
  int16_t* sB = reinterpret_cast<int16_t*>(sRow + ATT_MAXTOK);
  uint8_t* sRid = reinterpret_cast<uint8_t*>(sB + ATT_MAXTOK);

  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * QROWS, h = blockIdx.y, g = blockIdx.z;
  const bool window = p.mode == 1;
  const int Lq = p.Lq, Lk = p.Lk;
  long long qbase = static_cast<long long>(g) * Lq, kbase = static_cast<long long>(g) * Lk;
  const int ws = p.ws, tw2 = 2 * ws - 1;
  bool has_mask = false;  // only windows in the last window row / column contain masked pairs

  if (window) {
    const int nWw = p.W / ws, nW = (p.H / ws) * nWw;
    const int b = g / nW, w = g % nW, wh = w / nWw, ww = w % nWw;
    has_mask = p.shift > 0 && (wh == p.H / ws - 1 || ww == nWw - 1);
    for (int i = tid; i < ATT_MAXTOK; i += blockDim.x) {
      int row = 0, bidx = 0, rid = 0;
      if (i < Lq) {
        const int th = i / ws, tw = i % ws;
        const int hp = wh * ws + th, wp = ww * ws + tw;
        row = b * p.H * p.W + ((hp + p.shift) % p.H) * p.W + (wp + p.shift) % p.W;
        rid = 3 * ((hp >= p.H - ws) + (hp >= p.H - p.shift)) + (wp >= p.W - ws) + (wp >= p.W - p.shift);
        bidx = th * tw2 + tw;
      }
      sRow[i] = row; sB[i] = static_cast<int16_t>(bidx); sRid[i] = rid;
    }
    for (int t = tid; t < tw2 * tw2; t += blockDim.x) sTbl[t] = p.bias_table[t * p.nH + h];
    __syncthreads();
  }

  // ---- Q tile -> smem -> A fragments ----
  for (int c = tid; c < QROWS * CPR; c += blockDim.x) {
    const int r = c / CPR, cc = c % CPR, qi = q0 + r;
    const bool valid = qi < Lq;
    const long long grow = valid ? (window ? sRow[qi] : qbase + qi) : 0;
    cp_async16(smem_u32(sQ + r * PITCH + cc * 8), p.q + grow * p.ldq + h * HD + cc * 8, valid);
  }
  cp_async_wait_all();
  __syncthreads();
  uint32_t qf[HD / 16][4];
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks) {
    const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
    const int col = ks * 16 + (lane >> 4) * 8;
    ldsm_x4(smem_u32(sQ + row * PITCH + col), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
  }

  float m_run[2] = {-1e30f, -1e30f}, l_run[2] = {0.f, 0.f};
  float oacc[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;

  const int r_lo = lane >> 2;           // fragment row (and r_lo + 8)
  const int qi0 = q0 + warp * 16 + r_lo;  // global query index of fragment row 0
  const float keep_inv = p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.0f;

  for (int kc0 = 0; kc0 < Lk; kc0 += ATT_SKEYS) {
    const int nk = min(ATT_SKEYS, Lk - kc0);
    const int nk_pad = ((nk + ATT_KCHUNK - 1) / ATT_KCHUNK) * ATT_KCHUNK;
    __syncthreads();
    for (int c = tid; c < nk_pad * CPR; c += blockDim.x) {
      const int r = c / CPR, cc = c % CPR, kj = kc0 + r;
      const bool valid = kj < Lk;
      const long long grow = valid ? (window ? sRow[kj] : kbase + kj) : 0;
      cp_async16(smem_u32(sK + r * PITCH + cc * 8), p.k + grow * p.ldk + h * HD + cc * 8, valid);
      cp_async16(smem_u32(sV + r * PITCH + cc * 8), p.v + grow * p.ldv + h * HD + cc * 8, valid);
    }
    if (!window) {
      for (int j = tid; j < nk_pad; j += blockDim.x)
        sMask[j] = (p.key_mask && kc0 + j < Lk) ? p.key_mask[kbase + kc0 + j] : 0.f;
    }
    cp_async_wait_all();
    __syncthreads();

    for (int sub = 0; sub < nk_pad / ATT_KCHUNK; ++sub) {
      float s[6][4];
#pragma unroll
      for (int i = 0; i < 6; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
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
        }
      }
      // ---- scale + bias + mask, online softmax ----
      float mx[2] = {-1e30f, -1e30f};
      int aq[2] = {0, 0}, ridq[2] = {0, 0};
      if (window) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int qi = qi0 + r * 8;
          const int qq = qi < Lq ? qi : 0;
          aq[r] = sB[qq] + (ws - 1) * (tw2 + 1);
          ridq[r] = sRid[qq];
        }
      }
#pragma unroll
      for (int nt = 0; nt < 6; ++nt) {
        const int jl0 = sub * ATT_KCHUNK + nt * 8 + (lane & 3) * 2;
        int bj[2] = {0, 0}, ridj[2] = {0, 0};
        float mk[2] = {0.f, 0.f};
        if (window) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int jj = kc0 + jl0 + c < Lk ? kc0 + jl0 + c : 0;
            bj[c] = sB[jj];
            if (has_mask) ridj[c] = sRid[jj];
          }
        } else {
          mk[0] = sMask[jl0]; mk[1] = sMask[jl0 + 1];
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = kc0 + jl0 + (e & 1);
          float v = s[nt][e] * p.scale;
          if (window) {
            v += sTbl[aq[e >> 1] - bj[e & 1]];
            if (has_mask && ridq[e >> 1] != ridj[e & 1]) v += -100.0f;
          } else {
            v += mk[e & 1];
          }
          if (j >= Lk) v = -1e30f;
          s[nt][e] = v;
          mx[e >> 1] = fmaxf(mx[e >> 1], v);
        }
      }
      float corr[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        const float m_new = fmaxf(m_run[r], mx[r]);
        corr[r] = fast_exp(m_run[r] - m_new);
        m_run[r] = m_new;
        l_run[r] *= corr[r];
      }
#pragma unroll
      for (int nt = 0; nt < 6; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float pv = fast_exp(s[nt][e] - m_run[e >> 1]);
          l_run[e >> 1] += pv;
          if (p.drop_p > 0.f) {
            const int j = kc0 + sub * ATT_KCHUNK + nt * 8 + (lane & 3) * 2 + (e & 1);
            const int qi = qi0 + (e >> 1) * 8;
            const unsigned long long idx =
                ((static_cast<unsigned long long>(g) * p.nH + h) * Lq + qi) * Lk + j;
            pv = dropout_keep(p.seed, idx, p.drop_p) ? pv * keep_inv : 0.f;
          }
          s[nt][e] = pv;
        }
      }
#pragma unroll
      for (int dt = 0; dt < HD / 8; ++dt) {
        oacc[dt][0] *= corr[0]; oacc[dt][1] *= corr[0];
        oacc[dt][2] *= corr[1]; oacc[dt][3] *= corr[1];
      }
#pragma unroll
      for (int kk = 0; kk < 3; ++kk) {
        uint32_t a[4];
        a[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
        a[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
        a[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        a[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int dt2 = 0; dt2 < HD / 16; ++dt2) {
          const int row = sub * ATT_KCHUNK + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          const int col = dt2 * 16 + (lane >> 4) * 8;
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(smem_u32(sV + row * PITCH + col), b0, b1, b2, b3);
          mma16816(oacc[2 * dt2], a, b0, b1);
          mma16816(oacc[2 * dt2 + 1], a, b2, b3);
        }
      }
    }
  }

  // ---- finalise: normalise, LSE, store through this warp's own sQ rows (coalesced) ----
  float inv[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float l = l_run[r];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    inv[r] = 1.0f / l;
    const int qi = qi0 + r * 8;
    if ((lane & 3) == 0 && qi < Lq && p.lse)
      p.lse[(static_cast<long long>(g) * p.nH + h) * Lq + qi] = m_run[r] + __logf(l);
  }
  __syncwarp();
#pragma unroll
  for (int dt = 0; dt < HD / 8; ++dt) {
    const int col = dt * 8 + (lane & 3) * 2;
    *reinterpret_cast<uint32_t*>(sQ + (warp * 16 + r_lo) * PITCH + col) =
        pack_bf16(oacc[dt][0] * inv[0], oacc[dt][1] * inv[0]);
    *reinterpret_cast<uint32_t*>(sQ + (warp * 16 + r_lo + 8) * PITCH + col) =
        pack_bf16(oacc[dt][2] * inv[1], oacc[dt][3] * inv[1]);
  }
  __syncwarp();
  for (int c = lane; c < 16 * CPR; c += 32) {
    const int r = c / CPR, cc = c % CPR, qi = q0 + warp * 16 + r;
    if (qi < Lq) {
      const long long grow = window ? sRow[qi] : qbase + qi;
      *reinterpret_cast<uint4*>(p.o + grow * p.ldo + h * HD + cc * 8) =
          *reinterpret_cast<const uint4*>(sQ + (warp * 16 + r) * PITCH + cc * 8);
    }
  }
}

template <int HD, int NWARPS>
static int launch_fwd(const AttnParams& p, cudaStream_t stream) {
  constexpr int QROWS = 16 * NWARPS;
  constexpr int PITCH = HD + 8;
  const size_t smem = (QROWS + 2 * ATT_SKEYS) * PITCH * 2 + ATT_SKEYS * 4 + ATT_MAXTBL * 4 +
                      ATT_MAXTOK * 4 + 3 * ATT_MAXTOK + 16;  // sRow + sB (int16) + sRid
  auto kern = attn_fwd_kernel<HD, NWARPS>;
  static bool attr_set = false;
  if (!attr_set) {
    FIBER_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid((p.Lq + QROWS - 1) / QROWS, p.nH, p.mode == 1 ? p.G * (p.H / p.ws) * (p.W / p.ws) : p.G);
  FIBER_CUDA(launch_k(kern, grid, dim3(NWARPS * 32), smem, stream, p));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int attn_check(const AttnParams& p, int hd) {
  FIBER_CHECK(hd == 32 || hd == 64, "head_dim must be 32 or 64 (got %d)", hd);
  FIBER_CHECK(p.G > 0 && p.nH > 0 && p.Lq > 0 && p.Lk > 0, "bad attention shape");
  FIBER_CHECK(p.ldq % 8 == 0 && p.ldk % 8 == 0 && p.ldv % 8 == 0 && p.ldo % 8 == 0,
              "attention row strides must be multiples of 8 elements");
  if (p.mode == 1) {
    FIBER_CHECK(p.ws > 0 && p.H % p.ws == 0 && p.W % p.ws == 0, "window size must divide H and W");
    FIBER_CHECK(p.Lq == p.ws * p.ws && p.Lk == p.Lq, "window mode needs Lq == Lk == ws*ws");
    FIBER_CHECK(p.Lq <= ATT_MAXTOK && (2 * p.ws - 1) * (2 * p.ws - 1) <= ATT_MAXTBL, "window too large");
    FIBER_CHECK(p.shift >= 0 && p.shift < p.ws, "shift must be in [0, ws)");
    FIBER_CHECK(p.bias_table != nullptr, "window mode needs the relative position bias table");
  }
  FIBER_CHECK(p.drop_p >= 0.f && p.drop_p < 1.f, "dropout p must be in [0,1)");
  return 0;
}

bool win_attn_supported(const AttnParams& p, int hd);             // window_attn.cu
int launch_win_fwd(const AttnParams& p, cudaStream_t stream);  // window_attn.cu
bool win_attn_tc_supported(const AttnParams& p, int hd, bool bwd);  // window_attn_tc.cu
int launch_win_tc_fwd(const AttnParams& p, cudaStream_t stream);    // window_attn_tc.cu
int option_winattn_tc();                                            // capi.cu
bool attn_sk_supported(const AttnParams& p, int hd);                // attention_sk.cu
int launch_attn_sk_fwd(const AttnParams& p, int hd, cudaStream_t stream);
bool attn_pk_shape(const AttnParams& p);                            // attention_sk.cu
int option_attn_sk();                                               // capi.cu
void count_attn_sk_launch();

int attn_fwd_dispatch(const AttnParams& p, int hd, cudaStream_t stream) {
  if (attn_check(p, hd)) return -1;
  // tcgen05 + TMA forward for at most 64 keys per group.  Bit 0: where a 128-query tile is mostly full (image -> text:
  // 576 / 144 / 1296 queries per sample; measured 0.255 -> 0.158 ms at stage 2); bit 2: self-attention shapes that pack
  // two or three samples into a tile (RoBERTa at 40 / 48 / 64 tokens: as fast as mma.sync forward, 20 % faster backward);
  // bit 4 (off by default): every other short query sequence, unpacked (40 of 128 tile rows used: slower than mma.sync)
  const int sk_opt = option_attn_sk();
  if (attn_sk_supported(p, hd) &&
      (((sk_opt & 1) && p.Lq >= 96) || ((sk_opt & 4) && p.Lq < 96 && (attn_pk_shape(p) || (sk_opt & 16))))) {
    count_attn_sk_launch();
    return launch_attn_sk_fwd(p, hd, stream);
  }
  if ((option_winattn_tc() & 1) && win_attn_tc_supported(p, hd, false)) return launch_win_tc_fwd(p, stream);  // opt-in
  if (win_attn_supported(p, hd)) return launch_win_fwd(p, stream);  // ws*ws <= 144 tokens, head_dim 32
  if (hd == 32) {
    if (p.Lq <= 48) return launch_fwd<32, 3>(p, stream);
    if (p.Lq <= 64) return launch_fwd<32, 4>(p, stream);
    return launch_fwd<32, 9>(p, stream);
  }
  if (p.Lq <= 48) return launch_fwd<64, 3>(p, stream);
  if (p.Lq <= 64) return launch_fwd<64, 4>(p, stream);
  return launch_fwd<64, 9>(p, stream);
}

}  // namespace fiber
