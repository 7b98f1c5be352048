// fiber_b200 — definitions shared by the Swin window-attention kernels (window_attn.cu: mma.sync generation;
// window_attn_tc.cu: tcgen05 generation): tile constants, the static per-token tables of the relative-position
// bias / SW-MSA mask (swin_transformer.py:165-176, :363-387) and the closed-form cyclic-shift + partition map.
#pragma once
#include "attention.cuh"

namespace fiber {
namespace {

constexpr int WA_ROWS = 144;               // tokens per window tile (9 MMA row tiles)
constexpr int WA_HD = 32;
constexpr int WA_PITCH = WA_HD + 8;        // bf16 elements per tile row (80 B: conflict-free ldmatrix)
constexpr int WA_TILE = WA_ROWS * WA_PITCH;  // elements per Q / K / V / dO tile
constexpr int WA_MAXTBL = 23 * 23;         // (2*12-1)^2
constexpr int WA_SP = WA_ROWS + 8;         // pitch of the P / dS tiles (elements)
constexpr float WA_LOG2E = 1.4426950408889634f;
constexpr float WA_LN2 = 0.6931471805599453f;
constexpr float WA_MASK2 = -100.0f * WA_LOG2E;

__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cp_async4(uint32_t smem_addr, const void* gptr, bool valid) {
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_addr), "l"(gptr), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Static (window-independent) per-token tables shared by the forward and the backward kernel.
struct WinTables {
  float* tbl2;  // [2 * WA_MAXTBL]: log2e * bias table of this head, then -1e30 sentinels
  int* aq4;     // [144] byte offset of A_i = (i/ws)*(2ws-1) + i%ws + (ws-1)*2ws   (query side)
  int* bj4;     // [144] byte offset of B_j = (j/ws)*(2ws-1) + j%ws; padded keys: -4*WA_MAXTBL
  int* code;    // [144] SW-MSA region code, see the file header
  int* tok;     // [144] th | tw << 8
};

__device__ __forceinline__ void fill_tables(const WinTables& t, const float* __restrict__ bias_table, int nH, int h,
                                            int ws, int shift, int N, int tid, int nthreads) {
  const int tw2 = 2 * ws - 1;
  for (int i = tid; i < WA_MAXTBL; i += nthreads) {
    t.tbl2[i] = i < tw2 * tw2 ? bias_table[i * nH + h] * WA_LOG2E : 0.f;
    t.tbl2[WA_MAXTBL + i] = -1e30f;
  }
  for (int i = tid; i < WA_ROWS; i += nthreads) {
    const bool valid = i < N;
    const int th = valid ? i / ws : 0, tw = valid ? i % ws : 0;
    const int bidx = th * tw2 + tw;
    t.aq4[i] = 4 * (bidx + (ws - 1) * (tw2 + 1));
    t.bj4[i] = valid ? 4 * bidx : -4 * WA_MAXTBL;
    t.code[i] = valid ? ((th >= ws - shift) ? 1 : 0) | ((tw >= ws - shift) ? 2 : 0) : 0;
    t.tok[i] = th | (tw << 8);
  }
}

struct WinGeo {
  int H, W, ws, shift, nWw, nWh, nW;
  // window g -> (image row base, wh*ws + shift, ww*ws + shift, mask enable bits)
  __device__ __forceinline__ void decode(int g, long long& img_base, int& h0, int& w0, int& emask) const {
    const int b = g / nW, w = g - b * nW;
    const int wh = w / nWw, ww = w - wh * nWw;
    img_base = static_cast<long long>(b) * H * W;
    h0 = wh * ws + shift;
    w0 = ww * ws + shift;
    emask = shift > 0 ? ((wh == nWh - 1) ? 1 : 0) | ((ww == nWw - 1) ? 2 : 0) : 0;
  }
  // activation row of the token at window coordinates (th, tw)
  __device__ __forceinline__ long long row(long long img_base, int h0, int w0, int th, int tw) const {
    int hp = h0 + th, wp = w0 + tw;
    hp -= hp >= H ? H : 0;
    wp -= wp >= W ? W : 0;
    return img_base + hp * W + wp;
  }
};

}  // namespace
}  // namespace fiber
