// fiber_b200 — shared-memory layouts and UMMA descriptors of the tcgen05 window-attention kernels
// (window_attn_tc.cu).  Everything here is plain integer arithmetic, __host__ __device__, so that the CPU test
// tests/native/winattn_tc_layout_check.cu can replay every tcgen05.mma operand fetch of the kernels against the
// canonical UMMA layouts (cute/atom/mma_traits_sm100.hpp, "make_umma_desc") without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TCL_HD __host__ __device__ __forceinline__
#else
#define TCL_HD inline
#endif

namespace fiber {
namespace tcl {

constexpr int N = 144;                    // tokens per window (12 x 12)
constexpr int WS = 12;
constexpr int HD = 32;                    // head dim
constexpr int TILE = N * 64;              // bytes of one [144][32] bf16 tile (64-byte rows, SWIZZLE_64B)
constexpr int F_PCHUNK = 128 * 128;       // forward P chunk: 128 query rows x 64 keys (128-byte rows, SWIZZLE_128B)
constexpr int B_PCHUNK = N * 128;         // backward P / dS chunk: 144 query rows x 64 keys
constexpr uint32_t LAYOUT_SW128 = 2, LAYOUT_SW64 = 4;  // UMMA::LayoutType

// byte offset of 16-byte piece `piece` (8 bf16) of row `row`
TCL_HD uint32_t sw64_off(int row, int piece) { return row * 64 + ((piece ^ ((row >> 1) & 3)) << 4); }
TCL_HD uint32_t sw128_off(int row, int piece) { return row * 128 + ((piece ^ (row & 7)) << 4); }
// P / dS element pair (key j0, j0 + 1), j0 even, of query row `row`: three 64-key chunks `chunk_bytes` apart
TCL_HD uint32_t pds_off(int row, int j0, int chunk_bytes) {
  return (j0 >> 6) * chunk_bytes + sw128_off(row, (j0 & 63) >> 3) + (j0 & 7) * 2;
}
// 16-byte piece p8 = key / 8 (0..17) of query row `row`
TCL_HD uint32_t pds_piece_off(int row, int p8, int chunk_bytes) { return (p8 >> 3) * chunk_bytes + sw128_off(row, p8 & 7); }

// UMMA shared-memory matrix descriptor (sm_100, version 1)
TCL_HD uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}

// ---- operand descriptors; `step` = index of the K = 16 MMA step ----
// K-major use of a [144][32] tile (A or B of S = Q K^T and dP = dO V^T; K = head dim, two steps):
//   rows 64 B apart, 8-row groups 512 B apart (SBO), step = +32 B inside the swizzled row
TCL_HD uint64_t desc_tile_kmajor(uint32_t tile_addr, int step) { return umma_desc(tile_addr + step * 32, 16, 512, LAYOUT_SW64); }
// MN-major B use of a [144][32] tile (V in P V, dO in P^T dO, Q in dS^T Q, K in dS K; K = tokens, nine steps):
//   the 32 head-dim elements are the contiguous MN extent (one 64-byte block: LBO unused), 8 token rows = one
//   512-byte atom (SBO), step of 16 tokens = +1024 B
TCL_HD uint64_t desc_tile_mnmajor(uint32_t tile_addr, int step) { return umma_desc(tile_addr + step * 1024, 512, 512, LAYOUT_SW64); }
// K-major A use of the P / dS chunks (P V, dS K; M = queries 0..127, K = keys, nine steps): chunk = step / 4,
//   +32 B per step inside the chunk, 8-row groups 1024 B apart
TCL_HD uint64_t desc_pds_kmajor(uint32_t base, int step, int chunk_bytes) {
  return umma_desc(base + (step >> 2) * chunk_bytes + (step & 3) * 32, 16, 1024, LAYOUT_SW128);
}
// MN-major A use of the backward P / dS chunks (P^T dO, dS^T Q; M = keys 0..127 = chunks 0 and 1, K = the 144
//   query rows, nine steps): LBO = chunk stride (next 64 keys), SBO = 1024 (8 query rows), step = +2048 B
TCL_HD uint64_t desc_pds_mnmajor(uint32_t base, int step) { return umma_desc(base + step * 2048, B_PCHUNK, 1024, LAYOUT_SW128); }

}  // namespace tcl
}  // namespace fiber
