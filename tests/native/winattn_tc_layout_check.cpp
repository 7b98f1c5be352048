// CPU replay of every tcgen05.mma operand fetch of the tcgen05 window-attention kernels (fiber_b200/csrc/
// window_attn_tc.cu) — TEST INFRASTRUCTURE, built and run by tests/test_winattn_tc_layout.py with g++.
//
// The kernels' shared-memory writers and descriptor builders live in window_tc_layout.cuh as host/device
// functions; this program includes that header, fills a byte image of shared memory through the WRITER functions
// with a unique id per logical element, and then reads operands back the way the tensor core does: through an
// independent model of the canonical UMMA layouts (CUTLASS cute/atom/mma_traits_sm100.hpp, "make_umma_desc"):
//
//   K-major,  SWIZZLE_{32,64,128}B:  ((8,m),(T,2)) : ((row_bytes, SBO), (1, T))        T = 8 bf16 = 16 bytes
//   MN-major, SWIZZLE_{32,64,128}B:  ((row_bytes/16 x T, n),(8,k)) : ((1, LBO), (row_bytes, SBO))
//   followed by Swizzle<B,4,3> on the byte address (16-byte piece index ^= address bits [7, 7+B)).
//
// The model is anchored on the two descriptor forms gemm_sm100.cu uses, which are validated on hardware
// (K-major SWIZZLE_128B tiles written by TMA, and the MN-major wgrad operands).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../fiber_b200/csrc/window_tc_layout.cuh"

using namespace fiber::tcl;

static int g_fail = 0;
#define CHECK(cond, ...)                         \
  do {                                           \
    if (!(cond)) {                               \
      if (g_fail < 20) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } \
      ++g_fail;                                  \
    }                                            \
  } while (0)

struct Desc { uint32_t start, lbo, sbo, layout, version; };
static Desc decode(uint64_t d) {
  Desc r;
  r.start = static_cast<uint32_t>(d & 0x3FFF) << 4;
  r.lbo = static_cast<uint32_t>((d >> 16) & 0x3FFF) << 4;
  r.sbo = static_cast<uint32_t>((d >> 32) & 0x3FFF) << 4;
  r.version = static_cast<uint32_t>((d >> 46) & 3);
  r.layout = static_cast<uint32_t>(d >> 61);
  return r;
}
static int row_bytes(uint32_t layout) { return layout == 2 ? 128 : (layout == 4 ? 64 : (layout == 6 ? 32 : 16)); }
static uint32_t swizzle(uint32_t addr, uint32_t layout) {
  const int bits = layout == 2 ? 3 : (layout == 4 ? 2 : (layout == 6 ? 1 : 0));
  return addr ^ (((addr >> 7) & ((1u << bits) - 1)) << 4);
}
// element (r, k) of a K-major operand, k in [0, 16): the K = 16 slice of one tcgen05.mma
static uint32_t addr_kmajor(const Desc& d, int r, int k) {
  const uint32_t off = (r % 8) * row_bytes(d.layout) + (r / 8) * d.sbo + (k / 8) * 16 + (k % 8) * 2;
  return swizzle(d.start + off, d.layout);
}
// element (mn, k) of an MN-major operand
static uint32_t addr_mnmajor(const Desc& d, int mn, int k) {
  const int per_row = row_bytes(d.layout) / 2;
  const uint32_t off = (mn % per_row) * 2 + (mn / per_row) * d.lbo + (k % 8) * row_bytes(d.layout) + (k / 8) * d.sbo;
  return swizzle(d.start + off, d.layout);
}

static std::vector<uint8_t> smem(256 * 1024);
static uint16_t rd(uint32_t a) { uint16_t v; memcpy(&v, &smem[a], 2); return v; }
static void wr(uint32_t a, uint16_t v) { memcpy(&smem[a], &v, 2); }

static uint16_t id(int row, int col, int salt) { return static_cast<uint16_t>((row * 151 + col * 7 + salt * 9973 + 1) & 0xFFFF); }

// [144][32] tile written the way the loader does: 16-byte piece (row, piece) at sw64_off
static void fill_tile(uint32_t base, int salt) {
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < HD; ++c) wr(base + sw64_off(r, c / 8) + (c % 8) * 2, id(r, c, salt));
}
// P / dS chunks written the way the element-wise warps (16-byte pieces) and the remainder warps (4-byte pairs) do
static void fill_pds(uint32_t base, int rows, int chunk, int salt) {
  for (int r = 0; r < rows; ++r)
    for (int j = 0; j < N; ++j) {
      const uint32_t a = base + pds_piece_off(r, j / 8, chunk) + (j % 8) * 2;
      const uint32_t b = base + pds_off(r, j & ~1, chunk) + (j & 1) * 2;
      CHECK(a == b, "pds writers disagree at row %d key %d: %u vs %u", r, j, a, b);
      wr(a, id(r, j, salt));
    }
}

int main() {
  // ---- anchors: the two forms of gemm_sm100.cu (validated on hardware) ----
  {
    // K-major: TMA box of 64 bf16 x R rows, SWIZZLE_128B -> row r at r*128, 16-byte chunk c at c ^ (r & 7)
    const uint32_t base = 4096;
    for (int r = 0; r < 128; ++r)
      for (int c = 0; c < 64; ++c) wr(base + r * 128 + (((c / 8) ^ (r & 7)) << 4) + (c % 8) * 2, id(r, c, 1));
    for (int k4 = 0; k4 < 4; ++k4) {
      const Desc d = decode(umma_desc(base + k4 * 32, 16, 1024, LAYOUT_SW128));
      for (int r = 0; r < 128; ++r)
        for (int k = 0; k < 16; ++k) CHECK(rd(addr_kmajor(d, r, k)) == id(r, k4 * 16 + k, 1), "anchor K-major r=%d k=%d", r, k);
    }
    // MN-major: TMA boxes of 64 MN x 64 k-rows (8192 B each), operand M = 128 = two boxes
    for (int j = 0; j < 2; ++j)
      for (int kr = 0; kr < 64; ++kr)
        for (int c = 0; c < 64; ++c)
          wr(base + j * 8192 + kr * 128 + (((c / 8) ^ (kr & 7)) << 4) + (c % 8) * 2, id(kr, j * 64 + c, 2));
    for (int k4 = 0; k4 < 4; ++k4) {
      const Desc d = decode(umma_desc(base + k4 * 2048, 8192, 1024, LAYOUT_SW128));
      for (int mn = 0; mn < 128; ++mn)
        for (int k = 0; k < 16; ++k) CHECK(rd(addr_mnmajor(d, mn, k)) == id(k4 * 16 + k, mn, 2), "anchor MN-major mn=%d k=%d", mn, k);
    }
  }

  const uint32_t tileA = 1024, tileB = tileA + TILE, pbase = 32768, dsbase = pbase + 3 * B_PCHUNK;
  fill_tile(tileA, 3);
  fill_tile(tileB, 4);

  // ---- S = Q K^T / dP = dO V^T: A = tile rows 0..127 (M = 128), B = tile rows 0..143 (N = 144), K = head dim ----
  for (int ks = 0; ks < 2; ++ks) {
    const Desc a = decode(desc_tile_kmajor(tileA, ks)), b = decode(desc_tile_kmajor(tileB, ks));
    CHECK(a.version == 1 && a.layout == LAYOUT_SW64, "tile K-major descriptor fields");
    for (int k = 0; k < 16; ++k) {
      for (int m = 0; m < 128; ++m) CHECK(rd(addr_kmajor(a, m, k)) == id(m, ks * 16 + k, 3), "S: A m=%d k=%d step %d", m, k, ks);
      for (int n = 0; n < N; ++n) CHECK(rd(addr_kmajor(b, n, k)) == id(n, ks * 16 + k, 4), "S: B n=%d k=%d step %d", n, k, ks);
    }
  }
  // ---- B = tile as [K = tokens][N = head dim] (P V, P^T dO, dS^T Q, dS K): nine steps of 16 tokens ----
  for (int kk = 0; kk < 9; ++kk) {
    const Desc b = decode(desc_tile_mnmajor(tileB, kk));
    CHECK(b.version == 1 && b.layout == LAYOUT_SW64, "tile MN-major descriptor fields");
    for (int k = 0; k < 16; ++k)
      for (int n = 0; n < HD; ++n) CHECK(rd(addr_mnmajor(b, n, k)) == id(kk * 16 + k, n, 4), "tile MN-major n=%d k=%d step %d", n, k, kk);
  }
  // ---- forward P (128 query rows): A K-major, M = queries, K = keys ----
  fill_pds(pbase, 128, F_PCHUNK, 5);
  for (int kk = 0; kk < 9; ++kk) {
    const Desc a = decode(desc_pds_kmajor(pbase, kk, F_PCHUNK));
    CHECK(a.version == 1 && a.layout == LAYOUT_SW128, "P K-major descriptor fields");
    for (int m = 0; m < 128; ++m)
      for (int k = 0; k < 16; ++k) CHECK(rd(addr_kmajor(a, m, k)) == id(m, kk * 16 + k, 5), "fwd P m=%d k=%d step %d", m, k, kk);
  }
  // ---- backward P / dS (144 query rows) ----
  fill_pds(pbase, N, B_PCHUNK, 6);
  fill_pds(dsbase, N, B_PCHUNK, 7);
  for (int kk = 0; kk < 9; ++kk) {
    // dQ = dS K: A K-major, M = queries 0..127, K = keys
    const Desc a = decode(desc_pds_kmajor(dsbase, kk, B_PCHUNK));
    for (int m = 0; m < 128; ++m)
      for (int k = 0; k < 16; ++k) CHECK(rd(addr_kmajor(a, m, k)) == id(m, kk * 16 + k, 7), "bwd dS K-major m=%d k=%d step %d", m, k, kk);
    // dV = P^T dO, dK = dS^T Q: A MN-major, M = keys 0..127, K = queries (all 144 rows)
    const Desc t = decode(desc_pds_mnmajor(pbase, kk));
    CHECK(t.version == 1 && t.layout == LAYOUT_SW128 && t.lbo == (uint32_t)B_PCHUNK, "P MN-major descriptor fields");
    for (int m = 0; m < 128; ++m)
      for (int k = 0; k < 16; ++k) CHECK(rd(addr_mnmajor(t, m, k)) == id(kk * 16 + k, m, 6), "bwd P MN-major m=%d k=%d step %d", m, k, kk);
  }
  // ---- descriptor field ranges ----
  CHECK((B_PCHUNK >> 4) < (1 << 14) && B_PCHUNK % 1024 == 0 && F_PCHUNK % 1024 == 0 && TILE % 1024 == 0, "strides");

  if (g_fail) {
    printf("%d mismatches\n", g_fail);
    return 1;
  }
  printf("winattn_tc layouts OK\n");
  return 0;
}
