// fiber_b200 — hardware probe #2 for the fourth-generation window-attention kernels (TEST / DEBUG TOOL).
//
//   A. tcgen05.ld throughput per SM (4 / 8 warps, x8 .. x64 shapes): is the element-wise phase bound by the TMEM read port?
//   B. TMA 4-D box loads (32 channels x 6 x 6 tokens, SWIZZLE_64B) into a [144][32] tile at row offsets 0 / 36 / 72 / 108
//      (destinations that are 128-byte but not 512-byte aligned), wrapped (cyclic-shift) coordinates, and the 4-D box
//      stores back: does the swizzle follow the absolute shared-memory address, i.e. tcl::sw64_off(row, piece)?
//   C. tcgen05.mma with N = 72 / 80 / 64 (M = 128): which key splits are legal instruction shapes?
//   D. MUFU.EX2 vs FMA-pipe exp2 polynomial throughput (8 warps): how much a software exp2 can offload.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/tc_probe2.bin tools/tc_probe2.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../fiber_b200/csrc/common.cuh"
#include "../fiber_b200/csrc/window_tc_layout.cuh"

namespace fiber {
void set_last_error(const char*, ...) {}
}  // namespace fiber

using namespace fiber;
using namespace fiber::tcl;

#define CK(e)                                                                          \
  do {                                                                                 \
    cudaError_t _e = (e);                                                              \
    if (_e != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(_e), __FILE__, __LINE__);  \
      return 2;                                                                        \
    }                                                                                  \
  } while (0)

// ---------------------------------------------------------------------------------------------------------------
// A. TMEM read throughput
// ---------------------------------------------------------------------------------------------------------------
template <int X>
__device__ __forceinline__ void ld_x(uint32_t taddr, uint32_t& sink) {
  if constexpr (X == 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
    sink ^= r[0] ^ r[7];
  } else if constexpr (X == 16) {
    uint32_t r[16];
    tmem_ld16(taddr, r);
    sink ^= r[0] ^ r[15];
  } else if constexpr (X == 32) {
    uint32_t r[32];
    tmem_ld32(taddr, r);
    sink ^= r[0] ^ r[31];
  } else {
    uint32_t r[64];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
        "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]),
          "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]),
          "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]),
          "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr) : "memory");
    sink ^= r[0] ^ r[63];
  }
}

template <int X>
__global__ void __launch_bounds__(512, 1) tmem_bw_kernel(int iters, long long* cycles, uint32_t* sinks) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tmem_slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >> 2) & 3) * 128;
  uint32_t sink = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 128; c += X) ld_x<X>(base + c, sink);
    tmem_ld_wait();
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sinks[blockIdx.x * blockDim.x + threadIdx.x] = sink;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_slot, 512);
  }
}

template <int X>
static int run_tmem_bw(int nwarps, long long* d_cyc, uint32_t* d_sink) {
  const int iters = 2000;
  tmem_bw_kernel<X><<<148, nwarps * 32>>>(iters, d_cyc, d_sink);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  long long cyc[148];
  CK(cudaMemcpy(cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost));
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += cyc[i];
  avg /= 148;
  const double bytes = (double)nwarps * iters * 128 * 32 * 4;
  printf("A  tcgen05.ld x%-3d %2d warps: %8.0f cycles  -> %6.1f B/clk/SM   (%.1f clk per warp-instruction)\n", X, nwarps, avg,
         bytes / avg, avg / (iters * (128 / X)));
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// B. TMA 4-D box loads / stores
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

__global__ void __launch_bounds__(128, 1) tma_box_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                                                          int c0, int h0, int w0, int H, int W, int b, uint8_t* dump) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  for (int i = tid; i < TILE / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar, TILE);
    for (int q = 0; q < 4; ++q) {
      int hq = h0 + 6 * (q >> 1), wq = w0 + 6 * (q & 1);
      hq -= hq >= H ? H : 0;
      wq -= wq >= W ? W : 0;
      tma_load_4d(smem_u32(smem) + q * 36 * 64, &tm_in, &bar, c0, wq, hq, b);
    }
  }
  mbar_wait(&bar, 0);
  for (int i = tid; i < TILE / 16; i += 128) reinterpret_cast<uint4*>(dump)[i] = reinterpret_cast<const uint4*>(smem)[i];
  fence_proxy_async_smem();
  __syncthreads();
  if (tid == 0) {
    for (int q = 0; q < 4; ++q) {
      int hq = h0 + 6 * (q >> 1), wq = w0 + 6 * (q & 1);
      hq -= hq >= H ? H : 0;
      wq -= wq >= W ? W : 0;
      tma_store_4d(&tm_out, smem_u32(smem) + q * 36 * 64, c0, wq, hq, b);
    }
    tma_store_commit();
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// C. MMA shapes: S[128][n] = X[128][32] . Y[n][32]^T with n in {64, 72, 80, 144}
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) mma_n_kernel(const uint8_t* __restrict__ img, int n, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 2 * TILE / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(img)[i];
  fence_proxy_async_smem();
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(&bar, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (tid == 0) {
    const uint32_t a = smem_u32(smem), b = a + TILE;
    const uint32_t idesc = umma_idesc_bf16(128, n, 0, 0);
    for (int ks = 0; ks < 2; ++ks) umma_f16_ss(tmem_base, desc_tile_kmajor(a, ks), desc_tile_kmajor(b, ks), idesc, ks);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < n; c0 += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0) : "memory");
    tmem_ld_wait();
    for (int e = 0; e < 8; ++e) out[row * 144 + c0 + e] = __uint_as_float(r[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_slot, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// D. exp2 throughput: MUFU only, polynomial only, and a 3:1 mix (8 warps per SM, dependent-free streams)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float exp2_poly(float x) {
  // 2^x for x <= 0 (softmax exponents): round-to-nearest split x = n + f, f in [-0.5, 0.5], cubic for 2^f, exponent add
  const float t = x + 12582912.0f;         // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float n = t - 12582912.0f;
  const float f = x - n;
  float p = fmaf(f, 0.0555041086648216f, 0.2402264923172690f);
  p = fmaf(p, f, 0.6931471805599453f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) exp2_kernel(int iters, float seed, long long* cycles, float* sinks) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = -seed * (threadIdx.x + 1 + i) * 1e-3f;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y;
      if (MODE == 0 || (MODE == 2 && (i & 3) != 3)) y = ex2_approx(x[i]);
      else y = exp2_poly(x[i]);
      acc += y;
      x[i] -= 0.001f;
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sinks[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

static float bf2f(uint16_t b) {
  uint32_t u = static_cast<uint32_t>(b) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFF + ((u >> 16) & 1);
  return static_cast<uint16_t>(u >> 16);
}

int main() {
  long long* d_cyc;
  uint32_t* d_sink;
  CK(cudaMalloc(&d_cyc, 148 * sizeof(long long)));
  CK(cudaMalloc(&d_sink, 148 * 512 * sizeof(uint32_t)));
  // ---- A ----
  for (int nw : {4, 8, 16}) {
    if (run_tmem_bw<8>(nw, d_cyc, d_sink)) return 2;
    if (run_tmem_bw<16>(nw, d_cyc, d_sink)) return 2;
    if (run_tmem_bw<32>(nw, d_cyc, d_sink)) return 2;
    if (run_tmem_bw<64>(nw, d_cyc, d_sink)) return 2;
  }
  // ---- D ----
  {
    const int iters = 4000;
    const char* nm[3] = {"MUFU.EX2 only", "polynomial only", "3 MUFU : 1 polynomial"};
    for (int mode = 0; mode < 3; ++mode) {
      if (mode == 0) exp2_kernel<0><<<148, 256>>>(iters, 1.0f, d_cyc, reinterpret_cast<float*>(d_sink));
      if (mode == 1) exp2_kernel<1><<<148, 256>>>(iters, 1.0f, d_cyc, reinterpret_cast<float*>(d_sink));
      if (mode == 2) exp2_kernel<2><<<148, 256>>>(iters, 1.0f, d_cyc, reinterpret_cast<float*>(d_sink));
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
      long long cyc[148];
      CK(cudaMemcpy(cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost));
      double avg = 0;
      for (int i = 0; i < 148; ++i) avg += cyc[i];
      avg /= 148;
      printf("D  %-24s 8 warps: %.2f exp2 per clk per SM\n", nm[mode], 256.0 * iters * 8 / avg);
    }
  }
  // ---- B ----
  {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres));
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(sym);
    const int B = 2, H = 24, W = 24, LD = 96;  // [B*H*W rows][96 channels]
    const size_t nel = (size_t)B * H * W * LD;
    std::vector<uint16_t> g(nel);
    for (size_t i = 0; i < nel; ++i) g[i] = (uint16_t)((i * 2654435761u) >> 16);
    uint16_t *d_in, *d_out;
    uint8_t* d_dump;
    CK(cudaMalloc(&d_in, nel * 2));
    CK(cudaMalloc(&d_out, nel * 2));
    CK(cudaMalloc(&d_dump, TILE));
    CK(cudaMemcpy(d_in, g.data(), nel * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_out, 0, nel * 2));
    CUtensorMap tm_in, tm_out;
    const cuuint64_t dims[4] = {(cuuint64_t)LD, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)LD * 2, (cuuint64_t)W * LD * 2, (cuuint64_t)H * W * LD * 2};
    const cuuint32_t box[4] = {32, 6, 6, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r1 = enc(&tm_in, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d_in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&tm_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d_out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
      printf("B  cuTensorMapEncodeTiled failed: %d %d\n", (int)r1, (int)r2);
    } else {
      const int c0 = 32, h0 = 18, w0 = 18, b = 1;  // shifted last window: all four quadrants wrap differently
      CK(cudaFuncSetAttribute(tma_box_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE + 1024));
      tma_box_kernel<<<1, 128, TILE + 1024>>>(tm_in, tm_out, c0, h0, w0, H, W, b, d_dump);
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
      std::vector<uint8_t> dump(TILE);
      std::vector<uint16_t> back(nel);
      CK(cudaMemcpy(dump.data(), d_dump, TILE, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(back.data(), d_out, nel * 2, cudaMemcpyDeviceToHost));
      int bad = 0, bad_linear = 0, bad_store = 0;
      size_t expect_nonzero = 0;
      for (int q = 0; q < 4; ++q)
        for (int r6 = 0; r6 < 6; ++r6)
          for (int c6 = 0; c6 < 6; ++c6) {
            const int hq = (h0 + 6 * (q >> 1) + r6) % H, wq = (w0 + 6 * (q & 1) + c6) % W;
            const int row = 36 * q + r6 * 6 + c6;
            for (int c = 0; c < 32; ++c) {
              const size_t gi = (((size_t)b * H + hq) * W + wq) * LD + c0 + c;
              uint16_t got, got_lin;
              memcpy(&got, &dump[sw64_off(row, c / 8) + (c % 8) * 2], 2);
              memcpy(&got_lin, &dump[row * 64 + c * 2], 2);
              if (got != g[gi]) ++bad;
              if (got_lin != g[gi]) ++bad_linear;
              if (back[gi] != g[gi]) ++bad_store;
              ++expect_nonzero;
            }
          }
      size_t nonzero = 0;
      for (size_t i = 0; i < nel; ++i) nonzero += back[i] != 0;
      printf("B  TMA 4-D box loads into rows 0/36/72/108 (SWIZZLE_64B): %d mismatches vs sw64_off (absolute-address swizzle)  %s\n", bad,
             bad ? "FAILED" : "ok");
      printf("B  (same data read as an unswizzled tile: %d mismatches — expected to be many)\n", bad_linear);
      printf("B  TMA 4-D box stores back: %d mismatches, %zu elements written (expected about %zu)  %s\n", bad_store, nonzero,
             expect_nonzero, (bad_store || nonzero > expect_nonzero) ? "FAILED" : "ok");
    }
  }
  // ---- C ----
  {
    std::vector<float> X(N * HD), Y(N * HD);
    srand(99);
    auto rnd = [] { return bf2f(f2bf((rand() % 2001 - 1000) / 500.0f)); };
    for (auto& v : X) v = rnd();
    for (auto& v : Y) v = rnd();
    std::vector<uint8_t> img(2 * TILE, 0);
    for (int r = 0; r < N; ++r)
      for (int c = 0; c < HD; ++c) {
        uint16_t a = f2bf(X[r * HD + c]), b = f2bf(Y[r * HD + c]);
        memcpy(&img[sw64_off(r, c / 8) + (c % 8) * 2], &a, 2);
        memcpy(&img[TILE + sw64_off(r, c / 8) + (c % 8) * 2], &b, 2);
      }
    uint8_t* d_img;
    float* d_o;
    CK(cudaMalloc(&d_img, 2 * TILE));
    CK(cudaMalloc(&d_o, 128 * 144 * sizeof(float)));
    CK(cudaMemcpy(d_img, img.data(), 2 * TILE, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(mma_n_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE + 1024));
    for (int n : {144, 64, 80, 72, 48, 40}) {
      CK(cudaMemset(d_o, 0xFF, 128 * 144 * sizeof(float)));
      mma_n_kernel<<<1, 128, 2 * TILE + 1024>>>(d_img, n, d_o);
      cudaError_t e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("C  tcgen05.mma M=128 N=%-3d: launch failed (%s) — stopping the shape probe\n", n, cudaGetErrorString(e));
        break;
      }
      std::vector<float> out(128 * 144);
      CK(cudaMemcpy(out.data(), d_o, out.size() * sizeof(float), cudaMemcpyDeviceToHost));
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int c = 0; c < n; ++c) {
          double ref = 0;
          for (int k = 0; k < HD; ++k) ref += (double)X[m * HD + k] * Y[c * HD + k];
          if (!(fabs(out[m * 144 + c] - ref) <= 1e-3 * (1.0 + fabs(ref)))) ++bad;
        }
      printf("C  tcgen05.mma M=128 N=%-3d: %d wrong elements  %s\n", n, bad, bad ? "FAILED" : "ok");
    }
  }
  return 0;
}
