// fiber_b200 — write-path probe (TEST / DEBUG TOOL): how fast can a B200 SM array WRITE a [M, N] bf16 matrix
//   (a) with coalesced 16-byte stores, (b) with TMA 2-D box stores of the shapes a GEMM epilogue can produce?
// The two-output GEMM epilogues plateau near 3.3 TB/s of algorithmic traffic; this separates "HBM write limit" from
// "box shape" from "TMA issue rate".
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/bw_probe.bin tools/bw_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../fiber_b200/csrc/common.cuh"

namespace fiber {
void set_last_error(const char*, ...) {}
}  // namespace fiber
using namespace fiber;

#define CK(e)                                                                          \
  do {                                                                                 \
    cudaError_t _e = (e);                                                              \
    if (_e != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(_e), __FILE__, __LINE__);  \
      return 2;                                                                        \
    }                                                                                  \
  } while (0)

__global__ void __launch_bounds__(512) stg_kernel(uint4* out, size_t n16) {
  const uint4 v = make_uint4(1, 2, 3, 4);
  for (size_t i = blockIdx.x * 512ull + threadIdx.x; i < n16; i += gridDim.x * 512ull) out[i] = v;
}

// every warp stores boxes of ROWS x COLS bf16 (COLS * 2 bytes per row) from its own smem staging area, walking the
// [M, N] matrix tile by tile (128 x 256 tiles dealt to CTAs like the GEMM does); INFLIGHT = bulk groups kept pending
template <int ROWS, int COLS, int INFLIGHT>
__global__ void __launch_bounds__(512, 1) tma_store_kernel(const __grid_constant__ CUtensorMap tm, int M, int N) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int BOX_BYTES = ROWS * COLS * 2;
  uint8_t* box = smem + warp * BOX_BYTES * (INFLIGHT + 1);
  for (int i = lane; i < BOX_BYTES * (INFLIGHT + 1) / 16; i += 32) reinterpret_cast<uint4*>(box)[i] = make_uint4(i, 1, 2, 3);
  fence_proxy_async_smem();
  __syncwarp();
  const int tiles_n = N / 256, tiles = (M / 128) * tiles_n;
  // a 128 x 256 tile = (128 / ROWS) x (256 / COLS) boxes, dealt round-robin to the 16 warps
  constexpr int BR = 128 / ROWS, BC = 256 / COLS, NB = BR * BC;
  int k = 0;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int m0 = (t / tiles_n) * 128, n0 = (t % tiles_n) * 256;
    for (int b = warp; b < NB; b += 16) {
      if (lane == 0) {
        if (INFLIGHT == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (INFLIGHT == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        if (INFLIGHT == 3) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
        tma_store_2d(&tm, box + (k % (INFLIGHT + 1)) * BOX_BYTES, n0 + (b % BC) * COLS, m0 + (b / BC) * ROWS);
        tma_store_commit();
      }
      ++k;
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int ROWS, int COLS, int INFLIGHT>
static int run(EncodeTiledFn enc, void* out, int M, int N, CUtensorMapSwizzle sw, const char* name) {
  CUtensorMap tm;
  const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
  const cuuint64_t strides[1] = {(cuuint64_t)N * 2};
  const cuuint32_t box[2] = {COLS, ROWS};
  const cuuint32_t estr[2] = {1, 1};
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    printf("%s: encode failed\n", name);
    return 0;
  }
  const int smem = 16 * ROWS * COLS * 2 * (INFLIGHT + 1) + 1024;
  CK(cudaFuncSetAttribute(tma_store_kernel<ROWS, COLS, INFLIGHT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e9f;
  for (int it = 0; it < 5; ++it) {
    cudaEventRecord(e0);
    tma_store_kernel<ROWS, COLS, INFLIGHT><<<148, 512, smem>>>(tm, M, N);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  printf("%-58s %7.1f us  %6.0f GB/s\n", name, best * 1e3, (double)M * N * 2 / best / 1e6);
  return 0;
}

int main() {
  const int M = 147456, N = 4096;  // 1.2 GB of bf16: the two outputs of the stage-2 fc1 GEMM
  void* out;
  CK(cudaMalloc(&out, (size_t)M * N * 2));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e9f;
  for (int it = 0; it < 5; ++it) {
    cudaEventRecord(e0);
    stg_kernel<<<148 * 4, 512>>>(reinterpret_cast<uint4*>(out), (size_t)M * N * 2 / 16);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  printf("%-58s %7.1f us  %6.0f GB/s\n", "coalesced STG.128 (grid 592 x 512)", best * 1e3, (double)M * N * 2 / best / 1e6);
  best = 1e9f;
  for (int it = 0; it < 5; ++it) {
    cudaEventRecord(e0);
    CK(cudaMemsetAsync(out, 0, (size_t)M * N * 2));
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  printf("%-58s %7.1f us  %6.0f GB/s\n", "cudaMemsetAsync", best * 1e3, (double)M * N * 2 / best / 1e6);
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(sym);
  run<32, 32, 0>(enc, out, M, N, CU_TENSOR_MAP_SWIZZLE_64B, "TMA store 32 x 32 (64 B rows), wait for read each");
  run<32, 32, 1>(enc, out, M, N, CU_TENSOR_MAP_SWIZZLE_64B, "TMA store 32 x 32 (64 B rows), 1 in flight");
  run<32, 32, 3>(enc, out, M, N, CU_TENSOR_MAP_SWIZZLE_64B, "TMA store 32 x 32 (64 B rows), 3 in flight");
  run<32, 64, 0>(enc, out, M, N, CU_TENSOR_MAP_SWIZZLE_128B, "TMA store 32 x 64 (128 B rows), wait for read each");
  run<32, 64, 1>(enc, out, M, N, CU_TENSOR_MAP_SWIZZLE_128B, "TMA store 32 x 64 (128 B rows), 1 in flight");
  run<64, 64, 1>(enc, out, M, N, CU_TENSOR_MAP_SWIZZLE_128B, "TMA store 64 x 64 (128 B rows), 1 in flight");
  run<128, 64, 0>(enc, out, M, N, CU_TENSOR_MAP_SWIZZLE_128B, "TMA store 128 x 64 (128 B rows), wait for read each");
  return 0;
}
