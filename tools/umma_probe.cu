// fiber_b200 — hardware probe for the UMMA operand forms of the tcgen05 window-attention kernels.
//
// TEST / DEBUG TOOL (not part of the library).  window_attn_tc.cu relies on four descriptor forms that
// gemm_sm100.cu does not exercise (SWIZZLE_64B K-major and MN-major tiles, mixed operand majors, N = 144 / 32).
// This program runs ONE tcgen05.mma chain per form on operand images that the HOST writes through the kernels'
// own layout functions (fiber_b200/csrc/window_tc_layout.cuh) and compares the TMEM accumulator with a CPU
// product, so a wrong assumption shows up as "form X FAILED" instead of as a wrong attention output.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/umma_probe.bin tools/umma_probe.cu
//   timeout 120 tools/umma_probe.bin            (tools/gpu_round2a.sh does both)
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

constexpr int IMG_BYTES = 192 * 1024;  // shared-memory image (operands at fixed offsets, see main)

// forms: 0 S = A B^T (tile K-major x tile K-major, N = 144, 2 steps)
//        1 O = P V    (P K-major fwd chunks x tile MN-major, N = 32, 9 steps)
//        2 dV = P^T dO (P MN-major bwd chunks x tile MN-major, N = 32, 9 steps)
//        3 dQ = dS K  (dS K-major bwd chunks x tile MN-major, N = 32, 9 steps)
//        4 = form 0 with the accumulator at TMEM column 160 (the second S buffer of the forward kernel)
__global__ void __launch_bounds__(128, 1) probe_kernel(const uint8_t* __restrict__ img, int form, uint32_t a_off,
                                                        uint32_t b_off, float* __restrict__ out /* [128][N] */) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < IMG_BYTES / 16; i += 128)
    reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(img)[i];
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
  const uint32_t tmem_base = tmem_slot + (form == 4 ? 160u : 0u);
  const int n = (form == 0 || form == 4) ? N : HD;
  if (tid == 0) {
    const uint32_t a = smem_u32(smem) + a_off, b = smem_u32(smem) + b_off;
    if (form == 0 || form == 4) {
      const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
      for (int ks = 0; ks < 2; ++ks) umma_f16_ss(tmem_base, desc_tile_kmajor(a, ks), desc_tile_kmajor(b, ks), idesc, ks);
    } else if (form == 1) {
      const uint32_t idesc = umma_idesc_bf16(128, HD, 0, 1);
      for (int kk = 0; kk < 9; ++kk) umma_f16_ss(tmem_base, desc_pds_kmajor(a, kk, F_PCHUNK), desc_tile_mnmajor(b, kk), idesc, kk);
    } else if (form == 2) {
      const uint32_t idesc = umma_idesc_bf16(128, HD, 1, 1);
      for (int kk = 0; kk < 9; ++kk) umma_f16_ss(tmem_base, desc_pds_mnmajor(a, kk), desc_tile_mnmajor(b, kk), idesc, kk);
    } else {
      const uint32_t idesc = umma_idesc_bf16(128, HD, 0, 1);
      for (int kk = 0; kk < 9; ++kk) umma_f16_ss(tmem_base, desc_pds_kmajor(a, kk, B_PCHUNK), desc_tile_mnmajor(b, kk), idesc, kk);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < n; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int e = 0; e < 16; ++e) out[row * n + c0 + e] = __uint_as_float(r[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_slot, 512);
  }
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
static void put(std::vector<uint8_t>& img, uint32_t off, uint16_t v) { memcpy(&img[off], &v, 2); }

#define CK(e)                                                                          \
  do {                                                                                 \
    cudaError_t _e = (e);                                                              \
    if (_e != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(_e), __FILE__, __LINE__);  \
      return 2;                                                                        \
    }                                                                                  \
  } while (0)

int main() {
  // logical operands: X[144][32] and Y[144][32] tiles, W[144][144] "P / dS" matrix
  std::vector<float> X(N * HD), Y(N * HD), W(N * N);
  srand(1234);
  auto rnd = [] { return bf2f(f2bf((rand() % 2001 - 1000) / 500.0f)); };
  for (auto& v : X) v = rnd();
  for (auto& v : Y) v = rnd();
  for (auto& v : W) v = rnd();
  const uint32_t offX = 0, offY = TILE, offWf = 2 * TILE, offWb = offWf + 3 * F_PCHUNK;  // 1024-aligned offsets
  static_assert((2 * TILE) % 1024 == 0 && (3 * F_PCHUNK) % 1024 == 0, "alignment");
  if (offWb + 3 * B_PCHUNK > (uint32_t)IMG_BYTES) { printf("image too small\n"); return 2; }
  std::vector<uint8_t> img(IMG_BYTES, 0);
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < HD; ++c) {
      put(img, offX + sw64_off(r, c / 8) + (c % 8) * 2, f2bf(X[r * HD + c]));
      put(img, offY + sw64_off(r, c / 8) + (c % 8) * 2, f2bf(Y[r * HD + c]));
    }
  for (int r = 0; r < N; ++r)
    for (int j = 0; j < N; ++j) {
      if (r < 128) put(img, offWf + pds_piece_off(r, j / 8, F_PCHUNK) + (j % 8) * 2, f2bf(W[r * N + j]));
      put(img, offWb + pds_piece_off(r, j / 8, B_PCHUNK) + (j % 8) * 2, f2bf(W[r * N + j]));
    }
  uint8_t* d_img;
  float* d_out;
  CK(cudaMalloc(&d_img, IMG_BYTES));
  CK(cudaMalloc(&d_out, 128 * N * sizeof(float)));
  CK(cudaMemcpy(d_img, img.data(), IMG_BYTES, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IMG_BYTES + 1024));
  const char* names[5] = {"S = X Y^T   (SW64 K-major x SW64 K-major, N=144)", "O = W Y     (SW128 K-major x SW64 MN-major, N=32)",
                          "dV = W^T Y  (SW128 MN-major x SW64 MN-major, N=32)", "dQ = W Y    (bwd chunks K-major x SW64 MN-major)",
                          "S = X Y^T   (accumulator at TMEM column 160)"};
  int bad_forms = 0;
  for (int form = 0; form < 5; ++form) {
    const int n = (form == 0 || form == 4) ? N : HD;
    const uint32_t a_off = (form == 0 || form == 4) ? offX : (form == 1 ? offWf : offWb);
    CK(cudaMemset(d_out, 0xFF, 128 * N * sizeof(float)));
    probe_kernel<<<1, 128, IMG_BYTES + 1024>>>(d_img, form, a_off, offY, d_out);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> out(128 * n);
    CK(cudaMemcpy(out.data(), d_out, out.size() * sizeof(float), cudaMemcpyDeviceToHost));
    double max_err = 0;
    int bad = 0;
    for (int m = 0; m < 128; ++m)
      for (int c = 0; c < n; ++c) {
        double ref = 0;
        if (form == 0 || form == 4) for (int k = 0; k < HD; ++k) ref += (double)X[m * HD + k] * Y[c * HD + k];
        else if (form == 2) for (int k = 0; k < N; ++k) ref += (double)W[k * N + m] * Y[k * HD + c];
        else for (int k = 0; k < N; ++k) ref += (double)W[m * N + k] * Y[k * HD + c];
        const double err = fabs(out[m * n + c] - ref);
        if (err > max_err) max_err = err;
        if (!(err <= 1e-3 * (1.0 + fabs(ref)))) ++bad;
      }
    printf("form %d %-58s max |err| %.3g  %s\n", form, names[form], max_err, bad ? "FAILED" : "ok");
    if (bad) ++bad_forms;
  }
  cudaFree(d_img);
  cudaFree(d_out);
  return bad_forms ? 1 : 0;
}
