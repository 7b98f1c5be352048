// fiber_b200 — image side of the input pipeline on the GPU (include/fiber_b200.h: fiber_image_transform).
//
// A batch of decoded, ragged RGB byte images -> the normalised float32 NCHW batch FIBERTransformerSS.infer consumes
// (coarse_grained/fiber/modules/fiber_module.py:237-241), with the arithmetic of the reference's per-image CPU
// transform (transforms/transform.py:10-17: PIL bicubic Resize, ToTensor, Normalize) reproduced bit for bit.
// HBM-bound byte work: per image the source box is read once (3 box_w box_h bytes), the horizontally resampled byte
// planes (3 box_h out_w bytes, L2-resident for a batch) written and read once, and 12 out_h out_w bytes written.
// Three launches per batch, grids sized by the largest image of the batch; per-element bodies live in
// image_resample.cuh so that the CPU test-suite can replay them.
#include "common.cuh"
#include "image_resample.cuh"

#include <atomic>
#include <cstdlib>

namespace fiber {
void count_launch(int n);
void set_image_variant(int v);
int get_image_variant();

namespace img {

struct Norm {
  float mean[3], stdv[3];
};

__global__ void __launch_bounds__(128) image_coeffs_kernel(const fiber_image_desc* descs, void* ws, Norm norm, int out_h,
                                                           int out_w) {
  pdl_trigger();
  pdl_wait();
  coeffs_body(descs, ws, out_h, out_w, blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x, norm.mean, norm.stdv);
}

template <int R>
__global__ void __launch_bounds__(256) image_hpass_kernel(const fiber_image_desc* descs, void* ws, int out_h, int out_w) {
  pdl_trigger();
  pdl_wait();
  hpass_body<R>(descs, ws, out_h, out_w, blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x);
}

template <int R>
__global__ void __launch_bounds__(256) image_hpass_words_kernel(const fiber_image_desc* descs, void* ws, int out_h, int out_w) {
  pdl_trigger();
  pdl_wait();
  hpass_words_body<R>(descs, ws, out_h, out_w, blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x);
}

template <int W>
__global__ void __launch_bounds__(256) image_vpass_kernel(const fiber_image_desc* descs, const void* ws, float* out, int out_h,
                                                          int out_w) {
  __shared__ float lut[3 * 256];
  pdl_trigger();
  pdl_wait();
  for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) lut[i] = static_cast<const float*>(ws)[i];
  __syncthreads();
  vpass_body<W>(descs, ws, lut, out, out_h, out_w, blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x);
}

template <int W>
__global__ void __launch_bounds__(256) image_vpass_staged_kernel(const fiber_image_desc* descs, const void* ws, float* out,
                                                                 int out_h, int out_w) {
  extern __shared__ __align__(16) uint8_t staged[];   // [3 * 256] float look-up table, then the band's plane rows
  float* lut = reinterpret_cast<float*>(staged);
  pdl_trigger();
  pdl_wait();
  for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) lut[i] = static_cast<const float*>(ws)[i];
  vstage_load(descs, ws, staged + kLutBytes, out_h, out_w, blockIdx.y, blockIdx.x, threadIdx.x, blockDim.x);
  __syncthreads();
  vstage_compute<W>(descs, ws, staged + kLutBytes, lut, out, out_h, out_w, blockIdx.y, blockIdx.x, threadIdx.x, blockDim.x);
}

// "image_variant": bit 0 = the horizontal pass reads the source as aligned words (hpass_words_body) instead of bytes,
// bit 1 = eight output columns per thread in the vertical pass instead of four, bit 2 = sixteen; bit 3 = eight source rows
// per thread in the word-form horizontal pass instead of four; bit 4 = the vertical pass per band of 16 output rows with
// its plane rows staged in shared memory (falls back to the global form when a band needs more than 200 KB).  Same bytes
// out either way.
// FIBER_IMAGE_VARIANT.
static std::atomic<int> g_variant{-1};
static int option_variant() {
  int v = g_variant.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("FIBER_IMAGE_VARIANT");
    v = e ? (atoi(e) & 31) : kDefaultVariant;
    g_variant.store(v, std::memory_order_relaxed);
  }
  return v;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace img
void set_image_variant(int v) { img::g_variant.store(v < 0 ? -1 : (v & 31), std::memory_order_relaxed); }
int get_image_variant() { return img::option_variant(); }
}  // namespace fiber

extern "C" {

size_t fiber_image_transform_plan(fiber_image_desc* d, int32_t n, int32_t out_h, int32_t out_w) {
  using namespace fiber::img;
  if (!d || n <= 0 || out_h <= 0 || out_w <= 0 || (out_w & 3)) {
    fiber::set_last_error("image_transform_plan: need n > 0, out_h > 0, out_w > 0 and out_w %% 4 == 0");
    return 0;
  }
  size_t off = kLutBytes;
  for (int i = 0; i < n; ++i) {
    fiber_image_desc& e = d[i];
    const bool layout_ok = e.planar ? (e.planar == 1 && e.stride >= e.w && e.chan_stride >= static_cast<int64_t>(e.h - 1) * e.stride + e.w)
                                    : e.stride >= 3LL * e.w;
    if (!e.src || e.h <= 0 || e.w <= 0 || !layout_ok || e.box_w <= 0 || e.box_h <= 0 || e.box_x < 0 || e.box_y < 0 ||
        e.box_x + e.box_w > e.w || e.box_y + e.box_h > e.h) {
      fiber::set_last_error("image_transform_plan: image %d: bad size, stride or crop box", i);
      return 0;
    }
    if (e.box_h > 100LL * e.box_w && out_h < e.box_h) {
      fiber::set_last_error("image_transform_plan: image %d: box taller than 100x its width (Pillow resamples rows first there)", i);
      return 0;
    }
    e.ksize_x = ksize_for(e.box_w, out_w);
    e.ksize_y = ksize_for(e.box_h, out_h);
    e.coef_off = static_cast<int64_t>(off);
    off = align_up(off + static_cast<size_t>(table_ints(e, out_h, out_w)) * sizeof(int32_t), 16);
  }
  for (int i = 0; i < n; ++i) {
    d[i].tmp_off = static_cast<int64_t>(off);
    off = align_up(off + 3ull * d[i].box_h * tmp_pitch(out_w), 16);
  }
  return off;
}

int fiber_image_transform(const fiber_image_desc* dh, const fiber_image_desc* dd, int32_t n, int32_t out_h, int32_t out_w,
                          const float* mean, const float* stdv, void* ws, size_t ws_bytes, float* out, fiber_stream_t s) {
  using namespace fiber::img;
  cudaStream_t stream = static_cast<cudaStream_t>(s);
  FIBER_CHECK(dh && dd && n > 0 && out_h > 0 && out_w > 0 && (out_w & 3) == 0 && mean && stdv && ws && out,
              "image_transform: null argument, n <= 0 or out_w %% 4 != 0");
  FIBER_CHECK((reinterpret_cast<uintptr_t>(ws) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "image_transform: workspace and out must be 16-byte aligned");
  FIBER_CHECK(n <= 65535, "image_transform: at most 65535 images per call");
  int max_box_h = 0;
  size_t need = 0;
  bool any_planar = false;
  for (int i = 0; i < n; ++i) {
    const fiber_image_desc& e = dh[i];
    any_planar = any_planar || e.planar != 0;
    FIBER_CHECK(e.ksize_x == ksize_for(e.box_w, out_w) && e.ksize_y == ksize_for(e.box_h, out_h) && e.coef_off >= kLutBytes &&
                    e.tmp_off >= 0,
                "image_transform: descriptor %d was not planned for this output size (fiber_image_transform_plan)", i);
    max_box_h = e.box_h > max_box_h ? e.box_h : max_box_h;
    const size_t end = static_cast<size_t>(e.tmp_off) + 3ull * e.box_h * tmp_pitch(out_w);
    need = end > need ? end : need;
  }
  FIBER_CHECK(ws_bytes >= need, "image_transform: workspace too small (%zu < %zu bytes)", ws_bytes, need);
  Norm norm;
  for (int c = 0; c < 3; ++c) {
    norm.mean[c] = mean[c];
    norm.stdv[c] = stdv[c];
  }
  const dim3 g1(((out_w + out_h > 768 ? out_w + out_h : 768) + 127) / 128, n);
  FIBER_CUDA(fiber::launch_k(image_coeffs_kernel, g1, dim3(128), 0, stream, dd, ws, norm, out_h, out_w));
  const int variant = any_planar ? (option_variant() & ~9) : option_variant();  // planar sources: byte-form horizontal pass
  const int R = (variant & 9) == 9 ? 8 : 4, W = (variant & 4) ? 4 : ((variant & 2) ? 2 : 1);
  const long long hwork = static_cast<long long>((max_box_h + R - 1) / R) * out_w;
  const long long vwork = static_cast<long long>(out_h) * ((out_w / 4 + W - 1) / W);
  FIBER_CHECK(hwork < (1LL << 31) && vwork < (1LL << 31), "image_transform: image too large for 32-bit indexing");
  const dim3 g2(static_cast<unsigned>((hwork + 255) / 256), n);
  if (R == 8)
    FIBER_CUDA(fiber::launch_k(image_hpass_words_kernel<8>, g2, dim3(256), 0, stream, dd, ws, out_h, out_w));
  else if (variant & 1)
    FIBER_CUDA(fiber::launch_k(image_hpass_words_kernel<4>, g2, dim3(256), 0, stream, dd, ws, out_h, out_w));
  else
    FIBER_CUDA(fiber::launch_k(image_hpass_kernel<4>, g2, dim3(256), 0, stream, dd, ws, out_h, out_w));
  if (variant & 16) {   // staged form: shared memory for the band that needs the most plane rows
    int max_rows = 0;
    for (int i = 0; i < n; ++i)
      for (int y0 = 0; y0 < out_h; y0 += kBandRows) {
        const int y1 = y0 + kBandRows < out_h ? y0 + kBandRows : out_h;
        int f0, c0, f1, c1;
        bounds_one(dh[i].box_h, out_h, y0, &f0, &c0);
        bounds_one(dh[i].box_h, out_h, y1 - 1, &f1, &c1);
        max_rows = f1 + c1 - f0 > max_rows ? f1 + c1 - f0 : max_rows;
      }
    const size_t smem = kLutBytes + 3ull * max_rows * tmp_pitch(out_w);
    if (smem <= 200 * 1024) {
      auto kern = W == 4 ? image_vpass_staged_kernel<4> : (W == 2 ? image_vpass_staged_kernel<2> : image_vpass_staged_kernel<1>);
      FIBER_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      const dim3 gs((out_h + kBandRows - 1) / kBandRows, n);
      FIBER_CUDA(fiber::launch_k(kern, gs, dim3(256), smem, stream, dd, static_cast<const void*>(ws), out, out_h, out_w));
      fiber::count_launch(3);
      return 0;
    }
  }
  const dim3 g3(static_cast<unsigned>((vwork + 255) / 256), n);
  if (W == 4)
    FIBER_CUDA(fiber::launch_k(image_vpass_kernel<4>, g3, dim3(256), 0, stream, dd, static_cast<const void*>(ws), out, out_h,
                               out_w));
  else if (W == 2)
    FIBER_CUDA(fiber::launch_k(image_vpass_kernel<2>, g3, dim3(256), 0, stream, dd, static_cast<const void*>(ws), out, out_h,
                               out_w));
  else
    FIBER_CUDA(fiber::launch_k(image_vpass_kernel<1>, g3, dim3(256), 0, stream, dd, static_cast<const void*>(ws), out, out_h,
                               out_w));
  fiber::count_launch(3);
  return 0;
}

}  // extern "C"
