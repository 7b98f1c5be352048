// CPU replay of the image-transform kernels (fiber_b200/csrc/image_pipeline.cu) — TEST INFRASTRUCTURE, built with g++
// and driven by tests/test_image_pipeline_cpu.py.  Includes the kernels' own per-element bodies
// (fiber_b200/csrc/image_resample.cuh) and calls them for every (block, thread) index of the three launches, with the
// launcher's grid arithmetic, on host memory.  Descriptors come planned from the real library
// (fiber_image_transform_plan is host-only), so offsets and tap counts are the product's.
#include <cstdint>
#include <algorithm>
#include <vector>

#include "../../fiber_b200/csrc/image_resample.cuh"

using namespace fiber::img;

template <int R, int W, bool WORDS, bool STAGED = false>
static int run(const fiber_image_desc* d, int n, int out_h, int out_w, const float* mean, const float* stdv, void* ws,
               float* out) {
  int max_box_h = 0;
  for (int i = 0; i < n; ++i) max_box_h = d[i].box_h > max_box_h ? d[i].box_h : max_box_h;
  const float* lut = static_cast<const float*>(ws);   // written by coeffs_body (the device copies it to shared memory)
  const int g1 = ((out_w + out_h > 768 ? out_w + out_h : 768) + 127) / 128;
  const long long hwork = static_cast<long long>((max_box_h + R - 1) / R) * out_w;
  const long long g2 = (hwork + 255) / 256;
  const long long vwork = static_cast<long long>(out_h) * ((out_w / 4 + W - 1) / W);
  const long long g3 = (vwork + 255) / 256;
  for (int img = 0; img < n; ++img)
    for (int b = 0; b < g1; ++b)
      for (int t = 0; t < 128; ++t) coeffs_body(d, ws, out_h, out_w, img, b * 128 + t, mean, stdv);
  for (int img = 0; img < n; ++img)
    for (long long b = 0; b < g2; ++b)
      for (int t = 0; t < 256; ++t) {
        if (WORDS) hpass_words_body<R>(d, ws, out_h, out_w, img, static_cast<int>(b * 256 + t));
        else hpass_body<R>(d, ws, out_h, out_w, img, static_cast<int>(b * 256 + t));
      }
  if (STAGED) {   // one block per (band, image): phase 1 for every thread, barrier, phase 2 for every thread
    std::vector<uint8_t> smem(256 * 1024 + 16);
    uint8_t* base = smem.data() + (16 - reinterpret_cast<uintptr_t>(smem.data()) % 16) % 16;
    const int bands = (out_h + kBandRows - 1) / kBandRows;
    for (int img = 0; img < n; ++img)
      for (int b = 0; b < bands; ++b) {
        std::fill(smem.begin(), smem.end(), 0xCD);
        for (int t = 0; t < 256; ++t) vstage_load(d, ws, base, out_h, out_w, img, b, t, 256);
        for (int t = 0; t < 256; ++t) vstage_compute<W>(d, ws, base, lut, out, out_h, out_w, img, b, t, 256);
      }
    return 0;
  }
  for (int img = 0; img < n; ++img)
    for (long long b = 0; b < g3; ++b)
      for (int t = 0; t < 256; ++t) vpass_body<W>(d, ws, lut, out, out_h, out_w, img, static_cast<int>(b * 256 + t));
  return 0;
}

// variant: the library's "image_variant" option (bit 0: word-form horizontal pass, bit 1 / 2: eight / sixteen
// columns per thread, bit 3: eight rows per thread, bit 4: vertical pass staged per band in shared memory)
extern "C" int emul_image_transform(const fiber_image_desc* d, int n, int out_h, int out_w, const float* mean,
                                    const float* stdv, void* ws, float* out, int variant) {
  switch (variant & 31) {
    case 0: return run<4, 1, false>(d, n, out_h, out_w, mean, stdv, ws, out);
    case 1: return run<4, 1, true>(d, n, out_h, out_w, mean, stdv, ws, out);
    case 2: return run<4, 2, false>(d, n, out_h, out_w, mean, stdv, ws, out);
    case 3: return run<4, 2, true>(d, n, out_h, out_w, mean, stdv, ws, out);
    case 5: case 7: return run<4, 4, true>(d, n, out_h, out_w, mean, stdv, ws, out);
    case 11: return run<8, 2, true>(d, n, out_h, out_w, mean, stdv, ws, out);
    case 13: case 15: return run<8, 4, true>(d, n, out_h, out_w, mean, stdv, ws, out);
    case 19: return run<4, 2, true, true>(d, n, out_h, out_w, mean, stdv, ws, out);
    case 18: return run<4, 2, false, true>(d, n, out_h, out_w, mean, stdv, ws, out);
    case 23: return run<4, 4, true, true>(d, n, out_h, out_w, mean, stdv, ws, out);
    default: return -1;
  }
}
