// fiber_b200 — per-element bodies of the image-transform kernels (csrc/image_pipeline.cu).
//
// The arithmetic of Pillow's 8-bit bicubic resize (libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc,
// ImagingResampleHorizontal_8bpc / Vertical_8bpc) and of torchvision's ToTensor + Normalize, which is what the reference's
// `albef` / `albef_randaug` transforms run per image (coarse_grained/fiber/transforms/transform.py:10-45,
// datasets/base_dataset.py:93-110).  Every body is __host__ __device__ and free of shared memory and synchronisation, so
// tests/native/image_pipeline_emul.cpp can replay the exact kernel index math on the CPU (no GPU in the build container).
// Double-precision coefficient arithmetic uses the round-to-nearest intrinsics on the device: a fused multiply-add
// would round once where the library's C code rounds twice and flip a fixed-point coefficient by one unit.
#pragma once

#include <math.h>
#include <stdint.h>

#include "../../include/fiber_b200.h"

#ifdef __CUDACC__
#define FIBER_HD __host__ __device__ __forceinline__
#else
#define FIBER_HD inline
#endif

namespace fiber {
namespace img {

constexpr int kPrecisionBits = 32 - 8 - 2;  // Resample.c: PRECISION_BITS
constexpr int kRowsPerThread = 4;            // horizontal pass: source rows per thread (same coefficients)

FIBER_HD double dmul(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
FIBER_HD double dadd(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
FIBER_HD double dsub(double a, double b) { return dadd(a, -b); }
FIBER_HD double ddiv(double a, double b) {
#ifdef __CUDA_ARCH__
  return __ddiv_rn(a, b);
#else
  return a / b;
#endif
}

// Keys cubic convolution kernel, a = -0.5 (Resample.c: bicubic_filter), same operation order.
FIBER_HD double bicubic(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return dadd(dmul(dmul(dsub(dmul(a + 2.0, x), a + 3.0), x), x), 1.0);
  if (x < 2.0) return dmul(dsub(dmul(dadd(dmul(dsub(x, 5.0), x), 8.0), x), 4.0), a);
  return 0.0;
}

// Taps per output sample for in_size -> out_size (Resample.c: ksize = (int)ceil(support) * 2 + 1).
FIBER_HD int ksize_for(int in_size, int out_size) {
  double scale = ddiv(static_cast<double>(in_size), static_cast<double>(out_size));
  double fs = scale < 1.0 ? 1.0 : scale;
  return static_cast<int>(ceil(dmul(2.0, fs))) * 2 + 1;
}

// Fixed-point coefficients of output sample xx (full-image box: in0 = 0, in1 = in_size), written tap-major:
// k[t * out_size] for t in [0, ksize); bounds[0] = first source sample, bounds[1] = tap count.
FIBER_HD void coeffs_one(int in_size, int out_size, int xx, int ksize, int32_t* k, int32_t* bounds) {
  const double scale = ddiv(static_cast<double>(in_size), static_cast<double>(out_size));
  const double fs = scale < 1.0 ? 1.0 : scale;
  const double support = dmul(2.0, fs);
  const double ss = ddiv(1.0, fs);
  const double center = dadd(0.0, dmul(static_cast<double>(xx) + 0.5, scale));
  int xmin = static_cast<int>(dadd(dsub(center, support), 0.5));
  if (xmin < 0) xmin = 0;
  int xmax = static_cast<int>(dadd(dadd(center, support), 0.5));
  if (xmax > in_size) xmax = in_size;
  const int n = xmax - xmin;
  double ww = 0.0;
  for (int x = 0; x < n; ++x)
    ww = dadd(ww, bicubic(dmul(dadd(dsub(static_cast<double>(x + xmin), center), 0.5), ss)));
  for (int x = 0; x < ksize; ++x) {
    int32_t q = 0;
    if (x < n) {
      double w = bicubic(dmul(dadd(dsub(static_cast<double>(x + xmin), center), 0.5), ss));
      if (ww != 0.0) w = ddiv(w, ww);
      const double f = dmul(w, static_cast<double>(1 << kPrecisionBits));
      q = w < 0 ? static_cast<int32_t>(dadd(-0.5, f)) : static_cast<int32_t>(dadd(0.5, f));
    }
    k[static_cast<int64_t>(x) * out_size] = q;
  }
  bounds[0] = xmin;
  bounds[1] = n;
}

FIBER_HD uint8_t clip8(int32_t acc) {
  int32_t v = acc >> kPrecisionBits;  // arithmetic shift, as the library's lookup index
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// ToTensor + Normalize of one byte value (torchvision: v / 255 then (x - mean) / std, float32, one rounding each).
FIBER_HD float normalize_one(int v, float mean, float stdv) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v), 255.0f), mean), stdv);
#else
  volatile float a = static_cast<float>(v) / 255.0f;
  volatile float b = a - mean;
  return b / stdv;
#endif
}

// Per-image slices of the workspace (offsets filled by fiber_image_transform_plan).
struct Tables {
  const int32_t* kx;  // [ksize_x][out_w]
  const int32_t* bx;  // [out_w][2]
  const int32_t* ky;  // [ksize_y][out_h]
  const int32_t* by;  // [out_h][2]
};
FIBER_HD int64_t table_ints(const fiber_image_desc& d, int out_h, int out_w) {
  return static_cast<int64_t>(d.ksize_x + 2) * out_w + static_cast<int64_t>(d.ksize_y + 2) * out_h;
}
FIBER_HD Tables tables_of(const fiber_image_desc& d, const void* ws, int out_h, int out_w) {
  const int32_t* base = reinterpret_cast<const int32_t*>(static_cast<const uint8_t*>(ws) + d.coef_off);
  Tables t;
  t.kx = base;
  t.bx = t.kx + static_cast<int64_t>(d.ksize_x) * out_w;
  t.ky = t.bx + 2 * out_w;
  t.by = t.ky + static_cast<int64_t>(d.ksize_y) * out_h;
  return t;
}
FIBER_HD int tmp_pitch(int out_w) { return (out_w + 15) & ~15; }
struct alignas(16) Float4 {
  float a, b, c, d;
};

// ---- kernel 1: coefficient tables.  idx in [0, out_w + out_h) per image. ------------------------------------------
FIBER_HD void coeffs_body(const fiber_image_desc* descs, void* ws, int out_h, int out_w, int image, int idx) {
  if (idx >= out_w + out_h) return;
  const fiber_image_desc d = descs[image];
  Tables t = tables_of(d, ws, out_h, out_w);
  if (idx < out_w)
    coeffs_one(d.box_w, out_w, idx, d.ksize_x, const_cast<int32_t*>(t.kx) + idx, const_cast<int32_t*>(t.bx) + 2 * idx);
  else
    coeffs_one(d.box_h, out_h, idx - out_w, d.ksize_y, const_cast<int32_t*>(t.ky) + (idx - out_w),
               const_cast<int32_t*>(t.by) + 2 * (idx - out_w));
}

// ---- kernel 2: horizontal pass.  idx in [0, ceil(box_h / kRowsPerThread) * out_w) per image: one output column of
// kRowsPerThread consecutive source rows; interleaved RGB bytes in, three byte planes [3][box_h][pitch] out. ----------
FIBER_HD void hpass_body(const fiber_image_desc* descs, void* ws, int out_h, int out_w, int image, int64_t idx) {
  const fiber_image_desc d = descs[image];
  const int groups = (d.box_h + kRowsPerThread - 1) / kRowsPerThread;
  if (idx >= static_cast<int64_t>(groups) * out_w) return;
  const int x = static_cast<int>(idx % out_w);
  const int y0 = static_cast<int>(idx / out_w) * kRowsPerThread;
  const Tables t = tables_of(d, ws, out_h, out_w);
  const int xmin = t.bx[2 * x], n = t.bx[2 * x + 1];
  const int pitch = tmp_pitch(out_w);
  uint8_t* tmp = static_cast<uint8_t*>(ws) + d.tmp_off;
  const uint8_t* src = d.src + static_cast<int64_t>(d.box_y + y0) * d.stride + static_cast<int64_t>(d.box_x + xmin) * 3;
  int32_t acc[kRowsPerThread][3];
#pragma unroll
  for (int r = 0; r < kRowsPerThread; ++r) acc[r][0] = acc[r][1] = acc[r][2] = 1 << (kPrecisionBits - 1);
  const int rows = d.box_h - y0 < kRowsPerThread ? d.box_h - y0 : kRowsPerThread;
  for (int tp = 0; tp < n; ++tp) {
    const int32_t k = t.kx[static_cast<int64_t>(tp) * out_w + x];
#pragma unroll
    for (int r = 0; r < kRowsPerThread; ++r) {
      if (r < rows) {
        const uint8_t* p = src + r * d.stride + tp * 3;
        acc[r][0] += static_cast<int32_t>(p[0]) * k;
        acc[r][1] += static_cast<int32_t>(p[1]) * k;
        acc[r][2] += static_cast<int32_t>(p[2]) * k;
      }
    }
  }
  const int64_t plane = static_cast<int64_t>(d.box_h) * pitch;
#pragma unroll
  for (int r = 0; r < kRowsPerThread; ++r) {
    if (r < rows) {
      uint8_t* o = tmp + static_cast<int64_t>(y0 + r) * pitch + x;
      o[0] = clip8(acc[r][0]);
      o[plane] = clip8(acc[r][1]);
      o[2 * plane] = clip8(acc[r][2]);
    }
  }
}

// ---- kernel 3: vertical pass + ToTensor + Normalize (+ horizontal flip).  idx in [0, 3 * out_h * out_w / 4) per
// image: four consecutive output columns of one (channel, row); float32 NCHW out.  lut = [3][256] normalised values. ----
FIBER_HD void vpass_body(const fiber_image_desc* descs, const void* ws, const float* lut, float* out, int out_h, int out_w,
                         int image, int idx) {
  const int w4 = out_w >> 2;
  if (idx >= 3 * out_h * w4) return;
  const fiber_image_desc d = descs[image];
  const int x = (idx % w4) * 4;
  const int yy = (idx / w4) % out_h;
  const int c = idx / (w4 * out_h);
  const Tables t = tables_of(d, ws, out_h, out_w);
  const int ymin = t.by[2 * yy], n = t.by[2 * yy + 1];
  const int pitch = tmp_pitch(out_w);
  const uint8_t* col = static_cast<const uint8_t*>(ws) + d.tmp_off +
                       (static_cast<int64_t>(c) * d.box_h + ymin) * pitch + x;
  int32_t a0, a1, a2, a3;
  a0 = a1 = a2 = a3 = 1 << (kPrecisionBits - 1);
  for (int tp = 0; tp < n; ++tp) {
    const int32_t k = t.ky[static_cast<int64_t>(tp) * out_h + yy];
    const uint32_t p = *reinterpret_cast<const uint32_t*>(col + static_cast<int64_t>(tp) * pitch);  // x % 4 == 0, pitch % 16 == 0
    a0 += static_cast<int32_t>(p & 0xff) * k;
    a1 += static_cast<int32_t>((p >> 8) & 0xff) * k;
    a2 += static_cast<int32_t>((p >> 16) & 0xff) * k;
    a3 += static_cast<int32_t>(p >> 24) * k;
  }
  const float* l = lut + c * 256;
  const float v0 = l[clip8(a0)], v1 = l[clip8(a1)], v2 = l[clip8(a2)], v3 = l[clip8(a3)];
  float* o = out + ((static_cast<int64_t>(image) * 3 + c) * out_h + yy) * out_w;   // out 16-byte aligned, out_w % 4 == 0
  Float4 v;
  if (d.flip) {
    o += out_w - 4 - x;
    v.a = v3; v.b = v2; v.c = v1; v.d = v0;
  } else {
    o += x;
    v.a = v0; v.b = v1; v.c = v2; v.d = v3;
  }
  *reinterpret_cast<Float4*>(o) = v;   // one 128-bit store
}

}  // namespace img
}  // namespace fiber
