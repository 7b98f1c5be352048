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
// k[t * out_size] for t in [0, ksize) (zero past the tap count; ksize may be the padded table height);
// bounds[0] = first source sample, bounds[1] = tap count.
// First source sample and tap count of output sample xx (the bounds[] of Resample.c: precompute_coeffs).
FIBER_HD void bounds_one(int in_size, int out_size, int xx, int* first, int* count) {
  const double scale = ddiv(static_cast<double>(in_size), static_cast<double>(out_size));
  const double support = dmul(2.0, scale < 1.0 ? 1.0 : scale);
  const double center = dadd(0.0, dmul(static_cast<double>(xx) + 0.5, scale));
  int xmin = static_cast<int>(dadd(dsub(center, support), 0.5));
  if (xmin < 0) xmin = 0;
  int xmax = static_cast<int>(dadd(dadd(center, support), 0.5));
  if (xmax > in_size) xmax = in_size;
  *first = xmin;
  *count = xmax - xmin;
}

FIBER_HD void coeffs_one(int in_size, int out_size, int xx, int ksize, int32_t* k, int32_t* bounds) {
  const double scale = ddiv(static_cast<double>(in_size), static_cast<double>(out_size));
  const double fs = scale < 1.0 ? 1.0 : scale;
  const double ss = ddiv(1.0, fs);
  const double center = dadd(0.0, dmul(static_cast<double>(xx) + 0.5, scale));
  int xmin, n;
  bounds_one(in_size, out_size, xx, &xmin, &n);
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

// Read-only global loads (ld.global.nc): source pixels, byte planes and tables are never written by the reading kernel.
template <typename T>
FIBER_HD T ldg(const T* p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
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
  const int32_t* kx;  // [ksize_x rounded up to a multiple of 4][out_w], zero beyond a sample's tap count
  const int32_t* bx;  // [out_w][2]
  const int32_t* ky;  // [ksize_y][out_h]
  const int32_t* by;  // [out_h][2]
};
FIBER_HD int kpad_x(const fiber_image_desc& d) { return (d.ksize_x + 3) & ~3; }  // horizontal taps, zero-padded to fours
FIBER_HD int64_t table_ints(const fiber_image_desc& d, int out_h, int out_w) {
  return static_cast<int64_t>(kpad_x(d) + 2) * out_w + static_cast<int64_t>(d.ksize_y + 2) * out_h;
}
FIBER_HD Tables tables_of(const fiber_image_desc& d, const void* ws, int out_h, int out_w) {
  const int32_t* base = reinterpret_cast<const int32_t*>(static_cast<const uint8_t*>(ws) + d.coef_off);
  Tables t;
  t.kx = base;
  t.bx = t.kx + static_cast<int64_t>(kpad_x(d)) * out_w;
  t.ky = t.bx + 2 * out_w;
  t.by = t.ky + static_cast<int64_t>(d.ksize_y) * out_h;
  return t;
}
FIBER_HD int tmp_pitch(int out_w) { return (out_w + 15) & ~15; }
struct alignas(16) Float4 {
  float a, b, c, d;
};

constexpr int kDefaultVariant = 3;            // "image_variant" when neither the option nor the environment sets it
constexpr int kLutBytes = 3 * 256 * 4;        // the workspace starts with the [3][256] float32 normalisation table

// ---- kernel 1: coefficient tables.  idx in [0, max(out_w + out_h, 768)) per image; image 0 also fills the table of
// normalised values (ToTensor + Normalize of every byte value, per channel) at the head of the workspace. -------------
FIBER_HD void coeffs_body(const fiber_image_desc* descs, void* ws, int out_h, int out_w, int image, int idx,
                          const float* mean, const float* stdv) {
  if (image == 0 && idx < 3 * 256)
    static_cast<float*>(ws)[idx] = normalize_one(idx & 255, mean[idx >> 8], stdv[idx >> 8]);
  if (idx >= out_w + out_h) return;
  const fiber_image_desc d = descs[image];
  Tables t = tables_of(d, ws, out_h, out_w);
  if (idx < out_w)
    coeffs_one(d.box_w, out_w, idx, kpad_x(d), const_cast<int32_t*>(t.kx) + idx, const_cast<int32_t*>(t.bx) + 2 * idx);
  else
    coeffs_one(d.box_h, out_h, idx - out_w, d.ksize_y, const_cast<int32_t*>(t.ky) + (idx - out_w),
               const_cast<int32_t*>(t.by) + 2 * (idx - out_w));
}

// ---- kernel 2: horizontal pass.  idx in [0, ceil(box_h / R) * out_w) per image: one output column of R consecutive
// source rows, all three channels (3 R accumulators per coefficient load); interleaved RGB bytes in, three byte planes
// [3][box_h][pitch] out.  Rows past the box are computed on a duplicate of the last row and not stored, so the tap
// loop carries no predicates.  Handles both source layouts (the word form below is for interleaved bytes only). ----------------------------------------------------
template <int R>
FIBER_HD void hpass_body(const fiber_image_desc* descs, void* ws, int out_h, int out_w, int image, int idx) {
  const fiber_image_desc d = descs[image];
  const int groups = (d.box_h + R - 1) / R;
  if (idx >= groups * out_w) return;
  const int g = idx / out_w;
  const int x = idx - g * out_w;
  const int y0 = g * R;
  const Tables t = tables_of(d, ws, out_h, out_w);
  const int xmin = ldg(t.bx + 2 * x), n = ldg(t.bx + 2 * x + 1);
  const int rows = d.box_h - y0 < R ? d.box_h - y0 : R;
  // byte addressing of both layouts: pixel step ps and channel step cs (3 / 1 interleaved, 1 / chan_stride planar)
  const int64_t ps = d.planar ? 1 : 3, cs = d.planar ? d.chan_stride : 1;
  const uint8_t* p[R];
#pragma unroll
  for (int r = 0; r < R; ++r)
    p[r] = d.src + static_cast<int64_t>(d.box_y + y0 + (r < rows ? r : rows - 1)) * d.stride +
           static_cast<int64_t>(d.box_x + xmin) * ps;
  int32_t acc[R][3];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r][0] = acc[r][1] = acc[r][2] = 1 << (kPrecisionBits - 1);
  const int32_t* kp = t.kx + x;
  for (int tp = 0; tp < n; ++tp, kp += out_w) {
    const int32_t k = ldg(kp);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      acc[r][0] += static_cast<int32_t>(ldg(p[r])) * k;
      acc[r][1] += static_cast<int32_t>(ldg(p[r] + cs)) * k;
      acc[r][2] += static_cast<int32_t>(ldg(p[r] + 2 * cs)) * k;
      p[r] += ps;
    }
  }
  const int pitch = tmp_pitch(out_w);
  const int64_t plane = static_cast<int64_t>(d.box_h) * pitch;
  uint8_t* o = static_cast<uint8_t*>(ws) + d.tmp_off + static_cast<int64_t>(y0) * pitch + x;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (r < rows) {
      o[0] = clip8(acc[r][0]);
      o[plane] = clip8(acc[r][1]);
      o[2 * plane] = clip8(acc[r][2]);
    }
    o += pitch;
  }
}

FIBER_HD uint32_t byte_at(uint32_t w, int i) {  // byte i of w, zero-extended: one PRMT on the device
#ifdef __CUDA_ARCH__
  return __byte_perm(w, 0u, 0x4440u + static_cast<uint32_t>(i));
#else
  return (w >> (8 * i)) & 0xffu;
#endif
}
FIBER_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {  // bytes of the pair (lo, hi) starting sh / 8 bytes in
#ifdef __CUDA_ARCH__
  return __funnelshift_r(lo, hi, sh);
#else
  return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}

// ---- kernel 2, word form ("image_variant" bit 0): the same sums, with the source row read as aligned 32-bit words
// instead of single bytes — a byte load of a warp touches the same two 128-byte lines as a word load, so the pass
// was bound by L1 wavefronts (three per pixel and tap), not by instruction issue.  Four taps = twelve bytes = three
// words per step and row; the row's byte stream starts `off` bytes into its first word, a funnel shift re-aligns it.
// Only words that hold at least one byte of the sample's taps are loaded (such a word lies in the same page as a valid
// byte even when it straddles the image's first or last byte); the coefficient rows are zero-padded to fours, so
// whatever shares a word with the last tap is multiplied by zero. ------------------------------------------------------
template <int R>
FIBER_HD void hpass_words_body(const fiber_image_desc* descs, void* ws, int out_h, int out_w, int image, int idx) {
  const fiber_image_desc d = descs[image];
  const int groups = (d.box_h + R - 1) / R;
  if (idx >= groups * out_w) return;
  const int g = idx / out_w;
  const int x = idx - g * out_w;
  const int y0 = g * R;
  const Tables t = tables_of(d, ws, out_h, out_w);
  const int xmin = ldg(t.bx + 2 * x), n = ldg(t.bx + 2 * x + 1);
  const int rows = d.box_h - y0 < R ? d.box_h - y0 : R;
  const uint32_t* wp[R];   // the next word to load of each row (advances three words per step)
  uint32_t sh[R], carry[R];
  int rem[R];              // words of the row's taps not loaded yet
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const uint8_t* p = d.src + static_cast<int64_t>(d.box_y + y0 + (r < rows ? r : rows - 1)) * d.stride +
                       static_cast<int64_t>(d.box_x + xmin) * 3;
    const uint32_t off = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(p) & 3);
    wp[r] = reinterpret_cast<const uint32_t*>(p - off);
    sh[r] = off * 8;
    rem[r] = (static_cast<int>(off + 3 * n + 3) >> 2) - 1;   // words holding the sample's 3 n bytes, minus the first
    carry[r] = n > 0 ? ldg(wp[r]) : 0u;
    ++wp[r];
  }
  int32_t acc[R][3];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r][0] = acc[r][1] = acc[r][2] = 1 << (kPrecisionBits - 1);
  const int32_t* kp = t.kx + x;
  for (int tp = 0; tp < n; tp += 4, kp += 4 * out_w) {
    const int32_t k0 = ldg(kp), k1 = ldg(kp + out_w), k2 = ldg(kp + 2 * out_w), k3 = ldg(kp + 3 * out_w);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const uint32_t w1 = rem[r] > 0 ? ldg(wp[r]) : 0u;
      const uint32_t w2 = rem[r] > 1 ? ldg(wp[r] + 1) : 0u;
      const uint32_t w3 = rem[r] > 2 ? ldg(wp[r] + 2) : 0u;
      wp[r] += 3;
      rem[r] -= 3;
      const uint32_t a0 = funnel_r(carry[r], w1, sh[r]);   // R0 G0 B0 R1
      const uint32_t a1 = funnel_r(w1, w2, sh[r]);         // G1 B1 R2 G2
      const uint32_t a2 = funnel_r(w2, w3, sh[r]);         // B2 R3 G3 B3
      carry[r] = w3;
      acc[r][0] += static_cast<int32_t>(byte_at(a0, 0)) * k0 + static_cast<int32_t>(byte_at(a0, 3)) * k1 +
                   static_cast<int32_t>(byte_at(a1, 2)) * k2 + static_cast<int32_t>(byte_at(a2, 1)) * k3;
      acc[r][1] += static_cast<int32_t>(byte_at(a0, 1)) * k0 + static_cast<int32_t>(byte_at(a1, 0)) * k1 +
                   static_cast<int32_t>(byte_at(a1, 3)) * k2 + static_cast<int32_t>(byte_at(a2, 2)) * k3;
      acc[r][2] += static_cast<int32_t>(byte_at(a0, 2)) * k0 + static_cast<int32_t>(byte_at(a1, 1)) * k1 +
                   static_cast<int32_t>(byte_at(a2, 0)) * k2 + static_cast<int32_t>(byte_at(a2, 3)) * k3;
    }
  }
  const int pitch = tmp_pitch(out_w);
  const int64_t plane = static_cast<int64_t>(d.box_h) * pitch;
  uint8_t* o = static_cast<uint8_t*>(ws) + d.tmp_off + static_cast<int64_t>(y0) * pitch + x;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (r < rows) {
      o[0] = clip8(acc[r][0]);
      o[plane] = clip8(acc[r][1]);
      o[2 * plane] = clip8(acc[r][2]);
    }
    o += pitch;
  }
}

// W consecutive plane words (4 W pixels) in one load; the address is 4 W-byte aligned by construction.  SMEM: the planes
// were staged in shared memory (plain loads; ld.global.nc is for global addresses only).
template <int W, bool SMEM = false>
FIBER_HD void load_words(const uint8_t* p, uint32_t (&v)[W]) {
#ifdef __CUDA_ARCH__
  if constexpr (W == 4) {
    const uint4 q = SMEM ? *reinterpret_cast<const uint4*>(p) : __ldg(reinterpret_cast<const uint4*>(p));
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else if constexpr (W == 2) {
    const uint2 q = SMEM ? *reinterpret_cast<const uint2*>(p) : __ldg(reinterpret_cast<const uint2*>(p));
    v[0] = q.x; v[1] = q.y;
  } else {
    v[0] = SMEM ? *reinterpret_cast<const uint32_t*>(p) : __ldg(reinterpret_cast<const uint32_t*>(p));
  }
#else
  for (int w = 0; w < W; ++w) v[w] = reinterpret_cast<const uint32_t*>(p)[w];
#endif
}

// The vertical taps of 4 W output columns x .. x + 4 W - 1 of one output row, three channels: col = first tap row of
// channel 0 at column x, plane = bytes between channels; look-up and 128-bit stores (reversed for a flip).
template <int W, bool SMEM>
FIBER_HD void vpass_core(const uint8_t* col, int64_t plane, int pitch, int n, const int32_t* kp, int out_h, int out_w,
                         const float* lut, float* row, int x, int flip) {
  int32_t a[3][W][4];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int w = 0; w < W; ++w) a[c][w][0] = a[c][w][1] = a[c][w][2] = a[c][w][3] = 1 << (kPrecisionBits - 1);
  for (int tp = 0; tp < n; ++tp, kp += out_h, col += pitch) {
    const int32_t k = ldg(kp);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      // x % (4 W) == 0 and pitch % 16 == 0: aligned; the pitch padding keeps a partial last group inside the plane row
      uint32_t pw[W];
      load_words<W, SMEM>(col + c * plane, pw);
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const uint32_t p = pw[w];
        a[c][w][0] += static_cast<int32_t>(p & 0xff) * k;
        a[c][w][1] += static_cast<int32_t>((p >> 8) & 0xff) * k;
        a[c][w][2] += static_cast<int32_t>((p >> 16) & 0xff) * k;
        a[c][w][3] += static_cast<int32_t>(p >> 24) * k;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {   // out 16-byte aligned, out_w % 4 == 0: one 128-bit store per channel and group
    const float* l = lut + c * 256;
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const int xw = x + 4 * w;
      if (xw < out_w) {
        const float v0 = l[clip8(a[c][w][0])], v1 = l[clip8(a[c][w][1])], v2 = l[clip8(a[c][w][2])], v3 = l[clip8(a[c][w][3])];
        Float4 v;
        if (flip) {
          v.a = v3; v.b = v2; v.c = v1; v.d = v0;
        } else {
          v.a = v0; v.b = v1; v.c = v2; v.d = v3;
        }
        *reinterpret_cast<Float4*>(row + static_cast<int64_t>(c) * out_h * out_w + (flip ? out_w - 4 - xw : xw)) = v;
      }
    }
  }
}

// ---- kernel 3: vertical pass + ToTensor + Normalize (+ horizontal flip).  idx in [0, out_h * ceil(out_w / (4 W)))
// per image: 4 W consecutive output columns of one row, all three channels (12 W accumulators per coefficient load);
// float32 NCHW out.  lut = the [3][256] table of normalised values (shared memory on the device).  W = 1 or 2
// ("image_variant" bit 1). -------------------------------------------------------------------------------------------
template <int W>
FIBER_HD void vpass_body(const fiber_image_desc* descs, const void* ws, const float* lut, float* out, int out_h, int out_w,
                         int image, int idx) {
  const int w4 = out_w >> 2;
  const int per_row = (w4 + W - 1) / W;
  if (idx >= out_h * per_row) return;
  const fiber_image_desc d = descs[image];
  const int yy = idx / per_row;
  const int x = (idx - yy * per_row) * (4 * W);
  const Tables t = tables_of(d, ws, out_h, out_w);
  const int ymin = ldg(t.by + 2 * yy), n = ldg(t.by + 2 * yy + 1);
  const int pitch = tmp_pitch(out_w);
  const int64_t plane = static_cast<int64_t>(d.box_h) * pitch;
  const uint8_t* col = static_cast<const uint8_t*>(ws) + d.tmp_off + static_cast<int64_t>(ymin) * pitch + x;
  float* row = out + (static_cast<int64_t>(image) * 3 * out_h + yy) * out_w;
  vpass_core<W, false>(col, plane, pitch, n, t.ky + yy, out_h, out_w, lut, row, x, d.flip);
}

// ---- kernel 3, staged form ("image_variant" bit 4): one CTA = one image and a band of kBandRows output rows.  The
// plane rows the band's taps read ([first, first + count): band rows x scale + support) are copied into shared memory
// with 128-bit loads first, so the tap loop waits on shared-memory latency instead of an L2 round trip per tap
// (the global form runs at 48 % issue-active with 7.8 cycles of long-scoreboard stall per instruction). --------------
constexpr int kBandRows = 16;

FIBER_HD void copy16(uint8_t* dst, const uint8_t* src) {   // 16 aligned bytes, global (read-only path) -> shared
#ifdef __CUDA_ARCH__
  *reinterpret_cast<uint4*>(dst) = __ldg(reinterpret_cast<const uint4*>(src));
#else
  for (int i = 0; i < 16; ++i) dst[i] = src[i];
#endif
}

FIBER_HD void band_span(const int32_t* by, int out_h, int band, int* y0, int* y1, int* first, int* count) {
  *y0 = band * kBandRows;
  *y1 = *y0 + kBandRows < out_h ? *y0 + kBandRows : out_h;
  *first = ldg(by + 2 * *y0);
  *count = ldg(by + 2 * (*y1 - 1)) + ldg(by + 2 * (*y1 - 1) + 1) - *first;
}

// phase 1: thread tid of nthreads copies its share of the band's plane rows ([3][count][pitch] bytes, 16 per load)
FIBER_HD void vstage_load(const fiber_image_desc* descs, const void* ws, uint8_t* smem, int out_h, int out_w, int image,
                          int band, int tid, int nthreads) {
  const fiber_image_desc d = descs[image];
  const Tables t = tables_of(d, ws, out_h, out_w);
  int y0, y1, first, count;
  band_span(t.by, out_h, band, &y0, &y1, &first, &count);
  const int pitch = tmp_pitch(out_w), vec_per_row = pitch / 16;
  const uint8_t* src = static_cast<const uint8_t*>(ws) + d.tmp_off;
  for (int i = tid; i < 3 * count * vec_per_row; i += nthreads) {
    const int v = i % vec_per_row, r = (i / vec_per_row) % count, c = i / (vec_per_row * count);
    copy16(smem + (static_cast<int64_t>(c) * count + r) * pitch + 16 * v,
           src + (static_cast<int64_t>(c) * d.box_h + first + r) * pitch + 16 * v);
  }
}

// phase 2 (after a barrier): the band's outputs, 4 W columns x three channels per item, from the staged rows
template <int W>
FIBER_HD void vstage_compute(const fiber_image_desc* descs, const void* ws, const uint8_t* smem, const float* lut, float* out,
                             int out_h, int out_w, int image, int band, int tid, int nthreads) {
  const fiber_image_desc d = descs[image];
  const Tables t = tables_of(d, ws, out_h, out_w);
  int y0, y1, first, count;
  band_span(t.by, out_h, band, &y0, &y1, &first, &count);
  const int pitch = tmp_pitch(out_w), per_row = ((out_w >> 2) + W - 1) / W;
  for (int i = tid; i < (y1 - y0) * per_row; i += nthreads) {
    const int yy = y0 + i / per_row, x = (i % per_row) * (4 * W);
    const int ymin = ldg(t.by + 2 * yy), n = ldg(t.by + 2 * yy + 1);
    float* row = out + (static_cast<int64_t>(image) * 3 * out_h + yy) * out_w;
    vpass_core<W, true>(smem + static_cast<int64_t>(ymin - first) * pitch + x, static_cast<int64_t>(count) * pitch, pitch, n,
                        t.ky + yy, out_h, out_w, lut, row, x, d.flip);
  }
}

}  // namespace img
}  // namespace fiber
