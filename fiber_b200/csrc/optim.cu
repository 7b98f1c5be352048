// fiber_b200 — fused multi-tensor AdamW: ONE launch updates every parameter of the model.
//
// Reference: coarse_grained/fiber/modules/fiber_utils.py:156-252 builds six parameter groups (weight decay on / off x
// backbone / head / cross-modal learning rates) for transformers.AdamW(betas = (0.9, 0.98), eps = 1e-8).  HF 4.6's
// AdamW.step does, per parameter and in this order,
//     m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  denom = sqrt(v) + eps;
//     p -= lr * sqrt(1 - b2^t) / (1 - b1^t) * m / denom;  p -= lr * wd * p          (decay AFTER the Adam update)
// — 754 tiny foreach groups in eager mode.  Here the host hands over a table of tensors (pointers, sizes, per-tensor lr
// and weight decay: the group a tensor belongs to is just its two scalars) and a table of fixed-size chunks; each CTA
// owns one chunk, reads p / g / m / v once with 16-byte loads and writes p / m / v once (28 bytes per parameter, the
// algorithmic minimum), and can write the bf16 copy of the updated parameter that the GEMMs read next step.
#include "common.cuh"
#include "../../include/fiber_b200.h"

namespace fiber {

void count_launch(int n = 1);

struct AdamWTensor {  // mirrors fiber_adamw_tensor (include/fiber_b200.h)
  float* p;
  const float* g;
  float* m;
  float* v;
  bf16* p_bf16;  // optional same-layout bf16 copy of the updated parameter (nullable)
  long long n;
  float lr, wd;
};
static_assert(sizeof(AdamWTensor) == 56, "table layout is part of the C-ABI");

constexpr int ADAMW_THREADS = 256;

__global__ void __launch_bounds__(ADAMW_THREADS) adamw_multi_kernel(const AdamWTensor* __restrict__ tensors,
                                                                    const int2* __restrict__ chunks, int chunk_elems,
                                                                    float b1, float b2, float eps, float bias_corr) {
  pdl_trigger();
  pdl_wait();
  const int2 ck = chunks[blockIdx.x];  // (tensor index, chunk index inside the tensor)
  const AdamWTensor t = tensors[ck.x];
  const long long lo = static_cast<long long>(ck.y) * chunk_elems;
  const long long hi = min(lo + chunk_elems, t.n);
  const float step = t.lr * bias_corr, decay = 1.0f - t.lr * t.wd;
  const float ob1 = 1.0f - b1, ob2 = 1.0f - b2;
  auto update = [&](float& p, float g, float& m, float& v) {
    m = fmaf(b1, m, ob1 * g);
    v = fmaf(b2, v, ob2 * g * g);
    p = (p - step * m / (sqrtf(v) + eps)) * decay;
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) | reinterpret_cast<uintptr_t>(t.m) |
                     reinterpret_cast<uintptr_t>(t.v)) & 15) == 0 &&
                   (t.p_bf16 == nullptr || (reinterpret_cast<uintptr_t>(t.p_bf16) & 7) == 0);
  long long i = lo + threadIdx.x * 4;  // lo is a multiple of 4 (chunk_elems is)
  if (vec) {
    for (; i + 4 <= hi; i += ADAMW_THREADS * 4) {
      float4 p = *reinterpret_cast<const float4*>(t.p + i);
      const float4 g = *reinterpret_cast<const float4*>(t.g + i);
      float4 m = *reinterpret_cast<const float4*>(t.m + i);
      float4 v = *reinterpret_cast<const float4*>(t.v + i);
      update(p.x, g.x, m.x, v.x);
      update(p.y, g.y, m.y, v.y);
      update(p.z, g.z, m.z, v.z);
      update(p.w, g.w, m.w, v.w);
      *reinterpret_cast<float4*>(t.p + i) = p;
      *reinterpret_cast<float4*>(t.m + i) = m;
      *reinterpret_cast<float4*>(t.v + i) = v;
      if (t.p_bf16) *reinterpret_cast<uint2*>(t.p_bf16 + i) = make_uint2(pack_bf16(p.x, p.y), pack_bf16(p.z, p.w));
    }
  }
  // tail of the chunk (fewer than 4 elements left for this thread), or the whole chunk for unaligned tensors
  const long long end4 = vec ? min(i + 4, hi) : hi;
  for (long long j = vec ? i : lo + threadIdx.x; j < end4; j += vec ? 1 : ADAMW_THREADS) {
    if (vec && j >= hi) break;
    float p = t.p[j], m = t.m[j], v = t.v[j];
    update(p, t.g[j], m, v);
    t.p[j] = p;
    t.m[j] = m;
    t.v[j] = v;
    if (t.p_bf16) t.p_bf16[j] = __float2bfloat16_rn(p);
  }
}

int adamw_multi_dispatch(const void* tensors, const void* chunks, int n_chunks, int chunk_elems, float b1, float b2, float eps,
                         int step, cudaStream_t stream) {
  FIBER_CHECK(tensors != nullptr && chunks != nullptr, "adamw: null table");
  FIBER_CHECK(chunk_elems > 0 && chunk_elems % (ADAMW_THREADS * 4) == 0, "adamw: chunk_elems must be a multiple of %d",
              ADAMW_THREADS * 4);
  FIBER_CHECK(step >= 1, "adamw: step counts from 1");
  if (n_chunks <= 0) return 0;
  // HF AdamW (correct_bias=True): step_size = lr * sqrt(1 - b2^t) / (1 - b1^t)
  const double bc = sqrt(1.0 - pow(static_cast<double>(b2), step)) / (1.0 - pow(static_cast<double>(b1), step));
  FIBER_CUDA(launch_k(adamw_multi_kernel, dim3(n_chunks), dim3(ADAMW_THREADS), 0, stream, reinterpret_cast<const AdamWTensor*>(tensors),
                                                             reinterpret_cast<const int2*>(chunks), chunk_elems, b1, b2, eps,
                                                             static_cast<float>(bc)));
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace fiber

extern "C" int fiber_adamw_multi(const void* tensors, const void* chunks, int32_t n_chunks, int32_t chunk_elems, float beta1,
                                 float beta2, float eps, int32_t step, fiber_stream_t stream) {
  return fiber::adamw_multi_dispatch(tensors, chunks, n_chunks, chunk_elems, beta1, beta2, eps, step,
                                     reinterpret_cast<cudaStream_t>(stream));
}
