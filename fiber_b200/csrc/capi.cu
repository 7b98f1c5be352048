// fiber_b200 — extern "C" surface (see include/fiber_b200.h) and process-wide plumbing.
#include "attention.cuh"
#include "../../include/fiber_b200.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace fiber {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

int gemm_dispatch(const fiber_gemm_args* a, cudaStream_t stream);
int attn_fwd_dispatch(const AttnParams& p, int hd, cudaStream_t stream);
int attn_bwd_dispatch(const AttnParams& p, int hd, cudaStream_t stream);

static AttnParams to_params(const fiber_attn_args* a) {
  AttnParams p;
  p.q = reinterpret_cast<const bf16*>(a->q); p.k = reinterpret_cast<const bf16*>(a->k);
  p.v = reinterpret_cast<const bf16*>(a->v); p.o = reinterpret_cast<bf16*>(a->o);
  p.lse = a->lse;
  p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv; p.ldo = a->ldo;
  p.mode = a->mode; p.G = a->groups; p.nH = a->heads; p.Lq = a->lq; p.Lk = a->lk;
  p.scale = a->scale; p.key_mask = a->key_mask;
  p.H = a->h; p.W = a->w; p.ws = a->ws; p.shift = a->shift; p.bias_table = a->bias_table;
  p.drop_p = a->drop_p; p.seed = a->seed;
  p.d_o = reinterpret_cast<const bf16*>(a->d_o); p.dq = reinterpret_cast<bf16*>(a->dq);
  p.dk = reinterpret_cast<bf16*>(a->dk); p.dv = reinterpret_cast<bf16*>(a->dv);
  p.lddo = a->lddo; p.lddq = a->lddq; p.lddk = a->lddk; p.lddv = a->lddv;
  p.dbias_table = a->dbias_table;
  return p;
}

}  // namespace fiber

extern "C" {

const char* fiber_last_error(void) { return fiber::g_err; }
int fiber_version(void) { return 100; }
int64_t fiber_launch_count(void) { return fiber::g_launches.load(); }

int fiber_init(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    fiber::set_last_error("cudaGetDevice failed: %s", cudaGetErrorString(e));
    return -2;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    fiber::set_last_error("cudaGetDeviceProperties failed: %s", cudaGetErrorString(e));
    return -2;
  }
  if (prop.major != 10) {
    fiber::set_last_error("fiber_b200 requires an sm_100a device (found sm_%d%d)", prop.major,
                          prop.minor);
    return -3;
  }
  return 0;
}

int fiber_gemm(const fiber_gemm_args* args, fiber_stream_t stream) {
  return fiber::gemm_dispatch(args, reinterpret_cast<cudaStream_t>(stream));
}

int fiber_attn_fwd(const fiber_attn_args* a, fiber_stream_t stream) {
  if (!a) { fiber::set_last_error("null args"); return -1; }
  return fiber::attn_fwd_dispatch(fiber::to_params(a), a->head_dim, reinterpret_cast<cudaStream_t>(stream));
}
int fiber_attn_bwd(const fiber_attn_args* a, fiber_stream_t stream) {
  if (!a) { fiber::set_last_error("null args"); return -1; }
  return fiber::attn_bwd_dispatch(fiber::to_params(a), a->head_dim, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
