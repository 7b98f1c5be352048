// fiber_b200 — persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = epilogue( A[M,K] . B[N,K]^T )      bf16 operands, fp32 accumulation in TMEM
//
// One CTA per SM, 10 warps:
//   warp 0      TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1      MMA issuer     (one lane issues tcgen05.mma 128 x BN x 16, commits to mbarriers)
//   warps 2..9  epilogue       (tcgen05.ld TMEM -> regs -> swizzled smem transpose -> coalesced
//                               fused bias / GELU / GELU' / gate / DropPath / residual -> global)
// TMEM holds two BN-column fp32 accumulators so the epilogue of tile i overlaps the MMAs of
// tile i+1.  Operands may be K-major (forward, dgrad) or MN-major (wgrad: dW = dY^T X, both
// operands read straight from their row-major activations, no transposes materialised).
// Split-K work units accumulate with fp32 atomics (wgrad only).
#include "common.cuh"
#include "../../include/fiber_b200.h"

#include <cstdlib>
#include <mutex>

namespace fiber {

void count_launch(int n = 1);

struct GemmParams {
  int M, N, K;
  int tiles_m, tiles_n, splits, kb_total, kb_per_split;
  void* c;
  long long ldc;
  const float* bias;
  const bf16* residual;
  long long ldr;
  const bf16* aux;
  long long ldaux;
  bf16* preact;
  long long ldp;
  const float* scale;
  const float* row_scale;
  int rows_per_scale;
  int act;
  int out_mode;
  int tma_store;  // epilogue variant: thread=row math -> swizzled smem box -> TMA store (no residual/aux)
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 320;  // TMA warp, MMA warp, 8 epilogue warps

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr uint32_t A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr uint32_t B_BYTES = BN * GEMM_BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t STAGING_BYTES = 8 * 32 * 32 * 4;  // 8 epilogue warps x 32x32 fp32
  static constexpr uint32_t SMEM_BYTES =
      1024 /*align slack*/ + STAGES * STAGE_BYTES + STAGING_BYTES + 256 /*barriers*/;
  static constexpr uint32_t TMEM_COLS = 2 * BN;
};

template <int BN, int MN_MAJOR>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmP,
                    const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* stage_base = smem;
  float* staging = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::STAGING_BYTES);
  uint64_t* full_bar = bars;                  // [STAGES]
  uint64_t* empty_bar = bars + STAGES;        // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;    // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tfull_bar[a], 1);
        mbar_init(&tempty_bar[a], 8);
      }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int total_units = p.tiles_m * p.tiles_n * p.splits;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        const int tile = unit % (p.tiles_m * p.tiles_n);
        const int split = unit / (p.tiles_m * p.tiles_n);
        const int m0 = (tile / p.tiles_n) * GEMM_BM;
        const int n0 = (tile % p.tiles_n) * BN;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = stage_base + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if (MN_MAJOR == 0) {
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * GEMM_BK, m0);
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * GEMM_BK, n0);
          } else {
#pragma unroll
            for (int j = 0; j < GEMM_BM / 64; ++j)
              tma_load_2d(sa + j * 8192, &tmA, &full_bar[stage], m0 + j * 64, kb * GEMM_BK);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sb + j * 8192, &tmB, &full_bar[stage], n0 + j * 64, kb * GEMM_BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN, MN_MAJOR, MN_MAJOR);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase[2] = {0, 0};
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      const int split = unit / (p.tiles_m * p.tiles_n);
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
      // wait until the epilogue has drained this accumulator
      mbar_wait(&tempty_bar[acc], acc_phase[acc] ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(stage_base + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            uint64_t adesc, bdesc;
            if (MN_MAJOR == 0) {
              adesc = umma_desc_sw128(sa + k * 32, 16, 1024);
              bdesc = umma_desc_sw128(sb + k * 32, 16, 1024);
            } else {
              adesc = umma_desc_sw128(sa + k * 2048, 8192, 1024);
              bdesc = umma_desc_sw128(sb + k * 2048, 8192, 1024);
            }
            umma_f16_ss(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);            // frees the smem slot when the MMAs retire
          if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);  // accumulator ready for the epilogue
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      acc_phase[acc] ^= 1;
      acc ^= 1;
    }
  } else {
    // ================= epilogue (warps 2..9) =================
    // Two warps per TMEM lane quadrant, each owning half of the accumulator columns.  Per 32-column
    // chunk: tcgen05.ld -> swizzled smem transpose -> all global loads of the chunk issued as one
    // batch -> math -> coalesced stores (8-byte per lane, 64 B contiguous per row).
    const int ew = warp - 2;
    const int q = warp & 3;    // TMEM lane quadrant this warp may access
    const int half = ew >> 2;  // which half of the BN columns
    float* st = staging + ew * (32 * 32);
    int acc = 0;
    uint32_t acc_phase = 0;  // bit a = phase of accumulator a
    const float scale = p.scale ? __ldg(p.scale) : 1.0f;
    const float* __restrict__ bias = p.bias;
    const bf16* __restrict__ residual = p.residual;
    const bf16* __restrict__ aux = p.aux;
    bf16* __restrict__ preact = p.preact;
    const float* __restrict__ row_scale = p.row_scale;
    const int ldc = static_cast<int>(p.ldc), ldr = static_cast<int>(p.ldr), ldaux = static_cast<int>(p.ldaux),
              ldp = static_cast<int>(p.ldp);
    const int act = p.act, out_mode = p.out_mode;
    const int rsub = lane >> 3, cj = lane & 7;
    constexpr int CHUNKS = BN / 64;  // 32-column chunks per warp
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      const int tile = unit % (p.tiles_m * p.tiles_n);
      const int m0 = (tile / p.tiles_n) * GEMM_BM;
      const int n0 = (tile % p.tiles_n) * BN;
      const int ncols = min(BN, p.N - n0);
      const long long row0 = static_cast<long long>(m0) + q * 32;  // first row of this warp
      const int rows_left = p.M - static_cast<int>(row0);           // rows of this warp inside M
      // DropPath / per-sample scale: a warp's 32 rows span at most two samples (rows_per_scale >= 32)
      float rs_lo = scale, rs_hi = scale;
      int rs_split = 1 << 30;
      if (row_scale) {
        const int s0 = static_cast<int>(row0 / p.rows_per_scale);
        rs_split = (s0 + 1) * p.rows_per_scale - static_cast<int>(row0);
        rs_lo = scale * __ldg(row_scale + s0);
        const long long last = row0 + 31 < p.M ? row0 + 31 : p.M - 1;
        rs_hi = scale * __ldg(row_scale + static_cast<int>(last / p.rows_per_scale));
      }
      bf16* c16 = reinterpret_cast<bf16*>(p.c) + row0 * p.ldc + n0;
      float* c32 = reinterpret_cast<float*>(p.c) + row0 * p.ldc + n0;
      const bf16* res_t = residual ? residual + row0 * p.ldr + n0 : nullptr;
      const bf16* aux_t = aux ? aux + row0 * p.ldaux + n0 : nullptr;
      bf16* pre_t = preact ? preact + row0 * p.ldp + n0 : nullptr;

      mbar_wait(&tfull_bar[acc], (acc_phase >> acc) & 1);
      tc_fence_after();
      if (p.tma_store) {
        // ---- thread = accumulator row: bias / GELU / scales on registers, bf16 pack, 16-byte stores into
        //      this warp's 32x64 128B-swizzled box, one TMA store per box (and per output tensor) ----
        uint8_t* box = reinterpret_cast<uint8_t*>(st);  // 4 KB, 1024-byte aligned
        const float sc = lane < rs_split ? rs_lo : rs_hi;
        constexpr int BOXES = BN / 128;  // 64-column boxes per warp
        const int passes = preact ? 2 : 1;
        for (int pass = 0; pass < passes; ++pass) {
          const bool write_pre = preact && pass == 0;
#pragma unroll 1
          for (int bx = 0; bx < BOXES; ++bx) {
            const int cb = (half * BOXES + bx) * 64;  // first column of the box inside the tile
            if (cb < ncols) {
              if (lane == 0) tma_store_wait_read();  // previous store has finished reading the box
              __syncwarp();
#pragma unroll
              for (int cc = 0; cc < 2; ++cc) {
                const int c0 = cb + cc * 32;
                // residual / GELU'-operand rows are read straight in the accumulator layout (one row per
                // thread, 64 contiguous bytes per 32-column chunk), issued before the TMEM load completes
                uint4 rv[4], av[4];
                const bool row_ok = lane < rows_left;
                if (res_t && !write_pre) {
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    rv[j] = (row_ok && c0 + j * 8 < ncols)
                                ? *reinterpret_cast<const uint4*>(res_t + static_cast<long long>(lane) * ldr + c0 + j * 8)
                                : make_uint4(0u, 0u, 0u, 0u);
                }
                if (act == 2) {
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    av[j] = (row_ok && c0 + j * 8 < ncols)
                                ? *reinterpret_cast<const uint4*>(aux_t + static_cast<long long>(lane) * ldaux + c0 + j * 8)
                                : make_uint4(0u, 0u, 0u, 0u);
                }
                uint32_t r[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j) {  // 16-byte chunk = 8 columns
                  float x[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(r[j * 8 + e]);
                  if (bias) {
                    const int col = n0 + c0 + j * 8;
                    if (col < p.N) {
                      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col));
                      const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + col + 4));
                      x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
                      x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
                    }
                  }
                  if (!write_pre) {
                    if (act == 1) {
#pragma unroll
                      for (int e = 0; e < 8; ++e) x[e] = gelu_erf(x[e]);
                    } else if (act == 2) {
                      const uint32_t* au = reinterpret_cast<const uint32_t*>(&av[j]);
#pragma unroll
                      for (int e = 0; e < 4; ++e) {
                        const float2 f = unpack_bf16(au[e]);
                        x[2 * e] *= gelu_erf_grad(f.x);
                        x[2 * e + 1] *= gelu_erf_grad(f.y);
                      }
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) x[e] *= sc;
                    if (res_t) {
                      const uint32_t* ru = reinterpret_cast<const uint32_t*>(&rv[j]);
#pragma unroll
                      for (int e = 0; e < 4; ++e) {
                        const float2 f = unpack_bf16(ru[e]);
                        x[2 * e] += f.x;
                        x[2 * e + 1] += f.y;
                      }
                    }
                  }
                  uint4 o;
                  o.x = pack_bf16(x[0], x[1]); o.y = pack_bf16(x[2], x[3]);
                  o.z = pack_bf16(x[4], x[5]); o.w = pack_bf16(x[6], x[7]);
                  const int chunk = cc * 4 + j;
                  *reinterpret_cast<uint4*>(box + lane * 128 + ((chunk ^ (lane & 7)) << 4)) = o;
                }
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(write_pre ? &tmP : &tmC, box, n0 + cb, m0 + q * 32);
                tma_store_commit();
              }
            }
          }
        }
        tc_fence_before();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        acc_phase ^= (1u << acc);
        acc ^= 1;
        continue;
      }
#pragma unroll 1
      for (int i = 0; i < CHUNKS; ++i) {
        const int c0 = (half * CHUNKS + i) * 32;
        const bool last_chunk = (i == CHUNKS - 1) || (c0 + 32 >= ncols);
        if (c0 < ncols) {
          uint32_t r[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c0, r);
          tmem_ld_wait();
          if (last_chunk) {  // this warp's last TMEM read of the accumulator: hand it back early
            tc_fence_before();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                   __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
            reinterpret_cast<float4*>(st)[lane * 8 + (j ^ (lane & 7))] = v;
          }
          __syncwarp();
          const int col = c0 + cj * 4;
          const bool cvalid = col < ncols;
          float4 v[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + rsub;
            v[it] = reinterpret_cast<const float4*>(st)[rr * 8 + (cj ^ (rr & 7))];
          }
          __syncwarp();  // staging may be overwritten by the next chunk from here on
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias && cvalid) b4 = __ldg(reinterpret_cast<const float4*>(bias + n0 + col));
          uint2 rres[8], raux[8];
          if (res_t) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int rr = it * 4 + rsub;
              rres[it] = (cvalid && rr < rows_left) ? *reinterpret_cast<const uint2*>(res_t + rr * ldr + col)
                                                    : make_uint2(0u, 0u);
            }
          }
          if (act == 2) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int rr = it * 4 + rsub;
              raux[it] = (cvalid && rr < rows_left) ? *reinterpret_cast<const uint2*>(aux_t + rr * ldaux + col)
                                                    : make_uint2(0u, 0u);
            }
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + rsub;
            if (cvalid && rr < rows_left) {
              float x[4] = {v[it].x + b4.x, v[it].y + b4.y, v[it].z + b4.z, v[it].w + b4.w};
              if (pre_t)
                *reinterpret_cast<uint2*>(pre_t + rr * ldp + col) =
                    make_uint2(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]));
              if (act == 1) {
#pragma unroll
                for (int e = 0; e < 4; ++e) x[e] = gelu_erf(x[e]);
              } else if (act == 2) {
                const float2 a0 = unpack_bf16(raux[it].x), a1 = unpack_bf16(raux[it].y);
                x[0] *= gelu_erf_grad(a0.x); x[1] *= gelu_erf_grad(a0.y);
                x[2] *= gelu_erf_grad(a1.x); x[3] *= gelu_erf_grad(a1.y);
              }
              const float sc = rr < rs_split ? rs_lo : rs_hi;
#pragma unroll
              for (int e = 0; e < 4; ++e) x[e] *= sc;
              if (res_t) {
                const float2 a0 = unpack_bf16(rres[it].x), a1 = unpack_bf16(rres[it].y);
                x[0] += a0.x; x[1] += a0.y; x[2] += a1.x; x[3] += a1.y;
              }
              if (out_mode == 0) {
                *reinterpret_cast<uint2*>(c16 + rr * ldc + col) =
                    make_uint2(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]));
              } else if (out_mode == 1) {
                *reinterpret_cast<float4*>(c32 + rr * ldc + col) = make_float4(x[0], x[1], x[2], x[3]);
              } else {
                float* dst = c32 + rr * ldc + col;
#pragma unroll
                for (int e = 0; e < 4; ++e) atomicAdd(dst + e, x[e]);
              }
            }
          }
        } else if (i == 0) {
          // nothing to read for this warp (narrow last tile): still release the accumulator once
          tc_fence_before();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
      }
      acc_phase ^= (1u << acc);
      acc ^= 1;
    }
    if (p.tma_store && lane == 0) tma_store_wait_read();  // smem must outlive the last bulk store
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

// 2-D bf16 tensor map: dim0 = contiguous dimension (inner), dim1 = rows; 128B swizzle.
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer,
                      uint64_t row_pitch_elems, uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  FIBER_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  FIBER_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16B aligned");
  FIBER_CHECK((row_pitch_elems * 2) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes (ld=%llu)",
              (unsigned long long)row_pitch_elems);
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {row_pitch_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FIBER_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return 0;
}

template <int BN, int MN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tp,
                       const GemmParams& p, int grid, cudaStream_t stream) {
  auto kern = gemm_tcgen05_kernel<BN, MN>;
  static bool attr_set = false;
  if (!attr_set) {
    FIBER_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    GemmCfg<BN>::SMEM_BYTES));
    attr_set = true;
  }
  kern<<<grid, GEMM_THREADS, GemmCfg<BN>::SMEM_BYTES, stream>>>(ta, tb, tc, tp, p);
  FIBER_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int gemm_dispatch(const fiber_gemm_args* a, cudaStream_t stream) {
  FIBER_CHECK(a != nullptr, "null args");
  FIBER_CHECK(a->m > 0 && a->n > 0 && a->k > 0, "bad GEMM shape %d x %d x %d", a->m, a->n, a->k);
  FIBER_CHECK(a->a_major == a->b_major, "mixed operand majors are not supported");
  FIBER_CHECK(a->n % 4 == 0, "N must be a multiple of 4 (got %d)", a->n);
  FIBER_CHECK(a->out_mode >= 0 && a->out_mode <= 2, "bad out_mode");
  FIBER_CHECK(a->act != 2 || a->aux != nullptr, "act=2 (GELU grad) needs aux");
  const int mn = a->a_major;
  const int BN = (a->n > 128) ? 256 : 128;

  GemmParams p;
  p.M = a->m; p.N = a->n; p.K = a->k;
  p.tiles_m = (a->m + GEMM_BM - 1) / GEMM_BM;
  p.tiles_n = (a->n + BN - 1) / BN;
  p.kb_total = (a->k + GEMM_BK - 1) / GEMM_BK;
  const int sms = num_sms();
  int splits = a->splits;
  const int tiles = p.tiles_m * p.tiles_n;
  if (a->out_mode != 2) {
    splits = 1;
  } else if (splits <= 0) {
    // enough work units to fill the machine, at least 4 k-blocks each
    splits = (2 * sms + tiles - 1) / tiles;
    const int max_splits = (p.kb_total + 3) / 4;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.c = a->c; p.ldc = a->ldc;
  p.bias = a->bias;
  p.residual = reinterpret_cast<const bf16*>(a->residual); p.ldr = a->ldr;
  p.aux = reinterpret_cast<const bf16*>(a->aux); p.ldaux = a->ldaux;
  p.preact = reinterpret_cast<bf16*>(a->preact); p.ldp = a->ldp;
  p.scale = a->scale; p.row_scale = a->row_scale;
  p.rows_per_scale = a->rows_per_scale > 0 ? a->rows_per_scale : 1;
  p.act = a->act; p.out_mode = a->out_mode;

  CUtensorMap ta, tb;
  if (mn == 0) {
    // A[M, K] (lda), B[N, K] (ldb): inner = K
    if (make_tmap_bf16_2d(&ta, a->a, a->k, a->m, a->lda, GEMM_BK, GEMM_BM)) return -1;
    if (make_tmap_bf16_2d(&tb, a->b, a->k, a->n, a->ldb, GEMM_BK, BN)) return -1;
  } else {
    // A stored [K, M] (lda), B stored [K, N] (ldb): inner = M / N, 64-wide chunks
    if (make_tmap_bf16_2d(&ta, a->a, a->m, a->k, a->lda, 64, GEMM_BK)) return -1;
    if (make_tmap_bf16_2d(&tb, a->b, a->n, a->k, a->ldb, 64, GEMM_BK)) return -1;
  }
  // TMA-store epilogue (thread = accumulator row) whenever the output is bf16 with TMA-compatible pitch
  // and the fused operands are 16-byte addressable; fp32 / atomic outputs use the transposing epilogue
  CUtensorMap tc = ta, tp = ta;
  p.tma_store = 0;
  static const int tma_mode = [] {  // FIBER_TMA_STORE=0 never, 1 only without residual/aux, 2 (default) always
    const char* e = getenv("FIBER_TMA_STORE");
    return e ? atoi(e) : 2;
  }();
  if (tma_mode > 0 && (tma_mode > 1 || (a->residual == nullptr && a->act != 2)) &&
      a->out_mode == 0 && (a->ldc * 2) % 16 == 0 &&
      (a->residual == nullptr || (a->ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(a->residual) & 15) == 0)) &&
      (a->act != 2 || (a->ldaux % 8 == 0 && (reinterpret_cast<uintptr_t>(a->aux) & 15) == 0)) &&
      (reinterpret_cast<uintptr_t>(a->c) & 15) == 0 &&
      (a->preact == nullptr || ((a->ldp * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(a->preact) & 15) == 0)) &&
      (a->row_scale == nullptr || p.rows_per_scale >= 32) && a->n % 8 == 0) {
    if (make_tmap_bf16_2d(&tc, a->c, a->n, a->m, a->ldc, 64, 32)) return -1;
    if (a->preact && make_tmap_bf16_2d(&tp, a->preact, a->n, a->m, a->ldp, 64, 32)) return -1;
    p.tma_store = 1;
  }
  const int units = tiles * p.splits;
  const int grid = units < sms ? units : sms;
  if (BN == 256) {
    return mn ? launch_gemm<256, 1>(ta, tb, tc, tp, p, grid, stream) : launch_gemm<256, 0>(ta, tb, tc, tp, p, grid, stream);
  }
  return mn ? launch_gemm<128, 1>(ta, tb, tc, tp, p, grid, stream) : launch_gemm<128, 0>(ta, tb, tc, tp, p, grid, stream);
}

}  // namespace fiber
