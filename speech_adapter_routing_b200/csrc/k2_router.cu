// k2_router.cu — K2: language-ID router head + adapter-index bookkeeping (HBM-bound, CUDA cores).
//
// Restates LanguageClassifier.forward / _pool_features / predict of the reference
// (src/models/adapter_router.py:251-312; default architecture: LayerNorm(d) per frame -> mean over T ->
//  Linear(d,h1) LN ReLU -> Linear(h1,h2) LN ReLU -> Linear(h2,C) -> softmax -> argmax) and replaces the Python
// list bookkeeping of AdapterRouter.detect_language (:550-566) by device-side idx / perm / seg_starts.
//
// Pass 1 (k2_pool): one read of h [B,T,d].  A warp owns a frame: 16-byte vectorised coalesced loads, fp32
//   statistics by warp-shuffle reduction, accumulates (x-mean)*rstd per lane.  mean_T(gamma*n+beta) is
//   rewritten as gamma*mean_T(n)+beta, so the affine is applied once per utterance in pass 2.
//   Per-CTA partial sums go to the workspace in a fixed order (deterministic, no atomics on data).
// Pass 2 (k2_head): one CTA per utterance: fixed-order reduction of the partials, the 3-layer MLP, softmax,
//   first-max argmax; the last CTA to finish does the stable counting sort (perm, seg_starts).
#include "sar_internal.h"

#include <cuda_bf16.h>

namespace sar {

constexpr int K2_THREADS = 256;
constexpr int K2_WARPS = K2_THREADS / 32;
constexpr float K2_EPS = 1e-5f;
#ifndef K2_FR_PLAIN
#define K2_FR_PLAIN 2
#endif

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Raw 8-element vector as loaded from HBM (converted to fp32 only when consumed, so the NEXT pair of frames can be
// in flight in few registers while the current pair is being reduced).
template <bool FP32>
struct Raw8;
template <>
struct Raw8<false> {
  uint4 v;
  __device__ __forceinline__ void load(const void* row, int elem) {
    v = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(row) + elem));
  }
  __device__ __forceinline__ void zero() { v = make_uint4(0, 0, 0, 0); }
  __device__ __forceinline__ void to_float(float (&out)[8]) const {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      out[2 * i] = __uint_as_float(w[i] << 16);
      out[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
};
template <>
struct Raw8<true> {
  float4 a, b;
  __device__ __forceinline__ void load(const void* row, int elem) {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(row) + elem);
    a = __ldg(p);
    b = __ldg(p + 1);
  }
  __device__ __forceinline__ void zero() { a = b = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void to_float(float (&out)[8]) const {
    out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w;
    out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
  }
};

// grid = (chunks, B); each CTA reduces `rows_per_chunk` frames of one utterance to a d-vector of
// sum_t (x - mean_t) * rstd_t, written to partial[b][chunk][d].
//
// HBM-bound (one read of h).  The frames of a chunk are contiguous in memory, so a producer warp streams them through
// a shared-memory ring with 1-D bulk async copies (cp.async.bulk, K2_ROWS_PER_STAGE frames = 12 KB per copy at
// d = 768 bf16) that complete on mbarriers; up to S stages (72 KB) per CTA are in flight whatever the consumers'
// register pressure — the register-prefetch version (2 frames per warp in flight, 2 CTAs/SM) reached 3.2 TB/s.
// Each of the 8 consumer warps owns one frame of a stage: conflict-free 16-byte smem reads, two-pass fp32
// statistics by warp shuffles, per-lane accumulation; fixed warp/frame assignment and fixed-order reductions keep the
// result deterministic.
constexpr int K2_ROWS_PER_STAGE = 2 * K2_WARPS;   // each consumer warp owns TWO frames of a stage (two independent chains)
constexpr int K2_MAX_STAGES = 4;
constexpr int K2_POOL_THREADS = K2_THREADS + 32;   // 8 consumer warps + 1 producer warp

__device__ __forceinline__ uint32_t k2_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void k2_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(k2_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void k2_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(k2_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void k2_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(k2_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void k2_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(k2_smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void k2_bulk_load(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(k2_smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(k2_smem_u32(bar))
               : "memory");
}

template <int NV, bool FP32>
__global__ void __launch_bounds__(K2_POOL_THREADS)
k2_pool_kernel(const void* __restrict__ h, float* __restrict__ partial, int* __restrict__ done_counter, int T, int d,
               int rows_per_chunk, int n_stages) {
  extern __shared__ __align__(128) uint8_t k2_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, chunk = blockIdx.x;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *done_counter = 0;
  const int t0 = chunk * rows_per_chunk;
  const int t1 = min(T, t0 + rows_per_chunk);
  const size_t esz = FP32 ? 4 : 2;
  const size_t row_bytes = static_cast<size_t>(d) * esz;
  const size_t stage_bytes = K2_ROWS_PER_STAGE * row_bytes;
  uint8_t* ring = k2_smem;                                                  // [n_stages][8 rows][d]
  float* red = reinterpret_cast<float*>(k2_smem + n_stages * stage_bytes);  // [K2_WARPS][d]
  uint64_t* full = reinterpret_cast<uint64_t*>(red + K2_WARPS * d);         // [n_stages]
  uint64_t* empty = full + K2_MAX_STAGES;                                   // [n_stages]
  const uint8_t* base = reinterpret_cast<const uint8_t*>(h) + (static_cast<size_t>(b) * T + t0) * row_bytes;
  const int n_rows = t1 - t0;
  const int n_groups = (n_rows + K2_ROWS_PER_STAGE - 1) / K2_ROWS_PER_STAGE;

  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) {
      k2_mbar_init(&full[s], 1);
      k2_mbar_init(&empty[s], K2_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == K2_WARPS) {
    // ---------------------------------------------------------------- producer
    if (lane == 0) {
      for (int g = 0; g < n_groups; ++g) {
        const int s = g % n_stages;
        if (g >= n_stages) k2_mbar_wait(&empty[s], ((g / n_stages) - 1) & 1);
        const int rows = min(K2_ROWS_PER_STAGE, n_rows - g * K2_ROWS_PER_STAGE);
        const uint32_t bytes = static_cast<uint32_t>(rows * row_bytes);
        k2_mbar_expect_tx(&full[s], bytes);
        k2_bulk_load(ring + s * stage_bytes, base + static_cast<size_t>(g) * stage_bytes, bytes, &full[s]);
      }
    }
    return;
  }

  // ------------------------------------------------------------------ consumers: warp w owns frame g*8 + w
  const float inv_d = 1.0f / static_cast<float>(d);
  float acc[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[v][i] = 0.f;

  for (int g = 0; g < n_groups; ++g) {
    const int s = g % n_stages;
    k2_mbar_wait(&full[s], (g / n_stages) & 1);
    // frames warp and warp + 8 of this stage: the two statistic chains (24 adds + 5 shuffle rounds, twice) are
    // independent, so their latencies overlap — ncu showed a single chain per warp stalled 50 % on wait / short_sb
    const int r0 = g * K2_ROWS_PER_STAGE + warp, r1 = r0 + K2_WARPS;
    const bool live0 = r0 < n_rows, live1 = r1 < n_rows;
    if (live0) {
      const uint8_t* row0 = ring + s * stage_bytes + warp * row_bytes;
      const uint8_t* row1 = live1 ? row0 + K2_WARPS * row_bytes : row0;
      float x0[NV][8], x1[NV][8];
      float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int e = (v * 32 + lane) * 8;
        if (e < d) {
          if (FP32) {
            const float4 a = *reinterpret_cast<const float4*>(row0 + static_cast<size_t>(e) * 4);
            const float4 c = *reinterpret_cast<const float4*>(row0 + static_cast<size_t>(e) * 4 + 16);
            const float4 a1 = *reinterpret_cast<const float4*>(row1 + static_cast<size_t>(e) * 4);
            const float4 c1 = *reinterpret_cast<const float4*>(row1 + static_cast<size_t>(e) * 4 + 16);
            x0[v][0] = a.x; x0[v][1] = a.y; x0[v][2] = a.z; x0[v][3] = a.w;
            x0[v][4] = c.x; x0[v][5] = c.y; x0[v][6] = c.z; x0[v][7] = c.w;
            x1[v][0] = a1.x; x1[v][1] = a1.y; x1[v][2] = a1.z; x1[v][3] = a1.w;
            x1[v][4] = c1.x; x1[v][5] = c1.y; x1[v][6] = c1.z; x1[v][7] = c1.w;
          } else {
            const uint4 q0 = *reinterpret_cast<const uint4*>(row0 + static_cast<size_t>(e) * 2);
            const uint4 q1 = *reinterpret_cast<const uint4*>(row1 + static_cast<size_t>(e) * 2);
            const uint32_t w0[4] = {q0.x, q0.y, q0.z, q0.w};
            const uint32_t w1[4] = {q1.x, q1.y, q1.z, q1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              x0[v][2 * i] = __uint_as_float(w0[i] << 16);
              x0[v][2 * i + 1] = __uint_as_float(w0[i] & 0xFFFF0000u);
              x1[v][2 * i] = __uint_as_float(w1[i] << 16);
              x1[v][2 * i + 1] = __uint_as_float(w1[i] & 0xFFFF0000u);
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) x0[v][i] = x1[v][i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          sum0 += x0[v][i];
          sum1 += x1[v][i];
        }
      }
      const float mean0 = warp_sum(sum0) * inv_d, mean1 = warp_sum(sum1) * inv_d;
      float sq0 = 0.f, sq1 = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const bool live = (v * 32 + lane) * 8 < d;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float d0 = live ? x0[v][i] - mean0 : 0.f, d1 = live ? x1[v][i] - mean1 : 0.f;
          x0[v][i] = d0;
          x1[v][i] = d1;
          sq0 += d0 * d0;
          sq1 += d1 * d1;
        }
      }
      const float rstd0 = 1.0f / sqrtf(warp_sum(sq0) * inv_d + K2_EPS);
      const float rstd1 = live1 ? 1.0f / sqrtf(warp_sum(sq1) * inv_d + K2_EPS) : 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[v][i] += x0[v][i] * rstd0 + x1[v][i] * rstd1;
    }
    __syncwarp();
    if (lane == 0) k2_mbar_arrive(&empty[s]);
  }

  // fixed-order cross-warp reduction (consumer warps only: named barrier over 256 threads)
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int e = (v * 32 + lane) * 8;
    if (e < d) {
#pragma unroll
      for (int i = 0; i < 8; ++i) red[warp * d + e + i] = acc[v][i];
    }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(K2_THREADS) : "memory");
  float* out = partial + (static_cast<size_t>(b) * gridDim.x + chunk) * d;
  for (int j = threadIdx.x; j < d; j += K2_THREADS) {
    float sacc = 0.f;
#pragma unroll
    for (int w = 0; w < K2_WARPS; ++w) sacc += red[w * d + j];
    out[j] = sacc;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// k2_pool2 — the bf16 pooling pass, second generation (and, with PRE_LN, the fusion SURVEY §8(f)-4 asks for: encoder
// final LayerNorm -> LID LayerNorm -> mean over T straight from the last encoder layer's residual stream, so the LID
// pass never writes and re-reads the [B, 1500, d] encoder output; reference: $HF/modeling_whisper.py:643 +
// src/models/adapter_router.py:268, :229).
//
// The ring version above ran at 3.1 TB/s (47 % of the measured HBM peak): ncu showed its 8 consumer warps per CTA
// stalled on their own statistic chains.  What the plain LayerNorm kernel (5.6 TB/s) has and the ring did not is
// OCCUPANCY with cheap rows: here a warp owns a frame, loads it straight into 12 registers (d = 768; 16-byte
// lane-strided loads, full 128-byte lines per request), two frames per warp in flight, all arithmetic on the packed
// fp32 pipe (add / fma .f32x2: one instruction per bf16 pair — ~130 instructions per frame instead of ~350), 16
// warps per SM.  Per-lane accumulators, fixed-order cross-warp reduction, fixed-order partials: deterministic.
__device__ __forceinline__ float2 k2_unpack2(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
}
__device__ __forceinline__ float2 k2_fadd2(float2 a, float2 b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 k2_fmul2(float2 a, float2 b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 k2_ffma2(float2 a, float2 b, float2 c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(d)
      : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)), "l"(*reinterpret_cast<uint64_t*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ uint32_t k2_pack2(float2 v) {
  __nv_bfloat162 t = __floats2bfloat162_rn(v.x, v.y);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <int NV, bool PRE_LN>
__global__ void __launch_bounds__(K2_THREADS, 2)
k2_pool2_kernel(const uint4* __restrict__ h, const uint4* __restrict__ g1, const uint4* __restrict__ b1, float eps1,
                float* __restrict__ partial, int* __restrict__ done_counter, int T, int nvec, float inv_d,
                int rows_per_chunk) {
  extern __shared__ __align__(16) float k2_red[];                 // [K2_WARPS][nvec * 8]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, chunk = blockIdx.x;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *done_counter = 0;
  const int t0 = chunk * rows_per_chunk;
  const int t1 = min(T, t0 + rows_per_chunk);
  const int d = nvec * 8;
  uint4 gq[PRE_LN ? NV : 1], bq[PRE_LN ? NV : 1];
  if constexpr (PRE_LN) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      gq[i] = c < nvec ? __ldg(g1 + c) : make_uint4(0u, 0u, 0u, 0u);
      bq[i] = c < nvec ? __ldg(b1 + c) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  float2 acc[NV][4];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[i][e] = make_float2(0.f, 0.f);
  const uint4* base = h + static_cast<size_t>(b) * T * nvec;

  // FR frames per warp are processed in LOCKSTEP: their statistic chains (a serial packed-fp32 sum and a 5-step shuffle
  // reduction, twice per LayerNorm) are independent, so the scheduler interleaves them.  Processing the frames one after
  // the other — whatever number of loads was in flight — left a warp issuing one instruction per ~8 cycles and the pass
  // at 3.4 TB/s (ncu: warps active 21 %, issue-active 38-41 %).  d <= 768 rows are 12 registers per frame: four at a
  // time; with the encoder LayerNorm's gamma / beta resident two; wider rows one.
  constexpr int FR = (NV <= 3) ? (PRE_LN ? 1 : K2_FR_PLAIN) : 1;
  // mean and 1/sqrt(var + eps) of FR frames held as packed bf16 pairs (two-pass, fp32)
  auto stats = [&](const uint4 (&v)[FR][NV], float eps, float2 (&nm)[FR], float2 (&rs)[FR]) {
    float2 sum[FR];
#pragma unroll
    for (int u = 0; u < FR; ++u) sum[u] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {          // lanes past the row hold zeros: they add nothing
#pragma unroll
      for (int u = 0; u < FR; ++u) {
        sum[u] = k2_fadd2(sum[u], k2_unpack2(v[u][i].x));
        sum[u] = k2_fadd2(sum[u], k2_unpack2(v[u][i].y));
        sum[u] = k2_fadd2(sum[u], k2_unpack2(v[u][i].z));
        sum[u] = k2_fadd2(sum[u], k2_unpack2(v[u][i].w));
      }
    }
    float m[FR];
#pragma unroll
    for (int u = 0; u < FR; ++u) m[u] = sum[u].x + sum[u].y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < FR; ++u) m[u] += __shfl_xor_sync(0xffffffffu, m[u], o);
    float2 q[FR];
#pragma unroll
    for (int u = 0; u < FR; ++u) {
      nm[u] = make_float2(-m[u] * inv_d, -m[u] * inv_d);
      q[u] = make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + 32 * i < nvec) {
#pragma unroll
        for (int u = 0; u < FR; ++u) {
          float2 t;
          t = k2_fadd2(k2_unpack2(v[u][i].x), nm[u]); q[u] = k2_ffma2(t, t, q[u]);
          t = k2_fadd2(k2_unpack2(v[u][i].y), nm[u]); q[u] = k2_ffma2(t, t, q[u]);
          t = k2_fadd2(k2_unpack2(v[u][i].z), nm[u]); q[u] = k2_ffma2(t, t, q[u]);
          t = k2_fadd2(k2_unpack2(v[u][i].w), nm[u]); q[u] = k2_ffma2(t, t, q[u]);
        }
      }
    }
    float qs[FR];
#pragma unroll
    for (int u = 0; u < FR; ++u) qs[u] = q[u].x + q[u].y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < FR; ++u) qs[u] += __shfl_xor_sync(0xffffffffu, qs[u], o);
#pragma unroll
    for (int u = 0; u < FR; ++u) {
      const float rstd = 1.0f / sqrtf(qs[u] * inv_d + eps);
      rs[u] = make_float2(rstd, rstd);
    }
  };

  // register double buffering: the loads of the NEXT group of FR frames are issued before this group is reduced, so a
  // warp always has a group in flight (without it a warp alternated between waiting ~2 us for its loads and ~450
  // instructions of arithmetic, and 3-4 such warps per scheduler did not keep HBM busy)
  auto load_group = [&](int t, uint4 (&v)[FR][NV]) {
#pragma unroll
    for (int u = 0; u < FR; ++u) {
      const int tu = t + u * K2_WARPS;
      const uint4* r = base + static_cast<size_t>(tu < t1 ? tu : t0) * nvec;   // dead slot: any valid frame, weight 0
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        v[u][i] = c < nvec ? __ldcs(r + c) : make_uint4(0u, 0u, 0u, 0u);
      }
    }
  };
  uint4 vn[FR][NV];
  if (t0 + warp < t1) load_group(t0 + warp, vn);
#pragma unroll 1
  for (int t = t0 + warp; t < t1; t += FR * K2_WARPS) {
    uint4 v[FR][NV];
    bool live[FR];
#pragma unroll
    for (int u = 0; u < FR; ++u) {
      live[u] = t + u * K2_WARPS < t1;                     // warp-uniform
#pragma unroll
      for (int i = 0; i < NV; ++i) v[u][i] = vn[u][i];
    }
    if (t + FR * K2_WARPS < t1) load_group(t + FR * K2_WARPS, vn);
    float2 nm[FR], rs[FR];
    if constexpr (PRE_LN) {
      // the encoder's final LayerNorm: y = bf16((x - mean) * rstd * gamma + beta) — exactly the tensor the unfused
      // path would have written (sar_layernorm_fwd) and this kernel would have read back
      stats(v, eps1, nm, rs);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (lane + 32 * i < nvec) {
#pragma unroll
          for (int u = 0; u < FR; ++u) {
            v[u][i].x = k2_pack2(k2_ffma2(k2_fmul2(k2_fadd2(k2_unpack2(v[u][i].x), nm[u]), rs[u]), k2_unpack2(gq[i].x), k2_unpack2(bq[i].x)));
            v[u][i].y = k2_pack2(k2_ffma2(k2_fmul2(k2_fadd2(k2_unpack2(v[u][i].y), nm[u]), rs[u]), k2_unpack2(gq[i].y), k2_unpack2(bq[i].y)));
            v[u][i].z = k2_pack2(k2_ffma2(k2_fmul2(k2_fadd2(k2_unpack2(v[u][i].z), nm[u]), rs[u]), k2_unpack2(gq[i].z), k2_unpack2(bq[i].z)));
            v[u][i].w = k2_pack2(k2_ffma2(k2_fmul2(k2_fadd2(k2_unpack2(v[u][i].w), nm[u]), rs[u]), k2_unpack2(gq[i].w), k2_unpack2(bq[i].w)));
          }
        }
      }
    }
    stats(v, K2_EPS, nm, rs);               // the LID head's LayerNorm (affine applied once per utterance, in the head)
#pragma unroll
    for (int u = 0; u < FR; ++u)
      if (!live[u]) rs[u] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + 32 * i < nvec) {
#pragma unroll
        for (int u = 0; u < FR; ++u) {
          acc[i][0] = k2_ffma2(k2_fadd2(k2_unpack2(v[u][i].x), nm[u]), rs[u], acc[i][0]);
          acc[i][1] = k2_ffma2(k2_fadd2(k2_unpack2(v[u][i].y), nm[u]), rs[u], acc[i][1]);
          acc[i][2] = k2_ffma2(k2_fadd2(k2_unpack2(v[u][i].z), nm[u]), rs[u], acc[i][2]);
          acc[i][3] = k2_ffma2(k2_fadd2(k2_unpack2(v[u][i].w), nm[u]), rs[u], acc[i][3]);
        }
      }
    }
  }

  // fixed-order cross-warp reduction
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int e = (lane + 32 * i) * 8;
    if (e < d) {
      float4* dst = reinterpret_cast<float4*>(k2_red + warp * d + e);
      dst[0] = make_float4(acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y);
      dst[1] = make_float4(acc[i][2].x, acc[i][2].y, acc[i][3].x, acc[i][3].y);
    }
  }
  __syncthreads();
  float* out = partial + (static_cast<size_t>(b) * gridDim.x + chunk) * d;
  for (int j = threadIdx.x; j < d; j += K2_THREADS) {
    float sacc = 0.f;
#pragma unroll
    for (int w = 0; w < K2_WARPS; ++w) sacc += k2_red[w * d + j];
    out[j] = sacc;
  }
}

// The head is latency-bound (one CTA per utterance walks W1 [h1, d] fp32 through L2): 32 warps per CTA put four times as
// many independent row pairs in flight as 8 did (48 us -> see DESIGN.md §4).
constexpr int K2H_THREADS = 1024;
constexpr int K2H_WARPS = K2H_THREADS / 32;

__device__ __forceinline__ float block_sum(float v, float* scratch) {
  // scratch: K2H_WARPS floats.  All threads receive the same fixed-order sum.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < K2H_WARPS; ++w) s += scratch[w];
  return s;
}

// y[o] = W[o,:]·x + bias[o].  A warp owns two output rows per iteration and keeps 8 independent float4 loads in
// flight (the layers are tiny; latency, not bandwidth, is what matters).  x lives in shared memory.
__device__ __forceinline__ void dense_rows(const float* __restrict__ W, const float* __restrict__ bias,
                                           const float* x, float* y, int n_out, int n_in) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((n_in & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0) {
    const int n4 = n_in >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (int o = 2 * warp; o < n_out; o += 2 * K2H_WARPS) {
      const bool two = (o + 1) < n_out;
      const float4* w0 = reinterpret_cast<const float4*>(W + static_cast<size_t>(o) * n_in);
      const float4* w1 = reinterpret_cast<const float4*>(W + static_cast<size_t>(two ? o + 1 : o) * n_in);
      float s0 = 0.f, s1 = 0.f;
      for (int j0 = 0; j0 < n4; j0 += 128) {
        float4 a[4], c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + u * 32 + lane;
          const bool in = j < n4;
          a[u] = in ? __ldg(w0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          c[u] = in ? __ldg(w1 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + u * 32 + lane;
          if (j < n4) {
            const float4 xv = x4[j];
            s0 += a[u].x * xv.x + a[u].y * xv.y + a[u].z * xv.z + a[u].w * xv.w;
            s1 += c[u].x * xv.x + c[u].y * xv.y + c[u].z * xv.z + c[u].w * xv.w;
          }
        }
      }
      s0 = warp_sum(s0);
      s1 = warp_sum(s1);
      if (lane == 0) {
        y[o] = s0 + __ldg(bias + o);
        if (two) y[o + 1] = s1 + __ldg(bias + o + 1);
      }
    }
  } else {
    for (int o = warp; o < n_out; o += K2H_WARPS) {
      const float* w = W + static_cast<size_t>(o) * n_in;
      float s = 0.f;
      for (int j = lane; j < n_in; j += 32) s += __ldg(w + j) * x[j];
      s = warp_sum(s);
      if (lane == 0) y[o] = s + __ldg(bias + o);
    }
  }
}

// In-place LayerNorm (biased variance, eps) + ReLU over v[0..n).
__device__ __forceinline__ void ln_relu(float* v, const float* __restrict__ g, const float* __restrict__ be, int n,
                                        float* scratch) {
  float s = 0.f;
  for (int j = threadIdx.x; j < n; j += K2H_THREADS) s += v[j];
  const float mean = block_sum(s, scratch) / static_cast<float>(n);
  float q = 0.f;
  for (int j = threadIdx.x; j < n; j += K2H_THREADS) {
    const float dlt = v[j] - mean;
    q += dlt * dlt;
  }
  const float rstd = 1.0f / sqrtf(block_sum(q, scratch) / static_cast<float>(n) + K2_EPS);
  for (int j = threadIdx.x; j < n; j += K2H_THREADS) {
    const float y = (v[j] - mean) * rstd * __ldg(g + j) + __ldg(be + j);
    v[j] = y > 0.f ? y : 0.f;
  }
  __syncthreads();
}

struct K2HeadParams {
  const float* partial;
  int chunks;
  const float *ln_w, *ln_b, *W1, *b1, *g1, *be1, *W2, *b2, *g2, *be2, *W3, *b3;
  int B, T, d, h1, h2, C;
  float* logits;
  float* probs;
  int32_t* idx;
  int32_t* perm;
  int32_t* seg_starts;
  int* done_counter;
};

__global__ void __launch_bounds__(K2H_THREADS) k2_head_kernel(const K2HeadParams p) {
  extern __shared__ __align__(16) float sm[];  // pooled[d] | a1[h1] | a2[h2] | lg[64] | scratch[K2H_WARPS] | counts
  float* pooled = sm;
  float* a1 = pooled + ((p.d + 3) & ~3);
  float* a2 = a1 + ((p.h1 + 3) & ~3);
  float* lg = a2 + ((p.h2 + 3) & ~3);
  float* scratch = lg + 64;
  __shared__ int is_last;
  const int b = blockIdx.x;

  const float inv_T = 1.0f / static_cast<float>(p.T);
  for (int j = threadIdx.x; j < p.d; j += K2H_THREADS) {
    const float* src = p.partial + static_cast<size_t>(b) * p.chunks * p.d + j;
    float s = 0.f;
    for (int c = 0; c < p.chunks; ++c) s += src[static_cast<size_t>(c) * p.d];
    pooled[j] = __ldg(p.ln_w + j) * (s * inv_T) + __ldg(p.ln_b + j);
  }
  __syncthreads();
  dense_rows(p.W1, p.b1, pooled, a1, p.h1, p.d);
  __syncthreads();
  ln_relu(a1, p.g1, p.be1, p.h1, scratch);
  dense_rows(p.W2, p.b2, a1, a2, p.h2, p.h1);
  __syncthreads();
  ln_relu(a2, p.g2, p.be2, p.h2, scratch);
  dense_rows(p.W3, p.b3, a2, lg, p.C, p.h2);
  __syncthreads();
  if (threadIdx.x == 0) {
    float mx = lg[0];
    for (int c = 1; c < p.C; ++c) mx = fmaxf(mx, lg[c]);
    float den = 0.f;
    for (int c = 0; c < p.C; ++c) den += expf(lg[c] - mx);
    // argmax is taken over probs like the reference (adapter_router.py:311); strict > keeps the first maximal
    // index (torch.argmax tie rule).
    float best = -1.f;
    int parg = 0;
    for (int c = 0; c < p.C; ++c) {
      const float pr = expf(lg[c] - mx) / den;
      p.logits[static_cast<size_t>(b) * p.C + c] = lg[c];
      p.probs[static_cast<size_t>(b) * p.C + c] = pr;
      if (pr > best) {
        best = pr;
        parg = c;
      }
    }
    p.idx[b] = parg;
    __threadfence();
    is_last = (atomicAdd(p.done_counter, 1) == p.B - 1);
  }
  __syncthreads();
  if (is_last) {
    // Stable counting sort of the utterances by class, by the whole CTA (one thread walking idx[] through L2 cost
    // ~40 us at B = 64): for each class, a ballot-based block scan over the utterances in index order.
    __threadfence();
    __shared__ int warp_tot[K2H_WARPS];
    const volatile int32_t* vidx = p.idx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int running = 0;
    for (int c = 0; c < p.C; ++c) {
      if (threadIdx.x == 0) p.seg_starts[c] = running;
      for (int i0 = 0; i0 < p.B; i0 += K2H_THREADS) {
        const int i = i0 + threadIdx.x;
        const bool flag = i < p.B && vidx[i] == c;
        const unsigned ballot = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) warp_tot[warp] = __popc(ballot);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < K2H_WARPS; ++w) {
          before += w < warp ? warp_tot[w] : 0;
          total += warp_tot[w];
        }
        if (flag) p.perm[running + before + __popc(ballot & ((1u << lane) - 1u))] = i;
        running += total;
        __syncthreads();
      }
    }
    if (threadIdx.x == 0) p.seg_starts[p.C] = running;
  }
}

int64_t k2_workspace_bytes(int64_t B, int64_t T, int64_t d) {
  if (B <= 0 || T <= 0 || d <= 0) return SAR_EINVAL;
  const int64_t max_chunks = (T + 15) / 16;  // rows_per_chunk >= 16
  return 256 + B * max_chunks * d * 4;
}

static int k2_rows_per_chunk(int B, int T, int num_sms) {
  // One wave of long-lived CTAs (2 per SM: the shared-memory ring of each holds up to 72 KB in flight): as many chunks
  // per utterance as fit in 2*num_sms CTAs, so every CTA streams dozens of 8-frame groups and its ring stays full.
  // Short chunks (the register-prefetch version used 16..128 frames) pay the pipeline fill once per CTA and measured
  // 2.3 TB/s with this kernel.
  const int slots = num_sms * 2;
  int chunks = slots / (B > 0 ? B : 1);
  const int max_chunks = (T + K2_ROWS_PER_STAGE - 1) / K2_ROWS_PER_STAGE;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  int rows = (T + chunks - 1) / chunks;
  rows = (rows + K2_ROWS_PER_STAGE - 1) / K2_ROWS_PER_STAGE * K2_ROWS_PER_STAGE;
  if (rows < 16) rows = 16;   // workspace sizing assumes >= 16 frames per chunk
  return rows;
}

static int k2_rows_per_chunk2(int B, int T, int num_sms) {
  // k2_pool2: two CTAs per SM.  Pick the number of chunks per utterance (1x..4x the slot count) that fills whole waves
  // best — B = 64 on 148 SMs: 4 chunks leave 13 % of the slots idle (256 CTAs on 296 slots), 9 chunks give 576 CTAs =
  // 1.95 waves — preferring fewer, longer chunks on ties.
  const int slots = num_sms * 2;
  int best_c = 1;
  double best_eff = 0.0;
  const int c_lo = slots / B > 1 ? slots / B : 1, c_hi = 4 * slots / B > 1 ? 4 * slots / B : 1;
  for (int c = c_lo; c <= c_hi; ++c) {
    if ((T + c - 1) / c < 32) break;                       // at least 32 frames per chunk
    const long long ctas = static_cast<long long>(B) * c;
    const double eff = static_cast<double>(ctas) / (static_cast<double>((ctas + slots - 1) / slots) * slots);
    if (eff > best_eff + 0.02) {
      best_eff = eff;
      best_c = c;
    }
  }
  int rows = (T + best_c - 1) / best_c;
  if (rows < 16) rows = 16;                                // workspace sizing assumes >= 16 frames per chunk
  return rows;
}

template <bool FP32>
static int k2_launch_pool(const K2Args& a, float* partial, int* counter, int chunks, int rpc, cudaStream_t stream) {
  const int nv = (a.d + 255) / 256;
  const dim3 grid(chunks, a.B);
  const DeviceInfo& dev = device_info();
  const size_t stage_bytes = static_cast<size_t>(K2_ROWS_PER_STAGE) * a.d * (FP32 ? 4 : 2);
  const size_t fixed = static_cast<size_t>(K2_WARPS) * a.d * sizeof(float) + 2 * K2_MAX_STAGES * sizeof(uint64_t) + 128;
  // two CTAs per SM: each may use half of the shared memory
  const size_t budget = static_cast<size_t>(dev.max_smem_optin) / 2 - 2048;
  int n_stages = fixed < budget ? static_cast<int>((budget - fixed) / stage_bytes) : 0;
  if (n_stages > K2_MAX_STAGES) n_stages = K2_MAX_STAGES;
  if (n_stages < 2) {   // very wide fp32 rows: one CTA per SM
    n_stages = static_cast<int>((static_cast<size_t>(dev.max_smem_optin) - 2048 - fixed) / stage_bytes);
    if (n_stages > K2_MAX_STAGES) n_stages = K2_MAX_STAGES;
    if (n_stages < 1) return fail(SAR_EINVAL, "k2: row too wide for the shared-memory ring");
  }
  const size_t smem = n_stages * stage_bytes + fixed;
#define SAR_K2_CASE(N)                                                                                          \
  case N:                                                                                                       \
    if (smem > 48 * 1024)                                                                                       \
      cudaFuncSetAttribute(k2_pool_kernel<N, FP32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    k2_pool_kernel<N, FP32><<<grid, K2_POOL_THREADS, smem, stream>>>(a.h, partial, counter, a.T, a.d, rpc, n_stages); \
    break;
  switch (nv) {
    SAR_K2_CASE(1) SAR_K2_CASE(2) SAR_K2_CASE(3) SAR_K2_CASE(4) SAR_K2_CASE(5) SAR_K2_CASE(6) SAR_K2_CASE(7)
    SAR_K2_CASE(8)
    default: return fail(SAR_EINVAL, "k2: d must be <= 2048");
  }
#undef SAR_K2_CASE
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "k2: pool launch");
  return SAR_OK;
}

static int k2_launch_pool2(const K2Args& a, float* partial, int* counter, int chunks, int rpc, cudaStream_t stream) {
  const int nvec = a.d / 8;
  const int nv = (nvec + 31) / 32;
  const dim3 grid(chunks, a.B);
  const size_t smem = static_cast<size_t>(K2_WARPS) * a.d * sizeof(float);
  const uint4* hp = static_cast<const uint4*>(a.h);
  const uint4* gp = static_cast<const uint4*>(a.pre_ln_w);
  const uint4* bp = static_cast<const uint4*>(a.pre_ln_b);
  const float inv_d = 1.0f / static_cast<float>(a.d);
  const bool pre = a.pre_ln_w != nullptr;
  if (pre && (!a.pre_ln_b || ((reinterpret_cast<uintptr_t>(a.pre_ln_w) | reinterpret_cast<uintptr_t>(a.pre_ln_b)) & 15)))
    return fail(SAR_EINVAL, "k2: fused encoder LayerNorm needs 16-byte aligned bf16 gamma and beta");
#define SAR_K2P_CASE(N)                                                                                                  \
  case N:                                                                                                                \
    if (pre) {                                                                                                           \
      if (smem > 48 * 1024)                                                                                              \
        cudaFuncSetAttribute(k2_pool2_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
      k2_pool2_kernel<N, true><<<grid, K2_THREADS, smem, stream>>>(hp, gp, bp, a.pre_ln_eps, partial, counter, a.T, nvec, \
                                                                   inv_d, rpc);                                          \
    } else {                                                                                                             \
      if (smem > 48 * 1024)                                                                                              \
        cudaFuncSetAttribute(k2_pool2_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
      k2_pool2_kernel<N, false><<<grid, K2_THREADS, smem, stream>>>(hp, nullptr, nullptr, 0.f, partial, counter, a.T,    \
                                                                    nvec, inv_d, rpc);                                   \
    }                                                                                                                    \
    break;
  switch (nv) {
    SAR_K2P_CASE(1) SAR_K2P_CASE(2) SAR_K2P_CASE(3) SAR_K2P_CASE(4) SAR_K2P_CASE(5) SAR_K2P_CASE(6) SAR_K2P_CASE(7)
    SAR_K2P_CASE(8)
    default: return fail(SAR_EINVAL, "k2: d must be <= 2048");
  }
#undef SAR_K2P_CASE
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "k2: pool launch");
  return SAR_OK;
}

int k2_router_fwd(const K2Args& a, cudaStream_t stream) {
  if (!a.h || !a.ln_w || !a.ln_b || !a.W1 || !a.b1 || !a.g1 || !a.be1 || !a.W2 || !a.b2 || !a.g2 || !a.be2 ||
      !a.W3 || !a.b3 || !a.logits || !a.probs || !a.idx || !a.perm || !a.seg_starts || !a.ws)
    return fail(SAR_EINVAL, "k2: null pointer");
  if (a.B <= 0 || a.T <= 0) return fail(SAR_EINVAL, "k2: B and T must be positive");
  if (a.d % 8 || a.d <= 0 || a.d > 2048) return fail(SAR_EINVAL, "k2: d must be a multiple of 8, <= 2048");
  if (a.h1 <= 0 || a.h2 <= 0 || a.h1 > 1024 || a.h2 > 1024 || a.C <= 0 || a.C > 64)
    return fail(SAR_EINVAL, "k2: h1,h2 must be in [1,1024], C in [1,64]");
  if ((reinterpret_cast<uintptr_t>(a.h) | reinterpret_cast<uintptr_t>(a.ws)) & 15)
    return fail(SAR_EINVAL, "k2: h and ws must be 16-byte aligned");
  const DeviceInfo& dev = device_info();
  const int rpc = a.h_is_fp32 ? k2_rows_per_chunk(a.B, a.T, dev.num_sms) : k2_rows_per_chunk2(a.B, a.T, dev.num_sms);
  const int chunks = (a.T + rpc - 1) / rpc;
  int* counter = reinterpret_cast<int*>(a.ws);
  float* partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a.ws) + 256);
  int rc;
  if (a.h_is_fp32) {
    if (a.pre_ln_w) return fail(SAR_EINVAL, "k2: the fused encoder LayerNorm needs bf16 states");
    rc = k2_launch_pool<true>(a, partial, counter, chunks, rpc, stream);
  } else {
    rc = k2_launch_pool2(a, partial, counter, chunks, rpc, stream);
  }
  if (rc) return rc;

  K2HeadParams p{};
  p.partial = partial; p.chunks = chunks;
  p.ln_w = a.ln_w; p.ln_b = a.ln_b; p.W1 = a.W1; p.b1 = a.b1; p.g1 = a.g1; p.be1 = a.be1;
  p.W2 = a.W2; p.b2 = a.b2; p.g2 = a.g2; p.be2 = a.be2; p.W3 = a.W3; p.b3 = a.b3;
  p.B = a.B; p.T = a.T; p.d = a.d; p.h1 = a.h1; p.h2 = a.h2; p.C = a.C;
  p.logits = a.logits; p.probs = a.probs; p.idx = a.idx; p.perm = a.perm; p.seg_starts = a.seg_starts;
  p.done_counter = counter;
  const size_t smem = (static_cast<size_t>(a.d) + a.h1 + a.h2 + 12 + 64 + K2H_WARPS + 80) * sizeof(float);
  k2_head_kernel<<<a.B, K2H_THREADS, smem, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "k2: head launch");
  return SAR_OK;
}

}  // namespace sar
