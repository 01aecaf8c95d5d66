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

__device__ __forceinline__ float block_sum(float v, float* scratch) {
  // scratch: K2_WARPS floats.  All threads receive the same fixed-order sum.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < K2_WARPS; ++w) s += scratch[w];
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
    for (int o = 2 * warp; o < n_out; o += 2 * K2_WARPS) {
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
    for (int o = warp; o < n_out; o += K2_WARPS) {
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
  for (int j = threadIdx.x; j < n; j += K2_THREADS) s += v[j];
  const float mean = block_sum(s, scratch) / static_cast<float>(n);
  float q = 0.f;
  for (int j = threadIdx.x; j < n; j += K2_THREADS) {
    const float dlt = v[j] - mean;
    q += dlt * dlt;
  }
  const float rstd = 1.0f / sqrtf(block_sum(q, scratch) / static_cast<float>(n) + K2_EPS);
  for (int j = threadIdx.x; j < n; j += K2_THREADS) {
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

__global__ void __launch_bounds__(K2_THREADS) k2_head_kernel(const K2HeadParams p) {
  extern __shared__ __align__(16) float sm[];  // pooled[d] | a1[h1] | a2[h2] | lg[64] | scratch[K2_WARPS] | counts
  float* pooled = sm;
  float* a1 = pooled + ((p.d + 3) & ~3);
  float* a2 = a1 + ((p.h1 + 3) & ~3);
  float* lg = a2 + ((p.h2 + 3) & ~3);
  float* scratch = lg + 64;
  __shared__ int is_last;
  const int b = blockIdx.x;

  const float inv_T = 1.0f / static_cast<float>(p.T);
  for (int j = threadIdx.x; j < p.d; j += K2_THREADS) {
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
    __shared__ int warp_tot[K2_WARPS];
    const volatile int32_t* vidx = p.idx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int running = 0;
    for (int c = 0; c < p.C; ++c) {
      if (threadIdx.x == 0) p.seg_starts[c] = running;
      for (int i0 = 0; i0 < p.B; i0 += K2_THREADS) {
        const int i = i0 + threadIdx.x;
        const bool flag = i < p.B && vidx[i] == c;
        const unsigned ballot = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) warp_tot[warp] = __popc(ballot);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < K2_WARPS; ++w) {
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
  const int rpc = k2_rows_per_chunk(a.B, a.T, dev.num_sms);
  const int chunks = (a.T + rpc - 1) / rpc;
  int* counter = reinterpret_cast<int*>(a.ws);
  float* partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a.ws) + 256);
  int rc = a.h_is_fp32 ? k2_launch_pool<true>(a, partial, counter, chunks, rpc, stream)
                       : k2_launch_pool<false>(a, partial, counter, chunks, rpc, stream);
  if (rc) return rc;

  K2HeadParams p{};
  p.partial = partial; p.chunks = chunks;
  p.ln_w = a.ln_w; p.ln_b = a.ln_b; p.W1 = a.W1; p.b1 = a.b1; p.g1 = a.g1; p.be1 = a.be1;
  p.W2 = a.W2; p.b2 = a.b2; p.g2 = a.g2; p.be2 = a.be2; p.W3 = a.W3; p.b3 = a.b3;
  p.B = a.B; p.T = a.T; p.d = a.d; p.h1 = a.h1; p.h2 = a.h2; p.C = a.C;
  p.logits = a.logits; p.probs = a.probs; p.idx = a.idx; p.perm = a.perm; p.seg_starts = a.seg_starts;
  p.done_counter = counter;
  const size_t smem = (static_cast<size_t>(a.d) + a.h1 + a.h2 + 12 + 64 + K2_WARPS + 80) * sizeof(float);
  k2_head_kernel<<<a.B, K2_THREADS, smem, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "k2: head launch");
  return SAR_OK;
}

}  // namespace sar
