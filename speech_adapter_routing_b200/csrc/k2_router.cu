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
template <int NV, bool FP32>
__global__ void __launch_bounds__(K2_THREADS)
k2_pool_kernel(const void* __restrict__ h, float* __restrict__ partial, int* __restrict__ done_counter, int T, int d,
               int rows_per_chunk) {
  extern __shared__ float red[];  // [K2_WARPS][d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, chunk = blockIdx.x;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *done_counter = 0;
  const int t0 = chunk * rows_per_chunk;
  const int t1 = min(T, t0 + rows_per_chunk);
  const size_t esz = FP32 ? 4 : 2;
  const uint8_t* base = reinterpret_cast<const uint8_t*>(h) + static_cast<size_t>(b) * T * d * esz;
  const float inv_d = 1.0f / static_cast<float>(d);

  float acc[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[v][i] = 0.f;

  Raw8<FP32> na[NV], nb[NV];
  auto issue = [&](int t) {  // loads of frames t and t + K2_WARPS (the second may not exist)
    const bool has_b = (t + K2_WARPS) < t1;
    const uint8_t* ra = base + static_cast<size_t>(t) * d * esz;
    const uint8_t* rb = base + static_cast<size_t>(has_b ? t + K2_WARPS : t) * d * esz;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int e = (v * 32 + lane) * 8;
      if (e < d) {
        na[v].load(ra, e);
        nb[v].load(rb, e);
      } else {
        na[v].zero();
        nb[v].zero();
      }
    }
  };

  int t = t0 + warp;
  if (t < t1) issue(t);
  for (; t < t1; t += 2 * K2_WARPS) {
    const bool has_b = (t + K2_WARPS) < t1;
    float xa[NV][8], xb[NV][8];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      na[v].to_float(xa[v]);
      nb[v].to_float(xb[v]);
    }
    if (t + 2 * K2_WARPS < t1) issue(t + 2 * K2_WARPS);  // next pair in flight while this pair is reduced
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sa += xa[v][i];
        sb += xb[v][i];
      }
    const float ma = warp_sum(sa) * inv_d, mb = warp_sum(sb) * inv_d;
    float qa = 0.f, qb = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const bool live = (v * 32 + lane) * 8 < d;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float da = live ? xa[v][i] - ma : 0.f, db = live ? xb[v][i] - mb : 0.f;
        xa[v][i] = da;
        xb[v][i] = db;
        qa += da * da;
        qb += db * db;
      }
    }
    const float ra_std = 1.0f / sqrtf(warp_sum(qa) * inv_d + K2_EPS);
    const float rb_std = has_b ? 1.0f / sqrtf(warp_sum(qb) * inv_d + K2_EPS) : 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[v][i] += xa[v][i] * ra_std + xb[v][i] * rb_std;
  }

  // fixed-order cross-warp reduction
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int e = (v * 32 + lane) * 8;
    if (e < d) {
#pragma unroll
      for (int i = 0; i < 8; ++i) red[warp * d + e + i] = acc[v][i];
    }
  }
  __syncthreads();
  float* out = partial + (static_cast<size_t>(b) * gridDim.x + chunk) * d;
  for (int j = threadIdx.x; j < d; j += K2_THREADS) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < K2_WARPS; ++w) s += red[w * d + j];
    out[j] = s;
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
  if (is_last && threadIdx.x == 0) {
    // stable counting sort of utterances by class: B is a batch size (<= a few thousand)
    __threadfence();
    int* counts = reinterpret_cast<int*>(scratch + K2_WARPS);
    for (int c = 0; c <= p.C; ++c) counts[c] = 0;
    const volatile int32_t* vidx = p.idx;
    for (int i = 0; i < p.B; ++i) counts[vidx[i] + 1]++;
    for (int c = 0; c < p.C; ++c) counts[c + 1] += counts[c];
    for (int c = 0; c <= p.C; ++c) p.seg_starts[c] = counts[c];
    for (int i = 0; i < p.B; ++i) p.perm[counts[vidx[i]]++] = i;
  }
}

int64_t k2_workspace_bytes(int64_t B, int64_t T, int64_t d) {
  if (B <= 0 || T <= 0 || d <= 0) return SAR_EINVAL;
  const int64_t max_chunks = (T + 15) / 16;  // rows_per_chunk >= 16
  return 256 + B * max_chunks * d * 4;
}

static int k2_rows_per_chunk(int B, int T, int num_sms) {
  // Rows per CTA: a multiple of 16 (8 warps x 2 frames per iteration) in [16, 128].  Small batches get small chunks
  // (more CTAs); large batches pick the size whose CTA count wastes the least of the last wave (2 CTAs per SM).
  const int64_t slots = static_cast<int64_t>(num_sms) * 2;
  int best = 16;
  double best_cost = 1e30;
  for (int rows = 16; rows <= 128; rows += 16) {
    const int64_t ctas = static_cast<int64_t>(B) * ((T + rows - 1) / rows);
    const int64_t waves = (ctas + slots - 1) / slots;
    // time ~ waves * (rows + fixed per-CTA overhead of ~12 rows' worth of latency)
    const double cost = static_cast<double>(waves) * (rows + 12);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = rows;
    }
  }
  return best;
}

template <bool FP32>
static int k2_launch_pool(const K2Args& a, float* partial, int* counter, int chunks, int rpc, cudaStream_t stream) {
  const int nv = (a.d + 255) / 256;
  const dim3 grid(chunks, a.B);
  const size_t smem = static_cast<size_t>(K2_WARPS) * a.d * sizeof(float);
#define SAR_K2_CASE(N)                                                                                          \
  case N:                                                                                                       \
    if (smem > 48 * 1024)                                                                                       \
      cudaFuncSetAttribute(k2_pool_kernel<N, FP32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    k2_pool_kernel<N, FP32><<<grid, K2_THREADS, smem, stream>>>(a.h, partial, counter, a.T, a.d, rpc);          \
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
