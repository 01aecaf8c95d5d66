// decode_attn.cu — self-attention of ONE decode step over the static KV cache (head dim 64).
//
//   cache_k[b,h,pos,:] = k_new[b,h,:];  cache_v[b,h,pos,:] = v_new[b,h,:]
//   out[b,h,:] = softmax_{t <= pos}( q[b,h,:] · cache_k[b,h,t,:] ) · cache_v[b,h,t,:]          (q is pre-scaled)
//
// Replaces, per decoder layer and token, the DynamicCache.update + causal mask + SDPA of WhisperAttention.forward at
// tgt_len = 1 ($HF/models/whisper/modeling_whisper.py:326-350) — in the captured decode graph that was two index_copy
// kernels, three mask-building kernels and a library SDPA launch.  The position is read from DEVICE memory, so the
// same captured launch serves every token of the loop.  Latency-bound (a (b, h) pair touches <= 2·448·128 B of cache):
// one CTA per (b, h), scores in shared memory, fp32 softmax, coalesced V reads.
#include <cuda_bf16.h>

#include "sar_internal.h"

namespace sar {

constexpr int DA_THREADS = 128;
constexpr int DA_HD = 64;

__global__ void __launch_bounds__(DA_THREADS)
decode_self_attn_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k_new,
                        const __nv_bfloat16* __restrict__ v_new, __nv_bfloat16* __restrict__ cache_k,
                        __nv_bfloat16* __restrict__ cache_v, const long long* __restrict__ pos_ptr,
                        __nv_bfloat16* __restrict__ out, int t_max) {
  extern __shared__ float da_smem[];   // q[64] | scores[t_max] | red[4] | part[2][64]
  float* qs = da_smem;
  float* sc = qs + DA_HD;
  float* red = sc + t_max;
  float* part = red + 4;
  const int bh = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  long long pos = *pos_ptr;
  if (pos < 0) pos = 0;
  if (pos >= t_max) pos = t_max - 1;
  const int n = static_cast<int>(pos) + 1;   // keys 0..pos
  __nv_bfloat16* ck = cache_k + static_cast<size_t>(bh) * t_max * DA_HD;
  __nv_bfloat16* cv = cache_v + static_cast<size_t>(bh) * t_max * DA_HD;
  if (tid < DA_HD) {
    qs[tid] = __bfloat162float(q[static_cast<size_t>(bh) * DA_HD + tid]);
    ck[static_cast<size_t>(pos) * DA_HD + tid] = k_new[static_cast<size_t>(bh) * DA_HD + tid];
    cv[static_cast<size_t>(pos) * DA_HD + tid] = v_new[static_cast<size_t>(bh) * DA_HD + tid];
  }
  __syncthreads();   // also makes this CTA's cache writes visible to its own reads below

  // scores: thread t handles keys t, t + 128, ...  (one 128-byte row = 8 x 16-byte loads)
  float mx = -INFINITY;
  for (int t = tid; t < n; t += DA_THREADS) {
    const uint4* kr = reinterpret_cast<const uint4*>(ck + static_cast<size_t>(t) * DA_HD);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 w = kr[j];
      const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        s = fmaf(__uint_as_float(ww[i] << 16), qs[8 * j + 2 * i], s);
        s = fmaf(__uint_as_float(ww[i] & 0xFFFF0000u), qs[8 * j + 2 * i + 1], s);
      }
    }
    sc[t] = s;
    mx = fmaxf(mx, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  float sum = 0.f;
  for (int t = tid; t < n; t += DA_THREADS) {
    const float e = __expf(sc[t] - mx);
    sc[t] = e;
    sum += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncthreads();   // everyone has read red[] (max) before it is reused; sc[] is complete
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  const float inv = 1.0f / (red[0] + red[1] + red[2] + red[3]);

  // out[d] = sum_t p[t] * V[t][d]: two halves of the keys on two groups of 64 threads, coalesced along d
  const int d = tid & (DA_HD - 1), half = tid >> 6;
  float acc = 0.f;
  for (int t = half; t < n; t += 2) acc = fmaf(sc[t], __bfloat162float(cv[static_cast<size_t>(t) * DA_HD + d]), acc);
  part[half * DA_HD + d] = acc;
  __syncthreads();
  if (tid < DA_HD) out[static_cast<size_t>(bh) * DA_HD + tid] = __float2bfloat16_rn((part[tid] + part[DA_HD + tid]) * inv);
}

int decode_self_attn(const void* q, const void* k_new, const void* v_new, void* cache_k, void* cache_v,
                     const long long* pos, void* out, int B, int H, int head_dim, int t_max, cudaStream_t stream) {
  if (!q || !k_new || !v_new || !cache_k || !cache_v || !pos || !out) return fail(SAR_EINVAL, "decode_self_attn: null pointer");
  if (head_dim != DA_HD) return fail(SAR_EINVAL, "decode_self_attn: head dim must be 64");
  if (B <= 0 || H <= 0 || t_max <= 0 || t_max > 8192) return fail(SAR_EINVAL, "decode_self_attn: bad sizes");
  if ((reinterpret_cast<uintptr_t>(cache_k) | reinterpret_cast<uintptr_t>(cache_v)) & 15)
    return fail(SAR_EINVAL, "decode_self_attn: caches must be 16-byte aligned");
  const size_t smem = (DA_HD + t_max + 4 + 2 * DA_HD) * sizeof(float);
  decode_self_attn_kernel<<<B * H, DA_THREADS, smem, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(q), reinterpret_cast<const __nv_bfloat16*>(k_new),
      reinterpret_cast<const __nv_bfloat16*>(v_new), reinterpret_cast<__nv_bfloat16*>(cache_k),
      reinterpret_cast<__nv_bfloat16*>(cache_v), pos, reinterpret_cast<__nv_bfloat16*>(out), t_max);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "decode_self_attn: launch");
  return SAR_OK;
}

}  // namespace sar
