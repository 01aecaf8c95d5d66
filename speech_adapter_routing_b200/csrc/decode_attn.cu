// decode_attn.cu — self-attention of ONE decode step over the static KV cache (head dim 64).
//
//   cache_k[b,h,pos,:] = k_new[b,h,:];  cache_v[b,h,pos,:] = v_new[b,h,:]
//   out[b,h,:] = softmax_{t <= pos}( q[b,h,:] · cache_k[b,h,t,:] ) · cache_v[b,h,t,:]          (q is pre-scaled)
//
// Replaces, per decoder layer and token, the DynamicCache.update + causal mask + SDPA of WhisperAttention.forward at
// tgt_len = 1 ($HF/models/whisper/modeling_whisper.py:326-350) — in the captured decode graph that was two index_copy
// kernels, three mask-building kernels and a library SDPA launch.  The position is read from DEVICE memory, so the
// same captured launch serves every token of the loop.  Latency-bound (a (b, h) pair touches <= 2·448·128 B of cache):
// one CTA per (b, h), scores in shared memory, fp32 softmax.  The same kernel without the cache write and over all Tk
// keys is the decoder's cross-attention at one token per utterance (sar_decode_cross_attn): HBM-bound, it streams the
// 2 x Tk x 128 B of encoder K / V of every (b, h) once per token.  K and then V arrive in 64-key chunks (8 KB) through a
// 3-stage shared-memory ring of 1-D bulk async copies (cp.async.bulk + mbarrier complete_tx): 24 KB in flight per CTA
// with no register cost, all B*H CTAs resident at once (6 per SM) — register-staged 16-byte loads left a CTA
// latency-bound at 3.2-3.9 TB/s (93 / 75 us for B = 64, h = 12, Tk = 1500; cuDNN 48 us).
#include <cuda_bf16.h>

#include "sar_internal.h"
#include "sar_ptx.cuh"

namespace sar {

constexpr int DA_THREADS = 128;
constexpr int DA_HD = 64;
constexpr int DA_CHUNK_KEYS = 64;
constexpr int DA_CHUNK_BYTES = DA_CHUNK_KEYS * DA_HD * 2;   // 8 KB
constexpr int DA_STAGES = 3;
constexpr int DA_GROUPS = DA_THREADS / 8;                   // 8 threads x 16 B cover one 128-byte row

__device__ __forceinline__ void da_bulk_load(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// SELF: append k_new / v_new at *pos_ptr and attend to keys 0..pos.  !SELF: attend to all t_max keys of a read-only K / V
// (the decoder's cross-attention over the encoder states, one token per utterance).
template <bool SELF>
__global__ void __launch_bounds__(DA_THREADS)
decode_attn_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k_new,
                   const __nv_bfloat16* __restrict__ v_new, __nv_bfloat16* __restrict__ cache_k,
                   __nv_bfloat16* __restrict__ cache_v, const long long* __restrict__ pos_ptr,
                   __nv_bfloat16* __restrict__ out, int t_max) {
  // ring[DA_STAGES][8 KB] | full[DA_STAGES] | red[4] | part[DA_GROUPS][64] | scores[t_max]
  extern __shared__ __align__(128) uint8_t da_smem[];
  uint8_t* ring = da_smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + DA_STAGES * DA_CHUNK_BYTES);
  float* red = reinterpret_cast<float*>(full + DA_STAGES + 1);
  float* part = red + 4;
  float* sc = part + DA_GROUPS * DA_HD;
  const int bh = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int n = t_max;
  __nv_bfloat16* ck = cache_k + static_cast<size_t>(bh) * t_max * DA_HD;
  __nv_bfloat16* cv = cache_v + static_cast<size_t>(bh) * t_max * DA_HD;
  if (SELF) {
    long long pos = *pos_ptr;
    if (pos < 0) pos = 0;
    if (pos >= t_max) pos = t_max - 1;
    n = static_cast<int>(pos) + 1;   // keys 0..pos
    if (tid < DA_HD) {
      ck[static_cast<size_t>(pos) * DA_HD + tid] = k_new[static_cast<size_t>(bh) * DA_HD + tid];
      cv[static_cast<size_t>(pos) * DA_HD + tid] = v_new[static_cast<size_t>(bh) * DA_HD + tid];
      asm volatile("fence.proxy.async;" ::: "memory");   // the bulk copies below (async proxy) must see these rows
    }
  }
  if (tid == 0) {
    for (int s = 0; s < DA_STAGES; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  // this thread's 8 query dims (8 threads x 8 dims = one key row per quarter-warp... per 8 lanes)
  const int sub = lane & 7;
  float qr[8];
  {
    const uint4 w = *reinterpret_cast<const uint4*>(q + static_cast<size_t>(bh) * DA_HD + sub * 8);
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      qr[2 * i] = __uint_as_float(ww[i] << 16);
      qr[2 * i + 1] = __uint_as_float(ww[i] & 0xFFFF0000u);
    }
  }
  __syncthreads();

  const int n_chunks = (n + DA_CHUNK_KEYS - 1) / DA_CHUNK_KEYS;
  const int total = 2 * n_chunks;   // K chunks, then V chunks, through the same ring
  auto issue = [&](int c) {         // thread 0 only
    const int cc = c < n_chunks ? c : c - n_chunks;
    const __nv_bfloat16* src = (c < n_chunks ? ck : cv) + static_cast<size_t>(cc) * DA_CHUNK_KEYS * DA_HD;
    const uint32_t bytes = static_cast<uint32_t>(min(DA_CHUNK_KEYS, n - cc * DA_CHUNK_KEYS)) * DA_HD * 2;
    const int st = c % DA_STAGES;
    mbar_arrive_expect_tx(&full[st], bytes);
    da_bulk_load(ring + st * DA_CHUNK_BYTES, src, bytes, &full[st]);
  };
  if (tid == 0)
    for (int c = 0; c < min(total, DA_STAGES); ++c) issue(c);

  // ---- scores: 8 lanes per key (16 bytes each, conflict-free), 16 keys per pass, 4 passes per chunk
  float mx = -INFINITY;
  for (int c = 0; c < n_chunks; ++c) {
    const int st = c % DA_STAGES;
    mbar_wait(&full[st], (c / DA_STAGES) & 1);
    const uint8_t* chunk = ring + st * DA_CHUNK_BYTES;
#pragma unroll
    for (int pass = 0; pass < DA_CHUNK_KEYS / DA_GROUPS; ++pass) {
      const int key = pass * DA_GROUPS + (tid >> 3);
      const uint4 w = *reinterpret_cast<const uint4*>(chunk + key * (DA_HD * 2) + sub * 16);
      const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        s = fmaf(__uint_as_float(ww[i] << 16), qr[2 * i], s);
        s = fmaf(__uint_as_float(ww[i] & 0xFFFF0000u), qr[2 * i + 1], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      const int t = c * DA_CHUNK_KEYS + key;
      if (t < n) {                      // rows past n in the last chunk are stale shared memory
        if (sub == 0) sc[t] = s;
        mx = fmaxf(mx, s);
      }
    }
    __syncthreads();                    // every warp is done with this stage
    if (tid == 0 && c + DA_STAGES < total) issue(c + DA_STAGES);
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

  // ---- out[d] = sum_t p[t] * V[t][d]: 8 lanes x 8 dims cover a V row, DA_GROUPS rows per pass
  const int g = tid >> 3;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int c = n_chunks; c < total; ++c) {
    const int st = c % DA_STAGES;
    mbar_wait(&full[st], (c / DA_STAGES) & 1);
    const uint8_t* chunk = ring + st * DA_CHUNK_BYTES;
#pragma unroll
    for (int pass = 0; pass < DA_CHUNK_KEYS / DA_GROUPS; ++pass) {
      const int key = pass * DA_GROUPS + g;
      const int t = (c - n_chunks) * DA_CHUNK_KEYS + key;
      if (t < n) {
        const uint4 w = *reinterpret_cast<const uint4*>(chunk + key * (DA_HD * 2) + sub * 16);
        const float pt = sc[t];
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[2 * i] = fmaf(pt, __uint_as_float(ww[i] << 16), acc[2 * i]);
          acc[2 * i + 1] = fmaf(pt, __uint_as_float(ww[i] & 0xFFFF0000u), acc[2 * i + 1]);
        }
      }
    }
    __syncthreads();
    if (tid == 0 && c + DA_STAGES < total) issue(c + DA_STAGES);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) part[g * DA_HD + sub * 8 + i] = acc[i];
  __syncthreads();
  if (tid < DA_HD) {
    float o = 0.f;
#pragma unroll
    for (int gg = 0; gg < DA_GROUPS; ++gg) o += part[gg * DA_HD + tid];   // fixed order: deterministic
    out[static_cast<size_t>(bh) * DA_HD + tid] = __float2bfloat16_rn(o * inv);
  }
}

static size_t da_smem_bytes(int t_max) {
  return DA_STAGES * DA_CHUNK_BYTES + (DA_STAGES + 1) * 8 + (4 + DA_GROUPS * DA_HD + t_max) * sizeof(float);
}

template <bool SELF>
static int da_launch(const void* q, const void* k_new, const void* v_new, void* cache_k, void* cache_v,
                     const long long* pos, void* out, int BH, int t_max, cudaStream_t stream, const char* who) {
  const size_t smem = da_smem_bytes(t_max);
  static thread_local size_t smem_set[64] = {};
  const DeviceInfo& dev = device_info();
  size_t& cur = smem_set[dev.device & 63];
  if (smem > 48 * 1024 && cur < smem) {
    cudaError_t e = cudaFuncSetAttribute(decode_attn_kernel<SELF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return fail_cuda(e, who);
    cur = smem;
  }
  decode_attn_kernel<SELF><<<BH, DA_THREADS, smem, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(q), reinterpret_cast<const __nv_bfloat16*>(k_new),
      reinterpret_cast<const __nv_bfloat16*>(v_new), reinterpret_cast<__nv_bfloat16*>(cache_k),
      reinterpret_cast<__nv_bfloat16*>(cache_v), pos, reinterpret_cast<__nv_bfloat16*>(out), t_max);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, who);
  return SAR_OK;
}

int decode_self_attn(const void* q, const void* k_new, const void* v_new, void* cache_k, void* cache_v,
                     const long long* pos, void* out, int B, int H, int head_dim, int t_max, cudaStream_t stream) {
  if (!q || !k_new || !v_new || !cache_k || !cache_v || !pos || !out) return fail(SAR_EINVAL, "decode_self_attn: null pointer");
  if (head_dim != DA_HD) return fail(SAR_EINVAL, "decode_self_attn: head dim must be 64");
  if (B <= 0 || H <= 0 || t_max <= 0 || t_max > 8192) return fail(SAR_EINVAL, "decode_self_attn: bad sizes");
  if ((reinterpret_cast<uintptr_t>(cache_k) | reinterpret_cast<uintptr_t>(cache_v) | reinterpret_cast<uintptr_t>(q)) & 15)
    return fail(SAR_EINVAL, "decode_self_attn: q and the caches must be 16-byte aligned");
  return da_launch<true>(q, k_new, v_new, cache_k, cache_v, pos, out, B * H, t_max, stream, "decode_self_attn: launch");
}

int decode_cross_attn(const void* q, const void* k, const void* v, void* out, int B, int H, int head_dim, int Tk,
                      cudaStream_t stream) {
  if (!q || !k || !v || !out) return fail(SAR_EINVAL, "decode_cross_attn: null pointer");
  if (head_dim != DA_HD) return fail(SAR_EINVAL, "decode_cross_attn: head dim must be 64");
  if (B <= 0 || H <= 0 || Tk <= 0 || Tk > 8192) return fail(SAR_EINVAL, "decode_cross_attn: bad sizes");
  if ((reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(q)) & 15)
    return fail(SAR_EINVAL, "decode_cross_attn: q, k and v must be 16-byte aligned");
  // K / V are only read (SELF = false never writes through the cache pointers)
  return da_launch<false>(q, nullptr, nullptr, const_cast<void*>(k), const_cast<void*>(v), nullptr, out, B * H, Tk,
                          stream, "decode_cross_attn: launch");
}

}  // namespace sar
