// logmel.cu — Whisper's log-mel front-end on the device: waveform [B, 480000] fp32 -> input_features [B, n_mels, 3000].
//
// Replaces the per-example CPU call `processor.feature_extractor(audio_array, ...)` of the reference's data path
// (src/data/dataset.py:124-128 -> $HF/models/whisper/feature_extraction_whisper.py:105-135): reflect-pad n_fft/2,
// 400-sample periodic-Hann frames every 160 samples (the 3001st frame is dropped), |DFT|^2 over 201 bins, Slaney mel
// filterbank, log10 with a 1e-10 floor, clamp to (clip maximum - 8), (x + 4) / 4.
//
// fp32 CUDA-core arithmetic on purpose: the clamp keeps bins down to 1e-8 of the clip's peak power, i.e. an amplitude
// resolution of 1e-4 of full scale — bf16 / tf32 tensor-core operands (2^-8 / 2^-11) would put their rounding noise
// above that floor.  Kernel 1 owns 32 frames of one clip: the windowed frames sit transposed in shared memory (one
// 16-byte load feeds four frames), thread k accumulates bin k for all 32 frames in registers (64 accumulators, twiddles
// from a 400-entry shared table walked with stride k), then 80 / 128 threads contract the 32 x 201 power block with the
// filterbank.  123 GFLOP per 64-clip batch; off the routed forward's timed path.
// Kernel 2 applies the per-clip clamp (needs the clip maximum, collected with an ordered-integer atomicMax) and casts.
#include <cuda_bf16.h>

#include "sar_internal.h"

namespace sar {

constexpr int LM_NFFT = 400;
constexpr int LM_HOP = 160;
constexpr int LM_BINS = LM_NFFT / 2 + 1;   // 201
constexpr int LM_FR = 32;                  // frames per CTA
constexpr int LM_THREADS = 256;

// order-preserving float <-> int map for atomicMax on signed ints
__device__ __forceinline__ int lm_float_to_ordered(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float lm_ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void lm_init_kernel(int* clip_max, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) clip_max[i] = lm_float_to_ordered(-INFINITY);
}

__global__ void __launch_bounds__(LM_THREADS)
lm_spec_kernel(const float* __restrict__ wave, const float* __restrict__ window, const float* __restrict__ cos_t,
               const float* __restrict__ sin_t, const float* __restrict__ filters /* [201, n_mels] */,
               float* __restrict__ raw /* [B, n_mels, n_frames] */, int* __restrict__ clip_max, int n_samples,
               int n_frames, int n_mels) {
  extern __shared__ __align__(16) float lm_smem[];
  float* at = lm_smem;                          // [400][32]  windowed frames, transposed
  float* pw = at + LM_NFFT * LM_FR;             // [32][201]  power spectrum
  float* ct = pw + LM_FR * LM_BINS;             // [400]
  float* st = ct + LM_NFFT;                     // [400]
  float* red = st + LM_NFFT;                    // [8]
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * LM_FR;
  const int tid = threadIdx.x;
  const float* x = wave + static_cast<size_t>(b) * n_samples;

  for (int i = tid; i < LM_NFFT; i += LM_THREADS) {
    ct[i] = cos_t[i];
    st[i] = sin_t[i];
  }
  // frame f, tap n reads sample (f0 + f) * 160 + n - 200, reflected at both ends (np.pad mode="reflect")
  for (int i = tid; i < LM_NFFT * LM_FR; i += LM_THREADS) {
    const int n = i / LM_FR, f = i - n * LM_FR;
    float v = 0.f;
    if (f0 + f < n_frames) {
      int s = (f0 + f) * LM_HOP + n - LM_NFFT / 2;
      if (s < 0) s = -s;
      if (s >= n_samples) s = 2 * (n_samples - 1) - s;
      v = x[s] * window[n];
    }
    at[i] = v;
  }
  __syncthreads();

  if (tid < LM_BINS) {
    const int k = tid;
    float re[LM_FR], im[LM_FR];
#pragma unroll
    for (int f = 0; f < LM_FR; ++f) re[f] = im[f] = 0.f;
    int idx = 0;                                 // (k * n) mod 400
    for (int n = 0; n < LM_NFFT; ++n) {
      const float c = ct[idx], s = st[idx];
      const float4* row = reinterpret_cast<const float4*>(at + n * LM_FR);
#pragma unroll
      for (int g = 0; g < LM_FR / 4; ++g) {
        const float4 v = row[g];                 // same address for every thread: broadcast
        re[4 * g + 0] = fmaf(v.x, c, re[4 * g + 0]); im[4 * g + 0] = fmaf(v.x, s, im[4 * g + 0]);
        re[4 * g + 1] = fmaf(v.y, c, re[4 * g + 1]); im[4 * g + 1] = fmaf(v.y, s, im[4 * g + 1]);
        re[4 * g + 2] = fmaf(v.z, c, re[4 * g + 2]); im[4 * g + 2] = fmaf(v.z, s, im[4 * g + 2]);
        re[4 * g + 3] = fmaf(v.w, c, re[4 * g + 3]); im[4 * g + 3] = fmaf(v.w, s, im[4 * g + 3]);
      }
      idx += k;
      if (idx >= LM_NFFT) idx -= LM_NFFT;
    }
#pragma unroll
    for (int f = 0; f < LM_FR; ++f) pw[f * LM_BINS + k] = re[f] * re[f] + im[f] * im[f];
  }
  __syncthreads();

  // mel: thread m contracts its filter column with the 32 power rows; log10 with the 1e-10 floor
  float mx = -INFINITY;
  if (tid < n_mels) {
    const int m = tid;
    float acc[LM_FR];
#pragma unroll
    for (int f = 0; f < LM_FR; ++f) acc[f] = 0.f;
    for (int k = 0; k < LM_BINS; ++k) {
      const float w = filters[k * n_mels + m];   // coalesced across m
      if (w != 0.f) {
#pragma unroll
        for (int f = 0; f < LM_FR; ++f) acc[f] = fmaf(w, pw[f * LM_BINS + k], acc[f]);
      }
    }
    float* dst = raw + (static_cast<size_t>(b) * n_mels + m) * n_frames + f0;
#pragma unroll
    for (int f = 0; f < LM_FR; ++f) {
      if (f0 + f < n_frames) {
        const float v = log10f(fmaxf(acc[f], 1e-10f));
        dst[f] = v;
        mx = fmaxf(mx, v);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  if (tid == 0) {
    float m = red[0];
    for (int w = 1; w < LM_THREADS / 32; ++w) m = fmaxf(m, red[w]);
    atomicMax(clip_max + b, lm_float_to_ordered(m));
  }
}

template <bool BF16>
__global__ void lm_finish_kernel(const float* __restrict__ raw, const int* __restrict__ clip_max, void* __restrict__ out,
                                 long long per_clip, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float floor_v = lm_ordered_to_float(clip_max[i / per_clip]) - 8.0f;
  const float v = (fmaxf(raw[i], floor_v) + 4.0f) * 0.25f;
  if (BF16)
    reinterpret_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16_rn(v);
  else
    reinterpret_cast<float*>(out)[i] = v;
}

int logmel_fwd(const float* wave, const float* window, const float* cos_t, const float* sin_t, const float* filters,
               float* raw_ws, int* clip_max_ws, void* out, int B, int n_samples, int n_mels, int out_bf16,
               cudaStream_t stream) {
  if (!wave || !window || !cos_t || !sin_t || !filters || !raw_ws || !clip_max_ws || !out)
    return fail(SAR_EINVAL, "logmel_fwd: null pointer");
  if (B <= 0 || n_samples < LM_NFFT || n_samples % LM_HOP) return fail(SAR_EINVAL, "logmel_fwd: n_samples must be a positive multiple of 160");
  if (n_mels <= 0 || n_mels > LM_THREADS) return fail(SAR_EINVAL, "logmel_fwd: n_mels must be in 1..256");
  const int n_frames = n_samples / LM_HOP;
  const size_t smem = (LM_NFFT * LM_FR + LM_FR * LM_BINS + 2 * LM_NFFT + 8) * sizeof(float);
  static thread_local bool attr_set[64] = {};
  const DeviceInfo& dev = device_info();
  if (!attr_set[dev.device & 63]) {
    cudaError_t e = cudaFuncSetAttribute(lm_spec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return fail_cuda(e, "logmel_fwd: cudaFuncSetAttribute");
    attr_set[dev.device & 63] = true;
  }
  lm_init_kernel<<<(B + 255) / 256, 256, 0, stream>>>(clip_max_ws, B);
  const dim3 grid((n_frames + LM_FR - 1) / LM_FR, B);
  lm_spec_kernel<<<grid, LM_THREADS, smem, stream>>>(wave, window, cos_t, sin_t, filters, raw_ws, clip_max_ws, n_samples,
                                                     n_frames, n_mels);
  const long long per_clip = static_cast<long long>(n_mels) * n_frames, total = per_clip * B;
  const int blocks = static_cast<int>((total + 255) / 256);
  if (out_bf16)
    lm_finish_kernel<true><<<blocks, 256, 0, stream>>>(raw_ws, clip_max_ws, out, per_clip, total);
  else
    lm_finish_kernel<false><<<blocks, 256, 0, stream>>>(raw_ws, clip_max_ws, out, per_clip, total);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "logmel_fwd: launch");
  return SAR_OK;
}

}  // namespace sar
