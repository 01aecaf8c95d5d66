// rows_bgmv.cu — row-indexed variant of K1 for decode steps (one token per utterance, so rows of different
// adapters share a tile).  Base projection runs on the tcgen05 K1 kernel (no adapter), then a gathered
// batched matrix-vector kernel adds  (scale·x_m·A_kᵀ)·B_kᵀ  for k = row_adapter[m].
//
// Replaces the per-sample adapter.generate loop of the reference (src/models/adapter_router.py:744-750) for the
// decoder q/v projections at T = 1.
#include <cuda_bf16.h>

#include "sar_internal.h"

namespace sar {

constexpr int ROWS_THREADS = 256;

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// one CTA per row
__global__ void __launch_bounds__(ROWS_THREADS)
rows_bgmv_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ A_stack,
                 const __nv_bfloat16* __restrict__ Bp_stack, const int32_t* __restrict__ row_adapter,
                 __nv_bfloat16* __restrict__ y, int d_in, int d_out, int r, int n_adapters, float scale) {
  __shared__ float u_s[SAR_RPAD];
  const int m = blockIdx.x;
  const int k = row_adapter[m];
  if (k < 0 || k >= n_adapters) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16* xr = x + static_cast<size_t>(m) * d_in;
  const __nv_bfloat16* A = A_stack + static_cast<size_t>(k) * r * d_in;
  // u[j] = scale * <x_m, A_k[j,:]>, rounded to bf16 like the fused kernel's MMA operand
  for (int j = warp; j < r; j += ROWS_THREADS / 32) {
    const uint4* a4 = reinterpret_cast<const uint4*>(A + static_cast<size_t>(j) * d_in);
    const uint4* x4 = reinterpret_cast<const uint4*>(xr);
    float s = 0.f;
    for (int i = lane; i < d_in / 8; i += 32) {
      const uint4 av = __ldg(a4 + i), xv = __ldg(x4 + i);
      s += bf16_lo(av.x) * bf16_lo(xv.x) + bf16_hi(av.x) * bf16_hi(xv.x);
      s += bf16_lo(av.y) * bf16_lo(xv.y) + bf16_hi(av.y) * bf16_hi(xv.y);
      s += bf16_lo(av.z) * bf16_lo(xv.z) + bf16_hi(av.z) * bf16_hi(xv.z);
      s += bf16_lo(av.w) * bf16_lo(xv.w) + bf16_hi(av.w) * bf16_hi(xv.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) u_s[j] = __bfloat162float(__float2bfloat16_rn(s * scale));
  }
  __syncthreads();
  const __nv_bfloat16* Bp = Bp_stack + static_cast<size_t>(k) * d_out * SAR_RPAD;
  __nv_bfloat16* yr = y + static_cast<size_t>(m) * d_out;
  for (int n = threadIdx.x; n < d_out; n += ROWS_THREADS) {
    const uint4* b4 = reinterpret_cast<const uint4*>(Bp + static_cast<size_t>(n) * SAR_RPAD);
    float s = 0.f;
    for (int i = 0; i < r / 8; ++i) {
      const uint4 bv = __ldg(b4 + i);
      const float* uu = u_s + 8 * i;
      s += bf16_lo(bv.x) * uu[0] + bf16_hi(bv.x) * uu[1] + bf16_lo(bv.y) * uu[2] + bf16_hi(bv.y) * uu[3] +
           bf16_lo(bv.z) * uu[4] + bf16_hi(bv.z) * uu[5] + bf16_lo(bv.w) * uu[6] + bf16_hi(bv.w) * uu[7];
    }
    yr[n] = __float2bfloat16_rn(__bfloat162float(yr[n]) + s);
  }
}

// Multi-segment form for the fused decode-step projections: blockIdx.y = LoRA'd output segment, each with its own
// adapter set (A / Bp stacks concatenated set-major) and output tensor; y += oscale * (scale·x_m·A_kᵀ)·B_kᵀ.
struct RowsSegs {
  __nv_bfloat16* y[3];
  int set[3];
  float oscale[3];
};

__global__ void __launch_bounds__(ROWS_THREADS)
rows_bgmv_seg_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ A_cat,
                     const __nv_bfloat16* __restrict__ Bp_cat, const int32_t* __restrict__ row_adapter, RowsSegs segs,
                     int d_in, int d_out, int r, int n_adapters, float scale) {
  __shared__ float u_s[SAR_RPAD];
  const int m = blockIdx.x;
  const int sg = blockIdx.y;
  const int k = row_adapter[m];
  if (k < 0 || k >= n_adapters) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t ak = static_cast<size_t>(segs.set[sg]) * n_adapters + k;
  const __nv_bfloat16* xr = x + static_cast<size_t>(m) * d_in;
  const __nv_bfloat16* A = A_cat + ak * r * d_in;
  for (int j = warp; j < r; j += ROWS_THREADS / 32) {
    const uint4* a4 = reinterpret_cast<const uint4*>(A + static_cast<size_t>(j) * d_in);
    const uint4* x4 = reinterpret_cast<const uint4*>(xr);
    float s = 0.f;
    for (int i = lane; i < d_in / 8; i += 32) {
      const uint4 av = __ldg(a4 + i), xv = __ldg(x4 + i);
      s += bf16_lo(av.x) * bf16_lo(xv.x) + bf16_hi(av.x) * bf16_hi(xv.x);
      s += bf16_lo(av.y) * bf16_lo(xv.y) + bf16_hi(av.y) * bf16_hi(xv.y);
      s += bf16_lo(av.z) * bf16_lo(xv.z) + bf16_hi(av.z) * bf16_hi(xv.z);
      s += bf16_lo(av.w) * bf16_lo(xv.w) + bf16_hi(av.w) * bf16_hi(xv.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) u_s[j] = __bfloat162float(__float2bfloat16_rn(s * scale));
  }
  __syncthreads();
  const __nv_bfloat16* Bp = Bp_cat + ak * d_out * SAR_RPAD;
  __nv_bfloat16* yr = segs.y[sg] + static_cast<size_t>(m) * d_out;
  const float os = segs.oscale[sg];
  for (int n = threadIdx.x; n < d_out; n += ROWS_THREADS) {
    const uint4* b4 = reinterpret_cast<const uint4*>(Bp + static_cast<size_t>(n) * SAR_RPAD);
    float s = 0.f;
    for (int i = 0; i < r / 8; ++i) {
      const uint4 bv = __ldg(b4 + i);
      const float* uu = u_s + 8 * i;
      s += bf16_lo(bv.x) * uu[0] + bf16_hi(bv.x) * uu[1] + bf16_lo(bv.y) * uu[2] + bf16_hi(bv.y) * uu[3] +
           bf16_lo(bv.z) * uu[4] + bf16_hi(bv.z) * uu[5] + bf16_lo(bv.w) * uu[6] + bf16_hi(bv.w) * uu[7];
    }
    yr[n] = __float2bfloat16_rn(__bfloat162float(yr[n]) + s * os);
  }
}

int attn_proj_fwd_rows(const K1Args& a0, const int32_t* row_adapter, int M, cudaStream_t stream) {
  if (M <= 0) return fail(SAR_EINVAL, "attn_proj_rows: M must be positive");
  if (a0.n_seg < 1 || a0.n_seg > 3) return fail(SAR_EINVAL, "attn_proj_rows: n_seg must be 1, 2 or 3");
  K1Args a = a0;                      // base projections of all segments: one dense launch over the flattened rows
  a.B = 1; a.T = M; a.n_adapters = 0; a.utt_adapter = nullptr; a.A_stack = nullptr; a.Bp_stack = nullptr;
  a.x_head_major = 0; a.y_head_major = 0;
  for (int s = 0; s < 3; ++s) a.seg_set[s] = -1;
  int rc = attn_proj_fwd(a, stream);
  if (rc) return rc;
  if (a0.n_adapters <= 0 || !row_adapter || !a0.A_stack || !a0.Bp_stack) return SAR_OK;
  if (a0.r % 8 || a0.r <= 0 || a0.r > SAR_RPAD) return fail(SAR_EINVAL, "attn_proj_rows: r must be a multiple of 8, <= 64");
  RowsSegs segs{};
  int n_lora = 0;
  for (int s = 0; s < a0.n_seg; ++s) {
    if (a0.seg_set[s] < 0) continue;
    if (a0.seg_set[s] >= a0.n_sets) return fail(SAR_EINVAL, "attn_proj_rows: segment refers to a missing LoRA set");
    segs.y[n_lora] = reinterpret_cast<__nv_bfloat16*>(a0.y_seg[s]);
    segs.set[n_lora] = a0.seg_set[s];
    segs.oscale[n_lora] = a0.seg_scale[s];
    ++n_lora;
  }
  if (n_lora == 0) return SAR_OK;
  rows_bgmv_seg_kernel<<<dim3(M, n_lora), ROWS_THREADS, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(a0.x), reinterpret_cast<const __nv_bfloat16*>(a0.A_stack),
      reinterpret_cast<const __nv_bfloat16*>(a0.Bp_stack), row_adapter, segs, a0.d_in, a0.d_out, a0.r, a0.n_adapters,
      a0.scale);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "attn_proj_rows: bgmv launch");
  return SAR_OK;
}

int64_t rows_workspace_bytes(int64_t M, int64_t d, int64_t r) {
  (void)M; (void)d; (void)r;
  return 0;
}

int rows_qv_lora_fwd(const void* x, const void* W, const void* bias, const void* A_stack, const void* Bp_stack,
                     const int32_t* row_adapter, void* y, int M, int d_in, int d_out, int r, int n_adapters,
                     float scale, void* ws, cudaStream_t stream) {
  (void)ws;
  if (M <= 0) return fail(SAR_EINVAL, "rows: M must be positive");
  K1Args a{};
  a.x = x; a.W = W; a.bias = bias; a.y = y;
  a.B = 1; a.T = M; a.d_in = d_in; a.d_out = d_out; a.r = 16; a.n_adapters = 0; a.scale = 0.f;
  int rc = k1_qv_lora_fwd(a, stream);
  if (rc) return rc;
  if (n_adapters <= 0 || !row_adapter || !A_stack || !Bp_stack) return SAR_OK;
  if (r % 8 || r <= 0 || r > SAR_RPAD) return fail(SAR_EINVAL, "rows: r must be a multiple of 8, <= 64");
  rows_bgmv_kernel<<<M, ROWS_THREADS, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(A_stack),
      reinterpret_cast<const __nv_bfloat16*>(Bp_stack), row_adapter, reinterpret_cast<__nv_bfloat16*>(y), d_in, d_out,
      r, n_adapters, scale);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "rows: bgmv launch");
  return SAR_OK;
}

}  // namespace sar
