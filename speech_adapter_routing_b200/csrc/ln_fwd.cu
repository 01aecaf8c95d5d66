// ln_fwd.cu — LayerNorm over the last dimension for the Whisper blocks (bf16 in/out, fp32 statistics).
//
// HBM-bound: algorithmic traffic = one read + one write of x (4·M·d bytes).  One warp owns one row: the row lives in
// registers (16-byte vector loads, lane-strided so every warp request is a run of full 128-byte lines), mean and
// variance are two register passes with warp-shuffle reductions (same two-pass arithmetic as torch's fp32 LayerNorm
// on bf16 input), gamma/beta are read once per warp per row from L1/L2.
#include "sar_internal.h"
#include "sar_ptx.cuh"

namespace sar {

constexpr int LN_WARPS = 8;
constexpr int LN_MAXV = 8;   // 16-byte vectors per lane: d <= 8 * 32 * 8 = 2048

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NV>   // NV = ceil(d / 256) vectors per lane
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ gamma,
                                                               const uint4* __restrict__ beta, uint4* __restrict__ y,
                                                               float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                               long long M, int nvec, float inv_d, float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * LN_WARPS + (threadIdx.x >> 5);
  if (row >= M) return;
  const uint4* xr = x + row * nvec;
  float v[NV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    if (c < nvec) q = __ldcs(xr + c);   // streamed: every byte of x is read once
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[i][2 * j] = __uint_as_float(w[j] << 16);
      v[i][2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
      s += v[i][2 * j] + v[i][2 * j + 1];
    }
  }
  const float mean = warp_sum(s) * inv_d;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (lane + 32 * i < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = v[i][j] - mean;
        ss += t * t;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) * inv_d + eps);
  if (mean_out != nullptr && lane == 0) {   // training: the statistics ATen's LayerNorm backward consumes
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
  uint4* yr = y + row * nvec;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      const uint4 g = __ldg(gamma + c);
      const uint4 bt = __ldg(beta + c);
      const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
      const uint32_t bw[4] = {bt.x, bt.y, bt.z, bt.w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a0 = (v[i][2 * j] - mean) * rstd * __uint_as_float(gw[j] << 16) + __uint_as_float(bw[j] << 16);
        const float a1 = (v[i][2 * j + 1] - mean) * rstd * __uint_as_float(gw[j] & 0xFFFF0000u) +
                         __uint_as_float(bw[j] & 0xFFFF0000u);
        o[j] = pack_bf16x2(a0, a1);
      }
      yr[c] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

int layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean_out, float* rstd_out, int64_t M,
                  int d, float eps, cudaStream_t stream) {
  if (!x || !gamma || !beta || !y) return fail(SAR_EINVAL, "layernorm: null pointer");
  if ((mean_out == nullptr) != (rstd_out == nullptr)) return fail(SAR_EINVAL, "layernorm: mean and rstd come together");
  if (M <= 0) return fail(SAR_EINVAL, "layernorm: M must be positive");
  if (d <= 0 || d % 8 || d > LN_MAXV * 256) return fail(SAR_EINVAL, "layernorm: d must be a multiple of 8 and <= 2048");
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gamma) |
       reinterpret_cast<uintptr_t>(beta)) & 15)
    return fail(SAR_EINVAL, "layernorm: pointers must be 16-byte aligned");
  const int nvec = d / 8;
  const int nv = (nvec + 31) / 32;
  const long long blocks = (M + LN_WARPS - 1) / LN_WARPS;
  if (blocks > 0x7fffffffLL) return fail(SAR_EINVAL, "layernorm: too many rows");
  const dim3 grid(static_cast<unsigned>(blocks));
  const uint4* xp = static_cast<const uint4*>(x);
  const uint4* gp = static_cast<const uint4*>(gamma);
  const uint4* bp = static_cast<const uint4*>(beta);
  uint4* yp = static_cast<uint4*>(y);
  const float inv_d = 1.0f / static_cast<float>(d);
#define SAR_LN_CASE(NV) \
  case NV: ln_fwd_kernel<NV><<<grid, LN_WARPS * 32, 0, stream>>>(xp, gp, bp, yp, mean_out, rstd_out, M, nvec, inv_d, eps); break;
  switch (nv) {
    SAR_LN_CASE(1) SAR_LN_CASE(2) SAR_LN_CASE(3) SAR_LN_CASE(4) SAR_LN_CASE(5) SAR_LN_CASE(6) SAR_LN_CASE(7)
    SAR_LN_CASE(8)
    default: return fail(SAR_EINVAL, "layernorm: unsupported width");
  }
#undef SAR_LN_CASE
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "layernorm: launch");
  return SAR_OK;
}

}  // namespace sar
