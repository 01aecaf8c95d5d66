// operand_refresh.cu — one launch that re-derives every cached bf16 LoRA operand from the live adapter parameters.
//
// The kernels read LoRA weights from derived bf16 stacks (rank-padded A, 64-column B, their transposes for K3, the
// per-call concatenations of the fused projections).  Eagerly these are rebuilt on the host side whenever a parameter's
// version changes; inside a captured training step (CUDA graph) no host code runs between optimizer steps, so the
// rebuild has to be a graph node itself.  A descriptor table (built once, device resident) lists every
// (parameter block -> operand block) copy: dst[i, j] = bf16(scale · src[i, j]) with arbitrary element strides on both
// sides (transposes, column sub-blocks of padded stacks).  HBM-bound and tiny (whisper-small, r = 16: ~7 MB per step);
// one CTA per (descriptor, 4096-element tile), rows of the destination mapped to consecutive threads.
#include "sar_internal.h"
#include "sar_ptx.cuh"

#include <cuda_bf16.h>

namespace sar {

constexpr int OR_THREADS = 256;
constexpr int OR_TILE = 4096;

__global__ void __launch_bounds__(OR_THREADS) operand_refresh_kernel(const sar_refresh_desc* __restrict__ desc, int n_desc) {
  const int di = blockIdx.x;
  if (di >= n_desc) return;
  const sar_refresh_desc d = desc[di];
  const long long total = static_cast<long long>(d.rows) * d.cols;
  const long long base = static_cast<long long>(blockIdx.y) * OR_TILE;
  if (base >= total) return;
  const long long end = base + OR_TILE < total ? base + OR_TILE : total;
  // the faster-varying index follows whichever side is contiguous along columns on the destination (writes coalesce)
  const bool col_fast = d.dst_cs == 1 || d.dst_rs != 1;
  __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(d.dst);
  for (long long e = base + threadIdx.x; e < end; e += OR_THREADS) {
    long long i, j;
    if (col_fast) { i = e / d.cols; j = e - i * d.cols; }
    else          { j = e / d.rows; i = e - j * d.rows; }
    const long long so = i * d.src_rs + j * d.src_cs;
    float v = d.src_dtype == SAR_DTYPE_F32 ? static_cast<const float*>(d.src)[so]
                                           : __bfloat162float(static_cast<const __nv_bfloat16*>(d.src)[so]);
    // fp32 multiply then one rounding to bf16: the same arithmetic as the host-side builders (scale·B in fp32 -> bf16;
    // a later power-of-two fold only shifts the exponent)
    dst[i * d.dst_rs + j * d.dst_cs] = __float2bfloat16_rn(v * d.scale);
  }
}

int operand_refresh(const void* desc, int n_desc, int max_elems, cudaStream_t stream) {
  if (n_desc == 0) return SAR_OK;
  if (desc == nullptr || n_desc < 0 || max_elems <= 0) return fail(SAR_EINVAL, "operand_refresh: bad descriptor table");
  dim3 grid(n_desc, (max_elems + OR_TILE - 1) / OR_TILE);
  if (grid.y > 65535) return fail(SAR_EINVAL, "operand_refresh: block too large");
  operand_refresh_kernel<<<grid, OR_THREADS, 0, stream>>>(static_cast<const sar_refresh_desc*>(desc), n_desc);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "operand_refresh: launch");
  return SAR_OK;
}

}  // namespace sar
