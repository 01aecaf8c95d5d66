// k3_lora_bwd.cu — K3: LoRA-only backward of K1 (base W frozen).
//
//   dx    = dy·W + (scale·dy·B_k)·A_k         -> the tcgen05 K1 kernel on transposed operands (tensor-bound)
//   dA_k += (scale·dy·B_k)ᵀ · x  = vᵀ·x       -> skinny reduction over rows (HBM-bound: one read of x)
//   dB_k += dyᵀ · (scale·x·A_kᵀ) = (uᵀ·dy)ᵀ   -> skinny reduction over rows (HBM-bound: one read of dy)
//
// Replaces autograd through PEFT lora.Linear as triggered by the reference trainer
// (src/training/trainer.py:251-256).  dA/dB are ACCUMULATED into caller-provided fp32 slices of one flat
// gradient bucket (the NCCL all-reduce buffer of the data-parallel trainer), in a fixed order: every CTA writes
// a private partial, a second kernel reduces partials chunk by chunk — no floating-point atomics.
#include <cuda_bf16.h>

#include "sar_internal.h"

namespace sar {

constexpr int K3_THREADS = 256;
constexpr int K3_COLS = 128;        // columns of x / dy per CTA
constexpr int K3_SUB = 64;          // rows per pipeline sub-tile
constexpr int K3_CHUNK_ROWS = 512;  // rows per CTA (within one utterance)
constexpr int K3_QPITCH = K3_COLS * 2 + 16;  // bytes, padded against ldmatrix bank conflicts
constexpr int64_t K3_SPLIT_MIN_ROWS = 4096;  // dx: split LoRA path from this many rows up (as ops.SPLIT_MIN_ROWS forward)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
  const int sz = pred ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct K3SkinnyParams {
  const __nv_bfloat16* P[2];  // [B*T, r]   z=0: v (for dA), z=1: u (for dB)
  const __nv_bfloat16* Q[2];  // [B*T, d]   z=0: x,          z=1: dy
  int d[2];
  float* partial[2];          // [chunks][r][d]
  const int32_t* utt_adapter;
  int T, r, chunks_per_utt, n_adapters;
};

// partial[chunk][i][c] = sum_{m in chunk} P[m][i] * Q[m][c]     (bf16 inputs, fp32 accumulate, HMMA m16n8k16)
template <int RT>  // RT = r / 16
__global__ void __launch_bounds__(K3_THREADS) k3_skinny_kernel(const K3SkinnyParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int z = blockIdx.z;
  const int d = p.d[z];
  const int c0 = blockIdx.x * K3_COLS;
  if (c0 >= d) return;
  const int chunk = blockIdx.y;
  const int b = chunk / p.chunks_per_utt;
  const int k = p.utt_adapter[b];
  if (k < 0 || k >= p.n_adapters) return;
  const int t0 = (chunk - b * p.chunks_per_utt) * K3_CHUNK_ROWS;
  const int t1 = min(p.T, t0 + K3_CHUNK_ROWS);
  const int r = RT * 16;
  const int ppitch = r * 2 + 16;
  const int stage_bytes = K3_SUB * K3_QPITCH + K3_SUB * ppitch;
  const __nv_bfloat16* Q = p.Q[z] + (static_cast<size_t>(b) * p.T) * d + c0;
  const __nv_bfloat16* P = p.P[z] + (static_cast<size_t>(b) * p.T) * r;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = static_cast<uint32_t>(__cvta_generic_to_shared(smem));

  auto load_sub = [&](int stage, int ts) {
    const uint32_t qs = smem_base + stage * stage_bytes;
    const uint32_t ps = qs + K3_SUB * K3_QPITCH;
    // Q sub-tile: 64 rows x 128 cols = 64 x 16 segments of 16 B
    for (int i = threadIdx.x; i < K3_SUB * 16; i += K3_THREADS) {
      const int row = i >> 4, seg = i & 15;
      const int t = ts + row;
      const bool live = t < t1;
      cp_async16(qs + row * K3_QPITCH + seg * 16, Q + static_cast<size_t>(live ? t : t0) * d + seg * 8, live);
    }
    const int psegs = r / 8;
    for (int i = threadIdx.x; i < K3_SUB * psegs; i += K3_THREADS) {
      const int row = i / psegs, seg = i - row * psegs;
      const int t = ts + row;
      const bool live = t < t1;
      cp_async16(ps + row * ppitch + seg * 16, P + static_cast<size_t>(live ? t : t0) * r + seg * 8, live);
    }
  };

  float acc[RT][2][4];
#pragma unroll
  for (int i = 0; i < RT; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

  const int n_sub = (t1 - t0 + K3_SUB - 1) / K3_SUB;
  load_sub(0, t0);
  cp_async_commit();
  for (int s = 0; s < n_sub; ++s) {
    if (s + 1 < n_sub) load_sub((s + 1) & 1, t0 + (s + 1) * K3_SUB);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const uint32_t qs = smem_base + (s & 1) * stage_bytes;
    const uint32_t ps = qs + K3_SUB * K3_QPITCH;
#pragma unroll
    for (int ks = 0; ks < K3_SUB / 16; ++ks) {
      const int m0 = ks * 16;
      // B fragments for this warp's 16 columns (two n-tiles)
      uint32_t bq[4];
      ldmatrix_x4_trans(qs + (m0 + (lane & 7) + ((lane >> 3) & 1) * 8) * K3_QPITCH + (warp * 16 + (lane >> 4) * 8) * 2,
                        bq);
#pragma unroll
      for (int rt = 0; rt < RT; ++rt) {
        uint32_t ap[4];
        ldmatrix_x4_trans(ps + (m0 + (lane & 7) + (lane >> 4) * 8) * ppitch + (rt * 16 + ((lane >> 3) & 1) * 8) * 2, ap);
        mma_bf16_16816(acc[rt][0], ap, bq[0], bq[1]);
        mma_bf16_16816(acc[rt][1], ap, bq[2], bq[3]);
      }
    }
    __syncthreads();
  }

  float* out = p.partial[z] + static_cast<size_t>(chunk) * r * d;
  const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
  for (int rt = 0; rt < RT; ++rt)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int col = c0 + warp * 16 + j * 8 + t4 * 2;
      float* o0 = out + static_cast<size_t>(rt * 16 + g) * d + col;
      float* o1 = out + static_cast<size_t>(rt * 16 + g + 8) * d + col;
      *reinterpret_cast<float2*>(o0) = make_float2(acc[rt][j][0], acc[rt][j][1]);
      *reinterpret_cast<float2*>(o1) = make_float2(acc[rt][j][2], acc[rt][j][3]);
    }
}

struct K3ReduceParams {
  const float* partial[2];
  float* out[2];  // z=0: dA [n_adapters][r][d_in]; z=1: dB [n_adapters][d_out][r]
  int d[2];
  const int32_t* utt_adapter;
  int B, r, chunks_per_utt, n_adapters;
};

// grid = (ceil(r*d/K3R_ELEMS), n_adapters, 2), block = K3R_ELEMS x K3R_GROUPS threads.  Thread (e, g) sums the partials of
// chunks j = g, g + G, g + 2G, ... (j = utterance * chunks_per_utt + chunk) of element e; the G group sums are then added
// in group order by one thread per element: a fixed summation order (deterministic), with G-fold shorter dependent-load
// chains than one thread per element (the one-thread version spent 23 us on 4.7 MB at B = 16: pure latency).
constexpr int K3R_ELEMS = 32;
constexpr int K3R_GROUPS = 8;

__global__ void __launch_bounds__(K3R_ELEMS * K3R_GROUPS) k3_reduce_kernel(const K3ReduceParams p) {
  __shared__ float part[K3R_GROUPS][K3R_ELEMS];
  __shared__ int any_s[K3R_GROUPS];
  const int z = blockIdx.z, k = blockIdx.y;
  const int d = p.d[z];
  const int el = threadIdx.x % K3R_ELEMS, g = threadIdx.x / K3R_ELEMS;
  const int e = blockIdx.x * K3R_ELEMS + el;
  const bool live = e < p.r * d;
  const int total = p.B * p.chunks_per_utt;
  const size_t stride = static_cast<size_t>(p.r) * d;
  float s = 0.f;
  int any = 0;
  for (int j = g; j < total; j += K3R_GROUPS) {
    const int b = j / p.chunks_per_utt;
    if (p.utt_adapter[b] != k) continue;
    any = 1;
    if (live) s += p.partial[z][static_cast<size_t>(j) * stride + e];
  }
  part[g][el] = s;
  if (el == 0) any_s[g] = any;
  __syncthreads();
  if (g != 0 || !live) return;
  int seen = 0;
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < K3R_GROUPS; ++i) {
    tot += part[i][el];
    seen |= any_s[i];
  }
  if (!seen) return;
  const int i = e / d, c = e - i * d;
  if (z == 0)
    p.out[0][(static_cast<size_t>(k) * p.r + i) * d + c] += tot;
  else
    p.out[1][(static_cast<size_t>(k) * d + c) * p.r + i] += tot;
}

static inline int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }

int64_t k3_workspace_bytes(int64_t rows, int64_t T, int64_t d, int64_t r, int64_t n_adapters) {
  (void)n_adapters;
  if (rows <= 0 || T <= 0 || d <= 0 || r <= 0) return SAR_EINVAL;
  const int64_t B = rows / T;
  const int64_t chunks = B * ((T + K3_CHUNK_ROWS - 1) / K3_CHUNK_ROWS);
  // v [rows, r] bf16 | scratch dx [rows, d] bf16 | two partial buffers [chunks][r][d] fp32
  return align256(rows * r * 2) + align256(rows * d * 2) + 2 * align256(chunks * r * d * 4);
}

int k3_qv_lora_bwd(const K3Args& a, cudaStream_t stream) {
  if (!a.dy || !a.x || !a.u || !a.Wt || !a.At_stack || !a.Bt_stack || !a.utt_adapter || !a.dA || !a.dB || !a.ws)
    return fail(SAR_EINVAL, "k3: null pointer");
  if (a.B <= 0 || a.T <= 0 || a.n_adapters <= 0) return fail(SAR_EINVAL, "k3: B, T, n_adapters must be positive");
  if (a.d_in % 128 || a.d_out % 128) return fail(SAR_EINVAL, "k3: d_in and d_out must be multiples of 128");
  if (a.r % 16 || a.r < 16 || a.r > 64) return fail(SAR_EINVAL, "k3: r must be one of 16, 32, 48, 64");
  const int64_t rows = static_cast<int64_t>(a.B) * a.T;
  const int64_t dmax = a.d_in > a.d_out ? a.d_in : a.d_out;
  const int cpu = (a.T + K3_CHUNK_ROWS - 1) / K3_CHUNK_ROWS;
  const int chunks = a.B * cpu;
  uint8_t* ws = reinterpret_cast<uint8_t*>(a.ws);
  __nv_bfloat16* v = reinterpret_cast<__nv_bfloat16*>(ws);
  ws += align256(rows * a.r * 2);
  void* dx_scratch = ws;
  ws += align256(rows * dmax * 2);
  float* partA = reinterpret_cast<float*>(ws);
  ws += align256(static_cast<int64_t>(chunks) * a.r * dmax * 4);
  float* partB = reinterpret_cast<float*>(ws);

  // dx (+ v = scale·dy·B_k saved as the "u" of the transposed problem): K1 on transposed operands.
  //   * dx == NULL (the input needs no gradient: first encoder layer): only v is needed -> the U pass alone, no d x d GEMM;
  //   * >= K3_SPLIT_MIN_ROWS rows: the split path (U pass -> v, then the dense 256-wide kernel with the low-rank term as
  //     one extra K block): the single-launch LoRA kernel is limited to 128/192-wide tiles by its TMEM budget
  //     (ncu: 52 % vs 74 % tensor-active at M = 96 000);
  //   * fewer rows: the single-launch kernel (one launch is cheaper than two below ~4 k rows).
  K1Args k1{};
  k1.x = a.dy; k1.W = a.Wt; k1.bias = nullptr; k1.A_stack = a.Bt_stack; k1.Bp_stack = a.At_stack;
  k1.utt_adapter = a.utt_adapter; k1.y = a.dx ? a.dx : dx_scratch; k1.u_out = v;
  k1.B = a.B; k1.T = a.T; k1.d_in = a.d_out; k1.d_out = a.d_in; k1.r = a.r; k1.n_adapters = a.n_adapters;
  k1.scale = a.scale;
  int rc;
  const bool pair_ok = a.d_out % 64 == 0 && a.d_in % 128 == 0;
  if (pair_ok && (!a.dx || rows >= K3_SPLIT_MIN_ROWS)) {
    k1.u_ws = v;                       // [1][B, T, r] == the [B*T, r] layout the skinny kernel reads
    k1.u_out = nullptr;
    k1.n_sets = 1;
    k1.u_phase = a.dx ? 0 : 1;
    int bn = (a.d_in % 256 == 0) ? 256 : ((a.d_in % 192 == 0) ? 192 : 128);
    rc = k1v2_qv_lora_fwd(k1, bn, stream);
  } else {
    rc = k1_qv_lora_fwd(k1, stream);
  }
  if (rc) return rc;

  K3SkinnyParams sp{};
  sp.P[0] = v; sp.Q[0] = reinterpret_cast<const __nv_bfloat16*>(a.x); sp.d[0] = a.d_in; sp.partial[0] = partA;
  sp.P[1] = reinterpret_cast<const __nv_bfloat16*>(a.u); sp.Q[1] = reinterpret_cast<const __nv_bfloat16*>(a.dy);
  sp.d[1] = a.d_out; sp.partial[1] = partB;
  sp.utt_adapter = a.utt_adapter; sp.T = a.T; sp.r = a.r; sp.chunks_per_utt = cpu; sp.n_adapters = a.n_adapters;
  const dim3 grid(static_cast<unsigned>(dmax / K3_COLS), chunks, 2);
  const int smem = 2 * (K3_SUB * K3_QPITCH + K3_SUB * (a.r * 2 + 16));
  switch (a.r / 16) {
    case 1: k3_skinny_kernel<1><<<grid, K3_THREADS, smem, stream>>>(sp); break;
    case 2: k3_skinny_kernel<2><<<grid, K3_THREADS, smem, stream>>>(sp); break;
    case 3:
      cudaFuncSetAttribute(k3_skinny_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      k3_skinny_kernel<3><<<grid, K3_THREADS, smem, stream>>>(sp);
      break;
    case 4:
      cudaFuncSetAttribute(k3_skinny_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      k3_skinny_kernel<4><<<grid, K3_THREADS, smem, stream>>>(sp);
      break;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "k3: skinny launch");

  K3ReduceParams rp{};
  rp.partial[0] = partA; rp.partial[1] = partB; rp.out[0] = a.dA; rp.out[1] = a.dB;
  rp.d[0] = a.d_in; rp.d[1] = a.d_out; rp.utt_adapter = a.utt_adapter; rp.B = a.B; rp.r = a.r;
  rp.chunks_per_utt = cpu; rp.n_adapters = a.n_adapters;
  const dim3 rgrid(static_cast<unsigned>((a.r * dmax + K3R_ELEMS - 1) / K3R_ELEMS), a.n_adapters, 2);
  k3_reduce_kernel<<<rgrid, K3R_ELEMS * K3R_GROUPS, 0, stream>>>(rp);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "k3: reduce launch");
  return SAR_OK;
}

}  // namespace sar
