// skinny_fwd.cu — dense layers with at most 128 rows (decode steps: one token per utterance).
//
//   y_seg = act(x·W_segᵀ + bias_seg) + residual          x [M <= 128, K],  W_cat [n_seg*d_out, K]
//
// With so few rows a dense layer is a WEIGHT-STREAMING problem: 2·d_out·K bytes of weights per call against
// 2·M·(K + d_out) bytes of activations, ~0 reuse.  The 256-row CTA-pair kernel gives such a call d_out/128 tile steps,
// i.e. a handful of SMs pulling the whole weight matrix one K block after the other (24 us for fc2 at K = 3072).  Here
// every CTA owns a narrow N tile (32 or 64 columns) over the full K, so d_out/32 CTAs stream disjoint weight slices
// concurrently: single-CTA tcgen05 (cta_group::1, M = 128, N = 32/64), TMA-fed smem ring deep enough to hold a whole
// K = 768 slice, accumulator in TMEM, epilogue (bias / GELU / residual) straight from registers to global memory.
// Replaces, inside the captured decode step, out_proj / fc1 / fc2 / the base q|k|v projections and the lm head slices
// of one token per utterance (src/models/adapter_router.py:744-750 loop; HF modeling_whisper.py:310-353, :403-407).
#include "sar_internal.h"
#include "sar_ptx.cuh"

namespace sar {

constexpr int SK_THREADS = 192;   // warps 0-3 epilogue, warp 4 TMA producer, warp 5 TMEM alloc + MMA issue
constexpr int SK_BLOCK_K = 64;
constexpr int SK_X_BYTES = 128 * SK_BLOCK_K * 2;   // 16 KB
constexpr int SK_MAX_STAGES = 10;

struct SkinnyParams {
  int M, K, d_out, k_blocks, num_stages, nt_per_seg;
  int act;
  long long ldy, ldr;
  const __nv_bfloat16* bias;       // [n_seg*d_out] or null
  const __nv_bfloat16* residual;   // [M, ldr] or null (single segment)
  __nv_bfloat16* y[3];
};

template <int BN>
__global__ void __launch_bounds__(SK_THREADS, 1)
skinny_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, const SkinnyParams p) {
  constexpr int W_BYTES = BN * 128;
  constexpr int STAGE_BYTES = SK_X_BYTES + W_BYTES;
  constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.num_stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + SK_MAX_STAGES;
  uint64_t* acc_full = bars + 2 * SK_MAX_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = blockIdx.x;                      // N tile over all segments
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(tmem_ptr, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int KB = p.k_blocks;

  if (warp == 4) {
    if (lane == 0) {
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % S;
        if (kb >= S) mbar_wait(&empty[s], ((kb / S) - 1) & 1);
        uint8_t* st = smem + s * STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
        tma_load_2d(st, &tm_x, &full[s], kb * SK_BLOCK_K, 0);
        tma_load_2d(st + SK_X_BYTES, &tm_w, &full[s], kb * SK_BLOCK_K, nt * BN);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, BN);
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % S;
        mbar_wait(&full[s], (kb / S) & 1);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t xd = umma_desc_sw128(st), wd = umma_desc_sw128(st + SK_X_BYTES);
#pragma unroll
        for (int kk = 0; kk < SK_BLOCK_K / 16; ++kk) umma_bf16(tmem_base, xd + 2 * kk, wd + 2 * kk, idesc, (kb | kk) != 0);
        umma_commit(&empty[s]);
      }
      umma_commit(acc_full);
    }
  } else {
    // ---------------------------------------------------------------- epilogue: thread = row, BN columns
    const int row = warp * 32 + lane;
    const int seg = nt / p.nt_per_seg;
    const int n0_seg = (nt - seg * p.nt_per_seg) * BN;
    const int n0 = nt * BN;
    const bool live = row < p.M;
    uint4 bb[BN / 8], rr[BN / 8];
#pragma unroll
    for (int j = 0; j < BN / 8; ++j)
      bb[j] = p.bias ? __ldg(reinterpret_cast<const uint4*>(p.bias + n0) + j) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int j = 0; j < BN / 8; ++j) {
      rr[j] = (p.residual && live)
                  ? __ldg(reinterpret_cast<const uint4*>(p.residual + static_cast<size_t>(row) * p.ldr + n0_seg) + j)
                  : make_uint4(0u, 0u, 0u, 0u);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(taddr + c * 32, v);
      tmem_ld_wait();
      if (live) {
        uint4* dst = reinterpret_cast<uint4*>(p.y[seg] + static_cast<size_t>(row) * p.ldy + n0_seg + c * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 b4 = bb[c * 4 + j], r4 = rr[c * 4 + j];
          const uint32_t bw[4] = {b4.x, b4.y, b4.z, b4.w}, rw[4] = {r4.x, r4.y, r4.z, r4.w};
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float a0 = __uint_as_float(v[8 * j + 2 * i]) + __uint_as_float(bw[i] << 16);
            float a1 = __uint_as_float(v[8 * j + 2 * i + 1]) + __uint_as_float(bw[i] & 0xFFFF0000u);
            if (p.act == SAR_ACT_GELU) {
              const float2 ge = gelu_erf2(make_float2(a0, a1));
              a0 = ge.x;
              a1 = ge.y;
            }
            a0 += __uint_as_float(rw[i] << 16);
            a1 += __uint_as_float(rw[i] & 0xFFFF0000u);
            pk[i] = pack_bf16x2(a0, a1);
          }
          dst[j] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BN>
static int skinny_launch(const K1Args& a, int M, cudaStream_t stream) {
  const DeviceInfo& dev = device_info();
  const int n_seg = a.n_seg > 0 ? a.n_seg : 1;
  SkinnyParams p{};
  p.M = M; p.K = a.d_in; p.d_out = a.d_out;
  p.k_blocks = (a.d_in + SK_BLOCK_K - 1) / SK_BLOCK_K;
  p.nt_per_seg = a.d_out / BN;
  p.act = a.act;
  p.ldy = a.ldy > 0 ? a.ldy : a.d_out;
  p.ldr = a.ldr > 0 ? a.ldr : a.d_out;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(a.bias);
  p.residual = reinterpret_cast<const __nv_bfloat16*>(a.residual);
  for (int s = 0; s < 3; ++s) p.y[s] = reinterpret_cast<__nv_bfloat16*>(s < n_seg ? a.y_seg[s] : nullptr);
  const int stage_bytes = SK_X_BYTES + BN * 128;
  int S = (dev.max_smem_optin - 2048 - 512) / stage_bytes;
  if (S > SK_MAX_STAGES) S = SK_MAX_STAGES;
  if (S > p.k_blocks) S = p.k_blocks;
  if (S < 1) return fail(SAR_EINVAL, "skinny: shared memory budget too small");
  p.num_stages = S;
  const int smem_bytes = 1024 + S * stage_bytes + 512;
  CUtensorMap tm_x, tm_w;
  int rc;
  {
    const uint64_t ldx = a.ldx > 0 ? a.ldx : a.d_in;
    const uint64_t dims[2] = {(uint64_t)a.d_in, (uint64_t)M};
    const uint64_t strides[1] = {ldx * 2};
    const uint32_t box[2] = {SK_BLOCK_K, 128};
    if ((rc = make_tmap_bf16(&tm_x, a.x, 2, dims, strides, box))) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a.d_in, (uint64_t)n_seg * a.d_out};
    const uint64_t strides[1] = {(uint64_t)a.d_in * 2};
    const uint32_t box[2] = {SK_BLOCK_K, BN};
    if ((rc = make_tmap_bf16(&tm_w, a.W, 2, dims, strides, box))) return rc;
  }
  auto kern = skinny_kernel<BN>;
  static thread_local int smem_set_dev[64] = {};
  int& smem_set = smem_set_dev[dev.device & 63];
  if (smem_set < smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dev.max_smem_optin);
    if (e != cudaSuccess) return fail_cuda(e, "skinny: cudaFuncSetAttribute");
    smem_set = dev.max_smem_optin;
  }
  // (A programmatic-dependent-launch variant that prefetched the weight slices before griddepcontrol.wait was measured:
  // no gain — back-to-back grids fill every SM — and it would read stale weights if the preceding kernel wrote them.)
  kern<<<n_seg * p.nt_per_seg, SK_THREADS, smem_bytes, stream>>>(tm_x, tm_w, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "skinny: launch");
  return SAR_OK;
}

// true if the call is a dense, row-major, <= 128-row problem this kernel covers
bool skinny_applicable(const K1Args& a, int M) {
  if (M <= 0 || M > 128) return false;
  if (a.x_head_major || a.y_head_major || a.res_broadcast) return false;
  if (a.n_adapters > 0 && a.utt_adapter && a.A_stack && a.Bp_stack) return false;
  if (a.d_in % 64 || a.d_out % 64) return false;
  const int n_seg = a.n_seg > 0 ? a.n_seg : 1;
  if (a.residual && n_seg != 1) return false;
  for (int s = 0; s < n_seg; ++s)
    if (a.n_seg > 0 && a.seg_scale[s] != 1.0f) return false;
  return true;
}

int skinny_fwd(const K1Args& a, int M, cudaStream_t stream) {
  // 32-column tiles unless that makes more CTAs than ~2 per SM (lm head): then 64
  const int n_seg = a.n_seg > 0 ? a.n_seg : 1;
  const long long tiles32 = static_cast<long long>(n_seg) * (a.d_out / 32);
  // every CTA re-reads all of x from L2, so do not make more CTAs than SMs (whisper-large fc1: 160 tiles of 32 were
  // 18.7 us against 12.3 us on the CTA-pair kernel)
  if (a.d_out % 32 == 0 && tiles32 <= device_info().num_sms) return skinny_launch<32>(a, M, stream);
  return skinny_launch<64>(a, M, stream);
}

}  // namespace sar
