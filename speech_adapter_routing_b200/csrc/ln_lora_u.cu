// ln_lora_u.cu — LayerNorm fused with the LoRA down-projection of the projection that consumes it.
//
//   x      = LayerNorm(h)                           bf16 [B, T, d]            (what sar_layernorm_fwd writes)
//   U_s    = bf16(scale · x · A_{s,k(b)}ᵀ)          bf16 [n_sets][B, T, r]    (what the U pass of the split LoRA path writes)
//
// in ONE pass over h.  In a Whisper block the LayerNorm output feeds q|k|v (reference: PEFT lora_A at every q_proj /
// v_proj, src/models/whisper_lora.py:88-98; LayerNorm at $HF/modeling_whisper.py:392, :470, :483), and the split LoRA
// path needs U = x·Aᵀ before the dense kernel can add the low-rank term as one extra K block.  Computing U in a
// separate pass re-reads all of x (147 MB at whisper-small / 64 clips: 22 % more DRAM traffic than the op's algorithmic
// bytes and ~46 us per call); here the normalised rows are still on chip when they are multiplied by A_k.
//
// HBM-bound by design (4·M·d bytes + 2·M·r·n_sets for U); the LoRA flops (2·M·d·r·n_sets: 4.7 GF at M = 96 000, r = 16,
// two sets) ride on mma.sync m16n8k16 (bf16 -> fp32).  Structure, and the measurement behind it (first version: a warp
// kept 8 rows in registers as the mma B operand and streamed A_k fragments from shared memory for every 8 rows — 91 us
// against 52 us for the plain LayerNorm: 49 KB of shared-memory fragment loads per 8 rows and 8 warps per SM at 255
// registers left it bound by the shared-memory pipe and by issue latency, ncu: issue-active 47 %, short-scoreboard stalls):
//
//   * a CTA of 8 warps owns blocks of 32 rows of one utterance.  The raw rows arrive in shared memory by 1-D bulk async
//     copies (one per row, mbarrier complete_tx), two blocks deep, so the next block streams in while this one computes;
//   * LayerNorm: one warp per row, the row in 12 registers (d = 768), two-pass fp32 statistics by warp shuffles, all
//     arithmetic on the packed fp32 pipe (add / mul / fma .f32x2: one instruction per bf16 pair), gamma / beta resident
//     in registers; x goes to global memory (coalesced 16-byte stores) and back into the staging buffer in place;
//   * U: the contraction dimension is SPLIT ACROSS THE 8 WARPS — warp w owns columns [w·d/8, (w+1)·d/8).  Its slice of
//     A_k (the mma B operand, all rank columns of every set) lives in REGISTERS for as long as the adapter does not
//     change (48 registers at d = 768, r = 16 x 2 sets), the normalised rows are the A operand, read once with
//     ldmatrix (row pitch 2d + 16 bytes: conflict-free).  No operand is ever re-read from shared memory;
//   * the 8 partial [32 x 32] fp32 products meet in shared memory and are summed in a fixed order (deterministic),
//     scaled, rounded to bf16 once and stored.
//
// U differs from the tcgen05 U pass only in fp32 summation order (same bf16 operands, same single rounding).
#include "sar_internal.h"
#include "sar_ptx.cuh"

namespace sar {

constexpr int LU_WARPS = 8;
constexpr int LU_ROWS = 32;   // rows per block: two m16 tiles

__device__ __forceinline__ void lu_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                       uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void lu_ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr)
               : "memory");
}

__device__ __forceinline__ void lu_bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ float2 unpack2(uint32_t w) {   // bf16 pair -> fp32 pair (low half = first element)
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
}

__device__ __forceinline__ float lu_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct LnLoraUParams {
  const __nv_bfloat16* h;
  const uint4* gamma;
  const uint4* beta;
  uint4* x;
  const __nv_bfloat16* A_cat;     // [n_sets * n_adapters, r, d]
  const int32_t* utt_adapter;     // [B], < 0 or >= n_adapters: base weights only (no U rows written)
  __nv_bfloat16* u_out;           // [n_sets][B, T, r]
  int B, T, d, r, n_sets, n_adapters;
  int blocks_per_utt;             // ceil(T / 32)
  long long total_blocks;         // B * blocks_per_utt
  float scale, inv_d, eps;
};

// KS = d / 128: k16-steps of one warp's column slice;  NT = n_sets * r / 8: n-tiles (8 rank columns) over all sets
template <int KS, int NT>
__global__ void __launch_bounds__(LU_WARPS * 32, 2) ln_lora_u_kernel(const LnLoraUParams p) {
  constexpr int D = KS * 128;
  constexpr int NVEC = D / 8;                   // 16-byte chunks per row
  constexpr int NV = (NVEC + 31) / 32;          // chunks per lane (row-per-warp layout)
  constexpr int ROWB = D * 2 + 16;              // staged row pitch: +16 bytes -> ldmatrix rows fall on distinct banks
  constexpr int NC = NT * 8;                    // rank columns over all sets
  constexpr int RED_LD = NC + 8;                // partial-product row pitch in floats (bank spread)
  constexpr int RED_BYTES = LU_WARPS * LU_ROWS * RED_LD * 4;
  // d = 768: two CTAs per SM leave no room for a separate reduction buffer — the partial products overwrite the
  // block's own staging slot once every warp has read its rows (x is already in global memory by then)
  constexpr bool RED_IN_SLOT = LU_ROWS * ROWB >= RED_BYTES;
  extern __shared__ __align__(16) uint8_t lu_smem[];
  uint8_t* buf = lu_smem;                                                     // [2][LU_ROWS][ROWB]
  float* red_sep = reinterpret_cast<float*>(lu_smem + 2 * LU_ROWS * ROWB);    // [LU_WARPS][LU_ROWS][RED_LD] if separate
  uint64_t* bars = reinterpret_cast<uint64_t*>(lu_smem + 2 * LU_ROWS * ROWB + (RED_IN_SLOT ? 0 : RED_BYTES));   // [2]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // gamma / beta of this lane's chunks, resident for the whole kernel
  uint4 gq[NV], bq[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    gq[i] = c < NVEC ? __ldg(p.gamma + c) : make_uint4(0u, 0u, 0u, 0u);
    bq[i] = c < NVEC ? __ldg(p.beta + c) : make_uint4(0u, 0u, 0u, 0u);
  }
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  const long long blk_begin = p.total_blocks * blockIdx.x / gridDim.x;
  const long long blk_end = p.total_blocks * (blockIdx.x + 1) / gridDim.x;
  // warp 0: stream the rows of block `blk` into staging slot `slot` (one bulk copy per row)
  auto issue = [&](long long blk, int slot) {
    const int b = static_cast<int>(blk / p.blocks_per_utt);
    const int t0 = static_cast<int>(blk - static_cast<long long>(b) * p.blocks_per_utt) * LU_ROWS;
    const int n_valid = min(LU_ROWS, p.T - t0);
    if (lane == 0) mbar_arrive_expect_tx(&bars[slot], static_cast<uint32_t>(n_valid) * D * 2);
    __syncwarp();
    if (lane < n_valid)
      lu_bulk_load(smem_u32(buf + (slot * LU_ROWS + lane) * ROWB), p.h + (static_cast<size_t>(b) * p.T + t0 + lane) * D,
                   D * 2, smem_u32(&bars[slot]));
  };
  if (warp == 0) {
    if (blk_begin < blk_end) issue(blk_begin, 0);
    if (blk_begin + 1 < blk_end) issue(blk_begin + 1, 1);
  }

  uint32_t bfrag[KS][NT][2];                    // this warp's K slice of A_k as mma B fragments
  int loaded_k = -1;
  int k_next = blk_begin < blk_end ? p.utt_adapter[blk_begin / p.blocks_per_utt] : -1;
  uint32_t phase = 0;                           // bit s = parity to wait for on slot s
  const float2 zero2 = make_float2(0.f, 0.f);

#pragma unroll 1
  for (long long blk = blk_begin; blk < blk_end; ++blk) {
    const int slot = static_cast<int>(blk - blk_begin) & 1;
    const int b = static_cast<int>(blk / p.blocks_per_utt);
    const int t0 = static_cast<int>(blk - static_cast<long long>(b) * p.blocks_per_utt) * LU_ROWS;
    const int n_valid = min(LU_ROWS, p.T - t0);
    int k = k_next;                            // fetched one block ahead: the global load is off the critical path
    if (blk + 1 < blk_end) k_next = p.utt_adapter[(blk + 1) / p.blocks_per_utt];
    if (k < 0 || k >= p.n_adapters) k = -1;
    if (k >= 0 && k != loaded_k) {
      // B fragment (K x N, "col"): b0 = {B[k0 + 2q][n], B[k0 + 2q + 1][n]}, b1 = the same at k0 + 8, with n = lane / 4,
      // q = lane % 4 and B[kk][n] = A_k[rank n][column kk]
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int col = nt * 8 + (lane >> 2);                     // rank column over all sets
        const int set = col / p.r, rho = col - set * p.r;
        const __nv_bfloat16* arow =
            p.A_cat + ((static_cast<size_t>(set) * p.n_adapters + k) * p.r + rho) * D + warp * (KS * 16) + 2 * (lane & 3);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          bfrag[ks][nt][0] = __ldg(reinterpret_cast<const uint32_t*>(arow + ks * 16));
          bfrag[ks][nt][1] = __ldg(reinterpret_cast<const uint32_t*>(arow + ks * 16 + 8));
        }
      }
      loaded_k = k;
    }
    mbar_wait(&bars[slot], (phase >> slot) & 1u);
    phase ^= 1u << slot;
    const uint32_t sbuf = smem_u32(buf + slot * LU_ROWS * ROWB);

    // ---- LayerNorm: rows warp, warp + 8, warp + 16, warp + 24 of the block (two rows interleaved per warp was
    // measured slower: 100 us vs 92 us — the extra 12 registers spill at the 128-register budget of two CTAs per SM)
#pragma unroll 1
    for (int rr = warp; rr < n_valid; rr += LU_WARPS) {
      const uint32_t src = sbuf + rr * ROWB + lane * 16;
      uint4 v[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (lane + 32 * i < NVEC)
          ld_shared_v4(src + i * 512, v[i].x, v[i].y, v[i].z, v[i].w);
        else
          v[i] = make_uint4(0u, 0u, 0u, 0u);
      }
      float2 s = zero2;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        s = fadd2(s, unpack2(v[i].x));
        s = fadd2(s, unpack2(v[i].y));
        s = fadd2(s, unpack2(v[i].z));
        s = fadd2(s, unpack2(v[i].w));
      }
      const float mean = lu_warp_sum(s.x + s.y) * p.inv_d;
      const float2 nm = make_float2(-mean, -mean);
      float2 qq = zero2;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (lane + 32 * i < NVEC) {
          float2 t;
          t = fadd2(unpack2(v[i].x), nm); qq = ffma2(t, t, qq);
          t = fadd2(unpack2(v[i].y), nm); qq = ffma2(t, t, qq);
          t = fadd2(unpack2(v[i].z), nm); qq = ffma2(t, t, qq);
          t = fadd2(unpack2(v[i].w), nm); qq = ffma2(t, t, qq);
        }
      }
      const float rstd = rsqrtf(lu_warp_sum(qq.x + qq.y) * p.inv_d + p.eps);
      const float2 rs = make_float2(rstd, rstd);
      uint4* xr = p.x + (static_cast<size_t>(b) * p.T + t0 + rr) * NVEC;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (lane + 32 * i < NVEC) {
          const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
          const uint32_t gw[4] = {gq[i].x, gq[i].y, gq[i].z, gq[i].w};
          const uint32_t bw[4] = {bq[i].x, bq[i].y, bq[i].z, bq[i].w};
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 y = ffma2(fmul2(fadd2(unpack2(w[e]), nm), rs), unpack2(gw[e]), unpack2(bw[e]));
            o[e] = pack_bf16x2(y.x, y.y);
          }
          __stcs(xr + lane + 32 * i, make_uint4(o[0], o[1], o[2], o[3]));
          st_shared_v4(src + i * 512, o[0], o[1], o[2], o[3]);    // in place: the mma A operand
        }
      }
    }
    __syncthreads();                            // the block is normalised

    // ---- U partial of this warp's K slice: [32 rows] x [NC rank columns], rows past n_valid hold stale data that
    // only ever reaches their own (never stored) output rows
    float* red = RED_IN_SLOT ? reinterpret_cast<float*>(buf + slot * LU_ROWS * ROWB) : red_sep;
    float acc[2][NT][4];
    if (k >= 0) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
        const uint32_t a_addr = sbuf + (mt * 16 + (lane & 15)) * ROWB + (warp * (KS * 16) + (lane >> 4) * 8) * 2;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t a0, a1, a2, a3;
          lu_ldmatrix_x4(a_addr + ks * 32, a0, a1, a2, a3);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) lu_mma(acc[mt][nt], a0, a1, a2, a3, bfrag[ks][nt][0], bfrag[ks][nt][1]);
        }
      }
    }
    if constexpr (RED_IN_SLOT) __syncthreads();   // every ldmatrix read of this slot is done: it may be overwritten
    if (k >= 0) {
      // D fragment: c0, c1 at (row lane / 4, columns 2 (lane % 4) + {0, 1}); c2, c3 at row + 8
      float* my_red = red + warp * LU_ROWS * RED_LD;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        float* dst = my_red + (mt * 16 + (lane >> 2)) * RED_LD + 2 * (lane & 3);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          *reinterpret_cast<float2*>(dst + nt * 8) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
          *reinterpret_cast<float2*>(dst + 8 * RED_LD + nt * 8) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
        }
      }
    }
    __syncthreads();                            // partials written (and, if separate, every read of the slot is done)
    if (!RED_IN_SLOT && warp == 0 && blk + 2 < blk_end) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(blk + 2, slot);
    }
    if (k >= 0) {
      // fixed-order sum over the 8 K slices: thread t -> row t / 8, NT consecutive rank columns
      const int row = threadIdx.x >> 3, c0 = (threadIdx.x & 7) * NT;
      float sum[NT];
#pragma unroll
      for (int e = 0; e < NT; ++e) sum[e] = 0.f;
#pragma unroll
      for (int w = 0; w < LU_WARPS; ++w) {
        const float* src = red + (w * LU_ROWS + row) * RED_LD + c0;
        if constexpr (NT == 4) {
          const float4 t = *reinterpret_cast<const float4*>(src);
          sum[0] += t.x; sum[1] += t.y; sum[2] += t.z; sum[3] += t.w;
        } else {
          const float2 t = *reinterpret_cast<const float2*>(src);
          sum[0] += t.x; sum[1] += t.y;
        }
      }
      if (row < n_valid) {
        const int set = c0 / p.r, rho = c0 - set * p.r;
        __nv_bfloat16* dst = p.u_out + ((static_cast<size_t>(set) * p.B + b) * p.T + t0 + row) * p.r + rho;
        if constexpr (NT == 4) {
          *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(sum[0] * p.scale, sum[1] * p.scale),
                                                      pack_bf16x2(sum[2] * p.scale, sum[3] * p.scale));
        } else {
          *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(sum[0] * p.scale, sum[1] * p.scale);
        }
      }
    }
    if constexpr (RED_IN_SLOT) {
      __syncthreads();                          // the partials are consumed: the slot can take block blk + 2
      if (warp == 0 && blk + 2 < blk_end) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(blk + 2, slot);
      }
    }
    // separate buffer: no barrier needed, the next write to `red` comes after the next block's post-LayerNorm barrier
  }
}

bool ln_lora_u_supported(int d, int r, int n_sets) {
  if (r % 16 || r <= 0 || n_sets < 1 || n_sets > 2) return false;
  // d <= 768: the staging buffers (2 x 32 rows), the partial products and two CTAs per SM fit in shared memory, and a
  // warp's slice of A_k fits in registers;  n_sets·r in {16, 32}
  if (!(d == 256 || d == 384 || d == 512 || d == 768)) return false;
  const int nt = n_sets * r / 8;
  return nt == 2 || nt == 4;
}

template <int KS, int NT>
static int ln_lora_u_launch(const LnLoraUParams& p, cudaStream_t stream) {
  const DeviceInfo& dev = device_info();
  constexpr int D = KS * 128;
  const int slot_bytes = LU_ROWS * (D * 2 + 16), red_bytes = LU_WARPS * LU_ROWS * (NT * 8 + 8) * 4;
  const int smem = 2 * slot_bytes + (slot_bytes >= red_bytes ? 0 : red_bytes) + 16;
  auto kern = ln_lora_u_kernel<KS, NT>;
  static thread_local int smem_set_dev[64] = {};
  int& smem_set = smem_set_dev[dev.device & 63];
  if (smem > 48 * 1024 && smem_set < smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail_cuda(e, "ln_lora_u: cudaFuncSetAttribute");
    smem_set = smem;
  }
  static thread_local int per_sm_dev[64] = {};
  int& per_sm = per_sm_dev[dev.device & 63];
  cudaError_t e = cudaSuccess;
  if (per_sm < 1) {
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, LU_WARPS * 32, smem);
    if (e != cudaSuccess || per_sm < 1) return fail_cuda(e, "ln_lora_u: occupancy query");
  }
  long long grid = static_cast<long long>(dev.num_sms) * per_sm;
  if (grid * 2 > p.total_blocks) grid = (p.total_blocks + 1) / 2;   // at least two 32-row blocks per CTA
  if (grid < 1) grid = 1;
  kern<<<static_cast<unsigned>(grid), LU_WARPS * 32, smem, stream>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "ln_lora_u: launch");
  return SAR_OK;
}

int ln_lora_u_fwd(const void* h, const void* gamma, const void* beta, void* x, const void* A_cat,
                  const int32_t* utt_adapter, void* u_out, int B, int T, int d, int r, int n_sets, int n_adapters,
                  float scale, float eps, cudaStream_t stream) {
  if (!h || !gamma || !beta || !x || !A_cat || !utt_adapter || !u_out) return fail(SAR_EINVAL, "ln_lora_u: null pointer");
  if (B <= 0 || T <= 0 || n_adapters <= 0) return fail(SAR_EINVAL, "ln_lora_u: B, T and n_adapters must be positive");
  if (!ln_lora_u_supported(d, r, n_sets))
    return fail(SAR_EINVAL, "ln_lora_u: unsupported geometry (d in {256,384,512,768}, n_sets*r in {16,32})");
  if ((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gamma) |
       reinterpret_cast<uintptr_t>(beta) | reinterpret_cast<uintptr_t>(A_cat) | reinterpret_cast<uintptr_t>(u_out)) & 15)
    return fail(SAR_EINVAL, "ln_lora_u: pointers must be 16-byte aligned");
  LnLoraUParams p{};
  p.h = static_cast<const __nv_bfloat16*>(h); p.gamma = static_cast<const uint4*>(gamma); p.beta = static_cast<const uint4*>(beta);
  p.x = static_cast<uint4*>(x); p.A_cat = static_cast<const __nv_bfloat16*>(A_cat); p.utt_adapter = utt_adapter;
  p.u_out = static_cast<__nv_bfloat16*>(u_out);
  p.B = B; p.T = T; p.d = d; p.r = r; p.n_sets = n_sets; p.n_adapters = n_adapters;
  p.blocks_per_utt = (T + LU_ROWS - 1) / LU_ROWS;
  p.total_blocks = static_cast<long long>(B) * p.blocks_per_utt;
  p.scale = scale; p.inv_d = 1.0f / static_cast<float>(d); p.eps = eps;
  const int ks = d / 128, nt = n_sets * r / 8;
#define SAR_LU_CASE(KS)                                                              \
  case KS:                                                                           \
    return nt == 4 ? ln_lora_u_launch<KS, 4>(p, stream) : ln_lora_u_launch<KS, 2>(p, stream);
  switch (ks) {
    SAR_LU_CASE(2) SAR_LU_CASE(3) SAR_LU_CASE(4) SAR_LU_CASE(6)
    default: break;
  }
#undef SAR_LU_CASE
  return fail(SAR_EINVAL, "ln_lora_u: unsupported geometry");
}

}  // namespace sar
