// k1_qv_lora_fwd.cu — K1: fused base q/v GEMM + routed low-rank (LoRA) epilogue for sm_100a.
//
//   y[b,t,:] = x[b,t,:]·Wᵀ + bias + (scale·x[b,t,:]·A_kᵀ)·B_kᵀ,   k = utt_adapter[b]
//
// Replaces PEFT lora.Linear.forward at the q_proj/v_proj slots (reference src/models/whisper_lora.py:88-98,
// call sites $HF/models/whisper/modeling_whisper.py:310,332) and the per-utterance adapter loop of
// src/models/adapter_router.py:610-622.
//
// Design (one CTA per SM, persistent over "units"):
//   unit      = one 128-row tile of ONE utterance (3-D TMA map [B,T,d], box [1,128,64]: rows past T are
//               zero-filled on load and clipped on store), so a tile never mixes adapters.
//   per unit  : the CTA walks all N tiles (BLOCK_N columns each) with a double-buffered fp32 accumulator in TMEM.
//               While accumulating N-tile 0 it also accumulates U = X·A_kᵀ (N = r) from the SAME smem X stages.
//               The epilogue warps turn U into scale·U (bf16) in a 128B-swizzled smem tile, and every N tile then
//               gets r/16 extra MMAs  acc += U·B_k[n-tile]ᵀ.  U never leaves the SM; A_k is streamed once per unit,
//               B_k once per (unit, N tile) from L2.
//   warps     : 0 = TMA producer, 1 = TMEM allocator + MMA issuer (one thread), 2..5 = epilogue
//               (TMEM → regs → +bias → bf16 → swizzled smem → TMA store, each warp stores its own 32 rows).
#include "sar_internal.h"
#include "sar_ptx.cuh"

namespace sar {

constexpr int K1_BLOCK_M = 128;
constexpr int K1_BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int K1_THREADS = 192;
constexpr int K1_X_BYTES = K1_BLOCK_M * K1_BLOCK_K * 2;  // 16 KB
constexpr int K1_U_BYTES = K1_BLOCK_M * 128;             // 16 KB (128 rows x 64 bf16, first r columns used)
constexpr int K1_STG_BYTES = 32 * 128;                   // one epilogue warp's 32 rows x 64 bf16
constexpr int K1_MAX_STAGES = 8;

struct K1Params {
  int B, T, d_in, d_out, r;
  int tiles_per_utt, num_units, n_tiles, k_blocks, num_stages, n_adapters;
  float scale;
  const int32_t* utt_adapter;  // nullable
  const __nv_bfloat16* bias;   // nullable
  __nv_bfloat16* u_out;        // nullable: [B*T, r]
};

template <int BLOCK_N>
struct K1Smem {
  static constexpr int W_BYTES = BLOCK_N * 128;
  static constexpr int BP_BYTES = BLOCK_N * 128;
  static __host__ __device__ int a_bytes(int r) { return r * 128; }
  static __host__ __device__ int stage_bytes(int r) { return K1_X_BYTES + W_BYTES + a_bytes(r); }
  static __host__ __device__ int fixed_bytes() { return K1_U_BYTES + BP_BYTES + 4 * 2 * K1_STG_BYTES + 512; }
};

template <int BLOCK_N>
__global__ void __launch_bounds__(K1_THREADS, 1)
k1_qv_lora_fwd_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                      const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                      const __grid_constant__ CUtensorMap tm_y, const K1Params p) {
  using L = K1Smem<BLOCK_N>;
  constexpr int TMEM_COLS = 512;
  constexpr int U_COL = 2 * BLOCK_N;  // U accumulator columns [U_COL, U_COL + r)
  static_assert(2 * BLOCK_N + 64 <= TMEM_COLS, "TMEM budget");
  static_assert(BLOCK_N % 64 == 0 && BLOCK_N <= 256, "BLOCK_N");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.num_stages;
  const int stage_bytes = L::stage_bytes(p.r);
  uint8_t* stages = smem;
  uint8_t* u_tile = stages + S * stage_bytes;
  uint8_t* bp_tile = u_tile + K1_U_BYTES;
  uint8_t* stg = bp_tile + L::BP_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg + 4 * 2 * K1_STG_BYTES);
  uint64_t* full = bars;                       // [S]
  uint64_t* empty = bars + K1_MAX_STAGES;      // [S]
  uint64_t* tmem_full = bars + 2 * K1_MAX_STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;             // [2]
  uint64_t* u_full = tmem_empty + 2;
  uint64_t* u_ready = u_full + 1;
  uint64_t* b_full = u_ready + 1;
  uint64_t* b_empty = b_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(b_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool has_lora = (p.n_adapters > 0) && (p.utt_adapter != nullptr);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_y);
    if (has_lora) {
      tma_prefetch_desc(&tm_a);
      tma_prefetch_desc(&tm_b);
    }
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    mbar_init(u_full, 1);
    mbar_init(u_ready, 4);
    mbar_init(b_full, 1);
    mbar_init(b_empty, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int KB = p.k_blocks;
  const int NT = p.n_tiles;
  const int bp_issue_kb = KB > 2 ? KB / 2 : 0;

  if (warp == 0) {
    // =============================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t b_uses = 0;
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
        const int b = unit / p.tiles_per_utt;
        const int m0 = (unit - b * p.tiles_per_utt) * K1_BLOCK_M;
        int k = has_lora ? p.utt_adapter[b] : -1;
        if (k < 0 || k >= p.n_adapters) k = -1;
        for (int nt = 0; nt < NT; ++nt) {
          const int n0 = nt * BLOCK_N;
          const bool with_a = (k >= 0) && (nt == 0);
          const uint32_t tx = K1_X_BYTES + L::W_BYTES + (with_a ? L::a_bytes(p.r) : 0);
          for (int kb = 0; kb < KB; ++kb) {
            if (k >= 0 && kb == bp_issue_kb) {
              mbar_wait(b_empty, (b_uses & 1) ^ 1);
              mbar_arrive_expect_tx(b_full, L::BP_BYTES);
              tma_load_2d(bp_tile, &tm_b, b_full, 0, k * p.d_out + n0);
              ++b_uses;
            }
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* st = stages + stage * stage_bytes;
            mbar_arrive_expect_tx(&full[stage], tx);
            tma_load_3d(st, &tm_x, &full[stage], kb * K1_BLOCK_K, m0, b);
            tma_load_2d(st + K1_X_BYTES, &tm_w, &full[stage], kb * K1_BLOCK_K, n0);
            if (with_a) tma_load_2d(st + K1_X_BYTES + L::W_BYTES, &tm_a, &full[stage], kb * K1_BLOCK_K, k * p.r);
            if (++stage == S) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =============================================================== MMA issuer (single thread)
    if (lane == 0) {
      const uint32_t idesc_main = umma_idesc_bf16(K1_BLOCK_M, BLOCK_N);
      const uint32_t idesc_u = umma_idesc_bf16(K1_BLOCK_M, p.r);
      const uint32_t u_desc_base = smem_u32(u_tile);
      const uint32_t bp_desc_base = smem_u32(bp_tile);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tile_iter = 0, lora_units = 0, b_uses = 0;
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
        const int b = unit / p.tiles_per_utt;
        int k = has_lora ? p.utt_adapter[b] : -1;
        if (k < 0 || k >= p.n_adapters) k = -1;
        for (int nt = 0; nt < NT; ++nt, ++tile_iter) {
          const uint32_t buf = tile_iter & 1;
          const uint32_t acc = tmem_base + buf * BLOCK_N;
          const bool with_a = (k >= 0) && (nt == 0);
          mbar_wait(&tmem_empty[buf], ((tile_iter >> 1) & 1) ^ 1);
          tc_fence_after();
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t st = smem_u32(stages + stage * stage_bytes);
            const uint64_t xd = umma_desc_sw128(st);
            const uint64_t wd = umma_desc_sw128(st + K1_X_BYTES);
#pragma unroll
            for (int kk = 0; kk < K1_BLOCK_K / 16; ++kk)
              umma_bf16(acc, xd + 2 * kk, wd + 2 * kk, idesc_main, (kb | kk) != 0);
            if (with_a) {
              const uint64_t ad = umma_desc_sw128(st + K1_X_BYTES + L::W_BYTES);
#pragma unroll
              for (int kk = 0; kk < K1_BLOCK_K / 16; ++kk)
                umma_bf16(tmem_base + U_COL, xd + 2 * kk, ad + 2 * kk, idesc_u, (kb | kk) != 0);
            }
            umma_commit(&empty[stage]);
            if (++stage == S) {
              stage = 0;
              phase ^= 1;
            }
          }
          if (k >= 0) {
            if (nt == 0) {
              umma_commit(u_full);  // U accumulator complete -> epilogue converts it
              mbar_wait(u_ready, lora_units & 1);
              tc_fence_after();
            }
            mbar_wait(b_full, b_uses & 1);
            tc_fence_after();
            const uint64_t ud = umma_desc_sw128(u_desc_base);
            const uint64_t bd = umma_desc_sw128(bp_desc_base);
            const int ksteps = p.r >> 4;
            for (int kk = 0; kk < ksteps; ++kk) umma_bf16(acc, ud + 2 * kk, bd + 2 * kk, idesc_main, 1u);
            umma_commit(b_empty);
            ++b_uses;
          }
          umma_commit(&tmem_full[buf]);
        }
        if (k >= 0) ++lora_units;
      }
    }
    __syncwarp();
  } else {
    // =============================================================== epilogue warps (2..5)
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    uint8_t* my_stg = stg + q * (2 * K1_STG_BYTES);
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    uint32_t tile_iter = 0, lora_units = 0, stg_idx = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      const int b = unit / p.tiles_per_utt;
      const int m0 = (unit - b * p.tiles_per_utt) * K1_BLOCK_M;
      int k = has_lora ? p.utt_adapter[b] : -1;
      if (k < 0 || k >= p.n_adapters) k = -1;
      if (k >= 0) {
        // ---- U: TMEM fp32 -> scale -> bf16 -> 128B-swizzled smem A-operand tile (+ optional global save)
        mbar_wait(u_full, lora_units & 1);
        tc_fence_after();
        const uint32_t u_row = smem_u32(u_tile) + row * 128;
        const bool save = (p.u_out != nullptr) && (m0 + row < p.T);
        uint4* u_dst = save ? reinterpret_cast<uint4*>(p.u_out + (static_cast<size_t>(b) * p.T + m0 + row) * p.r)
                            : nullptr;
        for (int j = 0; j < (p.r >> 4); ++j) {
          uint32_t v[16];
          tmem_ld_32x16(tmem_base + lane_addr + U_COL + j * 16, v);
          tmem_ld_wait();
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            pk[i] = pack_bf16x2(__uint_as_float(v[2 * i]) * p.scale, __uint_as_float(v[2 * i + 1]) * p.scale);
          st_shared_v4(u_row + (((2 * j) ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
          st_shared_v4(u_row + (((2 * j + 1) ^ sw) << 4), pk[4], pk[5], pk[6], pk[7]);
          if (save) {
            u_dst[2 * j] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            u_dst[2 * j + 1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(u_ready);
        ++lora_units;
      }
      for (int nt = 0; nt < NT; ++nt, ++tile_iter) {
        const uint32_t buf = tile_iter & 1;
        const int n0 = nt * BLOCK_N;
        mbar_wait(&tmem_full[buf], (tile_iter >> 1) & 1);
        tc_fence_after();
        const bool rows_live = (m0 + q * 32) < p.T;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 64; ++c) {
          uint32_t v0[32], v1[32];
          const uint32_t taddr = tmem_base + lane_addr + buf * BLOCK_N + c * 64;
          tmem_ld_32x32(taddr, v0);
          tmem_ld_32x32(taddr + 32, v1);
          // make sure the TMA store that last used this staging buffer has finished reading it
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          tmem_ld_wait();
          uint8_t* sbuf = my_stg + (stg_idx & 1) * K1_STG_BYTES;
          const uint32_t srow = smem_u32(sbuf) + lane * 128;
          const uint4* bias4 =
              p.bias ? reinterpret_cast<const uint4*>(p.bias + n0 + c * 64) : nullptr;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float bf[8];
            if (bias4) {
              const uint4 bb = __ldg(bias4 + j);
              const uint32_t w[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                bf[2 * i] = __uint_as_float(w[i] << 16);
                bf[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) bf[i] = 0.f;
            }
            uint32_t pk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int e = 8 * j + 2 * i;
              const float a0 = __uint_as_float(e < 32 ? v0[e & 31] : v1[e & 31]);
              const float a1 = __uint_as_float(e < 32 ? v0[(e + 1) & 31] : v1[(e + 1) & 31]);
              pk[i] = pack_bf16x2(a0 + bf[2 * i], a1 + bf[2 * i + 1]);
            }
            st_shared_v4(srow + ((static_cast<uint32_t>(j) ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && rows_live) {
            tma_store_3d(&tm_y, sbuf, n0 + c * 64, m0 + q * 32, b);
            tma_store_commit();
          }
          ++stg_idx;
        }
        // all TMEM reads of this accumulator buffer are complete (tmem_ld_wait above)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[buf]);
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ host
template <int BLOCK_N>
static int k1_launch(const K1Args& a, cudaStream_t stream) {
  using L = K1Smem<BLOCK_N>;
  const DeviceInfo& dev = device_info();
  const bool has_lora = a.n_adapters > 0 && a.utt_adapter != nullptr && a.A_stack != nullptr && a.Bp_stack != nullptr;

  K1Params p{};
  p.B = a.B; p.T = a.T; p.d_in = a.d_in; p.d_out = a.d_out; p.r = has_lora ? a.r : 16;
  p.tiles_per_utt = (a.T + K1_BLOCK_M - 1) / K1_BLOCK_M;
  p.num_units = a.B * p.tiles_per_utt;
  p.n_tiles = a.d_out / BLOCK_N;
  p.k_blocks = a.d_in / K1_BLOCK_K;
  p.n_adapters = has_lora ? a.n_adapters : 0;
  p.scale = a.scale;
  p.utt_adapter = has_lora ? a.utt_adapter : nullptr;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(a.bias);
  p.u_out = has_lora ? reinterpret_cast<__nv_bfloat16*>(a.u_out) : nullptr;

  const int stage_bytes = L::stage_bytes(p.r);
  const int budget = dev.max_smem_optin - 1024 - L::fixed_bytes();
  int S = budget / stage_bytes;
  if (S > K1_MAX_STAGES) S = K1_MAX_STAGES;
  if (S < 2) return fail(SAR_EINVAL, "k1: shared memory budget too small for this shape");
  p.num_stages = S;
  const int smem_bytes = 1024 + S * stage_bytes + L::fixed_bytes();

  CUtensorMap tm_x, tm_w, tm_a, tm_b, tm_y;
  memset(&tm_a, 0, sizeof(tm_a));
  memset(&tm_b, 0, sizeof(tm_b));
  int rc;
  {
    const uint64_t dims[3] = {(uint64_t)a.d_in, (uint64_t)a.T, (uint64_t)a.B};
    const uint64_t strides[2] = {(uint64_t)a.d_in * 2, (uint64_t)a.T * a.d_in * 2};
    const uint32_t box[3] = {K1_BLOCK_K, K1_BLOCK_M, 1};
    if ((rc = make_tmap_bf16(&tm_x, a.x, 3, dims, strides, box))) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a.d_out, (uint64_t)a.T, (uint64_t)a.B};
    const uint64_t strides[2] = {(uint64_t)a.d_out * 2, (uint64_t)a.T * a.d_out * 2};
    const uint32_t box[3] = {64, 32, 1};
    if ((rc = make_tmap_bf16(&tm_y, a.y, 3, dims, strides, box))) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a.d_in, (uint64_t)a.d_out};
    const uint64_t strides[1] = {(uint64_t)a.d_in * 2};
    const uint32_t box[2] = {K1_BLOCK_K, BLOCK_N};
    if ((rc = make_tmap_bf16(&tm_w, a.W, 2, dims, strides, box))) return rc;
  }
  if (has_lora) {
    {
      const uint64_t dims[2] = {(uint64_t)a.d_in, (uint64_t)a.n_adapters * a.r};
      const uint64_t strides[1] = {(uint64_t)a.d_in * 2};
      const uint32_t box[2] = {K1_BLOCK_K, (uint32_t)a.r};
      if ((rc = make_tmap_bf16(&tm_a, a.A_stack, 2, dims, strides, box))) return rc;
    }
    {
      const uint64_t dims[2] = {(uint64_t)SAR_RPAD, (uint64_t)a.n_adapters * a.d_out};
      const uint64_t strides[1] = {(uint64_t)SAR_RPAD * 2};
      const uint32_t box[2] = {64, BLOCK_N};
      if ((rc = make_tmap_bf16(&tm_b, a.Bp_stack, 2, dims, strides, box))) return rc;
    }
  }

  auto kern = k1_qv_lora_fwd_kernel<BLOCK_N>;
  // the opt-in shared-memory limit is a per-device attribute of the kernel: remember it per device (and per thread:
  // no lock needed, the call is idempotent)
  static thread_local int smem_set_dev[64] = {};
  int& smem_set = smem_set_dev[dev.device & 63];
  if (smem_set < smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dev.max_smem_optin);
    if (e != cudaSuccess) return fail_cuda(e, "k1: cudaFuncSetAttribute");
    smem_set = dev.max_smem_optin;
  }
  int grid = p.num_units < dev.num_sms ? p.num_units : dev.num_sms;
  if (a.grid_override > 0 && a.grid_override < grid) grid = a.grid_override;
  kern<<<grid, K1_THREADS, smem_bytes, stream>>>(tm_x, tm_w, tm_a, tm_b, tm_y, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "k1: launch");
  return SAR_OK;
}

int k1_qv_lora_fwd(const K1Args& a, cudaStream_t stream) {
  if (!a.x || !a.W || !a.y) return fail(SAR_EINVAL, "k1: null x/W/y");
  if (a.B <= 0 || a.T <= 0) return fail(SAR_EINVAL, "k1: B and T must be positive");
  if (a.d_in % 64 || a.d_out % 64 || a.d_in <= 0 || a.d_out <= 0)
    return fail(SAR_EINVAL, "k1: d_in and d_out must be positive multiples of 64");
  if (a.n_adapters > 0 && a.utt_adapter && (a.r % 16 || a.r < 16 || a.r > 64))
    return fail(SAR_EINVAL, "k1: r must be one of 16, 32, 48, 64");
  if ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.W) | reinterpret_cast<uintptr_t>(a.y) |
       reinterpret_cast<uintptr_t>(a.A_stack) | reinterpret_cast<uintptr_t>(a.Bp_stack) |
       reinterpret_cast<uintptr_t>(a.bias) | reinterpret_cast<uintptr_t>(a.u_out)) & 15)
    return fail(SAR_EINVAL, "k1: pointers must be 16-byte aligned");
  int bn = a.block_n_override;
  if (bn == 0) bn = (a.d_out % 192 == 0) ? 192 : (a.d_out % 128 == 0) ? 128 : 64;
  if (a.d_out % bn) return fail(SAR_EINVAL, "k1: d_out not divisible by BLOCK_N");
  // CTA-pair kernel (cta_group::2, 256-row units) whenever an utterance has more than one 128-row tile; the
  // single-CTA kernel keeps short sequences (decoder T_dec <= 128, decode steps) from wasting half a pair.
  const bool pair_ok = (bn == 128 || bn == 192);
  const bool want_pair = a.kernel_override == 2 || (a.kernel_override == 0 && a.T > 128);
  if (pair_ok && want_pair) return k1v2_qv_lora_fwd(a, bn, stream);
  switch (bn) {
    case 64: return k1_launch<64>(a, stream);
    case 128: return k1_launch<128>(a, stream);
    case 192: return k1_launch<192>(a, stream);
    default: return fail(SAR_EINVAL, "k1: unsupported BLOCK_N");
  }
}

}  // namespace sar
