// attn_fwd.cu — softmax(Q Kᵀ) V for head dim 64 on tcgen05 / TMEM / TMA (flash-attention style, forward only).
//
//   O[b,h,i,:] = sum_j softmax_j( Q[b,h,i,:]·K[b,h,j,:] ) V[b,h,j,:]        Q pre-scaled; optional causal mask (j <= i)
//
// Replaces the attention_interface call of WhisperAttention.forward ($HF/models/whisper/modeling_whisper.py:341-350)
// on the head-major tensors that sar_attn_proj_fwd writes.  One CTA owns a 128-query tile of one (b, h) and walks the
// keys in tiles of 64:
//   warp 4  TMA producer: Q once, then K_j / V_j tiles through a 2-stage ring (3-D maps (64, T, B*h): a tile never
//           crosses a head, rows past T are zero-filled)
//   warp 5  single-thread MMA issuer: S_{j+1} = Q·K_{j+1}ᵀ (128x64x64, K-major operands) is issued BEFORE the softmax of
//           tile j has finished (two S buffers in TMEM), then PV_j = P_j·V_j (A = P from shared memory, two buffers;
//           B = the V tile as loaded, MN-major; two PV buffers) — the softmax warps never wait for an MMA round trip
//   warps 0-3  softmax + accumulation, thread = query row (TMEM lane): S_j to registers, running max / sum (exp2
//           domain), P_j as bf16 into a 128-byte-swizzled smem tile, O kept in fp32 REGISTERS and rescaled as
//           O = alpha·O + PV (no TMEM read-modify-write), final O / l stored as bf16.
// Two CTAs fit an SM (64 KB smem, 128 TMEM columns each), so one CTA's softmax overlaps the other's MMAs.  The kernel is
// bound by MUFU.EX2 (one exponential per score: 8 clk per warp-instruction per SMSP against 2·128 clk of MMA per tile).
#include "sar_internal.h"
#include "sar_ptx.cuh"

namespace sar {

constexpr int FA_THREADS = 192;
constexpr int FA_BQ = 128;
constexpr int FA_BK = 64;
constexpr int FA_HD = 64;
constexpr int FA_Q_BYTES = FA_BQ * FA_HD * 2;    // 16 KB
constexpr int FA_KV_BYTES = FA_BK * FA_HD * 2;   // 8 KB
constexpr int FA_P_BYTES = FA_BQ * FA_BK * 2;    // 16 KB
constexpr int FA_STAGES = 3;

struct FaParams {
  int Tq, Tk, causal;
  __nv_bfloat16* out;   // [B*h, Tq, 64]
};

// idesc with B operand MN-major (bit 16): B tile is [K rows][N contiguous]
__host__ __device__ constexpr uint32_t umma_idesc_bf16_bmn(uint32_t M, uint32_t N) {
  return umma_idesc_bf16(M, N) | (1u << 16);
}

__global__ void __launch_bounds__(FA_THREADS, 2)
fa_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
              const __grid_constant__ CUtensorMap tm_v, const FaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_tile = smem;
  uint8_t* k_tiles = q_tile + FA_Q_BYTES;                      // [FA_STAGES][8 KB]
  uint8_t* v_tiles = k_tiles + FA_STAGES * FA_KV_BYTES;        // [FA_STAGES][8 KB]
  uint8_t* p_tile = v_tiles + FA_STAGES * FA_KV_BYTES;         // [2][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_tile + 2 * FA_P_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;        // [FA_STAGES]
  uint64_t* kv_empty = bars + 5;       // [FA_STAGES]
  uint64_t* s_full = bars + 9;         // [2]
  uint64_t* s_free = bars + 11;        // [2] count 4 (one arrive per softmax warp)
  uint64_t* p_full = bars + 13;        // [2] count 4
  uint64_t* pv_full = bars + 15;       // [2]
  uint64_t* pv_free = bars + 17;       // [2] count 4
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 19);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * FA_BQ;
  const int bh = blockIdx.y;
  int n_kv = (p.Tk + FA_BK - 1) / FA_BK;
  if (p.causal) {   // keys beyond the last query row of this tile are never visible
    const int last = min(p.Tq, q0 + FA_BQ) - 1;
    n_kv = min(n_kv, last / FA_BK + 1);
  }

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < FA_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_full[i], 1);
      mbar_init(&pv_free[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(tmem_ptr, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // TMEM columns: S buffers at 0 / 64, PV buffers at 128 / 192
  auto s_col = [](int j) { return static_cast<uint32_t>((j & 1) * 64); };
  auto pv_col = [](int j) { return static_cast<uint32_t>(128 + (j & 1) * 64); };

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, FA_Q_BYTES);
      tma_load_3d(q_tile, &tm_q, q_full, 0, q0, bh);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % FA_STAGES;
        if (j >= FA_STAGES) mbar_wait(&kv_empty[st], ((j / FA_STAGES) - 1) & 1);
        mbar_arrive_expect_tx(&kv_full[st], 2 * FA_KV_BYTES);
        tma_load_3d(k_tiles + st * FA_KV_BYTES, &tm_k, &kv_full[st], 0, j * FA_BK, bh);
        tma_load_3d(v_tiles + st * FA_KV_BYTES, &tm_v, &kv_full[st], 0, j * FA_BK, bh);
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer (software-pipelined by one tile)
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(FA_BQ, FA_BK);        // S = Q K^T : both operands K-major
      const uint32_t idesc_o = umma_idesc_bf16_bmn(FA_BQ, FA_HD);    // PV = P V  : A K-major (smem), B = V MN-major
      const uint64_t qd = umma_desc_sw128(smem_u32(q_tile));
      mbar_wait(q_full, 0);
      auto issue_s = [&](int j) {     // S_j = Q K_j^T into S buffer j&1
        const int st = j % FA_STAGES;
        mbar_wait(&kv_full[st], (j / FA_STAGES) & 1);
        if (j >= 2) mbar_wait(&s_free[j & 1], ((j >> 1) - 1) & 1);   // softmax has S_{j-2} in registers
        tc_fence_after();
        const uint64_t kd = umma_desc_sw128(smem_u32(k_tiles + st * FA_KV_BYTES));
#pragma unroll
        for (int kk = 0; kk < FA_HD / 16; ++kk)
          umma_bf16(tmem_base + s_col(j), qd + 2 * kk, kd + 2 * kk, idesc_s, kk != 0);
        umma_commit(&s_full[j & 1]);
      };
      if (n_kv > 0) issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) issue_s(j + 1);                       // overlaps the softmax of tile j
        const int st = j % FA_STAGES;
        mbar_wait(&p_full[j & 1], (j >> 1) & 1);                // P_j is in smem
        if (j >= 2) mbar_wait(&pv_free[j & 1], ((j >> 1) - 1) & 1);   // softmax has read PV_{j-2}
        tc_fence_after();
        const uint64_t pd = umma_desc_sw128(smem_u32(p_tile + (j & 1) * FA_P_BYTES));
        const uint64_t vd = umma_desc_sw128(smem_u32(v_tiles + st * FA_KV_BYTES));
#pragma unroll
        for (int kk = 0; kk < FA_BK / 16; ++kk)                 // K = keys: 16 keys = 16 rows of 128 B in the V tile
          umma_bf16(tmem_base + pv_col(j), pd + 2 * kk, vd + 128 * kk, idesc_o, kk != 0);
        umma_commit(&pv_full[j & 1]);
        umma_commit(&kv_empty[st]);                             // K_j and V_j are no longer needed
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / accumulation: thread = query row
    const int row = warp * 32 + lane;
    const int qi = q0 + row;
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    const uint32_t p_row = smem_u32(p_tile) + row * 128;
    constexpr float LOG2E = 1.4426950408889634f;
    float o[FA_HD];
#pragma unroll
    for (int i = 0; i < FA_HD; ++i) o[i] = 0.f;
    float m = -INFINITY, l = 0.f, alpha_prev = 0.f;

    auto add_pv = [&](int i) {   // o = alpha_i * o + PV_i
      mbar_wait(&pv_full[i & 1], (i >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + lane_addr + pv_col(i) + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) o[c * 32 + e] = fmaf(o[c * 32 + e], alpha_prev, __uint_as_float(v[e]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pv_free[i & 1]);
    };

    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld_32x32(tmem_base + lane_addr + s_col(j), s0);
      tmem_ld_32x32(tmem_base + lane_addr + s_col(j) + 32, s1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[j & 1]);

      const int k0 = j * FA_BK;
      int k_lim = p.Tk - k0;              // keys k0 + e with e >= k_lim do not exist
      if (p.causal) k_lim = min(k_lim, qi - k0 + 1);
      float mx = m;
      if (k_lim >= FA_BK) {               // interior tile: nothing to mask
#pragma unroll
        for (int e = 0; e < 32; ++e) mx = fmaxf(mx, fmaxf(__uint_as_float(s0[e]), __uint_as_float(s1[e])));
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float a = e < k_lim ? __uint_as_float(s0[e]) : -INFINITY;
          const float b = (e + 32) < k_lim ? __uint_as_float(s1[e]) : -INFINITY;
          s0[e] = __float_as_uint(a);
          s1[e] = __float_as_uint(b);
          mx = fmaxf(mx, fmaxf(a, b));
        }
      }
      const float mc = mx == -INFINITY ? 0.f : mx * LOG2E;   // fully masked row so far: keep everything at zero
      float alpha;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(alpha) : "f"(fmaf(m, LOG2E, -mc)));
      if (m == -INFINITY) alpha = 0.f;
      float sum0 = 0.f, sum1 = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        float p0, p1, p2, p3;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(__uint_as_float(s0[e]), LOG2E, -mc)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(__uint_as_float(s0[e + 1]), LOG2E, -mc)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p2) : "f"(fmaf(__uint_as_float(s1[e]), LOG2E, -mc)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p3) : "f"(fmaf(__uint_as_float(s1[e + 1]), LOG2E, -mc)));
        sum0 += p0 + p1;
        sum1 += p2 + p3;
        pk[e >> 1] = pack_bf16x2(p0, p1);
        pk[16 + (e >> 1)] = pack_bf16x2(p2, p3);
      }
      l = fmaf(l, alpha, sum0 + sum1);
      m = mx;
      // P row: 64 bf16 = 8 chunks of 16 B, 128-byte swizzle (A operand of the PV MMA); buffer j&1 was last read by the
      // PV MMA of tile j-2, whose result this thread consumed (add_pv) during iteration j-1
      const uint32_t prow = p_row + (j & 1) * FA_P_BYTES;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        st_shared_v4(prow + ((static_cast<uint32_t>(c) ^ sw) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[j & 1]);
      if (j > 0) add_pv(j - 1);           // deferred by one tile: the PV MMA of tile j-1 ran during this softmax
      alpha_prev = alpha;
    }
    if (n_kv > 0) add_pv(n_kv - 1);
    if (qi < p.Tq) {
      const float inv = l > 0.f ? 1.0f / l : 0.f;
      uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(bh) * p.Tq + qi) * FA_HD);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = pack_bf16x2(o[8 * c + 2 * i] * inv, o[8 * c + 2 * i + 1] * inv);
        dst[c] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

int attn_fwd(const void* q, const void* k, const void* v, void* out, int BH, int Tq, int Tk, int head_dim, int causal,
             cudaStream_t stream) {
  if (!q || !k || !v || !out) return fail(SAR_EINVAL, "attn_fwd: null pointer");
  if (head_dim != FA_HD) return fail(SAR_EINVAL, "attn_fwd: head dim must be 64");
  if (BH <= 0 || Tq <= 0 || Tk <= 0) return fail(SAR_EINVAL, "attn_fwd: sizes must be positive");
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
       reinterpret_cast<uintptr_t>(out)) & 15)
    return fail(SAR_EINVAL, "attn_fwd: pointers must be 16-byte aligned");
  if (causal && Tq != Tk) return fail(SAR_EINVAL, "attn_fwd: the causal mask is defined for Tq == Tk");
  const DeviceInfo& dev = device_info();
  CUtensorMap tm_q, tm_k, tm_v;
  int rc;
  {
    const uint64_t dims[3] = {FA_HD, (uint64_t)Tq, (uint64_t)BH};
    const uint64_t strides[2] = {FA_HD * 2, (uint64_t)Tq * FA_HD * 2};
    const uint32_t box[3] = {FA_HD, FA_BQ, 1};
    if ((rc = make_tmap_bf16(&tm_q, q, 3, dims, strides, box))) return rc;
  }
  {
    const uint64_t dims[3] = {FA_HD, (uint64_t)Tk, (uint64_t)BH};
    const uint64_t strides[2] = {FA_HD * 2, (uint64_t)Tk * FA_HD * 2};
    const uint32_t box[3] = {FA_HD, FA_BK, 1};
    if ((rc = make_tmap_bf16(&tm_k, k, 3, dims, strides, box))) return rc;
    if ((rc = make_tmap_bf16(&tm_v, v, 3, dims, strides, box))) return rc;
  }
  FaParams p{};
  p.Tq = Tq; p.Tk = Tk; p.causal = causal;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  const int smem_bytes = 1024 + FA_Q_BYTES + 2 * FA_STAGES * FA_KV_BYTES + 2 * FA_P_BYTES + 256;
  static thread_local int smem_set_dev[64] = {};
  int& smem_set = smem_set_dev[dev.device & 63];
  if (smem_set < smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(fa_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return fail_cuda(e, "attn_fwd: cudaFuncSetAttribute");
    smem_set = smem_bytes;
  }
  const dim3 grid((Tq + FA_BQ - 1) / FA_BQ, BH);
  fa_fwd_kernel<<<grid, FA_THREADS, smem_bytes, stream>>>(tm_q, tm_k, tm_v, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "attn_fwd: launch");
  return SAR_OK;
}

}  // namespace sar
