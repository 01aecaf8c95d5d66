// attn_fwd.cu — softmax(Q Kᵀ) V for head dim 64 on tcgen05 / TMEM / TMA (flash-attention style, forward only).
//
//   O[b,h,i,:] = sum_j softmax_j( Q[b,h,i,:]·K[b,h,j,:] ) V[b,h,j,:]        Q pre-scaled; optional causal mask (j <= i)
//
// Replaces the attention_interface call of WhisperAttention.forward ($HF/models/whisper/modeling_whisper.py:341-350)
// on the head-major tensors that sar_attn_proj_fwd writes.  One CTA owns a 128-query tile of one (b, h) and walks the
// keys in tiles of 64; two CTAs share an SM (98 KB smem, 256 TMEM columns each):
//   warp 8   TMA producer: Q once, then K_j / V_j tiles through a 5-stage ring (3-D maps (64, T, B*h): a tile never
//            crosses a head, rows past T are zero-filled)
//   warp 9   S issuer: S_j = Q·K_jᵀ (128x64x64, K-major operands) into one of two TMEM buffers, up to two tiles ahead
//            of the softmax
//   warp 10  PV issuer: O += P_j·V_j with A = P read from TENSOR MEMORY (two 32-column buffers) and B = the V tile as
//            loaded (MN-major); O accumulates IN TMEM across key tiles.  Keeping P out of shared memory matters: with P
//            staged in smem a 128x64 tile moves 80 KB through the 128 B/clk shared-memory pipe (UMMA operand reads 48 KB,
//            P stores 16 KB, TMA fills 16 KB) = 625 clk, more than the 512 clk of exponentials; without it, 48 KB.
//   warps 0-7  softmax, TWO threads per query row: warp w owns TMEM lanes 32·(w & 3).. and score columns 32·(w >> 2)..
//            Both threads of a row read all 64 scores for the row maximum (so every decision is taken identically with
//            no communication) and exponentiate their own 32: FFMA2 → MUFU.EX2 → FADD2 / bf16 pack → tcgen05.st
//            into the P buffer (row = lane, two bf16 per column: the layout the MMA reads A in).  O is rescaled (each thread its own 32 columns) only when the row maximum has grown by more than
//            2^8 since the reference was last moved — P stays <= 256 and the final O / l is exact — so the steady state
//            has no TMEM round trip for O.  The partial row sums meet once, in the epilogue.
// Why eight softmax warps: the kernel is bound by MUFU.EX2 (8 clk per warp-instruction per SMSP: 512 clk per SM for a
// 128x64 tile against 2 x 128 clk of MMA), and with one thread per row (2 warps per SMSP) the ~250 non-MUFU
// instructions of a tile issue at the 4-6 clk dependent-issue latency and leave the MUFU pipe half idle (measured: XU
// 50 %, issue slots 48 %); four warps per SMSP interleave them.
#include <cstdlib>

#include "sar_internal.h"
#include "sar_ptx.cuh"

namespace sar {

int attn_fwd2(const void* q, const void* k, const void* v, void* out, int BH, int Tq, int Tk, cudaStream_t stream);

constexpr int FA_SOFTMAX_WARPS = 8;
constexpr int FA_THREADS = (FA_SOFTMAX_WARPS + 3) * 32;   // + TMA producer, S issuer, PV issuer
constexpr int FA_BQ = 128;
constexpr int FA_BK = 64;
constexpr int FA_HD = 64;
constexpr int FA_Q_BYTES = FA_BQ * FA_HD * 2;    // 16 KB
constexpr int FA_KV_BYTES = FA_BK * FA_HD * 2;   // 8 KB
constexpr int FA_STAGES = 5;

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
  float* lsum = reinterpret_cast<float*>(v_tiles + FA_STAGES * FA_KV_BYTES);   // [2][128] row-sum exchange (epilogue)
  uint64_t* bars = reinterpret_cast<uint64_t*>(lsum + 2 * FA_BQ);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                      // [FA_STAGES]
  uint64_t* kv_empty = kv_full + FA_STAGES;          // [FA_STAGES]
  uint64_t* s_full = kv_empty + FA_STAGES;           // [2]
  uint64_t* s_free = s_full + 2;                     // [2] one arrive per softmax warp
  uint64_t* p_full = s_free + 2;                     // [2] one arrive per softmax warp
  uint64_t* o_done = p_full + 2;                     // [2] PV MMA of tile j complete -> o_done[j & 1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * FA_BQ;
  const int bh = blockIdx.y;
  int n_kv = (p.Tk + FA_BK - 1) / FA_BK;
  if (p.causal) {   // keys beyond the last query row of this tile are never visible
    const int last = min(p.Tq, q0 + FA_BQ) - 1;
    n_kv = min(n_kv, last / FA_BK + 1);
  }

  if (warp == FA_SOFTMAX_WARPS && lane == 0) {
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
      mbar_init(&s_free[i], FA_SOFTMAX_WARPS);
      mbar_init(&p_full[i], FA_SOFTMAX_WARPS);
      mbar_init(&o_done[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == FA_SOFTMAX_WARPS + 1) tmem_alloc(tmem_ptr, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // TMEM columns: S buffers at 0 / 64, O accumulator at 128, P buffers (bf16 pairs) at 192 / 224
  auto s_col = [](int j) { return static_cast<uint32_t>((j & 1) * 64); };
  auto p_col = [](int j) { return static_cast<uint32_t>(192 + (j & 1) * 32); };
  constexpr uint32_t O_COL = 128;

  if (warp == FA_SOFTMAX_WARPS) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, FA_Q_BYTES);
      tma_load_3d(q_tile, &tm_q, q_full, 0, q0, bh);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % FA_STAGES;
        if (j >= FA_STAGES) mbar_wait_backoff(&kv_empty[st], ((j / FA_STAGES) - 1) & 1, 128);
        mbar_arrive_expect_tx(&kv_full[st], 2 * FA_KV_BYTES);
        tma_load_3d(k_tiles + st * FA_KV_BYTES, &tm_k, &kv_full[st], 0, j * FA_BK, bh);
        tma_load_3d(v_tiles + st * FA_KV_BYTES, &tm_v, &kv_full[st], 0, j * FA_BK, bh);
      }
    }
  } else if (warp == FA_SOFTMAX_WARPS + 1) {
    // ------------------------------------------------------------------ S = Q K^T issuer
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(FA_BQ, FA_BK);        // both operands K-major
      const uint64_t qd = umma_desc_sw128(smem_u32(q_tile));
      mbar_wait(q_full, 0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % FA_STAGES;
        mbar_wait_backoff(&kv_full[st], (j / FA_STAGES) & 1, 32);
        if (j >= 2) mbar_wait_backoff(&s_free[j & 1], ((j >> 1) - 1) & 1, 32);   // softmax has S_{j-2} in registers
        tc_fence_after();
        const uint64_t kd = umma_desc_sw128(smem_u32(k_tiles + st * FA_KV_BYTES));
#pragma unroll
        for (int kk = 0; kk < FA_HD / 16; ++kk)
          umma_bf16(tmem_base + s_col(j), qd + 2 * kk, kd + 2 * kk, idesc_s, kk != 0);
        umma_commit(&s_full[j & 1]);
      }
    }
  } else if (warp == FA_SOFTMAX_WARPS + 2) {
    // ------------------------------------------------------------------ O += P V issuer
    if (lane == 0) {
      const uint32_t idesc_o = umma_idesc_bf16_bmn(FA_BQ, FA_HD);    // A = P (TMEM), B = V MN-major
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % FA_STAGES;
        mbar_wait_backoff(&p_full[j & 1], (j >> 1) & 1, 32);    // P_j is in TMEM (and O has been rescaled if needed)
        mbar_wait(&kv_full[st], (j / FA_STAGES) & 1);           // long complete (S_j used it); observed for V's visibility
        tc_fence_after();
        const uint64_t vd = umma_desc_sw128(smem_u32(v_tiles + st * FA_KV_BYTES));
#pragma unroll
        for (int kk = 0; kk < FA_BK / 16; ++kk)   // K = keys: 16 keys = 8 TMEM columns of P = 16 rows of 128 B in the V tile
          umma_bf16_ts(tmem_base + O_COL, tmem_base + p_col(j) + 8 * kk, vd + 128 * kk, idesc_o, (j | kk) != 0);
        umma_commit(&o_done[j & 1]);
        umma_commit(&kv_empty[st]);                             // S_j finished before P_j existed: K_j and V_j are free
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax: two threads per query row
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;
    const int qi = q0 + row;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t own_col = static_cast<uint32_t>(half * 32), other_col = 32u - own_col;
    constexpr float LOG2E = 1.4426950408889634f;
    constexpr float RESCALE_LOG2 = 8.0f;   // move the reference maximum only when it is off by more than 2^8
    float mref = -INFINITY;   // reference maximum (natural-log domain, like the scores); identical in both threads
    float l = 0.f;            // sum of this thread's 32 columns

    for (int j = 0; j < n_kv; ++j) {
      const int k0 = j * FA_BK;
      int k_lim = p.Tk - k0;              // keys k0 + e with e >= k_lim do not exist
      if (p.causal) k_lim = min(k_lim, qi - k0 + 1);
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      float mx0 = -INFINITY, mx1 = -INFINITY;
      {   // the partner's 32 scores: only their maximum is needed
        uint32_t t[32];
        tmem_ld_32x32(tmem_base + lane_addr + s_col(j) + other_col, t);
        tmem_ld_wait();
        if (k_lim < FA_BK) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (static_cast<int>(other_col) + e >= k_lim) t[e] = 0xff800000u;
        }
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          mx0 = fmaxf(mx0, __uint_as_float(t[e]));
          mx1 = fmaxf(mx1, __uint_as_float(t[e + 1]));
        }
      }
      uint32_t s[32];
      tmem_ld_32x32(tmem_base + lane_addr + s_col(j) + own_col, s);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[j & 1]);
      if (k_lim < FA_BK) {                 // edge tile: masked scores become -inf, their P is exactly 0
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (static_cast<int>(own_col) + e >= k_lim) s[e] = 0xff800000u;
      }
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        mx0 = fmaxf(mx0, __uint_as_float(s[e]));
        mx1 = fmaxf(mx1, __uint_as_float(s[e + 1]));
      }
      const float mx = fmaxf(mx0, mx1);
      // lazy rescaling: keep the old reference unless this tile's maximum exceeds it by more than 2^RESCALE_LOG2
      const bool move = (mx - mref) * LOG2E > RESCALE_LOG2;     // also true for the first finite maximum (mref = -inf)
      float alpha = 1.0f;
      if (move) {
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(alpha) : "f"((mref - mx) * LOG2E));   // 0 when mref = -inf
        mref = mx;
      }
      const bool rescale = j > 0 && __any_sync(0xffffffffu, move);
      const float mc = mref == -INFINITY ? 0.f : mref * LOG2E;   // fully masked row so far: everything stays zero
      const float2 l2e = make_float2(LOG2E, LOG2E), nmc = make_float2(-mc, -mc);
      float2 sum0 = make_float2(0.f, 0.f), sum1 = make_float2(0.f, 0.f);
      uint32_t pk[16];
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        // packed FFMA2 / FADD2: half the issue slots of the scalar forms; the exponentials stay one MUFU each
        const float2 a = ffma2(make_float2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), l2e, nmc);
        const float2 b = ffma2(make_float2(__uint_as_float(s[e + 2]), __uint_as_float(s[e + 3])), l2e, nmc);
        float2 pa, pb;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pa.x) : "f"(a.x));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pa.y) : "f"(a.y));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pb.x) : "f"(b.x));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pb.y) : "f"(b.y));
        sum0 = fadd2(sum0, pa);
        sum1 = fadd2(sum1, pb);
        pk[e >> 1] = pack_bf16x2(pa.x, pa.y);
        pk[(e >> 1) + 1] = pack_bf16x2(pb.x, pb.y);
      }
      l = fmaf(l, alpha, (sum0.x + sum0.y) + (sum1.x + sum1.y));
      // P buffer j&1 was last read by the PV MMA of tile j-2
      if (j >= 2) {
        mbar_wait(&o_done[j & 1], ((j >> 1) - 1) & 1);
        tc_fence_after();
      }
      tmem_st_32x16(tmem_base + lane_addr + p_col(j) + half * 16, pk);
      if (rescale) {
        // O (TMEM) holds sum_{i<j} P_i V_i relative to the old reference: wait for the PV MMA of tile j-1, then this
        // thread rescales its own 32 columns of the row (its partner took the same decision for the other 32)
        mbar_wait(&o_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[16];
          tmem_ld_32x16(tmem_base + lane_addr + O_COL + own_col + c * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
          tmem_st_32x16(tmem_base + lane_addr + O_COL + own_col + c * 16, v);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[j & 1]);
    }
    if (n_kv > 0) {
      mbar_wait(&o_done[(n_kv - 1) & 1], ((n_kv - 1) >> 1) & 1);   // every PV MMA has landed: the P buffers are free too
      tc_fence_after();
    }
    // the two partial row sums meet in shared memory
    lsum[half * FA_BQ + row] = l;
    asm volatile("bar.sync 1, %0;" ::"n"(FA_SOFTMAX_WARPS * 32) : "memory");
    const float lt = lsum[row] + lsum[FA_BQ + row];
    const float inv = lt > 0.f ? 1.0f / lt : 0.f;
    uint32_t v[32];
    if (n_kv > 0) {
      tmem_ld_32x32(tmem_base + lane_addr + O_COL + own_col, v);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e) v[e] = 0u;
    }
    if (qi < p.Tq) {
      uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(bh) * p.Tq + qi) * FA_HD + own_col);
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          w[i] = pack_bf16x2(__uint_as_float(v[8 * g4 + 2 * i]) * inv, __uint_as_float(v[8 * g4 + 2 * i + 1]) * inv);
        dst[g4] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == FA_SOFTMAX_WARPS + 1) {
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
  // long non-causal sequences (the encoder, long cross-attention): two query tiles per CTA, one thread per row
  // (attn_fwd2.cu); SAR_ATTN_V2=0 keeps this kernel for A/B measurements
  static const bool v2_on = [] {
    const char* e = getenv("SAR_ATTN_V2");
    return !(e && e[0] == '0');
  }();
  if (v2_on && !causal && Tq >= 384) return attn_fwd2(q, k, v, out, BH, Tq, Tk, stream);
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
  const int smem_bytes = 1024 + FA_Q_BYTES + 2 * FA_STAGES * FA_KV_BYTES + 2 * FA_BQ * 4 + 256;
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
