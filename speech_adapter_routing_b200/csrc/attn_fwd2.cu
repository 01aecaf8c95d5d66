// attn_fwd2.cu — softmax(Q Kᵀ) V for head dim 64, long non-causal sequences (the encoder's 1500 x 1500 self-attention
// and long cross-attention): second generation of attn_fwd.cu.
//
// The first kernel (one 128-query tile per CTA, two threads per query row) runs the encoder shape in 762 us against
// 540 us for cuDNN.  Its softmax warps execute ~170 instructions per 32 exponentials: both threads of a row read all 64
// scores to agree on the row maximum, the partner half's scores are loaded from TMEM a second time, and the P buffer,
// the S buffers and O each have their own hand-shake.  The kernel is bound by MUFU.EX2 in principle (8192 exponentials
// per 128 x 64 tile = 512 clk per SM against 2 x 128 clk of MMA) but needed 66 % of the issue slots to keep that pipe
// busy and got ~50 %.
//
// Here a CTA owns TWO 128-query tiles of one (b, h) and every query row has ONE thread:
//   warps 0-3 / 4-7   softmax of query tile 0 / 1: thread = row (TMEM lane), all 64 scores of a key tile: one pass over
//                     TMEM for the maximum, one for the exponentials (32 columns at a time: ~70 live registers), P (bf16)
//                     stored back OVER the score columns it came from, no partner, no row-sum exchange;
//   warp 8            TMA producer: both Q tiles once, then K_j / V_j through a 4-stage ring SHARED by the two query tiles
//                     (half the shared-memory fill per query of the first kernel);
//   warp 9            S issuer: S_t = Q_t·K_jᵀ for t = 0, 1 (128 x 64 x 64) into the tile's single S buffer, as soon as the
//                     PV MMA of the previous key tile has consumed the P that lived there;
//   warp 10           PV issuer: O_t += P_t·V_j with A = P from tensor memory and V MN-major as loaded.
// TMEM: per query tile 64 columns S/P + 64 columns O = 256 per CTA, two CTAs per SM: four query tiles per SM share the
// MUFU pipe — while one tile waits for its PV and next S, the others exponentiate.
//
// Measured (B*h = 768, 1500 x 1500): 709 us against 766 us for the first kernel and 540 us for cuDNN; XU (MUFU) pipe 63 %
// busy (ncu).  What is left, from the source-level stall samples: a softmax warp spends 31 % of its time waiting for its
// next scores — PV -> barrier -> next S -> barrier is a ~1300-cycle round trip during which nothing else of that tile can
// run, and the two tiles of a CTA move in phase because one thread issues both their MMAs.  Tried and rejected (all
// parity-green): a single pass over the scores with a speculative reference maximum (743 us: the row maximum is not what
// costs); two independent 32-key half-chains per tile, S double buffering without more tensor memory (911 us in program
// order, 1046 us with readiness-ordered issuers: twice the barrier traffic and N = 32 MMAs cost more than the overlap
// returns).
//
// Same arithmetic as attn_fwd.cu: scores in fp32, exp2 with a lazily moved reference maximum (moved only when the row
// maximum has grown by more than 2^8: P <= 256, O / l exact for any reference), P rounded to bf16 for the PV MMA, O and l
// in fp32, one bf16 rounding of O / l.
#include "sar_internal.h"
#include "sar_ptx.cuh"

namespace sar {

constexpr int F2_SOFTMAX_WARPS = 8;
constexpr int F2_THREADS = (F2_SOFTMAX_WARPS + 3) * 32;
constexpr int F2_BQ = 128;                       // rows per query tile
constexpr int F2_QT = 2;                         // query tiles per CTA
constexpr int F2_BK = 64;
constexpr int F2_HD = 64;
constexpr int F2_Q_BYTES = F2_BQ * F2_HD * 2;    // 16 KB per query tile
constexpr int F2_KV_BYTES = F2_BK * F2_HD * 2;   // 8 KB
constexpr int F2_STAGES = 4;

struct Fa2Params {
  int Tq, Tk;
  __nv_bfloat16* out;   // [B*h, Tq, 64]
};

__host__ __device__ constexpr uint32_t f2_idesc_bmn(uint32_t M, uint32_t N) {   // B operand MN-major (bit 16)
  return umma_idesc_bf16(M, N) | (1u << 16);
}

__global__ void __launch_bounds__(F2_THREADS, 2)
fa2_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
               const __grid_constant__ CUtensorMap tm_v, const Fa2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_tiles = smem;                                       // [F2_QT][16 KB]
  uint8_t* k_tiles = q_tiles + F2_QT * F2_Q_BYTES;               // [F2_STAGES][8 KB]
  uint8_t* v_tiles = k_tiles + F2_STAGES * F2_KV_BYTES;          // [F2_STAGES][8 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_tiles + F2_STAGES * F2_KV_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                      // [F2_STAGES]
  uint64_t* kv_empty = kv_full + F2_STAGES;          // [F2_STAGES]  PV of both query tiles done with the stage
  uint64_t* s_full = kv_empty + F2_STAGES;           // [F2_QT]      S_t of key tile j is in TMEM
  uint64_t* p_full = s_full + F2_QT;                 // [F2_QT]      one arrive per softmax warp of the tile
  uint64_t* o_done = p_full + F2_QT;                 // [F2_QT]      PV MMA of key tile j has completed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_done + F2_QT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (F2_QT * F2_BQ);
  const int bh = blockIdx.y;
  const int n_kv = (p.Tk + F2_BK - 1) / F2_BK;

  if (warp == F2_SOFTMAX_WARPS && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < F2_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int t = 0; t < F2_QT; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], F2_SOFTMAX_WARPS / F2_QT);
      mbar_init(&o_done[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == F2_SOFTMAX_WARPS + 1) tmem_alloc(tmem_ptr, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // TMEM columns of query tile t: S (fp32, 64) / P (bf16 pairs, 32, over the first half of S) at 128 t, O at 128 t + 64
  auto s_col = [](int t) { return static_cast<uint32_t>(t * 128); };
  auto o_col = [](int t) { return static_cast<uint32_t>(t * 128 + 64); };

  if (warp == F2_SOFTMAX_WARPS) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, F2_QT * F2_Q_BYTES);
      for (int t = 0; t < F2_QT; ++t) tma_load_3d(q_tiles + t * F2_Q_BYTES, &tm_q, q_full, 0, q0 + t * F2_BQ, bh);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % F2_STAGES;
        if (j >= F2_STAGES) mbar_wait_backoff(&kv_empty[st], ((j / F2_STAGES) - 1) & 1, 128);
        mbar_arrive_expect_tx(&kv_full[st], 2 * F2_KV_BYTES);
        tma_load_3d(k_tiles + st * F2_KV_BYTES, &tm_k, &kv_full[st], 0, j * F2_BK, bh);
        tma_load_3d(v_tiles + st * F2_KV_BYTES, &tm_v, &kv_full[st], 0, j * F2_BK, bh);
      }
    }
  } else if (warp == F2_SOFTMAX_WARPS + 1) {
    // ------------------------------------------------------------------ S = Q K^T issuer (both query tiles)
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(F2_BQ, F2_BK);        // both operands K-major
      mbar_wait(q_full, 0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % F2_STAGES;
        mbar_wait_backoff(&kv_full[st], (j / F2_STAGES) & 1, 32);
        const uint64_t kd = umma_desc_sw128(smem_u32(k_tiles + st * F2_KV_BYTES));
        for (int t = 0; t < F2_QT; ++t) {
          // the tile's S buffer also held P_{j-1}: free once the PV MMA of key tile j-1 has completed
          if (j >= 1) mbar_wait_backoff(&o_done[t], (j - 1) & 1, 32);
          tc_fence_after();
          const uint64_t qd = umma_desc_sw128(smem_u32(q_tiles + t * F2_Q_BYTES));
#pragma unroll
          for (int kk = 0; kk < F2_HD / 16; ++kk)
            umma_bf16(tmem_base + s_col(t), qd + 2 * kk, kd + 2 * kk, idesc_s, kk != 0);
          umma_commit(&s_full[t]);
        }
      }
    }
  } else if (warp == F2_SOFTMAX_WARPS + 2) {
    // ------------------------------------------------------------------ O += P V issuer (both query tiles)
    if (lane == 0) {
      const uint32_t idesc_o = f2_idesc_bmn(F2_BQ, F2_HD);           // A = P (TMEM), B = V MN-major
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % F2_STAGES;
        mbar_wait(&kv_full[st], (j / F2_STAGES) & 1);                // long complete (S_j used it); observed for V's visibility
        const uint64_t vd = umma_desc_sw128(smem_u32(v_tiles + st * F2_KV_BYTES));
        for (int t = 0; t < F2_QT; ++t) {
          mbar_wait_backoff(&p_full[t], j & 1, 32);                  // P_j is in TMEM (and O has been rescaled if needed)
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < F2_BK / 16; ++kk)   // 16 keys = 8 TMEM columns of P = 16 rows of 128 B in the V tile
            umma_bf16_ts(tmem_base + o_col(t), tmem_base + s_col(t) + 8 * kk, vd + 128 * kk, idesc_o, (j | kk) != 0);
          umma_commit(&o_done[t]);
        }
        umma_commit(&kv_empty[st]);   // every MMA that read K_j / V_j (S_j long before, PV_j of both tiles) has completed
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax: one thread per query row
    const int t = warp >> 2, quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int qi = q0 + t * F2_BQ + row;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + s_col(t);
    const uint32_t o_addr = tmem_base + lane_addr + o_col(t);
    constexpr float LOG2E = 1.4426950408889634f;
    constexpr float RESCALE_LOG2 = 8.0f;   // move the reference maximum only when it is off by more than 2^8
    float mref = -INFINITY;   // reference maximum (natural-log domain, like the scores)
    float l = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      const int k_lim = p.Tk - j * F2_BK;              // keys at or past k_lim do not exist (last tile only)
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      // pass 1: the row maximum (two 32-column loads; the values are dropped)
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t sc[32];
        tmem_ld_32x32(s_addr + hh * 32, sc);
        tmem_ld_wait();
        if (k_lim < F2_BK) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (hh * 32 + e >= k_lim) sc[e] = 0xff800000u;
        }
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          mx0 = fmaxf(mx0, __uint_as_float(sc[e]));
          mx1 = fmaxf(mx1, __uint_as_float(sc[e + 1]));
        }
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
      // pass 2: exponentials, 32 columns at a time; P (bf16 pairs) is kept in registers until both halves are done —
      // it overwrites score columns 0..31, which the second half still has to read as 32..63
      uint32_t pk[32];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t sc[32];
        tmem_ld_32x32(s_addr + hh * 32, sc);
        tmem_ld_wait();
        if (k_lim < F2_BK) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (hh * 32 + e >= k_lim) sc[e] = 0xff800000u;
        }
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float2 a = ffma2(make_float2(__uint_as_float(sc[e]), __uint_as_float(sc[e + 1])), l2e, nmc);
          const float2 b = ffma2(make_float2(__uint_as_float(sc[e + 2]), __uint_as_float(sc[e + 3])), l2e, nmc);
          float2 pa, pb;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pa.x) : "f"(a.x));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pa.y) : "f"(a.y));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pb.x) : "f"(b.x));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pb.y) : "f"(b.y));
          sum0 = fadd2(sum0, pa);
          sum1 = fadd2(sum1, pb);
          pk[hh * 16 + (e >> 1)] = pack_bf16x2(pa.x, pa.y);
          pk[hh * 16 + (e >> 1) + 1] = pack_bf16x2(pb.x, pb.y);
        }
      }
      l = fmaf(l, alpha, (sum0.x + sum0.y) + (sum1.x + sum1.y));
      tmem_st_32x32(s_addr, pk);   // P_j over S_j's first 32 columns: 64 keys x bf16
      if (rescale) {
        // O (TMEM) holds sum_{i<j} P_i V_i relative to the old reference.  The PV MMA of key tile j-1 has completed: S_j
        // was only issued after it.  Every lane scales (alpha = 1 where the row's reference did not move).
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(o_addr + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
          tmem_st_32x32(o_addr + c * 32, v);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
    }
    if (n_kv > 0) {
      mbar_wait(&o_done[t], (n_kv - 1) & 1);   // the last PV MMA of this query tile has landed
      tc_fence_after();
    }
    const float inv = l > 0.f ? 1.0f / l : 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c) {                       // 32 output columns at a time (register budget: 80)
      uint32_t v[32];
      if (n_kv > 0) {
        tmem_ld_32x32(o_addr + c * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = 0u;
      }
      if (qi < p.Tq) {
        uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(bh) * p.Tq + qi) * F2_HD + c * 32);
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == F2_SOFTMAX_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

int attn_fwd2(const void* q, const void* k, const void* v, void* out, int BH, int Tq, int Tk, cudaStream_t stream) {
  const DeviceInfo& dev = device_info();
  CUtensorMap tm_q, tm_k, tm_v;
  int rc;
  {
    const uint64_t dims[3] = {F2_HD, (uint64_t)Tq, (uint64_t)BH};
    const uint64_t strides[2] = {F2_HD * 2, (uint64_t)Tq * F2_HD * 2};
    const uint32_t box[3] = {F2_HD, F2_BQ, 1};
    if ((rc = make_tmap_bf16(&tm_q, q, 3, dims, strides, box))) return rc;
  }
  {
    const uint64_t dims[3] = {F2_HD, (uint64_t)Tk, (uint64_t)BH};
    const uint64_t strides[2] = {F2_HD * 2, (uint64_t)Tk * F2_HD * 2};
    const uint32_t box[3] = {F2_HD, F2_BK, 1};
    if ((rc = make_tmap_bf16(&tm_k, k, 3, dims, strides, box))) return rc;
    if ((rc = make_tmap_bf16(&tm_v, v, 3, dims, strides, box))) return rc;
  }
  Fa2Params p{};
  p.Tq = Tq; p.Tk = Tk;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  const int smem_bytes = 1024 + F2_QT * F2_Q_BYTES + 2 * F2_STAGES * F2_KV_BYTES + 512;
  static thread_local int smem_set_dev[64] = {};
  int& smem_set = smem_set_dev[dev.device & 63];
  if (smem_set < smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(fa2_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return fail_cuda(e, "attn_fwd2: cudaFuncSetAttribute");
    smem_set = smem_bytes;
  }
  const dim3 grid((Tq + F2_QT * F2_BQ - 1) / (F2_QT * F2_BQ), BH);
  fa2_fwd_kernel<<<grid, F2_THREADS, smem_bytes, stream>>>(tm_q, tm_k, tm_v, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "attn_fwd2: launch");
  return SAR_OK;
}

}  // namespace sar
