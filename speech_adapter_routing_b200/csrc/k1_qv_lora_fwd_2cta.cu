// k1_qv_lora_fwd_2cta.cu — K1 v2: the fused q/v GEMM + routed LoRA epilogue on CTA PAIRS (tcgen05 cta_group::2).
//
// Same math and same per-unit structure as k1_qv_lora_fwd.cu (see the design notes there), re-tiled so that one
// tcgen05.mma spans two SMs: a pair owns 256 rows of ONE utterance (128 per CTA) and every B operand (W tile, A_k
// tile, B_k tile) is split in halves between the two CTAs, which halves the per-SM operand ingest from L2 — the
// limiter of the single-CTA kernel (ncu: tensor pipe 50 % active at 104 B/clk/SM demanded vs ~64 B/clk/SM served).
//
// Scheduling: work is the flat list of (unit, N-tile) steps.
//   LORA kernels: pair p owns a contiguous, balanced range of it, so the low-rank intermediate U of a unit is built
//     once and reused by the following N tiles, and the last wave is never more than one tile-step long.  A range
//     that starts in the middle of a unit first runs a U-only pass (loads X and A_k, issues only the N=r MMAs).
//   dense kernels (no low-rank term: out_proj / fc1 / fc2 / base-only projections): steps are dealt round-robin
//     (pair p takes g = p, p + P, ...), so at any moment the P pairs work on ~P/NT neighbouring units and all N
//     tiles of each — the X rows of a unit are fetched from HBM once and shared through L2 even when K is large
//     (fc2: a 256-row X tile is 1.5 MB; with contiguous ranges 74 of them thrash L2 and X is re-read NT times).
//
// The kernel is specialised at compile time on <BLOCK_N, LORA, EPI> so that each variant's epilogue is a few hundred
// instructions: ncu showed the one-size-fits-all epilogue (runtime GELU / residual / scale branches, ~2300
// instructions per 64-column chunk) stalled on instruction fetch (stall_no_inst) and paced the tensor pipe at 30 %.
//
// Roles per CTA (192 threads): warps 0-3 epilogue on the CTA's own 128 TMEM lanes, warp 4 TMA producer (both CTAs
// load their halves; completion bytes are signalled on the LEADER's mbarriers), warp 5 TMEM alloc + (leader only)
// single-thread MMA issue with multicast commits.
#include "sar_internal.h"
#include "sar_ptx.cuh"

namespace sar {

constexpr int V2_ROWS_PER_CTA = 128;
constexpr int V2_BLOCK_K = 64;

constexpr int V2_X_BYTES = V2_ROWS_PER_CTA * V2_BLOCK_K * 2;  // 16 KB
constexpr int V2_U_BYTES = V2_ROWS_PER_CTA * 128;             // 16 KB
constexpr int V2_STG_BYTES = 32 * 128;
constexpr int V2_MAX_STAGES = 8;

struct K1V2Params {
  int B, T, d_in, d_out, r;
  int tiles_per_utt;   // 256-row tiles per utterance
  int n_tiles;         // N tiles over ALL segments (n_seg * nt_per_seg)
  int nt_per_seg;      // N tiles per output segment (d_out / BLOCK_N)
  int n_seg;           // output segments sharing x: 1 (q or v), 2 (k|v), 3 (q|k|v)
  int n_sets;          // LoRA sets (A/B stacks) in use: 1 or 2
  int seg_set[3];      // LoRA set of each segment, -1 = none (k_proj)
  float seg_scale[3];  // epilogue scale of each segment (head_dim^-0.5 for q)
  int x_head_major;    // x is [B, h, T, 64] (SDPA output): K block kb <-> head kb
  int y_head_major;    // y is [B, h, T, 64] (SDPA input layout)
  int k_blocks, num_stages, n_adapters;
  long long total_steps;  // num_units * n_tiles
  int num_pairs;
  int swap_halves;     // debug: which CTA of the pair supplies the upper half of every B operand
  float scale;
  const int32_t* utt_adapter;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* residual;   // [B, T, d_out] added in the epilogue (n_seg == 1, row-major y), or null
  long long ldr, res_bs;           // residual row stride / batch stride in elements (res_bs = 0: broadcast over b)
  int act;                         // SAR_ACT_*: 0 none, 1 erf-GELU applied to (acc + bias) before the residual
  __nv_bfloat16* u_out;
  int u_only;        // LORA kernels: compute and save U = scale·X·A_kᵀ only (no output tiles): the split path's first launch
  int u_ld;          // > 0: u_out is [n_sets][B, T, u_ld = r] (one compact plane per LoRA set); 0: legacy [B*T, r], set 0 only
  // weighted multi-adapter mix (soft_fused routing): the "adapter" is the concatenation of n language adapters along the
  // rank (A [n*r0, d], B [d_out, n*r0]); rank columns [g*r0, (g+1)*r0) of utterance b's U are scaled by u_w[b*u_w_ld + g]
  const float* u_w;
  int u_w_ld, u_w_group;
};

enum : int { EPI_RES = 1, EPI_GELU = 2, EPI_SCALE = 4, EPI_DGELU = 8 };   // DGELU: y = acc * GELU'(residual operand)
// Epilogue warps per CTA.  The GELU epilogue is latency-bound with one warp per SMSP (ncu: 31 % wait + 12 % MUFU
// scoreboard stalls) and the residual epilogue waits on its global loads: two warps per TMEM lane quadrant interleave
// their dependency chains, and each then owns at most two column chunks whose residual rows are both requested before
// the accumulator is awaited.  Measured: fc1+GELU 488 -> 365 us, out_proj+residual (K = 768) 165 -> 137 us; with a long
// K loop (fc2, K = 3072) the epilogue is hidden anyway and the 8-warp variant is ~8 % slower, so the host picks by K.


template <int BLOCK_N, bool LORA>
struct V2Smem {
  static constexpr int WH_BYTES = (BLOCK_N / 2) * 128;
  static constexpr int BPH_BYTES = LORA ? (BLOCK_N / 2) * 128 : 0;
  static __host__ __device__ int ah_bytes(int r) { return (r / 2) * 128; }                      // bytes TMA writes
  static __host__ __device__ int ah_slot(int r) { return ((r / 2) * 128 + 1023) & ~1023; }      // 1 KB-aligned slot
  static __host__ __device__ int stage_bytes(int r, int n_sets) {
    return V2_X_BYTES + WH_BYTES + (LORA ? n_sets * ah_slot(r) : 0);
  }
  static __host__ __device__ int fixed_bytes(int n_sets) {
    return (LORA ? n_sets * V2_U_BYTES : 0) + BPH_BYTES + 4 * 2 * V2_STG_BYTES + 512;
  }
};

template <int BLOCK_N, bool LORA, int EPI, int EW, bool AUG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(32 * (EW + 2), 1)
k1v2_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
            const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
            const __grid_constant__ CUtensorMap tm_y0, const __grid_constant__ CUtensorMap tm_y1,
            const __grid_constant__ CUtensorMap tm_y2, const K1V2Params p) {
  using L = V2Smem<BLOCK_N, LORA>;
  constexpr int TMEM_COLS = 512;
  constexpr int U_COL = 2 * BLOCK_N;   // U accumulators: set s at columns [U_COL + 64 s, U_COL + 64 s + r)
  static_assert(2 * BLOCK_N + (LORA ? 2 * 64 : 0) <= TMEM_COLS, "TMEM budget");
  static_assert(BLOCK_N % 64 == 0 && BLOCK_N <= 256, "BLOCK_N");
  // Epilogue warps: TMEM lane quadrant q = warp % 4 is fixed by the hardware.  EW = 4, or 8 (two warps per quadrant,
  // column chunks split even / odd, one staging buffer each) for the math-heavy GELU epilogue.
  static_assert(EW == 4 || (EW == 8 && !LORA), "epilogue warps");
  static_assert(!(AUG && LORA), "AUG (low-rank term as one extra K block fed from a precomputed U) is a dense-kernel mode");
  constexpr int NGRP = EW / 4;
  constexpr int NBUF = EW == 8 ? 1 : 2;
  // Warp roles: epilogue = warps 0..EW-1, TMA producer = warp EW, MMA issuer = warp EW+1.  The SMSP arbiter picks
  // the highest warp id among eligible warps, so the two single-lane control warps — whose few instructions gate
  // the whole tensor pipe — must sit ABOVE the math-heavy epilogue warps they share an SMSP with; with the roles the
  // other way round every extra epilogue instruction (GELU, residual) delayed TMA and MMA issue.
  constexpr int W_TMA = EW, W_MMA = EW + 1;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.num_stages;
  const int stage_bytes = L::stage_bytes(p.r, p.n_sets);
  const int ah_slot = L::ah_slot(p.r);
  uint8_t* stages = smem;
  uint8_t* u_tile = stages + S * stage_bytes;          // [n_sets][16 KB]   (LORA only)
  uint8_t* bp_tile = u_tile + (LORA ? p.n_sets * V2_U_BYTES : 0);
  uint8_t* stg = bp_tile + L::BPH_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg + 4 * 2 * V2_STG_BYTES);
  uint64_t* full = bars;                          // [S]  leader only (count 1 + tx of BOTH CTAs)
  uint64_t* empty = bars + V2_MAX_STAGES;         // [S]  per CTA (count 1, multicast commit)
  uint64_t* tmem_full = bars + 2 * V2_MAX_STAGES; // [2]  per CTA (multicast commit)
  uint64_t* tmem_empty = tmem_full + 2;           // [2]  leader only (count 8: 4 epilogue warps x 2 CTAs)
  uint64_t* u_full = tmem_empty + 2;              //      per CTA (multicast commit)
  uint64_t* u_ready = u_full + 1;                 //      leader only (count 8)
  uint64_t* b_full = u_ready + 1;                 //      leader only (count 1 + tx of both)
  uint64_t* b_empty = b_full + 1;                 //      per CTA (multicast commit)
  uint64_t* u_full2 = b_empty + 1;                //      second U buffer of the U-only pass (double-buffered)
  uint64_t* u_ready2 = u_full2 + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(u_ready2 + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const bool has_lora = (LORA || AUG) && (p.n_adapters > 0) && (p.utt_adapter != nullptr);
  const uint32_t half = p.swap_halves ? (rank ^ 1u) : rank;   // which half of every B operand this CTA supplies

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_y0);
    if (p.n_seg > 1) tma_prefetch_desc(&tm_y1);
    if (p.n_seg > 2) tma_prefetch_desc(&tm_y2);
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
      mbar_init(&tmem_empty[i], 2 * EW);
    }
    mbar_init(u_full, 1);
    mbar_init(u_ready, 8);
    mbar_init(u_full2, 1);
    mbar_init(u_ready2, 8);
    mbar_init(b_full, 1);
    mbar_init(b_empty, 1);
    fence_mbar_init();
  }
  if (warp == W_MMA) {
    tmem_alloc_2sm(tmem_ptr, TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();   // barrier inits + TMEM allocation of BOTH CTAs are visible before any cross-CTA traffic
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int KB = p.k_blocks;
  const int NT = p.n_tiles;
  const int NTS = p.nt_per_seg;
  const int bp_issue_kb = KB > 2 ? KB / 2 : 0;
  // LORA: contiguous range [g0, g1);  dense: g = pair, pair + P, ... (see the scheduling note at the top)
  const long long g0 = LORA ? p.total_steps * pair / p.num_pairs : pair;
  const long long g1 = LORA ? p.total_steps * (pair + 1) / p.num_pairs : p.total_steps;
  const long long g_stride = p.num_pairs;

  // leader-side barrier addresses as seen from this CTA (shared::cluster window of rank 0)
  auto leader_addr = [&](uint64_t* bar) { return mapa_u32(smem_u32(bar), 0); };

  if (warp == W_TMA) {
    // =============================================================== TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t b_uses = 0;
      const uint32_t b_full_leader = leader_addr(b_full);
      // nt = N tile over all segments (-1 = U-only pass); bp_set = LoRA set whose B_k tile this N tile needs (-1 none)
      auto k_loop = [&](int b, int m0, int k, int nt, bool with_a, int bp_set) {
        const bool main = nt >= 0;
        const uint32_t tx_cta =
            V2_X_BYTES + (main ? L::WH_BYTES : 0) + (with_a ? p.n_sets * L::ah_bytes(p.r) : 0);
        for (int kb = 0; kb < KB; ++kb) {
          if (bp_set >= 0 && kb == bp_issue_kb) {
            mbar_wait(b_empty, (b_uses & 1) ^ 1);
            if (leader) mbar_arrive_expect_tx(b_full, 2 * L::BPH_BYTES);
            tma_load_2d_2sm(bp_tile, &tm_b, b_full_leader, 0,
                            (bp_set * p.n_adapters + k) * p.d_out + (nt % NTS) * BLOCK_N + half * (BLOCK_N / 2));
            ++b_uses;
          }
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* st = stages + stage * stage_bytes;
          const uint32_t full_leader = leader_addr(&full[stage]);
          if (leader) mbar_arrive_expect_tx(&full[stage], 2 * tx_cta);
          if (p.x_head_major)
            tma_load_4d_2sm(st, &tm_x, full_leader, 0, m0, kb, b);
          else
            tma_load_3d_2sm(st, &tm_x, full_leader, kb * V2_BLOCK_K, m0, b);
          if (main)
            tma_load_2d_2sm(st + V2_X_BYTES, &tm_w, full_leader, kb * V2_BLOCK_K, nt * BLOCK_N + half * (BLOCK_N / 2));
          if (with_a)
            for (int s = 0; s < p.n_sets; ++s)
              tma_load_2d_2sm(st + V2_X_BYTES + L::WH_BYTES + s * ah_slot, &tm_a, full_leader, kb * V2_BLOCK_K,
                              (s * p.n_adapters + k) * p.r + half * (p.r / 2));
          if (++stage == S) {
            stage = 0;
            phase ^= 1;
          }
        }
      };
      long long g = g0;
      while (g < g1) {
        const int unit = static_cast<int>(g / NT);
        const int nt_first = static_cast<int>(g - static_cast<long long>(unit) * NT);
        const int nt_last = LORA ? static_cast<int>(min(static_cast<long long>(NT), nt_first + (g1 - g))) : nt_first + 1;
        const int b = unit / p.tiles_per_utt;
        const int m0 = (unit - b * p.tiles_per_utt) * 256 + rank * V2_ROWS_PER_CTA;
        int k = has_lora ? p.utt_adapter[b] : -1;
        if (k < 0 || k >= p.n_adapters) k = -1;
        if constexpr (LORA) {
          if (k >= 0 && (nt_first > 0 || p.u_only)) k_loop(b, m0, k, -1, true, -1);   // U-only pass
          if (!p.u_only)
            for (int nt = nt_first; nt < nt_last; ++nt)
              k_loop(b, m0, k, nt, k >= 0 && nt == 0, k >= 0 ? p.seg_set[nt / NTS] : -1);
        } else {
          k_loop(b, m0, -1, nt_first, false, -1);
          if constexpr (AUG) {
            // low-rank term as ONE extra K block: X slot <- the unit's rows of the precomputed U (set's 64 columns),
            // W slot <- B_k's rows of this N tile (same 128-byte-swizzled shapes as a regular stage)
            const int set = k >= 0 ? p.seg_set[nt_first / NTS] : -1;
            if (set >= 0) {
              mbar_wait(&empty[stage], phase ^ 1);
              uint8_t* st = stages + stage * stage_bytes;
              const uint32_t full_leader = leader_addr(&full[stage]);
              if (leader) mbar_arrive_expect_tx(&full[stage], 2 * (V2_X_BYTES + L::WH_BYTES));
              // U plane of this set: rows of r elements; the box is 64 wide, columns >= r are out of bounds -> zero-filled
              // in shared memory WITHOUT being fetched (the extra K block costs r/64 of a regular stage's L2 traffic)
              tma_load_3d_2sm(st, &tm_a, full_leader, 0, m0, set * p.B + b);
              tma_load_2d_2sm(st + V2_X_BYTES, &tm_b, full_leader, 0,
                              (set * p.n_adapters + k) * p.d_out + (nt_first % NTS) * BLOCK_N + half * (BLOCK_N / 2));
              if (++stage == S) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
        g += LORA ? static_cast<long long>(nt_last - nt_first) : g_stride;
      }
    }
    __syncwarp();
  } else if (warp == W_MMA) {
    // =============================================================== MMA issuer (leader CTA, single thread)
    if (leader && lane == 0) {
      const uint32_t idesc_main = umma_idesc_bf16(256, BLOCK_N);
      const uint32_t idesc_u = umma_idesc_bf16(256, p.r);
      const uint32_t u_desc_base = smem_u32(u_tile);
      const uint32_t bp_desc_base = smem_u32(bp_tile);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tile_iter = 0, lora_units = 0, b_uses = 0;
      uint32_t u_col = U_COL;   // TMEM column of U set 0 (the U-only pass alternates between two buffers)
      auto k_loop = [&](uint32_t acc, bool main, bool with_a) {
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t st = smem_u32(stages + stage * stage_bytes);
          const uint64_t xd = umma_desc_sw128(st);
          if (main) {
            const uint64_t wd = umma_desc_sw128(st + V2_X_BYTES);
#pragma unroll
            for (int kk = 0; kk < V2_BLOCK_K / 16; ++kk)
              umma_bf16_2sm(acc, xd + 2 * kk, wd + 2 * kk, idesc_main, (kb | kk) != 0);
          }
          if (with_a) {
            for (int s = 0; s < p.n_sets; ++s) {
              const uint64_t ad = umma_desc_sw128(st + V2_X_BYTES + L::WH_BYTES + s * ah_slot);
#pragma unroll
              for (int kk = 0; kk < V2_BLOCK_K / 16; ++kk)
                umma_bf16_2sm(tmem_base + u_col + 64 * s, xd + 2 * kk, ad + 2 * kk, idesc_u, (kb | kk) != 0);
            }
          }
          umma_commit_2sm(&empty[stage], 0b11);
          if (++stage == S) {
            stage = 0;
            phase ^= 1;
          }
        }
      };
      auto publish_u = [&]() {   // U accumulator complete -> both CTAs' epilogue warps convert their rows
        umma_commit_2sm(u_full, 0b11);
        mbar_wait(u_ready, lora_units & 1);
        tc_fence_after();
        ++lora_units;
      };
      long long g = g0;
      while (g < g1) {
        const int unit = static_cast<int>(g / NT);
        const int nt_first = static_cast<int>(g - static_cast<long long>(unit) * NT);
        const int nt_last = LORA ? static_cast<int>(min(static_cast<long long>(NT), nt_first + (g1 - g))) : nt_first + 1;
        const int b = unit / p.tiles_per_utt;
        int k = has_lora ? p.utt_adapter[b] : -1;
        if (k < 0 || k >= p.n_adapters) k = -1;
        if (LORA && p.u_only) {
          // U-only pass: two U buffers in TMEM (the accumulator columns are free), so the MMAs of unit j+1 run while the
          // epilogue warps convert and store unit j; a buffer is re-used once its previous conversion has arrived
          if (k >= 0) {
            const uint32_t ub = lora_units & 1, nb = lora_units >> 1;
            if (nb >= 1) {
              mbar_wait(ub ? u_ready2 : u_ready, (nb - 1) & 1);
              tc_fence_after();
            }
            u_col = ub ? 0u : static_cast<uint32_t>(U_COL);
            k_loop(0, false, true);
            umma_commit_2sm(ub ? u_full2 : u_full, 0b11);
            ++lora_units;
          }
          g += nt_last - nt_first;
          continue;
        }
        if (LORA && k >= 0 && nt_first > 0) {
          k_loop(0, false, true);
          publish_u();
        }
        for (int nt = nt_first; nt < nt_last; ++nt, ++tile_iter) {
          const uint32_t buf = tile_iter & 1;
          const uint32_t acc = tmem_base + buf * BLOCK_N;
          mbar_wait(&tmem_empty[buf], ((tile_iter >> 1) & 1) ^ 1);
          tc_fence_after();
          k_loop(acc, true, LORA && k >= 0 && nt == 0);
          if (LORA && k >= 0 && nt == 0) publish_u();
          const int set = k >= 0 ? p.seg_set[nt / NTS] : -1;
          if constexpr (AUG) {
            if (set >= 0) {   // the extra K block: acc += U[:, set] · B_k[n tile]ᵀ, r/16 MMAs
              mbar_wait(&full[stage], phase);
              tc_fence_after();
              const uint32_t st = smem_u32(stages + stage * stage_bytes);
              const uint64_t ud = umma_desc_sw128(st);
              const uint64_t bd = umma_desc_sw128(st + V2_X_BYTES);
              const int ksteps = p.r >> 4;
              for (int kk = 0; kk < ksteps; ++kk) umma_bf16_2sm(acc, ud + 2 * kk, bd + 2 * kk, idesc_main, 1u);
              umma_commit_2sm(&empty[stage], 0b11);
              if (++stage == S) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
          if (LORA && set >= 0) {
            mbar_wait(b_full, b_uses & 1);
            tc_fence_after();
            const uint64_t ud = umma_desc_sw128(u_desc_base + set * V2_U_BYTES);
            const uint64_t bd = umma_desc_sw128(bp_desc_base);
            const int ksteps = p.r >> 4;
            for (int kk = 0; kk < ksteps; ++kk) umma_bf16_2sm(acc, ud + 2 * kk, bd + 2 * kk, idesc_main, 1u);
            umma_commit_2sm(b_empty, 0b11);
            ++b_uses;
          }
          umma_commit_2sm(&tmem_full[buf], 0b11);
        }
        g += LORA ? static_cast<long long>(nt_last - nt_first) : g_stride;
      }
    }
    __syncwarp();
  } else {
    // =============================================================== epilogue warps (0..EW-1), both CTAs
    const int q = warp & 3;
    const int grp = warp >> 2;   // 0 with EW = 4; 0 / 1 with EW = 8
    const int row = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    uint8_t* my_stg = stg + warp * (NBUF * V2_STG_BYTES);
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    uint32_t tile_iter = 0, lora_units = 0, stg_idx = 0;
    const uint32_t u_ready_leader = leader_addr(u_ready);
    const uint32_t u_ready2_leader = leader_addr(u_ready2);
    const uint32_t tmem_empty_leader[2] = {leader_addr(&tmem_empty[0]), leader_addr(&tmem_empty[1])};
    long long g = g0;
    while (g < g1) {
      const int unit = static_cast<int>(g / NT);
      const int nt_first = static_cast<int>(g - static_cast<long long>(unit) * NT);
      const int nt_last = LORA ? static_cast<int>(min(static_cast<long long>(NT), nt_first + (g1 - g))) : nt_first + 1;
      const int b = unit / p.tiles_per_utt;
      const int m0 = (unit - b * p.tiles_per_utt) * 256 + rank * V2_ROWS_PER_CTA;
      int k = has_lora ? p.utt_adapter[b] : -1;
      if (k < 0 || k >= p.n_adapters) k = -1;
      if (LORA && k >= 0) {
        // ---- U: TMEM fp32 -> scale -> bf16 -> swizzled smem A-operand tile of THIS CTA (+ optional global save)
        // (U-only pass: alternating TMEM buffers / barriers, see the MMA issuer)
        const bool second = p.u_only && (lora_units & 1);
        const uint32_t u_col = second ? 0u : static_cast<uint32_t>(U_COL);
        mbar_wait(second ? u_full2 : u_full, p.u_only ? ((lora_units >> 1) & 1) : (lora_units & 1));
        tc_fence_after();
        // the unit's rows are saved exactly once: by the range that owns its N-tile 0 (single-set calls only)
        const bool save = (p.u_out != nullptr) && (nt_first == 0) && (m0 + row < p.T);
        const size_t u_row_ld = p.u_ld > 0 ? static_cast<size_t>(p.u_ld) : static_cast<size_t>(p.r);
        uint4* u_dst = save ? reinterpret_cast<uint4*>(p.u_out + (static_cast<size_t>(b) * p.T + m0 + row) * u_row_ld)
                            : nullptr;
        for (int s = 0; s < p.n_sets; ++s) {
          const uint32_t u_row = smem_u32(u_tile) + s * V2_U_BYTES + row * 128;
          for (int j = 0; j < (p.r >> 4); ++j) {
            uint32_t v[16];
            tmem_ld_32x16(tmem_base + lane_addr + u_col + 64 * s + j * 16, v);
            tmem_ld_wait();
            uint32_t pk[8];
            const float sc = p.u_w ? p.scale * p.u_w[static_cast<size_t>(b) * p.u_w_ld + (j * 16) / p.u_w_group] : p.scale;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              pk[i] = pack_bf16x2(__uint_as_float(v[2 * i]) * sc, __uint_as_float(v[2 * i + 1]) * sc);
            if (!p.u_only) {   // the smem operand tile is only needed when this kernel also runs the output tiles
              st_shared_v4(u_row + (((2 * j) ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
              st_shared_v4(u_row + (((2 * j + 1) ^ sw) << 4), pk[4], pk[5], pk[6], pk[7]);
            }
            if (save && (s == 0 || p.u_ld > 0)) {   // u_ld layout: set s is plane s of [n_sets][B, T, r]
              uint4* ud = u_dst + static_cast<size_t>(s) * p.B * p.T * (u_row_ld >> 3);
              ud[2 * j] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              ud[2 * j + 1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
        }
        tc_fence_before();
        asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy smem writes -> async proxy (peer-issued UMMA)
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(second ? u_ready2_leader : u_ready_leader);
        ++lora_units;
      }
      if (LORA && p.u_only) {
        g += nt_last - nt_first;
        continue;
      }
      for (int nt = nt_first; nt < nt_last; ++nt, ++tile_iter) {
        const uint32_t buf = tile_iter & 1;
        const int n0 = nt * BLOCK_N;                 // column in the concatenated N space (W_cat / bias_cat rows)
        const int seg = nt / NTS;
        const int n0_seg = (nt - seg * NTS) * BLOCK_N;   // column inside this segment's output tensor
        const CUtensorMap* ty = seg == 0 ? &tm_y0 : (seg == 1 ? &tm_y1 : &tm_y2);
        float oscale = 1.0f;
        if constexpr ((EPI & EPI_SCALE) != 0) oscale = p.seg_scale[seg];
        const bool rows_live = (m0 + q * 32) < p.T;
        // residual: the warp's 32 x 64 block of a chunk is fetched COALESCED (request j: lane i reads 16 B of row
        // 4j + i/8 — four full 128-byte lines per request instead of 32 partial ones), parked in registers while the
        // accumulator is awaited (first chunk) / while the previous chunk is converted, then transposed to the
        // thread-per-row layout through the swizzled staging buffer that the output is about to overwrite anyway.
        constexpr int NC = BLOCK_N / 64;
        uint4 rs[8];
        const __nv_bfloat16* res_blk = nullptr;   // row m0 + q*32, column n0_seg of this warp's block
        const int r_sub = lane >> 3, r_chk = lane & 7;
        auto load_res = [&](uint4 (&dst)[8], int c) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int rr = 4 * j + r_sub;
            dst[j] = (res_blk != nullptr && m0 + q * 32 + rr < p.T)
                         ? __ldg(reinterpret_cast<const uint4*>(res_blk + static_cast<size_t>(rr) * p.ldr + c * 64) + r_chk)
                         : make_uint4(0u, 0u, 0u, 0u);
          }
        };
        if constexpr ((EPI & EPI_RES) != 0) {
          if (grp < NC)
            res_blk = p.residual + static_cast<size_t>(b) * p.res_bs + static_cast<size_t>(m0 + q * 32) * p.ldr + n0_seg;
          load_res(rs, grp);
        }
        uint4 rn[8];
        if constexpr ((EPI & EPI_RES) != 0 && EW == 8) {
          if (grp + NGRP < NC) load_res(rn, grp + NGRP);
        }
        mbar_wait(&tmem_full[buf], (tile_iter >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = grp; c < NC; c += NGRP) {
          uint32_t v0[32], v1[32];
          const uint32_t taddr = tmem_base + lane_addr + buf * BLOCK_N + c * 64;
          tmem_ld_32x32(taddr, v0);
          tmem_ld_32x32(taddr + 32, v1);
          if constexpr ((EPI & EPI_RES) != 0 && EW == 4) {
            if (c + NGRP < NC) load_res(rn, c + NGRP);
          }
          const uint4* bias4 = p.bias ? reinterpret_cast<const uint4*>(p.bias + n0 + c * 64) : nullptr;
          uint4 bb[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) bb[j] = bias4 ? __ldg(bias4 + j) : make_uint4(0u, 0u, 0u, 0u);
          // the staging buffer is reused every NBUF chunks: its previous TMA store must have finished READING it
          if (lane == 0) tma_store_wait_read<NBUF - 1>();
          __syncwarp();
          uint8_t* sbuf = my_stg + (stg_idx % NBUF) * V2_STG_BYTES;
          const uint32_t srow = smem_u32(sbuf) + lane * 128;
          if constexpr ((EPI & EPI_RES) != 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {   // row 4j + r_sub, 16-byte chunk r_chk, 128B-swizzled like the output
              const uint32_t rr = static_cast<uint32_t>(4 * j + r_sub);
              st_shared_v4(smem_u32(sbuf) + rr * 128 + ((static_cast<uint32_t>(r_chk) ^ (rr & 7u)) << 4), rs[j].x, rs[j].y,
                           rs[j].z, rs[j].w);
            }
            __syncwarp();
          }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t bw[4] = {bb[j].x, bb[j].y, bb[j].z, bb[j].w};
            uint32_t rw[4] = {0u, 0u, 0u, 0u};
            const uint32_t saddr = srow + ((static_cast<uint32_t>(j) ^ sw) << 4);
            if constexpr ((EPI & EPI_RES) != 0) ld_shared_v4(saddr, rw[0], rw[1], rw[2], rw[3]);
            uint32_t pk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int e = 8 * j + 2 * i;
              // bias add, activation, scale and residual add on the packed fp32 pipe (the epilogue is issue-bound)
              float2 av = fadd2(make_float2(__uint_as_float(e < 32 ? v0[e & 31] : v1[e & 31]),
                                            __uint_as_float(e < 32 ? v0[(e + 1) & 31] : v1[(e + 1) & 31])),
                                make_float2(__uint_as_float(bw[i] << 16), __uint_as_float(bw[i] & 0xFFFF0000u)));
              if constexpr ((EPI & EPI_GELU) != 0) av = gelu_erf2(av);
              if constexpr ((EPI & EPI_SCALE) != 0) av = fmul2(av, make_float2(oscale, oscale));
              if constexpr ((EPI & EPI_DGELU) != 0)   // GELU backward: the "residual" operand is the pre-activation
                av = fmul2(av, gelu_grad2(make_float2(__uint_as_float(rw[i] << 16), __uint_as_float(rw[i] & 0xFFFF0000u))));
              else if constexpr ((EPI & EPI_RES) != 0)
                av = fadd2(av, make_float2(__uint_as_float(rw[i] << 16), __uint_as_float(rw[i] & 0xFFFF0000u)));
              const float a0 = av.x, a1 = av.y;
              pk[i] = pack_bf16x2(a0, a1);
            }
            st_shared_v4(saddr, pk[0], pk[1], pk[2], pk[3]);
          }
          if constexpr ((EPI & EPI_RES) != 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) rs[j] = rn[j];
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && rows_live) {
            if (p.y_head_major)
              tma_store_4d(ty, sbuf, 0, m0 + q * 32, (n0_seg >> 6) + c, b);
            else
              tma_store_3d(ty, sbuf, n0_seg + c * 64, m0 + q * 32, b);
            tma_store_commit();
          }
          ++stg_idx;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(tmem_empty_leader[buf]);
      }
      g += LORA ? static_cast<long long>(nt_last - nt_first) : g_stride;
    }
    if (lane == 0) tma_store_wait_all<0>();
    __syncwarp();
  }

  tc_fence_before();
  cluster_sync_all();   // neither CTA may free TMEM / exit while its peer still reads its smem or signals its barriers
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ host
template <int BLOCK_N, bool LORA, int EPI, int EW = 4, bool AUG = false>
static int k1v2_launch(const K1Args& a, cudaStream_t stream) {
  using L = V2Smem<BLOCK_N, LORA>;
  const DeviceInfo& dev = device_info();
  const bool has_lora = LORA || AUG;
  const int n_seg = a.n_seg > 0 ? a.n_seg : 1;
  const int n_sets = (has_lora && a.n_sets > 0) ? a.n_sets : 1;

  K1V2Params p{};
  p.B = a.B; p.T = a.T; p.d_in = a.d_in; p.d_out = a.d_out; p.r = has_lora ? a.r : 16;
  p.tiles_per_utt = (a.T + 255) / 256;
  p.nt_per_seg = (a.d_out + BLOCK_N - 1) / BLOCK_N;   // ragged last tile: W rows past d_out are zero-filled by TMA
  p.n_seg = n_seg;
  p.n_tiles = n_seg * p.nt_per_seg;
  p.n_sets = n_sets;
  for (int s = 0; s < 3; ++s) {
    p.seg_set[s] = a.n_seg > 0 ? a.seg_set[s] : 0;
    p.seg_scale[s] = a.n_seg > 0 ? a.seg_scale[s] : 1.0f;
    if (p.seg_set[s] >= n_sets) return fail(SAR_EINVAL, "k1v2: segment refers to a LoRA set that does not exist");
  }
  p.x_head_major = a.x_head_major;
  p.y_head_major = a.y_head_major;
  p.k_blocks = (a.d_in + V2_BLOCK_K - 1) / V2_BLOCK_K;   // ragged K: columns past d_in are zero-filled by TMA
  p.n_adapters = has_lora ? a.n_adapters : 0;
  p.u_only = (LORA && a.u_only) ? 1 : 0;
  p.u_ld = a.u_ld;
  if (p.u_only) {           // one "step" per unit: the roles run the U-only pass and skip the output tiles
    p.n_tiles = 1;
    p.nt_per_seg = 1;
  }
  p.total_steps = static_cast<long long>(a.B) * p.tiles_per_utt * p.n_tiles;
  p.scale = a.scale;
  p.u_w = has_lora ? a.u_w : nullptr;
  p.u_w_ld = a.u_w_ld;
  p.u_w_group = a.u_w_group > 0 ? a.u_w_group : 16;
  p.swap_halves = a.swap_halves;
  p.utt_adapter = has_lora ? a.utt_adapter : nullptr;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(a.bias);
  p.residual = reinterpret_cast<const __nv_bfloat16*>(a.residual);
  p.ldr = a.ldr > 0 ? a.ldr : a.d_out;
  p.res_bs = a.res_broadcast ? 0 : (a.res_batch_stride > 0 ? a.res_batch_stride : static_cast<long long>(a.T) * p.ldr);
  const uint64_t ldx = a.ldx > 0 ? a.ldx : a.d_in;
  const uint64_t x_bs = a.x_batch_stride > 0 ? a.x_batch_stride : static_cast<uint64_t>(a.T) * ldx;
  const uint64_t ldy = a.ldy > 0 ? a.ldy : a.d_out;
  const uint64_t y_bs = a.y_batch_stride > 0 ? a.y_batch_stride : static_cast<uint64_t>(a.T) * ldy;
  p.act = a.act;
  p.u_out = (LORA && (n_sets == 1 || a.u_ld > 0)) ? reinterpret_cast<__nv_bfloat16*>(a.u_out) : nullptr;
  if (p.u_only && (!p.u_out || a.u_ld != p.r)) return fail(SAR_EINVAL, "k1v2: U-only pass needs u_out [n_sets][B,T,r]");

  const int stage_bytes = L::stage_bytes(p.r, n_sets);
  const int budget = dev.max_smem_optin - 1024 - L::fixed_bytes(n_sets);
  int S = budget / stage_bytes;
  if (S > V2_MAX_STAGES) S = V2_MAX_STAGES;
  if (S < 2) return fail(SAR_EINVAL, "k1v2: shared memory budget too small for this shape");
  p.num_stages = S;
  const int smem_bytes = 1024 + S * stage_bytes + L::fixed_bytes(n_sets);

  int pairs = dev.num_sms / 2;
  if (p.total_steps < pairs) pairs = static_cast<int>(p.total_steps);
  if (a.grid_override > 0 && a.grid_override / 2 >= 1 && a.grid_override / 2 < pairs) pairs = a.grid_override / 2;
  p.num_pairs = pairs;

  CUtensorMap tm_x, tm_w, tm_a, tm_b, tm_y[3];
  memset(&tm_a, 0, sizeof(tm_a));
  memset(&tm_b, 0, sizeof(tm_b));
  memset(tm_y, 0, sizeof(tm_y));
  int rc;
  if (a.x_head_major) {
    const uint64_t h = a.d_in / 64;
    const uint64_t dims[4] = {64, (uint64_t)a.T, h, (uint64_t)a.B};
    const uint64_t strides[3] = {128, (uint64_t)a.T * 128, h * a.T * 128};
    const uint32_t box[4] = {64, V2_ROWS_PER_CTA, 1, 1};
    if ((rc = make_tmap_bf16(&tm_x, a.x, 4, dims, strides, box))) return rc;
  } else {
    const uint64_t dims[3] = {(uint64_t)a.d_in, (uint64_t)a.T, (uint64_t)a.B};
    const uint64_t strides[2] = {ldx * 2, x_bs * 2};   // rows may overlap (ldx < d_in): conv-as-GEMM windows
    const uint32_t box[3] = {V2_BLOCK_K, V2_ROWS_PER_CTA, 1};
    if ((rc = make_tmap_bf16(&tm_x, a.x, 3, dims, strides, box))) return rc;
  }
  for (int s = 0; s < n_seg; ++s) {
    void* yp = a.n_seg > 0 ? a.y_seg[s] : a.y;
    if (!yp) return fail(SAR_EINVAL, "k1v2: null output segment");
    if (a.y_head_major) {
      const uint64_t h = a.d_out / 64;
      const uint64_t dims[4] = {64, (uint64_t)a.T, h, (uint64_t)a.B};
      const uint64_t strides[3] = {128, (uint64_t)a.T * 128, h * a.T * 128};
      const uint32_t box[4] = {64, 32, 1, 1};
      if ((rc = make_tmap_bf16(&tm_y[s], yp, 4, dims, strides, box))) return rc;
    } else {
      const uint64_t dims[3] = {(uint64_t)a.d_out, (uint64_t)a.T, (uint64_t)a.B};
      const uint64_t strides[2] = {ldy * 2, y_bs * 2};
      const uint32_t box[3] = {64, 32, 1};
      if ((rc = make_tmap_bf16(&tm_y[s], yp, 3, dims, strides, box))) return rc;
    }
  }
  {
    const uint64_t dims[2] = {(uint64_t)a.d_in, (uint64_t)n_seg * a.d_out};
    const uint64_t strides[1] = {(uint64_t)a.d_in * 2};
    const uint32_t box[2] = {V2_BLOCK_K, BLOCK_N / 2};
    if ((rc = make_tmap_bf16(&tm_w, a.W, 2, dims, strides, box))) return rc;
  }
  if (has_lora) {
    if constexpr (AUG) {   // tm_a = the precomputed U, [n_sets][B, T, r]: inner extent r < the 64-wide box
      if (!a.u_out || a.u_ld != a.r) return fail(SAR_EINVAL, "k1v2: AUG needs u [n_sets][B,T,r]");
      const uint64_t dims[3] = {(uint64_t)a.r, (uint64_t)a.T, (uint64_t)n_sets * a.B};
      const uint64_t strides[2] = {(uint64_t)a.r * 2, (uint64_t)a.T * a.r * 2};
      const uint32_t box[3] = {64, V2_ROWS_PER_CTA, 1};
      if ((rc = make_tmap_bf16(&tm_a, a.u_out, 3, dims, strides, box, /*l2_promotion_bytes=*/0))) return rc;
    } else {
      const uint64_t dims[2] = {(uint64_t)a.d_in, (uint64_t)n_sets * a.n_adapters * a.r};
      const uint64_t strides[1] = {(uint64_t)a.d_in * 2};
      const uint32_t box[2] = {V2_BLOCK_K, (uint32_t)a.r / 2};
      if ((rc = make_tmap_bf16(&tm_a, a.A_stack, 2, dims, strides, box))) return rc;
    }
    {
      // lora_B is stored rank-padded to 64 (one swizzle layout for every rank); the map's inner extent is the true
      // rank, so the padding columns are zero-filled on chip instead of being read
      const uint64_t dims[2] = {(uint64_t)a.r, (uint64_t)n_sets * a.n_adapters * a.d_out};
      const uint64_t strides[1] = {(uint64_t)SAR_RPAD * 2};
      const uint32_t box[2] = {64, BLOCK_N / 2};
      if ((rc = make_tmap_bf16(&tm_b, a.Bp_stack, 2, dims, strides, box, /*l2_promotion_bytes=*/AUG ? 0 : 256))) return rc;
    }
  }

  auto kern = k1v2_kernel<BLOCK_N, LORA, EPI, EW, AUG>;
  // the opt-in shared-memory limit is a per-device attribute of the kernel: remember it per device (and per thread:
  // no lock needed, the call is idempotent)
  static thread_local int smem_set_dev[64] = {};
  int& smem_set = smem_set_dev[dev.device & 63];
  if (smem_set < smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dev.max_smem_optin);
    if (e != cudaSuccess) return fail_cuda(e, "k1v2: cudaFuncSetAttribute");
    smem_set = dev.max_smem_optin;
  }
  kern<<<2 * pairs, 32 * (EW + 2), smem_bytes, stream>>>(tm_x, tm_w, tm_a, tm_b, tm_y[0], tm_y[1], tm_y[2], p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "k1v2: launch");
  return SAR_OK;
}

template <int BLOCK_N, bool LORA>
static int k1v2_dispatch_epi(const K1Args& a, int epi, cudaStream_t stream) {
  switch (epi) {
    case 0: return k1v2_launch<BLOCK_N, LORA, 0>(a, stream);
    case EPI_SCALE: return k1v2_launch<BLOCK_N, LORA, EPI_SCALE>(a, stream);
    case EPI_RES:
      if constexpr (!LORA) {
        if (a.d_in <= 1536) return k1v2_launch<BLOCK_N, false, EPI_RES, 8>(a, stream);
        return k1v2_launch<BLOCK_N, false, EPI_RES, 4>(a, stream);
      }
      break;
    case EPI_GELU:
      if constexpr (!LORA) return k1v2_launch<BLOCK_N, false, EPI_GELU, 8>(a, stream);
      break;
    case EPI_GELU | EPI_RES:
      if constexpr (!LORA) return k1v2_launch<BLOCK_N, false, EPI_GELU | EPI_RES, 8>(a, stream);
      break;
    case EPI_DGELU | EPI_RES:
      if constexpr (!LORA) return k1v2_launch<BLOCK_N, false, EPI_DGELU | EPI_RES, 8>(a, stream);
      break;
    default: break;
  }
  return fail(SAR_EINVAL, "k1v2: unsupported epilogue combination (residual / GELU are dense-only)");
}

int k1v2_qv_lora_fwd(const K1Args& a, int block_n, cudaStream_t stream) {
  const bool lora = a.n_adapters > 0 && a.utt_adapter != nullptr && a.A_stack != nullptr && a.Bp_stack != nullptr;
  const int n_seg = a.n_seg > 0 ? a.n_seg : 1;
  int epi = 0;
  if (a.residual) epi |= EPI_RES;
  if (a.act == SAR_ACT_GELU) epi |= EPI_GELU;
  if (a.act == SAR_ACT_GELU_BWD) {
    if (!a.residual) return fail(SAR_EINVAL, "k1v2: SAR_ACT_GELU_BWD reads the pre-activation through the residual operand");
    epi |= EPI_DGELU;
  }
  if (a.n_seg > 0)
    for (int s = 0; s < n_seg; ++s)
      if (a.seg_scale[s] != 1.0f) epi |= EPI_SCALE;
  if (a.residual && (n_seg != 1 || a.y_head_major))
    return fail(SAR_EINVAL, "k1v2: residual needs one row-major output segment");
  if (lora && a.u_ws != nullptr && !a.u_only) {
    // Split path (chosen by the caller passing a workspace): launch 1 = U-only pass of the LoRA kernel, U for every set
    // to HBM ([n_sets][B,T,r] bf16: r/d_in of x's bytes per set); launch 2 = the DENSE kernel with the low-rank term as
    // one extra K block per tile (AUG).  Measured against the single-launch kernel that keeps U in shared memory
    // (which is limited to 128/192-wide tiles by the TMEM budget and stalls once per unit on the U hand-over):
    // whisper-large-v3 q|k|v r64: 1460 us -> see DESIGN.md §4.
    if (a.residual || a.act != SAR_ACT_NONE) return fail(SAR_EINVAL, "k1v2: split LoRA path has no residual / activation");
    const int n_sets = a.n_sets > 0 ? a.n_sets : 1;
    if (a.u_phase != 2) {
      K1Args u = a;
      u.u_only = 1; u.u_out = a.u_ws; u.u_ld = a.r; u.x_head_major = a.x_head_major;
      if (a.u_phase == 1)   // U only: the output maps are built but never used; point them at the workspace
        for (int s = 0; s < 3; ++s) u.y_seg[s] = a.u_ws;
      int rc = k1v2_launch<128, true, 0, 4>(u, stream);
      if (rc || a.u_phase == 1) return rc;
    }
    K1Args m = a;
    m.u_out = a.u_ws; m.u_ld = a.r;
    int bn = a.block_n_override;
    if (bn != 128 && bn != 192 && bn != 256) bn = (a.d_out % 256 == 0) ? 256 : ((a.d_out % 192 == 0) ? 192 : 128);
    if (a.d_out % bn) return fail(SAR_EINVAL, "k1v2: d_out not divisible by BLOCK_N");
    const bool sc = (epi & EPI_SCALE) != 0;
    switch (bn) {
      case 128: return sc ? k1v2_launch<128, false, EPI_SCALE, 4, true>(m, stream) : k1v2_launch<128, false, 0, 4, true>(m, stream);
      case 192: return sc ? k1v2_launch<192, false, EPI_SCALE, 4, true>(m, stream) : k1v2_launch<192, false, 0, 4, true>(m, stream);
      default: return sc ? k1v2_launch<256, false, EPI_SCALE, 4, true>(m, stream) : k1v2_launch<256, false, 0, 4, true>(m, stream);
    }
  }
  if (lora) {
    if (block_n == 256) block_n = (a.d_out % 192 == 0) ? 192 : 128;
    switch (block_n) {
      case 128: return k1v2_dispatch_epi<128, true>(a, epi, stream);
      case 192: return k1v2_dispatch_epi<192, true>(a, epi, stream);
      default: return fail(SAR_EINVAL, "k1v2: unsupported BLOCK_N for the LoRA kernel (128 or 192)");
    }
  }
  switch (block_n) {
    case 128: return k1v2_dispatch_epi<128, false>(a, epi, stream);
    case 192: return k1v2_dispatch_epi<192, false>(a, epi, stream);
    case 256: return k1v2_dispatch_epi<256, false>(a, epi, stream);
    default: return fail(SAR_EINVAL, "k1v2: unsupported BLOCK_N");
  }
}

// Fused attention-projection entry (sar_attn_proj_fwd): validates and always runs on the pair kernel.
int attn_proj_fwd(const K1Args& a, cudaStream_t stream) {
  if (!a.x || !a.W) return fail(SAR_EINVAL, "attn_proj: null x/W");
  if (a.B <= 0 || a.T <= 0) return fail(SAR_EINVAL, "attn_proj: B and T must be positive");
  if (a.n_seg < 1 || a.n_seg > 3) return fail(SAR_EINVAL, "attn_proj: n_seg must be 1, 2 or 3");
  if (a.d_in <= 0 || a.d_out <= 0) return fail(SAR_EINVAL, "attn_proj: d_in and d_out must be positive");
  const bool plain = a.n_seg == 1 && !a.x_head_major && !a.y_head_major &&
                     !(a.n_adapters > 0 && a.utt_adapter && a.A_stack && a.Bp_stack);
  if (plain) {
    // single dense segment: ragged K (zero-filled by TMA) and a ragged last N tile are allowed
    const long long ldx = a.ldx > 0 ? a.ldx : a.d_in, ldy = a.ldy > 0 ? a.ldy : a.d_out;
    if (a.d_in % 8 || ldx % 8 || ldy % 8 || (a.x_batch_stride % 8) || (a.y_batch_stride % 8))
      return fail(SAR_EINVAL, "dense: d_in and every stride must be a multiple of 8 elements (16 bytes)");
    if (a.d_out % 64 && (a.bias || a.residual))
      return fail(SAR_EINVAL, "dense: bias / residual need d_out to be a multiple of 64");
    if (a.residual && ((a.ldr > 0 ? a.ldr : a.d_out) % 8 || (a.res_batch_stride > 0 && a.res_batch_stride % 8)))
      return fail(SAR_EINVAL, "dense: residual strides must be multiples of 8 elements");
  } else {
    if (a.d_in % 64 || a.d_out % 128)
      return fail(SAR_EINVAL, "attn_proj: d_in must be a multiple of 64 and d_out of 128");
    if (a.ldx || a.ldy || a.x_batch_stride || a.y_batch_stride)
      return fail(SAR_EINVAL, "attn_proj: custom strides are supported for single dense row-major segments only");
  }
  const bool lora = a.n_adapters > 0 && a.utt_adapter && a.A_stack && a.Bp_stack;
  if (lora && (a.r % 16 || a.r < 16 || a.r > 64)) return fail(SAR_EINVAL, "attn_proj: r must be 16, 32, 48 or 64");
  if (lora && (a.n_sets < 1 || a.n_sets > 2)) return fail(SAR_EINVAL, "attn_proj: n_sets must be 1 or 2");
  uintptr_t al = reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.W) |
                 reinterpret_cast<uintptr_t>(a.A_stack) | reinterpret_cast<uintptr_t>(a.Bp_stack) |
                 reinterpret_cast<uintptr_t>(a.bias);
  for (int s = 0; s < a.n_seg; ++s) al |= reinterpret_cast<uintptr_t>(a.y_seg[s]);
  if (al & 15) return fail(SAR_EINVAL, "attn_proj: pointers must be 16-byte aligned");
  // <= 128 rows, dense, row-major: weight-streaming problem -> narrow single-CTA tiles on every SM (skinny_fwd.cu)
  if (a.B == 1 && !a.block_n_override && !a.grid_override && !a.x_batch_stride && !a.y_batch_stride &&
      a.act != SAR_ACT_GELU_BWD && skinny_applicable(a, a.T))
    return skinny_fwd(a, a.T, stream);
  int bn = a.block_n_override;
  if (!bn) {
    bn = (!lora && (a.d_out % 256 == 0 || (plain && a.d_out > 2048))) ? 256 : ((a.d_out % 192 == 0) ? 192 : 128);
    // Few rows (decode steps: one 256-row unit): the call is a weight-streaming problem, and with 256-wide tiles only
    // d_out/256 pairs would pull the weights.  Narrow tiles put more SMs on the stream.
    const long long units = static_cast<long long>(a.B) * ((a.T + 255) / 256);
    const int n_seg = a.n_seg > 0 ? a.n_seg : 1;
    const int pairs = device_info().num_sms / 2;
    if (!lora && a.d_out % 128 == 0 && units * n_seg * ((a.d_out + bn - 1) / bn) < pairs) bn = 128;
  }
  if (a.d_out % bn && !plain) return fail(SAR_EINVAL, "attn_proj: d_out not divisible by BLOCK_N");
  return k1v2_qv_lora_fwd(a, bn, stream);
}

}  // namespace sar
