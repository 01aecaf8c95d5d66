// sar_internal.h — shared host-side declarations of libsar (not part of the public ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/sar.h"

namespace sar {

struct DeviceInfo {
  int ok;  // 1 = sm_100, 0 = other GPU, <0 = no device
  int device;
  int cc_major, cc_minor;
  int num_sms;
  int max_smem_optin;
  int l2_bytes;
};

// Cached per-process properties of the current device (queried once per device, thread-safe).
const DeviceInfo& device_info();

// Record an error message for sar_last_error() and return `code`.
int fail(int code, const char* msg);
int fail_cuda(cudaError_t e, const char* where);
// Entry-point guard: returns SAR_OK only on an sm_100 device.
int require_sm100();

// Encode a bf16 tiled tensor map with 128-byte swizzle.  dims/box are innermost-first; strides (bytes) has rank-1
// entries for dims 1..rank-1.  Uses cuTensorMapEncodeTiled resolved through cudaGetDriverEntryPoint, so libsar has
// no link-time dependency on libcuda.
// l2_promotion_bytes: 256 (default, streaming operands), 128, or 0 = none (narrow boxes whose rows are 32-128 bytes).
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int l2_promotion_bytes = 256);

struct K1Args {
  const void* x;
  const void* W;
  const void* bias;
  const void* A_stack;
  const void* Bp_stack;
  const int32_t* utt_adapter;
  void* y;
  void* u_out;
  int B, T, d_in, d_out, r, n_adapters;
  float scale;
  int block_n_override;  // 0 = auto
  int grid_override;     // 0 = auto
  int kernel_override;   // 0 = auto, 1 = single-CTA kernel, 2 = CTA-pair (cta_group::2) kernel
  int swap_halves;       // debug switch of the pair kernel's B-operand half assignment
  // ---- fused attention-projection extensions (pair kernel only); n_seg == 0 means "plain single projection"
  int n_seg;             // output segments sharing x (W, bias concatenated along d_out): q|k|v = 3, k|v = 2
  int n_sets;            // LoRA sets stacked in A_stack / Bp_stack ([n_sets*n_adapters, ...])
  int seg_set[3];        // LoRA set per segment (-1 = none)
  float seg_scale[3];    // epilogue scale per segment
  void* y_seg[3];        // one output tensor per segment
  int x_head_major;      // x is [B, d_in/64, T, 64]
  int y_head_major;      // every y is [B, d_out/64, T, 64]
  const void* residual;  // bf16 [B, T, d_out] added in the epilogue (single row-major segment only)
  int act;               // SAR_ACT_*
  // ---- strided dense-layer extensions (single row-major segment, no LoRA); 0 = contiguous default
  long long ldx, x_batch_stride;     // x row / batch stride in elements (rows may overlap: conv-as-GEMM windows)
  long long ldy, y_batch_stride;     // y row / batch stride in elements
  long long ldr, res_batch_stride;   // residual row stride (0 = d_out) / batch stride (0 = T*ldr)
  int res_broadcast;                 // 1: the same [T, d_out] residual for every b (positional embedding)
  // ---- split LoRA path: U to a caller workspace, then the dense kernel with one extra K block per tile
  void* u_ws;        // bf16 [n_sets][B, T, r] workspace, or null = single-launch kernel (U stays in shared memory)
  const float* u_w;  // fp32 [B, u_w_ld] per-utterance weights of the rank-column groups of U (null = none): soft_fused mix
  int u_w_ld, u_w_group;   // groups per utterance / rank columns per group (the language adapters' own rank, % 16 == 0)
  int u_phase;       // split path: 0 = both launches, 1 = U pass only (y may be null), 2 = dense launch only (ws holds U)
  int u_only;        // internal: run only the U pass
  int u_ld;          // internal: > 0 = u_out is [n_sets][B, T, u_ld] with u_ld == r (one plane per LoRA set)
};
int k1_qv_lora_fwd(const K1Args& a, cudaStream_t stream);
int k1v2_qv_lora_fwd(const K1Args& a, int block_n, cudaStream_t stream);
int attn_proj_fwd(const K1Args& a, cudaStream_t stream);
bool skinny_applicable(const K1Args& a, int M);
int skinny_fwd(const K1Args& a, int M, cudaStream_t stream);
int attn_proj_fwd_rows(const K1Args& a, const int32_t* row_adapter, int M, cudaStream_t stream);
int attn_fwd(const void* q, const void* k, const void* v, void* out, int BH, int Tq, int Tk, int head_dim, int causal,
             cudaStream_t stream);
int decode_self_attn(const void* q, const void* k_new, const void* v_new, void* cache_k, void* cache_v,
                     const long long* pos, void* out, int B, int H, int head_dim, int t_max, cudaStream_t stream);
int decode_cross_attn(const void* q, const void* k, const void* v, void* out, int B, int H, int head_dim, int Tk,
                      cudaStream_t stream);
int logmel_fwd(const float* wave, const float* window, const float* cos_t, const float* sin_t, const float* filters,
               float* raw_ws, int* clip_max_ws, void* out, int B, int n_samples, int n_mels, int out_bf16,
               cudaStream_t stream);
int layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean_out, float* rstd_out, int64_t M,
                  int d, float eps, cudaStream_t stream);
int ln_lora_u_fwd(const void* h, const void* gamma, const void* beta, void* x, const void* A_cat,
                  const int32_t* utt_adapter, void* u_out, int B, int T, int d, int r, int n_sets, int n_adapters,
                  float scale, float eps, cudaStream_t stream);
bool ln_lora_u_supported(int d, int r, int n_sets);
int operand_refresh(const void* desc, int n_desc, int max_elems, cudaStream_t stream);

struct K2Args {
  const void* h;
  int h_is_fp32;
  const float *ln_w, *ln_b, *W1, *b1, *g1, *be1, *W2, *b2, *g2, *be2, *W3, *b3;
  int B, T, d, h1, h2, C;
  float* logits;
  float* probs;
  int32_t* idx;
  int32_t* perm;
  int32_t* seg_starts;
  void* ws;
  // optional: h is the encoder's residual stream BEFORE its final LayerNorm; apply that LayerNorm (bf16 gamma / beta,
  // result rounded to bf16 like the stored encoder output) on the fly.  Null = h is the encoder output itself.
  const void* pre_ln_w;
  const void* pre_ln_b;
  float pre_ln_eps;
};
int64_t k2_workspace_bytes(int64_t B, int64_t T, int64_t d);
int k2_router_fwd(const K2Args& a, cudaStream_t stream);

// Row-indexed (decode-step) variant: base GEMM through K1 + per-row gathered low-rank update.
int64_t rows_workspace_bytes(int64_t M, int64_t d, int64_t r);
int rows_qv_lora_fwd(const void* x, const void* W, const void* bias, const void* A_stack, const void* Bp_stack,
                     const int32_t* row_adapter, void* y, int M, int d_in, int d_out, int r, int n_adapters,
                     float scale, void* ws, cudaStream_t stream);

struct K3Args {
  const void *dy, *x, *u, *Wt, *At_stack, *Bt_stack, *Bp_stack;
  const int32_t* utt_adapter;
  void* dx;
  float *dA, *dB;
  int B, T, d_in, d_out, r, n_adapters;
  float scale;
  void* ws;
};
int64_t k3_workspace_bytes(int64_t rows, int64_t T, int64_t d, int64_t r, int64_t n_adapters);
int k3_qv_lora_bwd(const K3Args& a, cudaStream_t stream);

}  // namespace sar
