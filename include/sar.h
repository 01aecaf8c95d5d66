/*
 * sar.h — C ABI of libsar.so: the B200 (sm_100a) hot path of routed multi-adapter LoRA Whisper.
 *
 * The reference (dhruv0811/speech-adapter-routing) has no native/FFI layer at all: its hot path is
 * Python delegating to HF transformers + PEFT.  This header therefore *defines* the boundary a
 * maintainer would bind (ctypes stub in INTEGRATION.md).  Each entry point names the reference
 * code it replaces (paths relative to the reference repo; $HF = transformers/models/whisper).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only.  No torch types, no C++ exceptions, no abort().
 *   - Every data pointer is a DEVICE pointer owned by the caller (PyTorch).  The library never
 *     allocates, frees or retains caller memory.  `stream` is a cudaStream_t passed as void*;
 *     every call is asynchronous on that stream and never synchronises.
 *   - bf16 tensors are `uint16_t`-sized storage (torch.bfloat16), row-major, last dim contiguous.
 *   - Return value: 0 = ok, <0 = sar_status.  sar_last_error() returns a thread-local message.
 *   - There is NO CPU fallback: on a device that is not sm_100 every compute call returns
 *     SAR_EARCH; without a CUDA device it returns SAR_ECUDA.
 */
#ifndef SAR_H_
#define SAR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAR_VERSION_MAJOR 0
#define SAR_VERSION_MINOR 2 /* 0.2: + layernorm_lora_u, layernorm_fwd_stats, router_fwd_fused_ln, attn_proj_fwd_mix, operand_refresh, SAR_ACT_GELU_BWD */

/* Rank padding of the packed lora_B stack (see sar_qv_lora_fwd). */
#define SAR_RPAD 64

typedef enum sar_status {
  SAR_OK = 0,
  SAR_EINVAL = -1, /* bad shape / alignment / null pointer */
  SAR_EARCH = -2,  /* device is not sm_100 (B200) */
  SAR_ECUDA = -3,  /* CUDA runtime / launch failure */
  SAR_EWORKSPACE = -4 /* workspace too small */
} sar_status;

/* flags for sar_qv_lora_fwd* */
#define SAR_FLAG_NONE 0u
#define SAR_FLAG_SAVE_U 1u /* also write u = scale*(x·A_k^T) (bf16 [B*T, r]) to ws for the backward */
/* sar_attn_proj_fwd, split path (ws != NULL): */
#define SAR_FLAG_U_READY 4u /* ws already holds U (e.g. written by sar_layernorm_lora_u_fwd): run the dense launch only */
#define SAR_FLAG_U_ONLY 8u  /* run only the U pass (U = scale·x·A_kᵀ of every set into ws); y is not written */

/* ops for sar_workspace_bytes */
typedef enum sar_op {
  SAR_OP_QV_LORA_FWD = 0,
  SAR_OP_ROUTER_FWD = 1,
  SAR_OP_QV_LORA_BWD = 2,
  SAR_OP_QV_LORA_FWD_ROWS = 3,
  SAR_OP_ATTN_PROJ_FWD = 4 /* rows = B*T, n = number of LoRA sets: the U workspace of the split path */
} sar_op;

/* Library version as major*1000+minor. */
int sar_version(void);

/* Thread-local description of the last error returned on this thread ("" if none). */
const char* sar_last_error(void);

/* 1 if the current CUDA device is sm_100 (B200), 0 if another GPU, <0 (SAR_ECUDA) if no device. */
int sar_device_ok(void);

/* Bytes of caller-provided workspace an op needs.  rows = B*T (or B for the router),
 * d = model width, r = LoRA rank, n = n_adapters (router: n classes). Returns <0 on bad op. */
int64_t sar_workspace_bytes(int op, int64_t rows, int64_t T, int64_t d, int64_t r, int64_t n);

/*
 * K1 — fused q/v projection with routed low-rank epilogue.
 *   y[b,t,:] = x[b,t,:]·Wᵀ + bias + scale · (x[b,t,:]·A_kᵀ)·B_kᵀ ,  k = utt_adapter[b]  (k = -1 → base only)
 * Replaces PEFT lora.Linear.forward at the q_proj / v_proj slots the reference injects in
 * src/models/whisper_lora.py:88-98, invoked from $HF/modeling_whisper.py:310 (q) and :332 (v),
 * and the per-utterance adapter selection of src/models/adapter_router.py:610-622.
 *
 *   x        bf16 [B, T, d_in]
 *   W        bf16 [d_out, d_in]          (nn.Linear weight)
 *   bias     bf16 [d_out] or NULL
 *   A_stack  bf16 [n_adapters, r, d_in]  (lora_A of every adapter, stacked)
 *   Bp_stack bf16 [n_adapters, d_out, SAR_RPAD]  (lora_B, columns >= r are zero padding)
 *   utt_adapter int32 [B] or NULL (NULL → base only)
 *   y        bf16 [B, T, d_out]
 *   u_out    bf16 [B*T, r] or NULL — scale*(x·A_kᵀ) rounded to bf16, saved for the backward
 * Constraints: d_in % 64 == 0, d_out % 64 == 0, r in {16,32,48,64}; all pointers 16-byte aligned.
 * The rank-r intermediate u is rounded to bf16 once (it is an MMA operand) with `scale` already applied.
 */
int sar_qv_lora_fwd(const void* x, const void* W, const void* bias, const void* A_stack,
                    const void* Bp_stack, const int32_t* utt_adapter, void* y, void* u_out,
                    int B, int T, int d_in, int d_out, int r, int n_adapters, float scale,
                    uint32_t flags, void* stream);

/*
 * Fused attention-projection stage: up to three projections of the SAME input in one launch,
 * with the routed low-rank term on any of them, the query scale folded into the epilogue and, optionally, the
 * [B,T,d] -> [B,h,T,64] head-major transpose done by the TMA store / load instead of separate copy kernels.
 *   seg s:  y_s = ( x·W_sᵀ + bias_s + (scale·x·A_{set,k}ᵀ)·B_{set,k}ᵀ ) * seg_scale[s],  set = seg_set[s] (-1: no LoRA)
 * Replaces, for one WhisperAttention.forward call ($HF/models/whisper/modeling_whisper.py:310-336): q_proj (LoRA) and
 * `* self.scaling` (:310), k_proj (:331), v_proj (LoRA, :332) and the three `.transpose(1, 2).contiguous()` copies
 * (:311-312, :333-334); with x_head_major = 1 and n_seg = 1 it is out_proj reading SDPA's output in place (:352-353).
 *
 *   x         bf16 [B, T, d_in]  or, if x_head_major, [B, d_in/64, T, 64]
 *   W_cat     bf16 [n_seg*d_out, d_in]   weights of the segments concatenated along the output dimension
 *   bias_cat  bf16 [n_seg*d_out] or NULL (k_proj has no bias: pass zeros in its slice)
 *   A_cat     bf16 [n_sets*n_adapters, r, d_in];  Bp_cat bf16 [n_sets*n_adapters, d_out, SAR_RPAD]
 *   y         n_seg output pointers, each bf16 [B, T, d_out] or, if y_head_major, [B, d_out/64, T, 64]
 * Constraints: head dim 64 for the head-major modes; d_out % 128 == 0; r in {16,32,48,64}; n_sets <= 2; n_seg <= 3.
 *
 *   ws   NULL: single launch, the rank-r intermediate U never leaves the SM (TMEM -> shared memory -> MMA operand);
 *        the TMEM budget (two accumulator buffers + U) then limits the tile to 128/192 columns.
 *        non-NULL (sar_workspace_bytes(SAR_OP_ATTN_PROJ_FWD, B*T, T, d_in, r, n_sets) bytes): split path — launch 1
 *        writes U = scale·x·A_kᵀ of every set to ws (bf16 [n_sets][B,T,r]), launch 2 is the dense 256-wide kernel with
 *        the low-rank term as one extra K block per tile.  Same rounding points, bit-identical results; measured
 *        ~1.9x faster at whisper-large-v3 shapes.
 */
int sar_attn_proj_fwd(const void* x, int x_head_major, const void* W_cat, const void* bias_cat, const void* A_cat,
                      const void* Bp_cat, const int32_t* utt_adapter, void* const* y, const int32_t* seg_set,
                      const float* seg_scale, int n_seg, int n_sets, int y_head_major, int B, int T, int d_in,
                      int d_out, int r, int n_adapters, float scale, uint32_t flags, void* ws, void* stream);

/*
 * sar_attn_proj_fwd with a per-utterance WEIGHTED MIX of adapters inside the low-rank term (an opt-in
 * alternative to the reference's soft routing, src/models/adapter_router.py:627-670, which runs every adapter's whole
 * model and mixes logits):   y = x·Wᵀ + b + Σ_g mix_w[b, g] · scale · (x·A_gᵀ)·B_gᵀ.
 * The caller stacks the language adapters ALONG THE RANK into one merged adapter per set (A_cat [n_sets, r, d_in] with
 * r = mix_groups * mix_group_rank <= 64, Bp_cat [n_sets, d_out, SAR_RPAD], n_adapters = 1, utt_adapter[b] = 0); the U
 * pass scales rank columns [g*mix_group_rank, (g+1)*mix_group_rank) of utterance b by mix_w[b*mix_groups + g] before the
 * single bf16 rounding.  Split path only (ws != NULL).  mix_w: fp32 [B, mix_groups] device pointer.
 */
int sar_attn_proj_fwd_mix(const void* x, int x_head_major, const void* W_cat, const void* bias_cat, const void* A_cat,
                          const void* Bp_cat, const int32_t* utt_adapter, void* const* y, const int32_t* seg_set,
                          const float* seg_scale, int n_seg, int n_sets, int y_head_major, int B, int T, int d_in,
                          int d_out, int r, int n_adapters, float scale, uint32_t flags, void* ws, const float* mix_w,
                          int mix_groups, int mix_group_rank, void* stream);

/*
 * Row-indexed form of sar_attn_proj_fwd for decode steps (one token per utterance: rows of different adapters share
 * a tile).  The base projections of all segments run as ONE dense launch over the M rows; one gathered BGMV launch
 * adds seg_scale[s]·(scale·x_m·A_{set,k}ᵀ)·B_{set,k}ᵀ, k = row_adapter[m], to every LoRA'd segment.
 *   x bf16 [M, d_in];  y: n_seg pointers, each bf16 [M, d_out] (for head dim 64 this IS [M, d_out/64, 1, 64]).
 * Replaces the q/k/v projections of WhisperAttention.forward at one token per utterance inside the reference's
 * per-sample adapter.generate loop (src/models/adapter_router.py:744-750).
 */
int sar_attn_proj_fwd_rows(const void* x, const void* W_cat, const void* bias_cat, const void* A_cat,
                           const void* Bp_cat, const int32_t* row_adapter, void* const* y,
                           const int32_t* seg_set, const float* seg_scale, int n_seg, int n_sets, int M,
                           int d_in, int d_out, int r, int n_adapters, float scale, uint32_t flags,
                           void* stream);

/*
 * softmax(Q·Kᵀ)·V for head dim 64 (flash-attention style forward on tcgen05 / TMEM / TMA), optional causal mask.
 * Q is already scaled.  Replaces the attention_interface call of WhisperAttention.forward
 * ($HF/models/whisper/modeling_whisper.py:341-350) on the head-major tensors sar_attn_proj_fwd writes.
 *   q, out  bf16 [B, H, Tq, 64];  k, v  bf16 [B, H, Tk, 64];  causal needs Tq == Tk.
 */
int sar_attn_fwd(const void* q, const void* k, const void* v, void* out, int B, int H, int Tq, int Tk,
                 int head_dim, int causal, void* stream);

/*
 * Self-attention of one decode step over a static KV cache (head dim 64), position read from device memory:
 *   cache_k[b,h,*pos,:] = k_new[b,h,:];  cache_v likewise;  out[b,h,:] = softmax_{t <= *pos}(q·cache_kᵀ)·cache_v
 * q is already scaled (sar_attn_proj_fwd* folds head_dim^-0.5 into the projection).  Replaces DynamicCache.update +
 * causal mask + SDPA of WhisperAttention.forward at tgt_len = 1 ($HF/models/whisper/modeling_whisper.py:326-350); a
 * device-side position lets one captured CUDA graph serve every token.
 *   q, k_new, v_new, out  bf16 [B, H, 64] (= [B, d] row-major);  cache_k, cache_v  bf16 [B, H, t_max, 64];  pos int64[1]
 */
int sar_decode_self_attn(const void* q, const void* k_new, const void* v_new, void* cache_k, void* cache_v,
                         const int64_t* pos, void* out, int B, int H, int head_dim, int t_max, void* stream);

/*
 * Cross-attention of one decode step (one query token per utterance, head dim 64) over read-only encoder K / V:
 *   out[b,h,:] = softmax_t( q[b,h,:]·k[b,h,t,:] ) · v[b,h,t,:],  t < Tk            (q is pre-scaled)
 * Replaces the attention_interface call of WhisperAttention.forward for encoder_attn at tgt_len = 1
 * ($HF/models/whisper/modeling_whisper.py:341-350) inside the per-sample decode loop of
 * src/models/adapter_router.py:744-750.  HBM-bound: streams 2·Tk·128 B per (b, h) per token.
 *   q, out  bf16 [B, H, 64];  k, v  bf16 [B, H, Tk, 64]
 */
int sar_decode_cross_attn(const void* q, const void* k, const void* v, void* out, int B, int H, int head_dim, int Tk,
                          void* stream);

/*
 * Whisper's log-mel front-end: waveform -> input_features, one launch sequence for the whole batch.
 *   frames of 400 samples every 160 (reflect-padded by 200, the last frame dropped), periodic Hann window, |DFT|^2 over
 *   201 bins, mel filterbank, log10(max(., 1e-10)), clamp to (clip maximum - 8), (x + 4) / 4
 * Replaces the per-example CPU call processor.feature_extractor(audio_array, ...) of src/data/dataset.py:124-128
 * ($HF/models/whisper/feature_extraction_whisper.py:105-135).  fp32 arithmetic on CUDA cores (see logmel.cu for why).
 *   wave         f32 [B, n_samples]   already padded / cut to the clip length (480000 for 30 s); n_samples % 160 == 0
 *   window       f32 [400]            0.5 - 0.5 cos(2 pi n / 400)
 *   cos_table, sin_table  f32 [400]   cos / sin(2 pi j / 400)
 *   mel_filters  f32 [201, n_mels]    as WhisperFeatureExtractor.mel_filters (Slaney scale and norm)
 *   raw_ws       f32 [B, n_mels, n_samples / 160]   workspace;  clip_max_ws  int32 [B]  workspace
 *   out          bf16 or f32 [B, n_mels, n_samples / 160]
 */
int sar_logmel_fwd(const float* wave, const float* window, const float* cos_table, const float* sin_table,
                   const float* mel_filters, float* raw_ws, int32_t* clip_max_ws, void* out, int B, int n_samples,
                   int n_mels, int out_bf16, void* stream);

/* epilogue activations of sar_linear_fwd */
#define SAR_ACT_NONE 0
#define SAR_ACT_GELU 1 /* erf-form GELU, HF ACT2FN["gelu"] */
/* sar_linear_fwd only: y = (x·Wᵀ) * GELU'(residual) — the GELU backward of the training layers fused into the dX GEMM of
 * fc2 (autograd through ACT2FN["gelu"], $HF/modeling_whisper.py:403); `residual` carries the saved pre-activation. */
#define SAR_ACT_GELU_BWD 2

/*
 * Self-attention q‖v pair: one x, two LoRA'd projections, two row-major outputs.
 *   y_q = x·W_qᵀ + b_q + (scale·x·A_{q,k}ᵀ)·B_{q,k}ᵀ ;  y_v likewise.  Thin wrapper over sar_attn_proj_fwd
 *   (n_seg = 2, n_sets = 2); W_cat = [W_q; W_v], A_cat = [A_q stack; A_v stack], Bp_cat likewise.
 * Replaces the q_proj + v_proj calls of one WhisperAttention.forward ($HF/modeling_whisper.py:310, :332).
 */
int sar_qv_lora_fwd_pair(const void* x, const void* W_cat, const void* bias_cat, const void* A_cat,
                         const void* Bp_cat, const int32_t* utt_adapter, void* y_q, void* y_v, int B,
                         int T, int d_in, int d_out, int r, int n_adapters, float scale, uint32_t flags,
                         void* stream);

/*
 * Plain dense layer of the Whisper block on the same tcgen05 pair kernel, with the neighbouring elementwise ops
 * folded into the epilogue:
 *   y = act(x·Wᵀ + bias) + residual
 * Replaces, per layer, out_proj + `residual + h` ($HF/modeling_whisper.py:352-353 + :399 / :481 / :495), fc1 + GELU
 * (:403 / :500) and fc2 + `residual + h` (:405-407 / :502-504), including the `.transpose(1,2).contiguous()` of
 * SDPA's output when x_head_major = 1.
 *   x        bf16 [B, T, d_in]  or, if x_head_major, [B, d_in/64, T, 64]
 *   W        bf16 [d_out, d_in];  bias bf16 [d_out] or NULL
 *   residual bf16 [B, T, d_out] or NULL (may alias y)
 *   y        bf16 [B, T, d_out]
 * Constraints: d_in % 64 == 0, d_out % 128 == 0, pointers 16-byte aligned.
 */
int sar_linear_fwd(const void* x, int x_head_major, const void* W, const void* bias, const void* residual,
                   void* y, int B, int T, int d_in, int d_out, int act, uint32_t flags, void* stream);

/*
 * General strided form of sar_linear_fwd (same kernel):  y = act(x·Wᵀ + bias) + residual  with explicit strides, a
 * ragged K (d_in need not be a multiple of 64: the tail is zero-filled by TMA) and a ragged last N tile.  It exists
 * for the layers either side of the transformer blocks:
 *   - lm_head (proj_out, d_out = 51865/51866, no bias): y row stride padded to a multiple of 8 elements
 *     ($HF/modeling_whisper.py:1135);
 *   - the convolutional front-end as GEMMs over OVERLAPPING rows of a channels-last, zero-padded frame buffer:
 *     conv1 (k=3, s=1) = rows of 3*n_mels elements at stride n_mels; conv2 (k=3, s=2) = rows of 3*d at stride 2*d,
 *     GELU in the epilogue, conv2 also adds the positional embedding as a batch-broadcast residual (:626-633).
 *   x        bf16, row t of utterance b at x + b*x_batch_stride + t*ldx, d_in elements (rows may overlap)
 *   W        bf16 [d_out, d_in] contiguous;  bias bf16 [d_out] or NULL
 *   residual bf16, row at residual + b*res_batch_stride + t*ldr, or NULL; res_broadcast=1: same rows for every b
 *   y        bf16, row at y + b*y_batch_stride + t*ldy
 * Strides are in elements; 0 selects the contiguous default (ldx=d_in, x_batch_stride=T*ldx, ...).
 * When d_out is not a multiple of 8, columns [d_out, round_up(d_out, 8)) of each y row may be overwritten with zeros
 * (TMA clips stores at 16-byte granularity): give y a row stride of at least round_up(d_out, 8).
 * Constraints: d_in and every stride multiples of 8; bias / residual need d_out % 64 == 0; 16-byte aligned pointers.
 */
int sar_dense_fwd(const void* x, int64_t ldx, int64_t x_batch_stride, const void* W, const void* bias,
                  const void* residual, int64_t ldr, int64_t res_batch_stride, int res_broadcast, void* y,
                  int64_t ldy, int64_t y_batch_stride, int B, int T, int d_in, int d_out, int act,
                  uint32_t flags, void* stream);

/*
 * LayerNorm over the last dimension, fp32 statistics, bf16 in / out (HBM-bound: one read + one write of x):
 *   y[m,:] = (x[m,:] - mean) * rsqrt(var + eps) * gamma + beta
 * Replaces nn.LayerNorm at self_attn_layer_norm / encoder_attn_layer_norm / final_layer_norm / layer_norm
 * ($HF/modeling_whisper.py:391, :401, :470, :486, :498, :643, :797).
 *   x, y bf16 [M, d];  gamma, beta bf16 [d].  Constraints: d % 8 == 0, d <= 2048.
 */
int sar_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, int64_t M, int d, float eps,
                      void* stream);
/* The same, also writing the row statistics (fp32 mean[M] and rstd[M] = 1/sqrt(var + eps)) that a LayerNorm backward
 * needs: the training layers' forward (autograd through nn.LayerNorm in src/training/trainer.py:251-256). */
int sar_layernorm_fwd_stats(const void* x, const void* gamma, const void* beta, void* y, float* mean, float* rstd, int64_t M,
                            int d, float eps, void* stream);

/*
 * LayerNorm fused with the LoRA down-projection of the attention projections that consume it (one pass over h):
 *   x[b,t,:]    = LayerNorm(h[b,t,:])                                   bf16 [B, T, d]   (as sar_layernorm_fwd)
 *   u[s][b,t,:] = bf16(scale · x[b,t,:] · A_cat[s*n_adapters + k]ᵀ)     bf16 [n_sets][B, T, r],  k = utt_adapter[b]
 * u is exactly the workspace layout of sar_attn_proj_fwd's split path: pass it as `ws` with SAR_FLAG_U_READY and the
 * U pass (a second read of all of x) disappears.  Replaces nn.LayerNorm ($HF/modeling_whisper.py:392, :470, :483) +
 * PEFT's lora_A(x) at q_proj / v_proj (src/models/whisper_lora.py:88-98) for the layers' pre-attention norms.
 * Utterances with k < 0 or k >= n_adapters get x only (their u rows are not written and never read).
 * Constraints: d / 32 in {8, 12, 16, 24}; n_sets * r / 16 in {1, 2}; sar_layernorm_lora_u_supported() tells.
 */
int sar_layernorm_lora_u_fwd(const void* h, const void* gamma, const void* beta, void* x, const void* A_cat,
                             const int32_t* utt_adapter, void* u, int B, int T, int d, int r, int n_sets,
                             int n_adapters, float scale, float eps, void* stream);
int sar_layernorm_lora_u_supported(int d, int r, int n_sets);

/*
 * Operand refresh for captured training steps.  PEFT reads lora_A / lora_B straight from the parameters on every
 * forward (peft/tuners/lora/layer.py Linear.forward, called from src/models/whisper_lora.py:88-98), so an optimizer
 * step is visible to the next forward by construction.  libsar's kernels read derived bf16 operands (rank-padded A
 * stacks, 64-column B stacks, their transposes for sar_qv_lora_bwd, the concatenations of sar_attn_proj_fwd); when the
 * training step is replayed as a CUDA graph no host code runs between optimizer steps, so this launch — a node of
 * that graph — re-derives every operand block from the live fp32 (or bf16) parameters:
 *   for each descriptor:  dst[i*dst_rs + j*dst_cs] = bf16(scale * src[i*src_rs + j*src_cs]),  i < rows, j < cols
 * `desc` is a DEVICE array of n_desc descriptors (strides in elements); max_elems = max over descriptors of rows*cols.
 * Padding rows / columns of the destination stacks are not touched (they stay zero).
 */
#define SAR_DTYPE_F32 0
#define SAR_DTYPE_BF16 1
typedef struct sar_refresh_desc {
  const void* src; /* parameter block (device) */
  void* dst;       /* bf16 operand block (device) */
  int32_t rows, cols;
  int64_t src_rs, src_cs; /* element strides of src along i, j */
  int64_t dst_rs, dst_cs; /* element strides of dst along i, j */
  float scale;
  int32_t src_dtype; /* SAR_DTYPE_F32 | SAR_DTYPE_BF16 */
} sar_refresh_desc;
int sar_operand_refresh(const sar_refresh_desc* desc, int n_desc, int max_elems, void* stream);

/*
 * Row-indexed variant for decode steps (T = 1 per utterance, rows of different adapters share a tile):
 *   y[m,:] = x[m,:]·Wᵀ + bias + scale·(x[m,:]·A_kᵀ)·B_kᵀ, k = row_adapter[m].
 * Replaces the per-sample adapter.generate loop of src/models/adapter_router.py:744-750 for the
 * decoder self-attn q/v and cross-attn q projections at one token per utterance.
 *   ws: workspace of sar_workspace_bytes(SAR_OP_QV_LORA_FWD_ROWS, M, 1, d, r, n) bytes.
 */
int sar_qv_lora_fwd_rows(const void* x, const void* W, const void* bias, const void* A_stack,
                         const void* Bp_stack, const int32_t* row_adapter, void* y, int M,
                         int d_in, int d_out, int r, int n_adapters, float scale, void* ws,
                         void* stream);

/*
 * K2 — language-ID router: LayerNorm(d) per frame → mean over T → MLP(d→h1→h2→C, LN+ReLU between)
 * → softmax → argmax, plus the adapter-index bookkeeping (stable counting sort by class).
 * Replaces LanguageClassifier.forward/_pool_features/predict (src/models/adapter_router.py:251-312,
 * default config: pooling="mean", use_layer_norm=True, use_cnn=False, two hidden layers) and the
 * Python list bookkeeping of AdapterRouter.detect_language (:550-566).
 *
 *   h          bf16 (h_is_fp32=0) or fp32 (h_is_fp32=1) [B, T, d] encoder hidden states
 *   ln_w,ln_b  fp32 [d]      layer_norm.{weight,bias}
 *   W1,b1      fp32 [h1,d],[h1]   classifier.0     g1,be1 fp32 [h1]  classifier.1 (LayerNorm)
 *   W2,b2      fp32 [h2,h1],[h2]  classifier.4     g2,be2 fp32 [h2]  classifier.5 (LayerNorm)
 *   W3,b3      fp32 [C,h2],[C]    classifier.8
 *   logits_out, probs_out fp32 [B,C];  idx_out int32 [B] (argmax, first maximal index on ties)
 *   perm_out   int32 [B]   utterance indices stably sorted by idx
 *   seg_starts_out int32 [C+1]  segment k = perm[seg_starts[k] : seg_starts[k+1]]
 *   ws         workspace of sar_workspace_bytes(SAR_OP_ROUTER_FWD, B, T, d, 0, C) bytes
 * Constraints: d % 8 == 0, d <= 2048, h1,h2 <= 1024, C <= 64, eps = 1e-5 for all LayerNorms.
 */
int sar_router_fwd(const void* h, int h_is_fp32, const float* ln_w, const float* ln_b,
                   const float* W1, const float* b1, const float* g1, const float* be1,
                   const float* W2, const float* b2, const float* g2, const float* be2,
                   const float* W3, const float* b3, int B, int T, int d, int h1, int h2, int C,
                   float* logits_out, float* probs_out, int32_t* idx_out, int32_t* perm_out,
                   int32_t* seg_starts_out, void* ws, void* stream);

/*
 * K2 with the encoder's final LayerNorm folded in: h_pre is the last encoder layer's residual stream
 * (bf16 [B, T, d]) BEFORE WhisperEncoder.layer_norm ($HF/modeling_whisper.py:643); the kernel applies that LayerNorm
 * (enc_ln_w / enc_ln_b bf16 [d], result rounded to bf16 exactly like the stored encoder output), then the LID head's
 * LayerNorm and the mean over T (src/models/adapter_router.py:268, :229) — one read of h_pre instead of LayerNorm
 * write + K2 read of the [B, 1500, d] encoder output.  Everything else as sar_router_fwd.
 */
int sar_router_fwd_fused_ln(const void* h_pre, const void* enc_ln_w, const void* enc_ln_b, float enc_ln_eps,
                            const float* ln_w, const float* ln_b, const float* W1, const float* b1, const float* g1,
                            const float* be1, const float* W2, const float* b2, const float* g2, const float* be2,
                            const float* W3, const float* b3, int B, int T, int d, int h1, int h2, int C,
                            float* logits_out, float* probs_out, int32_t* idx_out, int32_t* perm_out,
                            int32_t* seg_starts_out, void* ws, void* stream);

/*
 * K3 — LoRA-only backward of K1 (base W frozen):
 *   dx      = dy·W + (scale·dy·B_k)·A_k                (skipped when dx == NULL)
 *   dA_k   += scale · (dy·B_k)ᵀ · x      fp32 [n_adapters, r, d_in]
 *   dB_k   += (dy)ᵀ · u                  fp32 [n_adapters, d_out, r]   (u = scale·x·A_kᵀ from the forward)
 * Replaces autograd through PEFT lora.Linear triggered at src/training/trainer.py:251-256.
 * dA / dB point into one flat fp32 gradient bucket that the host all-reduces with NCCL.
 *   Wt        bf16 [d_in, d_out]  (W transposed, cached by the caller: W is frozen)
 *   At_stack  bf16 [n_adapters, d_in, SAR_RPAD] (lora_A transposed + rank-padded)
 *   Bt_stack  bf16 [n_adapters, r, d_out]       (lora_B transposed)
 *   ws        workspace of sar_workspace_bytes(SAR_OP_QV_LORA_BWD, B*T, T, max(d_in,d_out), r, n_adapters)
 */
int sar_qv_lora_bwd(const void* dy, const void* x, const void* u, const void* Wt,
                    const void* At_stack, const void* Bt_stack, const void* Bp_stack,
                    const int32_t* utt_adapter, void* dx, float* dA, float* dB, int B, int T,
                    int d_in, int d_out, int r, int n_adapters, float scale, void* ws,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SAR_H_ */
