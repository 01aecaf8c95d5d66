// sar_api.cu — the extern "C" surface of libsar.so (see include/sar.h) plus device/tensor-map plumbing.
#include <mutex>
#include <stdio.h>

#include "sar_internal.h"

namespace sar {

static thread_local char g_err[512] = "";

int fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
int fail_cuda(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return SAR_ECUDA;
}

static std::mutex g_dev_mu;
static DeviceInfo g_dev[64];
static bool g_dev_init[64];
static DeviceInfo g_nodev = {-3, -1, 0, 0, 0, 0, 0};

const DeviceInfo& device_info() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    cudaGetLastError();
    return g_nodev;
  }
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (!g_dev_init[dev]) {
    DeviceInfo d{};
    d.device = dev;
    cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev);
    cudaDeviceGetAttribute(&d.num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&d.l2_bytes, cudaDevAttrL2CacheSize, dev);
    d.ok = (d.cc_major == 10 && d.cc_minor == 0) ? 1 : 0;
    g_dev[dev] = d;
    g_dev_init[dev] = true;
  }
  return g_dev[dev];
}

int require_sm100() {
  const DeviceInfo& d = device_info();
  if (d.ok < 0) return fail(SAR_ECUDA, "no CUDA device available (libsar has no CPU fallback)");
  if (d.ok == 0) return fail(SAR_EARCH, "device is not sm_100 (B200); libsar kernels are sm_100a only");
  return SAR_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_encode_once;

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int l2_promotion_bytes) {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  });
  if (!g_encode) return fail(SAR_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                        gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        l2_promotion_bytes >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                                  : (l2_promotion_bytes >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                                               : CU_TENSOR_MAP_L2_PROMOTION_NONE),
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled failed (CUresult %d, rank %d, dims %llu x %llu)", (int)r, rank,
             (unsigned long long)dims[0], (unsigned long long)dims[1]);
    return SAR_EINVAL;
  }
  return SAR_OK;
}

}  // namespace sar

using namespace sar;

extern "C" {

int sar_version(void) { return SAR_VERSION_MAJOR * 1000 + SAR_VERSION_MINOR; }

const char* sar_last_error(void) { return g_err; }

int sar_device_ok(void) {
  const DeviceInfo& d = device_info();
  return d.ok < 0 ? SAR_ECUDA : d.ok;
}

int64_t sar_workspace_bytes(int op, int64_t rows, int64_t T, int64_t d, int64_t r, int64_t n) {
  switch (op) {
    case SAR_OP_QV_LORA_FWD: return 0;
    case SAR_OP_ROUTER_FWD: return k2_workspace_bytes(rows, T, d);
    case SAR_OP_QV_LORA_FWD_ROWS: return rows_workspace_bytes(rows, d, r);
    case SAR_OP_QV_LORA_BWD: return k3_workspace_bytes(rows, T, d, r, n);
    case SAR_OP_ATTN_PROJ_FWD: return rows * (r > 0 ? r : 64) * (n > 0 ? n : 1) * 2;   // U: bf16 [n_sets][rows, r]
    default: fail(SAR_EINVAL, "sar_workspace_bytes: unknown op"); return SAR_EINVAL;
  }
}

int sar_qv_lora_fwd(const void* x, const void* W, const void* bias, const void* A_stack, const void* Bp_stack,
                    const int32_t* utt_adapter, void* y, void* u_out, int B, int T, int d_in, int d_out, int r,
                    int n_adapters, float scale, uint32_t flags, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  K1Args a{};
  a.x = x; a.W = W; a.bias = bias; a.A_stack = A_stack; a.Bp_stack = Bp_stack; a.utt_adapter = utt_adapter;
  a.y = y; a.u_out = (flags & SAR_FLAG_SAVE_U) ? u_out : nullptr;
  a.B = B; a.T = T; a.d_in = d_in; a.d_out = d_out; a.r = r; a.n_adapters = n_adapters; a.scale = scale;
  a.block_n_override = static_cast<int>((flags >> 8) & 0x3FF);   // debug/tuning: bits [8,18) = BLOCK_N
  a.grid_override = static_cast<int>((flags >> 18) & 0x3FF);     // debug/tuning: bits [18,28) = grid size
  a.kernel_override = static_cast<int>((flags >> 28) & 0x3);     // debug/tuning: bits [28,30): 1 = single-CTA, 2 = pair
  a.swap_halves = static_cast<int>((flags >> 1) & 0x1);          // debug: bit 1
  if ((flags & SAR_FLAG_SAVE_U) && !u_out) return fail(SAR_EINVAL, "sar_qv_lora_fwd: SAVE_U without u_out");
  return k1_qv_lora_fwd(a, static_cast<cudaStream_t>(stream));
}

static int attn_proj_entry(const void* x, int x_head_major, const void* W_cat, const void* bias_cat, const void* A_cat,
                           const void* Bp_cat, const int32_t* utt_adapter, void* const* y, const int32_t* seg_set,
                           const float* seg_scale, int n_seg, int n_sets, int y_head_major, int B, int T, int d_in,
                           int d_out, int r, int n_adapters, float scale, uint32_t flags, void* ws, const float* mix_w,
                           int mix_groups, int mix_group_rank, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  if (!y || !seg_set || !seg_scale || n_seg < 1 || n_seg > 3) return fail(SAR_EINVAL, "sar_attn_proj_fwd: bad segments");
  if (reinterpret_cast<uintptr_t>(ws) & 15) return fail(SAR_EINVAL, "sar_attn_proj_fwd: ws must be 16-byte aligned");
  K1Args a{};
  a.x = x; a.W = W_cat; a.bias = bias_cat; a.A_stack = A_cat; a.Bp_stack = Bp_cat; a.utt_adapter = utt_adapter;
  a.B = B; a.T = T; a.d_in = d_in; a.d_out = d_out; a.r = r; a.n_adapters = n_adapters; a.scale = scale;
  a.block_n_override = static_cast<int>((flags >> 8) & 0x3FF);
  a.grid_override = static_cast<int>((flags >> 18) & 0x3FF);
  a.n_seg = n_seg; a.n_sets = n_sets; a.x_head_major = x_head_major; a.y_head_major = y_head_major;
  a.u_ws = ws;
  a.u_w = mix_w; a.u_w_ld = mix_groups; a.u_w_group = mix_group_rank;
  if (mix_w && (mix_groups < 1 || mix_group_rank < 16 || mix_group_rank % 16 || mix_groups * mix_group_rank != r))
    return fail(SAR_EINVAL, "sar_attn_proj_fwd_mix: r must equal mix_groups * mix_group_rank (a multiple of 16)");
  if (mix_w && !ws) return fail(SAR_EINVAL, "sar_attn_proj_fwd_mix: needs the split path (ws)");
  a.u_phase = (flags & SAR_FLAG_U_ONLY) ? 1 : ((flags & SAR_FLAG_U_READY) ? 2 : 0);
  if (a.u_phase && !ws) return fail(SAR_EINVAL, "sar_attn_proj_fwd: U_ONLY / U_READY need the U workspace");
  for (int s = 0; s < n_seg; ++s) {
    a.seg_set[s] = seg_set[s];
    a.seg_scale[s] = seg_scale[s];
    a.y_seg[s] = y[s];
  }
  for (int s = n_seg; s < 3; ++s) {
    a.seg_set[s] = -1;
    a.seg_scale[s] = 1.0f;
  }
  return attn_proj_fwd(a, static_cast<cudaStream_t>(stream));
}

int sar_attn_proj_fwd(const void* x, int x_head_major, const void* W_cat, const void* bias_cat, const void* A_cat,
                      const void* Bp_cat, const int32_t* utt_adapter, void* const* y, const int32_t* seg_set,
                      const float* seg_scale, int n_seg, int n_sets, int y_head_major, int B, int T, int d_in,
                      int d_out, int r, int n_adapters, float scale, uint32_t flags, void* ws, void* stream) {
  return attn_proj_entry(x, x_head_major, W_cat, bias_cat, A_cat, Bp_cat, utt_adapter, y, seg_set, seg_scale, n_seg, n_sets,
                         y_head_major, B, T, d_in, d_out, r, n_adapters, scale, flags, ws, nullptr, 0, 0, stream);
}

int sar_attn_proj_fwd_mix(const void* x, int x_head_major, const void* W_cat, const void* bias_cat, const void* A_cat,
                          const void* Bp_cat, const int32_t* utt_adapter, void* const* y, const int32_t* seg_set,
                          const float* seg_scale, int n_seg, int n_sets, int y_head_major, int B, int T, int d_in,
                          int d_out, int r, int n_adapters, float scale, uint32_t flags, void* ws, const float* mix_w,
                          int mix_groups, int mix_group_rank, void* stream) {
  if (!mix_w) return fail(SAR_EINVAL, "sar_attn_proj_fwd_mix: null mix_w");
  return attn_proj_entry(x, x_head_major, W_cat, bias_cat, A_cat, Bp_cat, utt_adapter, y, seg_set, seg_scale, n_seg, n_sets,
                         y_head_major, B, T, d_in, d_out, r, n_adapters, scale, flags, ws, mix_w, mix_groups, mix_group_rank,
                         stream);
}

int sar_attn_proj_fwd_rows(const void* x, const void* W_cat, const void* bias_cat, const void* A_cat,
                           const void* Bp_cat, const int32_t* row_adapter, void* const* y, const int32_t* seg_set,
                           const float* seg_scale, int n_seg, int n_sets, int M, int d_in, int d_out, int r,
                           int n_adapters, float scale, uint32_t flags, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  if (!y || !seg_set || !seg_scale || n_seg < 1 || n_seg > 3)
    return fail(SAR_EINVAL, "sar_attn_proj_fwd_rows: bad segments");
  K1Args a{};
  a.x = x; a.W = W_cat; a.bias = bias_cat; a.A_stack = A_cat; a.Bp_stack = Bp_cat;
  a.d_in = d_in; a.d_out = d_out; a.r = r; a.n_adapters = n_adapters; a.scale = scale;
  a.block_n_override = static_cast<int>((flags >> 8) & 0x3FF);
  a.grid_override = static_cast<int>((flags >> 18) & 0x3FF);
  a.n_seg = n_seg; a.n_sets = n_sets;
  for (int s = 0; s < 3; ++s) {
    a.seg_set[s] = s < n_seg ? seg_set[s] : -1;
    a.seg_scale[s] = s < n_seg ? seg_scale[s] : 1.0f;
    a.y_seg[s] = s < n_seg ? y[s] : nullptr;
  }
  return attn_proj_fwd_rows(a, row_adapter, M, static_cast<cudaStream_t>(stream));
}

int sar_qv_lora_fwd_pair(const void* x, const void* W_cat, const void* bias_cat, const void* A_cat,
                         const void* Bp_cat, const int32_t* utt_adapter, void* y_q, void* y_v, int B, int T, int d_in,
                         int d_out, int r, int n_adapters, float scale, uint32_t flags, void* stream) {
  void* ys[2] = {y_q, y_v};
  const int32_t sets[2] = {0, 1};
  const float scales[2] = {1.0f, 1.0f};
  return sar_attn_proj_fwd(x, 0, W_cat, bias_cat, A_cat, Bp_cat, utt_adapter, ys, sets, scales, 2, 2, 0, B, T, d_in,
                           d_out, r, n_adapters, scale, flags, nullptr, stream);
}

int sar_linear_fwd(const void* x, int x_head_major, const void* W, const void* bias, const void* residual, void* y,
                   int B, int T, int d_in, int d_out, int act, uint32_t flags, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  if (act != SAR_ACT_NONE && act != SAR_ACT_GELU && act != SAR_ACT_GELU_BWD)
    return fail(SAR_EINVAL, "sar_linear_fwd: unknown activation");
  if (reinterpret_cast<uintptr_t>(residual) & 15) return fail(SAR_EINVAL, "sar_linear_fwd: residual must be 16-byte aligned");
  K1Args a{};
  a.x = x; a.W = W; a.bias = bias; a.B = B; a.T = T; a.d_in = d_in; a.d_out = d_out; a.r = 16; a.scale = 0.f;
  a.block_n_override = static_cast<int>((flags >> 8) & 0x3FF);
  a.grid_override = static_cast<int>((flags >> 18) & 0x3FF);
  a.n_seg = 1; a.n_sets = 1; a.x_head_major = x_head_major; a.y_head_major = 0;
  a.seg_set[0] = a.seg_set[1] = a.seg_set[2] = -1;
  a.seg_scale[0] = a.seg_scale[1] = a.seg_scale[2] = 1.0f;
  a.y_seg[0] = y;
  a.residual = residual; a.act = act;
  return attn_proj_fwd(a, static_cast<cudaStream_t>(stream));
}

int sar_dense_fwd(const void* x, int64_t ldx, int64_t x_batch_stride, const void* W, const void* bias,
                  const void* residual, int64_t ldr, int64_t res_batch_stride, int res_broadcast, void* y,
                  int64_t ldy, int64_t y_batch_stride, int B, int T, int d_in, int d_out, int act, uint32_t flags,
                  void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  if (act != SAR_ACT_NONE && act != SAR_ACT_GELU) return fail(SAR_EINVAL, "sar_dense_fwd: unknown activation");
  if (ldx < 0 || x_batch_stride < 0 || ldy < 0 || y_batch_stride < 0 || ldr < 0 || res_batch_stride < 0)
    return fail(SAR_EINVAL, "sar_dense_fwd: negative stride");
  if (reinterpret_cast<uintptr_t>(residual) & 15) return fail(SAR_EINVAL, "sar_dense_fwd: residual must be 16-byte aligned");
  K1Args a{};
  a.x = x; a.W = W; a.bias = bias; a.B = B; a.T = T; a.d_in = d_in; a.d_out = d_out; a.r = 16; a.scale = 0.f;
  a.block_n_override = static_cast<int>((flags >> 8) & 0x3FF);
  a.grid_override = static_cast<int>((flags >> 18) & 0x3FF);
  a.n_seg = 1; a.n_sets = 1;
  a.seg_set[0] = a.seg_set[1] = a.seg_set[2] = -1;
  a.seg_scale[0] = a.seg_scale[1] = a.seg_scale[2] = 1.0f;
  a.y_seg[0] = y;
  a.residual = residual; a.act = act;
  a.ldx = ldx; a.x_batch_stride = x_batch_stride; a.ldy = ldy; a.y_batch_stride = y_batch_stride;
  a.ldr = ldr; a.res_batch_stride = res_batch_stride; a.res_broadcast = res_broadcast;
  return attn_proj_fwd(a, static_cast<cudaStream_t>(stream));
}

int sar_attn_fwd(const void* q, const void* k, const void* v, void* out, int B, int H, int Tq, int Tk, int head_dim,
                 int causal, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  if (B <= 0 || H <= 0) return fail(SAR_EINVAL, "sar_attn_fwd: B and H must be positive");
  return attn_fwd(q, k, v, out, B * H, Tq, Tk, head_dim, causal, static_cast<cudaStream_t>(stream));
}

int sar_decode_self_attn(const void* q, const void* k_new, const void* v_new, void* cache_k, void* cache_v,
                         const int64_t* pos, void* out, int B, int H, int head_dim, int t_max, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  return decode_self_attn(q, k_new, v_new, cache_k, cache_v, reinterpret_cast<const long long*>(pos), out, B, H,
                          head_dim, t_max, static_cast<cudaStream_t>(stream));
}

int sar_decode_cross_attn(const void* q, const void* k, const void* v, void* out, int B, int H, int head_dim, int Tk,
                          void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  return decode_cross_attn(q, k, v, out, B, H, head_dim, Tk, static_cast<cudaStream_t>(stream));
}

int sar_logmel_fwd(const float* wave, const float* window, const float* cos_table, const float* sin_table,
                   const float* mel_filters, float* raw_ws, int32_t* clip_max_ws, void* out, int B, int n_samples,
                   int n_mels, int out_bf16, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  return logmel_fwd(wave, window, cos_table, sin_table, mel_filters, raw_ws, reinterpret_cast<int*>(clip_max_ws), out, B,
                    n_samples, n_mels, out_bf16, static_cast<cudaStream_t>(stream));
}

int sar_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, int64_t M, int d, float eps,
                      void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  return layernorm_fwd(x, gamma, beta, y, nullptr, nullptr, M, d, eps, static_cast<cudaStream_t>(stream));
}

int sar_layernorm_fwd_stats(const void* x, const void* gamma, const void* beta, void* y, float* mean, float* rstd, int64_t M,
                            int d, float eps, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  if (!mean || !rstd) return fail(SAR_EINVAL, "layernorm_stats: null mean / rstd");
  return layernorm_fwd(x, gamma, beta, y, mean, rstd, M, d, eps, static_cast<cudaStream_t>(stream));
}

int sar_layernorm_lora_u_fwd(const void* h, const void* gamma, const void* beta, void* x, const void* A_cat,
                             const int32_t* utt_adapter, void* u, int B, int T, int d, int r, int n_sets,
                             int n_adapters, float scale, float eps, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  return ln_lora_u_fwd(h, gamma, beta, x, A_cat, utt_adapter, u, B, T, d, r, n_sets, n_adapters, scale, eps,
                       static_cast<cudaStream_t>(stream));
}

int sar_layernorm_lora_u_supported(int d, int r, int n_sets) { return ln_lora_u_supported(d, r, n_sets) ? 1 : 0; }

int sar_operand_refresh(const sar_refresh_desc* desc, int n_desc, int max_elems, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  return operand_refresh(desc, n_desc, max_elems, static_cast<cudaStream_t>(stream));
}

int sar_qv_lora_fwd_rows(const void* x, const void* W, const void* bias, const void* A_stack, const void* Bp_stack,
                         const int32_t* row_adapter, void* y, int M, int d_in, int d_out, int r, int n_adapters,
                         float scale, void* ws, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  return rows_qv_lora_fwd(x, W, bias, A_stack, Bp_stack, row_adapter, y, M, d_in, d_out, r, n_adapters, scale, ws,
                          static_cast<cudaStream_t>(stream));
}

int sar_router_fwd(const void* h, int h_is_fp32, const float* ln_w, const float* ln_b, const float* W1,
                   const float* b1, const float* g1, const float* be1, const float* W2, const float* b2,
                   const float* g2, const float* be2, const float* W3, const float* b3, int B, int T, int d, int h1,
                   int h2, int C, float* logits_out, float* probs_out, int32_t* idx_out, int32_t* perm_out,
                   int32_t* seg_starts_out, void* ws, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  K2Args a{};
  a.h = h; a.h_is_fp32 = h_is_fp32;
  a.ln_w = ln_w; a.ln_b = ln_b; a.W1 = W1; a.b1 = b1; a.g1 = g1; a.be1 = be1;
  a.W2 = W2; a.b2 = b2; a.g2 = g2; a.be2 = be2; a.W3 = W3; a.b3 = b3;
  a.B = B; a.T = T; a.d = d; a.h1 = h1; a.h2 = h2; a.C = C;
  a.logits = logits_out; a.probs = probs_out; a.idx = idx_out; a.perm = perm_out; a.seg_starts = seg_starts_out;
  a.ws = ws;
  return k2_router_fwd(a, static_cast<cudaStream_t>(stream));
}

int sar_router_fwd_fused_ln(const void* h_pre, const void* enc_ln_w, const void* enc_ln_b, float enc_ln_eps,
                            const float* ln_w, const float* ln_b, const float* W1, const float* b1, const float* g1,
                            const float* be1, const float* W2, const float* b2, const float* g2, const float* be2,
                            const float* W3, const float* b3, int B, int T, int d, int h1, int h2, int C,
                            float* logits_out, float* probs_out, int32_t* idx_out, int32_t* perm_out,
                            int32_t* seg_starts_out, void* ws, void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  if (!enc_ln_w || !enc_ln_b) return fail(SAR_EINVAL, "sar_router_fwd_fused_ln: null encoder LayerNorm parameters");
  K2Args a{};
  a.h = h_pre; a.h_is_fp32 = 0;
  a.pre_ln_w = enc_ln_w; a.pre_ln_b = enc_ln_b; a.pre_ln_eps = enc_ln_eps;
  a.ln_w = ln_w; a.ln_b = ln_b; a.W1 = W1; a.b1 = b1; a.g1 = g1; a.be1 = be1;
  a.W2 = W2; a.b2 = b2; a.g2 = g2; a.be2 = be2; a.W3 = W3; a.b3 = b3;
  a.B = B; a.T = T; a.d = d; a.h1 = h1; a.h2 = h2; a.C = C;
  a.logits = logits_out; a.probs = probs_out; a.idx = idx_out; a.perm = perm_out; a.seg_starts = seg_starts_out;
  a.ws = ws;
  return k2_router_fwd(a, static_cast<cudaStream_t>(stream));
}

int sar_qv_lora_bwd(const void* dy, const void* x, const void* u, const void* Wt, const void* At_stack,
                    const void* Bt_stack, const void* Bp_stack, const int32_t* utt_adapter, void* dx, float* dA,
                    float* dB, int B, int T, int d_in, int d_out, int r, int n_adapters, float scale, void* ws,
                    void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  K3Args a{};
  a.dy = dy; a.x = x; a.u = u; a.Wt = Wt; a.At_stack = At_stack; a.Bt_stack = Bt_stack; a.Bp_stack = Bp_stack;
  a.utt_adapter = utt_adapter; a.dx = dx; a.dA = dA; a.dB = dB;
  a.B = B; a.T = T; a.d_in = d_in; a.d_out = d_out; a.r = r; a.n_adapters = n_adapters; a.scale = scale; a.ws = ws;
  return k3_qv_lora_bwd(a, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
