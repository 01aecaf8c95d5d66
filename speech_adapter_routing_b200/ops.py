"""Torch-facing wrappers over the libsar C ABI.  PyTorch only provides device memory and the stream; all
arithmetic of the hot path happens inside libsar's sm_100a kernels.  Every wrapper raises if libsar is missing or
the tensors are not CUDA tensors — there is no eager fallback.
"""
from __future__ import annotations

import ctypes
from typing import List, NamedTuple, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import SAR_FLAG_SAVE_U, SAR_RPAD, check, lib


# ---- instrumentation used by bench.py: kernel-launch counts and (optional) per-launch CUDA-event timing ----------
LAUNCHES = {"k1": 0, "k2": 0, "k3": 0, "rows": 0, "proj": 0, "linear": 0, "ln": 0, "attn": 0, "logmel": 0, "refresh": 0}   # kernels launched by libsar, by op (k2 = 2, k3 = 3, rows = 2)
SPLIT_MIN_ROWS = int(__import__("os").environ.get("SAR_SPLIT_MIN_ROWS", "4096"))   # fewer rows: single-launch LoRA kernel (M = 8192, 768 -> 2304: split 79 us, single launch 91 us)
K1_TIMELINE = None   # set to a list to record (B*T, d_in, d_out, r, has_lora, start_event, end_event) per K1 call


def reset_counters() -> None:
    for k in LAUNCHES:
        LAUNCHES[k] = 0


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _need_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("libsar ops need CUDA tensors (sm_100a only, no CPU fallback)")


def _bf16c(t: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.bfloat16:
        raise TypeError(f"{name} must be torch.bfloat16, got {t.dtype}")
    return t.contiguous()


def pack_lora_b(B_stack: torch.Tensor) -> torch.Tensor:
    """[n_adapters, d_out, r] -> bf16 [n_adapters, d_out, SAR_RPAD] with zero rank padding (kernel operand layout)."""
    n, d_out, r = B_stack.shape
    if r > SAR_RPAD:
        raise ValueError(f"rank {r} > {SAR_RPAD} is not supported")
    out = torch.zeros(n, d_out, SAR_RPAD, dtype=torch.bfloat16, device=B_stack.device)
    out[:, :, :r] = B_stack.to(torch.bfloat16)
    return out


def pad_rank16(A_stack: torch.Tensor, B_stack: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """Zero-pad the rank to a multiple of 16 (MMA K granularity).  A [n,r,d_in], B [n,d_out,r]."""
    r = A_stack.shape[1]
    rp = (r + 15) // 16 * 16
    if rp == r:
        return A_stack, B_stack, r
    A2 = torch.zeros(A_stack.shape[0], rp, A_stack.shape[2], dtype=A_stack.dtype, device=A_stack.device)
    A2[:, :r] = A_stack
    B2 = torch.zeros(B_stack.shape[0], B_stack.shape[1], rp, dtype=B_stack.dtype, device=B_stack.device)
    B2[:, :, :r] = B_stack
    return A2, B2, rp


def qv_lora_fwd(x: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], A_stack: Optional[torch.Tensor],
                Bp_stack: Optional[torch.Tensor], utt_adapter: Optional[torch.Tensor], scale: float,
                save_u: bool = False, block_n: int = 0, grid: int = 0,
                out: Optional[torch.Tensor] = None, kernel: int = 0,
                swap_halves: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """K1: y = x·Wᵀ + bias + (scale·x·A_kᵀ)·B_kᵀ with k = utt_adapter[b].  x is [B, T, d_in] bf16."""
    _need_cuda(x, W, bias, A_stack, Bp_stack, utt_adapter)
    if x.dim() != 3:
        raise ValueError("x must be [B, T, d_in]")
    x = _bf16c(x, "x"); W = _bf16c(W, "W"); bias = _bf16c(bias, "bias")
    A_stack = _bf16c(A_stack, "A_stack"); Bp_stack = _bf16c(Bp_stack, "Bp_stack")
    B, T, d_in = x.shape
    d_out = W.shape[0]
    if W.shape[1] != d_in:
        raise ValueError("W must be [d_out, d_in]")
    n_adapters, r = 0, 16
    if A_stack is not None and utt_adapter is not None:
        n_adapters, r = A_stack.shape[0], A_stack.shape[1]
        if A_stack.shape[2] != d_in or tuple(Bp_stack.shape) != (n_adapters, d_out, SAR_RPAD):
            raise ValueError("A_stack must be [n,r,d_in] and Bp_stack [n,d_out,64]")
        if utt_adapter.dtype != torch.int32 or utt_adapter.numel() != B:
            raise ValueError("utt_adapter must be int32 [B]")
        utt_adapter = utt_adapter.contiguous()
    y = out if out is not None else torch.empty(B, T, d_out, dtype=torch.bfloat16, device=x.device)
    u = torch.empty(B * T, r, dtype=torch.bfloat16, device=x.device) if (save_u and n_adapters) else None
    flags = ((SAR_FLAG_SAVE_U if u is not None else 0) | ((block_n & 0x3FF) << 8) | ((grid & 0x3FF) << 18) |
             ((kernel & 0x3) << 28) | (2 if swap_halves else 0))
    def launch():
        check(lib().sar_qv_lora_fwd(_ptr(x), _ptr(W), _ptr(bias), _ptr(A_stack), _ptr(Bp_stack),
                                    _ptr(utt_adapter) if n_adapters else None, _ptr(y), _ptr(u), B, T, d_in, d_out, r,
                                    n_adapters, float(scale), flags, _stream(x)))
    flops = 2.0 * B * T * d_in * d_out + (2.0 * B * T * r * (d_in + d_out) if n_adapters else 0.0)
    _time_k1(K1_TIMELINE, "k1", B * T, d_in, d_out, flops, launch)
    LAUNCHES["k1"] += 1
    return y, u


def _time_k1(tl, kind, M, d_in, d_out, flops, fn):
    """Run ``fn`` (one tcgen05 GEMM launch) bracketed by CUDA events when bench.py asked for a timeline.
    ``flops`` = ALGORITHMIC flops of the call: 2·M·d_in·d_out + 2·M·r·(d_in + d_out) per LoRA'd segment."""
    if tl is None:
        fn()
        return
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record()
    fn()
    ev1.record()
    tl.append((kind, M, d_in, d_out, flops, ev0, ev1))


def attn_proj_fwd(x: torch.Tensor, W_cat: torch.Tensor, bias_cat: Optional[torch.Tensor],
                  A_cat: Optional[torch.Tensor], Bp_cat: Optional[torch.Tensor], utt_adapter: Optional[torch.Tensor],
                  seg_set: Sequence[int], seg_scale: Sequence[float], n_sets: int, scale: float,
                  x_head_major: bool = False, y_head_major: bool = True, block_n: int = 0,
                  grid: int = 0, split: Optional[bool] = None, u: Optional[torch.Tensor] = None,
                  mix_w: Optional[torch.Tensor] = None, mix_group_rank: int = 0) -> List[torch.Tensor]:
    """Fused attention projections (sar_attn_proj_fwd): up to three projections of the same x in one launch.

    x [B,T,d_in] (or [B,d_in/64,T,64] if ``x_head_major``); W_cat [n_seg*d_out, d_in]; A_cat [n_sets*n, r, d_in];
    Bp_cat [n_sets*n, d_out, 64].  Returns n_seg tensors, [B,d_out/64,T,64] if ``y_head_major`` else [B,T,d_out].
    ``split=True``: U = scale·x·A_kᵀ goes through a [B,T,64*n_sets] workspace and the projections run on the dense
    256-wide kernel with one extra K block; ``split=False``: the single-launch kernel that keeps U in shared memory;
    ``None`` (default): split from SPLIT_MIN_ROWS rows up (bit-identical results either way).
    ``u``: U already computed for this x ([n_sets, B, T, r], e.g. by ``layernorm_lora_u_fwd``): only the dense launch runs.
    ``mix_w`` (fp32 [B, groups]) with ``mix_group_rank``: weighted mix of ``groups`` adapters stacked along the rank of ONE
    merged adapter (sar_attn_proj_fwd_mix): rank columns of group g of utterance b are scaled by mix_w[b, g].
    """
    _need_cuda(x, W_cat, bias_cat, A_cat, Bp_cat, utt_adapter)
    x = _bf16c(x, "x"); W_cat = _bf16c(W_cat, "W_cat"); bias_cat = _bf16c(bias_cat, "bias_cat")
    A_cat = _bf16c(A_cat, "A_cat"); Bp_cat = _bf16c(Bp_cat, "Bp_cat")
    n_seg = len(seg_set)
    if x_head_major:
        B, hh, T, hd = x.shape
        d_in = hh * hd
    else:
        B, T, d_in = x.shape
    if W_cat.shape[1] != d_in or W_cat.shape[0] % n_seg:
        raise ValueError("W_cat must be [n_seg*d_out, d_in]")
    d_out = W_cat.shape[0] // n_seg
    n_adapters, r = 0, 16
    if A_cat is not None and utt_adapter is not None and any(s >= 0 for s in seg_set):
        if A_cat.shape[0] % n_sets:
            raise ValueError("A_cat must be [n_sets*n_adapters, r, d_in]")
        n_adapters, r = A_cat.shape[0] // n_sets, A_cat.shape[1]
        if A_cat.shape[2] != d_in or tuple(Bp_cat.shape) != (n_sets * n_adapters, d_out, SAR_RPAD):
            raise ValueError("A_cat must be [n_sets*n,r,d_in] and Bp_cat [n_sets*n,d_out,64]")
        if utt_adapter.dtype != torch.int32 or utt_adapter.numel() != B:
            raise ValueError("utt_adapter must be int32 [B]")
        utt_adapter = utt_adapter.contiguous()
    if y_head_major:
        if d_out % 64:
            raise ValueError("head-major output needs d_out % 64 == 0")
        ys = [torch.empty(B, d_out // 64, T, 64, dtype=torch.bfloat16, device=x.device) for _ in range(n_seg)]
    else:
        ys = [torch.empty(B, T, d_out, dtype=torch.bfloat16, device=x.device) for _ in range(n_seg)]
    yp = (ctypes.c_void_p * n_seg)(*[y.data_ptr() for y in ys])
    ss = (ctypes.c_int32 * n_seg)(*[int(s) for s in seg_set])
    sc = (ctypes.c_float * n_seg)(*[float(s) for s in seg_scale])
    flags = ((block_n & 0x3FF) << 8) | ((grid & 0x3FF) << 18)
    has_lora = n_adapters > 0
    ws = None
    if has_lora and u is not None:
        if u.dtype != torch.bfloat16 or not u.is_contiguous() or u.numel() != n_sets * B * T * r:
            raise ValueError("u must be contiguous bf16 [n_sets, B, T, r]")
        ws = u
        flags |= _lib.SAR_FLAG_U_READY
    elif has_lora and (mix_w is not None or (split if split is not None else B * T >= SPLIT_MIN_ROWS)):
        ws = torch.empty(n_sets * B * T * r, dtype=torch.bfloat16, device=x.device)   # U: [n_sets][B, T, r]
    if mix_w is not None and has_lora:
        if u is not None:
            raise ValueError("mix_w and a precomputed u are mutually exclusive")
        if mix_w.dtype != torch.float32 or not mix_w.is_contiguous() or mix_w.shape[0] != B or not mix_w.is_cuda:
            raise ValueError("mix_w must be contiguous fp32 CUDA [B, groups]")
        if mix_group_rank <= 0 or mix_w.shape[1] * mix_group_rank != r:
            raise ValueError("mix_w.shape[1] * mix_group_rank must equal the merged rank r")

    def launch():
        common = (_ptr(x), int(x_head_major), _ptr(W_cat), _ptr(bias_cat), _ptr(A_cat) if has_lora else None,
                  _ptr(Bp_cat) if has_lora else None, _ptr(utt_adapter) if has_lora else None, yp, ss, sc, n_seg,
                  n_sets if has_lora else 1, int(y_head_major), B, T, d_in, d_out, r, n_adapters, float(scale), flags,
                  _ptr(ws))
        if mix_w is not None and has_lora:
            check(lib().sar_attn_proj_fwd_mix(*common, _ptr(mix_w), int(mix_w.shape[1]), int(mix_group_rank), _stream(x)))
        else:
            check(lib().sar_attn_proj_fwd(*common, _stream(x)))
    n_lora = sum(1 for s in seg_set if s >= 0) if has_lora else 0
    # algorithmic flops of THIS call: with U precomputed (fused into the LayerNorm kernel) the down-projection
    # 2·M·r·d_in per set is not done here and is not credited to this kernel
    flops = 2.0 * B * T * d_in * d_out * n_seg + 2.0 * B * T * r * ((0 if u is not None else n_sets * d_in) +
                                                                     n_lora * d_out) * (n_lora > 0)
    _time_k1(K1_TIMELINE, "proj", B * T, d_in, n_seg * d_out, flops, launch)
    LAUNCHES["proj"] += 2 if (ws is not None and u is None) else 1
    return ys


def lora_u_fwd(x: torch.Tensor, A_cat: torch.Tensor, utt_adapter: torch.Tensor, n_sets: int, scale: float,
               d_out: int) -> torch.Tensor:
    """The U pass of the split path alone (SAR_FLAG_U_ONLY): U = scale·x·A_kᵀ, bf16 [n_sets, B, T, r], on the tcgen05
    pair kernel.  ``d_out`` only sizes tensor maps that the pass never touches."""
    _need_cuda(x, A_cat, utt_adapter)
    x = _bf16c(x, "x"); A_cat = _bf16c(A_cat, "A_cat")
    B, T, d_in = x.shape
    n_adapters, r = A_cat.shape[0] // n_sets, A_cat.shape[1]
    u = torch.empty(n_sets, B, T, r, dtype=torch.bfloat16, device=x.device)
    yp = (ctypes.c_void_p * 1)(u.data_ptr())
    ss = (ctypes.c_int32 * 1)(0)
    sc = (ctypes.c_float * 1)(1.0)
    dummy = torch.empty(16, dtype=torch.bfloat16, device=x.device)
    check(lib().sar_attn_proj_fwd(_ptr(x), 0, _ptr(dummy), None, _ptr(A_cat), _ptr(dummy), _ptr(utt_adapter.contiguous()),
                                  yp, ss, sc, 1, n_sets, 0, B, T, d_in, d_out, r, n_adapters, float(scale),
                                  _lib.SAR_FLAG_U_ONLY, _ptr(u), _stream(x)))
    LAUNCHES["proj"] += 1
    return u


def attn_proj_fwd_rows(x: torch.Tensor, W_cat: torch.Tensor, bias_cat: Optional[torch.Tensor],
                       A_cat: Optional[torch.Tensor], Bp_cat: Optional[torch.Tensor], row_adapter: Optional[torch.Tensor],
                       seg_set: Sequence[int], seg_scale: Sequence[float], n_sets: int, scale: float) -> List[torch.Tensor]:
    """Decode-step form of ``attn_proj_fwd`` (sar_attn_proj_fwd_rows): x [M, d_in], one token per utterance, adapter
    ``row_adapter[m]`` per row.  Returns n_seg tensors [M, d_out]."""
    _need_cuda(x, W_cat, bias_cat, A_cat, Bp_cat, row_adapter)
    x = _bf16c(x, "x"); W_cat = _bf16c(W_cat, "W_cat"); bias_cat = _bf16c(bias_cat, "bias_cat")
    A_cat = _bf16c(A_cat, "A_cat"); Bp_cat = _bf16c(Bp_cat, "Bp_cat")
    n_seg = len(seg_set)
    M, d_in = x.shape
    d_out = W_cat.shape[0] // n_seg
    n_adapters, r = 0, 16
    if A_cat is not None and row_adapter is not None and any(s >= 0 for s in seg_set):
        n_adapters, r = A_cat.shape[0] // n_sets, A_cat.shape[1]
        if row_adapter.dtype != torch.int32 or row_adapter.numel() != M:
            raise ValueError("row_adapter must be int32 [M]")
        row_adapter = row_adapter.contiguous()
    ys = [torch.empty(M, d_out, dtype=torch.bfloat16, device=x.device) for _ in range(n_seg)]
    yp = (ctypes.c_void_p * n_seg)(*[y.data_ptr() for y in ys])
    ss = (ctypes.c_int32 * n_seg)(*[int(s) for s in seg_set])
    sc = (ctypes.c_float * n_seg)(*[float(s) for s in seg_scale])
    has_lora = n_adapters > 0
    check(lib().sar_attn_proj_fwd_rows(_ptr(x), _ptr(W_cat), _ptr(bias_cat), _ptr(A_cat) if has_lora else None,
                                       _ptr(Bp_cat) if has_lora else None, _ptr(row_adapter) if has_lora else None,
                                       yp, ss, sc, n_seg, n_sets if has_lora else 1, M, d_in, d_out, r, n_adapters,
                                       float(scale), 0, _stream(x)))
    LAUNCHES["proj"] += 2 if has_lora else 1
    return ys


def linear_fwd(x: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], residual: Optional[torch.Tensor] = None,
               act: int = _lib.SAR_ACT_NONE, x_head_major: bool = False, out: Optional[torch.Tensor] = None,
               block_n: int = 0, grid: int = 0) -> torch.Tensor:
    """y = act(x·Wᵀ + bias) + residual on the tcgen05 pair kernel (sar_linear_fwd).  x [B,T,d_in] or head-major
    [B,d_in/64,T,64]; residual / y [B,T,d_out] (``out`` may be the residual tensor: in-place residual update)."""
    _need_cuda(x, W, bias, residual)
    x = _bf16c(x, "x"); W = _bf16c(W, "W"); bias = _bf16c(bias, "bias"); residual = _bf16c(residual, "residual")
    if x_head_major:
        B, hh, T, hd = x.shape
        d_in = hh * hd
    else:
        B, T, d_in = x.shape
    d_out = W.shape[0]
    if W.shape[1] != d_in:
        raise ValueError("W must be [d_out, d_in]")
    if residual is not None and tuple(residual.shape) != (B, T, d_out):
        raise ValueError("residual must be [B, T, d_out]")
    y = out if out is not None else torch.empty(B, T, d_out, dtype=torch.bfloat16, device=x.device)
    flags = ((block_n & 0x3FF) << 8) | ((grid & 0x3FF) << 18)

    def launch():
        check(lib().sar_linear_fwd(_ptr(x), int(x_head_major), _ptr(W), _ptr(bias), _ptr(residual), _ptr(y), B, T,
                                   d_in, d_out, int(act), flags, _stream(x)))
    _time_k1(K1_TIMELINE, "linear", B * T, d_in, d_out, 2.0 * B * T * d_in * d_out, launch)
    LAUNCHES["linear"] += 1
    return y


def dense_fwd(x_base: torch.Tensor, ldx: int, x_batch_stride: int, W: torch.Tensor, bias: Optional[torch.Tensor],
              y_base: torch.Tensor, ldy: int, y_batch_stride: int, B: int, T: int, d_in: int, d_out: int,
              act: int = _lib.SAR_ACT_NONE, residual: Optional[torch.Tensor] = None, ldr: int = 0,
              res_batch_stride: int = 0, res_broadcast: bool = False, block_n: int = 0, grid: int = 0) -> None:
    """General strided dense layer (sar_dense_fwd): y[b,t,:] = act(x[b,t,:]·Wᵀ + bias) + residual[b,t,:] where row
    (b, t) of x starts at ``x_base.data_ptr() + 2*(b*x_batch_stride + t*ldx)`` (rows may overlap) and likewise for y /
    residual.  ``x_base`` / ``y_base`` are tensors whose first element is the origin; the caller owns bounds."""
    _need_cuda(x_base, W, bias, y_base, residual)
    for t, name in ((x_base, "x"), (W, "W"), (bias, "bias"), (y_base, "y"), (residual, "residual")):
        if t is not None and t.dtype != torch.bfloat16:
            raise TypeError(f"{name} must be torch.bfloat16")
    if not W.is_contiguous() or tuple(W.shape) != (d_out, d_in):
        raise ValueError("W must be contiguous [d_out, d_in]")
    flags = ((block_n & 0x3FF) << 8) | ((grid & 0x3FF) << 18)

    def launch():
        check(lib().sar_dense_fwd(_ptr(x_base), int(ldx), int(x_batch_stride), _ptr(W), _ptr(bias), _ptr(residual),
                                  int(ldr), int(res_batch_stride), int(res_broadcast), _ptr(y_base), int(ldy),
                                  int(y_batch_stride), B, T, d_in, d_out, int(act), flags, _stream(x_base)))
    _time_k1(K1_TIMELINE, "dense", B * T, d_in, d_out, 2.0 * B * T * d_in * d_out, launch)
    LAUNCHES["linear"] += 1


def attn_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool = False) -> torch.Tensor:
    """softmax(q·kᵀ)·v (sar_attn_fwd): q [B,H,Tq,64] (pre-scaled), k / v [B,H,Tk,64], bf16 contiguous -> [B,H,Tq,64]."""
    _need_cuda(q, k, v)
    q = _bf16c(q, "q"); k = _bf16c(k, "k"); v = _bf16c(v, "v")
    B, H, Tq, hd = q.shape
    Tk = k.shape[2]
    if tuple(k.shape) != (B, H, Tk, hd) or tuple(v.shape) != (B, H, Tk, hd):
        raise ValueError("k and v must be [B, H, Tk, 64]")
    out = torch.empty_like(q)
    check(lib().sar_attn_fwd(_ptr(q), _ptr(k), _ptr(v), _ptr(out), B, H, Tq, Tk, hd, int(causal), _stream(q)))
    LAUNCHES["attn"] += 1
    return out


def decode_self_attn(q: torch.Tensor, k_new: torch.Tensor, v_new: torch.Tensor, cache_k: torch.Tensor,
                     cache_v: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
    """One decode step of self-attention over the static cache (sar_decode_self_attn): writes k_new / v_new at ``pos``
    (int64 device scalar) and returns softmax(q·Kᵀ)·V over positions 0..pos.  q / k_new / v_new: bf16 [B, H, 64] (any
    contiguous view with B*H*64 elements); caches bf16 [B, H, Tmax, 64].  Returns bf16 [B, H*64]."""
    _need_cuda(q, k_new, v_new, cache_k, cache_v, pos)
    B, H, Tmax, hd = cache_k.shape
    for t, name in ((q, "q"), (k_new, "k_new"), (v_new, "v_new")):
        if t.dtype != torch.bfloat16 or not t.is_contiguous() or t.numel() != B * H * hd:
            raise ValueError(f"{name} must be contiguous bf16 with B*H*64 elements")
    if cache_k.dtype != torch.bfloat16 or not cache_k.is_contiguous() or not cache_v.is_contiguous():
        raise ValueError("caches must be contiguous bf16 [B, H, Tmax, 64]")
    if pos.dtype != torch.int64 or pos.numel() != 1:
        raise ValueError("pos must be an int64 device scalar")
    out = torch.empty(B, H * hd, dtype=torch.bfloat16, device=q.device)
    check(lib().sar_decode_self_attn(_ptr(q), _ptr(k_new), _ptr(v_new), _ptr(cache_k), _ptr(cache_v), _ptr(pos),
                                     _ptr(out), B, H, hd, Tmax, _stream(q)))
    LAUNCHES["attn"] += 1
    return out


def decode_cross_attn(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """Cross-attention of one decode step (sar_decode_cross_attn): q bf16 with B*H*64 elements (pre-scaled), k / v bf16
    [B, H, Tk, 64] contiguous.  Returns bf16 [B, H*64]."""
    _need_cuda(q, k, v)
    B, H, Tk, hd = k.shape
    if q.dtype != torch.bfloat16 or not q.is_contiguous() or q.numel() != B * H * hd:
        raise ValueError("q must be contiguous bf16 with B*H*64 elements")
    for t, name in ((k, "k"), (v, "v")):
        if t.dtype != torch.bfloat16 or not t.is_contiguous() or t.shape != k.shape:
            raise ValueError(f"{name} must be contiguous bf16 [B, H, Tk, 64]")
    out = torch.empty(B, H * hd, dtype=torch.bfloat16, device=q.device)
    check(lib().sar_decode_cross_attn(_ptr(q), _ptr(k), _ptr(v), _ptr(out), B, H, hd, Tk, _stream(q)))
    LAUNCHES["attn"] += 1
    return out


def logmel_fwd(wave: torch.Tensor, window: torch.Tensor, cos_table: torch.Tensor, sin_table: torch.Tensor,
               mel_filters: torch.Tensor, out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """Log-mel front-end (sar_logmel_fwd): wave fp32 [B, n_samples] (n_samples a multiple of 160, already padded / cut),
    tables fp32 on the same device; returns [B, n_mels, n_samples / 160] in bf16 or fp32."""
    _need_cuda(wave, window, cos_table, sin_table, mel_filters)
    for t, name in ((wave, "wave"), (window, "window"), (cos_table, "cos_table"), (sin_table, "sin_table"),
                    (mel_filters, "mel_filters")):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError(f"{name} must be contiguous fp32")
    if wave.dim() != 2 or wave.shape[1] % 160 or window.numel() != 400 or cos_table.numel() != 400 or sin_table.numel() != 400:
        raise ValueError("wave must be [B, n_samples] with n_samples % 160 == 0; window / tables have 400 entries")
    if mel_filters.dim() != 2 or mel_filters.shape[0] != 201:
        raise ValueError("mel_filters must be [201, n_mels]")
    if out_dtype not in (torch.bfloat16, torch.float32):
        raise ValueError("out_dtype must be bfloat16 or float32")
    B, n_samples = wave.shape
    n_mels, n_frames = mel_filters.shape[1], n_samples // 160
    raw = torch.empty(B, n_mels, n_frames, dtype=torch.float32, device=wave.device)
    clip_max = torch.empty(B, dtype=torch.int32, device=wave.device)
    out = torch.empty(B, n_mels, n_frames, dtype=out_dtype, device=wave.device)
    check(lib().sar_logmel_fwd(_ptr(wave), _ptr(window), _ptr(cos_table), _ptr(sin_table), _ptr(mel_filters), _ptr(raw),
                               _ptr(clip_max), _ptr(out), B, n_samples, n_mels, int(out_dtype == torch.bfloat16),
                               _stream(wave)))
    LAUNCHES["logmel"] += 3
    return out


def layernorm_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LayerNorm over the last dim (sar_layernorm_fwd): bf16 in/out, fp32 statistics."""
    _need_cuda(x, gamma, beta)
    x = _bf16c(x, "x"); gamma = _bf16c(gamma, "gamma"); beta = _bf16c(beta, "beta")
    d = x.shape[-1]
    M = x.numel() // d
    y = out if out is not None else torch.empty_like(x)
    check(lib().sar_layernorm_fwd(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(y), M, d, float(eps), _stream(x)))
    LAUNCHES["ln"] += 1
    return y


def layernorm_fwd_stats(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                        eps: float = 1e-5) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """LayerNorm that also returns the row statistics (sar_layernorm_fwd_stats): y bf16 like x, mean / rstd fp32
    [..., 1] — the three results of ``torch.native_layer_norm``, for ATen's LayerNorm backward."""
    _need_cuda(x, gamma, beta)
    x = _bf16c(x, "x"); gamma = _bf16c(gamma, "gamma"); beta = _bf16c(beta, "beta")
    d = x.shape[-1]
    M = x.numel() // d
    y = torch.empty_like(x)
    mean = torch.empty(*x.shape[:-1], 1, dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    check(lib().sar_layernorm_fwd_stats(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(y), _ptr(mean), _ptr(rstd), M, d, float(eps),
                                        _stream(x)))
    LAUNCHES["ln"] += 1
    return y, mean, rstd


def layernorm_lora_u_supported(d: int, r: int, n_sets: int) -> bool:
    return bool(lib().sar_layernorm_lora_u_supported(int(d), int(r), int(n_sets)))


def layernorm_lora_u_fwd(h: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, A_cat: torch.Tensor,
                         utt_adapter: torch.Tensor, n_sets: int, scale: float,
                         eps: float = 1e-5) -> Tuple[torch.Tensor, torch.Tensor]:
    """x = LayerNorm(h) and U = scale·x·A_kᵀ for every LoRA set in one pass over h (sar_layernorm_lora_u_fwd).
    h [B,T,d] bf16; A_cat [n_sets*n_adapters, r, d] bf16; returns (x [B,T,d], U [n_sets,B,T,r])."""
    _need_cuda(h, gamma, beta, A_cat, utt_adapter)
    h = _bf16c(h, "h"); gamma = _bf16c(gamma, "gamma"); beta = _bf16c(beta, "beta"); A_cat = _bf16c(A_cat, "A_cat")
    if h.dim() != 3:
        raise ValueError("h must be [B, T, d]")
    B, T, d = h.shape
    if A_cat.shape[0] % n_sets or A_cat.shape[2] != d:
        raise ValueError("A_cat must be [n_sets*n_adapters, r, d]")
    n_adapters, r = A_cat.shape[0] // n_sets, A_cat.shape[1]
    if utt_adapter.dtype != torch.int32 or utt_adapter.numel() != B:
        raise ValueError("utt_adapter must be int32 [B]")
    x = torch.empty_like(h)
    u = torch.empty(n_sets, B, T, r, dtype=torch.bfloat16, device=h.device)
    check(lib().sar_layernorm_lora_u_fwd(_ptr(h), _ptr(gamma), _ptr(beta), _ptr(x), _ptr(A_cat),
                                         _ptr(utt_adapter.contiguous()), _ptr(u), B, T, d, r, n_sets, n_adapters,
                                         float(scale), float(eps), _stream(h)))
    LAUNCHES["ln"] += 1
    return x, u


def qv_lora_fwd_rows(x: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], A_stack: Optional[torch.Tensor],
                     Bp_stack: Optional[torch.Tensor], row_adapter: Optional[torch.Tensor],
                     scale: float) -> torch.Tensor:
    """Row-indexed K1 variant (decode steps): x [M, d_in], row_adapter int32 [M]."""
    _need_cuda(x, W, bias, A_stack, Bp_stack, row_adapter)
    x = _bf16c(x, "x"); W = _bf16c(W, "W"); bias = _bf16c(bias, "bias")
    A_stack = _bf16c(A_stack, "A_stack"); Bp_stack = _bf16c(Bp_stack, "Bp_stack")
    M, d_in = x.shape
    d_out = W.shape[0]
    n_adapters, r = 0, 16
    if A_stack is not None and row_adapter is not None:
        n_adapters, r = A_stack.shape[0], A_stack.shape[1]
        if row_adapter.dtype != torch.int32 or row_adapter.numel() != M:
            raise ValueError("row_adapter must be int32 [M]")
        row_adapter = row_adapter.contiguous()
    y = torch.empty(M, d_out, dtype=torch.bfloat16, device=x.device)
    check(lib().sar_qv_lora_fwd_rows(_ptr(x), _ptr(W), _ptr(bias), _ptr(A_stack), _ptr(Bp_stack),
                                     _ptr(row_adapter) if n_adapters else None, _ptr(y), M, d_in, d_out, r,
                                     n_adapters, float(scale), None, _stream(x)))
    LAUNCHES["rows"] += 2 if n_adapters else 1
    return y


class RouterOut(NamedTuple):
    logits: torch.Tensor      # fp32 [B, C]
    probs: torch.Tensor       # fp32 [B, C]
    idx: torch.Tensor         # int32 [B]  argmax class = adapter index
    perm: torch.Tensor        # int32 [B]  utterances stably sorted by idx
    seg_starts: torch.Tensor  # int32 [C+1]


class RouterParams(NamedTuple):
    """fp32 contiguous CUDA tensors of the default LanguageClassifier (reference state-dict keys in comments)."""
    ln_w: torch.Tensor   # layer_norm.weight
    ln_b: torch.Tensor   # layer_norm.bias
    W1: torch.Tensor     # classifier.0.weight
    b1: torch.Tensor     # classifier.0.bias
    g1: torch.Tensor     # classifier.1.weight
    be1: torch.Tensor    # classifier.1.bias
    W2: torch.Tensor     # classifier.4.weight
    b2: torch.Tensor     # classifier.4.bias
    g2: torch.Tensor     # classifier.5.weight
    be2: torch.Tensor    # classifier.5.bias
    W3: torch.Tensor     # classifier.8.weight
    b3: torch.Tensor     # classifier.8.bias

    @staticmethod
    def from_state_dict(sd, device) -> "RouterParams":
        keys = ["layer_norm.weight", "layer_norm.bias", "classifier.0.weight", "classifier.0.bias",
                "classifier.1.weight", "classifier.1.bias", "classifier.4.weight", "classifier.4.bias",
                "classifier.5.weight", "classifier.5.bias", "classifier.8.weight", "classifier.8.bias"]
        return RouterParams(*[sd[k].detach().to(device=device, dtype=torch.float32).contiguous() for k in keys])


def router_fwd(h: torch.Tensor, p: RouterParams, pre_ln: Optional[Tuple[torch.Tensor, torch.Tensor, float]] = None) -> RouterOut:
    """K2: LayerNorm → mean over T → MLP → softmax → argmax → (idx, perm, seg_starts).  h is [B,T,d] bf16 or fp32.
    ``pre_ln = (gamma, beta, eps)`` (bf16): h is the encoder's residual stream BEFORE its final LayerNorm, which the
    kernel applies on the fly (sar_router_fwd_fused_ln)."""
    _need_cuda(h, *p)
    if h.dim() != 3:
        raise ValueError("h must be [B, T, d]")
    if h.dtype not in (torch.bfloat16, torch.float32):
        raise TypeError("h must be bf16 or fp32")
    h = h.contiguous()
    B, T, d = h.shape
    h1, h2, C = p.W1.shape[0], p.W2.shape[0], p.W3.shape[0]
    dev = h.device
    logits = torch.empty(B, C, dtype=torch.float32, device=dev)
    probs = torch.empty(B, C, dtype=torch.float32, device=dev)
    idx = torch.empty(B, dtype=torch.int32, device=dev)
    perm = torch.empty(B, dtype=torch.int32, device=dev)
    seg = torch.empty(C + 1, dtype=torch.int32, device=dev)
    nbytes = lib().sar_workspace_bytes(_lib.SAR_OP_ROUTER_FWD, B, T, d, 0, C)
    if nbytes < 0:
        check(int(nbytes))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    if pre_ln is not None:
        g1, b1, eps1 = pre_ln
        if h.dtype != torch.bfloat16:
            raise TypeError("the fused encoder LayerNorm needs bf16 states")
        g1, b1 = _bf16c(g1, "pre_ln gamma"), _bf16c(b1, "pre_ln beta")
        check(lib().sar_router_fwd_fused_ln(_ptr(h), _ptr(g1), _ptr(b1), float(eps1), *[_ptr(t) for t in p], B, T, d, h1,
                                            h2, C, _ptr(logits), _ptr(probs), _ptr(idx), _ptr(perm), _ptr(seg), _ptr(ws),
                                            _stream(h)))
    else:
        check(lib().sar_router_fwd(_ptr(h), int(h.dtype == torch.float32), *[_ptr(t) for t in p], B, T, d, h1, h2, C,
                                   _ptr(logits), _ptr(probs), _ptr(idx), _ptr(perm), _ptr(seg), _ptr(ws), _stream(h)))
    LAUNCHES["k2"] += 2
    return RouterOut(logits, probs, idx, perm, seg)


def qv_lora_bwd(dy: torch.Tensor, x: torch.Tensor, u: torch.Tensor, Wt: torch.Tensor, At_stack: torch.Tensor,
                Bt_stack: torch.Tensor, utt_adapter: torch.Tensor, dA: torch.Tensor, dB: torch.Tensor, scale: float,
                need_dx: bool = True) -> Optional[torch.Tensor]:
    """K3: dx = dy·W + (scale·dy·B_k)·A_k; dA += vᵀx; dB += dyᵀu (accumulating into fp32 dA [n,r,d_in], dB [n,d_out,r]).

    Wt [d_in,d_out] = Wᵀ, At_stack [n,d_in,64] = rank-padded lora_Aᵀ, Bt_stack [n,r,d_out] = lora_Bᵀ (all bf16).
    """
    _need_cuda(dy, x, u, Wt, At_stack, Bt_stack, utt_adapter, dA, dB)
    dy = _bf16c(dy, "dy"); x = _bf16c(x, "x"); u = _bf16c(u, "u")
    B, T, d_out = dy.shape
    d_in = x.shape[-1]
    n, r = Bt_stack.shape[0], Bt_stack.shape[1]
    if dA.dtype != torch.float32 or dB.dtype != torch.float32 or not dA.is_contiguous() or not dB.is_contiguous():
        raise TypeError("dA/dB must be contiguous fp32")
    if tuple(dA.shape) != (n, r, d_in) or tuple(dB.shape) != (n, d_out, r):
        raise ValueError("dA must be [n,r,d_in], dB [n,d_out,r]")
    dx = torch.empty(B, T, d_in, dtype=torch.bfloat16, device=dy.device) if need_dx else None
    nbytes = lib().sar_workspace_bytes(_lib.SAR_OP_QV_LORA_BWD, B * T, T, max(d_in, d_out), r, n)
    if nbytes < 0:
        check(int(nbytes))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dy.device)
    check(lib().sar_qv_lora_bwd(_ptr(dy), _ptr(x), _ptr(u), _ptr(Wt.contiguous()), _ptr(At_stack.contiguous()),
                                _ptr(Bt_stack.contiguous()), None, _ptr(utt_adapter.contiguous()), _ptr(dx), _ptr(dA),
                                _ptr(dB), B, T, d_in, d_out, r, n, float(scale), _ptr(ws), _stream(dy)))
    LAUNCHES["k3"] += 3
    return dx
