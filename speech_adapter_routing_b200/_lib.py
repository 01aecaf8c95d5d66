"""ctypes binding of libsar.so (C ABI declared in include/sar.h).

The library is built in-tree (``speech_adapter_routing_b200/csrc/libsar.so``) by ``__graft_entry__.build()`` /
``make -C speech_adapter_routing_b200/csrc``.  There is deliberately no fallback: if the shared object is
missing, or the device is not an sm_100 GPU, every op raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_char_p, c_float, c_int, c_int64, c_uint32, c_void_p
from pathlib import Path

CSRC_DIR = Path(__file__).resolve().parent / "csrc"
LIB_PATH = CSRC_DIR / "libsar.so"

SAR_OK, SAR_EINVAL, SAR_EARCH, SAR_ECUDA, SAR_EWORKSPACE = 0, -1, -2, -3, -4
SAR_FLAG_SAVE_U = 1
SAR_FLAG_U_READY, SAR_FLAG_U_ONLY = 4, 8
SAR_OP_QV_LORA_FWD, SAR_OP_ROUTER_FWD, SAR_OP_QV_LORA_BWD, SAR_OP_QV_LORA_FWD_ROWS, SAR_OP_ATTN_PROJ_FWD = 0, 1, 2, 3, 4
SAR_RPAD = 64
SAR_ACT_NONE, SAR_ACT_GELU, SAR_ACT_GELU_BWD = 0, 1, 2
SAR_DTYPE_F32, SAR_DTYPE_BF16 = 0, 1

# name -> (restype, argtypes); mirrors include/sar.h one to one
_SIGNATURES = {
    "sar_version": (c_int, []),
    "sar_last_error": (c_char_p, []),
    "sar_device_ok": (c_int, []),
    "sar_workspace_bytes": (c_int64, [c_int, c_int64, c_int64, c_int64, c_int64, c_int64]),
    "sar_qv_lora_fwd": (c_int, [c_void_p] * 5 + [c_void_p, c_void_p, c_void_p] + [c_int] * 6 + [c_float, c_uint32, c_void_p]),
    "sar_attn_proj_fwd": (c_int, [c_void_p, c_int] + [c_void_p] * 8 + [c_int] * 9 + [c_float, c_uint32, c_void_p, c_void_p]),
    "sar_attn_proj_fwd_mix": (c_int, [c_void_p, c_int] + [c_void_p] * 8 + [c_int] * 9 + [c_float, c_uint32, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "sar_attn_proj_fwd_rows": (c_int, [c_void_p] * 9 + [c_int] * 7 + [c_float, c_uint32, c_void_p]),
    "sar_qv_lora_fwd_pair": (c_int, [c_void_p] * 8 + [c_int] * 6 + [c_float, c_uint32, c_void_p]),
    "sar_linear_fwd": (c_int, [c_void_p, c_int] + [c_void_p] * 4 + [c_int] * 5 + [c_uint32, c_void_p]),
    "sar_dense_fwd": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p,
                              c_int64, c_int64] + [c_int] * 5 + [c_uint32, c_void_p]),
    "sar_attn_fwd": (c_int, [c_void_p] * 4 + [c_int] * 6 + [c_void_p]),
    "sar_decode_self_attn": (c_int, [c_void_p] * 7 + [c_int] * 4 + [c_void_p]),
    "sar_decode_cross_attn": (c_int, [c_void_p] * 4 + [c_int] * 4 + [c_void_p]),
    "sar_logmel_fwd": (c_int, [c_void_p] * 8 + [c_int] * 4 + [c_void_p]),
    "sar_layernorm_fwd": (c_int, [c_void_p] * 4 + [c_int64, c_int, c_float, c_void_p]),
    "sar_layernorm_fwd_stats": (c_int, [c_void_p] * 6 + [c_int64, c_int, c_float, c_void_p]),
    "sar_layernorm_lora_u_fwd": (c_int, [c_void_p] * 7 + [c_int] * 6 + [c_float, c_float, c_void_p]),
    "sar_layernorm_lora_u_supported": (c_int, [c_int] * 3),
    "sar_operand_refresh": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "sar_qv_lora_fwd_rows": (c_int, [c_void_p] * 5 + [c_void_p, c_void_p] + [c_int] * 5 + [c_float, c_void_p, c_void_p]),
    "sar_router_fwd": (c_int, [c_void_p, c_int] + [c_void_p] * 12 + [c_int] * 6 + [c_void_p] * 5 + [c_void_p, c_void_p]),
    "sar_router_fwd_fused_ln": (c_int, [c_void_p, c_void_p, c_void_p, c_float] + [c_void_p] * 12 + [c_int] * 6 + [c_void_p] * 5 + [c_void_p, c_void_p]),
    "sar_qv_lora_bwd": (c_int, [c_void_p] * 7 + [c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 6 + [c_float, c_void_p, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class SarError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsar error {code}: {msg}")
        self.code = code


def build(verbose: bool = False) -> Path:
    """Compile libsar.so for sm_100a with nvcc (cross-compiles on a GPU-less box)."""
    out = subprocess.run(["make", "-C", str(CSRC_DIR), "-j8"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:])
        print(out.stderr[-4000:])
    if out.returncode != 0:
        raise RuntimeError("building libsar.so failed")
    return LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    """Load libsar.so (once).  Raises if it has not been built — there is no CPU/eager fallback."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            if os.environ.get("SAR_AUTOBUILD", "0") == "1":
                build()
            else:
                raise FileNotFoundError(
                    f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "or `make -C speech_adapter_routing_b200/csrc` (libsar has no fallback path)")
        handle = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != SAR_OK:
        raise SarError(rc, lib().sar_last_error().decode())
