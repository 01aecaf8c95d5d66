"""The C-ABI shared library loads on a GPU-less box and exports every symbol include/sar.h declares; compute entry
points fail loudly (error code + message) without a device — there is no CPU fallback to fall into."""
import ctypes
import re
from pathlib import Path

import torch

from speech_adapter_routing_b200 import _lib

HEADER = Path(__file__).resolve().parents[1] / "include" / "sar.h"


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(sar_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_documented_entry_points():
    fns = declared_functions()
    for name in ["sar_version", "sar_last_error", "sar_device_ok", "sar_workspace_bytes", "sar_qv_lora_fwd",
                 "sar_qv_lora_fwd_rows", "sar_router_fwd", "sar_qv_lora_bwd", "sar_qv_lora_fwd_pair",
                 "sar_attn_proj_fwd", "sar_attn_proj_fwd_rows", "sar_linear_fwd", "sar_dense_fwd", "sar_layernorm_fwd",
                 "sar_attn_fwd", "sar_decode_self_attn", "sar_decode_cross_attn"]:
        assert name in fns


def test_ctypes_bindings_have_the_arity_the_header_declares():
    """Every prototype in include/sar.h has as many parameters as the ctypes argtypes bound in _lib.py (a drifted
    binding would otherwise corrupt the call silently)."""
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    protos = dict(re.findall(r"\b(sar_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text))
    assert set(protos) == set(_lib._SIGNATURES)
    for name, params in protos.items():
        params = params.strip()
        n = 0 if params in ("", "void") else len(params.split(","))
        assert n == len(_lib._SIGNATURES[name][1]), f"{name}: header has {n} parameters, ctypes binds {len(_lib._SIGNATURES[name][1])}"


def test_library_exports_every_declared_symbol(libsar):
    for name in declared_functions():
        assert hasattr(libsar, name), f"libsar.so does not export {name}"
    assert set(declared_functions()) == set(_lib.EXPORTED_SYMBOLS)


def test_version_and_error_string(libsar):
    assert libsar.sar_version() == 2
    assert isinstance(libsar.sar_last_error(), bytes)


def test_signatures_use_plain_c_types_only():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    assert "torch" not in text and "at::" not in text and "std::" not in text
    assert 'extern "C"' in HEADER.read_text()


def test_workspace_queries_need_no_device(libsar):
    assert libsar.sar_workspace_bytes(_lib.SAR_OP_QV_LORA_FWD, 96000, 1500, 768, 16, 4) == 0
    n = libsar.sar_workspace_bytes(_lib.SAR_OP_ROUTER_FWD, 64, 1500, 768, 0, 4)
    assert n >= 64 * 24 * 768 * 4
    assert libsar.sar_workspace_bytes(_lib.SAR_OP_QV_LORA_BWD, 16 * 1500, 1500, 768, 16, 1) > 0
    assert libsar.sar_workspace_bytes(99, 1, 1, 1, 1, 1) == _lib.SAR_EINVAL


def test_compute_calls_fail_loudly_without_a_gpu(libsar):
    if torch.cuda.is_available():
        return  # on a GPU box the parity tests cover the calls
    assert libsar.sar_device_ok() == _lib.SAR_ECUDA
    rc = libsar.sar_qv_lora_fwd(None, None, None, None, None, None, None, None, 1, 1, 64, 64, 16, 0, 1.0, 0, None)
    assert rc == _lib.SAR_ECUDA
    assert b"no CPU fallback" in libsar.sar_last_error()
    rc = libsar.sar_router_fwd(None, 0, *([None] * 12), 1, 1, 8, 1, 1, 1, *([None] * 5), None, None)
    assert rc == _lib.SAR_ECUDA
    assert libsar.sar_linear_fwd(None, 0, None, None, None, None, 1, 1, 64, 128, 0, 0, None) == _lib.SAR_ECUDA
    assert libsar.sar_dense_fwd(None, 0, 0, None, None, None, 0, 0, 0, None, 0, 0, 1, 1, 64, 128, 0, 0, None) == _lib.SAR_ECUDA
    assert libsar.sar_layernorm_fwd(None, None, None, None, 1, 64, 1e-5, None) == _lib.SAR_ECUDA
    assert libsar.sar_attn_proj_fwd(None, 0, None, None, None, None, None, None, None, None, 1, 1, 0, 1, 1, 64, 128, 16,
                                    0, 1.0, 0, None, None) == _lib.SAR_ECUDA


def test_python_ops_refuse_cpu_tensors():
    from speech_adapter_routing_b200 import ops
    import pytest

    x = torch.zeros(1, 4, 64, dtype=torch.bfloat16)
    W = torch.zeros(64, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.qv_lora_fwd(x, W, None, None, None, None, 1.0)
