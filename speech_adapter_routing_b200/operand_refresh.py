"""In-graph refresh of the cached LoRA operands (sar_operand_refresh, include/sar.h).

PEFT's LoRA layers read lora_A / lora_B from the parameters on every forward (src/models/whisper_lora.py:88-98 injects
them), so the reference's ``optimizer.step()`` (src/training/trainer.py:262-268) is visible to the next forward for free.
Here the kernels read derived bf16 operand stacks that the host rebuilds when a parameter's version changes
(RoutedLoRALinear._stacks, whisper_blocks._ProjPack.get).  A CUDA-graph replay runs no host code, so ``OperandRefresh``
lists every (parameter block -> operand block) copy once in a device-resident descriptor table and re-derives all of
them with ONE kernel launch that is captured as the first node of the step.
"""
from __future__ import annotations

import ctypes
import math
from typing import List

import torch

from . import _lib, ops
from ._lib import check, lib
from .lora_linear import RoutedLoRALinear
from .peft_compat import lora_modules


class _Desc(ctypes.Structure):      # mirrors sar_refresh_desc (include/sar.h)
    _fields_ = [("src", ctypes.c_void_p), ("dst", ctypes.c_void_p), ("rows", ctypes.c_int32), ("cols", ctypes.c_int32),
                ("src_rs", ctypes.c_int64), ("src_cs", ctypes.c_int64), ("dst_rs", ctypes.c_int64),
                ("dst_cs", ctypes.c_int64), ("scale", ctypes.c_float), ("src_dtype", ctypes.c_int32)]


def _proj_packs(model: torch.nn.Module):
    from .whisper_blocks import _ProjPack

    for mod in model.modules():
        pk = getattr(mod, "_sar_pack", None)
        if isinstance(pk, dict):
            for v in pk.values():
                for p in vars(v).values() if hasattr(v, "__dict__") else ():
                    if isinstance(p, _ProjPack):
                        yield p


def cached_tensors(model: torch.nn.Module) -> List[torch.Tensor]:
    """Every tensor held by the operand caches under ``model`` (RoutedLoRALinear._cache, the fused blocks' packs).  A
    captured graph reads these by address; whoever replays it has to keep them alive across host-side cache rebuilds."""
    out: List[torch.Tensor] = []
    seen = set()

    def walk(o, depth=0):
        if id(o) in seen or depth > 6:
            return
        seen.add(id(o))
        if isinstance(o, torch.Tensor):
            out.append(o)
        elif isinstance(o, dict):
            for v in o.values():
                walk(v, depth + 1)
        elif isinstance(o, (list, tuple)):
            for v in o:
                walk(v, depth + 1)
        elif type(o).__module__.startswith(__package__) and hasattr(o, "__dict__") and not isinstance(o, torch.nn.Module):
            walk(vars(o), depth + 1)

    for mod in model.modules():
        for attr in ("_cache", "_sar_pack"):
            c = mod.__dict__.get(attr)
            if c is not None:
                walk(c)
    return out


class OperandRefresh:
    """Descriptor table over every cached LoRA operand reachable from ``model``.  ``run()`` launches the refresh on the
    current stream.  The operand tensors are kept alive here: a later host-side rebuild (eager use after parameter
    updates) rebinds the caches to new tensors, the captured graph keeps reading — and refreshing — these."""

    def __init__(self, model: torch.nn.Module):
        self.keep: List[torch.Tensor] = []
        descs: List[_Desc] = []
        self.max_elems = 1
        device = None

        def add(src: torch.Tensor, dst: torch.Tensor, scale: float) -> None:
            if tuple(src.shape) != tuple(dst.shape) or dst.dtype != torch.bfloat16:
                raise RuntimeError("operand refresh: shape / dtype mismatch between a parameter and its operand block")
            if src.dtype == torch.float32:
                dt = _lib.SAR_DTYPE_F32
            elif src.dtype == torch.bfloat16:
                dt = _lib.SAR_DTYPE_BF16
            else:
                raise TypeError(f"operand refresh: LoRA parameters must be fp32 or bf16, got {src.dtype}")
            rows, cols = dst.shape
            descs.append(_Desc(src.data_ptr(), dst.data_ptr(), rows, cols, src.stride(0), src.stride(1), dst.stride(0),
                               dst.stride(1), float(scale), dt))
            self.max_elems = max(self.max_elems, rows * cols)
            self.keep.append(dst)

        for m in lora_modules(model).values():
            names = m.adapter_order
            if not names or not m.base_layer.weight.is_cuda:
                continue
            device = m.base_layer.weight.device
            st = m._stacks(backward=True)
            uniform = all(abs(m.scaling[n] - m.scaling[names[0]]) < 1e-12 for n in names)
            for k, n in enumerate(names):
                wA, wB, r = m.lora_A[n].weight.detach(), m.lora_B[n].weight.detach(), m.r[n]
                sB = 1.0 if uniform else float(m.scaling[n])
                add(wA, st["A"][k, :r, :], 1.0)                 # [n, rp, d_in]
                add(wB, st["Bp"][k, :, :r], sB)                 # [n, d_out, 64]
                add(wA.t(), st["At"][k, :, :r], 1.0)            # [n, d_in, 64]
                add(wB.t(), st["Bt"][k, :r, :], sB)             # [n, rp, d_out]
            self.keep += [st["A"], st["Bp"], st["At"], st["Bt"]]
        for p in _proj_packs(model):
            p.get()
            if getattr(p, "A", None) is None or not p.ok:
                continue
            si = 0
            for m, s in zip(p.mods, p.seg_scale):
                if not (isinstance(m, RoutedLoRALinear) and m.adapter_order):
                    continue
                names = m.adapter_order
                st = m._stacks()
                n, rp = st["A"].shape[0], st["A"].shape[1]
                fold = s != 1.0 and math.frexp(s)[0] == 0.5     # power-of-two output scale folded into the operands
                uniform = all(abs(m.scaling[x] - m.scaling[names[0]]) < 1e-12 for x in names)
                for k, nm in enumerate(names):
                    wA, wB, r = m.lora_A[nm].weight.detach(), m.lora_B[nm].weight.detach(), m.r[nm]
                    sB = (1.0 if uniform else float(m.scaling[nm])) * (s if fold else 1.0)
                    add(wA, p.A[si * n + k, :r, :], 1.0)
                    add(wB, p.Bp[si * n + k, :, :r], sB)
                si += 1
            self.keep += [p.A, p.Bp]
        # ---- eager use (maybe_refresh): what a refresh covers (LoRA parameter versions) and what it does not
        self._mods = [m for m in lora_modules(model).values() if m.adapter_order and m.base_layer.weight.is_cuda]
        self._packs = [p for p in _proj_packs(model) if getattr(p, "A", None) is not None and p.ok]
        self._mod_tensors = [(m, m._cache.get("A"), m._cache.get("At")) for m in self._mods]
        self._pack_tensors = [(p, p.A) for p in self._packs]
        self._lora_ws = [w for m in self._mods for n in m.adapter_order for w in (m.lora_A[n].weight, m.lora_B[n].weight)]
        self._frozen0 = self._frozen_sig()
        self._versions = self._lora_versions()
        self.n = len(descs)
        self.table = None
        if self.n:
            raw = bytes((_Desc * self.n)(*descs))
            self.table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)

    # ---- eager training loop: one launch after optimizer.step() instead of ~25 small kernels per module and pack
    def _frozen_sig(self) -> tuple:
        from .routing import operand_epoch
        from .whisper_blocks import _lin_params

        sig = [operand_epoch()]
        for m in self._mods:
            b = m.base_layer
            sig.append((b.weight.data_ptr(), b.weight._version, None if b.bias is None else (b.bias.data_ptr(), b.bias._version),
                        tuple(m.adapter_order)))
        for p in self._packs:
            for mod in p.mods:
                W, b = _lin_params(mod)
                sig.append((W.data_ptr(), W._version, None if b is None else (b.data_ptr(), b._version)))
        sig.append(tuple(w.data_ptr() for w in self._lora_ws))
        return tuple(sig)

    def _lora_versions(self) -> tuple:
        return tuple(w._version for w in self._lora_ws)

    def maybe_refresh(self) -> bool:
        """Eager counterpart of the in-graph refresh: if only LoRA parameter values changed since the operands were
        derived (an optimizer step), re-derive them with one launch and mark every cache current.  False = something
        else changed (frozen weights, adapter set, epoch, or a cache was rebuilt elsewhere): the caller drops this object
        and the caches rebuild the ordinary way."""
        if self._frozen_sig() != self._frozen0:
            return False
        for m, A, At in self._mod_tensors:
            c = m._cache
            if c.get("A") is not A or c.get("At") is not At:
                return False
        for p, A in self._pack_tensors:
            if p.A is not A:
                return False
        v = self._lora_versions()
        if v != self._versions:
            self.run()
            self._versions = v
            for m in self._mods:
                m._cache["key"] = m._key()
                m._cache["_Bm"] = None          # fp32 source stack: rebuilt from the parameters if anything asks for it
            for p in self._packs:
                p.key = p._current_key()
        return True

    def run(self) -> None:
        if self.n:
            stream = torch.cuda.current_stream(self.table.device).cuda_stream
            check(lib().sar_operand_refresh(self.table.data_ptr(), self.n, self.max_elems, stream))
            ops.LAUNCHES["refresh"] = ops.LAUNCHES.get("refresh", 0) + 1
