"""Checkpoint loading and weight packing for the CUDA path.

Replaces the ``torch.load`` + ``load_state_dict`` boiler-plate of the reference plugins
(rendering/brdf_measured_disk.py:43-51) and ``load_pytorch_model_to_tinycuda``
(learning_repo_cleanup/utils/utils.py:13-23): a flow net (bias-free MLP) is packed ONCE into a
device-resident blob (fp32 image + fp16 tensor-core shared-memory image, see csrc/common.cuh) and a
base net into 308 floats.  Packing is cached per ``nn.Module`` and invalidated when a parameter is
modified in place (``_version``) or replaced.
"""
from __future__ import annotations

import ctypes
import weakref
from dataclasses import dataclass
from typing import Dict, List, Sequence

import struct

import numpy as np
import torch

from . import _lib


@dataclass
class PackedFlow:
    blob: torch.Tensor       # uint8, device
    in_dim: int
    hidden: int
    n_hidden: int

    @property
    def domain(self) -> int:
        """Domain of the sampler nets; generic tcnn-style nets (any other in_dim) have none and only run through
        ``ops.mlp_forward``."""
        if self.in_dim == 25:
            return _lib.DISK
        if self.in_dim == 26:
            return _lib.SPHERICAL
        raise ValueError(f"flow net takes {self.in_dim} inputs: only the 25 (disk) or 26 (spherical) input sampler "
                         "nets have a domain")

    def to(self, device) -> "PackedFlow":
        return PackedFlow(self.blob.to(device), self.in_dim, self.hidden, self.n_hidden)

    def set_fixup(self, sample: float, pdf: float) -> "PackedFlow":
        """Write this material's OWN fix-up thresholds into the blob header (bytes 40..51: magic "FIXT", sample, pdf): a
        multi-material launch called with a negative threshold recomputes, per material, only the rows below THAT material's
        threshold (``plugins.MultiMaterialSampler`` does this when its samplers carry calibrated thresholds)."""
        word = np.frombuffer(struct.pack("<Iff", 0x54584946, float(sample), float(pdf)), dtype=np.uint8)
        self.blob[40:52] = torch.from_numpy(word.copy()).to(self.blob.device)
        return self

    def get_fixup(self):
        """(sample, pdf) thresholds stored in the blob header, or None."""
        magic, ts, tp = struct.unpack("<Iff", self.blob[40:52].cpu().numpy().tobytes())
        return (ts, tp) if magic == 0x54584946 else None


def _np32(t) -> np.ndarray:
    if isinstance(t, torch.Tensor):
        t = t.detach().to("cpu", torch.float32).numpy()
    return np.ascontiguousarray(t, dtype=np.float32)


def flow_layers_from_state_dict(sd: Dict[str, torch.Tensor]) -> List[np.ndarray]:
    """[linear1.weight, linear2.weight, ..., output.weight] in layer order."""
    keys = sorted((k for k in sd if k.startswith("linear") and k.endswith(".weight")),
                  key=lambda k: int(k[len("linear"):-len(".weight")]))
    return [_np32(sd[k]) for k in keys] + [_np32(sd["output.weight"])]


def pack_flow_layers(layers: Sequence, device="cuda") -> PackedFlow:
    """Pack [W1 [H,in], W2 [H,H], ..., Wout [2,H]] (row-major, as nn.Linear stores them)."""
    ws = [_np32(w) for w in layers]
    n = len(ws)
    H, in_dim = ws[0].shape
    nbytes = _lib.lib.bsdfdiff_packed_flow_bytes(int(in_dim), int(H), n - 1)
    if nbytes == 0:
        raise _lib.BsdfDiffError(f"unsupported flow-net shape: in={in_dim} hidden={H} n_hidden={n - 1}")
    out = np.zeros(nbytes, np.uint8)
    ptrs = (ctypes.c_void_p * n)(*[w.ctypes.data for w in ws])
    rows = (ctypes.c_int * n)(*[w.shape[0] for w in ws])
    cols = (ctypes.c_int * n)(*[w.shape[1] for w in ws])
    _lib.check(_lib.lib.bsdfdiff_pack_flow(ptrs, rows, cols, n, out.ctypes.data), "bsdfdiff_pack_flow")
    return PackedFlow(torch.from_numpy(out).to(device), int(in_dim), int(H), n - 1)


def pack_flow_tcnn(params, in_dim: int, out_dim: int, hidden: int, n_hidden: int, device="cuda") -> PackedFlow:
    """Pack from a tinycudann-layout flat parameter vector (what load_pytorch_model_to_tinycuda writes)."""
    p = _np32(params)
    nbytes = _lib.lib.bsdfdiff_packed_flow_bytes(in_dim, hidden, n_hidden)
    if nbytes == 0:
        raise _lib.BsdfDiffError(f"unsupported tcnn net shape: in={in_dim} hidden={hidden} n_hidden={n_hidden}")
    in_pad = in_dim + 16 - in_dim % 16
    need = hidden * in_pad + (n_hidden - 1) * hidden * hidden + (out_dim + 16 - out_dim % 16) * hidden
    if p.size < need:
        raise _lib.BsdfDiffError(f"tcnn params too short: {p.size} < {need}")
    out = np.zeros(nbytes, np.uint8)
    _lib.check(_lib.lib.bsdfdiff_pack_flow_tcnn(p.ctypes.data, in_dim, out_dim, hidden, n_hidden, out.ctypes.data),
               "bsdfdiff_pack_flow_tcnn")
    return PackedFlow(torch.from_numpy(out).to(device), in_dim, hidden, n_hidden)


def pack_base_arrays(w1, b1, wo, bo, device="cuda") -> torch.Tensor:
    flat = np.concatenate([_np32(w1).ravel(), _np32(b1).ravel(), _np32(wo).ravel(), _np32(bo).ravel()])
    if flat.size != _lib.BASE_FLOATS:
        raise _lib.BsdfDiffError(f"base net must be 14->16->4 with biases (308 floats), got {flat.size}")
    return torch.from_numpy(flat).to(device)


def pack_base_state_dict(sd, device="cuda") -> torch.Tensor:
    return pack_base_arrays(sd["linear1.weight"], sd["linear1.bias"], sd["output.weight"], sd["output.bias"], device)


def load_checkpoint(path: str) -> Dict[str, torch.Tensor]:
    """The reference's checkpoints were pickled with CUDA storages; always map to CPU first."""
    return torch.load(path, map_location="cpu")


# ---- per-module cache -----------------------------------------------------------------------
_cache: "weakref.WeakKeyDictionary[torch.nn.Module, tuple]" = weakref.WeakKeyDictionary()


def invalidate(module: torch.nn.Module) -> None:
    """Drop the cached packing of ``module``.  Needed only after writing parameters through ``.data`` (which does not
    bump the version counter the cache is keyed on); in-place ops, ``load_state_dict`` and optimiser steps do."""
    _cache.pop(module, None)
    if hasattr(module, "_packed"):
        module._packed = None
        module._packed_sig = None


def _signature(module: torch.nn.Module, device) -> tuple:
    return (str(device),) + tuple((id(p), p._version, p.data_ptr()) for p in module.parameters())


def packed_flow_of(module: torch.nn.Module, device) -> PackedFlow:
    sig = _signature(module, device)
    hit = _cache.get(module)
    if hit is not None and hit[0] == sig:
        return hit[1]
    packed = pack_flow_layers(flow_layers_from_state_dict(module.state_dict()), device)
    _cache[module] = (sig, packed)
    return packed


def packed_base_of(module: torch.nn.Module, device) -> torch.Tensor:
    sig = _signature(module, device)
    hit = _cache.get(module)
    if hit is not None and hit[0] == sig:
        return hit[1]
    packed = pack_base_state_dict(module.state_dict(), device)
    _cache[module] = (sig, packed)
    return packed
