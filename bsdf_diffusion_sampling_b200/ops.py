"""Tensor-level operators over the C-ABI (``torch.ops.bsdfdiff.*``).

Each op is a thin ``torch.library`` custom op whose CUDA implementation hands raw device pointers and
the current CUDA stream to libbsdfdiff.so -- no host synchronisation, no allocation outside torch's
caching allocator, CUDA-graph capturable (same contract as tiny-cuda-nn's torch binding,
tiny-cuda-nn/bindings/torch/tinycudann/bindings.cpp:95-96).  There is no CPU implementation: CPU
tensors raise.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import (DISK, SPHERICAL, EPI_RAW, EPI_DISK, EPI_SPHERICAL, EPI_BSDF,  # noqa: F401
                   PREC_FP32, PREC_TC16, PREC_TC16_EXP)

_PREC_NAMES = {"fp32": PREC_FP32, "tc16": PREC_TC16, "tc16_exp": PREC_TC16_EXP}
_default_precision = _PREC_NAMES[os.environ.get("BSDFDIFF_PRECISION", "tc16")]

# Philox counter words consumed per query by one sample call (1 Box-Muller draw + <= 64 von Mises rounds),
# rounded to torch's granularity of 4.
PHILOX_OFFSET_PER_CALL = 68


def set_default_precision(name: str) -> None:
    """"tc16" (tcgen05, fp16 operands / fp32 accumulate) or "fp32" (CUDA cores, parity path)."""
    global _default_precision
    _default_precision = _PREC_NAMES[name]


def get_default_precision() -> int:
    return _default_precision


def _resolve_precision(precision) -> int:
    if precision is None:
        return _default_precision
    if isinstance(precision, str):
        return _PREC_NAMES[precision]
    return int(precision)


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _f32c(t: torch.Tensor, device=None) -> torch.Tensor:
    if device is not None and t.device != device:
        t = t.to(device)
    return t.detach().to(torch.float32).contiguous()


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"bsdfdiff.{what}: expected a CUDA tensor; this package has no CPU path")


def next_philox(device) -> Tuple[int, int]:
    """(seed, offset) from torch's CUDA generator, advancing it -- ``torch.manual_seed`` keeps
    controlling the sampler, like the reference's ``torch.randn_like`` (rendering/utils/model.py:390).

    The pair is read on the host and handed to the kernel BY VALUE, so it must not be baked into a CUDA graph
    (every replay would return the same base samples, where torch's own ``randn`` advances a device-side
    offset): under stream capture this raises; capture with an explicit ``seed=``/``offset=`` or a replayed
    ``x0=`` instead."""
    if torch.cuda.is_current_stream_capturing():
        raise RuntimeError("bsdfdiff: drawing (seed, offset) from torch's generator is not CUDA-graph capturable; "
                           "pass seed=/offset= (or x0=) explicitly when capturing")
    gen = torch.cuda.default_generators[torch.device(device).index or 0]
    seed, offset = gen.initial_seed(), gen.get_offset()
    gen.set_offset(offset + PHILOX_OFFSET_PER_CALL)
    return seed & (2 ** 63 - 1), offset


# ---- conditioning-triggered fp32 fix-up of the tc16 path (include/bsdfdiff.h) ---------------------------------
# Default conditioning thresholds (a query whose weight w falls below is recomputed in fp32), per flow family and call:
# the largest-waste-free values for which EVERY material the reference ships (27 measured-disk, 25 measured-spherical and
# 25 bsdf checkpoints) meets the raw BASELINE.md section-5 bars on the shipped path -- calibrated on the measured errors of
# the tensor-core launch (profiles/flag_dump.py + flag_study.py -> profiles/r2s_flag_study.txt, checked by
# profiles/material_sweep.py and tests/test_all_materials.py).  One material dictates each value (disk sample:
# ilm_solo_m_68's MEDIAN error; spherical pdf(): chm_orange's reverse, expanding flow); a material calibrated on its own
# (plugins.NeuralBSDFSampler.calibrate_fixup, stored by materials.MaterialPack) usually needs no fix-up at all.
_FIX_DEFAULT = {
    ("disk", "sample"): 0.15, ("disk", "pdf"): 1.0 / 60.0,
    ("spherical", "sample"): 0.25, ("spherical", "pdf"): 0.5,
    ("bsdf", "sample"): 0.125, ("bsdf", "pdf"): 0.25,
}
_fixup_threshold: Optional[float] = (float(os.environ["BSDFDIFF_FIXUP"]) if "BSDFDIFF_FIXUP" in os.environ else None)
_last_fix_scratch = {}


def set_fixup_threshold(thr: Optional[float]) -> None:
    """Override the conditioning weight below which a tc16 query is recomputed in fp32 for every call that does not pass
    ``fixup=`` (0 disables the second launch; None restores the per-family defaults)."""
    global _fixup_threshold
    _fixup_threshold = None if thr is None else float(thr)


def default_fixup_threshold(domain: int, epilogue: int = EPI_RAW, mode: str = "sample") -> float:
    """The threshold a call without ``fixup=`` uses: the process-wide override (``set_fixup_threshold`` / BSDFDIFF_FIXUP)
    if there is one, else the calibrated default of the flow family (raw-epilogue spherical calls cannot tell the
    measured-spherical from the bsdf plugin and take the stricter measured-spherical value)."""
    if _fixup_threshold is not None:
        return _fixup_threshold
    family = "disk" if domain == DISK else ("bsdf" if epilogue == EPI_BSDF else "spherical")
    return _FIX_DEFAULT[(family, mode)]


def get_fixup_threshold(domain: int = DISK, epilogue: int = EPI_RAW, mode: str = "sample") -> float:
    return default_fixup_threshold(domain, epilogue, mode)


def last_fixup_count(device=None) -> int:
    """Rows the most recent tc16 call on ``device`` recomputed in fp32 (synchronises; diagnostics only)."""
    dev = torch.device(device if device is not None else "cuda")
    t = _last_fix_scratch.get(dev.index or 0)
    return int(t[0].item()) if t is not None else 0


def _fix_scratch(n: int, device: torch.device, precision: int, fix_thr: float, T: int) -> Optional[torch.Tensor]:
    if not (fix_thr > 0.0) or precision == PREC_FP32 or T == 0:
        return None
    t = torch.empty((_lib.lib.bsdfdiff_fixup_scratch_bytes(n) + 3) // 4, dtype=torch.int32, device=device)
    if not torch.cuda.is_current_stream_capturing():
        _last_fix_scratch[device.index or 0] = t
    return t


# ------------------------------------------------------------------------------------------------
# custom ops.  int arguments: precision, domain, epilogue, T, hidden, n_hidden, seed, offset, first_index
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("bsdfdiff::sample", mutates_args=(), device_types="cuda")
def _sample_op(wi: torch.Tensor, flow_blob: Optional[torch.Tensor], base: torch.Tensor, x0: Optional[torch.Tensor],
               u: Optional[torch.Tensor], precision: int, domain: int, epilogue: int, T: int, hidden: int, n_hidden: int,
               seed: int, offset: int, first_index: int,
               want_x0: bool, fix_thr: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    n = wi.shape[0]
    out_dir = torch.empty((n, 2 if epilogue == EPI_RAW else 3), dtype=torch.float32, device=wi.device)
    out_pdf = torch.empty((n,), dtype=torch.float32, device=wi.device)
    scratch = _fix_scratch(n, wi.device, precision, fix_thr, T)
    need_x0 = want_x0
    out_x0 = torch.empty((n if need_x0 else 0, 2), dtype=torch.float32, device=wi.device)
    with torch.cuda.device(wi.device):
        rc = _lib.lib.bsdfdiff_sample(precision, domain, epilogue, T, n, wi.data_ptr(),
                                      flow_blob.data_ptr() if flow_blob is not None else None,
                                      hidden, n_hidden, base.data_ptr(), x0.data_ptr() if x0 is not None else None,
                                      u.data_ptr() if u is not None else None,
                                      seed, offset, first_index, out_dir.data_ptr(), out_pdf.data_ptr(),
                                      out_x0.data_ptr() if need_x0 else None,
                                      fix_thr if scratch is not None else 0.0,
                                      scratch.data_ptr() if scratch is not None else None, _stream(wi))
    _lib.check(rc, "bsdfdiff_sample")
    return out_dir, out_pdf, (out_x0 if want_x0 else out_x0[:0])


@_sample_op.register_fake
def _(wi, flow_blob, base, x0, u, precision, domain, epilogue, T, hidden, n_hidden, seed, offset, first_index, want_x0,
      fix_thr):
    n = wi.shape[0]
    return (wi.new_empty((n, 2 if epilogue == EPI_RAW else 3)), wi.new_empty((n,)),
            wi.new_empty((n if want_x0 else 0, 2)))


@torch.library.custom_op("bsdfdiff::sample_out", mutates_args=("out_dir", "out_pdf", "scratch"), device_types="cuda")
def _sample_out_op(wi: torch.Tensor, flow_blob: Optional[torch.Tensor], base: torch.Tensor, x0: Optional[torch.Tensor],
                   u: Optional[torch.Tensor], out_dir: torch.Tensor, out_pdf: torch.Tensor, scratch: Optional[torch.Tensor],
                   precision: int, domain: int, epilogue: int, T: int, hidden: int, n_hidden: int,
                   seed: int, offset: int, first_index: int, fix_thr: float) -> None:
    """As ``bsdfdiff::sample`` but writes into caller-owned buffers (no allocation: streaming pipelines).
    ``scratch`` (int32, ``sample_scratch_elems(n)`` elements) enables the fix-up: [count, pad x3][row list n, even-padded][base samples of the listed rows 2 n]."""
    n = wi.shape[0]
    fix = scratch is not None and fix_thr > 0.0 and precision != PREC_FP32 and T > 0
    with torch.cuda.device(wi.device):
        rc = _lib.lib.bsdfdiff_sample(precision, domain, epilogue, T, n, wi.data_ptr(),
                                      flow_blob.data_ptr() if flow_blob is not None else None,
                                      hidden, n_hidden, base.data_ptr(), x0.data_ptr() if x0 is not None else None,
                                      u.data_ptr() if u is not None else None,
                                      seed, offset, first_index, out_dir.data_ptr(), out_pdf.data_ptr(), None,
                                      fix_thr if fix else 0.0, scratch.data_ptr() if fix else None, _stream(wi))
    _lib.check(rc, "bsdfdiff_sample")


@torch.library.custom_op("bsdfdiff::pdf", mutates_args=(), device_types="cuda")
def _pdf_op(wo: torch.Tensor, wi: torch.Tensor, flow_blob: Optional[torch.Tensor], base: torch.Tensor,
            precision: int, domain: int, epilogue: int, T: int, hidden: int, n_hidden: int,
            fix_thr: float) -> torch.Tensor:
    n = wi.shape[0]
    out = torch.empty((n,), dtype=torch.float32, device=wi.device)
    scratch = _fix_scratch(n, wi.device, precision, fix_thr, T)
    with torch.cuda.device(wi.device):
        rc = _lib.lib.bsdfdiff_pdf(precision, domain, epilogue, T, n, wo.data_ptr(), wi.data_ptr(),
                                   flow_blob.data_ptr() if flow_blob is not None else None, hidden, n_hidden,
                                   base.data_ptr(), out.data_ptr(), fix_thr if scratch is not None else 0.0,
                                   scratch.data_ptr() if scratch is not None else None, _stream(wi))
    _lib.check(rc, "bsdfdiff_pdf")
    return out


@_pdf_op.register_fake
def _(wo, wi, flow_blob, base, precision, domain, epilogue, T, hidden, n_hidden, fix_thr):
    return wi.new_empty((wi.shape[0],))


@torch.library.custom_op("bsdfdiff::base_log_prob", mutates_args=(), device_types="cuda")
def _base_log_prob_op(x: torch.Tensor, wi: torch.Tensor, base: torch.Tensor, domain: int) -> torch.Tensor:
    out = torch.empty((wi.shape[0],), dtype=torch.float32, device=wi.device)
    with torch.cuda.device(wi.device):
        rc = _lib.lib.bsdfdiff_base_log_prob(domain, wi.shape[0], x.data_ptr(), wi.data_ptr(), base.data_ptr(),
                                             out.data_ptr(), _stream(wi))
    _lib.check(rc, "bsdfdiff_base_log_prob")
    return out


@_base_log_prob_op.register_fake
def _(x, wi, base, domain):
    return wi.new_empty((wi.shape[0],))


@torch.library.custom_op("bsdfdiff::flow_forward", mutates_args=(), device_types="cuda")
def _flow_forward_op(wi: torch.Tensor, wi_repeat: int, n: int, flow_blob: torch.Tensor, base: Optional[torch.Tensor],
                     x0: Optional[torch.Tensor], precision: int, domain: int, T: int, hidden: int, n_hidden: int,
                     seed: int, offset: int, first_index: int) -> Tuple[torch.Tensor, torch.Tensor]:
    out_x = torch.empty((n, 2), dtype=torch.float32, device=wi.device)
    out_x0 = torch.empty((n, 2), dtype=torch.float32, device=wi.device)
    with torch.cuda.device(wi.device):
        rc = _lib.lib.bsdfdiff_flow_forward(precision, domain, T, n, wi.data_ptr(), wi_repeat, flow_blob.data_ptr(),
                                            hidden, n_hidden, base.data_ptr() if base is not None else None,
                                            x0.data_ptr() if x0 is not None else None, seed, offset, first_index,
                                            out_x.data_ptr(), out_x0.data_ptr(), _stream(wi))
    _lib.check(rc, "bsdfdiff_flow_forward")
    return out_x, out_x0


@_flow_forward_op.register_fake
def _(wi, wi_repeat, n, flow_blob, base, x0, precision, domain, T, hidden, n_hidden, seed, offset, first_index):
    return wi.new_empty((n, 2)), wi.new_empty((n, 2))


@torch.library.custom_op("bsdfdiff::mlp_forward", mutates_args=(), device_types="cuda")
def _mlp_forward_op(x: torch.Tensor, flow_blob: torch.Tensor, precision: int, hidden: int,
                    n_hidden: int) -> torch.Tensor:
    n, in_dim = x.shape
    out = torch.empty((n, 2), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib.bsdfdiff_mlp_forward(precision, n, x.data_ptr(), in_dim, flow_blob.data_ptr(), hidden, n_hidden,
                                           out.data_ptr(), _stream(x))
    _lib.check(rc, "bsdfdiff_mlp_forward")
    return out


@_mlp_forward_op.register_fake
def _(x, flow_blob, precision, hidden, n_hidden):
    return x.new_empty((x.shape[0], 2))


# ---- planar directions (three component arrays per tensor: Dr.Jit Vector3f through DLPack, zero copy) -----------------
def _ptr3(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor):
    import ctypes
    return (ctypes.c_void_p * 3)(a.data_ptr(), b.data_ptr(), c.data_ptr())


@torch.library.custom_op("bsdfdiff::sample_planar", mutates_args=(), device_types="cuda")
def _sample_planar_op(wx: torch.Tensor, wy: torch.Tensor, wz: torch.Tensor, flow_blob: torch.Tensor, base: torch.Tensor,
                      x0: Optional[torch.Tensor], u: Optional[torch.Tensor], precision: int, domain: int, epilogue: int,
                      T: int, hidden: int, n_hidden: int, seed: int, offset: int, first_index: int,
                      fix_thr: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    n = wx.shape[0]
    ox, oy, oz, pdf = (torch.empty((n,), dtype=torch.float32, device=wx.device) for _ in range(4))
    scratch = _fix_scratch(n, wx.device, precision, fix_thr, T)
    with torch.cuda.device(wx.device):
        rc = _lib.lib.bsdfdiff_sample_planar(precision, domain, epilogue, T, n, _ptr3(wx, wy, wz), flow_blob.data_ptr(),
                                             hidden, n_hidden, base.data_ptr(),
                                             x0.data_ptr() if x0 is not None else None,
                                             u.data_ptr() if u is not None else None, seed, offset, first_index,
                                             _ptr3(ox, oy, oz), pdf.data_ptr(), None,
                                             fix_thr if scratch is not None else 0.0,
                                             scratch.data_ptr() if scratch is not None else None, _stream(wx))
    _lib.check(rc, "bsdfdiff_sample_planar")
    return ox, oy, oz, pdf


@_sample_planar_op.register_fake
def _(wx, wy, wz, flow_blob, base, x0, u, precision, domain, epilogue, T, hidden, n_hidden, seed, offset, first_index, fix_thr):
    return tuple(wx.new_empty((wx.shape[0],)) for _ in range(4))


@torch.library.custom_op("bsdfdiff::pdf_planar", mutates_args=(), device_types="cuda")
def _pdf_planar_op(ox: torch.Tensor, oy: torch.Tensor, oz: torch.Tensor, wx: torch.Tensor, wy: torch.Tensor, wz: torch.Tensor,
                   flow_blob: torch.Tensor, base: torch.Tensor, precision: int, domain: int, epilogue: int, T: int,
                   hidden: int, n_hidden: int, fix_thr: float) -> torch.Tensor:
    n = wx.shape[0]
    out = torch.empty((n,), dtype=torch.float32, device=wx.device)
    scratch = _fix_scratch(n, wx.device, precision, fix_thr, T)
    with torch.cuda.device(wx.device):
        rc = _lib.lib.bsdfdiff_pdf_planar(precision, domain, epilogue, T, n, _ptr3(ox, oy, oz), _ptr3(wx, wy, wz),
                                          flow_blob.data_ptr(), hidden, n_hidden, base.data_ptr(), out.data_ptr(),
                                          fix_thr if scratch is not None else 0.0,
                                          scratch.data_ptr() if scratch is not None else None, _stream(wx))
    _lib.check(rc, "bsdfdiff_pdf_planar")
    return out


@_pdf_planar_op.register_fake
def _(ox, oy, oz, wx, wy, wz, flow_blob, base, precision, domain, epilogue, T, hidden, n_hidden, fix_thr):
    return wx.new_empty((wx.shape[0],))


def from_dlpack(obj) -> torch.Tensor:
    """A torch view of any DLPack exporter (Dr.Jit arrays, CuPy, JAX, ...) without a copy; torch tensors pass through."""
    return obj if isinstance(obj, torch.Tensor) else torch.from_dlpack(obj)


def _planar3(comps, what: str):
    cs = [from_dlpack(c) for c in comps]
    if len(cs) != 3:
        raise ValueError(f"bsdfdiff: {what} must be three component arrays")
    n = cs[0].shape[0]
    for c in cs:
        if (not c.is_cuda) or c.dtype != torch.float32 or c.dim() != 1 or c.shape[0] != n or not c.is_contiguous() \
                or c.device != cs[0].device:
            raise ValueError(f"bsdfdiff: {what} components must be contiguous 1-D fp32 CUDA arrays of one length on one device")
    return cs


def sample_planar(wi_xyz, flow, base: torch.Tensor, T: int, *, epilogue: int, x0=None, u=None, seed: Optional[int] = None,
                  offset: int = 0, first_index: int = 0, precision=None, fixup=None):
    """``sample`` on three component arrays (anything that speaks DLPack) -> (wo_x, wo_y, wo_z, pdf), each [n]."""
    wx, wy, wz = _planar3(wi_xyz, "wi")
    if epilogue == EPI_RAW:
        raise ValueError("planar directions exist for the plugin epilogues only")
    _check_flow(flow, T, "sample_planar")
    if T < 1:
        raise ValueError("T must be >= 1")
    x0, u, seed, offset = _noise(wx.unsqueeze(1), x0, u, seed, offset)
    return _sample_planar_op(wx, wy, wz, flow.blob, base, x0, u, _resolve_precision(precision), flow.domain, epilogue,
                             int(T), flow.hidden, flow.n_hidden, int(seed), int(offset), int(first_index), _fix_thr(fixup, flow.domain, epilogue, "sample"))


def pdf_planar(wo_xyz, wi_xyz, flow, base: torch.Tensor, T: int, *, epilogue: int, precision=None, fixup=None) -> torch.Tensor:
    wx, wy, wz = _planar3(wi_xyz, "wi")
    ox, oy, oz = _planar3(wo_xyz, "wo")
    if ox.shape[0] != wx.shape[0]:
        raise ValueError("bsdfdiff: wi and wo must have the same number of rows")
    if epilogue == EPI_RAW:
        raise ValueError("planar directions exist for the plugin epilogues only")
    _check_flow(flow, T, "pdf_planar")
    return _pdf_planar_op(ox, oy, oz, wx, wy, wz, flow.blob, base, _resolve_precision(precision), flow.domain, epilogue,
                          int(T), flow.hidden, flow.n_hidden, _fix_thr(fixup, flow.domain, epilogue, "pdf"))


# ---- one wavefront, several materials (include/bsdfdiff.h: bsdfdiff_multi_plan / _sample_multi / _pdf_multi) ----------
@torch.library.custom_op("bsdfdiff::multi_plan", mutates_args=(), device_types="cuda")
def _multi_plan_op(material_id: torch.Tensor, n_materials: int) -> torch.Tensor:
    n = material_id.shape[0]
    scratch = torch.empty((_lib.lib.bsdfdiff_multi_scratch_bytes(n, n_materials) + 3) // 4, dtype=torch.int32,
                          device=material_id.device)
    with torch.cuda.device(material_id.device):
        rc = _lib.lib.bsdfdiff_multi_plan(n, material_id.data_ptr(), n_materials, scratch.data_ptr(),
                                          _stream(material_id))
    _lib.check(rc, "bsdfdiff_multi_plan")
    return scratch


@_multi_plan_op.register_fake
def _(material_id, n_materials):
    return material_id.new_empty((1280 + 4 * (material_id.shape[0] // 128 + n_materials + 1) + 4 * material_id.shape[0],),
                                 dtype=torch.int32)


@torch.library.custom_op("bsdfdiff::sample_multi", mutates_args=("plan",), device_types="cuda")
def _sample_multi_op(wi: torch.Tensor, plan: torch.Tensor, flows: torch.Tensor, bases: torch.Tensor,
                     x0: Optional[torch.Tensor], u: Optional[torch.Tensor], precision: int, domain: int, epilogue: int, T: int, hidden: int,
                     n_hidden: int, seed: int, offset: int, first_index: int,
                     fix_thr: float) -> Tuple[torch.Tensor, torch.Tensor]:
    n = wi.shape[0]
    out_dir = torch.empty((n, 2 if epilogue == EPI_RAW else 3), dtype=torch.float32, device=wi.device)
    out_pdf = torch.empty((n,), dtype=torch.float32, device=wi.device)
    with torch.cuda.device(wi.device):
        rc = _lib.lib.bsdfdiff_sample_multi(precision, domain, epilogue, T, n, wi.data_ptr(), plan.data_ptr(),
                                            flows.shape[0], flows.data_ptr(), bases.data_ptr(), hidden, n_hidden,
                                            x0.data_ptr() if x0 is not None else None,
                                            u.data_ptr() if u is not None else None, seed, offset, first_index,
                                            out_dir.data_ptr(), out_pdf.data_ptr(), None, fix_thr, _stream(wi))
    _lib.check(rc, "bsdfdiff_sample_multi")
    return out_dir, out_pdf


@_sample_multi_op.register_fake
def _(wi, plan, flows, bases, x0, u, precision, domain, epilogue, T, hidden, n_hidden, seed, offset, first_index, fix_thr):
    n = wi.shape[0]
    return wi.new_empty((n, 2 if epilogue == EPI_RAW else 3)), wi.new_empty((n,))


@torch.library.custom_op("bsdfdiff::pdf_multi", mutates_args=("plan",), device_types="cuda")
def _pdf_multi_op(wo: torch.Tensor, wi: torch.Tensor, plan: torch.Tensor, flows: torch.Tensor, bases: torch.Tensor,
                  precision: int, domain: int, epilogue: int, T: int, hidden: int, n_hidden: int,
                  fix_thr: float) -> torch.Tensor:
    n = wi.shape[0]
    out = torch.empty((n,), dtype=torch.float32, device=wi.device)
    with torch.cuda.device(wi.device):
        rc = _lib.lib.bsdfdiff_pdf_multi(precision, domain, epilogue, T, n, wo.data_ptr(), wi.data_ptr(),
                                         plan.data_ptr(), flows.shape[0], flows.data_ptr(), bases.data_ptr(), hidden,
                                         n_hidden, out.data_ptr(), fix_thr, _stream(wi))
    _lib.check(rc, "bsdfdiff_pdf_multi")
    return out


@_pdf_multi_op.register_fake
def _(wo, wi, plan, flows, bases, precision, domain, epilogue, T, hidden, n_hidden, fix_thr):
    return wi.new_empty((wi.shape[0],))


class MultiPlan:
    """Device-side bucketing of one wavefront's rows by material id (no host read).  Reusable by any number of
    ``sample_multi`` / ``pdf_multi`` calls on the same ``material_id`` column (they must be stream-ordered: the plan
    buffer also holds the per-call fix-up lists)."""

    def __init__(self, material_id: torch.Tensor, n_materials: int):
        _require_cuda(material_id, "multi_plan")
        if material_id.dim() != 1 or material_id.dtype not in (torch.int32, torch.int64):
            raise TypeError("material_id must be a 1-D int32 / int64 tensor")
        if not 1 <= int(n_materials) <= 255:
            raise ValueError("1 <= n_materials <= 255")
        if material_id.shape[0] >= 2 ** 31:
            raise ValueError("a multi-material wavefront holds fewer than 2^31 rows")
        self.n, self.n_materials = material_id.shape[0], int(n_materials)
        self.scratch = _multi_plan_op(material_id.to(torch.int32).contiguous(), self.n_materials)

    def counts(self) -> torch.Tensor:
        """Rows per material [n_materials] plus the inactive rows (last entry); device tensor (diagnostics)."""
        return self.scratch[: self.n_materials + 1]

    def fixup_counts(self) -> torch.Tensor:
        """Rows the last tc16 call on this plan recomputed in fp32, per material (device tensor; diagnostics)."""
        return self.scratch[1024:1024 + self.n_materials]


class MaterialTable:
    """Device pointer tables of a scene's material set: packed flow blobs and base-net blobs of one shape."""

    def __init__(self, flows, bases):
        flows, bases = list(flows), list(bases)
        if not flows or len(flows) != len(bases):
            raise ValueError("need one base net per flow net")
        f0 = flows[0]
        for f in flows:
            if (f.domain, f.hidden, f.n_hidden, f.in_dim) != (f0.domain, f0.hidden, f0.n_hidden, f0.in_dim):
                raise ValueError("all materials of one table must share the flow-net shape")
        dev = f0.blob.device
        for t in [f.blob for f in flows] + bases:
            if t.device != dev:
                raise ValueError("all material blobs must live on one device")
        self.flows, self.bases = flows, bases         # keep the blobs alive
        self.domain, self.hidden, self.n_hidden, self.in_dim = f0.domain, f0.hidden, f0.n_hidden, f0.in_dim
        self.flow_ptrs = torch.tensor([f.blob.data_ptr() for f in flows], dtype=torch.int64, device=dev)
        self.base_ptrs = torch.tensor([b.data_ptr() for b in bases], dtype=torch.int64, device=dev)

    def __len__(self):
        return len(self.flows)


def _check_multi(table: MaterialTable, plan: MultiPlan, wi: torch.Tensor, T: int, what: str) -> None:
    if T < 1:
        raise ValueError("T must be >= 1")
    if table.in_dim not in (25, 26):
        raise ValueError(f"bsdfdiff.{what}: the sampler nets take 25 (disk) or 26 (spherical) inputs")
    if plan.n != wi.shape[0] or plan.n_materials != len(table):
        raise ValueError(f"bsdfdiff.{what}: the plan was built for {plan.n} rows / {plan.n_materials} materials, "
                         f"got {wi.shape[0]} rows / {len(table)} materials")
    if plan.scratch.device != wi.device or table.flow_ptrs.device != wi.device:
        raise ValueError(f"bsdfdiff.{what}: wavefront, plan and material table must live on one device")


def sample_multi(wi: torch.Tensor, plan: MultiPlan, table: MaterialTable, T: int, *, epilogue: int = EPI_RAW,
                 x0: Optional[torch.Tensor] = None, u: Optional[torch.Tensor] = None, seed: Optional[int] = None,
                 offset: int = 0, first_index: int = 0, precision=None, fixup=None):
    """One launch for a wavefront with a material id per row -> (dir [n,2|3], pdf [n]) in wavefront order."""
    _require_cuda(wi, "sample_multi")
    wi = _f32c(wi)
    _check_rows(wi, wi.shape[0], 2 if epilogue == EPI_RAW else 3, "wi")
    _check_multi(table, plan, wi, T, "sample_multi")
    x0, u, seed, offset = _noise(wi, x0, u, seed, offset)
    return _sample_multi_op(wi, plan.scratch, table.flow_ptrs, table.base_ptrs, x0, u, _resolve_precision(precision),
                            table.domain, epilogue, int(T), table.hidden, table.n_hidden, int(seed), int(offset),
                            int(first_index), _fix_thr(fixup, table.domain, epilogue, "sample"))


def pdf_multi(wo: torch.Tensor, wi: torch.Tensor, plan: MultiPlan, table: MaterialTable, T: int, *,
              epilogue: int = EPI_RAW, precision=None, fixup=None) -> torch.Tensor:
    _require_cuda(wi, "pdf_multi")
    wi = _f32c(wi)
    wo = _f32c(wo, wi.device)
    cols = 2 if epilogue == EPI_RAW else 3
    _check_rows(wi, wi.shape[0], cols, "wi")
    _check_rows(wo, wi.shape[0], cols, "wo")
    _check_multi(table, plan, wi, T, "pdf_multi")
    return _pdf_multi_op(wo, wi, plan.scratch, table.flow_ptrs, table.base_ptrs, _resolve_precision(precision),
                         table.domain, epilogue, int(T), table.hidden, table.n_hidden, _fix_thr(fixup, table.domain, epilogue, "pdf"))


# ------------------------------------------------------------------------------------------------
# friendly wrappers
# ------------------------------------------------------------------------------------------------
class NullFlow:
    """No flow net: with T = 0 the sampler returns the base sample / base density alone."""
    blob = None
    hidden, n_hidden = 32, 1

    def __init__(self, domain: int):
        self.domain = domain
        self.in_dim = 25 if domain == DISK else 26


def _check_flow(flow, T: int, what: str) -> None:
    if T < 0 or (T == 0) != (flow.blob is None):
        raise ValueError("T must be >= 1 (T == 0 only with NullFlow: base distribution alone)")
    if flow.in_dim not in (25, 26):
        raise ValueError(f"bsdfdiff.{what}: the sampler nets take 25 (disk) or 26 (spherical) inputs, got {flow.in_dim}")


def _check_rows(t: torch.Tensor, n: int, cols: int, what: str) -> None:
    if t.dim() != 2 or t.shape[0] != n or t.shape[1] != cols:
        raise ValueError(f"bsdfdiff: {what} must have shape ({n}, {cols}), got {tuple(t.shape)}")


def _fix_thr(fixup, domain: int = DISK, epilogue: int = EPI_RAW, mode: str = "sample") -> float:
    """``fixup``: None (process-wide override or the family default), a threshold, or {"sample": t, "pdf": t}."""
    if isinstance(fixup, dict):
        fixup = fixup.get(mode)
    return default_fixup_threshold(domain, epilogue, mode) if fixup is None else float(fixup)


def _noise(wi: torch.Tensor, x0, u, seed, offset):
    """-> (x0, u, seed, offset): exactly one noise source.  ``x0`` = replayed base samples [n,2]; ``u`` = renderer
    uniforms [n,3] in [0,1) (Mitsuba's sample2.x, sample2.y, sample1); else Philox (seed, offset), drawn from torch's
    CUDA generator when no seed is given."""
    n = wi.shape[0]
    if x0 is not None and u is not None:
        raise ValueError("bsdfdiff: pass either x0= (replayed base samples) or u= (renderer uniforms), not both")
    if x0 is not None:
        x0 = _f32c(x0, wi.device)
        _check_rows(x0, n, 2, "x0")
        return x0, None, 0, 0
    if u is not None:
        u = _f32c(u, wi.device)
        _check_rows(u, n, 3, "u")
        return None, u, 0, 0
    if seed is None:
        seed, offset = next_philox(wi.device)
    return None, None, int(seed), int(offset)


def sample(wi: torch.Tensor, flow, base: torch.Tensor, T: int, *, epilogue: int = EPI_RAW,
           x0: Optional[torch.Tensor] = None, u: Optional[torch.Tensor] = None, seed: Optional[int] = None, offset: int = 0,
           first_index: int = 0, precision=None, return_x0: bool = True, fixup=None):
    """-> (dir [n,2|3], pdf [n], x0 [n,2]).  ``flow`` is a ``weights.PackedFlow``.
    Noise: ``x0=`` replays base samples, ``u=`` [n,3] uses the renderer's uniforms (sample2.x, sample2.y, sample1), else
    Philox.  ``return_x0=False`` skips returning the base sample; x0 is then empty.  ``fixup`` = conditioning threshold
    of the fp32 fix-up pass of the tc16 path (None: the module default, 0: off)."""
    _require_cuda(wi, "sample")
    wi = _f32c(wi)
    _check_flow(flow, T, "sample")
    _check_rows(wi, wi.shape[0], 2 if epilogue == EPI_RAW else 3, "wi")
    x0, u, seed, offset = _noise(wi, x0, u, seed, offset)
    return _sample_op(wi, flow.blob, base, x0, u, _resolve_precision(precision), flow.domain, epilogue, int(T),
                      flow.hidden, flow.n_hidden, int(seed), int(offset), int(first_index), bool(return_x0),
                      _fix_thr(fixup, flow.domain, epilogue, "sample"))


def sample_scratch_elems(n: int) -> int:
    """int32 elements of the scratch buffer ``sample_into`` needs for ``n`` rows when the fix-up is on."""
    return 4 + ((int(n) + 1) & ~1) + 2 * int(n)


def sample_into(wi: torch.Tensor, flow, base: torch.Tensor, T: int, out_dir: torch.Tensor, out_pdf: torch.Tensor, *,
                epilogue: int = EPI_RAW, x0: Optional[torch.Tensor] = None, u: Optional[torch.Tensor] = None,
                seed: Optional[int] = None, offset: int = 0,
                first_index: int = 0, precision=None, scratch: Optional[torch.Tensor] = None, fixup=None) -> None:
    """``sample`` into caller-owned contiguous fp32 CUDA buffers ``out_dir`` [n,2|3] and ``out_pdf`` [n].
    ``scratch``: caller-owned int32 CUDA buffer of ``sample_scratch_elems(n)`` elements; without it the call is the
    single tensor-core launch (no fp32 fix-up)."""
    _require_cuda(wi, "sample_into")
    wi = _f32c(wi)
    n, cols = wi.shape[0], (2 if epilogue == EPI_RAW else 3)
    _check_rows(wi, n, cols, "wi")
    for t, shape in ((out_dir, (n, cols)), (out_pdf, (n,))):
        if (not t.is_cuda) or t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != shape:
            raise ValueError(f"bsdfdiff.sample_into: output must be a contiguous fp32 CUDA tensor of shape {shape}")
    if scratch is not None and (scratch.dtype != torch.int32 or not scratch.is_cuda or not scratch.is_contiguous()
                                or scratch.numel() < sample_scratch_elems(n)):
        raise ValueError(f"bsdfdiff.sample_into: scratch must be a contiguous int32 CUDA tensor of >= "
                         f"{sample_scratch_elems(n)} elements")
    _check_flow(flow, T, "sample_into")
    x0, u, seed, offset = _noise(wi, x0, u, seed, offset)
    _sample_out_op(wi, flow.blob, base, x0, u, out_dir, out_pdf, scratch, _resolve_precision(precision), flow.domain,
                   epilogue, int(T), flow.hidden, flow.n_hidden, int(seed), int(offset), int(first_index),
                   _fix_thr(fixup, flow.domain, epilogue, "sample"))


def pdf(wo: torch.Tensor, wi: torch.Tensor, flow, base: torch.Tensor, T: int, *, epilogue: int = EPI_RAW,
        precision=None, fixup=None) -> torch.Tensor:
    _require_cuda(wi, "pdf")
    wi = _f32c(wi)
    wo = _f32c(wo, wi.device)
    _check_flow(flow, T, "pdf")
    cols = 2 if epilogue == EPI_RAW else 3
    _check_rows(wi, wi.shape[0], cols, "wi")
    _check_rows(wo, wi.shape[0], cols, "wo")
    return _pdf_op(wo, wi, flow.blob, base, _resolve_precision(precision), flow.domain, epilogue, int(T),
                   flow.hidden, flow.n_hidden, _fix_thr(fixup, flow.domain, epilogue, "pdf"))


def base_log_prob(x: torch.Tensor, wi: torch.Tensor, base: torch.Tensor, domain: int) -> torch.Tensor:
    """log p_base(x | wi) [n] for domain coordinates x, wi [n,2] (``D_base.log_prob``)."""
    _require_cuda(wi, "base_log_prob")
    wi = _f32c(wi)
    x = _f32c(x, wi.device)
    _check_rows(wi, wi.shape[0], 2, "wi")
    _check_rows(x, wi.shape[0], 2, "x")
    return _base_log_prob_op(x, wi, base, int(domain))


def flow_forward(wi: torch.Tensor, flow, T: int, *, n: Optional[int] = None, wi_repeat: int = 1,
                 base: Optional[torch.Tensor] = None, x0: Optional[torch.Tensor] = None,
                 seed: Optional[int] = None, offset: int = 0, first_index: int = 0, precision=None):
    """Forward-only T-step flow (dosampling).  -> (x_T [n,2], x0 [n,2])."""
    _require_cuda(wi, "flow_forward")
    wi = _f32c(wi)
    _check_rows(wi, wi.shape[0], 2, "wi")
    if flow.in_dim not in (25, 26):
        raise ValueError(f"bsdfdiff.flow_forward: the flow nets take 25 or 26 inputs, got {flow.in_dim}")
    if n is None:
        n = wi.shape[0] * wi_repeat
    if n > wi.shape[0] * wi_repeat:
        raise ValueError(f"bsdfdiff.flow_forward: n = {n} exceeds rows * wi_repeat = {wi.shape[0] * wi_repeat}")
    if x0 is not None:
        x0 = _f32c(x0, wi.device)
        _check_rows(x0, int(n), 2, "x0")
        seed, offset = 0, 0
    else:
        if base is None:
            raise ValueError("either x0 or the base net must be given")
        if seed is None:
            seed, offset = next_philox(wi.device)
    return _flow_forward_op(wi, int(wi_repeat), int(n), flow.blob, base, x0, _resolve_precision(precision),
                            flow.domain, int(T), flow.hidden, flow.n_hidden, int(seed), int(offset), int(first_index))


def mlp_forward(x: torch.Tensor, flow, precision=None) -> torch.Tensor:
    _require_cuda(x, "mlp_forward")
    x = _f32c(x)
    if x.shape[1] != flow.in_dim:
        raise RuntimeError(f"bsdfdiff.mlp_forward: expected {flow.in_dim} input columns, got {x.shape[1]}")
    return _mlp_forward_op(x, flow.blob, _resolve_precision(precision), flow.hidden, flow.n_hidden)
