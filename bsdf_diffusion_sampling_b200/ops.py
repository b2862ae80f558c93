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
    controlling the sampler, like the reference's ``torch.randn_like`` (rendering/utils/model.py:390)."""
    gen = torch.cuda.default_generators[torch.device(device).index or 0]
    seed, offset = gen.initial_seed(), gen.get_offset()
    gen.set_offset(offset + PHILOX_OFFSET_PER_CALL)
    return seed & (2 ** 63 - 1), offset


# ------------------------------------------------------------------------------------------------
# custom ops.  int arguments: precision, domain, epilogue, T, hidden, n_hidden, seed, offset, first_index
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("bsdfdiff::sample", mutates_args=(), device_types="cuda")
def _sample_op(wi: torch.Tensor, flow_blob: Optional[torch.Tensor], base: torch.Tensor, x0: Optional[torch.Tensor],
               precision: int, domain: int, epilogue: int, T: int, hidden: int, n_hidden: int,
               seed: int, offset: int, first_index: int,
               want_x0: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    n = wi.shape[0]
    out_dir = torch.empty((n, 2 if epilogue == EPI_RAW else 3), dtype=torch.float32, device=wi.device)
    out_pdf = torch.empty((n,), dtype=torch.float32, device=wi.device)
    out_x0 = torch.empty((n if want_x0 else 0, 2), dtype=torch.float32, device=wi.device)
    with torch.cuda.device(wi.device):
        rc = _lib.lib.bsdfdiff_sample(precision, domain, epilogue, T, n, wi.data_ptr(),
                                      flow_blob.data_ptr() if flow_blob is not None else None,
                                      hidden, n_hidden, base.data_ptr(), x0.data_ptr() if x0 is not None else None,
                                      seed, offset, first_index, out_dir.data_ptr(), out_pdf.data_ptr(),
                                      out_x0.data_ptr() if want_x0 else None, _stream(wi))
    _lib.check(rc, "bsdfdiff_sample")
    return out_dir, out_pdf, out_x0


@_sample_op.register_fake
def _(wi, flow_blob, base, x0, precision, domain, epilogue, T, hidden, n_hidden, seed, offset, first_index, want_x0):
    n = wi.shape[0]
    return (wi.new_empty((n, 2 if epilogue == EPI_RAW else 3)), wi.new_empty((n,)),
            wi.new_empty((n if want_x0 else 0, 2)))


@torch.library.custom_op("bsdfdiff::sample_out", mutates_args=("out_dir", "out_pdf"), device_types="cuda")
def _sample_out_op(wi: torch.Tensor, flow_blob: Optional[torch.Tensor], base: torch.Tensor, x0: Optional[torch.Tensor],
                   out_dir: torch.Tensor, out_pdf: torch.Tensor,
                   precision: int, domain: int, epilogue: int, T: int, hidden: int, n_hidden: int,
                   seed: int, offset: int, first_index: int) -> None:
    """As ``bsdfdiff::sample`` but writes into caller-owned buffers (no allocation: streaming pipelines)."""
    n = wi.shape[0]
    with torch.cuda.device(wi.device):
        rc = _lib.lib.bsdfdiff_sample(precision, domain, epilogue, T, n, wi.data_ptr(),
                                      flow_blob.data_ptr() if flow_blob is not None else None,
                                      hidden, n_hidden, base.data_ptr(), x0.data_ptr() if x0 is not None else None,
                                      seed, offset, first_index, out_dir.data_ptr(), out_pdf.data_ptr(), None,
                                      _stream(wi))
    _lib.check(rc, "bsdfdiff_sample")


@torch.library.custom_op("bsdfdiff::pdf", mutates_args=(), device_types="cuda")
def _pdf_op(wo: torch.Tensor, wi: torch.Tensor, flow_blob: Optional[torch.Tensor], base: torch.Tensor,
            precision: int, domain: int, epilogue: int, T: int, hidden: int, n_hidden: int) -> torch.Tensor:
    n = wi.shape[0]
    out = torch.empty((n,), dtype=torch.float32, device=wi.device)
    with torch.cuda.device(wi.device):
        rc = _lib.lib.bsdfdiff_pdf(precision, domain, epilogue, T, n, wo.data_ptr(), wi.data_ptr(),
                                   flow_blob.data_ptr() if flow_blob is not None else None, hidden, n_hidden,
                                   base.data_ptr(), out.data_ptr(),
                                   _stream(wi))
    _lib.check(rc, "bsdfdiff_pdf")
    return out


@_pdf_op.register_fake
def _(wo, wi, flow_blob, base, precision, domain, epilogue, T, hidden, n_hidden):
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


def sample(wi: torch.Tensor, flow, base: torch.Tensor, T: int, *, epilogue: int = EPI_RAW,
           x0: Optional[torch.Tensor] = None, seed: Optional[int] = None, offset: int = 0, first_index: int = 0,
           precision=None, return_x0: bool = True):
    """-> (dir [n,2|3], pdf [n], x0 [n,2]).  ``flow`` is a ``weights.PackedFlow``.
    ``return_x0=False`` skips materialising the base sample (8 B/query of HBM writes); x0 is then empty."""
    _require_cuda(wi, "sample")
    wi = _f32c(wi)
    if T < 0 or (T == 0) != (flow.blob is None):
        raise ValueError("T must be >= 1 (T == 0 only with NullFlow: base distribution alone)")
    if x0 is not None:
        x0 = _f32c(x0, wi.device)
        seed, offset = 0, 0
    elif seed is None:
        seed, offset = next_philox(wi.device)
    return _sample_op(wi, flow.blob, base, x0, _resolve_precision(precision), flow.domain, epilogue, int(T),
                      flow.hidden, flow.n_hidden, int(seed), int(offset), int(first_index), bool(return_x0))


def sample_into(wi: torch.Tensor, flow, base: torch.Tensor, T: int, out_dir: torch.Tensor, out_pdf: torch.Tensor, *,
                epilogue: int = EPI_RAW, x0: Optional[torch.Tensor] = None, seed: Optional[int] = None, offset: int = 0,
                first_index: int = 0, precision=None) -> None:
    """``sample`` into caller-owned contiguous fp32 CUDA buffers ``out_dir`` [n,2|3] and ``out_pdf`` [n]."""
    _require_cuda(wi, "sample_into")
    wi = _f32c(wi)
    n, cols = wi.shape[0], (2 if epilogue == EPI_RAW else 3)
    for t, shape in ((out_dir, (n, cols)), (out_pdf, (n,))):
        if (not t.is_cuda) or t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != shape:
            raise ValueError(f"bsdfdiff.sample_into: output must be a contiguous fp32 CUDA tensor of shape {shape}")
    if T < 0 or (T == 0) != (flow.blob is None):
        raise ValueError("T must be >= 1 (T == 0 only with NullFlow: base distribution alone)")
    if x0 is not None:
        x0 = _f32c(x0, wi.device)
        seed, offset = 0, 0
    elif seed is None:
        seed, offset = next_philox(wi.device)
    _sample_out_op(wi, flow.blob, base, x0, out_dir, out_pdf, _resolve_precision(precision), flow.domain, epilogue,
                   int(T), flow.hidden, flow.n_hidden, int(seed), int(offset), int(first_index))


def pdf(wo: torch.Tensor, wi: torch.Tensor, flow, base: torch.Tensor, T: int, *, epilogue: int = EPI_RAW,
        precision=None) -> torch.Tensor:
    _require_cuda(wi, "pdf")
    wi = _f32c(wi)
    wo = _f32c(wo, wi.device)
    if T < 0 or (T == 0) != (flow.blob is None):
        raise ValueError("T must be >= 1 (T == 0 only with NullFlow: base distribution alone)")
    if wo.shape[0] != wi.shape[0]:
        raise ValueError("wo and wi must have the same number of rows")
    return _pdf_op(wo, wi, flow.blob, base, _resolve_precision(precision), flow.domain, epilogue, int(T),
                   flow.hidden, flow.n_hidden)


def flow_forward(wi: torch.Tensor, flow, T: int, *, n: Optional[int] = None, wi_repeat: int = 1,
                 base: Optional[torch.Tensor] = None, x0: Optional[torch.Tensor] = None,
                 seed: Optional[int] = None, offset: int = 0, first_index: int = 0, precision=None):
    """Forward-only T-step flow (dosampling).  -> (x_T [n,2], x0 [n,2])."""
    _require_cuda(wi, "flow_forward")
    wi = _f32c(wi)
    if n is None:
        n = wi.shape[0] * wi_repeat
    if x0 is not None:
        x0 = _f32c(x0, wi.device)
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
