"""ctypes binding of libbsdfdiff.so (include/bsdfdiff.h).

The library is the product: there is no Python/PyTorch fallback.  If it has not been built
(``python -m bsdf_diffusion_sampling_b200.build``) importing this module raises ``ImportError``;
if it is called without a CUDA device the entry points raise ``RuntimeError``.
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BSDFDIFF_LIB") or os.path.join(HERE, "libbsdfdiff.so")   # env override: tuning builds only

# constants mirrored from include/bsdfdiff.h
DISK, SPHERICAL = 0, 1
EPI_RAW, EPI_DISK, EPI_SPHERICAL, EPI_BSDF = 0, 1, 2, 3
PREC_FP32, PREC_TC16, PREC_TC16_EXP = 0, 1, 2
BASE_FLOATS = 308
ABI_VERSION = 3
OK_FP32_REROUTE = 1

EXPORTS = [
    "bsdfdiff_abi_version", "bsdfdiff_error_string", "bsdfdiff_last_cuda_error", "bsdfdiff_debug_timeout_flag", "bsdfdiff_debug_trace",
    "bsdfdiff_device_info",
    "bsdfdiff_packed_flow_bytes", "bsdfdiff_pack_flow", "bsdfdiff_pack_flow_tcnn", "bsdfdiff_fixup_scratch_bytes",
    "bsdfdiff_sample", "bsdfdiff_pdf", "bsdfdiff_base_log_prob", "bsdfdiff_flow_forward", "bsdfdiff_mlp_forward",
    "bsdfdiff_measured_blob_bytes", "bsdfdiff_measured_pack", "bsdfdiff_measured_eval", "bsdfdiff_measured_weight",
    "bsdfdiff_multi_scratch_bytes", "bsdfdiff_multi_plan", "bsdfdiff_sample_multi", "bsdfdiff_pdf_multi",
    "bsdfdiff_sample_planar", "bsdfdiff_pdf_planar",
    "bsdfdiff_flow_param_count", "bsdfdiff_flow_matching_step", "bsdfdiff_base_nll_step",
]

_c = ctypes
_vp, _i, _i64, _u64, _f = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_uint64, _c.c_float


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m bsdf_diffusion_sampling_b200.build` "
            "(nvcc, sm_100a). There is no CPU / PyTorch fallback for this package.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.bsdfdiff_abi_version.restype = _i
    lib.bsdfdiff_error_string.restype = _c.c_char_p
    lib.bsdfdiff_error_string.argtypes = [_i]
    lib.bsdfdiff_last_cuda_error.restype = _i
    lib.bsdfdiff_debug_trace.restype = _i
    lib.bsdfdiff_debug_trace.argtypes = [_vp, _i]
    lib.bsdfdiff_device_info.argtypes = [_c.POINTER(_i)] * 3
    lib.bsdfdiff_packed_flow_bytes.restype = _c.c_size_t
    lib.bsdfdiff_packed_flow_bytes.argtypes = [_i, _i, _i]
    lib.bsdfdiff_pack_flow.argtypes = [_c.POINTER(_vp), _c.POINTER(_i), _c.POINTER(_i), _i, _vp]
    lib.bsdfdiff_pack_flow_tcnn.argtypes = [_vp, _i, _i, _i, _i, _vp]
    lib.bsdfdiff_fixup_scratch_bytes.restype = _c.c_size_t
    lib.bsdfdiff_fixup_scratch_bytes.argtypes = [_i64]
    lib.bsdfdiff_sample.argtypes = [_i, _i, _i, _i, _i64, _vp, _vp, _i, _i, _vp, _vp, _vp, _u64, _u64, _i64,
                                    _vp, _vp, _vp, _f, _vp, _vp]
    lib.bsdfdiff_pdf.argtypes = [_i, _i, _i, _i, _i64, _vp, _vp, _vp, _i, _i, _vp, _vp, _f, _vp, _vp]
    lib.bsdfdiff_base_log_prob.argtypes = [_i, _i64, _vp, _vp, _vp, _vp, _vp]
    lib.bsdfdiff_flow_forward.argtypes = [_i, _i, _i, _i64, _vp, _i64, _vp, _i, _i, _vp, _vp, _u64, _u64, _i64,
                                          _vp, _vp, _vp]
    lib.bsdfdiff_mlp_forward.argtypes = [_i, _i64, _vp, _i, _vp, _i, _i, _vp, _vp]
    lib.bsdfdiff_measured_blob_bytes.restype = _c.c_size_t
    lib.bsdfdiff_measured_blob_bytes.argtypes = [_i] * 10
    lib.bsdfdiff_measured_pack.argtypes = [_vp, _i, _vp, _i, _vp, _i, _i, _vp, _i, _i, _vp, _i, _i, _vp, _i, _i, _i, _vp]
    lib.bsdfdiff_measured_eval.argtypes = [_vp, _i64, _vp, _vp, _vp, _vp]
    lib.bsdfdiff_measured_weight.argtypes = [_vp, _i, _i64, _vp, _vp, _vp, _f, _f, _f, _f, _vp, _vp, _vp]
    lib.bsdfdiff_multi_scratch_bytes.restype = _c.c_size_t
    lib.bsdfdiff_multi_scratch_bytes.argtypes = [_i64, _i]
    lib.bsdfdiff_multi_plan.argtypes = [_i64, _vp, _i, _vp, _vp]
    lib.bsdfdiff_sample_multi.argtypes = [_i, _i, _i, _i, _i64, _vp, _vp, _i, _vp, _vp, _i, _i, _vp, _vp, _u64, _u64, _i64,
                                          _vp, _vp, _vp, _f, _vp]
    lib.bsdfdiff_pdf_multi.argtypes = [_i, _i, _i, _i, _i64, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _vp, _f, _vp]
    lib.bsdfdiff_flow_param_count.restype = _c.c_size_t
    lib.bsdfdiff_flow_param_count.argtypes = [_i, _i, _i]
    lib.bsdfdiff_flow_matching_step.restype = _i
    lib.bsdfdiff_flow_matching_step.argtypes = [_i, _i, _i, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _f,
                                                _i64, _i, _vp, _vp, _vp]
    lib.bsdfdiff_base_nll_step.restype = _i
    lib.bsdfdiff_base_nll_step.argtypes = [_i, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _i64, _i, _vp, _vp, _vp]
    _p3 = _c.POINTER(_vp)
    lib.bsdfdiff_sample_planar.argtypes = [_i, _i, _i, _i, _i64, _p3, _vp, _i, _i, _vp, _vp, _vp, _u64, _u64, _i64,
                                           _p3, _vp, _vp, _f, _vp, _vp]
    lib.bsdfdiff_pdf_planar.argtypes = [_i, _i, _i, _i, _i64, _p3, _p3, _vp, _i, _i, _vp, _vp, _f, _vp, _vp]
    for name in ("bsdfdiff_multi_plan", "bsdfdiff_sample_multi", "bsdfdiff_pdf_multi", "bsdfdiff_sample_planar",
                 "bsdfdiff_pdf_planar"):
        getattr(lib, name).restype = _i
    for name in ("bsdfdiff_measured_pack", "bsdfdiff_measured_eval", "bsdfdiff_measured_weight"):
        getattr(lib, name).restype = _i
    for name in ("bsdfdiff_pack_flow", "bsdfdiff_pack_flow_tcnn", "bsdfdiff_sample", "bsdfdiff_pdf",
                 "bsdfdiff_base_log_prob",
                 "bsdfdiff_flow_forward", "bsdfdiff_mlp_forward", "bsdfdiff_device_info"):
        getattr(lib, name).restype = _i
    if lib.bsdfdiff_abi_version() != ABI_VERSION:
        raise ImportError("libbsdfdiff.so ABI version mismatch; rebuild")
    return lib


lib = _load()


class BsdfDiffError(RuntimeError):
    pass


# number of calls the library answered with BSDFDIFF_OK_FP32_REROUTE (a tensor-core request whose net shape only the
# fp32 CUDA-core kernel covers); BSDFDIFF_STRICT_TC=1 turns the reroute into an error
fp32_reroutes = 0


def check(rc: int, what: str) -> None:
    global fp32_reroutes
    if rc == OK_FP32_REROUTE:
        fp32_reroutes += 1
        if os.environ.get("BSDFDIFF_STRICT_TC") == "1":
            raise BsdfDiffError(f"{what}: {lib.bsdfdiff_error_string(rc).decode()} (BSDFDIFF_STRICT_TC=1)")
        return
    if rc != 0:
        msg = lib.bsdfdiff_error_string(rc).decode()
        extra = f" (cudaError {lib.bsdfdiff_last_cuda_error()})" if rc == -3 else ""
        raise BsdfDiffError(f"{what}: {msg}{extra}")
