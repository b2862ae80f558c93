"""ctypes front-end of the C/OpenMP oracle (oracle/bsdf_oracle.c).

TEST INFRASTRUCTURE ONLY (see the header of bsdf_oracle.c).  ``build()`` compiles the shared
library next to the source with gcc; the functions take/return numpy float32 arrays.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "bsdf_oracle.c")
_LIB = os.path.join(_HERE, "libbsdf_oracle.so")
_lib = None

EPI_RAW, EPI_DISK, EPI_SPHERICAL, EPI_BSDF = 0, 1, 2, 3


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", _SRC, "-o", _LIB, "-lm"])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _flow_args(flow):
    layers = [np.asarray(w, np.float32) for w in flow.layers]
    flat = np.ascontiguousarray(np.concatenate([w.ravel() for w in layers]))
    H = layers[0].shape[0]
    return flat, int(layers[0].shape[1]), int(H), len(layers) - 1


def num_threads() -> int:
    return int(lib().bsdf_oracle_num_threads())


def sample(flow, base, wi, T, x0, epilogue=EPI_RAW, with_mindet=False):
    """-> (dir [n,2] raw or [n,3] plugin, pdf [n]) [, conditioning weight in (0,1] [n] if with_mindet: 1 = well-conditioned pdf, see bsdf_oracle.c euler()]."""
    wi, x0 = _f(wi), _f(x0)
    n = wi.shape[0]
    flat, in_dim, H, nh = _flow_args(flow)
    b = base.flat()
    out_dir = np.empty((n, 2 if epilogue == EPI_RAW else 3), np.float32)
    out_pdf = np.empty(n, np.float32)
    mindet = np.empty(n, np.float32)
    rc = lib().bsdf_oracle_sample(ctypes.c_int(flow.domain), ctypes.c_int(epilogue), ctypes.c_int(T),
                                  ctypes.c_int64(n), _p(wi), _p(flat), ctypes.c_int(in_dim), ctypes.c_int(H),
                                  ctypes.c_int(nh), _p(b), _p(x0), _p(out_dir), _p(out_pdf),
                                  _p(mindet) if with_mindet else None)
    if rc != 0:
        raise RuntimeError(f"bsdf_oracle_sample failed: {rc}")
    return (out_dir, out_pdf, mindet) if with_mindet else (out_dir, out_pdf)


def pdf(flow, base, wo, wi, T, epilogue=EPI_RAW, with_mindet=False):
    wo, wi = _f(wo), _f(wi)
    n = wi.shape[0]
    flat, in_dim, H, nh = _flow_args(flow)
    b = base.flat()
    out_pdf = np.empty(n, np.float32)
    mindet = np.empty(n, np.float32)
    rc = lib().bsdf_oracle_pdf(ctypes.c_int(flow.domain), ctypes.c_int(epilogue), ctypes.c_int(T), ctypes.c_int64(n),
                               _p(wo), _p(wi), _p(flat), ctypes.c_int(in_dim), ctypes.c_int(H), ctypes.c_int(nh),
                               _p(b), _p(out_pdf), _p(mindet) if with_mindet else None)
    if rc != 0:
        raise RuntimeError(f"bsdf_oracle_pdf failed: {rc}")
    return (out_pdf, mindet) if with_mindet else out_pdf


def reflow(flow, x0, wi, T):
    x0, wi = _f(x0), _f(wi)
    n = wi.shape[0]
    flat, in_dim, H, nh = _flow_args(flow)
    out = np.empty((n, 2), np.float32)
    rc = lib().bsdf_oracle_reflow(ctypes.c_int(flow.domain), ctypes.c_int(T), ctypes.c_int64(n), _p(x0), _p(wi),
                                  _p(flat), ctypes.c_int(in_dim), ctypes.c_int(H), ctypes.c_int(nh), _p(out))
    if rc != 0:
        raise RuntimeError(f"bsdf_oracle_reflow failed: {rc}")
    return out
