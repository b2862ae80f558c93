"""Ground-truth measured BSDF on the GPU: the RGL material-database tensor files the reference renders against.

Replaces, for the two measured plugins, Mitsuba's ``measured`` BSDF object
(``mi.load_dict({"type": "measured", "filename": "./measuredbsdfs/<mat>.bsdf"})``, rendering/brdf_measured_disk.py:36-42,
rendering/brdf_measured_spherical.py:45-51):

* ``MeasuredBSDF.from_file(path)`` parses the tensor file (fields ``theta_i, phi_i, ndf, sigma, vndf, rgb, jacobian``) and packs
  it once into a device blob (``bsdfdiff_measured_pack``: VNDF normalisation and CDFs as Mitsuba's ``Marginal2D`` builds them);
* ``eval(wi, wo)`` = ``self.bsdf.eval(ctx, si, wo)`` on [n,3] local-frame torch tensors (one kernel launch);
* ``weight_and_clamp(kind, wi, wo, bs_pdf, albedo)`` = the tail of ``MyBSDF.sample`` -- throughput weight, firefly clamp
  (luminance < 30) and the final masks (brdf_measured_disk.py:92-101, brdf_measured_spherical.py:100-109) -- fused into one
  launch, so a ``sample`` call needs no Dr.Jit <-> torch round trip after the sampler kernel.

The full-sphere ``bsdf`` plugin kind evaluates Mitsuba's analytic ``principled`` / ``roughdielectric`` models instead
(rendering/bsdf_myresult.py:46, utils/bsdf_dict.py) and is not covered here.
"""
from __future__ import annotations

import struct
from typing import Dict, Sequence, Tuple

import numpy as np
import torch

from . import _lib

_DTYPES = {1: np.uint8, 2: np.int8, 3: np.uint16, 4: np.int16, 5: np.uint32, 6: np.int32, 7: np.uint64, 8: np.int64,
           9: np.float16, 10: np.float32, 11: np.float64}
FIREFLY_LUMINANCE = 30.0          # brdf_measured_disk.py:98, brdf_measured_spherical.py:106


def read_tensor_file(path: str) -> Dict[str, np.ndarray]:
    """Mitsuba ``TensorFile`` container: 12-byte magic, version 1.0, field table (name, ndim, dtype, offset, shape)."""
    with open(path, "rb") as fh:
        b = fh.read()
    if b[:12] != b"tensor_file\0" or struct.unpack("<BB", b[12:14]) != (1, 0):
        raise ValueError(f"{path}: not a version-1.0 tensor file")
    n = struct.unpack("<I", b[14:18])[0]
    off, fields = 18, {}
    for _ in range(n):
        ln = struct.unpack("<H", b[off:off + 2])[0]
        name = b[off + 2:off + 2 + ln].decode()
        off += 2 + ln
        nd, dt = struct.unpack("<HB", b[off:off + 3])
        o = struct.unpack("<Q", b[off + 3:off + 11])[0]
        shape = struct.unpack("<%dQ" % nd, b[off + 11:off + 11 + 8 * nd])
        off += 11 + 8 * nd
        if dt not in _DTYPES:
            raise ValueError(f"{path}: field {name!r} has unknown dtype {dt}")
        fields[name] = np.frombuffer(b, _DTYPES[dt], int(np.prod(shape)) if nd else 1, o).reshape(shape)
    return fields


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


class MeasuredBSDF:
    def __init__(self, fields: Dict[str, np.ndarray], device="cuda"):
        need = ("theta_i", "phi_i", "ndf", "sigma", "vndf", "rgb", "jacobian")
        missing = [k for k in need if k not in fields]
        if missing:
            raise ValueError(f"measured BSDF: tensor file lacks the fields {missing} (rgb-variant RGL file expected)")
        phi, theta = _f32(fields["phi_i"]), _f32(fields["theta_i"])
        ndf, sigma, vndf, rgb = (_f32(fields[k]) for k in ("ndf", "sigma", "vndf", "rgb"))
        if vndf.shape[:2] != (phi.size, theta.size) or rgb.shape[:3] != (phi.size, theta.size, 3) or ndf.ndim != 2 \
                or sigma.ndim != 2:
            raise ValueError("measured BSDF: inconsistent tensor shapes")
        dims = (phi.size, theta.size, ndf.shape[1], ndf.shape[0], sigma.shape[1], sigma.shape[0], vndf.shape[3],
                vndf.shape[2], rgb.shape[4], rgb.shape[3])
        nbytes = _lib.lib.bsdfdiff_measured_blob_bytes(*dims)
        if nbytes == 0:
            raise _lib.BsdfDiffError(f"measured BSDF: unsupported table sizes {dims}")
        out = np.zeros(nbytes, np.uint8)
        jac = int(np.asarray(fields["jacobian"]).ravel()[0])
        _lib.check(_lib.lib.bsdfdiff_measured_pack(phi.ctypes.data, dims[0], theta.ctypes.data, dims[1], ndf.ctypes.data,
                                                   dims[2], dims[3], sigma.ctypes.data, dims[4], dims[5], vndf.ctypes.data,
                                                   dims[6], dims[7], rgb.ctypes.data, dims[8], dims[9], jac,
                                                   out.ctypes.data), "bsdfdiff_measured_pack")
        self.blob = torch.from_numpy(out).to(device)
        self.isotropic = phi.size <= 2
        self.n_phi, self.n_theta = int(phi.size), int(theta.size)

    @classmethod
    def from_file(cls, path: str, device="cuda") -> "MeasuredBSDF":
        return cls(read_tensor_file(path), device)

    def _check(self, *ts: torch.Tensor) -> Sequence[torch.Tensor]:
        out = []
        for t in ts:
            if not t.is_cuda:
                raise RuntimeError("measured BSDF: expected CUDA tensors; this package has no CPU path")
            out.append(t.detach().to(self.blob.device, torch.float32).contiguous())
        n = out[0].shape[0]
        for t in out[:2]:
            if t.dim() != 2 or t.shape != (n, 3):
                raise ValueError(f"measured BSDF: directions must have shape ({n}, 3), got {tuple(t.shape)}")
        return out

    def eval(self, wi: torch.Tensor, wo: torch.Tensor) -> torch.Tensor:
        """``mi.BSDF.eval`` of the measured model (f cos(theta_o); 0 outside the upper hemispheres): [n,3] rgb."""
        wi, wo = self._check(wi, wo)
        out = torch.empty_like(wi)
        with torch.cuda.device(wi.device):
            rc = _lib.lib.bsdfdiff_measured_eval(self.blob.data_ptr(), wi.shape[0], wi.data_ptr(), wo.data_ptr(),
                                                 out.data_ptr(), torch.cuda.current_stream(wi.device).cuda_stream)
        _lib.check(rc, "bsdfdiff_measured_eval")
        return out

    def weight_and_clamp(self, epilogue: int, wi: torch.Tensor, wo: torch.Tensor, bs_pdf: torch.Tensor,
                         albedo=(1.0, 1.0, 1.0), clamp: float = FIREFLY_LUMINANCE) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (weight [n,3], pdf [n]): ``value = eval / bs_pdf * albedo``, firefly clamp on its luminance, final masks."""
        wi, wo, bs_pdf = self._check(wi, wo, bs_pdf)
        n = wi.shape[0]
        if bs_pdf.shape != (n,):
            raise ValueError(f"measured BSDF: bs_pdf must have shape ({n},)")
        w = torch.empty_like(wi)
        pdf = torch.empty_like(bs_pdf)
        with torch.cuda.device(wi.device):
            rc = _lib.lib.bsdfdiff_measured_weight(self.blob.data_ptr(), int(epilogue), n, wi.data_ptr(), wo.data_ptr(),
                                                   bs_pdf.data_ptr(), float(albedo[0]), float(albedo[1]), float(albedo[2]),
                                                   float(clamp), w.data_ptr(), pdf.data_ptr(),
                                                   torch.cuda.current_stream(wi.device).cuda_stream)
        _lib.check(rc, "bsdfdiff_measured_weight")
        return w, pdf
