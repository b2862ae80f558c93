"""CPU restatement (numpy, float64) of the measured-BSDF ground truth the reference's plugins evaluate.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's checking legs; the product path
(bsdf_diffusion_sampling_b200/measured.py + csrc/measured.cu) never imports it.

What it restates.  The reference does not contain this code: its plugins call Mitsuba 3's ``measured`` BSDF
(``mi.load_dict({"type": "measured", "filename": ...})``, rendering/brdf_measured_disk.py:36-42,
rendering/brdf_measured_spherical.py:45-51) and use ``self.bsdf.eval(ctx, si, wo)`` for the throughput weight and the
firefly clamp (brdf_measured_disk.py:92-100, brdf_measured_spherical.py:100-108); the training targets use the same
call (learning_repo_cleanup/utils/mitsuba_brdf_scalar.py:75-89).  Mitsuba 3 is a third-party dependency that is NOT
vendored and whose version the reference does not pin (no requirements file); it is absent from the build container and
cannot be installed.  This file therefore restates the PUBLISHED algorithm:
  * the RGL material database tensor-file container (Mitsuba ``TensorFile``: magic "tensor_file\\0", version 1.0, field
    table of name / ndim / dtype / offset / shape) with fields ``theta_i, phi_i, ndf, sigma, vndf, luminance, rgb,
    jacobian`` (rendering/measuredbsdfs/*.bsdf);
  * Dupuy & Jakob, "An Adaptive Parameterization for Efficient Material Acquisition and Rendering" (SIGGRAPH Asia 2018),
    as implemented by Mitsuba 3 ``src/bsdfs/measured.cpp`` (eval) and ``include/mitsuba/core/distr_2d.h``
    (``Marginal2D<Float, Dimension, Continuous = true>``: bilinear patches, trapezoid CDFs accumulated in double,
    multilinear interpolation over the (phi_i, theta_i[, channel]) parameter axes):
        eval(wi, wo) = rgb(invert_vndf(u_m) | phi_i, theta_i) * ndf(u_m) / (4 sigma(u_wi))
    with u = (sqrt(2 theta / pi), (phi + pi) / (2 pi)), m = normalize(wi + wo), phi_m relative to phi_i for isotropic data.
    Cosine convention: the tabulated rgb values are the measured reflectance INCLUDING the foreshortening factor, so eval
    (which by Mitsuba's convention returns f cos(theta_o)) is the table expression as it stands, with no further cos
    factor.  Mitsuba cannot be consulted here; profiles/measured_convention_check.py decides it with the reference's own
    trained samplers (their density follows eval / cos(theta_o), not eval alone: mean TV 0.110 vs 0.132, 16 of 18 cases).

PARITY UNPINNED against Mitsuba itself (it cannot run here, and no output of it is stored in the reference).  What pins
this restatement instead (tests/test_measured.py): (a) structural identities of the model -- the VNDF warp integrates
to one, invert() is the inverse of the warp's sample(), the hemispherical albedo is physical; (b) the reference's OWN
shipped networks, which were trained on Mitsuba's eval of these very files: the flow's sampling density in disk
coordinates must follow lum(f cos) * clamp(1/cos, 1, 1e6) normalised (the reference's training target,
mitsuba_brdf_scalar.py:83-89) -- a restatement with a wrong parameterisation, Jacobian or cosine convention fails that
test by a wide margin.
"""
from __future__ import annotations

import struct

import numpy as np

_DTYPES = {1: np.uint8, 2: np.int8, 3: np.uint16, 4: np.int16, 5: np.uint32, 6: np.int32, 7: np.uint64, 8: np.int64,
           9: np.float16, 10: np.float32, 11: np.float64}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}


def read_tensor_file(path: str) -> dict:
    """Mitsuba TensorFile reader (mitsuba/core/tensor.h): {name: ndarray}."""
    b = open(path, "rb").read()
    if b[:12] != b"tensor_file\0":
        raise ValueError(f"{path}: not a tensor file")
    major, minor = struct.unpack("<BB", b[12:14])
    if (major, minor) != (1, 0):
        raise ValueError(f"{path}: unsupported tensor-file version {major}.{minor}")
    n = struct.unpack("<I", b[14:18])[0]
    off, fields = 18, {}
    for _ in range(n):
        ln = struct.unpack("<H", b[off:off + 2])[0]; off += 2
        name = b[off:off + ln].decode(); off += ln
        nd = struct.unpack("<H", b[off:off + 2])[0]; off += 2
        dt = b[off]; off += 1
        o = struct.unpack("<Q", b[off:off + 8])[0]; off += 8
        shape = struct.unpack("<%dQ" % nd, b[off:off + 8 * nd]); off += 8 * nd
        cnt = int(np.prod(shape)) if nd else 1
        fields[name] = np.frombuffer(b, _DTYPES[dt], cnt, o).reshape(shape)
    return fields


def write_tensor_file(path: str, fields: dict) -> None:
    """Inverse of read_tensor_file (used by tests to build synthetic anisotropic materials)."""
    names = list(fields)
    arrs = [np.ascontiguousarray(fields[k]) for k in names]
    head = 18 + sum(2 + len(k.encode()) + 2 + 1 + 8 + 8 * a.ndim for k, a in zip(names, arrs))
    offs, pos = [], (head + 63) // 64 * 64
    for a in arrs:
        offs.append(pos)
        pos = (pos + a.nbytes + 63) // 64 * 64
    out = bytearray(pos)
    out[:12] = b"tensor_file\0"
    out[12:14] = struct.pack("<BB", 1, 0)
    out[14:18] = struct.pack("<I", len(names))
    p = 18
    for k, a, o in zip(names, arrs, offs):
        kb = k.encode()
        out[p:p + 2] = struct.pack("<H", len(kb)); p += 2
        out[p:p + len(kb)] = kb; p += len(kb)
        out[p:p + 2] = struct.pack("<H", a.ndim); p += 2
        out[p] = _DTYPE_IDS[a.dtype]; p += 1
        out[p:p + 8] = struct.pack("<Q", o); p += 8
        out[p:p + 8 * a.ndim] = struct.pack("<%dQ" % a.ndim, *a.shape); p += 8 * a.ndim
        out[o:o + a.nbytes] = a.tobytes()
    open(path, "wb").write(bytes(out))


class Marginal2D:
    """``Marginal2D<Float, len(param_values), Continuous=true>`` of mitsuba/core/distr_2d.h.

    data: [*param_res, H, W]; the distribution lives on [0,1]^2 with x along W and y along H; values are interpolated
    bilinearly inside the (W-1) x (H-1) patches and multilinearly across the parameter axes."""

    def __init__(self, data, param_values=(), normalize=False):
        data = np.asarray(data, np.float64)
        self.param_values = [np.asarray(p, np.float64) for p in param_values]
        self.nparam = len(self.param_values)
        assert data.ndim == self.nparam + 2 and tuple(data.shape[:self.nparam]) == tuple(len(p) for p in self.param_values)
        self.h, self.w = data.shape[-2:]
        d = data.reshape(-1, self.h, self.w).copy()
        # conditional CDF along x per row, marginal CDF along y (trapezoid rule in units of one patch, in double)
        cond = np.cumsum(0.5 * (d[:, :, :-1] + d[:, :, 1:]), axis=2)                    # [S, H, W-1]
        marg = np.cumsum(0.5 * (cond[:, :-1, -1] + cond[:, 1:, -1]), axis=1)              # [S, H-1]
        if normalize:
            norm = 1.0 / marg[:, -1]
            cond *= norm[:, None, None]
            marg *= norm[:, None]
            d *= (norm * (self.w - 1) * (self.h - 1))[:, None, None]                      # pdf w.r.t. the unit square
        # Mitsuba stores these tables in single precision
        self.data = d.astype(np.float32).astype(np.float64)
        self.cond = cond.astype(np.float32).astype(np.float64)
        self.marg = marg.astype(np.float32).astype(np.float64)
        self.normalized = normalize

    # -- parameter axes ---------------------------------------------------------------------------------------------
    def _slices(self, params):
        """[(slice index [n], weight [n])] for the 2^nparam corners of the parameter cell."""
        n = len(params[0]) if self.nparam else None
        combos = [(np.zeros(n if n is not None else 1, np.int64), np.ones(n if n is not None else 1))]
        stride = 1
        strides = []
        for pv in reversed(self.param_values):
            strides.append(stride)
            stride *= len(pv)
        strides = strides[::-1]
        for dim, pv in enumerate(self.param_values):
            x = np.asarray(params[dim], np.float64)
            if len(pv) == 1:
                continue
            # find_interval: largest i in [0, size-2] with pv[i] <= x
            i = np.clip(np.searchsorted(pv, x, side="right") - 1, 0, len(pv) - 2)
            w1 = np.clip((x - pv[i]) / (pv[i + 1] - pv[i]), 0.0, 1.0)
            new = []
            for idx, wgt in combos:
                new.append((idx + strides[dim] * i, wgt * (1.0 - w1)))
                new.append((idx + strides[dim] * (i + 1), wgt * w1))
            combos = new
        return combos

    def _lookup(self, table, combos, *index):
        out = 0.0
        for idx, wgt in combos:
            out = out + wgt * table[(idx,) + index]
        return out

    # -- evaluation ----------------------------------------------------------------------------------------------------
    def eval(self, pos, params=()):
        pos = np.asarray(pos, np.float64)
        combos = self._slices(params) if self.nparam else [(np.zeros(pos.shape[0], np.int64), np.ones(pos.shape[0]))]
        px, py = pos[:, 0] * (self.w - 1), pos[:, 1] * (self.h - 1)
        ox = np.minimum(np.maximum(px, 0).astype(np.int64), self.w - 2)
        oy = np.minimum(np.maximum(py, 0).astype(np.int64), self.h - 2)
        wx1, wy1 = px - ox, py - oy
        v00 = self._lookup(self.data, combos, oy, ox)
        v10 = self._lookup(self.data, combos, oy, ox + 1)
        v01 = self._lookup(self.data, combos, oy + 1, ox)
        v11 = self._lookup(self.data, combos, oy + 1, ox + 1)
        return (1 - wy1) * ((1 - wx1) * v00 + wx1 * v10) + wy1 * ((1 - wx1) * v01 + wx1 * v11)

    def invert(self, pos, params=()):
        """Position in the distribution's domain -> the uniform sample that the warp maps there, and the pdf.
        u_y = marginal CDF at y, u_x = conditional CDF at x given y (both of the bilinear interpolant)."""
        pos = np.asarray(pos, np.float64)
        combos = self._slices(params) if self.nparam else [(np.zeros(pos.shape[0], np.int64), np.ones(pos.shape[0]))]
        px, py = pos[:, 0] * (self.w - 1), pos[:, 1] * (self.h - 1)
        ox = np.minimum(np.maximum(px, 0).astype(np.int64), self.w - 2)
        oy = np.minimum(np.maximum(py, 0).astype(np.int64), self.h - 2)
        sx, sy = px - ox, py - oy
        v00 = self._lookup(self.data, combos, oy, ox)
        v10 = self._lookup(self.data, combos, oy, ox + 1)
        v01 = self._lookup(self.data, combos, oy + 1, ox)
        v11 = self._lookup(self.data, combos, oy + 1, ox + 1)
        inv_area = 1.0 / ((self.w - 1) * (self.h - 1)) if self.normalized else 1.0      # tables hold pdf values; CDFs patch units
        c0 = (1 - sy) * v00 + sy * v01                 # the interpolant on the patch's left / right edge at this y
        c1 = (1 - sy) * v10 + sy * v11
        pdf = (1 - sx) * c0 + sx * c1
        # x: integral of the row interpolant from the patch's left edge to x, plus the full patches to the left
        part_x = (sx * c0 + 0.5 * sx * sx * (c1 - c0)) * inv_area
        has_left = ox > 0
        oxl = np.maximum(ox - 1, 0)
        left0 = np.where(has_left, self._lookup(self.cond, combos, oy, oxl), 0.0)
        left1 = np.where(has_left, self._lookup(self.cond, combos, oy + 1, oxl), 0.0)
        row0 = self._lookup(self.cond, combos, oy, np.full_like(ox, self.w - 2))           # full-row integrals
        row1 = self._lookup(self.cond, combos, oy + 1, np.full_like(ox, self.w - 2))
        num = (1 - sy) * left0 + sy * left1 + part_x
        den = (1 - sy) * row0 + sy * row1
        ux = num / den
        # y: integral of the row integral from the patch's lower edge to y, plus the full patch rows below
        part_y = sy * row0 + 0.5 * sy * sy * (row1 - row0)
        has_below = oy > 0
        below = np.where(has_below, self._lookup(self.marg, combos, np.maximum(oy - 1, 0)), 0.0)
        uy = below + part_y
        return np.stack([ux, uy], 1), pdf

    def sample(self, u, params=()):
        """The warp itself (uniform square -> domain) by bisection on invert(); O(40 n) -- tests only."""
        u = np.asarray(u, np.float64)
        n = u.shape[0]
        lo, hi = np.zeros(n), np.ones(n)
        for _ in range(48):                                               # y from the marginal
            mid = 0.5 * (lo + hi)
            uy = self.invert(np.stack([np.zeros(n), mid], 1), params)[0][:, 1]
            big = uy > u[:, 1]
            hi, lo = np.where(big, mid, hi), np.where(big, lo, mid)
        y = 0.5 * (lo + hi)
        lo, hi = np.zeros(n), np.ones(n)
        for _ in range(48):                                               # x from the conditional at y
            mid = 0.5 * (lo + hi)
            ux = self.invert(np.stack([mid, y], 1), params)[0][:, 0]
            big = ux > u[:, 0]
            hi, lo = np.where(big, mid, hi), np.where(big, lo, mid)
        return np.stack([0.5 * (lo + hi), y], 1)


def elevation(d):
    """Numerically robust polar angle (mitsuba/core/warp.h / vector.h: 2 asin(|d - z| / 2))."""
    return 2.0 * np.arcsin(np.clip(0.5 * np.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2 + (d[:, 2] - 1.0) ** 2), 0.0, 1.0))


def theta2u(theta):
    return np.sqrt(theta * (2.0 / np.pi))


def phi2u(phi):
    return (phi + np.pi) * (0.5 / np.pi)


class MeasuredBSDF:
    """Mitsuba 3 ``measured`` BSDF, rgb variant: ``eval(wi, wo)`` (= f cos(theta_o)) for wi, wo in the local frame."""

    def __init__(self, fields: dict):
        f = fields
        self.theta_i = np.asarray(f["theta_i"], np.float64)
        self.phi_i = np.asarray(f["phi_i"], np.float64)
        self.isotropic = self.phi_i.shape[0] <= 2
        self.jacobian = bool(np.asarray(f["jacobian"]).ravel()[0])
        self.reduction = 0 if self.isotropic else int(np.rint(2 * np.pi / (self.phi_i[-1] - self.phi_i[0])))
        self.ndf = Marginal2D(f["ndf"])
        self.sigma = Marginal2D(f["sigma"])
        self.vndf = Marginal2D(f["vndf"], (self.phi_i, self.theta_i), normalize=True)
        # (the `luminance` warp only drives Mitsuba's own sample(); eval does not use it)
        self.rgb = Marginal2D(f["rgb"], (self.phi_i, self.theta_i, np.arange(3.0)))
        self.fields = f

    @classmethod
    def from_file(cls, path: str) -> "MeasuredBSDF":
        return cls(read_tensor_file(path))

    def eval(self, wi, wo):
        wi, wo = np.asarray(wi, np.float64).copy(), np.asarray(wo, np.float64).copy()
        active = (wi[:, 2] > 0) & (wo[:, 2] > 0)
        wi[~active] = wo[~active] = (0.0, 0.0, 1.0)
        if self.reduction >= 2:                       # data covers half / a quarter of the azimuth range: fold
            sy = wi[:, 1].copy()
            sx = wi[:, 0].copy() if self.reduction == 4 else sy
            for v in (wi, wo):
                v[:, 0] = np.where(sx < 0, -v[:, 0], v[:, 0])
                v[:, 1] = np.where(sy < 0, -v[:, 1], v[:, 1])
        m = wi + wo
        m /= np.linalg.norm(m, axis=1, keepdims=True)
        theta_i, phi_i = elevation(wi), np.arctan2(wi[:, 1], wi[:, 0])
        theta_m, phi_m = elevation(m), np.arctan2(m[:, 1], m[:, 0])
        u_wi = np.stack([theta2u(theta_i), phi2u(phi_i)], 1)
        u_m = np.stack([theta2u(theta_m), phi2u(phi_m - phi_i if self.isotropic else phi_m)], 1)
        u_m[:, 1] -= np.floor(u_m[:, 1])
        params = (phi_i, theta_i)
        sample, _ = self.vndf.invert(u_m, params)
        spec = np.stack([self.rgb.eval(sample, params + (np.full_like(phi_i, float(c)),)) for c in range(3)], 1)
        if self.jacobian:
            spec = spec * (self.ndf.eval(u_m) / (4.0 * self.sigma.eval(u_wi)))[:, None]
        out = spec.copy()               # no explicit cos(theta_o) factor: see the module docstring ("cosine convention")
        out[~active] = 0.0
        return out


def rgb2lum(v):
    """rendering/utils/mitsuba_brdf_draw.py:36-38."""
    return 0.2126 * v[..., 0] + 0.7152 * v[..., 1] + 0.0722 * v[..., 2]


def disk_to_dir(xy):
    """utils/mitsuba_brdf_draw.py:40-43 (disk_to_cart)."""
    xy = np.asarray(xy, np.float64)
    z = np.sqrt(np.maximum(1.0 - (xy * xy).sum(1), 0.0))
    return np.concatenate([xy, z[:, None]], 1)


def target_density_disk(bsdf: MeasuredBSDF, wi_xy, wo_xy):
    """The (unnormalised) density the disk-domain nets were trained on, learning_repo_cleanup/utils/mitsuba_brdf_scalar.py
    :83-89: lum(eval) * clamp(1 / cos(theta_o), 1, 1e6) over the projected disk."""
    wi, wo = disk_to_dir(wi_xy), disk_to_dir(wo_xy)
    with np.errstate(divide="ignore"):
        inv_cos = np.clip(1.0 / wo[:, 2], 1.0, 1e6)
    inside = (np.asarray(wo_xy) ** 2).sum(1) < 1.0
    return np.where(inside, rgb2lum(bsdf.eval(wi, wo)) * inv_cos, 0.0)
