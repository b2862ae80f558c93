"""CPU oracle for the neural BSDF sampler hot path (numpy restatement).

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker / the CPU baseline.

This file restates, in closed form (forward-mode tangents instead of the reference's
two autograd sweeps), the arithmetic of the reference functions below.  Citations are
``file:line`` relative to the reference checkout (fzy28/BSDF_diffusion_sampling @ f2b615c):

* ``positional_encoding``        <- rendering/utils/model.py:9-57   (positional_encoding_1)
* ``base_forward``               <- rendering/utils/model.py:382-386 / 289-292 (forward of the *_one nets)
* ``base_logprob_disk``          <- rendering/utils/model.py:393-398
* ``base_params_spherical``      <- rendering/utils/model.py:293-298
* ``base_logprob_spherical``     <- rendering/utils/model.py:308-317 (+ torch/distributions/von_mises.py
                                    _log_modified_bessel_fn / VonMises.log_prob, torch 2.11.0)
* ``flow_velocity``              <- rendering/utils/model.py:490-501 (NN_cond_pos_simpler, 2nd def),
                                    :435-446 (NN_cond_pos), :465-477 (NN_cond_pos_spherical_complicate)
* ``sample_disk``                <- rendering/utils/mlp_brdf_sampling.py:17-51
* ``pdf_disk``                   <- rendering/utils/mlp_brdf_sampling.py:69-103
* ``sample_spherical``           <- rendering/utils/mlp_brdf_sampling.py:106-140
* ``pdf_spherical``              <- rendering/utils/mlp_brdf_sampling.py:144-181
* ``plugin_*``                   <- rendering/brdf_measured_disk.py:59-124,
                                    rendering/brdf_measured_spherical.py:31-39,69-137,
                                    rendering/bsdf_myresult.py:59-133,
                                    rendering/utils/mitsuba_brdf_draw.py:40-43
* ``reflow_forward``             <- learning_repo_cleanup/disk_domain_sampling.py:93-110,
                                    learning_repo_cleanup/spherical_domain_sampling.py:147-166
                                    (fp16 MLP emulation of tiny-cuda-nn's FullyFusedMLP:
                                    tiny-cuda-nn/src/fully_fused_mlp.cu:47-129, fp16 accumulate;
                                    SiLU in fp32 on the fp16 pre-activation,
                                    tiny-cuda-nn/include/tiny-cuda-nn/common_device.h:139-145)

Parity pin: the reference ships no tests or golden vectors for this path.  The oracle is
pinned against outputs of the reference's own Python functions run in the build container
(``tests/golden/make_golden.py`` imports ``/root/reference`` and writes ``tests/golden/*.npz``);
``tests/test_oracle_golden.py`` checks every function here against those files.
The Mitsuba plugin epilogues (``plugin_*``) cannot be run (Mitsuba/Dr.Jit absent): for
those, parity is UNPINNED and rests on the line-by-line restatement cited above.

Arithmetic is carried in float32 by default (the reference runs fp32 eager); pass
``dtype=np.float64`` for an error-analysis "truth".
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

DISK = 0
SPHERICAL = 1

# torch/distributions/von_mises.py (torch 2.11.0): Abramowitz & Stegun 9.8.1 / 9.8.2
_I0_COEF_SMALL = [1.0, 3.5156229, 3.0899424, 1.2067492, 0.2659732, 0.360768e-1, 0.45813e-2]
_I0_COEF_LARGE = [0.39894228, 0.1328592e-1, 0.225319e-2, -0.157565e-2, 0.916281e-2,
                  -0.2057706e-1, 0.2635537e-1, -0.1647633e-1, 0.392377e-2]


@dataclass
class FlowWeights:
    """Bias-free flow (velocity) MLP: ``layers`` = [W1 [H,in], W2 [H,H], ..., Wout [2,H]]."""
    layers: List[np.ndarray]

    @property
    def in_dim(self) -> int:
        return int(self.layers[0].shape[1])

    @property
    def domain(self) -> int:
        return DISK if self.in_dim == 25 else SPHERICAL

    def astype(self, dt) -> "FlowWeights":
        return FlowWeights([np.asarray(w, dtype=dt) for w in self.layers])


@dataclass
class BaseWeights:
    """Base-distribution net 14 -> 16 -> 4 with biases (model.py:374-381 / 277-288)."""
    w1: np.ndarray  # [16,14]
    b1: np.ndarray  # [16]
    wo: np.ndarray  # [4,16]
    bo: np.ndarray  # [4]

    def astype(self, dt) -> "BaseWeights":
        return BaseWeights(*(np.asarray(a, dtype=dt) for a in (self.w1, self.b1, self.wo, self.bo)))

    def flat(self) -> np.ndarray:
        """308-float blob in the order the C-ABI expects: w1 row-major, b1, wo row-major, bo."""
        return np.concatenate([self.w1.ravel(), self.b1.ravel(), self.wo.ravel(), self.bo.ravel()]
                              ).astype(np.float32)


# ------------------------------------------------------------------------------------------------
# building blocks
# ------------------------------------------------------------------------------------------------
def silu(z):
    with np.errstate(over="ignore"):
        return z / (1.0 + np.exp(-z))


def silu_and_grad(z):
    """silu(z) and silu'(z) = s (1 + z (1 - s)),  s = sigmoid(z)."""
    with np.errstate(over="ignore"):
        s = 1.0 / (1.0 + np.exp(-z))
    return z * s, s * (1.0 + z * (1.0 - s))


def positional_encoding(v: np.ndarray, L: int) -> np.ndarray:
    """[v, sin(2^0 v), cos(2^0 v), ..., sin(2^(L-1) v), cos(2^(L-1) v)]  (model.py:26,48-57)."""
    out = [v]
    for k in range(L):
        f = v.dtype.type(2.0 ** k)
        out.append(np.sin(v * f))
        out.append(np.cos(v * f))
    return np.concatenate(out, axis=-1)


def base_forward(base: BaseWeights, wi: np.ndarray) -> np.ndarray:
    """p = Wo silu(W1 PE_3(wi) + b1) + bo   (model.py:382-386)."""
    pe = positional_encoding(wi, 3)
    h = silu(pe @ base.w1.T + base.b1)
    return h @ base.wo.T + base.bo


def base_logprob_disk(base: BaseWeights, x: np.ndarray, wi: np.ndarray) -> np.ndarray:
    """model.py:393-398."""
    p = base_forward(base, wi)
    loc, ls = p[:, :2], p[:, 2:]
    eps = (x - loc) / np.exp(ls)
    dt = x.dtype.type
    return dt(-0.5 * 2 * math.log(2 * math.pi)) - ls.sum(1) - dt(0.5) * (eps ** 2).sum(1)


def softplus(x):
    """torch.nn.Softplus(beta=1, threshold=20)."""
    return np.where(x > 20.0, x, np.log1p(np.exp(np.minimum(x, 20.0))))


def base_params_spherical(base: BaseWeights, wi: np.ndarray):
    """loc, log_scale, loc_von, concentration  (model.py:293-298)."""
    p = base_forward(base, wi)
    dt = p.dtype.type
    return p[:, 0], p[:, 1], p[:, 2], (softplus(p[:, 3]) + dt(1e-3)).astype(p.dtype)


def log_i0(x: np.ndarray) -> np.ndarray:
    """torch _log_modified_bessel_fn(order=0): polynomial switch at 3.75."""
    dt = x.dtype.type

    def poly(y, coef):
        coef = list(coef)
        r = dt(coef.pop())
        while coef:
            r = dt(coef.pop()) + y * r
        return r

    y = x / dt(3.75)
    y = y * y
    small = np.log(poly(y, _I0_COEF_SMALL))
    with np.errstate(divide="ignore", invalid="ignore"):
        y = dt(3.75) / x
        large = x - dt(0.5) * np.log(x) + np.log(poly(y, _I0_COEF_LARGE))
    return np.where(x < dt(3.75), small, large)


def base_logprob_spherical(base: BaseWeights, x: np.ndarray, wi: np.ndarray) -> np.ndarray:
    """model.py:308-317.  NB the Gaussian normaliser uses log_scale although the scale is
    exp(log_scale)+1e-3 -- reference behaviour, preserved."""
    loc, ls, mu, kappa = base_params_spherical(base, wi)
    dt = x.dtype.type
    eps = (x[:, 0] - loc) / (np.exp(ls) + dt(1e-3))
    loggau = dt(-0.5 * math.log(2 * math.pi)) - ls - dt(0.5) * eps ** 2
    logvon = kappa * np.cos(x[:, 1] - mu) - dt(math.log(2 * math.pi)) - log_i0(kappa)
    return loggau + logvon


def flow_velocity(flow: FlowWeights, x: np.ndarray, alpha: float, pe_wi: np.ndarray):
    """Velocity d(x, alpha | wi) and its 2x2 Jacobian w.r.t. the 2-D state.

    Disk:       net input [x0, x1, alpha, PE_5(wi)]                       (model.py:494-495)
    Spherical:  net input [theta, sin phi, cos phi, alpha, PE_5(wi)]      (mlp_brdf_sampling.py:119-121)
    Returns d [N,2], du = dd/dx0 [N,2], dv = dd/dx1 [N,2].
    """
    W = flow.layers
    n = x.shape[0]
    dt = x.dtype
    a = np.full((n, 1), alpha, dtype=dt)
    if flow.domain == DISK:
        inp = np.concatenate([x, a, pe_wi], axis=1)
        t_u = np.broadcast_to(W[0][:, 0], (n, W[0].shape[0]))
        t_v = np.broadcast_to(W[0][:, 1], (n, W[0].shape[0]))
    else:
        s, c = np.sin(x[:, 1:2]), np.cos(x[:, 1:2])
        inp = np.concatenate([x[:, 0:1], s, c, a, pe_wi], axis=1)
        t_u = np.broadcast_to(W[0][:, 0], (n, W[0].shape[0]))
        t_v = c * W[0][:, 1][None, :] - s * W[0][:, 2][None, :]
    z = inp @ W[0].T
    h, g = silu_and_grad(z)
    u, v = g * t_u, g * t_v
    for Wk in W[1:-1]:
        z = h @ Wk.T
        h, g = silu_and_grad(z)
        u = g * (u @ Wk.T)
        v = g * (v @ Wk.T)
    Wo = W[-1]
    return h @ Wo.T, u @ Wo.T, v @ Wo.T


def _euler(flow: FlowWeights, x: np.ndarray, wi: np.ndarray, T: int, reverse: bool):
    """T explicit Euler steps with the running Jacobian-determinant product.

    forward  (mlp_brdf_sampling.py:26-47):  alpha=t/T,   x += d/T, J = I + dd/dx / T, R /= det J
    reverse  (mlp_brdf_sampling.py:77-99):  alpha=1-t/T, x -= d/T, J = I - dd/dx / T, R *= det J
    """
    dt = x.dtype.type
    pe = positional_encoding(wi, 5)
    R = np.ones(x.shape[0], dtype=x.dtype)
    inv_t = dt(1.0 / T)
    for t in range(T):
        alpha = dt((1 - t / T) if reverse else (t / T))
        d, du, dv = flow_velocity(flow, x, alpha, pe)
        sgn = dt(-1.0) if reverse else dt(1.0)
        # J_1 = e1 + s/T * grad(d[0]),  J_2 = e2 + s/T * grad(d[1]);  det = J1[0] J2[1] - J1[1] J2[0]
        j00 = dt(1.0) + sgn * inv_t * du[:, 0]
        j01 = sgn * inv_t * dv[:, 0]
        j10 = sgn * inv_t * du[:, 1]
        j11 = dt(1.0) + sgn * inv_t * dv[:, 1]
        det = j00 * j11 - j01 * j10
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            R = R * det if reverse else R / det
        x = x + sgn * inv_t * d
    return x, R


# ------------------------------------------------------------------------------------------------
# base samplers (used when no x0 is replayed; statistically, not bitwise, equal to torch's)
# ------------------------------------------------------------------------------------------------
def draw_x0_disk(base: BaseWeights, wi: np.ndarray, rng: np.random.Generator) -> np.ndarray:
    p = base_forward(base, wi)
    eps = rng.standard_normal(p[:, :2].shape).astype(wi.dtype)
    return p[:, :2] + eps * np.exp(p[:, 2:])


def draw_x0_spherical(base: BaseWeights, wi: np.ndarray, rng: np.random.Generator) -> np.ndarray:
    """model.py:299-307: theta0 ~ N(loc, exp(ls)+1e-3), phi0 ~ vonMises(mu, kappa) wrapped to [-pi,pi)."""
    loc, ls, mu, kappa = base_params_spherical(base, wi)
    dt = wi.dtype.type
    th = loc + rng.standard_normal(loc.shape).astype(wi.dtype) * (np.exp(ls) + dt(1e-3))
    ph = rng.vonmises(mu.astype(np.float64), kappa.astype(np.float64))
    ph = (ph + math.pi) % (2 * math.pi) - math.pi
    return np.stack([th, ph.astype(wi.dtype)], axis=1)


# ------------------------------------------------------------------------------------------------
# the four sampler entry points (raw domain coordinates, as mlp_brdf_sampling.py returns them)
# ------------------------------------------------------------------------------------------------
def _prep(flow, base, dtype, *arrs):
    return (flow.astype(dtype), base.astype(dtype)) + tuple(np.ascontiguousarray(a, dtype=dtype) for a in arrs)


def sample_disk(flow, base, wi, T=4, x0=None, rng=None, dtype=np.float32):
    """-> (x [N,2], pdf [N], x0 [N,2]);  mlp_brdf_sampling.py:17-51."""
    flow, base, wi = _prep(flow, base, dtype, wi)
    if x0 is None:
        x0 = draw_x0_disk(base, wi, rng or np.random.default_rng(0))
    x0 = np.asarray(x0, dtype=dtype)
    with np.errstate(over="ignore"):
        p0 = np.exp(base_logprob_disk(base, x0, wi))
    x, R = _euler(flow, x0, wi, T, reverse=False)
    return x, p0 * R, x0


def pdf_disk(flow, base, wo, wi, T=4, dtype=np.float32):
    """-> pdf [N];  mlp_brdf_sampling.py:69-103."""
    flow, base, wo, wi = _prep(flow, base, dtype, wo, wi)
    x, R = _euler(flow, wo, wi, T, reverse=True)
    with np.errstate(over="ignore"):
        return np.exp(base_logprob_disk(base, x, wi)) * R


def sample_spherical(flow, base, wi, T=8, x0=None, rng=None, dtype=np.float32):
    """-> (x=(theta,phi) [N,2], pdf [N], x0);  mlp_brdf_sampling.py:106-140."""
    flow, base, wi = _prep(flow, base, dtype, wi)
    if x0 is None:
        x0 = draw_x0_spherical(base, wi, rng or np.random.default_rng(0))
    x0 = np.asarray(x0, dtype=dtype)
    with np.errstate(over="ignore"):
        p0 = np.exp(base_logprob_spherical(base, x0, wi))
    x, R = _euler(flow, x0, wi, T, reverse=False)
    return x, p0 * R, x0


def pdf_spherical(flow, base, wo, wi, T=8, dtype=np.float32):
    """-> pdf [N];  mlp_brdf_sampling.py:144-181."""
    flow, base, wo, wi = _prep(flow, base, dtype, wo, wi)
    x, R = _euler(flow, wo, wi, T, reverse=True)
    with np.errstate(over="ignore"):
        return np.exp(base_logprob_spherical(base, x, wi)) * R


# ------------------------------------------------------------------------------------------------
# plugin-level pre/post-processing (tensor part of MyBSDF.sample / MyBSDF.pdf).  PARITY UNPINNED
# (Mitsuba is not runnable here); restated line by line.
# ------------------------------------------------------------------------------------------------
PLUGIN_DISK = 0         # rendering/brdf_measured_disk.py
PLUGIN_SPHERICAL = 1    # rendering/brdf_measured_spherical.py
PLUGIN_BSDF = 2         # rendering/bsdf_myresult.py

FLT_MAX = np.float32(np.finfo(np.float32).max)


def cart_to_spher(w3: np.ndarray) -> np.ndarray:
    """brdf_measured_spherical.py:35-39."""
    dt = w3.dtype.type
    r = np.sqrt((w3 * w3).sum(1))
    theta = np.arccos(w3[:, 2] / (r + dt(1e-8)))
    phi = np.arctan2(w3[:, 1], w3[:, 0])
    return np.stack([theta, phi], axis=1)


def _inv_sin_clamped(sin_t, use_abs):
    """dr.clamp(1/sin_theta, 1, FLT_MAX)  (brdf_measured_spherical.py:89-91; bsdf_myresult.py:81 adds abs)."""
    s = np.abs(sin_t) if use_abs else sin_t
    with np.errstate(divide="ignore"):
        inv = np.float32(1.0) / s
    return np.minimum(np.maximum(inv, np.float32(1.0)), FLT_MAX)


def _frame_sin_theta(wo3):
    """Mitsuba Frame3f.sin_theta: safe_sqrt(x^2 + y^2) of the direction."""
    return np.sqrt(np.maximum(wo3[:, 0] ** 2 + wo3[:, 1] ** 2, 0)).astype(wo3.dtype)


def plugin_sample(kind, flow, base, wi3, T=None, x0=None, rng=None, dtype=np.float32):
    """Tensor part of MyBSDF.sample *before* the ground-truth-eval firefly clamp.

    -> (wo3 [N,3], pdf_omega [N], x0).  pdf_omega is ``bs.pdf`` as first assigned
    (brdf_measured_disk.py:82, brdf_measured_spherical.py:92, bsdf_myresult.py:84).
    """
    wi3 = np.ascontiguousarray(wi3, dtype=dtype)
    if kind == PLUGIN_DISK:
        x, pdf, x0 = sample_disk(flow, base, wi3[:, :2], T or 4, x0, rng, dtype)
        valid = (x[:, 0] ** 2 + x[:, 1] ** 2) < dtype(0.995)            # :69
        x = np.where(valid[:, None], x, 0)                               # :70
        pdf = np.where(valid, pdf, 0)                                    # :71
        rr = (x ** 2).sum(1)
        z = np.sqrt(np.maximum(1 - rr, 0)).astype(x.dtype)               # disk_to_cart
        wo3 = np.concatenate([x, z[:, None]], axis=1)
        return wo3, pdf * z, x0                                          # :82
    wis = cart_to_spher(wi3)
    x, pdf, x0 = sample_spherical(flow, base, wis, T or 8, x0, rng, dtype)
    pdf = np.where(np.sin(x[:, 0]) > dtype(0.00005), pdf, 0)             # :79 / bsdf :69
    if kind == PLUGIN_SPHERICAL:
        pdf = np.where(np.cos(x[:, 0]) > 0, pdf, 0)                      # :80
    st, ct = np.sin(x[:, 0]), np.cos(x[:, 0])
    sp, cp = np.sin(x[:, 1]), np.cos(x[:, 1])
    wo3 = np.stack([cp * st, sp * st, ct], axis=1)                       # sph_to_dir :31-34
    inv = _inv_sin_clamped(_frame_sin_theta(wo3), use_abs=(kind == PLUGIN_BSDF))
    with np.errstate(over="ignore", invalid="ignore"):
        return wo3, (pdf * inv).astype(x.dtype), x0


def plugin_pdf(kind, flow, base, wi3, wo3, T=None, dtype=np.float32):
    """Tensor part of MyBSDF.pdf.  -> pdf_omega [N]."""
    wi3 = np.ascontiguousarray(wi3, dtype=dtype)
    wo3 = np.ascontiguousarray(wo3, dtype=dtype)
    if kind == PLUGIN_DISK:                                              # brdf_measured_disk.py:112-124
        pdf = pdf_disk(flow, base, wo3[:, :2], wi3[:, :2], T or 4, dtype)
        ok = (wi3[:, 2] > 0) & (wo3[:, 2] > 0)
        return np.where(ok, pdf * wo3[:, 2], 0).astype(dtype)
    wis, wos = cart_to_spher(wi3), cart_to_spher(wo3)
    pdf = pdf_spherical(flow, base, wos, wis, T or 8, dtype)
    inv = _inv_sin_clamped(_frame_sin_theta(wo3), use_abs=(kind == PLUGIN_BSDF))
    with np.errstate(over="ignore", invalid="ignore"):
        if kind == PLUGIN_SPHERICAL:                                     # brdf_measured_spherical.py:122-137
            pdf = np.where(np.sin(wos[:, 0]) > dtype(0.00005), pdf, 0)
            ok = (wi3[:, 2] > 0) & (wo3[:, 2] > 0)
            return np.where(ok, pdf * inv, 0).astype(dtype)
        return (pdf * inv).astype(dtype)                                 # bsdf_myresult.py:115-130


# ------------------------------------------------------------------------------------------------
# reflow forward-only sampling (dosampling) with tiny-cuda-nn fp16 semantics
# ------------------------------------------------------------------------------------------------
def tcnn_mlp_forward_fp16(layers: Sequence[np.ndarray], inp: np.ndarray, fp16_accumulate=True) -> np.ndarray:
    """FullyFusedMLP inference: fp16 operands; accumulators fp16 (reference kernel) or fp32;
    SiLU evaluated in fp32 on the stored fp16 pre-activation, rounded to fp16."""
    h = inp.astype(np.float16)
    acc = np.float16 if fp16_accumulate else np.float32
    for k, W in enumerate(layers):
        z = (h.astype(np.float32) @ W.astype(np.float16).astype(np.float32).T).astype(acc)
        if k == len(layers) - 1:
            return z.astype(np.float16)
        zf = z.astype(np.float16).astype(np.float32)
        h = silu(zf).astype(np.float16)
    raise AssertionError


def reflow_forward(flow: FlowWeights, x0: np.ndarray, wi: np.ndarray, T: int, mode="fp32"):
    """x_T of disk_domain_sampling.py:100-108 / spherical_domain_sampling.py:154-164.

    mode "fp32": the PyTorch module in fp32 (what tcnn approximates; tcnn's own agreement bar is
    rtol=atol=1e-2, tiny-cuda-nn/tmp.py:59).  mode "fp16": fp16 operands + fp32 accumulate;
    mode "tcnn": fp16 accumulate like the reference's kernel.
    """
    x = np.asarray(x0, dtype=np.float32).copy()
    wi = np.asarray(wi, dtype=np.float32)
    pe = positional_encoding(wi, 5)
    W = [np.asarray(w, np.float32) for w in flow.layers]
    n = x.shape[0]
    inv_t = np.float32(1.0 / T)
    for t in range(T):
        a = np.full((n, 1), np.float32(t / T), np.float32)
        if flow.domain == DISK:
            inp = np.concatenate([x, a, pe], axis=1)
        else:
            inp = np.concatenate([x[:, 0:1], np.sin(x[:, 1:2]), np.cos(x[:, 1:2]), a, pe], axis=1)
        if mode == "fp32":
            h = inp
            for Wk in W[:-1]:
                h = silu(h @ Wk.T)
            d = h @ W[-1].T
        else:
            d = tcnn_mlp_forward_fp16(W, inp, fp16_accumulate=(mode == "tcnn")).astype(np.float32)
        x = x + inv_t * d
    return x


def pack_tcnn_params(layers: Sequence[np.ndarray], input_dims: int, output_dims: int) -> np.ndarray:
    """Flat fp16 parameter vector of load_pytorch_model_to_tinycuda
    (learning_repo_cleanup/utils/utils.py:13-23): first layer padded by 16-(in%16) zero columns,
    last layer by 16-(out%16) zero rows, all row-major, concatenated."""
    parts = []
    w0 = np.asarray(layers[0], np.float32)
    parts.append(np.pad(w0, ((0, 0), (0, 16 - input_dims % 16))).ravel())
    for w in layers[1:-1]:
        parts.append(np.asarray(w, np.float32).ravel())
    wl = np.asarray(layers[-1], np.float32)
    parts.append(np.pad(wl, ((0, 16 - output_dims % 16), (0, 0))).ravel())
    return np.concatenate(parts).astype(np.float16)


# ------------------------------------------------------------------------------------------------
# fixtures
# ------------------------------------------------------------------------------------------------
def load_material_npz(path) -> Tuple[FlowWeights, BaseWeights, dict]:
    """Load a ``tests/golden/<material>.npz`` fixture (weights + reference outputs)."""
    z = np.load(path)
    n_layers = int(z["n_flow_layers"])
    flow = FlowWeights([z[f"flow_w{i}"] for i in range(n_layers)])
    base = BaseWeights(z["base_w1"], z["base_b1"], z["base_wo"], z["base_bo"])
    return flow, base, {k: z[k] for k in z.files}


def stratified_wi_disk(n_side: int, rng: Optional[np.random.Generator] = None, scale=0.95) -> np.ndarray:
    """Synthetic wi generator: jittered grid mapped to the unit disk by the concentric map
    (the reference's own wi generator, learning_repo_cleanup/utils/utils_sampling_torch_disk.py:79-114),
    scaled to r < ``scale``.  -> [n_side^2, 2] float32."""
    rng = rng or np.random.default_rng(0)
    i, j = np.meshgrid(np.arange(n_side), np.arange(n_side), indexing="ij")
    u = (i.ravel() + rng.random(n_side * n_side)) / n_side
    v = (j.ravel() + rng.random(n_side * n_side)) / n_side
    a, b = 2 * u - 1, 2 * v - 1
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.where(np.abs(a) > np.abs(b), a, b)
        phi = np.where(np.abs(a) > np.abs(b), (math.pi / 4) * (b / a), math.pi / 2 - (math.pi / 4) * (a / b))
    phi = np.where((a == 0) & (b == 0), 0.0, phi)
    return (scale * np.stack([r * np.cos(phi), r * np.sin(phi)], axis=1)).astype(np.float32)
