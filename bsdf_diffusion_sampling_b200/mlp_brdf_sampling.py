"""Drop-in replacements for the reference's sampler functions (same names, arguments, returns).

Reference: rendering/utils/mlp_brdf_sampling.py
    network_sampling_disk(D_base, D_sample, omega_i, T=4)          :17   -> (x [N,2], pdf [N])
    network_sampling_disk_tiny(x_alpha, D_sample, omega_i, T=4)    :54   -> (x [N,2], ones [N])
    network_pdf_disk(D_base, D_sample, omega_o, omega_i, T=4)      :69   -> pdf [N]
    network_sampling_spherical(D_base, D_sample, omega_i, T=8)     :106  -> (x=(theta,phi) [N,2], pdf [N])
    network_pdf_spherical(D_base, D_sample, omega_o, omega_i, T=8) :144  -> pdf [N]

``D_base`` / ``D_sample`` are ``nn.Module``s (this package's ``model`` classes or the reference's own --
only ``state_dict()`` is used); their weights are packed once and cached.  Each call is ONE fused
kernel launch on the current CUDA stream; outputs are new, detached fp32 tensors on ``omega_i``'s
device.  Extra keyword arguments (all optional) expose what the reference hard-codes:
``x0=`` replays an externally supplied base sample (parity testing), ``seed=/offset=`` pin the Philox
stream (default: torch's CUDA generator, so ``torch.manual_seed`` controls the sampler as it does the
reference), ``precision=`` selects "tc16" (tcgen05) or "fp32".
"""
from __future__ import annotations

import torch

from . import ops, weights


def _packed(D_base, D_sample, device):
    return weights.packed_flow_of(D_sample, device), weights.packed_base_of(D_base, device)


def _as_cuda(t: torch.Tensor) -> torch.Tensor:
    # the reference hard-codes device='cuda' (mlp_brdf_sampling.py:21-23,71)
    return t if t.is_cuda else t.to("cuda")


def network_sampling_disk(D_base, D_sample, omega_i, T=4, *, x0=None, seed=None, offset=0, precision=None):
    omega_i = _as_cuda(omega_i)
    flow, base = _packed(D_base, D_sample, omega_i.device)
    x, pdf, _ = ops.sample(omega_i, flow, base, T, x0=x0, seed=seed, offset=offset, precision=precision,
                           return_x0=False)
    return x, pdf


def network_pdf_disk(D_base, D_sample, omega_o, omega_i, T=4, *, precision=None):
    omega_i = _as_cuda(omega_i)
    flow, base = _packed(D_base, D_sample, omega_i.device)
    return ops.pdf(omega_o, omega_i, flow, base, T, precision=precision)


def network_sampling_spherical(D_base, D_sample, omega_i, T=8, *, x0=None, seed=None, offset=0, precision=None):
    omega_i = _as_cuda(omega_i)
    flow, base = _packed(D_base, D_sample, omega_i.device)
    x, pdf, _ = ops.sample(omega_i, flow, base, T, x0=x0, seed=seed, offset=offset, precision=precision,
                           return_x0=False)
    return x, pdf


def network_pdf_spherical(D_base, D_sample, omega_o, omega_i, T=8, *, precision=None):
    omega_i = _as_cuda(omega_i)
    flow, base = _packed(D_base, D_sample, omega_i.device)
    return ops.pdf(omega_o, omega_i, flow, base, T, precision=precision)


def network_sampling_disk_tiny(x_alpha, D_sample, omega_i, T=4, *, precision=None):
    """Forward-only Euler with a tcnn-style net fed ``cat[x, alpha, omega_i]``; returns pdf == 1
    (mlp_brdf_sampling.py:54-68; the reference evaluates the net twice per step and keeps the second
    result -- one evaluation gives the same numbers).  ``omega_i`` is whatever the caller concatenates
    after alpha, i.e. the already-encoded conditioning (22 columns for the shipped nets)."""
    x = _as_cuda(x_alpha).to(torch.float32)
    omega_i = _as_cuda(omega_i).to(torch.float32)
    n = x.shape[0]
    ones = torch.ones(n, 1, device=x.device)
    for t in range(T):
        d = D_sample(torch.cat([x, (t / T) * ones, omega_i], dim=1))
        x = x + 1 / T * d.to(torch.float32)
    return x, torch.ones(n, device=x.device)
