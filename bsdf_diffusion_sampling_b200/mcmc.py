"""Training-data generation on the GPU: an affine-invariant ensemble sampler for the (omega_i, omega_o) pairs
the three training stages consume, replacing the reference's emcee runs.

Reference (learning_repo_cleanup/utils/emcee_sampling.py:84-170, called from disk_domain_sampling.py:167-179): for each of
10 rings of omega_i, 49 emcee walkers x (10 000 burn-in + 40 000) steps of the stretch move on the 4-D density
    p(omega_i, omega_o)  ~  lum(f(omega_i, omega_o) cos theta_o) * clamp(1 / cos theta_o, 1, 1e6)     (disk coordinates)
                         ~  lum(f cos theta_o) * sin theta_o                                           ((theta, phi))
(utils/mitsuba_brdf_scalar.py:83-89), every density evaluation a Python call into Mitsuba through a multiprocessing pool
-- hours per material; the result is cached as brdf_samples_emcee<mat>.npy (``materials.load_emcee_cache``).

Here the SAME move (Goodman & Weare 2010 stretch move, a = 2, the ensemble split into two halves that update against
each other -- emcee's ``StretchMove`` / ``RedBlueMove``) runs for tens of thousands of walkers at once on the device: the
density is one launch of the CUDA measured-BSDF evaluator (``measured.MeasuredBSDF.eval``) per half-step over all
walkers of all rings, the proposal / accept logic a handful of elementwise torch kernels.  With 16 384 walkers per ring
the reference's 1.96 M samples per ring are 120 post-burn-in steps.

Quirk not reproduced: ``lnprob_brdf_disk`` rejects a state when ``mask_omegai == 0`` where ``mask_omegai`` is True OUTSIDE
the ring (emcee_sampling.py:13-17), i.e. as written it keeps omega_i outside its ring; the rings are meant to stratify
omega_i (the spherical variants at :21-29 constrain it to the ring) and that is what this module does.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Tuple

import torch


def stretch_move_ensemble(log_prob: Callable[..., torch.Tensor], p0: torch.Tensor, n_steps: int,
                          burn_in: int = 0, a: float = 2.0, generator: Optional[torch.Generator] = None,
                          groups: int = 1, thin: int = 1) -> Tuple[torch.Tensor, torch.Tensor, float]:
    """Affine-invariant ensemble sampling (emcee's default move) of ``exp(log_prob)`` from the walkers ``p0`` [W, d].
    ``log_prob(x [m,d], idx [m])`` -> [m]: ``idx`` are the walker numbers of the rows (per-walker constraints such as the
    omega_i ring of a walker's group key on it).

    ``groups`` > 1 runs that many INDEPENDENT ensembles in one call: walkers ``g * W/groups .. (g+1) * W/groups`` only ever
    pair with each other (one ensemble per omega_i ring).  -> (chain [n_kept * W, d] in step-major order like
    ``EnsembleSampler.get_chain(flat=True)``, final walkers [W, d], mean acceptance fraction)."""
    W, d = p0.shape
    if W % (2 * groups):
        raise ValueError("the walker count must be a multiple of 2 * groups")
    per, half = W // groups, W // groups // 2
    dev = p0.device
    x = p0.clone()
    lp = log_prob(x, torch.arange(W, device=dev))
    if not bool(torch.isfinite(lp).all()):
        raise ValueError("every initial walker needs a finite log-probability (emcee's requirement as well)")
    base = (torch.arange(groups, device=dev) * per).repeat_interleave(half)         # first walker of each half-row's group
    kept, accepted, proposed = [], 0.0, 0
    for step in range(burn_in + n_steps):
        # random split of every group into two halves
        order = torch.rand(groups, per, device=dev, generator=generator).argsort(1)
        first = (order[:, :half] + (torch.arange(groups, device=dev) * per)[:, None]).reshape(-1)
        second = (order[:, half:] + (torch.arange(groups, device=dev) * per)[:, None]).reshape(-1)
        for S, C in ((first, second), (second, first)):
            pick = torch.randint(0, half, (groups * half,), device=dev, generator=generator)
            partner = C.view(groups, half).gather(1, pick.view(groups, half)).reshape(-1)
            u = torch.rand(groups * half, device=dev, generator=generator)
            z = ((a - 1.0) * u + 1.0) ** 2 / a                                        # g(z) ~ 1/sqrt(z) on [1/a, a]
            xs, xc = x[S], x[partner]
            y = xc + z[:, None] * (xs - xc)
            lpy = log_prob(y, S)
            log_acc = (d - 1) * torch.log(z) + lpy - lp[S]
            acc = torch.log(torch.rand(groups * half, device=dev, generator=generator)) < log_acc
            acc &= torch.isfinite(lpy)
            x[S] = torch.where(acc[:, None], y, xs)
            lp[S] = torch.where(acc, lpy, lp[S])
            accepted += float(acc.float().sum())
            proposed += acc.numel()
        if step >= burn_in and (step - burn_in) % thin == 0:
            kept.append(x.clone())
    del base
    chain = torch.cat(kept, 0) if kept else x.new_zeros((0, d))
    return chain, x, accepted / max(proposed, 1)


def _disk_to_dir(xy: torch.Tensor) -> torch.Tensor:
    z = torch.sqrt(torch.clamp(1.0 - (xy * xy).sum(1), min=0.0))
    return torch.cat([xy, z[:, None]], 1)


def _sph_to_dir(tp: torch.Tensor) -> torch.Tensor:
    st = torch.sin(tp[:, 0])
    return torch.stack([st * torch.cos(tp[:, 1]), st * torch.sin(tp[:, 1]), torch.cos(tp[:, 0])], 1)


def brdf_target_density(bsdf, p: torch.Tensor, domain: str = "disk") -> torch.Tensor:
    """The density the nets are trained on, unnormalised (utils/mitsuba_brdf_scalar.py:83-89): ``p`` [n,4] =
    (omega_i, omega_o) in disk coordinates or (theta, phi); 0 outside the domain."""
    wi, wo = p[:, 0:2].contiguous(), p[:, 2:4].contiguous()
    if domain == "disk":
        ok = ((wi * wi).sum(1) < 1.0) & ((wo * wo).sum(1) < 1.0)
        wi3, wo3 = _disk_to_dir(wi), _disk_to_dir(wo)
        jac = torch.clamp(1.0 / torch.clamp(wo3[:, 2], min=1e-12), 1.0, 1e6)
    elif domain == "spherical":
        pi = math.pi
        ok = (wi[:, 0] > 0) & (wi[:, 0] < pi / 2) & (wo[:, 0] > 0) & (wo[:, 0] < pi / 2) & (wi[:, 1].abs() < pi) & (wo[:, 1].abs() < pi)
        wi3, wo3 = _sph_to_dir(wi), _sph_to_dir(wo)
        jac = torch.sqrt(torch.clamp(1.0 - wo3[:, 2] ** 2, min=0.0))
    else:
        raise ValueError("domain must be 'disk' or 'spherical'")
    v = bsdf.eval(wi3.float().contiguous(), wo3.float().contiguous())
    lum = 0.2126 * v[:, 0] + 0.7152 * v[:, 1] + 0.0722 * v[:, 2]
    return torch.where(ok, torch.nan_to_num(lum * jac, nan=0.0, posinf=0.0, neginf=0.0), torch.zeros_like(lum))


def sample_brdf_pairs(bsdf, n_samples: int, domain: str = "disk", piecewise: int = 10, walkers_per_ring: int = 16384,
                      burn_in: int = 300, generator: Optional[torch.Generator] = None, thin: int = 1) -> torch.Tensor:
    """GPU replacement of ``emcee_mcmc_brdf_disk`` / ``emcee_mcmc_brdf_spherical``: about ``n_samples`` (omega_i, omega_o)
    pairs [N,4], one tenth per omega_i ring (radius rings of the unit disk / theta_i bands of [0, pi/2]), ring-major like
    the reference's concatenation.  ``bsdf`` = ``measured.MeasuredBSDF`` (or anything with ``eval(wi3, wo3) -> [n,3]``)."""
    dev = bsdf.blob.device if hasattr(bsdf, "blob") else torch.device("cuda")
    G, per = int(piecewise), int(walkers_per_ring)
    W = G * per
    ring = torch.arange(G, device=dev).repeat_interleave(per)
    lo, hi = ring.float() / G, (ring.float() + 1.0) / G

    def rnd(*shape):
        return torch.rand(*shape, device=dev, generator=generator)

    def in_ring(wi, idx):
        r = torch.sqrt((wi * wi).sum(1)) if domain == "disk" else wi[:, 0] / (math.pi / 2)
        return (r > lo[idx]) & (r < hi[idx])

    all_idx = torch.arange(W, device=dev)

    def log_prob(p, idx):
        d = brdf_target_density(bsdf, p, domain)
        ok = in_ring(p[:, 0:2], idx) & (d > 0)
        return torch.where(ok, torch.log(torch.clamp(d, min=1e-38)), torch.full_like(d, -float("inf")))

    # initial walkers: omega_i uniform in its ring, omega_o redrawn until the density is non-zero (find_omegao, :46-58)
    def draw_i():
        if domain == "disk":
            r = torch.sqrt(lo * lo + rnd(W) * (hi * hi - lo * lo)) * 0.999 + 1e-4 * (hi - lo)
            a = rnd(W) * 2 * math.pi
            return torch.stack([r * torch.cos(a), r * torch.sin(a)], 1)
        return torch.stack([(lo + (0.001 + 0.998 * rnd(W)) * (hi - lo)) * (math.pi / 2), (rnd(W) * 2 - 1) * math.pi * 0.999], 1)

    def draw_o():
        if domain == "disk":
            r, a = torch.sqrt(rnd(W)) * 0.999, rnd(W) * 2 * math.pi
            return torch.stack([r * torch.cos(a), r * torch.sin(a)], 1)
        return torch.stack([(0.001 + 0.998 * rnd(W)) * (math.pi / 2), (rnd(W) * 2 - 1) * math.pi * 0.999], 1)

    p0 = torch.cat([draw_i(), draw_o()], 1)
    for _ in range(64):
        bad = ~torch.isfinite(log_prob(p0, all_idx))
        if not bool(bad.any()):
            break
        p0 = torch.where(bad[:, None], torch.cat([draw_i(), draw_o()], 1), p0)
    else:
        raise RuntimeError("could not find a walker start with non-zero density in every ring")
    steps = max(1, -(-int(n_samples) // W)) * thin
    chain, _, acc = stretch_move_ensemble(log_prob, p0, steps, burn_in=burn_in, generator=generator, groups=G, thin=thin)
    # step-major [steps, G, per, 4] -> ring-major like np.concatenate(all_samples)
    out = chain.view(-1, G, per, 4).permute(1, 0, 2, 3).reshape(-1, 4).contiguous()
    sample_brdf_pairs.last_acceptance = acc
    return out


__all__ = ["stretch_move_ensemble", "brdf_target_density", "sample_brdf_pairs"]
