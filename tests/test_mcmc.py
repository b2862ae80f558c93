"""GPU replacement of the reference's emcee runs: the stretch-move ensemble on an analytic target (runs on
CPU: it is plain torch), and -- on the GPU -- the (omega_i, omega_o) sampler against the measured BSDF's own density."""
import math
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR

SILK = os.path.join(GOLDEN_DIR, "measured", "measured_vch_silk_blue_rgb.npz")


def test_stretch_move_samples_a_correlated_gaussian_in_a_box(built_lib):
    """emcee's move on a known target: N(mu, Sigma) restricted to x_0 > -1 (a hard wall like the domain masks)."""
    mc = built_lib.mcmc
    g = torch.Generator().manual_seed(0)
    mu = torch.tensor([0.5, -1.0, 2.0, 0.0])
    A = torch.tensor([[1.0, 0, 0, 0], [0.8, 0.6, 0, 0], [0, 0, 0.3, 0], [0.2, 0, 0, 1.5]])
    P = torch.linalg.inv(A @ A.T)

    def log_prob(x, idx):
        d = x - mu
        lp = -0.5 * ((d @ P) * d).sum(1)
        return torch.where(x[:, 0] > -1.0, lp, torch.full_like(lp, -float("inf")))

    p0 = mu + 0.1 * torch.randn(512, 4, generator=g)
    chain, last, acc = mc.stretch_move_ensemble(log_prob, p0, 600, burn_in=300, generator=g, groups=2)
    assert chain.shape == (600 * 512, 4) and 0.2 < acc < 0.9
    assert bool((chain[:, 0] > -1.0).all())
    # reference moments of the truncated Gaussian by brute force
    ref = mu + torch.randn(2_000_000, 4, generator=g) @ A.T
    ref = ref[ref[:, 0] > -1.0]
    # (autocorrelated chain: ~1e4 effective draws of a target with standard deviations up to 1.5)
    assert (chain.mean(0) - ref.mean(0)).abs().max() < 0.08
    assert (torch.cov(chain.T) - torch.cov(ref.T)).abs().max() < 0.15
    with pytest.raises(ValueError, match="multiple of 2"):
        mc.stretch_move_ensemble(log_prob, p0[:511], 1)
    with pytest.raises(ValueError, match="finite log-probability"):
        mc.stretch_move_ensemble(log_prob, p0 - 10.0, 1)


@pytest.mark.gpu
@pytest.mark.parametrize("domain", ["disk", "spherical"])
def test_brdf_pairs_follow_the_measured_density(built_lib, domain):
    """Samples of one omega_i ring: (a) omega_i stays in its ring, everything inside the domain; (b) binned over
    (|omega_i| band, omega_o cell) the sample counts match the density integrated over the same cells (midpoint rule on
    the GPU evaluator): Pearson chi-square / dof close to 1 after thinning the chain to near-independent draws."""
    pkg = built_lib
    z = np.load(SILK)
    m = pkg.measured.MeasuredBSDF({k: z[k] for k in ("theta_i", "phi_i", "ndf", "sigma", "vndf", "rgb", "jacobian")})
    g = torch.Generator(device="cuda").manual_seed(3)
    G, per = 4, 8192
    s = pkg.mcmc.sample_brdf_pairs(m, G * per * 12, domain=domain, piecewise=G, walkers_per_ring=per, burn_in=400,
                                   generator=g, thin=12)
    assert s.shape == (G * per * 12, 4) and bool(torch.isfinite(s).all())
    assert 0.15 < pkg.mcmc.sample_brdf_pairs.last_acceptance < 0.9
    ring = torch.arange(G, device="cuda").repeat_interleave(per * 12)
    r = s[:, :2].norm(dim=1) if domain == "disk" else s[:, 0] / (math.pi / 2)
    assert bool(((r > ring / G) & (r < (ring + 1) / G)).all())
    d = pkg.mcmc.brdf_target_density(m, s, domain)
    assert bool((d > 0).all())
    # isotropic material: the density depends on (r_i or theta_i, omega_o relative to omega_i's azimuth); test ring 1
    sel = s[ring == 1]
    if domain == "disk":
        phi_i = torch.atan2(sel[:, 1], sel[:, 0])
        c, sn = torch.cos(-phi_i), torch.sin(-phi_i)
        ox, oy = c * sel[:, 2] - sn * sel[:, 3], sn * sel[:, 2] + c * sel[:, 3]       # omega_o rotated so that omega_i = (r, 0)
        ri = sel[:, :2].norm(dim=1)
        nb = 8
        H = torch.histogramdd(torch.stack([ox, oy], 1).cpu().double(), bins=[nb, nb], range=[-1.0, 1.0, -1.0, 1.0]).hist
        # expected: integrate the density over r_i in the ring (weight r_i dr_i: area element) and the omega_o cell
        sub, nr = 8, 24
        cc = (torch.arange(nb * sub, device="cuda") + 0.5) / (nb * sub) * 2 - 1
        X, Y = torch.meshgrid(cc, cc, indexing="ij")
        rr = (torch.arange(nr, device="cuda") + 0.5) / nr * (1.0 / G) + 1.0 / G
        E = torch.zeros(nb * sub * nb * sub, device="cuda", dtype=torch.float64)
        for rv in rr:
            p = torch.stack([torch.full_like(X.reshape(-1), float(rv)), torch.zeros_like(X.reshape(-1)), X.reshape(-1), Y.reshape(-1)], 1)
            E += pkg.mcmc.brdf_target_density(m, p, domain).double() * float(rv)
        E = E.view(nb, sub, nb, sub).sum((1, 3)).cpu()
    else:
        dphi = torch.remainder(sel[:, 3] - sel[:, 1] + math.pi, 2 * math.pi) - math.pi
        nb = 8
        H = torch.histogramdd(torch.stack([sel[:, 2], dphi], 1).cpu().double(), bins=[nb, nb],
                              range=[0.0, math.pi / 2, -math.pi, math.pi]).hist
        sub, nr = 8, 24
        tt = (torch.arange(nb * sub, device="cuda") + 0.5) / (nb * sub) * (math.pi / 2)
        pp = (torch.arange(nb * sub, device="cuda") + 0.5) / (nb * sub) * 2 * math.pi - math.pi
        X, Y = torch.meshgrid(tt, pp, indexing="ij")
        rr = ((torch.arange(nr, device="cuda") + 0.5) / nr * (1.0 / G) + 1.0 / G) * (math.pi / 2)
        E = torch.zeros(nb * sub * nb * sub, device="cuda", dtype=torch.float64)
        for tv in rr:
            # walkers keep |phi| < pi for BOTH angles; with phi_i = 0 every dphi in (-pi, pi) is admissible
            p = torch.stack([torch.full_like(X.reshape(-1), float(tv)), torch.zeros_like(X.reshape(-1)), X.reshape(-1), Y.reshape(-1)], 1)
            E += pkg.mcmc.brdf_target_density(m, p, domain).double()
        E = E.view(nb, sub, nb, sub).sum((1, 3)).cpu()
    E = E / E.sum() * H.sum()
    keep = E >= 20
    chi2 = float((((H - E) ** 2) / torch.where(keep, E, torch.ones_like(E)))[keep].sum())
    dof = int(keep.sum()) - 1
    print(f"[mcmc {domain}] ring 1: {int(H.sum())} samples, chi2/dof = {chi2 / dof:.2f} (dof {dof})")
    assert chi2 / dof < 4.0, (chi2, dof)               # thinned chain, midpoint quadrature: near 1, far from a mismatch (>> 50)
    flat = torch.full_like(E, float(H.sum()) / E.numel())
    assert float((((H - flat) ** 2) / flat).sum()) / dof > 50.0      # power: a uniform omega_o is rejected
