"""Training step: the fused forward + backward + Adam launch against fixtures generated from the reference's
own nn.Modules with torch autograd + torch.optim.Adam (tests/golden/make_train_golden.py).

Tolerances: fp32 arithmetic on both sides; the gradient is a sum over the batch accumulated in a different order (tile
sums + atomics), so gradients agree to 2e-5 of the layer's largest entry, the loss to 1e-6 relative, and the weights after
three Adam steps (lr 1e-3: every weight moves by ~3e-3) to 2e-6 absolute.
"""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR

TRAIN_FILES = sorted(f for f in glob.glob(os.path.join(GOLDEN_DIR, "train", "train_*.npz")) if "train_base_" not in f)
BASE_FILES = sorted(glob.glob(os.path.join(GOLDEN_DIR, "train", "train_base_*.npz")))


def ids():
    return [os.path.basename(f)[6:-4] for f in TRAIN_FILES]


def test_train_fixtures_present():
    assert len(TRAIN_FILES) == 3 and len(BASE_FILES) == 2
    for f in TRAIN_FILES + BASE_FILES:
        z = np.load(f)
        assert z["losses"].shape == (3,) and np.all(np.diff(z["losses"]) < 0)      # Adam on a fixed batch: the loss goes down


def test_trainer_host_checks(built_lib):
    T = built_lib.training.FlowMatchingTrainer
    z = np.load(TRAIN_FILES[0])
    layers = [torch.from_numpy(z[f"w{i}"]) for i in range(int(z["n_layers"]))]
    t = T(layers, device="cpu")
    assert (t.in_dim, t.hidden, t.n_hidden) == (25, 32, 3) and t.weights.numel() == 25 * 32 + 2 * 1024 + 64
    assert list(t.state_dict().keys()) == ["linear1.weight", "linear2.weight", "linear3.weight", "output.weight"]
    assert all(torch.equal(a, b) for a, b in zip(t.layers(), layers))
    with pytest.raises(RuntimeError, match="no CPU path"):
        t.step(torch.zeros(4, 2), torch.zeros(4, 2), torch.zeros(4, 2))
    with pytest.raises(ValueError):
        T([torch.zeros(32, 24), torch.zeros(2, 32)], device="cpu")
    with pytest.raises(ValueError):
        T([torch.zeros(32, 25), torch.zeros(32, 16), torch.zeros(2, 32)], device="cpu")
    L = built_lib._lib.lib
    assert L.bsdfdiff_flow_param_count(26, 64, 6) == 26 * 64 + 5 * 4096 + 128
    assert L.bsdfdiff_flow_param_count(25, 48, 3) == 0
    assert L.bsdfdiff_flow_matching_step(0, 32, 3, 16, None, 1, 1, None, 1, 1, 1, 1, 1e-3, 0.9, 0.999, 1e-8, 1, 1, 1, 1,
                                         None) == -1


@pytest.mark.gpu
@pytest.mark.parametrize("path", TRAIN_FILES, ids=ids())
def test_flow_matching_step_matches_reference_autograd(built_lib, path):
    z = np.load(path)
    nl = int(z["n_layers"])
    layers = [torch.from_numpy(z[f"w{i}"]) for i in range(nl)]
    x_0, omega_o, omega_i = (torch.from_numpy(z[k]).cuda() for k in ("x_0", "omega_o", "omega_i"))
    tr = built_lib.training.FlowMatchingTrainer(layers, lr=0.001)
    loss, grads = tr.loss_and_grad(x_0, omega_o, omega_i)
    assert abs(float(loss) - z["losses"][0]) <= 1e-6 * abs(z["losses"][0]) + 1e-7
    for i, g in enumerate(grads):
        ref = z[f"g{i}"]
        err = np.abs(g.cpu().numpy() - ref).max()
        assert err <= 2e-5 * np.abs(ref).max() + 1e-9, f"layer {i}: max |dg| = {err}, max |g| = {np.abs(ref).max()}"
    assert all(torch.equal(a.cpu(), b) for a, b in zip(tr.layers(), layers))        # loss_and_grad leaves the weights alone
    for k in range(3):
        loss = tr.step(x_0, omega_o, omega_i)
        assert abs(float(loss) - z["losses"][k]) <= 2e-6 * abs(z["losses"][k]) + 1e-7, (k, float(loss), z["losses"][k])
    for i, w in enumerate(tr.layers()):
        err = np.abs(w.cpu().numpy() - z[f"w_after{i}"]).max()
        assert err <= 2e-6, f"layer {i}: max |dw| after 3 Adam steps = {err}"
    assert float(tr.grad.abs().max()) == 0.0 and int(tr._ticket[0]) == 0            # buffers handed back clean
    # the trained weights drive the sampler
    pf = tr.packed()
    assert (pf.in_dim, pf.hidden, pf.n_hidden) == (tr.in_dim, tr.hidden, tr.n_hidden)
    sd = tr.state_dict()
    assert sd["output.weight"].shape == (2, tr.hidden)


@pytest.mark.gpu
def test_explicit_alpha_and_batch_tail(built_lib):
    """alpha= overrides the linspace; batches that are not a multiple of the 128-row tile and a single-row batch work."""
    z = np.load(TRAIN_FILES[0])
    layers = [torch.from_numpy(z[f"w{i}"]) for i in range(int(z["n_layers"]))]
    x_0, omega_o, omega_i = (torch.from_numpy(z[k]).cuda() for k in ("x_0", "omega_o", "omega_i"))
    tr = built_lib.training.FlowMatchingTrainer(layers)
    n = x_0.shape[0]
    a, ga = tr.loss_and_grad(x_0, omega_o, omega_i)
    b, gb = tr.loss_and_grad(x_0, omega_o, omega_i, alpha=torch.linspace(0, 1, n))
    assert abs(float(a) - float(b)) < 1e-6 and all(torch.allclose(u, v, atol=1e-7) for u, v in zip(ga, gb))
    # the loss of the whole batch is the mean of its per-row losses: split at a non-tile boundary, with explicit alpha
    al = torch.linspace(0, 1, n)
    k = 1000
    l1, g1 = tr.loss_and_grad(x_0[:k], omega_o[:k], omega_i[:k], alpha=al[:k])
    l2, g2 = tr.loss_and_grad(x_0[k:], omega_o[k:], omega_i[k:], alpha=al[k:])
    assert abs((float(l1) * k + float(l2) * (n - k)) / n - float(a)) < 2e-6
    for u, v, w in zip(g1, g2, ga):
        assert torch.allclose((u * k + v * (n - k)) / n, w, atol=2e-7)
    l0, _ = tr.loss_and_grad(x_0[:1], omega_o[:1], omega_i[:1])
    assert np.isfinite(float(l0))


@pytest.mark.gpu
def test_diffusion_stage_loop_reduces_the_loss(built_lib):
    """A short diffusion_stage run on synthetic (omega_i, omega_o) pairs: device-side batch draw, base sample from the
    sampler library, one training launch per iteration; the loss falls and the trained net samples."""
    from oracle import bsdf_oracle as O
    from conftest import DISK_FILE
    pkg = built_lib
    flow, base, _ = O.load_material_npz(DISK_FILE)
    pb = pkg.weights.pack_base_arrays(base.w1, base.b1, base.wo, base.bo, "cuda")
    g = torch.Generator(device="cuda").manual_seed(5)
    wi = (torch.rand(200_000, 2, device="cuda", generator=g) - 0.5) * 1.2
    wo = 0.6 * wi + 0.1 * torch.randn(200_000, 2, device="cuda", generator=g)      # a narrow lobe around 0.6 wi
    data = torch.cat([wi, wo], 1)
    torch.manual_seed(0)
    net = pkg.model.NN_cond_pos_simpler(input_dim=5, output_dim=2, N_NEURONS=32, POSITIONAL_ENCODING_BASIS_NUM=5)
    tr = pkg.training.FlowMatchingTrainer.from_module(net, lr=2e-3)
    first = float(pkg.training.diffusion_stage_step(tr, pb, data, 65_536, generator=g))
    for _ in range(150):
        last = pkg.training.diffusion_stage_step(tr, pb, data, 65_536, generator=g)
    assert float(last) < 0.5 * first, (first, float(last))
    x, pdf, _ = pkg.ops.sample(wi[:4096], tr.packed(), pb, 4, seed=1)
    assert torch.isfinite(x).all() and torch.isfinite(pdf).all()
    # samples moved from the base distribution towards the lobe
    assert float((x - 0.6 * wi[:4096]).pow(2).mean()) < float((wi[:4096] * 0).add(1).mean())


def test_pretrainer_host_checks(built_lib):
    B = built_lib.training.BasePretrainer
    z = np.load(BASE_FILES[0])
    t = B(z["w"], 0, device="cpu")
    assert list(t.state_dict().keys()) == ["linear1.weight", "linear1.bias", "output.weight", "output.bias"]
    assert t.state_dict()["linear1.weight"].shape == (16, 14) and t.blob().numel() == 308
    net = built_lib.model.NN_cond_pretrain_disk_one(input_dim=2, N_NEURONS=16, POSITIONAL_ENCODING_BASIS_NUM=3)
    net.load_state_dict(t.state_dict())
    assert torch.equal(B.from_module(net, 0, device="cpu").weights, t.weights)
    with pytest.raises(ValueError):
        B(np.zeros(300, np.float32), 0, device="cpu")
    with pytest.raises(ValueError):
        B(z["w"], 2, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU path"):
        t.step(torch.zeros(4, 2), torch.zeros(4, 2))
    assert built_lib._lib.lib.bsdfdiff_base_nll_step(0, 16, None, 1, 1, 1, 1, 1, 3e-4, 0.9, 0.999, 1e-8, 1, 1, 1, 1, None) == -1


@pytest.mark.gpu
@pytest.mark.parametrize("path", BASE_FILES, ids=[os.path.basename(f)[6:-4] for f in BASE_FILES])
def test_base_nll_step_matches_reference_autograd(built_lib, path):
    """Pretrain stage: -mean(log_prob) of the reference's base nets (disk Gaussian; spherical Gaussian x von Mises incl.
    the derivative of torch's log I0 polynomials and of softplus), its gradient, and three Adam(lr 3e-4) steps."""
    z = np.load(path)
    domain = 0 if "disk" in path else 1
    omega_o, omega_i = torch.from_numpy(z["omega_o"]).cuda(), torch.from_numpy(z["omega_i"]).cuda()
    tr = built_lib.training.BasePretrainer(z["w"], domain, lr=0.0003)
    loss, g = tr.loss_and_grad(omega_o, omega_i)
    assert abs(float(loss) - z["losses"][0]) <= 2e-6 * abs(z["losses"][0])
    err = np.abs(g.cpu().numpy() - z["g"])
    assert err.max() <= 3e-5 * np.abs(z["g"]).max(), (err.max(), np.abs(z["g"]).max())
    for k in range(3):
        loss = tr.step(omega_o, omega_i)
        assert abs(float(loss) - z["losses"][k]) <= 3e-6 * abs(z["losses"][k]), (k, float(loss), z["losses"][k])
    assert np.abs(tr.weights.cpu().numpy() - z["w_after"]).max() <= 2e-6
    assert float(tr.grad.abs().max()) == 0.0
    # the trained blob is what the sampler kernels take as `base`; log_prob through the library agrees with the loss
    lp = built_lib.ops.base_log_prob(omega_o, omega_i, tr.blob(), domain)
    nxt, _ = tr.loss_and_grad(omega_o, omega_i)
    assert abs(float(-lp.mean()) - float(nxt)) <= 1e-5 * abs(float(nxt))
