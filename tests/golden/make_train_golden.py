"""Golden fixtures for the training step (tests/golden/train/train_*.npz), generated from the UNMODIFIED reference.

Runs only in the build container (/root/reference).  For each net shape the reference's own nn.Module
(learning_repo_cleanup/utils/model.py) is initialised with a fixed seed, one batch is pushed through the loss of the
reference's diffusion / rectify stage (disk_domain_sampling.py:49-58; spherical_domain_sampling.py:55-76, copied here as
the few tensor lines they are) with torch autograd and torch.optim.Adam on the CPU in fp32, and the inputs, the loss, the
gradients of the first step and the weights after 3 steps are stored.

Usage:  python tests/golden/make_train_golden.py
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, os.path.join(REF, "learning_repo_cleanup"))
    import utils.model as M  # noqa  (the reference's module, unmodified)

    torch.set_num_threads(8)
    cases = [
        ("disk", lambda: M.NN_cond_pos_simpler(input_dim=5, output_dim=2, N_NEURONS=32, POSITIONAL_ENCODING_BASIS_NUM=5), 4099),
        ("spherical", lambda: M.NN_cond_pos(input_dim=6, output_dim=2, N_NEURONS=32, POSITIONAL_ENCODING_BASIS_NUM=5), 4099),
        ("spherical_complicate", lambda: M.NN_cond_pos_spherical_complicate(input_dim=6, output_dim=2, N_NEURONS=64,
                                                                             POSITIONAL_ENCODING_BASIS_NUM=5), 1031),
    ]
    for name, make, n in cases:
        torch.manual_seed(1234)
        net = make()
        g = torch.Generator().manual_seed(99)
        if name == "disk":
            omega_i = torch.rand(n, 2, generator=g) * 1.2 - 0.6
            omega_o = torch.rand(n, 2, generator=g) * 1.6 - 0.8
            x_0 = omega_i * 0.5 + 0.3 * torch.randn(n, 2, generator=g)
        else:
            omega_i = torch.stack([torch.rand(n, generator=g) * 1.5, torch.rand(n, generator=g) * 6.2 - 3.1], 1)
            omega_o = torch.stack([torch.rand(n, generator=g) * 1.5, torch.rand(n, generator=g) * 6.2 - 3.1], 1)
            x_0 = torch.stack([torch.rand(n, generator=g) * 1.5, torch.rand(n, generator=g) * 6.2 - 3.1], 1)
        w_before = [p.detach().clone().numpy() for p in net.parameters()]
        opt = torch.optim.Adam(net.parameters(), lr=0.001)
        losses, grads = [], None
        for it in range(3):
            alpha = torch.linspace(0, 1, n).reshape(-1, 1)
            if name == "disk":
                # disk_domain_sampling.py:49-58
                x_alpha = (1 - alpha) * x_0 + alpha * omega_o
                pred = net(x_alpha, alpha, omega_i)
                loss = torch.mean((pred - (omega_o - x_0)) ** 2)
            else:
                # spherical_domain_sampling.py:59-75 (omega_o is modified in place there: work on a copy)
                twopi = np.pi * 2
                oo = omega_o.clone()
                tmp_dirc = oo[:, 1] - x_0[:, 1]
                oo[:, 1] = torch.where(tmp_dirc < -np.pi, oo[:, 1] + twopi, torch.where(tmp_dirc > np.pi, oo[:, 1] - twopi, oo[:, 1]))
                x_alpha = (1 - alpha) * x_0 + alpha * oo
                x_alpha_predioc = torch.cat([torch.sin(x_alpha[:, 1]).reshape(-1, 1), torch.cos(x_alpha[:, 1]).reshape(-1, 1)], dim=1)
                x_alpha_2d = torch.cat([x_alpha[:, 0].reshape(-1, 1), x_alpha_predioc], dim=1)
                pred = net(x_alpha_2d, alpha, omega_i)
                thetadirc = oo[:, 0] - x_0[:, 0]
                phidirc = torch.where(tmp_dirc < -np.pi, tmp_dirc + twopi, torch.where(tmp_dirc > np.pi, tmp_dirc - twopi, tmp_dirc))
                twod_dirc = torch.cat([thetadirc.reshape(-1, 1), phidirc.reshape(-1, 1)], dim=1)
                loss = torch.mean((pred - twod_dirc) ** 2)
            loss.backward()
            if it == 0:
                grads = [p.grad.detach().clone().numpy() for p in net.parameters()]
            losses.append(float(loss))
            opt.step()
            opt.zero_grad()
        w_after = [p.detach().clone().numpy() for p in net.parameters()]
        out = {"n_layers": len(w_before), "x_0": x_0.numpy(), "omega_o": omega_o.numpy(), "omega_i": omega_i.numpy(),
               "losses": np.array(losses, np.float64)}
        for i, (a, b, c) in enumerate(zip(w_before, grads, w_after)):
            out[f"w{i}"], out[f"g{i}"], out[f"w_after{i}"] = a, b, c
        path = os.path.join(OUT, "train", f"train_{name}.npz")
        np.savez_compressed(path, **out)
        print(name, "n =", n, "losses", losses, "->", path, os.path.getsize(path), "bytes")


def pretrain():
    """Pretrain-stage fixtures: the reference's base nets, -mean(log_prob) + autograd + Adam(lr 3e-4), 3 steps."""
    sys.path.insert(0, os.path.join(REF, "learning_repo_cleanup"))
    import utils.model as M  # noqa
    for name, make, n in (("base_disk", lambda: M.NN_cond_pretrain_disk_one(input_dim=2, N_NEURONS=16, POSITIONAL_ENCODING_BASIS_NUM=3), 4099),
                          ("base_spherical", lambda: M.NN_cond_pretrain_spherical_one(input_dim=2, N_NEURONS=16, POSITIONAL_ENCODING_BASIS_NUM=3), 4099)):
        torch.manual_seed(4321)
        net = make()
        with torch.no_grad():                       # spread the predicted parameters (kappa on both sides of 3.75)
            net.output.bias.add_(torch.tensor([0.1, -0.3, 0.2, 1.5]))
            net.output.weight.mul_(3.0)
        g = torch.Generator().manual_seed(7)
        if name == "base_disk":
            omega_i = torch.rand(n, 2, generator=g) * 1.2 - 0.6
            omega_o = torch.rand(n, 2, generator=g) * 1.6 - 0.8
        else:
            omega_i = torch.stack([torch.rand(n, generator=g) * 1.5, torch.rand(n, generator=g) * 6.2 - 3.1], 1)
            omega_o = torch.stack([torch.rand(n, generator=g) * 1.5, torch.rand(n, generator=g) * 6.2 - 3.1], 1)
        order = ("linear1.weight", "linear1.bias", "output.weight", "output.bias")
        flat = lambda sd: np.concatenate([sd[k].detach().numpy().ravel() for k in order])
        w_before = flat(net.state_dict())
        opt = torch.optim.Adam(net.parameters(), lr=0.0003)
        losses, grad = [], None
        for it in range(3):
            logp = net.log_prob(omega_o, omega_i)
            loss = -torch.mean(logp)
            loss.backward()
            if it == 0:
                grad = flat({k: dict(net.named_parameters())[k].grad for k in order})
            losses.append(float(loss.detach()))
            opt.step()
            net.zero_grad()
        path = os.path.join(OUT, "train", f"train_{name}.npz")
        np.savez_compressed(path, w=w_before, g=grad, w_after=flat(net.state_dict()), omega_o=omega_o.numpy(),
                            omega_i=omega_i.numpy(), losses=np.array(losses, np.float64))
        print(name, "losses", losses, "->", path)


if __name__ == "__main__":
    main()
    pretrain()
