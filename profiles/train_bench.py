#!/usr/bin/env python
"""Training-step timing: one diffusion / rectify-stage iteration at the reference's batch size
(4.9 M rows, learning_repo_cleanup/disk_domain_sampling.py:144-148) on one B200 -- the fused forward + backward + Adam
launch against the reference's own nn.Module + autograd + torch.optim.Adam in eager PyTorch on the SAME GPU (the
reference's real deployment; module taken from oracle/_ref when it is there).

    python profiles/train_bench.py [--rows 4900000] [--workload disk|spherical]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bsdf_diffusion_sampling_b200 as pkg      # noqa: E402


def timed(fn, steps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4_900_000)
    ap.add_argument("--workload", default="disk", choices=["disk", "spherical"])
    args = ap.parse_args()
    n, dev = args.rows, torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(1)
    if args.workload == "disk":
        make = lambda M: M.NN_cond_pos_simpler(input_dim=5, output_dim=2, N_NEURONS=32, POSITIONAL_ENCODING_BASIS_NUM=5)
        omega_i = torch.rand(n, 2, device=dev, generator=g) * 1.2 - 0.6
        omega_o = torch.rand(n, 2, device=dev, generator=g) * 1.6 - 0.8
        x_0 = omega_i * 0.5 + 0.3 * torch.randn(n, 2, device=dev, generator=g)
        flops = 3 * 2 * (25 * 32 + 2 * 1024 + 64)
    else:
        make = lambda M: M.NN_cond_pos(input_dim=6, output_dim=2, N_NEURONS=32, POSITIONAL_ENCODING_BASIS_NUM=5)
        omega_i, omega_o, x_0 = (torch.stack([torch.rand(n, device=dev, generator=g) * 1.5,
                                              torch.rand(n, device=dev, generator=g) * 6.2 - 3.1], 1) for _ in range(3))
        flops = 3 * 2 * (26 * 32 + 3 * 1024 + 64)
    torch.manual_seed(0)
    ours_net = make(pkg.model)
    tr = pkg.training.FlowMatchingTrainer.from_module(ours_net, lr=1e-3)
    ms_ours = timed(lambda: tr.step(x_0, omega_o, omega_i), 10)
    out = {"workload": args.workload, "rows": n, "ms_per_step_fused": ms_ours, "rows_per_s_fused": n / (ms_ours * 1e-3),
           "algorithmic_tflops_fused": n * flops / (ms_ours * 1e-3) / 1e12,
           "note": "algorithmic FLOPs = 3 x forward (forward, activation gradients, weight gradients), unpadded shapes"}
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        from _ref import ref_model as M             # the reference's rendering/utils/model.py, byte for byte
        torch.manual_seed(0)
        net = make(M).to(dev)
        opt = torch.optim.Adam(net.parameters(), lr=1e-3)

        def ref_step():
            alpha = torch.linspace(0, 1, n, device=dev).reshape(-1, 1)
            if args.workload == "disk":
                x_alpha = (1 - alpha) * x_0 + alpha * omega_o
                pred = net(x_alpha, alpha, omega_i)
                loss = torch.mean((pred - (omega_o - x_0)) ** 2)
            else:
                twopi = np.pi * 2
                oo = omega_o.clone()
                tmp = oo[:, 1] - x_0[:, 1]
                oo[:, 1] = torch.where(tmp < -np.pi, oo[:, 1] + twopi, torch.where(tmp > np.pi, oo[:, 1] - twopi, oo[:, 1]))
                x_alpha = (1 - alpha) * x_0 + alpha * oo
                emb = torch.cat([x_alpha[:, 0:1], torch.sin(x_alpha[:, 1:2]), torch.cos(x_alpha[:, 1:2])], dim=1)
                pred = net(emb, alpha, omega_i)
                ph = torch.where(tmp < -np.pi, tmp + twopi, torch.where(tmp > np.pi, tmp - twopi, tmp))
                loss = torch.mean((pred - torch.stack([oo[:, 0] - x_0[:, 0], ph], 1)) ** 2)
            loss.backward()
            opt.step()
            opt.zero_grad()
        ms_ref = timed(ref_step, 5)
        out.update({"ms_per_step_reference_eager_gpu": ms_ref, "speedup_vs_reference_eager_gpu": ms_ref / ms_ours,
                    "reference": "oracle/_ref/ref_model.py (rendering/utils/model.py) + autograd + torch.optim.Adam, eager, same GPU"})
    except Exception as ex:                          # noqa: BLE001
        out["reference_error"] = repr(ex)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
