#!/usr/bin/env python
"""Small driver for ncu captures of the round-2 kernels (one launch of each after a warm-up):
    ncu --set full --clock-control none --import-source on -k regex:'flow_tc_kernel|flow_lane8|flow_matching|base_nll|multi_' \\
        -f -o gpurun_out/r2k_targets python profiles/ncu_targets.py
  1. multi_count / multi_tiles / multi_scatter + flow_tc_kernel<.., MULTI> : 12 materials, 4 M rows (disk, sample)
  2. flow_lane8_kernel : full-batch fp32 path, 1 M rows (disk, sample)
  3. flow_matching_step_kernel / base_nll_step_kernel : 4.9 M-row training steps
"""
import glob
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                    # noqa: E402
import bsdf_diffusion_sampling_b200 as pkg      # noqa: E402

dev = torch.device("cuda")
mats = []
files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "disk_*.npz")))
for j in range(12):
    z = np.load(files[j % len(files)])
    mats.append(pkg.plugins.NeuralBSDFSampler(
        "disk", pkg.weights.pack_flow_layers([z[f"flow_w{i}"] for i in range(int(z["n_flow_layers"]))], dev),
        pkg.weights.pack_base_arrays(z["base_w1"], z["base_b1"], z["base_wo"], z["base_bo"], dev), precision="tc16"))
mm = pkg.plugins.MultiMaterialSampler(mats)
wi = torch.from_numpy(bench.synth_wi3("disk", 2048, 1)).to(dev)
mid = torch.randint(0, 12, (wi.shape[0],), device=dev, dtype=torch.int32)
for _ in range(2):
    plan = mm.plan(mid)
    mm.sample(wi, plan=plan, seed=1)
torch.cuda.synchronize()

s32 = pkg.plugins.NeuralBSDFSampler("disk", mats[0].flow, mats[0].base, precision="fp32")
w1m = wi[: 1 << 20].contiguous()
for _ in range(2):
    s32.sample(w1m, seed=2)
torch.cuda.synchronize()

n = 4_900_000
g = torch.Generator(device=dev).manual_seed(1)
omega_i = torch.rand(n, 2, device=dev, generator=g) * 1.2 - 0.6
omega_o = torch.rand(n, 2, device=dev, generator=g) * 1.6 - 0.8
x_0 = omega_i * 0.5 + 0.3 * torch.randn(n, 2, device=dev, generator=g)
torch.manual_seed(0)
tr = pkg.training.FlowMatchingTrainer.from_module(
    pkg.model.NN_cond_pos_simpler(input_dim=5, output_dim=2, N_NEURONS=32, POSITIONAL_ENCODING_BASIS_NUM=5))
pre = pkg.training.BasePretrainer(mats[0].base, 0)
for _ in range(2):
    tr.step(x_0, omega_o, omega_i)
    pre.step(omega_o, omega_i)
torch.cuda.synchronize()
print("done")
