#!/usr/bin/env python
"""GPU half of the fix-up flag study (profiles/flag_study.py is the CPU half): for every material of oracle/_ref/all_*.bsdfpack
and n stratified incident directions, the RAW errors of the tensor-core launch WITHOUT the fix-up against the fp32 kernel
(sample pdf, pdf(), |dx|), with the inputs needed to recompute per-row conditioning features on the CPU.
    python profiles/flag_dump.py gpurun_out/flag_dump.npz [n_side]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "profiles"))
import bsdf_diffusion_sampling_b200 as pkg          # noqa: E402
from bsdf_diffusion_sampling_b200.materials import MaterialPack   # noqa: E402
from material_sweep import domain_wi                # noqa: E402

out, n_side = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 128
d = {}
for kind in ("disk", "spherical", "bsdf"):
    pack = MaterialPack.load(os.path.join(ROOT, "oracle", "_ref", f"all_{kind}.bsdfpack"))
    for e in pack.entries:
        T = e["T"]
        pf = pkg.weights.pack_flow_layers(e["flow"], "cuda")
        pb = torch.from_numpy(np.array(e["base"])).cuda()
        wi = torch.from_numpy(domain_wi(kind, n_side, 11)).cuda()
        x32, p32, x0 = pkg.ops.sample(wi, pf, pb, T, seed=20261018, offset=0, precision="fp32")
        x16, p16, _ = pkg.ops.sample(wi, pf, pb, T, x0=x0, precision="tc16", fixup=0.0)
        q32 = pkg.ops.pdf(x32, wi, pf, pb, T, precision="fp32")
        q16 = pkg.ops.pdf(x32, wi, pf, pb, T, precision="tc16", fixup=0.0)
        k = f"{kind}/{e['name']}"
        d[k + "/wi"] = wi.cpu().numpy()
        d[k + "/x0"] = x0.cpu().numpy()
        d[k + "/x32"] = x32.cpu().numpy()
        with np.errstate(all="ignore"):
            d[k + "/rs"] = np.minimum(np.abs(p16.cpu().numpy() / p32.cpu().numpy() - 1.0), 6e4).astype(np.float16)
            d[k + "/rp"] = np.minimum(np.abs(q16.cpu().numpy() / q32.cpu().numpy() - 1.0), 6e4).astype(np.float16)
        d[k + "/dx"] = np.abs(x16.cpu().numpy() - x32.cpu().numpy()).max(1).astype(np.float16)
np.savez_compressed(out, **d)
print(out, os.path.getsize(out))
