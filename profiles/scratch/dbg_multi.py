import sys, os, glob
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import bsdf_diffusion_sampling_b200 as pkg
from oracle import bsdf_oracle as O
files = sorted(glob.glob("tests/golden/spherical_*.npz"))
for prec in ("fp32", "tc16"):
    mats = []
    for path in files:
        flow, base, _ = O.load_material_npz(path)
        mats.append(pkg.plugins.NeuralBSDFSampler("spherical", pkg.weights.pack_flow_layers(flow.layers, "cuda"),
                    pkg.weights.pack_base_arrays(base.w1, base.b1, base.wo, base.bo, "cuda"), precision=prec))
    mm = pkg.plugins.MultiMaterialSampler(mats)
    rng = np.random.default_rng(5)
    n = 50001
    w = rng.normal(size=(n, 3)).astype(np.float32); w[:, 2] = np.abs(w[:, 2]) + 0.05
    wi = torch.from_numpy(w / np.linalg.norm(w, axis=1, keepdims=True)).cuda()
    mid = torch.from_numpy(rng.integers(0, 3, n).astype(np.int32)).cuda()
    plan = mm.plan(mid)
    wo, pdf = mm.sample(wi, plan=plan, seed=11, offset=40)
    p2 = mm.pdf(wi, wo, plan=plan)
    for m, s in enumerate(mats):
        sel = (mid == m).nonzero().squeeze(1)
        q = s.pdf(wi, wo)[sel]; r = p2[sel]
        bad = (q.view(torch.int32) != r.view(torch.int32))
        print(prec, m, "mismatch", int(bad.sum()), "nan", int(torch.isnan(q).sum()), int(torch.isnan(r).sum()),
              "maxrel", float(((q - r).abs() / q.abs().clamp_min(1e-20))[bad].max()) if bad.any() else 0.0)
        if bad.any():
            i = bad.nonzero()[:5, 0]
            print("   ", q[i].tolist(), r[i].tolist())
