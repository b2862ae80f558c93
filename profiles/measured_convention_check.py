#!/usr/bin/env python
"""Which cosine convention does Mitsuba's `measured` eval use?  (CPU, needs /root/reference for the .bsdf files.)

Mitsuba is absent here, so the oracle's eval cannot be compared with it.  The reference's shipped disk-domain networks were
trained on  lum(mitsuba_eval(wi, wo)) * clamp(1/cos(theta_o), 1, 1e6)  (learning_repo_cleanup/utils/mitsuba_brdf_scalar.py:
83-89), so their sampling density tells which candidate is right.  With S = rgb(vndf^-1(u_m)) ndf(u_m) / (4 sigma(u_wi)) (the
tabulated value): eval = S gives the target S/cos, eval = S cos gives S.  Total-variation distance between the histogram of the
reference algorithm's samples (oracle port, 400 K samples per wi) and each normalised candidate, 3 materials x 6 random wi."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bsdf_oracle as O, c_oracle as C, measured_oracle as M  # noqa: E402

tot = {}
for mat in ("vch_silk_blue_rgb", "cc_nothern_aurora_rgb", "aniso_brushed_aluminium_1_rgb"):
    b = M.MeasuredBSDF.from_file(f"/root/reference/rendering/measuredbsdfs/{mat}.bsdf")
    flow, base, z = O.load_material_npz(os.path.join(ROOT, "tests", "golden", f"disk_{mat}.npz"))
    rng = np.random.default_rng(1)
    for k in range(6):
        r, a = 0.85 * np.sqrt(rng.random()), rng.random() * 2 * np.pi
        wi_xy = [r * np.cos(a), r * np.sin(a)]
        g = 160
        xs = (np.arange(g) + 0.5) / g * 2 - 1
        X, Y = np.meshgrid(xs, xs, indexing="ij")
        wo = np.stack([X.ravel(), Y.ravel()], 1)
        inside = (wo ** 2).sum(1) < 0.98
        wi = np.tile(np.array([wi_xy]), (wo.shape[0], 1))
        cz = np.maximum(M.disk_to_dir(wo)[:, 2], 1e-3)
        S = np.where(inside, M.rgb2lum(b.eval(M.disk_to_dir(wi), M.disk_to_dir(wo))), 0)      # the oracle's eval returns S
        cell = (2 / g) ** 2
        n = 400000
        wis = np.tile(np.array([wi_xy], np.float32), (n, 1))
        x, _ = C.sample(flow, base, wis, 4, O.draw_x0_disk(base, wis, rng))
        H, _, _ = np.histogram2d(x[:, 0], x[:, 1], bins=g, range=[[-1, 1], [-1, 1]])
        pn = H.ravel() * inside
        pn = pn / (pn.sum() * cell)
        for name, t in (("S / cos^2", S / cz ** 2), ("S / cos   (eval = S)", S / cz), ("S         (eval = S cos)", S), ("S cos", S * cz)):
            tn = t / (t.sum() * cell)
            tot.setdefault(name, []).append(0.5 * np.abs(tn - pn).sum() * cell)
for k, v in tot.items():
    print(f"target {k:26s} mean TV {np.mean(v):.4f}   best in {sum(v[i] == min(tot[q][i] for q in tot) for i in range(len(v)))} of {len(v)} cases")
