#!/usr/bin/env python
"""Package the reference's OWN PyTorch implementation of the hot path into oracle/_ref/ (git-ignored, travels to the GPU box).

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference's path is Python, so "building" it means extracting the files where
they lie under /root/reference (nothing is copied into the tracked tree):

    oracle/_ref/ref_model.py      = rendering/utils/model.py, byte for byte
    oracle/_ref/ref_sampler.py    = rendering/utils/mlp_brdf_sampling.py from `def network_sampling_disk(` on (the five
                                    sampler functions, :17-181) with the hard-coded device literal 'cuda' replaced by
                                    'cpu' -- the file's top imports (imageio, matplotlib, OpenEXR) do not exist on the box
                                    and are not used by these functions
    oracle/_ref/checkpoints/...   = the rectify + pretrain .pth files of the bench materials
                                    (rendering/checkpoints_new/<mat>_{disk,spherical}/)

`bench.py --impl reference` and bench.py's cpu_baseline leg import oracle/_ref to time the reference's PyTorch code on the
box's host cores (cpu_baseline.kind = "reference"); when oracle/_ref is absent they fall back to the C/OpenMP port.
Run by __graft_entry__.build() whenever /root/reference is present.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
MATERIALS = ["aniso_brushed_aluminium_1_rgb"]


def make(reference_root: str = "/root/reference") -> bool:
    src_model = os.path.join(reference_root, "rendering", "utils", "model.py")
    src_sampler = os.path.join(reference_root, "rendering", "utils", "mlp_brdf_sampling.py")
    if not (os.path.exists(src_model) and os.path.exists(src_sampler)):
        return False
    os.makedirs(OUT, exist_ok=True)
    shutil.copyfile(src_model, os.path.join(OUT, "ref_model.py"))
    text = open(src_sampler).read()
    body = text[text.index("def network_sampling_disk("):]
    body = body.replace("'cuda'", "'cpu'").replace('"cuda"', '"cpu"')
    with open(os.path.join(OUT, "ref_sampler.py"), "w") as f:
        f.write("# extracted by oracle/make_ref.py from rendering/utils/mlp_brdf_sampling.py:17-181 ('cuda' -> 'cpu'); do not edit\n"
                "import torch\n\n" + body)
    open(os.path.join(OUT, "__init__.py"), "w").close()
    ck = os.path.join(reference_root, "rendering", "checkpoints_new")
    for mat in MATERIALS:
        for dom in ("disk", "spherical"):
            d = os.path.join(OUT, "checkpoints", f"{mat}_{dom}")
            os.makedirs(d, exist_ok=True)
            for kind in ("rectify", "pretrain"):
                name = f"brdf_{kind}_network{mat}.pth"
                p = os.path.join(ck, f"{mat}_{dom}", name)
                if os.path.exists(p):
                    shutil.copyfile(p, os.path.join(d, name))
    make_packs(ck)
    return True


def make_packs(ck: str) -> dict:
    """Every material the reference ships (rendering/checkpoints_new: measured disk, measured spherical, bsdf_<k>) as three
    .bsdfpack files under oracle/_ref/ -- input of profiles/material_sweep.py and tests/test_all_materials.py, which check the
    shipped tensor-core path against the fp32 kernel and the oracle on ALL of them, not only on the eight goldens."""
    repo = os.path.dirname(HERE)
    if repo not in sys.path:
        sys.path.insert(0, repo)
    from bsdf_diffusion_sampling_b200 import plugins
    from bsdf_diffusion_sampling_b200.materials import MaterialPack
    dirs = sorted(os.listdir(ck))
    sets = {"disk": [d[:-5] for d in dirs if d.endswith("_disk")],
            "spherical": [d[:-10] for d in dirs if d.endswith("_spherical") and not d.startswith("bsdf_")],
            "bsdf": sorted((d[5:-10] for d in dirs if d.startswith("bsdf_") and d.endswith("_spherical")), key=int)}
    out = {}
    for kind, mats in sets.items():
        ok = [m for m in mats if all(os.path.exists(p) for p in plugins.checkpoint_paths(kind, m, ck))]
        pack = MaterialPack.from_checkpoints(kind, ok, ck)
        pack.save(os.path.join(OUT, f"all_{kind}.bsdfpack"))
        out[kind] = len(ok)
    return out


def load(workload: str):
    """-> (sample_fn(wi [n,2] cpu tensor) -> (x, pdf), pdf_fn(wo, wi) -> pdf) running the reference's functions on the
    reference's modules with the reference's checkpoints, or None when oracle/_ref has not been built."""
    if not os.path.exists(os.path.join(OUT, "ref_sampler.py")):
        return None
    import torch
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    from _ref import ref_model as M
    from _ref import ref_sampler as S
    mat = MATERIALS[0]
    root = os.path.join(OUT, "checkpoints")
    if workload == "disk":
        D = M.NN_cond_pos_simpler(input_dim=5, output_dim=2, N_NEURONS=32, POSITIONAL_ENCODING_BASIS_NUM=5)                       # rendering/brdf_measured_disk.py:43
        B = M.NN_cond_pretrain_disk_one(input_dim=2, N_NEURONS=16, POSITIONAL_ENCODING_BASIS_NUM=3)                    # :49
        fs, fp, T = S.network_sampling_disk, S.network_pdf_disk, 4
        base_dir = f"{mat}_disk"
    else:
        D = M.NN_cond_pos(input_dim=6, output_dim=2, N_NEURONS=32, POSITIONAL_ENCODING_BASIS_NUM=5)                               # rendering/brdf_measured_spherical.py:52
        B = M.NN_cond_pretrain_spherical_one(input_dim=2, N_NEURONS=16)                  # :58 (loads the *_disk* pretrain checkpoint, :59)
        fs, fp, T = S.network_sampling_spherical, S.network_pdf_spherical, 8
        base_dir = f"{mat}_disk"
    dom = "disk" if workload == "disk" else "spherical"
    D.load_state_dict(torch.load(os.path.join(root, f"{mat}_{dom}", f"brdf_rectify_network{mat}.pth"), map_location="cpu"))
    B.load_state_dict(torch.load(os.path.join(root, base_dir, f"brdf_pretrain_network{mat}.pth"), map_location="cpu"))
    D.eval()
    B.eval()
    return (lambda wi: fs(B, D, wi, T)), (lambda wo, wi: fp(B, D, wo, wi, T))


if __name__ == "__main__":
    print("oracle/_ref built" if make(*sys.argv[1:2]) else "reference checkout not found; nothing built")
