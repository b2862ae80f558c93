"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Runs only in the build container, where the reference checkout is mounted read-only at
/root/reference (it does not exist on the GPU box; nothing at test time imports this file).

Recipe:
  * ``rendering/utils/model.py`` imports as-is (torch + numpy only).
  * ``rendering/utils/mlp_brdf_sampling.py`` cannot be imported (imageio / matplotlib / OpenEXR at the
    top, device hard-coded to 'cuda'); its five functions are exec'd from the source text starting
    at ``def network_sampling_disk(`` with the literal device string replaced by 'cpu'.
  * ``D_base.sample`` is wrapped to capture the base sample x0 so it can be replayed.
  * checkpoints: ``rendering/checkpoints_new/...`` (``map_location='cpu'``; saved from CUDA).

Output: one ``<name>.npz`` per material holding the network weights (these checkpoints are the
only thing that pins the reference's results) and the reference's outputs in fp32.

Usage:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
CKPT = os.path.join(REF, "rendering", "checkpoints_new")


def load_reference():
    sys.path.insert(0, os.path.join(REF, "rendering"))
    import utils.model as ref_model  # noqa  (the reference's own module, unmodified)
    src = open(os.path.join(REF, "rendering", "utils", "mlp_brdf_sampling.py")).read()
    src = src[src.index("def network_sampling_disk("):]
    src = src.replace("'cuda'", "'cpu'").replace('"cuda"', '"cpu"')
    ns = {"torch": torch}
    exec(compile(src, "mlp_brdf_sampling.py[patched device]", "exec"), ns)
    return ref_model, ns


def capture_x0(D_base):
    box = {}
    orig = D_base.sample

    def wrapped(x_co, numsamples=1):
        x0 = orig(x_co, numsamples)
        box["x0"] = x0.detach().clone()
        return x0

    D_base.sample = wrapped
    return box


def wi_sets(domain, kind, g):
    """fixed wi x64, stratified 64, random 2048 (domain coordinates)."""
    if domain == "disk":
        fixed = torch.tensor([[0.3, -0.2]]).repeat(64, 1)
        u = (torch.stack(torch.meshgrid(torch.arange(8), torch.arange(8), indexing="ij"), -1).reshape(-1, 2)
             + torch.rand(64, 2, generator=g)) / 8
        r, ph = 0.95 * torch.sqrt(u[:, 0]), 2 * np.pi * u[:, 1]
        strat = torch.stack([r * torch.cos(ph), r * torch.sin(ph)], 1)
        r, ph = 0.97 * torch.sqrt(torch.rand(2048, generator=g)), 2 * np.pi * torch.rand(2048, generator=g)
        rnd = torch.stack([r * torch.cos(ph), r * torch.sin(ph)], 1)
    else:
        tmax = np.pi if kind == "bsdf" else np.pi / 2
        fixed = torch.tensor([[0.6, 0.9]]).repeat(64, 1)
        u = (torch.stack(torch.meshgrid(torch.arange(8), torch.arange(8), indexing="ij"), -1).reshape(-1, 2)
             + torch.rand(64, 2, generator=g)) / 8
        strat = torch.stack([u[:, 0] * tmax, u[:, 1] * 2 * np.pi - np.pi], 1)
        rnd = torch.stack([torch.rand(2048, generator=g) * tmax,
                           torch.rand(2048, generator=g) * 2 * np.pi - np.pi], 1)
    return torch.cat([fixed, strat, rnd], 0).float()


def to_np(t):
    return t.detach().cpu().numpy().astype(np.float32)


def make(ref_model, ns, name, domain, kind, flow_path, base_path, complex_path=None):
    g = torch.Generator().manual_seed(1234)
    torch.manual_seed(0)
    if domain == "disk":
        D_sample = ref_model.NN_cond_pos_simpler(input_dim=5, output_dim=2, N_NEURONS=32,
                                                 POSITIONAL_ENCODING_BASIS_NUM=5)
        D_base = ref_model.NN_cond_pretrain_disk_one(input_dim=2, N_NEURONS=16, POSITIONAL_ENCODING_BASIS_NUM=3)
        f_sample, f_pdf, T = ns["network_sampling_disk"], ns["network_pdf_disk"], 4
    else:
        D_sample = ref_model.NN_cond_pos(input_dim=6, output_dim=2, N_NEURONS=32, POSITIONAL_ENCODING_BASIS_NUM=5)
        D_base = ref_model.NN_cond_pretrain_spherical_one(input_dim=2, N_NEURONS=16)
        f_sample, f_pdf, T = ns["network_sampling_spherical"], ns["network_pdf_spherical"], 8
    D_sample.load_state_dict(torch.load(flow_path, map_location="cpu"))
    D_base.load_state_dict(torch.load(base_path, map_location="cpu"))
    D_sample.eval()
    box = capture_x0(D_base)

    wi = wi_sets(domain, kind, g)
    x, pdf = f_sample(D_base, D_sample, wi, T=T)
    x0 = box["x0"]
    # pdf evaluation points: the reference's own samples, plus uniform points of the domain
    n = wi.shape[0]
    if domain == "disk":
        r, ph = 0.99 * torch.sqrt(torch.rand(n, generator=g)), 2 * np.pi * torch.rand(n, generator=g)
        wo_u = torch.stack([r * torch.cos(ph), r * torch.sin(ph)], 1)
    else:
        tmax = np.pi if kind == "bsdf" else np.pi / 2
        wo_u = torch.stack([torch.rand(n, generator=g) * tmax, torch.rand(n, generator=g) * 2 * np.pi - np.pi], 1)
    wo_eval = torch.cat([x.detach(), wo_u.float()], 0)
    wi_eval = torch.cat([wi, wi], 0)
    pdf_eval = f_pdf(D_base, D_sample, wo_eval, wi_eval, T=T)

    # a second T (the reference's functions take T as an argument)
    T2 = 16 if domain == "disk" else 3
    torch.manual_seed(1)
    x_t2, pdf_t2 = f_sample(D_base, D_sample, wi[:256], T=T2)
    x0_t2 = box["x0"]
    pdf_eval_t2 = f_pdf(D_base, D_sample, wo_eval[:256], wi_eval[:256], T=T2)

    out = {"domain": np.int32(0 if domain == "disk" else 1), "T": np.int32(T), "T2": np.int32(T2),
           "kind": np.array(kind)}
    sd = D_sample.state_dict()
    keys = [k for k in sd if k.startswith("linear")] + ["output.weight"]
    out["n_flow_layers"] = np.int32(len(keys))
    for i, k in enumerate(keys):
        out[f"flow_w{i}"] = to_np(sd[k])
    bd = D_base.state_dict()
    out.update(base_w1=to_np(bd["linear1.weight"]), base_b1=to_np(bd["linear1.bias"]),
               base_wo=to_np(bd["output.weight"]), base_bo=to_np(bd["output.bias"]))
    out.update(wi=to_np(wi), x0=to_np(x0), x=to_np(x), pdf_sample=to_np(pdf),
               wo_eval=to_np(wo_eval), wi_eval=to_np(wi_eval), pdf_eval=to_np(pdf_eval),
               x0_t2=to_np(x0_t2), x_t2=to_np(x_t2), pdf_sample_t2=to_np(pdf_t2), pdf_eval_t2=to_np(pdf_eval_t2))

    # reflow (dosampling) in fp32 with the PyTorch module the tcnn net mirrors, T=32, N=1024:
    # disk: the same 32-wide net;  spherical: the 64-wide 6-hidden "complex" net.
    with torch.no_grad():
        Tr = 32
        wir, x0r = wi[128:128 + 1024], x0[128:128 + 1024]
        if domain == "disk":
            net = D_sample
        else:
            net = ref_model.NN_cond_pos_spherical_complicate(input_dim=6, output_dim=2, N_NEURONS=64,
                                                             POSITIONAL_ENCODING_BASIS_NUM=5)
            net.load_state_dict(torch.load(complex_path, map_location="cpu"))
            csd = net.state_dict()
            ckeys = [k for k in csd if k.startswith("linear")] + ["output.weight"]
            out["n_reflow_layers"] = np.int32(len(ckeys))
            for i, k in enumerate(ckeys):
                out[f"reflow_w{i}"] = to_np(csd[k])
        xa = x0r.clone()
        for t in range(Tr):
            alpha = t / Tr * torch.ones(xa.shape[0], 1)
            if domain == "disk":
                d = net(xa, alpha, wir)
            else:
                x2d = torch.cat([xa[:, 0:1], torch.sin(xa[:, 1:2]), torch.cos(xa[:, 1:2])], 1)
                d = net(x2d, alpha, wir)
            xa = xa + 1 / Tr * d
        out.update(reflow_T=np.int32(Tr), reflow_wi=to_np(wir), reflow_x0=to_np(x0r), reflow_x=to_np(xa))

    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}: N={n} x range [{x.min():.3f},{x.max():.3f}] pdf median {pdf.median():.4f} "
          f"neg pdfs {(pdf <= 0).sum().item()}")


def main():
    ref_model, ns = load_reference()
    for m in ["aniso_brushed_aluminium_1_rgb", "cc_nothern_aurora_rgb", "vch_silk_blue_rgb"]:
        make(ref_model, ns, f"disk_{m}", "disk", "measured",
             f"{CKPT}/{m}_disk/brdf_rectify_network{m}.pth", f"{CKPT}/{m}_disk/brdf_pretrain_network{m}.pth")
    for m in ["aniso_brushed_aluminium_1_rgb", "chm_mint_rgb", "ilm_solo_m_68_rgb"]:
        # quirk preserved: the measured-spherical plugin loads the *_disk* pretrain checkpoint as
        # the spherical base net (rendering/brdf_measured_spherical.py:59)
        make(ref_model, ns, f"spherical_{m}", "spherical", "measured",
             f"{CKPT}/{m}_spherical/brdf_rectify_network{m}.pth", f"{CKPT}/{m}_disk/brdf_pretrain_network{m}.pth",
             f"{CKPT}/{m}_spherical/brdf_diffusion_network_complex{m}.pth")
    for k in [0, 12]:
        make(ref_model, ns, f"bsdf_{k}", "spherical", "bsdf",
             f"{CKPT}/bsdf_{k}_spherical/brdf_rectify_network{k}.pth",
             f"{CKPT}/bsdf_{k}_spherical/brdf_pretrain_network{k}.pth",
             f"{CKPT}/bsdf_{k}_spherical/brdf_diffusion_network_complex{k}.pth")


if __name__ == "__main__":
    main()
