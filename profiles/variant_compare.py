#!/usr/bin/env python
"""A/B harness for tuning builds of libbsdfdiff.so (GPU box).  For every library given on the command line it runs,
in a fresh subprocess (BSDFDIFF_LIB selects the library):
  * device-resident timing of the fused sample+pdf kernel on the two bench workloads (CUDA events, best of 5)
    and of the pdf() kernel,
  * sha1 of the outputs (wo, pdf) of a 2M-query Philox run -> two builds that should be arithmetically identical
    (e.g. the aliased and the non-aliased TMEM map) must print the same hashes,
  * the error statistics against two goldens.

    python profiles/variant_compare.py variants/libbsdfdiff_old.so bsdf_diffusion_sampling_b200/libbsdfdiff.so ...
"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    import bench
    import bsdf_diffusion_sampling_b200 as pkg
    from oracle import bsdf_oracle as O

    out = {"lib": os.environ.get("BSDFDIFF_LIB", "default")}
    n_side = int(os.environ.get("VC_NSIDE", "4096"))
    for workload, T, F in (("disk", 4, bench.F_DISK), ("spherical", 8, bench.F_SPH)):
        layers, base = bench.load_fixture(workload)
        pf = pkg.weights.pack_flow_layers(layers, "cuda")
        pb = pkg.weights.pack_base_arrays(*base, "cuda")
        wi = torch.from_numpy(bench.synth_wi3(workload, n_side, 0)).cuda()
        n = wi.shape[0]
        epi = "disk" if workload == "disk" else "spherical"
        sampler = pkg.plugins.NeuralBSDFSampler(epi, pf, pb, T=T, precision="tc16")
        wo, pdf = sampler.sample(wi, seed=1234, offset=0)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            wo, pdf = sampler.sample(wi, seed=1234, offset=0)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[f"{workload}_sample_ms"] = best
        out[f"{workload}_qps"] = n / best * 1e3
        out[f"{workload}_frac"] = n / best * 1e3 * F / 1685.5e12
        bestp = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            p2 = sampler.pdf(wi, wo)
            e1.record()
            torch.cuda.synchronize()
            bestp = min(bestp, e0.elapsed_time(e1))
        out[f"{workload}_pdf_ms"] = bestp
        m = 1 << 21
        h = hashlib.sha1()
        h.update(wo[:m].cpu().numpy().tobytes()); h.update(pdf[:m].cpu().numpy().tobytes()); h.update(p2[:m].cpu().numpy().tobytes())
        out[f"{workload}_sha1"] = h.hexdigest()[:16]
        pn = pdf.double()
        out[f"{workload}_pdf_sum"] = float(torch.nan_to_num(pn, nan=0.0, posinf=0.0, neginf=0.0).sum())
        out[f"{workload}_nonfinite"] = int((~torch.isfinite(pdf)).sum())
        # golden error statistics (raw epilogue, replayed x0)
        flow, basew, z = O.load_material_npz(os.path.join(ROOT, "tests", "golden", bench.MATERIAL[workload] + ".npz"))
        x, pg, _ = pkg.ops.sample(torch.from_numpy(z["wi"]).cuda(), pf, pb, T, x0=torch.from_numpy(z["x0"]).cuda(), precision="tc16")
        pe = pkg.ops.pdf(torch.from_numpy(z["wo_eval"]).cuda(), torch.from_numpy(z["wi_eval"]).cuda(), pf, pb, T, precision="tc16")
        x, pg, pe = x.cpu().numpy(), pg.cpu().numpy(), pe.cpu().numpy()
        rs = np.abs(pg - z["pdf_sample"]) / np.maximum(np.abs(z["pdf_sample"]), 1e-6)
        rp = np.abs(pe - z["pdf_eval"]) / np.maximum(np.abs(z["pdf_eval"]), 1e-6)
        dx = np.abs(x - z["x"]).ravel()
        out[f"{workload}_err"] = "dx med %.2e p99 %.2e | pdf_s med %.2e p99 %.2e | pdf_e med %.2e p99 %.2e" % (
            np.median(dx), np.quantile(dx, 0.99), np.nanmedian(rs), np.nanquantile(rs, 0.99), np.nanmedian(rp),
            np.nanquantile(rp, 0.99))
    out["timeout_flag"] = int(pkg._lib.lib.bsdfdiff_debug_timeout_flag())
    print("VC " + json.dumps(out), flush=True)


def main():
    if os.environ.get("VC_CHILD"):
        return child()
    for lib in sys.argv[1:]:
        env = dict(os.environ, VC_CHILD="1", BSDFDIFF_LIB=os.path.abspath(lib))
        r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True, timeout=600)
        lines = [l for l in r.stdout.splitlines() if l.startswith("VC ")]
        if not lines:
            print(f"{lib}: FAILED rc={r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-3000:]}", flush=True)
            continue
        d = json.loads(lines[0][3:])
        print(f"== {lib}")
        for k, v in d.items():
            if k != "lib":
                print(f"   {k}: {v}")
        sys.stdout.flush()


if __name__ == "__main__":
    main()
