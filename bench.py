#!/usr/bin/env python
"""bench.py -- headline benchmark of the fused neural-BSDF sampler (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload disk|spherical] [--mode sample|pdf] [--queries Q] [--precision tc16|fp32]
                    [--no-e2e] [--no-cpu] [--no-extra]

A "step" is one pass of the hot path over one batch of synthetic queries.  Headline (N=1) = BASELINE.json
configs[1], "measured BRDF disk-domain sampler: 16M (wi, noise) queries, full diffusion steps (T=4, the plugin's
setting), 1 GPU", mode sample (= the fused sample+pdf kernel; the pdf is a by-product of the same launch); for N>1
every rank gets its own 16M-query shard (weak scaling, no collective on the data path; Philox counters are global
row indices).  The shipped product path is timed: the tcgen05 kernel PLUS its conditioning-triggered fp32 fix-up
launch.  Prints ONE JSON line (rank 0):
  value        whole-job samples/s, inputs resident in HBM, CUDA events, max over ranks
  roofline     achieved algorithmic TFLOP/s of the launch pair vs the measured dense-fp16/bf16 tensor peak;
               traffic = DRAM bytes per launch read from the committed ncu summary under profiles/
  e2e          the same metric through plugins.NeuralBSDFSampler.sample_host (pinned host wi in, wo+pdf out, copies
               inside the timed region) + the copy-only ceiling of the same pipeline (kernels skipped)
  extra        the other BASELINE configs as short lines: configs[2] (spherical, T=8: sample, then pdf() of the produced
               wo) and disk pdf()
  cpu_baseline the reference's own PyTorch code (oracle/_ref, built by oracle/make_ref.py) on the host cores on a bounded
               sample, with the C/OpenMP port's figure beside it

``--impl reference`` times the reference's CPU implementation of the path on all host threads: exactly --warmup + --steps
steps of a FIXED, stated sample of the workload (REF_QUERIES queries per step), through the reference's own
network_sampling_* functions and nn.Modules (cpu_baseline.kind = "reference") when oracle/_ref exists, else through the
C/OpenMP port (kind = "port").
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_DISK = 57_664        # algorithmic FLOPs / query, disk T=4   (SURVEY.md 8d / BASELINE.md 3)
F_SPH = 164_416        # spherical / bsdf T=8
MATERIAL = {"disk": "disk_aniso_brushed_aluminium_1_rgb", "spherical": "spherical_aniso_brushed_aluminium_1_rgb"}


def load_fixture(workload):
    z = np.load(os.path.join(ROOT, "tests", "golden", MATERIAL[workload] + ".npz"))
    layers = [z[f"flow_w{i}"] for i in range(int(z["n_flow_layers"]))]
    base = (z["base_w1"], z["base_b1"], z["base_wo"], z["base_bo"])
    return layers, base


def synth_wi3(workload, n_side, seed):
    """Synthetic incident directions [n,3] (local frame): disk = the reference's stratified concentric-disk
    generator scaled to r<0.95 (utils_sampling_torch_disk.py:99-114); spherical = stratified (theta,phi)
    with theta in (0, pi/2) (spherical_domain_sampling.py:173-175)."""
    rng = np.random.default_rng(seed)
    n = n_side * n_side
    i, j = np.divmod(np.arange(n), n_side)
    u = (i + rng.random(n)) / n_side
    v = (j + rng.random(n)) / n_side
    if workload == "disk":
        a, b = 2 * u - 1, 2 * v - 1
        with np.errstate(divide="ignore", invalid="ignore"):
            big = np.abs(a) > np.abs(b)
            r = np.where(big, a, b)
            phi = np.where(big, (np.pi / 4) * (b / a), np.pi / 2 - (np.pi / 4) * (a / b))
        phi = np.nan_to_num(phi)
        x, y = 0.95 * r * np.cos(phi), 0.95 * r * np.sin(phi)
        zc = np.sqrt(np.maximum(1 - x * x - y * y, 0))
    else:
        th, ph = u * (np.pi / 2) * 0.98 + 0.01, v * 2 * np.pi - np.pi
        x, y, zc = np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)
    return np.stack([x, y, zc], 1).astype(np.float32)


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                 "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["bf16_tflops"]), "measured burst bf16 (MEASURED_PEAKS.json)"
    return 1590.0, "fallback (B200_PROFILING.md)"


REF_QUERIES = 262_144          # queries per step of the CPU arms (fixed and stated; a 16.7 M-query step would take ~40 s)


def domain_wi(workload, wi3):
    """plugin-frame wi [n,3] -> the [n,2] conditioning the reference's sampler functions take
    (brdf_measured_disk.py:66-67 / brdf_measured_spherical.py:35-39,76-77)."""
    if workload == "disk":
        return np.ascontiguousarray(wi3[:, :2])
    r = np.sqrt((wi3 * wi3).sum(1))
    return np.stack([np.arccos(wi3[:, 2] / (r + 1e-8)), np.arctan2(wi3[:, 1], wi3[:, 0])], 1).astype(np.float32)


class CpuArm:
    """The reference's CPU implementation of the path, fixed sample per step.  kind "reference" = the reference's own
    PyTorch functions/modules/checkpoints from oracle/_ref; kind "port" = oracle/bsdf_oracle.c (C/OpenMP)."""

    def __init__(self, workload, mode, prefer_reference=True, n=REF_QUERIES):
        self.workload, self.mode = workload, mode
        self.T = 4 if workload == "disk" else 8
        side = int(round(np.sqrt(n)))
        self.n = side * side
        self.wi3 = synth_wi3(workload, side, 1)
        self.threads = os.cpu_count() or 1
        os.environ["OMP_NUM_THREADS"] = str(self.threads)       # torchrun pins it to 1 for every rank
        self.kind = "port"
        if prefer_reference:
            try:
                from oracle import make_ref
                fns = make_ref.load(workload)
            except Exception:                                    # noqa: BLE001
                fns = None
            if fns is not None:
                import torch
                torch.set_num_threads(self.threads)
                self.torch = torch
                self.fs, self.fp = fns
                self.wi2 = torch.from_numpy(domain_wi(workload, self.wi3))
                self.kind = "reference"
                torch.manual_seed(0)
                self.wo = self.fs(self.wi2)[0].detach() if mode == "pdf" else None
        if self.kind == "port":
            from oracle import bsdf_oracle as O
            from oracle import c_oracle as C
            self.C = C
            self.flow, self.base, _ = O.load_material_npz(os.path.join(ROOT, "tests", "golden", MATERIAL[workload] + ".npz"))
            wi2 = domain_wi(workload, self.wi3)
            self.x0 = (O.draw_x0_disk if workload == "disk" else O.draw_x0_spherical)(self.base, wi2, np.random.default_rng(0))
            self.epi = 1 if workload == "disk" else 2
            self.threads = C.num_threads()
            self.wo3 = C.sample(self.flow, self.base, self.wi3, self.T, self.x0, epilogue=self.epi)[0] if mode == "pdf" else None

    def step(self):
        t = time.perf_counter()
        if self.kind == "reference":
            if self.mode == "sample":
                self.fs(self.wi2)
            else:
                self.fp(self.wo, self.wi2)
        elif self.mode == "sample":
            self.C.sample(self.flow, self.base, self.wi3, self.T, self.x0, epilogue=self.epi)
        else:
            self.C.pdf(self.flow, self.base, self.wo3, self.wi3, self.T, epilogue=self.epi)
        return time.perf_counter() - t

    def run(self, steps, warmup):
        for _ in range(warmup):
            self.step()
        dts = [self.step() for _ in range(steps)]
        return sum(dts)

    def describe(self, steps, total_s):
        what = ("the reference's network_%s_%s + nn.Modules + checkpoint (oracle/_ref), torch %d threads"
                % ("sampling" if self.mode == "sample" else "pdf", self.workload, self.threads)) if self.kind == "reference" \
            else "C/OpenMP port of the same algorithm (oracle/bsdf_oracle.c), %d threads" % self.threads
        return (f"{steps} steps x {self.n} queries (stratified wi, material aniso_brushed_aluminium_1_rgb, T={self.T}, "
                f"mode {self.mode}), {total_s:.1f} s, {what}")


def ncu_traffic(workload, mode):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the newest committed `ncu --set full`
    summary of this configuration under profiles/, or (None, None)."""
    import glob
    import re
    tag = {"sample": "", "pdf": "_pdf"}[mode]
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_ncu_tc16_{workload}{tag}_summary.txt")))
    for f in reversed(files):
        txt = open(f).read()
        m = re.findall(r"dram__bytes_(?:read|write)\.sum = ([0-9.]+) (\w+)", txt)
        if len(m) >= 2:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            return sum(float(v) * scale.get(u, 1.0) for v, u in m[:2]), os.path.relpath(f, ROOT)
    return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="disk", choices=["disk", "spherical"])
    ap.add_argument("--mode", default="sample", choices=["sample", "pdf"])
    ap.add_argument("--queries", type=int, default=4096 * 4096, help="queries per GPU per step")
    ap.add_argument("--precision", default=os.environ.get("BSDFDIFF_PRECISION", "tc16"), choices=["tc16", "fp32"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--cpu-port", action="store_true", help="reference arm / cpu_baseline: force the C/OpenMP port")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    T = 4 if args.workload == "disk" else 8
    F = F_DISK if args.workload == "disk" else F_SPH
    n_side = int(round(np.sqrt(args.queries)))
    n = n_side * n_side
    what = "fused sample+pdf" if args.mode == "sample" else "pdf() of given (wi, wo)"
    config = {"workload": f"measured BRDF {args.workload}-domain sampler, {what}, "
                          f"{n} (wi, Philox noise) queries per GPU, T={T}, material aniso_brushed_aluminium_1_rgb",
              "queries_per_gpu": n, "T": T, "mode": args.mode, "precision": args.precision,
              "l2": f"inputs+outputs {n * 28 / 1e6:.0f} MB per step > 126 MB L2 (no reuse between steps)",
              "sharding": f"dp{world}: contiguous row blocks, no data-path collective"}

    if args.impl == "reference":
        if rank != 0:
            return
        arm = CpuArm(args.workload, args.mode, prefer_reference=not args.cpu_port)
        total = arm.run(args.steps, args.warmup)
        v = arm.n * args.steps / total
        config["reference_sample"] = f"{arm.n} queries per step (fixed), {args.steps} timed steps after {args.warmup} warm-up steps"
        base = {"value": v, "unit": "samples/s", "cores": arm.threads, "kind": arm.kind,
                "sample": arm.describe(args.steps, total)}
        print(json.dumps({"impl": "reference", "metric": "BSDF samples/sec (sample+pdf)", "value": v,
                          "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * total / args.steps, "queries_per_step": arm.n, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": config, "cpu_baseline": base,
                          "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    import bsdf_diffusion_sampling_b200 as pkg

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = pkg.sharding.bind_to_gpu_numa_node(local_rank)
    config["numa_node"] = numa if numa is not None else "not exposed (single node / VM)"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    fix_thr = pkg.ops.default_fixup_threshold(pkg.ops.DISK if args.workload == "disk" else pkg.ops.SPHERICAL,
                                              pkg.ops.EPI_DISK if args.workload == "disk" else pkg.ops.EPI_SPHERICAL, args.mode)
    fix_on = args.precision == "tc16" and fix_thr > 0
    config["fixup_threshold"] = fix_thr if fix_on else 0.0
    launches_per_step = 2 if fix_on else 1
    first_index = rank * n

    def max_ms(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def time_device(workload, mode, steps, warmup, want_clocks=False, calibrate=False):
        """(ms total, queries per GPU, last output, rows the last fix-up pass recomputed, clocks) for `steps` timed steps.
        calibrate: use the material's OWN fix-up thresholds (what materials.MaterialPack.calibrate stores) instead of
        the flow family's defaults."""
        Tw = 4 if workload == "disk" else 8
        layers, base = load_fixture(workload)
        pf = pkg.weights.pack_flow_layers(layers, dev)
        pb = pkg.weights.pack_base_arrays(*base, dev)
        s = pkg.plugins.NeuralBSDFSampler(workload, pf, pb, T=Tw, precision=args.precision)
        if calibrate:
            s.calibrate_fixup()
        wi_np = synth_wi3(workload, n_side, seed=1000 + rank)
        wi = torch.from_numpy(wi_np).to(dev)
        wo = s.sample(wi, seed=7, first_index=first_index)[0] if mode == "pdf" else None

        def step(k):
            if mode == "sample":
                return s.sample(wi, seed=2024, offset=4 * k, first_index=first_index)
            return s.pdf(wi, wo)

        for k in range(warmup):
            out = step(k)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        clocks = ClockSampler(local_rank) if want_clocks else None
        if clocks:
            clocks.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(steps):
            out = step(warmup + k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if clocks:
            clocks.stop_flag = True
            clocks.join(1.0)
        fixed = pkg.ops.last_fixup_count(dev) if (fix_on and not (calibrate and s.fixup[mode] == 0)) else 0
        return max_ms(ms), s, wi_np, out, fixed, clocks

    ms, sampler, wi_np, out, fixed_rows, clocks = time_device(args.workload, args.mode, args.steps, args.warmup, True)
    value = world * n * args.steps / (ms * 1e-3)
    pdf_out = out[1] if args.mode == "sample" else out
    checksum = float(torch.nan_to_num(pdf_out.double(), nan=0.0, posinf=0.0, neginf=0.0).sum().item())

    # ---- e2e: host buffers through the plugin-level call, and the copy-only ceiling of the same pipeline ----------
    e2e = None
    if not args.no_e2e and args.mode == "sample":
        wi_host = torch.from_numpy(wi_np).pin_memory()
        wo_host = torch.empty((n, 3), dtype=torch.float32).pin_memory()
        pdf_host = torch.empty((n,), dtype=torch.float32).pin_memory()
        ke = max(3, min(args.steps, 10))

        def host_leg(copy_only):
            launches = 0
            for k in range(2):
                sampler.sample_host(wi_host, wo_host, pdf_host, seed=2024, offset=4 * k, first_index=first_index,
                                    copy_only=copy_only)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for k in range(ke):
                launches += sampler.sample_host(wi_host, wo_host, pdf_host, seed=2024, offset=4 * k,
                                                first_index=first_index, copy_only=copy_only)
            t1.record()
            torch.cuda.synchronize()
            return max_ms(t0.elapsed_time(t1)), launches

        ms_c, _ = host_leg(True)
        ms_e, launches_e2e = host_leg(False)
        bytes_step = n * 28
        e2e = {"value": world * n * ke / (ms_e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": n * 12,
               "d2h_bytes_per_step": n * 16, "steps": ke, "ms_per_step": ms_e / ke, "gpu_launches": launches_e2e,
               "api": "plugins.NeuralBSDFSampler.sample_host (pinned host wi -> wo,pdf), 3-stream chunked pipeline, "
                      "slot events persistent across calls",
               "copy_only_value": world * n * ke / (ms_c * 1e-3), "copy_only_ms_per_step": ms_c / ke,
               "copy_only_gbs_per_gpu": bytes_step * ke / (ms_c * 1e-3) / 1e9,
               "frac_of_copy_ceiling": ms_c / ms_e,
               "note": "copy_only = the same chunks, streams and pinned buffers with the kernels skipped: the host<->device "
                       "ceiling of this box at this N; e2e / ceiling says how much the kernels add on top of the copies",
               "checksum_pdf": float(torch.nan_to_num(pdf_host.double(), nan=0.0, posinf=0.0, neginf=0.0).sum().item())}

    # ---- the other BASELINE configs as short lines ---------------------------------------------------------------
    extra = []
    if not args.no_extra and args.precision == "tc16":
        for wl, md in (("spherical", "sample"), ("spherical", "pdf"), ("disk", "pdf"), ("disk", "sample")):
            if (wl, md) == (args.workload, args.mode):
                continue
            ks = max(3, min(args.steps, 5))
            ms_x, _, _, _, fx, _ = time_device(wl, md, ks, 3)
            Fx = F_DISK if wl == "disk" else F_SPH
            qps = n * ks / (ms_x * 1e-3)
            extra.append({"workload": wl, "mode": md, "T": 4 if wl == "disk" else 8, "value": world * qps,
                          "unit": "samples/s", "ms_per_step": ms_x / ks, "steps": ks,
                          "roofline_frac": qps * Fx / 1e12 / measured_peak()[0],
                          "fixup_rows_last_step": fx})
        # the same four calls with the material's own calibrated fix-up thresholds (a MaterialPack user's configuration)
        for wl, md in (("disk", "sample"), ("disk", "pdf"), ("spherical", "sample"), ("spherical", "pdf")):
            ks = max(3, min(args.steps, 5))
            ms_x, s_x, _, _, fx, _ = time_device(wl, md, ks, 3, calibrate=True)
            Fx = F_DISK if wl == "disk" else F_SPH
            qps = n * ks / (ms_x * 1e-3)
            extra.append({"workload": wl, "mode": md, "T": 4 if wl == "disk" else 8, "value": world * qps,
                          "unit": "samples/s", "ms_per_step": ms_x / ks, "steps": ks,
                          "roofline_frac": qps * Fx / 1e12 / measured_peak()[0], "fixup_rows_last_step": fx,
                          "fixup": "per-material calibrated thresholds (NeuralBSDFSampler.calibrate_fixup)",
                          "fixup_thresholds": s_x.fixup})

    # ---- twelve materials in ONE wavefront (disney_bsdf_array0_envmap.xml:35-335), single launch over
    #      the device-built plan vs Mitsuba's per-instance dispatch (boolean-mask gather, one launch per material, scatter)
    # ---- BASELINE configs[3]: matpreview 1920x1080 @ 64 spp through the bsdf_myresult plugin, replayed wavefront ----------
    if not args.no_extra and args.precision == "tc16":
        def timed(fn, steps, warm=3):
            for _ in range(warm):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return max_ms(a.elapsed_time(b)) / steps

        import glob
        mats = []
        disk_files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "disk_*.npz")))
        for j in range(12):                                   # 12 table entries: the three disk fixtures, own blobs each
            z = np.load(disk_files[j % len(disk_files)])
            mats.append(pkg.plugins.NeuralBSDFSampler(
                "disk", pkg.weights.pack_flow_layers([z[f"flow_w{i}"] for i in range(int(z["n_flow_layers"]))], dev),
                pkg.weights.pack_base_arrays(z["base_w1"], z["base_b1"], z["base_wo"], z["base_bo"], dev),
                precision=args.precision))
        mm = pkg.plugins.MultiMaterialSampler(mats)
        wi_m = torch.from_numpy(synth_wi3("disk", n_side, seed=77 + rank)).to(dev)
        mid = torch.randint(0, 12, (n,), device=dev, dtype=torch.int32)
        ms_plan = timed(lambda: mm.plan(mid), 5)
        plan = mm.plan(mid)
        ms_multi = timed(lambda: mm.sample(wi_m, plan=plan, seed=5, first_index=first_index), 5)
        ms_both = timed(lambda: mm.sample(wi_m, mid, seed=5, first_index=first_index), 5)
        ms_inst = timed(lambda: mm.sample_per_material(wi_m, mid, seed=5, first_index=first_index), 3, warm=1)
        extra.append({"workload": "disk, 12 materials in one wavefront (random material id per row)", "mode": "sample", "T": 4,
                      "value": world * n / (ms_both * 1e-3), "unit": "samples/s", "ms_per_step": ms_both,
                      "ms_plan": ms_plan, "ms_kernels_given_plan": ms_multi,
                      "ms_per_instance_dispatch": ms_inst, "roofline_frac": n / (ms_both * 1e-3) * F_DISK / 1e12 / measured_peak()[0],
                      "note": "single launch over a device-built plan (bsdfdiff_multi_plan + bsdfdiff_sample_multi) vs the "
                              "per-instance dispatch Mitsuba performs (mask gather + one launch per material + scatter)"})
        del mm, mats, wi_m, mid, plan

        zb = np.load(os.path.join(ROOT, "tests", "golden", "bsdf_0.npz"))
        sb = pkg.plugins.NeuralBSDFSampler(
            "bsdf", pkg.weights.pack_flow_layers([zb[f"flow_w{i}"] for i in range(int(zb["n_flow_layers"]))], dev),
            pkg.weights.pack_base_arrays(zb["base_w1"], zb["base_b1"], zb["base_wo"], zb["base_bo"], dev),
            precision=args.precision)
        lanes, spp = 1920 * 1080, 64
        g = torch.Generator(device=dev).manual_seed(3 + rank)
        v = torch.randn((lanes, 3), device=dev, generator=g)
        wi_f = v / v.norm(dim=1, keepdim=True)                # bsdf_myresult takes both hemispheres (bsdf_myresult.py:55-57)
        wo_l = torch.randn((lanes, 3), device=dev, generator=g)
        wo_l = wo_l / wo_l.norm(dim=1, keepdim=True)          # emitter-sampling direction of the MIS pdf() query

        def one_pass(k=[0]):
            k[0] += 1
            sb.sample(wi_f, seed=9, offset=68 * k[0], first_index=rank * lanes)
            sb.pdf(wi_f, wo_l)
        ms_pass = timed(one_pass, 16)
        extra.append({"workload": "matpreview replay: bsdf_myresult plugin (bsdf_0), 1920x1080 lanes per pass, full-sphere wi, "
                                  "sample() + pdf() per pass, 64 spp = 64 passes per frame", "mode": "sample+pdf", "T": 8,
                      "value": world * 2 * lanes / (ms_pass * 1e-3), "unit": "queries/s", "ms_per_pass": ms_pass,
                      "passes_per_s": 1e3 / ms_pass, "ms_per_64spp_frame_bounce": ms_pass * spp,
                      "roofline_frac": 2 * lanes / (ms_pass * 1e-3) * F_SPH / 1e12 / measured_peak()[0],
                      "note": "synthetic wavefront (Mitsuba is absent): one bounce of every pixel's path per pass; the "
                              "reference runs the same two calls as ~700 eager launches + 4 Dr.Jit<->torch crossings per pass"})

        # ---- one diffusion / rectify-stage iteration at the reference's batch (4.9 M rows), fused launch ----
        n_tr = 4_900_000
        gt = torch.Generator(device=dev).manual_seed(11 + rank)
        om_i = torch.rand(n_tr, 2, device=dev, generator=gt) * 1.2 - 0.6
        om_o = torch.rand(n_tr, 2, device=dev, generator=gt) * 1.6 - 0.8
        x_b = om_i * 0.5 + 0.3 * torch.randn(n_tr, 2, device=dev, generator=gt)
        lay, _ = load_fixture("disk")
        trainer = pkg.training.FlowMatchingTrainer(lay, lr=1e-3, device=dev)
        ms_tr = timed(lambda: trainer.step(x_b, om_o, om_i), 5, warm=2)
        extra.append({"workload": "training step: diffusion / rectify stage body (forward + backward + Adam, one launch), "
                                  "4.9 M rows, 32-wide disk flow net, fp32", "mode": "train", "value": world * n_tr / (ms_tr * 1e-3),
                      "unit": "rows/s", "ms_per_step": ms_tr,
                      "algorithmic_tflops": n_tr * 3 * 2 * (25 * 32 + 2 * 1024 + 64) / (ms_tr * 1e-3) / 1e12,
                      "note": "CUDA-core kernel (train.cu); the reference's eager PyTorch step on the same GPU: "
                              "profiles/r2i_train_bench.jsonl"})
        del trainer, om_i, om_o, x_b

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    per_gpu_qps = n * args.steps / (ms * 1e-3)
    achieved = per_gpu_qps * F / 1e12
    traffic, traffic_src = ncu_traffic(args.workload, args.mode) if (n == 4096 * 4096 and args.precision == "tc16") else (None, None)
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel": ("flow_tc_kernel (+ flow_simt_kernel fix-up pass)" if fix_on else "flow_tc_kernel")
                if args.precision == "tc16" else "flow_simt_kernel",
                "flops_per_query": F, "avg_launch_ms": ms / args.steps,
                "fixup_rows_last_step": fixed_rows, "fixup_share": fixed_rows / n,
                "hbm_algorithmic_bytes_per_query": 28, "hbm_achieved_gbs": per_gpu_qps * 28 / 1e9,
                "mufu_ceiling_frac": 0.43 if args.workload == "disk" else 0.46,
                "note": "algorithmic FLOPs (unpadded layer shapes, value + 2 tangent columns) x queries per step / "
                        "CUDA-event time per step (tensor-core launch + fix-up launch); per GPU. mufu_ceiling_frac = the "
                        "fraction of the tensor roofline at which the 16 tanh/clk/SM MUFU rate saturates (one tanh per "
                        "activation; DESIGN.md 4.1)"}
    line = {"metric": "BSDF samples/sec (sample+pdf)", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands / f32 accumulate+state" if args.precision == "tc16" else "f32",
            "data": "synthetic", "config": config, "roofline": roofline, "clocks": clocks.summary(),
            "gpu_launches": args.steps * launches_per_step, "checksum_pdf": checksum}
    if e2e:
        line["e2e"] = e2e
    if extra:
        line["extra"] = extra
    if not args.no_cpu and world == 1:
        try:
            arm = CpuArm(args.workload, args.mode, prefer_reference=not args.cpu_port)
            ks = 12 if arm.kind == "reference" else 6
            total = arm.run(ks, 1)
            line["cpu_baseline"] = {"value": arm.n * ks / total, "unit": "samples/s", "cores": arm.threads,
                                    "kind": arm.kind, "sample": arm.describe(ks, total)}
            if arm.kind == "reference":
                port = CpuArm(args.workload, args.mode, prefer_reference=False)
                tp = port.run(4, 1)
                line["cpu_baseline"]["port_value"] = port.n * 4 / tp
                line["cpu_baseline"]["port_note"] = "C/OpenMP port of the same algorithm (oracle/bsdf_oracle.c) on the same sample"
        except Exception as ex:                                      # noqa: BLE001
            line["cpu_baseline"] = {"error": repr(ex)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on
    # stdout when NCCL_DEBUG is VERSION/WARN), so file descriptor 1 is pointed at stderr for the whole run and the
    # result line goes to the saved descriptor.
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    _emit = []
    _print = print

    def print(*a, **k):                                  # noqa: A001  (result lines only; everything else -> stderr)
        _emit.append(" ".join(str(x) for x in a))

    try:
        main()
    finally:
        sys.stdout.flush()
        for _l in _emit:
            os.write(_real_stdout, (_l + "\n").encode())
