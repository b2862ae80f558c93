#!/usr/bin/env python
"""bench.py -- headline benchmark of the fused neural-BSDF sampler (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload disk|spherical] [--queries Q] [--precision tc16|fp32]

A "step" is one pass of the hot path (fused sample + pdf) over one batch of synthetic queries:
N=1 workload = BASELINE.json configs[1], "measured BRDF disk-domain sampler: 16M (wi, noise) queries,
full diffusion steps (T=4, the plugin's setting), 1 GPU"; for N>1 every rank gets its own 16M-query shard
(weak scaling, no collective on the data path; Philox counters are global row indices).
Prints ONE JSON line (rank 0).  ``value`` = whole-job samples/s with inputs resident in HBM;
``e2e`` = the same metric through the plugin-level host-buffer call (pinned host wi in, wo+pdf out,
copies inside the timed region); ``roofline`` = achieved algorithmic TFLOP/s of the fused kernel vs
the measured dense-fp16/bf16 tensor peak; ``cpu_baseline`` = the CPU oracle port on the host cores.

``--impl reference`` times the reference's CPU implementation of the path.  The reference is
Python/PyTorch and cannot travel to the GPU box, so this arm runs the oracle's C/OpenMP port
(oracle/bsdf_oracle.c, pinned against the reference's own outputs) on all host threads.
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


def cpu_leg(workload, T, n_target_s=12.0):
    """Time the CPU oracle port (C + OpenMP, all host threads) on a bounded sample of the same workload."""
    from oracle import bsdf_oracle as O
    from oracle import c_oracle as C
    flow, base, _ = O.load_material_npz(os.path.join(ROOT, "tests", "golden", MATERIAL[workload] + ".npz"))
    epi = 1 if workload == "disk" else 2
    rng = np.random.default_rng(0)

    def run(n_side):
        wi3 = synth_wi3(workload, n_side, 1)
        wi2 = wi3[:, :2] if workload == "disk" else O.cart_to_spher(wi3)
        x0 = (O.draw_x0_disk if workload == "disk" else O.draw_x0_spherical)(base, wi2, rng)
        t = time.perf_counter()
        C.sample(flow, base, wi3, T, x0, epilogue=epi)
        return wi3.shape[0], time.perf_counter() - t

    n, dt = run(256)                                    # 65 536 queries: calibrate
    rate = n / dt
    side = int(min(4096, max(256, np.sqrt(rate * n_target_s))))
    n, dt = run(side)
    return {"value": n / dt, "unit": "samples/s", "cores": C.num_threads(), "kind": "port",
            "sample": f"{n} queries ({side}x{side} stratified wi, same material/T), {dt:.1f} s, C/OpenMP oracle"}, n, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="disk", choices=["disk", "spherical"])
    ap.add_argument("--queries", type=int, default=4096 * 4096, help="queries per GPU per step")
    ap.add_argument("--precision", default=os.environ.get("BSDFDIFF_PRECISION", "tc16"), choices=["tc16", "fp32"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    T = 4 if args.workload == "disk" else 8
    F = F_DISK if args.workload == "disk" else F_SPH
    n_side = int(round(np.sqrt(args.queries)))
    n = n_side * n_side
    config = {"workload": f"measured BRDF {args.workload}-domain sampler, fused sample+pdf, "
                          f"{n} (wi, Philox noise) queries per GPU, T={T}, material aniso_brushed_aluminium_1_rgb",
              "queries_per_gpu": n, "T": T, "precision": args.precision,
              "l2": f"inputs+outputs {n * 28 / 1e6:.0f} MB per step > 126 MB L2 (no reuse between steps)",
              "sharding": f"dp{world}: contiguous row blocks, no data-path collective"}

    if args.impl == "reference":
        if rank != 0:
            return
        # torchrun exports OMP_NUM_THREADS=1 for every rank; the reference arm runs on rank 0 alone and is meant
        # to use every host thread (the OpenMP runtime reads the variable when the oracle library is first loaded)
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
        vals = []
        for _ in range(args.warmup and 1):
            cpu_leg(args.workload, T, 2.0)
        last = None
        t_tot, n_tot = 0.0, 0
        for _ in range(max(1, min(args.steps, 3))):
            last, nq, dt = cpu_leg(args.workload, T, 8.0)
            t_tot += dt
            n_tot += nq
        v = n_tot / t_tot
        last["value"] = v
        print(json.dumps({"impl": "reference", "metric": "BSDF samples/sec (sample+pdf)", "value": v,
                          "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * t_tot / max(1, min(args.steps, 3)), "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": config, "cpu_baseline": last,
                          "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "note": "reference = Python/PyTorch, cannot travel to the GPU box; this arm times the "
                                  "oracle's C/OpenMP port of the same algorithm on all host threads"}))
        return

    import torch
    import torch.distributed as dist
    import bsdf_diffusion_sampling_b200 as pkg

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = pkg.sharding.bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    layers, base = load_fixture(args.workload)
    pf = pkg.weights.pack_flow_layers(layers, dev)
    pb = pkg.weights.pack_base_arrays(*base, dev)
    sampler = pkg.plugins.NeuralBSDFSampler(args.workload, pf, pb, T=T, precision=args.precision)
    wi_np = synth_wi3(args.workload, n_side, seed=1000 + rank)
    wi = torch.from_numpy(wi_np).to(dev)
    first_index = rank * n

    def step(k):
        return sampler.sample(wi, seed=2024, offset=4 * k, first_index=first_index)

    for k in range(args.warmup):
        out = step(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local_rank)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        out = step(args.warmup + k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    clocks.stop_flag = True
    clocks.join(1.0)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * n * args.steps / (ms * 1e-3)
    checksum = float(out[1].double().sum().item())

    # ---- e2e: host buffers through the plugin-level call ---------------------------------------
    e2e = None
    if not args.no_e2e:
        wi_host = torch.from_numpy(wi_np).pin_memory()
        wo_host = torch.empty((n, 3), dtype=torch.float32).pin_memory()
        pdf_host = torch.empty((n,), dtype=torch.float32).pin_memory()
        launches_e2e = 0
        for k in range(2):
            sampler.sample_host(wi_host, wo_host, pdf_host, seed=2024, offset=4 * k, first_index=first_index)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ke = max(3, min(args.steps, 10))
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for k in range(ke):
            launches_e2e += sampler.sample_host(wi_host, wo_host, pdf_host, seed=2024, offset=4 * k,
                                                first_index=first_index)
        t1.record()
        torch.cuda.synchronize()
        ms_e = t0.elapsed_time(t1)
        if world > 1:
            t = torch.tensor([ms_e], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e = float(t.item())
        e2e = {"value": world * n * ke / (ms_e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": n * 12,
               "d2h_bytes_per_step": n * 16, "steps": ke, "ms_per_step": ms_e / ke,
               "api": "plugins.NeuralBSDFSampler.sample_host (pinned host wi -> wo,pdf), 3-stream chunked pipeline",
               "checksum_pdf": float(pdf_host.double().sum().item())}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    per_gpu_qps = n * args.steps / (ms * 1e-3)
    achieved = per_gpu_qps * F / 1e12
    # DRAM traffic of ONE launch from the committed ncu --set full capture of this exact configuration
    # (dram__bytes_read.sum + dram__bytes_write.sum, profiles/r1n_ncu_tc16_disk_summary.txt); other configs: null
    traffic = 423.12e6 if (args.workload == "disk" and n == 4096 * 4096 and args.precision == "tc16") else None
    traffic_src = "profiles/r1n_ncu_tc16_disk_summary.txt (bytes per launch)"
    if args.workload == "spherical" and n == 4096 * 4096 and args.precision == "tc16":
        traffic, traffic_src = 422.80e6, "profiles/r1n_ncu_tc16_spherical_summary.txt (bytes per launch)"
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src if traffic else None, "peak_source": peak_src,
                "kernel": "flow_tc_kernel" if args.precision == "tc16" else "flow_simt_kernel",
                "flops_per_query": F, "avg_launch_ms": ms / args.steps,
                "hbm_algorithmic_bytes_per_query": 28, "hbm_achieved_gbs": per_gpu_qps * 28 / 1e9,
                "mufu_ceiling_frac": 0.43 if args.workload == "disk" else 0.46,
                "note": "algorithmic FLOPs (unpadded layer shapes, value + 2 tangent columns) x queries per launch "
                        "/ CUDA-event time per launch; per GPU. mufu_ceiling_frac = the fraction of the tensor roofline at "
                        "which the 16 tanh/clk/SM MUFU rate saturates (one tanh per activation; DESIGN.md 4.1)"}
    line = {"metric": "BSDF samples/sec (sample+pdf)", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands / f32 accumulate+state" if args.precision == "tc16" else "f32",
            "data": "synthetic", "config": config, "roofline": roofline, "clocks": clocks.summary(),
            "gpu_launches": args.steps * (2 if (args.precision == "tc16" and pkg.ops.get_fixup_threshold() > 0) else 1),
            "checksum_pdf": checksum}
    if e2e:
        line["e2e"] = e2e
    if numa is not None:
        line["config"]["numa_node_rank0"] = numa
    if not args.no_cpu and world == 1:
        try:
            line["cpu_baseline"] = cpu_leg(args.workload, T)[0]
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
