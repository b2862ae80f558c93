#!/usr/bin/env python
"""BASELINE.json config 5: query-count x Euler-step sweep of the fused sample+pdf kernel (device-resident, CUDA
events, best of 3 after a warm-up), reported as queries/s and as a fraction of the tensor roofline with the
algorithmic FLOPs of SURVEY 8(d) scaled to T.  One GPU per process; under torchrun every rank runs the same sweep
on its own shard (weak scaling, no collective) and rank 0 prints the max-over-ranks time.

    python profiles/sweep.py [--workload disk|spherical] [--mq 1,4,16,64] [--T 4,8,32,128,256]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                    # noqa: E402
import bsdf_diffusion_sampling_b200 as pkg      # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="disk", choices=["disk", "spherical"])
    ap.add_argument("--mq", "--n", dest="n", default="1,4,16,64", help="millions of queries per GPU (--mq under torchrun: its parser claims --n)")
    ap.add_argument("--T", default="4,8,32,128,256")
    ap.add_argument("--budget-gflop", type=float, default=6.0e5, help="skip cells above this many GFLOP per launch")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    layers, base = bench.load_fixture(args.workload)
    pf = pkg.weights.pack_flow_layers(layers, dev)
    pb = pkg.weights.pack_base_arrays(*base, dev)
    peak, _ = bench.measured_peak()
    per_step = (bench.F_DISK - 576) / 4 if args.workload == "disk" else (bench.F_SPH - 576) / 8
    rows = []
    for nm in [int(v) for v in args.n.split(",")]:
        n_side = int(round(np.sqrt(nm * (1 << 20))))
        wi = torch.from_numpy(bench.synth_wi3(args.workload, n_side, seed=1000 + rank)).to(dev)
        n = wi.shape[0]
        for T in [int(v) for v in args.T.split(",")]:
            F = per_step * T + 576
            if n * F / 1e9 > args.budget_gflop:
                continue
            s = pkg.plugins.NeuralBSDFSampler(args.workload, pf, pb, T=T, precision="tc16")
            s.sample(wi, seed=1, offset=0, first_index=rank * n)
            torch.cuda.synchronize()
            best = 1e30
            for k in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                s.sample(wi, seed=1, offset=4 * k, first_index=rank * n)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            if world > 1:
                t = torch.tensor([best], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                best = float(t.item())
            qps = world * n / (best * 1e-3)
            rows.append({"workload": args.workload, "n_per_gpu": n, "T": T, "n_gpus": world, "ms": best,
                         "queries_per_s": qps, "tensor_roofline_frac": qps / world * F / (peak * 1e12)})
            if rank == 0:
                print(json.dumps(rows[-1]), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
