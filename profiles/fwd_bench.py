import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
import bsdf_diffusion_sampling_b200 as pkg
layers, base = bench.load_fixture("disk")
pf = pkg.weights.pack_flow_layers(layers, "cuda"); pb = pkg.weights.pack_base_arrays(*base, "cuda")
wi = torch.from_numpy(bench.synth_wi3("disk", 4096, 1)[:, :2].copy()).cuda()
x0 = torch.randn(wi.shape[0], 2, device="cuda") * 0.1
for T in (4,):
    for _ in range(3): pkg.ops.flow_forward(wi, pf, T, x0=x0, precision="tc16")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): pkg.ops.flow_forward(wi, pf, T, x0=x0, precision="tc16")
    e1.record(); torch.cuda.synchronize()
    print("forward-only T=%d: %.3f ms per 16.7M queries" % (T, e0.elapsed_time(e1) / 5))
