import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
import bsdf_diffusion_sampling_b200 as pkg
layers, base = bench.load_fixture("disk")
pf = pkg.weights.pack_flow_layers(layers, "cuda"); pb = pkg.weights.pack_base_arrays(*base, "cuda")
s = pkg.plugins.NeuralBSDFSampler("disk", pf, pb, T=4, precision="tc16")
wi_np = bench.synth_wi3("disk", 4096, 1)
wi = torch.from_numpy(wi_np).cuda()
def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for nq in (1 << 16, 1 << 19, 1 << 21, 1 << 24):
    w = wi[:nq]
    print("device-resident n=%d: %.3f ms" % (nq, timeit(lambda: s.sample(w, seed=1))))
n = wi.shape[0]
wi_host = torch.from_numpy(wi_np).pin_memory()
wo_host = torch.empty((n, 3), dtype=torch.float32).pin_memory()
pdf_host = torch.empty((n,), dtype=torch.float32).pin_memory()
print("pinned:", wi_host.is_pinned(), wo_host.is_pinned())
for chunk in (1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22):
    if hasattr(s, "_host_pipe"): del s._host_pipe
    print("sample_host chunk=%d: %.3f ms" % (chunk, timeit(lambda: s.sample_host(wi_host, wo_host, pdf_host, seed=1, chunk=chunk), 3)))
d = torch.empty_like(wi)
print("H2D 201MB: %.3f ms" % timeit(lambda: d.copy_(wi_host, non_blocking=True)))
wo_d = torch.empty((n, 3), device="cuda")
print("D2H 201MB: %.3f ms" % timeit(lambda: wo_host.copy_(wo_d, non_blocking=True)))
