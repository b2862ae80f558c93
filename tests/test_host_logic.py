"""CPU tests: the C-ABI library loads and exports every symbol include/bsdfdiff.h declares, the host-side
packer produces the documented layouts, argument validation / error behaviour, and the sharding logic
(world_size-2 gloo).  No compute call is made without a GPU."""
import ctypes
import os
import re
import socket

import numpy as np
import pytest
import torch

from conftest import BSDF_FILE, DISK_FILE, GOLDEN_FILES, ROOT
from oracle import bsdf_oracle as O


def test_every_declared_symbol_is_exported(built_lib):
    hdr = open(os.path.join(ROOT, "include", "bsdfdiff.h")).read()
    declared = set(re.findall(r"\b(bsdfdiff_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 11
    lib = ctypes.CDLL(built_lib._lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/bsdfdiff.h but not exported"
    assert set(built_lib._lib.EXPORTS) == declared
    assert built_lib._lib.lib.bsdfdiff_abi_version() == 3
    assert built_lib._lib.lib.bsdfdiff_error_string(-2).decode().startswith("shape not supported")


def _unpack_header(blob):
    import struct
    magic, in_dim, H, nh, dom, off32, n32, off16, n16, total = struct.unpack("<I4i5I", blob[:40])
    assert blob[40:64] == bytes(24)                                  # reserved words stay zero
    return dict(magic=magic, in_dim=in_dim, H=H, nh=nh, dom=dom, off32=off32, n32=n32, off16=off16, n16=n16,
                total=total)


@pytest.mark.parametrize("path", [DISK_FILE, BSDF_FILE])
def test_pack_flow_layout(built_lib, path):
    """fp32 image is the per-layer transpose; fp16 image is the UMMA K-major core-matrix layout with the
    hidden layers pre-scaled by 0.5 and the output layer padded to 16 rows."""
    flow, base, z = O.load_material_npz(path)
    packed = built_lib.weights.pack_flow_layers(flow.layers, device="cpu")
    blob = packed.blob.numpy().tobytes()
    h = _unpack_header(blob)
    H, in_dim, nh = flow.layers[0].shape[0], flow.in_dim, len(flow.layers) - 1
    assert (h["magic"], h["in_dim"], h["H"], h["nh"]) == (0xB5DFD1F0, in_dim, H, nh)
    assert h["total"] == len(blob) == built_lib._lib.lib.bsdfdiff_packed_flow_bytes(in_dim, H, nh)
    f32 = np.frombuffer(blob, np.float32, h["n32"] // 4, h["off32"])
    off = 0
    for w in flow.layers:
        assert np.array_equal(f32[off: off + w.size].reshape(w.shape[1], w.shape[0]), w.T)
        off += w.size
    f16 = np.frombuffer(blob, np.float16, h["n16"] // 2, h["off16"])

    def at(img, n, k, N):
        return img[((k // 8) * (N // 8) + n // 8) * 64 + (n % 8) * 8 + k % 8]

    def hi_lo(w):
        hi = np.float16(w)
        return hi, np.float16(np.float32(w) - np.float32(hi))

    w1 = flow.layers[0]
    img_hi, img_lo = f16[: H * 32], f16[H * 32: 2 * H * 32]
    # layer-1 K order: 8 state slots (hi parts then lo parts), PE5(wi) at k = 8..29, two zero columns
    state = [0, 1, 2, 2, 0, 1, None, None] if in_dim == 25 else [0, 1, 2, 3, 0, 1, 2, 3]
    first_pe = 3 if in_dim == 25 else 4
    for n in (0, 7, 13, H - 1):
        for k in range(32):
            src = state[k] if k < 8 else (first_pe + k - 8 if k < 30 else None)
            want = hi_lo(np.float32(0.5) * w1[n, src]) if src is not None else (np.float16(0), np.float16(0))
            assert (at(img_hi, n, k, H), at(img_lo, n, k, H)) == want
    l2 = f16[2 * H * 32: 2 * H * 32 + 2 * H * H]
    assert (at(l2[: H * H], 5, 9, H), at(l2[H * H:], 5, 9, H)) == hi_lo(np.float32(0.5) * flow.layers[1][5, 9])
    out = f16[2 * H * 32 + 2 * (nh - 1) * H * H:]
    assert (at(out[: 16 * H], 1, 17, 16), at(out[16 * H:], 1, 17, 16)) == hi_lo(flow.layers[-1][1, 17])
    assert at(out, 2, 3, 16) == 0 and out.size == 2 * 16 * H
    assert h["off16"] + -(-h["n16"] // 128) * 128 == h["total"]       # nothing follows the fp16 image


def test_pack_flow_tcnn_equals_pack_from_layers(built_lib):
    flow, _, _ = O.load_material_npz(DISK_FILE)
    layers16 = [w.astype(np.float16).astype(np.float32) for w in flow.layers]
    a = built_lib.weights.pack_flow_layers(layers16, device="cpu").blob
    params = O.pack_tcnn_params(flow.layers, flow.in_dim, 2).astype(np.float32)
    b = built_lib.weights.pack_flow_tcnn(params, flow.in_dim, 2, 32, len(flow.layers) - 1, device="cpu").blob
    assert torch.equal(a, b)


def test_argument_validation_without_gpu(built_lib):
    L = built_lib._lib
    assert L.lib.bsdfdiff_packed_flow_bytes(25, 48, 3) == 0          # hidden must be 32 | 64
    assert L.lib.bsdfdiff_packed_flow_bytes(40, 32, 3) == 0          # in_dim <= 32
    with pytest.raises(L.BsdfDiffError):
        built_lib.weights.pack_flow_layers([np.zeros((48, 25), np.float32), np.zeros((2, 48), np.float32)], "cpu")
    with pytest.raises(L.BsdfDiffError):
        built_lib.weights.pack_base_arrays(np.zeros((16, 14)), np.zeros(16), np.zeros((4, 16)), np.zeros(3), "cpu")
    # invalid arguments are rejected before any CUDA call
    assert L.lib.bsdfdiff_sample(0, 0, 0, 4, 16, None, None, 32, 3, None, None, None, 0, 0, 0, None, None, None, 0.0, None,
                                 None) == -1
    assert L.lib.bsdfdiff_sample(0, 0, 2, 4, 16, 1, 1, 32, 3, 1, None, None, 0, 0, 0, 1, 1, None, 0.0, None,
                                 None) == -1                          # spherical epilogue on a disk net
    assert L.lib.bsdfdiff_pdf(0, 1, 0, -1, 16, 1, 1, 1, 32, 4, 1, 1, 0.0, None, None) == -1
    # the fix-up needs its scratch buffer
    assert L.lib.bsdfdiff_sample(1, 0, 0, 4, 16, 1, 1, 32, 3, 1, None, None, 0, 0, 0, 1, 1, 1, 0.25, None, None) == -1
    assert L.lib.bsdfdiff_sample(0, 0, 0, 4, 16, 1, 1, 32, 3, 1, 1, 1, 0, 0, 0, 1, 1, None, 0.0, None, None) == -1   # x0 AND u_noise
    assert L.lib.bsdfdiff_base_log_prob(2, 16, 1, 1, 1, 1, None) == -1
    assert L.lib.bsdfdiff_fixup_scratch_bytes(1000) == 16 + 12000 and L.lib.bsdfdiff_fixup_scratch_bytes(1001) == 16 + 4008 + 8008
    assert L.lib.bsdfdiff_error_string(1).decode().startswith("ok (")
    # CPU tensors raise: there is no CPU path
    flow, base, z = O.load_material_npz(DISK_FILE)
    pf = built_lib.weights.pack_flow_layers(flow.layers, "cpu")
    with pytest.raises(RuntimeError, match="no CPU path"):
        built_lib.ops.sample(torch.zeros(4, 2), pf, torch.zeros(308), 4, x0=torch.zeros(4, 2))
    # generic tcnn-style nets (in_dim not 25 / 26) are mlp_forward-only: PackedFlow.domain refuses to guess
    odd = built_lib.weights.pack_flow_layers([np.zeros((32, 24), np.float32), np.zeros((2, 32), np.float32)], "cpu")
    with pytest.raises(ValueError, match="25 .disk. or 26"):
        odd.domain
    # checkpoint_paths: the measured-spherical plugin loads the *_disk* pretrain checkpoint (brdf_measured_spherical.py:59)
    cp = built_lib.plugins.checkpoint_paths
    assert cp("disk", "m", "r") == ("r/m_disk/brdf_rectify_networkm.pth", "r/m_disk/brdf_pretrain_networkm.pth")
    assert cp("spherical", "m", "r") == ("r/m_spherical/brdf_rectify_networkm.pth", "r/m_disk/brdf_pretrain_networkm.pth")
    assert cp("bsdf", 3, "r") == ("r/bsdf_3_spherical/brdf_rectify_network3.pth", "r/bsdf_3_spherical/brdf_pretrain_network3.pth")
    with pytest.raises(ValueError):
        cp("nope", "m")


def test_multi_material_host_checks_without_gpu(built_lib):
    """MultiMaterialSampler / ops.MaterialTable host logic: shape, kind, T checks and the scratch-size formula
    (no kernel call)."""
    P = built_lib.plugins
    flow, base, _ = O.load_material_npz(DISK_FILE)
    pf = built_lib.weights.pack_flow_layers(flow.layers, "cpu")
    pb = built_lib.weights.pack_base_arrays(base.w1, base.b1, base.wo, base.bo, "cpu")
    a, b = P.NeuralBSDFSampler("disk", pf, pb), P.NeuralBSDFSampler("disk", pf, pb, T=8)
    mm = P.MultiMaterialSampler([a, a])
    assert len(mm.table) == 2 and mm.table.flow_ptrs.tolist() == [pf.blob.data_ptr()] * 2 and mm.T == 4
    with pytest.raises(ValueError, match="share plugin kind, T"):
        P.MultiMaterialSampler([a, b])
    with pytest.raises(ValueError):
        P.MultiMaterialSampler([])
    sflow, sbase, _ = O.load_material_npz(BSDF_FILE)
    sph = P.NeuralBSDFSampler("bsdf", built_lib.weights.pack_flow_layers(sflow.layers, "cpu"),
                              built_lib.weights.pack_base_arrays(sbase.w1, sbase.b1, sbase.wo, sbase.bo, "cpu"))
    with pytest.raises(ValueError, match="share plugin kind"):
        P.MultiMaterialSampler([a, sph])
    with pytest.raises(RuntimeError, match="no CPU path"):
        mm.plan(torch.tensor([0, 1, 0], dtype=torch.int32))
    L = built_lib._lib.lib
    n, M = 1000, 12
    assert L.bsdfdiff_multi_scratch_bytes(n, M) == 4 * 1280 + 16 * (n // 128 + M + 1) + 8 * n + 8 * n
    assert L.bsdfdiff_multi_scratch_bytes(n, 0) == 0 and L.bsdfdiff_multi_scratch_bytes(n, 256) == 0


def test_model_classes_load_reference_state_dict_keys(built_lib):
    m = built_lib.model
    flow, base, z = O.load_material_npz(DISK_FILE)
    net = m.NN_cond_pos_simpler(input_dim=5, output_dim=2, N_NEURONS=32, POSITIONAL_ENCODING_BASIS_NUM=5)
    assert list(net.state_dict().keys()) == ["linear1.weight", "linear2.weight", "linear3.weight", "output.weight"]
    net.load_state_dict({k: torch.from_numpy(w) for k, w in zip(net.state_dict().keys(), flow.layers)})
    sph = m.NN_cond_pos(input_dim=6, output_dim=2, N_NEURONS=32, POSITIONAL_ENCODING_BASIS_NUM=5)
    assert sph.linear1.weight.shape == (32, 26) and hasattr(sph, "linear4") and not hasattr(sph, "linear5")
    cx = m.NN_cond_pos_spherical_complicate(input_dim=6, output_dim=2, N_NEURONS=64, POSITIONAL_ENCODING_BASIS_NUM=5)
    assert cx.linear6.weight.shape == (64, 64)
    b = m.NN_cond_pretrain_disk_one(input_dim=2, N_NEURONS=16, POSITIONAL_ENCODING_BASIS_NUM=3)
    assert b.linear1.weight.shape == (16, 14) and b.output.weight.shape == (4, 16)
    packed = built_lib.weights.packed_flow_of(net, "cpu")
    assert built_lib.weights.packed_flow_of(net, "cpu") is packed                    # cached
    with torch.no_grad():
        net.linear1.weight.mul_(2.0)
    assert built_lib.weights.packed_flow_of(net, "cpu") is not packed                # invalidated by _version
    pe = m.positional_encoding_1(torch.tensor([[0.3, -0.2]]), 5)
    assert np.allclose(pe.numpy(), O.positional_encoding(np.array([[0.3, -0.2]], np.float32), 5), atol=1e-6)


def test_tcnn_network_shim_layout(built_lib):
    r = built_lib.reflow
    net = r.Network(25, 2, {"otype": "FullyFusedMLP", "activation": "SiLU", "output_activation": "None",
                            "n_neurons": 32, "n_hidden_layers": 3})
    assert net.params.numel() == 32 * 32 + 2 * 1024 + 16 * 32
    flow, _, _ = O.load_material_npz(DISK_FILE)
    sd = {f"l{i}": torch.from_numpy(w) for i, w in enumerate(flow.layers)}
    r.load_pytorch_model_to_tinycuda(net, sd, 25, 2)
    want = O.pack_tcnn_params(flow.layers, 25, 2).astype(np.float32)
    assert np.array_equal(net.params.detach().numpy(), want)
    # loading a second material into the same Network must not serve the stale packed blob
    first = net.packed("cpu")
    assert net.packed("cpu") is first
    sd2 = {k: 2.0 * v for k, v in sd.items()}
    r.load_pytorch_model_to_tinycuda(net, sd2, 25, 2)
    second = net.packed("cpu")
    assert second is not first and not torch.equal(second.blob, first.blob)
    net.params.data[0] = 123.0                                      # a raw .data write does not bump the version ...
    built_lib.weights.invalidate(net)                               # ... so the explicit helper drops the cache
    assert net.packed("cpu") is not second
    with pytest.raises(RuntimeError):
        r.Network(25, 2, {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None",
                          "n_neurons": 32, "n_hidden_layers": 3})


def test_model_modules_are_inference_only(built_lib):
    """forward / sample / log_prob are detached library calls: with autograd on and trainable parameters they raise
    instead of returning values a training loop could not backpropagate through."""
    m = built_lib.model
    net = m.NN_cond_pos_simpler(input_dim=5, output_dim=2, N_NEURONS=32, POSITIONAL_ENCODING_BASIS_NUM=5)
    b = m.NN_cond_pretrain_disk_one(input_dim=2, N_NEURONS=16, POSITIONAL_ENCODING_BASIS_NUM=3)
    x = torch.zeros(4, 2)
    with pytest.raises(RuntimeError, match="inference-only"):
        net(x, torch.zeros(4, 1), x)
    with pytest.raises(RuntimeError, match="inference-only"):
        b.log_prob(x, x)
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU path"):     # past the guard: CPU tensors still raise
        b.log_prob(x, x)


def test_ops_shape_validation_without_gpu(built_lib):
    """Shape mismatches are rejected on the host before any pointer reaches the library (they would be silent
    out-of-bounds device reads otherwise)."""
    ops = built_lib.ops
    flow, base, _ = O.load_material_npz(DISK_FILE)
    pf = built_lib.weights.pack_flow_layers(flow.layers, "cpu")

    class FakeCuda(torch.Tensor):
        is_cuda = True

    def fake(*shape):
        return torch.zeros(*shape).as_subclass(FakeCuda)

    with pytest.raises(ValueError, match="wi must have shape"):
        ops.sample(fake(8, 3), pf, torch.zeros(308), 4, x0=fake(8, 2))               # raw epilogue wants [n,2]
    with pytest.raises(ValueError, match="wi must have shape"):
        ops.sample(fake(8, 2), pf, torch.zeros(308), 4, epilogue=ops.EPI_DISK, x0=fake(8, 2))
    with pytest.raises(ValueError, match="x0 must have shape"):
        ops.sample(fake(8, 2), pf, torch.zeros(308), 4, x0=fake(7, 2))
    with pytest.raises(ValueError, match="wo must have shape"):
        ops.pdf(fake(7, 2), fake(8, 2), pf, torch.zeros(308), 4)
    with pytest.raises(ValueError, match="exceeds rows"):
        ops.flow_forward(fake(8, 2), pf, 4, n=100, wi_repeat=2, x0=fake(100, 2))
    with pytest.raises(ValueError, match="T must be"):
        ops.sample(fake(8, 2), pf, torch.zeros(308), 0, x0=fake(8, 2))


def test_shard_range_partitions(built_lib):
    s = built_lib.sharding
    for n in (0, 1, 7, 16, 1000003):
        for world in (1, 2, 3, 8):
            spans = [s.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, n, q):
    import torch.distributed as dist
    from bsdf_diffusion_sampling_b200 import sharding
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    full = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3)
    local = sharding.shard_rows(full, rank, world) * 2.0            # stand-in for the per-shard kernel
    out = sharding.gather_rows(local, n)
    ok = torch.equal(out, full * 2.0)
    ss = sharding.ShardedSampler(sampler=None)
    q.put((rank, ok, ss.local_range(n), ss.world))
    dist.destroy_process_group()


def test_gather_rows_gloo_world2(built_lib):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n = 1001                                                        # ragged: 501 + 500
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert [r[1] for r in res] == [True, True]
    assert res[0][2] == (0, 501) and res[1][2] == (501, 1001) and res[0][3] == 2
