"""Packed material sets + MCMC sample cache: file round trip is bit exact, the loader replaces the per-plugin
torch.load calls, and (GPU) samplers built from a pack equal samplers built from the reference-layout .pth files."""
import os
import struct

import numpy as np
import pytest
import torch

from conftest import BSDF_FILE, DISK_FILE, GOLDEN_FILES, SPH_FILE
from oracle import bsdf_oracle as O


def _write_pth(pkg, root, kind, mat, path):
    flow, base, _ = O.load_material_npz(path)
    fp, bp = pkg.plugins.checkpoint_paths(kind, mat, root)
    for f in (fp, bp):
        os.makedirs(os.path.dirname(f), exist_ok=True)
    sd = {f"linear{i + 1}.weight": torch.from_numpy(w) for i, w in enumerate(flow.layers[:-1])}
    sd["output.weight"] = torch.from_numpy(flow.layers[-1])
    torch.save(sd, fp)
    torch.save({"linear1.weight": torch.from_numpy(base.w1), "linear1.bias": torch.from_numpy(base.b1),
                "output.weight": torch.from_numpy(base.wo), "output.bias": torch.from_numpy(base.bo)}, bp)
    return flow, base


def test_pack_round_trip_and_checks(built_lib, tmp_path):
    M = built_lib.materials
    root = str(tmp_path / "ckpt")
    disk_files = [f for f in GOLDEN_FILES if os.path.basename(f).startswith("disk_")]
    mats = [os.path.basename(f)[5:-4] for f in disk_files]
    refs = [_write_pth(built_lib, root, "disk", m, f) for m, f in zip(mats, disk_files)]
    pack = M.MaterialPack.from_checkpoints("disk", mats, root)
    path = str(tmp_path / "scene.bsdfpack")
    size = pack.save(path)
    raw = open(path, "rb").read()
    assert raw[:8] == b"BSDFPK01" and size == len(raw) and size < 3 * 16_000 + 4096
    jl = struct.unpack("<I", raw[8:12])[0]
    assert b'"version":1' in raw[12:12 + jl]
    back = M.MaterialPack.load(path)
    assert back.names() == mats
    for (flow, base), e in zip(refs, back.entries):
        assert e["kind"] == "disk" and e["T"] == 4
        assert all(np.array_equal(a, b) for a, b in zip(e["flow"], flow.layers))           # bit exact
        assert np.array_equal(e["base"], np.concatenate([base.w1.ravel(), base.b1, base.wo.ravel(), base.bo]))
    fsd, bsd = back.state_dicts(mats[0])
    assert list(fsd.keys()) == ["linear1.weight", "linear2.weight", "linear3.weight", "output.weight"]
    assert bsd["linear1.weight"].shape == (16, 14) and bsd["output.bias"].shape == (4,)
    net = built_lib.model.NN_cond_pos_simpler(input_dim=5, output_dim=2, N_NEURONS=32, POSITIONAL_ENCODING_BASIS_NUM=5)
    net.load_state_dict(fsd)                                                                # the reference's keys
    # mixed kinds in one pack, custom T
    sflow, sbase, _ = O.load_material_npz(BSDF_FILE)
    pack.add("bsdf0", "bsdf", sflow.layers, np.concatenate([sbase.w1.ravel(), sbase.b1, sbase.wo.ravel(), sbase.bo]), T=6)
    pack.save(path)
    assert M.MaterialPack.load(path).entries[-1]["T"] == 6
    with pytest.raises(ValueError, match="do not fit kind"):
        pack.add("wrong", "disk", sflow.layers, np.zeros(308, np.float32))
    with pytest.raises(ValueError, match="already in the pack"):
        pack.add("bsdf0", "bsdf", sflow.layers, np.zeros(308, np.float32))
    with pytest.raises(ValueError, match="308 floats"):
        pack.add("short", "bsdf", sflow.layers, np.zeros(300, np.float32))
    open(str(tmp_path / "bad.bsdfpack"), "wb").write(b"NOTAPACK" + raw[8:])
    with pytest.raises(ValueError, match="not a BSDFPK01"):
        M.MaterialPack.load(str(tmp_path / "bad.bsdfpack"))
    open(str(tmp_path / "cut.bsdfpack"), "wb").write(raw[: len(raw) // 2])
    with pytest.raises(ValueError, match="truncated"):
        M.MaterialPack.load(str(tmp_path / "cut.bsdfpack"))


def test_emcee_cache_round_trip(built_lib, tmp_path):
    M = built_lib.materials
    p = M.emcee_cache_path(str(tmp_path), "matA", "disk")
    assert p.endswith(os.path.join("matA_disk", "brdf_samples_emcee" + "matA.npy"))
    s = np.random.default_rng(0).normal(size=(1000, 4))                     # emcee chains are float64
    M.save_emcee_cache(p, s)
    t = M.load_emcee_cache(p, device="cpu")
    assert t.dtype == torch.float32 and t.shape == (1000, 4) and np.array_equal(t.numpy(), s.astype(np.float32))
    np.save(p, s[:, :3])
    with pytest.raises(ValueError, match=r"\[N,4\]"):
        M.load_emcee_cache(p, device="cpu")


@pytest.mark.gpu
def test_samplers_from_pack_equal_samplers_from_checkpoints(built_lib, tmp_path):
    pkg = built_lib
    root = str(tmp_path / "ckpt")
    disk_files = [f for f in GOLDEN_FILES if os.path.basename(f).startswith("disk_")]
    mats = [os.path.basename(f)[5:-4] for f in disk_files]
    for m, f in zip(mats, disk_files):
        _write_pth(pkg, root, "disk", m, f)
    path = str(tmp_path / "scene.bsdfpack")
    pkg.materials.MaterialPack.from_checkpoints("disk", mats, root).save(path)
    pack = pkg.materials.MaterialPack.load(path)
    rng = np.random.default_rng(2)
    w = rng.normal(size=(20_000, 3)).astype(np.float32)
    w[:, 2] = np.abs(w[:, 2]) + 0.05
    wi = torch.from_numpy(w / np.linalg.norm(w, axis=1, keepdims=True)).cuda()
    for m in mats:
        a = pack.sampler(m).sample(wi, seed=3, offset=8)
        b = pkg.plugins.NeuralBSDFSampler.from_checkpoints("disk", m, root).sample(wi, seed=3, offset=8)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    mm = pack.multi_sampler()
    mid = torch.from_numpy(rng.integers(0, len(mats), wi.shape[0]).astype(np.int32)).cuda()
    wo, pdf = mm.sample(wi, mid, seed=3, offset=8)
    for k, m in enumerate(mats):
        sel = (mid == k).nonzero().squeeze(1)
        ref = pack.sampler(m).sample(wi, seed=3, offset=8)
        assert torch.equal(wo[sel], ref[0][sel]) and torch.equal(pdf[sel], ref[1][sel])
