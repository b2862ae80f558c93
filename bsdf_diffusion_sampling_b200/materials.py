"""Packed material sets and training-data caches.

The reference keeps every material as two pickled ``.pth`` state dicts under a CWD-relative directory tree
(``./checkpoints_new/<mat>_disk/brdf_{rectify,pretrain}_network<mat>.pth``, loaded with ``torch.load`` from each plugin
instance: rendering/brdf_measured_disk.py:45,51, brdf_measured_spherical.py:54,59, bsdf_myresult.py:50,53) -- 290 files
for the shipped scenes, each needing Python's unpickler and a CUDA context just to be read -- and caches its MCMC
training samples as ``brdf_samples_emcee<mat>.npy`` next to them (learning_repo_cleanup/disk_domain_sampling.py:167-179).

``MaterialPack`` is the single-file replacement: ONE ``.bsdfpack`` per scene / material set, no pickle, memory-mappable.

    bytes 0..7    magic  b"BSDFPK01"
    bytes 8..11   uint32 little-endian: length J of the JSON index
    bytes 12..    JSON index (UTF-8): {"version": 1, "materials": [{"name", "kind", "T", "in_dim", "hidden", "n_hidden",
                  "flow": [{"shape": [r, c], "offset": o}...], "base": {"offset": o}}...]}   (offsets relative to the payload)
    payload       starts at the next multiple of 64 after the index: fp32 little-endian arrays, each 64-byte aligned:
                  the flow net's weight matrices in checkpoint order (linear1.weight .. output.weight, row-major as
                  nn.Linear stores them) and the base net as 308 floats (linear1.weight [16,14], linear1.bias [16],
                  output.weight [4,16], output.bias [4])

The arrays are the fp32 master weights, bit for bit what the ``.pth`` files hold; the device blobs of the kernels
(``weights.pack_flow_layers``) are derived at load time, so the format does not depend on the kernels' internal layout.
"""
from __future__ import annotations

import json
import os
import struct
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

from . import plugins, weights

MAGIC = b"BSDFPK01"
_ALIGN = 64


def _pad(n: int) -> int:
    return (n + _ALIGN - 1) // _ALIGN * _ALIGN


class MaterialPack:
    """An ordered set of materials: name -> (plugin kind, T, flow-net layers, base-net 308 floats), all fp32 numpy."""

    def __init__(self):
        self.entries: List[dict] = []

    # -- building --------------------------------------------------------------------------------------------------
    def add(self, name, kind: str, flow_layers: Sequence, base308, T: Optional[int] = None,
            fixup: Optional[Dict[str, float]] = None) -> "MaterialPack":
        if kind not in plugins._KINDS:
            raise ValueError(f"unknown plugin kind {kind!r}")
        layers = [np.ascontiguousarray(weights._np32(w)) for w in flow_layers]
        base = np.ascontiguousarray(weights._np32(base308)).ravel()
        H, in_dim = layers[0].shape
        want_in = 25 if plugins._KINDS[kind][0] == 0 else 26
        if in_dim != want_in or layers[-1].shape != (2, H) or any(w.shape != (H, H) for w in layers[1:-1]):
            raise ValueError(f"material {name!r}: flow-net shapes {[w.shape for w in layers]} do not fit kind {kind!r}")
        if base.size != 308:
            raise ValueError(f"material {name!r}: the base net is 308 floats, got {base.size}")
        if any(e["name"] == str(name) for e in self.entries):
            raise ValueError(f"material {name!r} is already in the pack")
        self.entries.append({"name": str(name), "kind": kind, "T": int(T or plugins._KINDS[kind][2]), "flow": layers,
                             "base": base, "fixup": None if fixup is None else {k: float(fixup[k]) for k in ("sample", "pdf")}})
        return self

    def calibrate(self, device="cuda", **kw) -> Dict[str, Dict[str, float]]:
        """Per-material fix-up thresholds of the tensor-core path (``NeuralBSDFSampler.calibrate_fixup``, needs the GPU),
        kept in the pack: ``sampler(name)`` then recomputes in fp32 only what THIS material needs to meet the parity bars
        instead of what the worst material of its family needs.  -> {name: {"sample": t, "pdf": t}}."""
        for e in self.entries:
            e["fixup"] = None
            e["fixup"] = self.sampler(e["name"], device).calibrate_fixup(**kw)
        return {e["name"]: dict(e["fixup"]) for e in self.entries}

    def add_state_dicts(self, name, kind: str, flow_sd: Dict, base_sd: Dict, T: Optional[int] = None) -> "MaterialPack":
        base = np.concatenate([weights._np32(base_sd[k]).ravel() for k in
                               ("linear1.weight", "linear1.bias", "output.weight", "output.bias")])
        return self.add(name, kind, weights.flow_layers_from_state_dict(flow_sd), base, T)

    @classmethod
    def from_checkpoints(cls, kind: str, materials: Iterable, root: str = "./checkpoints_new") -> "MaterialPack":
        """Collect the reference's per-material ``.pth`` pairs (``plugins.checkpoint_paths``, including the measured-spherical
        plugin's *_disk* pretrain quirk) into one pack."""
        pack = cls()
        for m in materials:
            fp, bp = plugins.checkpoint_paths(kind, m, root)
            pack.add_state_dicts(m, kind, weights.load_checkpoint(fp), weights.load_checkpoint(bp))
        return pack

    # -- file ------------------------------------------------------------------------------------------------------
    def save(self, path: str) -> int:
        index, off = {"version": 1, "materials": []}, 0
        blobs: List[np.ndarray] = []
        for e in self.entries:
            H, in_dim = e["flow"][0].shape
            rec = {"name": e["name"], "kind": e["kind"], "T": e["T"], "in_dim": int(in_dim), "hidden": int(H),
                   "n_hidden": len(e["flow"]) - 1, "flow": []}
            for w in e["flow"]:
                rec["flow"].append({"shape": list(w.shape), "offset": off})
                blobs.append(w)
                off += _pad(w.nbytes)
            if e.get("fixup") is not None:
                rec["fixup"] = e["fixup"]
            rec["base"] = {"offset": off}
            blobs.append(e["base"])
            off += _pad(e["base"].nbytes)
            index["materials"].append(rec)
        j = json.dumps(index, separators=(",", ":")).encode()
        head = MAGIC + struct.pack("<I", len(j)) + j
        with open(path, "wb") as fh:
            fh.write(head + b"\0" * (_pad(len(head)) - len(head)))
            for a in blobs:
                raw = a.astype("<f4", copy=False).tobytes()
                fh.write(raw + b"\0" * (_pad(len(raw)) - len(raw)))
        return os.path.getsize(path)

    @classmethod
    def load(cls, path: str) -> "MaterialPack":
        buf = np.memmap(path, dtype=np.uint8, mode="r")
        if buf.size < 12 or bytes(buf[:8]) != MAGIC:
            raise ValueError(f"{path}: not a BSDFPK01 material pack")
        jl = struct.unpack("<I", bytes(buf[8:12]))[0]
        index = json.loads(bytes(buf[12:12 + jl]).decode())
        if index.get("version") != 1:
            raise ValueError(f"{path}: unsupported pack version {index.get('version')}")
        base_off = _pad(12 + jl)

        def arr(off, count):
            a, b = base_off + off, base_off + off + 4 * count
            if b > buf.size:
                raise ValueError(f"{path}: truncated (array at {a}..{b}, file has {buf.size} bytes)")
            return np.frombuffer(buf, dtype="<f4", count=count, offset=a)

        pack = cls()
        for rec in index["materials"]:
            layers = [arr(f["offset"], int(np.prod(f["shape"]))).reshape(f["shape"]) for f in rec["flow"]]
            pack.add(rec["name"], rec["kind"], layers, arr(rec["base"]["offset"], 308), rec["T"], rec.get("fixup"))
        return pack

    # -- use -------------------------------------------------------------------------------------------------------
    def names(self) -> List[str]:
        return [e["name"] for e in self.entries]

    def sampler(self, name, device="cuda", **kw) -> "plugins.NeuralBSDFSampler":
        e = next((e for e in self.entries if e["name"] == str(name)), None)
        if e is None:
            raise KeyError(name)
        kw.setdefault("T", e["T"])
        if e.get("fixup") is not None:
            kw.setdefault("fixup", dict(e["fixup"]))
        return plugins.NeuralBSDFSampler(e["kind"], weights.pack_flow_layers(e["flow"], device),
                                         torch.from_numpy(np.array(e["base"])).to(device), **kw)

    def multi_sampler(self, device="cuda", **kw) -> "plugins.MultiMaterialSampler":
        """All materials of the pack as ONE single-launch wavefront sampler (material id = position in ``names()``)."""
        return plugins.MultiMaterialSampler([self.sampler(n, device, **kw) for n in self.names()])

    def state_dicts(self, name):
        """(flow state_dict, base state_dict) with the reference's checkpoint keys -- loads into its nn.Modules."""
        e = next(e for e in self.entries if e["name"] == str(name))
        flow = {f"linear{i + 1}.weight": torch.from_numpy(np.array(w)) for i, w in enumerate(e["flow"][:-1])}
        flow["output.weight"] = torch.from_numpy(np.array(e["flow"][-1]))
        b = np.array(e["base"])
        base = {"linear1.weight": torch.from_numpy(b[:224].reshape(16, 14)), "linear1.bias": torch.from_numpy(b[224:240]),
                "output.weight": torch.from_numpy(b[240:304].reshape(4, 16)), "output.bias": torch.from_numpy(b[304:308])}
        return flow, base


# ---- MCMC training-sample cache (disk_domain_sampling.py:167-179) -----------------------------------------------------
def emcee_cache_path(save_dir: str, material: str, domain: str = "disk") -> str:
    """Where the reference's training scripts keep the (omega_i, omega_o) samples of a material."""
    return os.path.join(save_dir, f"{material}_{domain}", f"brdf_samples_emcee{material}.npy")


def load_emcee_cache(path: str, device="cuda") -> torch.Tensor:
    """``brdf_samples_emcee<mat>.npy`` -> [N,4] fp32 device tensor (omega_i, omega_o), as the training stages index it
    (``torch.from_numpy(np.load(path)).to("cuda").type(torch.float32)``, disk_domain_sampling.py:174,181)."""
    a = np.load(path, mmap_mode="r")
    if a.ndim != 2 or a.shape[1] != 4:
        raise ValueError(f"{path}: expected an [N,4] array of (omega_i, omega_o) samples, got {a.shape}")
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)


def save_emcee_cache(path: str, samples) -> None:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    a = samples.detach().cpu().numpy() if isinstance(samples, torch.Tensor) else np.asarray(samples)
    if a.ndim != 2 or a.shape[1] != 4:
        raise ValueError("samples must be [N,4]")
    np.save(path, a)


__all__ = ["MaterialPack", "emcee_cache_path", "load_emcee_cache", "save_emcee_cache"]
