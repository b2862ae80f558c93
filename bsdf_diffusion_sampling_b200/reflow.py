"""Reflow ("rectify") sampling entry points of learning_repo_cleanup/*_sampling.py on the fused kernels.

* ``Network`` -- a ``tinycudann.Network``-compatible inference shim (FullyFusedMLP / SiLU / no output
  activation, 2 outputs): same constructor, ``.params`` flat parameter vector in tiny-cuda-nn's padded
  layout, ``forward(x[N,in] f32) -> [N,out] f16``
  (tiny-cuda-nn/bindings/torch/tinycudann/modules.py:162-192, 251-284).
* ``load_pytorch_model_to_tinycuda`` -- learning_repo_cleanup/utils/utils.py:13-23, verbatim semantics.
* ``dosampling`` -- the closure ``rectify_stage.dosampling(batchsize, omega_i, T)`` exposed as a function
  (learning_repo_cleanup/disk_domain_sampling.py:93-110, spherical_domain_sampling.py:147-166,
  bsdf_correct_sampling.py:147-166).  The reference runs T x (cat + identity-cast + tcnn MLP + axpy)
  through HBM; here the whole T-step loop is ONE launch (``bsdfdiff_flow_forward``): positional encoding,
  periodic re-embedding, MLP and Euler update stay on chip, ``repeat_interleave`` is an index division.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import ops, weights


class Network(nn.Module):
    def __init__(self, n_input_dims: int, n_output_dims: int, network_config: dict, seed: int = 1337):
        super().__init__()
        cfg = dict(network_config)
        otype = cfg.get("otype", "FullyFusedMLP")
        if otype not in ("FullyFusedMLP", "CutlassMLP"):
            raise RuntimeError(f"Network: unsupported otype {otype!r}")
        if str(cfg.get("activation", "ReLU")).lower() != "silu":
            raise RuntimeError("Network: only the SiLU activation (the reference's tiny-cuda-nn patch) is supported")
        if str(cfg.get("output_activation", "None")).lower() != "none":
            raise RuntimeError("Network: only output_activation 'None' is supported")
        self.n_input_dims, self.n_output_dims = int(n_input_dims), int(n_output_dims)
        self.n_neurons, self.n_hidden_layers = int(cfg["n_neurons"]), int(cfg["n_hidden_layers"])
        if self.n_neurons not in (32, 64) or self.n_output_dims != 2 or not (1 <= self.n_input_dims <= 32):
            raise RuntimeError("Network: supported shapes are in<=32, n_neurons in {32,64}, out=2")
        self.network_config = cfg
        self.seed = seed
        self.dtype = torch.float16
        in_pad = self.n_input_dims + 16 - self.n_input_dims % 16
        out_pad = self.n_output_dims + 16 - self.n_output_dims % 16
        n = self.n_neurons * in_pad + (self.n_hidden_layers - 1) * self.n_neurons ** 2 + out_pad * self.n_neurons
        g = torch.Generator().manual_seed(seed)
        scale = (6.0 / (2 * self.n_neurons)) ** 0.5          # xavier-uniform-like, as tcnn initialises
        self.params = nn.Parameter((torch.rand(n, generator=g) * 2 - 1) * scale, requires_grad=True)
        self._packed: Optional[weights.PackedFlow] = None
        self._packed_sig = None

    def packed(self, device) -> weights.PackedFlow:
        sig = (str(device), self.params._version, self.params.data_ptr())
        if self._packed is None or self._packed_sig != sig:
            # tcnn casts params to fp16 before use (modules.py:188)
            p = self.params.detach().to(torch.float16).to(torch.float32)
            self._packed = weights.pack_flow_tcnn(p, self.n_input_dims, self.n_output_dims, self.n_neurons,
                                                  self.n_hidden_layers, device)
            self._packed_sig = sig
        return self._packed

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            x = x.cuda()
        if x.shape[1] != self.n_input_dims:
            raise RuntimeError(f"Network: expected [N,{self.n_input_dims}] input, got {tuple(x.shape)}")
        return ops.mlp_forward(x, self.packed(x.device)).to(torch.float16)


def load_pytorch_model_to_tinycuda(model: Network, state_dict, input_dims: int, output_dims: int) -> None:
    """Write a PyTorch flow-net state_dict into ``model.params`` in tiny-cuda-nn's layout
    (learning_repo_cleanup/utils/utils.py:13-23): first layer zero-padded by 16-(in%16) columns, last
    by 16-(out%16) rows, everything row-major, concatenated, rounded to fp16."""
    sd = list(state_dict.values())
    parts = [nn.functional.pad(sd[0], pad=(0, 16 - (input_dims % 16), 0, 0)).flatten()]
    parts += [w.flatten() for w in sd[1:-1]]
    parts.append(nn.functional.pad(sd[-1], pad=(0, 0, 0, 16 - (output_dims % 16))).flatten())
    flat = torch.cat([p.detach().to("cpu", torch.float32) for p in parts]).half()
    with torch.no_grad():
        # copy_ on the Parameter itself bumps its version counter (a write through .data would not), and the packed
        # blob is dropped explicitly: the next forward / dosampling call re-packs the new weights
        model.params.copy_(flat.to(model.params.device, model.params.dtype))
    model._packed = None
    model._packed_sig = None


def dosampling(batchsize: int, omega_i: torch.Tensor, T: int, pretrain_network, rectify_net, *,
               x0: Optional[torch.Tensor] = None, seed: Optional[int] = None, offset: int = 0, precision=None):
    """-> (x_alpha [N,2], x_base_samples [N,2], x_target_y [N,2]) with N = batchsize * len(omega_i).

    ``pretrain_network`` is the base net (NN_cond_pretrain_*_one), ``rectify_net`` the flow net: either a
    ``reflow.Network`` loaded with ``load_pytorch_model_to_tinycuda`` or a PyTorch flow module
    (NN_cond_pos_simpler / NN_cond_pos_spherical_complicate)."""
    if not omega_i.is_cuda:
        omega_i = omega_i.cuda()
    omega_i = omega_i.detach().to(torch.float32).contiguous()
    dev = omega_i.device
    flow = rectify_net.packed(dev) if isinstance(rectify_net, Network) else weights.packed_flow_of(rectify_net, dev)
    base = weights.packed_base_of(pretrain_network, dev)
    n = int(batchsize) * omega_i.shape[0]
    x, x_base = ops.flow_forward(omega_i, flow, T, n=n, wi_repeat=int(batchsize), base=base, x0=x0, seed=seed,
                                 offset=offset, precision=precision)
    return x, x_base, omega_i.repeat_interleave(int(batchsize), 0)
