"""Host-side network definitions with the reference's names and parameter layout.

Only the six symbols that are live on the reference's hot path are provided:
``positional_encoding_1``, ``NN_cond_pos_simpler`` (the 3-hidden-layer definition that wins in
rendering/utils/model.py:479-501), ``NN_cond_pos`` (:422-446), ``NN_cond_pos_spherical_complicate``
(:449-477), ``NN_cond_pretrain_disk_one`` (:374-398) and ``NN_cond_pretrain_spherical_one`` (:277-317).

INFERENCE ONLY: ``forward`` / ``sample`` / ``log_prob`` are detached CUDA-library calls with no autograd, so the
reference's TRAINING stages (losses on ``log_prob`` / ``forward``) cannot backpropagate through them -- they raise
instead of silently returning constants when gradients are expected (call them under ``torch.no_grad()``).

These modules are *parameter containers*: their ``state_dict`` keys (``linear{k}.weight``,
``output.weight`` [+ ``.bias`` for the base nets]) match the reference so its checkpoints load
unchanged, and the samplers in ``mlp_brdf_sampling`` pack them for the CUDA kernels.  ``forward``,
``sample`` and ``log_prob`` run through the same CUDA library (no PyTorch re-implementation of the
math lives in this package); the reference's own ``nn.Module`` instances are accepted by the
samplers as well (duck-typed on ``state_dict()``).
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops, weights


def positional_encoding_1(tensor: torch.Tensor, num_encoding_functions: int = 6, include_input: bool = True,
                          log_sampling: bool = True) -> torch.Tensor:
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)] along the last axis
    (rendering/utils/model.py:9-57).  Host-side helper for callers that pre-encode their own inputs
    (e.g. ``network_sampling_disk_tiny``); the fused kernels evaluate the encoding on chip."""
    if not log_sampling:
        raise NotImplementedError("only log-spaced frequency bands are used by the reference")
    parts = [tensor] if include_input else []
    for k in range(num_encoding_functions):
        f = 2.0 ** k
        parts.append(torch.sin(tensor * f))
        parts.append(torch.cos(tensor * f))
    return parts[0] if len(parts) == 1 else torch.cat(parts, dim=-1)


def _inference_only(module: nn.Module, what: str, *tensors) -> None:
    if torch.is_grad_enabled() and (any(p.requires_grad for p in module.parameters())
                                    or any(t.requires_grad for t in tensors)):
        raise RuntimeError(
            f"{type(module).__name__}.{what}: this drop-in is inference-only (a detached CUDA-library call, no "
            "autograd); wrap the call in torch.no_grad(), or train with the reference's own nn.Module")


class _FlowNet(nn.Module):
    """Bias-free SiLU MLP  input_dim + 4*PE  ->  H  -> ... -> output_dim."""
    N_HIDDEN = 0

    def __init__(self, input_dim=3, output_dim=1, N_NEURONS=32, POSITIONAL_ENCODING_BASIS_NUM=5):
        super().__init__()
        self.input_dim = input_dim + 2 * POSITIONAL_ENCODING_BASIS_NUM * 2
        self.pos_num = POSITIONAL_ENCODING_BASIS_NUM
        self.linear1 = nn.Linear(self.input_dim, N_NEURONS, bias=False)
        for k in range(2, self.N_HIDDEN + 1):
            setattr(self, f"linear{k}", nn.Linear(N_NEURONS, N_NEURONS, bias=False))
        self.output = nn.Linear(N_NEURONS, output_dim, bias=False)

    def forward(self, x, alpha, x_co):
        """Velocity D(x, alpha | x_co): one fused MLP forward on the GPU."""
        _inference_only(self, "forward", x, x_co)
        packed = weights.packed_flow_of(self, x.device)
        inp = torch.cat([x, alpha, positional_encoding_1(x_co, self.pos_num)], dim=1)
        return ops.mlp_forward(inp, packed)


class NN_cond_pos_simpler(_FlowNet):
    """Disk flow net, 3 hidden layers (rendering/utils/model.py:479-501)."""
    N_HIDDEN = 3


class NN_cond_pos(_FlowNet):
    """Spherical flow net, 4 hidden layers (rendering/utils/model.py:422-446)."""
    N_HIDDEN = 4

    def __init__(self, input_dim=3, output_dim=1, N_NEURONS=64, POSITIONAL_ENCODING_BASIS_NUM=5):
        super().__init__(input_dim, output_dim, N_NEURONS, POSITIONAL_ENCODING_BASIS_NUM)


class NN_cond_pos_spherical_complicate(_FlowNet):
    """Reflow teacher, 6 hidden layers, 64 wide (rendering/utils/model.py:449-477)."""
    N_HIDDEN = 6

    def __init__(self, input_dim=3, output_dim=1, N_NEURONS=64, POSITIONAL_ENCODING_BASIS_NUM=5):
        super().__init__(input_dim, output_dim, N_NEURONS, POSITIONAL_ENCODING_BASIS_NUM)


class _BaseNet(nn.Module):
    """14 -> 16 -> 4 conditional base-distribution net (with biases)."""
    DOMAIN = ops.DISK

    def __init__(self, input_dim=2, N_NEURONS=16, POSITIONAL_ENCODING_BASIS_NUM=3):
        super().__init__()
        self.input_dim = input_dim + 2 * POSITIONAL_ENCODING_BASIS_NUM * 2
        self.POSITIONAL_ENCODING_BASIS_NUM = POSITIONAL_ENCODING_BASIS_NUM
        self.linear1 = nn.Linear(self.input_dim, N_NEURONS)
        self.output = nn.Linear(N_NEURONS, 4)


class NN_cond_pretrain_disk_one(_BaseNet):
    """Diagonal-Gaussian base on the disk (rendering/utils/model.py:374-398)."""
    DOMAIN = ops.DISK

    def __init__(self, input_dim=3, output_dim=4, N_NEURONS=16, POSITIONAL_ENCODING_BASIS_NUM=5):
        super().__init__(input_dim, N_NEURONS, POSITIONAL_ENCODING_BASIS_NUM)


class NN_cond_pretrain_spherical_one(_BaseNet):
    """Gaussian(theta) x von Mises(phi) base (rendering/utils/model.py:277-317)."""
    DOMAIN = ops.SPHERICAL

    def __init__(self, input_dim=3, N_NEURONS=16, POSITIONAL_ENCODING_BASIS_NUM=3):
        super().__init__(input_dim, N_NEURONS, POSITIONAL_ENCODING_BASIS_NUM)


def _base_sample(self, x_co, numsamples=1):
    """x0 ~ p_base(. | x_co), drawn on the GPU with Philox (T = 0 call of the fused sampler)."""
    _inference_only(self, "sample", x_co)
    base = weights.packed_base_of(self, x_co.device)
    x0, _, _ = ops.sample(x_co, ops.NullFlow(self.DOMAIN), base, 0)
    return x0


def _base_log_prob(self, x, x_co):
    """log p_base(x | x_co): the kernel returns the log density itself (finite where the density underflows)."""
    _inference_only(self, "log_prob", x, x_co)
    base = weights.packed_base_of(self, x_co.device)
    return ops.base_log_prob(x, x_co, base, self.DOMAIN)


_BaseNet.sample = _base_sample
_BaseNet.log_prob = _base_log_prob
