"""Training step of the flow nets on the GPU: one fused forward + backward + Adam launch per iteration.

Replaces the body of the reference's ``diffusion_stage`` / ``rectify_stage`` loops
(learning_repo_cleanup/disk_domain_sampling.py:35-67, 112-136; spherical_domain_sampling.py:40-110):

    alpha   = torch.linspace(0, 1, N)
    x_alpha = (1 - alpha) * x_0 + alpha * omega_o           # spherical: phi unwrapped towards x_0, (theta, sin, cos) embedding
    pred    = diffusion_network(x_alpha, alpha, omega_i)
    loss    = torch.mean((pred - (omega_o - x_0)) ** 2);  loss.backward();  optimizer.step();  optimizer.zero_grad()

``FlowMatchingTrainer`` owns the fp32 master weights, the gradient buffer and Adam's moments as flat device vectors in
checkpoint order (``linear1.weight`` ... ``output.weight``); ``step(x_0, omega_o, omega_i)`` is ONE kernel launch
(``bsdfdiff_flow_matching_step``) that returns the loss as a device scalar -- no autograd graph, no [N, H] activation in
HBM, no separate optimizer pass, no host synchronisation.  ``state_dict()`` gives the reference's checkpoint keys, so the
result loads into the reference's ``nn.Module`` classes (and ``save`` writes what ``save_model`` writes);
``packed()`` is the sampler-ready blob for ``ops.sample`` / ``plugins.NeuralBSDFSampler`` / ``reflow.dosampling``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import _lib, ops, weights


class FlowMatchingTrainer:
    def __init__(self, layers: Sequence[torch.Tensor], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 device="cuda"):
        """``layers`` = [linear1.weight [H,in], linear2.weight [H,H], ..., output.weight [2,H]] (any device / dtype)."""
        ls = [torch.as_tensor(w).detach().to(torch.float32) for w in layers]
        self.shapes = [tuple(w.shape) for w in ls]
        H, in_dim = self.shapes[0]
        self.hidden, self.in_dim, self.n_hidden = int(H), int(in_dim), len(ls) - 1
        if in_dim not in (25, 26):
            raise ValueError(f"the flow nets take 25 (disk) or 26 (spherical) inputs, got {in_dim}")
        self.domain = _lib.DISK if in_dim == 25 else _lib.SPHERICAL
        n_params = _lib.lib.bsdfdiff_flow_param_count(self.in_dim, self.hidden, self.n_hidden)
        if n_params == 0 or self.shapes[-1] != (2, H) or any(s != (H, H) for s in self.shapes[1:-1]):
            raise ValueError(f"unsupported flow-net shape {self.shapes}")
        dev = torch.device(device)
        self.weights = torch.cat([w.reshape(-1) for w in ls]).to(dev).contiguous()
        assert self.weights.numel() == n_params
        self.grad = torch.zeros_like(self.weights)
        self.m = torch.zeros_like(self.weights)
        self.v = torch.zeros_like(self.weights)
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.steps = 0
        self._ticket = torch.zeros(4, dtype=torch.int32, device=dev)
        self._packed = None

    @classmethod
    def from_module(cls, module: torch.nn.Module, **kw) -> "FlowMatchingTrainer":
        """From an ``NN_cond_pos_simpler`` / ``NN_cond_pos`` / ``NN_cond_pos_spherical_complicate`` (reference or ours)."""
        return cls([torch.from_numpy(w) for w in weights.flow_layers_from_state_dict(module.state_dict())], **kw)

    # -- one optimisation step ---------------------------------------------------------------------------------
    def _launch(self, x_0, omega_o, omega_i, alpha, apply_update: bool) -> torch.Tensor:
        dev = self.weights.device
        x_0, omega_o, omega_i = (t.detach().to(dev, torch.float32).contiguous() for t in (x_0, omega_o, omega_i))
        n = x_0.shape[0]
        for t, name in ((x_0, "x_0"), (omega_o, "omega_o"), (omega_i, "omega_i")):
            if t.dim() != 2 or tuple(t.shape) != (n, 2):
                raise ValueError(f"{name} must have shape ({n}, 2), got {tuple(t.shape)}")
        if n < 1:
            raise ValueError("empty batch")
        if alpha is not None:
            alpha = alpha.detach().to(dev, torch.float32).reshape(-1).contiguous()
            if alpha.numel() != n:
                raise ValueError(f"alpha must have {n} elements")
        if not x_0.is_cuda:
            raise RuntimeError("bsdfdiff.training: expected CUDA tensors; this package has no CPU path")
        loss = torch.empty((), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.lib.bsdfdiff_flow_matching_step(
                self.domain, self.hidden, self.n_hidden, n, x_0.data_ptr(), omega_o.data_ptr(), omega_i.data_ptr(),
                alpha.data_ptr() if alpha is not None else None, self.weights.data_ptr(), self.grad.data_ptr(),
                self.m.data_ptr(), self.v.data_ptr(), self.lr, self.betas[0], self.betas[1], self.eps,
                self.steps + 1, 1 if apply_update else 0, loss.data_ptr(), self._ticket.data_ptr(),
                torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "bsdfdiff_flow_matching_step")
        return loss

    def step(self, x_0: torch.Tensor, omega_o: torch.Tensor, omega_i: torch.Tensor,
             alpha: Optional[torch.Tensor] = None) -> torch.Tensor:
        """loss.backward(); optimizer.step(); optimizer.zero_grad() of one reference iteration -> loss (device scalar)."""
        loss = self._launch(x_0, omega_o, omega_i, alpha, True)
        self.steps += 1
        self._packed = None
        return loss

    def loss_and_grad(self, x_0, omega_o, omega_i, alpha=None):
        """(loss, [dloss/dW per layer]) without touching the weights (what ``loss.backward()`` leaves in ``.grad``)."""
        self.grad.zero_()
        loss = self._launch(x_0, omega_o, omega_i, alpha, False)
        grads = self._split(self.grad.clone())
        self.grad.zero_()
        return loss, grads

    # -- weights out ---------------------------------------------------------------------------------------------
    def _split(self, flat: torch.Tensor) -> List[torch.Tensor]:
        out, a = [], 0
        for s in self.shapes:
            k = s[0] * s[1]
            out.append(flat[a:a + k].view(s))
            a += k
        return out

    def layers(self) -> List[torch.Tensor]:
        return self._split(self.weights)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """The reference's checkpoint keys (``save_model`` = ``torch.save(model.state_dict(), ...)``, utils/utils.py)."""
        ls = self.layers()
        sd = {f"linear{i + 1}.weight": w.clone() for i, w in enumerate(ls[:-1])}
        sd["output.weight"] = ls[-1].clone()
        return sd

    def save(self, path: str) -> None:
        torch.save({k: v.cpu() for k, v in self.state_dict().items()}, path)

    def packed(self) -> "weights.PackedFlow":
        """Sampler-ready blob of the current weights (re-packed lazily after a step)."""
        if self._packed is None:
            self._packed = weights.pack_flow_layers([w.cpu() for w in self.layers()], self.weights.device)
        return self._packed


class BasePretrainer:
    """Pretrain stage (disk_domain_sampling.py:14-33, spherical_domain_sampling.py:16-35): maximum likelihood of the base
    distribution, ``loss = -mean(D_base.log_prob(omega_o, omega_i))`` with Adam (lr 3e-4 in the reference), one launch per
    step (``bsdfdiff_base_nll_step``).  The master weights are the 308-float base blob the sampler kernels read."""

    def __init__(self, base308, domain: int, lr: float = 3e-4, betas=(0.9, 0.999), eps: float = 1e-8, device="cuda"):
        b = torch.as_tensor(base308).detach().to(torch.float32).reshape(-1)
        if b.numel() != _lib.BASE_FLOATS:
            raise ValueError(f"the base net is {_lib.BASE_FLOATS} floats (14 -> 16 -> 4 with biases), got {b.numel()}")
        if domain not in (_lib.DISK, _lib.SPHERICAL):
            raise ValueError("domain must be DISK or SPHERICAL")
        dev = torch.device(device)
        self.domain = int(domain)
        self.weights = b.to(dev).contiguous().clone()
        self.grad, self.m, self.v = (torch.zeros_like(self.weights) for _ in range(3))
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.steps = 0
        self._ticket = torch.zeros(4, dtype=torch.int32, device=dev)

    @classmethod
    def from_module(cls, module: torch.nn.Module, domain: int, **kw) -> "BasePretrainer":
        sd = module.state_dict()
        flat = torch.cat([sd[k].detach().reshape(-1).float().cpu() for k in
                          ("linear1.weight", "linear1.bias", "output.weight", "output.bias")])
        return cls(flat, domain, **kw)

    def _launch(self, omega_o, omega_i, apply_update: bool) -> torch.Tensor:
        dev = self.weights.device
        omega_o, omega_i = (t.detach().to(dev, torch.float32).contiguous() for t in (omega_o, omega_i))
        n = omega_o.shape[0]
        if n < 1 or tuple(omega_o.shape) != (n, 2) or tuple(omega_i.shape) != (n, 2):
            raise ValueError("omega_o and omega_i must both have shape (n, 2), n >= 1")
        if not omega_o.is_cuda:
            raise RuntimeError("bsdfdiff.training: expected CUDA tensors; this package has no CPU path")
        loss = torch.empty((), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.lib.bsdfdiff_base_nll_step(self.domain, n, omega_o.data_ptr(), omega_i.data_ptr(),
                                                 self.weights.data_ptr(), self.grad.data_ptr(), self.m.data_ptr(),
                                                 self.v.data_ptr(), self.lr, self.betas[0], self.betas[1], self.eps,
                                                 self.steps + 1, 1 if apply_update else 0, loss.data_ptr(),
                                                 self._ticket.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "bsdfdiff_base_nll_step")
        return loss

    def step(self, omega_o: torch.Tensor, omega_i: torch.Tensor) -> torch.Tensor:
        loss = self._launch(omega_o, omega_i, True)
        self.steps += 1
        return loss

    def loss_and_grad(self, omega_o, omega_i):
        self.grad.zero_()
        loss = self._launch(omega_o, omega_i, False)
        g = self.grad.clone()
        self.grad.zero_()
        return loss, g

    def blob(self) -> torch.Tensor:
        """The 308-float device blob ``ops.sample`` / ``NeuralBSDFSampler`` take as ``base`` (live view of the weights)."""
        return self.weights

    def state_dict(self) -> Dict[str, torch.Tensor]:
        w = self.weights
        return {"linear1.weight": w[:224].view(16, 14).clone(), "linear1.bias": w[224:240].clone(),
                "output.weight": w[240:304].view(4, 16).clone(), "output.bias": w[304:308].clone()}

    def save(self, path: str) -> None:
        torch.save({k: v.cpu() for k, v in self.state_dict().items()}, path)


def diffusion_stage_step(trainer: FlowMatchingTrainer, base_blob: torch.Tensor, brdf_samples: torch.Tensor,
                         batch: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """One iteration of ``diffusion_stage`` (disk_domain_sampling.py:46-58) on device-resident data: draw the batch rows
    of ``brdf_samples`` [Ndata, 4] = (omega_i, omega_o), x_0 ~ base(.|omega_i) with the sampler library, one training
    launch.  -> loss."""
    idx = torch.randint(0, brdf_samples.shape[0], (batch,), device=brdf_samples.device, generator=generator)
    x_1 = brdf_samples.index_select(0, idx)
    omega_i, omega_o = x_1[:, 0:2].contiguous(), x_1[:, 2:4].contiguous()
    x_0 = ops.sample(omega_i, ops.NullFlow(trainer.domain), base_blob, 0)[2]
    return trainer.step(x_0, omega_o, omega_i)


__all__ = ["FlowMatchingTrainer", "BasePretrainer", "diffusion_stage_step"]
