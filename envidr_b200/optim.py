"""Fused Adam for the trainable state of the path, drop-in for the optimizer the reference builds in
main_nerf.py:150:

    optimizer = lambda model: torch.optim.Adam(model.get_params(opt.lr, ...), betas=(0.9, 0.99), eps=1e-15)
 -> optimizer = lambda model: envidr_b200.optim.FusedAdam(model.get_params(opt.lr, ...), betas=(0.9, 0.99), eps=1e-15)

Same constructor arguments (the subset torch.optim.Adam accepts that the fused kernel implements: no weight decay, no amsgrad,
no maximize -- anything else raises), same param_groups (so LambdaLR / GradScaler work unchanged), same state layout
(`step` as a CPU float tensor, `exp_avg`, `exp_avg_sq`: state_dict() is interchangeable with torch.optim.Adam's).
step() is ONE kernel launch over all parameter tensors (csrc/optim.cu); with zero_grad=True it also clears the gradients in the
same pass (the reference calls optimizer.zero_grad() before every backward, nerf/utils.py:1079).
"""
from __future__ import annotations

import ctypes
from typing import Iterable

import torch

from ._lib import check, lib, stream


class AdamTensor(ctypes.Structure):
    _fields_ = [("param", ctypes.c_void_p), ("grad", ctypes.c_void_p), ("exp_avg", ctypes.c_void_p), ("exp_avg_sq", ctypes.c_void_p),
                ("n", ctypes.c_uint64), ("step_size", ctypes.c_float), ("bias_correction2_sqrt", ctypes.c_float)]


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0,
                 amsgrad: bool = False, *, maximize: bool = False, zero_grad: bool = False, div_mode: int = 0):
        if weight_decay != 0 or amsgrad or maximize:
            raise ValueError("FusedAdam implements the configuration the reference trains with (no weight decay / amsgrad / maximize)")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("invalid Adam hyper-parameter")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None, capturable=False,
                        differentiable=False, fused=None)
        super().__init__(params, defaults)
        self.fused_zero_grad = zero_grad
        self.div_mode = div_mode
        b, e = self.param_groups[0]["betas"], self.param_groups[0]["eps"]
        for g in self.param_groups:
            if tuple(g["betas"]) != tuple(b) or g["eps"] != e:
                raise ValueError("FusedAdam: betas / eps must be the same in every parameter group (lr may differ)")

    # Host cost matters here: the kernel runs for ~0.1 ms, so the per-step Python work is kept to one data_ptr() per
    # parameter.  Step counts live in Python ints; the `step` tensors torch.optim.Adam keeps in its state are refreshed
    # when the state is read (state_dict) and re-read after load_state_dict.
    def _init_param(self, p):
        if p.grad.is_sparse or p.dtype != torch.float32 or not p.is_cuda:
            raise RuntimeError("FusedAdam: dense fp32 CUDA parameters only (no CPU fallback)")
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
        if not (p.is_contiguous() and st["exp_avg"].is_contiguous() and st["exp_avg_sq"].is_contiguous()):
            raise RuntimeError("FusedAdam: contiguous tensors only")
        self._fast[p] = [p.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel(), int(float(st["step"]))]

    def _sync_step_tensors(self):
        for p, f in getattr(self, "_fast", {}).items():
            self.state[p]["step"].fill_(float(f[4]))

    def state_dict(self):
        self._sync_step_tensors()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._fast = {}

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if not hasattr(self, "_fast"):
            self._fast = {}
        fast = self._fast
        beta1, beta2 = self.param_groups[0]["betas"]
        eps = self.param_groups[0]["eps"]
        n_max = sum(len(g["params"]) for g in self.param_groups)
        arr = getattr(self, "_arr", None)
        if arr is None or len(arr) < n_max:
            arr = self._arr = (AdamTensor * n_max)()
        k = 0
        for group in self.param_groups:
            lr = group["lr"]
            for p in group["params"]:
                g = p.grad
                if g is None:
                    continue
                f = fast.get(p)
                if f is None or f[0] != p.data_ptr():
                    self._init_param(p)
                    f = fast[p]
                if not g.is_contiguous():
                    raise RuntimeError("FusedAdam: contiguous gradients only")
                f[4] += 1
                step = f[4]
                d = arr[k]
                d.param, d.grad, d.exp_avg, d.exp_avg_sq, d.n = f[0], g.data_ptr(), f[1], f[2], f[3]
                d.step_size = lr / (1 - beta1 ** step)                      # step_size = lr / bias_correction1
                d.bias_correction2_sqrt = (1 - beta2 ** step) ** 0.5
                k += 1
        if k:
            check(lib().envidr_adam_step(arr, k, beta1, beta2, eps, int(self.fused_zero_grad), int(self.div_mode), stream()), "adam_step")
        return loss
