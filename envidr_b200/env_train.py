"""env_net of the training branch as ONE forward and ONE backward kernel (csrc/env_train_tc.cu, csrc/field_tc.cu).

Reference: `get_color_mlp_extra_params` / `forward_color` (nerf/network.py:527-541, 589-607) evaluate
    f_n = unitNorm(env_net(IDE(n,  diffuse_kappa_inv) * light_intensity_scale))
    f_r = unitNorm(env_net(IDE(w_r, roughness)        * light_intensity_scale))
as nn.Linear / ReLU stacks under autograd.  `env_features(...)` below returns the same pair from one fused forward kernel
(IDE -> all layers -> unit norm, activations on chip) and differentiates it with one fused backward kernel (normalize backward ->
data-gradient chain through all layers with the ReLU masks -> gradient w.r.t. the IDE features), the IDE backward kernel, and one
tensor-core weight-gradient GEMM per layer.  Once differentiable, like `linear_tc` (no shipped loss differentiates env_net twice).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch
from torch.autograd import Function

from ._lib import ENVIDR_MAX_LAYERS, EnvidrError, check, lib, ptr, stream
from .linear_tc import wgrad_tc


class EnvMlpDesc(ctypes.Structure):
    """envidr_env_mlp (include/envidr_b200.h)."""
    _fields_ = [("n_layers", ctypes.c_uint32), ("dims", ctypes.c_uint32 * (ENVIDR_MAX_LAYERS + 1)),
                ("weight", ctypes.c_void_p * ENVIDR_MAX_LAYERS), ("bias", ctypes.c_void_p * ENVIDR_MAX_LAYERS),
                ("ide_degree", ctypes.c_uint32), ("diffuse_kappa_inv", ctypes.c_float), ("light_intensity_scale", ctypes.c_float)]


def _desc(Ws: Sequence[torch.Tensor], bs: Sequence[Optional[torch.Tensor]], deg: int, kappa: float, lis: float) -> EnvMlpDesc:
    d = EnvMlpDesc()
    d.n_layers = len(Ws)
    d.dims[0] = Ws[0].shape[1]
    for i, (W, b) in enumerate(zip(Ws, bs)):
        d.dims[i + 1] = W.shape[0]
        d.weight[i] = W.data_ptr()
        d.bias[i] = 0 if b is None else b.data_ptr()
    d.ide_degree, d.diffuse_kappa_inv, d.light_intensity_scale = int(deg), float(kappa), float(lis)
    return d


def supported(Ws: Sequence[torch.Tensor], deg: int) -> bool:
    """Shapes the fused kernels take: IDE input of degree `deg`, 2..4 layers, hidden widths multiples of 32 in 64..256, env_feat <= 12."""
    if not (2 <= len(Ws) <= 4) or not Ws[0].is_cuda:
        return False
    bs = [None] * len(Ws)
    return lib().envidr_env_mlp_blob_bytes(ctypes.byref(_desc(Ws, bs, deg, 0.64, 1.0))) > 0


class _EnvNetFused(Function):
    @staticmethod
    def forward(ctx, normals, w_r, roughness, deg, kappa, lis, n_layers, *wb):
        dev = normals.device
        f32 = dict(dtype=torch.float32, device=dev)
        Ws = [wb[2 * i].detach().float().contiguous() for i in range(n_layers)]
        bs = [None if wb[2 * i + 1] is None else wb[2 * i + 1].detach().float().contiguous() for i in range(n_layers)]
        n = normals.detach().float().reshape(-1, 3)
        r = w_r.detach().float().reshape(-1, 3)
        rough = roughness.detach().float().reshape(-1)
        M = n.shape[0]
        d = _desc(Ws, bs, deg, kappa, lis)
        nbytes = lib().envidr_env_mlp_blob_bytes(ctypes.byref(d))
        if nbytes == 0:
            raise EnvidrError("env_net outside the fused training kernels (envidr_env_mlp_blob_bytes)")
        blob = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        check(lib().envidr_env_mlp_pack(ctypes.byref(d), ptr(blob), nbytes, stream()), "env_mlp_pack")
        rec = torch.zeros(M, 32, **f32)                          # the inference path's sample record: 20 roughness, 22.. normal, 25.. reflected
        rec[:, 20] = rough
        rec[:, 22:25] = n
        rec[:, 25:28] = r
        feat = torch.empty(M, 32, **f32)
        hidden = [int(W.shape[0]) for W in Ws[:-1]]
        acts = [torch.empty(2 * M, h, **f32) for h in hidden]
        masks = [torch.empty(2 * M, h // 32, dtype=torch.int32, device=dev) for h in hidden]
        pa = [ptr(a) for a in acts] + [None] * (3 - len(acts))
        pm = [ptr(m) for m in masks] + [None] * (3 - len(masks))
        check(lib().envidr_env_mlp_forward(ctypes.byref(d), ptr(blob), ptr(rec), M, ptr(feat), pa[0], pa[1], pa[2], pm[0], pm[1], pm[2], stream()),
              "env_mlp_forward")
        E = int(Ws[-1].shape[0])
        ctx.save_for_backward(n, r, rough, feat, blob, *Ws, *acts, *masks)
        ctx.meta = (int(deg), float(kappa), float(lis), n_layers, [b is not None for b in bs], tuple(roughness.shape))
        return feat[:, :E].contiguous(), feat[:, 16:16 + E].contiguous()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_n, g_r):
        deg, kappa, lis, nl, has_b, rough_shape = ctx.meta
        sv = ctx.saved_tensors
        n, r, rough, feat, blob = sv[:5]
        Ws = list(sv[5:5 + nl])
        acts = list(sv[5 + nl:5 + nl + nl - 1])
        masks = list(sv[5 + nl + nl - 1:])
        dev = n.device
        f32 = dict(dtype=torch.float32, device=dev)
        M, E = n.shape[0], Ws[-1].shape[0]
        d = _desc(Ws, [None] * nl, deg, kappa, lis)
        cols = int(lib().envidr_env_mlp_input_cols(ctypes.byref(d)))
        gfeat = torch.zeros(M, 32, **f32)
        gfeat[:, :E] = g_n
        gfeat[:, 16:16 + E] = g_r
        gacts = [torch.empty_like(a) for a in acts]
        gy = torch.empty(2 * M, 16, **f32)
        gx0 = torch.empty(2 * M, cols, **f32)
        pm = [ptr(m) for m in masks] + [None] * (3 - len(masks))
        pg = [ptr(g) for g in gacts] + [None] * (3 - len(gacts))
        check(lib().envidr_env_mlp_backward(ctypes.byref(d), ptr(blob), ptr(gfeat), ptr(feat), pm[0], pm[1], pm[2], M, pg[0], pg[1], pg[2], ptr(gy),
                                            ptr(gx0), stream()), "env_mlp_backward")
        # ---- inputs: de-interleave d x0 ([Re_0, Im_0, Re_1, ...] -> [Re | Im]) and run the IDE backward kernel over the [2M] batch
        need_in = any(ctx.needs_input_grad[:3])
        need_w = any(ctx.needs_input_grad[7:])
        dirs = torch.cat([n, r], 0)
        kap = torch.cat([torch.full((M,), kappa, **f32), rough], 0)
        P2 = Ws[0].shape[1]
        g_n_out = g_r_out = g_rough = None
        if need_in:
            ge = gx0[:, :P2].reshape(2 * M, P2 // 2, 2)
            g_enc = (torch.cat([ge[..., 0], ge[..., 1]], -1) * lis).contiguous()
            gd = torch.empty(2 * M, 3, **f32)
            gk = torch.empty(2 * M, **f32)
            check(lib().envidr_ide_encode_backward(ptr(dirs), ptr(kap), 0.0, 2 * M, deg, 1.0, ptr(g_enc), ptr(gd), ptr(gk), stream()),
                  "ide_encode_backward")
            g_n_out, g_r_out, g_rough = gd[:M], gd[M:], gk[M:].reshape(rough_shape)
        # ---- parameters: dW_l = gz_l^T a_{l-1} (tensor-core weight-gradient GEMM), db_l = column sums of gz_l
        grads: List[Optional[torch.Tensor]] = []
        if need_w:
            x0 = torch.empty(2 * M, P2, **f32)
            check(lib().envidr_ide_encode_forward(ptr(dirs), ptr(kap), 0.0, 2 * M, deg, lis, ptr(x0), stream()), "ide_encode_forward")
            a_in = [x0] + acts
            gz = gacts + [gy]
            for l in range(nl):
                want_w, want_b = ctx.needs_input_grad[7 + 2 * l], has_b[l] and ctx.needs_input_grad[8 + 2 * l]
                dW = db = None
                if want_w or want_b:
                    dW, db = wgrad_tc(gz[l], a_in[l], with_bias=True)
                    if l == nl - 1:
                        dW, db = dW[:E], db[:E]
                grads += [dW if want_w else None, db if want_b else None]
        else:
            grads = [None] * (2 * nl)
        return (g_n_out, g_r_out, g_rough, None, None, None, None, *grads)


def env_features(normals: torch.Tensor, w_r: torch.Tensor, roughness: torch.Tensor, layers: Sequence[Tuple[torch.Tensor, Optional[torch.Tensor]]],
                 deg: int, diffuse_kappa_inv: float, light_intensity_scale: float):
    """(f_n, f_r) = unit-norm env features of the normal and the reflected direction, [M, env_feat] each (see the module docstring).
    layers: [(W [out, in], b [out] or None)] of env_net."""
    wb = []
    for W, b in layers:
        wb += [W, b]
    return _EnvNetFused.apply(normals, w_r, roughness, int(deg), float(diffuse_kappa_inv), float(light_intensity_scale), len(layers), *wb)


# ---------------------------------------------------------------------------------------------------------------------
# The small ReLU heads (diffuse_net, color_net, renv_net) through the same pair of chain kernels
# ---------------------------------------------------------------------------------------------------------------------

def mlp_supported(Ws: Sequence[torch.Tensor]) -> bool:
    """2..4 layers, inputs <= 64, hidden widths multiples of 32 (<= 256), outputs <= 16."""
    if not (2 <= len(Ws) <= 4) or not Ws[0].is_cuda:
        return False
    return lib().envidr_mlp_blob_bytes(ctypes.byref(_desc(Ws, [None] * len(Ws), 1, 0.0, 1.0))) > 0


class _MlpFused(Function):
    @staticmethod
    def forward(ctx, x, n_layers, *wb):
        dev = x.device
        f32 = dict(dtype=torch.float32, device=dev)
        Ws = [wb[2 * i].detach().float().contiguous() for i in range(n_layers)]
        bs = [None if wb[2 * i + 1] is None else wb[2 * i + 1].detach().float().contiguous() for i in range(n_layers)]
        x2 = x.detach().float().contiguous()
        rows, K0 = x2.shape
        d = _desc(Ws, bs, 1, 0.0, 1.0)
        nbytes = lib().envidr_mlp_blob_bytes(ctypes.byref(d))
        if nbytes == 0:
            raise EnvidrError("MLP outside the fused chain kernels (envidr_mlp_blob_bytes)")
        blob = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        check(lib().envidr_mlp_pack(ctypes.byref(d), ptr(blob), nbytes, stream()), "mlp_pack")
        hidden = [int(W.shape[0]) for W in Ws[:-1]]
        trains = any(ctx.needs_input_grad[2:])                 # some weight / bias wants a gradient: keep the activations
        masks = [torch.empty(rows, h // 32, dtype=torch.int32, device=dev) for h in hidden]
        acts = [torch.empty(rows, h, **f32) for h in hidden] if trains else []
        y = torch.empty(rows, 16, **f32)
        pm = [ptr(m) for m in masks] + [None] * (3 - len(masks))
        pa = [ptr(a) for a in acts] + [None] * (3 - len(acts))
        check(lib().envidr_mlp_forward(ctypes.byref(d), ptr(blob), ptr(x2), K0, rows, ptr(y), pm[0], pm[1], pm[2], pa[0], pa[1], pa[2], stream()),
              "mlp_forward")
        ctx.save_for_backward(x2, blob, *Ws, *masks, *acts)
        ctx.meta = (n_layers, [b is not None for b in bs], trains)
        return y[:, :Ws[-1].shape[0]].contiguous()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        nl, has_b, trains = ctx.meta
        sv = ctx.saved_tensors
        x2, blob = sv[:2]
        Ws = list(sv[2:2 + nl])
        masks = list(sv[2 + nl:2 + nl + nl - 1])
        acts = list(sv[2 + nl + nl - 1:])
        dev = x2.device
        f32 = dict(dtype=torch.float32, device=dev)
        rows, K0 = x2.shape
        N = Ws[-1].shape[0]
        d = _desc(Ws, [None] * nl, 1, 0.0, 1.0)
        gy16 = torch.zeros(rows, 16, **f32)
        gy16[:, :N] = gy
        gzs = [torch.empty(rows, int(W.shape[0]), **f32) for W in Ws[:-1]] if trains else []
        cols = (K0 + 15) // 16 * 16
        gx = torch.empty(rows, cols, **f32)
        pm = [ptr(m) for m in masks] + [None] * (3 - len(masks))
        pg = [ptr(g) for g in gzs] + [None] * (3 - len(gzs))
        check(lib().envidr_mlp_backward(ctypes.byref(d), ptr(blob), ptr(gy16), 16, pm[0], pm[1], pm[2], rows, pg[0], pg[1], pg[2], ptr(gx), stream()),
              "mlp_backward")
        grads: List[Optional[torch.Tensor]] = []
        if trains:
            a_in = [x2] + acts
            gz = gzs + [gy16]
            for l in range(nl):
                want_w, want_b = ctx.needs_input_grad[2 + 2 * l], has_b[l] and ctx.needs_input_grad[3 + 2 * l]
                dW = db = None
                if want_w or want_b:
                    dW, db = wgrad_tc(gz[l], a_in[l], with_bias=True)
                    if l == nl - 1:
                        dW, db = dW[:N], db[:N]
                grads += [dW if want_w else None, db if want_b else None]
        else:
            grads = [None] * (2 * nl)
        return (gx[:, :K0] if ctx.needs_input_grad[0] else None, None, *grads)


def mlp_fused(x: torch.Tensor, layers: Sequence[Tuple[torch.Tensor, Optional[torch.Tensor]]]) -> torch.Tensor:
    """relu-MLP(x) for x [rows, K] (CUDA fp32): one forward kernel, one backward kernel for the data-gradient chain, one tensor-core
    weight-gradient GEMM per layer that trains."""
    wb = []
    for W, b in layers:
        wb += [W, b]
    return _MlpFused.apply(x, len(layers), *wb)
