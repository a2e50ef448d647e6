#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own Python code on the CPU.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Nothing here is imported by the product.  The reference modules are imported from where they lie;
missing third-party imports of the reference (trimesh, open3d, ...) are stubbed in sys.modules, numpy.math
is aliased to math (removed in numpy 2) and the reference's CUDA extension modules are resolved to the
prebuilt oracle/_ref/*.so (import only -- no GPU is needed, no kernel is launched).

Outputs (all float32, seeded):
  ide.npz           IntegratedDirEncoder (ide_encoder/ide_encoder.py) for deg_view 4 and 5
  demo_sphere.npz   BASELINE config 1: demo.ipynb cell 17 on 64x64 rays with the shipped demo/ weights
  field_glue.npz    NeRFNetwork.forward_sigma / get_color_mlp_extra_params / forward_color of the reference
                    (nerf/network.py, nerf/renderer.py) under configs/scenes/toaster.ini (env width 64 to keep
                    the fixture small), driven through a differentiable stand-in position encoder
  relight_mlps.npz  shipped ckpts/rendering_mlps.pth + ckpts/env_ckpts/env_net_3.pth, plus reference
                    forward_color outputs on seeded inputs at those dims (env 160 / IDE deg 4)
"""
import argparse
import math
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("ENVIDR_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))


def install_shims():
    """Third-party stubs, numpy.math, configargparse and the <pkg>._ext modules: shared with the GPU-side checker (oracle/ref_model.py)."""
    if REPO not in sys.path:
        sys.path.insert(0, REPO)
    from oracle import ref_model
    ref_model.install_shims()


def sd_to_np(sd, prefix=""):
    return {prefix + k.replace(".", "_"): v.detach().cpu().numpy().astype(np.float32) for k, v in sd.items()}


def gen_ide():
    from ide_encoder.ide_encoder import IntegratedDirEncoder
    g = torch.Generator().manual_seed(0)
    out = {}
    for deg in (4, 5):
        enc = IntegratedDirEncoder(deg_view=deg)
        d = torch.nn.functional.normalize(torch.randn(64, 3, generator=g), dim=-1)
        d[0] = torch.tensor([0.0, 0.0, 1.0]); d[1] = torch.tensor([0.0, 0.0, -1.0]); d[2] = torch.tensor([1.0, 0.0, 0.0])
        d[3] = torch.tensor([0.0, 0.0, 0.6])       # un-normalised, x == y == 0 (ide_encoder.py:114-115)
        rough = torch.rand(64, 1, generator=g) * 0.3
        rough[4] = 0.0
        out[f"dirs{deg}"] = d.numpy()
        out[f"rough{deg}"] = rough.numpy()
        out[f"ide{deg}_var"] = enc(d, rough).numpy()
        out[f"ide{deg}_const"] = enc(d, 0.64).numpy()
        out[f"mat{deg}"] = enc.mat.numpy(); out[f"ml{deg}"] = enc.ml_array.numpy(); out[f"sigma{deg}"] = enc.sigma.numpy()
    np.savez_compressed(os.path.join(HERE, "ide.npz"), **out)
    print("ide.npz", {k: v.shape for k, v in out.items()})


def gen_demo():
    """demo.ipynb cells 3-17 (the reference's CPU-runnable path), W=H=64."""
    import torch.nn as nn
    import torch.nn.functional as F
    from ide_encoder.ide_encoder import IntegratedDirEncoder

    def get_net(i, o, h, n):
        net = []
        for _ in range(n - 1):
            net += [nn.Linear(i, h), nn.ReLU(inplace=True)]
            i = h
        net.append(nn.Linear(i, o))
        return nn.Sequential(*net)

    sdf_net, env_net = get_net(37, 14, 64, 3), get_net(38, 12, 160, 4)
    diffuse_net, specular_net = get_net(24, 3, 32, 2), get_net(28, 3, 64, 3)
    ld = lambda p: torch.load(os.path.join(REF, p), map_location="cpu")
    sdf_net.load_state_dict(ld("demo/sdf_net.pth")); diffuse_net.load_state_dict(ld("demo/diffuse_net.pth"))
    specular_net.load_state_dict(ld("demo/specular_net.pth")); env_net.load_state_dict(ld("demo/envs/env_net_2.pth"))
    xyz_encoding = torch.from_numpy(np.loadtxt(os.path.join(REF, "demo/xyz_encoding.txt"))).float()
    encoder_dir = IntegratedDirEncoder(deg_view=4)

    W = H = 64
    theta, phi, radius = 123.0, 0.0, 4.0
    roughness, metallic, base_color = 0.0, 0.2, [20 / 255., 70 / 255., 160 / 255.]
    cam = 0.6194058656692505
    focal = W / (2 * np.tan(cam / 2))
    trans_t = lambda t: np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, t], [0, 0, 0, 1]])
    rot_phi = lambda p: np.array([[1, 0, 0, 0], [0, np.cos(p), -np.sin(p), 0], [0, np.sin(p), np.cos(p), 0], [0, 0, 0, 1]])
    rot_theta = lambda t: np.array([[np.cos(t), 0, -np.sin(t), 0], [0, 1, 0, 0], [np.sin(t), 0, np.cos(t), 0], [0, 0, 0, 1]])
    c2w = trans_t(radius); c2w = rot_phi(-phi / 180. * np.pi) @ c2w; c2w = rot_theta(theta / 180. * np.pi) @ c2w
    c2w = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]) @ c2w
    p = c2w
    pose = np.array([[p[1, 0], -p[1, 1], -p[1, 2], p[1, 3]], [p[2, 0], -p[2, 1], -p[2, 2], p[2, 3]],
                     [p[0, 0], -p[0, 1], -p[0, 2], p[0, 3]], [0, 0, 0, 1]], dtype=np.float32)
    pose = torch.from_numpy(pose)[None]
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing="ij")
    i = i.t().reshape(1, H * W) + 0.5; j = j.t().reshape(1, H * W) + 0.5
    zs = torch.ones_like(i)
    dirs = torch.stack(((i - W / 2) / focal * zs, (j - H / 2) / focal * zs, zs), -1)
    dirs = dirs / torch.norm(dirs, dim=-1, keepdim=True)
    rays_d = (dirs @ pose[:, :3, :3].transpose(-1, -2)).view(-1, 3)
    rays_o = pose[..., :3, 3][..., None, :].expand(1, H * W, 3).reshape(-1, 3)
    dot = torch.bmm(rays_d.view(-1, 1, 3), rays_o.view(-1, 3, 1)).squeeze(-1)
    nabla = dot ** 2 - (rays_o.norm(2, 1, keepdim=True) ** 2 - 1.0)
    near = -dot - torch.sqrt(nabla.clamp_min(0.0))
    mask = (nabla >= -1e-4)[..., 0]
    d = rays_d[mask]; xyz = rays_o[mask] + d * near[mask]; normals = xyz
    with torch.no_grad():
        h = torch.cat([xyz_encoding, torch.tensor([roughness, metallic, *base_color])])[None]
        h = sdf_net(h)
        geo = F.normalize(h[..., 1:13], dim=-1).repeat(xyz.shape[0], 1)
        kappa_inv = 1.0 * F.softplus(h[..., -1] - 1)[0]
        n_enc = encoder_dir(normals, 0.64)
        w_o = -d
        w_r = 2 * torch.sum(w_o * normals, -1, keepdim=True) * normals - w_o
        w_r_enc = encoder_dir(w_r, kappa_inv)
        n_dot_v = torch.sum(normals * w_o, -1, keepdim=True)
        f_n = F.normalize(env_net(n_enc), dim=-1)
        c_d = diffuse_net(torch.cat([geo, f_n], -1)).sigmoid()
        f_r = F.normalize(env_net(w_r_enc), dim=-1)
        c_s = specular_net(torch.cat([geo, normals, f_r, n_dot_v], -1)).sigmoid()
    out = dict(rays_o=rays_o.numpy(), rays_d=rays_d.numpy(), mask=mask.numpy(), xyz=xyz.numpy(), dirs=d.numpy(),
               sdf_out=h.numpy(), geo=geo[0].numpy(), kappa_inv=np.float32(kappa_inv.item()),
               diffuse=c_d.numpy(), specular=c_s.numpy(), xyz_encoding=xyz_encoding.numpy(),
               material=np.array([roughness, metallic, *base_color], np.float32))
    out.update(sd_to_np(sdf_net.state_dict(), "sdf_")); out.update(sd_to_np(env_net.state_dict(), "env_"))
    out.update(sd_to_np(diffuse_net.state_dict(), "diffuse_")); out.update(sd_to_np(specular_net.state_dict(), "specular_"))
    np.savez_compressed(os.path.join(HERE, "demo_sphere.npz"), **out)
    print("demo_sphere.npz: hit rays", int(mask.sum()), "of", H * W)


class StandInEncoder(torch.nn.Module):
    """Differentiable CPU stand-in for the (CUDA-only) hash encoder: enc = sin(xyz @ A + b)."""
    def __init__(self, A, b):
        super().__init__()
        self.A, self.b = A, b

    def forward(self, xyz, bound=1):
        return torch.sin(xyz @ self.A + self.b)


def build_model(extra_args, cuda_ray=False):
    from nerf.options import config_parser
    argv = sys.argv
    sys.argv = ["x", "--config", os.path.join(REF, "configs/scenes/toaster.ini")] + extra_args
    opt = config_parser()
    sys.argv = argv
    opt.cuda_ray = cuda_ray                 # density-grid buffers only for gen_density_grid
    from nerf.network import NeRFNetwork
    torch.manual_seed(0)
    model = NeRFNetwork(encoding="hashgrid", encoding_dir=opt.encoding_dir, bound=opt.bound, cuda_ray=cuda_ray,
                        density_scale=1, min_near=opt.min_near, density_thresh=opt.density_thresh, bg_radius=opt.bg_radius,
                        use_sdf=opt.use_sdf, hidden_dim=opt.hidden_dim, num_layers=opt.num_layers,
                        num_layers_color=opt.num_layers_color, hidden_dim_color=opt.hidden_dim_color,
                        num_layers_bg=opt.num_layers_bg, num_levels=opt.num_levels, geo_feat_dim=opt.geo_feat_dim,
                        opt=opt, env_opt=None)
    return model, opt


def run_field(model, opt, xyz, dirs, r_images=None, env_rot=None):
    xyz = xyz.clone().requires_grad_(True)
    sdfs, sigmas, geo, normals, _ = model.forward_sigma(xyz, use_sdf_sigma_grad=True, dirs=dirs, dists=None)
    roughness = model.roughness
    n_enc, w_r_enc, n_dot, n_env_enc = model.get_color_mlp_extra_params(normals, dirs, roughness, env_rot)
    rgb = model.forward_color(geo, dirs, n_enc, w_r_enc, n_dot, True, n_env_enc=n_env_enc, r_images=r_images, roughness=roughness)
    f = lambda t: t.detach().numpy().astype(np.float32)
    blend = getattr(model, "blend_weight", None)
    return dict(sdf=f(sdfs), sigma=f(sigmas), geo=f(geo), normal=f(normals), roughness=f(roughness), rgb=f(rgb),
                c_diffuse=f(model.c_diffuse), c_specular=f(model.c_specular), w_r_enc=f(w_r_enc), n_env_enc=f(n_env_enc),
                n_dot=f(n_dot), blend=f(blend) if blend is not None else np.zeros(0, np.float32))


def model_weights(model):
    out = {}
    for name in ["sdf_net", "env_net", "diffuse_net", "color_net", "renv_net"]:
        net = getattr(model, name, None)
        if net is None:
            continue
        for i, lin in enumerate(net):
            out[f"{name}_{i}_weight"] = lin.weight.detach().numpy().astype(np.float32)
            out[f"{name}_{i}_bias"] = lin.bias.detach().numpy().astype(np.float32)
    out["beta"] = np.float32(model.sdf_density.beta.item())
    return out


def gen_field_glue():
    model, opt = build_model(["--hidden_dim_env", "64"])
    g = torch.Generator().manual_seed(1)
    M = 96
    A = torch.randn(3, model.in_dim, generator=g) * 1.5
    b = torch.rand(model.in_dim, generator=g) * 6.28
    model.encoder = StandInEncoder(A, b)
    # xavier init gives near-zero SDF head; scale it so sigma spans both signs of the Laplace CDF
    with torch.no_grad():
        model.sdf_net[-1].weight[0] *= 0.2
        model.sdf_density.beta.fill_(0.05)
    xyz = torch.rand(M, 3, generator=g) * 1.6 - 0.8
    dirs = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1)
    r_images = torch.rand(1, M, 4, generator=g)
    r_images[0, ::3, 3] = 0.95 + 0.05 * r_images[0, ::3, 3]          # some visible reflections (vis > 0.9)
    out = dict(A=A.numpy(), b=b.numpy(), xyz=xyz.numpy(), dirs=dirs.numpy(), r_images=r_images[0].numpy())
    out.update(model_weights(model))
    for tag, kw in [("plain", {}), ("rot", dict(env_rot=0.7)), ("renv", dict(r_images=r_images[0]))]:
        # the renv branch indexes r_images[renv_mask] on the sample axis -> pass [M,4]
        res = run_field(model, opt, xyz, dirs, **kw)
        out.update({f"{tag}_{k}": v for k, v in res.items()})
    out["opt_ide_degree"] = np.int32(opt.sh_degree)
    out["opt_roughness_act_scale"] = np.float32(opt.roughness_act_scale)
    out["opt_indir_roughness_thresh"] = np.float32(opt.indir_roughness_thresh)
    out["opt_learn_indir_blend"] = np.int32(opt.learn_indir_blend)
    out["opt_beta_min"] = np.float32(opt.beta_min); out["opt_beta_max"] = np.float32(opt.beta_max)
    np.savez_compressed(os.path.join(HERE, "field_glue.npz"), **out)
    print("field_glue.npz: renv-masked samples",
          int(((out["renv_roughness"][:, 0] < 0.1) & (out["r_images"][:, 3] > 0.9)).sum()), "of", M)


def gen_relight():
    model, opt = build_model(["--hidden_dim_env", "160", "--sh_degree", "4"])
    sd = torch.load(os.path.join(REF, "ckpts/rendering_mlps.pth"), map_location="cpu")["model"]
    env = torch.load(os.path.join(REF, "ckpts/env_ckpts/env_net_3.pth"), map_location="cpu")["model"]
    # shipped env checkpoints spell keys 'env_net0.weight' (nerf/sph_loader.py:372) -> 'env_net.0.weight'
    env = {k.replace("env_net", "env_net."): v for k, v in env.items()}
    missing = model.load_state_dict({**sd, **env}, strict=False)
    assert not [k for k in missing.unexpected_keys], missing.unexpected_keys
    g = torch.Generator().manual_seed(2)
    M = 96
    A = torch.randn(3, model.in_dim, generator=g) * 1.5
    b = torch.rand(model.in_dim, generator=g) * 6.28
    model.encoder = StandInEncoder(A, b)
    with torch.no_grad():
        model.sdf_net[-1].weight[0] *= 0.2
        model.sdf_density.beta.fill_(0.05)
    xyz = torch.rand(M, 3, generator=g) * 1.6 - 0.8
    dirs = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1)
    out = dict(A=A.numpy(), b=b.numpy(), xyz=xyz.numpy(), dirs=dirs.numpy())
    out.update(model_weights(model))
    res = run_field(model, opt, xyz, dirs)
    out.update({f"plain_{k}": v for k, v in res.items()})
    out["opt_ide_degree"] = np.int32(4)
    np.savez_compressed(os.path.join(HERE, "relight_mlps.npz"), **out)
    print("relight_mlps.npz written")


def gen_density_grid():
    """NeRFRenderer.update_extra_state / mark_untrained_grid of the reference (nerf/renderer.py:200-352) on the CPU.
    The three CUDA-only helpers it calls (raymarching.morton3D / morton3D_invert / packbits) are bound to the oracle's C
    restatements (pinned bit-exact against the reference kernels on the GPU by tests/test_gpu_ops.py); the hash encoder is the
    stand-in encoder.  Sequence: 2 full updates (iter_density 0, 1), then one partial update (iter_density forced to 16), all from
    torch.manual_seed(3) -- the test replays the same torch-CPU random calls.  Stored: the grid at every 16th cell after each
    stage, mean_density, the full bit fields; the partial update's coords and noise (they depend on the grid values)."""
    sys.path.insert(0, REPO)
    from oracle import oracle as O
    import nerf.renderer as R
    model, opt = build_model(["--hidden_dim_env", "64"], cuda_ray=True)
    g = torch.Generator().manual_seed(4)
    A = torch.randn(3, model.in_dim, generator=g) * 1.5
    b = torch.rand(model.in_dim, generator=g) * 6.28
    model.encoder = StandInEncoder(A, b)
    with torch.no_grad():
        model.sdf_net[-1].weight[0] *= 0.2
        model.sdf_net[-1].bias[0] += 0.05
        model.sdf_density.beta.fill_(0.05)
    R.raymarching.morton3D = lambda c: torch.from_numpy(O.morton3D(c.numpy()))
    R.raymarching.morton3D_invert = lambda i: torch.from_numpy(O.morton3D_invert(i.numpy().astype(np.int32)))

    def packbits(grid, thresh, bitfield=None):
        return torch.from_numpy(O.packbits(grid.detach().numpy().reshape(-1), float(thresh)))
    R.raymarching.packbits = packbits
    out = dict(A=A.numpy(), b=b.numpy(), seed=np.int32(3), cascade=np.int32(model.cascade), bound=np.float32(model.bound),
               density_thresh=np.float32(model.density_thresh), density_scale=np.float32(model.density_scale))
    out.update(model_weights(model))
    out["opt_beta_min"] = np.float32(opt.beta_min); out["opt_beta_max"] = np.float32(opt.beta_max)
    # mark_untrained_grid with 6 cameras on a ring
    sys.path.insert(0, REPO)
    from envidr_b200 import scene
    poses = np.stack([scene.nerf_matrix_to_ngp(scene.pose_spherical(th, -30.0, 4.0), scale=0.65) for th in range(0, 360, 60)])
    intr = scene.intrinsics_from_fov(800, 800, 0.35)                # narrow field of view: part of the grid is never seen
    model.mark_untrained_grid(poses, intr)
    out["poses"] = poses.astype(np.float32); out["intrinsic"] = np.asarray(intr, np.float64)
    out["marked"] = np.packbits((model.density_grid.numpy().reshape(-1) < 0))
    torch.manual_seed(3)
    sub = slice(None, None, 16)
    for stage in range(2):
        model.update_extra_state()
        out[f"full{stage}_grid_sub"] = model.density_grid.numpy().reshape(-1)[sub].copy()
        out[f"full{stage}_mean"] = np.float64(model.mean_density)
        out[f"full{stage}_bits"] = model.density_bitfield.numpy().copy()
    model.iter_density = 16
    # record the random draws of the partial update by replaying them from the same generator state afterwards
    st = torch.get_rng_state()
    model.update_extra_state()
    out["part_grid_sub"] = model.density_grid.numpy().reshape(-1)[sub].copy()
    out["part_mean"] = np.float64(model.mean_density)
    out["part_bits"] = model.density_bitfield.numpy().copy()
    out["part_rng_state"] = st.numpy()
    np.savez_compressed(os.path.join(HERE, "density_grid.npz"), **out)
    print("density_grid.npz: marked", int((model.density_grid < 0).sum()), "means", out["full0_mean"], out["full1_mean"], out["part_mean"],
          "occupied", int(np.unpackbits(out["part_bits"]).sum()))


def gen_train_epilogue():
    """SURVEY 8 f-2 -- the reference's own code for the steps either side of the training branch, run on the CPU:
      * nerf.utils.get_rays (utils.py:110-209) for random pixel indices and for the full image;
      * the auxiliary block of run_cuda (cuda_ray.py:173-211), executed from the reference's source lines where they lie
        (the block has no function boundary) on a seeded sample buffer;
      * Trainer.train_step's loss terms (utils.py:560-808) with `self` stubbed: model.render returns the prepared outputs, so
        the reference computes colour L1 + mask BCE + back-sdf + Cauchy + eikonal and autograd gives the gradients.
    -> tests/golden/train_epilogue.npz"""
    import textwrap
    import nerf.utils as U
    from nerf.options import config_parser
    out = {}
    # ---- get_rays -------------------------------------------------------------------------------------------
    sys.path.insert(0, REPO)
    from envidr_b200 import scene
    poses = np.stack([scene.nerf_matrix_to_ngp(scene.pose_spherical(th, -30.0, 4.0), scale=0.65) for th in (10.0, 200.0)])
    H, W = 37, 53
    intr = scene.intrinsics_from_fov(W, H, 0.69)
    torch.manual_seed(7)
    r = U.get_rays(torch.from_numpy(poses), intr, H, W, N=300)
    out.update(rays_poses=poses.astype(np.float32), rays_intrinsics=np.asarray(intr, np.float64), rays_HW=np.array([H, W]),
               rays_inds=r["inds"][0].numpy(), rays_o=r["rays_o"].numpy(), rays_d=r["rays_d"].numpy())
    r = U.get_rays(torch.from_numpy(poses[:1]), intr, H, W, N=-1)
    out.update(rays_full_o=r["rays_o"].numpy(), rays_full_d=r["rays_d"].numpy())
    # ---- auxiliary block + losses -----------------------------------------------------------------------------
    argv = sys.argv
    sys.argv = ["x", "--config", os.path.join(REF, "configs/scenes/toaster.ini")]
    opt = config_parser()
    sys.argv = argv
    g = torch.Generator().manual_seed(8)
    N, M = 96, 1200
    cnt = torch.randint(0, 24, (N,), generator=g); cnt[5] = 0
    off = torch.cumsum(cnt, 0) - cnt
    total = int(cnt.sum())
    assert total < M
    rays = torch.stack([torch.arange(N), off, cnt], -1).int()
    deltas = torch.zeros(M, 2)
    deltas[:total, 0] = 2 * 3 ** 0.5 / 1024
    deltas[:total, 1] = deltas[:total, 0] * (1 + (torch.rand(total, generator=g) < 0.15).float() * 3)     # some gaps (empty cells skipped)
    sdfs = (torch.randn(M, generator=g) * 0.03).requires_grad_(True)
    sigmas = torch.rand(M, generator=g)
    weights = torch.rand(M, generator=g) * 0.04
    dirs = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1)
    normals = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1)
    xyzs = torch.rand(M, 3, generator=g) * 2 - 1
    src = open(os.path.join(REF, "nerf/render_func/cuda_ray.py")).read().splitlines()
    assert src[172].strip().startswith("if main_pass and (use_relsdf_loss or use_backsdf_loss or use_orientation_loss):"), src[172]
    block = textwrap.dedent("\n".join(src[172:212]))

    class _S:
        obj_aabb = None
        use_sdf = True
    results = {"sdfs": sdfs}
    ns = dict(torch=torch, self=_S(), main_pass=True, use_relsdf_loss=False, use_backsdf_loss=True, use_orientation_loss=False,
              sigmas=sigmas, rays=rays, deltas=deltas, xyzs=xyzs, sdfs=sdfs, dirs=dirs, normals=normals, weights=weights, results=results)
    exec(compile(block, "cuda_ray.py[173:212]", "exec"), ns)
    image = torch.rand(1, N, 3, generator=g).requires_grad_(True)
    weights_sum = torch.rand(1, N, generator=g)
    weights_sum[0, :4] = torch.tensor([0.0, 1.0, 5e-4, 0.9995])                                            # outside the BCE clip
    weights_sum.requires_grad_(True)
    sdf_gradients = (torch.randn(M, 3, generator=g) * 0.7).requires_grad_(True)
    with torch.no_grad():
        sdf_gradients[3] = 0                                                                              # norm at the origin
    results.update(image=image, weights_sum=weights_sum, sdf_gradients=sdf_gradients, sigmas=sigmas)
    images = torch.rand(1, N, 4, generator=g)
    images[..., 3] = (images[..., 3] > 0.4).float()
    beta = 0.02

    class _Density:
        def get_beta(self):
            return torch.tensor(beta)

        def density_func(self, sdf, beta=None, alpha=None):
            a = 1 / beta if alpha is None else alpha
            return a * (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() / beta))                           # network.py:32-37

    class _Model:
        bg_radius = -1
        sdf_density = _Density()

        def render(self, rays_o, rays_d, **kw):
            return results

    class _Self:
        pass
    fake = _Self()
    opt.color_space = "srgb"                      # keep gt as given (the srgb->linear conversion is data preparation)
    opt.alpha_bg_mode = "white"
    opt.eikonal_loss = True; opt.cauchy_loss = True; opt.backsdf_loss = True; opt.mask_loss = True
    opt.relsdf_loss = False; opt.orientation_loss = False; opt.dist_bound = False; opt.diffuse_loss = False
    opt.entropy_loss_weight = 0; opt.env_sph_mode = False; opt.cauchy_roughness_weighted = False
    fake.opt, fake.model, fake.device, fake.error_map = opt, _Model(), torch.device("cpu"), None
    fake.criterion = torch.nn.L1Loss(reduction="none")
    fake.epoch, fake.global_step = 500, 500
    data = dict(rays_o=torch.zeros(1, N, 3), rays_d=torch.zeros(1, N, 3), images=images.clone())
    pred, gt_rgb, loss, ld = U.Trainer.train_step(fake, data)
    gi, gw, gs, gg = torch.autograd.grad(loss, [image, weights_sum, sdfs, sdf_gradients])
    f = lambda t: t.detach().numpy().astype(np.float32)
    out.update(loss_rays=rays.numpy(), loss_deltas=f(deltas), loss_sdfs=f(sdfs), loss_weights=f(weights), loss_image=f(image[0]),
               loss_weights_sum=f(weights_sum[0]), loss_sdf_gradients=f(sdf_gradients), loss_gt_rgb=f(gt_rgb[0]), loss_gt_mask=f(images[0, :, 3]),
               loss_beta=np.float32(beta), loss_total=np.float64(loss.item()),
               loss_w=np.array([opt.color_loss_weight, opt.mask_loss_weight, opt.cauchy_loss_weight, opt.eikonal_loss_weight,
                                opt.backsdf_loss_weight, opt.backsdf_thresh], np.float64),
               loss_backsdf_mean=np.int32(opt.backsdf_mode != "sum"),
               **{f"loss_term_{k}": np.float64(v.item()) for k, v in ld.items()},
               grad_image=f(gi[0]), grad_weights_sum=f(gw[0]), grad_sdfs=f(gs), grad_sdf_gradients=f(gg),
               aux_point_count=np.int32(results["sdfs"].shape[0]), aux_relsdf=f(results["relsdf"]), aux_sdf_dist=f(results["sdf_dist"]))
    np.savez_compressed(os.path.join(HERE, "train_epilogue.npz"), **out)
    print("train_epilogue.npz:", {k: float(v) for k, v in ld.items()}, "total", loss.item(), "aux points", int(results["sdfs"].shape[0]), "of", M)


def gen_state_dict_keys():
    """Key names and shapes of the reference NeRFNetwork.state_dict() under configs/scenes/toaster.ini (cuda_ray on) -> the
    contract envidr_b200.checkpoint.state_from_field / field_from_state must meet.  -> tests/golden/state_dict_keys.json"""
    import json
    model, opt = build_model([], cuda_ray=True)
    keys = {k: list(v.shape) for k, v in model.state_dict().items()}
    json.dump(dict(keys=keys, per_level_scale=float(model.encoder.per_level_scale), base_resolution=int(model.encoder.base_resolution),
                   beta_min=float(opt.beta_min), beta_max=float(opt.beta_max)), open(os.path.join(HERE, "state_dict_keys.json"), "w"), indent=1)
    print("state_dict_keys.json:", len(keys), "keys")


def gen_neus():
    """NeuSDensity of the reference (nerf/network.py:46-102) on seeded inputs, with autograd gradients -> tests/golden/neus.npz"""
    from nerf.network import NeuSDensity
    g = torch.Generator().manual_seed(5)
    M = 512
    out = {}
    for tag, ratio, with_grad, tensor_dist, var in (("a", 1.0, True, False, 0.3), ("b", 0.3, True, True, 0.45), ("c", 1.0, False, False, 0.3),
                                                    ("d", 0.0, True, True, 1.5)):
        mod = NeuSDensity(var)
        sdf = (torch.randn(M, generator=g) * 0.02).requires_grad_(True)
        dirs = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1)
        grads = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1) * (0.8 + 0.4 * torch.rand(M, 1, generator=g))
        grads.requires_grad_(True)
        dists = (torch.rand(M, generator=g) * 0.004 + 0.002) if tensor_dist else mod.base_dist
        ga = torch.randn(M, generator=g)
        alpha = mod(sdf, dirs, dists, grads if with_grad else None, cos_anneal_ratio=ratio)
        leaves = [sdf, mod.variance] + ([grads] if with_grad else [])
        gr = torch.autograd.grad((alpha * ga).sum(), leaves)
        f = lambda t: t.detach().numpy().astype(np.float32)
        out.update({f"{tag}_sdf": f(sdf), f"{tag}_dirs": f(dirs), f"{tag}_grads": f(grads), f"{tag}_dists": f(dists) if tensor_dist else np.float32(dists),
                    f"{tag}_ga": f(ga), f"{tag}_alpha": f(alpha), f"{tag}_g_sdf": f(gr[0]), f"{tag}_g_var": np.float64(gr[1].item()),
                    f"{tag}_g_grads": f(gr[2]) if with_grad else np.zeros(0, np.float32), f"{tag}_ratio": np.float32(ratio),
                    f"{tag}_var": np.float32(var), f"{tag}_with_grad": np.int32(with_grad)})
    np.savez_compressed(os.path.join(HERE, "neus.npz"), **out)
    print("neus.npz: alpha ranges", {t: (float(out[f"{t}_alpha"].min()), float(out[f"{t}_alpha"].max())) for t in "abcd"})


def gen_train_branch():
    """The reference's OWN run_cuda training branch (nerf/render_func/cuda_ray.py:64-168) through NeRFRenderer.render and
    NeRFNetwork.forward_sigma / compute_normal / get_color_mlp_extra_params / forward_color, run on the CPU: the CUDA-only operators it
    calls are bound to the oracle's restatements (raymarching.near_far_from_aabb / march_rays_train / composite_rays_train /
    get_scatter_idx -> oracle C + its autograd wrapper; HashEncoder.forward -> the oracle's hash_encode with second-order backward, on
    the reference module's own embeddings / offsets / per_level_scale).  Two cases: single pass, and main pass with r_images (renv
    branch).  Stored: weights, rays, outputs and the gradients of a fixed functional of the outputs.  -> tests/golden/train_branch.npz"""
    sys.path.insert(0, REPO)
    from envidr_b200 import scene
    from oracle import oracle as O
    from oracle import train_oracle as TO
    import nerf.render_func.cuda_ray as CR
    model, opt = build_model(["--hidden_dim_env", "64", "--num_levels", "8", "--log2_hashmap_size", "12", "--desired_resolution", "256",
                              "--max_steps", "256"], cuda_ray=True)
    g = torch.Generator().manual_seed(6)
    enc = model.encoder
    with torch.no_grad():
        enc.embeddings.copy_(torch.rand(enc.embeddings.shape, generator=g) - 0.5)
        model.sdf_density.beta.fill_(0.05)
    offs = enc.offsets.numpy().astype(np.int32)

    def enc_forward(inputs, bound=1):                                      # hashgrid.py:157-168
        x01 = ((inputs + bound) / (2 * bound)).view(-1, 3)
        return TO.hash_encode(x01, enc.embeddings, offs, float(enc.per_level_scale), int(enc.base_resolution), x01.requires_grad)
    enc.forward = enc_forward

    class RM:
        @staticmethod
        def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
            n, f = O.near_far_from_aabb(rays_o.detach().numpy(), rays_d.detach().numpy(), aabb.numpy(), min_near)
            return torch.from_numpy(n), torch.from_numpy(f)

        @staticmethod
        def march_rays_train(rays_o, rays_d, bound, bitfield, C, H, nears, fars, step_counter=None, mean_count=-1, perturb=False, align=-1,
                             force_all_rays=False, dt_gamma=0, max_steps=1024, early_stop_steps=-1):
            assert not perturb and (force_all_rays or mean_count <= 0)
            N = rays_o.shape[0]
            x, d, dl, rays, cnt = O.march_rays_train(rays_o.detach().numpy(), rays_d.detach().numpy(), bound, bitfield.numpy(), C, H, nears.numpy(),
                                                     fars.numpy(), N * max_steps, dt_gamma=dt_gamma, max_steps=max_steps,
                                                     early_stop_steps=early_stop_steps)
            m = int(cnt[0])
            if align > 0:
                m += align - m % align                                     # raymarching.py:235-241
            return torch.from_numpy(x[:m]), torch.from_numpy(d[:m]), torch.from_numpy(dl[:m]), torch.from_numpy(rays)

        @staticmethod
        def composite_rays_train(sigmas, rgbs, deltas, rays, T_thresh=1e-4, ret_weights=False, input_alpha=False):
            return TO.composite_rays_train(sigmas, rgbs, deltas.numpy(), rays.numpy(), T_thresh, ret_weights, input_alpha)

        @staticmethod
        def get_scatter_idx(rays, dummy):
            return torch.from_numpy(O.get_scatter_idx(rays.numpy(), dummy.shape[0]))
    CR.raymarching = RM
    bf = scene.make_bitfield()
    model.density_bitfield = torch.from_numpy(bf)
    model.train()
    opt.indir_ref = False
    opt.eikonal_loss = True
    opt.backsdf_loss = False; opt.relsdf_loss = False; opt.orientation_loss = False; opt.dist_bound = False
    ro, rd = scene.camera_rays(12, 12)
    N = ro.shape[0]
    r_img = torch.rand(1, N, 4, generator=g)
    r_img[0, ::2, 3] = 0.95 + 0.05 * r_img[0, ::2, 3]
    out = dict(rays_o=ro.numpy(), rays_d=rd.numpy(), r_images=r_img[0].numpy(), offsets=offs, per_level_scale=np.float64(enc.per_level_scale),
               base_resolution=np.int32(enc.base_resolution), embeddings=enc.embeddings.detach().numpy().copy())
    out.update(model_weights(model))
    out["opt_beta_min"] = np.float32(opt.beta_min); out["opt_beta_max"] = np.float32(opt.beta_max)
    kw = {k: v for k, v in vars(opt).items()}
    names = ["sdf_net.0.weight", "sdf_net.2.bias", "env_net.1.weight", "renv_net.0.weight", "encoder.embeddings", "sdf_density.beta"]
    params = dict(model.named_parameters())
    for tag, ri in (("single", None), ("renv", r_img)):
        model.zero_grad()
        res = model.render(ro[None], rd[None], staged=False, bg_color=1, perturb=False, force_all_rays=True, get_normal_image=False,
                           r_images=ri, **kw)
        sdfs, sg = res["sdfs"], res["sdf_gradients"]
        loss = 0.3 * res["image"].sum() + (res["weights_sum"] ** 2).sum() + 5 * (sdfs ** 2).mean() + ((sg.norm(dim=-1) - 1) ** 2).mean()
        loss.backward()
        f = lambda t: t.detach().numpy().astype(np.float32)
        out.update({f"{tag}_image": f(res["image"][0]), f"{tag}_weights_sum": f(res["weights_sum"][0]), f"{tag}_depth": f(res["depth"][0]),
                    f"{tag}_sdfs": f(sdfs), f"{tag}_sdf_gradients": f(sg), f"{tag}_sigmas": f(res["sigmas"]), f"{tag}_loss": np.float64(loss.item())})
        for n in names:
            gr = params[n].grad
            out[f"{tag}_grad_{n.replace('.', '_')}"] = np.zeros(0, np.float32) if gr is None else f(gr)
    np.savez_compressed(os.path.join(HERE, "train_branch.npz"), **out)
    print("train_branch.npz: samples", out["single_sdfs"].shape[0], "loss", float(out["single_loss"]), float(out["renv_loss"]),
          "ws mean", float(out["single_weights_sum"].mean()))


def gen_infer_branch():
    """The reference's OWN inference path on the CPU: NeRFRenderer.render (renderer.py:364-531: single pass and the three-pass
    indirect-reflection frame) -> run_cuda inference while-loop (cuda_ray.py:238-359, get_normal_image, visual items) -> NeRFNetwork
    methods, with raymarching.near_far_from_aabb / march_rays / composite_rays bound to the oracle's C restatements (bit-exact against
    the reference kernels on the GPU, tests/test_gpu_ops.py) and HashEncoder.forward to the oracle's hash_encode.
    -> tests/golden/infer_branch.npz (weights, rays, all returned images of both frames)"""
    sys.path.insert(0, REPO)
    from envidr_b200 import scene
    from oracle import oracle as O
    from oracle import train_oracle as TO
    import nerf.render_func.cuda_ray as CR
    model, opt = build_model(["--hidden_dim_env", "64", "--num_levels", "8", "--log2_hashmap_size", "12", "--desired_resolution", "256",
                              "--max_steps", "256", "--indir_max_steps", "256"], cuda_ray=True)
    g = torch.Generator().manual_seed(7)
    enc = model.encoder
    # an SDF-like field: the analytic toaster through the synthetic-field construction, restricted to this small grid
    fp = scene.make_synthetic_field(0, hidden_dim_env=64, ide_degree=5, num_levels=8, log2_hashmap_size=12, desired_resolution=256, beta=0.02)
    assert tuple(fp.embeddings.shape) == tuple(enc.embeddings.shape) and np.array_equal(fp.offsets.numpy(), enc.offsets.numpy())
    with torch.no_grad():
        enc.embeddings.copy_(fp.embeddings)
        for name in ("sdf", "env", "diffuse", "color", "renv"):
            for lin, (W, b) in zip(getattr(model, name + "_net"), getattr(fp, name)):
                lin.weight.copy_(W); lin.bias.copy_(b)
        model.sdf_density.beta.fill_(0.02)
    offs = enc.offsets.numpy().astype(np.int32)

    def enc_forward(inputs, bound=1):
        x01 = ((inputs + bound) / (2 * bound)).view(-1, 3)
        return TO.hash_encode(x01, enc.embeddings, offs, float(enc.per_level_scale), int(enc.base_resolution), x01.requires_grad)
    enc.forward = enc_forward

    class RM:
        @staticmethod
        def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
            n, f = O.near_far_from_aabb(rays_o.detach().numpy(), rays_d.detach().numpy(), aabb.numpy(), min_near)
            return torch.from_numpy(n), torch.from_numpy(f)

        @staticmethod
        def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, C, H, near, far, align=-1, perturb=False,
                       dt_gamma=0, max_steps=1024):
            assert not perturb
            x, d, dl, _ = O.march_rays(n_alive, n_step, rays_alive.numpy(), rays_t.numpy(), rays_o.detach().numpy(), rays_d.detach().numpy(), bound,
                                       bitfield.numpy(), C, H, near.numpy(), far.numpy(), align=align, dt_gamma=dt_gamma, max_steps=max_steps)
            return torch.from_numpy(x), torch.from_numpy(d), torch.from_numpy(dl)

        @staticmethod
        def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2,
                           input_alpha=False, accum_deltas=True):
            O.composite_rays(n_alive, n_step, rays_alive.numpy(), rays_t.numpy(), sigmas.detach().numpy(), rgbs.detach().numpy(),
                             deltas.detach().numpy(), weights_sum.numpy(), depth.numpy(), image.numpy(), T_thresh=T_thresh,
                             input_alpha=bool(input_alpha), accum_deltas=bool(accum_deltas))      # in place through the shared memory
            return tuple()
    CR.raymarching = RM
    bf = scene.make_bitfield()
    model.density_bitfield = torch.from_numpy(bf)
    model.eval()
    W = 20
    ro, rd = scene.camera_rays(W, W)
    out = dict(rays_o=ro.numpy(), rays_d=rd.numpy(), offsets=offs, per_level_scale=np.float64(enc.per_level_scale),
               base_resolution=np.int32(enc.base_resolution), embeddings=enc.embeddings.detach().numpy().copy())
    out.update(model_weights(model))
    out["opt_beta_min"] = np.float32(opt.beta_min); out["opt_beta_max"] = np.float32(opt.beta_max)
    for tag, indir, rot in (("one", False, None), ("three", True, None), ("rot", True, 0.7)):      # rot: one step of the relight sweep
        opt.indir_ref = indir
        kw = {k: v for k, v in vars(opt).items()}
        res = model.render(ro[None], rd[None], staged=True, bg_color=1, perturb=False, get_normal_image=True, env_rot_radian=rot, **kw)
        f = lambda t: t.detach().numpy().astype(np.float32)
        for k in ("image", "depth", "weights_sum", "normal_image", "diffuse_image", "specular_image", "roughness_image"):
            if k in res and res[k] is not None:
                out[f"{tag}_{k}"] = f(res[k]).reshape(W * W, -1)
    np.savez_compressed(os.path.join(HERE, "infer_branch.npz"), **out)
    print("infer_branch.npz: hit pixels", int((out["one_weights_sum"] > 0.5).sum()), "of", W * W, "keys", sorted(k for k in out if k.startswith("three_")))


def gen_neus_field():
    """BASELINE config 4 -- the reference's OWN NeRFNetwork built with use_neus_sdf / encoding_pos=frequency / geometric_init /
    8 x 256 layers / skip_layers [4] (weight-normed layers, Softplus(beta=100); network.py:154-222, 415-448), run on the CPU:
    forward_sigma (normals by autograd, NeuSDensity alpha) + get_color_mlp_extra_params + forward_color on seeded samples, and
    NeRFRenderer.render -> run_cuda inference loop (input_alpha compositing) on a 16 x 16 frame with raymarching bound to the oracle's
    C restatements.  The CUDA-only FreqEncoder is replaced by the torch formula of its kernel (oracle.neus_oracle.freq_encode, checked
    against the C restatement of freqencoder.cu in tests/test_oracle_golden.py).  -> tests/golden/neus_field.npz"""
    sys.path.insert(0, REPO)
    from envidr_b200 import scene
    from oracle import oracle as O
    from oracle import neus_oracle as NO
    import nerf.render_func.cuda_ray as CR
    flags = ["--use_neus_sdf", "--encoding_pos", "frequency", "--multires", "6", "--geometric_init", "--num_layers", "8", "--hidden_dim", "256",
             "--skip_layers", "4", "--init_variance", "0.6", "--geo_init_bias", "0.5", "--hidden_dim_env", "64", "--sh_degree", "4",
             "--max_steps", "256"]
    model, opt = build_model(flags, cuda_ray=True)
    nf = scene.make_neus_field(0, hidden_dim_env=64, ide_degree=4)
    with torch.no_grad():
        for lin, (W, b) in zip(model.sdf_net, nf.sdf):
            assert tuple(lin.weight_v.shape) == tuple(W.shape), (lin.weight_v.shape, W.shape)
            lin.weight_v.copy_(W); lin.weight_g.copy_(W.norm(dim=1, keepdim=True)); lin.bias.copy_(b)
        sh = nf.shading
        for name in ("env", "diffuse", "color", "renv"):
            for lin, (W, b) in zip(getattr(model, name + "_net"), getattr(sh, name)):
                lin.weight.copy_(W); lin.bias.copy_(b)

    class Enc(torch.nn.Module):
        def forward(self, x, bound=1, **kw):
            return NO.freq_encode(x, 6)
    model.encoder = Enc()
    model.eval()
    g = torch.Generator().manual_seed(11)
    M = 256
    u = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1)
    x = (u * (0.25 + 0.4 * torch.rand(M, 1, generator=g))).requires_grad_(True)          # around the surface (r = 0.37 .. 0.48)
    d = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1)
    dists = torch.full((M,), 2 * 3 ** 0.5 / 256)
    sdfs, alpha, geo, normals, _ = model.forward_sigma(x, use_sdf_sigma_grad=True, dirs=d, dists=dists)
    rough = model.roughness
    n_enc, w_r_enc, n_dot, n_env_enc = model.get_color_mlp_extra_params(normals, d, rough, None)
    rgb = model.forward_color(geo, d, n_enc, w_r_enc, n_dot, True, n_env_enc=n_env_enc, r_images=None, roughness=rough)
    f = lambda t: t.detach().numpy().astype(np.float32)
    out = dict(x=f(x), d=f(d), dists=f(dists), sdf=f(sdfs), alpha=f(alpha).reshape(-1), geo=f(geo), normal=f(normals), roughness=f(rough), rgb=f(rgb),
               c_diffuse=f(model.c_diffuse), c_specular=f(model.c_specular), blend=f(model.blend_weight))
    # ---- the inference loop through NeRFRenderer.render
    class RM:
        @staticmethod
        def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
            n, fr = O.near_far_from_aabb(rays_o.detach().numpy(), rays_d.detach().numpy(), aabb.numpy(), min_near)
            return torch.from_numpy(n), torch.from_numpy(fr)

        @staticmethod
        def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, C, H, near, far, align=-1, perturb=False,
                       dt_gamma=0, max_steps=1024):
            xx, dd, dl, _ = O.march_rays(n_alive, n_step, rays_alive.numpy(), rays_t.numpy(), rays_o.detach().numpy(), rays_d.detach().numpy(), bound,
                                         bitfield.numpy(), C, H, near.numpy(), far.numpy(), align=align, dt_gamma=dt_gamma, max_steps=max_steps)
            return torch.from_numpy(xx), torch.from_numpy(dd), torch.from_numpy(dl)

        @staticmethod
        def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2,
                           input_alpha=False, accum_deltas=True):
            O.composite_rays(n_alive, n_step, rays_alive.numpy(), rays_t.numpy(), sigmas.detach().numpy(), rgbs.detach().numpy(),
                             deltas.detach().numpy(), weights_sum.numpy(), depth.numpy(), image.numpy(), T_thresh=T_thresh,
                             input_alpha=bool(input_alpha), accum_deltas=bool(accum_deltas))
            return tuple()
    CR.raymarching = RM
    bf = scene.make_sphere_bitfield()
    model.density_bitfield = torch.from_numpy(bf)
    Wd = 16
    ro, rd = scene.camera_rays(Wd, Wd)
    opt.indir_ref = False
    kw = {k: v for k, v in vars(opt).items()}
    res = model.render(ro[None], rd[None], staged=True, bg_color=1, perturb=False, get_normal_image=True, **kw)
    out.update(rays_o=ro.numpy(), rays_d=rd.numpy())
    for k in ("image", "depth", "weights_sum", "normal_image"):
        out[f"frame_{k}"] = f(res[k]).reshape(Wd * Wd, -1)
    np.savez_compressed(os.path.join(HERE, "neus_field.npz"), **out)
    print("neus_field.npz: alpha range", float(out["alpha"].min()), float(out["alpha"].max()), "hit pixels", int((out["frame_weights_sum"] > 0.5).sum()), "of", Wd * Wd)


if __name__ == "__main__":
    install_shims()
    if "neus_field" in sys.argv[1:]:
        torch.set_num_threads(8)
        gen_neus_field()
        sys.exit(0)
    if "infer" in sys.argv[1:]:
        torch.set_num_threads(8)
        gen_infer_branch()
        sys.exit(0)
    if "train" in sys.argv[1:]:
        torch.set_num_threads(8)
        gen_train_branch()
        sys.exit(0)
    if "neus" in sys.argv[1:]:
        gen_neus()
        sys.exit(0)
    if "keys" in sys.argv[1:]:
        gen_state_dict_keys()
        sys.exit(0)
    if "epilogue" in sys.argv[1:]:
        gen_train_epilogue()
        sys.exit(0)
    if "density" in sys.argv[1:]:
        torch.set_num_threads(8)
        gen_density_grid()
        sys.exit(0)
    torch.set_num_threads(1)
    gen_ide()
    gen_demo()
    gen_field_glue()
    gen_relight()
