"""GPU parity tests for the operator-level C ABI (called through the reference-shaped Python surface).

Each op is compared on identical seeded inputs with
  (a) the reference's own kernel, rebuilt unmodified for sm_100a (oracle/_ref/_<pkg>.so), when present, and
  (b) the CPU oracle (oracle/envidr_oracle.c).
Bar: bit-exact for integer / index / march outputs; floats within the tolerance written in each test.
"""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))


def _ref(name):
    try:
        return __import__(name)
    except Exception as e:  # pragma: no cover
        pytest.skip(f"oracle/_ref/{name}.so not available: {e}")


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def scene_data(dev):
    from envidr_b200 import scene
    bf = scene.make_bitfield()
    ro, rd = scene.camera_rays(96, 96)
    return dict(bitfield=bf, bitfield_t=torch.from_numpy(bf).to(dev), rays_o=ro, rays_d=rd, rays_o_t=ro.to(dev), rays_d_t=rd.to(dev))


# ---------------------------------------------------------------------------------------------
# support ops
# ---------------------------------------------------------------------------------------------

def test_near_far_and_sph(dev, scene_data):
    from envidr_b200 import raymarching as rm
    from oracle import oracle as O
    R = _ref("_raymarching")
    ro, rd = scene_data["rays_o_t"].clone(), scene_data["rays_d_t"].clone()
    N = ro.shape[0]
    g = torch.Generator().manual_seed(0)
    rd[:64] = torch.nn.functional.normalize(torch.randn(64, 3, generator=g), dim=-1).to(dev)     # random directions: most miss the box
    aabb = torch.tensor([-1, -1, -1, 1, 1, 1.0], device=dev)
    nears, fars = rm.near_far_from_aabb(ro, rd, aabb, 0.2)
    n_ref, f_ref = torch.empty(N, device=dev), torch.empty(N, device=dev)
    R.near_far_from_aabb(ro, rd, aabb, N, 0.2, n_ref, f_ref)
    assert torch.equal(nears, n_ref) and torch.equal(fars, f_ref)            # bit-exact vs reference kernel
    n_o, f_o = O.near_far_from_aabb(ro.cpu().numpy(), rd.cpu().numpy(), aabb.cpu().numpy(), 0.2)
    np.testing.assert_array_equal(nears.cpu().numpy(), n_o)                   # and vs the CPU oracle
    np.testing.assert_array_equal(fars.cpu().numpy(), f_o)
    assert (nears == torch.finfo(torch.float32).max).any(), "some rays must miss the box"
    # rays from inside a radius-4 sphere
    coords = rm.sph_from_ray(ro, rd, 4.0)
    c_ref = torch.empty(N, 2, device=dev)
    R.sph_from_ray(ro, rd, 4.0, N, c_ref)
    torch.testing.assert_close(coords, c_ref, atol=2e-6, rtol=0)
    np.testing.assert_allclose(coords.cpu().numpy(), O.sph_from_ray(ro.cpu().numpy(), rd.cpu().numpy(), 4.0), atol=2e-6)


def test_morton_packbits_scatter(dev):
    from envidr_b200 import raymarching as rm
    from oracle import oracle as O
    R = _ref("_raymarching")
    g = torch.Generator().manual_seed(0)
    coords = torch.randint(0, 128, (5000, 3), generator=g, dtype=torch.int32)
    coords[0] = torch.tensor([0, 0, 0]); coords[1] = torch.tensor([127, 127, 127]); coords[2] = torch.tensor([1023, 0, 1023])
    idx = rm.morton3D(coords.to(dev))
    idx_ref = torch.empty_like(idx)
    R.morton3D(coords.to(dev), coords.shape[0], idx_ref)
    assert torch.equal(idx, idx_ref)
    np.testing.assert_array_equal(idx.cpu().numpy(), O.morton3D(coords.numpy()))
    back = rm.morton3D_invert(idx)
    assert torch.equal(back.cpu(), coords)                                    # round trip
    # packbits: thresholds incl. exact ties, ragged tail handled by N
    grid = torch.rand(2, 32 ** 3, generator=g)
    grid[0, :8] = 0.5
    for thresh in (0.5, 0.01, -1.0, 2.0):
        bits = rm.packbits(grid.to(dev), thresh)
        b_ref = torch.empty_like(bits)
        R.packbits(grid.to(dev), bits.numel(), thresh, b_ref)
        assert torch.equal(bits, b_ref)
        np.testing.assert_array_equal(bits.cpu().numpy(), O.packbits(grid.numpy(), thresh))
    # scatter idx with empty rays
    counts = torch.randint(0, 40, (300,), generator=g, dtype=torch.int32)
    counts[::7] = 0
    offs = torch.cumsum(counts, 0, dtype=torch.int32) - counts
    rays = torch.stack([torch.randperm(300, generator=g).int(), offs, counts], -1).contiguous()
    M = int(counts.sum())
    out = rm.get_scatter_idx(rays.to(dev), torch.zeros(M, dtype=torch.int32, device=dev))
    np.testing.assert_array_equal(out.cpu().numpy(), O.get_scatter_idx(rays.numpy(), M))


# ---------------------------------------------------------------------------------------------
# march (bit-exact) and composite
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("dt_gamma,max_steps", [(0.0, 1024), (1.0 / 128, 512), (0.0, 1000)])
def test_march_rays_inference_bit_exact(dev, scene_data, dt_gamma, max_steps):
    """One march_rays call from the ray origins and a second one continuing from advanced rays_t, n_step 1 and 8."""
    from envidr_b200 import raymarching as rm
    from oracle import oracle as O
    R = _ref("_raymarching")
    ro, rd, bf = scene_data["rays_o_t"], scene_data["rays_d_t"], scene_data["bitfield_t"]
    N = ro.shape[0]
    aabb = torch.tensor([-1, -1, -1, 1, 1, 1.0], device=dev)
    nears, fars = rm.near_far_from_aabb(ro, rd, aabb, 0.2)
    alive = torch.arange(N, dtype=torch.int32, device=dev)
    rays_t = nears.clone()
    for n_step in (1, 8, 3):
        x, d, dl = rm.march_rays(N, n_step, alive, rays_t, ro, rd, 1.0, bf, 1, 128, nears, fars, 128, False, dt_gamma, max_steps)
        M = x.shape[0]
        xr, dr, dlr = torch.zeros(M, 3, device=dev), torch.zeros(M, 3, device=dev), torch.zeros(M, 2, device=dev)
        R.march_rays(N, n_step, alive, rays_t, ro, rd, 1.0, dt_gamma, max_steps, 1, 128, bf, nears, fars, xr, dr, dlr,
                     torch.zeros(N, device=dev))
        assert torch.equal(x, xr) and torch.equal(d, dr) and torch.equal(dl, dlr), f"march_rays differs from reference (n_step={n_step})"
        xo, do, dlo, cnt = O.march_rays(N, n_step, alive.cpu().numpy(), rays_t.cpu().numpy(), ro.cpu().numpy(), rd.cpu().numpy(), 1.0,
                                        scene_data["bitfield"], 1, 128, nears.cpu().numpy(), fars.cpu().numpy(), align=128,
                                        dt_gamma=dt_gamma, max_steps=max_steps)
        np.testing.assert_array_equal(x.cpu().numpy(), xo)
        np.testing.assert_array_equal(dl.cpu().numpy(), dlo)
        assert cnt.sum() > 0
        # advance every ray by its accumulated deltas, as composite_rays would
        adv = dl[: N * n_step, 1].view(N, n_step).sum(-1)
        rays_t = rays_t + adv


def test_march_rays_train_counts_and_samples(dev, scene_data):
    from envidr_b200 import raymarching as rm
    from oracle import oracle as O
    R = _ref("_raymarching")
    ro, rd, bf = scene_data["rays_o_t"], scene_data["rays_d_t"], scene_data["bitfield_t"]
    N = ro.shape[0]
    aabb = torch.tensor([-1, -1, -1, 1, 1, 1.0], device=dev)
    nears, fars = rm.near_far_from_aabb(ro, rd, aabb, 0.2)
    g = torch.Generator().manual_seed(3)
    noises = torch.rand(N, generator=g).to(dev)
    for early, dtg in ((1024, 0.0), (24, 0.0), (1024, 1.0 / 128)):       # dt_gamma = 0: specialised probe; > 0: general probe
        M = N * 64
        from envidr_b200.backend import _raymarching as B
        xyzs, dirs, deltas = torch.zeros(M, 3, device=dev), torch.zeros(M, 3, device=dev), torch.zeros(M, 2, device=dev)
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        counter = torch.zeros(2, dtype=torch.int32, device=dev)
        B.march_rays_train(ro, rd, bf, 1.0, dtg, 1024, early, N, 1, 128, M, nears, fars, xyzs, dirs, deltas, rays, counter, noises)
        xr, dr, dlr = torch.zeros(M, 3, device=dev), torch.zeros(M, 3, device=dev), torch.zeros(M, 2, device=dev)
        rays_r = torch.empty(N, 3, dtype=torch.int32, device=dev)
        counter_r = torch.zeros(2, dtype=torch.int32, device=dev)
        R.march_rays_train(ro, rd, bf, 1.0, dtg, 1024, early, N, 1, 128, M, nears, fars, xr, dr, dlr, rays_r, counter_r, noises)
        assert torch.equal(counter, counter_r)                       # total samples / rays
        # reference slot order is scheduling dependent: compare ray id -> (count, samples)
        rr = rays_r.cpu().numpy(); mine = rays.cpu().numpy()
        cnt_ref = np.zeros(N, np.int64); off_ref = np.zeros(N, np.int64)
        cnt_ref[rr[:, 0]] = rr[:, 2]; off_ref[rr[:, 0]] = rr[:, 1]
        np.testing.assert_array_equal(mine[:, 0], np.arange(N))
        np.testing.assert_array_equal(mine[:, 2], cnt_ref)
        assert (mine[:, 2] <= early).all()
        X, Xr, DL, DLr = xyzs.cpu().numpy(), xr.cpu().numpy(), deltas.cpu().numpy(), dlr.cpu().numpy()
        for n in np.nonzero(cnt_ref)[0][:: max(1, N // 400)]:
            a, b, c = mine[n, 1], off_ref[n], cnt_ref[n]
            np.testing.assert_array_equal(X[a:a + c], Xr[b:b + c])
            np.testing.assert_array_equal(DL[a:a + c], DLr[b:b + c])
        # and the deterministic CPU oracle reproduces our layout exactly
        xo, do, dlo, rays_o_, ctr_o = O.march_rays_train(ro.cpu().numpy(), rd.cpu().numpy(), 1.0, scene_data["bitfield"], 1, 128,
                                                         nears.cpu().numpy(), fars.cpu().numpy(), M, noises=noises.cpu().numpy(),
                                                         dt_gamma=dtg, early_stop_steps=early)
        np.testing.assert_array_equal(mine, rays_o_)
        np.testing.assert_array_equal(X, xo)
        np.testing.assert_array_equal(DL, dlo)
    # dropped rays when M is too small: every kept ray still fits, totals still counted
    M = 4096
    xyzs, dirs, deltas, rays = rm.march_rays_train(ro, rd, 1.0, bf, 1, 128, nears, fars, torch.zeros(2, dtype=torch.int32, device=dev), M,
                                                   False, 128, False, 0.0, 1024)
    assert xyzs.shape[0] == M + 128 - M % 128 or xyzs.shape[0] == M


def _random_train_batch(seed, N=700, M_pad=64, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    counts = torch.poisson(torch.full((N,), 32.0), generator=g).clamp(0, 1024).int()
    counts[::11] = 0
    counts[5] = 300
    offs = (torch.cumsum(counts, 0) - counts).int()
    perm = torch.randperm(N, generator=g)
    rays = torch.stack([perm.int(), offs, counts], -1).contiguous()
    M = int(counts.sum()) + M_pad
    sigmas = torch.nn.functional.softplus(torch.randn(M, generator=g)) * 50
    rgbs = torch.rand(M, 3, generator=g)
    deltas = torch.stack([torch.full((M,), 2 * 3 ** 0.5 / 1024), torch.rand(M, generator=g) * 0.01], -1).contiguous()
    return rays, sigmas, rgbs, deltas


@pytest.mark.parametrize("input_alpha,accum", [(False, True), (True, True), (False, False)])
def test_composite_rays_train_fwd_bwd(dev, input_alpha, accum):
    from envidr_b200 import raymarching as rm
    from oracle import oracle as O
    R = _ref("_raymarching")
    rays, sigmas, rgbs, deltas = _random_train_batch(0)
    if input_alpha:
        sigmas = torch.rand_like(sigmas) * 0.6
    N, M = rays.shape[0], sigmas.shape[0]
    rays[3, 1] = M - 2; rays[3, 2] = 50                                    # a ray that overflows M -> dropped, outputs zero
    s, c, dl, ry = sigmas.to(dev).requires_grad_(True), rgbs.to(dev).requires_grad_(True), deltas.to(dev), rays.to(dev)
    ws, depth, image, weights = rm.composite_rays_train(s, c, dl, ry, 1e-4, True, input_alpha, accum)
    ws_r, d_r, im_r, w_r = torch.empty(N, device=dev), torch.empty(N, device=dev), torch.empty(N, 3, device=dev), torch.zeros(M, device=dev)
    R.composite_rays_train_forward(s.detach(), c.detach(), dl, ry, M, N, 1e-4, int(accum), int(input_alpha), ws_r, d_r, im_r, w_r)
    tol = dict(atol=2e-6, rtol=2e-5)          # warp scan vs serial product: rounding only
    torch.testing.assert_close(ws, ws_r, **tol); torch.testing.assert_close(image, im_r, **tol)
    torch.testing.assert_close(depth, d_r, **tol); torch.testing.assert_close(weights, w_r, **tol)
    wo, do, io, wwo = O.composite_rays_train_forward(sigmas.numpy(), rgbs.numpy(), deltas.numpy(), rays.numpy(), N, 1e-4, True, input_alpha, accum)
    np.testing.assert_allclose(image.detach().cpu().numpy(), io, atol=5e-6, rtol=1e-4)          # __expf vs expf
    np.testing.assert_allclose(ws.detach().cpu().numpy(), wo, atol=5e-6, rtol=1e-4)
    assert float(ws[rays[3, 0]]) == 0.0
    # no-weights variant gives the same image
    ws2, _, image2, w2 = rm.composite_rays_train(s.detach(), c.detach(), dl, ry, 1e-4, False, input_alpha, accum)
    assert torch.equal(image2, image.detach()) and w2.numel() == 0
    # backward
    g = torch.Generator().manual_seed(1)
    g_ws, g_im = torch.randn(N, generator=g).to(dev), torch.randn(N, 3, generator=g).to(dev)
    (ws * g_ws).sum().add((image * g_im).sum()).backward()
    gs_r, gc_r = torch.zeros(M, device=dev), torch.zeros(M, 3, device=dev)
    R.composite_rays_train_backward(g_ws, g_im, torch.zeros(N, device=dev), s.detach(), c.detach(), dl, ry, ws_r, im_r, d_r, M, N, 1e-4,
                                    gs_r, gc_r, int(accum), int(input_alpha))
    torch.testing.assert_close(c.grad, gc_r, atol=2e-6, rtol=2e-5)
    torch.testing.assert_close(s.grad, gs_r, atol=5e-5, rtol=1e-3)         # differences of O(1) prefix sums, scaled by 1/(1-alpha)


@pytest.mark.parametrize("n_step", [1, 4, 8])
def test_composite_rays_inference(dev, n_step):
    from envidr_b200 import raymarching as rm
    from oracle import oracle as O
    R = _ref("_raymarching")
    g = torch.Generator().manual_seed(n_step)
    N, n_alive = 900, 500
    alive = torch.randperm(N, generator=g)[:n_alive].int()
    M = n_alive * n_step
    sig = torch.nn.functional.softplus(torch.randn(M, generator=g)) * 80
    rgb = torch.rand(M, 3, generator=g)
    dl = torch.stack([torch.full((M,), 2 * 3 ** 0.5 / 1024), torch.rand(M, generator=g) * 0.01], -1)
    kill = torch.rand(n_alive, generator=g) < 0.3                             # rays that ran out of samples: zero-delta tail
    for n in torch.nonzero(kill).flatten().tolist():
        k = int(torch.randint(0, n_step, (1,), generator=g))
        dl[n * n_step + k:(n + 1) * n_step] = 0
    state = lambda: [alive.clone().to(dev), torch.rand(N, generator=torch.Generator().manual_seed(5)).to(dev),
                     (torch.rand(N, generator=torch.Generator().manual_seed(6)) * 0.999).to(dev), torch.rand(N, generator=torch.Generator().manual_seed(7)).to(dev),
                     torch.rand(N, 3, generator=torch.Generator().manual_seed(8)).to(dev)]
    a1, t1, w1, d1, i1 = state()
    a2, t2, w2, d2, i2 = state()
    w1[alive[:20].long()] = 0.99995; w2[alive[:20].long()] = 0.99995          # T < T_thresh on entry
    for accum in (True, False):
        rm.composite_rays(n_alive, n_step, a1, t1, sig.to(dev), rgb.to(dev), dl.contiguous().to(dev), w1, d1, i1, 1e-4, False, accum)
        R.composite_rays(n_alive, n_step, 1e-4, int(accum), 0, a2, t2, sig.to(dev), rgb.to(dev), dl.contiguous().to(dev), w2, d2, i2)
        assert torch.equal(a1, a2) and torch.equal(t1, t2)
        assert torch.equal(w1, w2) and torch.equal(d1, d2) and torch.equal(i1, i2)      # same serial arithmetic -> bit-exact
        a1 = a1.clone(); a2 = a2.clone()
        a1[a1 < 0] = alive.to(dev)[a1 < 0]; a2[a2 < 0] = alive.to(dev)[a2 < 0]
    # CPU oracle (expf instead of __expf): tolerance
    a3, t3, w3, d3, i3 = [x.cpu().numpy().copy() for x in state()]
    O.composite_rays(n_alive, n_step, a3, t3, sig.numpy(), rgb.numpy(), dl.contiguous().numpy(), w3, d3, i3, 1e-4)
    a4, t4, w4, d4, i4 = state()
    rm.composite_rays(n_alive, n_step, a4, t4, sig.to(dev), rgb.to(dev), dl.contiguous().to(dev), w4, d4, i4, 1e-4)
    np.testing.assert_array_equal(a4.cpu().numpy(), a3)
    np.testing.assert_allclose(i4.cpu().numpy(), i3, atol=3e-6, rtol=1e-5)


# ---------------------------------------------------------------------------------------------
# encoders
# ---------------------------------------------------------------------------------------------

def _pin_oracle_level_scales(dev, S, H, L):
    """exp2f on the device is ex2.approx (<= 2 ulp from libm): give the CPU oracle the device's per-level scale so that
    both sides interpolate in exactly the same cells with the same fractions (see orc_set_level_scales)."""
    from envidr_b200._lib import check, lib, ptr, stream
    from oracle import oracle as O
    sc = torch.empty(L, device=dev)
    check(lib().envidr_debug_level_scales(float(S), H, L, ptr(sc), stream()))
    sc = sc.cpu().numpy()
    exact = np.exp2(np.arange(L, dtype=np.float32) * np.float32(S)).astype(np.float32) * np.float32(H) - np.float32(1)
    assert np.abs(sc - exact).max() <= 4 * np.spacing(np.float32(exact.max())), (sc, exact)     # a couple of ulp at most
    assert (np.ceil(sc) == np.ceil(exact)).all()                                                # same resolutions
    O.set_level_scales(sc)


def _enc_inputs(B, D, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, D, generator=g)
    x[0] = 0.0; x[1] = 1.0; x[2] = 0.5; x[3, 0] = -0.1; x[4, -1] = 1.2          # corners, centre, out of range
    x[5] = 1.0 / 15.0                                                           # exactly on a level-0 vertex
    return x


@pytest.mark.parametrize("D,C", [(3, 2), (3, 4), (2, 2), (3, 1), (3, 8)])
def test_hash_encode_forward_backward(dev, D, C):
    from envidr_b200.backend import _hashencoder as B
    from oracle import oracle as O
    R = _ref("_hashencoder")
    L, H = 16, 16
    offsets, pls = O.hash_offsets(D, L, H, 19 if D == 3 else 16, 2048)
    T = int(offsets[-1])
    g = torch.Generator().manual_seed(D * 10 + C)
    emb = (torch.rand(T, C, generator=g) * 2 - 1)
    Bn = 4099
    x = _enc_inputs(Bn, D, 1)
    S = float(np.log2(pls))
    off_t = torch.from_numpy(offsets).to(dev)
    xd, ed = x.to(dev), emb.to(dev)
    out, jac = torch.empty(L, Bn, C, device=dev), torch.empty(Bn, L * D * C, device=dev)
    B.hash_encode_forward(xd, ed, off_t, out, Bn, D, C, L, S, H, True, jac)
    out_r, jac_r = torch.empty_like(out), torch.empty_like(jac)
    R.hash_encode_forward(xd, ed, off_t, out_r, Bn, D, C, L, S, H, True, jac_r)
    torch.testing.assert_close(out, out_r, atol=1e-6, rtol=1e-5)                 # same cells (bit-exact indices), rounding-level blend
    torch.testing.assert_close(jac, jac_r, atol=2e-3, rtol=1e-4)                 # |dy_dx| up to scale*|table| ~ 2e3: relative check
    _pin_oracle_level_scales(dev, S, H, L)
    out_o, jac_o = O.hash_encode_forward(x.numpy(), emb.numpy(), offsets, pls, H, True)
    np.testing.assert_allclose(out.cpu().numpy(), out_o, atol=1e-6, rtol=1e-5)
    np.testing.assert_allclose(jac.cpu().numpy(), jac_o, atol=2e-3, rtol=1e-4)
    assert float(out[:, 3].abs().max()) == 0 and float(jac[3].abs().max()) == 0  # out-of-range input -> zeros
    # no-grad variant
    out2 = torch.empty_like(out)
    B.hash_encode_forward(xd, ed, off_t, out2, Bn, D, C, L, S, H, False, torch.empty(1, device=dev))
    assert torch.equal(out2, out)
    # backward (atomics: summation order differs -> tolerance)
    grad = torch.randn(L, Bn, C, generator=g).to(dev)
    ge, gi = torch.zeros_like(ed), torch.zeros_like(xd)
    B.hash_encode_backward(grad, xd, ed, off_t, ge, Bn, D, C, L, S, H, True, jac, gi)
    ge_r, gi_r = torch.zeros_like(ed), torch.zeros_like(xd)
    R.hash_encode_backward(grad, xd, ed, off_t, ge_r, Bn, D, C, L, S, H, True, jac_r, gi_r)
    torch.testing.assert_close(ge, ge_r, atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(gi, gi_r, atol=5e-2, rtol=1e-4)                   # sums of 32*C terms of size up to 1e3
    ge_o, gi_o = O.hash_encode_backward(grad.cpu().numpy(), x.numpy(), emb.numpy(), offsets, pls, H, jac_o)
    np.testing.assert_allclose(ge.cpu().numpy(), ge_o, atol=2e-5, rtol=1e-4)
    # linearity of the gather in the table (size-independent property)
    out3 = torch.empty_like(out)
    B.hash_encode_forward(xd, 2 * ed, off_t, out3, Bn, D, C, L, S, H, False, torch.empty(1, device=dev))
    torch.testing.assert_close(out3, 2 * out, atol=1e-6, rtol=1e-6)
    if C == 1:
        O.set_level_scales(None)
        with pytest.raises(RuntimeError):
            B.hash_encode_second_backward(grad, xd, ed, off_t, Bn, D, C, L, S, H, True, jac, gi, torch.zeros_like(grad), torch.zeros_like(ed))
        return
    # second-order backward
    ggx = torch.randn(Bn, D, generator=g).to(dev)
    gg, g2 = torch.zeros_like(grad), torch.zeros_like(ed)
    B.hash_encode_second_backward(grad, xd, ed, off_t, Bn, D, C, L, S, H, True, jac, ggx, gg, g2)
    gg_r, g2_r = torch.zeros_like(grad), torch.zeros_like(ed)
    R.hash_encode_second_backward(grad, xd, ed, off_t, Bn, D, C, L, S, H, True, jac_r, ggx, gg_r, g2_r)
    torch.testing.assert_close(gg, gg_r, atol=5e-3, rtol=1e-4)
    torch.testing.assert_close(g2, g2_r, atol=5e-2, rtol=1e-3)
    gg_o, g2_o = O.hash_encode_second_backward(grad.cpu().numpy(), x.numpy(), emb.numpy(), offsets, pls, H, jac_o, ggx.cpu().numpy())
    np.testing.assert_allclose(g2.cpu().numpy(), g2_o, atol=5e-2, rtol=1e-3)
    O.set_level_scales(None)


def test_hash_encode_unsupported_dims(dev):
    from envidr_b200.backend import _hashencoder as B
    x = torch.rand(8, 4, device=dev)
    emb = torch.rand(64, 2, device=dev)
    off = torch.tensor([0, 64], dtype=torch.int32, device=dev)
    with pytest.raises(RuntimeError):
        B.hash_encode_forward(x, emb, off, torch.empty(1, 8, 2, device=dev), 8, 4, 2, 1, 1.0, 16, False, torch.empty(1, device=dev))
    with pytest.raises(RuntimeError):      # C = 3
        B.hash_encode_forward(x[:, :3].contiguous(), torch.rand(64, 3, device=dev), off, torch.empty(1, 8, 3, device=dev), 8, 3, 3, 1, 1.0,
                              16, False, torch.empty(1, device=dev))
    with pytest.raises(RuntimeError):      # CPU tensor
        B.hash_encode_forward(x.cpu(), emb, off, torch.empty(1, 8, 2, device=dev), 8, 3, 2, 1, 1.0, 16, False, torch.empty(1, device=dev))


@pytest.mark.parametrize("gridtype,align,D,C", [(0, False, 3, 2), (1, False, 3, 2), (0, True, 3, 4), (1, True, 2, 8), (0, False, 2, 1)])
def test_grid_encode(dev, gridtype, align, D, C):
    from envidr_b200.backend import _gridencoder as B
    from oracle import oracle as O
    R = _ref("_gridencoder")
    L, H = 16, 16
    offsets, pls = O.grid_offsets(D, L, H, 19 if D == 3 else 16, 2048, align_corners=align)
    T = int(offsets[-1])
    g = torch.Generator().manual_seed(gridtype * 7 + D + C)
    emb = (torch.rand(T, C, generator=g) * 2 - 1)
    Bn = 3001
    x = _enc_inputs(Bn, D, 2)
    S = float(np.log2(pls))
    off_t = torch.from_numpy(offsets).to(dev)
    xd, ed = x.to(dev), emb.to(dev)
    out, jac = torch.empty(L, Bn, C, device=dev), torch.empty(Bn, L * D * C, device=dev)
    B.grid_encode_forward(xd, ed, off_t, out, Bn, D, C, L, S, H, jac, gridtype, align)
    out_r, jac_r = torch.empty_like(out), torch.empty_like(jac)
    R.grid_encode_forward(xd, ed, off_t, out_r, Bn, D, C, L, S, H, jac_r, gridtype, align)
    torch.testing.assert_close(out, out_r, atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(jac, jac_r, atol=2e-3, rtol=1e-4)
    _pin_oracle_level_scales(dev, S, H, L)
    out_o, jac_o = O.grid_encode_forward(x.numpy(), emb.numpy(), offsets, pls, H, True, gridtype, align)
    O.set_level_scales(None)
    np.testing.assert_allclose(out.cpu().numpy(), out_o, atol=1e-6, rtol=1e-5)
    out2 = torch.empty_like(out)
    B.grid_encode_forward(xd, ed, off_t, out2, Bn, D, C, L, S, H, None, gridtype, align)
    assert torch.equal(out2, out)
    grad = torch.randn(L, Bn, C, generator=g).to(dev)
    ge, gi = torch.zeros_like(ed), torch.zeros_like(xd)
    B.grid_encode_backward(grad, xd, ed, off_t, ge, Bn, D, C, L, S, H, jac, gi, gridtype, align)
    ge_r, gi_r = torch.zeros_like(ed), torch.zeros_like(xd)
    R.grid_encode_backward(grad, xd, ed, off_t, ge_r, Bn, D, C, L, S, H, jac_r, gi_r, gridtype, align)
    torch.testing.assert_close(ge, ge_r, atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(gi, gi_r, atol=5e-2, rtol=1e-4)


def test_freq_encode(dev):
    from envidr_b200.backend import _freqencoder as B
    from oracle import oracle as O
    R = _ref("_freqencoder")
    g = torch.Generator().manual_seed(0)
    for D, deg in ((3, 6), (3, 4), (2, 10), (3, 0)):
        Bn, C = 1025, D + 2 * D * deg
        x = (torch.rand(Bn, D, generator=g) * 2 - 1).to(dev)
        out, out_r = torch.empty(Bn, C, device=dev), torch.empty(Bn, C, device=dev)
        B.freq_encode_forward(x, Bn, D, deg, C, out)
        R.freq_encode_forward(x, Bn, D, deg, C, out_r)
        # same intrinsic (__sinf(scalbnf(x,f)+phase)); the reference TU is built with -use_fast_math (FTZ only differs on denormals)
        torch.testing.assert_close(out, out_r, atol=1e-6, rtol=0)
        if deg <= 6:   # __sinf (fast intrinsic) vs libm sinf: abs error grows with |2^f x|
            np.testing.assert_allclose(out.cpu().numpy(), O.freq_encode_forward(x.cpu().numpy(), deg), atol=2e-4)
        grad = torch.randn(Bn, C, generator=g).to(dev)
        gi, gi_r = torch.zeros(Bn, D, device=dev), torch.zeros(Bn, D, device=dev)
        B.freq_encode_backward(grad, out, Bn, D, deg, C, gi)
        R.freq_encode_backward(grad, out_r, Bn, D, deg, C, gi_r)
        torch.testing.assert_close(gi, gi_r, atol=1e-4, rtol=1e-5)


@pytest.mark.parametrize("degree", [1, 2, 4, 5, 8])
def test_sh_encode(dev, degree):
    from envidr_b200.backend import _shencoder as B
    from oracle import oracle as O
    R = _ref("_shencoder")
    g = torch.Generator().manual_seed(degree)
    Bn = 777
    x = torch.nn.functional.normalize(torch.randn(Bn, 3, generator=g), dim=-1)
    x[0] = torch.tensor([0.0, 0.0, 1.0]); x[1] = torch.tensor([0.3, -0.2, 0.5])      # pole; non-unit input (polynomials evaluated as-is)
    xd = x.to(dev)
    C2 = degree * degree
    out, jac = torch.empty(Bn, C2, device=dev), torch.empty(Bn, 3 * C2, device=dev)
    B.sh_encode_forward(xd, out, Bn, 3, degree, jac)
    out_r, jac_r = torch.empty_like(out), torch.empty_like(jac)
    R.sh_encode_forward(xd, out_r, Bn, 3, degree, jac_r)
    # recurrence vs hard-coded polynomials: same polynomials, fp32 rounding differs; derivatives reach ~50 at degree 8
    torch.testing.assert_close(out, out_r, atol=2e-5, rtol=2e-5)
    torch.testing.assert_close(jac, jac_r, atol=5e-4, rtol=2e-5)
    out_o, jac_o = O.sh_encode_forward(x.numpy(), degree, True)
    np.testing.assert_allclose(out.cpu().numpy(), out_o, atol=2e-5, rtol=2e-5)
    np.testing.assert_allclose(jac.cpu().numpy(), jac_o, atol=5e-4, rtol=2e-5)
    grad = torch.randn(Bn, C2, generator=g).to(dev)
    gi, gi_r = torch.ones(Bn, 3, device=dev), torch.ones(Bn, 3, device=dev)             # accumulates (+=)
    B.sh_encode_backward(grad, xd, Bn, 3, degree, jac, gi)
    R.sh_encode_backward(grad, xd, Bn, 3, degree, jac_r, gi_r)
    torch.testing.assert_close(gi, gi_r, atol=2e-3, rtol=1e-4)


@pytest.mark.parametrize("deg", [4, 5])
def test_ide_kernel_vs_reference_golden(dev, golden_dir, deg):
    from envidr_b200.ide_encoder import IntegratedDirEncoder
    z = np.load(os.path.join(golden_dir, "ide.npz"))
    enc = IntegratedDirEncoder(deg_view=deg).to(dev)
    d = torch.from_numpy(z[f"dirs{deg}"]).to(dev)
    from oracle import oracle as O
    l = np.concatenate([z[f"ml{deg}"][1]] * 2)
    atol = np.where(l >= 16, 1e-3, 1e-5)[None]       # l=16: fp32 cancellation noise of the reference formula itself (~5e-4)
    atol_t = np.where(l >= 16, 1e-3, 2e-5)[None]
    for rough, key in ((torch.from_numpy(z[f"rough{deg}"]).to(dev), "var"), (0.64, "const")):
        got = enc(d, rough).cpu().numpy()
        ref = z[f"ide{deg}_{key}"]
        assert got.shape == ref.shape
        assert (np.abs(got - ref) <= atol + 2e-5 * np.abs(ref)).all(), np.abs(got - ref).max()
        # against the exact encoding (float64 arithmetic AND float64 coefficients) the kernel is accurate in every band
        r64 = rough.double().cpu() if torch.is_tensor(rough) else rough
        exact = O.ide_encode(d.double().cpu(), r64, deg, exact_tables=True).numpy()
        assert (np.abs(got - exact) <= 2e-6 + 2e-5 * np.abs(exact)).all(), np.abs(got - exact).max()   # values reach 90 at z = +-1
    # autograd path (torch formulation) agrees with the kernel
    d2 = d.clone().requires_grad_(True)
    got_t = enc(d2, 0.1)
    got_k = enc(d, 0.1)
    assert (np.abs(got_t.detach().cpu().numpy() - got_k.cpu().numpy()) <= atol_t + 2e-5 * np.abs(got_k.cpu().numpy())).all()


def test_gather_rows_equals_index_select(dev):
    """envidr_gather_rows (frame assembly after the all-gather of a sharded render, envidr_b200/dist.py) against torch.index_select: bit-exact,
    rows of 8 floats (the packed per-ray outputs) and of 4, a count that is not a multiple of the block size; and through dist._deinterleave."""
    from envidr_b200 import dist as D
    from envidr_b200._lib import check, lib, ptr, stream
    g = torch.Generator().manual_seed(5)
    for n, rf in ((100_003, 8), (257, 4), (1, 8)):
        src = torch.rand(n + 17, rf, generator=g).to(dev)
        idx = torch.randint(0, n + 17, (n,), generator=g).to(dev)
        dst = torch.empty(n, rf, device=dev)
        check(lib().envidr_gather_rows(ptr(src), ptr(idx.to(torch.int32)), n, rf, ptr(dst), stream()), "gather_rows")
        assert torch.equal(dst, src.index_select(0, idx))
        assert torch.equal(D._deinterleave(src, idx), src.index_select(0, idx))
