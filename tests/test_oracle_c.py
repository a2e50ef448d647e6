"""CPU tests of the C restatement (oracle/envidr_oracle.c): internal consistency and domain properties.
(Its parity with the reference kernels themselves is checked on the GPU box, tests/test_gpu_ops.py.)"""
import numpy as np
import pytest
import torch

from envidr_b200 import scene
from oracle import oracle as O


def test_morton_roundtrip_and_packbits():
    rng = np.random.default_rng(0)
    c = rng.integers(0, 128, size=(4096, 3)).astype(np.int32)
    idx = O.morton3D(c)
    np.testing.assert_array_equal(O.morton3D_invert(idx), c)
    assert idx.max() < 128 ** 3 and len(np.unique(idx)) == len(np.unique(c, axis=0))
    assert O.morton3D(np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [127, 127, 127]], np.int32)).tolist() == [1, 2, 4, 128 ** 3 - 1]
    g = rng.random(64).astype(np.float32)
    bits = O.packbits(g, 0.5)
    np.testing.assert_array_equal(np.unpackbits(bits, bitorder="little"), (g > 0.5).astype(np.uint8))


def test_scene_bitfield_matches_morton_layout():
    bf = scene.make_bitfield()
    assert bf.shape == (128 ** 3 // 8,) and bf.dtype == np.uint8
    centre = O.morton3D(np.array([[64, 64, 64]], np.int32))[0]
    corner = O.morton3D(np.array([[0, 0, 0]], np.int32))[0]
    assert bf[centre // 8] & (1 << (centre % 8)) and not (bf[corner // 8] & (1 << (corner % 8)))


def test_march_inference_is_consistent_with_train_march():
    """Marching a ray in one go (train kernel) and in n_step chunks without re-synchronising t (inference kernel fed
    with the exact t) visit the same samples; counts match the occupancy structure."""
    bf = scene.make_bitfield()
    ro, rd = scene.camera_rays(40, 40)
    ro, rd = ro.numpy(), rd.numpy()
    N = ro.shape[0]
    nears, fars = O.near_far_from_aabb(ro, rd, np.array([-1, -1, -1, 1, 1, 1], np.float32), 0.2)
    xyzs, dirs, deltas, rays, ctr = O.march_rays_train(ro, rd, 1.0, bf, 1, 128, nears, fars, N * 256)
    assert ctr[1] == N and ctr[0] == rays[:, 2].sum() and ctr[0] > 1000
    np.testing.assert_array_equal(rays[:, 0], np.arange(N))
    np.testing.assert_array_equal(rays[:, 1], np.cumsum(rays[:, 2]) - rays[:, 2])
    # all samples lie inside occupied cells
    p = xyzs[: ctr[0]]
    cell = np.clip((0.5 * (p + 1) * 128).astype(np.int64), 0, 127).astype(np.int32)
    m = O.morton3D(cell)
    assert ((bf[m // 8] >> (m % 8)) & 1).all()
    # first n_step samples of the inference march equal the head of the train march
    alive = np.arange(N, dtype=np.int32)
    x8, d8, dl8, cnt = O.march_rays(N, 8, alive, nears.copy(), ro, rd, 1.0, bf, 1, 128, nears, fars)
    for n in np.nonzero(rays[:, 2])[0][:200]:
        k = min(8, rays[n, 2])
        assert cnt[n] == k
        np.testing.assert_array_equal(x8[n * 8: n * 8 + k], xyzs[rays[n, 1]: rays[n, 1] + k])
        np.testing.assert_array_equal(dl8[n * 8: n * 8 + k], deltas[rays[n, 1]: rays[n, 1] + k])
        assert (dl8[n * 8 + k: (n + 1) * 8] == 0).all()
    # early_stop_steps caps the count (raymarching.cu:388)
    _, _, _, rays24, _ = O.march_rays_train(ro, rd, 1.0, bf, 1, 128, nears, fars, N * 256, early_stop_steps=24)
    np.testing.assert_array_equal(rays24[:, 2], np.minimum(rays[:, 2], 24))


def test_composite_train_matches_closed_form_and_backward_matches_autograd():
    rng = np.random.default_rng(1)
    counts = np.array([0, 5, 40, 1, 17], np.int32)
    offs = (np.cumsum(counts) - counts).astype(np.int32)
    rays = np.stack([np.array([3, 0, 4, 1, 2], np.int32), offs, counts], -1)
    M = int(counts.sum())
    sig = (rng.random(M) * 30).astype(np.float32); rgb = rng.random((M, 3)).astype(np.float32)
    dl = np.stack([np.full(M, 0.0034, np.float32), (rng.random(M) * 0.01).astype(np.float32)], -1)
    ws, depth, img, w = O.composite_rays_train_forward(sig, rgb, dl, rays, 5, T_thresh=0.0)
    s, c = torch.tensor(sig, dtype=torch.float64, requires_grad=True), torch.tensor(rgb, dtype=torch.float64, requires_grad=True)
    tot = 0
    gi = rng.standard_normal((5, 3)); gw = rng.standard_normal(5)
    for r in range(5):
        idx, o, k = rays[r]
        if k == 0:
            assert ws[idx] == 0 and (img[idx] == 0).all()
            continue
        a = 1 - torch.exp(-s[o:o + k] * torch.tensor(dl[o:o + k, 0], dtype=torch.float64))
        T = torch.cumprod(torch.cat([torch.ones(1, dtype=torch.float64), 1 - a[:-1]]), 0)
        wt = a * T
        np.testing.assert_allclose(w[o:o + k], wt.detach().numpy(), rtol=2e-5, atol=1e-7)
        np.testing.assert_allclose(img[idx], (wt[:, None] * c[o:o + k]).sum(0).detach().numpy(), rtol=2e-5, atol=1e-6)
        t = np.cumsum(dl[o:o + k, 1].astype(np.float64))
        np.testing.assert_allclose(depth[idx], float((wt.detach().numpy() * t).sum()), rtol=2e-5, atol=1e-7)
        tot = tot + (torch.tensor(gi[idx]) * (wt[:, None] * c[o:o + k]).sum(0)).sum() + gw[idx] * wt.sum()
    tot.backward()
    gs, gc = O.composite_rays_train_backward(gw.astype(np.float32), gi.astype(np.float32), np.zeros(5, np.float32), sig, rgb, dl, rays, ws, img,
                                             depth, T_thresh=0.0)
    np.testing.assert_allclose(gc, c.grad.numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(gs, s.grad.numpy(), rtol=2e-3, atol=2e-6)


def test_composite_inference_equals_train_composite_when_chunked():
    """Chunked in-place compositing (inference kernel) reproduces the one-shot compositor, except that the inference
    kernel tests T *before* the sample (raymarching.cu:1009-1023) and so keeps one more sample after the threshold."""
    rng = np.random.default_rng(2)
    k, n_step = 37, 8
    sig = (rng.random(k) * 200).astype(np.float32); rgb = rng.random((k, 3)).astype(np.float32)
    dl = np.stack([np.full(k, 0.0034, np.float32), np.full(k, 0.0034, np.float32)], -1)
    rays = np.array([[0, 0, k]], np.int32)
    for thr in (0.0, 1e-4):
        ws1, d1, im1, w1 = O.composite_rays_train_forward(sig, rgb, dl, rays, 1, T_thresh=thr)
        alive, t, ws, d, im = np.array([0], np.int32), np.zeros(1, np.float32), np.zeros(1, np.float32), np.zeros(1, np.float32), np.zeros((1, 3), np.float32)
        used = 0
        for c0 in range(0, 40, n_step):
            s_, c_, d_ = np.zeros(n_step, np.float32), np.zeros((n_step, 3), np.float32), np.zeros((n_step, 2), np.float32)
            m = max(0, min(n_step, k - c0))
            s_[:m], c_[:m], d_[:m] = sig[c0:c0 + m], rgb[c0:c0 + m], dl[c0:c0 + m]
            O.composite_rays(1, n_step, alive, t, s_, c_, d_, ws, d, im, T_thresh=thr)
            used += m
            if alive[0] < 0:
                break
        n_train = int((w1 > 0).sum())
        if thr == 0.0:
            np.testing.assert_allclose(im, im1[None, 0], rtol=1e-5, atol=1e-6)
        else:
            assert n_train < k
            extra = (1 - ws1[0]) * (1 - np.exp(-sig[n_train] * 0.0034)) * rgb[n_train]
            np.testing.assert_allclose(im[0], im1[0] + extra, rtol=1e-3, atol=2e-6)


@pytest.mark.parametrize("D,C", [(3, 2), (2, 4)])
def test_hash_encode_jacobian_and_adjoints(D, C):
    """dy_dx equals a finite difference of the forward; backward is the adjoint of forward; second backward is the adjoint
    of input-backward w.r.t. (grad, table)."""
    rng = np.random.default_rng(3)
    offsets, pls = O.hash_offsets(D, 6, 4, 10, 64)
    emb = rng.standard_normal((int(offsets[-1]), C)).astype(np.float32)
    B = 64
    x = rng.uniform(0.05, 0.95, size=(B, D)).astype(np.float32)
    out, jac = O.hash_encode_forward(x, emb, offsets, pls, 4, True)
    L = out.shape[0]
    eps = 1e-3
    for d in range(D):
        xp, xm = x.copy(), x.copy(); xp[:, d] += eps; xm[:, d] -= eps
        fd = (O.hash_encode_forward(xp, emb, offsets, pls, 4)[0] - O.hash_encode_forward(xm, emb, offsets, pls, 4)[0]) / (2 * eps)
        an = jac.reshape(B, L, D, C)[:, :, d, :].transpose(1, 0, 2)
        ok = np.abs(fd - an) <= 0.05 * (np.abs(an) + 1) + 2.0      # cells crossed by +-eps break the finite difference
        assert ok.mean() > 0.97
    g = rng.standard_normal(out.shape).astype(np.float32)
    ge, gi = O.hash_encode_backward(g, x, emb, offsets, pls, 4, jac)
    emb2 = rng.standard_normal(emb.shape).astype(np.float32)
    lhs = float((O.hash_encode_forward(x, emb2, offsets, pls, 4)[0].astype(np.float64) * g).sum())
    rhs = float((ge.astype(np.float64) * emb2).sum())
    assert abs(lhs - rhs) <= 1e-4 * (abs(lhs) + 1)                    # <F(E2), g> == <E2, F^T g>
    np.testing.assert_allclose(gi, np.einsum("lbc,bldc->bd", g, jac.reshape(B, L, D, C)), rtol=1e-4, atol=1e-4)
    ggx = rng.standard_normal((B, D)).astype(np.float32)
    gg, g2 = O.hash_encode_second_backward(g, x, emb, offsets, pls, 4, jac, ggx)
    np.testing.assert_allclose(gg, np.einsum("bd,bldc->lbc", ggx, jac.reshape(B, L, D, C)), rtol=1e-4, atol=1e-4)
    # <ggx, grad_inputs(E2)> == <g2, E2>  (grad_inputs is linear in the table)
    _, jac2 = O.hash_encode_forward(x, emb2, offsets, pls, 4, True)
    gi2 = np.einsum("lbc,bldc->bd", g.astype(np.float64), jac2.reshape(B, L, D, C).astype(np.float64))
    lhs, rhs = float((gi2 * ggx).sum()), float((g2.astype(np.float64) * emb2).sum())
    assert abs(lhs - rhs) <= 1e-3 * (abs(lhs) + 1)
    # out-of-range inputs give zeros
    xo = x.copy(); xo[0, 0] = 1.5
    out_o, jac_o = O.hash_encode_forward(xo, emb, offsets, pls, 4, True)
    assert (out_o[:, 0] == 0).all() and (jac_o[0] == 0).all()


def test_hash_offsets_match_reference_table():
    offsets, pls = O.hash_offsets()
    assert offsets[-1] == 6098108 and abs(pls - 1.3819128) < 1e-6
    res = [int(np.ceil(np.float32(np.exp2(np.float32(l) * np.float32(np.log2(pls))) * 16 - 1))) + 1 for l in range(16)]
    assert res == [16, 23, 31, 43, 59, 81, 112, 154, 213, 295, 407, 562, 777, 1073, 1483, 2048]
    g_off, _ = O.grid_offsets()
    assert g_off[-1] == 6119864


def test_freq_and_sh_properties():
    rng = np.random.default_rng(4)
    x = rng.uniform(-1, 1, size=(50, 3)).astype(np.float32)
    out = O.freq_encode_forward(x, 4)
    assert out.shape == (50, 27)
    np.testing.assert_array_equal(out[:, :3], x)
    np.testing.assert_allclose(out[:, 3:6], np.sin(x), atol=1e-6)
    np.testing.assert_allclose(out[:, 6:9], np.cos(x), atol=1e-6)
    np.testing.assert_allclose(out[:, 21:24], np.sin(8 * x), atol=1e-5)
    g = rng.standard_normal(out.shape).astype(np.float32)
    gi = O.freq_encode_backward(g, out, 3, 4)
    eps = 1e-3
    xp = x.copy(); xp[:, 1] += eps
    fd = ((O.freq_encode_forward(xp, 4) - out) * g).sum(-1) / eps
    np.testing.assert_allclose(gi[:, 1], fd, rtol=5e-2, atol=5e-2)
    # SH: orthonormality on the sphere (Monte-Carlo) and the hard-coded low-order terms of the reference table
    d = rng.standard_normal((200000, 3)); d /= np.linalg.norm(d, axis=-1, keepdims=True)
    Y, J = O.sh_encode_forward(d.astype(np.float32), 4, True)
    gram = 4 * np.pi * (Y.T.astype(np.float64) @ Y) / len(d)
    np.testing.assert_allclose(gram, np.eye(16), atol=0.03)
    x_, y_, z_ = d[:, 0], d[:, 1], d[:, 2]
    np.testing.assert_allclose(Y[:, 0], 0.28209479177387814, atol=1e-6)
    np.testing.assert_allclose(Y[:, 1], -0.48860251190291987 * y_, atol=1e-6)       # shencoder.cu:52-54
    np.testing.assert_allclose(Y[:, 2], 0.48860251190291987 * z_, atol=1e-6)
    np.testing.assert_allclose(Y[:, 3], -0.48860251190291987 * x_, atol=1e-6)
    np.testing.assert_allclose(Y[:, 4], 1.0925484305920792 * x_ * y_, atol=1e-6)
    np.testing.assert_allclose(Y[:, 6], 0.94617469575755997 * z_ * z_ - 0.31539156525251999, atol=1e-6)
    np.testing.assert_allclose(Y[:, 8], 0.54627421529603959 * (x_ * x_ - y_ * y_), atol=1e-6)
    np.testing.assert_allclose(Y[:, 15], 0.59004358992664352 * x_ * (-x_ * x_ + 3 * y_ * y_), atol=1e-6)
    np.testing.assert_allclose(J[:, 16 * 2 + 6], 2 * 0.94617469575755997 * z_, atol=1e-5)   # d/dz of Y[6]
    np.testing.assert_allclose(J[:, 16 * 0 + 4], 1.0925484305920792 * y_, atol=1e-5)        # d/dx of Y[4]


def test_oracle_render_three_pass_runs_and_changes_reflective_pixels():
    fp = scene.make_synthetic_field(0, hidden_dim_env=32, ide_degree=4)
    bf = scene.make_bitfield()
    ro, rd = scene.camera_rays(48, 48)
    P = fp.to_oracle()
    st1, st3 = [], []
    a = O.render(P, ro.numpy(), rd.numpy(), bf, stats=st1)
    b = O.render(P, ro.numpy(), rd.numpy(), bf, indir_ref=True, stats=st3)
    assert len(st1) == 1 and len(st3) == 3
    assert a["image"].shape == (48 * 48, 3) and np.isfinite(b["image"]).all()
    hit = a["weights_sum"] > 0.9
    assert hit.sum() > 20 and np.abs(a["image"][hit] - b["image"][hit]).max() > 1e-3      # self-reflections at the box/sphere junction
    miss = a["weights_sum"] == 0          # (rays with 0 < ws <= 0.3 are dropped by ray_mask in the 3-pass scheme, renderer.py:452)
    assert np.abs(a["image"][miss] - b["image"][miss]).max() < 1e-6
