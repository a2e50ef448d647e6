"""FusedAdam (csrc/optim.cu through envidr_b200.optim) on the B200 against
  * the CPU oracle (oracle.adam_step, pinned to torch.optim.Adam by tests/test_optim_cpu.py),
  * torch.optim.Adam on the same GPU (foreach, the implementation the reference's training loop runs): identical moments and
    parameters to rounding over 30 steps with the reference's LambdaLR decay and per-group learning rates,
and its contract: ragged / unaligned / empty tensors, more than 64 tensors in one step, the fused zero_grad, untouched rows
bit-unchanged, state_dict interchange with torch.optim.Adam."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _params(dev, sizes, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(s, generator=g) * 0.1).to(dev) for s in sizes]


def _grads(sizes, step, seed=1):
    g = torch.Generator().manual_seed(seed * 1000 + step)
    out = []
    for s in sizes:
        n = int(np.prod(s)) if not isinstance(s, int) else s
        gr = torch.randn(n, generator=g) * (10.0 ** torch.randint(-8, 2, (n,), generator=g).float())
        gr[::3] = 0
        out.append(gr.view(s) if not isinstance(s, int) else gr)
    return out


def test_fused_adam_matches_torch_and_oracle(dev):
    from envidr_b200.optim import FusedAdam
    from oracle import oracle as O
    sizes = [(4099, 2), (64, 32), (64,), (15, 64), 1, (3, 7), 100_003]
    a = [p.clone().requires_grad_(True) for p in _params(dev, sizes)]
    b = [p.clone().requires_grad_(True) for p in _params(dev, sizes)]
    groups = lambda ps: [{"params": ps[:1], "lr": 2e-2}, {"params": ps[1:4], "lr": 1e-3}, {"params": ps[4:], "lr": 5e-3}]
    ours = FusedAdam(groups(a), betas=(0.9, 0.99), eps=1e-15)
    ref = torch.optim.Adam(groups(b), betas=(0.9, 0.99), eps=1e-15)
    lam = lambda it: 0.1 ** min(it / 20, 1)
    s1 = torch.optim.lr_scheduler.LambdaLR(ours, lam)
    s2 = torch.optim.lr_scheduler.LambdaLR(ref, lam)
    orc = [(p.detach().cpu().numpy().reshape(-1).copy(), np.zeros(p.numel(), np.float32), np.zeros(p.numel(), np.float32)) for p in a]
    lrs0 = [2e-2, 1e-3, 1e-3, 1e-3, 5e-3, 5e-3, 5e-3]
    for step in range(1, 31):
        gs = _grads(sizes, step)
        for pa, pb, g in zip(a, b, gs):
            pa.grad = g.to(dev).clone()
            pb.grad = g.to(dev).clone()
        lr_now = [g["lr"] for g in ours.param_groups]
        ours.step(); ref.step()
        s1.step(); s2.step()
        for i, (pa, pb, g) in enumerate(zip(a, b, gs)):
            lr = lrs0[i] * lam(step - 1)
            assert abs(lr - lr_now[[0, 1, 1, 1, 2, 2, 2][i]]) < 1e-12
            po, m, v = O.adam_step(orc[i][0], g.numpy().reshape(-1), orc[i][1], orc[i][2], step, lr)
            orc[i] = (po, m, v)
            sa, sb = ours.state[pa], ref.state[pb]
            for name, o in (("exp_avg", m), ("exp_avg_sq", v)):
                np.testing.assert_allclose(sa[name].cpu().numpy().reshape(-1), sb[name].cpu().numpy().reshape(-1), rtol=2e-6, atol=1e-38)
                np.testing.assert_allclose(sa[name].cpu().numpy().reshape(-1), o, rtol=2e-6, atol=1e-38)
            # each step rounds the parameter once (<= 1/2 ulp) and the update itself agrees to ~1e-6 relative
            tol = 2e-6 * lrs0[i] * step
            np.testing.assert_allclose(pa.detach().cpu().numpy(), pb.detach().cpu().numpy(), rtol=6e-7, atol=tol)
            np.testing.assert_allclose(pa.detach().cpu().numpy().reshape(-1), po, rtol=6e-7, atol=tol)
    sd = ours.state_dict()                       # refreshes the `step` tensors torch.optim.Adam keeps per parameter
    assert all(float(st["step"]) == 30 for st in sd["state"].values())
    assert all(float(ref.state[pb]["step"]) == 30 for pb in b)
    # rows whose gradient was always zero never moved
    p0 = _params(dev, sizes)
    for pa, p in zip(a, p0):
        assert torch.equal(pa.detach().reshape(-1)[::3], p.reshape(-1)[::3])
    # state_dict interchange: torch.optim.Adam loads ours and continues identically
    ref2 = torch.optim.Adam(groups([p.clone().detach().requires_grad_(True) for p in a]), betas=(0.9, 0.99), eps=1e-15)
    import copy
    ref2.load_state_dict(copy.deepcopy(ours.state_dict()))     # load_state_dict shares tensors that need no cast: copy first
    gs = _grads(sizes, 99)
    for pa, pc, g in zip(a, [p for gr in ref2.param_groups for p in gr["params"]], gs):
        pa.grad = g.to(dev).clone(); pc.grad = g.to(dev).clone()
    ours.step(); ref2.step()
    for pa, pc in zip(a, [p for gr in ref2.param_groups for p in gr["params"]]):
        np.testing.assert_allclose(pa.detach().cpu().numpy(), pc.detach().cpu().numpy(), rtol=6e-7, atol=1e-6)


def test_fused_adam_edge_cases(dev):
    from envidr_b200.optim import FusedAdam
    # > 64 tensors (two launches), an empty tensor, unaligned views, fused zero_grad, tensors without a gradient
    base = torch.randn(70 * 33 + 1, device=dev)
    ps = [base[1 + 33 * i: 1 + 33 * (i + 1)].detach().clone().requires_grad_(True) for i in range(70)]
    big = torch.randn(1_000_003, device=dev)
    view = big[3:].detach()                      # 12-byte offset: not 16-byte aligned -> scalar path
    view.requires_grad_(True)
    empty = torch.zeros(0, device=dev, requires_grad=True)
    nograd = torch.ones(5, device=dev, requires_grad=True)
    allp = ps + [view, empty, nograd]
    ref_p = [p.detach().clone().requires_grad_(True) for p in allp]
    ours = FusedAdam(allp, lr=1e-2, betas=(0.9, 0.99), eps=1e-15, zero_grad=True)
    ref = torch.optim.Adam(ref_p, lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    for step in range(3):
        for pa, pb in zip(allp[:-1], ref_p[:-1]):
            g = torch.randn_like(pa)
            if pa.grad is None:
                pa.grad = g.clone()
            else:
                assert float(pa.grad.abs().sum()) == 0.0 if pa.numel() else True       # cleared by the fused pass
                pa.grad.copy_(g)
            pb.grad = g.clone()
        ours.step(); ref.step()
        for pa, pb in zip(allp, ref_p):
            np.testing.assert_allclose(pa.detach().cpu().numpy(), pb.detach().cpu().numpy(), rtol=3e-7, atol=1e-7 * (step + 1))
    assert torch.equal(nograd.detach(), torch.ones(5, device=dev)) and len(ours.state[nograd]) == 0
    with pytest.raises(ValueError):
        FusedAdam(ps, lr=1e-2, weight_decay=0.1)
    cpu_p = torch.zeros(4, requires_grad=True)
    cpu_p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        FusedAdam([cpu_p], lr=1e-2).step()       # no CPU fallback
