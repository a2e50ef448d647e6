"""The oracle's Adam restatement (oracle.adam_step) against torch.optim.Adam itself on the CPU -- PyTorch is the third-party
dependency whose arithmetic the reference's training loop uses (main_nerf.py:150: betas=(0.9, 0.99), eps=1e-15), and it is
installed here, so the oracle is pinned against the real thing.  Covers: several steps with a decaying learning rate (the
reference's LambdaLR), zero gradients (untouched hash rows: 0 / 1e-15 must stay 0), tiny and large gradients."""
import numpy as np
import torch

from oracle import oracle as O


def test_adam_oracle_matches_torch_adam():
    g = torch.Generator().manual_seed(0)
    n = 20_000
    p = torch.randn(n, generator=g) * 0.1
    p_ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=1e-2, betas=(0.9, 0.99), eps=1e-15, foreach=False)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda it: 0.1 ** min(it / 50, 1))
    po, m, v = p.numpy().copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    for step in range(1, 21):
        grad = torch.randn(n, generator=g) * (10.0 ** torch.randint(-8, 2, (n,), generator=g).float())
        grad[::3] = 0                                                   # rows never touched keep m = v = 0
        if step > 10:
            grad[1::3] = 0                                              # rows touched early, then decaying moments only
        p_ref.grad = grad.clone()
        lr = opt.param_groups[0]["lr"]
        opt.step()
        sched.step()
        po, m, v = O.adam_step(po, grad.numpy(), m, v, step, lr)
        st = opt.state[p_ref]
        np.testing.assert_allclose(m, st["exp_avg"].numpy(), rtol=1e-6, atol=1e-30)
        np.testing.assert_allclose(v, st["exp_avg_sq"].numpy(), rtol=1e-6, atol=1e-38)
        np.testing.assert_allclose(po, p_ref.detach().numpy(), rtol=0, atol=2e-6 * lr / 1e-2 * step)
    assert (po[::3] == p.numpy()[::3]).all()                            # zero-gradient rows are bit-unchanged
