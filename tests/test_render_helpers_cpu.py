"""Host-side helpers of the batched frame schedule that need no GPU: the synchronisation-free mask compaction of render.render and the
CPU path of the frame de-interleave (envidr_b200/dist.py)."""
import torch


def test_compact_equals_nonzero():
    from envidr_b200.render import _compact
    g = torch.Generator().manual_seed(0)
    for n, p in ((1000, 0.3), (17, 1.0), (64, 0.0), (1, 0.5), (4096, 0.01)):
        m = torch.rand(n, generator=g) < p
        want = m.nonzero().squeeze(-1)
        got = _compact(m, int(m.sum()))
        assert got.dtype == torch.int64 and torch.equal(got, want)


def test_deinterleave_cpu_path_is_index_select():
    from envidr_b200 import dist as D
    g = torch.Generator().manual_seed(1)
    out = torch.rand(50, 8, generator=g)
    inv = torch.randperm(50, generator=g)
    assert torch.equal(D._deinterleave(out, inv), out.index_select(0, inv))


def test_gather_index_inverts_the_tile_sharding():
    """rank-ordered rows of an all-gathered frame -> pixel order: every pixel appears exactly once, ragged borders included."""
    from envidr_b200 import dist as D
    for H, W, ws in ((40, 24, 4), (19, 21, 3), (16, 16, 8)):
        sizes = D.shard_sizes(H, W, ws)
        n_max = max(sizes)
        buf = torch.full((ws * n_max,), -1, dtype=torch.int64)
        for r in range(ws):
            idx = D.tile_shard_indices(H, W, r, ws)
            buf[r * n_max: r * n_max + idx.numel()] = idx
        inv = D._gather_index(H, W, ws, "cpu")
        assert torch.equal(buf[inv], torch.arange(H * W))
