"""envidr_b200 -- B200 (sm_100a) implementation of the ENVIDR volumetric-render hot path.

Layout (only what the path needs):
    csrc/            hand-written CUDA kernels + the C ABI (include/envidr_b200.h) -> _lib/libenvidr_b200.so
    backend.py       drop-in `_backend` objects with the reference's pybind signatures
    raymarching.py, hashencoder.py, gridencoder.py, freqencoder.py, shencoder.py, ide_encoder.py
                     the reference's torch.autograd.Function / nn.Module operator surface
    field.py         fused per-sample field (hash grid + SDF/env/diffuse/colour MLPs) host side
    render.py        fused inference loop, run_cuda / NeRFRenderer.render mirrors, install()
    train.py         run_cuda training branch on the library's operators, CUDA-graph step; linear_tc.py: dense layers on tcgen05
    density.py       occupancy-grid maintenance (update_extra_state, mark_untrained_grid), install(model)
    epilogue.py      ray generation (get_rays) and the fused loss epilogue of a training step
    optim.py         FusedAdam (one launch over all parameter tensors), drop-in for torch.optim.Adam as the reference configures it
    checkpoint.py    the reference's checkpoint / published-weight formats, EMA (host logic only)
    dist.py          ray sharding across GPUs + NCCL all-gather of the image
    scene.py         seeded synthetic scenes / cameras for tests and bench
There is no CPU fallback: every compute entry raises if the CUDA library is missing.
"""
__version__ = "0.1.0"
