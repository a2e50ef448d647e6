"""Tensor-core path on hardware: the tcgen05 probe pins the UMMA shared-memory descriptor / operand layout conventions
(csrc/tc_common.cuh) against a plain fp32 matmul of the fp16-rounded operands."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("N,K", [(256, 16), (256, 64), (64, 32), (16, 256), (128, 80)])
def test_tcgen05_probe_matches_matmul(dev, N, K):
    from envidr_b200._lib import check, lib, ptr, stream
    g = torch.Generator().manual_seed(N + K)
    A = torch.randn(128, K, generator=g).to(dev)
    B = torch.randn(N, K, generator=g).to(dev)
    ref = A.half().float() @ B.half().float().T
    D = torch.full((128, N), float("nan"), device=dev)
    check(lib().envidr_tc_probe(ptr(A), ptr(B), ptr(D), N, K, 0, stream()), "tc_probe")
    torch.cuda.synchronize()
    err = (D - ref).abs().max().item()
    if not err < 1e-3:
        D1 = torch.full((128, N), float("nan"), device=dev)
        check(lib().envidr_tc_probe(ptr(A), ptr(B), ptr(D1), N, K, 1, stream()), "tc_probe")
        torch.cuda.synchronize()
        err1 = (D1 - ref).abs().max().item()
        pytest.fail(f"variant 0 (LBO = K-chunk stride, SBO = 8-row stride) max err {err}; swapped variant max err {err1}; "
                    f"D[0,:4]={D[0,:4].tolist()} ref[0,:4]={ref[0,:4].tolist()}")
