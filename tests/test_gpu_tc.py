"""GPU: the tcgen05 building blocks in isolation (operand image, UMMA descriptors, TMEM read-back)."""
import numpy as np
import pytest
import torch

from swnerf_b200 import _lib
from swnerf_b200._lib import call, stream

pytestmark = pytest.mark.gpu


def _run(mode, A, B, N, K):
    D = torch.empty((128, N), dtype=torch.float32, device="cuda")
    scratch = torch.empty(512 * 1024, dtype=torch.uint8, device="cuda")
    call("swnerf_tc_selftest", mode, A.data_ptr(), B.data_ptr(), D.data_ptr(), N, K, scratch.data_ptr(), stream())
    torch.cuda.synchronize()
    return D


@pytest.mark.parametrize("N,K", [(256, 64), (256, 256), (144, 64), (144, 256), (16, 128)])
def test_umma_kmajor(N, K):
    g = torch.Generator(device="cuda").manual_seed(N + K)
    A = torch.randn(128, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g)
    D = _run(0, A, B, N, K)
    ref = A.half().double() @ B.half().double().t()            # exact products of the fp16-rounded operands
    err = float((D.double() - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, err                                       # fp32 accumulation only


@pytest.mark.parametrize("N,K", [(128, 64), (128, 256), (144, 256), (256, 128)])
def test_umma_a_from_tmem(N, K):
    """A operand written to tensor memory with tcgen05.st (two fp16 per 32-bit column) and consumed from there."""
    g = torch.Generator(device="cuda").manual_seed(N + K + 1)
    A = torch.randn(128, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g)
    D = _run(2, A, B, N, K)
    ref = A.half().double() @ B.half().double().t()
    err = float((D.double() - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, err


@pytest.mark.parametrize("N", [256, 128, 64, 16])
def test_umma_mnmajor(N):
    g = torch.Generator(device="cuda").manual_seed(N)
    P = torch.randn(128, 128, device="cuda", generator=g)        # [samples, channels]
    Q = torch.randn(128, N, device="cuda", generator=g)
    D = _run(1, P, Q, N, 128)
    ref = P.half().double().t() @ Q.half().double()
    err = float((D.double() - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, err


@pytest.mark.parametrize("N,K", [(256, 64), (256, 256), (128, 128)])
def test_umma_cta_pair(N, K):
    """cta_group::2: M = 256 across a cluster of two CTAs, each CTA holding half of B."""
    g = torch.Generator(device="cuda").manual_seed(N + K + 2)
    A = torch.randn(256, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g)
    D = torch.empty((256, N), dtype=torch.float32, device="cuda")
    scratch = torch.empty(1024 * 1024, dtype=torch.uint8, device="cuda")
    call("swnerf_tc_selftest_pair", A.data_ptr(), B.data_ptr(), D.data_ptr(), N, K, 1, 1, None, scratch.data_ptr(), stream())
    torch.cuda.synchronize()
    ref = A.half().double() @ B.half().double().t()
    err = float((D.double() - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, err


@pytest.mark.parametrize("N,K", [(256, 64), (256, 128), (128, 64)])
def test_umma_cta_pair_mnmajor(N, K):
    """cta_group::2 with MN-major operands (the weight-gradient shape): D = P^T Q, K = samples."""
    g = torch.Generator(device="cuda").manual_seed(N + K + 3)
    P = torch.randn(K, 256, device="cuda", generator=g)
    Q = torch.randn(K, N, device="cuda", generator=g)
    D = torch.empty((256, N), dtype=torch.float32, device="cuda")
    scratch = torch.empty(1024 * 1024, dtype=torch.uint8, device="cuda")
    call("swnerf_tc_selftest_pair", P.data_ptr(), Q.data_ptr(), D.data_ptr(), N, K, 0, 1, None, scratch.data_ptr(), stream())
    torch.cuda.synchronize()
    ref = P.half().double().t() @ Q.half().double()
    err = float((D.double() - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, err
