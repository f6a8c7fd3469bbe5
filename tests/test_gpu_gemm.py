"""GPU: the tcgen05 split-precision GEMM core (gemm3.cuh) on plain matrices vs float64 matmul.
Bring-up gate for the STFT / conv / dense kernels that share the core."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(got, ref):
    return float((got.double() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("mode", [1, 0])
@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (128, 128, 128), (100, 16, 64), (256, 256, 256), (1000, 512, 2048),
                                   (20000, 128, 576), (37, 48, 6144)])
def test_gemm3_matches_fp64(engine3s, M, N, K, mode):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + mode)
    A = torch.randn(M, K, generator=g, device="cuda")
    B = torch.randn(N, K, generator=g, device="cuda")
    if mode == 0:                      # fp16 hi part: keep operands inside the fp16 range
        A, B = A * 0.25, B * 0.25
    C = engine3s.dbg_gemm(A, B, mode)
    ref = A.double() @ B.double().T
    err = _rel(C, ref)
    # hi*hi + lo*hi + hi*lo: bf16 pairs carry 16 mantissa bits (~2^-16 relative per operand), fp16 pairs 22;
    # fp32 accumulation over K
    assert err < 6e-5, (M, N, K, mode, err)   # at K = 6144 the fp32 accumulation in TMEM dominates (2e-5) for both formats
    # and it is much better than a single 16-bit pass, i.e. the lo terms are really applied
    one_pass = _rel((A.bfloat16().float() @ B.bfloat16().float().T), ref)
    assert err < one_pass / 20 or one_pass < 1e-6


def test_gemm3_structure(engine3s):
    """Identity / one-hot operands locate layout bugs (row/column permutations) exactly."""
    K, N, M = 128, 64, 256
    A = torch.zeros(M, K, device="cuda")
    A[torch.arange(M), torch.arange(M) % K] = 1.0
    B = torch.arange(N * K, device="cuda", dtype=torch.float32).reshape(N, K) / 64.0
    C = engine3s.dbg_gemm(A, B, 1)
    ref = (A.double() @ B.double().T).float()
    assert torch.equal(C, ref)
