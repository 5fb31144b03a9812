"""tcgen05 split-bf16 linear layer against torch fp64 on the same device (kernel-level parity)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPES = [(128, 16, 16), (4096, 80, 40), (20000, 40, 80), (5000, 80, 120), (3000, 120, 80), (2500, 40, 40),
          (9000, 480, 40), (7000, 40, 480), (2048, 48, 160), (33000, 80, 80)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tc_gemm_matches_fp64(cuda_lib, M, N, K):
    import torch
    from clsr_b200.engine import Engine
    eng = Engine(100, 10, 10, max_rows=8, seq_len=50, training=False)
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(K, N, device="cuda", generator=g) * 0.3
    b = torch.randn(N, device="cuda", generator=g)
    ref = (A.double() @ W.double() + b.double())
    for mode, tol in ((0, 2e-6), (1, 2e-5)):
        C = torch.full((M, N), float("nan"), device="cuda")
        eng._check(eng.lib.clsr_debug_gemm(eng.h, M, N, K, A.data_ptr(), K, W.data_ptr(), N, b.data_ptr(),
                                            C.data_ptr(), N, mode))
        eng.synchronize()
        err = ((C.double() - ref).abs().max() / ref.abs().max()).item()
        assert err < tol, (mode, err)


DW_SHAPES = [(4096, 40, 80), (20000, 80, 40), (5000, 120, 80), (9000, 40, 480), (7001, 40, 40), (33000, 80, 120),
             (2048, 40, 160)]


@pytest.mark.parametrize("M,K,N", DW_SHAPES)
def test_tc_dwgemm_matches_fp64(cuda_lib, M, K, N):
    import torch
    from clsr_b200.engine import Engine
    eng = Engine(100, 10, 10, max_rows=8, seq_len=50, training=False)
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(M, N, device="cuda", generator=g)
    ref = A.double().t() @ B.double()
    refc = B.double().sum(0)
    for mode, tol in ((0, 2e-5), (1, 5e-5)):
        dW = torch.zeros(K, N, device="cuda")
        cs = torch.zeros(N, device="cuda")
        eng._check(eng.lib.clsr_debug_dwgemm(eng.h, M, K, N, A.data_ptr(), K, B.data_ptr(), N, dW.data_ptr(), N,
                                              cs.data_ptr(), mode))
        eng.synchronize()
        err = ((dW.double() - ref).abs().max() / ref.abs().max()).item()
        errc = ((cs.double() - refc).abs().max() / refc.abs().max()).item()
        assert err < tol and errc < tol, (mode, err, errc)
