"""tcgen05 building blocks (csrc/umma.cuh): descriptor / TMEM / layout conventions pinned against a float64 GEMM.
The 3xTF32 split must deliver fp32-level accuracy (that is what keeps the tensor-core path inside the parity tolerances)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N", [(8, 16), (64, 32), (64, 64), (128, 64), (96, 32), (64, 192)])
def test_umma_3xtf32_gemm(cuda, K, N):
    from etch_b200 import _lib as L
    g = torch.Generator().manual_seed(K * 1000 + N)
    A = torch.randn(128, K, generator=g).to(cuda)
    B = torch.randn(N, K, generator=g).to(cuda)
    C = torch.full((128, N), float("nan"), device=cuda)
    L.call("umma_selftest", L.ptr(A), L.ptr(B), L.ptr(C), K, N)
    torch.cuda.synchronize()
    ref = (A.double() @ B.double().t())
    err = (C.double() - ref).abs().max().item() / ref.abs().max().item()
    print("K=%d N=%d rel err %.3e" % (K, N, err))
    assert err < 5e-6, err
