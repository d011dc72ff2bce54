"""tcgen05 building blocks through the C ABI test hooks: UMMA descriptor conventions and the input-projection GEMM."""
import pytest
import torch

from las_pytorch_b200 import _cabi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("a_sw,b_sw", [(0, 0), (1, 1), (1, 0), (0, 1)])
@pytest.mark.parametrize("N,K", [(16, 256), (64, 128), (256, 64)])
def test_umma_layouts(N, K, a_sw, b_sw):
    lib = _cabi.load_library()
    g = torch.Generator().manual_seed(N * 1000 + K)
    a = torch.randn(128, K, generator=g).cuda().to(torch.bfloat16)
    b = torch.randn(N, K, generator=g).cuda().to(torch.bfloat16)
    d = torch.full((128, N), float("nan"), device="cuda")
    _cabi.check(lib.las_debug_umma_probe(_cabi.ptr(a), _cabi.ptr(b), _cabi.ptr(d), N, K, a_sw, b_sw, 0, _cabi.current_stream_ptr()))
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t()
    assert float((d - ref).abs().max()) <= 1e-3 * float(ref.abs().max())


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 128, 80), (1000, 2048, 1024), (389, 280, 96), (2048, 1024, 512)])
def test_input_projection_gemm(M, N, K):
    lib = _cabi.load_library()
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).cuda().to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda().to(torch.bfloat16)
    bias = torch.randn(N, generator=g).cuda()
    c = torch.full((M, N), float("nan"), device="cuda")
    _cabi.check(lib.las_debug_gemm_bf16(_cabi.ptr(a), _cabi.ptr(w), _cabi.ptr(bias), _cabi.ptr(c), M, N, K, _cabi.current_stream_ptr()))
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias
    assert float((c - ref).abs().max()) <= 2e-3


@pytest.mark.parametrize("M,N,K", [(300, 128, 80), (389, 280, 96), (4096, 2048, 256), (1000, 136, 512)])
def test_gemm_tma_store_epilogue_matches_direct_stores(M, N, K):
    """The epilogue stages 32x32 blocks in shared memory and writes them with TMA stores (M / N tails clipped by the tensor map);
    las_debug_set_option(8, 1) selects the row-per-thread global stores that remain as the fallback: bit-identical outputs, and
    nothing is written outside [M, N]."""
    lib = _cabi.load_library()
    g = torch.Generator().manual_seed(7 * M + N + K)
    a = torch.randn(M, K, generator=g).cuda().to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda().to(torch.bfloat16)
    bias = torch.randn(N, generator=g).cuda()
    outs = []
    for direct in (1, 0):
        buf = torch.full((M + 64, N), float("nan"), device="cuda")  # guard rows behind the output
        try:
            lib.las_debug_set_option(8, direct)
            _cabi.check(lib.las_debug_gemm_bf16(_cabi.ptr(a), _cabi.ptr(w), _cabi.ptr(bias), _cabi.ptr(buf), M, N, K, _cabi.current_stream_ptr()))
            torch.cuda.synchronize()
        finally:
            lib.las_debug_set_option(8, 0)
        assert torch.isnan(buf[M:]).all()
        outs.append(buf[:M].clone())
    assert torch.equal(outs[0], outs[1])
    ref = a.float() @ w.float().t() + bias
    assert float((outs[1] - ref).abs().max()) <= 2e-3
