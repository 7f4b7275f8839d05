"""GPU parity of vdk_gemm_tn (tcgen05/TMA GEMM) against a plain fp32 torch reference of the same op.

Inputs are drawn already rounded to the 16-bit input type, so the only differences are the fp32
accumulation order (tolerance 2e-3 relative to the row scale) and, for 16-bit outputs, one final rounding.
"""
import pytest
import torch

from visiondk_b200 import _lib

pytestmark = pytest.mark.gpu

DT = {"bf16": (torch.bfloat16, _lib.DTYPE_BF16), "fp16": (torch.float16, _lib.DTYPE_FP16),
      "fp32": (torch.float32, _lib.DTYPE_FP32)}


def run_gemm(a, b, out_dtype, epilogue=_lib.EPI_NONE, bias=None, gamma=None, residual=None):
    lib = _lib.load()
    M, K = a.shape
    N = b.shape[0]
    d = torch.full((M, N), float("nan"), dtype=DT[out_dtype][0], device=a.device)
    in_code = _lib.DTYPE_BF16 if a.dtype == torch.bfloat16 else _lib.DTYPE_FP16
    rc = lib.vdk_gemm_tn(a.data_ptr(), b.data_ptr(), d.data_ptr(), M, N, K, a.stride(0), b.stride(0), d.stride(0),
                         in_code, DT[out_dtype][1], epilogue, _lib.ptr(bias), _lib.ptr(gamma), _lib.ptr(residual),
                         residual.stride(0) if residual is not None else 0, _lib.stream_ptr())
    _lib.check(rc, "vdk_gemm_tn")
    torch.cuda.synchronize()
    return d


def describe_mismatch(got, ref, tol):
    bad = (got.float() - ref).abs() > tol
    idx = bad.nonzero()
    rows = idx[:, 0]
    cols = idx[:, 1]
    return (f"{int(bad.sum())}/{bad.numel()} wrong; rows%8 hist {torch.bincount(rows % 8, minlength=8).tolist()} "
            f"cols%8 hist {torch.bincount(cols % 8, minlength=8).tolist()} "
            f"row tiles {torch.unique(rows // 128).tolist()[:8]} col tiles {torch.unique(cols // 128).tolist()[:8]} "
            f"first {idx[:4].tolist()} got {got[bad][:4].tolist()} ref {ref[bad][:4].tolist()} nan {int(torch.isnan(got.float()).sum())}")


@pytest.mark.parametrize("in_dtype", ["bf16", "fp16"])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 256, 64), (128, 128, 256), (256, 512, 128),
                                   (1000, 384, 200), (300, 1000, 512), (4096, 512, 128), (77, 8, 8),
                                   (20000, 1024, 256)])
def test_gemm_plain_fp32_out(lib, in_dtype, M, N, K):
    torch.manual_seed(M * 7 + N * 3 + K)
    dt = DT[in_dtype][0]
    a = torch.randn(M, K, device="cuda").to(dt)
    b = torch.randn(N, K, device="cuda").to(dt)
    ref = a.float() @ b.float().t()
    got = run_gemm(a, b, "fp32")
    tol = 2e-3 * (K ** 0.5)
    assert torch.isfinite(got).all(), describe_mismatch(got, ref, tol)
    assert (got - ref).abs().max().item() <= tol, describe_mismatch(got, ref, tol)


def test_gemm_exact_small_integers(lib):
    # integer-valued operands: every product and partial sum is exact, so the result must be bit-exact
    torch.manual_seed(0)
    a = torch.randint(-4, 5, (256, 192), device="cuda").to(torch.bfloat16)
    b = torch.randint(-4, 5, (384, 192), device="cuda").to(torch.bfloat16)
    ref = a.float() @ b.float().t()
    got = run_gemm(a, b, "fp32")
    assert torch.equal(got, ref), describe_mismatch(got, ref, 0.0)


def test_gemm_strided_operands(lib):
    torch.manual_seed(1)
    abuf = torch.randn(500, 320, device="cuda").to(torch.bfloat16)
    bbuf = torch.randn(264, 448, device="cuda").to(torch.bfloat16)
    a, b = abuf[:, :256], bbuf[:, 64:320]  # pitches 320 / 448, 16-byte aligned starts
    ref = a.float() @ b.float().t()
    got = run_gemm(a, b, "fp32")
    assert (got - ref).abs().max().item() <= 2e-3 * 16, describe_mismatch(got, ref, 2e-3 * 16)


@pytest.mark.parametrize("out_dtype", ["bf16", "fp32"])
def test_gemm_bias_gelu(lib, out_dtype):
    torch.manual_seed(2)
    M, N, K = 1568, 512, 128
    a = (0.5 * torch.randn(M, K, device="cuda")).to(torch.bfloat16)
    b = (0.2 * torch.randn(N, K, device="cuda")).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    ref = torch.nn.functional.gelu(a.float() @ b.float().t() + bias)
    got = run_gemm(a, b, out_dtype, _lib.EPI_GELU, bias=bias).float()
    tol = 2e-2 if out_dtype == "bf16" else 2e-3
    assert (got - ref).abs().max().item() <= tol, describe_mismatch(got, ref, tol)


@pytest.mark.parametrize("out_dtype", ["bf16", "fp32"])
def test_gemm_layerscale_residual(lib, out_dtype):
    torch.manual_seed(3)
    M, N, K = 784, 256, 1024
    a = (0.3 * torch.randn(M, K, device="cuda")).to(torch.bfloat16)
    b = (0.1 * torch.randn(N, K, device="cuda")).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    gamma = torch.rand(N, device="cuda")
    res = torch.randn(M, N, device="cuda").to(DT[out_dtype][0])
    ref = res.float() + gamma * (a.float() @ b.float().t() + bias)
    got = run_gemm(a, b, out_dtype, _lib.EPI_SCALE_RESIDUAL, bias=bias, gamma=gamma, residual=res).float()
    tol = 4e-2 if out_dtype == "bf16" else 3e-3
    assert (got - ref).abs().max().item() <= tol, describe_mismatch(got, ref, tol)


def test_gemm_rejects_bad_arguments(lib):
    a = torch.zeros(8, 8, device="cuda", dtype=torch.bfloat16)
    d = torch.zeros(8, 8, device="cuda")
    rc = lib.vdk_gemm_tn(a.data_ptr(), a.data_ptr(), d.data_ptr(), 8, 7, 8, 8, 8, 8, 0, 2, 0, 0, 0, 0, 0, 0)
    assert rc == _lib.VDK_ERR_INVALID and "multiples of 8" in _lib.last_error()
