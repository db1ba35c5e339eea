"""GPU: the tcgen05/TMEM/TMA GEMM against a float64 torch reference of the same op (tolerances stated per mode)."""
from ctypes import c_int

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(A, B, M, N, K, precision=1, bf16=False, bias=None, act=0, out32=False, resid=None, split_off=0, ldc=None):
    from sampt_b200 import native
    ctx = native.get_context("cuda")
    ldc = ldc or N
    o16 = None if out32 else torch.zeros((M, ldc), device="cuda", dtype=torch.bfloat16 if bf16 else torch.float16)
    o32 = torch.zeros((M, ldc), device="cuda", dtype=torch.float32) if out32 else None
    native.check(native.lib().sampt_gemm_f16(
        ctx.handle, native.ptr(A), c_int(A.shape[1]), native.ptr(B), c_int(B.shape[1]), c_int(M), c_int(N), c_int(K),
        c_int(precision), c_int(1 if bf16 else 0), native.ptr(bias), c_int(act), native.ptr(o16), native.ptr(o32),
        native.ptr(resid), c_int(ldc), c_int(split_off), native.stream_ptr()), "gemm_f16")
    torch.cuda.synchronize()
    return o32 if out32 else o16


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 256), (256, 512, 128), (4900, 3840, 1280), (300, 384, 128),
                                    (4096, 1280, 5120), (77, 96, 64)])
def test_gemm_fp16_single_pass(M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn((M, K), generator=g)).half().cuda()
    B = (torch.randn((N, K), generator=g) / K ** 0.5).half().cuda()
    bias = torch.randn((N,), generator=g).cuda()
    out = _run(A, B, M, N, K, bias=bias, out32=True)
    ref = A.double().cpu() @ B.double().cpu().T + bias.double().cpu()
    err = (out.cpu().double() - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()) * 1e-1 + 1e-4, err  # fp32 accumulation of exact fp16 products


def test_gemm_bf16_and_gelu_fp16_out():
    M, N, K = 512, 512, 256
    g = torch.Generator().manual_seed(1)
    A = torch.randn((M, K), generator=g).bfloat16().cuda()
    B = (torch.randn((N, K), generator=g) / K ** 0.5).bfloat16().cuda()
    out = _run(A, B, M, N, K, bf16=True, act=1)
    ref = torch.nn.functional.gelu(A.double().cpu() @ B.double().cpu().T)
    assert (out.cpu().double() - ref).abs().max() < 3e-2  # bf16 output rounding


def test_gemm_residual_and_split_output():
    M, N, K = 384, 256, 128
    g = torch.Generator().manual_seed(2)
    A = torch.randn((M, K), generator=g).half().cuda()
    B = (torch.randn((N, K), generator=g) / K ** 0.5).half().cuda()
    resid = torch.randn((M, N), generator=g).cuda()
    out = _run(A, B, M, N, K, out32=True, resid=resid)
    ref = A.double().cpu() @ B.double().cpu().T + resid.double().cpu()
    assert (out.cpu().double() - ref).abs().max() < 1e-4
    # split fp16 output: hi|lo at column offset N reconstructs the fp32 value to ~2^-22
    o = _run(A, B, M, N, K, split_off=N, ldc=2 * N)
    rec = o[:, :N].double().cpu() + o[:, N:].double().cpu()
    ref = A.double().cpu() @ B.double().cpu().T
    assert (rec - ref).abs().max() < 2e-5


@pytest.mark.parametrize("precision", [2, 3])
def test_gemm_split_precision(precision):
    """operands carried as fp16 hi|lo: 3 passes reach ~fp32 accuracy for fp32 A and B; 2 passes are exact in B (weights)
    for fp16-representable A."""
    M, N, K = 640, 768, 512
    g = torch.Generator().manual_seed(3)
    A32 = torch.randn((M, K), generator=g)
    B32 = torch.randn((N, K), generator=g) / K ** 0.5
    if precision == 2:
        A32 = A32.half().float()  # activations are fp16 in this mode

    def split(x):
        hi = x.half()
        lo = (x - hi.float()).half()
        return torch.cat([hi, lo], dim=1).cuda()

    A = split(A32) if precision == 3 else A32.half().cuda()
    B = split(B32)
    out = _run(A, B, M, N, K, precision=precision, out32=True)
    ref = A32.double() @ B32.double().T
    err = (out.cpu().double() - ref).abs().max().item()
    assert err < 5e-5, err
    # and it must be far better than the single-pass result
    single = _run(A32.half().cuda(), B32.half().cuda(), M, N, K, out32=True)
    err1 = (single.cpu().double() - (A32.double() @ B32.double().T)).abs().max().item()
    assert err < err1 / 20
