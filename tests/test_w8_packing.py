"""CPU: the weight operand of the fp8-corrected GEMM (include/sampt_b200.h: sampt_gemm_f8c; host side
`ImageEncoderViT._w8`) decodes back to the weights, and the three segments emulated in float64 reproduce x @ w.T to ~1e-5 -- the
accuracy argument of DESIGN.md §5 checked without a GPU."""
import math
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sam-pt_b200"))


def _e4m3(t):
    return t.clamp(-448.0, 448.0).to(torch.float8_e4m3fn).float()


def _unpack(w8, K):
    raw = w8.contiguous().view(torch.uint8)                      # (N, 4K) bytes: [fp16 x K | e4m3 x K | e4m3 x K]
    hi16 = raw[:, : 2 * K].contiguous().view(torch.float16).float()
    hi8 = raw[:, 2 * K: 3 * K].contiguous().view(torch.float8_e4m3fn).float()
    lo8 = raw[:, 3 * K:].contiguous().view(torch.float8_e4m3fn).float()
    return hi16, hi8, lo8


@pytest.mark.parametrize("std,outlier", [(0.02, 0.0), (0.02, 0.9), (1.5, 0.0), (1e-4, 0.0)])
def test_w8_layout_and_scales(std, outlier):
    from segment_anything.modeling.image_encoder import ImageEncoderViT
    g = torch.Generator().manual_seed(3)
    N, K = 96, 256
    w = torch.randn((N, K), generator=g) * std
    if outlier:
        w[5, 7] = outlier
    w8, scale = ImageEncoderViT._w8(w)
    assert w8.dtype == torch.float16 and w8.shape == (N, 2 * K) and scale.shape == (1,)
    s = -math.log2(float(scale))
    assert s == int(s)                                            # a power of two: the epilogue's multiply is exact
    hi16, hi8, lo8 = _unpack(w8, K)
    S = 2.0 ** s
    assert float(w.abs().max()) * S <= 2.0 ** 15                 # fp16(w * 2^s) cannot overflow
    assert float(w.abs().max()) * S > 2.0 ** 13                  # ... and uses the range (largest exponent that fits)
    assert torch.equal(hi16, (w * S).half().float())
    rem = w * S - hi16
    assert (hi8 - w * 2.0 ** (s - 12)).abs().max() <= 2.0 ** -4 * (w.abs().max() * 2.0 ** (s - 12)) + 2.0 ** -10   # e4m3 rounding
    assert (lo8 - rem * 8.0).abs().max() <= 2.0 ** -4 * (rem.abs().max() * 8.0) + 2.0 ** -10
    # the fp16 block plus the scaled remainder block give the weight back to ~2^-16 relative
    rec = (hi16 + lo8 / 8.0) / S
    assert (rec - w).abs().max() <= 2.0 ** -15 * float(w.abs().max())


def test_three_segments_reproduce_the_product():
    from segment_anything.modeling.image_encoder import ImageEncoderViT
    g = torch.Generator().manual_seed(4)
    M, N, K = 128, 160, 1280
    x = torch.randn((M, K), generator=g)
    x[:, :40] *= 8.0
    w = torch.randn((N, K), generator=g) * 0.02
    w8, scale = ImageEncoderViT._w8(w)
    hi16, hi8, lo8 = _unpack(w8, K)
    xh = x.half().float()
    a_lo8, a_hi8 = _e4m3((x - xh) * 4096.0), _e4m3(x * 0.125)   # tc_api.cuh: F8_LO_SCALE, F8_HI_SCALE
    acc = xh.double() @ hi16.double().T + a_lo8.double() @ hi8.double().T + a_hi8.double() @ lo8.double().T
    y = acc * float(scale)
    ref = x.double() @ w.double().T
    rel = float((y - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())
    two_pass = float(((xh.double() @ w.double().T) - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())
    assert rel < 2e-5, rel
    assert rel < two_pass / 10, (rel, two_pass)
