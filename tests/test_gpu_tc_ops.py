"""GPU parity of the tcgen05 GEMM / implicit-GEMM conv against plain PyTorch fp32 on the same bf16 inputs.

Tolerance: inputs are exactly representable bf16, accumulation is fp32 on both sides, the output is rounded to bf16 once:
|err| <= 2^-8 * |ref| + small absolute slack for cancellation (atol scaled by sqrt(K))."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _check(out, ref, K):
    out, ref = out.float(), ref.float()
    scale = ref.abs().max().item() + 1e-6
    err = (out - ref).abs().max().item()
    assert err <= 1.0 / 128 * scale + 1e-3 * math.sqrt(K), (err, scale)
    # and on average much tighter than the worst case
    assert (out - ref).abs().mean().item() <= 4e-3 * ref.abs().mean().item() + 1e-5


@pytest.mark.parametrize('M,N,K', [(128, 256, 64), (256, 320, 320), (4096, 1280, 640), (4928, 640, 768), (300, 2560, 128),
                                   (128, 16, 64), (1000, 40, 192), (8192, 960, 320), (128, 64, 4096)])
def test_gemm_plain(M, N, K):
    from mvedit_b200 import tc_ops
    g = torch.Generator(device='cuda').manual_seed(M + N + K)
    a = torch.randn(M, K, device='cuda', generator=g).bfloat16()
    w = (torch.randn(N, K, device='cuda', generator=g) / math.sqrt(K)).bfloat16()
    out = tc_ops.gemm(a, w)
    _check(out, a.float() @ w.float().t(), K)


def test_gemm_epilogue():
    from mvedit_b200 import tc_ops
    g = torch.Generator(device='cuda').manual_seed(1)
    M, N, K, rpg = 1024, 320, 256, 256
    a = torch.randn(M, K, device='cuda', generator=g).bfloat16()
    w = (torch.randn(N, K, device='cuda', generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device='cuda', generator=g)
    rb = torch.randn(M // rpg, N, device='cuda', generator=g)
    res = torch.randn(M, N, device='cuda', generator=g).bfloat16()
    for act, f in (('silu', torch.nn.functional.silu), ('gelu', torch.nn.functional.gelu), (None, lambda t: t)):
        out = tc_ops.gemm(a, w, bias=bias, row_bias=rb, rows_per_group=rpg, residual=res, act=act, alpha=0.5)
        ref = f(a.float() @ w.float().t() + bias + rb.repeat_interleave(rpg, 0)) * 0.5 + res.float()
        _check(out, ref, K)


@pytest.mark.parametrize('B,H,W,Cin,Cout', [(2, 8, 8, 64, 64), (4, 16, 16, 128, 320), (2, 32, 32, 320, 640), (2, 64, 64, 64, 320),
                                            (1, 128, 128, 64, 32), (2, 64, 64, 320, 4), (1, 256, 256, 64, 16)])
def test_conv3x3(B, H, W, Cin, Cout):
    from mvedit_b200 import tc_ops
    g = torch.Generator(device='cuda').manual_seed(B + H + Cin + Cout)
    x = torch.randn(B, H, W, Cin, device='cuda', generator=g).bfloat16()
    w = (torch.randn(Cout, 3, 3, Cin, device='cuda', generator=g) / math.sqrt(9 * Cin)).bfloat16()
    bias = torch.randn(Cout, device='cuda', generator=g)
    rb = torch.randn(B, Cout, device='cuda', generator=g)
    out = tc_ops.conv3x3(x, w, bias=bias, row_bias=rb)
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, padding=1)
    ref = (ref + rb[:, :, None, None]).permute(0, 2, 3, 1)
    _check(out, ref, 9 * Cin)
