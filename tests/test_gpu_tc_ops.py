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


@pytest.mark.parametrize('B,heads,Sq,Skv,d', [(2, 8, 256, 256, 40), (1, 8, 1024, 1024, 80), (2, 8, 256, 256, 160), (3, 8, 64, 64, 160),
                                              (2, 8, 1024, 77, 40), (2, 8, 256, 77, 80), (1, 8, 4096, 4096, 40), (2, 2, 128, 128, 64),
                                              (1, 8, 512, 93, 160), (2, 1, 200, 300, 64)])
def test_attention(B, heads, Sq, Skv, d):
    """vs F.scaled_dot_product_attention in fp32 on the same bf16 inputs.  P is rounded to bf16 before P.V (as every flash
    kernel does): tolerance 2e-2 of the output scale."""
    from mvedit_b200 import tc_ops
    g = torch.Generator(device='cuda').manual_seed(Sq + Skv + d)
    C = heads * d
    # q comes as a slice of a wider fused projection (row stride 3C), like in the UNet
    qkv = torch.randn(B, Sq, 3 * C, device='cuda', generator=g).bfloat16()
    q = qkv[:, :, :C]
    k = torch.randn(B, Skv, C, device='cuda', generator=g).bfloat16()
    v = torch.randn(B, Skv, C, device='cuda', generator=g).bfloat16()
    out = tc_ops.attention(q, k, v, heads)
    sh = lambda t, S: t.float().reshape(B, S, heads, d).transpose(1, 2)
    ref = torch.nn.functional.scaled_dot_product_attention(sh(q, Sq), sh(k, Skv), sh(v, Skv)).transpose(1, 2).reshape(B, Sq, C)
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-2 * ref.abs().max().item() + 1e-3, (err, ref.abs().max().item())
    assert (out.float() - ref).abs().mean().item() <= 4e-3 * ref.abs().mean().item() + 1e-4


def test_attention_peaked_softmax_rescale():
    """large score range -> the running max moves in later key blocks, exercising the TMEM O-rescale path."""
    from mvedit_b200 import tc_ops
    g = torch.Generator(device='cuda').manual_seed(0)
    B, heads, S, d = 1, 8, 512, 40
    q = (torch.randn(B, S, heads * d, device='cuda', generator=g) * 4).bfloat16()
    k = (torch.randn(B, S, heads * d, device='cuda', generator=g) * 4).bfloat16()
    k[:, 300:] *= 2      # later keys dominate
    v = torch.randn(B, S, heads * d, device='cuda', generator=g).bfloat16()
    out = tc_ops.attention(q, k, v, heads)
    sh = lambda t: t.float().reshape(B, S, heads, d).transpose(1, 2)
    ref = torch.nn.functional.scaled_dot_product_attention(sh(q), sh(k), sh(v)).transpose(1, 2).reshape(B, S, heads * d)
    assert (out.float() - ref).abs().max().item() <= 3e-2 * ref.abs().max().item()


@pytest.mark.parametrize('B,HW,C', [(2, 64, 64), (3, 1024, 320), (2, 4096, 640), (2, 256, 1280), (1, 64, 2560), (2, 4096, 960)])
@pytest.mark.parametrize('silu', [False, True])
def test_groupnorm(B, HW, C, silu):
    from mvedit_b200 import tc_ops
    g = torch.Generator(device='cuda').manual_seed(C)
    x = (torch.randn(B, HW, C, device='cuda', generator=g) * 2 + 0.5).bfloat16()
    gamma, beta = torch.randn(C, device='cuda', generator=g), torch.randn(C, device='cuda', generator=g)
    y = tc_ops.groupnorm(x, gamma, beta, 32, 1e-5, silu)
    ref = torch.nn.functional.group_norm(x.float().transpose(1, 2), 32, gamma, beta, 1e-5).transpose(1, 2)
    if silu:
        ref = torch.nn.functional.silu(ref)
    assert (y.float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
    assert (y.float() - ref).abs().mean().item() <= 4e-3 * ref.abs().mean().item() + 1e-4


@pytest.mark.parametrize('M,F,K', [(300, 1280, 320), (4096 + 77, 2560, 640), (128, 128, 64)])
def test_gemm_fused_geglu_epilogue(M, F, K):
    """act='geglu': the feed-forward's first projection with GEGLU in the epilogue == unfused GEMM + GEGLU (torch fp32 reference)."""
    from mvedit_b200 import tc_ops
    g = torch.Generator(device='cuda').manual_seed(M + F)
    x = (torch.randn(M, K, device='cuda', generator=g) * 0.5).bfloat16()
    w = (torch.randn(2 * F, K, device='cuda', generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(2 * F, device='cuda', generator=g) * 0.1
    h = x.float() @ w.float().t() + b
    val, gate = h.chunk(2, dim=-1)
    ref = val * torch.nn.functional.gelu(gate)
    wi, bi = tc_ops.geglu_interleave(w, b)
    out = tc_ops.gemm(x, wi, bias=bi, act='geglu')
    assert out.shape == (M, F)
    assert (out.float() - ref).abs().max().item() <= 1.5e-2 * ref.abs().max().item()
    assert (out.float() - ref).abs().mean().item() <= 4e-3 * ref.abs().mean().item() + 1e-5


def test_layernorm_geglu_upsample_im2col_layout():
    from mvedit_b200 import tc_ops
    g = torch.Generator(device='cuda').manual_seed(0)
    for C in (64, 320, 640, 1280, 960):
        x = torch.randn(777, C, device='cuda', generator=g).bfloat16()
        gamma, beta = torch.randn(C, device='cuda', generator=g), torch.randn(C, device='cuda', generator=g)
        ref = torch.nn.functional.layer_norm(x.float(), (C,), gamma, beta, 1e-5)
        assert (tc_ops.layernorm(x, gamma, beta).float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
    h = torch.randn(500, 2 * 1280, device='cuda', generator=g).bfloat16()
    a, gate = h.float().chunk(2, dim=-1)
    ref = a * torch.nn.functional.gelu(gate)
    assert (tc_ops.geglu(h).float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    x = torch.randn(2, 8, 8, 64, device='cuda', generator=g).bfloat16()
    up = tc_ops.upsample2x(x)
    assert torch.equal(up, x.repeat_interleave(2, 1).repeat_interleave(2, 2))
    cols = tc_ops.im2col3x3s2(x)
    unf = torch.nn.functional.unfold(x.float().permute(0, 3, 1, 2), 3, padding=1, stride=2)      # [B, C*9, L]  (c-major, tap-minor)
    unf = unf.view(2, 64, 9, 16).permute(0, 3, 2, 1).reshape(2 * 16, 9 * 64)                      # -> tap-major, c-minor
    assert torch.equal(cols.float(), unf)
    lat = torch.randn(2, 4, 8, 8, device='cuda', generator=g)
    nh = tc_ops.nchw_to_nhwc_pad(lat, 64)
    assert torch.equal(nh[..., :4].float(), lat.bfloat16().float().permute(0, 2, 3, 1)) and nh[..., 4:].abs().sum() == 0


@pytest.mark.parametrize('M,N,K', [(40000, 128, 128), (65536, 640, 1280), (41000, 320, 1280), (38000, 384, 64)])
def test_gemm_tall_tiles(M, N, K):
    """256-row CTA tiles (two 128-row MMAs per B stage): 256 x 128 with two accumulator stages and 256 x 160 with one; M not a
    multiple of 256 exercises the TMA zero fill / row guard of the second sub-tile; epilogue operands included."""
    from mvedit_b200 import tc_ops
    g = torch.Generator(device='cuda').manual_seed(M + N + K)
    a = torch.randn(M, K, device='cuda', generator=g).bfloat16()
    w = (torch.randn(N, K, device='cuda', generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device='cuda', generator=g)
    res = torch.randn(M, N, device='cuda', generator=g).bfloat16()
    out = tc_ops.gemm(a, w, bias=bias, residual=res, act='silu')
    _check(out, torch.nn.functional.silu(a.float() @ w.float().t() + bias) + res.float(), K)


@pytest.mark.parametrize('B,H,W,Cin,Cout', [(4, 128, 128, 128, 128), (16, 64, 64, 320, 320), (2, 256, 256, 64, 128), (3, 128, 128, 64, 640),
                                            (40, 32, 32, 128, 320)])
def test_conv3x3_tall_tiles(B, H, W, Cin, Cout):
    from mvedit_b200 import tc_ops
    g = torch.Generator(device='cuda').manual_seed(B + H + Cin + Cout)
    x = torch.randn(B, H, W, Cin, device='cuda', generator=g).bfloat16()
    w = (torch.randn(Cout, 3, 3, Cin, device='cuda', generator=g) / math.sqrt(9 * Cin)).bfloat16()
    bias = torch.randn(Cout, device='cuda', generator=g)
    rb = torch.randn(B, Cout, device='cuda', generator=g)
    res = torch.randn(B, H, W, Cout, device='cuda', generator=g).bfloat16()
    out = tc_ops.conv3x3(x, w, bias=bias, row_bias=rb, residual=res)
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, padding=1)
    ref = (ref + rb[:, :, None, None]).permute(0, 2, 3, 1) + res.float()
    _check(out, ref, 9 * Cin)


@pytest.mark.parametrize('cin,cout,stride,nchw', [(3, 16, 1, True), (16, 16, 1, False), (16, 32, 2, False), (32, 32, 1, False), (32, 96, 2, False),
                                                   (3, 16, 1, 'bf16')])
def test_conv3x3_direct_small_channels(cin, cout, stride, nchw):
    """mve_conv3x3_direct_bf16 (ControlNet hint front, CUDA cores, true channel counts) vs F.conv2d + SiLU."""
    import torch.nn.functional as F
    from mvedit_b200 import tc_ops as T
    g = torch.Generator(device='cuda').manual_seed(cin * 100 + cout)
    B, H, W = 3, 40, 56
    w = torch.randn(cout, cin, 3, 3, device='cuda', generator=g) / math.sqrt(9 * cin)
    bias = torch.randn(cout, device='cuda', generator=g)
    x = torch.randn(B, cin, H, W, device='cuda', generator=g)
    if nchw is True:
        xin, xr = x, x
    elif nchw == 'bf16':
        xin = x.bfloat16(); xr = xin.float()
    else:
        xin = x.permute(0, 2, 3, 1).contiguous().bfloat16(); xr = xin.float().permute(0, 3, 1, 2)
    out = T.conv3x3_direct(xin, T.pack_direct_weight(w), bias, cin, cout, stride, act='silu', nchw=bool(nchw), out_channels=(cout + 63) // 64 * 64)
    ref = F.silu(F.conv2d(xr, w, bias, stride=stride, padding=1)).permute(0, 2, 3, 1)
    assert out.shape == (B, H // stride, W // stride, (cout + 63) // 64 * 64)
    _check(out[..., :cout], ref, 9 * cin)
    assert float(out[..., cout:].abs().max()) == 0.0 if out.shape[-1] > cout else True
