"""LPIPS patch loss on the kernels (mvedit_b200.lpips) against the fp32 restatement in oracle/lpips_oracle.py (lpips==0.1.4 is not
installed: the oracle follows the package's published algorithm and is unpinned).

Tolerances.  The product runs VGG16 in bf16 like the reference (lpips_loss.py:30: ``lpips_dtype = torch.bfloat16``), the oracle in
fp32: 13 layers of bf16 rounding give ~1 % on the distance and a few % (relative L2) on the gradient; building blocks that are exact
in bf16 (pooling, ReLU gate) are compared exactly, the convolution epilogues against torch on bf16-rounded inputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp(min=1e-20)).item()


@pytest.mark.parametrize('B,H,W,Cin,Cout', [(2, 16, 16, 64, 128), (2, 8, 8, 512, 512), (1, 16, 16, 256, 512)])
def test_conv_relu_and_relu_gate_epilogues(B, H, W, Cin, Cout):
    """The second and third shapes take the split-K path (few tiles, K = 9 * Cin >= 1024: fp32 partials in a workspace + finishing
    kernel); calling twice checks that the workspace is left zero."""
    from mvedit_b200 import tc_ops as T
    g = torch.Generator(device='cuda').manual_seed(0)
    x = torch.randn(B, H, W, Cin, device='cuda', generator=g).to(torch.bfloat16)
    w = (torch.randn(Cout, 3, 3, Cin, device='cuda', generator=g) * (2 / (9 * Cin)) ** 0.5).to(torch.bfloat16)
    b = torch.randn(Cout, device='cuda', generator=g) * 0.1
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), b, padding=1).permute(0, 2, 3, 1)
    y = T.conv3x3(x, w, bias=b, act='relu', split_k=True)
    assert (y.float() - T.conv3x3(x, w, bias=b, act='relu', split_k=True).float()).abs().max() < 1e-2
    assert torch.equal(T.conv3x3(x, w, bias=b, act='relu'), T.conv3x3(x, w, bias=b, act='relu'))          # without the opt-in: deterministic
    torch.testing.assert_close(y.float(), ref.clamp(min=0), rtol=2e-2, atol=2e-2)
    assert float(y.min()) >= 0 and float((y == 0).float().mean()) > 0.2
    gate = torch.randn(B, H, W, Cout, device='cuda', generator=g).to(torch.bfloat16)
    z = T.conv3x3(x, w, act='relu_gate', residual=gate, alpha=0.5, split_k=True)
    ref2 = torch.where(gate.float() > 0, 0.5 * (ref - b), torch.zeros_like(ref))
    torch.testing.assert_close(z.float(), ref2, rtol=2e-2, atol=2e-2)
    assert float(z[gate <= 0].abs().max()) == 0.0


@pytest.mark.parametrize('P,H,W,Cin,Cout', [(1, 8, 8, 512, 512), (2, 32, 32, 128, 64), (1, 128, 128, 64, 64)])
def test_input_gradient_convolution(P, H, W, Cin, Cout):
    """conv3x3(g, wT) with wT[i,ky,kx,o] = w[o,i,2-ky,2-kx] is the input gradient of the stride-1 pad-1 convolution."""
    from mvedit_b200 import tc_ops as T
    g = torch.Generator(device='cuda').manual_seed(1)
    w = torch.randn(Cout, Cin, 3, 3, device='cuda', generator=g) * (2 / (9 * Cin)) ** 0.5
    w = w.to(torch.bfloat16).float()
    gy = torch.randn(P, H, W, Cout, device='cuda', generator=g).to(torch.bfloat16)
    x = torch.zeros(P, Cin, H, W, device='cuda', requires_grad=True)
    F.conv2d(x, w, padding=1).backward(gy.float().permute(0, 3, 1, 2))
    wT = w.flip(2, 3).permute(1, 2, 3, 0).contiguous().to(torch.bfloat16)
    out = T.conv3x3(gy, wT, split_k=True)
    assert rel(out, x.grad.permute(0, 2, 3, 1)) < 1e-2


def test_maxpool_forward_backward_exact():
    from mvedit_b200._lib import call, ptr, stream, c_u32
    g = torch.Generator(device='cuda').manual_seed(2)
    B, H, W, C = 3, 8, 16, 64
    x = torch.randn(B, H, W, C, device='cuda', generator=g).clamp(min=0).to(torch.bfloat16)          # a ReLU output (ties at 0)
    y = torch.empty(B, H // 2, W // 2, C, dtype=torch.bfloat16, device='cuda')
    call('mve_maxpool2x2_bf16', ptr(x), c_u32(B), c_u32(H), c_u32(W), c_u32(C), ptr(y), stream())
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    yr = F.max_pool2d(xr, 2, 2)
    assert torch.equal(y.float(), yr.detach().permute(0, 2, 3, 1))
    gy = torch.randn(B, H // 2, W // 2, C, device='cuda', generator=g).to(torch.bfloat16)
    g0 = torch.randn(B, H, W, C, device='cuda', generator=g).to(torch.bfloat16)
    gx = g0.clone()
    call('mve_maxpool2x2_relu_backward_bf16', ptr(x), ptr(gy), c_u32(B), c_u32(H), c_u32(W), c_u32(C), ptr(gx), stream())
    yr.backward(gy.float().permute(0, 3, 1, 2))
    want = torch.where(x.float() > 0, (g0.float() + xr.grad.permute(0, 2, 3, 1)).to(torch.bfloat16).float(), torch.zeros_like(g0.float()))
    assert torch.equal(gx.float(), want)


def _patches(P, S, seed):
    """target: smooth random images; pred: the target plus structured and white perturbations (what a half-fitted render looks like)."""
    g = torch.Generator(device='cuda').manual_seed(seed)
    low = torch.rand(P, 3, S // 8, S // 8, device='cuda', generator=g)
    tgt = F.interpolate(low, size=(S, S), mode='bicubic', align_corners=False).clamp(0, 1)
    low2 = torch.rand(P, 3, S // 4, S // 4, device='cuda', generator=g) - 0.5
    pred = (tgt + 0.25 * F.interpolate(low2, size=(S, S), mode='bilinear') + 0.05 * torch.randn(P, 3, S, S, device='cuda', generator=g)).clamp(0, 1)
    return pred.permute(0, 2, 3, 1).contiguous(), tgt.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize('P,S', [(1, 128), (2, 64), (3, 32)])
def test_lpips_value_and_gradient_vs_oracle(P, S):
    from oracle import lpips_oracle as lo
    from mvedit_b200.lpips import LPIPS, LPIPSLoss
    sd = lo.random_lpips_state_dict(0)
    sdc = {k: v.cuda() for k, v in sd.items()}
    pred, tgt = _patches(P, S, 7 + S)
    w = torch.linspace(0.7, 1.4, P, device='cuda')
    x = pred.permute(0, 3, 1, 2).clone().requires_grad_(True)
    d_ref = lo.lpips(sdc, x * 2 - 1, tgt.permute(0, 3, 1, 2) * 2 - 1)
    loss_ref = lo.lpips_loss(sdc, x, tgt.permute(0, 3, 1, 2), w, loss_weight=1.2) * 0.9
    loss_ref.backward()
    g_ref = x.grad.permute(0, 2, 3, 1)
    m = LPIPSLoss(sd, loss_weight=1.2)
    loss, g, d = m.loss_and_grad(pred, tgt, w, torch.tensor(0.9, device='cuda'))
    assert float((d / d_ref.detach() - 1).abs().max()) < 3e-2, (d, d_ref)
    assert abs(float(loss) / float(loss_ref) - 1) < 3e-2
    r = rel(g, g_ref)
    assert r < 8e-2, r
    cos = F.cosine_similarity(g.flatten(), g_ref.flatten(), dim=0).item()
    assert cos > 0.995, cos
    # the module-style call (NCHW in [0,1], as the reference calls nerf.patch_loss) and the bare distance agree with loss_and_grad
    # (two separate evaluations: the split-K convolutions of the deep levels sum their fp32 partials in atomic order)
    torch.testing.assert_close(m(pred.permute(0, 3, 1, 2), tgt.permute(0, 3, 1, 2), weight=w) * 0.9, loss, rtol=1e-3, atol=1e-7)
    torch.testing.assert_close(LPIPS(sd)(pred, tgt), d, rtol=1e-3, atol=1e-7)
    assert float(LPIPS(sd)(tgt, tgt).abs().max()) < 1e-4 * float(d.max())        # ~1e-8 vs ~2e-3: the two halves of the batch sum in different orders


def test_gradient_descends_the_oracle_distance():
    """A step along -g (from the kernels) must lower the ORACLE's distance: the gradient is usable, not just close in norm."""
    from oracle import lpips_oracle as lo
    from mvedit_b200.lpips import LPIPS
    sd = lo.random_lpips_state_dict(3)
    sdc = {k: v.cuda() for k, v in sd.items()}
    pred, tgt = _patches(1, 64, 11)
    _, g, d0 = LPIPS(sd).loss_and_grad(pred, tgt)
    step = 0.02 / float(g.abs().max())
    d = lambda p: float(lo.lpips(sdc, p.permute(0, 3, 1, 2) * 2 - 1, tgt.permute(0, 3, 1, 2) * 2 - 1))
    assert d((pred - step * g).clamp(0, 1)) < 0.97 * d(pred)


@pytest.mark.parametrize('shaded', [False, True])
def test_objective_with_lpips_term_matches_torch_chain(shaded):
    """mve_nerf_patch_out_rgb -> LPIPS -> mve_nerf_patch_loss(g_out_extra): d/d(image, alpha, depth) of L1 + alpha + TV + entropy + LPIPS
    against autograd of the oracle objective with the oracle's LPIPS term (mvedit_3d_pipeline.py:541-617)."""
    from tests.test_gpu_nerf_loss import torch_chain
    from oracle import lpips_oracle as lo
    from mvedit_b200.lpips import LPIPSLoss
    from mvedit_b200._lib import call, ptr, stream, c_u32, c_int, c_f32
    P, ps = 2, 32
    g = torch.Generator(device='cuda').manual_seed(5 + shaded)
    N = P * ps * ps
    R = lambda *s: torch.rand(*s, device='cuda', generator=g)
    pred, tgt = _patches(P, ps, 3)
    alpha = (0.3 + R(N) * 0.8).clamp(0, 1)
    image = pred.reshape(N, 3) * alpha[:, None]
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, ps, device='cuda'), torch.linspace(-1, 1, ps, device='cuda'), indexing='ij')
    depth = alpha * (0.3 + 0.05 * (xx * xx + yy * yy)).flatten().repeat(P)               # a smooth bowl: well-defined normals
    tgt_mask = R(P, ps, ps, 1)
    xs = (torch.arange(ps, device='cuda') + 0.5 - ps / 2) / (2.0 * ps)
    dirs = torch.stack([xs[None, :].expand(ps, ps), xs[:, None].expand(ps, ps), torch.ones(ps, ps, device='cuda')], -1)[None].repeat(P, 1, 1, 1).contiguous()
    patch_w, lights = 0.5 + R(P), F.normalize(torch.tensor([[0.2, 0.3, 1.0]] * P, device='cuda'), dim=-1)
    prw = 0.8
    sd = lo.random_lpips_state_dict(0)
    sdc = {k: v.cuda() for k, v in sd.items()}
    # ---- oracle: the torch chain + LPIPS on out_rgbs, one autograd graph
    inp = [t.clone().requires_grad_(True) for t in (image, alpha, depth)]
    ref = torch_chain(*inp, tgt, tgt_mask, dirs, patch_w, lights, ps, shaded, 0.2, 1.0, 0.015, 1.2, 1.0, 1.3, 0.02)
    from oracle.nerf_oracle import depth_to_normal
    od = inp[2].reshape(P, ps, ps) * torch.linalg.norm(dirs, dim=-1) / inp[1].reshape(P, ps, ps).clamp(min=1e-6)
    out = inp[0].reshape(P, ps, ps, 3)
    if shaded:
        n_fg = depth_to_normal(od, dirs)
        ncv = torch.cat([n_fg[..., :1] * 2 - 1, -n_fg[..., 1:3] * 2 + 1], dim=-1)
        out = out * ((lights[:, None, None, None, :] @ ncv[..., :, None]).clamp(min=0) * 0.8 + 0.2).squeeze(-1)
    out = out + 1.0 * (1 - inp[1].reshape(P, ps, ps, 1))
    l_lp = lo.lpips_loss(sdc, out.permute(0, 3, 1, 2), tgt.permute(0, 3, 1, 2), patch_w, loss_weight=1.2) * prw
    (ref[0] + l_lp).backward()
    # ---- product
    f32 = dict(dtype=torch.float32, device='cuda')
    scratch, out_rgb = torch.empty(N * 10, **f32), torch.empty(N, 3, **f32)
    call('mve_nerf_patch_out_rgb', ptr(image), ptr(alpha), ptr(depth), ptr(dirs), ptr(lights), c_u32(P), c_u32(ps), c_int(int(shaded)),
         c_f32(0.2), c_f32(1.0), ptr(scratch), ptr(out_rgb), None, c_u32(0), stream())
    torch.testing.assert_close(out_rgb.view(P, ps, ps, 3), out.detach(), rtol=1e-4, atol=1e-5)
    lp, g_extra, _ = LPIPSLoss(sd, loss_weight=1.2).loss_and_grad(out_rgb.view(P, ps, ps, 3), tgt, patch_w, prw)
    assert abs(float(lp) / float(l_lp) - 1) < 3e-2
    sc = [torch.tensor(v, device='cuda') for v in (1.0, 1.3, 0.02)]
    g_i, g_a, g_d, loss5 = torch.empty(N, 3, **f32), torch.empty(N, **f32), torch.empty(N, **f32), torch.empty(5, **f32)
    call('mve_nerf_patch_loss', ptr(image), ptr(alpha), ptr(depth), ptr(tgt), ptr(tgt_mask), ptr(dirs), ptr(patch_w), ptr(lights), c_u32(P),
         c_u32(ps), c_int(int(shaded)), c_f32(0.2), c_f32(1.0), c_f32(0.015), c_f32(1.2), ptr(sc[0]), ptr(sc[1]), ptr(sc[2]), ptr(scratch),
         ptr(g_i), ptr(g_a), ptr(g_d), ptr(loss5), ptr(g_extra), None, c_u32(0), stream())
    torch.testing.assert_close(loss5, ref.detach(), rtol=2e-4, atol=1e-6)
    for a, b, name in zip((g_i, g_a, g_d), inp, ('image', 'alpha', 'depth')):
        r = rel(a.view_as(b.grad), b.grad)
        assert r < 8e-2, (name, r)
