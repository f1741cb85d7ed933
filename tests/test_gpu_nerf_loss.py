"""Fused nerf_optim objective (mve_nerf_patch_loss + entropy folded into composite backward) vs the eager torch chain that restates
mvedit_3d_pipeline.py:541-603 with the oracle's loss modules (oracle/nerf_oracle.py; TVLoss / depth_to_normal pinned against the
reference's own functions by tests/test_reference_pins.py)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def torch_chain(image, alpha, depth, tgt_rgb, tgt_mask, dirs, patch_w, lights, ps, shaded, ambient, bg, bg_width, plw, alpha_mul, nreg, went,
                tone=None):
    from oracle.nerf_oracle import depth_to_normal, TVLoss, L1LossMod
    import torch.nn.functional as F
    P = alpha.numel() // (ps * ps)
    out_rgbs = image.reshape(P, ps, ps, 3)
    out_alphas = alpha.reshape(P, ps, ps, 1)
    out_depth = depth.reshape(P, ps, ps) * torch.linalg.norm(dirs, dim=-1)
    out_depth_fg = out_depth / out_alphas.reshape(P, ps, ps).clamp(min=1e-6)
    n_fg = depth_to_normal(out_depth_fg, dirs)
    fgw = -F.max_pool2d(-out_alphas.detach().squeeze(-1).unsqueeze(1), 3, stride=1, padding=1).squeeze(1).unsqueeze(-1)
    if shaded:
        ncv = torch.cat([n_fg[..., :1] * 2 - 1, -n_fg[..., 1:3] * 2 + 1], dim=-1)
        sh = ((lights[:, None, None, None, :] @ ncv[..., :, None]).clamp(min=0) * (1 - ambient) + ambient).squeeze(-1)
        if tone is None:
            out_rgbs = out_rgbs * sh + bg * (1 - out_alphas)
        else:        # mvedit_3d_pipeline.py:564-570
            out_rgbs = tone.lut(tone.inverse_lut(out_rgbs / out_alphas.clamp(min=1e-6)) + sh.clamp(min=1e-6).log2()) * out_alphas \
                + bg * (1 - out_alphas)
    else:
        out_rgbs = out_rgbs + bg * (1 - out_alphas)
    w = patch_w[:, None, None, None].expand(-1, ps, ps, 1)
    l1 = L1LossMod(plw)
    l_rgb = l1(out_rgbs, tgt_rgb, weight=w) * 4.5
    l_a = l1(out_alphas, tgt_mask, weight=w) * alpha_mul
    l_tv = TVLoss(power=1.5)(n_fg.permute(0, 3, 1, 2), None, weight=fgw.permute(0, 3, 1, 2)) * nreg
    bgw = 1 - alpha.flatten()
    l_e = -torch.sum(bgw * (torch.log(bgw.clamp(min=1e-6)) - math.log(bg_width))) * (went / alpha.numel())
    return torch.stack([l_rgb + l_a + l_tv + l_e, l_rgb, l_a, l_tv, l_e])


@pytest.mark.parametrize('shaded', [False, True, 'tone'])
@pytest.mark.parametrize('P,ps', [(1, 32), (3, 16)])
def test_fused_patch_loss_matches_torch_chain(shaded, P, ps):
    """shaded == 'tone': Lambert shading applied in tone-mapped space (the runner always passes a Tonemapping, adapter3d.py:88,780)."""
    from mvedit_b200.nerf import patch_loss
    from mvedit_b200.tonemapping import Tonemapping
    from oracle.nerf_oracle import Tonemapping as OracleTonemapping
    tone_o, tone_p = (OracleTonemapping().cuda(), Tonemapping()) if shaded == 'tone' else (None, None)
    shaded = bool(shaded)
    g = torch.Generator(device='cuda').manual_seed(P * ps + int(shaded))
    N = P * ps * ps
    R = lambda *s: torch.rand(*s, device='cuda', generator=g)
    alpha = (R(N) * 1.1).clamp(0, 1)
    alpha[::7] = 0.0          # background rays (clamps active)
    image = R(N, 3) * alpha[:, None]
    depth = alpha * (0.2 + 0.2 * R(N))
    tgt_rgb, tgt_mask = R(P, ps, ps, 3), R(P, ps, ps, 1)
    xs = (torch.arange(ps, device='cuda') + 0.5 - ps / 2) / (2.0 * ps)
    dirs = torch.stack([xs[None, :].expand(ps, ps), xs[:, None].expand(ps, ps), torch.ones(ps, ps, device='cuda')], -1)[None].repeat(P, 1, 1, 1).contiguous()
    patch_w, lights = 0.5 + R(P), torch.nn.functional.normalize(torch.randn(P, 3, device='cuda', generator=g), dim=-1)
    sc = [torch.tensor(v, device='cuda') for v in (5.0, 1.3, 0.02)]
    inp = [t.clone().requires_grad_(True) for t in (image, alpha, depth)]
    ref = torch_chain(*inp, tgt_rgb, tgt_mask, dirs, patch_w, lights, ps, shaded, 0.2, 1.0, 0.015, 1.2, 5.0, 1.3, 0.02, tone=tone_o)
    ref[0].backward()
    out, *grads = patch_loss(image, alpha, depth, tgt_rgb, tgt_mask, dirs, patch_w, lights, ps, shaded, 0.2, 1.0, 0.015, 1.2, *sc, tonemapping=tone_p)
    torch.testing.assert_close(out, ref.detach(), rtol=2e-4, atol=1e-6)
    for a, b, name in zip(grads, inp, ('image', 'alpha', 'depth')):
        err = (a.view_as(b.grad) - b.grad).abs().max().item()
        assert err <= 2e-3 * b.grad.abs().max().item() + 1e-7, (name, err, b.grad.abs().max().item())


def test_fused_sample_entropy_gradient():
    """composite backward with the fused entropy term == autograd of -c * sum w (log w - log dt) through composite_rays_train."""
    from mvedit_b200 import raymarching as rm
    g = torch.Generator(device='cuda').manual_seed(0)
    counts = torch.tensor([0, 5, 40, 33, 1, 64], dtype=torch.int32)
    rays = torch.stack([torch.cat([torch.zeros(1, dtype=torch.int32), counts.cumsum(0)[:-1].int()]), counts], -1).cuda()
    M, N = int(counts.sum()), counts.numel()
    sig = torch.exp(torch.randn(M, device='cuda', generator=g))
    rgb = torch.rand(M, 3, device='cuda', generator=g)
    ts = torch.stack([2 + torch.rand(M, device='cuda', generator=g).sort().values, 0.003 + 0.01 * torch.rand(M, device='cuda', generator=g)], -1)
    went = torch.tensor(0.37, device='cuda')
    s1, c1 = sig.clone().requires_grad_(True), rgb.clone().requires_grad_(True)
    w, ws, d, img = rm.composite_rays_train(s1, c1, ts, rays)
    ent = -(w * (torch.log(w.clamp(min=1e-6)) - torch.log(ts[:, 1].clamp(min=1e-6)))).sum() * (went / N)
    (img.sum() * 0.3 + ws.sum() + ent).backward()
    s2, c2 = sig.clone().requires_grad_(True), rgb.clone().requires_grad_(True)
    w2, ws2, d2, img2 = rm.composite_rays_train(s2, c2, ts, rays, 1e-4, False, None, (went, 1.0 / N))
    (img2.sum() * 0.3 + ws2.sum()).backward()
    # NB the reference's backward applies gw_i to the later samples of the ray (raymarching.cu:676); both paths go through
    # the same kernel formula, so they must agree exactly up to float noise
    torch.testing.assert_close(s2.grad, s1.grad, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(c2.grad, c1.grad, rtol=1e-5, atol=1e-7)
