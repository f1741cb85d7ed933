"""The fused nerf_optim objective (``mvedit_b200/csrc/nerf_loss.cu``) on the CPU: the kernel source compiled UNCHANGED as C++ through
tests/host_shim (launches -> serial loops) and driven through the product's own Python mirror (``mvedit_b200.nerf.patch_loss`` /
``patch_out_normal``, routed to that library), against the eager torch chain that restates mvedit_3d_pipeline.py:541-626 with the oracle's
loss modules -- incl. the optional terms of image-to-3D runs: target normals inside the TV term, the L1 depth term and the gradient of
the normal patch term w.r.t. the alpha-composited normals.  The same checks run on the GPU in tests/test_gpu_nerf_loss.py."""
import math

import pytest
import torch
import torch.nn.functional as F

from tests import host_harness
from mvedit_b200 import nerf as pnerf

NORMAL_BG = (0.5, 0.5, 1.0)


def torch_chain(image, alpha, depth, tgt_rgb, tgt_mask, dirs, patch_w, lights, ps, shaded, ambient, bg, bg_width, plw, alpha_mul, nreg, went,
                tone=None, tgt_normal=None, tgt_depth=None, w_depth=0.0, normal_probe=None):
    """-> ([total, rgb, alpha, normal_reg, entropy(, depth)], composited normals)."""
    from oracle.nerf_oracle import depth_to_normal, TVLoss, L1LossMod
    P = alpha.numel() // (ps * ps)
    out_rgbs = image.reshape(P, ps, ps, 3)
    out_alphas = alpha.reshape(P, ps, ps, 1)
    out_depth = depth.reshape(P, ps, ps) * torch.linalg.norm(dirs, dim=-1)
    out_depth_fg = out_depth / out_alphas.reshape(P, ps, ps).clamp(min=1e-6)
    n_fg = depth_to_normal(out_depth_fg, dirs)
    out_normals = n_fg * out_alphas + n_fg.new_tensor(NORMAL_BG) * (1 - out_alphas)
    fgw = -F.max_pool2d(-out_alphas.detach().squeeze(-1).unsqueeze(1), 3, stride=1, padding=1).squeeze(1).unsqueeze(-1)
    if shaded:
        ncv = torch.cat([n_fg[..., :1] * 2 - 1, -n_fg[..., 1:3] * 2 + 1], dim=-1)
        sh = ((lights[:, None, None, None, :] @ ncv[..., :, None]).clamp(min=0) * (1 - ambient) + ambient).squeeze(-1)
        if tone is None:
            out_rgbs = out_rgbs * sh + bg * (1 - out_alphas)
        else:
            out_rgbs = tone.lut(tone.inverse_lut(out_rgbs / out_alphas.clamp(min=1e-6)) + sh.clamp(min=1e-6).log2()) * out_alphas \
                + bg * (1 - out_alphas)
    else:
        out_rgbs = out_rgbs + bg * (1 - out_alphas)
    w = patch_w[:, None, None, None].expand(-1, ps, ps, 1)
    l1 = L1LossMod(plw)
    l_rgb = l1(out_rgbs, tgt_rgb, weight=w) * 4.5
    l_a = l1(out_alphas, tgt_mask, weight=w) * alpha_mul
    l_tv = TVLoss(power=1.5)(n_fg.permute(0, 3, 1, 2), None if tgt_normal is None else tgt_normal.permute(0, 3, 1, 2),
                             weight=fgw.permute(0, 3, 1, 2)) * nreg
    bgw = 1 - alpha.flatten()
    l_e = -torch.sum(bgw * (torch.log(bgw.clamp(min=1e-6)) - math.log(bg_width))) * (went / alpha.numel())
    terms = [l_rgb + l_a + l_tv + l_e, l_rgb, l_a, l_tv, l_e]
    if tgt_depth is not None:
        l_d = l1(out_depth.reshape(tgt_depth.size()), tgt_depth, weight=w) * w_depth
        terms[0] = terms[0] + l_d
        terms.append(l_d)
    if normal_probe is not None:                 # stands for the patch term: any function of the composited normals
        terms[0] = terms[0] + (out_normals * normal_probe).sum()
    return torch.stack(terms), out_normals


def make_inputs(P, ps, seed):
    g = torch.Generator().manual_seed(seed)
    N = P * ps * ps
    R = lambda *s: torch.rand(*s, generator=g)
    alpha = (0.05 + R(N) * 1.05).clamp(0, 1).view(P, ps, ps)
    # background rays (alpha = 0: the clamps put them at 1e6 x direction) come in REGIONS, as in a render.  An isolated background ray
    # between foreground rays would make its own normal a cross product of four nearly parallel 1e6-long vectors -- fp32 noise of tens
    # of per cent in any implementation, nothing a parity test can hold
    if ps >= 8:
        alpha[:, ps // 4:ps // 2, ps // 4:ps // 2 + 3] = 0.0
        alpha[0, :2, :] = 0.0
    else:
        alpha[-1] = 0.0
    alpha = alpha.reshape(N).contiguous()
    image = R(N, 3) * alpha[:, None]
    depth = alpha * (0.2 + 0.2 * R(N))
    tgt_rgb, tgt_mask = R(P, ps, ps, 3), R(P, ps, ps, 1)
    xs = (torch.arange(ps) + 0.5 - ps / 2) / (2.0 * ps)
    dirs = torch.stack([xs[None, :].expand(ps, ps), xs[:, None].expand(ps, ps), torch.ones(ps, ps)], -1)[None].repeat(P, 1, 1, 1).contiguous()
    patch_w, lights = 0.5 + R(P), F.normalize(torch.randn(P, 3, generator=g), dim=-1)
    tgt_normal = F.normalize(torch.randn(P, ps, ps, 3, generator=g), dim=-1) / 2 + 0.5
    tgt_depth = 0.2 + 0.3 * R(P, ps, ps, 1)
    probe = torch.randn(P, ps, ps, 3, generator=g) * 1e-3
    return image, alpha, depth, tgt_rgb, tgt_mask, dirs, patch_w, lights, tgt_normal, tgt_depth, probe


def check(grads, inp, alpha, tol=2e-3):
    """Foreground rays to ``tol``; background rays on the rim of a region (their normal mixes 1e6-long and unit-long differences, good to
    ~1 % in fp32 whatever the evaluation order) to 5 %."""
    bgr = alpha.flatten() == 0
    for a, b, name in zip(grads, inp, ('image', 'alpha', 'depth')):
        err = (a.view_as(b.grad) - b.grad).abs().reshape(bgr.numel(), -1).max(dim=1).values
        top = b.grad.abs().max().item()
        assert err[~bgr].max().item() <= tol * top + 1e-7, (name, err[~bgr].max().item(), top)
        assert err[bgr].max().item() <= 0.05 * top + 1e-7, (name, err[bgr].max().item(), top)


@pytest.mark.parametrize('shaded', [False, True, 'tone'])
@pytest.mark.parametrize('P,ps', [(1, 32), (3, 16), (2, 2)])
def test_objective_kernels_match_the_torch_chain(shaded, P, ps):
    from mvedit_b200.tonemapping import Tonemapping
    from oracle.nerf_oracle import Tonemapping as OracleTonemapping
    tone_o, tone_p = (OracleTonemapping(), Tonemapping()) if shaded == 'tone' else (None, None)
    shaded = bool(shaded)
    image, alpha, depth, tgt_rgb, tgt_mask, dirs, patch_w, lights, *_ = make_inputs(P, ps, P * ps + int(shaded))
    sc = [torch.tensor(v) for v in (5.0, 1.3, 0.02)]
    inp = [t.clone().requires_grad_(True) for t in (image, alpha, depth)]
    ref, _ = torch_chain(*inp, tgt_rgb, tgt_mask, dirs, patch_w, lights, ps, shaded, 0.2, 1.0, 0.015, 1.2, 5.0, 1.3, 0.02, tone=tone_o)
    ref[0].backward()
    with host_harness.routed(pnerf, host_harness.shimmed('nerf_loss.cu')):
        out, *grads = pnerf.patch_loss(image, alpha, depth, tgt_rgb, tgt_mask, dirs, patch_w, lights, ps, shaded, 0.2, 1.0, 0.015, 1.2, *sc,
                                       tonemapping=tone_p)
    torch.testing.assert_close(out, ref.detach(), rtol=2e-4, atol=1e-6)
    check(grads, inp, alpha)


@pytest.mark.parametrize('which', ['normal', 'depth', 'patch_normal', 'all', 'all_tone'])
def test_optional_targets_match_the_torch_chain(which):
    from mvedit_b200.tonemapping import Tonemapping
    from oracle.nerf_oracle import Tonemapping as OracleTonemapping
    P, ps = 3, 16
    image, alpha, depth, tgt_rgb, tgt_mask, dirs, patch_w, lights, tgt_normal, tgt_depth, probe = make_inputs(P, ps, 11)
    tone_o, tone_p = (OracleTonemapping(), Tonemapping()) if which == 'all_tone' else (None, None)
    use_n, use_d, use_p = which in ('normal', 'all', 'all_tone'), which in ('depth', 'all', 'all_tone'), which in ('patch_normal', 'all', 'all_tone')
    sc = [torch.tensor(v) for v in (1.0, 1.3, 0.02)]
    inp = [t.clone().requires_grad_(True) for t in (image, alpha, depth)]
    ref, ref_normals = torch_chain(*inp, tgt_rgb, tgt_mask, dirs, patch_w, lights, ps, True, 0.2, 1.0, 0.015, 1.2, 1.0, 1.3, 0.02, tone=tone_o,
                                   tgt_normal=tgt_normal if use_n else None, tgt_depth=tgt_depth if use_d else None, w_depth=0.7,
                                   normal_probe=probe if use_p else None)
    ref[0].backward()
    with host_harness.routed(pnerf, host_harness.shimmed('nerf_loss.cu')):
        out, *grads = pnerf.patch_loss(image, alpha, depth, tgt_rgb, tgt_mask, dirs, patch_w, lights, ps, True, 0.2, 1.0, 0.015, 1.2, *sc,
                                       tonemapping=tone_p, tgt_normal=tgt_normal.reshape(-1, 3) if use_n else None,
                                       g_normal_extra=probe.reshape(-1, 3) if use_p else None, normal_bg=NORMAL_BG,
                                       tgt_depth=tgt_depth.reshape(-1) if use_d else None, w_depth=torch.tensor(0.7) if use_d else None)
        normals = pnerf.patch_out_normal(alpha, depth, dirs, ps, NORMAL_BG)
    torch.testing.assert_close(normals.view_as(ref_normals), ref_normals.detach(), rtol=1e-4, atol=2e-5)
    expect = ref.detach().clone()
    if use_p:
        expect[0] -= (ref_normals.detach() * probe).sum()                      # the kernel reports its own terms; the patch term is the caller's
    torch.testing.assert_close(out, expect, rtol=2e-4, atol=1e-6)
    assert out.numel() == (6 if use_d else 5)
    check(grads, inp, alpha)


def test_converted_inputs_stay_alive_until_the_call():
    """Non-contiguous / non-fp32 inputs are converted to temporaries: the mirror must keep them referenced (a freed temporary could be
    reused by the next conversion before the launch reads it)."""
    P, ps = 2, 8
    image, alpha, depth, tgt_rgb, tgt_mask, dirs, patch_w, lights, *_ = make_inputs(P, ps, 3)
    sc = [torch.tensor(v) for v in (1.0, 1.3, 0.02)]
    with host_harness.routed(pnerf, host_harness.shimmed('nerf_loss.cu')):
        a = pnerf.patch_loss(image, alpha, depth, tgt_rgb, tgt_mask, dirs, patch_w, lights, ps, True, 0.2, 1.0, 0.015, 1.2, *sc)
        b = pnerf.patch_loss(image.double(), alpha.double(), depth.double(), tgt_rgb.double(), tgt_mask.double(), dirs.double(), patch_w.double(),
                             lights.double(), ps, True, 0.2, 1.0, 0.015, 1.2, *sc)
    for x, y in zip(a, b):
        torch.testing.assert_close(x, y, rtol=0, atol=0)
