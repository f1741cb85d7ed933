"""3DGS rasteriser (mvedit_b200.gs_renderer: torch projection + CUDA tile binning / blend fwd+bwd) vs the dense oracle
(oracle/gs_oracle.py).  No reference code exists for this row: parity is against the restated public algorithm.
Tolerances: forward 2e-4 absolute on colour / alpha (exp via __expf, different summation order), gradients 2e-3 of the tensor's max."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cloud(P, seed):
    g = torch.Generator(device='cuda').manual_seed(seed)
    R = lambda *s: torch.rand(*s, device='cuda', generator=g)
    means = torch.randn(P, 3, device='cuda', generator=g) * 0.3
    scales = torch.exp(R(P, 3) * 2.0 - 4.5)
    quats = torch.randn(P, 4, device='cuda', generator=g)
    opac = torch.sigmoid(torch.randn(P, device='cuda', generator=g) + 1.0)
    cols = R(P, 3)
    view = torch.eye(4, device='cuda')
    view[2, 3] = 2.2
    return means, scales, quats, opac, cols, view


@pytest.mark.parametrize('P,H,W', [(300, 48, 40), (1200, 64, 80)])
def test_gs_forward_backward_match_oracle(P, H, W):
    from oracle import gs_oracle as go
    from mvedit_b200.gs_renderer import GaussianRasterizer, GaussianRasterizationSettings
    means, scales, quats, opac, cols, view = _cloud(P, P)
    means[::23, 2] = -3.5                                        # culled by the near plane
    K = (1.4 * W, 1.35 * W, W / 2 + 0.7, H / 2 - 1.2)
    bg = (0.1, 0.5, 0.9)
    ins = [t.clone().requires_grad_(True) for t in (means, scales, quats, opac, cols)]
    ref_c, ref_d, ref_a = go.render(*ins[:3], ins[3], ins[4], view, K, H, W, torch.tensor(bg, device='cuda'))
    g = torch.Generator(device='cuda').manual_seed(1)
    wc, wd, wa = torch.randn(H, W, 3, device='cuda', generator=g), torch.randn(H, W, device='cuda', generator=g), torch.randn(H, W, device='cuda', generator=g)
    ((ref_c * wc).sum() + (ref_d * wd).sum() + (ref_a * wa).sum()).backward()
    ins2 = [t.clone().requires_grad_(True) for t in (means, scales, quats, opac, cols)]
    rast = GaussianRasterizer(GaussianRasterizationSettings(image_height=H, image_width=W, viewmatrix=view, intrinsics=K, bg=bg))
    color, depth, alpha = rast(ins2[0], ins2[3], ins2[4], ins2[1], ins2[2])
    assert color.shape == (3, H, W) and depth.shape == (1, H, W) and alpha.shape == (1, H, W)
    assert float(ref_a.max()) > 0.5
    assert (color.permute(1, 2, 0) - ref_c).abs().max().item() < 2e-4
    assert (alpha[0] - ref_a).abs().max().item() < 2e-4 and (depth[0] - ref_d).abs().max().item() < 1e-3
    ((color.permute(1, 2, 0) * wc).sum() + (depth[0] * wd).sum() + (alpha[0] * wa).sum()).backward()
    for a, b, name in zip(ins2, ins, ('means3D', 'scales', 'rotations', 'opacities', 'colors')):
        err, scale = (a.grad - b.grad).abs().max().item(), b.grad.abs().max().item()
        assert err <= 2e-3 * scale + 1e-6, (name, err, scale)


def test_gs_empty_and_offscreen():
    from mvedit_b200.gs_renderer import GaussianRasterizer, GaussianRasterizationSettings
    means, scales, quats, opac, cols, view = _cloud(64, 3)
    means[:, 0] += 100.0                                          # everything off screen
    rast = GaussianRasterizer(GaussianRasterizationSettings(image_height=32, image_width=32, viewmatrix=view, intrinsics=(40.0, 40.0, 16.0, 16.0),
                                                            bg=(0.2, 0.3, 0.4)))
    color, depth, alpha = rast(means.requires_grad_(True), opac, cols, scales, quats)
    assert float(alpha.abs().max()) == 0.0 and torch.allclose(color[:, 0, 0], torch.tensor([0.2, 0.3, 0.4], device='cuda'))
    color.sum().backward()
    assert float(means.grad.abs().max()) == 0.0
