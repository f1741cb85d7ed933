"""CPU checks of the 3DGS host side: the differentiable per-Gaussian projection of mvedit_b200.gs_renderer against the oracle's
restatement (oracle/gs_oracle.py: same public algorithm, written independently as a dense reference), and properties of the dense
oracle itself (a single isotropic Gaussian renders the closed-form alpha; depth ordering matters)."""
import math

import pytest
import torch


def _cloud(P, seed, dev='cpu'):
    g = torch.Generator().manual_seed(seed)
    means = torch.randn(P, 3, generator=g) * 0.3
    scales = torch.exp(torch.empty(P, 3).uniform_(-4.0, -2.0, generator=g))
    quats = torch.randn(P, 4, generator=g)
    opac = torch.sigmoid(torch.randn(P, generator=g))
    cols = torch.rand(P, 3, generator=g)
    view = torch.eye(4)
    view[2, 3] = 2.5                                 # camera 2.5 in front of the cloud, looking down +z
    return [t.to(dev) for t in (means, scales, quats, opac, cols, view)]


def test_projection_matches_oracle():
    from oracle import gs_oracle as go
    from mvedit_b200.gs_renderer import project_gaussians
    means, scales, quats, opac, cols, view = _cloud(500, 0)
    means[::17, 2] = -3.0                            # some Gaussians behind the near plane
    K, H, W = (120.0, 118.0, 33.0, 21.5), 40, 72
    pre = go.preprocess(means, scales, quats, view, K, H, W)
    xy, conic, depth, rect = project_gaussians(means, scales, quats, view, K, H, W)
    v = pre['valid']
    torch.testing.assert_close(xy[v], pre['xy'][v], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(conic[v], pre['conic'][v], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(depth[v], pre['depth'][v])
    assert torch.equal(rect, pre['rect']) and int((~v).sum()) >= 30
    assert (((rect[:, 2] - rect[:, 0]) * (rect[:, 3] - rect[:, 1]) > 0) == v).all()


def test_oracle_single_gaussian_closed_form():
    from oracle import gs_oracle as go
    s = 0.05
    means = torch.zeros(1, 3)
    scales, quats = torch.full((1, 3), s), torch.tensor([[1.0, 0, 0, 0]])
    opac, cols = torch.tensor([0.8]), torch.tensor([[0.2, 0.5, 0.9]])
    view = torch.eye(4); view[2, 3] = 2.0
    f, H, W = 100.0, 33, 33
    K = (f, f, W / 2, H / 2)
    color, depth, alpha = go.render(means, scales, quats, opac, cols, view, K, H, W, torch.tensor([1.0, 1.0, 1.0]))
    sig2 = (f * s / 2.0) ** 2 + 0.3                  # projected variance + low-pass
    c = H // 2                                       # pixel index 16 has its centre at the principal point (16.5 - 0.5)
    assert alpha[c, c].item() == pytest.approx(0.8, rel=1e-5)
    assert alpha[c, c + 3].item() == pytest.approx(0.8 * math.exp(-0.5 * 9 / sig2), rel=1e-4)
    assert depth[c, c].item() == pytest.approx(0.8 * 2.0, rel=1e-5)
    torch.testing.assert_close(color[c, c], 0.8 * cols[0] + 0.2 * torch.ones(3))
    # far outside the 3-sigma tile rect nothing is drawn
    assert alpha[0, 0].item() == 0.0


def test_oracle_depth_order_and_early_stop():
    from oracle import gs_oracle as go
    means = torch.tensor([[0.0, 0, 0.5], [0.0, 0, 0.0]])          # second one is nearer to the camera at z = +2
    scales, quats = torch.full((2, 3), 0.2), torch.tensor([[1.0, 0, 0, 0]] * 2)
    view = torch.eye(4); view[2, 3] = 2.0
    cols = torch.tensor([[1.0, 0, 0], [0, 1.0, 0]])
    K, H, W = (60.0, 60.0, 8.0, 8.0), 16, 16
    color, depth, alpha = go.render(means, scales, quats, torch.tensor([0.99, 0.99]), cols, view, K, H, W, torch.zeros(3))
    assert color[8, 8, 1] > 50 * color[8, 8, 0]                    # the near (green) Gaussian hides the far (red) one
    # opaque stack: blending stops before T < 1e-4, so the third layer never contributes
    means3 = torch.tensor([[0.0, 0, 0.0], [0.0, 0, 0.1], [0.0, 0, 0.2]])
    color3, _, alpha3 = go.render(means3, torch.full((3, 3), 0.2), torch.tensor([[1.0, 0, 0, 0]] * 3), torch.tensor([0.97] * 3),
                                  torch.tensor([[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0]]), view, K, H, W, torch.zeros(3))
    assert color3[8, 8, 2].item() == 0.0 and 0.998 < alpha3[8, 8].item() < 0.9995      # third layer: T would drop below 1e-4 -> never blended
