"""CPU sanity of the field oracle (oracle/field_oracle.py): level table vs SURVEY.md Appendix B, interpolation properties,
autograd.  (No tcnn vector exists offline -- 'parity unpinned', see the oracle's header.)"""
import numpy as np
import torch

from oracle import field_oracle as fo


def test_level_table():
    levels, n = fo.level_table(12, 16, 320)
    assert [l[1] for l in levels] == [16, 22, 28, 37, 48, 63, 82, 108, 142, 186, 244, 320]
    assert n == 3593720
    assert all(l[2] == 1 << 19 for l in levels[6:]) and all(l[2] < 1 << 19 for l in levels[:6])
    assert abs(fo.per_level_scale(320, 1.0, 16, 12) - 1.31303) < 1e-4
    levels, n = fo.level_table(14, 16, 512)
    assert n == 4594792 and levels[-1][1] == 512


def test_encode_interpolates_grid_values_at_cell_corners():
    levels, n = fo.level_table(12, 16, 320)
    g = torch.Generator().manual_seed(0)
    table = torch.randn(n, 2, generator=g, dtype=torch.float64)
    # level 0 is dense with res 16, scale 15: x01 = (k - 0.5)/15 puts pos exactly on integer k -> weight 1 on corner (k,k,k)
    scale, res, size, off = levels[0]
    k = 5
    x = torch.full((1, 3), (k - 0.5) / scale, dtype=torch.float64)
    enc = fo.hash_encode(x, table, levels)
    idx = (k + k * res + k * res * res) % size
    np.testing.assert_allclose(enc[0, :2].numpy(), table[off + idx].numpy(), rtol=1e-9)


def test_point_decode_autograd_and_ranges():
    levels, n = fo.level_table(12, 16, 320)
    params = [p.double().requires_grad_(True) for p in fo.init_params(levels, n, table_scale=0.5)]
    # seeded: an unlucky draw can put x[0] on a ReLU kink of the MLP, where the one-sided finite difference below disagrees
    x = (torch.rand(64, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(7)) * 2 - 1).requires_grad_(True)
    sigma, rgb = fo.point_decode(x, *params, levels)
    assert sigma.min() > 0 and rgb.min() >= -0.001 and rgb.max() <= 1.001
    (sigma.sum() + rgb.sum()).backward()
    assert all(p.grad is not None for p in params) and x.grad.abs().sum() > 0
    # finite-difference check of d sigma / d x on one coordinate (smoothstep is C1)
    eps = 1e-6
    xp = x.detach().clone(); xp[0, 0] += eps
    s2, _ = fo.point_decode(xp, *[p.detach() for p in params], levels)
    gs = torch.autograd.grad(fo.point_decode(x, *params, levels)[0][0], x)[0][0, 0]
    assert abs((s2[0] - sigma[0].detach()) / eps - gs) < 1e-3 * (1 + abs(gs))
