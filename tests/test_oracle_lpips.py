"""oracle/lpips_oracle.py on CPU: the properties LPIPS must have (the lpips package is not installed -> no golden vectors)."""
import torch

from oracle import lpips_oracle as lo


def test_identity_symmetry_and_shapes():
    sd = lo.random_lpips_state_dict(0)
    g = torch.Generator().manual_seed(1)
    a, b = torch.rand(2, 3, 32, 32, generator=g), torch.rand(2, 3, 32, 32, generator=g)
    taps = lo.features(sd, a * 2 - 1)
    assert [t.shape[1] for t in taps] == lo.CHANNELS and [t.shape[-1] for t in taps] == [32, 16, 8, 4, 2]
    d_ab, d_ba, d_aa = lo.lpips(sd, a * 2 - 1, b * 2 - 1), lo.lpips(sd, b * 2 - 1, a * 2 - 1), lo.lpips(sd, a * 2 - 1, a * 2 - 1)
    assert d_ab.shape == (2,) and torch.all(d_ab > 0)
    torch.testing.assert_close(d_ab, d_ba)
    assert float(d_aa.abs().max()) == 0.0
    # closer images are closer; the loss is the weighted mean times 1.2
    assert torch.all(lo.lpips(sd, a * 2 - 1, (0.9 * a + 0.1 * b) * 2 - 1) < d_ab)
    w = torch.tensor([0.5, 2.0])
    torch.testing.assert_close(lo.lpips_loss(sd, a, b, w), (d_ab * w).mean() * 1.2)


def test_scale_invariance_of_a_tap():
    """Unit-normalised features: scaling one conv stack's output channel-uniformly must not change the distance -- checked by
    scaling the last conv of slice 5 (weights and bias; ReLU is positively homogeneous)."""
    sd = lo.random_lpips_state_dict(2)
    g = torch.Generator().manual_seed(3)
    a, b = torch.rand(1, 3, 32, 32, generator=g) * 2 - 1, torch.rand(1, 3, 32, 32, generator=g) * 2 - 1
    d0 = lo.lpips(sd, a, b)
    sd2 = dict(sd)
    sd2['net.slice5.28.weight'] = sd['net.slice5.28.weight'] * 3.0
    sd2['net.slice5.28.bias'] = sd['net.slice5.28.bias'] * 3.0
    torch.testing.assert_close(lo.lpips(sd2, a, b), d0, rtol=1e-4, atol=1e-6)
