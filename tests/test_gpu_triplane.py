"""(1) the stand-alone hash-grid encoding (seam B4: ``HashGridEncoding.__call__`` -> mve_hashgrid_forward / _backward) vs the oracle's
restatement of tiny-cuda-nn's grid.h (oracle/field_oracle.hash_encode; PARITY UNPINNED: tcnn is not installable offline);
(2) ``TriPlaneiNGPDecoder.point_decode`` vs the oracle's restatement of
/root/reference/lib/models/decoders/triplane_ingp_decoder.py:142-212, forward and gradients;
(3) the tri-plane decoder through nerf_optim and BaseNeRF.render (reference protocol, no fused renderer)."""
import math

import pytest
import torch

from oracle import field_oracle as fo
from tests import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('L,max_res', [(12, 320), (14, 512)])
def test_hashgrid_encoding_forward_backward(L, max_res):
    from mvedit_b200.ingp_decoder import HashGridEncoding
    enc = HashGridEncoding(n_levels=L, max_resolution=max_res).cuda()
    g = torch.Generator(device='cuda').manual_seed(L)
    with torch.no_grad():
        enc.params.copy_((torch.rand(enc.params.shape, device='cuda', generator=g) * 2 - 1) * 0.3)
    levels, n_entries = fo.level_table(L, 16, max_res)
    assert n_entries * 2 == enc.params.numel() and enc.n_output_dims == 2 * L
    x = torch.rand(5000, 3, device='cuda', generator=g)
    x[:8] = torch.tensor([[0, 0, 0], [1, 1, 1], [0.5, 0.5, 0.5], [1, 0, 0.25], [0, 1, 0], [0.999, 0.001, 0.5], [0.25, 0.75, 1], [1, 1, 0]], device='cuda')
    w = torch.randn(5000, 2 * L, device='cuda', generator=g)
    x1 = x.clone().requires_grad_(True)
    out = enc(x1)
    (out * w).sum().backward()
    x2 = x.clone().requires_grad_(True)
    tab = enc.params.detach().clone().requires_grad_(True)
    ref = fo.hash_encode(x2, tab.view(-1, 2), levels)
    (ref * w).sum().backward()
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=5e-5)        # fma order differs from the torch chain
    assert (enc.params.grad - tab.grad).abs().max().item() <= 2e-4 * tab.grad.abs().max().item()
    assert (x1.grad - x2.grad).abs().max().item() <= 2e-4 * x2.grad.abs().max().item() + 1e-5


def _decoder():
    from mvedit_b200.triplane_ingp_decoder import TriPlaneiNGPDecoder
    torch.manual_seed(0)
    dec = TriPlaneiNGPDecoder(plane_cfg=['yx', 'yz', 'xz'], flip_z=True, base_layers=[48, 64], density_layers=[64, 1], color_layers=[64, 3],
                              max_steps=256, weight_culling_th=0.001).cuda()
    with torch.no_grad():
        dec.encoder.params.uniform_(-0.2, 0.2)
        torch.nn.init.xavier_uniform_(dec.ingp_base_net[-1].weight)          # the zero-initialised hash head would hide the encoding
    return dec


def test_triplane_point_decode_matches_oracle():
    dec = _decoder()
    g = torch.Generator(device='cuda').manual_seed(1)
    code = torch.randn(1, 3, 16, 20, 20, device='cuda', generator=g)
    xyz = (torch.rand(4000, 3, device='cuda', generator=g) * 2 - 1) * 0.98
    levels, _ = fo.level_table(12, 16, 320)
    assert set(dec.state_dict().keys()) == {'aabb', 'encoder.params', 'base_net.0.weight', 'base_net.0.bias', 'ingp_base_net.0.weight',
                                           'ingp_base_net.0.bias', 'density_net.0.weight', 'density_net.0.bias', 'color_net.0.weight', 'color_net.0.bias'}
    x1 = xyz.clone().requires_grad_(True)
    s1, c1, n = dec.point_decode([x1], None, code)
    ws, wc = torch.randn(4000, device='cuda', generator=g), torch.randn(4000, 3, device='cuda', generator=g)
    ((s1 * ws).sum() + (c1 * wc).sum()).backward()
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and k != 'aabb') for k, v in dec.state_dict().items()}
    x2 = xyz.clone().requires_grad_(True)
    s2, c2 = fo.triplane_point_decode(x2, code, sd, levels, plane_cfg=['yx', 'yz', 'xz'], flip_z=True)
    ((s2 * ws).sum() + (c2 * wc).sum()).backward()
    torch.testing.assert_close(s1, s2, rtol=2e-4, atol=1e-6)
    torch.testing.assert_close(c1, c2, rtol=2e-4, atol=1e-6)
    for name, p in dec.named_parameters():
        ref = sd[name].grad
        assert (p.grad - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-7, name
    assert (x1.grad - x2.grad).abs().max().item() <= 2e-3 * x2.grad.abs().max().item() + 1e-6
    sd_only, _, _ = dec.point_decode([xyz], None, code, density_only=True)
    torch.testing.assert_close(sd_only, s1.detach())


def test_triplane_decoder_through_nerf_optim_and_render():
    from mvedit_b200.nerf import BaseNeRF, nerf_optim
    from mvedit_b200.optim import FusedAdam
    from oracle.nerf_oracle import get_ray_directions, get_rays
    dec = _decoder()
    V, size, ps = 4, 64, 32
    nerf = BaseNeRF(grid_size=64, decoder=dec, patch_size=ps).cuda()
    poses = torch.from_numpy(synth.surround_poses(V, seed=3)).cuda()
    f = 0.5 * size / math.tan(math.radians(15))
    K = torch.tensor([[f, f, size / 2, size / 2]] * V, device='cuda')
    d = get_ray_directions(size, size, K[None], device='cuda')
    ro, rd = get_rays(d, poses[None], norm=True)
    b = (ro * rd).sum(-1)
    hit = (b * b - ((ro * ro).sum(-1) - 0.25)) > 0
    img = torch.where(hit[..., None], torch.full_like(ro, 0.3), torch.ones_like(ro))
    msk = hit[..., None].float()
    code = torch.randn(1, 3, 16, 20, 20, device='cuda') * 0.1
    grid, bits = nerf.get_init_density_grid(1, 'cuda'), nerf.get_init_density_bitfield(1, 'cuda')
    opt = FusedAdam(dec.parameters(), lr=0.01)
    log = nerf_optim(nerf, img, msk, None, optimizer=opt, lr=0.01, inverse_steps=80, n_inverse_rays=ps * ps * 2, patch_rgb_weight=0.0,
                     patch_normal_weight=0.0, alpha_soften=0.02, normal_reg_weight=0.1, entropy_weight=0.01, nerf_code=code, density_grid=grid,
                     density_bitfield=bits, render_size=size, intrinsics=K, intrinsics_size=size, camera_poses=poses,
                     cam_weights=torch.ones(V, device='cuda'), cam_lights=torch.nn.functional.normalize(torch.randn(V, 3, device='cuda'), dim=-1),
                     patch_size=ps, is_init=True, bg_width=0.015, ambient_light=0.2, dt_gamma_scale=0.5, init_shaded=False, debug=True)
    first = sum(l['pixel_rgb'] + l['alpha'] for l in log[:8]) / 8
    last = sum(l['pixel_rgb'] + l['alpha'] for l in log[-8:]) / 8
    assert last < 0.6 * first, (first, last)
    rgba, depth = nerf.render(dec, code, bits, size, size, K[None], poses[None], cfg=dict(dt_gamma_scale=0.5, return_rgba=True))
    assert rgba.shape == (1, V, size, size, 4) and torch.isfinite(rgba).all()
    assert (rgba[..., 3:] - msk).abs().mean().item() < 0.2
