"""The product's ``nerf_optim`` (``mvedit_b200/nerf.py``: patch-ray kernel, objective kernels, targets, patch terms, optimizer plumbing) run
END TO END ON THE CPU against ``oracle/nerf_oracle.nerf_optim`` -- the restatement that tests/test_nerf_loop_pins.py holds to the
reference's own ``nerf_optim`` code.  The kernels are the product's sources (``recon.cu``, ``nerf_loss.cu``) compiled unchanged as C++
(tests/host_shim); what cannot run without a GPU -- the hash-grid field and the ray marcher -- is replaced ON BOTH SIDES by the same
analytic, differentiable toy field (a soft sphere with a learnable colour map), so every loss term, target gather, weight schedule and
the Adam step are compared value for value over several iterations."""
import math

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from tests import host_harness, synth
from tests.test_mesh_stage_host import _FakePatchLoss
from mvedit_b200 import nerf as pnerf

V, RS, PS, ITERS, N_RAYS = 3, 32, 16, 4, 2 * 16 * 16


class ToyField(nn.Module):
    """rays -> (premultiplied rgb, alpha, sum w / t) of a soft sphere; the decoder interface of both loops."""
    supports_capacity, sample_capacity, max_steps, weight_culling_th, mlp_tf32, grad_sink, state_dict_bak = False, 0, 64, 0.0, False, None, None

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(3)
        self.center = nn.Parameter(torch.tensor([0.05, -0.04, 0.02]))
        self.radius = nn.Parameter(torch.tensor(0.55))
        self.sharp = nn.Parameter(torch.tensor(9.0))
        self.w = nn.Parameter(torch.randn(6, 3, generator=g))
        self.b = nn.Parameter(torch.randn(3, generator=g) * 0.1)

    def check_sample_overflow(self):
        pass

    note_sample_overflow = check_sample_overflow

    def update_extra_state(self, *a, **k):
        pass

    def forward(self, rays_o, rays_d, code, bitfield, grid_size, dt_gamma=None, perturb=False, noises=None, fused_entropy=None, **kw):
        o, d = rays_o.reshape(-1, 3) - self.center, rays_d.reshape(-1, 3)
        tstar = -(o * d).sum(-1)
        closest = o + tstar[:, None] * d
        b = (closest.square().sum(-1) + 1e-8).sqrt()
        alpha = torch.sigmoid((self.radius - b) * self.sharp)
        thit = tstar - (self.radius ** 2 - b ** 2).clamp(min=1e-3).sqrt()
        depth = alpha / thit.clamp(min=0.1)
        col = torch.sigmoid(torch.cat([closest * 3, d], dim=-1) @ self.w + self.b)
        image = col * alpha[:, None]
        return dict(image=image[None], weights_sum=alpha[None], depth=depth[None], weights=image.new_zeros(0), ts=[image.new_zeros(0, 2)])


class ToyNeRF:
    """The attributes ``nerf_optim`` reads from BaseNeRF."""

    def __init__(self, dec, batches):
        self.decoder, self.patch_size, self.grid_size, self.bg_color = dec, PS, 32, 1.0
        self.pixel_loss, self.patch_loss = pnerf.L1LossMod(1.2), _FakePatchLoss()
        self.update_extra_interval, self.update_extra_iters, self.use_cuda_graph, self.data_parallel = 2, 1, False, False
        self._batches = batches

    def get_raybatch_inds(self, imgs, n):
        return self._batches, len(self._batches)


def scene():
    g = torch.Generator().manual_seed(0)
    poses = torch.from_numpy(synth.surround_poses(V, seed=0)).float()
    f = 0.5 * RS / math.tan(math.radians(15))
    intr = torch.tensor([[f, f, RS / 2, RS / 2], [f * 1.1, f * 0.9, RS / 2 + 1, RS / 2 - 0.5], [f, f, RS / 2, RS / 2]])
    yy, xx = torch.meshgrid(torch.arange(RS), torch.arange(RS), indexing='ij')
    disc = (((xx - 15.5) ** 2 + (yy - 15.5) ** 2).float().sqrt() < 9).float()
    masks = disc[None, None, :, :, None].expand(1, V, -1, -1, -1).contiguous()
    images = (torch.rand(1, V, RS, RS, 3, generator=g) * 0.5 + 0.25) * masks + (1 - masks)
    nx, ny = (xx - 15.5) / 12.0, -(yy - 15.5) / 12.0
    n = F.normalize(torch.stack([nx, ny, (1 - nx ** 2 - ny ** 2).clamp(min=0.05).sqrt()], -1) + 0.05 * torch.randn(RS, RS, 3, generator=g), dim=-1)
    normals = (n / 2 + 0.5)[None, None].expand(1, V, -1, -1, -1).contiguous()
    depths = 0.25 + 0.1 * torch.rand(1, V, RS, RS, 1, generator=g)
    batches = torch.randperm(V * (RS // PS) ** 2, generator=g)[None].split(N_RAYS // PS ** 2, dim=1)
    return poses, intr, images, masks, normals, depths, torch.tensor([1.0, 0.5, 2.0]), F.normalize(torch.randn(V, 3, generator=g), dim=-1), batches


CASES = dict(
    init=dict(is_init=True, init_shaded=False, patch_rgb=0.0),
    shaded=dict(is_init=False, init_shaded=False, patch_rgb=0.4),
    tone=dict(is_init=False, init_shaded=False, patch_rgb=0.4, tone=True),
    normals=dict(is_init=False, init_shaded=False, patch_rgb=0.4, normals=True, patch_normal=0.0),
    normals_patch=dict(is_init=False, init_shaded=False, patch_rgb=0.4, normals=True, patch_normal=0.7),
    depths=dict(is_init=True, init_shaded=True, patch_rgb=0.0, depths=True, depth_weight=0.3),
    all=dict(is_init=False, init_shaded=False, patch_rgb=0.4, normals=True, patch_normal=0.7, depths=True, depth_weight=0.3, tone=True))


@pytest.mark.parametrize('case', list(CASES))
def test_product_nerf_optim_matches_the_oracle_loop(case):
    from oracle import nerf_oracle as no
    from mvedit_b200.tonemapping import Tonemapping
    c = CASES[case]
    poses, intr, images, masks, normals, depths, cam_w, cam_lights, batches = scene()
    tn = normals if c.get('normals') else None
    td = depths if c.get('depths') else None
    sched = dict(lr=0.02, alpha_soften=0.02, normal_reg=0.1, entropy=0.01, bg_width=0.015, ambient=0.2, dt_gamma_scale=1.0)

    grid, bits = torch.zeros(1, 32 ** 3, dtype=torch.float16), torch.full((1, 32 ** 3 // 8), 255, dtype=torch.uint8)

    def args(nerf, opt):
        return (nerf, images, masks, tn, opt, sched['lr'], ITERS, N_RAYS, c['patch_rgb'], c.get('patch_normal', 0.0), sched['alpha_soften'],
                sched['normal_reg'], sched['entropy'], [None], grid, bits, RS, intr, RS, poses, cam_w, cam_lights, PS, c['is_init'],
                sched['bg_width'], sched['ambient'], sched['dt_gamma_scale'], c['init_shaded'])
    # ---- oracle loop
    dec_o = ToyField()
    nerf_o = no.OracleNeRF(dec_o, grid_size=32, patch_size=PS, update_extra_interval=2)
    nerf_o.patch_loss = _FakePatchLoss()
    dec_o.update_extra_state = lambda *a, **k: None
    log_o = no.nerf_optim(*args(nerf_o, torch.optim.Adam(dec_o.parameters(), lr=0.01)), debug=True, tgt_depths=td,
                          depth_weight=c.get('depth_weight', 0.0), raybatch_inds=list(batches),
                          tonemapping=no.Tonemapping() if c.get('tone') else None)
    # ---- product loop on the host builds of its kernels
    dec_p = ToyField()
    nerf_p = ToyNeRF(dec_p, batches)
    libs = host_harness.Libraries(host_harness.shimmed('recon.cu'), host_harness.shimmed('nerf_loss.cu'))
    with host_harness.routed(pnerf, libs):
        log_p = pnerf.nerf_optim(*args(nerf_p, torch.optim.Adam(dec_p.parameters(), lr=0.01)), debug=True, tgt_depths=td,
                                 depth_weight=c.get('depth_weight', 0.0), tonemapping=Tonemapping() if c.get('tone') else None)
    assert len(log_p) == len(log_o) == ITERS
    for it, (a, b) in enumerate(zip(log_p, log_o)):
        for k in ('pixel_rgb', 'alpha', 'normal_reg', 'entropy'):
            assert abs(a[k] - b[k]) <= 2e-4 * abs(b[k]) + 1e-6, (it, k, a[k], b[k])
        assert abs(a['loss'] - b['loss']) <= 2e-4 * abs(b['loss']) + 1e-6, (it, a['loss'], b['loss'])
    moved = 0.0
    for (k, p), q in zip(dec_p.named_parameters(), dec_o.parameters()):
        assert (p - q).abs().max() <= 2e-4, (k, float((p - q).abs().max()))
        moved = max(moved, float((p.detach() - ToyField().state_dict()[k]).abs().max()))
    assert moved > 0.05                                                        # four Adam steps at lr 0.02
