"""The two glue kernels of a reconstruction iteration (csrc/recon.cu) against what they replace:
  mve_adam_step   vs torch.optim.Adam (+ zero_grad)                                   (mvedit_3d_pipeline.py:631-633)
  mve_patch_rays  vs the oracle's ray_sample / get_ray_directions / get_rays chain    (base_nerf.py:245-303, geometry_utils.py:18-55;
                                                                                       pinned by tests/test_reference_pins.py)"""
import math

import numpy as np
import pytest
import torch

from tests import synth

pytestmark = pytest.mark.gpu


def test_fused_adam_matches_torch_adam():
    from mvedit_b200.optim import FusedAdam
    g = torch.Generator(device='cuda').manual_seed(0)
    shapes = [(100003,), (64, 24), (64,), (4, 64), (4,)]
    pa = [torch.nn.Parameter(torch.randn(*s, device='cuda', generator=g)) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    ref = torch.optim.Adam(pa, lr=0.01)
    opt = FusedAdam(pb, lr=0.01)
    for it in range(7):
        lr = 0.01 - 0.001 * it
        ref.param_groups[0]['lr'] = lr
        opt.set_lr(lr)
        grads = [torch.randn(*s, device='cuda', generator=g) * (0.1 if it % 2 else 3.0) for s in shapes]
        ref.zero_grad()
        for p, q, gr in zip(pa, pb, grads):
            p.grad = gr.clone()
            opt.grad_sink(q).add_(gr)                   # what the backward kernels do: accumulate into the flat buffer
        ref.step()
        opt.step()
        assert float(opt.flat_grad.abs().max()) == 0.0  # consumed gradients are zeroed by the same launch
        for p, q in zip(pa, pb):
            torch.testing.assert_close(q, p, rtol=2e-6, atol=2e-7)
    assert int(opt._step) == 7
    # autograd-attached gradients (a parameter whose .grad was replaced) are folded in too
    pb[1].grad = torch.ones_like(pb[1])
    before = pb[1].detach().clone()
    opt.step()
    assert (pb[1] - before).abs().max() > 0 and pb[1].grad.data_ptr() == opt.grad_sink(pb[1]).data_ptr()


@pytest.mark.parametrize('row_range', [(0, 8), (2, 6)])
def test_patch_rays_matches_oracle_chain(row_range):
    from oracle import nerf_oracle as no
    from mvedit_b200._lib import call, ptr, stream, c_u32, c_f32
    V, rs, ps = 5, 32, 8
    poses = torch.from_numpy(synth.surround_poses(V, seed=2)).cuda()
    f = 0.5 * 48 / math.tan(math.radians(15))
    K = torch.tensor([[f, f * 1.1, 24.5, 23.0]] * V, device='cuda') * torch.linspace(1.0, 1.3, V, device='cuda')[:, None]   # at size 48
    g = torch.Generator(device='cuda').manual_seed(1)
    img, msk = torch.rand(1, V, rs, rs, 3, device='cuda', generator=g), torch.rand(1, V, rs, rs, 1, device='cuda', generator=g)
    camw, lights = 0.5 + torch.rand(V, device='cuda', generator=g), torch.randn(V, 3, device='cuda', generator=g)
    inds = torch.tensor([37, 3, 64], device='cuda')
    P = inds.numel()
    lo, hi = row_range
    n, nl = P * ps * ps, P * (hi - lo) * ps
    f32 = dict(dtype=torch.float32, device='cuda')
    ro, rd, dirs, trgb, tmsk = torch.empty(nl, 3, **f32), torch.empty(nl, 3, **f32), torch.empty(n, 3, **f32), torch.empty(n, 3, **f32), torch.empty(n, **f32)
    pw, pl, dtg = torch.empty(P, **f32), torch.empty(P, 3, **f32), torch.empty(1, **f32)
    Rm, Tm = poses[:, :3, :3].contiguous(), poses[:, :3, 3].contiguous()      # keep alive: ptr() of a temporary dangles once it is freed
    call('mve_patch_rays', ptr(inds), c_u32(P), c_u32(V), c_u32(rs), c_u32(ps), ptr(Rm), ptr(Tm),
         ptr(K), c_f32(rs / 48), ptr(img[0]), ptr(msk[0]), ptr(camw), ptr(lights), c_f32(0.7), c_u32(lo), c_u32(hi), ptr(ro), ptr(rd), ptr(dirs),
         ptr(trgb), ptr(tmsk), ptr(pw), ptr(pl), ptr(dtg), stream())
    # the reference chain (oracle restatement)
    d_all = no.get_ray_directions(rs, rs, K[None] * (rs / 48), norm=False, device='cuda')
    ro_all, rd_all = no.get_rays(d_all, poses[None], norm=True)
    cam_ids = torch.arange(V, device='cuda')[None, :, None, None, None].expand(-1, -1, rs, rs, -1).float()
    nerf = no.OracleNeRF(None, patch_size=ps)
    o_ro, o_rd, o_rgb, o_msk, o_dir, o_cam = nerf.ray_sample(ro_all, rd_all, img, n, sample_inds=inds[None], cond_extras=[msk, d_all, cam_ids])
    cam = o_cam[:, 0, 0, 0].long()
    strip = lambda t: t.reshape(P, ps, ps, -1)[:, lo:hi].reshape(nl, -1)
    torch.testing.assert_close(dirs, o_dir.reshape(n, 3), rtol=1e-6, atol=1e-7)
    assert torch.equal(trgb, o_rgb.reshape(n, 3)) and torch.equal(tmsk, o_msk.reshape(n))
    assert torch.equal(ro, strip(o_ro))
    torch.testing.assert_close(rd, strip(o_rd), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(pw, camw[cam] / camw.mean(), rtol=1e-6, atol=0)
    assert torch.equal(pl, lights[cam])
    ref_dtg = 0.7 / (K[cam, :2].mean(dim=-1) * rs / 48)
    assert float(dtg) == pytest.approx(float(ref_dtg[0]), rel=1e-6)


def test_im2col_bottom_right_padding():
    """pad_lo = 0 (AutoencoderKL encoder Downsample2D: F.pad (0,1,0,1) + stride-2 conv, padding 0) vs F.conv2d."""
    import torch.nn.functional as F
    from mvedit_b200 import tc_ops as T
    g = torch.Generator(device='cuda').manual_seed(0)
    x = torch.randn(2, 64, 16, 16, device='cuda', generator=g)
    w = torch.randn(64, 64, 3, 3, device='cuda', generator=g) / 24
    xb = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    wb = w.permute(0, 2, 3, 1).reshape(64, -1).to(torch.bfloat16)
    out = T.gemm(T.im2col3x3s2(xb, pad_lo=0), wb).view(2, 8, 8, 64).permute(0, 3, 1, 2)
    ref = F.conv2d(F.pad(xb.float().permute(0, 3, 1, 2), (0, 1, 0, 1)), wb.float().view(64, 3, 3, 64).permute(0, 3, 1, 2), stride=2)
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=2e-2)
