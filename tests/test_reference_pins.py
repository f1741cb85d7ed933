"""Oracles and host-side helpers against outputs of the REFERENCE'S OWN code (tests/golden/reference_pins.npz, generated in the build
container by tests/golden/make_reference_pins.py from /root/reference's torch-only modules, loaded by path / AST extraction).
CPU only.  These pin the restatements that SURVEY.md §8c listed as "parity unpinned although the source is in the tree"."""
import math
import os
import types

import numpy as np
import pytest
import torch

PINS = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_pins.npz'))
T = lambda k: torch.from_numpy(PINS[k])


def test_trunc_exp_oracle():
    from oracle.field_oracle import TruncExpFn
    x = T('truncexp_x').clone().requires_grad_(True)
    y = TruncExpFn.apply(x)
    y.backward(T('truncexp_gy'))
    np.testing.assert_allclose(y.detach().numpy(), PINS['truncexp_y'], rtol=1e-6)
    np.testing.assert_allclose(x.grad.numpy(), PINS['truncexp_gx'], rtol=1e-6)       # clamp [1e-6, 1e6] active at both ends
    assert PINS['truncexp_gx'][-1] == pytest.approx(1.5e6) and PINS['truncexp_gx'][0] == pytest.approx(0.5e-6)


def test_cross_image_attention_oracle():
    """oracle/unet_oracle.attention(num_cross_attn_imgs=2) == CrossImageAttnProcWrapper around an SDPA processor (joint_attn.py:11-37)."""
    from oracle import unet_oracle as uo
    sd = {'a.to_q.weight': T('ja_w_q'), 'a.to_k.weight': T('ja_w_k'), 'a.to_v.weight': T('ja_w_v'), 'a.to_out.0.weight': T('ja_w_o'),
          'a.to_out.0.bias': T('ja_bo')}
    hs, ctx = T('ja_hs'), T('ja_ctx')
    np.testing.assert_allclose(uo.attention(sd, 'a', hs, None, 2, 2).numpy(), PINS['ja_self'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(uo.attention(sd, 'a', hs, None, 2, 1).numpy(), PINS['ja_self_1'], rtol=1e-5, atol=1e-6)
    sd2 = dict(sd, **{'a.to_k.weight': T('ja_w_k2'), 'a.to_v.weight': T('ja_w_v2')})
    np.testing.assert_allclose(uo.attention(sd2, 'a', hs, ctx, 2, 2).numpy(), PINS['ja_cross'], rtol=1e-5, atol=1e-6)


def test_ip_adapter_attention_oracle():
    """oracle/unet_oracle.attention(ip_tokens / drop_tokens) == IPAttnProcessor2_0 / CNAttnProcessor2_0 of the reference."""
    from oracle import unet_oracle as uo
    sd = {'a.to_q.weight': T('ip_wq'), 'a.to_k.weight': T('ip_wk'), 'a.to_v.weight': T('ip_wv'), 'a.to_out.0.weight': T('ip_wo'),
          'a.to_out.0.bias': T('ip_bo'), 'a.to_k_ip.weight': T('ip_wk_ip'), 'a.to_v_ip.weight': T('ip_wv_ip')}
    hs, ctx = T('ip_hs'), T('ip_ctx')
    np.testing.assert_allclose(uo.attention(sd, 'a', hs, ctx, 2, ip_tokens=4, ip_scale=0.7).numpy(), PINS['ip_out'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(uo.attention(sd, 'a', hs, ctx, 2, drop_tokens=2).numpy(), PINS['cn_out'], rtol=1e-5, atol=1e-6)


def test_noise_scales():
    from oracle import nerf_oracle as no
    from mvedit_b200.pipeline import EulerAncestralScheduler
    ab = PINS['ns_alphas_bar']
    for key, t in (('ns_int', T('ns_t_int')), ('ns_flt', T('ns_t_flt'))):
        a, b = no.get_noise_scales(ab, t, 1000)
        np.testing.assert_allclose(np.stack([a.numpy(), b.numpy()]), PINS[key], rtol=1e-6)
    sch = EulerAncestralScheduler()
    np.testing.assert_allclose(sch.alphas_cumprod, ab, rtol=1e-12)
    for j, t in enumerate(T('ns_t_flt')):
        a, b = sch.noise_scales(t)
        assert float(a) == pytest.approx(float(PINS['ns_flt'][0, j]), rel=2e-6) and float(b) == pytest.approx(float(PINS['ns_flt'][1, j]), rel=2e-5, abs=1e-7)
    for j, t in enumerate(T('ns_t_int')):
        a, b = sch.noise_scales(t)
        assert float(a) == pytest.approx(float(PINS['ns_int'][0, j]), rel=2e-6) and float(b) == pytest.approx(float(PINS['ns_int'][1, j]), rel=2e-5)


def test_geometry_oracle_and_host_helpers():
    from oracle import nerf_oracle as no
    from mvedit_b200.nerf import pixel_directions
    K, poses = T('geo_K'), T('geo_poses')
    h = w = PINS['geo_dirs'].shape[2]
    d = no.get_ray_directions(h, w, K[None], norm=False)
    np.testing.assert_array_equal(d.numpy(), PINS['geo_dirs'])
    np.testing.assert_array_equal(no.get_ray_directions(h, w, K[None], norm=True).numpy(), PINS['geo_dirs_n'])
    ro, rd = no.get_rays(d, poses[None], norm=True)
    np.testing.assert_array_equal(ro.numpy(), PINS['geo_ro'])
    np.testing.assert_allclose(rd.numpy(), PINS['geo_rd'], rtol=1e-6, atol=1e-7)
    depth = T('geo_depth')
    np.testing.assert_allclose(no.depth_to_normal(depth, d).numpy(), PINS['geo_normal_gl'], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(no.depth_to_normal(depth, d, format='opencv').numpy(), PINS['geo_normal_cv'], rtol=1e-6, atol=1e-7)
    al = T('geo_alphas')
    np.testing.assert_allclose(no.normalize_depth(depth[0] * al.squeeze(-1), al).numpy(), PINS['geo_depth_norm'], rtol=1e-6, atol=1e-7)
    # the product's direction helper (feeds nothing on the GPU path but BaseNeRF.render's 1/r -> 1/z factor)
    np.testing.assert_allclose(pixel_directions(K[None], h, w).numpy(), PINS['geo_dirs'], rtol=1e-6, atol=1e-7)


def test_tv_loss_oracle():
    from oracle import nerf_oracle as no
    pred, tgt, w = T('tv_pred'), T('tv_tgt'), T('tv_w')
    tv = no.TVLoss(power=1.5)
    # TVLoss.forward = weighted_loss(tv_loss)(..., reduction='mean') * loss_weight: the mean of the per-(batch, channel) values
    assert float(tv(pred)) == pytest.approx(float(PINS['tv_plain'].mean()), rel=1e-6)
    assert float(tv(pred, tgt, weight=w)) == pytest.approx(float(PINS['tv_full'].mean()), rel=1e-6)


def test_ray_sample_and_raybatch():
    """BaseNeRF.ray_sample / get_raybatch_inds (base_nerf.py:245-322): oracle restatement AND the product's gather-only version."""
    from oracle import nerf_oracle as no
    from mvedit_b200.nerf import BaseNeRF
    from mvedit_b200.ingp_decoder import iNGPDecoder
    args = (T('rs_ro'), T('rs_rd'), T('rs_img'), 3 * 16)
    kw = dict(sample_inds=T('rs_inds'), cond_extras=[T('rs_ex0'), T('rs_ex1')])
    orc = no.OracleNeRF(decoder=None, patch_size=4)
    prod = BaseNeRF(grid_size=16, decoder=iNGPDecoder(n_levels=2, max_resolution=32), patch_size=4)
    for impl in (orc, prod):
        outs = impl.ray_sample(*args, **kw)
        assert len(outs) == 5
        for i, o in enumerate(outs):
            np.testing.assert_array_equal(o.numpy(), PINS['rs_out%d' % i])
        torch.manual_seed(int(PINS['rb_seed']))
        rb, nb = impl.get_raybatch_inds(T('rs_img'), 2 * 16)
        assert nb == int(PINS['rb_num']) and rb[0].shape[1] == int(PINS['rb_len0'])
        np.testing.assert_array_equal(torch.cat(list(rb), dim=1).numpy(), PINS['rb_cat'])


def test_camera_pruning_schedules_lights():
    from mvedit_b200 import mvedit_3d_pipeline as mp
    poses, cw = T('prune_poses'), T('prune_cw')
    d = mp.get_camera_dists(poses, cw)
    np.testing.assert_allclose(d.numpy(), PINS['prune_dists'], rtol=2e-5, atol=2e-5)
    keep, d2 = mp.prune_cameras(T('prune_dists'), 2, 5, pixel_dist=T('prune_pix'))
    np.testing.assert_array_equal(keep.numpy(), PINS['prune_keep'])
    np.testing.assert_allclose(d2.numpy(), PINS['prune_dists_after'], rtol=1e-6)
    keep0, _ = mp.prune_cameras(T('prune_dists'), 0, 4)
    np.testing.assert_array_equal(keep0.numpy(), PINS['prune_keep_nopix'])
    ps = PINS['sched_p']
    for n in ('default_lr_multiplier', 'default_max_num_views', 'default_render_size_p', 'default_lr_schedule', 'default_patch_rgb_weight',
              'default_patch_normal_weight', 'default_entropy_weight', 'default_normal_reg_weight'):
        f = getattr(mp, n)
        got = [f(float(p), 0.6) if n in ('default_lr_multiplier', 'default_max_num_views') else f(float(p)) for p in ps]
        np.testing.assert_allclose(np.array(got, np.float64), PINS['sched_' + n], rtol=1e-12, atol=1e-12)
    torch.manual_seed(int(PINS['light_seed']))
    wl, cl = mp.light_sampling(T('geo_poses'))
    np.testing.assert_allclose(cl.numpy(), PINS['light_cam'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(wl.numpy(), PINS['light_world'], rtol=1e-5, atol=1e-6)


def test_tonemapping_module_and_oracle():
    """The tone curve (tonemapping.py:5-52): knots, lut / inverse_lut in both modes (incl. the extrapolating end segments and values on
    the knots), smooth_forward, and shading in tone-mapped space (mvedit_3d_pipeline.py:418-422) -- product module and oracle."""
    from oracle.nerf_oracle import Tonemapping as OT
    from mvedit_b200.tonemapping import Tonemapping as PT
    t = lambda k: torch.from_numpy(PINS[k])
    for cls in (OT, PT):
        tm = cls()
        np.testing.assert_allclose(tm.lut_x.numpy(), PINS['tm_lut_x'], rtol=0, atol=0)
        np.testing.assert_allclose(tm.lut_y.numpy(), PINS['tm_lut_y'], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(tm.lut(t('tm_xs')).numpy(), PINS['tm_lut'], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(tm.lut(torch.exp2(t('tm_xs')), input_mode='linear').numpy(), PINS['tm_lut_lin'], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(tm.inverse_lut(t('tm_ys')).numpy(), PINS['tm_inv'], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(tm.inverse_lut(t('tm_ys'), output_mode='linear').numpy(), PINS['tm_inv_lin'], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(tm.smooth_forward(t('tm_xs')).numpy(), PINS['tm_smooth'], rtol=1e-6, atol=1e-7)
        shaded = tm.lut(tm.inverse_lut(t('tm_alb')) + t('tm_shd').clamp(min=1e-6).log2())
        np.testing.assert_allclose(shaded.numpy(), PINS['tm_shaded'], rtol=1e-5, atol=1e-6)
    arr, n = PT().knots()
    assert n == 16 and np.allclose(np.array(arr[:16]), PINS['tm_lut_x']) and np.allclose(np.array(arr[16:]), PINS['tm_lut_y'], atol=1e-7)


def test_srvgg_enhancer_oracle():
    """oracle/enhancer_oracle.srvgg_forward against the reference's SRVGGNetCompact (image_space_ss.py:8-75) on seeded weights."""
    from oracle.enhancer_oracle import srvgg_forward
    sd = {k[len('sr_sd.'):]: torch.from_numpy(PINS[k]) for k in PINS.files if k.startswith('sr_sd.')}
    assert len(sd) == 5 * 2 + 4                                       # 5 convolutions (weight + bias), 4 PReLUs
    y = srvgg_forward(sd, torch.from_numpy(PINS['sr_x']), num_conv=3, upscale=4)
    np.testing.assert_allclose(y.numpy(), PINS['sr_y'], rtol=1e-5, atol=1e-6)
