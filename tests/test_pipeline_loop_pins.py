"""The loop body of ``MVEdit3DPipeline.__call__`` against THE REFERENCE'S OWN ``__call__``:
tests/golden/make_pipeline_loop_pins.py ran mvedit_3d_pipeline.py:875-1500 unmodified (with the reference's own input loaders, denoiser
mixin, noise scales, light sampling, camera pruning, depth normalisation and schedules) around toy components and recorded every
``nerf_optim`` call.  Here the product's ``__call__`` runs around the same toys: at every step it must hand the same targets, cameras
(i.e. the same re-ordering and pruning decisions), lights, weights and flags to ``nerf_optim`` -- in the optimisation-only mode, in both
denoise modes, with dynamic blending, with reference-image pairs, from pure noise, with image-to-3D targets -- and, past
``progress_to_dmtet``, the same arguments to ``mesh_optim`` (tet initialisation, two-group optimiser, ``is_end``) with the mesh branch of
the per-step render feeding the next step.  CPU.

The product keeps the rendered conditions in bf16 where the reference run here used fp32: value comparisons carry that tolerance."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import nerf_oracle as no
from mvedit_b200 import mvedit_3d_pipeline as P
from mvedit_b200.schedulers import EulerAncestralScheduler

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location('make_pipeline_loop_pins', os.path.join(HERE, 'golden', 'make_pipeline_loop_pins.py'))
gen = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gen)
PINS = np.load(os.path.join(HERE, 'golden', 'pipeline_loop_pins.npz'))


class AdamLike(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, **kw):
        super().__init__(params, lr=lr)


class _FieldForOracleRender:
    """``oracle.nerf_oracle.render_views`` (the restatement of the reference's render + shading lines, :1363-1395) over the toy field."""

    def __init__(self, field):
        self.field, self.bg_color = field, field.bg_color

    def render(self, density_bitfield, h, w, intrinsics, poses, cfg=None, normal_bg=(0.5, 0.5, 1.0)):
        return self.field.render(None, None, density_bitfield, h, w, intrinsics, poses, cfg=cfg, normal_bg=normal_bg)


@pytest.mark.parametrize('case', list(gen.CASES))
def test_call_hands_nerf_optim_what_the_reference_loop_hands_it(case, monkeypatch):
    field, log = gen.ToyField(), []
    monkeypatch.setattr(P, 'FusedAdam', AdamLike)
    monkeypatch.setattr(P, 'nerf_optim', lambda nerf, *a, **k: gen.record_call(log, field, *a, **k))
    monkeypatch.setattr(P, 'mesh_optim', lambda pipe_, *a, **k: gen.record_mesh_call(log, field, *a, **k))
    monkeypatch.setattr(P, 'init_tet', gen.toy_init_tet)
    pipe = P.MVEdit3DPipeline(gen.ToyVAE(), None, None, gen.ToyUNet(), gen.mixin_gen.toy_nets(2), EulerAncestralScheduler(), field,
                              mesh_renderer=gen.ToyMeshRenderer(field), image_enhancer=gen.ToyEnhancer(), segmentation=gen.toy_segmentation)
    wrapped = _FieldForOracleRender(field)
    pipe.render_views = lambda bitfield, poses, intr, intr_size, rs, cam_lights, ambient, tdg, render_bs=None, **kw: no.render_views(
        wrapped, bitfield, poses, intr, intr_size, rs, cam_lights, ambient, tdg, render_bs=render_bs, out_dtype=torch.float32)
    poses, intr, init, embeds = gen.inputs()
    kw = gen.call_kwargs(case, poses, intr, init, embeds)
    torch.manual_seed(1234)
    mesh, state = pipe(prompt_embeds=embeds.clone(), **kw)
    assert state is not None, 'the run raised inside __call__ (traceback printed above)'
    assert (mesh is not None) == (case == 'dmtet')                              # the DMTet mesh comes back once that stage was entered
    assert len(log) == int(PINS[case + '_steps'])
    for i, rec in enumerate(log):
        for opt in ('tgt_normals', 'tgt_depths'):                                       # handed over exactly when the reference does
            assert (opt in rec) == ('%s_%d_%s_pooled' % (case, i, opt) in PINS.files), (i, opt)
        for k, v in rec.items():
            key = '%s_%d_%s' % (case, i, k)
            if k in gen.MAPS:
                assert tuple(v.shape) == tuple(PINS[key + '_shape']), (key, v.shape)
                x = v[0].permute(0, 3, 1, 2)
                d = np.abs(torch.nn.functional.avg_pool2d(x, 8).numpy() - PINS[key + '_pooled'])
                # the masks carry a hard colour threshold (do_segmentation's background rule): a bf16-rounded pixel may sit on the other side
                assert d.max() <= (0.1 if k == 'tgt_masks' else 1.5e-2) and (d > 1.5e-2).mean() <= 0.005, (key, float(d.max()), float((d > 1.5e-2).mean()))
                np.testing.assert_allclose(x.flatten(2).std(dim=2).numpy(), PINS[key + '_std'], rtol=0, atol=1.5e-2, err_msg=key)
            elif torch.is_tensor(v):
                np.testing.assert_allclose(v.numpy(), PINS[key], rtol=1e-5, atol=1e-6, err_msg=key)       # cameras, intrinsics, weights, lights
            else:
                assert float(v) == pytest.approx(float(PINS[key]), rel=1e-6), key                       # schedules, sizes, flags
    assert field.decoder.state_dict_bak is not None
    if kw.get('ip_adapter') is not None:                                         # what the IP-Adapter was shown: CLIP size, CLIP normalisation
        assert len(kw['ip_adapter'].seen) == 1
        np.testing.assert_allclose(torch.nn.functional.avg_pool2d(kw['ip_adapter'].seen[0], 16).numpy(), PINS[case + '_ipa_images'], rtol=1e-4, atol=1e-4)


def test_step_is_one_iteration_of_the_pinned_loop(monkeypatch):
    """``MVEdit3DStep.step`` -- the function bench.py times -- against the 2-pass iterations of ``__call__`` (which the test above holds to
    the reference's own loop): from the latents ``__call__`` had before a solver step, with the ancestral noise it drew, ``step`` must hand
    ``nerf_optim`` the same targets and return the same latents and conditions."""
    field, log, solver = gen.ToyField(), [], []
    field.patch_size = 64                                                        # (step() reads the patch size off the field)
    monkeypatch.setattr(P, 'FusedAdam', AdamLike)
    monkeypatch.setattr(P, 'nerf_optim', lambda nerf, *a, **k: gen.record_call(log, field, *a, **k))
    sch = EulerAncestralScheduler()
    pipe = P.MVEdit3DPipeline(gen.ToyVAE(), None, None, gen.ToyUNet(), gen.mixin_gen.toy_nets(2), sch, field, segmentation=gen.toy_segmentation)
    wrapped = _FieldForOracleRender(field)
    pipe.render_views = lambda bitfield, poses, intr, intr_size, rs, cam_lights, ambient, tdg, render_bs=None, **kw: no.render_views(
        wrapped, bitfield, poses, intr, intr_size, rs, cam_lights, ambient, tdg, render_bs=render_bs, out_dtype=torch.float32)
    real_step = sch.step

    def spy(model_output, t, sample, noise):
        out = real_step(model_output, t, sample, noise)
        solver.append(dict(t=t, latents_in=sample.clone(), noise=noise.clone(), latents_out=out.clone(), fits=field.fits))
        return out
    sch.step = spy
    poses, intr, init, embeds = gen.inputs()
    kw = gen.call_kwargs('two_pass', poses, intr, init, embeds)
    kw.update(keep_views=None, max_num_views=lambda p, q: gen.N, seg_padding=0)  # no re-ordering, no pruning, no padding: what step() does
    torch.manual_seed(99)
    mesh, state = pipe(prompt_embeds=embeds.clone(), **kw)
    assert state is not None and len(solver) == 2 and len(log) == 4              # the last iteration fits and stops: no solver step (:1409)
    sch.step = real_step
    import mvedit_b200.pipeline as S
    for k, sv in enumerate(solver):
        rec, step_log = log[k + 1], []                                           # log[0] is the initial fit (t is None)
        monkeypatch.setattr(S, 'nerf_optim', lambda nerf, *a, **kk: gen.record_call(step_log, field, *a, **kk))
        field.fits = sv['fits'] - 1                                              # the field as it was before this iteration's fit
        i = int((sch.timesteps == float(sv['t'])).nonzero()[0])
        latents, ctrl_images, ctrl_depths = pipe.step(
            i, sv['latents_in'], embeds.clone(), None, None, None, None, rec['camera_poses'], rec['intrinsics'], rec['intrinsics_size'],
            rec['cam_weights'], rec['cam_lights'], sv['noise'], guidance_scale=kw['guidance_scale'], render_size=rec['render_size'],
            n_inverse_steps=rec['inverse_steps'], n_inverse_rays=rec['n_inverse_rays'], lr=rec['lr'], alpha_soften=rec['alpha_soften'],
            normal_reg_weight=rec['normal_reg_weight'], entropy_weight=rec['entropy_weight'], patch_rgb_weight=rec['patch_rgb_weight'],
            bg_width=rec['bg_width'], ambient_light=rec['ambient_light'], dt_gamma_scale=rec['dt_gamma_scale'], render_bs=kw['render_bs'])
        assert len(step_log) == 1
        for name in ('tgt_images', 'tgt_masks', 'camera_poses', 'cam_lights'):
            torch.testing.assert_close(step_log[0][name], rec[name], rtol=1e-5, atol=1e-5)
        for name in ('lr', 'inverse_steps', 'n_inverse_rays', 'render_size', 'is_init', 'init_shaded'):
            assert step_log[0][name] == rec[name], name
        torch.testing.assert_close(latents, sv['latents_out'], rtol=1e-4, atol=1e-4)
        assert ctrl_images.shape == (gen.N, 3, 512, 512) and ctrl_depths.shape == (gen.N, 3, 512, 512)
