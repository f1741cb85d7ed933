"""``init_tet`` and ``Adapter3DMixin.load_init_mesh`` against THE REFERENCE'S OWN functions (tests/golden/make_stage_glue_pins.py ran
lib/pipelines/utils.py:156-184 and adapter3d_mixin.py:21-66 unmodified).  CPU."""
import importlib.util
import os
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location('make_stage_glue_pins', os.path.join(HERE, 'golden', 'make_stage_glue_pins.py'))
gen = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gen)
PINS = np.load(os.path.join(HERE, 'golden', 'stage_glue_pins.npz'))


def test_init_tet_matches_the_reference_function():
    from mvedit_b200.mesh_optim import init_tet
    from mvedit_b200.mesh_renderer import make_tet_grid
    verts, indices, sdf = init_tet(SimpleNamespace(decoder=gen.BlobDensity()), None, density_thresh=5.0, resolution=12, tets=make_tet_grid(12))
    np.testing.assert_allclose(verts.numpy(), PINS['tet_verts'], rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(indices.numpy(), PINS['tet_indices'])
    np.testing.assert_allclose(sdf.numpy(), PINS['tet_sdf'], rtol=1e-5, atol=1e-6)
    assert (sdf == -1).any() and (sdf == 1).any() and ((sdf > -1) & (sdf < 1)).any()


def test_load_init_mesh_matches_the_reference_method():
    from mvedit_b200.adapter3d_mixin import Adapter3DMixin

    class Pipe(Adapter3DMixin):
        pass
    poses, intr = gen.glue_inputs()
    pipe, rend = Pipe(), gen.RecordingRenderer()
    pipe.mesh_renderer, pipe.bg_color = rend, 0.7
    mesh = gen.ToyMesh()
    mesh.vn = 1                                               # (the product computes vertex normals when the input mesh has none)
    funs = ['f0', 'f1', 'f2']
    m, images, alphas, depths = pipe.load_init_mesh(mesh, poses, intr, 32, 2, funs, diff_size=48)
    assert m is mesh and rend.ssaa == 1
    for got, key in ((images, 'lim_images'), (alphas, 'lim_alphas'), (depths, 'lim_depths')):
        np.testing.assert_allclose(got.numpy(), PINS[key], rtol=1e-6, atol=1e-7)
    assert [c['ssaa'] for c in rend.calls] == PINS['lim_ssaa'].tolist() == [2, 2, 2]
    np.testing.assert_allclose(torch.cat([c['intrinsics'][0] for c in rend.calls]).numpy(), PINS['lim_intr'], rtol=1e-6)
    assert [[c['h'], c['w']] for c in rend.calls] == PINS['lim_sizes'].tolist()
    assert [funs.index(c['fun']) for c in rend.calls] == PINS['lim_funs'].tolist()
    rend2 = gen.RecordingRenderer()
    pipe.mesh_renderer = rend2
    pipe.load_init_mesh(mesh, poses, intr, 32, 4, None)
    assert [[c['h'], c['w']] for c in rend2.calls] == PINS['lim_default_sizes'].tolist() and all(c['fun'] is None for c in rend2.calls)


def test_field_composition_matches_the_reference_decoder():
    """``iNGPDecoder.point_decode`` (input normalisation, MLP, density blob, TruncExp, saturated sigmoid) run unmodified around the oracle's
    hash grid == ``oracle/field_oracle.point_decode`` (what the GPU tests hold the field kernels to)."""
    from oracle import field_oracle as fo
    xyz, (table, w1, b1, w2, b2), levels = gen.decode_inputs()
    sig, rgb = fo.point_decode(xyz, table, w1, b1, w2, b2, levels)
    np.testing.assert_allclose(sig.numpy(), PINS['dec_sigma'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(rgb.numpy(), PINS['dec_rgb'], rtol=1e-5, atol=1e-6)
    assert PINS['dec_sigma'].std() > 0.1 and PINS['dec_rgb'].std() > 0.02


def test_triplane_composition_matches_the_reference_decoder():
    """``TriPlaneiNGPDecoder.point_decode`` + ``xyz_transform`` run unmodified (two plane layouts, with and without the z flip) ==
    ``oracle/field_oracle.triplane_point_decode`` (what tests/test_gpu_triplane.py holds the kernels to)."""
    from oracle import field_oracle as fo
    xyz, code, sd, levels = gen.triplane_inputs()
    for name, plane_cfg, flip in (('a', ('xy', 'xz', 'yz'), False), ('b', ['yx', 'yz', 'xz'], True)):
        sig, rgb = fo.triplane_point_decode(xyz, code, sd, levels, plane_cfg=plane_cfg, flip_z=flip)
        np.testing.assert_allclose(sig.numpy(), PINS['tri_sigma_' + name], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(rgb.numpy(), PINS['tri_rgb_' + name], rtol=1e-5, atol=1e-6)
    assert np.abs(PINS['tri_rgb_a'] - PINS['tri_rgb_b']).max() > 0.05


def test_make_shading_fun_matches_the_reference_method():
    from mvedit_b200.mvedit_3d_pipeline import MVEdit3DPipeline
    from mvedit_b200.tonemapping import Tonemapping
    lights, albedo, normal, fg = gen.shading_inputs()
    for name, tone in (('plain', None), ('tone', Tonemapping())):
        pipe = object.__new__(MVEdit3DPipeline)
        pipe.tonemapping = tone
        got = pipe.make_shading_fun(lights, 0.2)(world_pos=None, albedo=albedo, world_normal=normal, fg_mask=fg)
        np.testing.assert_allclose(got.numpy(), PINS['shade_' + name], rtol=1e-5, atol=1e-6)


def test_load_init_nerf_matches_the_reference_method():
    from mvedit_b200.mvedit_3d_pipeline import MVEdit3DPipeline
    from mvedit_b200.tonemapping import Tonemapping
    field, poses, intr, l_world, l_cam = gen.init_nerf_inputs()
    for name, tone in (('plain', None), ('tone', Tonemapping())):
        pipe = object.__new__(MVEdit3DPipeline)
        pipe.nerf, pipe.tonemapping, pipe.normal_bg = field, tone, [0.5, 0.5, 1.0]
        im, al = pipe.load_init_nerf([None], None, poses, intr, 64, l_world, l_cam, 0.2, 2, 0.25, diff_size=48)
        np.testing.assert_allclose(im.numpy(), PINS['lin_images_' + name], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(al.numpy(), PINS['lin_alphas_' + name], rtol=1e-6, atol=1e-7)
