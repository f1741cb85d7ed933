"""Generates tests/golden/signatures.json: argument names and default values of the reference's call surface (seams B1-B5 of SURVEY §8b),
read from its source by AST.  The test holds the product's mirrors to them: same names in the same order, same defaults (a product
function may append arguments of its own after the reference's).

Run:  python tests/golden/make_signature_pins.py
"""
import ast
import json
import os

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'signatures.json')
TARGETS = {
    'lib/pipelines/mvedit_3d_pipeline.py': ['MVEdit3DPipeline.__init__', 'MVEdit3DPipeline.__call__', 'MVEdit3DPipeline.nerf_optim', 'MVEdit3DPipeline.mesh_optim',
                                            'MVEdit3DPipeline.load_init_nerf', 'MVEdit3DPipeline.load_init_images', 'MVEdit3DPipeline.load_depths',
                                            'MVEdit3DPipeline.enable_normals', 'MVEdit3DPipeline.load_cond_images',
                                            'default_lr_multiplier', 'default_max_num_views', 'default_render_size_p', 'default_lr_schedule',
                                            'default_patch_rgb_weight', 'default_patch_normal_weight', 'default_entropy_weight', 'default_normal_reg_weight'],
    'lib/pipelines/mvedit_texture_pipeline.py': ['MVEditTexturePipeline.__call__', 'MVEditTexturePipeline.texture_optim', 'camera_dense_weighting',
                                                 'default_patch_rgb_weight', 'default_max_num_views'],
    'lib/pipelines/mvedit_texture_superres_pipeline.py': ['MVEditTextureSuperResPipeline.__call__'],
    'lib/pipelines/adapter3d_mixin.py': ['Adapter3DMixin.get_noise_pred', 'Adapter3DMixin.get_noise_pred_p1', 'Adapter3DMixin.get_noise_pred_p2',
                                         'Adapter3DMixin.load_init_mesh'],
    'lib/pipelines/utils.py': ['init_tet', 'get_camera_dists', 'prune_cameras', 'highpass', 'join_prompts'],
    'lib/models/decoders/mesh_renderer/base_mesh_renderer.py': ['MeshRenderer.__init__', 'MeshRenderer.forward', 'MeshRenderer.bake_xyz_shading_fun',
                                                                'MeshRenderer.bake_multiview', 'MeshRenderer.get_cam_weights_uv', 'DMTet.__call__',
                                                                'normal_consistency', 'laplacian_smooth_loss', 'compute_edge_to_face_mapping'],
    'lib/models/autoencoders/base_nerf.py': ['BaseNeRF.render', 'BaseNeRF.ray_sample', 'BaseNeRF.get_raybatch_inds'],
    'lib/models/decoders/base_volume_renderer.py': ['VolumeRenderer.forward', 'VolumeRenderer.update_extra_state'],
    'lib/models/decoders/ingp_decoder.py': ['iNGPDecoder.__init__', 'iNGPDecoder.point_decode'],
    'lib/models/decoders/mesh_renderer/mesh_utils.py': ['Mesh.__init__', 'Mesh.load', 'Mesh.auto_normal', 'Mesh.to', 'Mesh.write'],
    'lib/ops/edge_dilation.py': ['edge_dilation'],
    'lib/models/decoders/tonemapping.py': ['Tonemapping.__init__', 'Tonemapping.lut', 'Tonemapping.inverse_lut', 'Tonemapping.smooth_forward'],
    'lib/core/utils/camera_utils.py': ['light_sampling'],
    'lib/ops/raymarching/raymarching.py': ['near_far_from_aabb', 'march_rays_train', 'composite_rays_train', 'march_rays', 'composite_rays',
                                           'morton3D', 'morton3D_invert', 'packbits'],
}


def describe(fn):
    a = fn.args
    names = [x.arg for x in a.posonlyargs + a.args]
    defaults = [None] * (len(names) - len(a.defaults)) + [ast.unparse(d) for d in a.defaults]
    out = [[n, d] for n, d in zip(names, defaults)]
    if a.vararg:
        out.append(['*' + a.vararg.arg, None])
    out += [[x.arg, None if d is None else ast.unparse(d)] for x, d in zip(a.kwonlyargs, a.kw_defaults)]
    if a.kwarg:
        out.append(['**' + a.kwarg.arg, None])
    return out


def main():
    out = {}
    for rel, names in TARGETS.items():
        tree = ast.parse(open(os.path.join(REF, rel)).read())
        scope = {n.name: n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef))}
        assign = {t.id: v for n in tree.body if isinstance(n, ast.Assign) for t, v in [(n.targets[0], n.value)] if isinstance(t, ast.Name)}
        for name in names:
            if '.' in name:
                cls, meth = name.split('.')
                fn = next(m for m in scope[cls].body if isinstance(m, ast.FunctionDef) and m.name == meth)
            elif name in scope:
                fn = scope[name]
            else:                                 # e.g. ``near_far_from_aabb = _near_far_from_aabb.apply``: the autograd Function's forward minus ctx
                target = ast.unparse(assign[name]).split('.')[0]
                fn = next(m for m in scope[target].body if isinstance(m, ast.FunctionDef) and m.name == 'forward')
            sig = describe(fn)
            if sig and sig[0][0] in ('self', 'ctx', 'cls'):
                sig = sig[1:]
            out[rel + '::' + name] = sig
    json.dump(out, open(OUT, 'w'), indent=1)
    print('wrote', OUT, len(out), 'signatures')


if __name__ == '__main__':
    main()
