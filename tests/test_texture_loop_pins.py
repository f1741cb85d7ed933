"""The loop bodies of the two texture pipelines against THE REFERENCE'S OWN ``MVEditTexturePipeline.__call__`` and
``MVEditTextureSuperResPipeline.__call__``: tests/golden/make_texture_loop_pins.py ran them unmodified around toy components and
recorded every hand-over to ``bake_multiview`` / ``texture_optim`` / ``bake_xyz_shading_fun`` / ``get_cam_weights_uv`` (and the texture
the super-resolution pipeline returns).  The product's ``__call__``s run around the same toys and must hand over the same things: targets,
dense camera weights, cameras after re-ordering / pruning, sizes, weights -- optimisation only, 1-pass, 2-pass with reference pairs and
weighted pruning, from noise, and the super-resolution variant with and without regulariser views / an input texture.  CPU."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from mvedit_b200 import mvedit_texture_pipeline as TP
from mvedit_b200.schedulers import EulerAncestralScheduler

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location('make_texture_loop_pins', os.path.join(HERE, 'golden', 'make_texture_loop_pins.py'))
gen = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gen)
PINS = np.load(os.path.join(HERE, 'golden', 'texture_loop_pins.npz'))


class AdamLike(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, **kw):
        super().__init__(params, lr=lr)


@pytest.mark.parametrize('case', list(gen.CASES))
def test_texture_pipelines_hand_over_what_the_reference_loops_hand_over(case, monkeypatch):
    monkeypatch.setattr(TP, 'FusedAdam', AdamLike)
    log = []
    renderer = gen.ToyTexRenderer(log)
    poses, intr, embeds = gen.inputs()
    kind, kw = gen.call_kwargs(case, poses, intr, embeds)
    cls = TP.MVEditTexturePipeline if kind == 'texture' else TP.MVEditTextureSuperResPipeline
    pipe = cls(gen.L.ToyVAE(), None, None, gen.L.ToyUNet(), gen.L.mixin_gen.toy_nets(2), EulerAncestralScheduler(), renderer.field, renderer)
    pipe.load_init_mesh = gen.toy_load_init_mesh(renderer)
    pipe.texture_optim = lambda *a, **k: gen.record_texture_optim(log, *a, **k)
    torch.manual_seed(4321)
    res = pipe(prompt_embeds=embeds.clone(), **kw)
    if kind == 'texture':
        mesh, state = res
        assert mesh is not None and state is not None, 'the run raised inside __call__ (traceback printed above)'
    else:
        assert res is not None, 'the run raised inside __call__ (traceback printed above)'
        log.append(dict(kind=4.0, maps=res.albedo[..., :3][None].clone()))
    assert len(log) == int(PINS[case + '_calls']), [r['kind'] for r in log]
    if kw.get('ip_adapter') is not None:
        assert len(kw['ip_adapter'].seen) == 1
        np.testing.assert_allclose(torch.nn.functional.avg_pool2d(kw['ip_adapter'].seen[0], 16).numpy(), PINS[case + '_ipa_images'], rtol=1e-4, atol=1e-4)
    for i, rec in enumerate(log):
        assert rec['kind'] == float(PINS['%s_%d_kind' % (case, i)]), (i, rec['kind'])
        for k, v in rec.items():
            key = '%s_%d_%s' % (case, i, k)
            if k in ('maps', 'weights'):
                assert tuple(v.shape) == tuple(PINS[key + '_shape']), (key, v.shape)
                x = v.reshape(-1, *v.shape[-3:]).permute(0, 3, 1, 2).float()
                pooled = torch.nn.functional.avg_pool2d(x, 16 if x.shape[-1] >= 128 else 4).numpy()
                np.testing.assert_allclose(pooled, PINS[key + '_pooled'], rtol=1e-4, atol=1.5e-2, err_msg=key)
                np.testing.assert_allclose(x.flatten(2).std(dim=2).numpy(), PINS[key + '_std'], rtol=1e-4, atol=1.5e-2, err_msg=key)
            elif torch.is_tensor(v):
                np.testing.assert_allclose(v.numpy(), PINS[key], rtol=1e-5, atol=1e-6, err_msg=key)
            else:
                assert float(v) == pytest.approx(float(PINS[key]), rel=1e-6), key
        assert {k for k in rec if k not in ('maps', 'weights')} == {f[len('%s_%d_' % (case, i)):] for f in PINS.files
                                                                    if f.startswith('%s_%d_' % (case, i)) and not f.endswith(('_pooled', '_std', '_shape'))}
