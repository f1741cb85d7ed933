"""``MVEdit3DPipeline.enable_normals`` / ``load_depths`` (the inputs of the image-to-3D targets) against THE REFERENCE'S OWN methods:
tests/golden/make_pipeline_input_pins.py ran mvedit_3d_pipeline.py:232-306 unmodified, with the reference's Tonemapping module and toy
stand-ins for the normal model / enhancer shared with this test.  CPU (pure host-side torch: these run once before the loop)."""
import importlib.util
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location('make_pipeline_input_pins', os.path.join(HERE, 'golden', 'make_pipeline_input_pins.py'))
gen = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gen)
PINS = np.load(os.path.join(HERE, 'golden', 'pipeline_input_pins.npz'))


def _pipe(tone=None, normal_model=gen.toy_normal_model):
    from mvedit_b200.mvedit_3d_pipeline import MVEdit3DPipeline
    pipe = object.__new__(MVEdit3DPipeline)
    pipe.unet = SimpleNamespace(device='cpu', dtype=torch.float32)
    pipe.normal_model, pipe.image_enhancer, pipe.tonemapping, pipe.normal_bg, pipe.bg_color = normal_model, gen.toy_enhancer, tone, [0.5, 0.5, 1.0], 1.0
    return pipe


@pytest.mark.parametrize('name', ['plain', 'tone'])
def test_enable_normals_matches_the_reference_method(name):
    from mvedit_b200.tonemapping import Tonemapping
    images, masks, lights, normals, _ = gen.inputs()
    im, nm = _pipe(Tonemapping() if name == 'tone' else None).enable_normals(images.clone(), masks, lights, 0.2, normals=normals)
    torch.testing.assert_close(nm, torch.from_numpy(PINS[name + '_normals']), rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(im, torch.from_numpy(PINS[name + '_images']), rtol=1e-5, atol=2e-6)
    assert (im - images).abs().max() > 0.05                                     # the inputs were re-shaded


def test_enable_normals_without_a_normal_model_needs_every_map():
    images, masks, lights, normals, _ = gen.inputs()
    pipe = _pipe(normal_model=None)
    with pytest.raises(NotImplementedError):
        pipe.enable_normals(images, masks, lights, 0.2, normals=normals)         # two views carry no map
    given = [normals[1]] * 4
    im, nm = pipe.enable_normals(images, masks, lights, 0.2, normals=given)
    assert nm.shape == images.shape and torch.isfinite(im).all()


def test_load_depths_matches_the_reference_method():
    *_, depths = gen.inputs()
    torch.testing.assert_close(_pipe().load_depths(depths, diff_size=gen.S), torch.from_numpy(PINS['depths']), rtol=1e-6, atol=1e-7)
