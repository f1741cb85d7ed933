"""Host logic of seam B2 (``mvedit_b200.adapter3d_mixin``) against THE REFERENCE'S OWN ``get_noise_pred{,_p1,_p2}``:
tests/golden/make_mixin_pins.py ran adapter3d_mixin.py:68-317 unmodified around toy networks; the product mixin, around the same toy
networks, must return the same noise predictions although it fuses the ``diff_bs`` chunks into one batch.  CPU."""
import importlib.util
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location('make_mixin_pins', os.path.join(HERE, 'golden', 'make_mixin_pins.py'))
gen = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gen)
PINS = np.load(os.path.join(HERE, 'golden', 'mixin_pins.npz'))


@pytest.mark.parametrize('case', ['plain', 'no_depth', 'extra_nets', 'reference'])
def test_product_mixin_matches_the_reference_methods(case):
    from mvedit_b200.adapter3d_mixin import Adapter3DMixin

    class Pipe(Adapter3DMixin):
        pass
    pipe = Pipe()
    pipe.unet = gen.ToyUNet()
    c = gen.cases()[case]
    pipe.controlnet = SimpleNamespace(nets=gen.toy_nets(c['nets']))
    got = gen.run(pipe, c)
    for k, v in got.items():
        ref = torch.from_numpy(PINS['%s_%s' % (case, k)])
        assert v.shape == ref.shape, (k, v.shape, ref.shape)
        torch.testing.assert_close(v, ref, rtol=1e-5, atol=1e-5)
    assert (got['p2'] - got['p1']).abs().max() > 1e-3                       # the second pass (tile ControlNet on the renders) changes the prediction
