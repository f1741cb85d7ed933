"""``oracle/unet_oracle.unet_enc`` / ``unet_dec`` (what the GPU tests hold the product's ``UNet.enc`` / ``UNet.dec`` to) against THE
REFERENCE'S OWN ``unet_enc`` / ``unet_dec`` (tests/golden/make_unet_split_pins.py ran lib/models/architecture/diffusers.py:57-164 unmodified
on a diffusers-shaped object made of the oracle's blocks): the residual hand-over between the halves, the ControlNet residual entry points
(only when BOTH residual kinds are given, as the reference decides) and the reference-pair kwargs.  CPU."""
import importlib.util
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location('make_unet_split_pins', os.path.join(HERE, 'golden', 'make_unet_split_pins.py'))
gen = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gen)
PINS = np.load(os.path.join(HERE, 'golden', 'unet_split_pins.npz'))


@pytest.mark.parametrize('name,kw', [('plain', None), ('pairs', dict(num_cross_attn_imgs=2))])
def test_oracle_split_matches_the_reference_functions(name, kw):
    from oracle import unet_oracle as uo
    cfg, sd, sample, ctx, down, mid = gen.inputs()
    close = lambda a, k: np.testing.assert_allclose(a.numpy(), PINS[k], rtol=1e-4, atol=1e-5, err_msg=k)
    with torch.no_grad():
        emb, res, s = uo.unet_enc(sd, cfg, sample, 500.0, ctx, cross_attention_kwargs=kw)
        assert len(res) == int(PINS[name + '_n_res'])
        close(emb, name + '_emb'); close(s, name + '_s')
        for k, r in enumerate(res):
            close(r.mean(dim=(2, 3)), '%s_res%d' % (name, k))
        close(uo.unet_dec(sd, cfg, emb, res, s, ctx, kw), name + '_dec')
        close(uo.unet_dec(sd, cfg, emb, res, s, ctx, kw, down, mid), name + '_dec_cn')
        close(uo.unet_dec(sd, cfg, emb, res, s, ctx, kw, down, None), name + '_dec_down_only')
    assert np.abs(PINS[name + '_dec_cn'] - PINS[name + '_dec']).max() > 1e-2
    np.testing.assert_array_equal(PINS[name + '_dec_down_only'], PINS[name + '_dec'])       # residuals are ignored unless both kinds are given
