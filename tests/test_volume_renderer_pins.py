"""``oracle/nerf_oracle.OracleDecoder`` (forward in both modes, ``update_extra_state``) against THE REFERENCE'S OWN ``VolumeRenderer`` code:
tests/golden/make_volume_renderer_pins.py ran base_volume_renderer.py:17-343 unmodified (cut out by AST) on the C restatement of the
ray-marching kernels and the plain-torch field, with the same noise.  CPU."""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location('make_volume_renderer_pins', os.path.join(HERE, 'golden', 'make_volume_renderer_pins.py'))
gen = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gen)
PINS = np.load(os.path.join(HERE, 'golden', 'volume_renderer_pins.npz'))


def test_oracle_decoder_matches_the_reference_volume_renderer():
    ro, rd, bitfield, dec, noise, dt_gamma = gen.inputs()
    grid = torch.zeros(1, gen.GRID ** 3, dtype=torch.float16)
    bits = torch.zeros(1, gen.GRID ** 3 // 8, dtype=torch.uint8)
    dec.update_extra_state(None, grid, bits, 0, density_thresh=0.1, noise=noise['grid'])
    np.testing.assert_allclose(grid.float().numpy(), PINS['ue_grid'].astype(np.float32), rtol=1e-3, atol=1e-3)
    assert (bits.numpy() == PINS['ue_bits']).all()
    dec.train(True)
    with torch.no_grad():
        r = dec(ro, rd, None, bitfield, gen.GRID, dt_gamma=dt_gamma, perturb=True, noises=noise['march'])
    assert (r['rays'][0].numpy() == PINS['tr_rays']).all()                      # same culling decisions, same re-indexing
    np.testing.assert_array_equal(r['ts'][0].numpy(), PINS['tr_ts'])
    for k, key in (('weights', 'tr_weights'),):
        np.testing.assert_allclose(r[k].numpy(), PINS[key], rtol=1e-5, atol=1e-7)
    for k in ('weights_sum', 'depth', 'image'):
        np.testing.assert_allclose(r[k][0].numpy(), PINS['tr_' + k], rtol=1e-5, atol=1e-6)
    dec.train(False)
    with torch.no_grad():
        e = dec(ro, rd, [None], bitfield, gen.GRID, dt_gamma=torch.tensor([dt_gamma]), perturb=False)
    for k in ('weights_sum', 'depth', 'image'):
        np.testing.assert_allclose(e[k][0].numpy(), PINS['ev_' + k], rtol=1e-5, atol=1e-6)
    assert 0.05 < float(e['weights_sum'][0].mean()) < 0.9


def test_oracle_nerf_render_matches_the_reference_method():
    """``OracleNeRF.render`` against the reference's own ``BaseNeRF.render`` (base_nerf.py:489-556) driving the same decoder."""
    from oracle import nerf_oracle as no
    _, _, bitfield, dec, _, _ = gen.inputs()
    poses, intr = gen.render_cameras()
    nerf = no.OracleNeRF(dec, grid_size=gen.GRID)
    with torch.no_grad():
        rgba, depth, normal, normal_fg = nerf.render(bitfield, gen.SIZE, gen.SIZE, intr[None], poses[None], cfg=dict(dt_gamma_scale=0.5))
    for got, key in ((rgba, 'rn_rgba'), (depth, 'rn_depth'), (normal, 'rn_normal'), (normal_fg, 'rn_normal_fg')):
        np.testing.assert_allclose(got.numpy(), PINS[key], rtol=1e-5, atol=1e-6)
    assert rgba.shape == (1, gen.VIEWS, gen.SIZE, gen.SIZE, 4)
