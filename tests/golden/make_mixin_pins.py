"""Generates tests/golden/mixin_pins.npz by RUNNING the reference's own ``Adapter3DMixin.get_noise_pred`` / ``get_noise_pred_p1`` /
``get_noise_pred_p2`` (lib/pipelines/adapter3d_mixin.py:68-317), cut out by AST and executed unmodified, around TOY networks: what is
pinned is the host logic of seam B2 -- chunk loop, reference-image reshapes (``num_cross_attn_imgs``, zero-interleaved ControlNet
residuals), which ControlNets run in which pass (``controlnet_skip``), residual sums of the second pass, CFG / adapter-scale combination --
not the networks (diffusers is absent; their kernels have their own oracles).  The toy UNet / ControlNets are deterministic functions
that depend on every input, accept both the reference's keyword call convention and the product's positional one, and mix the two
images of a reference pair when ``num_cross_attn_imgs=2``.

Run:  python tests/golden/make_mixin_pins.py      (CPU, seconds)
"""
import ast
import os
import sys
from copy import copy

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'mixin_pins.npz')


class ToyControlNet:
    def __init__(self, seed):
        g = torch.Generator().manual_seed(seed)
        self.a = torch.rand(3, generator=g) + 0.5
        self.b = torch.rand(3, generator=g) - 0.5

    def __call__(self, sample, t, encoder_hidden_states=None, controlnet_cond=None, conditioning_scale=1.0, accumulate=None, cond_repeat=1, **kw):
        pe = encoder_hidden_states.mean(dim=(1, 2))[:, None, None, None]
        cond = torch.nn.functional.adaptive_avg_pool2d(controlnet_cond.float(), sample.shape[-1]).mean(dim=1, keepdim=True)
        base = sample.mean(dim=1, keepdim=True)
        down = [(base * self.a[k] + cond * self.b[k] + pe + float(t) * 1e-3 * (k + 1)) * conditioning_scale for k in range(3)]
        mid = (base * self.b[0] - cond * self.a[1] + 2 * pe) * conditioning_scale
        if accumulate is not None:
            down, mid = [x + y for x, y in zip(down, accumulate[0])], mid + accumulate[1]
        return down, mid


class ToyUNet:
    config = None

    def enc(self, x, t, encoder_hidden_states=None, cross_attention_kwargs=None, added_cond_kwargs=None):
        pe = encoder_hidden_states.mean(dim=(1, 2))[:, None, None, None]
        s = x * 0.5 + pe
        if cross_attention_kwargs is not None and cross_attention_kwargs.get('num_cross_attn_imgs', 1) == 2:
            pair = s.reshape(-1, 2, *s.shape[1:])
            s = (pair + 0.25 * pair.flip(1)).reshape(s.shape)                 # the two images of a (reference, view) pair see each other
        res = (x.mean(dim=1, keepdim=True) * 1.1, x.mean(dim=1, keepdim=True) * 0.9 + pe, x[:, :1] * 0.3)
        return torch.full((x.shape[0], 1), float(t) * 1e-3), res, s

    def dec(self, emb, res, s, encoder_hidden_states=None, cross_attention_kwargs=None, down_block_additional_residuals=None,
            mid_block_additional_residual=None):
        out = s + emb[:, :, None, None]
        for k, r in enumerate(res):
            out = out + (k + 1) * 0.1 * (r if down_block_additional_residuals is None else r + down_block_additional_residuals[k])
        if mid_block_additional_residual is not None:
            out = out - 0.7 * mid_block_additional_residual
        if cross_attention_kwargs is not None and cross_attention_kwargs.get('num_cross_attn_imgs', 1) == 2:
            pair = out.reshape(-1, 2, *out.shape[1:])
            out = (pair - 0.1 * pair.flip(1)).reshape(out.shape)
        return out

    def __call__(self, x, t, encoder_hidden_states=None, cross_attention_kwargs=None, down_block_additional_residuals=None,
                 mid_block_additional_residual=None, added_cond_kwargs=None, return_dict=True):
        out = self.dec(*self.enc(x, t, encoder_hidden_states, cross_attention_kwargs), encoder_hidden_states=encoder_hidden_states,
                       cross_attention_kwargs=cross_attention_kwargs, down_block_additional_residuals=down_block_additional_residuals,
                       mid_block_additional_residual=mid_block_additional_residual)
        return (out,) if return_dict is False else out


def toy_nets(n):
    return [ToyControlNet(k + 1) for k in range(n)]


def cases():
    """name -> kwargs of the three calls; inputs drawn from one seeded generator.  N views, CFG: [uncond ; cond] chunks of <= 3."""
    g = torch.Generator().manual_seed(0)
    N, L, S = 5, 4, 32
    r = lambda *s: torch.randn(*s, generator=g)
    lat, pe = r(2 * N, 4, L, L), r(2 * N, 7, 6)
    ci, cd, ex = torch.rand(2 * N, 3, S, S, generator=g), torch.rand(2 * N, 3, S, S, generator=g), torch.rand(2 * N, 3, S, S, generator=g)
    sp = lambda x: list(x.split(3, dim=0))
    lat_ref = r(2 * N, 4, 2 * L, L)                                         # every latent is the (reference image ; view) pair
    return dict(
        plain=dict(nets=2, lat=sp(lat), pe=sp(pe), ci=sp(ci), cd=sp(cd), extra=None, t=500.0, tw=0.8, dw=0.6, g=7.0, adapter=None),
        no_depth=dict(nets=1, lat=sp(lat), pe=sp(pe), ci=sp(ci), cd=None, extra=None, t=321.0, tw=1.0, dw=None, g=3.0, adapter=None),
        extra_nets=dict(nets=3, lat=sp(lat), pe=sp(pe), ci=sp(ci), cd=sp(cd), extra=[sp(ex)], t=77.0, tw=0.5, dw=1.0, g=7.0, adapter=0.4),
        reference=dict(nets=2, lat=sp(lat_ref), pe=sp(pe), ci=sp(ci), cd=sp(cd), extra=None, t=650.0, tw=0.9, dw=1.0, g=5.0, adapter=None))


def run(obj, c):
    """The three entry points on one case -> dict of outputs (works for the reference functions bound to ``obj`` and for the product mixin)."""
    out = {}
    if c['cd'] is not None:                      # the reference's one-pass entry always hands the depth ControlNet an image
        out['one_pass'] = obj.get_noise_pred(c['lat'], c['pe'], c['ci'], c['cd'], c['t'], c['tw'], c['dw'], c['g'],
                                             extra_control_batches=c['extra'], adapter_scale=c['adapter'])
    n1, da, dk = obj.get_noise_pred_p1(c['lat'], c['pe'], c['t'], c['g'], c['cd'], c['dw'], extra_control_batches=c['extra'])
    out['p1'] = n1
    out['p2'] = obj.get_noise_pred_p2(c['lat'], c['pe'], da, dk, c['t'], c['g'], c['ci'], c['tw'], c['cd'], c['dw'], adapter_scale=c['adapter'])
    return out


def main():
    import types
    tree = ast.parse(open(os.path.join(REF, 'lib/pipelines/adapter3d_mixin.py')).read())
    want = ('get_noise_pred', 'get_noise_pred_p1', 'get_noise_pred_p2')
    nodes = [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name in want]

    class MultiControlNetModel:                  # diffusers' semantics: every net on its own condition image, residuals summed
        def __init__(self, nets):
            self.nets = list(nets)

        def __call__(self, sample, t, encoder_hidden_states=None, controlnet_cond=None, conditioning_scale=None, guess_mode=False,
                     added_cond_kwargs=None, return_dict=True):
            acc = None
            for net, cond, sc in zip(self.nets, controlnet_cond, conditioning_scale):
                acc = net(sample, t, encoder_hidden_states=encoder_hidden_states, controlnet_cond=cond, conditioning_scale=sc, accumulate=acc)
            return acc

    env = dict(torch=torch, copy=copy, MultiControlNetModel=MultiControlNetModel,
               unet_enc=lambda unet, *a, **k: unet.enc(*a, **k), unet_dec=lambda unet, *a, **k: unet.dec(*a, **k))
    mod = ast.Module(body=nodes, type_ignores=[])
    ast.fix_missing_locations(mod)
    exec(compile(mod, 'adapter3d_mixin.py', 'exec'), env)
    out = {}
    for name, c in cases().items():
        obj = types.SimpleNamespace(unet=ToyUNet(), controlnet=MultiControlNetModel(toy_nets(c['nets'])))
        for n in want:
            setattr(obj, n, types.MethodType(env[n], obj))
        for k, v in run(obj, c).items():
            out['%s_%s' % (name, k)] = v.numpy()
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, sorted(out))


if __name__ == '__main__':
    main()
