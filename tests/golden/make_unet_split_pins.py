"""Generates tests/golden/unet_split_pins.npz by RUNNING the reference's own ``unet_enc`` / ``unet_dec`` (lib/models/architecture/diffusers.py:57-164,
the split of the UNet forward that the 2-pass denoiser relies on) unmodified on a diffusers-SHAPED object whose blocks are the block functions of
oracle/unet_oracle.py (diffusers itself is absent: the blocks' internals stay "parity unpinned", what is pinned is the reference's carry logic --
which residuals are handed to which up block, where the ControlNet residuals enter, the reference-pair kwargs passing through).

Run:  python tests/golden/make_unet_split_pins.py      (CPU, seconds)
"""
import ast
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'unet_split_pins.npz')


def inputs():
    from oracle import unet_oracle as uo
    cfg = uo.TINY
    sd = uo.random_unet_state_dict(cfg, 0)
    g = torch.Generator().manual_seed(1)
    N = 4
    sample, ctx = torch.randn(N, 4, 16, 16, generator=g), torch.randn(N, 7, cfg.cross_attention_dim, generator=g)
    with torch.no_grad():
        e, r, s = uo.unet_enc(sd, cfg, sample, 500.0, ctx)
    down = [torch.randn(x.shape, generator=g) * 0.3 for x in r]
    mid = torch.randn(s.shape, generator=g) * 0.3
    return cfg, sd, sample, ctx, down, mid


def diffusers_shaped(sd, cfg):
    """An object with the attributes ``unet_enc`` / ``unet_dec`` touch, in diffusers' block calling conventions (CrossAttnDownBlock2D /
    DownBlock2D return (sample, res_samples); up blocks consume ``res_hidden_states_tuple`` from its END, one entry per resnet)."""
    from oracle import unet_oracle as uo
    nlev = len(cfg.block_out_channels)
    n_of = lambda kw: (kw or {}).get('num_cross_attn_imgs', 1)

    class Down:
        def __init__(self, i):
            self.i, self.has_cross_attention = i, cfg.attn_levels[i]

        def __call__(self, hidden_states, temb, encoder_hidden_states=None, cross_attention_kwargs=None):
            res = ()
            for j in range(cfg.layers_per_block):
                hidden_states = uo.resnet(sd, f'down_blocks.{self.i}.resnets.{j}', hidden_states, temb, cfg)
                if self.has_cross_attention:
                    hidden_states = uo.transformer(sd, f'down_blocks.{self.i}.attentions.{j}', hidden_states, encoder_hidden_states, cfg.num_heads[self.i],
                                                   cfg, n_of(cross_attention_kwargs))
                res += (hidden_states,)
            if self.i < nlev - 1:
                hidden_states = uo._c(sd, f'down_blocks.{self.i}.downsamplers.0.conv', hidden_states, stride=2)
                res += (hidden_states,)
            return hidden_states, res

    class Mid:
        has_cross_attention = True

        def __call__(self, sample, emb, encoder_hidden_states=None, cross_attention_kwargs=None):
            return uo.mid_forward(sd, cfg, sample, emb, encoder_hidden_states, n_of(cross_attention_kwargs))

    class Up:
        def __init__(self, i):
            self.i, self.has_cross_attention = i, list(reversed(cfg.attn_levels))[i]
            self.resnets = [None] * (cfg.layers_per_block + 1)

        def __call__(self, hidden_states, temb, res_hidden_states_tuple, encoder_hidden_states=None, cross_attention_kwargs=None):
            heads = list(reversed(cfg.num_heads))[self.i]
            for j in range(len(self.resnets)):
                hidden_states = torch.cat([hidden_states, res_hidden_states_tuple[-1]], dim=1)
                res_hidden_states_tuple = res_hidden_states_tuple[:-1]
                hidden_states = uo.resnet(sd, f'up_blocks.{self.i}.resnets.{j}', hidden_states, temb, cfg)
                if self.has_cross_attention:
                    hidden_states = uo.transformer(sd, f'up_blocks.{self.i}.attentions.{j}', hidden_states, encoder_hidden_states, heads, cfg,
                                                   n_of(cross_attention_kwargs))
            if self.i < nlev - 1:
                hidden_states = uo._c(sd, f'up_blocks.{self.i}.upsamplers.0.conv', F.interpolate(hidden_states, scale_factor=2.0, mode='nearest'))
            return hidden_states
    return types.SimpleNamespace(
        config=types.SimpleNamespace(center_input_sample=False),
        get_time_embed=lambda sample, timestep: uo.timestep_embedding(torch.as_tensor(timestep).reshape(-1).expand(sample.shape[0]),
                                                                      cfg.block_out_channels[0]).to(sample.dtype),
        time_embedding=lambda e: uo._l(sd, 'time_embedding.linear_2', F.silu(uo._l(sd, 'time_embedding.linear_1', e))),
        get_aug_embed=lambda **k: None, time_embed_act=None,
        process_encoder_hidden_states=lambda encoder_hidden_states, added_cond_kwargs: encoder_hidden_states,
        conv_in=lambda x: uo._c(sd, 'conv_in', x), down_blocks=[Down(i) for i in range(nlev)], mid_block=Mid(), up_blocks=[Up(i) for i in range(nlev)],
        conv_norm_out=lambda x: uo._gn(sd, 'conv_norm_out', x, cfg.norm_groups, 1e-5), conv_act=F.silu, conv_out=lambda x: uo._c(sd, 'conv_out', x))


def main():
    tree = ast.parse(open(os.path.join(REF, 'lib/models/architecture/diffusers.py')).read())
    env = dict(torch=torch)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ('unet_enc', 'unet_dec'):
            node.returns = None
            for a in node.args.args:
                a.annotation = None
            mod = ast.Module(body=[node], type_ignores=[])
            ast.fix_missing_locations(mod)
            exec(compile(mod, 'diffusers.py', 'exec'), env)
    cfg, sd, sample, ctx, down, mid = inputs()
    unet = diffusers_shaped(sd, cfg)
    out = {}
    with torch.no_grad():
        for name, kw, x, c in (('plain', None, sample, ctx), ('pairs', dict(num_cross_attn_imgs=2), sample, ctx)):
            emb, res, s = env['unet_enc'](unet, x, 500.0, c, cross_attention_kwargs=kw)
            out[name + '_emb'], out[name + '_s'] = emb.numpy(), s.numpy()
            out[name + '_n_res'] = np.array(len(res))
            for k, r in enumerate(res):
                out['%s_res%d' % (name, k)] = r.mean(dim=(2, 3)).numpy()          # per image and channel: the fixture stays small
            out[name + '_dec'] = env['unet_dec'](unet, emb, res, s, c, cross_attention_kwargs=kw).numpy()
            out[name + '_dec_cn'] = env['unet_dec'](unet, emb, res, s, c, cross_attention_kwargs=kw, down_block_additional_residuals=down,
                                                    mid_block_additional_residual=mid).numpy()
            out[name + '_dec_down_only'] = env['unet_dec'](unet, emb, res, s, c, cross_attention_kwargs=kw, down_block_additional_residuals=down).numpy()
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, sum(v.nbytes for v in out.values()) // 1024, 'KiB')


if __name__ == '__main__':
    main()
