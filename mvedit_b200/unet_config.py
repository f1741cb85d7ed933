"""Architecture description + random initialisation (diffusers state-dict key names) for the SD-1.5 UNet / ControlNet v1.1
(SURVEY.md Appendix A).  Checkpoints are not available offline (BASELINE.json: "random-init SD/ControlNet weights"), so benchmarks
and smoke tests build weights of the exact published shapes here, directly on the target device; a real diffusers state dict with the
same keys loads through the same ``UNet(state_dict, cfg)`` constructor.
"""
import math
from dataclasses import dataclass
from typing import Tuple

import torch


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    num_heads: Tuple[int, ...] = (8, 8, 8, 8)
    cross_attention_dim: int = 768
    attn_levels: Tuple[bool, ...] = (True, True, True, False)
    cond_embed_channels: Tuple[int, ...] = (16, 32, 96, 256)
    norm_groups: int = 32

    @property
    def time_embed_dim(self):
        return self.block_out_channels[0] * 4


SD15 = UNetConfig()
TINY = UNetConfig(block_out_channels=(64, 128, 128, 128), num_heads=(1, 2, 2, 2), cross_attention_dim=64)


class _Init:
    def __init__(self, seed, device):
        self.g = torch.Generator(device=device).manual_seed(seed)
        self.dev = device
        self.sd = {}

    def conv(self, name, cin, cout, k, scale=1.0):
        self.sd[name + '.weight'] = torch.randn(cout, cin, k, k, generator=self.g, device=self.dev) * (scale / math.sqrt(cin * k * k))
        self.sd[name + '.bias'] = torch.randn(cout, generator=self.g, device=self.dev) * 0.02

    def lin(self, name, cin, cout, bias=True):
        self.sd[name + '.weight'] = torch.randn(cout, cin, generator=self.g, device=self.dev) / math.sqrt(cin)
        if bias:
            self.sd[name + '.bias'] = torch.randn(cout, generator=self.g, device=self.dev) * 0.02

    def norm(self, name, c):
        self.sd[name + '.weight'] = 1 + 0.1 * torch.randn(c, generator=self.g, device=self.dev)
        self.sd[name + '.bias'] = 0.1 * torch.randn(c, generator=self.g, device=self.dev)

    def resnet(self, p, cin, cout, temb):
        self.norm(p + '.norm1', cin); self.conv(p + '.conv1', cin, cout, 3)
        self.lin(p + '.time_emb_proj', temb, cout)
        self.norm(p + '.norm2', cout); self.conv(p + '.conv2', cout, cout, 3)
        if cin != cout:
            self.conv(p + '.conv_shortcut', cin, cout, 1)

    def transformer(self, p, c, cross):
        self.norm(p + '.norm', c); self.conv(p + '.proj_in', c, c, 1)
        b = p + '.transformer_blocks.0'
        for n in ('norm1', 'norm2', 'norm3'):
            self.norm(b + '.' + n, c)
        for a, kv in (('attn1', c), ('attn2', cross)):
            self.lin(b + f'.{a}.to_q', c, c, bias=False); self.lin(b + f'.{a}.to_k', kv, c, bias=False)
            self.lin(b + f'.{a}.to_v', kv, c, bias=False); self.lin(b + f'.{a}.to_out.0', c, c)
        self.lin(b + '.ff.net.0.proj', c, 8 * c); self.lin(b + '.ff.net.2', 4 * c, c)
        self.conv(p + '.proj_out', c, c, 1)

    def encoder(self, cfg):
        boc, temb = cfg.block_out_channels, cfg.time_embed_dim
        self.conv('conv_in', cfg.in_channels, boc[0], 3)
        self.lin('time_embedding.linear_1', boc[0], temb); self.lin('time_embedding.linear_2', temb, temb)
        cin = boc[0]
        for i, cout in enumerate(boc):
            for j in range(cfg.layers_per_block):
                self.resnet(f'down_blocks.{i}.resnets.{j}', cin if j == 0 else cout, cout, temb)
                if cfg.attn_levels[i]:
                    self.transformer(f'down_blocks.{i}.attentions.{j}', cout, cfg.cross_attention_dim)
            if i < len(boc) - 1:
                self.conv(f'down_blocks.{i}.downsamplers.0.conv', cout, cout, 3)
            cin = cout
        c = boc[-1]
        self.resnet('mid_block.resnets.0', c, c, temb); self.transformer('mid_block.attentions.0', c, cfg.cross_attention_dim)
        self.resnet('mid_block.resnets.1', c, c, temb)


def skip_channels(cfg):
    ch = [cfg.block_out_channels[0]]
    for i, c in enumerate(cfg.block_out_channels):
        ch += [c] * cfg.layers_per_block
        if i < len(cfg.block_out_channels) - 1:
            ch.append(c)
    return ch


def random_unet_state_dict(cfg=SD15, seed=0, device='cuda'):
    it = _Init(seed, device)
    it.encoder(cfg)
    boc, temb = cfg.block_out_channels, cfg.time_embed_dim
    skips = skip_channels(cfg)
    rev, rev_attn = list(reversed(boc)), list(reversed(cfg.attn_levels))
    prev = rev[0]
    for i, cout in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            it.resnet(f'up_blocks.{i}.resnets.{j}', prev + skips.pop(), cout, temb)
            if rev_attn[i]:
                it.transformer(f'up_blocks.{i}.attentions.{j}', cout, cfg.cross_attention_dim)
            prev = cout
        if i < len(rev) - 1:
            it.conv(f'up_blocks.{i}.upsamplers.0.conv', cout, cout, 3)
    it.norm('conv_norm_out', boc[0]); it.conv('conv_out', boc[0], cfg.out_channels, 3)
    return it.sd


def random_controlnet_state_dict(cfg=SD15, seed=1, device='cuda'):
    """Zero convolutions get NON-zero weights so the residual path is exercised."""
    it = _Init(seed, device)
    it.encoder(cfg)
    ce = cfg.cond_embed_channels
    it.conv('controlnet_cond_embedding.conv_in', 3, ce[0], 3)
    k = 0
    for a, b in zip(ce[:-1], ce[1:]):
        it.conv(f'controlnet_cond_embedding.blocks.{k}', a, a, 3); k += 1
        it.conv(f'controlnet_cond_embedding.blocks.{k}', a, b, 3); k += 1
    it.conv('controlnet_cond_embedding.conv_out', ce[-1], cfg.block_out_channels[0], 3, scale=0.5)
    for i, c in enumerate(skip_channels(cfg)):
        it.conv(f'controlnet_down_blocks.{i}', c, c, 1, scale=0.5)
    it.conv('controlnet_mid_block', cfg.block_out_channels[-1], cfg.block_out_channels[-1], 1, scale=0.5)
    return it.sd
