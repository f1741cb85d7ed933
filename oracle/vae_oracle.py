"""fp32 PyTorch restatement of the AutoencoderKL the reference decodes / encodes with on the hot loop
(``vae.decode(pred_x0 / scaling_factor)``: /root/reference/lib/pipelines/mvedit_3d_pipeline.py:1258-1263; ``vae.encode(x * 2 - 1)``:
:1119-1120, :1440-1443; loaded at /root/reference/lib/apis/adapter3d.py:162-180).

TEST INFRASTRUCTURE ONLY (see oracle/raymarching_oracle.c header): imported by tests/, bench.py's reference legs and smoke() as the
checker -- never by mvedit_b200.

diffusers==0.27.2 (requirements.txt:13) and the SD-1.5 VAE weights are ABSENT offline: the published architecture of
``stable-diffusion-v1-5/vae`` is restated from SURVEY.md Appendix A (block_out_channels [128,256,512,512], layers_per_block 2 -> 3
resnets per decoder block, GroupNorm(32, eps 1e-6), mid block with one single-head attention over the 64x64 latent grid,
latent 4 channels, scaling_factor 0.18215), diffusers state-dict key names kept so that a real checkpoint would load.
==> PARITY UNPINNED (random-init weights).
"""
import math
from dataclasses import dataclass
from typing import Tuple

import torch
import torch.nn.functional as F


@dataclass
class VAEConfig:
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 4
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_groups: int = 32
    scaling_factor: float = 0.18215


SD15_VAE = VAEConfig()
TINY_VAE = VAEConfig(block_out_channels=(64, 64, 128, 128))


def _conv(sd, g, name, cin, cout, k):
    sd[name + '.weight'] = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    sd[name + '.bias'] = torch.randn(cout, generator=g) * 0.02


def _lin(sd, g, name, cin, cout):
    sd[name + '.weight'] = torch.randn(cout, cin, generator=g) / math.sqrt(cin)
    sd[name + '.bias'] = torch.randn(cout, generator=g) * 0.02


def _norm(sd, g, name, c):
    sd[name + '.weight'] = 1 + 0.1 * torch.randn(c, generator=g)
    sd[name + '.bias'] = 0.1 * torch.randn(c, generator=g)


def _resnet(sd, g, p, cin, cout):
    _norm(sd, g, p + '.norm1', cin); _conv(sd, g, p + '.conv1', cin, cout, 3)
    _norm(sd, g, p + '.norm2', cout); _conv(sd, g, p + '.conv2', cout, cout, 3)
    if cin != cout:
        _conv(sd, g, p + '.conv_shortcut', cin, cout, 1)


def _mid(sd, g, p, c):
    _resnet(sd, g, p + '.resnets.0', c, c)
    a = p + '.attentions.0'
    _norm(sd, g, a + '.group_norm', c)
    for n in ('to_q', 'to_k', 'to_v', 'to_out.0'):
        _lin(sd, g, a + '.' + n, c, c)
    _resnet(sd, g, p + '.resnets.1', c, c)


def random_vae_state_dict(cfg=SD15_VAE, seed=3):
    """AutoencoderKL key names (diffusers 0.27): encoder.*, decoder.*, quant_conv, post_quant_conv."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    boc = cfg.block_out_channels
    # encoder
    _conv(sd, g, 'encoder.conv_in', cfg.in_channels, boc[0], 3)
    cin = boc[0]
    for i, cout in enumerate(boc):
        for j in range(cfg.layers_per_block):
            _resnet(sd, g, f'encoder.down_blocks.{i}.resnets.{j}', cin if j == 0 else cout, cout)
        if i < len(boc) - 1:
            _conv(sd, g, f'encoder.down_blocks.{i}.downsamplers.0.conv', cout, cout, 3)
        cin = cout
    _mid(sd, g, 'encoder.mid_block', boc[-1])
    _norm(sd, g, 'encoder.conv_norm_out', boc[-1]); _conv(sd, g, 'encoder.conv_out', boc[-1], 2 * cfg.latent_channels, 3)
    _conv(sd, g, 'quant_conv', 2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)
    # decoder
    _conv(sd, g, 'post_quant_conv', cfg.latent_channels, cfg.latent_channels, 1)
    rev = list(reversed(boc))
    _conv(sd, g, 'decoder.conv_in', cfg.latent_channels, rev[0], 3)
    _mid(sd, g, 'decoder.mid_block', rev[0])
    cin = rev[0]
    for i, cout in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            _resnet(sd, g, f'decoder.up_blocks.{i}.resnets.{j}', cin if j == 0 else cout, cout)
        if i < len(rev) - 1:
            _conv(sd, g, f'decoder.up_blocks.{i}.upsamplers.0.conv', cout, cout, 3)
        cin = cout
    _norm(sd, g, 'decoder.conv_norm_out', boc[0]); _conv(sd, g, 'decoder.conv_out', boc[0], cfg.out_channels, 3)
    return sd


# ------------------------------------------------------------------------------------------------ layers
def _gn(sd, p, x, groups):
    return F.group_norm(x, groups, sd[p + '.weight'], sd[p + '.bias'], 1e-6)


def _c(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + '.weight'], sd[p + '.bias'], stride=stride, padding=padding)


def resnet(sd, p, x, cfg):
    """ResnetBlock2D with temb_channels=None (diffusers Encoder / Decoder)."""
    h = _c(sd, p + '.conv1', F.silu(_gn(sd, p + '.norm1', x, cfg.norm_groups)))
    h = _c(sd, p + '.conv2', F.silu(_gn(sd, p + '.norm2', h, cfg.norm_groups)))
    if (p + '.conv_shortcut.weight') in sd:
        x = _c(sd, p + '.conv_shortcut', x, padding=0)
    return x + h


def attention(sd, p, x, cfg):
    """diffusers Attention(heads=1, dim_head=C, residual_connection=True, group_norm) with AttnProcessor2_0."""
    B, C, H, W = x.shape
    h = _gn(sd, p + '.group_norm', x, cfg.norm_groups).reshape(B, C, H * W).transpose(1, 2)
    lin = lambda n, t: F.linear(t, sd[p + f'.{n}.weight'], sd[p + f'.{n}.bias'])
    q, k, v = lin('to_q', h), lin('to_k', h), lin('to_v', h)
    o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
    o = lin('to_out.0', o)
    return o.transpose(1, 2).reshape(B, C, H, W) + x


def mid_block(sd, p, x, cfg):
    x = resnet(sd, p + '.resnets.0', x, cfg)
    x = attention(sd, p + '.attentions.0', x, cfg)
    return resnet(sd, p + '.resnets.1', x, cfg)


def decode(sd, cfg, z):
    """AutoencoderKL.decode(z).sample: post_quant_conv -> Decoder.  z [B,4,L,L] (already divided by scaling_factor) -> [B,3,8L,8L]."""
    x = _c(sd, 'post_quant_conv', z, padding=0)
    x = _c(sd, 'decoder.conv_in', x)
    x = mid_block(sd, 'decoder.mid_block', x, cfg)
    n = len(cfg.block_out_channels)
    for i in range(n):
        for j in range(cfg.layers_per_block + 1):
            x = resnet(sd, f'decoder.up_blocks.{i}.resnets.{j}', x, cfg)
        if i < n - 1:
            x = F.interpolate(x, scale_factor=2.0, mode='nearest')
            x = _c(sd, f'decoder.up_blocks.{i}.upsamplers.0.conv', x)
    x = F.silu(_gn(sd, 'decoder.conv_norm_out', x, cfg.norm_groups))
    return _c(sd, 'decoder.conv_out', x)


def encode_moments(sd, cfg, x):
    """AutoencoderKL.encode(x).latent_dist parameters: (mean, logvar) each [B,4,L,L].  Downsample2D(padding=0): F.pad (0,1,0,1)."""
    h = _c(sd, 'encoder.conv_in', x)
    n = len(cfg.block_out_channels)
    for i in range(n):
        for j in range(cfg.layers_per_block):
            h = resnet(sd, f'encoder.down_blocks.{i}.resnets.{j}', h, cfg)
        if i < n - 1:
            h = _c(sd, f'encoder.down_blocks.{i}.downsamplers.0.conv', F.pad(h, (0, 1, 0, 1)), stride=2, padding=0)
    h = mid_block(sd, 'encoder.mid_block', h, cfg)
    h = _c(sd, 'encoder.conv_out', F.silu(_gn(sd, 'encoder.conv_norm_out', h, cfg.norm_groups)))
    m = _c(sd, 'quant_conv', h, padding=0)
    mean, logvar = m.chunk(2, dim=1)
    return mean, logvar.clamp(-30.0, 20.0)


def decode_targets(sd, cfg, pred_x0):
    """mvedit_3d_pipeline.py:1258-1263: (vae.decode(x0 / sf) / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1) -> fp32 [B,H,W,3]."""
    return (decode(sd, cfg, pred_x0 / cfg.scaling_factor) / 2 + 0.5).clamp(min=0, max=1).permute(0, 2, 3, 1).float()
