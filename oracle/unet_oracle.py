"""fp32 PyTorch restatement of the Stable-Diffusion-1.5 denoiser stack the reference drives: UNet2DConditionModel,
ControlNetModel (v1.1), unet_enc / unet_dec split, and Adapter3DMixin.get_noise_pred{,_p1,_p2}.

TEST INFRASTRUCTURE ONLY (see oracle/raymarching_oracle.c header).  Device-agnostic plain torch (runs on CPU, or in fp32 on the GPU
box as the checker for mid-size configs).

What it follows:
  * diffusers==0.27.2 (requirements.txt:13) is an ABSENT third-party dependency: its published architecture for
    ``stable-diffusion-v1-5`` / ``control_v11*_sd15_*`` is restated -- layer by layer, diffusers state-dict
    key names kept so a real checkpoint would load.   ==> PARITY UNPINNED (no diffusers, no weights offline): random-init weights.
  * in-tree call sites that ARE restated line by line:
      unet_enc / unet_dec                   /root/reference/lib/models/architecture/diffusers.py:57-164
      get_noise_pred / _p1 / _p2            /root/reference/lib/pipelines/adapter3d_mixin.py:68-317
      CrossImageAttnProcWrapper             /root/reference/lib/models/architecture/joint_attn.py:11-37
"""
import math
from dataclasses import dataclass, field
from typing import Tuple

import torch
import torch.nn.functional as F


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    num_heads: Tuple[int, ...] = (8, 8, 8, 8)          # SD1.5: attention_head_dim=8 means 8 heads at every level
    cross_attention_dim: int = 768
    attn_levels: Tuple[bool, ...] = (True, True, True, False)   # CrossAttnDownBlock2D x3 + DownBlock2D
    cond_embed_channels: Tuple[int, ...] = (16, 32, 96, 256)    # ControlNet hint embedding
    norm_groups: int = 32

    @property
    def time_embed_dim(self):
        return self.block_out_channels[0] * 4


SD15 = UNetConfig()
TINY = UNetConfig(block_out_channels=(64, 128, 128, 128), num_heads=(1, 2, 2, 2), cross_attention_dim=64)


# ------------------------------------------------------------------------------------------------ random weights (diffusers key names)
def _conv(sd, g, name, cin, cout, k, zero=False, scale=1.0):
    std = scale / math.sqrt(cin * k * k)
    sd[name + '.weight'] = torch.zeros(cout, cin, k, k) if zero else torch.randn(cout, cin, k, k, generator=g) * std
    sd[name + '.bias'] = torch.randn(cout, generator=g) * 0.02


def _lin(sd, g, name, cin, cout, bias=True, scale=1.0):
    sd[name + '.weight'] = torch.randn(cout, cin, generator=g) * (scale / math.sqrt(cin))
    if bias:
        sd[name + '.bias'] = torch.randn(cout, generator=g) * 0.02


def _norm(sd, g, name, c):
    sd[name + '.weight'] = 1 + 0.1 * torch.randn(c, generator=g)
    sd[name + '.bias'] = 0.1 * torch.randn(c, generator=g)


def _resnet(sd, g, p, cin, cout, temb):
    _norm(sd, g, p + '.norm1', cin); _conv(sd, g, p + '.conv1', cin, cout, 3)
    _lin(sd, g, p + '.time_emb_proj', temb, cout)
    _norm(sd, g, p + '.norm2', cout); _conv(sd, g, p + '.conv2', cout, cout, 3)
    if cin != cout:
        _conv(sd, g, p + '.conv_shortcut', cin, cout, 1)


def _transformer(sd, g, p, c, cross):
    _norm(sd, g, p + '.norm', c); _conv(sd, g, p + '.proj_in', c, c, 1)
    b = p + '.transformer_blocks.0'
    for n in ('norm1', 'norm2', 'norm3'):
        _norm(sd, g, b + '.' + n, c)
    for a, kv in (('attn1', c), ('attn2', cross)):
        _lin(sd, g, b + f'.{a}.to_q', c, c, bias=False); _lin(sd, g, b + f'.{a}.to_k', kv, c, bias=False)
        _lin(sd, g, b + f'.{a}.to_v', kv, c, bias=False); _lin(sd, g, b + f'.{a}.to_out.0', c, c)
    _lin(sd, g, b + '.ff.net.0.proj', c, 8 * c); _lin(sd, g, b + '.ff.net.2', 4 * c, c)
    _conv(sd, g, p + '.proj_out', c, c, 1)


def _encoder(sd, g, cfg):
    boc, temb = cfg.block_out_channels, cfg.time_embed_dim
    _conv(sd, g, 'conv_in', cfg.in_channels, boc[0], 3)
    _lin(sd, g, 'time_embedding.linear_1', boc[0], temb); _lin(sd, g, 'time_embedding.linear_2', temb, temb)
    cin = boc[0]
    for i, cout in enumerate(boc):
        for j in range(cfg.layers_per_block):
            _resnet(sd, g, f'down_blocks.{i}.resnets.{j}', cin if j == 0 else cout, cout, temb)
            if cfg.attn_levels[i]:
                _transformer(sd, g, f'down_blocks.{i}.attentions.{j}', cout, cfg.cross_attention_dim)
        if i < len(boc) - 1:
            _conv(sd, g, f'down_blocks.{i}.downsamplers.0.conv', cout, cout, 3)
        cin = cout
    c = boc[-1]
    _resnet(sd, g, 'mid_block.resnets.0', c, c, temb); _transformer(sd, g, 'mid_block.attentions.0', c, cfg.cross_attention_dim)
    _resnet(sd, g, 'mid_block.resnets.1', c, c, temb)


def skip_channels(cfg):
    boc = cfg.block_out_channels
    ch = [boc[0]]
    for i, c in enumerate(boc):
        ch += [c] * cfg.layers_per_block
        if i < len(boc) - 1:
            ch.append(c)
    return ch


def random_unet_state_dict(cfg=SD15, seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    _encoder(sd, g, cfg)
    boc, temb = cfg.block_out_channels, cfg.time_embed_dim
    skips = skip_channels(cfg)
    rev = list(reversed(boc))
    prev = rev[0]
    for i, cout in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            skip = skips.pop()
            _resnet(sd, g, f'up_blocks.{i}.resnets.{j}', prev + skip, cout, temb)
            if list(reversed(cfg.attn_levels))[i]:
                _transformer(sd, g, f'up_blocks.{i}.attentions.{j}', cout, cfg.cross_attention_dim)
            prev = cout
        if i < len(rev) - 1:
            _conv(sd, g, f'up_blocks.{i}.upsamplers.0.conv', cout, cout, 3)
    _norm(sd, g, 'conv_norm_out', boc[0]); _conv(sd, g, 'conv_out', boc[0], cfg.out_channels, 3)
    return sd


def random_controlnet_state_dict(cfg=SD15, seed=1, zero_convs_nonzero=True):
    """ControlNet v1.1.  The zero convolutions are given NON-zero weights so the path is exercised."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    _encoder(sd, g, cfg)
    ce = cfg.cond_embed_channels
    _conv(sd, g, 'controlnet_cond_embedding.conv_in', 3, ce[0], 3)
    k = 0
    for a, b in zip(ce[:-1], ce[1:]):
        _conv(sd, g, f'controlnet_cond_embedding.blocks.{k}', a, a, 3); k += 1
        _conv(sd, g, f'controlnet_cond_embedding.blocks.{k}', a, b, 3); k += 1
    _conv(sd, g, 'controlnet_cond_embedding.conv_out', ce[-1], cfg.block_out_channels[0], 3, zero=not zero_convs_nonzero, scale=0.5)
    for i, c in enumerate(skip_channels(cfg)):
        _conv(sd, g, f'controlnet_down_blocks.{i}', c, c, 1, zero=not zero_convs_nonzero, scale=0.5)
    _conv(sd, g, 'controlnet_mid_block', cfg.block_out_channels[-1], cfg.block_out_channels[-1], 1, zero=not zero_convs_nonzero, scale=0.5)
    return sd


# ------------------------------------------------------------------------------------------------ IP-Adapter (ip_adapter.py:85-113)
IP_OPTS = {}          # id(state dict) -> dict(ip_tokens=.., ip_scale=..) for a UNet, dict(drop_tokens=..) for a ControlNet


def transformer_paths(cfg=SD15):
    """The UNet's transformer blocks in diffusers ``attn_processors`` order (down_blocks, up_blocks, mid_block -- the module
    registration order of UNet2DConditionModel): processor 2i is attn1 of block i, 2i + 1 its attn2 -- the numbering of the
    ``ip_adapter`` checkpoint section ("1.to_k_ip.weight", "3.to_k_ip.weight", ..., ip_adapter.py:61-62)."""
    paths = []
    for i in range(len(cfg.block_out_channels)):
        if cfg.attn_levels[i]:
            paths += [f'down_blocks.{i}.attentions.{j}' for j in range(cfg.layers_per_block)]
    rev = list(reversed(cfg.attn_levels))
    for i in range(len(cfg.block_out_channels)):
        if rev[i]:
            paths += [f'up_blocks.{i}.attentions.{j}' for j in range(cfg.layers_per_block + 1)]
    return paths + ['mid_block.attentions.0']


def random_ip_adapter_state_dict(cfg=SD15, seed=5):
    """'ip_adapter' section of an IP-Adapter checkpoint: {2i+1}.to_k_ip.weight / to_v_ip.weight [hidden, cross_attention_dim]."""
    g = torch.Generator().manual_seed(seed)
    boc = cfg.block_out_channels
    sd = {}
    for i, p in enumerate(transformer_paths(cfg)):
        blk = int(p.split('.')[1]) if not p.startswith('mid') else len(boc) - 1
        c = boc[-1] if p.startswith('mid') else (boc[blk] if p.startswith('down') else list(reversed(boc))[blk])
        for n in ('to_k_ip', 'to_v_ip'):
            sd[f'{2 * i + 1}.{n}.weight'] = torch.randn(c, cfg.cross_attention_dim, generator=g) / math.sqrt(cfg.cross_attention_dim)
    return sd


def set_ip_adapter(unet_sd, ip_sd, cfg=SD15, num_tokens=16, scale=1.0, controlnet_sds=(), cn_num_tokens=4):
    """IPAdapter.set_ip_adapter (ip_adapter.py:85-113) for the functional oracle: merges the adapter weights into the UNet state dict
    under <block>.transformer_blocks.0.attn2.to_{k,v}_ip.weight and registers the processor options."""
    for i, p in enumerate(transformer_paths(cfg)):
        for n in ('to_k_ip', 'to_v_ip'):
            unet_sd[f'{p}.transformer_blocks.0.attn2.{n}.weight'] = ip_sd[f'{2 * i + 1}.{n}.weight'].to(next(iter(unet_sd.values())))
    IP_OPTS[id(unet_sd)] = dict(ip_tokens=num_tokens, ip_scale=scale)
    for c in controlnet_sds:
        IP_OPTS[id(c)] = dict(drop_tokens=cn_num_tokens)


# ------------------------------------------------------------------------------------------------ layers
def timestep_embedding(t, dim):
    """diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0)."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t.float()[:, None] * torch.exp(exponent)[None]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


def _gn(sd, p, x, groups, eps):
    return F.group_norm(x, groups, sd[p + '.weight'], sd[p + '.bias'], eps)


def _c(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + '.weight'], sd[p + '.bias'], stride=stride, padding=padding)


def _l(sd, p, x):
    return F.linear(x, sd[p + '.weight'], sd.get(p + '.bias'))


def resnet(sd, p, x, emb, cfg):
    h = _c(sd, p + '.conv1', F.silu(_gn(sd, p + '.norm1', x, cfg.norm_groups, 1e-5)))
    h = h + _l(sd, p + '.time_emb_proj', F.silu(emb))[:, :, None, None]
    h = _c(sd, p + '.conv2', F.silu(_gn(sd, p + '.norm2', h, cfg.norm_groups, 1e-5)))
    if (p + '.conv_shortcut.weight') in sd:
        x = _c(sd, p + '.conv_shortcut', x, padding=0)
    return x + h


def attention(sd, p, x, ctx, heads, num_cross_attn_imgs=1, ip_tokens=0, ip_scale=1.0, drop_tokens=0):
    """Attention with AttnProcessor2_0 semantics; with num_cross_attn_imgs=2 the CrossImageAttnProcWrapper view
    (joint_attn.py:13-33): self-attention spans the (ref, view) pair, text context is averaged over the pair.
    Cross-attention variants of the IP-Adapter (pinned against the reference classes by tests/test_reference_pins.py):
      ip_tokens > 0   IPAttnProcessor2_0 (ip_adapter/attention_processor.py:301-396): the last ip_tokens context tokens go through
                      to_k_ip / to_v_ip (keys p + '.to_k_ip.weight' / '.to_v_ip.weight') in a second attention, added with ip_scale;
      drop_tokens > 0 CNAttnProcessor2_0 (:472-556): the ControlNets attend to context[:, :-drop_tokens] only (NB num_tokens defaults
                      to 4 there although the plus adapter appends 16 tokens: 12 image tokens stay in -- quirk kept)."""
    B, S, C = x.shape
    if num_cross_attn_imgs > 1:
        x = x.reshape(B // num_cross_attn_imgs, num_cross_attn_imgs * S, C)
        if ctx is not None:
            ctx = ctx.reshape(B // num_cross_attn_imgs, num_cross_attn_imgs, *ctx.shape[1:]).mean(dim=1)
    ip_ctx = None
    if ctx is not None and drop_tokens > 0:
        ctx = ctx[:, :ctx.shape[1] - drop_tokens]
    if ctx is not None and ip_tokens > 0:
        ctx, ip_ctx = ctx[:, :ctx.shape[1] - ip_tokens], ctx[:, ctx.shape[1] - ip_tokens:]
    kv = x if ctx is None else ctx
    q, k, v = _l(sd, p + '.to_q', x), _l(sd, p + '.to_k', kv), _l(sd, p + '.to_v', kv)
    d = C // heads
    sh = lambda t: t.reshape(t.shape[0], t.shape[1], heads, d).transpose(1, 2)
    o = F.scaled_dot_product_attention(sh(q), sh(k), sh(v))
    o = o.transpose(1, 2).reshape(x.shape[0], x.shape[1], C)
    if ip_ctx is not None:
        o_ip = F.scaled_dot_product_attention(sh(q), sh(_l(sd, p + '.to_k_ip', ip_ctx)), sh(_l(sd, p + '.to_v_ip', ip_ctx)))
        o = o + ip_scale * o_ip.transpose(1, 2).reshape(x.shape[0], x.shape[1], C)
    o = _l(sd, p + '.to_out.0', o)
    return o.reshape(B, S, C)


def transformer(sd, p, x, ctx, heads, cfg, num_cross_attn_imgs=1):
    ip = IP_OPTS.get(id(sd), {})
    B, C, H, W = x.shape
    res = x
    h = _c(sd, p + '.proj_in', _gn(sd, p + '.norm', x, cfg.norm_groups, 1e-6), padding=0)
    h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
    b = p + '.transformer_blocks.0'
    ln = lambda n, t: F.layer_norm(t, (C,), sd[b + f'.{n}.weight'], sd[b + f'.{n}.bias'], 1e-5)
    h = h + attention(sd, b + '.attn1', ln('norm1', h), None, heads, num_cross_attn_imgs)
    h = h + attention(sd, b + '.attn2', ln('norm2', h), ctx, heads, num_cross_attn_imgs, **ip)
    f = _l(sd, b + '.ff.net.0.proj', ln('norm3', h))
    a, gate = f.chunk(2, dim=-1)
    h = h + _l(sd, b + '.ff.net.2', a * F.gelu(gate))
    h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
    return _c(sd, p + '.proj_out', h, padding=0) + res


def time_embed(sd, cfg, t, batch, device, dtype):
    t = torch.as_tensor(t, device=device).reshape(-1).expand(batch)
    e = timestep_embedding(t, cfg.block_out_channels[0]).to(dtype)
    return _l(sd, 'time_embedding.linear_2', F.silu(_l(sd, 'time_embedding.linear_1', e)))


def encoder_forward(sd, cfg, sample, emb, ctx, n_imgs=1):
    """down path shared by UNet and ControlNet: returns (res_samples tuple, sample)."""
    res = (sample,)
    for i in range(len(cfg.block_out_channels)):
        for j in range(cfg.layers_per_block):
            sample = resnet(sd, f'down_blocks.{i}.resnets.{j}', sample, emb, cfg)
            if cfg.attn_levels[i]:
                sample = transformer(sd, f'down_blocks.{i}.attentions.{j}', sample, ctx, cfg.num_heads[i], cfg, n_imgs)
            res += (sample,)
        if i < len(cfg.block_out_channels) - 1:
            sample = _c(sd, f'down_blocks.{i}.downsamplers.0.conv', sample, stride=2)
            res += (sample,)
    return res, sample


def mid_forward(sd, cfg, sample, emb, ctx, n_imgs=1):
    sample = resnet(sd, 'mid_block.resnets.0', sample, emb, cfg)
    sample = transformer(sd, 'mid_block.attentions.0', sample, ctx, cfg.num_heads[-1], cfg, n_imgs)
    return resnet(sd, 'mid_block.resnets.1', sample, emb, cfg)


def unet_enc(sd, cfg, sample, t, encoder_hidden_states, cross_attention_kwargs=None):
    """diffusers.py:57-99."""
    n = (cross_attention_kwargs or {}).get('num_cross_attn_imgs', 1)
    emb = time_embed(sd, cfg, t, sample.shape[0], sample.device, sample.dtype)
    sample = _c(sd, 'conv_in', sample)
    res, sample = encoder_forward(sd, cfg, sample, emb, encoder_hidden_states, n)
    return emb, res, sample


def unet_dec(sd, cfg, emb, down_block_res_samples, sample, encoder_hidden_states, cross_attention_kwargs=None,
             down_block_additional_residuals=None, mid_block_additional_residual=None):
    """diffusers.py:102-164."""
    n = (cross_attention_kwargs or {}).get('num_cross_attn_imgs', 1)
    is_controlnet = mid_block_additional_residual is not None and down_block_additional_residuals is not None
    if is_controlnet:
        down_block_res_samples = tuple(a + b for a, b in zip(down_block_res_samples, down_block_additional_residuals))
    sample = mid_forward(sd, cfg, sample, emb, encoder_hidden_states, n)
    if is_controlnet:
        sample = sample + mid_block_additional_residual
    res = list(down_block_res_samples)
    rev_attn = list(reversed(cfg.attn_levels))
    rev_heads = list(reversed(cfg.num_heads))
    nlev = len(cfg.block_out_channels)
    for i in range(nlev):
        for j in range(cfg.layers_per_block + 1):
            sample = torch.cat([sample, res.pop()], dim=1)
            sample = resnet(sd, f'up_blocks.{i}.resnets.{j}', sample, emb, cfg)
            if rev_attn[i]:
                sample = transformer(sd, f'up_blocks.{i}.attentions.{j}', sample, encoder_hidden_states, rev_heads[i], cfg, n)
        if i < nlev - 1:
            sample = F.interpolate(sample, scale_factor=2.0, mode='nearest')
            sample = _c(sd, f'up_blocks.{i}.upsamplers.0.conv', sample)
    sample = F.silu(_gn(sd, 'conv_norm_out', sample, cfg.norm_groups, 1e-5))
    return _c(sd, 'conv_out', sample)


def unet_forward(sd, cfg, sample, t, encoder_hidden_states, cross_attention_kwargs=None, down_block_additional_residuals=None,
                 mid_block_additional_residual=None):
    emb, res, s = unet_enc(sd, cfg, sample, t, encoder_hidden_states, cross_attention_kwargs)
    return unet_dec(sd, cfg, emb, res, s, encoder_hidden_states, cross_attention_kwargs, down_block_additional_residuals,
                    mid_block_additional_residual)


def controlnet_forward(sd, cfg, sample, t, encoder_hidden_states, controlnet_cond, conditioning_scale=1.0):
    """ControlNetModel.forward (guess_mode=False): 12 down residuals + mid residual, each scaled."""
    emb = time_embed(sd, cfg, t, sample.shape[0], sample.device, sample.dtype)
    sample = _c(sd, 'conv_in', sample)
    h = F.silu(_c(sd, 'controlnet_cond_embedding.conv_in', controlnet_cond))
    nb = 2 * (len(cfg.cond_embed_channels) - 1)
    for k in range(nb):
        h = F.silu(_c(sd, f'controlnet_cond_embedding.blocks.{k}', h, stride=2 if k % 2 == 1 else 1))
    sample = sample + _c(sd, 'controlnet_cond_embedding.conv_out', h)
    res, sample = encoder_forward(sd, cfg, sample, emb, encoder_hidden_states)
    sample = mid_forward(sd, cfg, sample, emb, encoder_hidden_states)
    down = [_c(sd, f'controlnet_down_blocks.{i}', r, padding=0) * conditioning_scale for i, r in enumerate(res)]
    mid = _c(sd, 'controlnet_mid_block', sample, padding=0) * conditioning_scale
    return down, mid


def multi_controlnet_forward(sds, cfg, sample, t, encoder_hidden_states, conds, scales):
    """diffusers MultiControlNetModel: sum of the nets' residuals."""
    down = mid = None
    for sd, cond, sc in zip(sds, conds, scales):
        d, m = controlnet_forward(sd, cfg, sample, t, encoder_hidden_states, cond, sc)
        down, mid = (d, m) if down is None else ([a + b for a, b in zip(down, d)], mid + m)
    return down, mid


# ------------------------------------------------------------------------------------------------ Adapter3DMixin (adapter3d_mixin.py:68-317)
def _split_ref(batch_latent, latent_size, prompt):
    shp = batch_latent.shape
    if shp[2] == 2 * shp[3]:
        unet_in = batch_latent.reshape(*shp[:2], 2, shp[3], shp[3]).permute(0, 2, 1, 3, 4).reshape(shp[0] * 2, shp[1], shp[3], shp[3])
        return dict(num_cross_attn_imgs=2), unet_in, batch_latent[:, :, -latent_size:], \
            prompt.unsqueeze(1).expand(-1, 2, -1, -1).reshape(-1, *prompt.shape[1:]), prompt
    return None, batch_latent, batch_latent, prompt, prompt


def _interleave_zero(t):
    return torch.stack([torch.zeros_like(t), t], dim=1).view(-1, *t.shape[1:])


def get_noise_pred(unet_sd, cn_sds, cfg, latent_batches, prompt_embeds_batches, ctrl_images_batches, ctrl_depths_batches, t,
                   tile_weight, depth_weight, guidance_scale):
    """adapter3d_mixin.py:68-135 with controlnet = [tile, depth]."""
    latent_size = latent_batches[0].size(-1)
    out = []
    for lat, pe, ci, cd in zip(latent_batches, prompt_embeds_batches, ctrl_images_batches, ctrl_depths_batches):
        cak, unet_in, cn_in, unet_pe, cn_pe = _split_ref(lat, latent_size, pe)
        down, mid = multi_controlnet_forward(cn_sds, cfg, cn_in, t, cn_pe, [ci, cd], [tile_weight, depth_weight])
        if cak is not None:
            down, mid = [_interleave_zero(d) for d in down], _interleave_zero(mid)
        o = unet_forward(unet_sd, cfg, unet_in, t, unet_pe, cak, down, mid)
        if cak is not None:
            o = o.view(lat.shape[0], 2, lat.shape[1], lat.shape[3], lat.shape[3])[:, 1]
        out.append(o)
    out = torch.cat(out, dim=0)
    uncond, text = out.chunk(2)
    return guidance_scale * text + (1 - guidance_scale) * uncond


def get_noise_pred_p1(unet_sd, cfg, latent_batches, prompt_embeds_batches, t, guidance_scale):
    """adapter3d_mixin.py:137-237 without extra ControlNets / cond_noisy latents (text-to-3D recipe)."""
    latent_size = latent_batches[0].size(-1)
    out, dec_args, dec_kwargs = [], [], []
    for lat, pe in zip(latent_batches, prompt_embeds_batches):
        cak, unet_in, cn_in, unet_pe, cn_pe = _split_ref(lat, latent_size, pe)
        emb, res, s = unet_enc(unet_sd, cfg, unet_in, t, unet_pe, cak)
        dec_args.append((emb, res, s))
        dec_kwargs.append(dict(encoder_hidden_states=unet_pe, cross_attention_kwargs=cak, down_block_additional_residuals=None,
                               mid_block_additional_residual=None))
        o = unet_dec(unet_sd, cfg, *dec_args[-1], **dec_kwargs[-1])
        if cak is not None:
            o = o.view(lat.shape[0], 2, lat.shape[1], lat.shape[3], lat.shape[3])[:, 1]
        out.append(o)
    out = torch.cat(out, dim=0)
    uncond, text = out.chunk(2)
    return guidance_scale * text + (1 - guidance_scale) * uncond, dec_args, dec_kwargs


def get_noise_pred_p2(unet_sd, cn_sds, cfg, latent_batches, prompt_embeds_batches, dec_args, dec_kwargs, t, guidance_scale,
                      ctrl_images_batches, tile_weight, ctrl_depths_batches=None, depth_weight=None):
    """adapter3d_mixin.py:239-317."""
    latent_size = latent_batches[0].size(-1)
    out = []
    cdb = [None] * len(latent_batches) if ctrl_depths_batches is None else ctrl_depths_batches
    for lat, pe, da, dk, ci, cd in zip(latent_batches, prompt_embeds_batches, dec_args, dec_kwargs, ctrl_images_batches, cdb):
        ref = lat.shape[2] == 2 * lat.shape[3]
        cn_in = lat[:, :, -latent_size:] if ref else lat
        nets = cn_sds[:1 if cd is None else 2]
        down, mid = multi_controlnet_forward(nets, cfg, cn_in, t, pe, [ci] if cd is None else [ci, cd],
                                             [tile_weight] if cd is None else [tile_weight, depth_weight])
        if ref:
            down, mid = [_interleave_zero(d) for d in down], _interleave_zero(mid)
        if dk['down_block_additional_residuals'] is not None:
            down = [a + b for a, b in zip(down, dk['down_block_additional_residuals'])]
            mid = mid + dk['mid_block_additional_residual']
        dk_ = dict(dk)
        dk_.update(down_block_additional_residuals=down, mid_block_additional_residual=mid)
        o = unet_dec(unet_sd, cfg, *da, **dk_)
        if ref:
            o = o.view(lat.shape[0], 2, lat.shape[1], lat.shape[3], lat.shape[3])[:, 1]
        out.append(o)
    out = torch.cat(out, dim=0)
    uncond, cond = out.chunk(2)
    return guidance_scale * cond + (1 - guidance_scale) * uncond
