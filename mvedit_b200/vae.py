"""AutoencoderKL (SD-1.5 VAE) on the tcgen05 kernels of libmvedit_b200 -- the B200 replacement for the diffusers module the reference
loads at /root/reference/lib/apis/adapter3d.py:162-180 and calls inside the hot loop:

    vae.decode(pred_x0 / vae.config.scaling_factor, return_dict=False)[0]      mvedit_3d_pipeline.py:1258-1263   (every step, N views)
    vae.encode(x * 2 - 1).latent_dist.sample() / .mean                          :1119-1120, :1440-1443            (init / dynamic blend)

Same call surface (``decode``, ``encode(...).latent_dist``, ``config.scaling_factor``, ``device``/``dtype``), diffusers state-dict keys.
Everything is bf16 NHWC on the kernels the UNet uses: 3x3 convolutions = implicit GEMM (mve_conv3x3_bf16), GroupNorm+SiLU
(mve_groupnorm_bf16), nearest x2 upsample, stride-2 convolution = im2col (right/bottom zero pad, as Downsample2D(padding=0)) + GEMM.
The mid-block attention has ONE head of d = 512 over S = 4096 tokens -- outside the flash kernel's head dims -- and is 1.4 % of the
decoder's FLOPs: it runs as score GEMM -> mve_softmax_rows_bf16 -> value GEMM, with V produced already transposed by a GEMM whose
A operand is the weight (V^T = Wv . X^T), so no transpose kernel exists; V's bias is added after the value GEMM (rows of P sum to 1).
``post_quant_conv`` / ``quant_conv`` (1x1): the first is a GEMM on the 64-padded latent, the second is folded into ``encoder.conv_out``
at load time (a 1x1 convolution after a 3x3 one is a 3x3 one).
"""
import math
from types import SimpleNamespace

import torch

from . import tc_ops as T
from .unet import _Weights, _bf, _f32, _pad64


class VAEConfig(SimpleNamespace):
    pass


def sd15_vae_config():
    return VAEConfig(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                     norm_groups=32, scaling_factor=0.18215)


def random_vae_state_dict(cfg=None, seed=3, device='cuda'):
    """Random AutoencoderKL weights of the published SD-1.5 shapes, diffusers key names (no checkpoints offline: BASELINE.json)."""
    cfg = cfg or sd15_vae_config()
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}

    def conv(name, cin, cout, k):
        sd[name + '.weight'] = torch.randn(cout, cin, k, k, generator=g, device=device) / math.sqrt(cin * k * k)
        sd[name + '.bias'] = torch.randn(cout, generator=g, device=device) * 0.02

    def lin(name, cin, cout):
        sd[name + '.weight'] = torch.randn(cout, cin, generator=g, device=device) / math.sqrt(cin)
        sd[name + '.bias'] = torch.randn(cout, generator=g, device=device) * 0.02

    def norm(name, c):
        sd[name + '.weight'] = 1 + 0.1 * torch.randn(c, generator=g, device=device)
        sd[name + '.bias'] = 0.1 * torch.randn(c, generator=g, device=device)

    def resnet(p, cin, cout):
        norm(p + '.norm1', cin); conv(p + '.conv1', cin, cout, 3)
        norm(p + '.norm2', cout); conv(p + '.conv2', cout, cout, 3)
        if cin != cout:
            conv(p + '.conv_shortcut', cin, cout, 1)

    def mid(p, c):
        resnet(p + '.resnets.0', c, c)
        norm(p + '.attentions.0.group_norm', c)
        for n in ('to_q', 'to_k', 'to_v', 'to_out.0'):
            lin(p + '.attentions.0.' + n, c, c)
        resnet(p + '.resnets.1', c, c)

    boc = cfg.block_out_channels
    conv('encoder.conv_in', cfg.in_channels, boc[0], 3)
    cin = boc[0]
    for i, cout in enumerate(boc):
        for j in range(cfg.layers_per_block):
            resnet(f'encoder.down_blocks.{i}.resnets.{j}', cin if j == 0 else cout, cout)
        if i < len(boc) - 1:
            conv(f'encoder.down_blocks.{i}.downsamplers.0.conv', cout, cout, 3)
        cin = cout
    mid('encoder.mid_block', boc[-1])
    norm('encoder.conv_norm_out', boc[-1]); conv('encoder.conv_out', boc[-1], 2 * cfg.latent_channels, 3)
    conv('quant_conv', 2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)
    conv('post_quant_conv', cfg.latent_channels, cfg.latent_channels, 1)
    rev = list(reversed(boc))
    conv('decoder.conv_in', cfg.latent_channels, rev[0], 3)
    mid('decoder.mid_block', rev[0])
    cin = rev[0]
    for i, cout in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            resnet(f'decoder.up_blocks.{i}.resnets.{j}', cin if j == 0 else cout, cout)
        if i < len(rev) - 1:
            conv(f'decoder.up_blocks.{i}.upsamplers.0.conv', cout, cout, 3)
        cin = cout
    norm('decoder.conv_norm_out', boc[0]); conv('decoder.conv_out', boc[0], cfg.out_channels, 3)
    return sd


class DiagonalGaussian:
    """diffusers DiagonalGaussianDistribution: the two members the reference touches (``sample``, ``mean``)."""

    def __init__(self, mean, logvar):
        self.mean = mean
        self.logvar = logvar.clamp(-30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None):
        return self.mean + self.std * torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)

    def mode(self):
        return self.mean


class AutoencoderKL:
    """decode / encode of diffusers' AutoencoderKL, bf16 on libmvedit_b200."""

    def __init__(self, state_dict, cfg=None, device='cuda'):
        self.config = cfg or sd15_vae_config()
        self.device = torch.device(device)
        self.dtype = torch.bfloat16
        self.w = _Weights(state_dict, self.device)
        self._scores = None

    # ------------------------------------------------------------------ blocks (NHWC bf16)
    def _resnet(self, p, x):
        w, G = self.w, self.config.norm_groups
        h = T.groupnorm(x, *w.gn(p + '.norm1'), G, 1e-6, silu=True)
        h = T.conv3x3(h, *w.conv3x3(p + '.conv1'))
        h = T.groupnorm(h, *w.gn(p + '.norm2'), G, 1e-6, silu=True)
        if (p + '.conv_shortcut.weight') in w.sd:
            ws, bs = w.linear(p + '.conv_shortcut')
            B, H, W_, C = x.shape
            x = T.gemm(x.view(-1, C), ws, bias=bs).view(B, H, W_, -1)
        wc2, bc2 = w.conv3x3(p + '.conv2')
        return T.conv3x3(h, wc2, bias=bc2, residual=x)

    def _attention(self, p, x):
        w, G = self.w, self.config.norm_groups
        B, H, W_, C = x.shape
        S = H * W_
        h = T.groupnorm(x, *w.gn(p + '.group_norm'), G, 1e-6, silu=False).view(B * S, C)
        wq, bq = w.linear(p + '.to_q')
        wk, bk = w.linear(p + '.to_k')
        wv, bv = w.linear(p + '.to_v')
        wo, bo = w.linear(p + '.to_out.0')
        q = T.gemm(h, wq, bias=bq)
        k = T.gemm(h, wk, bias=bk)
        vt = T.gemm(wv, h)                                          # V^T [C, B*S]  (bias deferred: softmax rows sum to 1)
        if self._scores is None or self._scores.shape != (S, S):
            self._scores = torch.empty(S, S, dtype=torch.bfloat16, device=x.device)
        o = torch.empty(B * S, C, dtype=torch.bfloat16, device=x.device)
        scale = C ** -0.5
        for b in range(B):
            r = slice(b * S, (b + 1) * S)
            T.gemm(q[r], k[r], out=self._scores)
            T.softmax_rows(self._scores, scale)
            T.gemm(self._scores, vt[:, r], bias=bv, out=o[r])
        return T.gemm(o, wo, bias=bo, residual=x.view(B * S, C)).view(B, H, W_, C)

    def _mid(self, p, x):
        x = self._resnet(p + '.resnets.0', x)
        x = self._attention(p + '.attentions.0', x)
        return self._resnet(p + '.resnets.1', x)

    # ------------------------------------------------------------------ decode
    def _decode_nhwc(self, z, alpha=1.0, bias_shift=0.0):
        """z [B,4,L,L] -> [B,8L,8L,3] bf16 NHWC = (decoder(z) + bias_shift) * alpha."""
        cfg, w = self.config, self.w
        B, C, L, _ = z.shape
        x = T.nchw_to_nhwc_pad(z, 64)
        key = 'post_quant_conv#pad'
        if key not in w.lin:
            wt = w.sd['post_quant_conv.weight'].reshape(cfg.latent_channels, -1).float()
            wp = torch.zeros(cfg.latent_channels, 64, device=wt.device)
            wp[:, :wt.shape[1]] = wt
            w.lin[key] = (_bf(wp, self.device), _f32(w.sd['post_quant_conv.bias'], self.device))
        wp, bp = w.lin[key]
        x2 = torch.zeros(B * L * L, 64, dtype=torch.bfloat16, device=x.device)
        T.gemm(x.view(-1, 64), wp, bias=bp, out=x2)
        x = T.conv3x3(x2.view(B, L, L, 64), *w.conv3x3('decoder.conv_in', 64))
        x = self._mid('decoder.mid_block', x)
        n = len(cfg.block_out_channels)
        for i in range(n):
            for j in range(cfg.layers_per_block + 1):
                x = self._resnet(f'decoder.up_blocks.{i}.resnets.{j}', x)
            if i < n - 1:
                x = T.conv3x3(T.upsample2x(x), *w.conv3x3(f'decoder.up_blocks.{i}.upsamplers.0.conv'))
        x = T.groupnorm(x, *w.gn('decoder.conv_norm_out'), cfg.norm_groups, 1e-6, silu=True)
        wo, bo = w.conv3x3('decoder.conv_out')
        return T.conv3x3(x, wo, bias=bo + bias_shift if bias_shift else bo, alpha=alpha)

    def decode(self, z, return_dict=False):
        """AutoencoderKL.decode: z [B,4,L,L] (already / scaling_factor) -> ([B,3,8L,8L] bf16 NCHW view,)  (return_dict=False form,
        the only one the reference uses)."""
        out = self._decode_nhwc(z).permute(0, 3, 1, 2)
        return (out,) if not return_dict else SimpleNamespace(sample=out)

    def decode_images(self, pred_x0):
        """The reference's decode tail in one go (mvedit_3d_pipeline.py:1258-1263):
        (vae.decode(x0 / scaling_factor) / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).float() -> [B,H,W,3] fp32.
        The affine map rides in conv_out's epilogue ((acc + bias + 1) * 0.5); the output is born NHWC, so the permute is free."""
        out = self._decode_nhwc(pred_x0 / self.config.scaling_factor, alpha=0.5, bias_shift=1.0)
        return out.float().clamp_(0, 1)

    # ------------------------------------------------------------------ encode
    def _down(self, p, x):
        key = p + '.__s2_ohwi__'
        if key not in self.w.lin:
            wt = self.w.sd[p + '.weight']
            self.w.lin[key] = (_bf(wt.float().permute(0, 2, 3, 1).reshape(wt.shape[0], -1), self.device), _f32(self.w.sd[p + '.bias'], self.device))
        wd, bd = self.w.lin[key]
        B, H, W_, C = x.shape
        return T.gemm(T.im2col3x3s2(x, pad_lo=0), wd, bias=bd).view(B, H // 2, W_ // 2, -1)

    def encode(self, x, return_dict=True):
        """AutoencoderKL.encode: x [B,3,H,W] in [-1,1] -> object with ``.latent_dist`` (``.sample()``, ``.mean``), fp32 [B,4,H/8,W/8]."""
        cfg, w = self.config, self.w
        B = x.shape[0]
        h = T.nchw_to_nhwc_pad(x, 64)
        h = T.conv3x3(h, *w.conv3x3('encoder.conv_in', 64))
        n = len(cfg.block_out_channels)
        for i in range(n):
            for j in range(cfg.layers_per_block):
                h = self._resnet(f'encoder.down_blocks.{i}.resnets.{j}', h)
            if i < n - 1:
                h = self._down(f'encoder.down_blocks.{i}.downsamplers.0.conv', h)
        h = self._mid('encoder.mid_block', h)
        h = T.groupnorm(h, *w.gn('encoder.conv_norm_out'), cfg.norm_groups, 1e-6, silu=True)
        key = 'encoder.conv_out#quant'
        if key not in w.conv3:
            wc, bc = w.sd['encoder.conv_out.weight'].float(), w.sd['encoder.conv_out.bias'].float()
            wq = w.sd['quant_conv.weight'].float().reshape(wc.shape[0], wc.shape[0])
            wf = torch.einsum('oi,ichw->ochw', wq, wc)                       # quant_conv (1x1) folded into conv_out (3x3)
            bf = wq @ bc + w.sd['quant_conv.bias'].float()
            w.conv3[key] = (_bf(wf.permute(0, 2, 3, 1), self.device), _f32(bf, self.device))
        wf, bf = w.conv3[key]
        m = T.conv3x3(h, wf, bias=bf).permute(0, 3, 1, 2).float()           # [B, 8, L, L]
        mean, logvar = m.chunk(2, dim=1)
        return SimpleNamespace(latent_dist=DiagonalGaussian(mean.contiguous(), logvar.contiguous()))
