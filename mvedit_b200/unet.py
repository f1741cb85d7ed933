"""SD-1.5 UNet / ControlNet v1.1 on the tcgen05 kernels of libmvedit_b200 -- the B200 replacement for the diffusers modules the
reference instantiates (``UNet2DConditionModel`` / ``ControlNetModel``, /root/reference/lib/apis/adapter3d.py:162-180,
lib/pipelines/utils.py:235-240) and splits at the skip connections (``unet_enc`` / ``unet_dec``,
/root/reference/lib/models/architecture/diffusers.py:57-164).

* Weights come from a diffusers-format state dict (same key names), re-laid once: conv OIHW -> OHWI (K index = tap*Cin + c) with
  the input channels zero-padded to a multiple of 64, q/k/v fused into one projection, all resnets' time-embedding projections
  fused into ONE GEMM per forward.
* Activations are bf16 NHWC: the [B*H*W, C] pixel matrix IS the transformer's token matrix, so no permutes exist on this path.
* Every matmul-shaped op is mve_gemm_bf16 / mve_conv3x3_bf16 (implicit GEMM) / mve_attention_bf16; bias, per-image time-embedding
  bias, SiLU, ControlNet scale and residual adds ride in the GEMM epilogues; GroupNorm/LayerNorm/GEGLU/upsample are single-pass
  vectorised kernels.  All N views (x CFG halves) go through as ONE batch -- the reference loops over ``diff_bs``-sized chunks.
"""
import math

import torch

from . import tc_ops as T


def _pad64(c):
    return (c + 63) // 64 * 64


def _bf(t, dev):
    return t.to(device=dev, dtype=torch.bfloat16).contiguous()


def _f32(t, dev):
    return t.to(device=dev, dtype=torch.float32).contiguous()


class _Weights:
    """Re-laid weights of one network, keyed by the diffusers module path."""

    def __init__(self, sd, device):
        self.dev = device
        self.sd = sd
        self.conv3 = {}
        self.lin = {}
        self.norm = {}

    def conv3x3(self, name, cin_pad=None):
        if name not in self.conv3:
            w = self.sd[name + '.weight']                      # [Cout, Cin, 3, 3]
            cout, cin = w.shape[:2]
            cp = _pad64(cin) if cin_pad is None else cin_pad
            wp = torch.zeros(cout, 3, 3, cp, dtype=torch.float32, device=w.device)
            wp[..., :cin] = w.float().permute(0, 2, 3, 1)
            self.conv3[name] = (_bf(wp, self.dev), _f32(self.sd[name + '.bias'], self.dev))
        return self.conv3[name]

    def linear(self, name):
        """Linear or 1x1 conv -> ([N,K] bf16, bias f32 | None)."""
        if name not in self.lin:
            w = self.sd[name + '.weight']
            w = w.reshape(w.shape[0], -1)
            b = self.sd.get(name + '.bias')
            self.lin[name] = (_bf(w, self.dev), None if b is None else _f32(b, self.dev))
        return self.lin[name]

    def geglu(self, name):
        """GEGLU projection re-ordered for the fused value * gelu(gate) epilogue (tc_ops.geglu_interleave)."""
        key = name + '#geglu'
        if key not in self.lin:
            w, b = self.linear(name)
            self.lin[key] = T.geglu_interleave(w, b)
        return self.lin[key]

    def fused(self, key, names):
        if key not in self.lin:
            ws = [self.sd[n + '.weight'].reshape(self.sd[n + '.weight'].shape[0], -1) for n in names]
            bs = [self.sd.get(n + '.bias') for n in names]
            b = None if bs[0] is None else _f32(torch.cat(bs), self.dev)
            self.lin[key] = (_bf(torch.cat(ws, dim=0), self.dev), b)
        return self.lin[key]

    def gn(self, name):
        if name not in self.norm:
            self.norm[name] = (_f32(self.sd[name + '.weight'], self.dev), _f32(self.sd[name + '.bias'], self.dev))
        return self.norm[name]


def timestep_embedding(t, dim, device):
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=device) / half
    emb = t.float()[:, None] * torch.exp(exponent)[None]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class _Net:
    """Shared encoder machinery of UNet and ControlNet."""

    def __init__(self, state_dict, cfg, device='cuda'):
        self.cfg = cfg
        self.device = torch.device(device)
        self.dtype = torch.bfloat16
        self.w = _Weights(state_dict, self.device)
        self._resnet_names = [k[:-len('.time_emb_proj.weight')] for k in state_dict if k.endswith('.time_emb_proj.weight')]
        self._temb_slices = {}
        off = 0
        for n in self._resnet_names:
            c = state_dict[n + '.time_emb_proj.weight'].shape[0]
            self._temb_slices[n] = (off, c)
            off += c
        self._temb_total = off
        self.ip_tokens, self.ip_scale, self.drop_tokens = 0, 1.0, 0

    # ---- IP-Adapter (lib/models/architecture/ip_adapter/ip_adapter.py:85-113)
    def transformer_paths(self):
        """Transformer blocks in diffusers ``attn_processors`` order (down, up, mid): block i owns processors 2i (attn1), 2i+1 (attn2)."""
        cfg, paths = self.cfg, []
        for i in range(len(cfg.block_out_channels)):
            if cfg.attn_levels[i]:
                paths += [f'down_blocks.{i}.attentions.{j}' for j in range(cfg.layers_per_block)]
        rev = list(reversed(cfg.attn_levels))
        for i in range(len(cfg.block_out_channels)):
            if rev[i]:
                paths += [f'up_blocks.{i}.attentions.{j}' for j in range(cfg.layers_per_block + 1)]
        return paths + ['mid_block.attentions.0']

    def set_ip_adapter(self, ip_state_dict, num_tokens=16, scale=1.0):
        """UNet side of IPAdapter.set_ip_adapter: every cross-attention becomes IPAttnProcessor2_0 (attention_processor.py:283-396).
        ``ip_state_dict`` is the ``ip_adapter`` section of an IP-Adapter checkpoint ("{2i+1}.to_k_ip.weight", "{2i+1}.to_v_ip.weight");
        to_k_ip / to_v_ip are fused into one projection per block."""
        for i, p in enumerate(self.transformer_paths()):
            wk, wv = ip_state_dict[f'{2 * i + 1}.to_k_ip.weight'], ip_state_dict[f'{2 * i + 1}.to_v_ip.weight']
            self.w.lin[p + '#ip_kv'] = (_bf(torch.cat([wk, wv], dim=0), self.device), None)
        self.ip_tokens, self.ip_scale = int(num_tokens), float(scale)

    def set_cn_attn_processor(self, num_tokens=4):
        """ControlNet side (CNAttnProcessor2_0, attention_processor.py:466-556): cross-attention sees context[:, :-num_tokens].
        The reference instantiates it with the DEFAULT num_tokens = 4 although the plus adapter appends 16 image tokens
        (ip_adapter.py:104-110): 12 image tokens stay in the ControlNets' context.  Quirk kept for parity."""
        self.drop_tokens = int(num_tokens)

    # ---- time embedding: emb [1|B, 4*C0] bf16 and ALL resnets' projections in one GEMM -> f32 [B, total]
    def time_embed(self, t, batch):
        cfg = self.cfg
        t = torch.as_tensor(t, device=self.device).reshape(-1)
        e = timestep_embedding(t, cfg.block_out_channels[0], self.device).to(torch.bfloat16)
        w1, b1 = self.w.linear('time_embedding.linear_1')
        w2, b2 = self.w.linear('time_embedding.linear_2')
        emb = T.gemm(T.gemm(e, w1, bias=b1, act='silu'), w2, bias=b2)          # [len(t), temb]
        wt, bt = self.w.fused('__temb_all__', [n + '.time_emb_proj' for n in self._resnet_names])
        proj = T.gemm(torch.nn.functional.silu(emb.float()).to(torch.bfloat16), wt, bias=bt).float()   # [len(t), total]
        if proj.shape[0] != batch:
            proj = proj.expand(batch, -1).contiguous()
        return emb, proj

    def _temb(self, proj, name):
        off, c = self._temb_slices[name]
        return proj[:, off:off + c]

    # ---- blocks
    def resnet(self, p, x, proj):
        cfg, w = self.cfg, self.w
        g1, b1 = w.gn(p + '.norm1')
        h = T.groupnorm(x, g1, b1, cfg.norm_groups, 1e-5, silu=True)
        wc1, bc1 = w.conv3x3(p + '.conv1')
        h = T.conv3x3(h, wc1, bias=bc1, row_bias=self._temb(proj, p))
        g2, b2 = w.gn(p + '.norm2')
        h = T.groupnorm(h, g2, b2, cfg.norm_groups, 1e-5, silu=True)
        if (p + '.conv_shortcut.weight') in w.sd:
            ws, bs = w.linear(p + '.conv_shortcut')
            B, H, W_, C = x.shape
            sc = T.gemm(x.view(-1, C), ws, bias=bs).view(B, H, W_, -1)
        else:
            sc = x
        wc2, bc2 = w.conv3x3(p + '.conv2')
        return T.conv3x3(h, wc2, bias=bc2, residual=sc)

    def transformer(self, p, x, ctx, heads, n_imgs=1):
        """Transformer2DModel (depth 1).  x [B,H,W,C]; ctx [B,T,Dc] bf16.  n_imgs=2: CrossImageAttnProcWrapper view
        (joint_attn.py:13-33) -- consecutive (ref, view) images attend jointly, text context averaged over the pair."""
        cfg, w = self.cfg, self.w
        B, H, W_, C = x.shape
        S = H * W_
        gn, bn = w.gn(p + '.norm')
        h = T.groupnorm(x, gn, bn, cfg.norm_groups, 1e-6, silu=False)
        wi, bi = w.linear(p + '.proj_in')
        h = T.gemm(h.view(-1, C), wi, bias=bi)                                   # tokens [B*S, C]
        b = p + '.transformer_blocks.0'
        Bj, Sj = B // n_imgs, S * n_imgs
        # self-attention
        n1 = T.layernorm(h, *w.gn(b + '.norm1'))
        wqkv, _ = w.fused(b + '.attn1.qkv', [b + '.attn1.to_q', b + '.attn1.to_k', b + '.attn1.to_v'])
        qkv = T.gemm(n1, wqkv).view(Bj, Sj, 3 * C)
        o = T.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], heads)
        wo, bo = w.linear(b + '.attn1.to_out.0')
        h = T.gemm(o.view(-1, C), wo, bias=bo, residual=h)
        # cross-attention
        n2 = T.layernorm(h, *w.gn(b + '.norm2'))
        wq, _ = w.linear(b + '.attn2.to_q')
        q = T.gemm(n2, wq).view(Bj, Sj, C)
        if n_imgs > 1:
            ctx = ctx.view(Bj, n_imgs, *ctx.shape[1:]).float().mean(dim=1).to(torch.bfloat16)
        if self.drop_tokens:
            ctx = ctx[:, :ctx.shape[1] - self.drop_tokens]
        ctx_ip = None
        if self.ip_tokens and (p + '#ip_kv') in w.lin:
            ctx, ctx_ip = ctx[:, :ctx.shape[1] - self.ip_tokens], ctx[:, ctx.shape[1] - self.ip_tokens:]
        Tn = ctx.shape[1]
        wkv, _ = w.fused(b + '.attn2.kv', [b + '.attn2.to_k', b + '.attn2.to_v'])
        kv = T.gemm(ctx.reshape(-1, ctx.shape[-1]), wkv).view(Bj, Tn, 2 * C)
        o = T.attention(q, kv[:, :, :C], kv[:, :, C:], heads)
        if ctx_ip is not None:                      # IPAttnProcessor2_0: second attention over the image tokens, added with ``scale``
            kv_ip = T.gemm(ctx_ip.reshape(-1, ctx_ip.shape[-1]), w.lin[p + '#ip_kv'][0]).view(Bj, self.ip_tokens, 2 * C)
            o.add_(T.attention(q, kv_ip[:, :, :C], kv_ip[:, :, C:], heads), alpha=self.ip_scale)
        wo2, bo2 = w.linear(b + '.attn2.to_out.0')
        h = T.gemm(o.view(-1, C), wo2, bias=bo2, residual=h)
        # feed-forward (GEGLU)
        n3 = T.layernorm(h, *w.gn(b + '.norm3'))
        wf2, bf2 = w.linear(b + '.ff.net.2')
        if w.sd[b + '.ff.net.0.proj.weight'].shape[0] % 256 == 0:
            wf1, bf1 = w.geglu(b + '.ff.net.0.proj')          # GEGLU inside the projection's epilogue
            g = T.gemm(n3, wf1, bias=bf1, act='geglu')
        else:
            wf1, bf1 = w.linear(b + '.ff.net.0.proj')
            g = T.geglu(T.gemm(n3, wf1, bias=bf1))
        h = T.gemm(g, wf2, bias=bf2, residual=h)
        wp, bp = w.linear(p + '.proj_out')
        return T.gemm(h, wp, bias=bp, residual=x.view(-1, C)).view(B, H, W_, C)

    def downsample(self, p, x):
        """Downsample2D: 3x3 stride-2 pad-1 conv = im2col (one vectorised pass) + GEMM."""
        key = p + '.__s2_ohwi__'
        if key not in self.w.lin:
            wt = self.w.sd[p + '.weight']
            self.w.lin[key] = (_bf(wt.float().permute(0, 2, 3, 1).reshape(wt.shape[0], -1), self.device),
                               _f32(self.w.sd[p + '.bias'], self.device))
        wd, bd = self.w.lin[key]
        B, H, W_, C = x.shape
        cols = T.im2col3x3s2(x)
        return T.gemm(cols, wd, bias=bd).view(B, H // 2, W_ // 2, -1)

    def encoder(self, sample, proj, ctx, n_imgs=1):
        cfg = self.cfg
        res = [sample]
        for i in range(len(cfg.block_out_channels)):
            for j in range(cfg.layers_per_block):
                sample = self.resnet(f'down_blocks.{i}.resnets.{j}', sample, proj)
                if cfg.attn_levels[i]:
                    sample = self.transformer(f'down_blocks.{i}.attentions.{j}', sample, ctx, cfg.num_heads[i], n_imgs)
                res.append(sample)
            if i < len(cfg.block_out_channels) - 1:
                sample = self.downsample(f'down_blocks.{i}.downsamplers.0.conv', sample)
                res.append(sample)
        return res, sample

    def mid(self, sample, proj, ctx, n_imgs=1):
        sample = self.resnet('mid_block.resnets.0', sample, proj)
        sample = self.transformer('mid_block.attentions.0', sample, ctx, self.cfg.num_heads[-1], n_imgs)
        return self.resnet('mid_block.resnets.1', sample, proj)

    def conv_in(self, x_nchw):
        x = T.nchw_to_nhwc_pad(x_nchw, 64)
        wc, bc = self.w.conv3x3('conv_in', 64)
        return x, wc, bc


class UNet(_Net):
    """UNet2DConditionModel forward, split as unet_enc / unet_dec (diffusers.py:57-164)."""

    def enc(self, sample, t, encoder_hidden_states, cross_attention_kwargs=None):
        """sample [B,4,L,L] (any float dtype, NCHW as the pipeline holds it) -> (emb pack, skips (NHWC bf16), sample)."""
        n = (cross_attention_kwargs or {}).get('num_cross_attn_imgs', 1)
        B = sample.shape[0]
        emb, proj = self.time_embed(t, B)
        x, wc, bc = self.conv_in(sample)
        x = T.conv3x3(x, wc, bias=bc)
        ctx = encoder_hidden_states.to(torch.bfloat16).contiguous()
        res, s = self.encoder(x, proj, ctx, n)
        return (emb, proj), tuple(res), s

    def dec(self, emb, down_block_res_samples, sample, encoder_hidden_states, cross_attention_kwargs=None,
            down_block_additional_residuals=None, mid_block_additional_residual=None):
        cfg = self.cfg
        n = (cross_attention_kwargs or {}).get('num_cross_attn_imgs', 1)
        _, proj = emb
        ctx = encoder_hidden_states.to(torch.bfloat16).contiguous()
        is_cn = mid_block_additional_residual is not None and down_block_additional_residuals is not None
        res = list(down_block_res_samples)
        if is_cn:
            res = [a + b for a, b in zip(res, down_block_additional_residuals)]
        sample = self.mid(sample, proj, ctx, n)
        if is_cn:
            sample = sample + mid_block_additional_residual
        rev_attn, rev_heads = list(reversed(cfg.attn_levels)), list(reversed(cfg.num_heads))
        nlev = len(cfg.block_out_channels)
        for i in range(nlev):
            for j in range(cfg.layers_per_block + 1):
                sample = torch.cat([sample, res.pop()], dim=-1)
                sample = self.resnet(f'up_blocks.{i}.resnets.{j}', sample, proj)
                if rev_attn[i]:
                    sample = self.transformer(f'up_blocks.{i}.attentions.{j}', sample, ctx, rev_heads[i], n)
            if i < nlev - 1:
                wu, bu = self.w.conv3x3(f'up_blocks.{i}.upsamplers.0.conv')
                sample = T.conv3x3(T.upsample2x(sample), wu, bias=bu)
        g, b = self.w.gn('conv_norm_out')
        sample = T.groupnorm(sample, g, b, cfg.norm_groups, 1e-5, silu=True)
        wo, bo = self.w.conv3x3('conv_out')
        out = T.conv3x3(sample, wo, bias=bo)                     # [B,L,L,4]
        return out.permute(0, 3, 1, 2)                           # NCHW view, as the pipeline expects

    def __call__(self, sample, t, encoder_hidden_states, cross_attention_kwargs=None, down_block_additional_residuals=None,
                 mid_block_additional_residual=None):
        emb, res, s = self.enc(sample, t, encoder_hidden_states, cross_attention_kwargs)
        return self.dec(emb, res, s, encoder_hidden_states, cross_attention_kwargs, down_block_additional_residuals,
                        mid_block_additional_residual)


class ControlNet(_Net):
    """ControlNetModel (v1.1) forward, guess_mode=False.  Residuals come back NHWC bf16, already scaled; ``accumulate`` lets a
    second net add into the first one's buffers in its zero-conv epilogues (MultiControlNetModel's sum)."""

    def cond_embedding(self, cond, base, cond_repeat=1):
        """cond [B,3,8L,8L]; base = conv_in(sample) NHWC.  Returns base + controlnet_cond_embedding(cond).
        cond_repeat = r: the caller guarantees cond = r stacked copies of the same images (the pipeline feeds
        torch.cat([ctrl_images] * 2) for the CFG halves, mvedit_3d_pipeline.py:1417-1421): the 512^2 hint convolutions then run once
        on B/r images and only the last (64^2) convolution sees all B."""
        w, ce = self.w, self.cfg.cond_embed_channels
        if cond_repeat > 1:
            cond = cond[:cond.shape[0] // cond_repeat]
        def conv(name, x, cin, cout, stride, act, residual=None):
            B, H, W_, cp = x.shape
            if stride == 1:
                wc, bc = w.conv3x3(name, cp)
                if residual is not None:
                    return T.conv3x3(x, wc, bias=bc, act=act, residual=residual)
                out = torch.zeros(B, H, W_, _pad64(cout), dtype=torch.bfloat16, device=x.device) if cout % 64 else \
                    torch.empty(B, H, W_, cout, dtype=torch.bfloat16, device=x.device)
                T.conv3x3(x, wc, bias=bc, act=act, out=out)
                return out
            key = name + '.__s2__'
            if key not in w.lin:
                wt = w.sd[name + '.weight']
                wp = torch.zeros(wt.shape[0], 3, 3, cp, device=wt.device)
                wp[..., :cin] = wt.float().permute(0, 2, 3, 1)
                w.lin[key] = (_bf(wp.reshape(wt.shape[0], -1), self.device), _f32(w.sd[name + '.bias'], self.device))
            wd, bd = w.lin[key]
            cols = T.im2col3x3s2(x)
            out = torch.zeros(B * (H // 2) * (W_ // 2), _pad64(cout), dtype=torch.bfloat16, device=x.device) if cout % 64 else \
                torch.empty(B * (H // 2) * (W_ // 2), cout, dtype=torch.bfloat16, device=x.device)
            T.gemm(cols, wd, bias=bd, act=act, out=out)
            return out.view(B, H // 2, W_ // 2, -1)

        # layer list of ControlNetConditioningEmbedding: conv_in, then (same-width stride 1, widening stride 2) pairs
        layers = [('controlnet_cond_embedding.conv_in', 3, ce[0], 1)]
        k = 0
        for a, b in zip(ce[:-1], ce[1:]):
            layers += [(f'controlnet_cond_embedding.blocks.{k}', a, a, 1), (f'controlnet_cond_embedding.blocks.{k + 1}', a, b, 2)]
            k += 2
        # the few-channel front (3->16->16->32->32->96 on 512^2 ... 128^2 images) runs on the CUDA cores at its true channel counts
        # (mve_conv3x3_direct_bf16); from the first layer that is wide enough on, the tensor-core path takes over on 64-padded channels
        h, direct, first = cond, True, True
        for idx, (name, cin, cout, stride) in enumerate(layers):
            if direct and (cin, cout, stride) in T.DIRECT_CONV_CONFIGS:
                key = name + '#direct'
                if key not in w.lin:
                    w.lin[key] = (T.pack_direct_weight(w.sd[name + '.weight']).to(self.device), _f32(w.sd[name + '.bias'], self.device))
                wp, bp = w.lin[key]
                nxt = layers[idx + 1] if idx + 1 < len(layers) else None
                hand_off = nxt is None or (nxt[1], nxt[2], nxt[3]) not in T.DIRECT_CONV_CONFIGS
                h = T.conv3x3_direct(h, wp, bp, cin, cout, stride, act='silu', nchw=first, out_channels=_pad64(cout) if hand_off else None)
            else:
                if first:
                    h = T.nchw_to_nhwc_pad(cond, 64)
                direct = False
                h = conv(name, h, cin, cout, stride, 'silu')
            first = False
        if cond_repeat > 1:
            h = h.repeat(cond_repeat, 1, 1, 1)
        return conv('controlnet_cond_embedding.conv_out', h, ce[-1], self.cfg.block_out_channels[0], 1, None, residual=base)

    def __call__(self, sample, t, encoder_hidden_states, controlnet_cond, conditioning_scale=1.0, accumulate=None, cond_repeat=1):
        B = sample.shape[0]
        emb, proj = self.time_embed(t, B)
        x, wc, bc = self.conv_in(sample)
        x = T.conv3x3(x, wc, bias=bc)
        x = self.cond_embedding(controlnet_cond, x, cond_repeat)
        ctx = encoder_hidden_states.to(torch.bfloat16).contiguous()
        res, s = self.encoder(x, proj, ctx)
        s = self.mid(s, proj, ctx)
        acc_down, acc_mid = accumulate if accumulate is not None else ([None] * len(res), None)
        down = []
        for i, r in enumerate(res):
            wz, bz = self.w.linear(f'controlnet_down_blocks.{i}')
            Bn, H, W_, C = r.shape
            a = acc_down[i]
            down.append(T.gemm(r.view(-1, C), wz, bias=bz, alpha=float(conditioning_scale),
                               residual=None if a is None else a.view(-1, C)).view(Bn, H, W_, C))
        wz, bz = self.w.linear('controlnet_mid_block')
        Bn, H, W_, C = s.shape
        mid = T.gemm(s.view(-1, C), wz, bias=bz, alpha=float(conditioning_scale),
                     residual=None if acc_mid is None else acc_mid.view(-1, C)).view(Bn, H, W_, C)
        return down, mid


class MultiControlNet:
    """diffusers MultiControlNetModel: the nets' residuals summed -- here by chaining the zero-conv epilogues."""

    def __init__(self, nets):
        self.nets = list(nets)

    def __call__(self, sample, t, encoder_hidden_states, controlnet_cond, conditioning_scale, cond_repeat=1):
        acc = None
        for net, cond, sc in zip(self.nets, controlnet_cond, conditioning_scale):
            acc = net(sample, t, encoder_hidden_states, cond, sc, accumulate=acc, cond_repeat=cond_repeat)
        return acc
