"""CPU oracle for the Instant-NGP field of the reference: multi-resolution hash grid
+ 2-layer MLP + activations.   TEST INFRASTRUCTURE ONLY (see oracle/raymarching_oracle.c header).

Follows:
  * the hash grid the reference instantiates through ``tinycudann.Encoding``
    (/root/reference/lib/models/decoders/ingp_decoder.py:62-74): HashGrid, n_levels, 2 features/level,
    log2_hashmap_size 19, base_resolution 16, Smoothstep interpolation, per_level_scale
    exp2(log2(max_res*bound/base_res)/(n_levels-1)), fp32 parameters.
    tiny-cuda-nn is an UN-VENDORED, UNPINNED dependency (requirements.txt:5: git HEAD) and is not installed
    here; its published algorithm (include/tiny-cuda-nn/encodings/grid.h: grid_scale, grid_resolution,
    pos_fract, grid_index / coherent-prime hash, kernel_grid) is restated below.   ==> PARITY UNPINNED:
    no tcnn output vector exists offline to check this restatement against.
  * iNGPDecoder.point_decode / density_blob / MLP (ingp_decoder.py:20-40,101-120) and TruncExp
    (/root/reference/lib/ops/activation.py:8-23), which ARE in the tree and are restated line by line.

Everything is plain differentiable torch, so autograd supplies the backward oracle
(d/d table, d/d MLP weights, d/d xyz).
"""
import math

import numpy as np
import torch

PRIMES = (1, 2654435761, 805459861)


def per_level_scale(max_resolution, bound, base_resolution, n_levels):
    # ingp_decoder.py:71
    return float(np.exp2(np.log2(max_resolution * bound / base_resolution) / (n_levels - 1)))


def level_table(n_levels=12, base_resolution=16, max_resolution=320, bound=1.0, log2_hashmap_size=19):
    """Per level: (scale, resolution, n_params_entries, offset_entries). grid.h: grid_scale / grid_resolution."""
    pls = per_level_scale(max_resolution, bound, base_resolution, n_levels)
    log2_pls = np.float32(math.log2(pls))
    out, offset = [], 0
    for l in range(n_levels):
        scale = np.float32(np.exp2(np.float32(l) * log2_pls, dtype=np.float32) * np.float32(base_resolution) - np.float32(1.0))
        res = int(math.ceil(float(scale))) + 1
        n = res ** 3
        n = (n + 7) // 8 * 8
        n = min(n, 1 << log2_hashmap_size)
        out.append((float(scale), res, n, offset))
        offset += n
    return out, offset


class TruncExpFn(torch.autograd.Function):
    """activation.py:8-23: forward exp(x); backward g * clamp(exp(x), 1e-6, 1e6)."""

    @staticmethod
    def forward(ctx, x):
        e = torch.exp(x)
        ctx.save_for_backward(e)
        return e

    @staticmethod
    def backward(ctx, g):
        return g * ctx.saved_tensors[0].clamp(min=1e-6, max=1e6)


def hash_encode(x01, table, levels):
    """x01 [M,3] in [0,1]; table [n_entries, 2]; -> [M, 2*L].  kernel_grid of grid.h with Smoothstep."""
    M = x01.shape[0]
    feats = []
    for (scale, res, n, off) in levels:
        pos = x01 * scale + 0.5
        g = torch.floor(pos)
        f = pos - g
        w = f * f * (3.0 - 2.0 * f)                      # smoothstep
        g = g.to(torch.int64)
        acc = x01.new_zeros(M, 2)
        for corner in range(8):
            wgt = x01.new_ones(M)
            idx_c = []
            for d in range(3):
                if corner & (1 << d):
                    wgt = wgt * w[:, d]
                    idx_c.append(g[:, d] + 1)
                else:
                    wgt = wgt * (1 - w[:, d])
                    idx_c.append(g[:, d])
            # grid_index
            stride, index = 1, torch.zeros(M, dtype=torch.int64, device=x01.device)
            for d in range(3):
                if stride <= n:
                    index = index + idx_c[d] * stride
                    stride *= res
            if n < stride:   # hashed level
                index = torch.zeros(M, dtype=torch.int64, device=x01.device)
                for d in range(3):
                    index = index ^ ((idx_c[d] * PRIMES[d]) & 0xFFFFFFFF)
            index = (index & 0xFFFFFFFF) % n
            acc = acc + wgt[:, None] * table[off + index]
        feats.append(acc)
    return torch.cat(feats, dim=-1)


def density_blob(x, blob_density=1.0, blob_radius=0.2):
    # ingp_decoder.py:101-104
    d = (x ** 2).sum(-1).clamp(min=0.2)
    return blob_density * torch.exp(-d / (2 * blob_radius ** 2))


def point_decode(xyz, table, w1, b1, w2, b2, levels, bound=1.0, sigmoid_saturation=0.001, blob_density=1.0, blob_radius=0.2):
    """iNGPDecoder.point_decode (ingp_decoder.py:106-120). xyz [M,3] in [-bound,bound] -> sigma [M], rgb [M,3]."""
    enc = hash_encode((xyz + bound) / (2 * bound), table, levels)
    h = torch.relu(enc @ w1.t() + b1) @ w2.t() + b2
    sigma = TruncExpFn.apply(h[..., 0] + density_blob(xyz, blob_density, blob_radius))
    rgb = torch.sigmoid(h[..., 1:])
    if sigmoid_saturation > 0:
        rgb = rgb * (1 + sigmoid_saturation * 2) - sigmoid_saturation
    return sigma, rgb


def init_params(levels, n_entries, hidden=64, seed=0, table_scale=1e-4, dtype=torch.float32):
    """iNGPDecoder.init_weights (ingp_decoder.py:87-91): table U(-1e-4,1e-4), Linear xavier-uniform, zero bias.
    table_scale can be raised for tests/benchmarks so densities are non-trivial."""
    g = torch.Generator().manual_seed(seed)
    in_dim = 2 * len(levels)
    table = (torch.rand(n_entries, 2, generator=g, dtype=dtype) * 2 - 1) * table_scale
    a1 = math.sqrt(6.0 / (in_dim + hidden))
    a2 = math.sqrt(6.0 / (hidden + 4))
    w1 = (torch.rand(hidden, in_dim, generator=g, dtype=dtype) * 2 - 1) * a1
    w2 = (torch.rand(4, hidden, generator=g, dtype=dtype) * 2 - 1) * a2
    return table, w1, torch.zeros(hidden, dtype=dtype), w2, torch.zeros(4, dtype=dtype)


def triplane_point_decode(xyz, code, sd, levels, plane_cfg=('xy', 'xz', 'yz'), flip_z=False, bound=1.0, sigmoid_saturation=0.001, interp_mode='bilinear'):
    """TriPlaneiNGPDecoder.point_decode (/root/reference/lib/models/decoders/triplane_ingp_decoder.py:142-212; xyz_transform:
    triplane_decoder.py:106-130) for one scene without view directions (MVEdit disables them, lib/apis/adapter3d.py:1362).
    xyz [M,3]; code (1,3,C,h,w); sd: state dict with the reference's keys (encoder.params, base_net.0.*, ingp_base_net.0.*,
    density_net.0.*, color_net.0.*).  SiLU activation, trunc_exp density, saturated sigmoid colour."""
    import torch.nn.functional as F
    enc = hash_encode((xyz + bound) / (2 * bound), sd['encoder.params'].view(-1, 2), levels)
    ax = dict(x=xyz[..., 0], y=xyz[..., 1], z=-xyz[..., 2] if flip_z else xyz[..., 2])
    grid = torch.stack([torch.stack([ax[a] for a in plane], dim=-1) for plane in plane_cfg], dim=0).unsqueeze(1)       # (3,1,M,2)
    C = code.shape[2]
    pc = F.grid_sample(code[0].float(), grid, mode=interp_mode, padding_mode='border', align_corners=False).squeeze(-2)  # (3,C,M)
    point_code = pc.permute(2, 1, 0).reshape(xyz.shape[0], C * 3)
    lin = lambda n, t: F.linear(t, sd[n + '.0.weight'], sd[n + '.0.bias'])
    base_x = lin('base_net', point_code) + lin('ingp_base_net', enc)
    act = F.silu(base_x)
    sigma = TruncExpFn.apply(lin('density_net', act).squeeze(-1))
    rgb = torch.sigmoid(lin('color_net', act))
    if sigmoid_saturation > 0:
        rgb = rgb * (1 + sigmoid_saturation * 2) - sigmoid_saturation
    return sigma, rgb
