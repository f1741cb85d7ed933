"""Python bindings of the tcgen05 building blocks (GEMM, implicit-GEMM conv, attention, norms) in libmvedit_b200.so.

These are the kernels under mvedit_b200.unet (the B200 replacement for the diffusers UNet/ControlNet call that
Adapter3DMixin.get_noise_pred* makes, adapter3d_mixin.py:101-125).  bf16 storage, fp32 accumulation.
"""
import torch

import ctypes

from ._lib import call, ptr, raw_ptr, stream, c_int, c_u32, c_f32

ACT = {None: 0, 'none': 0, 'silu': 1, 'gelu': 2, 'geglu': 3, 'relu': 4, 'relu_gate': 5, 'prelu': 6}


def gemm(a, w, bias=None, row_bias=None, rows_per_group=0, residual=None, act=None, alpha=1.0, out=None, act_param=None):
    """out[M,N] = act(a[M,K] @ w[N,K]^T + bias + row_bias[row // rows_per_group]) * alpha + residual.   bf16 in/out.
    act='geglu': w / bias must be ordered by geglu_interleave(); out[M, N/2] = value * gelu(gate)."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1]
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N // 2 if act == 'geglu' else N, dtype=torch.bfloat16, device=a.device)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
    if row_bias is not None:
        assert row_bias.dtype == torch.float32 and row_bias.shape[-1] == N
    assert a.stride(1) == 1 and w.stride(1) == 1 and out.stride(1) == 1
    call('mve_gemm_bf16', raw_ptr(a), raw_ptr(w), raw_ptr(out), c_u32(M), c_u32(N), c_u32(K), c_u32(a.stride(0)), c_u32(w.stride(0)),
         c_u32(out.stride(0)), ptr(bias), raw_ptr(row_bias), c_u32(rows_per_group),
         c_u32(row_bias.stride(0) if row_bias is not None else 0), raw_ptr(residual),
         c_u32(residual.stride(0) if residual is not None else 0), c_int(ACT[act]), c_f32(alpha), ptr(act_param), stream(),
         _meta=dict(flops=2.0 * M * N * K, shape='gemm M%d N%d K%d%s' % (M, N, K, ' geglu' if act == 'geglu' else '')))
    return out


def geglu_interleave(w, b, tile=256):
    """Reorder a diffusers GEGLU projection ([2F, K] weight: rows 0..F-1 values, F..2F-1 gates) so that every `tile` rows hold
    [tile/2 values | the tile/2 matching gates] -- the layout mve_gemm_bf16's act=3 epilogue expects."""
    F = w.shape[0] // 2
    h = tile // 2
    assert w.shape[0] % tile == 0
    idx = torch.arange(F, device=w.device).view(-1, h)
    perm = torch.cat([idx, idx + F], dim=1).reshape(-1)
    return w[perm].contiguous(), (None if b is None else b[perm].contiguous())


def conv3x3(x, w, bias=None, row_bias=None, residual=None, act=None, alpha=1.0, out=None, act_param=None, split_k=False, family=None):
    """x [B,H,W,Cin] bf16 NHWC, w [Cout,3,3,Cin] bf16 -> [B,H,W,Cout] bf16.  stride 1, pad 1.
    row_bias [B,Cout] f32 is added per image (time embedding); residual [B,H,W,Cout] is added after scaling.
    split_k: allow the split-K path for few-tile, long-K problems (fp32 atomics: not bit-reproducible run to run)."""
    assert x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.is_contiguous() and w.is_contiguous()
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    assert w.shape[1:] == (3, 3, Cin)
    if out is None:
        out = torch.empty(B, H, W, Cout, dtype=torch.bfloat16, device=x.device)
    call('mve_conv3x3_bf16', ptr(x), ptr(w), ptr(out), c_u32(B), c_u32(H), c_u32(W), c_u32(Cin), c_u32(Cout), c_u32(out.stride(2)),
         ptr(bias), raw_ptr(row_bias), c_u32(row_bias.stride(0) if row_bias is not None else 0), ptr(residual),
         c_u32(residual.stride(2) if residual is not None else 0), c_int(ACT[act] | (0x100 if split_k else 0)),
         c_f32(alpha), ptr(act_param), stream(), _meta=dict(flops=2.0 * B * H * W * Cout * 9 * Cin, family=family,
                                                            shape='conv B%d %dx%d Cin%d Cout%d' % (B, H, W, Cin, Cout)))
    return out


def attention(q, k, v, heads, scale=None, out=None):
    """q [B, Sq, heads*d] (row stride may exceed heads*d: slices of a fused projection), k/v [B, Skv, heads*d] -> [B, Sq, heads*d]."""
    B, Sq, C = q.shape
    Skv = k.shape[1]
    d = C // heads
    assert q.dtype == torch.bfloat16 and q.stride(2) == 1 and k.stride(2) == 1 and v.stride(2) == 1
    assert q.stride(0) == Sq * q.stride(1) and k.stride(0) == Skv * k.stride(1) and v.stride(0) == Skv * v.stride(1)
    if out is None:
        out = torch.empty(B, Sq, C, dtype=torch.bfloat16, device=q.device)
    call('mve_attention_bf16', raw_ptr(q), raw_ptr(k), raw_ptr(v), ptr(out), c_u32(B), c_u32(heads), c_u32(Sq), c_u32(Skv), c_u32(d),
         c_u32(q.stride(1)), c_u32(k.stride(1)), c_u32(v.stride(1)), c_u32(out.stride(1)), c_f32(scale if scale is not None else d ** -0.5),
         stream(), _meta=dict(flops=4.0 * B * heads * Sq * Skv * d, shape='attn B%d h%d Sq%d Skv%d d%d' % (B, heads, Sq, Skv, d)))
    return out


_gn_scratch = {}


def groupnorm(x, gamma, beta, groups=32, eps=1e-5, silu=False, out=None):
    """x [B, ..., C] bf16 channels-last -> same shape."""
    B, C = x.shape[0], x.shape[-1]
    HW = x.numel() // (B * C)
    assert x.is_contiguous() and x.dtype == torch.bfloat16
    if out is None:
        out = torch.empty_like(x)
    key = (x.device, B * groups)
    if key not in _gn_scratch:
        _gn_scratch[key] = torch.empty(B * groups * 2, dtype=torch.float32, device=x.device)
    call('mve_groupnorm_bf16', ptr(x), ptr(out), c_u32(B), c_u32(HW), c_u32(C), c_u32(groups), ptr(gamma), ptr(beta), c_f32(eps),
         c_int(int(silu)), ptr(_gn_scratch[key]), stream())
    return out


def layernorm(x, gamma, beta, eps=1e-5, out=None):
    C = x.shape[-1]
    rows = x.numel() // C
    assert x.is_contiguous() and x.dtype == torch.bfloat16
    if out is None:
        out = torch.empty_like(x)
    call('mve_layernorm_bf16', ptr(x), ptr(out), c_u32(rows), c_u32(C), ptr(gamma), ptr(beta), c_f32(eps), stream())
    return out


def geglu(h, out=None):
    F2 = h.shape[-1]
    M = h.numel() // F2
    assert h.is_contiguous() and h.dtype == torch.bfloat16
    if out is None:
        out = torch.empty(*h.shape[:-1], F2 // 2, dtype=torch.bfloat16, device=h.device)
    call('mve_geglu_bf16', ptr(h), ptr(out), ctypes.c_uint64(M), c_u32(F2 // 2), stream())
    return out


def upsample2x(x):
    B, H, W, C = x.shape
    out = torch.empty(B, 2 * H, 2 * W, C, dtype=torch.bfloat16, device=x.device)
    call('mve_upsample2x_bf16', ptr(x), ptr(out), c_u32(B), c_u32(H), c_u32(W), c_u32(C), stream())
    return out


def im2col3x3s2(x, pad_lo=1):
    B, H, W, C = x.shape
    out = torch.empty(B * (H // 2) * (W // 2), 9 * C, dtype=torch.bfloat16, device=x.device)
    call('mve_im2col3x3s2_bf16', ptr(x), ptr(out), c_u32(B), c_u32(H), c_u32(W), c_u32(C), c_int(pad_lo), stream())
    return out


def softmax_rows(x, scale=1.0, out=None):
    """x [rows, cols] bf16 (row stride may exceed cols) -> softmax(scale * x) per row; in place when out is None."""
    assert x.dim() == 2 and x.dtype == torch.bfloat16 and x.stride(1) == 1
    out = x if out is None else out
    call('mve_softmax_rows_bf16', raw_ptr(x), raw_ptr(out), c_u32(x.shape[0]), c_u32(x.shape[1]), c_u32(x.stride(0)), c_u32(out.stride(0)),
         c_f32(scale), stream())
    return out


def nchw_to_nhwc_pad(x, cpad):
    """x [B,C,H,W] f32/bf16 -> [B,H,W,cpad] bf16 (zero channel padding)."""
    B, C, H, W = x.shape
    x = x.contiguous()
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    out = torch.empty(B, H, W, cpad, dtype=torch.bfloat16, device=x.device)
    call('mve_nchw_to_nhwc_pad_bf16', ptr(x), c_int(int(x.dtype == torch.float32)), ptr(out), c_u32(B), c_u32(C), c_u32(H * W), c_u32(cpad),
         stream())
    return out


DIRECT_CONV_CONFIGS = {(3, 16, 1), (16, 16, 1), (16, 32, 2), (32, 32, 1), (32, 96, 2), (3, 8, 1), (8, 8, 1), (8, 16, 2), (16, 32, 1)}


def pack_direct_weight(w):
    """[Cout, Cin, 3, 3] -> f32 [9, Cin, Cout] (tap = ky * 3 + kx), the layout mve_conv3x3_direct_bf16 stages into shared memory."""
    return w.float().permute(2, 3, 1, 0).reshape(9, w.shape[1], w.shape[0]).contiguous()


def conv3x3_direct(x, w_packed, bias, cin, cout, stride=1, act=None, nchw=False, out_channels=None):
    """Few-channel 3x3 convolution on the CUDA cores (pad 1).  x: NHWC bf16 [B,H,W,cin], or with nchw=True an NCHW f32 / bf16 image
    [B,cin,H,W].  -> NHWC bf16 [B,H/stride,W/stride,out_channels or cout] (extra channels zero: the hand-off to a tensor-core layer)."""
    if nchw:
        B, _, H, W = x.shape
        assert x.dtype in (torch.float32, torch.bfloat16)
        fmt = 1 if x.dtype == torch.float32 else 2
    else:
        B, H, W, _ = x.shape
        assert x.dtype == torch.bfloat16
        fmt = 0
    x = x.contiguous()
    oc = out_channels or cout
    alloc = torch.zeros if oc != cout else torch.empty
    out = alloc(B, H // stride, W // stride, oc, dtype=torch.bfloat16, device=x.device)
    call('mve_conv3x3_direct_bf16', ptr(x), c_int(fmt), ptr(w_packed), ptr(bias), ptr(out), c_u32(B), c_u32(H), c_u32(W), c_u32(cin), c_u32(cout),
         c_u32(stride), c_u32(oc), c_int(ACT[act]), stream(), _meta=dict(flops=2.0 * B * (H // stride) * (W // stride) * cout * 9 * cin,
                                                                        shape='direct conv B%d %dx%d Cin%d Cout%d s%d' % (B, H, W, cin, cout, stride)))
    return out
