"""Python bindings of the tcgen05 building blocks (GEMM, implicit-GEMM conv, attention, norms) in libmvedit_b200.so.

These are the kernels under mvedit_b200.unet (the B200 replacement for the diffusers UNet/ControlNet call that
Adapter3DMixin.get_noise_pred* makes, adapter3d_mixin.py:101-125).  bf16 storage, fp32 accumulation.
"""
import torch

from ._lib import call, ptr, stream, c_int, c_u32, c_f32

ACT = {None: 0, 'none': 0, 'silu': 1, 'gelu': 2}


def gemm(a, w, bias=None, row_bias=None, rows_per_group=0, residual=None, act=None, alpha=1.0, out=None):
    """out[M,N] = act(a[M,K] @ w[N,K]^T + bias + row_bias[row // rows_per_group]) * alpha + residual.   bf16 in/out."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1]
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, dtype=torch.bfloat16, device=a.device)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
    if row_bias is not None:
        assert row_bias.dtype == torch.float32 and row_bias.shape[-1] == N
    call('mve_gemm_bf16', ptr(a), ptr(w), ptr(out), c_u32(M), c_u32(N), c_u32(K), c_u32(a.stride(0)), c_u32(w.stride(0)),
         c_u32(out.stride(0)), ptr(bias), ptr(row_bias), c_u32(rows_per_group), ptr(residual),
         c_u32(residual.stride(0) if residual is not None else 0), c_int(ACT[act]), c_f32(alpha), stream())
    return out


def conv3x3(x, w, bias=None, row_bias=None, residual=None, act=None, alpha=1.0, out=None):
    """x [B,H,W,Cin] bf16 NHWC, w [Cout,3,3,Cin] bf16 -> [B,H,W,Cout] bf16.  stride 1, pad 1.
    row_bias [B,Cout] f32 is added per image (time embedding); residual [B,H,W,Cout] is added after scaling."""
    assert x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.is_contiguous() and w.is_contiguous()
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    assert w.shape[1:] == (3, 3, Cin)
    if out is None:
        out = torch.empty(B, H, W, Cout, dtype=torch.bfloat16, device=x.device)
    call('mve_conv3x3_bf16', ptr(x), ptr(w), ptr(out), c_u32(B), c_u32(H), c_u32(W), c_u32(Cin), c_u32(Cout), c_u32(out.stride(2)),
         ptr(bias), ptr(row_bias), ptr(residual), c_u32(residual.stride(2) if residual is not None else 0), c_int(ACT[act]),
         c_f32(alpha), stream())
    return out
