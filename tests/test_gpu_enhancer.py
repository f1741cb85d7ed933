"""SRVGGNetCompact (the render enhancer, image_space_ss.py:8-75) on the conv kernels vs oracle/enhancer_oracle.py (fp32; itself pinned
against the reference class on CPU).  bf16 activations through 34 convolutions: the network output (dominated by the nearest-upsampled
input) agrees to ~1e-2 relative L2, the learned residual alone to a few per cent."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def test_prelu_epilogue_and_pixel_shuffle():
    from mvedit_b200 import tc_ops as T
    from mvedit_b200._lib import call, ptr, stream, c_u32
    g = torch.Generator(device='cuda').manual_seed(0)
    x = torch.randn(2, 32, 32, 64, device='cuda', generator=g).to(torch.bfloat16)
    w = (torch.randn(48, 3, 3, 64, device='cuda', generator=g) * 0.06).to(torch.bfloat16)
    b, slope = torch.randn(48, device='cuda', generator=g) * 0.1, torch.rand(48, device='cuda', generator=g) * 0.5
    ref = F.prelu(F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), b, padding=1), slope)
    y = T.conv3x3(x, w, bias=b, act='prelu', act_param=slope)
    torch.testing.assert_close(y.float(), ref.permute(0, 2, 3, 1), rtol=2e-2, atol=2e-2)
    out = torch.empty(2, 3, 128, 128, dtype=torch.bfloat16, device='cuda')
    call('mve_pixel_shuffle_add_bf16', ptr(y), ptr(x), c_u32(2), c_u32(32), c_u32(32), c_u32(3), c_u32(4), c_u32(48), c_u32(64), ptr(out), stream())
    want = F.pixel_shuffle(y.float().permute(0, 3, 1, 2), 4) + F.interpolate(x[..., :3].float().permute(0, 3, 1, 2), scale_factor=4.0, mode='nearest')
    assert torch.equal(out, want.to(torch.bfloat16))


@pytest.mark.parametrize('num_conv,B,S', [(32, 2, 128), (4, 3, 64)])
def test_enhancer_vs_oracle(num_conv, B, S):
    from mvedit_b200.enhancer import SRVGGNetCompact, random_srvgg_state_dict
    from oracle.enhancer_oracle import srvgg_forward
    sd = random_srvgg_state_dict(1, num_conv=num_conv, device='cuda')
    g = torch.Generator(device='cuda').manual_seed(2)
    x = F.interpolate(torch.rand(B, 3, S // 8, S // 8, device='cuda', generator=g), size=(S, S), mode='bicubic').clamp(0, 1)
    ref = srvgg_forward(sd, x, num_conv)
    net = SRVGGNetCompact(sd, num_conv=num_conv)
    out = net(x.to(torch.bfloat16))
    assert out.shape == (B, 3, 4 * S, 4 * S) and out.dtype == torch.bfloat16
    base = F.interpolate(x, scale_factor=4.0, mode='nearest')
    assert float((ref - base).abs().mean()) > 1e-3                         # the network does something
    assert rel(out, ref) < 1.5e-2, rel(out, ref)
    assert rel(out.float() - base, ref - base) < 8e-2, rel(out.float() - base, ref - base)
