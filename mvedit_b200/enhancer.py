"""``SRVGGNetCompact`` -- the render enhancer of the loop body (/root/reference/lib/models/decoders/image_space_ss.py:8-75; built by
``init_mvedit`` as ``SRVGGNetCompact(3, 3, num_feat=64, num_conv=32, upscale=4, act_type='prelu')`` in bf16, lib/pipelines/utils.py:212-215)
on the B200 kernels.  While the NeRF is rendered below 512^2 (``render_size_p``: 128 up to 30 % of the schedule, 256 up to 60 %) every
render goes through it before the ControlNets see it (mvedit_3d_pipeline.py:1398-1401).

34 convolutions 3x3 (3 -> 64, 32 x 64 -> 64, 64 -> 48) run on the tcgen05 implicit-GEMM kernel with bias + PReLU in the epilogue
(``act = 6``), then one kernel does PixelShuffle(4) + the nearest-upsampled input (``mve_pixel_shuffle_add_bf16``): 37 launches for a
whole batch of views, activations bf16 NHWC.  Weights: the reference module's ``state_dict`` (``body.{2k}.weight|bias`` convolutions,
``body.{2k+1}.weight`` PReLU slopes).  Pinned: oracle/enhancer_oracle.py against the reference class itself (tests/test_reference_pins.py)."""
import torch

from . import tc_ops as T
from ._lib import call, ptr, stream, c_u32


def random_srvgg_state_dict(seed=0, num_in_ch=3, num_out_ch=3, num_feat=64, num_conv=32, upscale=4, device='cpu'):
    """Random weights with the reference module's keys / shapes (no checkpoint offline)."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd, cin = {}, num_in_ch
    for k in range(num_conv + 2):
        cout = num_feat if k <= num_conv else num_out_ch * upscale * upscale
        sd[f'body.{2 * k}.weight'] = torch.randn(cout, cin, 3, 3, generator=g, device=device) * (1.6 / (9 * cin)) ** 0.5
        sd[f'body.{2 * k}.bias'] = torch.randn(cout, generator=g, device=device) * 0.02
        if k <= num_conv:
            sd[f'body.{2 * k + 1}.weight'] = 0.1 + 0.3 * torch.rand(cout, generator=g, device=device)
        cin = cout
    return sd


class SRVGGNetCompact:
    def __init__(self, state_dict, num_in_ch=3, num_out_ch=3, num_feat=64, num_conv=16, upscale=4, act_type='prelu', device='cuda'):
        assert act_type == 'prelu' and num_feat % 64 == 0 and num_in_ch <= 64, 'the reference builds it with PReLU and 64 features'
        dev = torch.device(device)
        self.num_in_ch, self.num_out_ch, self.num_feat, self.num_conv, self.upscale, self.device = \
            num_in_ch, num_out_ch, num_feat, num_conv, upscale, dev
        self.layers = []
        for k in range(num_conv + 2):
            w = state_dict[f'body.{2 * k}.weight'].to(dev, torch.float32)                      # [Cout,Cin,3,3]
            cin_p = max(64, w.shape[1])
            wf = torch.zeros(w.shape[0], 3, 3, cin_p, device=dev)
            wf[..., :w.shape[1]] = w.permute(0, 2, 3, 1)
            slope = state_dict.get(f'body.{2 * k + 1}.weight') if k <= num_conv else None
            self.layers.append(dict(w=wf.to(torch.bfloat16).contiguous(), b=state_dict[f'body.{2 * k}.bias'].to(dev, torch.float32).contiguous(),
                                    slope=None if slope is None else slope.to(dev, torch.float32).contiguous()))

    def parameters(self):            # the pipeline reads next(self.image_enhancer.parameters()).dtype (mvedit_3d_pipeline.py:1017)
        return iter([self.layers[0]['w']])

    @torch.no_grad()
    def __call__(self, x):
        """x [B, 3, H, W] (any float dtype) -> [B, 3, upscale H, upscale W] in x's dtype."""
        B, C, H, W = x.shape
        assert C == self.num_in_ch
        cur = inp = T.nchw_to_nhwc_pad(x, 64)
        for L in self.layers:
            cur = T.conv3x3(cur, L['w'], bias=L['b'], act='prelu' if L['slope'] is not None else None, act_param=L['slope'])
        r = self.upscale
        out = torch.empty(B, self.num_out_ch, H * r, W * r, dtype=torch.bfloat16, device=x.device)
        call('mve_pixel_shuffle_add_bf16', ptr(cur), ptr(inp), c_u32(B), c_u32(H), c_u32(W), c_u32(self.num_out_ch), c_u32(r),
             c_u32(cur.shape[-1]), c_u32(inp.shape[-1]), ptr(out), stream())
        return out.to(x.dtype)
