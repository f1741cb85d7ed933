"""``LPIPS`` -- the perceptual patch loss of the reconstruction objective on B200 kernels, forward and gradient in one call.

The reference: ``patch_loss = LPIPSLoss(net='vgg', loss_weight=1.2)`` (/root/reference/lib/pipelines/utils.py:231-232;
lib/models/losses/lpips_loss.py:14-43) wraps ``lpips.LPIPS(net='vgg')`` (lpips==0.1.4, run in bf16, :30) and is evaluated on the
rendered 128^2 patch against the target patch in EVERY reconstruction iteration (lib/pipelines/mvedit_3d_pipeline.py:611-617),
then differentiated by autograd back to the rendered pixels.

Here: VGG16 ``features`` up to relu5_3 over [pred ; target] as one batch -- 13 tcgen05 implicit-GEMM convolutions with the ReLU in
the epilogue (mve_conv3x3_bf16, act 4), 4 max pools -- five fused "normalise / difference / lin / spatial mean + gradient" kernels
(mve_lpips_layer), and the backward as 13 more convolutions with 180-degree-rotated, transposed weights whose epilogue applies the
ReLU gate of the layer below (act 5).  No autograd graph: ``loss_and_grad`` returns the loss and d loss / d pred directly, ~42
launches, all shapes static, so the whole thing is captured inside the iteration's CUDA graph (nerf.nerf_optim).

Weights: a ``lpips.LPIPS(net='vgg').state_dict()`` (keys ``net.slice{1..5}.{idx}.weight|bias`` with torchvision's vgg16.features
indices, ``lin{0..4}.model.1.weight``); the scaling-layer constants are lpips's.  Without network access the weights are random
(oracle/lpips_oracle.py::random_lpips_state_dict) and parity is against the fp32 restatement there -- the lpips package is not
installed, so that oracle is unpinned.
"""
import torch

from . import tc_ops as T
from ._lib import call, ptr, stream, c_u32

# (slice, torchvision vgg16.features index, Cin, Cout); a max pool precedes slices 2..5
VGG16_LAYERS = [(1, 0, 3, 64), (1, 2, 64, 64), (2, 5, 64, 128), (2, 7, 128, 128), (3, 10, 128, 256), (3, 12, 256, 256), (3, 14, 256, 256),
                (4, 17, 256, 512), (4, 19, 512, 512), (4, 21, 512, 512), (5, 24, 512, 512), (5, 26, 512, 512), (5, 28, 512, 512)]
LPIPS_CHANNELS = [64, 128, 256, 512, 512]


def random_lpips_state_dict(seed=0, device='cpu'):
    """Random weights with lpips.LPIPS(net='vgg')'s keys and shapes (no checkpoints offline; BASELINE.md mandates random init): He-normal
    convolutions, small positive biases, non-negative 'lin' weights."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    for sl, idx, cin, cout in VGG16_LAYERS:
        sd[f'net.slice{sl}.{idx}.weight'] = torch.randn(cout, cin, 3, 3, generator=g, device=device) * (2.0 / (9 * cin)) ** 0.5
        sd[f'net.slice{sl}.{idx}.bias'] = torch.randn(cout, generator=g, device=device) * 0.05 + 0.02
    for k, c in enumerate(LPIPS_CHANNELS):
        sd[f'lin{k}.model.1.weight'] = torch.rand(1, c, 1, 1, generator=g, device=device) * (2.0 / c)
    return sd


class LPIPS:
    def __init__(self, state_dict, device='cuda'):
        dev = torch.device(device)
        self.device = dev
        self.layers = []                 # per conv: dict(slice, w [Cout,3,3,Cin64] bf16, b f32, wT [Cin64,3,3,Cout] bf16)
        for sl, idx, cin, cout in VGG16_LAYERS:
            w = state_dict[f'net.slice{sl}.{idx}.weight'].to(dev, torch.float32)             # [Cout,Cin,3,3]
            b = state_dict[f'net.slice{sl}.{idx}.bias'].to(dev, torch.float32).contiguous()
            cin_p = max(cin, 64)
            wf = torch.zeros(cout, 3, 3, cin_p, device=dev)
            wf[..., :cin] = w.permute(0, 2, 3, 1)
            # input gradient of a stride-1 pad-1 correlation = correlation of the output gradient with the taps rotated by 180 degrees
            # and in / out channels swapped: wT[i, ky, kx, o] = w[o, i, 2 - ky, 2 - kx]; the 3-channel input is padded to 64 outputs
            wt = torch.zeros(cin_p, 3, 3, cout, device=dev)
            wt[:cin] = w.flip(2, 3).permute(1, 2, 3, 0)
            self.layers.append(dict(slice=sl, cin=cin_p, cout=cout, w=wf.to(torch.bfloat16).contiguous(), b=b,
                                    wT=wt.to(torch.bfloat16).contiguous()))
        self.lin = [state_dict[f'lin{k}.model.1.weight'].to(dev, torch.float32).reshape(-1).contiguous() for k in range(5)]
        assert [l.numel() for l in self.lin] == LPIPS_CHANNELS

    # ------------------------------------------------------------------ forward over [pred ; target]
    def _features(self, pred, target):
        """pred / target [P,h,w,3] fp32 in [0,1] -> (acts: output of every conv [2P,H,W,C] bf16, feats: the 5 slice outputs)."""
        P, h, w, _ = pred.shape
        x = torch.empty(2 * P, h, w, 64, dtype=torch.bfloat16, device=pred.device)
        call('mve_lpips_prep', ptr(pred), ptr(target), c_u32(P * h * w), ptr(x), stream())
        acts, feats, cur = [], [], x
        for k, L in enumerate(self.layers):
            if k > 0 and L['slice'] != self.layers[k - 1]['slice']:
                feats.append(cur)
                B, H, W, C = cur.shape
                pooled = torch.empty(B, H // 2, W // 2, C, dtype=torch.bfloat16, device=cur.device)
                call('mve_maxpool2x2_bf16', ptr(cur), c_u32(B), c_u32(H), c_u32(W), c_u32(C), ptr(pooled), stream())
                cur = pooled
            cur = T.conv3x3(cur, L['w'], bias=L['b'], act='relu', split_k=True, family='lpips')
            acts.append(cur)
        feats.append(cur)
        return x, acts, feats

    @torch.no_grad()
    def __call__(self, pred, target):
        """-> lpips distance per image [P] (fp32), as ``lpips.LPIPS(net='vgg')(pred * 2 - 1, target * 2 - 1).flatten()``."""
        return self.loss_and_grad(pred, target, None, 1.0, need_grad=False)[2]

    @torch.no_grad()
    def loss_and_grad(self, pred, target, weight=None, scale=1.0, need_grad=True):
        """``scale * mean_p(lpips_p * weight_p)`` and its gradient w.r.t. ``pred``.

        pred, target: [P,h,w,3] fp32 in [0,1] (h, w multiples of 16 that divide 128 or are multiples of 128: the conv kernel's
        tiling); weight [P] or None; scale: float or 0-dim device tensor; -> (loss 0-dim, g_pred [P,h,w,3] fp32 or None, lpips [P])."""
        P, h, w, _ = pred.shape
        assert pred.shape == target.shape and pred.dtype == torch.float32 and target.dtype == torch.float32
        assert h % 16 == 0 and w % 16 == 0, 'four 2x2 pools'
        pred, target = pred.contiguous(), target.contiguous()
        dev = pred.device
        wgt = torch.ones(P, device=dev) if weight is None else weight.to(dev, torch.float32)
        # ``scale`` may be a device scalar (a schedule value that must stay live inside a captured graph): tensor arithmetic only
        gscale = (wgt * scale / P).to(torch.float32).contiguous()
        x, acts, feats = self._features(pred, target)
        per_img = torch.zeros(P, dtype=torch.float32, device=dev)
        grads = []
        for f, lw in zip(feats, self.lin):
            _, H, W, C = f.shape
            g = torch.empty(P, H, W, C, dtype=torch.bfloat16, device=dev)
            call('mve_lpips_layer', ptr(f), c_u32(P), c_u32(H * W), c_u32(C), ptr(lw), ptr(gscale), ptr(per_img), ptr(g), stream())
            grads.append(g)
        loss = (per_img * gscale).sum()
        if not need_grad:
            return loss, None, per_img
        # ---- backward: from relu5_3 down; ``g`` is always the gradient w.r.t. a conv's PRE-activation (already ReLU-gated)
        g = grads[4]
        for k in range(len(self.layers) - 1, -1, -1):
            L = self.layers[k]
            first_of_slice = k == 0 or self.layers[k - 1]['slice'] != L['slice']
            if not first_of_slice:
                # d / d (previous conv's output), gated by that ReLU: its pre-activation gradient
                g = T.conv3x3(g, L['wT'], act='relu_gate', residual=acts[k - 1][:P], split_k=True, family='lpips')
            elif k > 0:
                gp = T.conv3x3(g, L['wT'], split_k=True, family='lpips')           # gradient w.r.t. the pooled features of the slice below
                below = feats[L['slice'] - 2]
                _, H, W, C = below.shape
                g = grads[L['slice'] - 2]
                call('mve_maxpool2x2_relu_backward_bf16', ptr(below), ptr(gp), c_u32(P), c_u32(H), c_u32(W), c_u32(C), ptr(g), stream())
            else:
                g64 = T.conv3x3(g, L['wT'], split_k=True, family='lpips')          # gradient of the normalised, padded input
                g_pred = torch.empty(P, h, w, 3, dtype=torch.float32, device=dev)
                call('mve_lpips_input_grad', ptr(g64), c_u32(P * h * w), ptr(g_pred), stream())
        return loss, g_pred, per_img


class LPIPSLoss:
    """``LPIPSLoss(net='vgg', loss_weight=1.2)`` (lib/models/losses/lpips_loss.py:14-43; built at lib/pipelines/utils.py:232) for
    ``BaseNeRF(patch_loss=...)``: ``mean(lpips(pred, target) * weight) * loss_weight``, inputs NCHW in [0,1] as the reference passes
    them.  ``loss_and_grad`` is what ``nerf_optim`` calls (NHWC patches, gradient returned instead of recorded)."""

    def __init__(self, state_dict, net='vgg', normalize_inputs=True, loss_weight=1.0, device='cuda'):
        assert net == 'vgg' and normalize_inputs, 'the reference builds LPIPSLoss(net="vgg") with normalised inputs'
        self.lpips = LPIPS(state_dict, device)
        self.loss_weight = loss_weight

    def loss_and_grad(self, pred_nhwc, target_nhwc, weight=None, scale=1.0):
        return self.lpips.loss_and_grad(pred_nhwc, target_nhwc, weight, scale * self.loss_weight)

    def __call__(self, pred, target, weight=None, avg_factor=None):
        assert avg_factor is None
        d = self.lpips(pred.permute(0, 2, 3, 1).float().contiguous(), target.permute(0, 2, 3, 1).float().contiguous())
        return ((d if weight is None else d * weight).mean() * self.loss_weight).to(pred.dtype)
