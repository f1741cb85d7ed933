"""CPU / fp32 restatement of the perceptual patch loss -- TEST INFRASTRUCTURE ONLY (never imported by mvedit_b200).

What it restates: ``LPIPSLoss(net='vgg', loss_weight=1.2)`` (/root/reference/lib/models/losses/lpips_loss.py:14-43) ->
``lpips.LPIPS(net='vgg', eval_mode=True, pnet_tune=False)`` of the pip package lpips==0.1.4 (requirements; NOT installed in this
image, so this oracle is **parity unpinned**: it follows the package's published algorithm -- Zhang et al. 2018, "The Unreasonable
Effectiveness of Deep Features as a Perceptual Metric", and the package's lpips.py -- and is anchored on the reference's call site):

    in0, in1 in [-1, 1]  (LPIPSLoss.normalize_inputs: x * 2 - 1, lpips_loss.py:37-39)
    ScalingLayer:  (x - shift) / scale,  shift = (-.030, -.088, -.188), scale = (.458, .448, .450)
    VGG16 features, taps after relu1_2, relu2_2, relu3_3, relu4_3, relu5_3
    per tap: f / (sqrt(sum_c f^2) + 1e-10) for both images, squared difference, 1x1 conv 'lin' (no bias), spatial mean
    sum over the 5 taps -> one distance per image

and the reduction of the reference's ``weighted_loss`` (mmgen): ``mean(lpips * weight) * loss_weight`` (lpips_loss.py:8-10,40-43).
Autograd provides the gradient the kernels are checked against."""
import torch
import torch.nn.functional as F

# (slice, torchvision vgg16.features index, Cin, Cout); slices 2..5 start with a 2x2 max pool
VGG16_LAYERS = [(1, 0, 3, 64), (1, 2, 64, 64), (2, 5, 64, 128), (2, 7, 128, 128), (3, 10, 128, 256), (3, 12, 256, 256), (3, 14, 256, 256),
                (4, 17, 256, 512), (4, 19, 512, 512), (4, 21, 512, 512), (5, 24, 512, 512), (5, 26, 512, 512), (5, 28, 512, 512)]
CHANNELS = [64, 128, 256, 512, 512]
SHIFT = (-0.030, -0.088, -0.188)
SCALE = (0.458, 0.448, 0.450)


def random_lpips_state_dict(seed=0):
    """A state dict with lpips.LPIPS(net='vgg')'s keys and shapes; He-normal convolutions (activations keep O(1) scale through 13
    layers), small positive biases so that ReLUs are neither all dead nor all open, non-negative 'lin' weights (as trained)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for sl, idx, cin, cout in VGG16_LAYERS:
        sd[f'net.slice{sl}.{idx}.weight'] = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
        sd[f'net.slice{sl}.{idx}.bias'] = torch.randn(cout, generator=g) * 0.05 + 0.02
    for k, c in enumerate(CHANNELS):
        sd[f'lin{k}.model.1.weight'] = torch.rand(1, c, 1, 1, generator=g) * (2.0 / c)
    return sd


def features(sd, x):
    """x [B,3,H,W] in [-1,1] -> the five taps."""
    x = (x - x.new_tensor(SHIFT).view(1, 3, 1, 1)) / x.new_tensor(SCALE).view(1, 3, 1, 1)
    taps, prev = [], 1
    for sl, idx, _, _ in VGG16_LAYERS:
        if sl != prev:
            taps.append(x)
            x = F.max_pool2d(x, 2, 2)
            prev = sl
        x = F.relu(F.conv2d(x, sd[f'net.slice{sl}.{idx}.weight'].to(x), sd[f'net.slice{sl}.{idx}.bias'].to(x), padding=1))
    taps.append(x)
    return taps


def lpips(sd, in0, in1):
    """in0, in1 [B,3,H,W] in [-1,1] -> [B] distances."""
    total = 0
    for k, (a, b) in enumerate(zip(features(sd, in0), features(sd, in1))):
        na = a / (a.square().sum(dim=1, keepdim=True).sqrt() + 1e-10)
        nb = b / (b.square().sum(dim=1, keepdim=True).sqrt() + 1e-10)
        total = total + F.conv2d((na - nb) ** 2, sd[f'lin{k}.model.1.weight'].to(a)).mean(dim=(2, 3))
    return total.flatten()


def lpips_loss(sd, pred, target, weight=None, loss_weight=1.2):
    """LPIPSLoss.forward: pred / target [B,3,H,W] in [0,1] -> mean(lpips * weight) * loss_weight."""
    d = lpips(sd, pred * 2 - 1, target * 2 - 1)
    if weight is not None:
        d = d * weight
    return d.mean() * loss_weight
