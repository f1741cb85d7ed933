"""fp32 restatement of SRVGGNetCompact.forward -- TEST INFRASTRUCTURE ONLY.

Follows /root/reference/lib/models/decoders/image_space_ss.py:8-75 (a Real-ESRGAN "compact" network: conv 3x3 -> PReLU, num_conv times
conv 3x3 -> PReLU, conv 3x3 to C r^2 channels, PixelShuffle(r), plus the nearest-upsampled input) as a function of the module's state
dict.  Pinned: tests/golden/make_reference_pins.py instantiates the REFERENCE class (cut out of its file by AST; the mmgen registry
decorator and the checkpoint loader are not needed) with seeded weights and stores input, weights and output; tests/test_reference_pins.py."""
import torch
import torch.nn.functional as F


def srvgg_forward(sd, x, num_conv, upscale=4):
    """sd: keys body.{2k}.weight|bias (convs), body.{2k+1}.weight (PReLU slopes); x [B,Cin,H,W] -> [B,Cout,rH,rW]."""
    out = x
    for k in range(num_conv + 2):
        out = F.conv2d(out, sd[f'body.{2 * k}.weight'].to(x), sd[f'body.{2 * k}.bias'].to(x), padding=1)
        if k <= num_conv:
            out = F.prelu(out, sd[f'body.{2 * k + 1}.weight'].to(x))
    return F.pixel_shuffle(out, upscale) + F.interpolate(x, scale_factor=float(upscale), mode='nearest')
