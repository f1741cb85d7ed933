"""``Tonemapping`` -- the reference's tone curve module (/root/reference/lib/models/decoders/tonemapping.py:5-52), which the runner
always hands to the pipelines (lib/apis/adapter3d.py:88,780: ``tonemapping=self.tonemapping``).

Same constructor, buffers (``lut_x``: 16 log2-exposure knots in [-9, 3]; ``lut_y = smooth_forward(lut_x)``) and methods.  The
pipelines use it to apply Lambert shading in tone-mapped space, ``lut(inverse_lut(albedo) + log2(shading))``
(mvedit_3d_pipeline.py:418-422,564-570,1377-1384); on the B200 path that expression lives INSIDE the fused kernels
(mve_nerf_patch_loss / mve_nerf_patch_out_rgb / mve_shade_views take the knots by value: ``knots()``), the torch methods here serve
the host-side call sites and the tests.  Pinned against the reference module by tests/test_reference_pins.py."""
import ctypes

import torch
import torch.nn as nn


class Tonemapping(nn.Module):
    def __init__(self, exposure=0.0, contrast=0.953, bias=0.088, sigmoid_gain=0.943, log_gain=0.011, lut_logx_min=-9, lut_logx_max=3,
                 lut_steps=16):
        super().__init__()
        self.exposure, self.contrast, self.bias, self.sigmoid_gain, self.log_gain = exposure, contrast, bias, sigmoid_gain, log_gain
        self.register_buffer('lut_x', torch.linspace(lut_logx_min, lut_logx_max, lut_steps))
        self.register_buffer('lut_y', self.smooth_forward(self.lut_x))
        self._knots = None

    def smooth_forward(self, x, input_mode='log'):
        assert input_mode in ('log', 'linear')
        if input_mode == 'linear':
            x = x.clamp(min=1e-6).log2()
        x = (x + self.exposure) * self.contrast
        return x.sigmoid() * self.sigmoid_gain + x * self.log_gain + self.bias

    @staticmethod
    def _interp(v, a, b):
        i = torch.bucketize(v, a, right=True).clamp(min=1, max=len(a) - 1)
        return b[i - 1] + (b[i] - b[i - 1]) * ((v - a[i - 1]) / (a[i] - a[i - 1]))

    def lut(self, x, input_mode='log'):
        assert input_mode in ('log', 'linear')
        dtype = x.dtype
        x = x.to(self.lut_x.dtype)
        if input_mode == 'linear':
            x = x.clamp(min=1e-6).log2()
        return self._interp(x, self.lut_x, self.lut_y).to(dtype)

    def inverse_lut(self, y, output_mode='log'):
        assert output_mode in ('log', 'linear')
        dtype = y.dtype
        x = self._interp(y.to(self.lut_y.dtype), self.lut_y, self.lut_x)
        return (torch.exp2(x) if output_mode == 'linear' else x).to(dtype)

    def knots(self):
        """(host float array [lut_x | lut_y], n) for the kernels' ``tonemap_knots, tonemap_n`` arguments (cached: one D2H copy)."""
        if self._knots is None:
            n = self.lut_x.numel()
            vals = torch.cat([self.lut_x.float().cpu(), self.lut_y.float().cpu()]).tolist()
            self._knots = ((ctypes.c_float * (2 * n))(*vals), n)
        return self._knots


def tone_args(tonemapping):
    """-> the two trailing kernel arguments (knots pointer or NULL, n)."""
    if tonemapping is None:
        return None, ctypes.c_uint32(0)
    arr, n = tonemapping.knots()
    return arr, ctypes.c_uint32(n)
