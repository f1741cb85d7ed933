"""ncu helper (round 2 evidence): ONE launch of each shipped hot kernel at its benchmark shape between cudaProfilerStart/Stop.
  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_kernels python tests/gpu_profile_kernels.py
Also prints CUDA-event times of the same launches (never taken under ncu) when run plain."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mvedit_b200 import tc_ops as T  # noqa: E402

dev = torch.device('cuda')
g = torch.Generator(device=dev).manual_seed(0)
bf = lambda *s: (torch.randn(*s, device=dev, generator=g) * 0.5).to(torch.bfloat16)
jobs = []

# tcgen05 GEMM / conv (k_gemm_tc): the K = 320 transformer projection, a 1280-wide projection, the UNet 64^2 conv, the VAE 512^2 conv
a1, w1 = bf(262144, 320), bf(320, 320)
jobs.append(('gemm 262144x320x320 (K=320 family)', lambda: T.gemm(a1, w1)))
a2, w2 = bf(65536, 1280), bf(1280, 1280)
jobs.append(('gemm 65536x1280x1280', lambda: T.gemm(a2, w2)))
x3, w3 = bf(64, 64, 64, 320), bf(320, 3, 3, 320)
jobs.append(('conv3x3 64x64^2 320->320 (UNet)', lambda: T.conv3x3(x3, w3)))
x4, w4 = bf(8, 512, 512, 128), bf(128, 3, 3, 128)
jobs.append(('conv3x3 8x512^2 128->128 (VAE)', lambda: T.conv3x3(x4, w4)))
# attention (k_attention_pp): 8 images x 8 heads x 4096^2 x d=40
qkv = bf(8, 4096, 960)
jobs.append(('attention 8x8x4096^2 d40', lambda: T.attention(qkv[:, :, :320], qkv[:, :, 320:640], qkv[:, :, 640:], 8)))
# GroupNorm on a VAE activation
gn_g, gn_b = torch.ones(128, device=dev), torch.zeros(128, device=dev)
jobs.append(('groupnorm 8x512^2x128', lambda: T.groupnorm(x4, gn_g, gn_b, 32, 1e-6, silu=True)))
# field forward / backward on 1 M samples clustered along rays
from mvedit_b200.ingp_decoder import iNGPDecoder  # noqa: E402
dec = iNGPDecoder(max_steps=1024).to(dev)
with torch.no_grad():
    dec.encoder.params.uniform_(-0.3, 0.3)
o = torch.nn.functional.normalize(torch.randn(10000, 3, device=dev, generator=g), dim=-1) * 0.5
d = torch.nn.functional.normalize(-o / 0.5 + 0.3 * torch.randn(10000, 3, device=dev, generator=g), dim=-1)
xyz = (o[:, None] + d[:, None] * (torch.arange(100, device=dev) * 0.0034)[None, :, None]).reshape(-1, 3).clamp(-1, 1).contiguous()


def field_fb():
    s, c, _ = dec.point_decode([xyz.clone()], None, None)
    (s.sum() + c.sum()).backward()


jobs.append(('field fwd+bwd 1M samples', field_fb))
# 3DGS blend forward / backward: 2^18 Gaussians, one 512^2 view
from tests import synth  # noqa: E402
from mvedit_b200.gs_renderer import GaussianRasterizer, GaussianRasterizationSettings  # noqa: E402
P = 1 << 18
means = (torch.randn(P, 3, device=dev, generator=g) * 0.3).requires_grad_(True)
scales = torch.exp(torch.rand(P, 3, device=dev, generator=g) * 2 - 5)
quats, opac, cols = torch.randn(P, 4, device=dev, generator=g), torch.sigmoid(torch.randn(P, device=dev, generator=g)), torch.rand(P, 3, device=dev, generator=g)
pose = torch.from_numpy(synth.surround_poses(1, seed=0)).to(dev)[0]
f = 0.5 * 512 / math.tan(math.radians(15))
rast = GaussianRasterizer(GaussianRasterizationSettings(image_height=512, image_width=512, viewmatrix=torch.linalg.inv(pose), intrinsics=(f, f, 256, 256), bg=(1, 1, 1)))


def gs_fb():
    c, dd, al = rast(means, opac, cols, scales, quats)
    (c.sum() + al.sum()).backward()


jobs.append(('3DGS fwd+bwd 2^18 gaussians 512^2', gs_fb))

for name, fn in jobs:                       # warm-up (allocator, lazy module loads)
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.start()
times = []
for name, fn in jobs:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    times.append((name, e0.elapsed_time(e1)))
torch.cuda.profiler.stop()
for name, ms in times:
    print('%-44s %8.3f ms' % (name, ms))
