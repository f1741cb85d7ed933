"""ncu helper: field forward (fp32 + TF32 density) and backward on 300k clustered samples."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvedit_b200.ingp_decoder import iNGPDecoder

torch.manual_seed(0)
dec = iNGPDecoder(max_steps=1024).cuda()
with torch.no_grad():
    dec.encoder.params.uniform_(-0.3, 0.3)
M = 300000
# samples along 3000 rays through a ball of radius 0.5 (100 consecutive samples each, dt = 0.0034)
o = torch.randn(3000, 3, device='cuda'); o = o / o.norm(dim=-1, keepdim=True) * 0.5
d = -o / 0.5 + 0.3 * torch.randn(3000, 3, device='cuda'); d = d / d.norm(dim=-1, keepdim=True)
t = torch.arange(100, device='cuda') * 0.0034
xyz = (o[:, None] + d[:, None] * t[None, :, None]).reshape(-1, 3).clamp(-1, 1).contiguous()


def run():
    x = xyz.clone()
    s, c, _ = dec.point_decode([x], None, None)
    (s.sum() + c.sum()).backward()
    dec.point_density_decode([x], None)


run(); torch.cuda.synchronize()
torch.cuda.profiler.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('fwd+bwd+density ms', e0.elapsed_time(e1))
