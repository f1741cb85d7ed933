"""ncu helper: the UNet's largest self-attention (64 images x 8 heads x 4096 tokens x d 40) once between cudaProfilerStart/Stop."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvedit_b200 import tc_ops

torch.manual_seed(0)
B, S, H, D = int(os.environ.get('PA_B', 64)), 4096, 8, 40
qkv = (torch.randn(B, S, 3 * H * D, device='cuda') * 0.5).bfloat16()
q, k, v = qkv[..., :H * D], qkv[..., H * D:2 * H * D], qkv[..., 2 * H * D:]
for _ in range(2):
    o = tc_ops.attention(q, k, v, H)
torch.cuda.synchronize()
torch.cuda.profiler.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); o = tc_ops.attention(q, k, v, H); e1.record(); torch.cuda.synchronize()
torch.cuda.profiler.stop()
ms = e0.elapsed_time(e1)
print('attention ms', ms, 'TFLOP/s (d=40)', 4.0 * B * H * S * S * D / ms / 1e9)
