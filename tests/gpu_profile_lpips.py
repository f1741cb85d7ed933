"""Where one LPIPS forward + gradient (1 patch of 128^2) spends its time.
  python tests/gpu_profile_lpips.py            -> CUDA-event time eager / as a graph replay, CUPTI per-kernel table
  ncu --set full --clock-control none --profile-from-start off -k regex:k_gemm_tc -o /tmp/lpips python tests/gpu_profile_lpips.py ncu"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mvedit_b200.lpips import LPIPS, random_lpips_state_dict  # noqa: E402

dev = torch.device('cuda')
m = LPIPS(random_lpips_state_dict(0, dev), dev)
g = torch.Generator(device=dev).manual_seed(0)
pred, tgt = torch.rand(1, 128, 128, 3, device=dev, generator=g), torch.rand(1, 128, 128, 3, device=dev, generator=g)
for _ in range(3):
    m.loss_and_grad(pred, tgt)
torch.cuda.synchronize()
if len(sys.argv) > 1 and sys.argv[1] == 'ncu':
    torch.cuda.profiler.start()
    m.loss_and_grad(pred, tgt)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)


def timed(fn, n=20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


print('eager        %.3f ms / call' % timed(lambda: m.loss_and_grad(pred, tgt)))
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    m.loss_and_grad(pred, tgt)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = m.loss_and_grad(pred, tgt)
torch.cuda.synchronize()
print('graph replay %.3f ms / call' % timed(graph.replay))
big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def cold():
    big.zero_()          # evict L2 (weights come from HBM, as inside a reconstruction iteration that streams ~150 MB in between)
    graph.replay()


t_flush = timed(lambda: big.zero_())
print('graph replay after an L2 flush %.3f ms / call' % (timed(cold) - t_flush))
from torch.profiler import profile, ProfilerActivity  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    graph.replay()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
print('kernels in launch order (us):')
for e in evs:
    print('  %8.1f  %s' % (e.device_time, e.name[:110]))
print('sum %.1f us over %d kernels' % (sum(e.device_time for e in evs), len(evs)))
