"""ncu helper: fit a NeRF briefly, then ONE fused render launch (8 views x 512^2) and one density pre-pass between cudaProfilerStart/Stop."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import synth
from bench import synth_targets
from mvedit_b200.nerf import BaseNeRF, nerf_optim
from mvedit_b200.ingp_decoder import iNGPDecoder

torch.manual_seed(0)
V, IMG = int(os.environ.get('PR_VIEWS', 8)), 512
ITERS = int(os.environ.get('PR_ITERS', 300))
dev = torch.device('cuda')
poses = torch.from_numpy(synth.surround_poses(V, seed=0)).to(dev)
f = 0.5 * IMG / math.tan(math.radians(15))
K = torch.tensor([[f, f, IMG / 2, IMG / 2]] * V, device=dev)
img, msk = synth_targets(poses, K, IMG, dev)
nerf = BaseNeRF(grid_size=128, decoder=iNGPDecoder(max_steps=1024, weight_culling_th=0.001), patch_size=128).to(dev)
nerf.decoder.sample_capacity = 16384 * 160
grid, bits = nerf.get_init_density_grid(1, dev), nerf.get_init_density_bitfield(1, dev)
opt = torch.optim.Adam(nerf.decoder.parameters(), lr=0.01)
with torch.no_grad():
    nerf_optim(nerf, img[None], msk[None], None, opt, 0.01, ITERS, 16384, 0.0, 0.0, 0.02, 0.1, 0.01, None, grid, bits, IMG, K, IMG, poses,
               torch.ones(V, device=dev), torch.nn.functional.normalize(torch.randn(V, 3, device=dev), dim=-1), 128, True, 0.015, 0.2, 1.0, False)
    torch.cuda.synchronize()
    nerf.render(nerf.decoder, None, bits, IMG, IMG, K[None], poses[None], cfg=dict(dt_gamma_scale=0.25, return_rgba=True))
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nerf.render(nerf.decoder, None, bits, IMG, IMG, K[None], poses[None], cfg=dict(dt_gamma_scale=0.25, return_rgba=True))
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    st = nerf.decoder.last_render_stats()
    for rep in range(int(os.environ.get('PR_REPS', 0))):
        e0.record()
        nerf.render(nerf.decoder, None, bits, IMG, IMG, K[None], poses[None], cfg=dict(dt_gamma_scale=0.25, return_rgba=True))
        e1.record()
        torch.cuda.synchronize()
        print('rep', rep, 'render ms', e0.elapsed_time(e1))
    print('rounds', st[2], 'shade rounds', st[1], 'dda warp trips', st[3], 'rays', V * IMG * IMG)
    print('render ms', e0.elapsed_time(e1), 'samples', st[0], 'Gsamples/s', st[0] / e0.elapsed_time(e1) / 1e6, 'lane util', st[0] / max(st[1] * 32, 1))
