"""Diagnostic: fused renderer vs the reference's stepwise inference protocol (march_rays -> point_decode -> composite_rays loop,
base_volume_renderer.py:264-329) run with this library's per-function kernels, on a trained scene with a carved occupancy grid."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import synth
from bench import synth_targets
from mvedit_b200.nerf import BaseNeRF, nerf_optim
from mvedit_b200 import raymarching as rm
from mvedit_b200.ingp_decoder import iNGPDecoder


def stepwise_render(dec, ro, rd, bits, H, dt_gamma, T_thresh=1e-2):
    N = ro.shape[0]
    dev = ro.device
    nears, fars = rm.near_far_from_aabb(ro, rd, dec.aabb, dec.min_near)
    ws = torch.zeros(N, device=dev); d = torch.zeros(N, device=dev); img = torch.zeros(N, 3, device=dev)
    alive = torch.arange(N, dtype=torch.int32, device=dev)
    rt = nears.clone()
    step, shaded = 0, 0
    while step < dec.max_steps:
        n_alive = alive.numel()
        if n_alive == 0:
            break
        n_step = min(max(N // n_alive, 1), 8)
        x, _, t = rm.march_rays(n_alive, n_step, alive, rt, ro, rd, dec.bound, bits, 1, H, nears, fars, dt_gamma=dt_gamma, max_steps=dec.max_steps)
        s_, c_, _ = dec.point_decode([x], None, None)
        shaded += int((t[:, 1] > 0).sum())
        rm.composite_rays(n_alive, n_step, alive, rt, s_, c_, t, ws, d, img, T_thresh=T_thresh)
        alive = alive[alive >= 0].contiguous()
        step += n_step
    return ws, d, img, shaded


def main():
    torch.manual_seed(0)
    V, IMG, R = int(os.environ.get('PR_VIEWS', 8)), 512, int(os.environ.get('PR_RENDER', 192))
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
        nerf_optim(nerf, img[None], msk[None], None, opt, 0.01, int(os.environ.get('PR_ITERS', 640)), 16384, 0.0, 0.0, 0.02, 0.1, 0.01, None, grid, bits,
                   IMG, K, IMG, poses, torch.ones(V, device=dev), torch.nn.functional.normalize(torch.randn(V, 3, device=dev), dim=-1), 128, True,
                   0.015, 0.2, 1.0, False)
        dec = nerf.decoder.eval()
        occ = int((((bits.view(-1).to(torch.int32).unsqueeze(-1) >> torch.arange(8, device=dev)) & 1).sum()).item())
        print('occupied cells', occ, 'of', 128 ** 3)
        Kr = K[:2] * (R / IMG)
        dtg = 0.25 / float(Kr[0, 0])
        ws_c, d_c, img_c = dec.render_cameras(poses[:2], Kr, R, R, bits[0], 128, dt_gamma=dtg)
        st = dec.last_render_stats()
        # explicit rays of the same cameras
        j, i = torch.meshgrid(torch.arange(R, device=dev, dtype=torch.float32) + 0.5, torch.arange(R, device=dev, dtype=torch.float32) + 0.5, indexing='ij')
        ro, rd = [], []
        for v in range(2):
            dirs = torch.stack([(i - Kr[v, 2]) / Kr[v, 0], (j - Kr[v, 3]) / Kr[v, 1], torch.ones_like(i)], -1)
            dw = dirs @ poses[v, :3, :3].T
            rd.append((dw / dw.norm(dim=-1, keepdim=True)).reshape(-1, 3))
            ro.append(poses[v, :3, 3].expand(R * R, 3))
        ro, rd = torch.cat(ro).contiguous(), torch.cat(rd).contiguous()
        ws, d, im, shaded = stepwise_render(dec, ro, rd, bits[0], 128, dtg)
        print('fused samples', st[0], 'stepwise samples', shaded)
        dw = (ws_c.reshape(-1) - ws).abs()
        print('weights_sum: max abs diff', float(dw.max()), 'frac > 5e-3', float((dw > 5e-3).float().mean()), 'mean ws', float(ws.mean()))
        print('image: max abs diff', float((img_c.reshape(-1, 3) - im).abs().max()))


if __name__ == '__main__':
    main()
