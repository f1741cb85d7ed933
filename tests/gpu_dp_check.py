"""2-rank check of the data-parallel reconstruction (run under torchrun on a 2-GPU box; not collected by pytest):
every rank marches its row strip, per-ray outputs are all-gathered, the flat gradient is all-reduced, Adam runs replicated.
Checks: replicas stay bit-identical; the fit converges like the single-GPU run.  Prints one line 'DP_CHECK ok ...'."""
import math
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import synth  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    from mvedit_b200.nerf import BaseNeRF, nerf_optim, pixel_directions
    from mvedit_b200.ingp_decoder import iNGPDecoder
    from mvedit_b200.optim import FusedAdam
    graph = os.environ.get('DP_GRAPH', '1') == '1'
    torch.manual_seed(0)
    V, size, ps = 6, 64, 32
    poses = torch.from_numpy(synth.surround_poses(V, seed=3)).to(dev)
    f = 0.5 * size / math.tan(math.radians(15))
    K = torch.tensor([[f, f, size / 2, size / 2]] * V, device=dev)
    d = pixel_directions(K[None], size, size)
    rd = torch.nn.functional.normalize(d @ poses[None, :, None, :3, :3].transpose(-1, -2), dim=-1)
    ro = poses[None, :, None, None, :3, 3].expand(rd.shape)
    b = (ro * rd).sum(-1)
    disc = b * b - ((ro * ro).sum(-1) - 0.25)
    hit = disc > 0
    p = ro + (-b - disc.clamp(min=0).sqrt())[..., None] * rd
    img = torch.where(hit[..., None], 0.5 + 0.5 * torch.sin(p * 6), torch.ones_like(p))
    msk = hit[..., None].float()
    nerf = BaseNeRF(grid_size=64, decoder=iNGPDecoder(max_steps=256, weight_culling_th=0.001), patch_size=ps).to(dev)
    nerf.use_cuda_graph, nerf.data_parallel = graph, True
    grid, bits = nerf.get_init_density_grid(1, dev), nerf.get_init_density_bitfield(1, dev)
    opt = FusedAdam(nerf.decoder.parameters(), lr=0.01)
    kw = dict(optimizer=opt, lr=0.01, n_inverse_rays=ps * ps * 2, patch_rgb_weight=0.0, patch_normal_weight=0.0, alpha_soften=0.02,
              normal_reg_weight=0.1, entropy_weight=0.01, nerf_code=None, density_grid=grid, density_bitfield=bits, render_size=size,
              intrinsics=K, intrinsics_size=size, camera_poses=poses, cam_weights=torch.ones(V, device=dev),
              cam_lights=torch.nn.functional.normalize(torch.randn(V, 3, device=dev), dim=-1), patch_size=ps, is_init=True,
              bg_width=0.015, ambient_light=0.2, dt_gamma_scale=0.5, init_shaded=False)
    t0 = time.time()
    for _ in range(3):
        nerf_optim(nerf, img, msk, None, inverse_steps=64, **kw)
    torch.cuda.synchronize()
    dt = time.time() - t0
    # replicas identical?
    flat = torch.cat([q.detach().reshape(-1) for q in nerf.decoder.parameters()])
    other = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(other, flat)
    same = all(torch.equal(o, other[0]) for o in other)
    bsame = [torch.empty_like(bits) for _ in range(world)]
    dist.all_gather(bsame, bits)
    same_bits = all(torch.equal(o, bsame[0]) for o in bsame)
    rgba, depth = nerf.render(nerf.decoder, None, bits, size, size, K[None], poses[None], cfg=dict(dt_gamma_scale=0.5, return_rgba=True))
    a_err = (rgba[..., 3:] - msk).abs().mean().item()
    c_err = (rgba[..., :3] + (1 - rgba[..., 3:]) - img).abs().mean().item()
    # ---- gradient equivalence: ONE iteration data-parallel vs the same iteration on one rank (same patch, same noise).  After the
    # first Adam step exp_avg = 0.1 * g, so the flat first-moment buffer IS the (all-reduced) gradient.
    def one_iter(dp):
        torch.manual_seed(123)
        n2 = BaseNeRF(grid_size=64, decoder=iNGPDecoder(max_steps=256, weight_culling_th=0.001), patch_size=ps).to(dev)
        n2.decoder.load_state_dict(nerf.decoder.state_dict())
        n2.use_cuda_graph, n2.data_parallel = False, dp
        o2 = FusedAdam(n2.decoder.parameters(), lr=0.01)
        g2, b2 = grid.clone(), bits.clone()
        n2.update_extra_iters = 0
        torch.manual_seed(77)
        nerf_optim(n2, img, msk, None, inverse_steps=1, **dict(kw, optimizer=o2, density_grid=g2, density_bitfield=b2))
        return o2._m.clone() * 10.0
    g_dp, g_one = one_iter(True), one_iter(False)
    g_rel = ((g_dp - g_one).norm() / g_one.norm()).item()
    ok = same and same_bits and a_err < 0.08 and c_err < 0.08 and g_rel < 2e-3
    if rank == 0:
        print('DP_CHECK %s graph=%s replicas_identical=%s bitfields_identical=%s alpha_err=%.4f rgb_err=%.4f dp_vs_single_grad_rel=%.2e wall=%.2fs max_kept=%d' % (
            'ok' if ok else 'FAILED', graph, same, same_bits, a_err, c_err, g_rel, dt, nerf.decoder.check_sample_overflow()), flush=True)
    dist.barrier()
    # captured graphs hold NCCL collectives of this process group: drop them before the communicator, and do not block on its teardown
    import gc
    nerf.__dict__.get('_recon_programs', {}).clear()
    gc.collect()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if ok else 1)


if __name__ == '__main__':
    main()
