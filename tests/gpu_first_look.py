"""First-look timings on the GPU box (not a bench line): config-5 ray-march/composite kernels, ours vs the reference's."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tests import synth
from oracle import build_ref
from mvedit_b200 import raymarching as rm


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)


def main():
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    H = 128
    grid = synth.sphere_density_grid(H=H, radius=0.5)
    bf = rm.packbits(cu(grid), 0.5)
    poses = synth.surround_poses(64, seed=0)
    ro, rd, f = synth.camera_rays(poses, 256)
    ro, rd = cu(ro), cu(rd)
    N = ro.shape[0]
    aabb = cu(np.array([-1, -1, -1, 1, 1, 1], np.float32))
    nears, fars = rm.near_far_from_aabb(ro, rd, aabb, 0.2)
    noises = torch.rand(N, device='cuda')
    x, d, t, rays = rm.march_rays_train(ro, rd, 1.0, bf, 1, H, nears, fars, perturb=True, dt_gamma=1 / f, max_steps=1024, noises=noises)
    M = x.shape[0]
    print('N rays', N, 'M samples', M, 'mean/ray', M / N)
    sig = torch.exp(torch.randn(M, device='cuda')); rgb = torch.rand(M, 3, device='cuda')
    print('near_far ms', timeit(lambda: rm.near_far_from_aabb(ro, rd, aabb, 0.2)))
    tm = timeit(lambda: rm.march_rays_train(ro, rd, 1.0, bf, 1, H, nears, fars, perturb=True, dt_gamma=1 / f, max_steps=1024, noises=noises, max_points=M + 16))
    print('march fused ms', tm, 'GB/s', (M * 32 + N * 44) / tm / 1e6)
    tf = timeit(lambda: rm.composite_rays_train(sig, rgb, t, rays))
    print('composite fwd ms (incl. torch allocs)', tf, 'GB/s', (M * 28 + N * 28) / tf / 1e6)
    s2, c2 = sig.clone().requires_grad_(True), rgb.clone().requires_grad_(True)
    w, ws, dep, img = rm.composite_rays_train(s2, c2, t, rays)
    gw, gws, gd, gi = torch.randn_like(w), torch.randn_like(ws), torch.randn_like(dep), torch.randn_like(img)
    def bwd():
        torch.autograd.backward([w, ws, dep, img], [gw, gws, gd, gi], retain_graph=True)
    tb = timeit(bwd)
    print('composite bwd ms (incl. torch allocs)', tb, 'GB/s', (M * 44 + N * 48) / tb / 1e6)
    ref = build_ref.load_ref()
    if ref is not None:
        counter = torch.zeros(1, dtype=torch.int32, device='cuda'); rr = torch.empty(N, 2, dtype=torch.int32, device='cuda')
        xr, dr, tr = torch.zeros(M, 3, device='cuda'), torch.zeros(M, 3, device='cuda'), torch.zeros(M, 2, device='cuda')
        def ref_march():
            counter.zero_()
            ref.march_rays_train(ro, rd, bf, 1.0, False, 1 / f, 1024, N, 1, H, nears, fars, None, None, None, rr, counter, noises)
            ref.march_rays_train(ro, rd, bf, 1.0, False, 1 / f, 1024, N, 1, H, nears, fars, xr, dr, tr, rr, counter, noises)
        print('REF march 2-pass ms', timeit(ref_march))
        wr, wsr, der, imr = torch.zeros(M, device='cuda'), torch.empty(N, device='cuda'), torch.empty(N, device='cuda'), torch.empty(N, 3, device='cuda')
        print('REF composite fwd ms', timeit(lambda: ref.composite_rays_train_forward(sig, rgb, tr, rr, M, N, 1e-4, False, wr, wsr, der, imr)))
        gs, gc = torch.zeros(M, device='cuda'), torch.zeros(M, 3, device='cuda')
        print('REF composite bwd ms', timeit(lambda: ref.composite_rays_train_backward(gw, gws, gd, gi, sig, rgb, tr, rr, wsr, der, imr, M, N, 1e-4, False, gs, gc)))
    # GEMM first look
    from mvedit_b200 import tc_ops
    for (Mm, Nn, Kk) in [(8192, 8192, 8192), (262144, 320, 320), (65536, 1280, 1280), (262144, 2560, 320)]:
        a = torch.randn(Mm, Kk, device='cuda').bfloat16(); w = torch.randn(Nn, Kk, device='cuda').bfloat16()
        out = torch.empty(Mm, Nn, device='cuda', dtype=torch.bfloat16)
        tg = timeit(lambda: tc_ops.gemm(a, w, out=out))
        tt = timeit(lambda: torch.matmul(a, w.t(), out=out))
        print('gemm', Mm, Nn, Kk, 'ours ms', tg, 'TF/s', 2 * Mm * Nn * Kk / tg / 1e9, '| cublas ms', tt, 'TF/s', 2 * Mm * Nn * Kk / tt / 1e9)
    for (B, Hh, C, Co) in [(64, 64, 320, 320), (64, 32, 640, 640), (64, 16, 1280, 1280), (8, 512, 128, 128), (8, 256, 256, 256), (16, 128, 512, 512)]:
        xx = torch.randn(B, Hh, Hh, C, device='cuda').bfloat16(); ww = torch.randn(Co, 3, 3, C, device='cuda').bfloat16()
        tg = timeit(lambda: tc_ops.conv3x3(xx, ww))
        xn = xx.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last); wn = ww.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        tt = timeit(lambda: torch.nn.functional.conv2d(xn, wn, padding=1))
        fl = 2 * B * Hh * Hh * C * Co * 9
        print('conv3x3', B, Hh, C, Co, 'ours ms', tg, 'TF/s', fl / tg / 1e9, '| cudnn ms', tt, 'TF/s', fl / tt / 1e9)
    # attention vs flash SDPA (torch picks the flash / cuDNN backend for bf16)
    for (B, heads, S, d) in [(16, 8, 4096, 40), (16, 8, 1024, 80), (16, 8, 256, 160)]:
        qkv = torch.randn(B, S, 3 * heads * d, device='cuda').bfloat16()
        C = heads * d
        tg = timeit(lambda: tc_ops.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], heads))
        q, k, v = (qkv[:, :, i * C:(i + 1) * C].reshape(B, S, heads, d).transpose(1, 2) for i in range(3))
        tt = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
        fl = 4 * B * heads * S * S * d
        print('attention', B, heads, S, d, 'ours ms', tg, 'TF/s', fl / tg / 1e9, '| torch SDPA ms', tt, 'TF/s', fl / tt / 1e9)


if __name__ == '__main__':
    main()
