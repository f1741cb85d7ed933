"""Times mve_field_backward as one launch and split (MLP backward + scatter kernel) on uniform and surface-concentrated samples.
Run on a GPU box:  python tests/gpu_field_bwd_split.py"""
import torch
import mvedit_b200.ingp_decoder as ing
from mvedit_b200.ingp_decoder import iNGPDecoder


def main():
    dec = iNGPDecoder().cuda()
    torch.manual_seed(0)
    with torch.no_grad():
        dec.encoder.params.uniform_(-0.1, 0.1)
    for name, M in (('uniform', 600_000), ('uniform', 2_000_000), ('shell', 2_000_000)):
        xyz = torch.rand(M, 3, device='cuda') * 2 - 1
        if name == 'shell':           # samples near a sphere of radius 0.5, the way a fitted field's kept samples cluster at a surface
            d = torch.nn.functional.normalize(torch.randn(M, 3, device='cuda'), dim=-1)
            xyz = d * (0.5 + 0.02 * torch.randn(M, 1, device='cuda'))
        gs, gr = torch.randn(M, device='cuda'), torch.randn(M, 3, device='cuda')
        for split in (False, True):
            ing._FIELD_SPLIT = split
            ts = []
            for it in range(6):
                for p in dec.parameters():
                    p.grad = None
                sig, rgb, _ = dec.point_decode([xyz], None, None)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                torch.autograd.backward([sig, rgb], [gs, gr])
                b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            print(f'{name:8s} M={M:8d} split={int(split)}  backward {min(ts[2:]):.3f} ms (autograd call incl. zeros_like of the table grad)')


if __name__ == '__main__':
    main()
