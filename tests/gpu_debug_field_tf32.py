import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import field_oracle as fo
from tests.test_gpu_field import make_decoder
from mvedit_b200.ingp_decoder import level_table

for L, R in ((12, 320), (14, 512)):
    dec, levels, (table, w1, b1, w2, b2) = make_decoder(L, R)
    g = torch.Generator().manual_seed(3)
    M = 3001
    xyz = (torch.rand(M, 3, generator=g) * 2 - 1) * 0.999
    pt = [t.clone().requires_grad_(True) for t in (table, w1, b1, w2, b2)]
    xo = xyz.clone().requires_grad_(True)
    sig_o, rgb_o = fo.point_decode(xo, *pt, levels)
    gs, gr = torch.randn(M, generator=g), torch.randn(M, 3, generator=g)
    (sig_o * gs).sum().add((rgb_o * gr).sum()).backward()
    res = {}
    for tf32 in (False, True):
        dec.mlp_tf32 = tf32
        for p in dec.parameters():
            p.grad = None
        xg = xyz.cuda().requires_grad_(True)
        sig, rgb, _ = dec.point_decode([xg], None, None)
        torch.autograd.backward([sig, rgb], [gs.cuda(), gr.cuda()])
        names = ['table', 'w1', 'b1', 'w2', 'b2']
        grads = [dec.encoder.params.grad, dec.mlp.net[0].weight.grad, dec.mlp.net[0].bias.grad, dec.mlp.net[1].weight.grad, dec.mlp.net[1].bias.grad]
        for n, a, b in zip(names, grads, [p.grad for p in pt]):
            a = a.detach().cpu().double().reshape(-1); b = b.detach().double().reshape(-1)
            print('tf32' if tf32 else 'fp32', n, 'max abs diff', float((a - b).abs().max()), 'ref max', float(b.abs().max()),
                  'rel-to-max', float((a - b).abs().max() / b.abs().max()), 'rel-L2', float((a - b).norm() / b.norm()))
        lt = level_table(L, 16, R)
        a = dec.encoder.params.grad.detach().cpu().double().reshape(-1, 2); b = pt[0].grad.double().reshape(-1, 2)
        offs = [int(o) for o in lt['off']] + [a.shape[0]]
        for l in range(L):
            da = (a[offs[l]:offs[l + 1]] - b[offs[l]:offs[l + 1]]).abs().max(); mb = b[offs[l]:offs[l + 1]].abs().max()
            pass
        dx = xg.grad.cpu().double() - xo.grad.double()
        print('  dxyz max', float(dx.abs().max()), float(xo.grad.abs().max()), 'rel-L2', float(dx.norm() / xo.grad.double().norm()))
