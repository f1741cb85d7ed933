"""CPU tests of host-side logic that the CUDA kernels rely on (no GPU, no library calls):
  * tc_ops.geglu_interleave -- the weight re-ordering the GEMM's GEGLU epilogue (act = 3) expects;
  * the fragment algebra of the tensor-core MLP backward (mvedit_b200/csrc/field_bwd_mma.cuh): an m16n8k8 mma.sync is emulated
    with numpy from the PTX fragment layouts and the kernel's index expressions are replayed lane by lane, so a wrong fragment
    index (the easiest mistake to make in that file) fails here, in seconds, without a GPU."""
import numpy as np
import torch


def test_geglu_interleave_layout_and_math():
    from mvedit_b200.tc_ops import geglu_interleave
    F, K, tile = 512, 24, 256
    g = torch.Generator().manual_seed(0)
    w, b = torch.randn(2 * F, K, generator=g), torch.randn(2 * F, generator=g)
    wi, bi = geglu_interleave(w, b, tile)
    assert wi.shape == w.shape and bi.shape == b.shape
    h = tile // 2
    for n in range(2 * F // tile):                       # tile n = [values n*h..(n+1)*h | gates F + n*h..]
        assert torch.equal(wi[n * tile:n * tile + h], w[n * h:(n + 1) * h])
        assert torch.equal(wi[n * tile + h:(n + 1) * tile], w[F + n * h:F + (n + 1) * h])
        assert torch.equal(bi[n * tile + h:(n + 1) * tile], b[F + n * h:F + (n + 1) * h])
    # what the epilogue computes per tile == diffusers GEGLU on the original layout
    x = torch.randn(5, K, generator=g)
    val, gate = (x @ w.t() + b).chunk(2, dim=-1)
    ref = val * torch.nn.functional.gelu(gate)
    hi = x @ wi.t() + bi
    out = torch.cat([hi[:, n * tile:n * tile + h] * torch.nn.functional.gelu(hi[:, n * tile + h:(n + 1) * tile]) for n in range(2 * F // tile)], 1)
    torch.testing.assert_close(out, ref)


# ---------------------------------------------------------------------------------------------- mma.sync.m16n8k8 emulation
LANES = np.arange(32)
G, T = LANES >> 2, LANES & 3


def mma(C, A, B):
    """Per-lane fragments (PTX ISA, m16n8k8 .tf32): A a0(row g,k t) a1(g+8,t) a2(g,t+4) a3(g+8,t+4); B b0(k t,n g) b1(k t+4,n g);
    C c0(g,2t) c1(g,2t+1) c2(g+8,2t) c3(g+8,2t+1)."""
    Am, Bm, Cm = np.zeros((16, 8)), np.zeros((8, 8)), np.zeros((16, 8))
    for l in range(32):
        g, t = G[l], T[l]
        Am[g, t], Am[g + 8, t], Am[g, t + 4], Am[g + 8, t + 4] = A[l]
        Bm[t, g], Bm[t + 4, g] = B[l]
        Cm[g, 2 * t], Cm[g, 2 * t + 1], Cm[g + 8, 2 * t], Cm[g + 8, 2 * t + 1] = C[l]
    D = Cm + Am @ Bm
    return np.array([[D[G[l], 2 * T[l]], D[G[l], 2 * T[l] + 1], D[G[l] + 8, 2 * T[l]], D[G[l] + 8, 2 * T[l] + 1]] for l in range(32)])


def c_to_a(c):      # field_bwd_mma.cuh c_to_a: logical k = t <-> column 2t, k = t+4 <-> column 2t+1
    return np.stack([c[:, 0], c[:, 2], c[:, 1], c[:, 3]], 1)


def test_mlp_backward_fragment_algebra():
    rng = np.random.default_rng(0)
    IN, HID, KS, NT, FT, WT, MT, LD = 24, 64, 3, 8, 3, 4, 4, 40          # BCfg<12>
    W1, b1, W2 = rng.normal(size=(HID, IN)), rng.normal(size=HID), rng.normal(size=(4, HID))
    Enc, dOut = rng.normal(size=(32, IN)), rng.normal(size=(32, 4))
    st = np.zeros(32 * LD)
    for f in range(IN):
        st[f * LD:f * LD + 32] = Enc[:, f]
    st[IN * LD:IN * LD + 32] = 1.0                                           # init_bwd_stage: ones row -> db1
    xch = dOut.reshape(-1)
    w1 = lambda h, f: W1[h, f] if f < IN else 0.0
    w2 = lambda o, h: W2[o, h] if o < 4 else 0.0
    # reference
    H = Enc @ W1.T + b1
    dH = (dOut @ W2) * (H > 0)
    ref = dict(dEnc=dH @ W1, dW1=dH.T @ Enc, db1=dH.sum(0), dW2=dOut.T @ np.maximum(H, 0))
    # pass A (sample-major forward, relu mask bits)
    a = np.zeros((2, KS, 32, 4))
    for mt in range(2):
        for s in range(KS):
            for l in range(32):
                g, t = G[l], T[l]
                a[mt, s, l] = [st[(8 * s + t) * LD + 16 * mt + g], st[(8 * s + t) * LD + 16 * mt + g + 8],
                               st[(8 * s + t + 4) * LD + 16 * mt + g], st[(8 * s + t + 4) * LD + 16 * mt + g + 8]]
    mask = np.zeros((2, 32, NT * 4), bool)
    for nt in range(NT):
        h = np.array([[[b1[8 * nt + 2 * T[l]], b1[8 * nt + 2 * T[l] + 1]] * 2 for l in range(32)]] * 2, dtype=float)
        for s in range(KS):
            B = np.array([[w1(8 * nt + G[l], 8 * s + T[l]), w1(8 * nt + G[l], 8 * s + T[l] + 4)] for l in range(32)])
            for mt in range(2):
                h[mt] = mma(h[mt], a[mt, s], B)
        mask[:, :, nt * 4:nt * 4 + 4] = h > 0
    # pass C (hidden-major, weight gradients)
    acc1, acc2 = np.zeros((MT, WT, 32, 4)), np.zeros((MT, 32, 4))
    bdo = np.array([[xch[(8 * ns + G[l]) * 4 + T[l]] for l in range(32)] for ns in range(4)])
    bd2 = np.array([[[xch[(8 * ns + 2 * T[l]) * 4 + G[l]] if G[l] < 4 else 0, xch[(8 * ns + 2 * T[l] + 1) * 4 + G[l]] if G[l] < 4 else 0]
                     for l in range(32)] for ns in range(4)])
    for mt in range(MT):
        hT = np.array([[[b1[16 * mt + G[l]]] * 2 + [b1[16 * mt + G[l] + 8]] * 2 for l in range(32)]] * 4, dtype=float)
        for ks in range(KS):
            aw = np.array([[w1(16 * mt + G[l], 8 * ks + T[l]), w1(16 * mt + G[l] + 8, 8 * ks + T[l]),
                            w1(16 * mt + G[l], 8 * ks + T[l] + 4), w1(16 * mt + G[l] + 8, 8 * ks + T[l] + 4)] for l in range(32)])
            for ns in range(4):
                hT[ns] = mma(hT[ns], aw, np.stack([a[ns >> 1, ks, :, ns & 1], a[ns >> 1, ks, :, 2 + (ns & 1)]], 1))
        a4 = np.array([[w2(T[l], 16 * mt + G[l]), w2(T[l], 16 * mt + G[l] + 8), 0, 0] for l in range(32)])
        for ns in range(4):
            dT = mma(np.zeros((32, 4)), a4, np.stack([bdo[ns], np.zeros(32)], 1))
            dT = np.where(hT[ns] > 0, dT, 0)
            hr = np.maximum(hT[ns], 0)
            for ft in range(WT):
                be = np.array([[st[(8 * ft + G[l]) * LD + 8 * ns + 2 * T[l]], st[(8 * ft + G[l]) * LD + 8 * ns + 2 * T[l] + 1]] for l in range(32)])
                acc1[mt, ft] = mma(acc1[mt, ft], c_to_a(dT), be)
            acc2[mt] = mma(acc2[mt], c_to_a(hr), bd2[ns])
    gW1, gb1, gW2 = np.zeros((HID, IN)), np.zeros(HID), np.zeros((4, HID))
    for mt in range(MT):                                                     # the CTA reduction's index map
        for l in range(32):
            for c in range(4):
                hid = 16 * mt + G[l] + 8 * (c >> 1)
                for ft in range(WT):
                    f = 8 * ft + 2 * T[l] + (c & 1)
                    if f < IN:
                        gW1[hid, f] += acc1[mt, ft, l, c]
                    elif f == IN:
                        gb1[hid] += acc1[mt, ft, l, c]
                o = 2 * T[l] + (c & 1)
                if o < 4:
                    gW2[o, hid] += acc2[mt, l, c]
    # pass B (sample-major dH -> dEnc)
    de = np.zeros((2, FT, 32, 4))
    for nt in range(NT):
        b4 = np.array([[w2(T[l], 8 * nt + G[l]), 0] for l in range(32)])
        for mt in range(2):
            ado = np.array([[xch[(16 * mt + G[l]) * 4 + T[l]], xch[(16 * mt + G[l] + 8) * 4 + T[l]], 0, 0] for l in range(32)])
            dh = np.where(mask[mt, :, nt * 4:nt * 4 + 4], mma(np.zeros((32, 4)), ado, b4), 0)
            for ft in range(FT):
                b3 = np.array([[w1(8 * nt + 2 * T[l], 8 * ft + G[l]), w1(8 * nt + 2 * T[l] + 1, 8 * ft + G[l])] for l in range(32)])
                de[mt, ft] = mma(de[mt, ft], c_to_a(dh), b3)
    gE = np.zeros((32, IN))
    for mt in range(2):
        for ft in range(FT):
            for l in range(32):
                f = 8 * ft + 2 * T[l]
                if f < IN:
                    gE[16 * mt + G[l], f], gE[16 * mt + G[l], f + 1], gE[16 * mt + G[l] + 8, f], gE[16 * mt + G[l] + 8, f + 1] = de[mt, ft, l]
    for name, got in (('dEnc', gE), ('dW1', gW1), ('db1', gb1), ('dW2', gW2)):
        np.testing.assert_allclose(got, ref[name], rtol=1e-10, atol=1e-10, err_msg=name)
