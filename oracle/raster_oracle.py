"""CPU oracle of the mesh rasteriser ops (seam B5) -- TEST INFRASTRUCTURE ONLY (imported by tests/, smoke() and bench.py's
cpu_baseline; never by the product).

The reference calls ``nvdiffrast.torch`` (``lib/models/decoders/mesh_renderer/base_mesh_renderer.py:5,204,241-298,407-410``); nvdiffrast is
an un-vendored dependency pinned by commit in ``requirements.txt:3`` and is not installed here, so this restates the published semantics
 -- **parity unpinned**: no golden vector of nvdiffrast exists offline.  What is anchored on the reference itself
are its call sites: argument layout, the (u, v, z/w, id + 1) output convention (``:240,244``), attribute broadcasting (``:251-252``),
``diff_attrs='all'`` (``:260-261``), the 8-channel antialias call (``:296-298``).

Conventions restated:
  * clip-space positions ``pos [B,V,4]``; pixel (row y, col x) has its centre at NDC ``((2x+1)/W - 1, (2y+1)/H - 1)``;
  * a pixel is covered when its centre is inside or on the triangle (both windings, no culling) and within the triangle's fp32 screen
    bounding box padded by one pixel (which pins down zero-area triangles, whose edge functions are rounding noise); nearest ``z/w`` within [-1, 1] wins,
    ties go to the lower triangle index; a triangle with a vertex at ``w <= 0`` is dropped (no near-plane clipping);
  * ``rast = (u, v, z/w, id + 1)`` with (u, v) the perspective-correct barycentrics of the triangle's vertices 0 and 1;
    ``rast_db = (du/dX, du/dY, dv/dX, dv/dY)`` per pixel;
  * ``interpolate``: ``u a0 + v a1 + (1-u-v) a2``; ``antialias``: blend across silhouette edges by sub-pixel coverage.

``rasterize`` is numpy float32 with the operation order of the kernel spelled out (no fused multiply-add), so ids and barycentrics
are compared bit for bit.  The differentiable pieces are torch functions of (pos, attr, color) given the discrete outputs, so autograd
supplies the reference gradients.
"""
import numpy as np
import torch

f32 = np.float32


def _edge_functions(p0, p1, p2, fx, fy):
    """Homogeneous edge functions a0, a1, a2 at NDC (fx, fy) (arrays), fp32, unfused; also the q vectors."""
    q0x = p0[0] - fx * p0[3]; q0y = p0[1] - fy * p0[3]
    q1x = p1[0] - fx * p1[3]; q1y = p1[1] - fy * p1[3]
    q2x = p2[0] - fx * p2[3]; q2y = p2[1] - fy * p2[3]
    a0 = q1x * q2y - q1y * q2x
    a1 = q2x * q0y - q2y * q0x
    a2 = q0x * q1y - q0y * q1x
    return (a0, a1, a2), (q0x, q0y, q1x, q1y, q2x, q2y)


def rasterize(pos, tri, resolution, grad_db=True):
    """pos [B,V,4] or [V,4] float32 numpy, tri [F,3] int -> rast [B,H,W,4], rast_db [B,H,W,4] (zeros when grad_db is False)."""
    pos = np.asarray(pos, f32)
    tri = np.asarray(tri, np.int64)
    if pos.ndim == 2:
        pos = pos[None]
    H, W = resolution
    B = pos.shape[0]
    best_zw = np.full((B, H, W), np.inf, f32)
    best_id = np.full((B, H, W), -1, np.int64)
    xs = (np.arange(W, dtype=f32) + f32(0.5)) * (f32(2.0) / f32(W)) - f32(1.0)
    ys = (np.arange(H, dtype=f32) + f32(0.5)) * (f32(2.0) / f32(H)) - f32(1.0)
    with np.errstate(all='ignore'):
        for b in range(B):
            for t in range(tri.shape[0]):
                i0, i1, i2 = tri[t]
                if min(i0, i1, i2) < 0 or max(i0, i1, i2) >= pos.shape[1]:
                    continue
                p0, p1, p2 = pos[b, i0], pos[b, i1], pos[b, i2]
                if not (p0[3] > 0 and p1[3] > 0 and p2[3] > 0):
                    continue
                # candidate pixels: the triangle's screen bounding box in fp32, one pixel of slack (part of the specification: a
                # zero-area triangle has rounding noise for edge functions, so WHERE they are evaluated must be pinned down too)
                hw_, hh_ = f32(0.5) * f32(W), f32(0.5) * f32(H)
                sx = np.array([p[0] / p[3] * hw_ + hw_ for p in (p0, p1, p2)], f32)
                sy = np.array([p[1] / p[3] * hh_ + hh_ for p in (p0, p1, p2)], f32)
                mnx, mxx = np.fmin.reduce(sx), np.fmax.reduce(sx)          # fmin / fmax ignore NaN, like fminf / fmaxf
                mny, mxy = np.fmin.reduce(sy), np.fmax.reduce(sy)
                if not (mxx >= 0 and mnx <= f32(W) and mxy >= 0 and mny <= f32(H)):
                    continue
                x0 = int(np.floor(np.fmax(mnx, f32(0)) - f32(0.5))); x1 = int(np.ceil(np.fmin(mxx, f32(W)) - f32(0.5)))
                y0 = int(np.floor(np.fmax(mny, f32(0)) - f32(0.5))); y1 = int(np.ceil(np.fmin(mxy, f32(H)) - f32(0.5)))
                x0, y0, x1, y1 = max(x0, 0), max(y0, 0), min(x1, W - 1) + 1, min(y1, H - 1) + 1      # [x0, x1) x [y0, y1)
                if x1 <= x0 or y1 <= y0:
                    continue
                fx, fy = np.meshgrid(xs[x0:x1], ys[y0:y1])
                (a0, a1, a2), _ = _edge_functions(p0, p1, p2, fx, fy)
                at = (a0 + a1) + a2
                inside = (((a0 >= 0) & (a1 >= 0) & (a2 >= 0)) | ((a0 <= 0) & (a1 <= 0) & (a2 <= 0))) & (at != 0)
                z = (p0[2] * a0 + p1[2] * a1) + p2[2] * a2
                w = (p0[3] * a0 + p1[3] * a1) + p2[3] * a2
                zw = z / w
                inside &= (zw >= -1) & (zw <= 1)
                cur = best_zw[b, y0:y1, x0:x1]
                cid = best_id[b, y0:y1, x0:x1]
                # monotonic-bits comparison == float comparison except that -0.0 < +0.0 in the kernel's key
                key_lt = (zw < cur) | ((zw == cur) & (np.signbit(zw) & ~np.signbit(cur)))
                take = inside & (key_lt | ((zw == cur) & (np.signbit(zw) == np.signbit(cur)) & (cid > t)) | (cid < 0))
                cur[take] = zw[take]
                cid[take] = t
    rast = np.zeros((B, H, W, 4), f32)
    rast_db = np.zeros((B, H, W, 4), f32)
    with np.errstate(all='ignore'):
        for b in range(B):
            yy, xx = np.nonzero(best_id[b] >= 0)
            if len(yy) == 0:
                continue
            t = best_id[b, yy, xx]
            p0, p1, p2 = (pos[b, tri[t, k]].T for k in range(3))          # [4, n]
            fx, fy = xs[xx], ys[yy]
            (a0, a1, a2), (q0x, q0y, q1x, q1y, q2x, q2y) = _edge_functions(p0, p1, p2, fx, fy)
            at = (a0 + a1) + a2
            z = (p0[2] * a0 + p1[2] * a1) + p2[2] * a2
            w = (p0[3] * a0 + p1[3] * a1) + p2[3] * a2
            rast[b, yy, xx, 0] = a0 / at
            rast[b, yy, xx, 1] = a1 / at
            rast[b, yy, xx, 2] = z / w
            rast[b, yy, xx, 3] = (t + 1).astype(f32)
            if grad_db:
                a0x = q1y * p2[3] - p1[3] * q2y; a0y = p1[3] * q2x - q1x * p2[3]
                a1x = q2y * p0[3] - p2[3] * q0y; a1y = p2[3] * q0x - q2x * p0[3]
                a2x = q0y * p1[3] - p0[3] * q1y; a2y = p0[3] * q1x - q0x * p1[3]
                atx = (a0x + a1x) + a2x; aty = (a0y + a1y) + a2y
                iat2 = f32(1.0) / (at * at); sx_, sy_ = f32(2.0) / f32(W), f32(2.0) / f32(H)
                rast_db[b, yy, xx, 0] = (a0x * at - a0 * atx) * iat2 * sx_
                rast_db[b, yy, xx, 1] = (a0y * at - a0 * aty) * iat2 * sy_
                rast_db[b, yy, xx, 2] = (a1x * at - a1 * atx) * iat2 * sx_
                rast_db[b, yy, xx, 3] = (a1y * at - a1 * aty) * iat2 * sy_
    return rast, rast_db


# ---- differentiable restatements (torch; any float dtype) -----------------------------------------------------------------------

def _pixel_ndc(H, W, dtype):
    xs = (torch.arange(W, dtype=dtype) + 0.5) * (2.0 / W) - 1.0
    ys = (torch.arange(H, dtype=dtype) + 0.5) * (2.0 / H) - 1.0
    return xs, ys


def barycentrics(pos, tri, tri_id):
    """(u, v, z/w) as a differentiable function of pos [B,V,4] for the given winners tri_id [B,H,W] (-1 = empty) -> [B,H,W,3]."""
    B, H, W = tri_id.shape
    if pos.dim() == 2:
        pos = pos[None].expand(B, -1, -1)
    xs, ys = _pixel_ndc(H, W, pos.dtype)
    fx = xs[None, None, :].expand(B, H, W)
    fy = ys[None, :, None].expand(B, H, W)
    t = tri_id.clamp(min=0)
    bidx = torch.arange(B)[:, None, None].expand(B, H, W)
    p = [pos[bidx, tri[t, k].long()] for k in range(3)]                   # [B,H,W,4]
    q = [(pk[..., 0] - fx * pk[..., 3], pk[..., 1] - fy * pk[..., 3]) for pk in p]
    a0 = q[1][0] * q[2][1] - q[1][1] * q[2][0]
    a1 = q[2][0] * q[0][1] - q[2][1] * q[0][0]
    a2 = q[0][0] * q[1][1] - q[0][1] * q[1][0]
    at = a0 + a1 + a2
    at = torch.where(tri_id >= 0, at, torch.ones_like(at))
    z = p[0][..., 2] * a0 + p[1][..., 2] * a1 + p[2][..., 2] * a2
    w = p[0][..., 3] * a0 + p[1][..., 3] * a1 + p[2][..., 3] * a2
    w = torch.where(tri_id >= 0, w, torch.ones_like(w))
    out = torch.stack([a0 / at, a1 / at, z / w], dim=-1)
    return torch.where((tri_id >= 0)[..., None], out, torch.zeros_like(out))     # where, not *: empty pixels pass no gradient (not even NaN)


def interpolate(attr, rast, tri, rast_db=None, diff_attrs=None):
    """attr [B or 1, Va, C], rast [B,H,W,4] (u, v differentiable), tri [F,3] -> out [B,H,W,C], out_da [B,H,W,2C] or None."""
    B, H, W, _ = rast.shape
    tri_id = rast[..., 3].long() - 1
    fg = tri_id >= 0
    t = tri_id.clamp(min=0)
    if attr.dim() == 2:
        attr = attr[None]
    ab = attr.expand(B, -1, -1) if attr.shape[0] == 1 else attr
    bidx = torch.arange(B)[:, None, None].expand(B, H, W)
    a = [ab[bidx, tri[t, k].long()] for k in range(3)]                    # [B,H,W,C]
    u, v = rast[..., 0:1], rast[..., 1:2]
    out = u * a[0] + v * a[1] + (1 - u - v) * a[2]
    out = torch.where(fg[..., None], out, torch.zeros_like(out))          # empty pixels neither produce nor receive anything
    out_da = None
    if diff_attrs is not None:
        assert diff_attrs == 'all' and rast_db is not None
        e0, e1 = a[0] - a[2], a[1] - a[2]
        dX = e0 * rast_db[..., 0:1] + e1 * rast_db[..., 2:3]
        dY = e0 * rast_db[..., 1:2] + e1 * rast_db[..., 3:4]
        out_da = torch.stack([dX, dY], dim=-1).reshape(B, H, W, -1)
        out_da = torch.where(fg[..., None], out_da, torch.zeros_like(out_da))
    return out, out_da


def edge_opposites(tri):
    """opp [F,3]: for the edge facing vertex k of triangle f, the vertex of the adjacent triangle that is not on the edge (-1: open edge).
    Edges shared by more than two triangles pair their first two users (in (triangle, edge) order); later users see the first."""
    tri = np.asarray(tri, np.int64)
    F = tri.shape[0]
    opp = np.full((F, 3), -1, np.int64)
    users = {}
    for f in range(F):
        for k in range(3):
            a, b = tri[f, (k + 1) % 3], tri[f, (k + 2) % 3]
            users.setdefault((min(a, b), max(a, b)), []).append((f, k))
    for lst in users.values():
        if len(lst) < 2:
            continue
        (f0, k0), (f1, k1) = lst[0], lst[1]
        opp[f0, k0] = tri[f1, k1]
        opp[f1, k1] = tri[f0, k0]
        for f, k in lst[2:]:
            opp[f, k] = tri[f0, k0]
    return opp


def _same_sign(a, b):
    return torch.signbit(a) == torch.signbit(b)


def antialias(color, rast, pos, tri, opp=None):
    """color [B,H,W,C], rast [B,H,W,4], pos [B,V,4] or [V,4], tri [F,3] -> antialiased color; differentiable w.r.t. color and pos.

    For every horizontal (d = 0) and vertical (d = 1) pixel pair with different ids: the nearer triangle (by z/w; a lone foreground
    pixel owns the pair) is examined in pixel units relative to its own pixel centre; an edge is a silhouette if it is open or its
    neighbour folds back (the wing has the triangle's own orientation); among the edges crossing the axis towards the neighbour the
    one crossing farthest is taken, must be a silhouette, steeper than 45 degrees w.r.t. that axis, and cross within [-1/16, 17/16]
    of the pitch; the crossing fraction c (clamped to [0, 1]) gives alpha = +-(0.5 - c): the pixel on the far side of the pixel
    boundary from the crossing receives alpha * (other colour - own colour)."""
    B, H, W, C = color.shape
    dt = color.dtype
    if pos.dim() == 2:
        pos = pos[None].expand(B, -1, -1)
    pos = pos.to(dt)
    tri_t = torch.as_tensor(np.asarray(tri), dtype=torch.long)
    opp_t = torch.as_tensor(edge_opposites(tri) if opp is None else np.asarray(opp), dtype=torch.long)
    ids = rast[..., 3].long() - 1
    zw = rast[..., 2]
    out = color.clone()
    hw, hh = 0.5 * W, 0.5 * H
    for d in (0, 1):
        if d == 0:
            b0, y0, x0 = torch.meshgrid(torch.arange(B), torch.arange(H), torch.arange(W - 1), indexing='ij')
            y1, x1 = y0, x0 + 1
        else:
            b0, y0, x0 = torch.meshgrid(torch.arange(B), torch.arange(H - 1), torch.arange(W), indexing='ij')
            y1, x1 = y0 + 1, x0
        b0, y0, x0, y1, x1 = (v.reshape(-1) for v in (b0, y0, x0, y1, x1))
        t0, t1 = ids[b0, y0, x0], ids[b0, y1, x1]
        sel = t0 != t1
        b0, y0, x0, y1, x1, t0, t1 = (v[sel] for v in (b0, y0, x0, y1, x1, t0, t1))
        if b0.numel() == 0:
            continue
        z0, z1 = zw[b0, y0, x0], zw[b0, y1, x1]
        own0 = torch.where((t0 >= 0) & (t1 >= 0), z0 < z1, t0 >= 0)
        t = torch.where(own0, t0, t1)
        qx = torch.where(own0, x0, x1).to(dt)
        qy = torch.where(own0, y0, y1).to(dt)
        vi = tri_t[t]                                                     # [n,3]
        ov = opp_t[t]
        P = pos[b0[:, None], vi]                                          # [n,3,4]
        O = pos[b0[:, None], ov.clamp(min=0)]
        cx, cy = (qx + 0.5 - hw)[:, None], (qy + 0.5 - hh)[:, None]
        x = P[..., 0] / P[..., 3] * hw - cx
        y = P[..., 1] / P[..., 3] * hh - cy
        ox = torch.where(ov >= 0, O[..., 0] / O[..., 3] * hw - cx, x)
        oy = torch.where(ov >= 0, O[..., 1] / O[..., 3] * hh - cy, y)
        bb = (x[:, 1] - x[:, 0]) * (y[:, 2] - y[:, 0]) - (x[:, 2] - x[:, 0]) * (y[:, 1] - y[:, 0])
        sil = []
        for k in range(3):
            ia, ib = (k + 1) % 3, (k + 2) % 3
            wing = (x[:, ia] - ox[:, k]) * (y[:, ib] - oy[:, k]) - (x[:, ib] - ox[:, k]) * (y[:, ia] - oy[:, k])
            sil.append(_same_sign(wing, bb))
        sil = torch.stack(sil, dim=1)
        if d == 1:
            x, y = y, x
        ds = torch.where(own0, torch.ones_like(qx), -torch.ones_like(qx))
        cs, ok = [], []
        for k in range(3):
            ia, ib = (k + 1) % 3, (k + 2) % 3
            dx, dy = x[:, ib] - x[:, ia], y[:, ib] - y[:, ia]
            crosses = ~_same_sign(y[:, ia], y[:, ib])
            dy_safe = torch.where(crosses, dy, torch.ones_like(dy))
            c = ds * (x[:, ia] * dy_safe - y[:, ia] * dx) / dy_safe
            cs.append(torch.where(crosses, c, torch.full_like(c, -float('inf'))))
            ok.append(crosses & (dy.abs() >= dx.abs()))
        cs, ok = torch.stack(cs, dim=1), torch.stack(ok, dim=1)
        # first maximum (ties -> lowest edge index)
        bk = torch.zeros_like(t)
        best = cs[:, 0]
        for k in (1, 2):
            gt = cs[:, k] > best
            bk = torch.where(gt, torch.full_like(bk, k), bk)
            best = torch.where(gt, cs[:, k], best)
        rows = torch.arange(t.numel())
        eps = 0.0625
        good = sil.any(dim=1) & torch.isfinite(best) & sil[rows, bk] & ok[rows, bk] & (best > -eps) & (best < 1 + eps)
        dc = best.clamp(0, 1)
        alpha = ds * (0.5 - dc)
        g = good.nonzero().reshape(-1)
        if g.numel() == 0:
            continue
        bg_, y0g, x0g, y1g, x1g, ag = b0[g], y0[g], x0[g], y1[g], x1[g], alpha[g]
        delta = ag[:, None] * (color[bg_, y1g, x1g] - color[bg_, y0g, x0g])
        to0 = ag > 0
        flat = out.reshape(-1, C)
        idx = torch.where(to0, (bg_ * H + y0g) * W + x0g, (bg_ * H + y1g) * W + x1g)
        flat = flat.index_add(0, idx, delta)
        out = flat.reshape(B, H, W, C)
    return out


def n_mip_levels(th, tw, max_mip_level=None):
    """Levels of the 2x2-average pyramid: a level is halved while both of its dimensions are even."""
    n = 1
    while th % 2 == 0 and tw % 2 == 0 and (max_mip_level is None or n <= max_mip_level):
        th, tw, n = th // 2, tw // 2, n + 1
    return n


def texture(tex, uv, uv_da=None, filter_mode='linear-mipmap-linear', max_mip_level=None):
    """``dr.texture`` restated (boundary mode 'wrap', nvdiffrast's default): tex [Bt,th,tw,C] (Bt = 1 or B), uv [B,H,W,2] with (0,0) the
    corner of texel (0,0), uv_da [B,H,W,4] = (du/dX, du/dY, dv/dX, dv/dY).  Bilinear taps on the two pyramid levels around
    ``0.5 log2(major axis^2 of the pixel footprint in texels)`` (clamped to the pyramid), linear blend between them; without uv_da or
    with filter_mode='linear': level 0.  Differentiable w.r.t. ``tex`` (torch ops)."""
    B, H, W, _ = uv.shape
    Bt, th, tw, C = tex.shape
    use_mip = filter_mode == 'linear-mipmap-linear' and uv_da is not None
    n_levels = n_mip_levels(th, tw, max_mip_level) if use_mip else 1
    pyr = [tex]
    for _ in range(1, n_levels):
        pyr.append(torch.nn.functional.avg_pool2d(pyr[-1].permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1))
    bidx = (torch.arange(B)[:, None, None].expand(B, H, W) if Bt == B and B > 1 else torch.zeros(B, H, W, dtype=torch.long))

    def tap(level):
        t = pyr[level]
        hl, wl = t.shape[1], t.shape[2]
        x, y = uv[..., 0] * wl - 0.5, uv[..., 1] * hl - 0.5
        x0, y0 = torch.floor(x), torch.floor(y)
        fx, fy = (x - x0)[..., None], (y - y0)[..., None]
        x0, y0 = x0.long(), y0.long()
        g = lambda yy, xx: t[bidx, yy % hl, xx % wl]
        return ((1 - fx) * (1 - fy) * g(y0, x0) + fx * (1 - fy) * g(y0, x0 + 1)) + (1 - fx) * fy * g(y0 + 1, x0) + fx * fy * g(y0 + 1, x0 + 1)

    if n_levels == 1:
        return tap(0)
    dsdx, dsdy, dtdx, dtdy = uv_da[..., 0] * tw, uv_da[..., 1] * tw, uv_da[..., 2] * th, uv_da[..., 3] * th
    A, Bq, Cq = dsdx ** 2 + dtdx ** 2, dsdy ** 2 + dtdy ** 2, dsdx * dsdy + dtdx * dtdy
    major = 0.5 * (A + Bq) + torch.sqrt(0.25 * (A - Bq) ** 2 + Cq ** 2)
    level = torch.where(major > 0, 0.5 * torch.log2(major.clamp(min=1e-30)), torch.zeros_like(major)).clamp(0, n_levels - 1)
    l0 = torch.floor(level).long()
    f = (level - l0)[..., None]
    l1 = (l0 + 1).clamp(max=n_levels - 1)
    out = torch.zeros(B, H, W, C, dtype=tex.dtype)
    for l in range(n_levels):
        s = None
        m0, m1 = (l0 == l)[..., None], ((l1 == l) & (l1 != l0))[..., None]
        if m0.any() or m1.any():
            s = tap(l)
            out = out + torch.where(m0, (1 - torch.where(l1[..., None] != l0[..., None], f, torch.zeros_like(f))) * s, torch.zeros_like(s))
            out = out + torch.where(m1, f * s, torch.zeros_like(s))
    return out
