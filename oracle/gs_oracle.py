"""Dense PyTorch restatement of the 3D Gaussian-splatting rasteriser -- the checker for mvedit_b200.gs_renderer.

TEST INFRASTRUCTURE ONLY (see oracle/raymarching_oracle.c header).

The reference snapshot does not contain its 3DGS code (README.md:121 names "3DGS" and ashawkey/diff-gaussian-rasterization as the
upstream of the withheld renderer): there is NO reference implementation to pin this against -- PARITY UNPINNED by
construction.  What is restated is the public algorithm of diff-gaussian-rasterization (Kerbl et al. 2023; the ashawkey fork adds the
depth and alpha outputs), with its constants: low-pass +0.3 px on the 2-D covariance, radius = ceil(3 sqrt(lambda_max)), 16 x 16 tiles
touched by the radius rect, near plane 0.2, alpha = min(0.99, o exp(power)), alpha < 1/255 skipped, blending stops BEFORE T would drop
below 1e-4, background composited with the final T.  Every pixel is evaluated against every Gaussian (depth order, masks for the tile
rect and the skip rules) with plain differentiable torch ops, so autograd supplies the backward oracle; the clamp at 0.99 passes the
gradient straight through, as the public backward does.
"""
import math

import torch
import torch.nn.functional as F

TILE = 16


def quat_to_rotmat(q):
    q = F.normalize(q, dim=-1)
    w, x, y, z = q.unbind(-1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=-1).reshape(q.shape[:-1] + (3, 3))


def preprocess(means3D, scales, quats, viewmat, K, H, W):
    """Per-Gaussian projection (preprocessCUDA of the public implementation, pinhole intrinsics instead of a projection matrix).
    viewmat [4,4] world -> camera (OpenCV: x right, y down, z forward); K = (fx, fy, cx, cy) with pixel centres at i + 0.5.
    -> dict(xy [P,2] in pixel-INDEX coordinates, conic [P,3] = (A, B, C), depth [P], radius [P] int, rect [P,4] int (x0,y0,x1,y1), valid [P])."""
    fx, fy, cx, cy = [float(v) for v in K]
    R, t = viewmat[:3, :3], viewmat[:3, 3]
    pc = means3D @ R.t() + t
    x, y, z = pc.unbind(-1)
    valid = z > 0.2
    zs = torch.where(valid, z, torch.ones_like(z))
    limx, limy = 1.3 * W / (2 * fx), 1.3 * H / (2 * fy)
    tx = torch.minimum(torch.maximum(x / zs, x.new_tensor(-limx)), x.new_tensor(limx)) * zs
    ty = torch.minimum(torch.maximum(y / zs, y.new_tensor(-limy)), y.new_tensor(limy)) * zs
    zero = torch.zeros_like(zs)
    J = torch.stack([fx / zs, zero, -fx * tx / (zs * zs), zero, fy / zs, -fy * ty / (zs * zs)], dim=-1).reshape(-1, 2, 3)
    M = quat_to_rotmat(quats) * scales[:, None, :]
    Sigma = M @ M.transpose(1, 2)
    Tm = J @ R
    cov = Tm @ Sigma @ Tm.transpose(1, 2)
    a, b, c = cov[:, 0, 0] + 0.3, cov[:, 0, 1], cov[:, 1, 1] + 0.3
    det = a * c - b * b
    valid = valid & (det > 0)
    dets = torch.where(valid, det, torch.ones_like(det))
    conic = torch.stack([c / dets, -b / dets, a / dets], dim=-1)
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius = torch.ceil(3.0 * torch.sqrt(lam)).to(torch.int32)
    xy = torch.stack([fx * x / zs + cx - 0.5, fy * y / zs + cy - 0.5], dim=-1)
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    r = radius.to(xy.dtype)
    tr = lambda v: torch.trunc(v).to(torch.int32)
    x0 = tr((xy[:, 0] - r) / TILE).clamp(0, gx); x1 = tr((xy[:, 0] + r + TILE - 1) / TILE).clamp(0, gx)
    y0 = tr((xy[:, 1] - r) / TILE).clamp(0, gy); y1 = tr((xy[:, 1] + r + TILE - 1) / TILE).clamp(0, gy)
    rect = torch.stack([x0, y0, x1, y1], dim=-1)
    rect = torch.where(valid[:, None], rect, torch.zeros_like(rect))
    valid = valid & ((rect[:, 2] - rect[:, 0]) * (rect[:, 3] - rect[:, 1]) > 0)
    return dict(xy=xy, conic=conic, depth=z, radius=radius, rect=rect, valid=valid)


def blend(xy, conic, depth, opacity, colors, rect, valid, bg, H, W):
    """Front-to-back blending of every pixel against every valid Gaussian in depth order (renderCUDA).
    -> color [H,W,3], depth [H,W], alpha [H,W]."""
    dev = xy.device
    order = torch.argsort(torch.where(valid, depth.detach(), depth.new_tensor(float('inf'))), stable=True)
    order = order[valid[order]]
    ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing='ij')
    fxp, fyp = xs.to(xy.dtype), ys.to(xy.dtype)
    tx, ty = xs // TILE, ys // TILE
    T = torch.ones(H, W, dtype=xy.dtype, device=dev)
    done = torch.zeros(H, W, dtype=torch.bool, device=dev)
    C = torch.zeros(H, W, 3, dtype=xy.dtype, device=dev)
    D = torch.zeros(H, W, dtype=xy.dtype, device=dev)
    for g in order.tolist():
        dx, dy = xy[g, 0] - fxp, xy[g, 1] - fyp
        power = -0.5 * (conic[g, 0] * dx * dx + conic[g, 2] * dy * dy) - conic[g, 1] * dx * dy
        raw = opacity[g] * torch.exp(power)
        alpha = raw + (torch.clamp(raw, max=0.99) - raw).detach()           # min(0.99, .) with a straight-through gradient
        in_rect = (tx >= rect[g, 0]) & (tx < rect[g, 2]) & (ty >= rect[g, 1]) & (ty < rect[g, 3])
        ok = in_rect & (power <= 0) & (alpha >= 1.0 / 255.0) & ~done
        test_T = T * (1 - alpha)
        stop = ok & (test_T < 1e-4)
        done = done | stop
        ok = ok & ~stop
        w = torch.where(ok, alpha * T, torch.zeros_like(T))
        C = C + w[..., None] * colors[g]
        D = D + w * depth[g]
        T = torch.where(ok, test_T, T)
    return C + T[..., None] * bg, D, 1 - T


def render(means3D, scales, quats, opacities, colors, viewmat, K, H, W, bg):
    pre = preprocess(means3D, scales, quats, viewmat, K, H, W)
    return blend(pre['xy'], pre['conic'], pre['depth'], opacities.reshape(-1), colors, pre['rect'], pre['valid'], bg, H, W)
