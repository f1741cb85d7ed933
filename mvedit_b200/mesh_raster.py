"""Triangle-mesh rasteriser ops on libmvedit_b200 -- seam B5: the ``nvdiffrast.torch`` surface that
``MeshRenderer`` uses (``lib/models/decoders/mesh_renderer/base_mesh_renderer.py:5,204,241-298,407-410,442-500,521-577``), same names,
argument order and return conventions, so that module reads ``from mvedit_b200 import mesh_raster as dr``:

    glctx = dr.RasterizeCudaContext()
    rast, rast_db = dr.rasterize(glctx, pos[B,V,4], tri[F,3] int32, resolution=(h, w), ranges=None, grad_db=True)
    tex_out       = dr.texture(tex[B or 1,th,tw,C], uv[B,h,w,2], uv_da=out_da, filter_mode='linear-mipmap-linear')
    out, out_da   = dr.interpolate(attr[B or 1,V,C], rast, tri, rast_db=None, diff_attrs=None)
    color_aa      = dr.antialias(color[B,h,w,C], rast, pos, tri)

All four are differentiable the way the reference relies on: ``interpolate`` w.r.t. attributes and ``rast`` (u, v); ``rasterize``
w.r.t. ``pos`` through (u, v, z/w); ``antialias`` w.r.t. colours and ``pos`` (the silhouette gradient of ``mesh_optim``'s alpha loss,
``mvedit_3d_pipeline.py:768-770``).  Not carried (they raise / are marked non-differentiable): gradients through ``rast_db`` /
``out_da`` (only a mip-mapped texture fetch under a geometry gradient would need them; MVEdit optimises either geometry or texture),
range mode (``ranges=``, only reached with ``num_scenes > 1``, ``base_mesh_renderer.py:301-381``).

CUDA only: tensors must live on the GPU and the library must be built; there is no CPU path in the product (the CPU test-suite
drives the same per-pixel code through ``tests/host_harness.py``, which compiles the kernel source as plain C++).
"""
import torch

from ._lib import call, ptr, stream, c_u32, c_int


class RasterizeCudaContext:
    """Owns the rasteriser's scratch (64-bit z-buffer, large-triangle queue), re-used across calls of the same size like nvdiffrast's
    context owns its buffers (``base_mesh_renderer.py:204``)."""

    def __init__(self, device=None):
        self.device = device
        self._zbuf = None
        self._queue = None

    def scratch(self, n_pix, n_tri_inst, device):
        if self._zbuf is None or self._zbuf.numel() < n_pix or self._zbuf.device != device:
            self._zbuf = torch.empty(n_pix, dtype=torch.int64, device=device)
        if self._queue is None or self._queue.numel() < n_tri_inst + 1 or self._queue.device != device:
            self._queue = torch.empty(n_tri_inst + 1, dtype=torch.int32, device=device)
        return self._zbuf, self._queue


RasterizeGLContext = RasterizeCudaContext      # the reference picks one by a flag (base_mesh_renderer.py:204); there is one rasteriser here


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


def _i32c(t):
    return t.detach().to(torch.int32).contiguous()


def edge_opposites(tri, method='sort'):
    """opp [F,3] int32: for the edge facing vertex k of triangle f, the vertex of the adjacent triangle that is not on the edge
    (-1 on an open edge) -- the topology ``antialias`` needs to tell silhouette edges from interior ones.  One stable sort of the
    3F edge keys; edges used by more than two triangles pair their first two users, later users see the first.
    ``method='hash'`` (opt-in): the same table from three launches of an open-addressing hash kernel instead of the sort
    (``mve_edge_opposites``; CPU-checked against this function, timed by bench.py's mesh child process)."""
    if method == 'hash':
        return _edge_opposites_hash(tri)
    t = tri.detach().long()
    F = t.shape[0]
    if F == 0:
        return torch.zeros(0, 3, dtype=torch.int32, device=tri.device)
    a, b = t[:, [1, 2, 0]].reshape(-1), t[:, [2, 0, 1]].reshape(-1)
    n_v = int(t.max()) + 1
    key = torch.minimum(a, b) * n_v + torch.maximum(a, b)
    ks, order = torch.sort(key, stable=True)
    n = ks.numel()
    first = torch.ones(n, dtype=torch.bool, device=t.device)
    first[1:] = ks[1:] != ks[:-1]
    starts = first.nonzero().reshape(-1)
    run = torch.cumsum(first.long(), 0) - 1
    run_len = torch.diff(torch.cat([starts, starts.new_tensor([n])]))[run]
    idx = torch.arange(n, device=t.device)
    in_run = idx - starts[run]
    partner = torch.where(in_run == 0, idx + 1, torch.where(in_run == 1, idx - 1, starts[run]))
    has = run_len >= 2
    own = t.reshape(-1)[order]                                   # vertex k of triangle f, in sorted order
    opp_sorted = torch.where(has, own[partner.clamp(max=n - 1)], torch.full_like(own, -1))
    opp = torch.empty(n, dtype=torch.long, device=t.device)
    opp[order] = opp_sorted
    return opp.reshape(F, 3).to(torch.int32)


def _edge_opposites_hash(tri):
    tri_c = _i32c(tri)
    F = tri_c.shape[0]
    dev = tri_c.device
    opp = torch.empty(F, 3, dtype=torch.int32, device=dev)
    if F == 0:
        return opp
    slots = 1 << max(int(6 * F - 1).bit_length(), 4)
    keys = torch.empty(slots, dtype=torch.int64, device=dev)
    first, second = torch.empty(slots, dtype=torch.int32, device=dev), torch.empty(slots, dtype=torch.int32, device=dev)
    slot_of = torch.empty(3 * F, dtype=torch.int32, device=dev)
    call('mve_edge_opposites', ptr(tri_c), c_u32(F), c_u32(slots), ptr(keys), ptr(first), ptr(second), ptr(slot_of), ptr(opp), stream())
    return opp


class _RasterizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, glctx, pos, tri, H, W, grad_db):
        B, V, _ = pos.shape
        F = tri.shape[0]
        pos_c, tri_c = _f32c(pos), _i32c(tri)
        zbuf, queue = glctx.scratch(B * H * W, B * F, pos.device)
        rast = torch.empty(B, H, W, 4, dtype=torch.float32, device=pos.device)
        rast_db = torch.empty(B, H, W, 4, dtype=torch.float32, device=pos.device)
        call('mve_rasterize_fwd', ptr(pos_c), ptr(tri_c), c_u32(B), c_u32(V), c_u32(F), c_u32(H), c_u32(W), c_int(1), ptr(zbuf), ptr(queue),
             ptr(rast), ptr(rast_db), stream())
        ctx.save_for_backward(pos_c, tri_c, rast)
        ctx.mark_non_differentiable(rast_db)
        return rast, rast_db

    @staticmethod
    def backward(ctx, g_rast, g_db):
        pos_c, tri_c, rast = ctx.saved_tensors
        B, H, W, _ = rast.shape
        g_pos = torch.zeros_like(pos_c)
        g_rast_c = _f32c(g_rast)          # a named reference: a temporary passed as ptr(...) would be freed before the kernel reads it
        call('mve_rasterize_bwd', ptr(pos_c), ptr(tri_c), c_u32(B), c_u32(pos_c.shape[1]), c_u32(tri_c.shape[0]), c_u32(H), c_u32(W), c_int(1),
             ptr(rast), ptr(g_rast_c), ptr(g_pos), stream())
        return None, g_pos, None, None, None, None


def rasterize(glctx, pos, tri, resolution, ranges=None, grad_db=True):
    """``dr.rasterize`` -> (rast [B,h,w,4] = (u, v, z/w, triangle id + 1), rast_db [B,h,w,4] = (du/dX, du/dY, dv/dX, dv/dY)).
    Like nvdiffrast's CUDA context, ``rast_db`` is produced whatever ``grad_db`` says (the flag only governs gradient flow THROUGH
    rast_db, which is not carried here at all): the reference rasterises texture space with ``grad_db=False`` and still feeds
    ``tex_rast_db`` to ``interpolate`` (base_mesh_renderer.py:442,496-497)."""
    if ranges is not None or pos.dim() != 3:
        raise NotImplementedError('mesh_raster.rasterize: range mode (2-D pos + ranges) is not built; MVEdit uses instance mode')
    if pos.shape[-1] != 4 or tri.dim() != 2 or tri.shape[1] != 3:
        raise ValueError('rasterize: pos must be [B,V,4] and tri [F,3]')
    H, W = int(resolution[0]), int(resolution[1])
    return _RasterizeFn.apply(glctx, pos, tri, H, W, bool(grad_db))


class _InterpolateFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, attr, rast, tri, rast_db, want_da):
        B, H, W, _ = rast.shape
        attr_c, rast_c, tri_c = _f32c(attr), _f32c(rast), _i32c(tri)
        Ba, Va, C = attr_c.shape
        if Ba not in (1, B):
            raise ValueError('interpolate: attr batch must be 1 or %d, got %d' % (B, Ba))
        out = torch.empty(B, H, W, C, dtype=torch.float32, device=rast.device)
        out_da = torch.empty(B, H, W, 2 * C, dtype=torch.float32, device=rast.device) if want_da else None
        db_c = _f32c(rast_db) if want_da else None
        call('mve_interpolate_fwd', ptr(attr_c), ptr(tri_c), ptr(rast_c), ptr(db_c), c_u32(B), c_u32(H), c_u32(W), c_u32(Va), c_u32(tri_c.shape[0]),
             c_u32(C), c_int(0 if Ba == 1 else 1), ptr(out), ptr(out_da), stream())
        ctx.save_for_backward(attr_c, rast_c, tri_c)
        if out_da is None:
            out_da = torch.zeros(B, H, W, 0, dtype=torch.float32, device=rast.device)
        ctx.mark_non_differentiable(out_da)
        return out, out_da

    @staticmethod
    def backward(ctx, g_out, g_da):
        attr_c, rast_c, tri_c = ctx.saved_tensors
        B, H, W, _ = rast_c.shape
        Ba, Va, C = attr_c.shape
        g_attr = torch.zeros_like(attr_c)
        g_rast = torch.empty_like(rast_c) if ctx.needs_input_grad[1] else None
        g_out_c = _f32c(g_out)
        call('mve_interpolate_bwd', ptr(attr_c), ptr(tri_c), ptr(rast_c), c_u32(B), c_u32(H), c_u32(W), c_u32(Va), c_u32(tri_c.shape[0]), c_u32(C),
             c_int(0 if Ba == 1 else 1), ptr(g_out_c), ptr(g_attr), ptr(g_rast), stream())
        return g_attr, g_rast, None, None, None


def interpolate(attr, rast, tri, rast_db=None, diff_attrs=None):
    """``dr.interpolate`` -> (out [B,h,w,C], out_da [B,h,w,2C] with (d/dX, d/dY) per attribute, or [B,h,w,0])."""
    if attr.dim() == 2:
        raise NotImplementedError('mesh_raster.interpolate: range mode (2-D attr) is not built')
    want_da = diff_attrs is not None
    if want_da:
        if diff_attrs != 'all' and list(diff_attrs) != list(range(attr.shape[-1])):
            raise NotImplementedError("interpolate: diff_attrs must be None or 'all'")
        if rast_db is None or rast_db.shape[-1] != 4:
            raise ValueError('interpolate: diff_attrs needs the rast_db of rasterize(grad_db=True)')
    return _InterpolateFn.apply(attr, rast, tri, rast_db if want_da else None, want_da)


class _AntialiasFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, color, rast, pos, tri, opp):
        B, H, W, C = color.shape
        color_c, rast_c, pos_c, tri_c, opp_c = _f32c(color), _f32c(rast), _f32c(pos), _i32c(tri), _i32c(opp)
        out = torch.empty_like(color_c)
        call('mve_antialias_fwd', ptr(color_c), ptr(rast_c), ptr(pos_c), ptr(tri_c), ptr(opp_c), c_u32(B), c_u32(H), c_u32(W), c_u32(C),
             c_u32(pos_c.shape[1]), c_u32(tri_c.shape[0]), c_int(1), ptr(out), stream())
        ctx.save_for_backward(color_c, rast_c, pos_c, tri_c, opp_c)
        return out

    @staticmethod
    def backward(ctx, g_out):
        color_c, rast_c, pos_c, tri_c, opp_c = ctx.saved_tensors
        B, H, W, C = color_c.shape
        g_color = torch.empty_like(color_c)
        g_pos = torch.zeros_like(pos_c) if ctx.needs_input_grad[2] else None
        g_out_c = _f32c(g_out)
        call('mve_antialias_bwd', ptr(color_c), ptr(rast_c), ptr(pos_c), ptr(tri_c), ptr(opp_c), c_u32(B), c_u32(H), c_u32(W), c_u32(C),
             c_u32(pos_c.shape[1]), c_u32(tri_c.shape[0]), c_int(1), ptr(g_out_c), ptr(g_color), ptr(g_pos), stream())
        return g_color, None, g_pos, None, None


def antialias(color, rast, pos, tri, topology_hash=None, pos_gradient_boost=1.0):
    """``dr.antialias``; ``topology_hash`` may carry a precomputed ``edge_opposites(tri)`` (nvdiffrast's hash has the same role)."""
    if pos.dim() != 3:
        raise NotImplementedError('mesh_raster.antialias: range mode (2-D pos) is not built')
    if pos_gradient_boost != 1.0:
        raise NotImplementedError('antialias: pos_gradient_boost != 1 is not used by the reference and not built')
    opp = edge_opposites(tri) if topology_hash is None else topology_hash
    return _AntialiasFn.apply(color, rast, pos, tri, opp)


def antialias_construct_topology_hash(tri):
    return edge_opposites(tri)


def _n_mip_levels(th, tw, max_mip_level=None):
    n = 1
    while th % 2 == 0 and tw % 2 == 0 and n < 16 and (max_mip_level is None or n <= max_mip_level):
        th, tw, n = th // 2, tw // 2, n + 1
    return n


class _TextureFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tex, uv, uv_da, n_levels):
        Bt, th, tw, C = tex.shape
        B, H, W, _ = uv.shape
        uv_c = _f32c(uv)
        da_c = _f32c(uv_da) if uv_da is not None else None
        lib_total = _pyramid_floats(Bt, th, tw, C, n_levels)
        pyr = torch.empty(lib_total, dtype=torch.float32, device=tex.device)
        pyr[:Bt * th * tw * C].copy_(tex.detach().to(torch.float32).reshape(-1))
        if n_levels > 1:
            call('mve_texture_mip_build', ptr(pyr), c_u32(Bt), c_u32(th), c_u32(tw), c_u32(C), c_u32(n_levels), stream())
        out = torch.empty(B, H, W, C, dtype=torch.float32, device=tex.device)
        call('mve_texture_fwd', ptr(pyr), c_u32(Bt), c_u32(th), c_u32(tw), c_u32(C), c_u32(n_levels), ptr(uv_c), ptr(da_c), c_u32(B), c_u32(H),
             c_u32(W), ptr(out), stream())
        ctx.save_for_backward(uv_c, da_c)
        ctx.dims = (Bt, th, tw, C, n_levels, lib_total)
        return out

    @staticmethod
    def backward(ctx, g_out):
        uv_c, da_c = ctx.saved_tensors
        Bt, th, tw, C, n_levels, total = ctx.dims
        B, H, W, _ = uv_c.shape
        g_pyr = torch.zeros(total, dtype=torch.float32, device=uv_c.device)
        g_out_c = _f32c(g_out)
        call('mve_texture_bwd', c_u32(Bt), c_u32(th), c_u32(tw), c_u32(C), c_u32(n_levels), ptr(uv_c), ptr(da_c), c_u32(B), c_u32(H), c_u32(W),
             ptr(g_out_c), ptr(g_pyr), stream())
        return g_pyr[:Bt * th * tw * C].view(Bt, th, tw, C), None, None, None


def _pyramid_floats(Bt, th, tw, C, n_levels):
    total = sum(Bt * (th >> l) * (tw >> l) * C for l in range(n_levels))
    if total >= 2 ** 31:
        raise ValueError('texture: the mip pyramid exceeds 2^31 floats')
    return total


def texture(tex, uv, uv_da=None, mip_level_bias=None, mip=None, filter_mode='auto', boundary_mode='wrap', max_mip_level=None):
    """``dr.texture``: tex [B or 1, th, tw, C], uv [B,h,w,2], uv_da [B,h,w,4] (``interpolate(..., diff_attrs='all')``'s second output)
    -> [B,h,w,C].  ``filter_mode`` 'linear' or 'linear-mipmap-linear' ('auto' = the latter when uv_da is given); boundary mode 'wrap'.
    Differentiable w.r.t. ``tex`` (the reference bakes textures through exactly that gradient, base_mesh_renderer.py:470-475); uv and
    uv_da receive no gradient."""
    if filter_mode == 'auto':
        filter_mode = 'linear-mipmap-linear' if uv_da is not None else 'linear'
    if filter_mode not in ('linear', 'linear-mipmap-linear') or boundary_mode != 'wrap' or mip is not None or mip_level_bias is not None:
        raise NotImplementedError("texture: only filter_mode 'linear' / 'linear-mipmap-linear' with boundary_mode='wrap' is built")
    if tex.dim() != 4 or uv.dim() != 4 or uv.shape[-1] != 2 or tex.shape[0] not in (1, uv.shape[0]):
        raise ValueError('texture: tex must be [B or 1, th, tw, C] and uv [B, h, w, 2]')
    if uv.requires_grad or (uv_da is not None and uv_da.requires_grad):
        raise NotImplementedError('texture: gradients w.r.t. uv / uv_da are not built (MVEdit optimises either geometry or texture)')
    use_mip = filter_mode == 'linear-mipmap-linear'
    if use_mip and (uv_da is None or uv_da.shape[-1] != 4):
        raise ValueError("texture: 'linear-mipmap-linear' needs uv_da [B,h,w,4]")
    n_levels = _n_mip_levels(tex.shape[1], tex.shape[2], max_mip_level) if use_mip else 1
    return _TextureFn.apply(tex, uv, uv_da if use_mip else None, n_levels)
