"""Host-side mirror of the reference's Instant-NGP decoder + volume renderer.

Mirrors, name for name and argument for argument,
  * ``iNGPDecoder`` / ``MLP``        /root/reference/lib/models/decoders/ingp_decoder.py:20-125
  * ``VolumeRenderer``               /root/reference/lib/models/decoders/base_volume_renderer.py:17-343
  * the ``tcnn.Encoding`` surface the reference touches (``.params``, ``.n_output_dims``; ingp_decoder.py:62-74,88)
so that ``decoder.state_dict()`` carries the reference's keys (``aabb``, ``encoder.params``, ``mlp.net.{0,1}.{weight,bias}`` --
the cross-stage ``ingp_states`` contract) and ``decoder(rays_o, rays_d, code, density_bitfield, grid_size, ...)``
returns the same dict.  The arithmetic is libmvedit_b200.so: hash grid + MLP + activations are ONE fused kernel per direction
(mve_field_forward/backward), marching/compositing are the B3 kernels, the inference while-loop is ONE kernel (mve_render_rays).
"""
import ctypes
import math
from copy import deepcopy

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from . import raymarching as rm
from ._lib import call, ptr, stream, get_lib, c_int, c_u32, c_f32


def level_table(n_levels=12, base_resolution=16, max_resolution=320, bound=1.0, log2_hashmap_size=19):
    """tiny-cuda-nn grid.h (grid_scale / grid_resolution / params_in_level) for the config of ingp_decoder.py:62-72."""
    pls = np.exp2(np.log2(max_resolution * bound / base_resolution) / (n_levels - 1))
    log2_pls = np.float32(math.log2(pls))
    scale, res, size, off = [], [], [], []
    o = 0
    for l in range(n_levels):
        s = np.float32(np.exp2(np.float32(l) * log2_pls, dtype=np.float32) * np.float32(base_resolution) - np.float32(1.0))
        r = int(math.ceil(float(s))) + 1
        n = min((r ** 3 + 7) // 8 * 8, 1 << log2_hashmap_size)
        scale.append(float(s)); res.append(r); size.append(n); off.append(o)
        o += n
    return dict(scale=np.array(scale, np.float32), res=np.array(res, np.uint32), size=np.array(size, np.uint32),
                off=np.array(off, np.uint32), n_entries=o, n_levels=n_levels)


class _LevelArgs:
    """ctypes views of the per-level host arrays expected by the C ABI."""

    def __init__(self, lt):
        self.lt = lt
        self.n = c_u32(lt['n_levels'])
        self.scale = lt['scale'].ctypes.data_as(ctypes.c_void_p)
        self.res = lt['res'].ctypes.data_as(ctypes.c_void_p)
        self.size = lt['size'].ctypes.data_as(ctypes.c_void_p)
        self.off = lt['off'].ctypes.data_as(ctypes.c_void_p)

    def args(self):
        return (self.n, self.scale, self.res, self.size, self.off)


class _HashGridFn(Function):
    """enc = HashGrid(x01): differentiable w.r.t. the table and the positions (mve_hashgrid_forward / _backward)."""

    @staticmethod
    def forward(ctx, x01, params, enc):
        x01 = x01.float().contiguous()
        M = x01.shape[0]
        out = torch.empty(M, enc.n_output_dims, dtype=torch.float32, device=x01.device)
        call('mve_hashgrid_forward', ptr(x01), c_u32(M), ptr(params), *enc._largs.args(), ptr(out), stream())
        ctx.save_for_backward(x01, params)
        ctx.enc = enc
        return out

    @staticmethod
    def backward(ctx, g):
        x01, params = ctx.saved_tensors
        g_table = torch.zeros_like(params) if ctx.needs_input_grad[1] else None
        g_x = torch.zeros_like(x01) if ctx.needs_input_grad[0] else None
        call('mve_hashgrid_backward', ptr(x01), c_u32(x01.shape[0]), ptr(params), *ctx.enc._largs.args(), ptr(g.float().contiguous()), ptr(g_table),
             ptr(g_x), stream())
        return g_x, g_table, None


class HashGridEncoding(nn.Module):
    """The slice of ``tinycudann.Encoding`` the reference uses (seam B4): a flat fp32 ``params`` vector, ``n_output_dims`` and the
    call ``enc(x [M,3] in [0,1]) -> [M, 2 * n_levels]``, differentiable w.r.t. ``params`` and ``x`` (ingp_decoder.py:62-74,112;
    triplane_ingp_decoder.py:102-114,150).  iNGPDecoder does not go through this call -- its encoding is fused with the MLP in
    mve_field_forward / _backward -- TriPlaneiNGPDecoder does."""

    def __init__(self, n_levels=12, base_resolution=16, max_resolution=320, bound=1.0, log2_hashmap_size=19):
        super().__init__()
        self.levels = level_table(n_levels, base_resolution, max_resolution, bound, log2_hashmap_size)
        self.n_output_dims = 2 * n_levels
        self.params = nn.Parameter(torch.zeros(self.levels['n_entries'] * 2, dtype=torch.float32))
        self._largs = _LevelArgs(self.levels)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError('mvedit_b200 ops need CUDA tensors (no CPU fallback)')
        return _HashGridFn.apply(x, self.params, self)


class MLP(nn.Module):
    """ingp_decoder.py:20-40 (kept as an nn.Module so the state-dict keys are mlp.net.{l}.{weight,bias})."""

    def __init__(self, dim_in, dim_out, dim_hidden, num_layers, bias=True):
        super().__init__()
        assert num_layers == 2 and dim_hidden == 64 and dim_out == 4 and bias, \
            'the fused sm_100a field kernel implements the configuration MVEdit instantiates (2 layers, 64 hidden, 4 out)'
        self.dim_in, self.dim_out, self.dim_hidden, self.num_layers = dim_in, dim_out, dim_hidden, num_layers
        self.net = nn.ModuleList([nn.Linear(dim_in, dim_hidden, bias=True), nn.Linear(dim_hidden, dim_out, bias=True)])


class _FieldFn(Function):
    """Fused point_decode: (xyz, table, w1, b1, w2, b2) -> (sigma, rgb)."""

    @staticmethod
    def forward(ctx, xyz, table, w1, b1, w2, b2, dec, density_only, m_dev=None):
        xyz = xyz.float().contiguous()
        M = xyz.shape[0]
        # capacity buffers (m_dev given): every consumer (composite fwd / bwd, cull, field backward) clamps to *m_dev, so the tail is
        # never read and needs no fill
        sigma = torch.empty(M, dtype=torch.float32, device=xyz.device)
        rgb = None if density_only else torch.empty(M, 3, dtype=torch.float32, device=xyz.device)
        # kernel mode: 0/1 = fp32 FFMA MLP (full / density), 2/3 = TF32 tensor-core MLP (density / full)
        tf32 = bool(getattr(dec, 'mlp_tf32', True))
        mode = (2 if tf32 or density_only == 2 else 1) if density_only else (3 if tf32 else 0)
        if M > 0:
            call('mve_field_forward', ptr(xyz), c_u32(M), ptr(m_dev), ptr(table), ptr(w1), ptr(b1), ptr(w2), ptr(b2),
                 *dec.encoder._largs.args(), c_f32(dec.bound), c_f32(dec.blob_density), c_f32(dec.blob_radius),
                 c_f32(dec.sigmoid_saturation), c_int(mode), ptr(sigma), ptr(rgb), stream())
        ctx.save_for_backward(xyz, table, w1, b1, w2, b2)
        ctx.dec = dec
        ctx.density_only = density_only
        ctx.tf32 = tf32
        ctx.m_dev = m_dev
        if density_only:
            empty = sigma.new_zeros(0)
            ctx.mark_non_differentiable(empty)
            return sigma, empty
        return sigma, rgb

    @staticmethod
    def backward(ctx, g_sigma, g_rgb):
        xyz, table, w1, b1, w2, b2 = ctx.saved_tensors
        dec = ctx.dec
        M = xyz.shape[0]
        sink = dec.grad_sink            # FusedAdam: the kernels accumulate straight into the optimizer's flat gradient buffer
        if sink is not None:
            prm = dec._field_params()
            g_table, g_w1, g_b1, g_w2, g_b2 = (sink.grad_sink(q) for q in prm)
        else:
            g_table = torch.zeros_like(table)
            g_w1, g_b1, g_w2, g_b2 = torch.empty_like(w1), torch.empty_like(b1), torch.empty_like(w2), torch.empty_like(b2)
        need_dx = ctx.needs_input_grad[0]
        g_xyz = torch.zeros_like(xyz) if need_dx else None
        ws = dec._workspace(xyz.device)
        g_sigma = g_sigma.float().contiguous()
        g_rgb = None if (ctx.density_only or g_rgb is None) else g_rgb.float().contiguous()
        call('mve_field_backward', ptr(xyz), c_u32(M), ptr(ctx.m_dev), ptr(table), ptr(w1), ptr(b1), ptr(w2), ptr(b2),
             *dec.encoder._largs.args(), c_f32(dec.bound), c_f32(dec.blob_density), c_f32(dec.blob_radius),
             c_f32(dec.sigmoid_saturation), ptr(g_sigma), ptr(g_rgb), ptr(g_table), ptr(g_w1), ptr(g_b1), ptr(g_w2), ptr(g_b2),
             c_int(int(sink is not None)), c_int(int(ctx.tf32)), ptr(ws), ptr(g_xyz), stream())
        if sink is not None:
            return g_xyz, None, None, None, None, None, None, None, None
        return g_xyz, g_table, g_w1, g_b1, g_w2, g_b2, None, None, None


class iNGPDecoder(nn.Module):
    """ingp_decoder.py:43-125 on top of VolumeRenderer (base_volume_renderer.py:17-343)."""

    def __init__(self, bound=1, min_near=0.2, bg_radius=-1, max_steps=256, weight_culling_th=0.0,
                 base_resolution=16, max_resolution=320, n_levels=12, num_layers=2, hidden_dim=64,
                 sigmoid_saturation=0.001, blob_density=1.0, blob_radius=0.2):
        super().__init__()
        assert bg_radius <= 0, 'background sphere is not on the MVEdit path (bg_radius=-1, SURVEY.md §2.2)'
        self.bound = bound
        self.min_near = min_near
        self.bg_radius = bg_radius
        self.max_steps = max_steps
        self.weight_culling_th = weight_culling_th
        self.base_resolution, self.max_resolution, self.n_levels = base_resolution, max_resolution, n_levels
        self.register_buffer('aabb', torch.FloatTensor([-bound, -bound, -bound, bound, bound, bound]))
        self.encoder = HashGridEncoding(n_levels, base_resolution, max_resolution, bound)
        self.in_dim = self.encoder.n_output_dims
        self.mlp = MLP(self.in_dim, 4, hidden_dim, num_layers, bias=True)
        self.sigmoid_saturation = sigmoid_saturation
        self.blob_density = blob_density
        self.blob_radius = blob_radius
        self.state_dict_bak = None
        self.sample_capacity = 0      # > 0: sync-free training forward with capacity-sized sample buffers
        # MLP matmuls on tensor cores in TF32 -- what the reference runs (torch.backends.cuda.matmul.allow_tf32 = True,
        # lib/apis/adapter3d.py:51-61); False selects the fp32 FFMA kernels (exact-parity tests)
        self.mlp_tf32 = True
        self._ws = None
        self._grid_cache = {}
        self.grad_sink = None          # set by nerf_optim to a FusedAdam: field backward accumulates into its flat gradient buffer
        self._ov = None                # device max of the post-cull sample count over the iterations since the last check
        self._ov_host = None
        self._jitter_gen = None        # generator of the occupancy-refresh jitter (seeded identically on every rank)
        self.test_noise = None         # parity tests: dict(march=[N] tensor, grid=[H^3,3] tensor) replaces the internal random draws
        self.init_weights()

    # ------------------------------------------------------------------ parameters / state
    def init_weights(self):
        """ingp_decoder.py:87-91 (mmcv xavier_init: xavier_uniform_, bias 0)."""
        self.encoder.params.data.uniform_(-1e-4, 1e-4)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight, gain=1)
                nn.init.constant_(m.bias, 0)

    def backup_state_dict(self):
        self.state_dict_bak = deepcopy(self.state_dict())

    def restore_state_dict(self):
        if self.state_dict_bak is None:
            raise RuntimeError("No backup state dict found")
        self.load_state_dict(self.state_dict_bak)

    def _workspace(self, device):
        if self._ws is None or self._ws.device != device:
            fn = get_lib().mve_field_backward_workspace_floats
            fn.restype = ctypes.c_uint32
            self._ws = torch.empty(int(fn(c_u32(self.n_levels))), dtype=torch.float32, device=device)
        return self._ws

    # ------------------------------------------------------------------ capacity bookkeeping (sync-free training path)
    def _track_overflow(self, counter):
        """counter: device int32 [1] = TRUE number of samples that passed the cull in this forward (it keeps counting past the
        capacity, mve_cull_samples).  A running device max is kept; nothing is read back here."""
        if self._ov is None or self._ov.device != counter.device:
            self._ov = torch.zeros(1, dtype=torch.int32, device=counter.device)
            self._ov_host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self._ov_event = None
        torch.maximum(self._ov, counter, out=self._ov)

    def note_sample_overflow(self):
        """End of a nerf_optim call: start an async copy of the running max to pinned host memory (no sync)."""
        if self._ov is None:
            return
        self._ov_host.copy_(self._ov, non_blocking=True)
        self._ov.zero_()                                   # the running max restarts: the next call may use another capacity
        self._ov_cap = int(self.sample_capacity)
        self._ov_event = torch.cuda.Event()
        self._ov_event.record()

    def check_sample_overflow(self, sync=False):
        """Raises if an earlier training forward produced more samples than ``sample_capacity`` (rays were dropped: the reference
        never drops samples).  Cheap: waits only on the event of the async copy issued by ``note_sample_overflow``.
        Returns the largest post-cull sample count of the most recent nerf_optim call."""
        if self._ov is None or self._ov_event is None:
            return 0
        self._ov_event.synchronize()
        m = int(self._ov_host[0])
        cap = getattr(self, '_ov_cap', 0)
        if cap and m > cap:
            raise RuntimeError('iNGPDecoder: %d samples survived the weight cull in one training forward but sample_capacity was %d -- '
                               'rays were dropped; raise decoder.sample_capacity' % (m, cap))
        return m

    def _field_params(self):
        return (self.encoder.params, self.mlp.net[0].weight, self.mlp.net[0].bias, self.mlp.net[1].weight, self.mlp.net[1].bias)

    def preproc(self, code):
        return code

    def loss(self):
        return None

    # ------------------------------------------------------------------ field
    def density_blob(self, x):
        d = (x ** 2).sum(-1).clamp(min=0.2)
        return self.blob_density * torch.exp(-d / (2 * self.blob_radius ** 2))

    def point_decode(self, xyzs, dirs, code, density_only=False, use_2nd_order=False, m_dev=None):
        """ingp_decoder.py:106-120.  xyzs: list with one [M,3] tensor (or a [1,M,3] tensor)."""
        assert len(xyzs) == 1, "Multiple scenes not implemented"
        assert not use_2nd_order
        sigmas, rgbs = _FieldFn.apply(xyzs[0], *self._field_params(), self, density_only, m_dev)
        return sigmas, (None if density_only else rgbs), [len(xyzs[0])]

    def point_density_decode(self, xyzs, code, **kwargs):
        sigmas, _, num_points = self.point_decode(xyzs, None, code, density_only=True, **kwargs)
        return sigmas, num_points

    # ------------------------------------------------------------------ occupancy grid
    def _morton_grid(self, grid_size, device):
        key = (grid_size, str(device))
        if key not in self._grid_cache:
            idx = torch.arange(grid_size ** 3, dtype=torch.int32, device=device)
            coords = rm.morton3D_invert(idx)                                   # cell of Morton index i
            centres = (coords.float() - (grid_size - 1) / 2) * (2 * self.bound / grid_size)
            # position of Morton index i in the reference's meshgrid enumeration (x-major), for externally supplied noise
            mesh_pos = (coords[:, 0].long() * grid_size + coords[:, 1].long()) * grid_size + coords[:, 2].long()
            self._grid_cache[key] = (centres, mesh_pos, torch.zeros(1, dtype=torch.float32, device=device))
        return self._grid_cache[key]

    def update_extra_state(self, code, density_grid, density_bitfield, iter_density, density_thresh=0.01, decay=0.9, S=128,
                           noise=None):
        """base_volume_renderer.py:105-177.  Full update (iter_density < 16: the only branch the MVEdit pipelines reach,
        SURVEY.md Appendix F): every cell is re-sampled at a jittered position, EMA'd and re-packed.
        ``noise`` (optional, [H^3,3] in [0,1), in the reference's meshgrid order) replaces the internal torch.rand draws."""
        with torch.no_grad():
            assert density_grid.dim() == 2 and density_grid.size(0) == 1, 'one scene'
            n_cells = density_grid.size(-1)
            grid_size = int(round(n_cells ** (1. / 3.)))
            device = density_grid.device
            if iter_density >= 16:
                raise NotImplementedError('partial occupancy update: not reachable from the MVEdit pipelines (iter_density stays 0)')
            centres, mesh_pos, scratch = self._morton_grid(grid_size, device)
            half_voxel_width = self.bound / grid_size
            if noise is None and self.test_noise is not None and self.test_noise.get('grid') is not None:
                noise = self.test_noise['grid']
            if noise is None:
                if self._jitter_gen is None or self._jitter_gen.device != centres.device:
                    self._jitter_gen = torch.Generator(device=centres.device)
                    self._jitter_gen.manual_seed(0x5eed)        # the same stream on every rank: replicas refresh identical grids
                u = torch.rand(centres.shape, generator=self._jitter_gen, device=centres.device)
            else:
                u = noise.to(device=device, dtype=torch.float32).reshape(-1, 3)[mesh_pos]
            xyzs = centres + (u * (2 * half_voxel_width) - half_voxel_width)
            sigmas, _ = self.point_density_decode([xyzs], code)
            assert density_grid.dtype == torch.float16 and density_grid.is_contiguous()
            call('mve_density_grid_update', ptr(density_grid), ptr(sigmas), ptr(None), c_u32(n_cells), c_f32(decay), ptr(scratch),
                 c_u32(n_cells), c_f32(density_thresh), ptr(density_bitfield), stream())
        return

    # ------------------------------------------------------------------ renderer
    def forward(self, rays_o, rays_d, code, density_bitfield, grid_size, dt_gamma=0.0, perturb=False, return_loss=False,
                compute_normal=False, update_extra_state=0, extra_args=None, extra_kwargs=None, noises=None, fused_entropy=None):
        """base_volume_renderer.py:179-343 (one scene).  rays_o/rays_d: (1, N, 3); density_bitfield: (1, H^3/8)."""
        assert not compute_normal, 'compute_normal is not used on the MVEdit path'
        for _ in range(update_extra_state):
            self.update_extra_state(code, *extra_args, **extra_kwargs)
        num_scenes = len(rays_o)
        assert num_scenes == 1, 'Multiple scenes not implemented (as ingp_decoder.py:110)'
        if isinstance(grid_size, (list, tuple)):
            grid_size = grid_size[0]
        dt_gamma_t = None
        if isinstance(dt_gamma, torch.Tensor):
            if self.training and self.sample_capacity:
                dt_gamma_t = dt_gamma.reshape(-1)[:1].float().contiguous()   # read on the device: no sync, graph-safe
                dt_gamma = 0.0
            else:
                dt_gamma = float(dt_gamma.reshape(-1)[0])          # only element 0 is used by the reference (:212-218)
        elif isinstance(dt_gamma, (list, tuple)):
            dt_gamma = float(dt_gamma[0])
        ro, rd = rays_o[0], rays_d[0]
        bitfield = density_bitfield[0]
        if self.training and self.test_noise is not None and self.test_noise.get('march') is not None:
            noises = self.test_noise['march']               # parity tests: supplied draws win over the caller's
        if self.training and self.sample_capacity:
            # B200-native protocol: fixed-capacity sample buffers + device-side counts -> no host sync anywhere in the iteration
            # (the reference syncs three times here: raymarching.py:290, base_volume_renderer.py:235-241), CUDA-graph capturable.
            cap = int(self.sample_capacity)                      # post-cull capacity (zero-filled tails: torch may reduce over them)
            nears, fars = rm.near_far_from_aabb(ro, rd, self.aabb, self.min_near)
            N = ro.shape[0]
            cap1 = N * int(self.max_steps) if self.weight_culling_th > 0 else cap   # the marcher can never overflow N*max_steps
            xyzs, dirs, ts, rays, counter = rm.march_rays_train(ro, rd, self.bound, bitfield, 1, grid_size, nears, fars, perturb=perturb,
                                                                dt_gamma=dt_gamma_t if dt_gamma_t is not None else dt_gamma,
                                                                max_steps=self.max_steps, noises=noises, max_points=cap1,
                                                                zero_tail=not self.weight_culling_th > 0, want_dirs=False)
            counter1 = counter
            if self.weight_culling_th > 0:
                with torch.no_grad():
                    sig0, _, _ = self.point_decode([xyzs], None, code, density_only=2, m_dev=counter)   # TF32 tensor-core MLP
                    w0 = torch.empty(cap1, dtype=torch.float32, device=xyzs.device)
                    scratch = torch.empty(N * 5, dtype=torch.float32, device=xyzs.device)
                    call('mve_composite_rays_train_forward', ptr(sig0), ptr(None), ptr(ts), ptr(rays), c_u32(cap1), ptr(counter), c_u32(N),
                         c_f32(1e-4), c_int(0), ptr(w0), ptr(scratch[:N]), ptr(scratch[N:2 * N]), ptr(scratch[2 * N:]), stream())
                    counter2 = torch.zeros(1, dtype=torch.int32, device=xyzs.device)
                    rays2 = torch.empty_like(rays)
                    xyzs2 = torch.empty(cap, 3, dtype=torch.float32, device=xyzs.device)
                    ts2 = torch.empty(cap, 2, dtype=torch.float32, device=xyzs.device)
                    call('mve_cull_samples', ptr(w0), c_f32(self.weight_culling_th), ptr(rays), ptr(xyzs), ptr(ts), c_u32(N), c_u32(cap1),
                         ptr(counter), c_u32(cap), ptr(rays2), ptr(xyzs2), ptr(ts2), ptr(counter2), stream())
                    xyzs, ts, rays, counter = xyzs2, ts2, rays2, counter2
                    self._track_overflow(counter2)
            sigmas, rgbs, num_points = self.point_decode([xyzs], None, code, m_dev=counter)
            weights, weights_sum, depth, image = rm.composite_rays_train(sigmas, rgbs, ts, rays, 1e-4, False, counter, fused_entropy)
            self.last_counts = (counter1, counter)       # device counters: marched / kept samples of the last training forward
            results = dict(weights=weights, weights_sum=weights_sum[None], depth=depth[None], image=image[None], rays=[rays], normal=None,
                           ts=[ts], num_samples=counter)
        elif self.training:
            nears, fars = rm.near_far_from_aabb(ro, rd, self.aabb, self.min_near)
            xyzs, dirs, ts, rays = rm.march_rays_train(ro, rd, self.bound, bitfield, 1, grid_size, nears, fars, perturb=perturb,
                                                       dt_gamma=dt_gamma, max_steps=self.max_steps, noises=noises)
            if self.weight_culling_th > 0:
                with torch.no_grad():
                    M, N = xyzs.shape[0], rays.shape[0]
                    sig0, _ = self.point_density_decode([xyzs], code)
                    w0 = torch.empty(M, dtype=torch.float32, device=xyzs.device)
                    scratch = torch.empty(N * 5, dtype=torch.float32, device=xyzs.device)
                    zeros_rgb = torch.zeros(M, 3, dtype=torch.float32, device=xyzs.device)
                    call('mve_composite_rays_train_forward', ptr(sig0), ptr(zeros_rgb), ptr(ts), ptr(rays), c_u32(M), ptr(None),
                         c_u32(N), c_f32(1e-4), c_int(0), ptr(w0), ptr(scratch[:N]), ptr(scratch[N:2 * N]), ptr(scratch[2 * N:]),
                         stream())
                    counter = torch.zeros(1, dtype=torch.int32, device=xyzs.device)
                    rays2 = torch.empty_like(rays)
                    xyzs2, ts2 = torch.empty_like(xyzs), torch.empty_like(ts)
                    call('mve_cull_samples', ptr(w0), c_f32(self.weight_culling_th), ptr(rays), ptr(xyzs), ptr(ts), c_u32(N),
                         c_u32(M), ptr(None), c_u32(M), ptr(rays2), ptr(xyzs2), ptr(ts2), ptr(counter), stream())
                    M2 = int(counter.item())
                    xyzs, ts, rays, dirs = xyzs2[:M2], ts2[:M2], rays2, None
            sigmas, rgbs, num_points = self.point_decode([xyzs], [dirs], code)
            if fused_entropy is not None:           # sample-entropy gradient folded into the composite backward (nerf_optim)
                weights, weights_sum, depth, image = rm.composite_rays_train(sigmas, rgbs, ts, rays, 1e-4, False, None, fused_entropy)
                weights_sum, depth, image = weights_sum[None], depth[None], image[None]
            else:
                weights, weights_sum, depth, image = rm.batch_composite_rays_train(sigmas, rgbs, [ts], [rays], num_points)
            results = dict(weights=weights, weights_sum=weights_sum, depth=depth, image=image, rays=[rays], normal=None, ts=[ts])
        else:
            N = ro.shape[0]
            ro_c, rd_c = ro.float().contiguous(), rd.float().contiguous()
            ws = torch.empty(N, dtype=torch.float32, device=ro.device)
            depth = torch.empty(N, dtype=torch.float32, device=ro.device)
            image = torch.empty(N, 3, dtype=torch.float32, device=ro.device)
            table, w1, b1, w2, b2 = self._field_params()
            call('mve_render_rays', ptr(ro_c), ptr(rd_c), ptr(None), ptr(None), ptr(None), c_u32(0), c_u32(0), c_u32(N), ptr(self.aabb),
                 c_f32(self.min_near), ptr(bitfield.contiguous()), c_f32(self.bound), c_f32(dt_gamma), c_u32(self.max_steps), c_u32(1),
                 c_u32(grid_size), c_f32(1e-2), ptr(table), ptr(w1), ptr(b1), ptr(w2), ptr(b2), *self.encoder._largs.args(),
                 c_f32(self.blob_density), c_f32(self.blob_radius), c_f32(self.sigmoid_saturation), ptr(ws), ptr(depth), ptr(image),
                 stream())
            results = dict(weights=None, weights_sum=[ws], depth=[depth], image=[image], rays=None, normal=[None], ts=None)
        if return_loss:
            results.update(decoder_reg_loss=self.loss())
        return results

    @staticmethod
    def last_render_stats():
        """(samples shaded, warp-rounds that shaded, warp-rounds, warp-level DDA search trips) of the most recent fused render
        launch (synchronises)."""
        import ctypes as _c
        buf = (_c.c_uint64 * 4)()
        call('mve_render_last_sample_count', buf)
        return int(buf[0]), int(buf[1]), int(buf[2]), int(buf[3])

    def render_cameras(self, poses, intrinsics, h, w, density_bitfield, grid_size, dt_gamma=0.0, dt_gamma_per_view=None):
        """Fused BaseNeRF.render core (base_nerf.py:489-556): rays are generated inside the kernel from (pose, intrinsics, pixel).
        poses [V,4,4] (or [V,3,4]) c2w, intrinsics [V,4] at the render size.  -> weights_sum [V,h,w], depth [V,h,w] (sum w/t),
        image [V,h,w,3] (premultiplied, no background)."""
        V = poses.shape[0]
        P = torch.zeros(V, 4, 4, dtype=torch.float32, device=poses.device)
        P[:, :poses.shape[1], :] = poses.float()
        K = intrinsics.float().contiguous()
        N = V * h * w
        ws = torch.empty(N, dtype=torch.float32, device=poses.device)
        depth = torch.empty(N, dtype=torch.float32, device=poses.device)
        image = torch.empty(N, 3, dtype=torch.float32, device=poses.device)
        table, w1, b1, w2, b2 = self._field_params()
        if dt_gamma_per_view is not None:
            dt_gamma_per_view = dt_gamma_per_view.float().contiguous()
            assert dt_gamma_per_view.numel() == V
        call('mve_render_rays', ptr(None), ptr(None), ptr(P), ptr(K), ptr(dt_gamma_per_view), c_u32(h), c_u32(w), c_u32(N), ptr(self.aabb),
             c_f32(self.min_near), ptr(density_bitfield.reshape(-1).contiguous()), c_f32(self.bound), c_f32(float(dt_gamma)),
             c_u32(self.max_steps), c_u32(1), c_u32(grid_size), c_f32(1e-2), ptr(table), ptr(w1), ptr(b1), ptr(w2), ptr(b2),
             *self.encoder._largs.args(), c_f32(self.blob_density), c_f32(self.blob_radius), c_f32(self.sigmoid_saturation),
             ptr(ws), ptr(depth), ptr(image), stream())
        return ws.view(V, h, w), depth.view(V, h, w), image.view(V, h, w, 3)
