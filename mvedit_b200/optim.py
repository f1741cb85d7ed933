"""``FusedAdam`` -- the optimizer of the reconstruction loops as ONE kernel per step.

The reference builds ``torch.optim.Adam(self.nerf.decoder.parameters(), lr=...)`` (and a multi-group Adam for the mesh stage) and
calls ``optimizer.zero_grad(); loss.backward(); optimizer.step()`` every iteration
(/root/reference/lib/pipelines/mvedit_3d_pipeline.py:1034,1313-1315,631-633).  Same constructor and the same arithmetic here
(torch.optim.Adam defaults: betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad), but:

  * all gradients live in ONE flat fp32 buffer (``p.grad`` are views into it) -- libmvedit_b200's backward kernels accumulate
    straight into it (``grad_sink``), and a data-parallel reconstruction all-reduces it with a single NCCL call;
  * ``step()`` is one launch of ``mve_adam_step`` over all tensors which also zeroes the gradients it consumed, so
    ``zero_grad()`` is free and the ~15 elementwise passes of the eager Adam over the 28.7 MB hash table become one;
  * ``lr`` and the step counter are device scalars: the update is CUDA-graph capturable as is.
"""
import ctypes

import torch

from ._lib import call, stream, c_int, c_u32, c_f32


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        defaults = dict(lr=lr, betas=betas, eps=eps, capturable=True)
        super().__init__(params, defaults)
        ps = [p for g in self.param_groups for p in g['params']]
        assert ps and all(p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() for p in ps), \
            'FusedAdam: contiguous fp32 CUDA parameters (no CPU fallback)'
        assert all(g['betas'] == self.param_groups[0]['betas'] and g['eps'] == self.param_groups[0]['eps'] for g in self.param_groups)
        dev = ps[0].device
        self._params = ps
        # 16-byte aligned segments in the flat buffers
        offs, o = [], 0
        for p in ps:
            offs.append(o)
            o += (p.numel() + 3) // 4 * 4
        self.flat_grad = torch.zeros(o, dtype=torch.float32, device=dev)
        self._m = torch.zeros(o, dtype=torch.float32, device=dev)
        self._v = torch.zeros(o, dtype=torch.float32, device=dev)
        self._step = torch.zeros(1, dtype=torch.int32, device=dev)
        self._views = []
        for p, off in zip(ps, offs):
            gv = self.flat_grad[off:off + p.numel()].view_as(p)
            p.grad = gv
            self._views.append((gv, self._m[off:off + p.numel()], self._v[off:off + p.numel()]))
            self.state[p] = dict(step=self._step, exp_avg=self._views[-1][1].view_as(p), exp_avg_sq=self._views[-1][2].view_as(p))
        for g in self.param_groups:
            g['lr'] = torch.tensor(float(g['lr']), dtype=torch.float32, device=dev)
        self._clean = True          # gradients are known to be zero (fresh, or zeroed by the last step)
        self._build_args()

    def _build_args(self):
        n = len(self._params)
        assert n <= 16, 'FusedAdam: at most 16 parameter tensors (mve_adam_step launches them together)'
        arr = ctypes.c_void_p * n
        lrs = [g['lr'] for g in self.param_groups for _ in g['params']]
        self._c = dict(p=arr(*[p.data_ptr() for p in self._params]), g=arr(*[v[0].data_ptr() for v in self._views]),
                       m=arr(*[v[1].data_ptr() for v in self._views]), v=arr(*[v[2].data_ptr() for v in self._views]),
                       n=(ctypes.c_uint32 * n)(*[p.numel() for p in self._params]), lr=arr(*[t.data_ptr() for t in lrs]))
        self._ptrs = tuple(p.data_ptr() for p in self._params)

    # ------------------------------------------------------------------ gradient sink for the backward kernels
    def grad_sink(self, p):
        """The persistent gradient view of ``p`` (kernels accumulate into it; it is zero after every ``step``)."""
        for q, v in zip(self._params, self._views):
            if q is p:
                self._clean = False
                return v[0]
        raise KeyError('parameter is not managed by this optimizer')

    def zero_grad(self, set_to_none=False):
        """``step`` zeroes the gradients it consumed, so inside a captured iteration this is free (nothing is recorded: the buffer is
        zero by construction).  Outside a capture the buffer is cleared explicitly (a 28.7 MB memset, ~5 us) -- gradients written by
        autograd between steps are not tracked."""
        for p, v in zip(self._params, self._views):
            if p.grad is None or p.grad.data_ptr() != v[0].data_ptr():
                p.grad = v[0]
        if not torch.cuda.is_current_stream_capturing():
            self.flat_grad.zero_()
        self._clean = True

    def hard_zero_grad(self):
        self.flat_grad.zero_()
        self._clean = True

    @torch.no_grad()
    def fold_grads(self):
        """Make ``flat_grad`` hold every gradient: a ``p.grad`` that autograd attached as a fresh tensor is added into its slot of the flat
        buffer (``step`` does this itself; a data-parallel caller does it before all-reducing ``flat_grad``)."""
        for p, v in zip(self._params, self._views):
            if p.grad is not None and p.grad.data_ptr() != v[0].data_ptr():
                v[0].add_(p.grad)
                p.grad = v[0]

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        assert closure is None
        if tuple(p.data_ptr() for p in self._params) != self._ptrs:
            self._build_args()       # a parameter was re-allocated (load_state_dict copies in place, so this is rare)
        self.fold_grads()
        g0 = self.param_groups[0]
        c = self._c
        call('mve_adam_step', c_u32(len(self._params)), c['p'], c['g'], c['m'], c['v'], c['n'], c['lr'], c_f32(g0['betas'][0]),
             c_f32(g0['betas'][1]), c_f32(g0['eps']), c_f32(grad_scale), ctypes.c_void_p(self._step.data_ptr()), c_int(1), stream())
        self._clean = True

    def set_lr(self, lr, group=None):
        for i, g in enumerate(self.param_groups):
            if group is None or group == i:
                g['lr'].fill_(float(lr))

    # ------------------------------------------------------------------ snapshots (bench: every timed step restarts from the same state)
    def snapshot(self):
        return dict(m=self._m.clone(), v=self._v.clone(), step=self._step.clone(), lr=[g['lr'].clone() for g in self.param_groups])

    def restore(self, snap):
        self._m.copy_(snap['m']); self._v.copy_(snap['v']); self._step.copy_(snap['step'])
        for g, l in zip(self.param_groups, snap['lr']):
            g['lr'].copy_(l)
        self.flat_grad.zero_()
        self._clean = True
