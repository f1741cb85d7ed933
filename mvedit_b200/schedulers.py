"""The solver side of the loop body: the schedulers ``Adapter3DRunner.load_scheduler`` can hand to ``MVEdit3DPipeline``
(/root/reference/lib/apis/adapter3d.py:279-303: ``<name>Scheduler.from_config(SD1.5 scheduler_config, use_karras_sigmas=False,
timestep_spacing='trailing')`` or, for the ``...Karras`` names, ``use_karras_sigmas=True, timestep_spacing='leading'``), restated from
the published update rules of diffusers==0.27.2 (requirements.txt:13; not installed here -> parity unpinned):

  * ``EulerAncestralScheduler``      -- the UI default (lib/core/webui/shared_opts.py:39)
  * ``DPMSolverMultistepScheduler``  -- image-to-3D (lib/core/webui/tab_img_to_3d.py:54): DPM-Solver++(2M), epsilon prediction,
                                        midpoint, ``lower_order_final``, ``final_sigmas_type='zero'``

  * ``DPMSolverSDEScheduler``        -- DPM-Solver++ SDE (k-diffusion's ``sample_dpmpp_sde``); the reference keeps one scheduler object per
                                        view because each owns a torchsde Brownian tree (mvedit_3d_pipeline.py:1176-1177,1456-1459);
                                        here the Brownian increments are built from the explicit noise tensors, one object for all views
  * ``DDIMScheduler``

Interface: what the loop body calls on ``self.scheduler`` (mvedit_3d_pipeline.py:1101-1109,1209-1213,1225,1461,1466,1478):
``betas``, ``order``, ``set_timesteps``, ``timesteps``, ``init_noise_sigma``, ``scale_model_input(sample, t)``,
``step(model_output, t, sample, ...)``, ``add_noise(x, noise, timesteps)``, ``model_outputs``.  ``t`` is an element of ``timesteps``
(a 0-dim tensor / float) or -- an extension used by bench.py -- a python ``int`` schedule index.  Two extensions, both because the
oracle and the kernels must see the same random draws: ``step`` takes the ancestral ``noise`` as a tensor, and ``prune(keep_ids)``
applies camera pruning to the multistep history (the reference indexes ``scheduler.model_outputs`` by hand, :1209-1211).
Everything here is O(latents) elementwise work on the device, a few launches per step; the schedule itself lives in numpy fp64.
"""
import numpy as np
import torch


class _SigmaSchedule:
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, timestep_spacing='trailing',
                 use_karras_sigmas=False, steps_offset=1):
        # SD1.5 scheduler_config.json: scaled_linear betas, steps_offset 1, epsilon prediction
        self.num_train_timesteps = num_train_timesteps
        self.timestep_spacing, self.use_karras_sigmas, self.steps_offset = timestep_spacing, use_karras_sigmas, steps_offset
        self.betas_np = np.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=np.float64) ** 2
        self.betas = torch.from_numpy(self.betas_np).float()
        self.alphas = 1.0 - self.betas_np
        self.alphas_cumprod = np.cumprod(self.alphas)
        self._train_sigmas = np.sqrt((1 - self.alphas_cumprod) / self.alphas_cumprod)
        self.init_noise_sigma = None
        self.timesteps = self.sigmas = None

    # ------------------------------------------------------------------ schedule
    def _spaced_timesteps(self, n):
        T = self.num_train_timesteps
        if self.timestep_spacing == 'trailing':
            return np.round(np.arange(T, 0, -T / n)) - 1
        if self.timestep_spacing == 'leading':
            return (np.arange(0, n) * (T // n)).round()[::-1].astype(np.float64) + self.steps_offset
        if self.timestep_spacing == 'linspace':
            return np.linspace(0, T - 1, n)[::-1].copy()
        raise ValueError(self.timestep_spacing)

    def _karras(self, n, rho=7.0):
        """Karras et al. 2022 eq. 5 between the training schedule's extreme sigmas."""
        smin, smax = self._train_sigmas[0], self._train_sigmas[-1]
        ramp = np.linspace(0, 1, n)
        return (smax ** (1 / rho) + ramp * (smin ** (1 / rho) - smax ** (1 / rho))) ** rho

    def _sigma_to_t(self, sigma):
        """Fractional training timestep whose (log-)sigma interpolates to ``sigma``."""
        log_sigmas = np.log(self._train_sigmas)
        ls = np.log(np.maximum(sigma, 1e-10))
        low = np.clip(np.cumsum(ls[None, :] - log_sigmas[:, None] >= 0, axis=0).argmax(axis=0), None, len(log_sigmas) - 2)
        lo, hi = log_sigmas[low], log_sigmas[low + 1]
        w = np.clip((lo - ls) / (lo - hi), 0, 1)
        return (1 - w) * low + w * (low + 1)

    def _schedule(self, n, round_karras_t):
        if self.use_karras_sigmas:
            sig = self._karras(n)
            ts = self._sigma_to_t(sig)
            return (ts.round() if round_karras_t else ts), sig
        ts = self._spaced_timesteps(n)
        return ts, np.interp(ts, np.arange(self.num_train_timesteps), self._train_sigmas)

    _cursor = 0          # index after the last ``step``: rounded Karras timesteps can repeat, the walk through them must not

    def index_of(self, t):
        """python int -> schedule index as is; anything else is a timestep value looked up in ``timesteps`` (the first match at or
        after the last step taken, else the first match)."""
        if isinstance(t, int):
            return t
        hit = (self.timesteps == float(t)).nonzero().reshape(-1).tolist()
        if not hit:
            raise ValueError(f'timestep {float(t)} is not in the schedule')
        later = [h for h in hit if h >= self._cursor]
        return later[0] if later else hit[0]

    def noise_scales(self, t):
        """(sqrt(alpha_bar_t), sqrt(1 - alpha_bar_t)) as 0-dim tensors on t's device; a fractional t (the 'trailing' / Karras
        timesteps are floats) interpolates the VE sigma between the two neighbouring integer timesteps -- what the reference's
        get_noise_scales does (lib/core/diffusion.py:4-21; checked against it in tests/test_reference_pins.py)."""
        tf = float(t)
        sig = self._train_sigmas
        lo = min(int(tf), self.num_train_timesteps - 1)
        hi = min(lo + 1, self.num_train_timesteps - 1)
        ve = sig[lo] + (sig[hi] - sig[lo]) * (tf - lo) if torch.is_floating_point(t) else sig[lo]
        a = 1.0 / np.sqrt(1.0 + ve * ve)
        return t.new_tensor(a, dtype=torch.float32), t.new_tensor(ve * a, dtype=torch.float32)

    def prune(self, keep_ids):
        """Camera pruning (lib/pipelines/utils.py:350-379) drops views: per-view solver state follows."""


class EulerAncestralScheduler(_SigmaSchedule):
    """diffusers EulerAncestralDiscreteScheduler as the reference configures it for SD1.5 (scaled_linear betas 0.00085 -> 0.012,
    1000 train steps, epsilon prediction, timestep_spacing='trailing': lib/apis/adapter3d.py:280-300).
    The ancestral noise is an explicit argument of ``step`` so that oracle and kernels see identical draws."""

    def set_timesteps(self, num_inference_steps, device='cpu'):
        ts, sigmas = self._schedule(num_inference_steps, round_karras_t=False)
        self.sigmas = torch.tensor(np.concatenate([sigmas, [0.0]]), dtype=torch.float32, device=device)
        self.timesteps = torch.tensor(ts, dtype=torch.float32, device=device)
        sm = float(self.sigmas.max())
        # 'trailing' / 'linspace' start from pure noise of std sigma_max; 'leading' from the VP-consistent sqrt(sigma_max^2 + 1)
        self.init_noise_sigma = sm if self.timestep_spacing in ('linspace', 'trailing') else (sm * sm + 1) ** 0.5
        self._cursor = 0

    def scale_model_input(self, sample, t):
        return sample / ((self.sigmas[self.index_of(t)] ** 2 + 1) ** 0.5)

    def add_noise(self, original_samples, noise, timesteps):
        """x + noise * sigma(t), t looked up in the current schedule."""
        idx = [self.index_of(t) for t in timesteps.reshape(-1)]
        sigma = self.sigmas[idx].to(original_samples.device)
        return original_samples + noise * sigma.view(-1, *([1] * (original_samples.dim() - 1)))

    def step(self, model_output, t, sample, noise):
        i = self.index_of(t)
        self._cursor = i + 1
        sigma = self.sigmas[i]
        pred_original_sample = sample - sigma * model_output
        sigma_to = self.sigmas[i + 1]
        sigma_up = (sigma_to ** 2 * (sigma ** 2 - sigma_to ** 2) / sigma ** 2) ** 0.5
        sigma_down = (sigma_to ** 2 - sigma_up ** 2) ** 0.5
        derivative = (sample - pred_original_sample) / sigma
        return sample + derivative * (sigma_down - sigma) + noise * sigma_up


class DDIMScheduler(_SigmaSchedule):
    """DDIM (Song et al. 2021, eq. 12) with eta = 0, as diffusers' DDIMScheduler on the SD1.5 config (clip_sample False,
    set_alpha_to_one False: the step past the last timestep lands on alphas_cumprod[0])."""

    def set_timesteps(self, num_inference_steps, device='cpu'):
        assert not self.use_karras_sigmas, 'DDIM has no Karras variant (shared_opts.py:40-42)'
        ts = self._spaced_timesteps(num_inference_steps)
        self._stride = self.num_train_timesteps // num_inference_steps
        self.timesteps = torch.tensor(ts, dtype=torch.int64, device=device)
        self.sigmas = None
        self.init_noise_sigma = 1.0
        self._cursor = 0

    def scale_model_input(self, sample, t):
        return sample

    def add_noise(self, original_samples, noise, timesteps):
        ac = original_samples.new_tensor(self.alphas_cumprod[[int(t) for t in timesteps.reshape(-1)]])
        shape = (-1, *([1] * (original_samples.dim() - 1)))
        return ac.sqrt().view(shape) * original_samples + (1 - ac).sqrt().view(shape) * noise

    def step(self, model_output, t, sample, noise=None):
        ti = int(self.timesteps[t]) if isinstance(t, int) else int(t)
        prev = ti - self._stride
        a_t = self.alphas_cumprod[ti]
        a_p = self.alphas_cumprod[prev] if prev >= 0 else self.alphas_cumprod[0]
        eps = model_output.float()
        x0 = (sample.float() - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
        return (a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * eps).to(sample.dtype)


class DPMSolverMultistepScheduler(_SigmaSchedule):
    """DPM-Solver++(2M) (Lu et al. 2022, arXiv:2211.01095 Alg. 2) in the sigma parameterisation of diffusers 0.27.2's
    DPMSolverMultistepScheduler with the defaults the reference leaves in place (``algorithm_type`` is deleted from the SD1.5 config,
    adapter3d.py:295-297 -> 'dpmsolver++'; solver_order 2, 'midpoint', lower_order_final, final_sigmas_type 'zero', no thresholding).

    The model is an epsilon predictor; the solver works on x0 = (x - sigma_t eps) / alpha_t with alpha_t = 1 / sqrt(sigma^2 + 1),
    sigma_t = sigma alpha_t, lambda = log(alpha_t / sigma_t):
        first order   x <- (sigma_t / sigma_s) x - alpha_t (e^{-h} - 1) D0
        second order  x <- (sigma_t / sigma_s) x - alpha_t (e^{-h} - 1) (D0 + D1 / 2),   D1 = (x0_s - x0_{s-1}) / r,  r = h_prev / h
    The first step, and the last one when the schedule has fewer than 15 steps or ends at sigma 0, are first order."""
    order = 1           # diffusers' ``order`` counts model evaluations per step (mvedit_3d_pipeline.py:1109 divides by it)

    def __init__(self, *args, solver_order=2, lower_order_final=True, **kwargs):
        super().__init__(*args, **kwargs)
        assert solver_order in (1, 2)
        self.solver_order, self.lower_order_final = solver_order, lower_order_final
        self.model_outputs = [None] * solver_order
        self.lower_order_nums = 0

    def set_timesteps(self, num_inference_steps, device='cpu'):
        ts, sigmas = self._schedule(num_inference_steps, round_karras_t=True)
        self._sig = np.concatenate([sigmas, [0.0]])                      # final_sigmas_type 'zero'
        self.sigmas = torch.tensor(self._sig, dtype=torch.float32, device=device)
        self.timesteps = torch.tensor(ts, dtype=torch.int64, device=device)
        self.init_noise_sigma = 1.0
        self.model_outputs = [None] * self.solver_order
        self.lower_order_nums = 0
        self._cursor = 0

    def scale_model_input(self, sample, t):
        return sample

    @staticmethod
    def _alpha_sigma(sigma):
        a = 1.0 / np.sqrt(sigma * sigma + 1.0)
        return a, sigma * a

    def add_noise(self, original_samples, noise, timesteps):
        idx = [self.index_of(t) for t in timesteps.reshape(-1)]
        a, s = self._alpha_sigma(self._sig[idx])
        shape = (-1, *([1] * (original_samples.dim() - 1)))
        return (original_samples * original_samples.new_tensor(a).view(shape) + noise * original_samples.new_tensor(s).view(shape))

    def prune(self, keep_ids):
        self.model_outputs = [m[keep_ids] if m is not None else None for m in self.model_outputs]

    def step(self, model_output, t, sample, noise=None):
        i = self.index_of(t)
        self._cursor = i + 1
        n = len(self.timesteps)
        final = i == n - 1                       # final sigma is 0: always first order
        second_last = i == n - 2 and self.lower_order_final and n < 15
        a_s, s_s = self._alpha_sigma(self._sig[i])
        x0 = (sample.float() - s_s * model_output.float()) / a_s
        self.model_outputs = self.model_outputs[1:] + [x0]
        a_t, s_t = self._alpha_sigma(self._sig[i + 1])
        with np.errstate(divide='ignore'):
            lam_t, lam_s = np.log(a_t) - np.log(s_t), np.log(a_s) - np.log(s_s)
        h = lam_t - lam_s
        c = -a_t * np.expm1(-h)                  # exp(-inf) = 0 at the final step: x <- x0
        if self.solver_order == 1 or self.lower_order_nums < 1 or final:
            out = (s_t / s_s) * sample.float() + c * x0
        else:
            del second_last                      # order 2 is the maximum: the lower_order_second switch only matters for order 3
            a_p, s_p = self._alpha_sigma(self._sig[i - 1])
            h0 = lam_s - (np.log(a_p) - np.log(s_p))
            d1 = (x0 - self.model_outputs[-2]) * (h / h0)
            out = (s_t / s_s) * sample.float() + c * x0 + (0.5 * c) * d1
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        return out.to(sample.dtype)


class DPMSolverSDEScheduler(_SigmaSchedule):
    """DPM-Solver++ SDE, the stochastic second-order sampler of Karras et al. / k-diffusion's ``sample_dpmpp_sde`` as diffusers 0.27.2
    packages it (``DPMSolverSDEScheduler``: epsilon prediction, midpoint ratio 1/2, eta = s_noise = 1).  The reference keeps ONE
    scheduler object per view because each owns a seeded torchsde Brownian tree (mvedit_3d_pipeline.py:1176-1177,1456-1459); here the
    Brownian increments are built from the explicit per-call ``noise`` tensor, so one object steps all views and ``prune`` drops the
    state of pruned views.

    Every sampler step is two model evaluations, so ``timesteps`` has 2n - 1 entries: t_0, m_0, t_1, m_1, ..., t_{n-1} with m_k the
    training timestep of the geometric-mean sigma sqrt(sigma_k sigma_{k+1}) (the midpoint in t = -log sigma); ``order`` = 2 as in
    diffusers (the pipeline's denoising-strength cut keeps that alignment, mvedit_3d_pipeline.py:1109).  With x0 = x - sigma eps:
        stage 1 (at t_k):  ancestral split of sigma_k -> sigma_mid into (down, up);  x_mid = (down / sigma_k) x + (1 - down / sigma_k) x0 + up * n1
        stage 2 (at m_k):  x0' from the model at x_mid;  split of sigma_k -> sigma_{k+1};  x_next = (down / sigma_k) x + (1 - down / sigma_k) x0' + up * n
    where n is the normalised Brownian increment over [sigma_{k+1}, sigma_k], which CONTAINS stage 1's interval:
    n = (sqrt(a) n1 + sqrt(b) n2) / sqrt(a + b), a = sigma_k - sigma_mid, b = sigma_mid - sigma_{k+1} (the tree's time is sigma itself).
    The last step (sigma_next = 0) is one Euler step to x0."""
    order = 2

    def set_timesteps(self, num_inference_steps, device='cpu'):
        ts, sigmas = self._schedule(num_inference_steps, round_karras_t=False)
        self._sig = np.concatenate([sigmas, [0.0]])
        mid = np.sqrt(self._sig[:-2] * self._sig[1:-1])                               # n - 1 midpoints (none for the final Euler step)
        self._mid = mid
        mts = self._sigma_to_t(mid)
        full = np.empty(2 * len(ts) - 1)
        full[0::2], full[1::2] = ts, mts
        per_index = np.empty(2 * len(ts) - 1)
        per_index[0::2], per_index[1::2] = sigmas, mid
        self.timesteps = torch.tensor(full, dtype=torch.float32, device=device)
        self.sigmas = torch.tensor(np.concatenate([per_index, [0.0]]), dtype=torch.float32, device=device)   # sigma the model sees at index i
        sm = float(sigmas.max())
        self.init_noise_sigma = sm if self.timestep_spacing in ('linspace', 'trailing') else (sm * sm + 1) ** 0.5
        self._sample = self._n1 = None
        self._cursor = 0

    def scale_model_input(self, sample, t):
        return sample / ((self.sigmas[self.index_of(t)] ** 2 + 1) ** 0.5)

    def add_noise(self, original_samples, noise, timesteps):
        idx = [self.index_of(t) for t in timesteps.reshape(-1)]
        sigma = self.sigmas[idx].to(original_samples.device)
        return original_samples + noise * sigma.view(-1, *([1] * (original_samples.dim() - 1)))

    def prune(self, keep_ids):
        if self._sample is not None:
            self._sample, self._n1 = self._sample[keep_ids], self._n1[keep_ids]

    @staticmethod
    def _split(sigma_from, sigma_to):
        """Ancestral split (eta = 1): the step lands on sigma_down and sigma_up of fresh noise restores the marginal at sigma_to."""
        up = min(sigma_to, (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5)
        return (sigma_to ** 2 - up ** 2) ** 0.5, up

    def step(self, model_output, t, sample, noise):
        i = self.index_of(t)
        self._cursor = i + 1
        k, second = i // 2, i % 2 == 1
        s_k, s_next = float(self._sig[k]), float(self._sig[k + 1])
        if s_next == 0.0:                                      # final step: Euler to x0
            return (sample.float() - s_k * model_output.float()).to(sample.dtype)
        s_mid = float(self._mid[k])
        if not second:
            x0 = sample.float() - s_k * model_output.float()
            down, up = self._split(s_k, s_mid)
            self._sample, self._n1 = sample.float(), noise.float()
            out = (down / s_k) * self._sample + (1 - down / s_k) * x0 + up * self._n1
        else:
            if self._sample is None:
                raise RuntimeError('DPMSolverSDEScheduler: second-stage step without its first stage (timesteps must be walked in order)')
            x0 = sample.float() - s_mid * model_output.float()         # the model saw x_mid at sigma_mid
            down, up = self._split(s_k, s_next)
            a, b = s_k - s_mid, s_mid - s_next
            n = (a ** 0.5 * self._n1 + b ** 0.5 * noise.float()) / (a + b) ** 0.5
            out = (down / s_k) * self._sample + (1 - down / s_k) * x0 + up * n
            self._sample = self._n1 = None
        return out.to(sample.dtype)
