"""mvedit_b200.schedulers (CPU): the solver restatements against properties that hold for the published update rules.

diffusers is not installed here, so these are not reference pins (parity unpinned for the schedulers); what is checked:
  * an exact epsilon model of a point-mass data distribution (eps = (x - alpha x0) / sigma) must be solved EXACTLY by DDIM and by
    DPM-Solver++ of either order (their updates are exact when x0-prediction is constant), and in expectation-free form by
    Euler-ancestral with zero noise;
  * DPM-Solver++ first order == DDIM on the same (non-Karras) schedule;
  * the second-order solver beats the first-order one on a model whose x0-prediction varies (two-point mixture posterior mean);
  * Karras sigmas are monotone, span the training range, and ``sigma_to_t`` inverts the training sigma table;
  * ``noise_scales`` / ``add_noise`` are consistent with the VP marginal; pruning keeps the multistep history aligned."""
import numpy as np
import pytest
import torch

from mvedit_b200.schedulers import EulerAncestralScheduler, DDIMScheduler, DPMSolverMultistepScheduler


def vp(sch, i):
    """(alpha, sigma) of the VP marginal at schedule index i."""
    if sch.sigmas is not None:
        s = float(sch.sigmas[i])
        a = 1 / np.sqrt(1 + s * s)
        return a, s * a
    ac = sch.alphas_cumprod[int(sch.timesteps[i])]
    return np.sqrt(ac), np.sqrt(1 - ac)


def run(sch, n, eps_model, x_T_unit, noise=False, skip_last=0):
    sch.set_timesteps(n)
    x = x_T_unit * sch.init_noise_sigma
    g = torch.Generator().manual_seed(1)
    for i, t in enumerate(sch.timesteps[:n - skip_last]):
        a, s = vp(sch, i)
        xin = sch.scale_model_input(x, t)
        eps = eps_model(xin.double(), a, s).float()
        nz = torch.randn(x.shape, generator=g) if noise else torch.zeros_like(x)
        x = sch.step(eps, t, x, nz)
    return x


@pytest.mark.parametrize('cls,kw', [(DDIMScheduler, {}), (DPMSolverMultistepScheduler, {}), (DPMSolverMultistepScheduler, dict(solver_order=1)),
                                    (DPMSolverMultistepScheduler, dict(use_karras_sigmas=True, timestep_spacing='leading')),
                                    (EulerAncestralScheduler, {})])
def test_point_mass_is_solved_exactly(cls, kw):
    x0 = torch.tensor([[0.7, -1.3, 0.2, 2.0]])
    xT = torch.tensor([[0.3, -0.5, 1.1, -2.0]])
    out = run(cls(**kw), 12, lambda x, a, s: (x - a * x0.double()) / s, xT)
    if cls is DDIMScheduler:       # set_alpha_to_one False: DDIM stops at alphas_cumprod[0] (sigma 0.029), not at sigma 0, on the exact path
        sch = cls(**kw)
        a0, s0 = np.sqrt(sch.alphas_cumprod[0]), np.sqrt(1 - sch.alphas_cumprod[0])
        sch.set_timesteps(12)
        aT, sT = vp(sch, 0)
        x0 = a0 * x0 + s0 * (xT - aT * x0) / sT
    # DDIM's integer stride (prev = t - T // n) does not land on the next 'trailing' timestep: exact only to ~1e-3
    assert float((out - x0).abs().max()) < (2e-3 if cls is DDIMScheduler else 1e-4), out


def test_first_order_dpm_equals_ddim_update():
    """x <- (sigma_t / sigma_s) x - alpha_t (e^{-h} - 1) x0  is  sqrt(a_prev) x0 + sqrt(1 - a_prev) eps  re-arranged."""
    d = DPMSolverMultistepScheduler(solver_order=1)
    d.set_timesteps(10)
    g = torch.Generator().manual_seed(0)
    x, eps = torch.randn(2, 4, 8, 8, generator=g), torch.randn(2, 4, 8, 8, generator=g)
    for i in (0, 3, 8):
        out = d.step(eps, i, x)
        a_s, s_s = vp(d, i)
        a_t, s_t = vp(d, i + 1)
        x0 = (x - s_s * eps) / a_s
        np.testing.assert_allclose(out.numpy(), (a_t * x0 + s_t * eps).numpy(), rtol=1e-5, atol=1e-5)


def test_second_order_is_more_accurate():
    """Data ~ N(0, c^2): the x0-prediction a c^2 x / (a^2 c^2 + s^2) varies along the trajectory and the probability-flow ODE has
    the closed-form solution x_t = x_T sqrt(a_t^2 c^2 + s_t^2) / sqrt(a_T^2 c^2 + s_T^2) (marginals stay Gaussian).  Compared
    before the last step (which jumps to sigma 0, h = inf: a denoise, not an ODE step): the first-order solver converges at O(h),
    DPM-Solver++(2M) at O(h^2)."""
    c = 0.5

    def eps_model(x, a, s):
        x0 = a * c * c * x / (a * a * c * c + s * s)
        return (x - a * x0) / s

    xT = torch.linspace(-2, 2, 41)[None]
    errs = {}
    for order in (1, 2):
        for n in (10, 20, 40):
            sch = DPMSolverMultistepScheduler(solver_order=order)
            out = run(sch, n, eps_model, xT, skip_last=1)
            (a, s), (al, sl) = vp(sch, 0), vp(sch, n - 1)
            exact = xT * np.sqrt(al * al * c * c + sl * sl) / np.sqrt(a * a * c * c + s * s)
            errs[order, n] = float((out - exact).abs().max())
    # measured: order 1 -> 0.078 / 0.054 / 0.034, order 2 -> 0.0081 / 0.0037 / (fp32 floor ~0.004)
    assert errs[2, 10] < 0.15 * errs[1, 10] and errs[2, 20] < 0.15 * errs[1, 20], errs
    assert errs[1, 40] < 0.75 * errs[1, 20] < 0.75 * 0.75 * errs[1, 10], errs


def test_karras_schedule_and_sigma_to_t():
    s = DPMSolverMultistepScheduler(use_karras_sigmas=True, timestep_spacing='leading')
    s.set_timesteps(24)
    sig = s.sigmas.numpy()
    assert sig[-1] == 0 and np.all(np.diff(sig) < 0)
    np.testing.assert_allclose(sig[0], s._train_sigmas[-1], rtol=1e-6)
    np.testing.assert_allclose(sig[-2], s._train_sigmas[0], rtol=1e-6)
    ts = s.timesteps.numpy()
    assert ts[0] == 999 and ts[-1] == 0 and np.all(np.diff(ts) <= 0)
    np.testing.assert_allclose(s._sigma_to_t(s._train_sigmas[[0, 17, 500, 998]]), [0, 17, 500, 998], atol=1e-6)
    # Euler with Karras sigmas keeps fractional timesteps
    e = EulerAncestralScheduler(use_karras_sigmas=True, timestep_spacing='leading')
    e.set_timesteps(24)
    assert e.timesteps.dtype == torch.float32 and abs(e.init_noise_sigma - (float(e.sigmas[0]) ** 2 + 1) ** 0.5) < 1e-5


def test_trailing_spacing_and_truncation_index():
    s = EulerAncestralScheduler()
    s.set_timesteps(24)
    assert float(s.timesteps[0]) == 999 and len(s.timesteps) == 24 and float(s.timesteps[-1]) == 41
    d = DPMSolverMultistepScheduler(timestep_spacing='leading')
    d.set_timesteps(20)
    assert d.timesteps.tolist() == [int(v) for v in (np.arange(20) * 50)[::-1] + 1]
    assert s.index_of(s.timesteps[5]) == 5 and s.index_of(5) == 5 and d.index_of(d.timesteps[7]) == 7
    with pytest.raises(ValueError):
        d.index_of(torch.tensor(2))


@pytest.mark.parametrize('cls', [EulerAncestralScheduler, DDIMScheduler, DPMSolverMultistepScheduler])
def test_add_noise_matches_noise_scales(cls):
    s = cls()
    s.set_timesteps(16)
    x, n = torch.full((2, 3), 2.0), torch.full((2, 3), -1.0)
    t = s.timesteps[4:5]
    out = s.scale_model_input(s.add_noise(x, n, t), t[0])
    a, sg = s.noise_scales(t[0])
    np.testing.assert_allclose(out.numpy(), (a * x + sg * n).numpy(), rtol=2e-5)


def test_prune_keeps_history_aligned():
    d = DPMSolverMultistepScheduler()
    d.set_timesteps(8)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(6, 4, generator=g)
    eps = [torch.randn(6, 4, generator=g) for _ in range(3)]
    keep = torch.tensor([0, 2, 5])
    full = x
    for i in range(3):
        full = d.step(eps[i], d.timesteps[i], full)
    d.set_timesteps(8)
    part = x
    for i in range(3):
        if i == 2:
            d.prune(keep); part = part[keep]
        part = d.step(eps[i] if i < 2 else eps[i][keep], d.timesteps[i], part)
    np.testing.assert_allclose(part.numpy(), full[keep].numpy(), rtol=1e-6, atol=1e-6)


def _sde(n=16, **kw):
    from mvedit_b200.schedulers import DPMSolverSDEScheduler
    s = DPMSolverSDEScheduler(**kw)
    s.set_timesteps(n)
    return s


def test_dpm_sde_schedule_layout():
    s = _sde(12)
    assert len(s.timesteps) == 23 and s.order == 2 and len(s.sigmas) == 24 and float(s.sigmas[-1]) == 0
    sig = s.sigmas[:-1].numpy()
    assert (np.diff(sig) < 0).all()                                         # t_0 > m_0 > t_1 > ... in sigma
    np.testing.assert_allclose(sig[1::2], np.sqrt(sig[0:-1:2] * sig[2::2]), rtol=1e-5)      # midpoints in log sigma
    assert (np.diff(s.timesteps.numpy()) < 0).all()
    assert abs(s.init_noise_sigma - float(s.sigmas[0])) < 1e-6


def _dpmpp_sde_direct(sig, denoiser, x, noises):
    """k-diffusion's ``sample_dpmpp_sde`` loop (r = 1/2, eta = s_noise = 1) written out directly; the Brownian path lives on the sigma
    axis: noises[2k] drives [sigma_mid, sigma_k], noises[2k+1] drives [sigma_{k+1}, sigma_mid]."""
    split = lambda s_from, s_to: ((s_to ** 2 - min(s_to, (s_to ** 2 * (s_from ** 2 - s_to ** 2) / s_from ** 2) ** 0.5) ** 2) ** 0.5,
                                  min(s_to, (s_to ** 2 * (s_from ** 2 - s_to ** 2) / s_from ** 2) ** 0.5))
    for k in range(len(sig) - 1):
        s, sn = sig[k], sig[k + 1]
        den = denoiser(x, s)
        if sn == 0:
            x = den
            continue
        sm = (s * sn) ** 0.5
        n1, n2 = noises[2 * k], noises[2 * k + 1]
        down, up = split(s, sm)
        x2 = (down / s) * x + (1 - down / s) * den + up * n1
        den2 = denoiser(x2, sm)
        down, up = split(s, sn)
        w = ((s - sm) ** 0.5 * n1 + (sm - sn) ** 0.5 * n2) / (s - sn) ** 0.5            # W(sigma_k) - W(sigma_next), normalised
        x = (down / s) * x + (1 - down / s) * den2 + up * w
    return x


def test_dpm_sde_equals_the_direct_loop_and_converges_on_a_gaussian():
    """Data ~ N(0, c^2): exact denoiser x c^2 / (c^2 + sigma^2).  (i) The two-calls-per-step scheduler reproduces the directly written
    k-diffusion loop given the same unit normals (which checks that stage 2's Brownian increment contains stage 1's); (ii) the final
    variance approaches c^2 as the step count grows; (iii) breaking the correlation between the stages makes it worse."""
    c = 0.6
    den = lambda x, s: x * (c * c / (c * c + s * s))
    errs = {}
    for n in (16, 32):
        s = _sde(n)
        g = torch.Generator().manual_seed(0)
        x = torch.randn(100000, generator=g).double() * (c * c + s.init_noise_sigma ** 2) ** 0.5
        noises = [torch.randn(x.shape, generator=g).double() for _ in range(len(s.timesteps))]
        ref = _dpmpp_sde_direct([float(v) for v in s._sig], den, x.clone(), noises)
        y = x.clone()
        for i, t in enumerate(s.timesteps):
            sigma = float(s.sigmas[i])
            y = s.step((y - den(y, sigma)) / sigma, t, y, noises[i])
        assert float((y - ref).abs().max()) < 1e-4
        errs[n] = abs(float(y.var()) / (c * c) - 1)
        if n == 32:                                                # (iii) independent stage-2 noise: a different, worse sampler
            s2 = _sde(n)
            z = x.clone()
            for i, t in enumerate(s2.timesteps):
                sigma = float(s2.sigmas[i])
                if i % 2 == 1 and s2._n1 is not None:
                    s2._n1 = torch.randn(x.shape, generator=g)
                z = s2.step((z - den(z, sigma)) / sigma, t, z, noises[i])
            assert abs(float(z.var()) / (c * c) - 1) > 1.5 * errs[32]
    assert errs[32] < 0.6 * errs[16] and errs[32] < 0.1, errs


def test_dpm_sde_point_mass_and_pruning():
    s = _sde(8)
    x0 = torch.tensor([[0.7], [-1.3], [0.2]])
    x = torch.tensor([[0.3], [-0.5], [1.1]]) * s.init_noise_sigma
    keep = torch.tensor([0, 2])
    g = torch.Generator().manual_seed(3)
    for i, t in enumerate(s.timesteps):
        if i == 5:                                   # prune between a first and a second stage: the stored state follows the views
            x, x0 = x[keep], x0[keep]
            s.prune(keep)
        eps = (x - x0) / float(s.sigmas[i])
        x = s.step(eps, t, x, torch.randn(x.shape, generator=g) * 0.0)
    assert x.shape == (2, 1) and float((x - x0).abs().max()) < 1e-5
    with pytest.raises(RuntimeError):
        s2 = _sde(4)
        s2.step(torch.zeros(1), s2.timesteps[1], torch.zeros(1), torch.zeros(1))
