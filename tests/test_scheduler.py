"""DDIM schedule: known answers of the SD config and product-vs-oracle agreement."""
import math

import pytest
import torch

from oracle.ddim import DDIMOracle
from powerpaint_b200.schedulers import DDIMScheduler


def test_timesteps_known_answers():
    for cls in (DDIMScheduler, DDIMOracle):
        s = cls()
        s.set_timesteps(50)
        assert s.timesteps.tolist() == list(range(981, 0, -20))
        s.set_timesteps(20)
        assert s.timesteps.tolist() == list(range(951, 0, -50))
        assert s.init_noise_sigma == 1.0 and s.order == 1


def test_alphas_cumprod_known_answers():
    s = DDIMScheduler()
    a = s.alphas_cumprod
    assert a.shape == (1000,)
    assert abs(float(a[0]) - 0.99915) < 1e-6          # 1 - 0.00085
    assert abs(float(a[999]) - 0.0046600) < 2e-6       # SD-1.5 scaled_linear endpoint
    assert float(s.final_alpha_cumprod) == float(a[0])  # set_alpha_to_one=False
    assert torch.equal(a, DDIMOracle().alphas_cumprod)
    assert torch.all(a[1:] < a[:-1])


def test_step_coefficients_match_oracle_step():
    sp, so = DDIMScheduler(), DDIMOracle()
    sp.set_timesteps(50)
    so.set_timesteps(50)
    g = torch.Generator().manual_seed(0)
    x, eps, nz = (torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64) for _ in range(3))
    for eta in (0.0, 0.5, 1.0):
        coef = sp.step_coefficients(eta=eta).double()
        for i, t in enumerate(sp.timesteps.tolist()):
            sa, s1a, sap, dirc, sigma = coef[i, :5]
            x0 = (x - s1a * eps) / sa
            mine = sap * x0 + dirc * eps + sigma * nz
            ref = so.step(eps, t, x, eta=eta, variance_noise=nz)
            assert torch.allclose(mine, ref, atol=1e-5), (eta, t)
    # last step uses final_alpha_cumprod = alphas_cumprod[0]
    last = sp.step_coefficients()[-1]
    assert abs(float(last[2]) - math.sqrt(float(sp.alphas_cumprod[0]))) < 1e-6


def test_config_surface_and_errors():
    s = DDIMScheduler()
    assert s.config.steps_offset == 1 and s.config["beta_schedule"] == "scaled_linear" and "clip_sample" in s.config
    s2 = DDIMScheduler.from_config(s.config)
    assert torch.equal(s2.alphas_cumprod, s.alphas_cumprod)
    with pytest.raises(NotImplementedError):
        DDIMScheduler(prediction_type="v_prediction")
    with pytest.raises(ValueError):
        s.step(torch.zeros(1, 4, 8, 8), 981, torch.zeros(1, 4, 8, 8))  # set_timesteps not called
    s.set_timesteps(10)
    with pytest.raises(RuntimeError):
        s.step(torch.zeros(1, 4, 8, 8), int(s.timesteps[0]), torch.zeros(1, 4, 8, 8))  # CPU tensors: no fallback
    x0, n = torch.ones(2, 4, 2, 2), torch.zeros(2, 4, 2, 2)
    out = s.add_noise(x0, n, torch.tensor([0, 999]))
    assert torch.allclose(out[0], x0[0] * float(s.alphas_cumprod[0]) ** 0.5)


def test_strength_subset_of_the_schedule_matches_oracle():
    """strength < 1 keeps the LAST int(n * strength) timesteps (ref:pipeline_PowerPaint.py:713-720); the
    coefficient rows for that suffix must drive the same trajectory as the oracle's DDIM step"""
    from powerpaint_b200.pipelines.pipeline_PowerPaint import StableDiffusionInpaintPipeline as P

    sp, so = DDIMScheduler(), DDIMOracle()
    sp.set_timesteps(20)
    so.set_timesteps(20)

    class _Stub:  # get_timesteps only touches self.scheduler
        scheduler = sp

    ts, n = P.get_timesteps(_Stub(), 20, 0.35, "cpu")
    assert n == 7 and ts.tolist() == sp.timesteps[13:].tolist() == list(range(301, 0, -50))
    ts1, n1 = P.get_timesteps(_Stub(), 20, 1.0, "cpu")
    assert n1 == 20 and ts1.tolist() == sp.timesteps.tolist()
    ts0, n0 = P.get_timesteps(_Stub(), 20, 0.01, "cpu")
    assert n0 == 0 and len(ts0) == 0
    g = torch.Generator().manual_seed(1)
    x0, nz, eps = (torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64) for _ in range(3))
    # initial latents: image latents noised to the first kept timestep (scheduler.add_noise == oracle.add_noise)
    x = sp.add_noise(x0, nz, ts[:1].repeat(2))
    assert torch.allclose(x, so.add_noise(x0, nz, ts[:1].repeat(2)))
    coef = sp.step_coefficients(ts).double()
    ref = x.clone()
    for i, t in enumerate(ts.tolist()):
        sa, s1a, sap, dirc, _ = coef[i, :5]
        x = sap * ((x - s1a * eps) / sa) + dirc * eps
        ref = so.step(eps, t, ref)
    assert torch.allclose(x, ref, atol=1e-5)


def test_unipc_folded_coefficients_match_stepwise_oracle():
    """UniPCMultistepScheduler.from_config(DDIM config) (ref app.py:197): timesteps and the 10 folded scalars per
    step reproduce the stepwise predictor-corrector restatement (oracle/unipc.py) on random tensors"""
    import torch

    from oracle.unipc import UniPCOracle
    from powerpaint_b200.schedulers import DDIMScheduler, UniPCMultistepScheduler

    for n in (50, 20, 7, 2, 1):
        o = UniPCOracle()
        o.set_timesteps(n)
        s = UniPCMultistepScheduler.from_config(DDIMScheduler().config)
        s.set_timesteps(n)
        assert s.config.timestep_spacing == "leading" and s.config.steps_offset == 1 and s.config.solver_order == 2
        assert [int(t) for t in o.timesteps] == [int(t) for t in s.timesteps]
        u = s.unipc_coefficients().double()
        assert u.shape == (n, 12) and float(u[0, 2]) == 0.0 and (n == 1 or float(u[1, 2]) == 1.0)
        g = torch.Generator().manual_seed(n)
        xo = torch.randn(2, 4, 8, 8, generator=g)
        xs = xo.double().clone()
        last, m1, m2 = torch.zeros_like(xs), torch.zeros_like(xs), torch.zeros_like(xs)
        for i, t in enumerate(o.timesteps):
            eps = torch.randn(2, 4, 8, 8, generator=g)
            xo = o.step(eps, int(t), xo)
            r = u[i]
            mt = r[0] * xs + r[1] * eps.double()
            xc = r[3] * last + r[4] * m1 + r[5] * m2 + r[6] * mt if r[2] != 0 else xs
            last, m2, m1, xs = xc, m1, mt, r[7] * xc + r[8] * mt + r[9] * m1
        assert ((xs - xo.double()).norm() / xo.double().norm()).item() < 1e-4
    s = UniPCMultistepScheduler.from_config(DDIMScheduler().config)
    s.set_timesteps(50)
    assert [int(t) for t in s.timesteps[:3]] == [951, 932, 913] and int(s.timesteps[-1]) == 20
    # strength < 1: the history starts empty at the first kept step
    u5 = s.unipc_coefficients(first=45)
    assert u5.shape == (5, 12) and float(u5[0, 2]) == 0.0 and float(u5[0, 9]) == 0.0 and float(u5[-1, 9]) == 0.0
    import pytest

    with pytest.raises(NotImplementedError):
        UniPCMultistepScheduler(solver_order=3)


def _gaussian_denoiser_eps(x, alpha, sigma, s2):
    """optimal epsilon-prediction for data ~ N(0, s2): E[x0 | x_t] = alpha s2 / (alpha^2 s2 + sigma^2) x_t"""
    return (x - alpha * (alpha * s2 / (alpha * alpha * s2 + sigma * sigma)) * x) / sigma


def test_unipc_has_the_published_order_of_accuracy():
    """No output of diffusers' UniPCMultistepScheduler can be produced here, so the stepwise restatement
    (oracle/unipc.py) AND the product's folded per-step scalars are held to the mathematics: for data ~ N(0, s^2) the
    probability-flow ODE has the closed form x_t = x_T sqrt(alpha_t^2 s^2 + sigma_t^2) / sqrt(alpha_T^2 s^2 + sigma_T^2)
    and the optimal denoiser is linear, so the global error of a run is measurable exactly. UniPC-p = UniP-p + UniC-p has
    order of accuracy p + 1 (Zhao et al. 2023, Thm 3.1): halving the step must cut the error ~8x for solver_order 2 (~4x
    for 1); any wrong coefficient (rho, B(h), r_k, the lambda differences) drops it to first order. The comparison point
    is t = 199 (the smooth part of the schedule; the last steps towards t = 0 have h = O(1) whatever the step count)."""
    import math

    import torch

    from oracle.unipc import UniPCOracle
    from powerpaint_b200.schedulers import UniPCMultistepScheduler

    s2 = 0.25
    x_init = torch.tensor([1.3, -0.7, 0.2], dtype=torch.float64)

    def exact(a0, s0, a1, s1):
        return x_init * math.sqrt(a1 * a1 * s2 + s1 * s1) / math.sqrt(a0 * a0 * s2 + s0 * s0)

    def oracle_error(n, k, order):
        o = UniPCOracle(timestep_spacing="trailing", solver_order=order)
        o.set_timesteps(n)
        o.sigmas = o.sigmas.double()
        x = x_init.clone()
        a0, s0 = [float(v) for v in o._alpha_sigma(o.sigmas[0])]
        for i in range(k):
            a, s = o._alpha_sigma(o.sigmas[o.step_index])
            x = o.step(_gaussian_denoiser_eps(x, a, s, s2), int(o.timesteps[i]), x)
        assert int(o.timesteps[k]) == 199
        a1, s1 = [float(v) for v in o._alpha_sigma(o.sigmas[k])]
        ref = exact(a0, s0, a1, s1)
        return ((x - ref).norm() / ref.norm()).item()

    def product_error(n, k):
        s = UniPCMultistepScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                                    timestep_spacing="trailing")
        s.set_timesteps(n)
        assert int(s.timesteps[k]) == 199
        u = s.unipc_coefficients().double()
        alpha = [1.0 / float(r[0]) for r in u]          # m = x / alpha - sigma / alpha * eps
        sigma = [-float(r[1]) / float(r[0]) for r in u]
        xs = x_init.clone()
        last, m1, m2 = torch.zeros_like(xs), torch.zeros_like(xs), torch.zeros_like(xs)
        for i in range(k):  # the update pp_unipc_step applies (same recursion as the folding test above)
            r = u[i]
            eps = _gaussian_denoiser_eps(xs, alpha[i], sigma[i], s2)
            mt = r[0] * xs + r[1] * eps
            xc = r[3] * last + r[4] * m1 + r[5] * m2 + r[6] * mt if r[2] != 0 else xs
            last, m2, m1, xs = xc, m1, mt, r[7] * xc + r[8] * mt + r[9] * m1
        ref = exact(alpha[0], sigma[0], alpha[k], sigma[k])
        return ((xs - ref).norm() / ref.norm()).item()

    grid = [(10, 8), (20, 16), (40, 32)]  # trailing spacing: steps of 100 / 50 / 25 from t = 999 down to t = 199
    e1 = [oracle_error(n, k, 1) for n, k in grid]
    e2 = [oracle_error(n, k, 2) for n, k in grid]
    ep = [product_error(n, k) for n, k in grid]
    order1 = [math.log2(e1[i] / e1[i + 1]) for i in range(2)]
    order2 = [math.log2(e2[i] / e2[i + 1]) for i in range(2)]
    orderp = [math.log2(ep[i] / ep[i + 1]) for i in range(2)]
    assert all(1.4 < o < 2.4 for o in order1), (e1, order1)          # UniPC-1: second order
    assert all(o > 2.7 for o in order2) and e2[-1] < 5e-5, (e2, order2)  # UniPC-2: third order (or better)
    assert all(o > 2.7 for o in orderp) and ep[-1] < 5e-5, (ep, orderp)
    # same iterates, folded (fp32 scalars) or stepwise (fp64 here)
    assert all(abs(a - b) <= 1e-2 * b for a, b in zip(ep, e2)), (ep, e2)


def test_ddim_converges_to_the_exact_flow_with_first_order():
    """the same closed-form problem for DDIM (eta = 0), oracle step and the product's coefficient rows: the error at
    t = 201 halves when the step halves (first-order exponential integrator) and the two agree"""
    import math

    import torch

    from oracle.ddim import DDIMOracle
    from powerpaint_b200.schedulers import DDIMScheduler

    s2 = 0.25
    x_init = torch.tensor([1.3, -0.7, 0.2], dtype=torch.float64)
    errs_o, errs_p = [], []
    for n in (10, 20, 40):
        so, sp = DDIMOracle(), DDIMScheduler()
        so.set_timesteps(n)
        sp.set_timesteps(n)
        ts = [int(t) for t in so.timesteps]
        k = ts.index(201)
        ac = so.alphas_cumprod.double()
        coef = sp.step_coefficients(sp.timesteps).double()
        xo, xp = x_init.clone(), x_init.clone()
        for i in range(k):
            a, s = float(ac[ts[i]]) ** 0.5, float(1 - ac[ts[i]]) ** 0.5
            xo = so.step(_gaussian_denoiser_eps(xo, a, s, s2), ts[i], xo)
            eps = _gaussian_denoiser_eps(xp, a, s, s2)
            sa, s1a, sap, dirc = [float(v) for v in coef[i, :4]]
            xp = sap * ((xp - s1a * eps) / sa) + dirc * eps
        a0, s0 = float(ac[ts[0]]) ** 0.5, float(1 - ac[ts[0]]) ** 0.5
        a1, s1 = float(ac[201]) ** 0.5, float(1 - ac[201]) ** 0.5
        ref = x_init * math.sqrt(a1 * a1 * s2 + s1 * s1) / math.sqrt(a0 * a0 * s2 + s0 * s0)
        errs_o.append(((xo - ref).norm() / ref.norm()).item())
        errs_p.append(((xp - ref).norm() / ref.norm()).item())
    for errs in (errs_o, errs_p):
        orders = [math.log2(errs[i] / errs[i + 1]) for i in range(2)]
        assert all(0.7 < o < 1.4 for o in orders) and errs[-1] < 2e-2, (errs, orders)
    assert all(abs(a - b) <= 1e-3 * b for a, b in zip(errs_p, errs_o)), (errs_p, errs_o)


def test_scheduler_from_pretrained_round_trip(tmp_path):
    """`scheduler/scheduler_config.json` of a checkpoint directory, and the app's swap-by-config (ref:app.py:197)"""
    from powerpaint_b200.schedulers import DDIMScheduler, UniPCMultistepScheduler

    s = DDIMScheduler()
    s.save_pretrained(str(tmp_path / "scheduler"))
    s2 = DDIMScheduler.from_pretrained(str(tmp_path), subfolder="scheduler")
    assert vars(s2.config) == vars(s.config)
    u = UniPCMultistepScheduler.from_config(s2.config)
    u.save_pretrained(str(tmp_path / "unipc"))
    u2 = UniPCMultistepScheduler.from_pretrained(str(tmp_path / "unipc"))
    assert vars(u2.config) == vars(u.config) and u2.config.solver_order == 2
    u2.set_timesteps(20)
    u.set_timesteps(20)
    assert [int(t) for t in u2.timesteps] == [int(t) for t in u.timesteps]
