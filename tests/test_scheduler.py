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
