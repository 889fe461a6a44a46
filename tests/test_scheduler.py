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
