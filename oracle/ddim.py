"""ORACLE (test infrastructure, not product): restatement of diffusers==0.27.0 DDIMScheduler
with the SD config the reference pipelines force (steps_offset=1, clip_sample=False;
powerpaint/pipelines/pipeline_PowerPaint.py:205-231), as called at :906 (set_timesteps),
:993 (scale_model_input), :1023 (step), :642 (init_noise_sigma). SURVEY.md App. A.8.

PARITY UNPINNED against the reference (diffusers not installable here); the known answers
pinned in tests/test_scheduler.py (timesteps for 50/20 steps, alphas_cumprod endpoints) are the
published values of this schedule.
"""
from __future__ import annotations

import numpy as np
import torch


class DDIMOracle:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                 beta_schedule="scaled_linear", clip_sample=False, set_alpha_to_one=False,
                 steps_offset=1, prediction_type="epsilon", timestep_spacing="leading"):
        assert beta_schedule == "scaled_linear" and prediction_type == "epsilon" and not clip_sample
        assert timestep_spacing == "leading"
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        step_ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        ts += self.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step(self, model_output, timestep, sample, eta: float = 0.0, generator=None, variance_noise=None):
        t = int(timestep)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        beta_t = 1 - a_t
        x0 = (sample - beta_t ** 0.5 * model_output) / a_t ** 0.5
        variance = (1 - a_prev) / (1 - a_t) * (1 - a_t / a_prev)
        std = eta * variance ** 0.5
        direction = (1 - a_prev - std ** 2) ** 0.5 * model_output
        prev = a_prev ** 0.5 * x0 + direction
        if eta > 0:
            if variance_noise is None:
                variance_noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype)
            prev = prev + std * variance_noise
        return prev

    def add_noise(self, original, noise, timesteps):
        a = self.alphas_cumprod[timesteps].to(original.dtype)
        while a.dim() < original.dim():
            a = a[..., None]
        return a ** 0.5 * original + (1 - a) ** 0.5 * noise
