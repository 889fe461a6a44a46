"""ORACLE (test infrastructure, not product): UniPCMultistepScheduler restated step by step.

The v2 app swaps the pipeline's scheduler for `UniPCMultistepScheduler.from_config(pipe.scheduler.config)`
(reference app.py:197). The class lives in the un-vendored dependency diffusers==0.27.0
(schedulers/scheduling_unipc_multistep.py); this restates its published algorithm (UniPC, Zhao et al. 2023,
"bh2" variant, data prediction) for the configuration that call produces: solver_order 2, predict_x0,
prediction_type epsilon, lower_order_final, no thresholding / Karras sigmas, and the beta schedule / timestep
spacing / steps_offset inherited from the DDIM config. Tensor arithmetic in fp32 torch like the original.
PARITY UNPINNED against diffusers (absent here: no golden vectors of the original exist). Held to the mathematics
instead: tests/test_scheduler.py::test_unipc_has_the_published_order_of_accuracy measures this class (and the product's
folded scalars) against the closed-form probability-flow solution for Gaussian data: third-order convergence for
solver_order 2, second for 1 — a wrong coefficient anywhere drops it to first order.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch


class UniPCOracle:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 solver_order=2, timestep_spacing="leading", steps_offset=1, lower_order_final=True):
        if beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        self.T = num_train_timesteps
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.solver_order = solver_order
        self.timestep_spacing = timestep_spacing
        self.steps_offset = steps_offset
        self.lower_order_final = lower_order_final
        self.init_noise_sigma = 1.0
        self.order = 1

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, n: int):
        if self.timestep_spacing == "linspace":
            ts = np.linspace(0, self.T - 1, n + 1).round()[::-1][:-1].copy().astype(np.int64)
        elif self.timestep_spacing == "leading":
            step_ratio = self.T // (n + 1)
            ts = (np.arange(0, n + 1) * step_ratio).round()[::-1][:-1].copy().astype(np.int64)
            ts += self.steps_offset
        elif self.timestep_spacing == "trailing":
            step_ratio = self.T / n
            ts = np.arange(self.T, 0, -step_ratio).round().copy().astype(np.int64) - 1
        else:
            raise ValueError(self.timestep_spacing)
        ac = self.alphas_cumprod.numpy()
        sigmas = np.array(((1 - ac) / ac) ** 0.5)
        sigmas = np.interp(ts, np.arange(0, len(sigmas)), sigmas)
        sigma_last = ((1 - ac[0]) / ac[0]) ** 0.5
        self.sigmas = torch.from_numpy(np.concatenate([sigmas, [sigma_last]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts)
        self.num_inference_steps = n
        self.model_outputs: List[Optional[torch.Tensor]] = [None] * self.solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self.step_index = 0
        self.this_order = 1

    @staticmethod
    def _alpha_sigma(sigma):
        alpha_t = 1 / ((sigma ** 2 + 1) ** 0.5)
        return alpha_t, sigma * alpha_t

    def _convert(self, eps, sample):
        alpha_t, sigma_t = self._alpha_sigma(self.sigmas[self.step_index])
        return (sample - sigma_t * eps) / alpha_t

    def _uni_p(self, sample, order):
        m0 = self.model_outputs[-1]
        x = sample
        alpha_t, sigma_t = self._alpha_sigma(self.sigmas[self.step_index + 1])
        alpha_s0, sigma_s0 = self._alpha_sigma(self.sigmas[self.step_index])
        lambda_t = torch.log(alpha_t) - torch.log(sigma_t)
        lambda_s0 = torch.log(alpha_s0) - torch.log(sigma_s0)
        h = lambda_t - lambda_s0
        rks, D1s = [], []
        for i in range(1, order):
            si = self.step_index - i
            mi = self.model_outputs[-(i + 1)]
            a_si, s_si = self._alpha_sigma(self.sigmas[si])
            rk = ((torch.log(a_si) - torch.log(s_si)) - lambda_s0) / h
            rks.append(rk)
            D1s.append((mi - m0) / rk)
        hh = -h
        h_phi_1 = torch.expm1(hh)
        B_h = torch.expm1(hh)
        x_t_ = sigma_t / sigma_s0 * x - alpha_t * h_phi_1 * m0
        if D1s:
            assert order == 2  # rhos_p = [0.5]
            return x_t_ - alpha_t * B_h * (0.5 * D1s[0])
        return x_t_

    def _uni_c(self, this_m, last_sample, this_sample, order):
        m0 = self.model_outputs[-1]
        x = last_sample
        alpha_t, sigma_t = self._alpha_sigma(self.sigmas[self.step_index])
        alpha_s0, sigma_s0 = self._alpha_sigma(self.sigmas[self.step_index - 1])
        lambda_t = torch.log(alpha_t) - torch.log(sigma_t)
        lambda_s0 = torch.log(alpha_s0) - torch.log(sigma_s0)
        h = lambda_t - lambda_s0
        rks, D1s = [], []
        for i in range(1, order):
            si = self.step_index - (i + 1)
            mi = self.model_outputs[-(i + 1)]
            a_si, s_si = self._alpha_sigma(self.sigmas[si])
            rk = ((torch.log(a_si) - torch.log(s_si)) - lambda_s0) / h
            rks.append(rk)
            D1s.append((mi - m0) / rk)
        rks.append(torch.tensor(1.0))
        rks = torch.stack([torch.as_tensor(r, dtype=torch.float32) for r in rks])
        hh = -h
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        B_h = torch.expm1(hh)
        R, b = [], []
        factorial_i = 1
        for i in range(1, order + 1):
            R.append(torch.pow(rks, i - 1))
            b.append(h_phi_k * factorial_i / B_h)
            factorial_i *= i + 1
            h_phi_k = h_phi_k / hh - 1 / factorial_i
        R = torch.stack(R)
        b = torch.stack([torch.as_tensor(v, dtype=torch.float32) for v in b])
        rhos_c = torch.tensor([0.5]) if order == 1 else torch.linalg.solve(R, b)
        x_t_ = sigma_t / sigma_s0 * x - alpha_t * h_phi_1 * m0
        corr = 0
        for k, D in enumerate(D1s):
            corr = corr + rhos_c[k] * D
        return x_t_ - alpha_t * B_h * (corr + rhos_c[-1] * (this_m - m0))

    def step(self, eps, timestep, sample, **unused):
        use_corrector = self.step_index > 0 and self.last_sample is not None
        m = self._convert(eps, sample)
        if use_corrector:
            sample = self._uni_c(m, self.last_sample, sample, self.this_order)
        for i in range(self.solver_order - 1):
            self.model_outputs[i] = self.model_outputs[i + 1]
        self.model_outputs[-1] = m
        this_order = min(self.solver_order, len(self.timesteps) - self.step_index) if self.lower_order_final \
            else self.solver_order
        self.this_order = min(this_order, self.lower_order_nums + 1)
        self.last_sample = sample
        prev = self._uni_p(sample, self.this_order)
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        self.step_index += 1
        return prev
