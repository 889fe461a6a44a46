"""DDIMScheduler and UniPCMultistepScheduler with the diffusers surface the reference pipelines duck-type
(`.config`, `.set_timesteps`, `.timesteps`, `.order`, `.init_noise_sigma`, `.scale_model_input`,
`.step(..., eta=, generator=, return_dict=False)[0]`, `.add_noise`, `from_config`),
restating diffusers==0.27.0 DDIMScheduler as called by
powerpaint/pipelines/pipeline_PowerPaint.py:906 (set_timesteps), :993 (scale_model_input),
:1023 (step), :642 (init_noise_sigma).

Host side only computes the schedule (numpy / CPU tensors). The arithmetic of `step` on CUDA
tensors runs in the fused CFG+DDIM kernel (`pp_cfg_ddim_step`); there is no CPU `step`.
`step_coefficients()` exports the per-step scalar table the captured CUDA graph indexes.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import List, Optional, Tuple, Union

import numpy as np
import torch


class _SchedConfig(SimpleNamespace):
    def __getitem__(self, k):
        return getattr(self, k)

    def __contains__(self, k):
        return hasattr(self, k)

    def get(self, k, default=None):
        return getattr(self, k, default)

    def keys(self):
        return self.__dict__.keys()


class DDIMSchedulerOutput(SimpleNamespace):
    pass


class DDIMScheduler:
    order = 1
    _compatibles: List[str] = []

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 beta_schedule: str = "scaled_linear", trained_betas=None, clip_sample: bool = False,
                 set_alpha_to_one: bool = False, steps_offset: int = 1, prediction_type: str = "epsilon",
                 thresholding: bool = False, dynamic_thresholding_ratio: float = 0.995, clip_sample_range: float = 1.0,
                 sample_max_value: float = 1.0, timestep_spacing: str = "leading", rescale_betas_zero_snr: bool = False,
                 **unused):
        if prediction_type != "epsilon":
            raise NotImplementedError("only prediction_type='epsilon' is on the PowerPaint hot path")
        if clip_sample or thresholding or rescale_betas_zero_snr:
            raise NotImplementedError("clip_sample / thresholding / zero-SNR rescale are off in the SD config")
        self.config = _SchedConfig(
            num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
            beta_schedule=beta_schedule, trained_betas=trained_betas, clip_sample=clip_sample,
            set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset, prediction_type=prediction_type,
            thresholding=thresholding, dynamic_thresholding_ratio=dynamic_thresholding_ratio,
            clip_sample_range=clip_sample_range, sample_max_value=sample_max_value,
            timestep_spacing=timestep_spacing, rescale_betas_zero_snr=rescale_betas_zero_snr)
        if trained_betas is not None:
            betas = torch.tensor(trained_betas, dtype=torch.float32)
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(f"beta_schedule {beta_schedule}")
        self.betas = betas
        self.alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    @classmethod
    def from_config(cls, config, **kwargs):
        d = dict(config.__dict__) if hasattr(config, "__dict__") else dict(config)
        d = {k: v for k, v in d.items() if not k.startswith("_")}
        d.update(kwargs)
        return cls(**d)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder=None, revision=None, local_files_only=False,
                        cache_dir=None, **kwargs):
        """`SchedulerMixin.from_pretrained`: `<dir>[/<subfolder>]/scheduler_config.json` of a checkpoint in the diffusers
        layout (keys of other scheduler classes that have no meaning here are dropped by the constructor)"""
        from .loading import load_json, resolve_checkpoint_dir

        d = resolve_checkpoint_dir(pretrained_model_name_or_path, subfolder, revision, local_files_only, cache_dir)
        return cls.from_config(load_json(d, "scheduler_config.json"), **kwargs)

    def save_pretrained(self, save_directory, **unused):
        import json
        import os

        os.makedirs(save_directory, exist_ok=True)
        cfg = {"_class_name": type(self).__name__, "_diffusers_version": "0.27.0"}
        cfg.update({k: v for k, v in vars(self.config).items() if not k.startswith("_")})
        with open(os.path.join(save_directory, "scheduler_config.json"), "w") as f:
            json.dump(cfg, f, indent=2, sort_keys=True)

    def scale_model_input(self, sample: torch.Tensor, timestep=None) -> torch.Tensor:
        return sample

    def set_timesteps(self, num_inference_steps: int, device: Union[str, torch.device, None] = None):
        c = self.config
        if num_inference_steps > c.num_train_timesteps:
            raise ValueError(f"`num_inference_steps`: {num_inference_steps} cannot be larger than "
                             f"`self.config.train_timesteps`: {c.num_train_timesteps}")
        self.num_inference_steps = num_inference_steps
        if c.timestep_spacing == "linspace":
            ts = np.linspace(0, c.num_train_timesteps - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        elif c.timestep_spacing == "leading":
            ratio = c.num_train_timesteps // num_inference_steps
            ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
            ts += c.steps_offset
        elif c.timestep_spacing == "trailing":
            ratio = c.num_train_timesteps / num_inference_steps
            ts = np.round(np.arange(c.num_train_timesteps, 0, -ratio)).astype(np.int64) - 1
        else:
            raise ValueError(f"{c.timestep_spacing} is not supported")
        self.timesteps = torch.from_numpy(ts).to(device)

    # ---- schedule scalars
    def _alphas(self, timestep: int) -> Tuple[float, float]:
        prev = timestep - self.config.num_train_timesteps // self.num_inference_steps
        a_t = float(self.alphas_cumprod[timestep])
        a_prev = float(self.alphas_cumprod[prev]) if prev >= 0 else float(self.final_alpha_cumprod)
        return a_t, a_prev

    def step_coefficients(self, timesteps=None, eta: float = 0.0) -> torch.Tensor:
        """[n_steps, 8] fp32 rows {sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev-sigma^2), sigma,
        0, 0, 0} consumed by pp_cfg_ddim_step (include/powerpaint_b200.h)."""
        if self.num_inference_steps is None:
            raise ValueError("call set_timesteps first")
        ts = self.timesteps if timesteps is None else timesteps
        rows = []
        for t in [int(x) for x in ts]:
            a_t, a_prev = self._alphas(t)
            var = (1 - a_prev) / (1 - a_t) * (1 - a_t / a_prev)
            sigma = eta * math.sqrt(max(var, 0.0))
            rows.append([math.sqrt(a_t), math.sqrt(1 - a_t), math.sqrt(a_prev),
                         math.sqrt(max(1 - a_prev - sigma * sigma, 0.0)), sigma, 0.0, 0.0, 0.0])
        return torch.tensor(rows, dtype=torch.float32)

    def step(self, model_output: torch.Tensor, timestep: int, sample: torch.Tensor, eta: float = 0.0,
             use_clipped_model_output: bool = False, generator=None, variance_noise: Optional[torch.Tensor] = None,
             return_dict: bool = True):
        """x_t -> x_{t-1} on CUDA tensors through the fused kernel (no CFG: model_output is used as is)."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after "
                             "creating the scheduler")
        if not sample.is_cuda:
            raise RuntimeError("DDIMScheduler.step runs in the CUDA kernel pp_cfg_ddim_step; "
                               "CPU tensors are not supported (no CPU fallback on the hot path)")
        from . import ops

        nb, c, h, w = sample.shape
        if c != 4:
            raise ValueError("the fused step kernel handles 4-channel latents")
        coef = self.step_coefficients([int(timestep)], eta).to(sample.device)
        lat = ops.nhwc_fp32_from_nchw(sample)
        eps = ops.nhwc_fp32_from_nchw(model_output)
        noise = None
        if eta > 0:
            if variance_noise is None:
                variance_noise = torch.randn(model_output.shape, generator=generator,
                                             device=generator.device if generator is not None else sample.device,
                                             dtype=torch.float32).to(sample.device)
            noise = ops.nhwc_fp32_from_nchw(variance_noise)
        ops.run(ops.cfg_ddim_desc(eps=eps, eps_fp32=True, eps_ld=4, latents=lat, coef=coef, step_idx=None,
                                  advance_step=False, noise=noise, guidance_scale=1.0, do_cfg=False, batch=nb,
                                  hw=h * w))
        prev = lat.view(nb, h, w, 4).permute(0, 3, 1, 2).contiguous().to(sample.dtype)
        if not return_dict:
            return (prev,)
        # x0 prediction of the output object (diffusers' `pred_original_sample`, epsilon parametrisation); the loop never
        # reads it: one elementwise expression on the caller's tensors, outside the fused kernel
        a_t, _ = self._alphas(int(timestep))
        x0 = ((sample.float() - math.sqrt(1 - a_t) * model_output.float()) / math.sqrt(a_t)).to(sample.dtype)
        return DDIMSchedulerOutput(prev_sample=prev, pred_original_sample=x0)

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        a = self.alphas_cumprod.to(original_samples.device)[timesteps.to(original_samples.device)]
        a = a.to(original_samples.dtype)
        while a.dim() < original_samples.dim():
            a = a.unsqueeze(-1)
        return a ** 0.5 * original_samples + (1 - a) ** 0.5 * noise

    def __len__(self):
        return self.config.num_train_timesteps


class UniPCMultistepScheduler(DDIMScheduler):
    """diffusers `UniPCMultistepScheduler` as the v2 app builds it — `UniPCMultistepScheduler.from_config(
    pipe.scheduler.config)` (reference app.py:197): solver_order 2, solver_type "bh2", predict_x0, epsilon
    prediction, lower_order_final, beta schedule / spacing / steps_offset inherited from the DDIM config.

    The multistep predictor-corrector update is linear in the tensors it touches (current sample, last corrected
    sample, the previous x0 predictions), so `unipc_coefficients()` folds the whole schedule (log-SNR steps, the
    B(h) terms, the 2x2 solve of the corrector) into 10 scalars per step, computed in float64 on the host; the
    arithmetic on the latents runs in the fused CFG + UniPC kernel (`pp_unipc_step`), like DDIM's."""

    order = 1
    kind = "unipc"

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas=None, solver_order: int = 2,
                 prediction_type: str = "epsilon", thresholding: bool = False, predict_x0: bool = True,
                 solver_type: str = "bh2", lower_order_final: bool = True, disable_corrector=(),
                 use_karras_sigmas: bool = False, timestep_spacing: str = "linspace", steps_offset: int = 0, **unused):
        if solver_order != 2 or not predict_x0 or solver_type != "bh2" or use_karras_sigmas or disable_corrector:
            raise NotImplementedError("only the app's UniPC configuration is built: solver_order=2, bh2, predict_x0, "
                                      "no Karras sigmas, corrector on")
        super().__init__(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                         beta_schedule=beta_schedule, trained_betas=trained_betas, steps_offset=steps_offset,
                         prediction_type=prediction_type, thresholding=thresholding, timestep_spacing=timestep_spacing)
        self.config.solver_order = solver_order
        self.config.predict_x0 = predict_x0
        self.config.solver_type = solver_type
        self.config.lower_order_final = lower_order_final
        self.config.use_karras_sigmas = use_karras_sigmas
        self.sigmas = None
        self._state = None

    def set_timesteps(self, num_inference_steps: int, device=None):
        c = self.config
        T = c.num_train_timesteps
        n = num_inference_steps
        if c.timestep_spacing == "linspace":
            ts = np.linspace(0, T - 1, n + 1).round()[::-1][:-1].copy().astype(np.int64)
        elif c.timestep_spacing == "leading":
            step_ratio = T // (n + 1)
            ts = (np.arange(0, n + 1) * step_ratio).round()[::-1][:-1].copy().astype(np.int64)
            ts += c.steps_offset
        elif c.timestep_spacing == "trailing":
            step_ratio = T / n
            ts = np.arange(T, 0, -step_ratio).round().copy().astype(np.int64) - 1
        else:
            raise ValueError(f"{c.timestep_spacing} is not supported")
        ac = self.alphas_cumprod.numpy().astype(np.float64)
        sig = ((1 - ac) / ac) ** 0.5
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        self.sigmas = np.concatenate([sig, [((1 - ac[0]) / ac[0]) ** 0.5]]).astype(np.float32).astype(np.float64)
        self.timesteps = torch.from_numpy(ts).to(device)
        self.num_inference_steps = n
        self._state = None

    def step_coefficients(self, timesteps=None, eta: float = 0.0) -> torch.Tensor:
        """the shared 8-float rows (guidance / side-net scale are filled in by the denoiser); DDIM columns unused"""
        n = len(self.timesteps if timesteps is None else timesteps)
        return torch.zeros(n, 8, dtype=torch.float32)

    def unipc_coefficients(self, first: int = 0) -> torch.Tensor:
        """[n, 12] fp32 rows for `pp_unipc_step` (include/powerpaint_b200.h), starting at schedule index `first`
        (strength < 1 starts part-way: the multistep history then starts empty there, like the reference)."""
        if self.sigmas is None:
            raise ValueError("call set_timesteps first")
        sg = self.sigmas
        n = len(sg) - 1

        def alpha_sigma(s):
            a = 1.0 / math.sqrt(s * s + 1.0)
            return a, s * a

        def lam(s):
            a, st = alpha_sigma(s)
            return math.log(a) - math.log(st)

        rows = []
        lower = 0
        this_order = 1
        for i in range(first, n):
            a_i, s_i = alpha_sigma(sg[i])
            u = [1.0 / a_i, -s_i / a_i, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]
            if i > first:  # corrector with the order chosen at the previous step
                order = this_order
                a_t, s_t = a_i, s_i
                _, s_s0 = alpha_sigma(sg[i - 1])
                h = lam(sg[i]) - lam(sg[i - 1])
                hh = -h
                h_phi_1 = math.expm1(hh)
                B_h = math.expm1(hh)
                h_phi_k = h_phi_1 / hh - 1.0
                if order == 1:
                    rho_last, rho0_over_rk = 0.5, 0.0
                else:
                    rk = (lam(sg[i - 2]) - lam(sg[i - 1])) / h
                    b1 = h_phi_k / B_h
                    b2 = (h_phi_k / hh - 0.5) * 2.0 / B_h
                    # R = [[1, 1], [rk, 1]] rhos = [b1, b2]
                    rho0 = (b1 - b2) / (1.0 - rk)
                    rho_last = b1 - rho0
                    rho0_over_rk = rho0 / rk
                u[2] = 1.0
                u[3] = s_t / s_s0
                u[4] = -a_t * h_phi_1 + a_t * B_h * (rho0_over_rk + rho_last)
                u[5] = -a_t * B_h * rho0_over_rk
                u[6] = -a_t * B_h * rho_last
            order = min(2, n - i) if self.config.lower_order_final else 2
            this_order = min(order, lower + 1)
            # predictor towards sigma[i + 1]
            a_t, s_t = alpha_sigma(sg[i + 1])
            h = lam(sg[i + 1]) - lam(sg[i])
            hh = -h
            h_phi_1 = math.expm1(hh)
            B_h = math.expm1(hh)
            u[7] = s_t / s_i
            u[8] = -a_t * h_phi_1
            if this_order == 2:
                rk = (lam(sg[i - 1]) - lam(sg[i])) / h
                u[8] += a_t * B_h * 0.5 / rk
                u[9] = -a_t * B_h * 0.5 / rk
            if lower < 2:
                lower += 1
            rows.append(u)
        return torch.tensor(rows, dtype=torch.float32)

    def step(self, model_output: torch.Tensor, timestep: int, sample: torch.Tensor, return_dict: bool = True, **unused):
        """eager x_t -> x_{t-1} on CUDA tensors through the fused kernel; keeps the multistep state like diffusers
        (reset by `set_timesteps`)"""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after "
                             "creating the scheduler")
        if not sample.is_cuda:
            raise RuntimeError("UniPCMultistepScheduler.step runs in the CUDA kernel pp_unipc_step; CPU tensors are "
                               "not supported (no CPU fallback on the hot path)")
        from . import ops

        nb, c, h, w = sample.shape
        if c != 4:
            raise ValueError("the fused step kernel handles 4-channel latents")
        st = self._state
        if st is None:
            z = lambda: torch.zeros(nb, h * w, 4, dtype=torch.float32, device=sample.device)  # noqa: E731
            st = self._state = dict(i=0, last=z(), m1=z(), m2=z(), ucoef=self.unipc_coefficients().to(sample.device),
                                    coef=torch.zeros(len(self.sigmas) - 1, 8, device=sample.device),
                                    idx=torch.zeros(1, dtype=torch.int32, device=sample.device))
        st["idx"].fill_(st["i"])
        lat = ops.nhwc_fp32_from_nchw(sample)
        eps = ops.nhwc_fp32_from_nchw(model_output)
        ops.run(ops.unipc_desc(eps=eps, eps_fp32=True, eps_ld=4, latents=lat, last_sample=st["last"], m1=st["m1"],
                               m2=st["m2"], coef=st["coef"], ucoef=st["ucoef"], step_idx=st["idx"], advance_step=False,
                               do_cfg=False, batch=nb, hw=h * w))
        st["i"] += 1
        prev = lat.view(nb, h, w, 4).permute(0, 3, 1, 2).contiguous().to(sample.dtype)
        if not return_dict:
            return (prev,)
        return DDIMSchedulerOutput(prev_sample=prev)
