"""ORACLE (test infrastructure, not product): the three denoising loops restated in plain fp32
PyTorch on already-prepared tensors (the part of `__call__` that is the hot path):

  v1          powerpaint/pipelines/pipeline_PowerPaint.py:988-1035
  v2 BrushNet powerpaint/pipelines/pipeline_PowerPaint_Brushnet_CA.py:1384-1449
  ControlNet  powerpaint/pipelines/pipeline_PowerPaint_ControlNet.py:1663-1735

PINNED for what the reference's own files decide: tests/golden/pipeline_{v1,brushnet,controlnet}_call.npz are the final
latents of the reference's own `__call__` (its pipeline files and its UNet / BrushNet imported unmodified over
tests/golden/diffusers_shim, generator: tests/golden/make_pipeline_golden.py); tests/test_pipeline_golden.py runs the
product's `__call__` over these loops' arithmetic against them. The arithmetic of the diffusers blocks inside the nets
stays unpinned (oracle/blocks.py).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .ddim import DDIMOracle


@torch.no_grad()
def loop_v1(unet, sched: DDIMOracle, latents, prompt_embeds, mask, masked_image_latents, guidance_scale: float,
            eta: float = 0.0, noise_fn: Optional[Callable[[int], torch.Tensor]] = None, record=None):
    """latents [B,4,h,w]; prompt_embeds [2B,77,D] (negative first, :516); mask [B,1,h,w];
    masked_image_latents [B,4,h,w]."""
    do_cfg = guidance_scale > 1.0
    if do_cfg:
        mask = torch.cat([mask] * 2)
        masked_image_latents = torch.cat([masked_image_latents] * 2)
    for i, t in enumerate(sched.timesteps):
        x = torch.cat([latents] * 2) if do_cfg else latents
        x = sched.scale_model_input(x, t)
        x = torch.cat([x, mask, masked_image_latents], dim=1)
        eps = unet(x, int(t), prompt_embeds)
        if do_cfg:
            u, c = eps.chunk(2)
            eps = u + guidance_scale * (c - u)
        latents = sched.step(eps, int(t), latents, eta=eta, variance_noise=noise_fn(i) if noise_fn else None)
        if record is not None:
            record.append(latents.clone())
    return latents


@torch.no_grad()
def loop_brushnet(unet, brushnet, sched: DDIMOracle, latents, prompt_embeds_task, prompt_embeds_u,
                  conditioning_latents, guidance_scale: float, conditioning_scale: float = 1.0, record=None,
                  keep=None):
    """conditioning_latents [2B,5,h,w] (already duplicated for CFG like the reference, :949-950);
    keep[i] = `brushnet_keep[i]` of the control_guidance window (:1369-1376, :1403-1409)"""
    do_cfg = guidance_scale > 1.0
    for i, t in enumerate(sched.timesteps):
        x = torch.cat([latents] * 2) if do_cfg else latents
        cs = conditioning_scale * (keep[i] if keep is not None else 1.0)
        d, m, u = brushnet(x, int(t), prompt_embeds_task, conditioning_latents, cs)
        eps = unet(x, int(t), prompt_embeds_u, down_block_add_samples=d, mid_block_add_sample=m,
                   up_block_add_samples=u)
        if do_cfg:
            a, c = eps.chunk(2)
            eps = a + guidance_scale * (c - a)
        latents = sched.step(eps, int(t), latents)
        if record is not None:
            record.append(latents.clone())
    return latents


@torch.no_grad()
def loop_controlnet(unet, controlnet, sched: DDIMOracle, latents, prompt_embeds, mask, masked_image_latents,
                    control_image, guidance_scale: float, conditioning_scale: float = 0.5, record=None, keep=None):
    """control_image [2B,3,H,W] in [0,1] (duplicated for CFG, :855-856); keep[i] = `controlnet_keep[i]`
    (:1652-1658, :1682-1684)"""
    do_cfg = guidance_scale > 1.0
    if do_cfg:
        mask = torch.cat([mask] * 2)
        masked_image_latents = torch.cat([masked_image_latents] * 2)
    for i, t in enumerate(sched.timesteps):
        x4 = torch.cat([latents] * 2) if do_cfg else latents
        d, m = controlnet(x4, int(t), prompt_embeds, control_image,
                          conditioning_scale * (keep[i] if keep is not None else 1.0))
        x9 = torch.cat([x4, mask, masked_image_latents], dim=1)
        eps = unet(x9, int(t), prompt_embeds, down_block_additional_residuals=d, mid_block_additional_residual=m)
        if do_cfg:
            a, c = eps.chunk(2)
            eps = a + guidance_scale * (c - a)
        latents = sched.step(eps, int(t), latents)
        if record is not None:
            record.append(latents.clone())
    return latents
