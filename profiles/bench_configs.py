"""Per-step device time of the fused denoising program for the BASELINE.json configs on ONE GPU
(per-GPU share of each config). Developer/profile script, not the judged bench. The torch-eager bf16
comparison point lives in tests/measure_torch_eager_bf16.py (it runs the oracle modules, which only test
infrastructure may import).

    python profiles/bench_configs.py [c2 c3 c4 c5]
"""
import sys
import time

import torch

sys.path.insert(0, ".")
from powerpaint_b200.denoise import FusedDenoiser  # noqa: E402
from powerpaint_b200.engine import NetConfig  # noqa: E402
from powerpaint_b200.models import BrushNetModel, ControlNetModel, UNet2DConditionModel  # noqa: E402
from powerpaint_b200.schedulers import DDIMScheduler  # noqa: E402

dev = torch.device("cuda:0")
which = set(sys.argv[1:]) or {"c2", "c3", "c4", "c5"}
sched = DDIMScheduler()
sched.set_timesteps(50)
coef = sched.step_coefficients()
FL = {64: 0.8034e12, 128: 4.674e12}


def timed(den, kw, B, label, flops_step):
    den.run(**kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    den.run(**kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{label}: {ms / 50:.2f} ms/step, {B / (ms / 1e3):.2f} images/s/GPU, {flops_step / (ms / 50 / 1e3) / 1e12:.0f} TFLOP/s, "
          f"launches/step {den.launches_per_step}", flush=True)


def inputs(B, h, cin_extra=5):
    g = torch.Generator(device=dev).manual_seed(0)
    return (torch.randn(B, 4, h, h, device=dev, generator=g), torch.randn(2 * B, 77, 768, device=dev, generator=g) * 0.5,
            torch.randn(B, cin_extra, h, h, device=dev, generator=g))


if "c2" in which or "c4" in which or "c5" in which:
    unet9 = UNet2DConditionModel.synthetic(NetConfig(in_channels=9)).to(dev)
if "c2" in which:
    lat, emb, extra = inputs(8, 64)
    timed(FusedDenoiser(unet9), dict(latents=lat, prompt_embeds=emb, timesteps=sched.timesteps, coef=coef,
                                     guidance_scale=7.5, extra=extra), 8, "C2 v1 8x512^2 (UNet batch 16)", 16 * FL[64])
if "c4" in which:
    lat, emb, extra = inputs(2, 128)
    timed(FusedDenoiser(unet9), dict(latents=lat, prompt_embeds=emb, timesteps=sched.timesteps, coef=coef,
                                     guidance_scale=7.5, extra=extra), 2, "C4 v1 2x1024^2 per GPU (UNet batch 4)", 4 * FL[128])
if "c5" in which:
    cn = ControlNetModel.synthetic(NetConfig(in_channels=4), seed=77).to(dev)
    lat, emb, extra = inputs(2, 64)
    ctrl = torch.rand(4, 3, 512, 512, device=dev)
    timed(FusedDenoiser(unet9, cn, "controlnet"),
          dict(latents=lat, prompt_embeds=emb, side_prompt_embeds=emb, control_image=ctrl, timesteps=sched.timesteps,
               coef=coef, guidance_scale=7.5, extra=extra, side_scale=0.5), 2,
          "C5 v1+ControlNet 2x512^2 per GPU (batch 4 + 4)", 4 * (FL[64] + 0.2686e12))
if "c3" in which:
    unet4 = UNet2DConditionModel.synthetic(NetConfig(in_channels=4)).to(dev)
    bn = BrushNetModel.synthetic(NetConfig(in_channels=4), seed=99).to(dev)
    lat, emb, extra = inputs(4, 64)
    timed(FusedDenoiser(unet4, bn, "brushnet"),
          dict(latents=lat, prompt_embeds=emb, side_prompt_embeds=emb, timesteps=sched.timesteps, coef=coef,
               guidance_scale=7.5, extra=torch.cat([extra, extra]), side_scale=1.0), 4,
          "C3 v2 BrushNet 4x512^2 per GPU (batch 8 + 8)", 8 * (FL[64] + 0.8262e12))
