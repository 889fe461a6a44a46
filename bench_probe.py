"""Developer probe (not the judged bench): time the fused denoising loop only, print per-step ms."""
import argparse
import time

import torch

from powerpaint_b200.denoise import FusedDenoiser
from powerpaint_b200.engine import NetConfig
from powerpaint_b200.models import BrushNetModel, UNet2DConditionModel
from powerpaint_b200.schedulers import DDIMScheduler

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--hw", type=int, default=64)
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--mode", default="v1")
ap.add_argument("--no-graph", action="store_true")
a = ap.parse_args()

torch.manual_seed(0)
dev = torch.device("cuda:0")
t0 = time.time()
if a.mode == "v1":
    unet = UNet2DConditionModel.synthetic(NetConfig(in_channels=9)).to(dev)
    den = FusedDenoiser(unet)
else:
    unet = UNet2DConditionModel.synthetic(NetConfig(in_channels=4)).to(dev)
    bn = BrushNetModel.synthetic(NetConfig(in_channels=4), seed=99).to(dev)
    den = FusedDenoiser(unet, bn, "brushnet")
print(f"model init {time.time() - t0:.1f}s")
B, h = a.batch, a.hw
sched = DDIMScheduler()
sched.set_timesteps(a.steps)
coef = sched.step_coefficients()
lat = torch.randn(B, 4, h, h, device=dev)
emb = torch.randn(2 * B, 77, 768, device=dev) * 0.5
extra = torch.randn(B, 5, h, h, device=dev)
kw = dict(latents=lat, prompt_embeds=emb, timesteps=sched.timesteps, coef=coef, guidance_scale=7.5, extra=extra,
          use_graph=not a.no_graph)
if a.mode != "v1":
    kw["side_prompt_embeds"] = emb
t0 = time.time()
out = den.run(**kw)
torch.cuda.synchronize()
print(f"first call (plan + graph capture) {time.time() - t0:.1f}s; launches/step {den.launches_per_step}; "
      f"finite {torch.isfinite(out).all().item()} std {out.std().item():.3f}")
for r in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    out = den.run(**kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"rep {r}: {ms:.1f} ms total, {ms / a.steps:.2f} ms/step, {B / (ms / 1e3):.2f} images/s (loop only)")
