#!/usr/bin/env python
"""bench.py — 512x512 50-step text-guided inpainting throughput of the PowerPaint-v1 hot path.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference [--gpus N ...]          CPU arm: the oracle port of the reference

Workload (BASELINE.json configs[1], "C2"): PowerPaint-v1 text-guided inpaint, batch 8 x 512x512 per
GPU, 50 DDIM steps, CFG 7.5 (UNet batch 16), bf16 storage / fp32 accumulate, synthetic seeded weights
(no checkpoint is reachable offline) and synthetic image + mask + prompt-embedding batches.
One bench "step" = one full 50-step denoise of the per-GPU batch.

  value  images/s, whole job (sum over GPUs), the denoising loop with inputs already resident in HBM
         (latents, mask/masked-image latents, prompt embeddings) — CUDA-event timed, max over ranks.
  e2e    the same metric through the reference-facing API: StableDiffusionInpaintPipeline.__call__
         with pinned HOST buffers (image, mask, prompt embeddings) -> ... -> decoded images read back to
         host; H2D / D2H, VAE encode/decode (torch library path, SURVEY.md §8f "next" row) included.
         With N > 1, rank 0 owns all inputs and NCCL scatters them / gathers the decoded images.
  roofline  tensor-bound: algorithmic FLOPs of one denoising step (12.85 TFLOP at C2)
         / mean device time of one recorded step program (one CUDA graph replay = all launches of the
         step), against the measured sustained bf16 peak of MEASURED_PEAKS.json.
  cpu_baseline  the fp32 oracle port of the same loop on the host cores (bounded sample).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "512x512 50-step inpaint images/sec"
UNIT = "images/s"
DDIM_STEPS = 50
BATCH = 8
LATENT = 64
GUIDANCE = 7.5
FLOP_PER_SAMPLE_FWD = 0.8034e12  # UNet forward per sample at 512^2


def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(tflops=float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1467.7))),
                    hbm=float(p.get("hbm_gbs", 6570.6)), src="MEASURED_PEAKS.json (bf16_tflops_sustained)")
    return dict(tflops=1400.0, hbm=6650.0, src="fallback of B200_PROFILING.md (sustained)")


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append((float(parts[0]), float(parts[1])))
                    for n, v in zip(names, parts[2:6]):
                        if v.lower().startswith("active"):
                            self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": sorted(self.reasons)}
        sm = sorted(s[0] for s in self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0][1], "reasons": sorted(self.reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------- CPU arm (oracle port)
_CPU_UNET = None
_CPU_THREADS = None


def best_cpu_threads():
    """torch's fp32 conv/GEMM throughput on a many-core host is not monotonic in the thread count
    (oversubscription, NUMA): probe a few counts on a representative conv and keep the fastest."""
    global _CPU_THREADS
    if _CPU_THREADS is not None:
        return _CPU_THREADS
    n = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, n) if c <= n})
    x = torch.randn(2, 320, 64, 64)
    w = torch.randn(320, 320, 3, 3)
    best, best_t = n, None
    for c in cands:
        torch.set_num_threads(c)
        torch.nn.functional.conv2d(x, w, padding=1)
        t0 = time.perf_counter()
        for _ in range(3):
            torch.nn.functional.conv2d(x, w, padding=1)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
    _CPU_THREADS = best
    return best


def cpu_oracle_rate(threads: int, ddim_steps_sample: int = 2, repeats: int = 1):
    """images/s of the fp32 oracle port of the v1 loop on the host cores, extrapolated from a bounded
    sample: 1 image (UNet batch 2 with CFG), `ddim_steps_sample` of the 50 steps."""
    from oracle.ddim import DDIMOracle
    from oracle.pipelines import loop_v1
    from oracle.unet import UNet2DConditionOracle, UNetConfig, init_synthetic_

    torch.set_num_threads(threads)
    torch.manual_seed(0)
    global _CPU_UNET
    if _CPU_UNET is None:
        _CPU_UNET = init_synthetic_(UNet2DConditionOracle(UNetConfig.sd15(9))).eval()
    unet = _CPU_UNET
    sched = DDIMOracle()
    sched.set_timesteps(DDIM_STEPS)
    sched.timesteps = sched.timesteps[:ddim_steps_sample]
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(1, 4, LATENT, LATENT, generator=g)
    emb = torch.randn(2, 77, 768, generator=g) * 0.5
    mask = (torch.rand(1, 1, LATENT, LATENT, generator=g) > 0.75).float()
    ml = torch.randn(1, 4, LATENT, LATENT, generator=g)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        loop_v1(unet, sched, lat, emb, mask, ml, GUIDANCE)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    per_ddim_step = best / ddim_steps_sample
    return 1.0 / (per_ddim_step * DDIM_STEPS), per_ddim_step


def run_reference_arm(args):
    rank, world, _ = _dist_env()
    if rank != 0:
        return 0
    threads = best_cpu_threads()
    vals = []
    for i in range(args.warmup + args.steps):
        v, per = cpu_oracle_rate(threads, ddim_steps_sample=1)
        if i >= args.warmup:
            vals.append((v, per))
    v = sum(x[0] for x in vals) / len(vals)
    per = sum(x[1] for x in vals) / len(vals)
    sample = "1 image x 512x512 (UNet batch 2, CFG), 1 of 50 DDIM steps per bench step, extrapolated x50"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per * DDIM_STEPS * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "PowerPaint-v1 text-guided inpaint 512x512, 50 DDIM steps, CFG 7.5 (C2 shape, "
                               "bounded CPU sample)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference's own diffusers pipeline cannot run here (diffusers==0.27.0 absent, no network); "
                "this times the fp32 oracle port of its loop (oracle/) on the host cores",
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------- GPU arm
def build_pipeline(dev, seed_offset=0):
    from powerpaint_b200.engine import NetConfig
    from powerpaint_b200.models import UNet2DConditionModel
    from powerpaint_b200.models.autoencoder_kl import AutoencoderKL
    from powerpaint_b200.pipelines import StableDiffusionInpaintPipeline
    from powerpaint_b200.schedulers import DDIMScheduler

    unet = UNet2DConditionModel.synthetic(NetConfig(in_channels=9), seed=1234).to(dev)
    vae = AutoencoderKL.synthetic(seed=4321).to(dev).to(torch.bfloat16)
    return StableDiffusionInpaintPipeline(vae=vae, text_encoder=None, tokenizer=None, unet=unet,
                                          scheduler=DDIMScheduler(), safety_checker=None)


def synth_inputs(B, seed):
    """host-side synthetic request batch: images in [-1,1], centred 25%-area rectangle masks,
    prompt embeddings ~ 0.5 N(0,1)"""
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(B, 3, 8 * LATENT, 8 * LATENT, generator=g) * 2 - 1
    mask = torch.zeros(B, 1, 8 * LATENT, 8 * LATENT)
    q = 8 * LATENT // 4
    mask[:, :, q:3 * q, q:3 * q] = 1.0
    pe = torch.randn(B, 77, 768, generator=g) * 0.5
    ne = torch.randn(B, 77, 768, generator=g) * 0.5
    return img, mask, pe, ne


def run_gpu_arm(args):
    rank, world, local = _dist_env()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            print(f"bench.py: --gpus {args.gpus} needs torchrun with {args.gpus} ranks", file=sys.stderr)
            return 2
    if not torch.cuda.is_available():
        print("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)",
              file=sys.stderr)
        return 2
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    from powerpaint_b200 import _native

    assert _native.lib().pp_device_supported() == 1, "bench needs an sm_100 (B200) device"
    pipe = build_pipeline(dev)
    den = pipe.denoiser()
    sched = pipe.scheduler
    sched.set_timesteps(DDIM_STEPS)
    coef = sched.step_coefficients()
    B = BATCH
    img, mask, pe, ne = synth_inputs(B, seed=rank)
    # ---- resident inputs for `value`
    g = torch.Generator(device=dev).manual_seed(rank)
    lat0 = torch.randn(B, 4, LATENT, LATENT, device=dev, generator=g)
    emb = torch.cat([ne, pe]).to(dev)
    extra = torch.cat([torch.nn.functional.interpolate(mask, size=(LATENT, LATENT)),
                       torch.randn(B, 4, LATENT, LATENT, generator=torch.Generator().manual_seed(7))], 1).to(dev)

    def loop_once():
        return den.run(latents=lat0, prompt_embeds=emb, timesteps=sched.timesteps, coef=coef,
                       guidance_scale=GUIDANCE, extra=extra)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1)):
        loop_once()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record()
        for _ in range(args.steps):
            out = loop_once()
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = B * world * args.steps / (ms_max / 1e3)
    ms_per_ddim = ms_max / args.steps / DDIM_STEPS
    launches_per_ddim = den.launches_per_step

    # ---- per-step program time for the roofline (events around each graph replay)
    st = next(iter(den._cache.values()))
    torch.cuda.synchronize()
    with torch.cuda.stream(den._stream):
        st["step_idx"].zero_()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(DDIM_STEPS + 1)]
        evs[0].record()
        for i in range(DDIM_STEPS):
            st["program"].launch()
            evs[i + 1].record()
    torch.cuda.synchronize()
    step_ms = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(5, DDIM_STEPS - 5))
    step_mean = sum(step_ms) / len(step_ms)
    flops_step = FLOP_PER_SAMPLE_FWD * 2 * B
    pk = _peaks()
    achieved = flops_step / (step_mean / 1e3) / 1e12

    # ---- e2e through the public pipeline API, host buffers
    img_h, mask_h = img.pin_memory(), mask.pin_memory()
    pe_h, ne_h = pe.pin_memory(), ne.pin_memory()
    if dist is not None and world > 1:
        all_in = [synth_inputs(B, seed=r) for r in range(world)] if rank == 0 else None

    def e2e_once():
        if dist is not None and world > 1:
            # rank 0 owns every request: H2D there, NCCL scatter, ..., NCCL gather, D2H on rank 0
            from powerpaint_b200.parallel import scatter_requests

            full = [torch.cat([a[k] for a in all_in]).pin_memory().to(dev, non_blocking=True) for k in range(4)] \
                if rank == 0 else None
            i_d, m_d, p_d, n_d = scatter_requests(
                full, [(B, 3, 8 * LATENT, 8 * LATENT), (B, 1, 8 * LATENT, 8 * LATENT), (B, 77, 768), (B, 77, 768)],
                [torch.float32] * 4, dev)
        else:
            i_d, m_d = img_h.to(dev, non_blocking=True), mask_h.to(dev, non_blocking=True)
            p_d, n_d = pe_h.to(dev, non_blocking=True), ne_h.to(dev, non_blocking=True)
        res = pipe(image=i_d, mask=m_d, prompt_embeds=p_d, negative_prompt_embeds=n_d, height=8 * LATENT,
                   width=8 * LATENT, num_inference_steps=DDIM_STEPS, guidance_scale=GUIDANCE,
                   generator=torch.Generator().manual_seed(rank), output_type="pt").images
        res = (res * 255).round().to(torch.uint8)
        if dist is not None and world > 1:
            from powerpaint_b200.parallel import gather_images

            allimg = gather_images(res)
            return allimg.cpu() if rank == 0 else None
        return res.cpu()

    e2e_once()
    barrier()
    k_e2e = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(k_e2e):
        r = e2e_once()
    barrier()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_value = B * world * k_e2e / float(tt.item())
    h2d = sum(x.numel() * x.element_size() for x in (img, mask, pe, ne)) * (world if world > 1 else 1)
    d2h = B * 3 * (8 * LATENT) ** 2 * (world if world > 1 else 1)

    line = None
    if rank == 0:
        cpu_v, cpu_per = cpu_oracle_rate(best_cpu_threads(), ddim_steps_sample=2) if world == 1 else (None, None)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 1), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "PowerPaint-v1 text-guided inpaint, batch 8 x 512x512 per GPU, 50 DDIM steps, "
                                   "CFG 7.5 (UNet batch 16), bf16, synthetic seeded weights (BASELINE.json configs[1])",
                       "per_gpu_batch": B, "global_batch": B * world, "ddim_steps": DDIM_STEPS,
                       "parallelism": f"batch-sharded x{world}, no collective inside the loop",
                       "l2": "working set per step (1.7 GB weights + activations) exceeds the 126 MB L2",
                       "unet_ms_per_ddim_step": step_mean},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "includes": "H2D of image/mask/prompt embeddings, VAE encode (torch), 50 fused steps, VAE decode "
                                "(torch), uint8 D2H" + (", NCCL scatter/gather via rank 0" if world > 1 else "")},
            "gpu_launches": int(launches_per_ddim * DDIM_STEPS * args.steps),
            "clocks": clocks.summary(),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                         "frac": achieved / pk["tflops"], "traffic": None, "peak_source": pk["src"] + ", of measured",
                         "kernel": "one denoising-step program (UNet forward + CFG/DDIM), "
                                   f"{launches_per_ddim} launches replayed as one CUDA graph",
                         "flops_per_launch": flops_step, "ms_per_launch": step_mean},
        }
        if cpu_v is not None:
            line["cpu_baseline"] = {"value": cpu_v, "unit": UNIT, "cores": best_cpu_threads(), "kind": "port",
                                    "host_cpus": os.cpu_count(),
                                    "sample": "fp32 oracle port, 1 image (UNet batch 2), 2 of 50 DDIM steps, "
                                              f"{cpu_per:.2f} s per DDIM step, extrapolated x50"}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
