"""GPU parity of the fused denoising loops and of the pipeline `__call__`s against the fp32 oracle
loops (oracle/pipelines.py) on identical seeded weights / inputs.

Stated tolerance (bf16 storage vs fp32 oracle): final latents rel-L2 <= 5e-2 and
cosine >= 0.998 for the step counts used here. The synthetic randomly-initialised nets amplify
rounding noise strongly (a single forward already differs by ~1.2e-2 from fp32 for BOTH this
implementation and torch-bf16 eager), so long trajectories are compared over 10 steps for the
tiny config; the per-step kernels are pinned much tighter in test_kernels_gpu.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def _cos(a, b):
    return torch.nn.functional.cosine_similarity(a.float().flatten(), b.float().flatten(), dim=0).item()


@pytest.fixture(scope="module", autouse=True)
def _fp32_exact():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def _nets(in_channels, kinds, tiny=True):
    from oracle.unet import BrushNetOracle, ControlNetOracle, UNet2DConditionOracle, UNetConfig
    from powerpaint_b200.engine import NetConfig
    from powerpaint_b200.models import BrushNetModel, ControlNetModel, UNet2DConditionModel, synthetic_state_dict

    out = {}
    for kind, cin, seed in kinds:
        o = UNetConfig.tiny(cin) if tiny else UNetConfig.sd15(cin)
        n = NetConfig(in_channels=cin, block_out_channels=o.block_out_channels, attention_head_dim=o.attention_head_dim,
                      cross_attention_dim=o.cross_attention_dim, norm_num_groups=o.norm_num_groups)
        sd = synthetic_state_dict(n, kind, seed)
        ocls = {"unet": UNet2DConditionOracle, "brushnet": BrushNetOracle, "controlnet": ControlNetOracle}[kind]
        pcls = {"unet": UNet2DConditionModel, "brushnet": BrushNetModel, "controlnet": ControlNetModel}[kind]
        om = ocls(o)
        om.load_state_dict(sd)
        out[kind] = (om.to(DEV).eval(), pcls.from_state_dict(n, sd).to(DEV), o)
    return out


def _scheds(steps, total=50):
    from oracle.ddim import DDIMOracle
    from powerpaint_b200.schedulers import DDIMScheduler

    so, sp = DDIMOracle(), DDIMScheduler()
    so.set_timesteps(total)
    sp.set_timesteps(total)
    so.timesteps = so.timesteps[:steps]
    return so, sp, sp.timesteps[:steps]


def test_loop_v1_tiny_and_callback_and_eta():
    from oracle.pipelines import loop_v1
    from powerpaint_b200.denoise import FusedDenoiser

    nets = _nets(9, [("unet", 9, 1234)])
    om, pm, o = nets["unet"]
    g = torch.Generator(device=DEV).manual_seed(0)
    B, h, w = 2, 16, 8
    lat = torch.randn(B, 4, h, w, device=DEV, generator=g)
    emb = torch.randn(2 * B, 77, o.cross_attention_dim, device=DEV, generator=g) * 0.5
    mask = (torch.rand(B, 1, h, w, device=DEV, generator=g) > 0.5).float()
    ml = torch.randn(B, 4, h, w, device=DEV, generator=g)
    so, sp, ts = _scheds(10)
    rec = []
    ref = loop_v1(om, so, lat, emb, mask, ml, 7.5, record=rec)
    den = FusedDenoiser(pm)
    extra = torch.cat([mask, ml], 1)
    got = den.run(latents=lat, prompt_embeds=emb, timesteps=ts, coef=sp.step_coefficients(ts), guidance_scale=7.5,
                  extra=extra)
    assert _rel(got, ref) < 5e-2 and _cos(got, ref) > 0.998, (_rel(got, ref), _cos(got, ref))
    # graph replay == plain launches == per-step callback path (all observe the same trajectory)
    seen = []
    got_cb = den.run(latents=lat, prompt_embeds=emb, timesteps=ts, coef=sp.step_coefficients(ts), guidance_scale=7.5,
                     extra=extra, callback=lambda i, t, x: seen.append((i, int(t), x.clone())) or None)
    assert len(seen) == 10 and [s[1] for s in seen] == [int(t) for t in ts]
    assert _rel(got_cb, ref) < 5e-2
    assert _rel(seen[0][2], rec[0]) < 2e-2, "first step latents"
    # inputs are not mutated
    assert torch.equal(lat, torch.randn(B, 4, h, w, device=DEV, generator=torch.Generator(device=DEV).manual_seed(0)))
    # guidance_scale <= 1: no CFG, eps is the conditional prediction (B-row embeddings)
    ref1 = loop_v1(om, so, lat, emb[B:], mask, ml, 1.0)
    got1 = den.run(latents=lat, prompt_embeds=emb[B:], timesteps=ts, coef=sp.step_coefficients(ts),
                   guidance_scale=1.0, extra=extra)
    assert _rel(got1, ref1) < 5e-2
    # eta > 0 with supplied variance noise
    noises = [torch.randn(B, 4, h, w, device=DEV, generator=g) for _ in range(10)]
    ref_e = loop_v1(om, so, lat, emb, mask, ml, 7.5, eta=0.7, noise_fn=lambda i: noises[i])
    got_e = den.run(latents=lat, prompt_embeds=emb, timesteps=ts, coef=sp.step_coefficients(ts, eta=0.7),
                    guidance_scale=7.5, extra=extra, noise_fn=lambda i: noises[i])
    assert _rel(got_e, ref_e) < 5e-2


def test_loop_brushnet_tiny():
    from oracle.pipelines import loop_brushnet
    from powerpaint_b200.denoise import FusedDenoiser

    nets = _nets(4, [("unet", 4, 1234), ("brushnet", 4, 99)])
    (om_u, pm_u, o), (om_b, pm_b, _) = nets["unet"], nets["brushnet"]
    g = torch.Generator(device=DEV).manual_seed(1)
    B, h = 2, 8
    lat = torch.randn(B, 4, h, h, device=DEV, generator=g)
    emb_t = torch.randn(2 * B, 77, o.cross_attention_dim, device=DEV, generator=g) * 0.5
    emb_u = torch.randn(2 * B, 77, o.cross_attention_dim, device=DEV, generator=g) * 0.5
    cond = torch.randn(2 * B, 5, h, h, device=DEV, generator=g)  # one set per CFG half, like the reference
    so, sp, ts = _scheds(8)
    ref = loop_brushnet(om_u, om_b, so, lat, emb_t, emb_u, cond, 7.5, 0.9)
    got = FusedDenoiser(pm_u, pm_b, "brushnet").run(latents=lat, prompt_embeds=emb_u, side_prompt_embeds=emb_t,
                                                    timesteps=ts, coef=sp.step_coefficients(ts), guidance_scale=7.5,
                                                    extra=cond, side_scale=0.9)
    assert _rel(got, ref) < 5e-2 and _cos(got, ref) > 0.998, (_rel(got, ref), _cos(got, ref))


def test_loop_controlnet_tiny():
    from oracle.pipelines import loop_controlnet
    from powerpaint_b200.denoise import FusedDenoiser

    nets = _nets(9, [("unet", 9, 1234), ("controlnet", 4, 77)])
    (om_u, pm_u, o), (om_c, pm_c, _) = nets["unet"], nets["controlnet"]
    g = torch.Generator(device=DEV).manual_seed(2)
    B, h = 2, 8
    lat = torch.randn(B, 4, h, h, device=DEV, generator=g)
    emb = torch.randn(2 * B, 77, o.cross_attention_dim, device=DEV, generator=g) * 0.5
    mask = (torch.rand(B, 1, h, h, device=DEV, generator=g) > 0.5).float()
    ml = torch.randn(B, 4, h, h, device=DEV, generator=g)
    ctrl = torch.rand(B, 3, 8 * h, 8 * h, device=DEV, generator=g)
    ctrl2 = torch.cat([ctrl] * 2)
    so, sp, ts = _scheds(8)
    ref = loop_controlnet(om_u, om_c, so, lat, emb, mask, ml, ctrl2, 7.5, 0.5)
    got = FusedDenoiser(pm_u, pm_c, "controlnet").run(latents=lat, prompt_embeds=emb, side_prompt_embeds=emb,
                                                      control_image=ctrl2, timesteps=ts,
                                                      coef=sp.step_coefficients(ts), guidance_scale=7.5,
                                                      extra=torch.cat([mask, ml], 1), side_scale=0.5)
    assert _rel(got, ref) < 5e-2 and _cos(got, ref) > 0.998, (_rel(got, ref), _cos(got, ref))


def test_pipeline_v1_call_tiny():
    """the public `__call__`: host tensors in, latents out, against the oracle fed the same prepared tensors"""
    from oracle.pipelines import loop_v1
    from oracle.vae import AutoencoderKLOracle as AutoencoderKL
    from powerpaint_b200.pipelines import StableDiffusionInpaintPipeline
    from powerpaint_b200.pipelines.common import prepare_mask_and_masked_image, randn_tensor, vae_encode
    from powerpaint_b200.schedulers import DDIMScheduler

    nets = _nets(9, [("unet", 9, 1234)])
    om, pm, o = nets["unet"]
    vae = AutoencoderKL.synthetic(tiny=True).to(DEV)
    pipe = StableDiffusionInpaintPipeline(vae=vae, text_encoder=None, tokenizer=None, unet=pm,
                                          scheduler=DDIMScheduler(), safety_checker=None)
    B, H = 2, 128
    g = torch.Generator().manual_seed(3)
    img = torch.rand(B, 3, H, H, generator=g) * 2 - 1
    mask = torch.zeros(B, 1, H, H)
    mask[:, :, 32:96, 32:96] = 1
    pe = torch.randn(B, 77, o.cross_attention_dim, generator=g) * 0.5
    ne = torch.randn(B, 77, o.cross_attention_dim, generator=g) * 0.5
    steps = 6
    out = pipe(image=img, mask=mask, prompt_embeds=pe, negative_prompt_embeds=ne, height=H, width=H,
               num_inference_steps=steps, guidance_scale=7.5, generator=torch.Generator().manual_seed(11),
               output_type="latent", return_dict=False)[0]
    assert out.shape == (B, 4, H // 8, H // 8)
    # the same host-side preparation, then the oracle loop
    gen = torch.Generator().manual_seed(11)
    m, mi = prepare_mask_and_masked_image(img, mask, H, H)
    lat = randn_tensor((B, 4, H // 8, H // 8), generator=gen, device=DEV, dtype=torch.float32)
    m_l = torch.nn.functional.interpolate(m, size=(H // 8, H // 8)).to(DEV)
    ml = vae_encode(vae, mi.to(DEV), gen)
    from oracle.ddim import DDIMOracle

    so = DDIMOracle()
    so.set_timesteps(steps)
    ref = loop_v1(om, so, lat, torch.cat([ne, pe]).to(DEV), m_l, ml, 7.5)
    assert _rel(out, ref) < 5e-2 and _cos(out, ref) > 0.998, (_rel(out, ref), _cos(out, ref))
    # decoded output types
    res = pipe(image=img, mask=mask, prompt_embeds=pe, negative_prompt_embeds=ne, height=H, width=H,
               num_inference_steps=2, generator=torch.Generator().manual_seed(11), output_type="pil")
    assert len(res.images) == B and res.images[0].size == (H, H) and res.nsfw_content_detected is None
    # reference error behaviour
    with pytest.raises(ValueError):
        pipe(image=img, mask=mask, prompt_embeds=pe, negative_prompt_embeds=ne, height=H + 4, width=H)
    with pytest.raises(ValueError):
        pipe(image=img, mask=mask, height=H, width=H)


def test_pipeline_v1_strength_below_one_tiny():
    """strength < 1 (ref:pipeline_PowerPaint.py:713-720,916-941): the loop starts part-way down the schedule
    from the VAE-encoded image noised to the first kept timestep"""
    from oracle.ddim import DDIMOracle
    from oracle.pipelines import loop_v1
    from oracle.vae import AutoencoderKLOracle as AutoencoderKL
    from powerpaint_b200.pipelines import StableDiffusionInpaintPipeline
    from powerpaint_b200.pipelines.common import prepare_mask_and_masked_image, randn_tensor, vae_encode
    from powerpaint_b200.schedulers import DDIMScheduler

    nets = _nets(9, [("unet", 9, 1234)])
    om, pm, o = nets["unet"]
    vae = AutoencoderKL.synthetic(tiny=True).to(DEV)
    pipe = StableDiffusionInpaintPipeline(vae=vae, text_encoder=None, tokenizer=None, unet=pm,
                                          scheduler=DDIMScheduler(), safety_checker=None)
    B, H = 2, 128
    g = torch.Generator().manual_seed(5)
    img = torch.rand(B, 3, H, H, generator=g) * 2 - 1
    mask = torch.zeros(B, 1, H, H)
    mask[:, :, 16:80, 40:120] = 1
    pe = torch.randn(B, 77, o.cross_attention_dim, generator=g) * 0.5
    ne = torch.randn(B, 77, o.cross_attention_dim, generator=g) * 0.5
    steps, strength = 8, 0.5
    out = pipe(image=img, mask=mask, prompt_embeds=pe, negative_prompt_embeds=ne, height=H, width=H,
               strength=strength, num_inference_steps=steps, guidance_scale=7.5,
               generator=torch.Generator().manual_seed(21), output_type="latent", return_dict=False)[0]
    assert out.shape == (B, 4, H // 8, H // 8)
    # the same host-side preparation in the reference's order (image latents, noise, masked-image latents)
    gen = torch.Generator().manual_seed(21)
    m, mi, init = prepare_mask_and_masked_image(img, mask, H, H, return_image=True)
    so, sp = DDIMOracle(), DDIMScheduler()
    so.set_timesteps(steps)
    sp.set_timesteps(steps)
    t_start = steps - min(int(steps * strength), steps)
    assert t_start == 4
    so.timesteps = so.timesteps[t_start:]
    image_latents = vae_encode(vae, init.to(DEV), gen)
    noise = randn_tensor((B, 4, H // 8, H // 8), generator=gen, device=DEV, dtype=torch.float32)
    lat = sp.add_noise(image_latents, noise, sp.timesteps[t_start:t_start + 1].repeat(B))
    m_l = torch.nn.functional.interpolate(m, size=(H // 8, H // 8)).to(DEV)
    ml = vae_encode(vae, mi.to(DEV), gen)
    ref = loop_v1(om, so, lat, torch.cat([ne, pe]).to(DEV), m_l, ml, 7.5)
    assert _rel(out, ref) < 5e-2 and _cos(out, ref) > 0.998, (_rel(out, ref), _cos(out, ref))
    # strength so small that no step is left: reference error
    with pytest.raises(ValueError):
        pipe(image=img, mask=mask, prompt_embeds=pe, negative_prompt_embeds=ne, height=H, width=H, strength=0.05,
             num_inference_steps=steps)


def test_pipeline_brushnet_call_tiny():
    from oracle.ddim import DDIMOracle
    from oracle.pipelines import loop_brushnet
    from oracle.vae import AutoencoderKLOracle as AutoencoderKL
    from powerpaint_b200.pipelines import StableDiffusionPowerPaintBrushNetPipeline
    from powerpaint_b200.pipelines.common import preprocess_image, randn_tensor
    from powerpaint_b200.schedulers import DDIMScheduler

    nets = _nets(4, [("unet", 4, 1234), ("brushnet", 4, 99)])
    (om_u, pm_u, o), (om_b, pm_b, _) = nets["unet"], nets["brushnet"]
    vae = AutoencoderKL.synthetic(tiny=True).to(DEV)
    pipe = StableDiffusionPowerPaintBrushNetPipeline(vae=vae, text_encoder=None, text_encoder_brushnet=None,
                                                     tokenizer=None, unet=pm_u, brushnet=pm_b,
                                                     scheduler=DDIMScheduler(), safety_checker=None)
    B, H = 1, 64
    g = torch.Generator().manual_seed(5)
    img = torch.rand(B, 3, H, H, generator=g) * 2 - 1
    mask = torch.ones(B, 3, H, H)
    mask[:, :, 16:48, 16:48] = -1.0  # preprocessed mask: sum over channels < 0 -> 1
    pe = torch.randn(B, 77, o.cross_attention_dim, generator=g) * 0.5
    ne = torch.randn(B, 77, o.cross_attention_dim, generator=g) * 0.5
    peU = torch.randn(2 * B, 77, o.cross_attention_dim, generator=g) * 0.5
    steps = 5
    with pytest.raises(TypeError):
        pipe(image=img, mask=mask, prompt_embeds=pe, negative_prompt_embeds=ne, prompt_embedsU=peU,
             brushnet_conditioning_scale=1)
    torch.manual_seed(123)  # the conditioning latents use the GLOBAL RNG like the reference
    out = pipe(image=img, mask=mask, prompt_embeds=pe, negative_prompt_embeds=ne, prompt_embedsU=peU, height=H,
               width=H, num_inference_steps=steps, guidance_scale=7.5, brushnet_conditioning_scale=1.0,
               generator=torch.Generator().manual_seed(9), output_type="latent", return_dict=False)[0]
    # same preparation for the oracle
    image_t = torch.cat([img.to(DEV)] * 2)
    om_mask = torch.cat([mask.to(DEV)] * 2)
    original_mask = (om_mask.sum(1)[:, None] < 0).float()
    lat = randn_tensor((B, 4, H // 8, H // 8), generator=torch.Generator().manual_seed(9), device=DEV,
                       dtype=torch.float32)
    torch.manual_seed(123)
    cl = vae.encode(image_t).latent_dist.sample() * vae.config.scaling_factor
    cond = torch.cat([cl, torch.nn.functional.interpolate(original_mask, size=cl.shape[-2:])], 1)
    so = DDIMOracle()
    so.set_timesteps(steps)
    ref = loop_brushnet(om_u, om_b, so, lat, torch.cat([ne, pe]).to(DEV), peU.to(DEV), cond, 7.5, 1.0)
    assert _rel(out, ref) < 5e-2 and _cos(out, ref) > 0.998, (_rel(out, ref), _cos(out, ref))


def test_loop_v1_sd15_5steps():
    """SD-1.5-size UNet, one image x CFG, 5 DDIM steps at 512^2"""
    from oracle.pipelines import loop_v1
    from powerpaint_b200.denoise import FusedDenoiser

    nets = _nets(9, [("unet", 9, 1234)], tiny=False)
    om, pm, o = nets["unet"]
    g = torch.Generator(device=DEV).manual_seed(0)
    B, h = 1, 64
    lat = torch.randn(B, 4, h, h, device=DEV, generator=g)
    emb = torch.randn(2 * B, 77, 768, device=DEV, generator=g) * 0.5
    mask = (torch.rand(B, 1, h, h, device=DEV, generator=g) > 0.75).float()
    ml = torch.randn(B, 4, h, h, device=DEV, generator=g)
    so, sp, ts = _scheds(5)
    ref = loop_v1(om, so, lat, emb, mask, ml, 7.5)
    got = FusedDenoiser(pm).run(latents=lat, prompt_embeds=emb, timesteps=ts, coef=sp.step_coefficients(ts),
                                guidance_scale=7.5, extra=torch.cat([mask, ml], 1))
    print(f"sd15 5-step loop: rel-L2 {_rel(got, ref):.3e} cos {_cos(got, ref):.5f}")
    assert _rel(got, ref) < 5e-2 and _cos(got, ref) > 0.998


# --------------------------------------------------------------------------- SD-1.5-size configs (BASELINE.json)
def _bf16_copy(om):
    import copy

    return copy.deepcopy(om).to(torch.bfloat16)


def _report(name, **kv):
    """numbers quoted in BASELINE.md are written next to the test log (gpurun_out/ travels back)"""
    import json
    import os

    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", "parity_report.jsonl"), "a") as f:
        f.write(json.dumps(dict(test=name, **kv)) + "\n")
    print(name, kv)


def test_loop_brushnet_sd15_c3():
    """C3 (PowerPaint-v2-1 BrushNet), SD-1.5 size: one image x CFG (UNet batch 2 + BrushNet batch 2), 64x64
    latents, 5 DDIM steps, brushnet scale 1.0 with a control_guidance window that drops the last step
    (ref:pipeline_PowerPaint_Brushnet_CA.py:1384-1449, keep flags :1369-1376)"""
    from oracle.pipelines import loop_brushnet
    from powerpaint_b200.denoise import FusedDenoiser

    nets = _nets(4, [("unet", 4, 1234), ("brushnet", 4, 99)], tiny=False)
    (om_u, pm_u, o), (om_b, pm_b, _) = nets["unet"], nets["brushnet"]
    g = torch.Generator(device=DEV).manual_seed(1)
    B, h = 1, 64
    lat = torch.randn(B, 4, h, h, device=DEV, generator=g)
    emb_t = torch.randn(2 * B, 77, 768, device=DEV, generator=g) * 0.5
    emb_u = torch.randn(2 * B, 77, 768, device=DEV, generator=g) * 0.5
    cond = torch.randn(2 * B, 5, h, h, device=DEV, generator=g)
    so, sp, ts = _scheds(5)
    den = FusedDenoiser(pm_u, pm_b, "brushnet")
    for keep in (None, [1.0, 1.0, 1.0, 1.0, 0.0]):
        ref = loop_brushnet(om_u, om_b, so, lat, emb_t, emb_u, cond, 7.5, 1.0, keep=keep)
        got = den.run(latents=lat, prompt_embeds=emb_u, side_prompt_embeds=emb_t, timesteps=ts,
                      coef=sp.step_coefficients(ts), guidance_scale=7.5, extra=cond, side_scale=1.0, side_keep=keep)
        _report("loop_brushnet_sd15_c3", keep=keep, rel=_rel(got, ref), cos=_cos(got, ref))
        assert torch.isfinite(got).all()
        assert _rel(got, ref) < 5e-2 and _cos(got, ref) > 0.999, (_rel(got, ref), _cos(got, ref))
    assert len(den._cache) == 1, "one recorded plan serves every conditioning scale / keep window"


def test_loop_controlnet_sd15_c5():
    """C5 (v1 + ControlNet), SD-1.5 size: one image x CFG, 64x64 latents, 5 DDIM steps, scale 0.5
    (ref:pipeline_PowerPaint_ControlNet.py:1663-1735)"""
    from oracle.pipelines import loop_controlnet
    from powerpaint_b200.denoise import FusedDenoiser

    nets = _nets(9, [("unet", 9, 1234), ("controlnet", 4, 77)], tiny=False)
    (om_u, pm_u, o), (om_c, pm_c, _) = nets["unet"], nets["controlnet"]
    g = torch.Generator(device=DEV).manual_seed(2)
    B, h = 1, 64
    lat = torch.randn(B, 4, h, h, device=DEV, generator=g)
    emb = torch.randn(2 * B, 77, 768, device=DEV, generator=g) * 0.5
    mask = (torch.rand(B, 1, h, h, device=DEV, generator=g) > 0.5).float()
    ml = torch.randn(B, 4, h, h, device=DEV, generator=g)
    ctrl2 = torch.cat([torch.rand(B, 3, 8 * h, 8 * h, device=DEV, generator=g)] * 2)
    so, sp, ts = _scheds(5)
    ref = loop_controlnet(om_u, om_c, so, lat, emb, mask, ml, ctrl2, 7.5, 0.5)
    got = FusedDenoiser(pm_u, pm_c, "controlnet").run(latents=lat, prompt_embeds=emb, side_prompt_embeds=emb,
                                                      control_image=ctrl2, timesteps=ts,
                                                      coef=sp.step_coefficients(ts), guidance_scale=7.5,
                                                      extra=torch.cat([mask, ml], 1), side_scale=0.5)
    _report("loop_controlnet_sd15_c5", rel=_rel(got, ref), cos=_cos(got, ref))
    assert torch.isfinite(got).all()
    assert _rel(got, ref) < 5e-2 and _cos(got, ref) > 0.999, (_rel(got, ref), _cos(got, ref))


def test_pipeline_controlnet_call_tiny():
    """the ControlNet pipeline's public `__call__` on the GPU (host tensors in, latents out) against the oracle
    loop fed the same prepared tensors, including a control_guidance window"""
    from oracle.ddim import DDIMOracle
    from oracle.pipelines import loop_controlnet
    from oracle.vae import AutoencoderKLOracle as AutoencoderKL
    from powerpaint_b200.pipelines import StableDiffusionControlNetInpaintPipeline
    from powerpaint_b200.pipelines.common import prepare_mask_and_masked_image, randn_tensor, vae_encode
    from powerpaint_b200.schedulers import DDIMScheduler

    nets = _nets(9, [("unet", 9, 1234), ("controlnet", 4, 77)])
    (om_u, pm_u, o), (om_c, pm_c, _) = nets["unet"], nets["controlnet"]
    vae = AutoencoderKL.synthetic(tiny=True).to(DEV)
    pipe = StableDiffusionControlNetInpaintPipeline(vae=vae, text_encoder=None, tokenizer=None, unet=pm_u,
                                                    controlnet=pm_c, scheduler=DDIMScheduler(), safety_checker=None)
    B, H = 2, 64
    g = torch.Generator().manual_seed(8)
    img = torch.rand(B, 3, H, H, generator=g) * 2 - 1
    mask = torch.zeros(B, 1, H, H)
    mask[:, :, 16:48, 8:40] = 1
    ctrl = torch.rand(B, 3, H, H, generator=g)
    pe = torch.randn(B, 77, o.cross_attention_dim, generator=g) * 0.5
    ne = torch.randn(B, 77, o.cross_attention_dim, generator=g) * 0.5
    steps = 6
    out = pipe(image=img, mask=mask, control_image=ctrl, prompt_embeds=pe, negative_prompt_embeds=ne, height=H,
               width=H, num_inference_steps=steps, guidance_scale=7.5, controlnet_conditioning_scale=0.5,
               control_guidance_start=0.0, control_guidance_end=0.7,
               generator=torch.Generator().manual_seed(13), output_type="latent", return_dict=False)[0]
    assert out.shape == (B, 4, H // 8, H // 8)
    gen = torch.Generator().manual_seed(13)
    m, mi = prepare_mask_and_masked_image(img, mask, H, H)
    lat = randn_tensor((B, 4, H // 8, H // 8), generator=gen, device=DEV, dtype=torch.float32)
    m_l = torch.nn.functional.interpolate(m, size=(H // 8, H // 8)).to(DEV)
    ml = vae_encode(vae, mi.to(DEV), gen)
    so = DDIMOracle()
    so.set_timesteps(steps)
    keep = [1.0 - float(i / steps < 0.0 or (i + 1) / steps > 0.7) for i in range(steps)]
    assert keep == [1.0, 1.0, 1.0, 1.0, 0.0, 0.0]
    ref = loop_controlnet(om_u, om_c, so, lat, torch.cat([ne, pe]).to(DEV), m_l, ml, torch.cat([ctrl] * 2).to(DEV), 7.5,
                          0.5, keep=keep)
    assert _rel(out, ref) < 5e-2 and _cos(out, ref) > 0.998, (_rel(out, ref), _cos(out, ref))


def test_loop_v1_c4_1024():
    """C4 (v1 outpainting at 1024^2): UNet batch 2, 128x128 latents (16384-token d = 40 self-attention),
    2 DDIM steps (ref:pipeline_PowerPaint.py:988-1035)"""
    from oracle.pipelines import loop_v1
    from powerpaint_b200.denoise import FusedDenoiser

    nets = _nets(9, [("unet", 9, 1234)], tiny=False)
    om, pm, o = nets["unet"]
    g = torch.Generator(device=DEV).manual_seed(4)
    B, h = 1, 128
    lat = torch.randn(B, 4, h, h, device=DEV, generator=g)
    emb = torch.randn(2 * B, 77, 768, device=DEV, generator=g) * 0.5
    mask = torch.ones(B, 1, h, h, device=DEV)
    mask[:, :, 32:96, 32:96] = 0  # outpainting: keep the centre
    ml = torch.randn(B, 4, h, h, device=DEV, generator=g)
    so, sp, ts = _scheds(2)
    ref = loop_v1(om, so, lat, emb, mask, ml, 7.5)
    got = FusedDenoiser(pm).run(latents=lat, prompt_embeds=emb, timesteps=ts, coef=sp.step_coefficients(ts),
                                guidance_scale=7.5, extra=torch.cat([mask, ml], 1))
    _report("loop_v1_c4_1024", rel=_rel(got, ref), cos=_cos(got, ref))
    assert torch.isfinite(got).all()
    assert _rel(got, ref) < 5e-2 and _cos(got, ref) > 0.999, (_rel(got, ref), _cos(got, ref))


def test_loop_v1_sd15_50steps_trajectory():
    """The full 50-step DDIM trajectory of C2's per-image work (SD-1.5 size, one image x CFG): SURVEY §8d asks
    final latents rel-L2 <= 5e-2 and cosine >= 0.999 against the fp32 oracle. The same trajectory is also run
    through the oracle modules in torch bf16 eager (the reference's own library path) as the sanity bound."""
    from oracle.ddim import DDIMOracle
    from oracle.pipelines import loop_v1
    from powerpaint_b200.denoise import FusedDenoiser

    nets = _nets(9, [("unet", 9, 1234)], tiny=False)
    om, pm, o = nets["unet"]
    g = torch.Generator(device=DEV).manual_seed(0)
    B, h = 1, 64
    lat = torch.randn(B, 4, h, h, device=DEV, generator=g)
    emb = torch.randn(2 * B, 77, 768, device=DEV, generator=g) * 0.5
    mask = (torch.rand(B, 1, h, h, device=DEV, generator=g) > 0.75).float()
    ml = torch.randn(B, 4, h, h, device=DEV, generator=g)
    so, sp, ts = _scheds(50)
    ref = loop_v1(om, so, lat, emb, mask, ml, 7.5)
    got = FusedDenoiser(pm).run(latents=lat, prompt_embeds=emb, timesteps=ts, coef=sp.step_coefficients(ts),
                                guidance_scale=7.5, extra=torch.cat([mask, ml], 1))
    om16 = _bf16_copy(om)
    so16 = DDIMOracle()
    so16.set_timesteps(50)
    e16 = loop_v1(om16, so16, lat.to(torch.bfloat16), emb.to(torch.bfloat16), mask.to(torch.bfloat16),
                  ml.to(torch.bfloat16), 7.5).float()
    _report("loop_v1_sd15_50steps", rel=_rel(got, ref), cos=_cos(got, ref), rel_torch_bf16_eager=_rel(e16, ref),
            cos_torch_bf16_eager=_cos(e16, ref))
    assert torch.isfinite(got).all()
    assert _rel(got, ref) < 5e-2 and _cos(got, ref) > 0.999, (_rel(got, ref), _cos(got, ref))


def test_odd_latent_size_v1_640x856():
    """the reference's canonical call recipe resizes the short side to 640 and rounds to multiples of 8 PIXELS
    (ref:app.py:258-269,317-321), e.g. 640 x 856 -> an 80 x 107 latent: odd intermediate resolutions
    (107 -> 54 -> 27 -> 14) through the stride-2 convs and `upsample_size` (ref:unet_2d_condition.py:1120-1126,
    1311-1312)"""
    from oracle.pipelines import loop_v1
    from powerpaint_b200.denoise import FusedDenoiser

    nets = _nets(9, [("unet", 9, 1234)], tiny=False)
    om, pm, o = nets["unet"]
    g = torch.Generator(device=DEV).manual_seed(6)
    B, h, w = 1, 80, 107
    lat = torch.randn(B, 4, h, w, device=DEV, generator=g)
    emb = torch.randn(2 * B, 77, 768, device=DEV, generator=g) * 0.5
    mask = (torch.rand(B, 1, h, w, device=DEV, generator=g) > 0.6).float()
    ml = torch.randn(B, 4, h, w, device=DEV, generator=g)
    so, sp, ts = _scheds(3)
    ref = loop_v1(om, so, lat, emb, mask, ml, 7.5)
    got = FusedDenoiser(pm).run(latents=lat, prompt_embeds=emb, timesteps=ts, coef=sp.step_coefficients(ts),
                                guidance_scale=7.5, extra=torch.cat([mask, ml], 1))
    _report("odd_latent_v1_640x856", rel=_rel(got, ref), cos=_cos(got, ref))
    assert got.shape == ref.shape and torch.isfinite(got).all()
    assert _rel(got, ref) < 5e-2 and _cos(got, ref) > 0.999, (_rel(got, ref), _cos(got, ref))


# --------------------------------------------------------------------------- UniPC and the 4-channel blend path
def test_loop_brushnet_unipc_tiny():
    """the v2 app's scheduler (UniPCMultistepScheduler.from_config(pipe.scheduler.config), ref app.py:197): the
    fused CFG + UniPC step kernel against the stepwise oracle scheduler inside the oracle BrushNet loop"""
    from oracle.pipelines import loop_brushnet
    from oracle.unipc import UniPCOracle
    from powerpaint_b200.denoise import FusedDenoiser
    from powerpaint_b200.schedulers import DDIMScheduler, UniPCMultistepScheduler

    nets = _nets(4, [("unet", 4, 1234), ("brushnet", 4, 99)])
    (om_u, pm_u, o), (om_b, pm_b, _) = nets["unet"], nets["brushnet"]
    g = torch.Generator(device=DEV).manual_seed(3)
    B, h = 2, 8
    lat = torch.randn(B, 4, h, h, device=DEV, generator=g)
    emb_t = torch.randn(2 * B, 77, o.cross_attention_dim, device=DEV, generator=g) * 0.5
    emb_u = torch.randn(2 * B, 77, o.cross_attention_dim, device=DEV, generator=g) * 0.5
    cond = torch.randn(2 * B, 5, h, h, device=DEV, generator=g)
    steps = 10
    so = UniPCOracle()
    so.set_timesteps(steps)
    sp = UniPCMultistepScheduler.from_config(DDIMScheduler().config)
    sp.set_timesteps(steps)
    assert [int(t) for t in sp.timesteps] == [int(t) for t in so.timesteps]
    ref = loop_brushnet(om_u, om_b, so, lat, emb_t, emb_u, cond, 7.5, 1.0)
    got = FusedDenoiser(pm_u, pm_b, "brushnet").run(latents=lat, prompt_embeds=emb_u, side_prompt_embeds=emb_t,
                                                    timesteps=sp.timesteps, coef=sp.step_coefficients(),
                                                    ucoef=sp.unipc_coefficients(), guidance_scale=7.5, extra=cond,
                                                    side_scale=1.0)
    assert _rel(got, ref) < 5e-2 and _cos(got, ref) > 0.998, (_rel(got, ref), _cos(got, ref))
    # the eager scheduler.step (multistep state kept like diffusers) == the oracle on random eps
    so.set_timesteps(6)
    sp.set_timesteps(6)
    x_o = x_p = lat
    for t in so.timesteps:
        eps = torch.randn(B, 4, h, h, device=DEV, generator=g)
        x_o = so.step(eps, int(t), x_o)
        x_p = sp.step(eps, int(t), x_p).prev_sample
    assert _rel(x_p, x_o) < 1e-4, _rel(x_p, x_o)


def test_pipeline_v1_four_channel_unet_blend():
    """a 4-channel (non-inpainting) UNet in the v1 pipeline: after every step the known region is reset to the
    original latents noised to the next timestep (ref pipeline_PowerPaint.py:1025-1035, incl. the `[:1]` indexing)"""
    from oracle.ddim import DDIMOracle
    from powerpaint_b200.pipelines import StableDiffusionInpaintPipeline
    from powerpaint_b200.pipelines.common import prepare_mask_and_masked_image, randn_tensor, vae_encode
    from powerpaint_b200.schedulers import DDIMScheduler
    from oracle.vae import AutoencoderKLOracle

    nets = _nets(4, [("unet", 4, 1234)])
    om, pm, o = nets["unet"]
    vae = AutoencoderKLOracle.synthetic(tiny=True).to(DEV)
    pipe = StableDiffusionInpaintPipeline(vae=vae, text_encoder=None, tokenizer=None, unet=pm,
                                          scheduler=DDIMScheduler(), safety_checker=None)
    B, H = 2, 64
    g = torch.Generator().manual_seed(4)
    img = torch.rand(B, 3, H, H, generator=g) * 2 - 1
    mask = torch.zeros(B, 1, H, H)
    mask[:, :, 16:48, 24:56] = 1
    pe = torch.randn(B, 77, o.cross_attention_dim, generator=g) * 0.5
    ne = torch.randn(B, 77, o.cross_attention_dim, generator=g) * 0.5
    steps = 5
    out = pipe(image=img, mask=mask, prompt_embeds=pe, negative_prompt_embeds=ne, height=H, width=H,
               num_inference_steps=steps, guidance_scale=7.5, generator=torch.Generator().manual_seed(17),
               output_type="latent", return_dict=False)[0]
    # the reference's order of RNG use (:916-962): image latents, noise, masked-image latents
    gen = torch.Generator().manual_seed(17)
    m, mi, init = prepare_mask_and_masked_image(img, mask, H, H, return_image=True)
    image_latents = vae_encode(vae, init.to(DEV), gen)
    noise = randn_tensor((B, 4, H // 8, H // 8), generator=gen, device=DEV, dtype=torch.float32)
    m_l = torch.nn.functional.interpolate(m, size=(H // 8, H // 8)).to(DEV)
    so = DDIMOracle()
    so.set_timesteps(steps)
    emb = torch.cat([ne, pe]).to(DEV)
    lat = noise
    ts = [int(t) for t in so.timesteps]
    with torch.no_grad():
        for i, t in enumerate(ts):
            eps = om(torch.cat([lat] * 2), t, emb)
            u, c = eps.chunk(2)
            lat = so.step(u + 7.5 * (c - u), t, lat)
            proper = image_latents[:1]
            if i < len(ts) - 1:
                a = so.alphas_cumprod[ts[i + 1]].to(DEV)
                proper = a ** 0.5 * proper + (1 - a) ** 0.5 * noise
            lat = (1 - m_l[:1]) * proper + m_l[:1] * lat
    assert _rel(out, lat) < 5e-2 and _cos(out, lat) > 0.998, (_rel(out, lat), _cos(out, lat))
