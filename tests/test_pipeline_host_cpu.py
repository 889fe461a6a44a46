"""CPU test of the v1 pipeline's HOST side (no kernels): `__call__` prepares timesteps, initial latents, mask /
masked-image latents and the DDIM coefficient table, then hands them to the fused denoiser. Here the denoiser
is replaced by a stand-in that evaluates the oracle UNet and applies the coefficient rows exactly as
pp_cfg_ddim_step does, and the result must equal the oracle loop fed the same preparation — for the full
schedule and for strength < 1 (ref:pipeline_PowerPaint.py:713-720,905-941,988-1035)."""
import pytest
import torch

from oracle.ddim import DDIMOracle
from oracle.pipelines import loop_v1
from oracle.unet import UNet2DConditionOracle, UNetConfig


class _CoefficientDenoiser:
    """what FusedDenoiser.run computes, in fp32 torch on the CPU (test stand-in only)"""

    def __init__(self, unet):
        self.unet = unet
        self.calls = []

    @torch.no_grad()
    def run(self, *, latents, prompt_embeds, timesteps, coef, guidance_scale, extra, noise_fn=None, callback=None,
            ucoef=None, blend=None):
        assert ucoef is None and blend is None  # DDIM, 9-channel UNet
        self.calls.append((tuple(latents.shape), [int(t) for t in timesteps], tuple(coef.shape)))
        do_cfg = guidance_scale > 1.0
        for i, t in enumerate(timesteps):
            x = torch.cat([latents, extra], dim=1)
            x = torch.cat([x] * 2) if do_cfg else x
            eps = self.unet(x, int(t), prompt_embeds)
            if do_cfg:
                u, c = eps.chunk(2)
                eps = u + guidance_scale * (c - u)
            sa, s1a, sap, dirc, sigma = [float(v) for v in coef[i, :5]]
            latents = sap * ((latents - s1a * eps) / sa) + dirc * eps
            if noise_fn is not None:
                latents = latents + sigma * noise_fn(i)
        return latents


@pytest.mark.parametrize("strength,steps,kept", [(1.0, 4, 4), (0.5, 8, 4), (0.3, 10, 3)])
def test_v1_call_host_side_matches_oracle_loop(strength, steps, kept):
    from powerpaint_b200.engine import NetConfig
    from powerpaint_b200.models import UNet2DConditionModel, synthetic_state_dict
    from oracle.vae import AutoencoderKLOracle as AutoencoderKL
    from powerpaint_b200.pipelines import StableDiffusionInpaintPipeline
    from powerpaint_b200.pipelines.common import prepare_mask_and_masked_image, randn_tensor, vae_encode
    from powerpaint_b200.schedulers import DDIMScheduler

    torch.manual_seed(0)
    o = UNetConfig.tiny(9)
    n = NetConfig(in_channels=9, block_out_channels=o.block_out_channels, attention_head_dim=o.attention_head_dim,
                  cross_attention_dim=o.cross_attention_dim, norm_num_groups=o.norm_num_groups)
    sd = synthetic_state_dict(n, "unet", 77)
    om = UNet2DConditionOracle(o)
    om.load_state_dict(sd)
    om.eval()
    pm = UNet2DConditionModel.from_state_dict(n, sd)
    vae = AutoencoderKL.synthetic(tiny=True)
    pipe = StableDiffusionInpaintPipeline(vae=vae, text_encoder=None, tokenizer=None, unet=pm,
                                          scheduler=DDIMScheduler(), safety_checker=None)
    fake = _CoefficientDenoiser(om)
    pipe.denoiser = lambda: fake
    B, H = 2, 64
    g = torch.Generator().manual_seed(9)
    img = torch.rand(B, 3, H, H, generator=g) * 2 - 1
    mask = torch.zeros(B, 1, H, H)
    mask[:, :, 8:40, 16:56] = 1
    pe = torch.randn(B, 77, o.cross_attention_dim, generator=g) * 0.5
    ne = torch.randn(B, 77, o.cross_attention_dim, generator=g) * 0.5
    out = pipe(image=img, mask=mask, prompt_embeds=pe, negative_prompt_embeds=ne, height=H, width=H,
               strength=strength, num_inference_steps=steps, guidance_scale=7.5,
               generator=torch.Generator().manual_seed(4), output_type="latent", return_dict=False)[0]
    (shape, ts, cshape), = fake.calls
    assert shape == (B, 4, H // 8, H // 8) and len(ts) == kept and cshape == (kept, 8)
    # the same preparation in the reference's order, then the oracle loop over the kept timesteps
    gen = torch.Generator().manual_seed(4)
    m, mi, init = prepare_mask_and_masked_image(img, mask, H, H, return_image=True)
    so, sp = DDIMOracle(), DDIMScheduler()
    so.set_timesteps(steps)
    sp.set_timesteps(steps)
    t_start = steps - kept
    assert ts == sp.timesteps[t_start:].tolist()
    so.timesteps = so.timesteps[t_start:]
    if strength == 1.0:
        lat = randn_tensor((B, 4, H // 8, H // 8), generator=gen, device="cpu", dtype=torch.float32)
    else:
        image_latents = vae_encode(vae, init, gen)
        noise = randn_tensor((B, 4, H // 8, H // 8), generator=gen, device="cpu", dtype=torch.float32)
        lat = sp.add_noise(image_latents, noise, sp.timesteps[t_start:t_start + 1].repeat(B))
    m_l = torch.nn.functional.interpolate(m, size=(H // 8, H // 8))
    ml = vae_encode(vae, mi, gen)
    ref = loop_v1(om, so, lat, torch.cat([ne, pe]), m_l, ml, 7.5)
    assert torch.allclose(out, ref, atol=2e-4, rtol=1e-4), (out - ref).abs().max()


class _ControlNetCoefficientDenoiser:
    """stand-in for FusedDenoiser(mode="controlnet").run: oracle ControlNet + UNet, coefficient-row DDIM"""

    def __init__(self, unet, controlnet):
        self.unet, self.controlnet = unet, controlnet

    @torch.no_grad()
    def run(self, *, latents, prompt_embeds, side_prompt_embeds, control_image, timesteps, coef, guidance_scale,
            extra, side_scale, side_keep=None, noise_fn=None, callback=None, ucoef=None):
        do_cfg = guidance_scale > 1.0
        for i, t in enumerate(timesteps):
            x4 = torch.cat([latents] * 2) if do_cfg else latents
            ex = torch.cat([extra] * 2) if do_cfg else extra
            d, m = self.controlnet(x4, int(t), side_prompt_embeds, control_image,
                                   side_scale * (side_keep[i] if side_keep is not None else 1.0))
            eps = self.unet(torch.cat([x4, ex], dim=1), int(t), prompt_embeds, down_block_additional_residuals=d,
                            mid_block_additional_residual=m)
            if do_cfg:
                u, c = eps.chunk(2)
                eps = u + guidance_scale * (c - u)
            sa, s1a, sap, dirc, _ = [float(v) for v in coef[i, :5]]
            latents = sap * ((latents - s1a * eps) / sa) + dirc * eps
        return latents


@pytest.mark.parametrize("strength,steps,kept", [(1.0, 3, 3), (0.5, 6, 3)])
def test_controlnet_call_host_side_matches_oracle_loop(strength, steps, kept):
    from oracle.pipelines import loop_controlnet
    from oracle.unet import ControlNetOracle
    from powerpaint_b200.engine import NetConfig
    from powerpaint_b200.models import ControlNetModel, UNet2DConditionModel, synthetic_state_dict
    from oracle.vae import AutoencoderKLOracle as AutoencoderKL
    from powerpaint_b200.pipelines import StableDiffusionControlNetInpaintPipeline
    from powerpaint_b200.pipelines.common import prepare_mask_and_masked_image, randn_tensor, vae_encode
    from powerpaint_b200.schedulers import DDIMScheduler

    o9, o4 = UNetConfig.tiny(9), UNetConfig.tiny(4)

    def cfg(o, cin):
        return NetConfig(in_channels=cin, block_out_channels=o.block_out_channels,
                         attention_head_dim=o.attention_head_dim, cross_attention_dim=o.cross_attention_dim,
                         norm_num_groups=o.norm_num_groups)

    sd_u, sd_c = synthetic_state_dict(cfg(o9, 9), "unet", 5), synthetic_state_dict(cfg(o4, 4), "controlnet", 6)
    ou, oc = UNet2DConditionOracle(o9), ControlNetOracle(o4)
    ou.load_state_dict(sd_u)
    oc.load_state_dict(sd_c)
    ou.eval()
    oc.eval()
    vae = AutoencoderKL.synthetic(tiny=True)
    pipe = StableDiffusionControlNetInpaintPipeline(
        vae=vae, text_encoder=None, tokenizer=None, unet=UNet2DConditionModel.from_state_dict(cfg(o9, 9), sd_u),
        controlnet=ControlNetModel.from_state_dict(cfg(o4, 4), sd_c), scheduler=DDIMScheduler())
    fake = _ControlNetCoefficientDenoiser(ou, oc)
    pipe.denoiser = lambda: fake
    B, H = 1, 64
    g = torch.Generator().manual_seed(2)
    img = torch.rand(B, 3, H, H, generator=g) * 2 - 1
    ctl = torch.rand(B, 3, H, H, generator=g)
    mask = torch.zeros(B, 1, H, H)
    mask[:, :, 16:48, 8:40] = 1
    pe = torch.randn(B, 77, o9.cross_attention_dim, generator=g) * 0.5
    ne = torch.randn(B, 77, o9.cross_attention_dim, generator=g) * 0.5
    out = pipe(image=img, mask=mask, control_image=ctl, prompt_embeds=pe, negative_prompt_embeds=ne, height=H, width=H,
               strength=strength, num_inference_steps=steps, guidance_scale=5.0, controlnet_conditioning_scale=0.5,
               generator=torch.Generator().manual_seed(8), output_type="latent", return_dict=False)[0]
    gen = torch.Generator().manual_seed(8)
    m, mi, init = prepare_mask_and_masked_image(img, mask, H, H, return_image=True)
    so, sp = DDIMOracle(), DDIMScheduler()
    so.set_timesteps(steps)
    sp.set_timesteps(steps)
    t_start = steps - kept
    so.timesteps = so.timesteps[t_start:]
    if strength == 1.0:
        lat = randn_tensor((B, 4, H // 8, H // 8), generator=gen, device="cpu", dtype=torch.float32)
    else:
        image_latents = vae_encode(vae, init, gen)
        noise = randn_tensor((B, 4, H // 8, H // 8), generator=gen, device="cpu", dtype=torch.float32)
        lat = sp.add_noise(image_latents, noise, sp.timesteps[t_start:t_start + 1].repeat(B))
    m_l = torch.nn.functional.interpolate(m, size=(H // 8, H // 8))
    ml = vae_encode(vae, mi, gen)
    ref = loop_controlnet(ou, oc, so, lat, torch.cat([ne, pe]), m_l, ml, torch.cat([ctl] * 2), 5.0, 0.5)
    assert torch.allclose(out, ref, atol=2e-4, rtol=1e-4), (out - ref).abs().max()


class _BrushNetCoefficientDenoiser:
    """stand-in for FusedDenoiser(mode="brushnet").run: oracle BrushNet + UNet, coefficient-row DDIM"""

    def __init__(self, unet, brushnet):
        self.unet, self.brushnet = unet, brushnet

    @torch.no_grad()
    def run(self, *, latents, prompt_embeds, side_prompt_embeds, timesteps, coef, guidance_scale, extra, side_scale,
            side_keep=None, noise_fn=None, callback=None, ucoef=None):
        do_cfg = guidance_scale > 1.0
        for i, t in enumerate(timesteps):
            x = torch.cat([latents] * 2) if do_cfg else latents
            d, m, u = self.brushnet(x, int(t), side_prompt_embeds, extra,
                                    side_scale * (side_keep[i] if side_keep is not None else 1.0))
            eps = self.unet(x, int(t), prompt_embeds, down_block_add_samples=d, mid_block_add_sample=m,
                            up_block_add_samples=u)
            if do_cfg:
                a, c = eps.chunk(2)
                eps = a + guidance_scale * (c - a)
            sa, s1a, sap, dirc, _ = [float(v) for v in coef[i, :5]]
            latents = sap * ((latents - s1a * eps) / sa) + dirc * eps
        return latents


def test_brushnet_call_host_side_matches_oracle_loop():
    from oracle.pipelines import loop_brushnet
    from oracle.unet import BrushNetOracle
    from powerpaint_b200.engine import NetConfig
    from powerpaint_b200.models import BrushNetModel, UNet2DConditionModel, synthetic_state_dict
    from oracle.vae import AutoencoderKLOracle as AutoencoderKL
    from powerpaint_b200.pipelines import StableDiffusionPowerPaintBrushNetPipeline
    from powerpaint_b200.pipelines.common import randn_tensor
    from powerpaint_b200.schedulers import DDIMScheduler

    o = UNetConfig.tiny(4)
    n = NetConfig(in_channels=4, block_out_channels=o.block_out_channels, attention_head_dim=o.attention_head_dim,
                  cross_attention_dim=o.cross_attention_dim, norm_num_groups=o.norm_num_groups)
    sd_u, sd_b = synthetic_state_dict(n, "unet", 11), synthetic_state_dict(n, "brushnet", 12)
    ou, ob = UNet2DConditionOracle(o), BrushNetOracle(o)
    ou.load_state_dict(sd_u)
    ob.load_state_dict(sd_b)
    ou.eval()
    ob.eval()
    vae = AutoencoderKL.synthetic(tiny=True)
    pipe = StableDiffusionPowerPaintBrushNetPipeline(
        vae=vae, text_encoder=None, text_encoder_brushnet=None, tokenizer=None,
        unet=UNet2DConditionModel.from_state_dict(n, sd_u), brushnet=BrushNetModel.from_state_dict(n, sd_b),
        scheduler=DDIMScheduler(), safety_checker=None)
    fake = _BrushNetCoefficientDenoiser(ou, ob)
    pipe.denoiser = lambda: fake
    B, H, steps = 1, 64, 3
    g = torch.Generator().manual_seed(5)
    img = torch.rand(B, 3, H, H, generator=g) * 2 - 1
    mask = torch.ones(B, 3, H, H)
    mask[:, :, 16:48, 16:48] = -1.0  # preprocessed mask: sum over channels < 0 -> 1
    pe = torch.randn(B, 77, o.cross_attention_dim, generator=g) * 0.5
    ne = torch.randn(B, 77, o.cross_attention_dim, generator=g) * 0.5
    peU = torch.randn(2 * B, 77, o.cross_attention_dim, generator=g) * 0.5
    torch.manual_seed(123)  # the conditioning latents use the GLOBAL RNG like the reference
    out = pipe(image=img, mask=mask, prompt_embeds=pe, negative_prompt_embeds=ne, prompt_embedsU=peU, height=H,
               width=H, num_inference_steps=steps, guidance_scale=7.5, brushnet_conditioning_scale=1.0,
               generator=torch.Generator().manual_seed(9), output_type="latent", return_dict=False)[0]
    image_t = torch.cat([img] * 2)
    original_mask = (torch.cat([mask] * 2).sum(1)[:, None] < 0).float()
    lat = randn_tensor((B, 4, H // 8, H // 8), generator=torch.Generator().manual_seed(9), device="cpu",
                       dtype=torch.float32)
    torch.manual_seed(123)
    cl = vae.encode(image_t).latent_dist.sample() * vae.config.scaling_factor
    cond = torch.cat([cl, torch.nn.functional.interpolate(original_mask, size=cl.shape[-2:])], 1)
    so = DDIMOracle()
    so.set_timesteps(steps)
    ref = loop_brushnet(ou, ob, so, lat, torch.cat([ne, pe]), peU, cond, 7.5, 1.0)
    assert torch.allclose(out, ref, atol=2e-4, rtol=1e-4), (out - ref).abs().max()


def test_pipeline_to_moves_every_registered_module():
    """`pipe = pipe.to("cuda")` (ref:app.py:113,135,200) is DiffusionPipeline.to: vae, both text encoders, unet, brushnet /
    controlnet all move (checked with the meta device on the CPU), the pipeline returns itself, a dtype selects the dtype
    of the tensors the hot-path nets return"""
    from oracle.vae import AutoencoderKLOracle
    from powerpaint_b200.engine import NetConfig
    from powerpaint_b200.models import BrushNetModel, ControlNetModel, UNet2DConditionModel
    from powerpaint_b200.pipelines import (StableDiffusionControlNetInpaintPipeline,
                                           StableDiffusionPowerPaintBrushNetPipeline)
    from powerpaint_b200.schedulers import DDIMScheduler

    o = UNetConfig.tiny(4)

    def cfg(cin):
        return NetConfig(in_channels=cin, block_out_channels=o.block_out_channels, attention_head_dim=o.attention_head_dim,
                         cross_attention_dim=o.cross_attention_dim, norm_num_groups=o.norm_num_groups)

    te, te_b = torch.nn.Linear(4, 4), torch.nn.Linear(4, 4)  # stand-ins: any nn.Module is an injected dependency
    pb = StableDiffusionPowerPaintBrushNetPipeline(
        vae=AutoencoderKLOracle.synthetic(tiny=True), text_encoder=te, text_encoder_brushnet=te_b, tokenizer=None,
        unet=UNet2DConditionModel(cfg(4)), brushnet=BrushNetModel(cfg(4)), scheduler=DDIMScheduler(), safety_checker=None)
    assert pb.to("meta") is pb
    for name in ("vae", "text_encoder", "text_encoder_brushnet", "unet", "brushnet"):
        assert next(getattr(pb, name).parameters()).device.type == "meta", name
    pb.to(torch.float16)
    assert pb.unet._out_dtype == torch.float16 and pb.brushnet._out_dtype == torch.float16
    pc = StableDiffusionControlNetInpaintPipeline(
        vae=AutoencoderKLOracle.synthetic(tiny=True), text_encoder=te, tokenizer=None, unet=UNet2DConditionModel(cfg(9)),
        controlnet=ControlNetModel(cfg(4)), scheduler=DDIMScheduler())
    pc.to("meta")
    for name in ("vae", "text_encoder", "unet", "controlnet"):
        assert next(getattr(pc, name).parameters()).device.type == "meta", name


def test_bench_style_uint8_requests_through_all_three_calls():
    """the argument kinds bench.py's e2e leg hands to the three `__call__`s (uint8 NCHW images / masks / control images,
    float prompt embeddings, output_type="uint8") on the CPU stand-ins: uint8 NHWC images come back, and the uint8
    request gives the same latents as its float form (image / 127.5 - 1, mask / 255, control / 255)"""
    from oracle.unet import BrushNetOracle, ControlNetOracle
    from oracle.vae import AutoencoderKLOracle
    from powerpaint_b200.engine import NetConfig
    from powerpaint_b200.models import BrushNetModel, ControlNetModel, UNet2DConditionModel, synthetic_state_dict
    from powerpaint_b200.pipelines import (StableDiffusionControlNetInpaintPipeline, StableDiffusionInpaintPipeline,
                                           StableDiffusionPowerPaintBrushNetPipeline)
    from powerpaint_b200.schedulers import DDIMScheduler

    def pair(cls_o, cls_p, cin, kind, seed):
        o = UNetConfig.tiny(cin)
        n = NetConfig(in_channels=cin, block_out_channels=o.block_out_channels,
                      attention_head_dim=o.attention_head_dim, cross_attention_dim=o.cross_attention_dim,
                      norm_num_groups=o.norm_num_groups)
        sd = synthetic_state_dict(n, kind, seed)
        m = cls_o(o).eval()
        m.load_state_dict(sd)
        return m, cls_p.from_state_dict(n, sd), o

    vae = AutoencoderKLOracle.synthetic(tiny=True)
    B, H = 2, 64
    g = torch.Generator().manual_seed(1)
    img = torch.randint(0, 256, (B, 3, H, H), generator=g, dtype=torch.uint8)
    mask = torch.zeros(B, 1, H, H, dtype=torch.uint8)
    mask[:, :, 16:48, 16:48] = 255
    ctl = (torch.rand(B, 3, H, H, generator=g) > 0.9).to(torch.uint8) * 255
    ou9, pu9, o = pair(UNet2DConditionOracle, UNet2DConditionModel, 9, "unet", 3)
    ou4, pu4, _ = pair(UNet2DConditionOracle, UNet2DConditionModel, 4, "unet", 11)
    ob, pb, _ = pair(BrushNetOracle, BrushNetModel, 4, "brushnet", 12)
    oc, pc, _ = pair(ControlNetOracle, ControlNetModel, 4, "controlnet", 6)
    pe = torch.randn(B, 77, o.cross_attention_dim, generator=g) * 0.5
    ne = torch.randn(B, 77, o.cross_attention_dim, generator=g) * 0.5
    peU = torch.randn(2 * B, 77, o.cross_attention_dim, generator=g) * 0.5

    def common(u8: bool, out: str):
        return dict(image=img if u8 else img.float() / 127.5 - 1, mask=mask if u8 else mask.float() / 255,
                    prompt_embeds=pe, negative_prompt_embeds=ne, height=H, width=H, num_inference_steps=3,
                    guidance_scale=7.5, generator=torch.Generator().manual_seed(0), output_type=out)

    v1 = StableDiffusionInpaintPipeline(vae=vae, text_encoder=None, tokenizer=None, unet=pu9,
                                        scheduler=DDIMScheduler(), safety_checker=None)
    f1 = _CoefficientDenoiser(ou9)
    v1.denoiser = lambda: f1
    bn = StableDiffusionPowerPaintBrushNetPipeline(vae=vae, text_encoder=None, text_encoder_brushnet=None,
                                                   tokenizer=None, unet=pu4, brushnet=pb, scheduler=DDIMScheduler(),
                                                   safety_checker=None)
    f2 = _BrushNetCoefficientDenoiser(ou4, ob)
    bn.denoiser = lambda: f2
    cn = StableDiffusionControlNetInpaintPipeline(vae=vae, text_encoder=None, tokenizer=None, unet=pu9, controlnet=pc,
                                                  scheduler=DDIMScheduler(), safety_checker=None)
    f3 = _ControlNetCoefficientDenoiser(ou9, oc)
    cn.denoiser = lambda: f3
    calls = [
        (v1, lambda u8: {}),
        (bn, lambda u8: dict(prompt_embedsU=peU, brushnet_conditioning_scale=1.0)),
        (cn, lambda u8: dict(control_image=ctl if u8 else ctl.float() / 255, controlnet_conditioning_scale=0.5)),
    ]
    for pipe, extra in calls:
        torch.manual_seed(7)  # the BrushNet conditioning latents draw from the global RNG
        res = pipe(**common(True, "uint8"), **extra(True)).images
        assert res.dtype == torch.uint8 and res.shape == (B, H, H, 3), type(pipe).__name__
        torch.manual_seed(7)
        a = pipe(**common(True, "latent"), **extra(True)).images
        torch.manual_seed(7)
        b = pipe(**common(False, "latent"), **extra(False)).images
        assert torch.allclose(a, b, atol=1e-5, rtol=1e-5), (type(pipe).__name__, (a - b).abs().max())


def test_brushnet_call_custom_timesteps_raise_like_retrieve_timesteps():
    """ref:pipeline_PowerPaint_Brushnet_CA.py:114-122: a scheduler whose `set_timesteps` takes no `timesteps` (DDIM and
    UniPC of diffusers 0.27.0, and the ones here) -> ValueError, raised after the prompt / image preparation"""
    from oracle.unet import BrushNetOracle
    from oracle.vae import AutoencoderKLOracle
    from powerpaint_b200.engine import NetConfig
    from powerpaint_b200.models import BrushNetModel, UNet2DConditionModel, synthetic_state_dict
    from powerpaint_b200.pipelines import StableDiffusionPowerPaintBrushNetPipeline
    from powerpaint_b200.schedulers import DDIMScheduler, UniPCMultistepScheduler

    o = UNetConfig.tiny(4)
    n = NetConfig(in_channels=4, block_out_channels=o.block_out_channels, attention_head_dim=o.attention_head_dim,
                  cross_attention_dim=o.cross_attention_dim, norm_num_groups=o.norm_num_groups)
    for sched in (DDIMScheduler(), UniPCMultistepScheduler()):
        pipe = StableDiffusionPowerPaintBrushNetPipeline(
            vae=AutoencoderKLOracle.synthetic(tiny=True), text_encoder=None, text_encoder_brushnet=None, tokenizer=None,
            unet=UNet2DConditionModel.from_state_dict(n, synthetic_state_dict(n, "unet", 1)),
            brushnet=BrushNetModel.from_state_dict(n, synthetic_state_dict(n, "brushnet", 2)), scheduler=sched,
            safety_checker=None)
        pipe.denoiser = lambda: (_ for _ in ()).throw(AssertionError("the loop must not be reached"))
        img = torch.zeros(1, 3, 64, 64)
        pe = torch.zeros(1, 77, o.cross_attention_dim)
        with pytest.raises(ValueError, match="does not support custom timestep schedules"):
            pipe(image=img, mask=torch.ones(1, 3, 64, 64), prompt_embeds=pe, negative_prompt_embeds=pe,
                 prompt_embedsU=torch.zeros(2, 77, o.cross_attention_dim), timesteps=[900, 500, 100])


def test_brushnet_encode_prompt_clip_skip_and_the_clip_skip_property():
    """`encode_prompt(clip_skip=k)` (ref:pipeline_PowerPaint_Brushnet_CA.py:537-552): the hidden state k layers before
    the last, through the final LayerNorm, for the positive prompt only (the negative branch :596-610 takes the last
    hidden state); the property returns what `__call__` was given (:1006-1007,:1227)"""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from synthetic_clip import make_text_encoder, make_tokenizer

    from powerpaint_b200.pipelines import StableDiffusionPowerPaintBrushNetPipeline

    tok = make_tokenizer()
    te = make_text_encoder(len(tok), seed=3).eval()
    pipe = StableDiffusionPowerPaintBrushNetPipeline.__new__(StableDiffusionPowerPaintBrushNetPipeline)
    pipe.tokenizer, pipe.text_encoder = tok, te
    assert pipe.clip_skip is None
    prompts = ["a photo of a cat", "a"]
    with torch.no_grad():
        plain = pipe.encode_prompt(prompts, "cpu", 1, True)
        skipped = pipe.encode_prompt(prompts, "cpu", 1, True, clip_skip=1)
        ids = tok(prompts, padding="max_length", max_length=tok.model_max_length, truncation=True,
                  return_tensors="pt").input_ids
        out = te(ids, output_hidden_states=True)
        want = te.text_model.final_layer_norm(out.hidden_states[-2])
    assert torch.equal(skipped[:2], plain[:2])  # negative half: unchanged
    assert torch.allclose(skipped[2:], want, atol=1e-6) and not torch.allclose(skipped[2:], plain[2:], atol=1e-4)
    pipe._clip_skip = 2
    assert pipe.clip_skip == 2
