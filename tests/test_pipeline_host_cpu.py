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
    def run(self, *, latents, prompt_embeds, timesteps, coef, guidance_scale, extra, noise_fn=None, callback=None):
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
    from powerpaint_b200.models.autoencoder_kl import AutoencoderKL
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
