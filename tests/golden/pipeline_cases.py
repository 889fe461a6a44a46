"""Invalid v1 calls shared by the golden generator (which records what the REFERENCE raises) and the test (which holds
the product to it): name -> kwargs."""
import numpy as np
import torch


def error_cases(img, mask, pe, ne, H, W):
    """name -> kwargs overriding a valid v1 call"""
    base = dict(image=img, mask=mask, prompt_embeds=pe, negative_prompt_embeds=ne, height=H, width=W,
                num_inference_steps=2, guidance_scale=7.5, output_type="latent", return_dict=False)

    def c(**kw):
        return {**base, **kw}

    return {
        "strength_range": c(strength=1.5),
        "height_not_multiple_of_8": c(height=60),
        "callback_steps_zero": c(callback_steps=0),
        "prompt_and_embeds": c(promptA="a cat", promptB="a cat"),
        "no_prompt": c(prompt_embeds=None, negative_prompt_embeds=None),
        "prompt_type": c(promptA=3, promptB=3, prompt_embeds=None, negative_prompt_embeds=None),
        "negative_prompt_and_embeds": c(negative_promptA="x", negative_promptB="x"),
        "embeds_shape_mismatch": c(negative_prompt_embeds=ne[:1]),
        "image_none": c(image=None),
        "mask_none": c(mask=None),
        "mask_not_tensor": c(mask=np.zeros((H, W), dtype=np.float32)),
        "image_range": c(image=img * 3),
        "mask_range": c(mask=mask + 1.5),
        "size_mismatch": c(mask=mask[..., :-8]),
        "batch_mismatch": c(mask=torch.cat([mask, mask[:1]])),
        "too_few_steps": c(strength=0.1, num_inference_steps=2),
        "generator_list_length": c(generator=[torch.Generator().manual_seed(0)] * 3),
    }


def sized_inputs(batch, h, w, cross, seed):
    """image in [-1, 1], two-rectangle mask, prompt / negative embeddings and a control image for an h x w request
    (the fixtures the GPU tests use: latent sizes the tiny GPU tests already exercise)"""
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(batch, 3, h, w, generator=g) * 2 - 1
    mask = torch.zeros(batch, 1, h, w)
    mask[0, :, h // 8:5 * h // 8, w // 4:5 * w // 8] = 1
    mask[1:, :, h // 4:, w // 16:w // 2] = 1
    pe = torch.randn(batch, 77, cross, generator=g) * 0.5
    ne = torch.randn(batch, 77, cross, generator=g) * 0.5
    ctl = torch.rand(batch, 3, h, w, generator=g)
    return img, mask, pe, ne, ctl


GPU_V1 = dict(size=128, seed=51, gen_seed=11, kw=dict(num_inference_steps=6, guidance_scale=7.5))
GPU_V1_STRENGTH = dict(size=128, seed=52, gen_seed=12, kw=dict(num_inference_steps=10, strength=0.6, guidance_scale=7.5))
GPU_CONTROLNET = dict(size=64, seed=53, gen_seed=13,
                      kw=dict(num_inference_steps=6, guidance_scale=7.5, controlnet_conditioning_scale=0.5,
                              control_guidance_start=0.0, control_guidance_end=0.7))


def brushnet_error_cases(img, mask, prompts, cross):
    """invalid v2 (BrushNet) calls: name -> kwargs"""
    base = dict(**prompts, tradoff=0.7, tradoff_nag=0.4, image=img, mask=mask, num_inference_steps=2,
                guidance_scale=7.5, output_type="latent", return_dict=False)
    pe = torch.randn(img.shape[0], 77, cross, generator=torch.Generator().manual_seed(3))

    def c(**kw):
        return {**base, **kw}

    return {
        "scale_int": c(brushnet_conditioning_scale=1),
        "scale_list": c(brushnet_conditioning_scale=[1.0]),
        "callback_steps_zero": c(callback_steps=0),
        "prompt_and_embeds": c(prompt_embeds=pe),
        "no_prompt": c(promptA=None, promptB=None),
        "prompt_type": c(promptA=3, promptB=3),
        "negative_prompt_and_embeds": c(negative_prompt_embeds=pe),
        "image_type": c(image="x"),
        "mask_type": c(mask=3.0),
        "image_batch_mismatch": c(image=torch.cat([img, img[:1]])),
        "guidance_start_ge_end": c(control_guidance_start=0.5, control_guidance_end=0.5),
        "guidance_start_negative": c(control_guidance_start=-0.1),
        "guidance_end_above_one": c(control_guidance_end=1.5),
        "guidance_length_mismatch": c(control_guidance_start=[0.0, 0.1], control_guidance_end=[1.0]),
        "callback_tensor_inputs": c(callback_on_step_end_tensor_inputs=["nope"]),
        "generator_list_length": c(generator=[torch.Generator().manual_seed(0)] * 3),
        "ip_adapter_both": c(ip_adapter_image=img, ip_adapter_image_embeds=[pe]),
    }


def controlnet_error_cases(img, mask, pe, ne, ctl, H, W):
    """invalid ControlNet calls: name -> kwargs (strength outside [0, 1] is NOT one: the reference does not check it)"""
    base = dict(image=img, mask=mask, control_image=ctl, prompt_embeds=pe, negative_prompt_embeds=ne, height=H, width=W,
                num_inference_steps=2, guidance_scale=7.5, output_type="latent", return_dict=False)

    def c(**kw):
        return {**base, **kw}

    return {
        "height_not_multiple_of_8": c(height=60),
        "callback_steps_zero": c(callback_steps=0),
        "prompt_and_embeds": c(promptA="a cat", promptB="a cat"),
        "no_prompt": c(prompt_embeds=None, negative_prompt_embeds=None),
        "negative_prompt_and_embeds": c(negative_promptA="x", negative_promptB="x"),
        "embeds_shape_mismatch": c(negative_prompt_embeds=ne[:1]),
        "scale_int": c(controlnet_conditioning_scale=1),
        "scale_list": c(controlnet_conditioning_scale=[0.5]),
        "control_image_type": c(control_image="x"),
        "control_image_none": c(control_image=None),
        "control_image_batch": c(control_image=torch.cat([ctl, ctl[:1]])),
        "guidance_start_ge_end": c(control_guidance_start=0.5, control_guidance_end=0.5),
        "guidance_start_negative": c(control_guidance_start=-0.1),
        "guidance_end_above_one": c(control_guidance_end=1.5),
        "image_none": c(image=None),
        "mask_range": c(mask=mask + 1.5),
        "generator_list_length": c(generator=[torch.Generator().manual_seed(0)] * 3),
    }


def prepare_input_kinds(H, W):
    """every input container `prepare_mask_and_masked_image` accepts (ref:pipeline_PowerPaint.py:39-153): PIL (resized
    with LANCZOS / NEAREST), lists, numpy HWC / HW, tensors of every rank, plus two rejected combinations"""
    import PIL.Image

    rng = np.random.default_rng(0)
    pil = PIL.Image.fromarray(rng.integers(0, 256, (100, 80, 3), dtype=np.uint8))
    pm = PIL.Image.fromarray((rng.random((100, 80)) > 0.5).astype(np.uint8) * 255).convert("L")
    npimg = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    npm = (rng.random((H, W)) > 0.5).astype(np.float32)
    g = torch.Generator().manual_seed(0)
    t3 = torch.rand(3, H, W, generator=g) * 2 - 1
    tm2 = (torch.rand(H, W, generator=g) > 0.5).float()
    return {
        "pil": (pil, pm),
        "pil_list": ([pil, pil.transpose(PIL.Image.FLIP_LEFT_RIGHT)], [pm, pm]),
        "np": (npimg, npm),
        "np_list": ([npimg, npimg[::-1].copy()], [npm, npm]),
        "np_batch4d": (np.stack([npimg, npimg]), np.stack([npm, npm])[:, None]),
        "tensor3d_mask2d": (t3, tm2),
        "tensor4d_mask3d": (t3[None], tm2[None]),
        "tensor_batch_mask_single": (torch.stack([t3, t3 * 0.5]), tm2[None, None]),
        "pil_image_np_mask": (pil.resize((W, H)), npm),
        "np_mask_255": (npimg, (npm * 255).astype(np.uint8)),
    }


def digest_prepare(fn, image, mask, H, W):
    """shapes + sha256 of the three outputs (bit-exact comparison without storing them), or the exception"""
    import hashlib

    try:
        outs = fn(image, mask, H, W, return_image=True)
    except Exception as e:  # noqa: BLE001
        return {"error": [type(e).__name__, str(e)]}
    return {"shapes": [list(o.shape) for o in outs],
            "sha256": [hashlib.sha256(o.contiguous().numpy().tobytes()).hexdigest() for o in outs]}
