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
