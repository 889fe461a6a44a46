"""VaeImageProcessor: only what `__call__` uses when it stops at latents / tensors (pipeline_PowerPaint.py:258, :1064)"""
import torch


class VaeImageProcessor:
    def __init__(self, vae_scale_factor=8, do_convert_rgb=False, do_normalize=True, **kw):
        self.vae_scale_factor = vae_scale_factor
        self.do_normalize = do_normalize

    def preprocess(self, image, height=None, width=None):
        """TENSOR inputs only (4-D, float). diffusers 0.27's rule for tensors, restated: size defaults to the tensor's
        own (rounded down to a multiple of vae_scale_factor) and resizing is `interpolate(size=...)` — the identity at
        that size, which is all the goldens use; `[0, 1]` tensors are normalised to `[-1, 1]` when `do_normalize`, a
        tensor that already has negative values is passed through (with a deprecation warning upstream)"""
        assert torch.is_tensor(image) and image.dim() == 4 and image.is_floating_point()
        h = (height or image.shape[2]) // self.vae_scale_factor * self.vae_scale_factor
        w = (width or image.shape[3]) // self.vae_scale_factor * self.vae_scale_factor
        assert tuple(image.shape[-2:]) == (h, w), "the shim does not resize"
        if self.do_normalize and image.min() >= 0:
            image = 2.0 * image - 1.0
        return image

    @staticmethod
    def denormalize(images):
        return (images / 2 + 0.5).clamp(0, 1)

    def postprocess(self, image, output_type="pil", do_denormalize=None):
        if output_type == "latent":
            return image
        if output_type != "pt":
            raise NotImplementedError("the golden generator only asks for 'latent' or 'pt'")
        if do_denormalize is None:
            do_denormalize = [True] * image.shape[0]
        return torch.stack([self.denormalize(image[i]) if do_denormalize[i] else image[i]
                            for i in range(image.shape[0])])


PipelineImageInput = object  # annotation only
