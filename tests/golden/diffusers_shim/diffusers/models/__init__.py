"""`diffusers.models` names the reference PIPELINE file imports (pipeline_PowerPaint.py:26): the VAE is an adapter over
oracle/vae.py (its arithmetic is not what the pipeline golden pins), the other two are annotation-only."""
from oracle.vae import AutoencoderKLOracle as AutoencoderKL  # noqa: F401  (has the `config` fields :257, :659, :921 read)


class AsymmetricAutoencoderKL:
    def __init__(self, *a, **k):
        raise NotImplementedError("AsymmetricAutoencoderKL is outside the hot path")


class UNet2DConditionModel:
    def __init__(self, *a, **k):
        raise NotImplementedError("annotation only: the reference ships its own UNet2DConditionModel")


from oracle.unet import ControlNetOracle as _ControlNetOracle  # noqa: E402


class ControlNetModel(_ControlNetOracle):
    """oracle ControlNet behind the diffusers call signature (pipeline_PowerPaint_ControlNet.py:1686-1694)"""

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def config(self):
        from ..configuration_utils import FrozenDict

        return FrozenDict(global_pool_conditions=False)

    def forward(self, sample, timestep, encoder_hidden_states, controlnet_cond, conditioning_scale=1.0,
                guess_mode=False, return_dict=True, **kw):
        assert not guess_mode, "guess_mode is outside the hot path"
        down, mid = super().forward(sample, timestep, encoder_hidden_states, controlnet_cond, conditioning_scale)
        return down, mid
