from . import StableDiffusionPipelineOutput  # noqa: F401
