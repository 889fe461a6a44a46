"""powerpaint_b200 — B200-native (sm_100a) implementation of PowerPaint's denoising hot path.

Public surface mirrors the reference package layout (powerpaint/{models,pipelines,utils}):
  powerpaint_b200.models     UNet2DConditionModel, BrushNetModel, ControlNetModel
  powerpaint_b200.pipelines  StableDiffusionInpaintPipeline, StableDiffusionPowerPaintBrushNetPipeline,
                             StableDiffusionControlNetInpaintPipeline
  powerpaint_b200.utils      TokenizerWrapper, EmbeddingLayerWithFixes, add_tokens
  powerpaint_b200.schedulers DDIMScheduler
The CUDA kernels live in csrc/ behind the C ABI of include/powerpaint_b200.h
(libpowerpaint_b200.so, built by `python -m powerpaint_b200.build`).
"""
__version__ = "0.1.0"
