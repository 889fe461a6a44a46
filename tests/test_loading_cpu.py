"""CPU tests of the checkpoint-directory side of the drop-in (ref:app.py:91-200): `from_pretrained` of the models and the
pipelines on directories in the diffusers layout, written here BY HAND in the files' own formats (config.json with the
keys of the published SD-1.5 configs, safetensors / .bin weights, transformers' text encoder + tokenizer folders, a PNDM
scheduler config like runwayml/stable-diffusion-inpainting ships) — then the assembled pipeline must reproduce the
reference's own `__call__` fixture."""
import json
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
BOC, HEADS, CROSS, GROUPS = (32, 64, 128, 128), 4, 64, 8

# the keys of runwayml/stable-diffusion-inpainting's unet/config.json (values shrunk to the tiny test config)
UNET_CONFIG = {
    "_class_name": "UNet2DConditionModel", "_diffusers_version": "0.6.0", "act_fn": "silu", "attention_head_dim": HEADS,
    "block_out_channels": list(BOC), "center_input_sample": False, "cross_attention_dim": CROSS,
    "down_block_types": ["CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"],
    "downsample_padding": 1, "flip_sin_to_cos": True, "freq_shift": 0, "in_channels": 9, "layers_per_block": 2,
    "mid_block_scale_factor": 1, "norm_eps": 1e-05, "norm_num_groups": GROUPS, "out_channels": 4, "sample_size": 8,
    "up_block_types": ["UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"]}
VAE_CONFIG = {
    "_class_name": "AutoencoderKL", "_diffusers_version": "0.6.0", "act_fn": "silu", "block_out_channels": [16, 32, 32, 32],
    "down_block_types": ["DownEncoderBlock2D"] * 4, "in_channels": 3, "latent_channels": 4, "layers_per_block": 1,
    "norm_num_groups": 8, "out_channels": 3, "sample_size": 64, "up_block_types": ["UpDecoderBlock2D"] * 4}
PNDM_CONFIG = {
    "_class_name": "PNDMScheduler", "_diffusers_version": "0.6.0", "beta_end": 0.012, "beta_schedule": "scaled_linear",
    "beta_start": 0.00085, "num_train_timesteps": 1000, "set_alpha_to_one": False, "skip_prk_steps": True,
    "steps_offset": 1, "trained_betas": None, "clip_sample": False}
MODEL_INDEX = {
    "_class_name": "StableDiffusionInpaintPipeline", "_diffusers_version": "0.6.0",
    "feature_extractor": ["transformers", "CLIPImageProcessor"],
    "safety_checker": ["stable_diffusion", "StableDiffusionSafetyChecker"], "scheduler": ["diffusers", "PNDMScheduler"],
    "text_encoder": ["transformers", "CLIPTextModel"], "tokenizer": ["transformers", "CLIPTokenizer"],
    "unet": ["diffusers", "UNet2DConditionModel"], "vae": ["diffusers", "AutoencoderKL"]}


def _cfg(cin):
    from powerpaint_b200.engine import NetConfig

    return NetConfig(in_channels=cin, block_out_channels=BOC, attention_head_dim=HEADS, cross_attention_dim=CROSS,
                     norm_num_groups=GROUPS)


def _write(directory, config, state_dict, config_name="config.json", weights="diffusion_pytorch_model.safetensors"):
    from safetensors.torch import save_file

    os.makedirs(directory, exist_ok=True)
    with open(os.path.join(directory, config_name), "w") as f:
        json.dump(config, f)
    if state_dict is not None:
        sd = {k: v.detach().contiguous() for k, v in state_dict.items()}
        if weights.endswith(".bin"):
            torch.save(sd, os.path.join(directory, weights))
        else:
            save_file(sd, os.path.join(directory, weights))


def _pipeline_directory(root):
    """a v1 checkpoint like the app loads (ref:app.py:91): fp16 UNet weights, a VAE with the DEPRECATED attention names
    old SD checkpoints carry, transformers folders, PNDM scheduler, a safety-checker entry that must be skipped"""
    from oracle.vae import AutoencoderKLOracle
    from powerpaint_b200.models import synthetic_state_dict
    from synthetic_clip import make_text_encoder, make_tokenizer

    with open(os.path.join(root, "model_index.json"), "w") as f:
        json.dump(MODEL_INDEX, f)
    _write(os.path.join(root, "unet"), UNET_CONFIG, synthetic_state_dict(_cfg(9), "unet", 77))
    vae_sd = {}
    for k, v in AutoencoderKLOracle.synthetic(tiny=True).state_dict().items():
        for new, old in ((".to_q.", ".query."), (".to_k.", ".key."), (".to_v.", ".value."), (".to_out.0.", ".proj_attn.")):
            k = k.replace(new, old)
        vae_sd[k] = v
    assert any(".query." in k for k in vae_sd)
    _write(os.path.join(root, "vae"), VAE_CONFIG, vae_sd)
    _write(os.path.join(root, "scheduler"), PNDM_CONFIG, None, config_name="scheduler_config.json")
    tok = make_tokenizer()
    tok.save_pretrained(os.path.join(root, "tokenizer"))
    make_text_encoder(len(tok), hidden=CROSS, seed=1).save_pretrained(os.path.join(root, "text_encoder"))


def test_pipeline_from_pretrained_assembles_a_checkpoint_directory_and_reproduces_the_reference_call(tmp_path):
    import test_pipeline_golden as G
    from oracle.unet import UNet2DConditionOracle
    from oracle.vae import AutoencoderKLOracle
    from powerpaint_b200.models import AutoencoderKL, UNet2DConditionModel
    from powerpaint_b200.pipelines import StableDiffusionInpaintPipeline as Pipeline
    from powerpaint_b200.schedulers import DDIMScheduler, UniPCMultistepScheduler
    from powerpaint_b200.utils import TokenizerWrapper
    from test_pipeline_host_cpu import _CoefficientDenoiser

    root = str(tmp_path / "stable-diffusion-inpainting")
    os.makedirs(root)
    _pipeline_directory(root)
    with pytest.warns(UserWarning, match="PNDMScheduler is replaced by DDIMScheduler"):
        half = Pipeline.from_pretrained(root, torch_dtype=torch.float16, local_files_only=True)
    assert half.unet.dtype == torch.float16 and next(half.text_encoder.parameters()).dtype == torch.float16
    with pytest.warns(UserWarning):
        pipe = Pipeline.from_pretrained(root, local_files_only=True)  # fp32 text encoder for the exact comparison below
    assert isinstance(pipe.unet, UNet2DConditionModel) and pipe.unet.dtype == torch.float32
    assert pipe.unet.config.in_channels == 9 and pipe.unet.config.sample_size == 8
    assert isinstance(pipe.vae, AutoencoderKL) and pipe.vae_scale_factor == 8 and pipe.safety_checker is None
    assert isinstance(pipe.scheduler, DDIMScheduler) and pipe.scheduler.config.steps_offset == 1
    assert pipe.scheduler.config.beta_schedule == "scaled_linear" and not pipe.scheduler.config.set_alpha_to_one
    # the app's next lines (ref:app.py:94-112,197): tokenizer wrapper from the same folder, scheduler swap by config
    pipe.tokenizer = TokenizerWrapper(from_pretrained=root, subfolder="tokenizer", revision=None, local_files_only=True)
    assert UniPCMultistepScheduler.from_config(pipe.scheduler.config).config.solver_order == 2
    # loaded weights == written weights (VAE: through the deprecated-name conversion)
    om, sd = G._oracle(UNet2DConditionOracle, 9, "unet", 77)
    assert all(torch.equal(pipe.unet.state_dict()[k].float(), v) for k, v in sd.items())
    ovae = AutoencoderKLOracle.synthetic(tiny=True)
    assert all(torch.equal(pipe.vae.state_dict()[k].float(), v) for k, v in ovae.state_dict().items())
    # run it: the kernel-backed VAE / fused denoiser cannot run on the CPU, so the same weights go through the fp32
    # stand-ins (attributes are plain and settable like upstream); the text encoder / tokenizer are the loaded ones.
    # The result is the reference's own `__call__` with string prompts (fixture `prompts_latents`).
    pipe.vae = ovae
    fake = _CoefficientDenoiser(om)
    pipe.denoiser = lambda: fake
    img, mask, _, _ = G._call_inputs()
    with torch.no_grad():
        out = pipe(promptA=G.PROMPTS["promptA"], promptB=G.PROMPTS["promptB"], tradoff=0.7, tradoff_nag=0.4,
                   negative_promptA=G.PROMPTS["negative_promptA"], negative_promptB=G.PROMPTS["negative_promptB"],
                   image=img, mask=mask, height=G.H, width=G.W, num_inference_steps=3, guidance_scale=7.5,
                   generator=G._generators(False), output_type="latent", return_dict=False)[0]
    ref = torch.from_numpy(np.load(os.path.join(HERE, "golden", "pipeline_v1_call.npz"))["prompts_latents"])
    assert G._rel(out, ref) < 5e-5, G._rel(out, ref)


def test_model_from_pretrained_variants_and_refusals(tmp_path):
    from powerpaint_b200.models import BrushNetModel, ControlNetModel, UNet2DConditionModel, synthetic_state_dict

    # .bin weights + a variant name + subfolder, as `UNet2DConditionModel.from_pretrained(path, subfolder="unet")`
    root = str(tmp_path / "ckpt")
    sd = synthetic_state_dict(_cfg(4), "unet", 3)
    _write(os.path.join(root, "unet"), {**UNET_CONFIG, "in_channels": 4}, {k: v.half() for k, v in sd.items()},
           weights="diffusion_pytorch_model.fp16.bin")
    m = UNet2DConditionModel.from_pretrained(root, subfolder="unet", revision=None, torch_dtype=torch.float16,
                                             variant="fp16", local_files_only=True)
    assert all(torch.equal(m.state_dict()[k].float(), v.half().float()) for k, v in sd.items())
    with pytest.raises(EnvironmentError):
        UNet2DConditionModel.from_pretrained(root, subfolder="unet")  # no non-variant weights there
    with pytest.raises(EnvironmentError):
        UNet2DConditionModel.from_pretrained(str(tmp_path / "missing"), local_files_only=True)
    # the ControlNet config of lllyasviel/sd-controlnet-canny (ref:app.py:121-123), the BrushNet one, save_pretrained
    cn_cfg = {k: v for k, v in UNET_CONFIG.items() if k not in ("up_block_types", "out_channels", "sample_size",
                                                                  "center_input_sample")}
    cn_cfg.update({"_class_name": "ControlNetModel", "in_channels": 4, "class_embed_type": None,
                   "conditioning_embedding_out_channels": [16, 32, 96, 256],
                   "controlnet_conditioning_channel_order": "rgb", "num_class_embeds": None,
                   "only_cross_attention": False, "projection_class_embeddings_input_dim": None,
                   "resnet_time_scale_shift": "default", "upcast_attention": False, "use_linear_projection": False})
    _write(os.path.join(root, "canny"), cn_cfg, synthetic_state_dict(_cfg(4), "controlnet", 6))
    cn = ControlNetModel.from_pretrained(os.path.join(root, "canny"), torch_dtype=torch.float16)
    assert cn.config.in_channels == 4 and tuple(cn.config.conditioning_embedding_out_channels) == (16, 32, 96, 256)
    bn = BrushNetModel.from_unet(m)
    bn.save_pretrained(os.path.join(root, "brushnet"))
    bn2 = BrushNetModel.from_pretrained(os.path.join(root, "brushnet"))
    assert bn2.config.conditioning_channels == 5
    assert all(torch.equal(a, b) for a, b in zip(bn.state_dict().values(), bn2.state_dict().values()))
    # options outside the SD-1.5 family are refused by name, not half-loaded
    for key, value in (("use_linear_projection", True), ("attention_head_dim", [5, 10, 20, 20]),
                       ("class_embed_type", "timestep"), ("act_fn", "gelu")):
        _write(os.path.join(root, "bad"), {**UNET_CONFIG, key: value}, None)
        with pytest.raises(NotImplementedError, match=key):
            UNet2DConditionModel.from_pretrained(os.path.join(root, "bad"))


def test_brushnet_pipeline_from_pretrained_with_passed_components(tmp_path):
    """ref:app.py:141-171: UNet by subfolder, BrushNet from it, the pipeline from a base-model directory with
    `brushnet=`, `text_encoder_brushnet=`, `safety_checker=None` passed in, then `pipe.unet` re-assigned"""
    from powerpaint_b200.models import BrushNetModel, UNet2DConditionModel, synthetic_state_dict
    from powerpaint_b200.pipelines import StableDiffusionPowerPaintBrushNetPipeline
    from synthetic_clip import make_text_encoder, make_tokenizer

    root = str(tmp_path / "realisticVision")
    os.makedirs(root)
    _pipeline_directory(root)
    _write(os.path.join(root, "unet"), {**UNET_CONFIG, "in_channels": 4}, synthetic_state_dict(_cfg(4), "unet", 11))
    unet = UNet2DConditionModel.from_pretrained(root, subfolder="unet", revision=None, torch_dtype=torch.float16,
                                                local_files_only=True)
    brushnet = BrushNetModel.from_unet(unet)
    te_b = make_text_encoder(len(make_tokenizer()), hidden=CROSS, seed=2)
    with pytest.warns(UserWarning, match="replaced by DDIMScheduler"):
        pipe = StableDiffusionPowerPaintBrushNetPipeline.from_pretrained(
            root, brushnet=brushnet, text_encoder_brushnet=te_b, torch_dtype=torch.float16, low_cpu_mem_usage=False,
            safety_checker=None)
    assert pipe.brushnet is brushnet and pipe.text_encoder_brushnet is te_b and pipe.unet.config.in_channels == 4
    pipe.unet = unet  # plain settable attribute (ref:app.py:165)
    with pytest.raises(ValueError, match="expected brushnet"):
        StableDiffusionPowerPaintBrushNetPipeline.from_pretrained(root, text_encoder_brushnet=te_b)
    with pytest.raises(TypeError, match="unexpected keyword"):
        StableDiffusionPowerPaintBrushNetPipeline.from_pretrained(root, brushnet=brushnet, text_encoder_brushnet=te_b,
                                                                  nonsense=1)
