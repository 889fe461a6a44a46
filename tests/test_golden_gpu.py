"""GPU parity against the REFERENCE's own code: the CUDA path (nets and pipeline `__call__`s on the tiny config) against
the committed fixtures tests/golden/unet_composition.npz and pipeline_*_call.npz, which hold outputs of the reference's
own model / pipeline files run unmodified in fp32 on the CPU (generators: tests/golden/make_unet_golden.py,
make_pipeline_golden.py). No oracle call is involved in what is compared here except the VAE module handed to the
pipelines (the same on both sides).

Tolerances are the bf16-vs-fp32 ones of tests/test_nets_gpu.py (single forward, rel-L2 <= 3e-2 on these randomly
initialised tiny nets) and tests/test_pipelines_gpu.py (trajectories: rel-L2 <= 5e-2, cosine >= 0.998)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
BOC, HEADS, CROSS, GROUPS = (32, 64, 128, 128), 4, 64, 8
DEV = "cuda"
NET_REL = 3e-2


@pytest.fixture(scope="module", autouse=True)
def _fp32_exact():
    """the torch VAE handed to the pipelines runs in true fp32 like the fixture's"""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def _gold(name):
    return np.load(os.path.join(HERE, "golden", name))


def _cfg(cin):
    from powerpaint_b200.engine import NetConfig

    return NetConfig(in_channels=cin, block_out_channels=BOC, attention_head_dim=HEADS, cross_attention_dim=CROSS,
                     norm_num_groups=GROUPS)


def _model(cls, cin, kind, seed):
    from powerpaint_b200.models import synthetic_state_dict

    return cls.from_state_dict(_cfg(cin), synthetic_state_dict(_cfg(cin), kind, seed)).to(DEV)


def _inputs(seed, cin, h, w):  # == make_unet_golden.inputs (CPU generator, then moved)
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(2, cin, h, w, generator=g).to(DEV), torch.randn(2, 77, CROSS, generator=g).to(DEV),
            torch.randn(2, 5, h, w, generator=g).to(DEV))


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def _cos(a, b):
    return torch.nn.functional.cosine_similarity(a.float().flatten(), b.float().flatten(), dim=0).item()


def _report(test, **kw):
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", "golden_parity_report.jsonl"), "a") as f:
        f.write(json.dumps({"test": test, **kw}) + "\n")


def _close(got, key, gold, what):
    ref = torch.from_numpy(gold[key]).to(DEV)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    assert torch.isfinite(got.float()).all(), what
    r = _rel(got, ref)
    _report(what, key=key, rel=r)
    assert r <= NET_REL, f"{what}: rel-L2 {r:.3e} vs the reference's own forward"


@pytest.mark.parametrize("tag,h,w", [("8x8", 8, 8), ("10x12", 10, 12)])
def test_unet_forward_cuda_vs_reference_forward(tag, h, w):
    """ref:powerpaint/models/unet_2d_condition.py:1040-1363 (10x12: odd sizes inside, the `upsample_size` path)"""
    from powerpaint_b200.models import UNet2DConditionModel

    pm = _model(UNet2DConditionModel, 9, "unet", 1234)
    x, ctx, _ = _inputs(11, 9, h, w)
    _close(pm(x, 321, ctx).sample, f"unet9_{tag}", _gold("unet_composition.npz"), f"unet9 {tag}")


def test_brushnet_and_unet_adds_cuda_vs_reference_forward():
    """ref:powerpaint/models/BrushNet_CA.py:690-952 (28 outputs) and the 28 add points inside the reference UNet"""
    from powerpaint_b200.models import BrushNetModel, UNet2DConditionModel

    gold = _gold("unet_composition.npz")
    pb = _model(BrushNetModel, 4, "brushnet", 77)
    pu = _model(UNet2DConditionModel, 4, "unet", 1234)
    x, ctx, cond = _inputs(13, 4, 8, 8)
    d, m, u = pb(x, 500, ctx, cond, 0.8, return_dict=False)
    assert len(d) == 12 and len(u) == 15
    for i, t in enumerate(list(d) + [m] + list(u)):
        _close(t, f"brushnet_{i:02d}", gold, f"brushnet output {i}")
    # the REFERENCE's residuals into the product UNet: the injection points alone
    refs = [torch.from_numpy(gold[f"brushnet_{i:02d}"]).to(DEV) for i in range(28)]
    dl, ul = refs[:12], refs[13:]
    got = pu(x, 500, ctx, down_block_add_samples=dl, mid_block_add_sample=refs[12], up_block_add_samples=ul).sample
    assert len(dl) == 0 and len(ul) == 0
    _close(got, "unet4_with_adds", gold, "unet4 with the reference's adds")
    _close(pu(x, 500, ctx).sample, "unet4_plain", gold, "unet4 plain")


def test_unet_controlnet_residuals_cuda_vs_reference_forward():
    """ref:powerpaint/models/unet_2d_condition.py:1263-1272, :1296-1297"""
    from powerpaint_b200.models import UNet2DConditionModel

    gold = _gold("unet_composition.npz")
    pm = _model(UNet2DConditionModel, 9, "unet", 1234)
    shapes = [gold[f"brushnet_{i:02d}"].shape for i in range(13)]
    g = torch.Generator().manual_seed(17)
    dres = tuple((torch.randn(s, generator=g) * 0.1).to(DEV) for s in shapes[:12])
    mres = (torch.randn(shapes[12], generator=g) * 0.1).to(DEV)
    x9, ctx9, _ = _inputs(19, 9, 8, 8)
    got = pm(x9, 500, ctx9, down_block_additional_residuals=dres, mid_block_additional_residual=mres).sample
    _close(got, "unet9_controlnet_residuals", gold, "unet9 + ControlNet residuals")


def _traj(out, key, gold, what):
    ref = torch.from_numpy(gold[key]).to(out.device)
    assert out.shape == ref.shape and torch.isfinite(out).all(), what
    r, c = _rel(out, ref), _cos(out, ref)
    _report(what, key=key, rel=r, cos=c)
    assert r < 5e-2 and c > 0.998, (what, r, c)


@pytest.mark.parametrize("name", ["gpu_v1", "gpu_v1_strength"])
def test_v1_call_cuda_vs_reference_call(name):
    """the public `__call__` on the GPU (host tensors in, latents out) against the final latents of the reference's own
    `StableDiffusionInpaintPipeline.__call__` (ref:pipeline_PowerPaint.py:855-1071)"""
    import pipeline_cases as pc
    from oracle.vae import AutoencoderKLOracle
    from powerpaint_b200.models import UNet2DConditionModel
    from powerpaint_b200.pipelines import StableDiffusionInpaintPipeline
    from powerpaint_b200.schedulers import DDIMScheduler

    case = {"gpu_v1": pc.GPU_V1, "gpu_v1_strength": pc.GPU_V1_STRENGTH}[name]
    pipe = StableDiffusionInpaintPipeline(vae=AutoencoderKLOracle.synthetic(tiny=True).to(DEV), text_encoder=None,
                                          tokenizer=None, unet=_model(UNet2DConditionModel, 9, "unet", 77),
                                          scheduler=DDIMScheduler(), safety_checker=None)
    s = case["size"]
    img, mask, pe, ne, _ = pc.sized_inputs(2, s, s, CROSS, case["seed"])
    out = pipe(image=img, mask=mask, prompt_embeds=pe, negative_prompt_embeds=ne, height=s, width=s,
               generator=torch.Generator().manual_seed(case["gen_seed"]), output_type="latent", return_dict=False,
               **case["kw"])[0]
    _traj(out, f"{name}_latents", _gold("pipeline_v1_call.npz"), f"v1 __call__ {name}")


def test_controlnet_call_cuda_vs_reference_call():
    """ref:pipeline_PowerPaint_ControlNet.py:1349-1770 (reference UNet behind the fixture; control-guidance window)"""
    import pipeline_cases as pc
    from oracle.vae import AutoencoderKLOracle
    from powerpaint_b200.models import ControlNetModel, UNet2DConditionModel
    from powerpaint_b200.pipelines import StableDiffusionControlNetInpaintPipeline
    from powerpaint_b200.schedulers import DDIMScheduler

    case = pc.GPU_CONTROLNET
    pipe = StableDiffusionControlNetInpaintPipeline(
        vae=AutoencoderKLOracle.synthetic(tiny=True).to(DEV), text_encoder=None, tokenizer=None,
        unet=_model(UNet2DConditionModel, 9, "unet", 5), controlnet=_model(ControlNetModel, 4, "controlnet", 6),
        scheduler=DDIMScheduler(), safety_checker=None)
    s = case["size"]
    img, mask, pe, ne, ctl = pc.sized_inputs(2, s, s, CROSS, case["seed"])
    out = pipe(image=img, mask=mask, control_image=ctl, prompt_embeds=pe, negative_prompt_embeds=ne, height=s, width=s,
               generator=torch.Generator().manual_seed(case["gen_seed"]), output_type="latent", return_dict=False,
               **case["kw"])[0]
    _traj(out, "gpu_controlnet_latents", _gold("pipeline_controlnet_call.npz"), "controlnet __call__")
