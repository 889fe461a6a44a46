"""CPU tests: parameter layout of the product models equals the oracle's (diffusers names)."""
import torch

from oracle.unet import BrushNetOracle, ControlNetOracle, UNet2DConditionOracle, UNetConfig
from powerpaint_b200.engine import NetConfig
from powerpaint_b200.models.spec import param_shapes, synthetic_state_dict


def _cfgs(tiny: bool, in_channels: int):
    o = UNetConfig.tiny(in_channels) if tiny else UNetConfig.sd15(in_channels)
    n = NetConfig(in_channels=in_channels, block_out_channels=o.block_out_channels,
                  attention_head_dim=o.attention_head_dim, cross_attention_dim=o.cross_attention_dim,
                  norm_num_groups=o.norm_num_groups)
    return o, n


def _check(oracle_module, cfg, kind):
    want = {k: tuple(v.shape) for k, v in oracle_module.state_dict().items()}
    got = dict(param_shapes(cfg, kind))
    assert set(got) == set(want), (sorted(set(got) ^ set(want))[:10])
    for k in want:
        assert got[k] == want[k], (k, got[k], want[k])


def test_param_layout_tiny():
    o, n = _cfgs(True, 9)
    _check(UNet2DConditionOracle(o), n, "unet")
    o4, n4 = _cfgs(True, 4)
    _check(BrushNetOracle(o4), n4, "brushnet")
    _check(ControlNetOracle(o4), n4, "controlnet")


def test_param_layout_sd15_counts():
    """SD-1.5 UNet has 859.5 M parameters with 9 input channels (857 M + conv_in) — SURVEY.md §6"""
    _, n = _cfgs(False, 9)
    total = 0
    for shape in param_shapes(n, "unet").values():
        k = 1
        for s in shape:
            k *= s
        total += k
    assert 855e6 < total < 865e6, total
    _, n4 = _cfgs(False, 4)
    sd = param_shapes(n4, "brushnet")
    assert len([k for k in sd if k.startswith("brushnet_down_blocks") and k.endswith("weight")]) == 12
    assert len([k for k in sd if k.startswith("brushnet_up_blocks") and k.endswith("weight")]) == 15


def test_synthetic_state_dict_loads_into_oracle():
    o, n = _cfgs(True, 9)
    sd = synthetic_state_dict(n, "unet", seed=7)
    m = UNet2DConditionOracle(o)
    m.load_state_dict(sd, strict=True)
    sd2 = synthetic_state_dict(n, "unet", seed=7)
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)


def test_vae_param_layout_matches_torch_restatement():
    """the kernel-backed AutoencoderKL carries exactly the diffusers state-dict names / shapes of the torch
    restatement (so `vae/diffusion_pytorch_model.safetensors` loads into either)"""
    from oracle.vae import AutoencoderKLOracle
    from powerpaint_b200.models.autoencoder_kl import AutoencoderKL, vae_param_shapes

    for kw in (dict(), dict(block_out_channels=(16, 32, 32, 32), norm_num_groups=8, layers_per_block=1)):
        o = AutoencoderKLOracle(**kw)
        want = {k: tuple(v.shape) for k, v in o.state_dict().items()}
        got = dict(vae_param_shapes(block_out_channels=kw.get("block_out_channels", (128, 256, 512, 512)),
                                    layers_per_block=kw.get("layers_per_block", 2)))
        assert got == want, set(got) ^ set(want)
    p = AutoencoderKL.synthetic(tiny=True)
    o = AutoencoderKLOracle(block_out_channels=(16, 32, 32, 32), norm_num_groups=8, layers_per_block=1)
    o.load_state_dict(p.state_dict(), strict=True)
    import pytest
    import torch

    with pytest.raises(RuntimeError):
        p.encode(torch.zeros(1, 3, 64, 64))  # CPU parameters: there is no CPU path


def test_safetensors_load_model_like_the_app(tmp_path):
    """the app loads checkpoints with `safetensors.torch.load_model(pipe.unet | pipe.brushnet, path)`
    (ref:app.py:111,188-191): diffusers-named tensors in, nothing missing or unexpected, the recorded programs of the
    previous weights invalidated"""
    import pytest

    st = pytest.importorskip("safetensors.torch")
    from powerpaint_b200.models import BrushNetModel, UNet2DConditionModel

    for cls, kind, cin in ((UNet2DConditionModel, "unet", 9), (BrushNetModel, "brushnet", 4)):
        _, n = _cfgs(True, cin)
        sd = synthetic_state_dict(n, kind, 3)
        path = str(tmp_path / f"{kind}.safetensors")
        st.save_file({k: v.contiguous() for k, v in sd.items()}, path)
        m = cls(n)
        before = m.generation
        missing, unexpected = st.load_model(m, path)
        assert not missing and not unexpected
        assert m.generation > before
        got = m.state_dict()
        assert all(torch.equal(got[k].float(), v) for k, v in sd.items())


def test_every_kernel_waits_for_its_predecessor_grid():
    """With PP_B200_PDL=1 every launch carries the programmatic-stream-serialization attribute (off by default: measured
    slower), so a kernel may start while its predecessor drains; `griddepcontrol.wait` (pdl_wait) is then what orders
    its memory accesses after the predecessor's. The planner's
    buffer recycling (a block handed to a later writer once its last reader has been RECORDED) additionally needs the
    order to be transitive: kernel N + 1 waits for N only, so N itself must have waited for N - 1. Hence: every
    __global__ function calls pdl_wait(), and before anything that could return."""
    import glob
    import os
    import re

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "powerpaint_b200", "csrc")
    seen = 0
    for path in sorted(glob.glob(os.path.join(root, "*.cu")) + glob.glob(os.path.join(root, "*.cuh"))):
        with open(path) as f:
            src = f.read()
        for m in re.finditer(r"__global__[^;{]*?\(", src):
            i = src.find("{", m.end())
            depth, j = 0, i
            while True:
                depth += (src[j] == "{") - (src[j] == "}")
                if depth == 0:
                    break
                j += 1
            body = src[i:j]
            seen += 1
            k = body.find("pdl_wait()")
            assert k >= 0, f"{os.path.basename(path)}: a kernel near offset {m.start()} never calls pdl_wait()"
            head = re.sub(r"//[^\n]*", "", body[:k])
            while True:  # lambda bodies (address helpers) return values, not the kernel
                lm = re.search(r"\[[&=]?\]\s*\([^)]*\)\s*(?:->\s*[\w:<>*& ]+?)?\s*\{", head)
                if not lm:
                    break
                d, e = 0, lm.end() - 1
                while e < len(head):
                    d += (head[e] == "{") - (head[e] == "}")
                    if d == 0:
                        break
                    e += 1
                head = head[:lm.start()] + head[e + 1:]
            assert "return" not in head, \
                f"{os.path.basename(path)}: a kernel near offset {m.start()} can return before pdl_wait()"
    assert seen >= 20
