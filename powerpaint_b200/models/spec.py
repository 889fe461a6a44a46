"""Parameter layout (diffusers state-dict names and shapes) of the three nets
on the hot path, derived from the config alone. Used to (a) build parameter containers whose
`state_dict()` keys equal the reference checkpoints' (`unet/unet.safetensors`,
`PowerPaint_Brushnet/diffusion_pytorch_model.safetensors`), (b) create deterministic synthetic
weights when no checkpoint is available.

Layout sources: powerpaint/models/unet_2d_condition.py:256-479 (UNet ctor),
powerpaint/models/BrushNet_CA.py:223-228,330-376,446-454 (conv_in_condition + zero-convs),
diffusers ControlNetModel.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict

import torch

from ..engine import NetConfig


def _conv(d, name, cout, cin, k):
    d[name + ".weight"] = (cout, cin, k, k)
    d[name + ".bias"] = (cout,)


def _lin(d, name, cout, cin, bias=True):
    d[name + ".weight"] = (cout, cin)
    if bias:
        d[name + ".bias"] = (cout,)


def _norm(d, name, c):
    d[name + ".weight"] = (c,)
    d[name + ".bias"] = (c,)


def _resnet(d, p, cin, cout, temb):
    _norm(d, p + ".norm1", cin)
    _conv(d, p + ".conv1", cout, cin, 3)
    _lin(d, p + ".time_emb_proj", cout, temb)
    _norm(d, p + ".norm2", cout)
    _conv(d, p + ".conv2", cout, cout, 3)
    if cin != cout:
        _conv(d, p + ".conv_shortcut", cout, cin, 1)


def _transformer(d, p, c, cross):
    _norm(d, p + ".norm", c)
    _conv(d, p + ".proj_in", c, c, 1)
    b = p + ".transformer_blocks.0"
    _norm(d, b + ".norm1", c)
    for a, kv in (("attn1", c), ("attn2", cross)):
        _lin(d, f"{b}.{a}.to_q", c, c, bias=False)
        _lin(d, f"{b}.{a}.to_k", c, kv, bias=False)
        _lin(d, f"{b}.{a}.to_v", c, kv, bias=False)
        _lin(d, f"{b}.{a}.to_out.0", c, c)
        if a == "attn1":
            _norm(d, b + ".norm2", c)
    _norm(d, b + ".norm3", c)
    _lin(d, b + ".ff.net.0.proj", 8 * c, c)
    _lin(d, b + ".ff.net.2", c, 4 * c)
    _conv(d, p + ".proj_out", c, c, 1)


def param_shapes(cfg: NetConfig, kind: str = "unet") -> "OrderedDict[str, Tuple[int, ...]]":
    d: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    boc = cfg.block_out_channels
    temb = boc[0] * 4
    cross = cfg.cross_attention_dim
    if kind == "brushnet":
        _conv(d, "conv_in_condition", boc[0], cfg.in_channels + cfg.conditioning_channels, 3)
    else:
        _conv(d, "conv_in", boc[0], cfg.in_channels, 3)
    _lin(d, "time_embedding.linear_1", temb, boc[0])
    _lin(d, "time_embedding.linear_2", temb, temb)
    if kind == "controlnet":
        ce = cfg.conditioning_embedding_out_channels
        pre = "controlnet_cond_embedding"
        _conv(d, pre + ".conv_in", ce[0], cfg.controlnet_cond_channels, 3)
        for i in range(len(ce) - 1):
            _conv(d, f"{pre}.blocks.{2 * i}", ce[i], ce[i], 3)
            _conv(d, f"{pre}.blocks.{2 * i + 1}", ce[i + 1], ce[i], 3)
        _conv(d, pre + ".conv_out", boc[0], ce[-1], 3)
    out_c = boc[0]
    for i, t in enumerate(cfg.down_block_types):
        in_c, out_c = out_c, boc[i]
        for j in range(cfg.layers_per_block):
            _resnet(d, f"down_blocks.{i}.resnets.{j}", in_c if j == 0 else out_c, out_c, temb)
        if t == "CrossAttnDownBlock2D":
            for j in range(cfg.layers_per_block):
                _transformer(d, f"down_blocks.{i}.attentions.{j}", out_c, cross)
        if i != len(boc) - 1:
            _conv(d, f"down_blocks.{i}.downsamplers.0.conv", out_c, out_c, 3)
    cm = boc[-1]
    _resnet(d, "mid_block.resnets.0", cm, cm, temb)
    _transformer(d, "mid_block.attentions.0", cm, cross)
    _resnet(d, "mid_block.resnets.1", cm, cm, temb)
    if kind == "controlnet":
        k = 0
        _conv(d, f"controlnet_down_blocks.{k}", boc[0], boc[0], 1)
        for i, c in enumerate(boc):
            for _ in range(cfg.layers_per_block):
                k += 1
                _conv(d, f"controlnet_down_blocks.{k}", c, c, 1)
            if i != len(boc) - 1:
                k += 1
                _conv(d, f"controlnet_down_blocks.{k}", c, c, 1)
        _conv(d, "controlnet_mid_block", cm, cm, 1)
        return d
    rev = list(reversed(boc))
    out_c = rev[0]
    n = cfg.layers_per_block + 1
    for i, t in enumerate(cfg.up_block_types):
        prev_c, out_c = out_c, rev[i]
        in_c = rev[min(i + 1, len(boc) - 1)]
        for j in range(n):
            skip = in_c if j == n - 1 else out_c
            rin = prev_c if j == 0 else out_c
            _resnet(d, f"up_blocks.{i}.resnets.{j}", rin + skip, out_c, temb)
        if t == "CrossAttnUpBlock2D":
            for j in range(n):
                _transformer(d, f"up_blocks.{i}.attentions.{j}", out_c, cross)
        if i != len(boc) - 1:
            _conv(d, f"up_blocks.{i}.upsamplers.0.conv", out_c, out_c, 3)
    if kind == "brushnet":
        k = 0
        _conv(d, f"brushnet_down_blocks.{k}", boc[0], boc[0], 1)
        for i, c in enumerate(boc):
            for _ in range(cfg.layers_per_block):
                k += 1
                _conv(d, f"brushnet_down_blocks.{k}", c, c, 1)
            if i != len(boc) - 1:
                k += 1
                _conv(d, f"brushnet_down_blocks.{k}", c, c, 1)
        _conv(d, "brushnet_mid_block", cm, cm, 1)
        k = 0
        for i, c in enumerate(rev):
            for _ in range(n):
                _conv(d, f"brushnet_up_blocks.{k}", c, c, 1)
                k += 1
            if i != len(boc) - 1:
                _conv(d, f"brushnet_up_blocks.{k}", c, c, 1)
                k += 1
    else:
        _norm(d, "conv_norm_out", boc[0])
        _conv(d, "conv_out", cfg.out_channels, boc[0], 3)
    return d


def synthetic_state_dict(cfg: NetConfig, kind: str = "unet", seed: int = 1234,
                         zero_conv_scale: float = 0.1, device="cpu") -> Dict[str, torch.Tensor]:
    """Deterministic synthetic weights (no checkpoints are reachable offline):
    conv/linear ~ N(0, 1/fan_in), biases ~ 0.02 N(0,1), norm gamma ~ 1 + 0.05 N(0,1),
    zero-convs ~ 0.1 N(0, 1/fan_in) so the injection path is exercised."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in param_shapes(cfg, kind).items():
        is_zero = name.startswith(("brushnet_", "controlnet_down", "controlnet_mid")) or name.startswith(
            "controlnet_cond_embedding.conv_out")
        if len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g) / fan_in ** 0.5 * (zero_conv_scale if is_zero else 1.0)
        elif "norm" in name and name.endswith("weight"):
            t = 1.0 + 0.05 * torch.randn(shape, generator=g)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
        out[name] = t.to(device)
    return out
