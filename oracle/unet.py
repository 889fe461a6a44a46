"""ORACLE (test infrastructure, not product): fp32 restatement of the reference
`UNet2DConditionModel` (powerpaint/models/unet_2d_condition.py:166-481 layout, :1040-1363 forward)
for the SD-1.5 family of configs, including every BrushNet hook (:1203-1207, :1220-1223,
:1232-1253, :1299-1300, :1316-1339) and the ControlNet residual path (:1263-1272, :1296-1297),
and of `BrushNetModel` (powerpaint/models/BrushNet_CA.py:139-454 layout, :456-542 from_unet,
:690-952 forward, non-guess-mode) and of diffusers' ControlNetModel.

COMPOSITION PINNED: tests/golden/unet_composition.npz holds outputs of the reference's own model files (imported
unmodified over tests/golden/diffusers_shim by tests/golden/make_unet_golden.py); tests/test_oracle_golden.py holds
these classes to them. The arithmetic of the diffusers blocks they are built from stays PARITY UNPINNED
(oracle/blocks.py header).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .blocks import (CrossAttnDownBlock2D, CrossAttnUpBlock2D, DownBlock2D, TimestepEmbedding,
                     UNetMidBlock2DCrossAttn, UpBlock2D, get_timestep_embedding)


@dataclass
class UNetConfig:
    """Subset of the diffusers config keys the hot path reads (SURVEY.md §5 'Config / flags')."""
    in_channels: int = 9
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    attention_head_dim: int = 8          # == number of heads (diffusers naming quirk, :232-238)
    cross_attention_dim: int = 768
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    sample_size: int = 64
    flip_sin_to_cos: bool = True
    freq_shift: int = 0
    mid_block_scale_factor: float = 1.0
    down_block_types: Tuple[str, ...] = ("CrossAttnDownBlock2D", "CrossAttnDownBlock2D",
                                         "CrossAttnDownBlock2D", "DownBlock2D")
    up_block_types: Tuple[str, ...] = ("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D",
                                       "CrossAttnUpBlock2D")
    time_cond_proj_dim: Optional[int] = None

    @staticmethod
    def sd15(in_channels: int = 9) -> "UNetConfig":
        return UNetConfig(in_channels=in_channels)

    @staticmethod
    def tiny(in_channels: int = 9) -> "UNetConfig":
        """small config for fast tests: 4 heads, d = 8/16/32/32, groups of 8"""
        return UNetConfig(in_channels=in_channels, block_out_channels=(32, 64, 128, 128),
                          attention_head_dim=4, cross_attention_dim=64, norm_num_groups=8, sample_size=8)


def _build_trunk(m: nn.Module, cfg: UNetConfig, with_up: bool = True):
    boc = cfg.block_out_channels
    temb_c = boc[0] * 4
    heads, cross, groups, eps = cfg.attention_head_dim, cfg.cross_attention_dim, cfg.norm_num_groups, cfg.norm_eps
    m.time_embedding = TimestepEmbedding(boc[0], temb_c)
    m.down_blocks = nn.ModuleList()
    out_c = boc[0]
    for i, t in enumerate(cfg.down_block_types):
        in_c, out_c = out_c, boc[i]
        final = i == len(boc) - 1
        if t == "CrossAttnDownBlock2D":
            m.down_blocks.append(CrossAttnDownBlock2D(in_c, out_c, temb_c, heads, cross, groups, eps,
                                                      cfg.layers_per_block, not final))
        elif t == "DownBlock2D":
            m.down_blocks.append(DownBlock2D(in_c, out_c, temb_c, groups, eps, cfg.layers_per_block, not final))
        else:
            raise ValueError(t)
    m.mid_block = UNetMidBlock2DCrossAttn(boc[-1], temb_c, heads, cross, groups, eps, cfg.mid_block_scale_factor)
    if not with_up:
        return
    m.up_blocks = nn.ModuleList()
    rev = list(reversed(boc))
    out_c = rev[0]
    for i, t in enumerate(cfg.up_block_types):
        prev_c, out_c = out_c, rev[i]
        in_c = rev[min(i + 1, len(boc) - 1)]
        final = i == len(boc) - 1
        if t == "CrossAttnUpBlock2D":
            m.up_blocks.append(CrossAttnUpBlock2D(in_c, out_c, prev_c, temb_c, heads, cross, groups, eps,
                                                  cfg.layers_per_block + 1, not final))
        elif t == "UpBlock2D":
            m.up_blocks.append(UpBlock2D(in_c, out_c, prev_c, temb_c, groups, eps, cfg.layers_per_block + 1, not final))
        else:
            raise ValueError(t)


def _timesteps_tensor(timestep, sample):
    """reference unet_2d_condition.py:918-931: python scalar / 0-d tensor -> [batch]"""
    t = timestep
    if not torch.is_tensor(t):
        t = torch.tensor([t], dtype=torch.float64 if isinstance(t, float) else torch.int64, device=sample.device)
    elif t.dim() == 0:
        t = t[None].to(sample.device)
    return t.expand(sample.shape[0])


class UNet2DConditionOracle(nn.Module):
    def __init__(self, cfg: UNetConfig):
        super().__init__()
        self.cfg = cfg
        boc = cfg.block_out_channels
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        _build_trunk(self, cfg)
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, boc[0], eps=cfg.norm_eps)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    def time_embed(self, sample, timestep):
        t = _timesteps_tensor(timestep, sample)
        t_emb = get_timestep_embedding(t, self.cfg.block_out_channels[0], self.cfg.flip_sin_to_cos, self.cfg.freq_shift)
        return self.time_embedding(t_emb.to(sample.dtype))

    def forward(self, sample, timestep, encoder_hidden_states,
                down_block_additional_residuals: Optional[Sequence[torch.Tensor]] = None,
                mid_block_additional_residual: Optional[torch.Tensor] = None,
                down_block_add_samples: Optional[List[torch.Tensor]] = None,
                mid_block_add_sample: Optional[torch.Tensor] = None,
                up_block_add_samples: Optional[List[torch.Tensor]] = None):
        ctx = encoder_hidden_states
        # unet_2d_condition.py:1112-1126: the upsample size is forwarded when a spatial dim is not a multiple
        # of 2 ** (number of upsamplers)
        factor = 2 ** sum(1 for b in self.up_blocks if b.upsamplers is not None)
        forward_upsample_size = any(d % factor != 0 for d in sample.shape[-2:])
        emb = self.time_embed(sample, timestep)
        sample = self.conv_in(sample)
        is_controlnet = mid_block_additional_residual is not None and down_block_additional_residuals is not None
        is_brushnet = (down_block_add_samples is not None and mid_block_add_sample is not None
                       and up_block_add_samples is not None)
        if is_brushnet:  # the reference consumes the caller's lists with pop(0); copy to stay pure
            down_block_add_samples = list(down_block_add_samples)
            up_block_add_samples = list(up_block_add_samples)
        res_samples = (sample,)  # skip 0 is captured BEFORE the first add (:1220 before :1223)
        if is_brushnet:
            sample = sample + down_block_add_samples.pop(0)
        for blk in self.down_blocks:
            adds = None
            if is_brushnet and len(down_block_add_samples) > 0:
                n = len(blk.resnets) + (blk.downsamplers is not None)
                adds = [down_block_add_samples.pop(0) for _ in range(n)]
            sample, outs = blk(sample, emb, ctx, adds)
            res_samples += outs
        if is_controlnet:
            res_samples = tuple(r + a for r, a in zip(res_samples, down_block_additional_residuals))
        sample = self.mid_block(sample, emb, ctx)
        if is_controlnet:
            sample = sample + mid_block_additional_residual
        if is_brushnet:
            sample = sample + mid_block_add_sample
        for i, blk in enumerate(self.up_blocks):
            n = len(blk.resnets)
            res = res_samples[-n:]
            res_samples = res_samples[:-n]
            adds = None
            if is_brushnet and len(up_block_add_samples) > 0:
                k = n + (blk.upsamplers is not None)
                adds = [up_block_add_samples.pop(0) for _ in range(k)]
            # :1310-1312 — the size of the next block's first skip
            upsample_size = (res_samples[-1].shape[2:] if forward_upsample_size and i != len(self.up_blocks) - 1
                             else None)
            sample = blk(sample, res, emb, ctx, upsample_size, adds)
        sample = F.silu(self.conv_norm_out(sample))
        return self.conv_out(sample)


def _zero_conv(c_in, c_out, k=1):
    conv = nn.Conv2d(c_in, c_out, k, padding=k // 2)
    nn.init.zeros_(conv.weight)
    nn.init.zeros_(conv.bias)
    return conv


class BrushNetOracle(nn.Module):
    """reference powerpaint/models/BrushNet_CA.py: full UNet trunk with cross-attention + 1x1
    zero-convs on 12 down / 1 mid / 15 up states; returns (down[12], mid, up[15]) x scale."""

    def __init__(self, cfg: UNetConfig, conditioning_channels: int = 5):
        super().__init__()
        self.cfg = cfg
        boc = cfg.block_out_channels
        self.conv_in_condition = nn.Conv2d(cfg.in_channels + conditioning_channels, boc[0], 3, padding=1)
        _build_trunk(self, cfg)
        self.brushnet_down_blocks = nn.ModuleList([_zero_conv(boc[0], boc[0])])
        for i, c in enumerate(boc):
            for _ in range(cfg.layers_per_block):
                self.brushnet_down_blocks.append(_zero_conv(c, c))
            if i != len(boc) - 1:
                self.brushnet_down_blocks.append(_zero_conv(c, c))
        self.brushnet_mid_block = _zero_conv(boc[-1], boc[-1])
        self.brushnet_up_blocks = nn.ModuleList()
        for i, c in enumerate(reversed(boc)):
            for _ in range(cfg.layers_per_block + 1):
                self.brushnet_up_blocks.append(_zero_conv(c, c))
            if i != len(boc) - 1:
                self.brushnet_up_blocks.append(_zero_conv(c, c))

    @classmethod
    def from_unet(cls, unet: UNet2DConditionOracle, conditioning_channels: int = 5):
        """BrushNet_CA.py:456-542: clone trunk weights; conv_in -> channels 0:4 AND 4:8, ch 8 zero."""
        cfg = UNetConfig(**{**unet.cfg.__dict__})
        bn = cls(cfg, conditioning_channels)
        w = torch.zeros_like(bn.conv_in_condition.weight)
        w[:, :4] = unet.conv_in.weight
        w[:, 4:8] = unet.conv_in.weight
        bn.conv_in_condition.weight = nn.Parameter(w)
        bn.conv_in_condition.bias = nn.Parameter(unet.conv_in.bias.detach().clone())
        bn.time_embedding.load_state_dict(unet.time_embedding.state_dict())
        bn.down_blocks.load_state_dict(unet.down_blocks.state_dict(), strict=False)
        bn.mid_block.load_state_dict(unet.mid_block.state_dict(), strict=False)
        bn.up_blocks.load_state_dict(unet.up_blocks.state_dict(), strict=False)
        return bn

    def forward(self, sample, timestep, encoder_hidden_states, brushnet_cond, conditioning_scale: float = 1.0):
        ctx = encoder_hidden_states
        t = _timesteps_tensor(timestep, sample)
        t_emb = get_timestep_embedding(t, self.cfg.block_out_channels[0], self.cfg.flip_sin_to_cos, self.cfg.freq_shift)
        emb = self.time_embedding(t_emb.to(sample.dtype))
        sample = self.conv_in_condition(torch.cat([sample, brushnet_cond], dim=1))
        res_samples = (sample,)
        for blk in self.down_blocks:
            sample, outs = blk(sample, emb, ctx)
            res_samples += outs
        down_out = [conv(r) for r, conv in zip(res_samples, self.brushnet_down_blocks)]
        sample = self.mid_block(sample, emb, ctx)
        mid_out = self.brushnet_mid_block(sample)
        up_states = ()
        for i, blk in enumerate(self.up_blocks):
            n = len(blk.resnets)
            res = res_samples[-n:]
            res_samples = res_samples[:-n]
            upsample_size = res_samples[-1].shape[2:] if i != len(self.up_blocks) - 1 else None
            sample, outs = blk(sample, res, emb, ctx, upsample_size, None, True)
            up_states += outs
        up_out = [conv(r) for r, conv in zip(up_states, self.brushnet_up_blocks)]
        return ([d * conditioning_scale for d in down_out], mid_out * conditioning_scale,
                [u * conditioning_scale for u in up_out])


class ControlNetOracle(nn.Module):
    """diffusers ControlNetModel (SD-1.5), used by the reference at
    powerpaint/pipelines/pipeline_PowerPaint_ControlNet.py:1686-1694."""

    def __init__(self, cfg: UNetConfig, conditioning_embedding_out_channels=(16, 32, 96, 256),
                 conditioning_channels: int = 3):
        super().__init__()
        self.cfg = cfg
        boc = cfg.block_out_channels
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        _build_trunk(self, cfg, with_up=False)
        ce = conditioning_embedding_out_channels
        emb = nn.Module()
        emb.conv_in = nn.Conv2d(conditioning_channels, ce[0], 3, padding=1)
        emb.blocks = nn.ModuleList()
        for i in range(len(ce) - 1):
            emb.blocks.append(nn.Conv2d(ce[i], ce[i], 3, padding=1))
            emb.blocks.append(nn.Conv2d(ce[i], ce[i + 1], 3, padding=1, stride=2))
        emb.conv_out = _zero_conv(ce[-1], boc[0], 3)
        self.controlnet_cond_embedding = emb
        self.controlnet_down_blocks = nn.ModuleList([_zero_conv(boc[0], boc[0])])
        for i, c in enumerate(boc):
            for _ in range(cfg.layers_per_block):
                self.controlnet_down_blocks.append(_zero_conv(c, c))
            if i != len(boc) - 1:
                self.controlnet_down_blocks.append(_zero_conv(c, c))
        self.controlnet_mid_block = _zero_conv(boc[-1], boc[-1])

    def cond_embedding(self, cond):
        e = self.controlnet_cond_embedding
        x = F.silu(e.conv_in(cond))
        for blk in e.blocks:
            x = F.silu(blk(x))
        return e.conv_out(x)

    def forward(self, sample, timestep, encoder_hidden_states, controlnet_cond, conditioning_scale: float = 1.0):
        ctx = encoder_hidden_states
        t = _timesteps_tensor(timestep, sample)
        t_emb = get_timestep_embedding(t, self.cfg.block_out_channels[0], self.cfg.flip_sin_to_cos, self.cfg.freq_shift)
        emb = self.time_embedding(t_emb.to(sample.dtype))
        sample = self.conv_in(sample) + self.cond_embedding(controlnet_cond)
        res_samples = (sample,)
        for blk in self.down_blocks:
            sample, outs = blk(sample, emb, ctx)
            res_samples += outs
        sample = self.mid_block(sample, emb, ctx)
        down = [conv(r) * conditioning_scale for r, conv in zip(res_samples, self.controlnet_down_blocks)]
        mid = self.controlnet_mid_block(sample) * conditioning_scale
        return down, mid


def init_synthetic_(module: nn.Module, seed: int = 1234, zero_conv_scale: float = 0.1):
    """SURVEY.md §8d synthetic weights: conv/linear ~ N(0, 1/fan_in), biases 0, norm gamma 1 /
    beta 0 (small perturbations so affine paths are exercised), zero-convs ~ N(0, 1/fan_in)*0.1."""
    g = torch.Generator().manual_seed(seed)
    for name, p in module.named_parameters():
        with torch.no_grad():
            is_zero = name.startswith(("brushnet_", "controlnet_down", "controlnet_mid")) or name.endswith(
                "controlnet_cond_embedding.conv_out.weight")
            if p.dim() >= 2:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) / fan_in ** 0.5 * (zero_conv_scale if is_zero else 1.0))
            elif "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.05 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
    return module


def build_synthetic(cls, cfg: "UNetConfig", seed: int = 1234, **kw) -> nn.Module:
    """`init_synthetic_(cls(cfg))` without torch's default initialisation: the module is laid out on the meta device
    and materialised empty, then every parameter is written by `init_synthetic_` (the oracle nets hold no buffers), so
    the result is the same tensor for tensor. Constructing the SD-1.5-size net the ordinary way spends minutes in
    `kaiming_uniform_` on a small host."""
    with torch.device("meta"):
        m = cls(cfg)
    assert not list(m.buffers()), "to_empty() would leave buffers uninitialised"
    return init_synthetic_(m.to_empty(device="cpu"), seed, **kw).eval()

