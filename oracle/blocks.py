"""ORACLE (test infrastructure, not product): fp32 PyTorch restatement of the diffusers==0.27.0
building blocks the PowerPaint hot path instantiates.

PARITY UNPINNED for these blocks: diffusers is an un-vendored dependency of the reference
(requirements/requirements.txt:3) that is not installable here and the reference ships no tests
or golden vectors. The restatement follows the published diffusers 0.27.0
algorithm and is anchored on the reference's own call sites, cited per class below.
Module / parameter names equal the diffusers state-dict names so a real
`unet.safetensors` loads unchanged.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------- embeddings
def get_timestep_embedding(timesteps: torch.Tensor, dim: int, flip_sin_to_cos: bool = True,
                           downscale_freq_shift: float = 0.0, max_period: int = 10000) -> torch.Tensor:
    """diffusers Timesteps / get_timestep_embedding (call site: reference
    powerpaint/models/unet_2d_condition.py:914-938 `get_time_embed`, ctor :554)."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - downscale_freq_shift)
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class TimestepEmbedding(nn.Module):
    """linear_2(silu(linear_1(x))) — reference ctor call unet_2d_condition.py:269."""

    def __init__(self, in_channels: int, time_embed_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


# ----------------------------------------------------------------------------- resnet
class ResnetBlock2D(nn.Module):
    """diffusers ResnetBlock2D, time_embedding_norm="default" (reference instantiations:
    powerpaint/models/unet_2d_blocks.py:789,1274,1428,2499,2672)."""

    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, groups: int = 32,
                 eps: float = 1e-5, output_scale_factor: float = 1.0):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None
        self.output_scale_factor = output_scale_factor

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return (x + h) / self.output_scale_factor


class Downsample2D(nn.Module):
    """conv 3x3 stride 2 pad 1 (use_conv=True, name="op"; reference unet_2d_blocks.py:1319)."""

    def __init__(self, channels: int, padding: int = 1):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=padding)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    """nearest x2 (or to `output_size`) then conv 3x3 (reference unet_2d_blocks.py:2542)."""

    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x, output_size=None):
        if output_size is None:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        else:
            x = F.interpolate(x, size=output_size, mode="nearest")
        return self.conv(x)


# ----------------------------------------------------------------------------- attention
class Attention(nn.Module):
    """diffusers Attention with AttnProcessor2_0 (no mask, scale 1/sqrt(d)); to_q/k/v bias-free,
    to_out[0] with bias (processors referenced at reference unet_2d_condition.py:24-31)."""

    def __init__(self, query_dim: int, heads: int, dim_head: int, cross_attention_dim: Optional[int] = None):
        super().__init__()
        inner = heads * dim_head
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(kv_dim, inner, bias=False)
        self.to_v = nn.Linear(kv_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def forward(self, x, encoder_hidden_states=None):
        ctx = x if encoder_hidden_states is None else encoder_hidden_states
        b, n, _ = x.shape
        q, k, v = self.to_q(x), self.to_k(ctx), self.to_v(ctx)
        d = q.shape[-1] // self.heads
        q = q.view(b, -1, self.heads, d).transpose(1, 2)
        k = k.view(b, -1, self.heads, d).transpose(1, 2)
        v = v.view(b, -1, self.heads, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, n, self.heads * d)
        return self.to_out[0](o)


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    """GEGLU(C -> 4C) -> Dropout(0) -> Linear(4C -> C)."""

    def __init__(self, dim: int, mult: int = 4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, dim_head: int, cross_attention_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, heads, dim_head, cross_attention_dim)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), encoder_hidden_states=ctx)
        x = x + self.ff(self.norm3(x))
        return x


class Transformer2DModel(nn.Module):
    """diffusers Transformer2DModel, use_linear_projection=False, num_layers=1 (reference
    instantiations: unet_2d_blocks.py:807,1289,2514)."""

    def __init__(self, heads: int, dim_head: int, in_channels: int, cross_attention_dim: int,
                 num_layers: int = 1, norm_num_groups: int = 32):
        super().__init__()
        inner = heads * dim_head
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim) for _ in range(num_layers)])
        self.proj_out = nn.Conv2d(inner, in_channels, 1)

    def forward(self, x, ctx):
        b, c, h, w = x.shape
        r = x
        x = self.proj_in(self.norm(x))
        inner = x.shape[1]
        x = x.permute(0, 2, 3, 1).reshape(b, h * w, inner)
        for blk in self.transformer_blocks:
            x = blk(x, ctx)
        x = x.reshape(b, h, w, inner).permute(0, 3, 1, 2).contiguous()
        return self.proj_out(x) + r


# ----------------------------------------------------------------------------- UNet blocks
class CrossAttnDownBlock2D(nn.Module):
    """reference powerpaint/models/unet_2d_blocks.py:1237-1402 (BrushNet adds :1388-1389,:1397-1398)."""
    has_cross_attention = True

    def __init__(self, in_c, out_c, temb_c, heads, cross_dim, groups, eps, num_layers=2, add_downsample=True):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(in_c if i == 0 else out_c, out_c, temb_c, groups, eps) for i in range(num_layers)])
        self.attentions = nn.ModuleList(
            [Transformer2DModel(heads, out_c // heads, out_c, cross_dim, norm_num_groups=groups) for _ in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_c)]) if add_downsample else None

    def forward(self, h, temb, ctx, adds=None):
        outs = ()
        for resnet, attn in zip(self.resnets, self.attentions):
            h = attn(resnet(h, temb), ctx)
            if adds is not None:
                h = h + adds.pop(0)
            outs += (h,)
        if self.downsamplers is not None:
            h = self.downsamplers[0](h)
            if adds is not None:
                h = h + adds.pop(0)
            outs += (h,)
        return h, outs


class DownBlock2D(nn.Module):
    """reference unet_2d_blocks.py:1405-1500."""
    has_cross_attention = False

    def __init__(self, in_c, out_c, temb_c, groups, eps, num_layers=2, add_downsample=True):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(in_c if i == 0 else out_c, out_c, temb_c, groups, eps) for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_c)]) if add_downsample else None

    def forward(self, h, temb, ctx=None, adds=None):
        outs = ()
        for resnet in self.resnets:
            h = resnet(h, temb)
            if adds is not None:
                h = h + adds.pop(0)
            outs += (h,)
        if self.downsamplers is not None:
            h = self.downsamplers[0](h)
            if adds is not None:
                h = h + adds.pop(0)
            outs += (h,)
        return h, outs


class UNetMidBlock2DCrossAttn(nn.Module):
    """reference unet_2d_blocks.py:756-899: resnet, then (attention, resnet) x num_layers."""
    has_cross_attention = True

    def __init__(self, c, temb_c, heads, cross_dim, groups, eps, output_scale_factor=1.0):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, temb_c, groups, eps, output_scale_factor) for _ in range(2)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, c // heads, c, cross_dim, norm_num_groups=groups)])

    def forward(self, h, temb, ctx):
        h = self.resnets[0](h, temb)
        for attn, resnet in zip(self.attentions, self.resnets[1:]):
            h = resnet(attn(h, ctx), temb)
        return h


class CrossAttnUpBlock2D(nn.Module):
    """reference unet_2d_blocks.py:2458-2643 (capture :2627-2631, adds :2630,:2638)."""
    has_cross_attention = True

    def __init__(self, in_c, out_c, prev_c, temb_c, heads, cross_dim, groups, eps, num_layers=3, add_upsample=True):
        super().__init__()
        resnets = []
        for i in range(num_layers):
            res_skip = in_c if i == num_layers - 1 else out_c
            res_in = prev_c if i == 0 else out_c
            resnets.append(ResnetBlock2D(res_in + res_skip, out_c, temb_c, groups, eps))
        self.resnets = nn.ModuleList(resnets)
        self.attentions = nn.ModuleList(
            [Transformer2DModel(heads, out_c // heads, out_c, cross_dim, norm_num_groups=groups) for _ in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_c)]) if add_upsample else None

    def forward(self, h, res_tuple, temb, ctx, upsample_size=None, adds=None, return_res=False):
        outs = ()
        for resnet, attn in zip(self.resnets, self.attentions):
            res = res_tuple[-1]
            res_tuple = res_tuple[:-1]
            h = attn(resnet(torch.cat([h, res], dim=1), temb), ctx)
            if return_res:
                outs += (h,)
            if adds is not None:
                h = h + adds.pop(0)
        if self.upsamplers is not None:
            h = self.upsamplers[0](h, upsample_size)
            if return_res:
                outs += (h,)
            if adds is not None:
                h = h + adds.pop(0)
        return (h, outs) if return_res else h


class UpBlock2D(nn.Module):
    """reference unet_2d_blocks.py:2646-2770."""
    has_cross_attention = False

    def __init__(self, in_c, out_c, prev_c, temb_c, groups, eps, num_layers=3, add_upsample=True):
        super().__init__()
        resnets = []
        for i in range(num_layers):
            res_skip = in_c if i == num_layers - 1 else out_c
            res_in = prev_c if i == 0 else out_c
            resnets.append(ResnetBlock2D(res_in + res_skip, out_c, temb_c, groups, eps))
        self.resnets = nn.ModuleList(resnets)
        self.upsamplers = nn.ModuleList([Upsample2D(out_c)]) if add_upsample else None

    def forward(self, h, res_tuple, temb, ctx=None, upsample_size=None, adds=None, return_res=False):
        outs = ()
        for resnet in self.resnets:
            res = res_tuple[-1]
            res_tuple = res_tuple[:-1]
            h = resnet(torch.cat([h, res], dim=1), temb)
            if return_res:
                outs += (h,)
            if adds is not None:
                h = h + adds.pop(0)
        if self.upsamplers is not None:
            h = self.upsamplers[0](h, upsample_size)
            if return_res:
                outs += (h,)
            if adds is not None:
                h = h + adds.pop(0)
        return (h, outs) if return_res else h
