"""ORACLE (test infrastructure, not product): AutoencoderKL (SD-1.5 VAE) restated as a plain PyTorch module.

Checker for `powerpaint_b200.models.autoencoder_kl.AutoencoderKL` (which runs encode / decode as recorded CUDA
programs on the repo's kernels) and the VAE the CPU host-logic tests inject into the pipelines. Reference call
sites: `_encode_vae_image` powerpaint/pipelines/pipeline_PowerPaint.py:657-669, decode :1051;
pipeline_PowerPaint_Brushnet_CA.py:1338-1341, :1476. Restates diffusers==0.27.0 `AutoencoderKL` (SURVEY.md App. A.10:
block_out_channels (128,256,512,512), latent_channels 4, scaling_factor 0.18215, GroupNorm eps 1e-6,
Downsample2D(padding=0) = F.pad (0,1,0,1) + stride-2 conv, one 512-channel attention head in the mid blocks) with
diffusers state-dict names. PARITY UNPINNED like the rest of oracle/: diffusers is absent here.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


class _Resnet(nn.Module):
    def __init__(self, cin, cout, groups=32, eps=1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class _Attn(nn.Module):
    """single-head spatial self-attention of the VAE mid block (diffusers Attention names)"""

    def __init__(self, c, groups=32, eps=1e-6):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=eps)
        self.to_q = nn.Linear(c, c)
        self.to_k = nn.Linear(c, c)
        self.to_v = nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Dropout(0.0)])

    def forward(self, x):
        b, c, h, w = x.shape
        r = x
        t = self.group_norm(x).view(b, c, h * w).transpose(1, 2)
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        o = self.to_out[0](o).transpose(1, 2).reshape(b, c, h, w)
        return o + r


class _Mid(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.resnets = nn.ModuleList([_Resnet(c, c, groups), _Resnet(c, c, groups)])
        self.attentions = nn.ModuleList([_Attn(c, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class _Down(nn.Module):
    def __init__(self, cin, cout, layers, add_down, groups):
        super().__init__()
        self.resnets = nn.ModuleList([_Resnet(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.downsamplers = None
        if add_down:
            ds = nn.Module()
            ds.conv = nn.Conv2d(cout, cout, 3, stride=2, padding=0)
            self.downsamplers = nn.ModuleList([ds])

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0].conv(F.pad(x, (0, 1, 0, 1)))  # diffusers Downsample2D(padding=0)
        return x


class _Up(nn.Module):
    def __init__(self, cin, cout, layers, add_up, groups):
        super().__init__()
        self.resnets = nn.ModuleList([_Resnet(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.upsamplers = None
        if add_up:
            us = nn.Module()
            us.conv = nn.Conv2d(cout, cout, 3, padding=1)
            self.upsamplers = nn.ModuleList([us])

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0].conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))
        return x


class _Encoder(nn.Module):
    def __init__(self, cin, latent, boc, layers, groups):
        super().__init__()
        self.conv_in = nn.Conv2d(cin, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        c = boc[0]
        for i, co in enumerate(boc):
            self.down_blocks.append(_Down(c, co, layers, i != len(boc) - 1, groups))
            c = co
        self.mid_block = _Mid(c, groups)
        self.conv_norm_out = nn.GroupNorm(groups, c, eps=1e-6)
        self.conv_out = nn.Conv2d(c, 2 * latent, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class _Decoder(nn.Module):
    def __init__(self, latent, cout, boc, layers, groups):
        super().__init__()
        rev = list(reversed(boc))
        self.conv_in = nn.Conv2d(latent, rev[0], 3, padding=1)
        self.mid_block = _Mid(rev[0], groups)
        self.up_blocks = nn.ModuleList()
        c = rev[0]
        for i, co in enumerate(rev):
            self.up_blocks.append(_Up(c, co, layers + 1, i != len(boc) - 1, groups))
            c = co
        self.conv_norm_out = nn.GroupNorm(groups, c, eps=1e-6)
        self.conv_out = nn.Conv2d(c, cout, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class DiagonalGaussianDistribution:
    def __init__(self, parameters: torch.Tensor):
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        dev = generator.device if generator is not None else self.mean.device
        noise = torch.randn(self.mean.shape, generator=generator, device=dev, dtype=self.mean.dtype)
        return self.mean + self.std * noise.to(self.mean.device)

    def mode(self) -> torch.Tensor:
        return self.mean


class AutoencoderKLOracle(nn.Module):
    def __init__(self, in_channels: int = 3, out_channels: int = 3, latent_channels: int = 4,
                 block_out_channels: Tuple[int, ...] = (128, 256, 512, 512), layers_per_block: int = 2,
                 norm_num_groups: int = 32, scaling_factor: float = 0.18215):
        super().__init__()
        self.config = SimpleNamespace(in_channels=in_channels, out_channels=out_channels,
                                      latent_channels=latent_channels, block_out_channels=block_out_channels,
                                      layers_per_block=layers_per_block, scaling_factor=scaling_factor,
                                      norm_num_groups=norm_num_groups)
        self.encoder = _Encoder(in_channels, latent_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.decoder = _Decoder(latent_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    def encode(self, x: torch.Tensor, return_dict: bool = True):
        dist = DiagonalGaussianDistribution(self.quant_conv(self.encoder(x)))
        return SimpleNamespace(latent_dist=dist) if return_dict else (dist,)

    def decode(self, z: torch.Tensor, return_dict: bool = True, generator=None):
        img = self.decoder(self.post_quant_conv(z))
        return SimpleNamespace(sample=img) if return_dict else (img,)

    @classmethod
    def synthetic(cls, seed: int = 4321, tiny: bool = False, **kw) -> "AutoencoderKLOracle":
        """deterministic random weights (no checkpoint reachable offline); `tiny` = small widths for tests"""
        if tiny:
            kw.setdefault("block_out_channels", (16, 32, 32, 32))
            kw.setdefault("norm_num_groups", 8)
            kw.setdefault("layers_per_block", 1)
        m = cls(**kw)
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for name, p in m.named_parameters():
                if p.dim() >= 2:
                    p.copy_(torch.randn(p.shape, generator=g) / p[0].numel() ** 0.5)
                elif "norm" in name and name.endswith("weight"):
                    p.fill_(1.0)
                else:
                    p.zero_()
        return m.eval()
