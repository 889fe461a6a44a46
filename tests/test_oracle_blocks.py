"""CPU tests: the oracle's restatements of diffusers' primitive blocks (oracle/blocks.py, "parity unpinned" against the
absent package) against INDEPENDENT implementations of the same definitions that do exist in this image: torch's
nn.MultiheadAttention for the head layout / scaling of `Attention`, closed forms for the sinusoidal timestep embedding,
torch.nn.functional for the resampling rules. Not a substitute for diffusers outputs — a guard against a restatement
that is self-consistent but wrong."""
import math

import torch
import torch.nn as nn

from oracle.blocks import (Attention, Downsample2D, FeedForward, ResnetBlock2D, TimestepEmbedding, Upsample2D,
                           get_timestep_embedding)


@torch.no_grad()
def test_attention_equals_torch_multihead_attention_self_and_cross():
    """heads are contiguous channel slices [B, N, heads, d], scores scaled by 1/sqrt(d), bias only on the output
    projection — torch's MultiheadAttention with the same matrices"""
    torch.manual_seed(0)
    for cross_dim in (None, 24):
        dim, heads = 32, 4
        att = Attention(dim, heads, dim // heads, cross_attention_dim=cross_dim).eval()
        for p in att.parameters():
            p.copy_(torch.randn_like(p) * 0.2)
        kv = cross_dim or dim
        mha = nn.MultiheadAttention(dim, heads, bias=True, kdim=kv, vdim=kv, batch_first=True).eval()
        if cross_dim is None:
            mha.in_proj_weight.copy_(torch.cat([att.to_q.weight, att.to_k.weight, att.to_v.weight]))
        else:
            mha.q_proj_weight.copy_(att.to_q.weight)
            mha.k_proj_weight.copy_(att.to_k.weight)
            mha.v_proj_weight.copy_(att.to_v.weight)
        mha.in_proj_bias.zero_()
        mha.out_proj.weight.copy_(att.to_out[0].weight)
        mha.out_proj.bias.copy_(att.to_out[0].bias)
        x = torch.randn(2, 10, dim)
        ctx = None if cross_dim is None else torch.randn(2, 7, cross_dim)
        ref = mha(x, x if ctx is None else ctx, x if ctx is None else ctx, need_weights=False)[0]
        assert torch.allclose(att(x, ctx), ref, atol=2e-6), (att(x, ctx) - ref).abs().max()


def test_timestep_embedding_closed_forms():
    """DDPM / Transformer sinusoid, `flip_sin_to_cos=True, freq_shift=0` (the SD config, ref:unet_2d_condition.py:554):
    [cos(t w_k) | sin(t w_k)], w_k = 10000^(-k / half)"""
    t = torch.tensor([0.0, 1.0, 321.0, 999.0])
    dim = 320
    half = dim // 2
    e = get_timestep_embedding(t, dim, flip_sin_to_cos=True, downscale_freq_shift=0.0)
    assert e.shape == (4, dim)
    w = torch.tensor([10000.0 ** (-k / half) for k in range(half)], dtype=torch.float64)
    want = torch.cat([torch.cos(t.double()[:, None] * w), torch.sin(t.double()[:, None] * w)], dim=1)
    assert torch.allclose(e.double(), want, atol=2e-4)  # fp32 arguments up to 999 rad
    assert torch.equal(e[0, :half], torch.ones(half)) and torch.equal(e[0, half:], torch.zeros(half))
    # unflipped order and the freq_shift = 1 variant (original DDPM code): denominators half - 1
    e2 = get_timestep_embedding(t, dim, flip_sin_to_cos=False, downscale_freq_shift=1.0)
    w2 = torch.tensor([10000.0 ** (-k / (half - 1)) for k in range(half)], dtype=torch.float64)
    want2 = torch.cat([torch.sin(t.double()[:, None] * w2), torch.cos(t.double()[:, None] * w2)], dim=1)
    assert torch.allclose(e2.double(), want2, atol=2e-4)


@torch.no_grad()
def test_timestep_mlp_resnet_and_feed_forward_definitions():
    """TimestepEmbedding = linear_2(silu(linear_1)); ResnetBlock2D (App. A.1) and FeedForward(GEGLU) (App. A.5) written
    out once more with torch.nn.functional on the modules' own parameters"""
    import torch.nn.functional as F

    torch.manual_seed(1)
    te = TimestepEmbedding(16, 32).eval()
    x = torch.randn(3, 16)
    assert torch.allclose(te(x), F.linear(F.silu(F.linear(x, te.linear_1.weight, te.linear_1.bias)),
                                          te.linear_2.weight, te.linear_2.bias), atol=1e-6)
    rb = ResnetBlock2D(16, 32, 24, groups=4, eps=1e-5).eval()
    h = torch.randn(2, 16, 6, 5)
    temb = torch.randn(2, 24)
    y = F.conv2d(F.silu(F.group_norm(h, 4, rb.norm1.weight, rb.norm1.bias, 1e-5)), rb.conv1.weight, rb.conv1.bias,
                 padding=1)
    y = y + F.linear(F.silu(temb), rb.time_emb_proj.weight, rb.time_emb_proj.bias)[:, :, None, None]
    y = F.conv2d(F.silu(F.group_norm(y, 4, rb.norm2.weight, rb.norm2.bias, 1e-5)), rb.conv2.weight, rb.conv2.bias,
                 padding=1)
    short = F.conv2d(h, rb.conv_shortcut.weight, rb.conv_shortcut.bias)  # 1x1 because in != out
    assert torch.allclose(rb(h, temb), short + y, atol=1e-5)
    ff = FeedForward(16).eval()
    z = torch.randn(2, 5, 16)
    proj = F.linear(z, ff.net[0].proj.weight, ff.net[0].proj.bias)
    hidden, gate = proj.chunk(2, dim=-1)  # value half first, gate half second
    want = F.linear(hidden * F.gelu(gate), ff.net[2].weight, ff.net[2].bias)
    assert torch.allclose(ff(z), want, atol=1e-6)


@torch.no_grad()
def test_resampling_rules():
    """Downsample2D = 3x3 conv, stride 2, padding 1 (h -> ceil(h / 2)); Upsample2D = nearest x2 — or to an explicit
    size (`upsample_size`, ref:unet_2d_condition.py:1120-1126) — then 3x3 conv"""
    import torch.nn.functional as F

    torch.manual_seed(2)
    d = Downsample2D(8).eval()
    x = torch.randn(1, 8, 7, 10)
    assert torch.allclose(d(x), F.conv2d(x, d.conv.weight, d.conv.bias, stride=2, padding=1), atol=1e-6)
    assert d(x).shape[-2:] == (math.ceil(7 / 2), 5)
    u = Upsample2D(8).eval()
    y = torch.randn(1, 8, 4, 5)
    assert torch.allclose(u(y), F.conv2d(F.interpolate(y, scale_factor=2.0, mode="nearest"), u.conv.weight, u.conv.bias,
                                         padding=1), atol=1e-6)
    assert torch.allclose(u(y, output_size=(7, 10)),
                          F.conv2d(F.interpolate(y, size=(7, 10), mode="nearest"), u.conv.weight, u.conv.bias, padding=1),
                          atol=1e-6)


def test_build_synthetic_equals_default_construction_then_init():
    """oracle.unet.build_synthetic (meta-device layout, no default initialisation) writes the same parameters as
    init_synthetic_ on an ordinarily constructed net — bench.py's CPU leg relies on it"""
    from oracle.unet import BrushNetOracle, UNet2DConditionOracle, UNetConfig, build_synthetic, init_synthetic_

    for cls, cin in ((UNet2DConditionOracle, 9), (BrushNetOracle, 4)):
        a = build_synthetic(cls, UNetConfig.tiny(cin), seed=7).state_dict()
        b = init_synthetic_(cls(UNetConfig.tiny(cin)), seed=7).state_dict()
        assert list(a) == list(b)
        assert all(torch.equal(a[k], b[k]) for k in a)
