"""Timesteps / TimestepEmbedding over the oracle's restatement (oracle/blocks.py: get_timestep_embedding,
TimestepEmbedding); the other embedding classes are never built by the SD-1.5 configuration"""
import torch.nn as nn

from oracle import blocks as _ob


class Timesteps(nn.Module):
    def __init__(self, num_channels, flip_sin_to_cos, downscale_freq_shift):
        super().__init__()
        self.num_channels, self.flip_sin_to_cos, self.downscale_freq_shift = num_channels, flip_sin_to_cos, downscale_freq_shift

    def forward(self, timesteps):
        return _ob.get_timestep_embedding(timesteps, self.num_channels, self.flip_sin_to_cos, self.downscale_freq_shift)


class TimestepEmbedding(_ob.TimestepEmbedding):
    def __init__(self, in_channels, time_embed_dim, act_fn="silu", post_act_fn=None, cond_proj_dim=None, out_dim=None):
        assert act_fn == "silu" and post_act_fn is None and cond_proj_dim is None and out_dim is None
        super().__init__(in_channels, time_embed_dim)

    def forward(self, sample, condition=None):
        assert condition is None
        return super().forward(sample)


def _placeholder(name):
    class _P:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name} is outside the hot path")
    _P.__name__ = name
    return _P


GaussianFourierProjection = _placeholder("GaussianFourierProjection")
GLIGENTextBoundingboxProjection = _placeholder("GLIGENTextBoundingboxProjection")
ImageHintTimeEmbedding = _placeholder("ImageHintTimeEmbedding")
ImageProjection = _placeholder("ImageProjection")
ImageTimeEmbedding = _placeholder("ImageTimeEmbedding")
TextImageProjection = _placeholder("TextImageProjection")
TextImageTimeEmbedding = _placeholder("TextImageTimeEmbedding")
TextTimeEmbedding = _placeholder("TextTimeEmbedding")
