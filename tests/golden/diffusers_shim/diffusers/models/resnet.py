"""ResnetBlock2D / Downsample2D / Upsample2D with the constructor and call signatures the reference uses
(unet_2d_blocks.py:1274-1285, 1319-1323, 2542), over the oracle's restatements"""
from oracle import blocks as _ob


class ResnetBlock2D(_ob.ResnetBlock2D):
    def __init__(self, *, in_channels, out_channels=None, temb_channels=512, eps=1e-6, groups=32, dropout=0.0,
                 time_embedding_norm="default", non_linearity="swish", output_scale_factor=1.0, pre_norm=True, **kw):
        assert time_embedding_norm == "default" and non_linearity in ("swish", "silu") and dropout == 0.0 and pre_norm
        assert not kw, kw
        super().__init__(in_channels, out_channels or in_channels, temb_channels, groups=groups, eps=eps,
                         output_scale_factor=output_scale_factor)

    def forward(self, x, temb, scale=1.0):
        return super().forward(x, temb)


class Downsample2D(_ob.Downsample2D):
    def __init__(self, channels, use_conv=False, out_channels=None, padding=1, name="conv"):
        assert use_conv and (out_channels in (None, channels)) and name == "op"
        super().__init__(channels, padding=padding)

    def forward(self, x, scale=1.0):
        return super().forward(x)


class Upsample2D(_ob.Upsample2D):
    def __init__(self, channels, use_conv=False, use_conv_transpose=False, out_channels=None, name="conv"):
        assert use_conv and not use_conv_transpose and (out_channels in (None, channels))
        super().__init__(channels)

    def forward(self, x, output_size=None, scale=1.0):
        return super().forward(x, output_size)


def _placeholder(name):
    class _P:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name} is outside the hot path")
    _P.__name__ = name
    return _P


FirDownsample2D = _placeholder("FirDownsample2D")
FirUpsample2D = _placeholder("FirUpsample2D")
KDownsample2D = _placeholder("KDownsample2D")
KUpsample2D = _placeholder("KUpsample2D")
ResnetBlockCondNorm2D = _placeholder("ResnetBlockCondNorm2D")
