"""Transformer2DModel with the constructor / call signature the reference uses (unet_2d_blocks.py:1289-1300 and the
forward calls :1378-1385), over the oracle's restatement"""
from oracle import blocks as _ob


class Transformer2DModel(_ob.Transformer2DModel):
    def __init__(self, num_attention_heads=16, attention_head_dim=88, in_channels=None, num_layers=1,
                 cross_attention_dim=None, norm_num_groups=32, use_linear_projection=False, only_cross_attention=False,
                 upcast_attention=False, attention_type="default", **kw):
        assert not use_linear_projection and not only_cross_attention and not upcast_attention and attention_type == "default"
        assert not kw, kw
        super().__init__(num_attention_heads, attention_head_dim, in_channels, cross_attention_dim, num_layers=num_layers,
                         norm_num_groups=norm_num_groups)

    def forward(self, hidden_states, encoder_hidden_states=None, cross_attention_kwargs=None, attention_mask=None,
                encoder_attention_mask=None, return_dict=True, **kw):
        assert attention_mask is None and encoder_attention_mask is None and not kw
        assert not return_dict
        return (super().forward(hidden_states, encoder_hidden_states),)
