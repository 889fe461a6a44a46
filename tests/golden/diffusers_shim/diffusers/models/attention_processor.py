"""names the reference imports; only `Attention` (through Transformer2DModel) is on the hot path, and the processors
are never constructed by the reference's SD-1.5 configuration"""


class AttentionProcessor:
    pass


class AttnProcessor(AttentionProcessor):
    pass


class AttnProcessor2_0(AttentionProcessor):
    pass


class AttnAddedKVProcessor(AttentionProcessor):
    pass


class AttnAddedKVProcessor2_0(AttentionProcessor):
    pass


ADDED_KV_ATTENTION_PROCESSORS = (AttnAddedKVProcessor, AttnAddedKVProcessor2_0)
CROSS_ATTENTION_PROCESSORS = (AttnProcessor, AttnProcessor2_0)


class Attention:
    def __init__(self, *a, **k):
        raise NotImplementedError("stand-alone Attention blocks (AttnDownBlock2D, UNetMidBlock2D ...) are outside the hot path")
