class AdaGroupNorm:
    def __init__(self, *a, **k):
        raise NotImplementedError("AdaGroupNorm (resnet_time_scale_shift='spatial') is outside the hot path")
