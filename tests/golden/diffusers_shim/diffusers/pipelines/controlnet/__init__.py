class MultiControlNetModel:
    def __init__(self, *a, **k):
        raise NotImplementedError("multi-ControlNet is outside the hot path (the reference app passes one)")
