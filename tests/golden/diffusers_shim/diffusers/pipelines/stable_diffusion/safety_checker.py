class StableDiffusionSafetyChecker:
    def __init__(self, *a, **k):
        raise NotImplementedError("the safety checker is outside the hot path")
