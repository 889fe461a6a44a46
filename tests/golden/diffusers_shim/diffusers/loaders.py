class PeftAdapterMixin:
    pass


class UNet2DConditionLoadersMixin:
    pass
