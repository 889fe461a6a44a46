"""ConfigMixin / register_to_config: keep the constructor arguments as `self.config.<name>` (attribute and item access),
which is all the reference model files use (`self.config.addition_embed_type`, `unet.config.in_channels`, ...)."""
import functools
import inspect


class FrozenDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class ConfigMixin:
    config_name = None

    def register_to_config(self, **kw):
        self._internal_dict = FrozenDict({**getattr(self, "_internal_dict", {}), **kw})

    @property
    def config(self):
        return self._internal_dict


def register_to_config(init):
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        sig = inspect.signature(init)
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = {k: v for k, v in bound.arguments.items() if k != "self" and not k.startswith("_")}
        if "kwargs" in cfg:
            cfg.update(cfg.pop("kwargs"))
        self.register_to_config(**cfg)
        init(self, *args, **kwargs)
    return inner
