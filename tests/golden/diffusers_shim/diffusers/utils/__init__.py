"""the handful of diffusers.utils names the reference model files import"""
import logging as _pylogging
from collections import OrderedDict
from dataclasses import fields, is_dataclass

import torch

USE_PEFT_BACKEND = False


class BaseOutput(OrderedDict):
    """dataclass-style outputs that are also tuples / dicts (only attribute and index access are used here)"""

    def __post_init__(self):
        if is_dataclass(self):
            for f in fields(self):
                v = getattr(self, f.name)
                if v is not None:
                    self[f.name] = v

    def __getitem__(self, k):
        if isinstance(k, str):
            return dict(self.items())[k]
        return self.to_tuple()[k]

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())


class _Logging:
    @staticmethod
    def get_logger(name):
        return _pylogging.getLogger(name)


logging = _Logging()


def deprecate(*args, **kwargs):
    return None


def scale_lora_layers(model, weight):
    return None


def unscale_lora_layers(model, weight=None):
    return None


def is_torch_version(op, version):
    import operator
    from packaging import version as V

    ops = {">": operator.gt, ">=": operator.ge, "<": operator.lt, "<=": operator.le, "==": operator.eq}
    return ops[op](V.parse(torch.__version__.split("+")[0]), V.parse(version))


def is_accelerate_available():
    return False


def is_accelerate_version(op, version):
    return False


def replace_example_docstring(example):
    return lambda fn: fn
