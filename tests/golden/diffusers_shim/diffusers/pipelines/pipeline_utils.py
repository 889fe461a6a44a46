"""DiffusionPipeline: module registry, config, device and progress bar — what `__call__` touches"""
import contextlib

import torch

from ..configuration_utils import ConfigMixin


class _Bar:
    def update(self, n=1):
        pass


class DiffusionPipeline(ConfigMixin):
    config_name = "model_index.json"

    def register_modules(self, **modules):
        for name, module in modules.items():
            setattr(self, name, module)

    @property
    def device(self):
        return torch.device("cpu")

    @property
    def _execution_device(self):
        return torch.device("cpu")

    def maybe_free_model_hooks(self):
        return None

    @contextlib.contextmanager
    def progress_bar(self, iterable=None, total=None):
        yield _Bar()


class StableDiffusionMixin:
    pass
