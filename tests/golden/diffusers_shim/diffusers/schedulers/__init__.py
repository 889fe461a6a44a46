"""`diffusers.schedulers`: the DDIM scheduler is an adapter over oracle/ddim.py (pinned separately by the known-answer
tests in tests/test_scheduler.py) exposing the diffusers call contract the pipeline relies on: `config`, `order`,
`init_noise_sigma`, `set_timesteps(n, device=)`, `timesteps`, `scale_model_input`, `step(..., return_dict=False)[0]`."""
import enum

from oracle.ddim import DDIMOracle

from ..configuration_utils import FrozenDict


class KarrasDiffusionSchedulers(enum.Enum):
    DDIMScheduler = 1


class DDIMScheduler(DDIMOracle):
    def __init__(self, **kw):
        super().__init__(**kw)
        self._internal_dict = FrozenDict(steps_offset=self.steps_offset, num_train_timesteps=self.num_train_timesteps)

    @property
    def config(self):
        return self._internal_dict

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output: bool = False,
             generator=None, variance_noise=None, return_dict: bool = True):
        prev = super().step(model_output, timestep, sample, eta=eta, generator=generator, variance_noise=variance_noise)
        return (prev,)
