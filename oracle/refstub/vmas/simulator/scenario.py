"""vmas.simulator.scenario.BaseScenario (the env_* wrappers of vmas 1.4.0)."""
import torch


class BaseScenario:
    def __init__(self):
        self._world = None

    @property
    def world(self):
        return self._world

    def to(self, device):
        return None

    def env_make_world(self, batch_dim, device, **kwargs):
        self._world = self.make_world(batch_dim, device, **kwargs)
        return self._world

    def env_reset_world_at(self, env_index):
        self.world.reset(env_index)
        self.reset_world_at(env_index)

    def done(self):
        return torch.zeros(self.world.batch_dim, dtype=torch.bool)

    def info(self, agent):
        return {}
