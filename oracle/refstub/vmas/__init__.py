"""CPU stand-in for the slice of vmas==1.4.0 the reference uses (TEST INFRASTRUCTURE, see ../README.md):
``make_env`` and the ``Environment`` facade (call sites train_gcn_dqn.py:280-290,149,153,165,168-169;
tests/test_*.py:29-41; simulator.py:51,68,70,102)."""
import random as _random

import numpy as _np
import torch as _torch

from oracle import swarm_oracle as _so

__version__ = "1.4.0+refstub"


class _Box:
    def __init__(self, shape):
        self.shape = tuple(shape)


class _Discrete:
    def __init__(self, n):
        self.n = n


class Environment:
    """vmas.simulator.environment.Environment, discrete actions, dict spaces; restated from vmas 1.4.0
    (``__init__`` -> ``reset(seed=seed)``; ``step``: ``_set_action`` per agent, ``world.step()``, then per agent
    ``reward(agent).clone()`` and ``observation(agent)``; ``done = scenario.done() | steps >= max_steps``)."""

    def __init__(self, scenario, num_envs=32, device="cpu", max_steps=None, continuous_actions=True, seed=None,
                 dict_spaces=False, **kwargs):
        assert not continuous_actions, "the reference uses the discrete 9-way action set"
        self.scenario = scenario
        self.num_envs = num_envs
        self.device = _torch.device(device)
        self.max_steps = max_steps
        self.continuous_actions = continuous_actions
        self.dict_spaces = dict_spaces
        self.world = scenario.env_make_world(num_envs, self.device, **kwargs)
        self.agents = self.world.agents
        self.n_agents = len(self.agents)
        self.steps = _torch.zeros(num_envs)
        self.reset(seed=seed)
        act = {a.name: _Discrete(9) for a in self.agents}
        obs = {a.name: _Box((int(scenario.observation(a).shape[-1]),)) for a in self.agents}
        self.action_space = act if dict_spaces else list(act.values())
        self.observation_space = obs if dict_spaces else list(obs.values())

    def seed(self, seed=None):
        if seed is None:
            seed = 0
        _torch.manual_seed(seed)
        _np.random.seed(seed)
        _random.seed(seed)
        return [seed]

    def _collect(self, fn):
        if self.dict_spaces:
            return {a.name: fn(a) for a in self.agents}
        return [fn(a) for a in self.agents]

    def reset(self, seed=None, return_observations=True, return_info=False, return_dones=False):
        if seed is not None:
            self.seed(seed)
        self.scenario.env_reset_world_at(env_index=None)
        self.steps = _torch.zeros(self.num_envs)
        return self._collect(self.scenario.observation) if return_observations else None

    def done(self):
        dones = self.scenario.done().clone()
        if self.max_steps is not None:
            dones = dones | (self.steps >= self.max_steps)
        return dones

    def step(self, actions):
        if isinstance(actions, dict):
            assert len(actions) == self.n_agents and all(a.name in actions for a in self.agents), \
                "Expecting actions for all agents"
            actions = [actions[a.name] for a in self.agents]
        assert len(actions) == self.n_agents
        for agent, act in zip(self.agents, actions):
            act = _torch.as_tensor(act).reshape(self.num_envs, -1)[:, 0]
            agent.action.u = _so.decode_action(act.to(_torch.int64))        # vmas Environment._set_action
        self.world.step()
        self.steps += 1
        rewards, obs, infos = [], [], []
        for agent in self.agents:
            rewards.append(self.scenario.reward(agent).clone())
            obs.append(self.scenario.observation(agent))
            infos.append(self.scenario.info(agent))
        if self.dict_spaces:
            names = [a.name for a in self.agents]
            return dict(zip(names, obs)), dict(zip(names, rewards)), self.done(), dict(zip(names, infos))
        return obs, rewards, self.done(), infos

    def render(self, *args, **kwargs):
        raise NotImplementedError("rendering is not part of the checked path")


def make_env(scenario, num_envs, device="cpu", continuous_actions=True, wrapper=None, max_steps=None, seed=None,
             dict_spaces=False, **kwargs):
    assert wrapper is None and not isinstance(scenario, str)
    kwargs.pop("scenario_name", None)
    return Environment(scenario, num_envs=num_envs, device=device, max_steps=max_steps,
                       continuous_actions=continuous_actions, seed=seed, dict_spaces=dict_spaces, **kwargs)
