"""``make_env`` and the environment facade the reference scripts drive.

Mirrors the slice of vmas==1.4.0 ``make_env`` / ``Environment`` used at
src/training/train_gcn_dqn.py:280-290,149,153,165,168-169 and tests/test_*.py:29-41 plus
src/simulation/simulator.py:51,68,70,102: ``reset()``, ``step(dict)``, ``n_agents``, ``agents``,
``max_steps``, ``observation_space``, ``action_space``, ``scenario``, ``device``, ``render``; seeding
re-seeds torch, numpy and ``random`` and the constructor performs one reset, exactly like vmas.
"""
from __future__ import annotations

import random as _py_random
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from .scenarios import BaseScenario


class Box:
    """Shape-only stand-in for gym.spaces.Box (the scripts read ``.shape[0]``, train:79)."""

    def __init__(self, low: float, high: float, shape):
        self.low, self.high, self.shape = low, high, tuple(shape)
        self.dtype = np.float32


class Discrete:
    """Stand-in for gym.spaces.Discrete (the scripts read ``.n``, train:80)."""

    def __init__(self, n: int):
        self.n = n
        self.shape = ()
        self.dtype = np.int64


class Environment:
    def __init__(self, scenario: BaseScenario, num_envs: int = 32, device="cuda", max_steps: Optional[int] = None,
                 continuous_actions: bool = True, seed: Optional[int] = None, dict_spaces: bool = False, **kwargs):
        if continuous_actions:
            raise NotImplementedError("the swarm_b200 kernels implement the discrete 9-way action set the reference "
                                      "uses (continuous_actions=False at train:284 / tests:33)")
        self.scenario = scenario
        self.num_envs = num_envs
        self.device = torch.device(device)
        self.max_steps = max_steps
        self.continuous_actions = continuous_actions
        self.dict_spaces = dict_spaces
        self.world = scenario.env_make_world(num_envs, self.device, **kwargs)
        self.agents = self.world.agents
        self.n_agents = len(self.agents)
        self.steps = torch.zeros(num_envs, device=self.device)
        act_spaces = {a.name: Discrete(9) for a in self.agents}
        self.action_space = act_spaces if dict_spaces else list(act_spaces.values())
        self.reset(seed=seed)
        # observation sizes come from the scenario (6 for the reference scenarios: pos, vel, goal; train:79)
        obs_spaces = {a.name: Box(-float("inf"), float("inf"), (int(scenario.observation(a).shape[-1]),))
                      for a in self.agents}
        self.observation_space = obs_spaces if dict_spaces else list(obs_spaces.values())

    # -- seeding (vmas Environment.seed) ------------------------------------------------------
    def seed(self, seed: Optional[int] = None) -> List[int]:
        if seed is None:
            seed = 0
        torch.manual_seed(seed)
        np.random.seed(seed)
        _py_random.seed(seed)
        return [seed]

    def reset(self, seed: Optional[int] = None, return_observations: bool = True, return_info: bool = False,
              return_dones: bool = False):
        if seed is not None:
            self.seed(seed)
        self.scenario.env_reset_world_at(None)
        self.steps = torch.zeros(self.num_envs, device=self.device)
        out = []
        if return_observations:
            out.append(self._collect(self.scenario.observation))
        if return_info:
            out.append(self._collect(self.scenario.info))
        if return_dones:
            out.append(self.done())
        return out[0] if len(out) == 1 else tuple(out)

    def reset_at(self, index: int, return_observations: bool = True):
        self.scenario.env_reset_world_at(index)
        self.steps[index] = 0
        return self._collect(self.scenario.observation) if return_observations else None

    def _collect(self, fn):
        # tensors cross the seam on the device the script asked for (host-seam mode: a D2H copy of what a kernel-backed
        # scenario returned; a no-op otherwise)
        def home(v):
            return v.to(self.device) if isinstance(v, torch.Tensor) and v.device != self.device else v
        if self.dict_spaces:
            return {a.name: home(fn(a)) for a in self.agents}
        return [home(fn(a)) for a in self.agents]

    def done(self) -> torch.Tensor:
        dones = self.scenario.done().clone().to(self.device)
        if self.max_steps is not None:
            dones = dones | (self.steps >= self.max_steps)
        return dones

    # -- step ----------------------------------------------------------------------------------
    def _gather_actions(self, actions: Union[Dict[str, torch.Tensor], Sequence[torch.Tensor], torch.Tensor]) -> torch.Tensor:
        """{'agent{i}': int[B] or [B,1]} / list / tensor [B,N] -> int32[B,N] on the device."""
        B, N = self.num_envs, self.n_agents
        if isinstance(actions, torch.Tensor):
            if actions.shape != (B, N):
                raise AssertionError(f"action tensor must be [{B}, {N}], got {list(actions.shape)}")
            if actions.is_cuda:
                torch._assert_async(((actions >= 0) & (actions <= 8)).all(), "Discrete actions must be in [0, 8]")
            return actions.to(device=self.world.compute_device, dtype=torch.int32).contiguous()
        if isinstance(actions, dict):
            if len(actions) != N or any(a.name not in actions for a in self.agents):
                raise AssertionError("Expecting actions for all agents")     # vmas Environment.step assertion
            cols = [actions[a.name] for a in self.agents]
        else:
            if len(actions) != N:
                raise AssertionError(f"Expecting actions for {N} agents, got {len(actions)} actions")
            cols = list(actions)
        cols = [torch.as_tensor(c).reshape(-1) for c in cols]
        for c in cols:
            if c.numel() != B:
                raise AssertionError(f"Actions used in input of env must be of len {B}, got {c.numel()}")
        out = torch.stack(cols, dim=1)
        if not out.is_cuda:
            lo, hi = int(out.min()), int(out.max())
            if lo < 0 or hi > 8:
                raise AssertionError(f"Discrete actions must be in [0, 8], got [{lo}, {hi}]")
        else:
            # device tensors: asserted on the device without a host sync (vmas raises here); the kernels themselves
            # map an out-of-range action to "no force" so nothing is indexed out of bounds either way
            torch._assert_async(((out >= 0) & (out <= 8)).all(), "Discrete actions must be in [0, 8]")
        return out.to(device=self.world.compute_device, dtype=torch.int32).contiguous()

    def step(self, actions):
        self.world.step(self._gather_actions(actions))
        self.steps += 1
        obs = self._collect(self.scenario.observation)
        rews = self._collect(lambda a: self.scenario.reward(a).clone())
        infos = self._collect(self.scenario.info)
        return obs, rews, self.done(), infos

    def render(self, *args, **kwargs):
        raise NotImplementedError("rendering is outside the B200 hot path (simulator.py:88-93 is optional)")


def make_env(scenario: BaseScenario, num_envs: int, device="cuda", continuous_actions: bool = True, wrapper=None,
             max_steps: Optional[int] = None, seed: Optional[int] = None, dict_spaces: bool = False, **kwargs) -> Environment:
    """vmas.make_env with the argument names the reference passes (train:280-290, tests:29-41).
    ``scenario_name`` (tests:31) is accepted and ignored, as vmas does for scenario objects."""
    if isinstance(scenario, str):
        raise NotImplementedError("pass a scenario object (GoToPositionScenario / ObstacleAvoidanceScenario)")
    if wrapper is not None:
        raise NotImplementedError("wrappers are not part of the hot path (the reference passes wrapper=None)")
    kwargs.pop("scenario_name", None)
    return Environment(scenario, num_envs=num_envs, device=device, max_steps=max_steps,
                       continuous_actions=continuous_actions, seed=seed, dict_spaces=dict_spaces, **kwargs)
