"""Single-env CPU oracle of the Flocking and Cohesion scenarios (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates, with one separately rounded float32 torch op per reference op and ``batch_dim = 1`` like every script of the
reference runs them,

  * ``FlockingScenario``   src/scenarios/flocking_scenario.py:9-205  (reset with its shaping memory :93-121, collective
                           reward :124-171, observation :173-182)
  * ``CohesionScenario``   src/scenarios/cohesion_scenario.py:9-101  (fixed start table :44-64, reward :66-85,
                           observation :87-94)

on the vmas world the GoTo oracle already restates (``swarm_oracle.OracleWorld``: colliding sphere agents, no
colliding landmark).

Parity status of THIS file: **pinned** (since round 2) against the reference's own source.  The reference ships no
golden vectors for these two scenarios and vmas cannot be imported here, so the reference's ``flocking_scenario.py`` /
``cohesion_scenario.py`` are executed UNMODIFIED on a CPU stand-in of the vmas API (``oracle/refstub``, whose world step
is the restatement that reproduces the reference's shipped GoTo / ObstacleAvoidance trajectories bit for bit --
``tests/test_refstub.py`` proves that with the reference's own Simulator file).  ``tests/golden/make_reference_runs.py``
records what their ``reset_world_at`` / ``reward()`` code produces over 60-tick runs with contacts, collision terms,
on-goal bonuses and both sigma branches (``tests/golden/reference_runs.npz``), and ``tests/test_oracle_scenarios.py``
checks this restatement against it bit for bit -- including cohesion:80's ``np.exp`` on a tensor (numpy's float32 exp,
not torch's: 216 of 960 rewards differ in the last bit if ``torch.exp`` is used instead).
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch

from . import swarm_oracle as so

GOAL_POS = (-0.8, 0.8)                 # flocking:96
DESIRED_DISTANCE = 0.15                # flocking:20
MIN_COLLISION_DISTANCE = 0.005         # flocking:21
AGENT_COLLISION_REWARD = -1            # flocking:19
ON_GOAL_BONUS = 50                     # flocking:142
SIGMA = 0.15                           # cohesion:23
COHESION_START = [[-1.0, -1.0], [0.0, -1.0], [0.0, 1.0], [0.0, 0.0], [1.0, 1.0], [1.0, -1.0], [-1.0, 1.0], [1.0, 0.0],
                  [-1.0, 0.0]]         # cohesion:46-56


def flocking_draw_center() -> torch.Tensor:
    """flocking:94-98: position_range (-1, 1) + N((-0.6, 0.6), 0.4^2), one draw from the global generator."""
    position_range = torch.tensor([-1, 1])
    return position_range + torch.normal(mean=torch.tensor([-0.6, 0.6]), std=torch.tensor([0.4, 0.4]))


def flocking_grid(center: torch.Tensor, num_points: int, distance: float = DESIRED_DISTANCE) -> torch.Tensor:
    """flocking:48-81: the GoTo grid, but every visited grid point also draws two (unused) normal deviates, which
    advances the global generator by 2 per point."""
    x_center, y_center = center
    cols = math.ceil(math.sqrt(num_points))
    rows = math.ceil(num_points / cols)
    pts = []
    for i in range(rows):
        for j in range(cols):
            torch.normal(mean=torch.tensor([0.0]), std=torch.tensor([0.1]))      # x_dev (unused, flocking:70)
            torch.normal(mean=torch.tensor([0.0]), std=torch.tensor([0.1]))      # y_dev (unused, flocking:71)
            pts.append([x_center + (j - (cols - 1) / 2) * distance, y_center + (i - (rows - 1) / 2) * distance])
            if len(pts) >= num_points:
                break
        if len(pts) >= num_points:
            break
    return torch.tensor(pts)


class FlockingOracle:
    """One env of FlockingScenario.  Positions / velocities are lists of f32[1, 2] like vmas entity states."""

    def __init__(self, n_agents: int, pos_shaping_factor: float = 10.0, dist_shaping_factor: float = 10.0):
        self.n = n_agents
        self.pos_shaping_factor = pos_shaping_factor
        self.dist_shaping_factor = dist_shaping_factor
        self.world = so.OracleWorld(so.GOTO, n_agents)
        self.goal_radius = so.SPHERE_RADIUS                      # Landmark default shape Sphere(0.05)
        self.previous_distance_to_goal: List[Optional[torch.Tensor]] = [None] * n_agents
        self.previous_distance_to_agents: List[Optional[torch.Tensor]] = [None] * n_agents
        self.pos_rew = [torch.zeros(1) for _ in range(n_agents)]
        self.dist_rew = [torch.zeros(1) for _ in range(n_agents)]
        self.distance_to_goal = [torch.zeros(1) for _ in range(n_agents)]

    # -- reset (vmas world.reset zeroes every state, then flocking:93-121) -------------------------------------
    def reset(self, center: Optional[torch.Tensor] = None) -> None:
        w = self.world
        w.steps = 0
        w.goal = torch.tensor(list(GOAL_POS)).unsqueeze(0)
        for i in range(self.n):
            w.pos[i] = torch.zeros(1, 2)
            w.vel[i] = torch.zeros(1, 2)
        if center is None:
            center = flocking_draw_center()
        grid = flocking_grid(center, self.n)
        for i in range(self.n):
            w.pos[i] = grid[i].unsqueeze(0).clone()
            # the shaping memory of agent i is taken right after ITS placement: agents i+1.. are still at the origin
            self.previous_distance_to_goal[i] = (torch.linalg.vector_norm(w.pos[i] - w.goal, dim=1)
                                                 * self.pos_shaping_factor)
            self.previous_distance_to_agents[i] = self._spacing(i)

    def _spacing(self, i: int) -> torch.Tensor:
        w = self.world
        d = torch.stack([torch.linalg.vector_norm(w.pos[i] - w.pos[j], dim=-1) for j in range(self.n) if j != i], dim=1)
        return (d - DESIRED_DISTANCE).pow(2).mean(-1) * self.dist_shaping_factor

    # -- reward (flocking:124-171) -----------------------------------------------------------------------------
    def reward(self) -> torch.Tensor:
        """The collective reward f32[1] every agent receives for the current state (updates the shaping memory)."""
        w = self.world
        collective = 0
        for i in range(self.n):
            # distance_to_goal_reward
            self.distance_to_goal[i] = torch.linalg.vector_norm(w.pos[i] - w.goal, dim=-1)
            on_goal = self.distance_to_goal[i] < self.goal_radius
            shaped = self.distance_to_goal[i] * self.pos_shaping_factor
            self.pos_rew[i] = self.previous_distance_to_goal[i] - shaped
            self.previous_distance_to_goal[i] = shaped
            goal_reward = self.pos_rew[i]
            if on_goal:
                goal_reward = goal_reward + ON_GOAL_BONUS
            # agent_avoidance_reward
            avoidance = sum(AGENT_COLLISION_REWARD for j in range(self.n)
                            if j != i and so.get_distance(w.pos[i], w.pos[j]) <= MIN_COLLISION_DISTANCE)
            # distance_to_agents_reward
            spacing = self._spacing(i)
            self.dist_rew[i] = self.previous_distance_to_agents[i] - spacing
            self.previous_distance_to_agents[i] = spacing
            collective += goal_reward + avoidance + self.dist_rew[i]
        return collective

    def step(self, actions: torch.Tensor) -> torch.Tensor:
        self.world.step(actions)               # GoTo physics; its own reward is discarded
        return self.reward()

    def observations(self) -> torch.Tensor:
        return self.world.observations()       # cat[pos, vel, goal] (flocking:173-182)


class CohesionOracle:
    """One env of CohesionScenario (no landmark; agents start on the fixed table)."""

    def __init__(self, n_agents: int):
        if n_agents > len(COHESION_START):
            raise IndexError("CohesionScenario places at most 9 agents (cohesion:46-64)")
        self.n = n_agents
        self.world = so.OracleWorld(so.GOTO, n_agents)

    def reset(self) -> None:
        w = self.world
        w.steps = 0
        table = torch.tensor(COHESION_START, dtype=torch.float32)
        for i in range(self.n):
            w.pos[i] = table[i].unsqueeze(0).clone()
            w.vel[i] = torch.zeros(1, 2)

    def reward(self) -> torch.Tensor:
        """f32[N]: cohesion:66-85 per agent."""
        w = self.world
        out = []
        for i in range(self.n):
            distances = torch.cat([so.get_distance(w.pos[i], w.pos[j]) for j in range(self.n) if j != i])
            mn, mx = torch.min(distances), torch.max(distances)
            collision = 0 if mn > SIGMA else np.exp(-(mn / SIGMA))     # cohesion:80: numpy's float32 exp on a tensor
            cohesion = 0 if mn < SIGMA else -(mx - SIGMA)
            out.append(torch.as_tensor(collision + cohesion, dtype=torch.float32).reshape(1))
        return torch.cat(out)

    def step(self, actions: torch.Tensor) -> torch.Tensor:
        self.world.step(actions)
        return self.reward()

    def observations(self) -> torch.Tensor:
        w = self.world                          # cat[pos, vel] (cohesion:87-94)
        return torch.cat([torch.cat([w.pos[i], w.vel[i]], dim=-1) for i in range(self.n)])
