"""GoTo-position and ObstacleAvoidance scenarios behind the VMAS ``BaseScenario`` surface.

Mirrors src/scenarios/go_to_position_scenario.py:7-149 and
src/scenarios/obstacle_avoidance_scenario.py:7-181 (``make_world / reset_world_at / observation /
reward / done / info`` plus the metric methods the Simulator calls on ``env.scenario``).  The arithmetic
is not here: ``World.step`` runs ``swarm_sim_step`` and the callbacks return views of what the kernel
wrote for the agent they are asked about.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import _lib, ops


class _EntityState:
    def __init__(self, pos: torch.Tensor, vel: Optional[torch.Tensor]):
        self.pos = pos
        self.vel = vel


class Landmark:
    def __init__(self, name: str, collide: bool, pos: torch.Tensor):
        self.name = name
        self.collide = collide
        self.movable = False
        self.state = _EntityState(pos, None)

    def set_pos(self, pos: torch.Tensor, batch_index: Optional[int] = None) -> None:
        self.state.pos[...] = pos.to(self.state.pos.device)


class Agent:
    """View of one agent column of the world's SoA state."""

    def __init__(self, name: str, index: int, world: "World"):
        self.name = name
        self.index = index
        self.collide = True
        self.movable = True
        self._world = world
        self.goal: Optional[Landmark] = None
        self.pos_rew = torch.zeros(world.batch_dim, device=world.device)

    @property
    def state(self) -> _EntityState:
        s = self._world.state
        return _EntityState(s[:, self.index, 0:2], s[:, self.index, 2:4])

    @property
    def distance_to_goal(self) -> torch.Tensor:
        return self._world.last["dist"][:, self.index, 0]


class World:
    """Batched world: state f32[B,N,4] in HBM, stepped by the CUDA kernel (vmas World.step)."""

    def __init__(self, batch_dim: int, device, scenario_id: int, n_agents: int):
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.SwarmError(f"swarm_b200 worlds live on a CUDA device (got {device}); there is no CPU fallback")
        self.batch_dim = batch_dim
        self.device = device
        self.cfg = ops.make_config(scenario_id, batch_dim, n_agents)
        self.state = torch.zeros(batch_dim, n_agents, 4, dtype=torch.float32, device=device)
        self.agents: List[Agent] = []
        self.landmarks: List[Landmark] = []
        self.last: Dict[str, torch.Tensor] = {}
        self._refresh_outputs()

    # vmas attribute names
    @property
    def dt(self) -> float:
        return self.cfg.dt

    @property
    def entities(self):
        return self.landmarks + self.agents

    def add_agent(self, agent: Agent) -> None:
        self.agents.append(agent)

    def add_landmark(self, landmark: Landmark) -> None:
        self.landmarks.append(landmark)

    def _refresh_outputs(self) -> None:
        """Observation / distance terms of the current state without stepping (used after reset)."""
        B, N = self.cfg.num_envs, self.cfg.n_agents
        goal = torch.tensor([self.cfg.goal_x, self.cfg.goal_y], dtype=torch.float32, device=self.device)
        self.last = {
            "obs": torch.cat([self.state, goal.view(1, 1, 2).expand(B, N, 2)], dim=2),
            "rewards": torch.zeros(B, N, dtype=torch.float32, device=self.device),
            "flags": torch.zeros(B, N, dtype=torch.uint8, device=self.device),
            "dist": torch.zeros(B, N, 2, dtype=torch.float32, device=self.device),
        }

    def reset_to_grid(self, centers: torch.Tensor, env_index: Optional[int] = None) -> None:
        """generate_grid + set_pos for every agent; velocities zero (vmas world.reset)."""
        centers = centers.to(device=self.device, dtype=torch.float32).reshape(-1, 2)
        if env_index is None:
            if centers.shape[0] == 1:
                centers = centers.expand(self.batch_dim, 2)
            ops.reset_grid(self.cfg, centers.contiguous(), out=self.state)
        else:
            one = ops.clone_config(self.cfg, num_envs=1)
            ops.reset_grid(one, centers[:1].contiguous(), out=self.state[env_index])
        self._refresh_outputs()

    def step(self, actions: torch.Tensor) -> None:
        """actions int32[B,N] -> in-place world step; results kept in ``self.last``."""
        self.last = ops.sim_step(self.cfg, self.state, actions, state_out=self.state, want_obs=True)

    def get_distance(self, a, b) -> torch.Tensor:
        """world.get_distance(agent, obstacle) of the last step (oa:149-150,168,171)."""
        agent = a if isinstance(a, Agent) else b
        return self.last["dist"][:, agent.index, 1]


class BaseScenario:
    """The slice of vmas.simulator.scenario.BaseScenario the reference scripts rely on."""

    scenario_id = -1

    def __init__(self):
        self._world: Optional[World] = None

    @property
    def world(self) -> World:
        return self._world

    def env_make_world(self, batch_dim: int, device, **kwargs) -> World:
        self._world = self.make_world(batch_dim, device, **kwargs)
        return self._world

    def env_reset_world_at(self, env_index: Optional[int]) -> None:
        self.reset_world_at(env_index)

    # -- common to both scenarios ----------------------------------------------------------
    def _build_world(self, batch_dim: int, device, with_obstacle: bool) -> World:
        world = World(batch_dim, device, self.scenario_id, self.n_agents)
        B = batch_dim
        goal = Landmark("goal", collide=False, pos=torch.zeros(B, 2, device=world.device))
        world.add_landmark(goal)
        if with_obstacle:
            world.add_landmark(Landmark("obstacle", collide=True, pos=torch.zeros(B, 2, device=world.device)))
        for i in range(self.n_agents):
            agent = Agent(f"agent{i}", i, world)
            agent.goal = goal
            world.add_agent(agent)
        self.pos_rew = torch.zeros(B, device=world.device)
        self.final_rew = self.pos_rew.clone()
        return world

    def observation(self, agent: Agent) -> torch.Tensor:
        """cat[pos, vel, goal] f32[B,6] (go_to:124-132, oa:154-162)."""
        return self.world.last["obs"][:, agent.index]

    def reward(self, agent: Agent) -> torch.Tensor:
        return self.world.last["rewards"][:, agent.index]

    def done(self) -> torch.Tensor:
        return torch.zeros(self.world.batch_dim, device=self.world.device, dtype=torch.bool)

    def info(self, agent: Agent) -> Dict[str, torch.Tensor]:
        return {"pos_rew": agent.pos_rew, "final_rew": self.final_rew}

    def average_distance_to_goal(self) -> torch.Tensor:
        # torch.mean(torch.stack([agent.distance_to_goal ...])) (go_to:134-135) -- reduced on the host like the
        # reference (device = cpu there) so the metric is bit-identical; this is a per-tick metric read, not hot path
        return torch.mean(self.world.last["dist"][:, :, 0].cpu().transpose(0, 1).contiguous())

    def set_start_centers(self, centers: Optional[torch.Tensor]) -> None:
        """Batched extension: explicit per-env start centres f32[B,2] used by the next resets instead of
        the reference's single shared draw (hard part 3)."""
        self._explicit_centers = centers


class GoToPositionScenario(BaseScenario):
    scenario_id = _lib.SCENARIO_GOTO

    def make_world(self, batch_dim: int, device, **kwargs) -> World:
        self.pos_shaping_factor = kwargs.get("pos_shaping_factor", 1.0)
        self.agent_radius = kwargs.get("agent_radius", 0.1)      # unused by the physics
        self.n_agents = kwargs.get("n_agents", 1)
        self.seed = kwargs.get("seed", 1)
        self.per_env_centers = kwargs.get("per_env_centers", False)
        self._explicit_centers = None
        self.desired_distance = 0.15
        torch.manual_seed(self.seed)                              # go_to:23 (make_env never forwards `seed`)
        if torch.cuda.is_available():
            torch.cuda.manual_seed_all(self.seed)
        return self._build_world(batch_dim, device, with_obstacle=False)

    def reset_world_at(self, env_index: Optional[int] = None) -> None:
        w = self.world
        w.landmarks[0].set_pos(torch.tensor([w.cfg.goal_x, w.cfg.goal_y]))
        if self._explicit_centers is not None:
            centers = self._explicit_centers if env_index is None else self._explicit_centers[env_index:env_index + 1]
        else:
            n = w.batch_dim if (self.per_env_centers and env_index is None) else 1
            position_range = -torch.tensor([-1.5, 1.5])            # go_to:84-88
            if n == 1:
                centers = position_range + torch.normal(mean=torch.tensor([-0.6, 0.6]), std=torch.tensor([0.4, 0.4]))
            else:
                centers = position_range + torch.normal(mean=torch.tensor([-0.6, 0.6]).expand(n, 2),
                                                        std=torch.tensor([0.4, 0.4]).expand(n, 2))
        w.reset_to_grid(centers, env_index)

    def average_distance_to_obstacles(self) -> torch.Tensor:
        return torch.tensor(0.0)

    def obstacles_hits(self) -> torch.Tensor:
        return torch.tensor(0.0)


class ObstacleAvoidanceScenario(BaseScenario):
    scenario_id = _lib.SCENARIO_OBSTACLE_AVOIDANCE

    def make_world(self, batch_dim: int, device, **kwargs) -> World:
        self.random = kwargs.get("random", False)
        self.pos_shaping_factor = kwargs.get("pos_shaping_factor", 10.0)
        self.dist_shaping_factor = kwargs.get("dist_shaping_factor", 10.0)
        self.agent_radius = kwargs.get("agent_radius", 0.1)
        self.n_agents = kwargs.get("n_agents", 1)
        self.per_env_centers = kwargs.get("per_env_centers", False)
        self._explicit_centers = None
        self.n_obstacles = 1
        self.desired_distance = 0.15
        self.min_collision_distance_reward = 1
        self.min_collision_distance_count = 0.2
        return self._build_world(batch_dim, device, with_obstacle=True)

    def reset_world_at(self, env_index: Optional[int] = None) -> None:
        w = self.world
        w.landmarks[0].set_pos(torch.tensor([w.cfg.goal_x, w.cfg.goal_y]))
        w.landmarks[1].set_pos(torch.tensor([w.cfg.obstacle_x, w.cfg.obstacle_y]))
        if self._explicit_centers is not None:
            centers = self._explicit_centers if env_index is None else self._explicit_centers[env_index:env_index + 1]
        else:
            n = w.batch_dim if (self.per_env_centers and env_index is None) else 1
            if self.random and n == 1:                             # oa:100-102
                delta = torch.normal(mean=torch.tensor([0.0, 0.0]), std=torch.tensor([0.1, 0.1]))
            elif self.random:
                delta = torch.normal(mean=torch.tensor([0.0, 0.0]).expand(n, 2), std=torch.tensor([0.1, 0.1]).expand(n, 2))
            else:
                delta = torch.zeros(n, 2)
            centers = torch.tensor([0.6, -0.6]) + delta
        w.reset_to_grid(centers, env_index)

    def average_distance_to_obstacles(self) -> torch.Tensor:
        return torch.mean(self.world.last["dist"][:, :, 1].cpu().transpose(0, 1).contiguous())

    def obstacles_hits(self) -> torch.Tensor:
        hits = (self.world.last["flags"] & _lib.FLAG_HIT) != 0       # oa:170-173
        return torch.sum(hits)
