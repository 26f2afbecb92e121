"""The VMAS scenario seam for USER-WRITTEN scenarios: a scenario file written against the vmas API (imported through
the shim's `vmas` module names) builds its own World / Landmark / Agent objects, computes rewards and observations
with its own torch code on views of the device state, and is stepped by the CUDA world step.  Physics is checked
against the oracle world step, the scenario's own reward / observation code against the same code on CPU tensors."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shim_modules():
    for p in (os.path.join(ROOT, "shim"), ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import vmas
    from vmas.simulator import core, scenario, utils
    return vmas, core, scenario, utils


def _make_scenario_class():
    vmas, core, scenario, utils = _shim_modules()

    class RendezvousScenario(scenario.BaseScenario):
        """N agents gather at a beacon next to a round obstacle: reward = -(own distance to the beacon) - 0.1 x mean
        distance to the other agents - 3 if closer than 0.1 to the obstacle surface; observation = own pos, vel and
        the vector to the beacon."""

        def make_world(self, batch_dim, device, **kwargs):
            self.n_agents = kwargs.get("n_agents", 4)
            world = core.World(batch_dim, device)
            beacon = core.Landmark(name="beacon", collide=False, color=utils.Color.BLACK)
            world.add_landmark(beacon)
            world.add_landmark(core.Landmark(name="rock", collide=True, shape=core.Sphere(radius=0.05), color=utils.Color.RED))
            for i in range(self.n_agents):
                agent = core.Agent(name=f"agent{i}", collide=True, color=utils.Color.GREEN, render_action=True)
                agent.goal = beacon
                agent.visits = torch.zeros(batch_dim, device=device)
                world.add_agent(agent)
            return world

        def reset_world_at(self, env_index=None):
            self.world.landmarks[0].set_pos(torch.tensor([-0.8, 0.8]), batch_index=env_index)
            self.world.landmarks[1].set_pos(torch.tensor([-0.1, 0.1]), batch_index=env_index)
            for i, agent in enumerate(self.world.agents):
                agent.set_pos(torch.tensor([0.05 + 0.11 * (i % 3), -0.05 - 0.11 * (i // 3)]), batch_index=env_index)

        def reward(self, agent):
            d_goal = torch.linalg.vector_norm(agent.state.pos - agent.goal.state.pos, dim=-1)
            others = [torch.linalg.vector_norm(agent.state.pos - a.state.pos, dim=-1) for a in self.world.agents if a is not agent]
            spread = torch.stack(others, dim=1).mean(dim=1)
            near = self.world.get_distance(agent, self.world.landmarks[1]) < 0.1
            return -d_goal - 0.1 * spread - 3.0 * near.float()

        def observation(self, agent):
            return torch.cat([agent.state.pos, agent.state.vel, agent.goal.state.pos - agent.state.pos], dim=-1)

    return vmas, RendezvousScenario


def test_user_scenario_on_the_cuda_world_step():
    from oracle import batched_oracle as bo, swarm_oracle as so
    vmas, RendezvousScenario = _make_scenario_class()
    N, B, T = 5, 3, 25
    env = vmas.make_env(RendezvousScenario(), num_envs=B, device="cuda:0", continuous_actions=False, dict_spaces=True,
                        wrapper=None, max_steps=T, seed=0, n_agents=N)
    assert env.n_agents == N and env.observation_space["agent0"].shape == (6,) and env.action_space["agent0"].n == 9
    obs = env.reset()
    start = torch.tensor([[0.05 + 0.11 * (i % 3), -0.05 - 0.11 * (i // 3)] for i in range(N)])
    pos = start.unsqueeze(0).expand(B, N, 2).contiguous()
    vel = torch.zeros(B, N, 2)
    goal, rock = torch.tensor(list(so.GOAL_POS)), torch.tensor(list(so.OBSTACLE_POS))
    assert torch.equal(obs["agent2"].cpu(), torch.cat([pos[:, 2], vel[:, 2], goal - pos[:, 2]], dim=-1))
    g = torch.Generator().manual_seed(4)
    contacts = 0
    for t in range(T):
        # drive the swarm towards the rock so that agent-agent and agent-obstacle contacts occur
        a = torch.randint(0, 9, (B, N), generator=g)
        a[:, :3] = 4 if t < 12 else a[:, :3]                                  # (-1, -1): up-left towards the rock / beacon
        obs, rews, done, info = env.step({f"agent{i}": a[:, i] for i in range(N)})
        ref = bo.step(so.OBSTACLE_AVOIDANCE, pos, vel, a)
        pos, vel = ref["pos"], ref["vel"]
        contacts += int((ref["contact"] != 0).sum() + (ref["flags"] & 1).sum())
        got = env.world.state.cpu()
        assert torch.allclose(got[..., :2], pos, rtol=0, atol=2e-6) and torch.allclose(got[..., 2:], vel, rtol=0, atol=2e-5)
        pos, vel = got[..., :2].contiguous(), got[..., 2:].contiguous()      # stay on the device trajectory
        for i in range(N):
            d_goal = torch.linalg.vector_norm(pos[:, i] - goal, dim=-1)
            spread = torch.stack([torch.linalg.vector_norm(pos[:, i] - pos[:, j], dim=-1) for j in range(N) if j != i], 1).mean(1)
            near = (torch.linalg.vector_norm(pos[:, i] - rock, dim=-1) - 0.05 - 0.05) < 0.1
            expect = -d_goal - 0.1 * spread - 3.0 * near.float()
            assert torch.allclose(rews[f"agent{i}"].cpu(), expect, rtol=1e-6, atol=1e-6)
            assert torch.allclose(obs[f"agent{i}"].cpu(), torch.cat([pos[:, i], vel[:, i], goal - pos[:, i]], -1), atol=1e-7)
        assert done.tolist() == [t == T - 1] * B
    assert contacts > 0, "the test trajectory should exercise contact forces"
    # per-env reset keeps the other envs untouched
    before = env.world.state.clone()
    env.reset_at(1)
    after = env.world.state
    assert torch.equal(after[0], before[0]) and torch.equal(after[2], before[2])
    assert torch.equal(after[1, :, :2].cpu(), start) and torch.count_nonzero(after[1, :, 2:]) == 0


def test_unsupported_worlds_fail_loudly():
    vmas, core, scenario, utils = _shim_modules()
    with pytest.raises(NotImplementedError):
        core.World(2, "cuda:0", substeps=2)
    with pytest.raises(NotImplementedError):
        core.World(2, "cuda:0", x_semidim=1.0)
    with pytest.raises(NotImplementedError):
        core.Agent(name="a", u_range=2.0)
    with pytest.raises(Exception, match="CUDA"):
        core.World(2, "cpu")

    class TwoRocks(scenario.BaseScenario):
        def make_world(self, batch_dim, device, **kwargs):
            w = core.World(batch_dim, device)
            w.add_landmark(core.Landmark(name="r1", collide=True))
            w.add_landmark(core.Landmark(name="r2", collide=True))
            w.add_agent(core.Agent(name="agent0"))
            return w

    with pytest.raises(NotImplementedError, match="at most one colliding"):
        vmas.make_env(TwoRocks(), num_envs=1, device="cuda:0", continuous_actions=False, dict_spaces=True)

    class MixedRadii(scenario.BaseScenario):
        def make_world(self, batch_dim, device, **kwargs):
            w = core.World(batch_dim, device)
            w.add_agent(core.Agent(name="agent0"))
            w.add_agent(core.Agent(name="agent1", shape=core.Sphere(radius=0.1)))
            return w

    with pytest.raises(NotImplementedError, match="one radius"):
        vmas.make_env(MixedRadii(), num_envs=1, device="cuda:0", continuous_actions=False, dict_spaces=True)
