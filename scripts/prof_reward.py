"""A few swarm_scenario_reward launches for ncu (Flocking then Cohesion, N = 12, 1 Mi envs)."""
import sys
import torch
sys.path.insert(0, '.')
import swarm_b200 as sb
from swarm_b200 import ops
dev = torch.device('cuda:0')
L = sb._lib
B, N = 1 << 20, 12
g = torch.Generator().manual_seed(0)
centers = (torch.tensor([-1.6, 1.6]) + 0.4 * torch.randn(B, 2, generator=g)).to(dev)
state = ops.reset_grid(ops.make_config(0, B, N), centers)
state[:, :, :2] += 0.02 * torch.randn(B, N, 2, device=dev)
shaping = torch.zeros(B, N, 2, device=dev)
fl = ops.reward_spec(L.REWARD_FLOCKING, B, N)
ops.scenario_reward(fl, state, shaping, reset=True)
for _ in range(3):
    ops.scenario_reward(fl, state, shaping)
    ops.scenario_reward(ops.reward_spec(L.REWARD_COHESION, B, N), state)
torch.cuda.synchronize()
