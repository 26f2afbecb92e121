"""Stacked-network rollout for profiling: python scripts/prof_stack.py [B N ticks]"""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
import swarm_b200 as sb
from swarm_b200 import ops
B, N, ticks = (int(x) for x in (sys.argv[1:4] + ['4096', '12', '10'][len(sys.argv) - 1:]))
dev = torch.device('cuda:0'); L = sb._lib
z = np.load('tests/golden/flocking_models.npz')
sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('0/')}
spec = ops.stack_spec(3, 8, 7)
w = ops.pack_stack_weights(sd, spec, dev)
cfg = ops.make_config(L.SCENARIO_GOTO, B, N, L.GRAPH_KNN, 5)
g = torch.Generator().manual_seed(0)
centers = (torch.tensor([0.9, -0.9]) + 0.1 * torch.randn(B, 2, generator=g)).to(dev)
rs = ops.reward_spec(L.REWARD_FLOCKING, B, N)
state = ops.reset_grid(cfg, centers)
shaping = torch.zeros(B, N, 2, device=dev)
ops.scenario_reward(rs, state, shaping, reset=True)
for rep in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.rollout_stack(cfg, spec, w, state, ticks, reward=rs, shaping=shaping); b.record(); torch.cuda.synchronize()
    print('stack rollout: %.1f us per tick' % (a.elapsed_time(b) * 1e3 / ticks))
