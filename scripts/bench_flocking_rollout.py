"""Fused rollout with the Flocking reward on the C2 shape (4 096 envs x 12 agents, complete graph, greedy, 100 ticks per
launch), next to the same launch with GoTo's reward.  usage: python scripts/bench_flocking_rollout.py [out.json]"""
import json, sys
import numpy as np, torch
sys.path.insert(0, '.')
import swarm_b200 as sb
from swarm_b200 import ops
dev = torch.device('cuda:0')
L = sb._lib
B, N, T = 4096, 12, 100
models = np.load('tests/golden/models.npz')
pre = 'GoTo/0/'
w = sb.pack_weights({k[len(pre):]: torch.from_numpy(models[k]) for k in models.files if k.startswith(pre)}, dev)
cfg = ops.make_config(L.SCENARIO_GOTO, B, N)
g = torch.Generator().manual_seed(0)
centers = (torch.tensor([-1.6, 1.6]) + 0.4 * torch.randn(B, 2, generator=g)).to(dev)
spec = ops.reward_spec(L.REWARD_FLOCKING, B, N)
shaping = torch.zeros(B, N, 2, device=dev)
res = {}
for name, kw in (('goto_reward', {}), ('flocking_reward', dict(flocking=spec, shaping=shaping))):
    ms = []
    for rep in range(6):
        state = ops.reset_grid(cfg, centers)
        if kw:
            ops.scenario_reward(spec, state, shaping, reset=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.rollout(cfg, w, state, T, **kw); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    best = float(np.median(ms[1:]))
    res[name] = {'ms_per_100_ticks': best, 'agent_steps_per_s': B * N * T / (best * 1e-3)}
print(json.dumps(res))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], 'w'), indent=1)
