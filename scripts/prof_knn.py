"""kNN rollout for profiling: python scripts/prof_knn.py [B N k ticks]   (ObstacleAvoidance, greedy, kNN graph)"""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
import swarm_b200 as sb
from swarm_b200 import ops
B, N, k, ticks = (int(x) for x in (sys.argv[1:5] + ['4096', '12', '5', '20'][len(sys.argv) - 1:]))
dev = torch.device('cuda:0')
models = np.load('tests/golden/models.npz'); pre = 'ObstacleAvoidance/0/'
w = sb.pack_weights({k_[len(pre):]: torch.from_numpy(models[k_]) for k_ in models.files if k_.startswith(pre)}, dev)
cfg = ops.make_config(1, B, N, sb._lib.GRAPH_KNN, k)
g = torch.Generator().manual_seed(0)
centers = (torch.tensor([0.6, -0.6]) + 0.1 * torch.randn(B, 2, generator=g)).to(dev)
for rep in range(3):
    state = ops.reset_grid(cfg, centers)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.rollout(cfg, w, state, ticks); b.record(); torch.cuda.synchronize()
    print('rollout %d ticks: %.3f ms, %.3e agent-steps/s' % (ticks, a.elapsed_time(b), B * N * ticks / (a.elapsed_time(b) * 1e-3)))
