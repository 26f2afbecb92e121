"""Large-swarm kernels for profiling: python scripts/prof_large.py [radius|complete|knn] [N B ticks]"""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
import swarm_b200 as sb
from swarm_b200 import ops
mode = sys.argv[1] if len(sys.argv) > 1 else "radius"
N, B, ticks = (int(x) for x in (sys.argv[2:5] + ['1024', '1024', '4'][len(sys.argv) - 2:]))
dev = torch.device('cuda:0'); L = sb._lib
models = np.load('tests/golden/models.npz'); pre = 'ObstacleAvoidance/0/'
w = sb.pack_weights({k_[len(pre):]: torch.from_numpy(models[k_]) for k_ in models.files if k_.startswith(pre)}, dev)
gm = {"radius": L.GRAPH_RADIUS, "complete": L.GRAPH_COMPLETE, "knn": L.GRAPH_KNN}[mode]
cfg = ops.make_config(L.SCENARIO_OBSTACLE_AVOIDANCE, B, N, gm, 10, graph_radius=0.35)
g = torch.Generator().manual_seed(9)
centers = (torch.tensor([0.6, -0.6]) + 0.1 * torch.randn(B, 2, generator=g)).to(dev)
state = ops.reset_grid(cfg, centers)
for rep in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.rollout_large(cfg, w, state, ticks); b.record(); torch.cuda.synchronize()
    print('%s %d x %d: %.3f ms per tick' % (mode, N, B, a.elapsed_time(b) / ticks))
