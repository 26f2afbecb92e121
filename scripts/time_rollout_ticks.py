"""Fixed cost vs per-tick cost of the fused rollout launch (C2): python scripts/time_rollout_ticks.py"""
import sys, torch
sys.path.insert(0, '.')
import numpy as np
import swarm_b200 as sb
from swarm_b200 import ops
dev = torch.device('cuda:0')
B, N = 4096, 12
cfg = ops.make_config(1, B, N)
g = torch.Generator().manual_seed(0)
centers = (torch.tensor([0.6, -0.6]) + 0.1 * torch.randn(B, 2, generator=g)).to(dev)
state = ops.reset_grid(cfg, centers)
models = np.load('tests/golden/models.npz')
pre = 'ObstacleAvoidance/0/'
w = sb.pack_weights({k[len(pre):]: torch.from_numpy(models[k]) for k in models.files if k.startswith(pre)}, dev)
returns = torch.zeros(B, N, device=dev); hits = torch.zeros(B, dtype=torch.int32, device=dev)
def graph_time(fn, reps=20):
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph(); s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(gr, stream=s):
        for _ in range(reps): fn()
    gr.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for T in (1, 2, 3, 5, 10, 50):
    print(T, 'ticks: %.1f us' % graph_time(lambda: ops.rollout(cfg, w, state, T, returns=returns, hits=hits)))
