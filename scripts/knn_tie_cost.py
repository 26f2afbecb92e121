import sys, torch, numpy as np
sys.path.insert(0, '.')
import swarm_b200 as sb
from swarm_b200 import ops
dev = torch.device('cuda:0')
B, N = 65536, 12
models = np.load('tests/golden/models.npz'); pre = 'ObstacleAvoidance/0/'
w = sb.pack_weights({k_[len(pre):]: torch.from_numpy(models[k_]) for k_ in models.files if k_.startswith(pre)}, dev)
cfg = ops.make_config(1, B, N, sb._lib.GRAPH_KNN, 5)
g = torch.Generator().manual_seed(0)
centers = (torch.tensor([0.6, -0.6]) + 0.1 * torch.randn(B, 2, generator=g)).to(dev)
for noise in (0.0, 1e-4):
    for ticks in (50,):
        ms = 0
        for rep in range(4):
            state = ops.reset_grid(cfg, centers)
            state[:, :, :2] += noise * torch.randn(B, N, 2, device=dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.rollout(cfg, w, state, ticks); b.record(); torch.cuda.synchronize()
            if rep: ms += a.elapsed_time(b)
        print('noise', noise, 'agent-steps/s %.3e' % (B * N * ticks * 3 / (ms * 1e-3)))
