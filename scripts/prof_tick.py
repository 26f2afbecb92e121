"""A few eager train ticks (G = 32 and G = 4096) for ncu launch lists: python scripts/prof_tick.py [G]"""
import sys, torch
sys.path.insert(0, '.')
import numpy as np
import swarm_b200 as sb
from swarm_b200 import ops
dev = torch.device('cuda:0')
B, N = 4096, 12
G = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = ops.make_config(1, B, N)
g = torch.Generator().manual_seed(0)
centers = (torch.tensor([0.6, -0.6]) + 0.1 * torch.randn(B, 2, generator=g)).to(dev)
state = ops.reset_grid(cfg, centers)
models = np.load('tests/golden/models.npz')
pre = 'ObstacleAvoidance/0/'
w = sb.pack_weights({k[len(pre):]: torch.from_numpy(models[k]) for k in models.files if k.startswith(pre)}, dev)
w_t = w.clone(); m = torch.zeros_like(w); v = torch.zeros_like(w)
ring = ops.ReplayRing(1 << 20, N, dev)
returns = torch.zeros(B, N, device=dev); hits = torch.zeros(B, dtype=torch.int32, device=dev)
tt = ops.TrainTick(cfg, ring, graphs_per_update=G)
tt.load_cursor(0, 0, 0.3)
for _ in range(12):
    tt.tick(w, w_t, m, v, state, returns, hits)
torch.cuda.synchronize()
print('ok', tt.read_cursor())
