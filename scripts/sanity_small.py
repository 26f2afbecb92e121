"""Small invocations of the round-2 kernels for compute-sanitizer: python scripts/sanity_small.py"""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
import swarm_b200 as sb
from swarm_b200 import ops
dev = torch.device('cuda:0')
models = np.load('tests/golden/models.npz'); pre = 'ObstacleAvoidance/0/'
w = sb.pack_weights({k[len(pre):]: torch.from_numpy(models[k]) for k in models.files if k.startswith(pre)}, dev)
g = torch.Generator().manual_seed(0)
for N, B, graph, k in ((12, 37, sb._lib.GRAPH_COMPLETE, 5), (12, 37, sb._lib.GRAPH_KNN, 5), (7, 20, sb._lib.GRAPH_KNN, 7),
                       (20, 9, sb._lib.GRAPH_KNN, 6)):
    cfg = ops.make_config(1, B, N, graph, k)
    centers = (torch.tensor([0.6, -0.6]) + 0.1 * torch.randn(B, 2, generator=g)).to(dev)
    state = ops.reset_grid(cfg, centers)
    ring = ops.ReplayRing(4096, N, dev)
    out = ops.rollout(cfg, w, state, 6, epsilon=0.3, replay=ring, trace=dict(q=True, edges=True))
    q, a = ops.gatq_forward(cfg, w, state, want_actions=True)
    G = min(len(ring), 50)
    gcfg = ops.clone_config(cfg, num_envs=G)
    grad, loss, td = ops.dqn_grad(gcfg, w, w.clone(), ring, torch.arange(G, device=dev), G, want_td=True)
    torch.cuda.synchronize()
    print(N, B, graph, float(out['returns'].sum()), float(loss), float(grad.abs().max()))
print('ok')
