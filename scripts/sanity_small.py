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
# ---- last session of round 2: grid broad phase, large-swarm forwards, grid world step, stacked networks, kNN memo ----
L = sb._lib
for N, B, gm, r in ((300, 3, L.GRAPH_RADIUS, 0.3), (1024, 2, L.GRAPH_RADIUS, 0.35), (130, 4, L.GRAPH_RADIUS, 0.0),
                    (200, 3, L.GRAPH_COMPLETE, 0.0), (640, 2, L.GRAPH_KNN, 0.0)):
    cfg = ops.make_config(1, B, N, gm, 10, graph_radius=r)
    centers = (torch.tensor([0.6, -0.6]) + 0.1 * torch.randn(B, 2, generator=g)).to(dev)
    state = ops.reset_grid(cfg, centers)
    state[..., :2] *= 0.6                                   # squeezed: contacts
    if gm == L.GRAPH_RADIUS:
        rp, src = ops.graph_build_radius_csr(cfg, state)
    if gm != L.GRAPH_KNN:
        q, a = ops.gatq_forward_large(cfg, w, state, want_q=True, want_actions=True)
    import os
    os.environ["SWARM_STEP_GRID"] = "1"
    out = ops.rollout_large(cfg, w, state, 2)
    os.environ.pop("SWARM_STEP_GRID")
    torch.cuda.synchronize()
    print("large", N, B, gm, float(out["returns"].sum()))
z = np.load('tests/golden/flocking_models.npz')
sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('0/')}
spec = ops.stack_spec(3, 8, 7)
ws = ops.pack_stack_weights(sd, spec, dev)
for N, B, gm in ((12, 37, L.GRAPH_KNN), (9, 20, L.GRAPH_COMPLETE), (100, 3, L.GRAPH_RADIUS), (16, 9, L.GRAPH_KNN)):
    cfg = ops.make_config(0, B, N, gm, 5, graph_radius=0.3)
    centers = (torch.tensor([0.6, -0.6]) + 0.1 * torch.randn(B, 2, generator=g)).to(dev)
    state = ops.reset_grid(cfg, centers)
    rs = ops.reward_spec(L.REWARD_FLOCKING, B, N)
    shaping = torch.zeros(B, N, 2, device=dev)
    ops.scenario_reward(rs, state, shaping, reset=True)
    out = ops.rollout_stack(cfg, spec, ws, state, 3, reward=rs, shaping=shaping)
    torch.cuda.synchronize()
    print("stack", N, B, gm, float(out["returns"].sum()))
table = torch.zeros(1 << 8, dtype=torch.int64, device=dev)
for N, B, k in ((12, 37, 5), (8, 30, 3), (16, 11, 6), (5, 9, 5)):
    cfg = ops.make_config(1, B, N, L.GRAPH_KNN, k)
    centers = (torch.tensor([0.6, -0.6]) + 0.1 * torch.randn(B, 2, generator=g)).to(dev)
    state = ops.reset_grid(cfg, centers)
    out = ops.rollout(cfg, w, state, 8, knn_memo=table)
    torch.cuda.synchronize()
    print("knn set", N, B, k, float(out["returns"].sum()), int((table != 0).sum()))
    table.zero_()
print('ok2')
