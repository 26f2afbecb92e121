"""C4 (1 024 agents x 1 024 envs) per-tick time with the three graphs: python scripts/c4_graphs.py"""
import sys, json
import numpy as np, torch
sys.path.insert(0, '.')
import swarm_b200 as sb
from swarm_b200 import ops
dev = torch.device('cuda:0')
L = sb._lib
models = np.load('tests/golden/models.npz'); pre = 'ObstacleAvoidance/0/'
w = sb.pack_weights({k[len(pre):]: torch.from_numpy(models[k]) for k in models.files if k.startswith(pre)}, dev)
g = torch.Generator().manual_seed(9)
res = {}
for N, B in ((1024, 1024), (4096, 256)):
    centers = (torch.tensor([0.6, -0.6]) + 0.1 * torch.randn(B, 2, generator=g)).to(dev)
    for name, gm in (("knn_k10", L.GRAPH_KNN), ("radius_r0.35", L.GRAPH_RADIUS), ("complete", L.GRAPH_COMPLETE)):
        cfg = ops.make_config(L.SCENARIO_OBSTACLE_AVOIDANCE, B, N, gm, 10, graph_radius=0.35)
        st = ops.reset_grid(cfg, centers)
        ops.rollout_large(cfg, w, st, 2)
        T = 10
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.rollout_large(cfg, w, st, T); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / T
        # forward alone
        if gm != L.GRAPH_KNN:
            ops.gatq_forward_large(cfg, w, st, want_q=False, want_actions=True)
            a.record()
            for _ in range(5): ops.gatq_forward_large(cfg, w, st, want_q=False, want_actions=True)
            b.record(); torch.cuda.synchronize()
            fwd = a.elapsed_time(b) / 5
        else:
            fwd = None
        res[f"{N}x{B}_{name}"] = {"ms_per_tick": ms, "agent_steps_per_s": B * N / (ms * 1e-3), "forward_ms": fwd}
        print(N, B, name, res[f"{N}x{B}_{name}"], flush=True)
    if N == 1024:
        cfg = ops.make_config(L.SCENARIO_OBSTACLE_AVOIDANCE, B, N, L.GRAPH_RADIUS, 10, graph_radius=0.35)
        st = ops.reset_grid(cfg, centers)
        rp, src = ops.graph_build_radius_csr(cfg, st)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): rp, src = ops.graph_build_radius_csr(cfg, st)
        b.record(); torch.cuda.synchronize()
        res["1024x1024_radius_csr_build"] = {"ms": a.elapsed_time(b) / 5, "edges": int(src.numel())}
        print(res["1024x1024_radius_csr_build"])
json.dump(res, open('gpurun_out/r2_c4_graphs.json', 'w'), indent=1)
