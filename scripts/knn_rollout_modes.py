"""C2 / C3 kNN rollouts with the row variants: python scripts/knn_rollout_modes.py"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, '.')
import swarm_b200 as sb
from swarm_b200 import ops
dev = torch.device('cuda:0')
L = sb._lib
models = np.load('tests/golden/models.npz')
def weights(pre):
    return sb.pack_weights({k[len(pre):]: torch.from_numpy(models[k]) for k in models.files if k.startswith(pre)}, dev)
w_oa, w_goto = weights('ObstacleAvoidance/0/'), weights('GoTo/0/')
res = {}
def rate(scen, B, N, graph, K, T, reps=3, fresh=False, **kw):
    cfg = ops.make_config(scen, B, N, graph, K)
    g = torch.Generator().manual_seed(7)
    base = torch.tensor([0.6, -0.6]) if scen == L.SCENARIO_OBSTACLE_AVOIDANCE else torch.tensor([0.9, -0.9])
    centers = (base + 0.1 * torch.randn(B, 2, generator=g)).to(dev)
    w = w_oa if scen == L.SCENARIO_OBSTACLE_AVOIDANCE else w_goto
    state = ops.reset_grid(cfg, centers)
    ops.rollout(cfg, w, state, T, **kw)
    ms = 0.0
    for _ in range(reps):
        ops.reset_grid(cfg, centers, out=state)
        if fresh: ops.knn_memo_table(dev, N, K, fresh=True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.rollout(cfg, w, state, T, **kw); b.record(); torch.cuda.synchronize()
        ms += a.elapsed_time(b)
    return B * N * T * reps / (ms * 1e-3)
OA, GOTO = L.SCENARIO_OBSTACLE_AVOIDANCE, L.SCENARIO_GOTO
for name, scen, B, N, T in (("c2", OA, 4096, 12, 100), ("c3_N12", GOTO, 65536, 12, 50), ("c3_N5", GOTO, 65536, 5, 50), ("c3_N8", GOTO, 65536, 8, 50)):
    r = {"complete": rate(scen, B, N, L.GRAPH_COMPLETE, 5, T)}
    os.environ["SWARM_KNN_ORDERED"] = "1"
    r["knn_ordered"] = rate(scen, B, N, L.GRAPH_KNN, 5, T, knn_memo=None)
    del os.environ["SWARM_KNN_ORDERED"]
    r["knn_set_no_table"] = rate(scen, B, N, L.GRAPH_KNN, 5, T, knn_memo=None)
    r["knn_set_cold_table"] = rate(scen, B, N, L.GRAPH_KNN, 5, T, fresh=True)
    r["knn_set_warm_table"] = rate(scen, B, N, L.GRAPH_KNN, 5, T)
    r["ratio_cold_vs_complete"] = r["knn_set_cold_table"] / r["complete"]
    res[name] = r
    print(name, {k: (f"{v:.3e}" if v > 10 else round(v, 3)) for k, v in r.items()}, flush=True)
json.dump(res, open('gpurun_out/r2_knn_rollout_modes.json', 'w'), indent=1)
