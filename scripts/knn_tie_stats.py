"""How often does torch.topk's tie handling matter in a C2 kNN rollout?  python scripts/knn_tie_stats.py
Per tick over B envs: fraction of rows (a) not strictly tie-free among the k+1 smallest (current shortcut test),
(b) whose k-smallest SET is ambiguous (a tie straddles the k boundary), and the hit rate of a memo of order patterns
(per thread = same agent slot previous tie row; global = any pattern seen in an earlier tick anywhere)."""
import sys, json
import numpy as np, torch
sys.path.insert(0, '.')
import swarm_b200 as sb
from swarm_b200 import ops
dev = torch.device('cuda:0')
L = sb._lib
models = np.load('tests/golden/models.npz'); pre = 'ObstacleAvoidance/0/'
w = sb.pack_weights({k[len(pre):]: torch.from_numpy(models[k]) for k in models.files if k.startswith(pre)}, dev)
B, N, K, T = 2048, 12, 5, 100
g = torch.Generator().manual_seed(0)
centers = (torch.tensor([0.6, -0.6]) + 0.1 * torch.randn(B, 2, generator=g)).to(dev)
cfg = ops.make_config(L.SCENARIO_OBSTACLE_AVOIDANCE, B, N, L.GRAPH_KNN, K)
state0 = ops.reset_grid(cfg, centers)
out = ops.rollout(cfg, w, state0.clone(), T, trace=dict(state=True))
states = torch.cat([state0[None], out["trace_state"][:-1]], 0)          # pre-step states of every tick [T,B,N,4]
seen = set()
last = {}
rows = []
for t in range(T):
    p = states[t, :, :, :2]
    d = torch.linalg.norm(p[:, None, :, :] - p[:, :, None, :], dim=-1)      # [B,N,N]  d[b,i,j]
    rank = (d[:, :, None, :] < d[:, :, :, None]).sum(-1)                    # rank[b,i,j] = #{l: d_l < d_j}
    srt = torch.sort(rank, dim=-1).values
    strict = (srt[..., :K + 1] == torch.arange(K + 1, device=dev)).all(-1)   # ranks 0..K all present
    amb = (rank < K).sum(-1) != K                                             # set ambiguous
    key = (rank.long() << (4 * torch.arange(N, device=dev))).sum(-1)         # nibble-packed order pattern
    keyc = key.cpu().numpy(); tie = (~strict).cpu().numpy(); ambc = amb.cpu().numpy()
    # memo simulations over rows needing the emulation (current rule: not strict; set rule: ambiguous)
    hit_thread = hit_global = n_tie = 0
    hit_thread_a = hit_global_a = n_amb = 0
    new = set()
    ks = keyc[tie]; idx = np.argwhere(tie)
    for (b, i), k_ in zip(idx, ks):
        n_tie += 1
        if last.get((b, i)) == k_: hit_thread += 1
        if k_ in seen: hit_global += 1
        last[(b, i)] = k_
        new.add(k_)
    seen |= new
    # per-CTA view: 10 envs per CTA; a CTA "pays" when any of its rows misses the per-thread memo
    rows.append(dict(tick=t, not_strict=float(tie.mean()), set_ambiguous=float(ambc.mean()), memo_thread=hit_thread / max(n_tie, 1),
                     memo_global=hit_global / max(n_tie, 1), distinct_patterns=len(seen)))
    if t % 10 == 0 or t == T - 1: print(rows[-1], flush=True)
json.dump(rows, open('gpurun_out/knn_tie_stats.json', 'w'))
