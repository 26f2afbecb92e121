"""Rollout throughput sweep over (B, N, graph, kernel variant).  usage: python scripts/sweep_rollout.py"""
import os, subprocess, sys, json
import numpy as np, torch
sys.path.insert(0, '.')
def run(B, N, mode, k, ticks=50, reps=5, scen=1):
    import swarm_b200 as sb
    from swarm_b200 import ops
    dev = torch.device('cuda:0')
    models = np.load('tests/golden/models.npz')
    pre = 'ObstacleAvoidance/0/' if scen == 1 else 'GoTo/0/'
    w = sb.pack_weights({k_[len(pre):]: torch.from_numpy(models[k_]) for k_ in models.files if k_.startswith(pre)}, dev)
    gm = sb._lib.GRAPH_KNN if mode == 'knn' else sb._lib.GRAPH_COMPLETE
    cfg = ops.make_config(scen, B, N, gm, k)
    g = torch.Generator().manual_seed(0)
    centers = (torch.tensor([0.6, -0.6]) + 0.1 * torch.randn(B, 2, generator=g)).to(dev)
    state = ops.reset_grid(cfg, centers)
    ret = torch.zeros(B, N, device=dev); hits = torch.zeros(B, dtype=torch.int32, device=dev)
    for _ in range(2):
        ops.reset_grid(cfg, centers, out=state); ops.rollout(cfg, w, state, ticks, returns=ret, hits=hits)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = 0
    for _ in range(reps):
        ops.reset_grid(cfg, centers, out=state)
        a.record(); ops.rollout(cfg, w, state, ticks, returns=ret, hits=hits); b.record(); torch.cuda.synchronize()
        ms += a.elapsed_time(b)
    return B * N * ticks * reps / (ms * 1e-3)
if __name__ == '__main__':
    if len(sys.argv) > 1:
        B, N, mode, k = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
        print(json.dumps({'B': B, 'N': N, 'graph': mode, 'k': k, 'tc': os.environ.get('SWARM_TC', '2'), 'agent_steps_per_s': run(B, N, mode, k)}))
    else:
        for B, N, mode, k in ((4096, 12, 'complete', 5), (65536, 12, 'complete', 5), (65536, 12, 'knn', 5), (65536, 5, 'complete', 5), (65536, 8, 'knn', 5), (4096, 12, 'knn', 5), (16384, 32, 'complete', 5)):
            for tc in ('0', '1', '2'):
                env = dict(os.environ, SWARM_TC=tc)
                r = subprocess.run([sys.executable, __file__, str(B), str(N), mode, str(k)], capture_output=True, text=True, env=env)
                print(r.stdout.strip() or r.stderr[-300:], flush=True)
