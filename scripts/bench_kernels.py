"""Stand-alone (memory-bound) kernels against the measured HBM peak: sim_step, graph_build, reset_grid,
replay push / gather at sizes larger than L2 (126 MB).  usage: python scripts/bench_kernels.py [out.json]"""
import json, os, sys
import torch
sys.path.insert(0, '.')
import swarm_b200 as sb
from swarm_b200 import ops
dev = torch.device('cuda:0')
PEAK = 6544.3
try:
    PEAK = float(json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'])
except Exception:
    pass

def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

res = []
def report(name, ms, bytes_alg, note=''):
    gbs = bytes_alg / (ms * 1e-3) / 1e9
    r = {'kernel': name, 'ms': ms, 'algorithmic_bytes': bytes_alg, 'GBps': gbs, 'frac_of_measured_hbm_peak': gbs / PEAK, 'note': note}
    res.append(r)
    print(json.dumps(r), flush=True)

N = 12
B = 1 << 22                      # 4 Mi envs x 12 agents = 50 M agents, state 805 MB >> L2
g = torch.Generator().manual_seed(0)
centers = (torch.tensor([0.6, -0.6]) + 0.1 * torch.randn(B, 2, generator=g)).to(dev)
for scen, sname in ((1, 'obstacle_avoidance'), (0, 'go_to')):
    cfg = ops.make_config(scen, B, N)
    state = ops.reset_grid(cfg, centers)
    out_state = torch.empty_like(state)
    actions = torch.randint(0, 9, (B, N), device=dev, dtype=torch.int32)
    rewards = torch.empty(B, N, device=dev); flags = torch.empty(B, N, dtype=torch.uint8, device=dev)
    import ctypes as C
    L = sb._lib
    def step_min():
        L.check(L.lib().swarm_sim_step(C.byref(cfg), L.ptr(state), L.ptr(actions), L.ptr(out_state), L.ptr(rewards), None, None, None, None, L.stream_ptr(dev)))
    report(f'sim_step[{sname}] state+action->state+reward', timeit(step_min), B * N * 40.0, 'SURVEY 8d: 40 B/agent-step')
    obs = torch.empty(B, N, 6, device=dev); dist = torch.empty(B, N, 2, device=dev)
    def step_full():
        L.check(L.lib().swarm_sim_step(C.byref(cfg), L.ptr(state), L.ptr(actions), L.ptr(out_state), L.ptr(rewards), L.ptr(flags), None, L.ptr(obs), L.ptr(dist), L.stream_ptr(dev)))
    report(f'sim_step[{sname}] + obs + dist + flags', timeit(step_full), B * N * (40.0 + 24 + 8 + 1), '73 B/agent-step')
    del obs, dist
    if scen == 1:
        report('reset_grid', timeit(lambda: ops.reset_grid(cfg, centers, out=state)), B * (8.0 + N * 16), '8 B/env + 16 B/agent')
    del state, out_state, actions, rewards, flags
torch.cuda.empty_cache()

B = 1 << 21
centers = centers[:B].contiguous()
for mode, k in (('knn', 5), ('knn', 10), ('complete', 0)):
    gm = L.GRAPH_KNN if mode == 'knn' else L.GRAPH_COMPLETE
    cfg = ops.make_config(0, B, N, gm, max(k, 1))
    state = ops.reset_grid(cfg, centers)
    state[:, :, :2] += 0.01 * torch.randn(B, N, 2, device=dev)      # break the grid ties (fast path) ...
    E = ops.edges_per_env(cfg)
    edges = torch.empty(B, 2, E, dtype=torch.int32, device=dev)
    def gb():
        L.check(L.lib().swarm_graph_build(C.byref(cfg), L.ptr(state), L.ptr(edges), None, L.stream_ptr(dev)))
    report(f'graph_build[{mode} k={k}] perturbed', timeit(gb, reps=5), B * (N * 16.0 + 2 * E * 4), 'read state 16 B/agent, write 8E B/env')
    if mode == 'knn':
        state = ops.reset_grid(cfg, centers)                           # ... and the tie-heavy exact grids (emulation path)
        report(f'graph_build[{mode} k={k}] exact grid (ties)', timeit(gb, reps=5), B * (N * 16.0 + 2 * E * 4))
    del edges, state
torch.cuda.empty_cache()

# replay ring
B = 1 << 20
cfg = ops.make_config(1, B, N)
ring = ops.ReplayRing(B * 2, N, dev)
state = torch.randn(B, N, 4, device=dev); nxt = torch.randn(B, N, 4, device=dev)
actions = torch.randint(0, 9, (B, N), device=dev, dtype=torch.int32); rewards = torch.randn(B, N, device=dev)
report('replay_push', timeit(lambda: ops.replay_push(cfg, ring, state, actions, rewards, nxt)), B * N * (16 + 16 + 4 + 4 + 37.0), 'read 40 B + write 37 B per agent')
idx = torch.randint(0, 2 * B, (B,), device=dev, dtype=torch.int64)
report('replay_gather (random slots)', timeit(lambda: ops.replay_gather(ring, idx)), B * N * (37 + 40.0) + B * 8, 'read 37 B + write 40 B per agent')
if len(sys.argv) > 1:
    json.dump({'peak_hbm_gbs': PEAK, 'results': res}, open(sys.argv[1], 'w'), indent=1)
