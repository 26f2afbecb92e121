"""swarm_scenario_reward (Flocking / Cohesion) against the measured HBM peak at a size larger than L2.
usage: python scripts/bench_reward.py [out.json]"""
import json, sys
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

import ctypes as C
L = sb._lib
res = []
for N in (5, 12):
    B = (1 << 22) if N == 12 else (1 << 23)
    g = torch.Generator().manual_seed(0)
    centers = (torch.tensor([-1.6, 1.6]) + 0.4 * torch.randn(B, 2, generator=g)).to(dev)
    cfg = ops.make_config(0, B, N)
    state = ops.reset_grid(cfg, centers)
    state[:, :, :2] += 0.02 * torch.randn(B, N, 2, device=dev)
    shaping = torch.zeros(B, N, 2, device=dev)
    reward = torch.zeros(B, device=dev)
    fl = ops.reward_spec(L.REWARD_FLOCKING, B, N)
    ops.scenario_reward(fl, state, shaping, reset=True)
    def flock():
        L.check(L.lib().swarm_scenario_reward(C.byref(fl), L.ptr(state), L.ptr(shaping), L.ptr(reward), None, L.stream_ptr(dev)))
    ms = timeit(flock)
    by = B * (N * 32.0 + 4)
    r = {'kernel': f'scenario_reward[flocking N={N}]', 'ms': ms, 'algorithmic_bytes': by, 'GBps': by / ms / 1e6,
         'frac_of_measured_hbm_peak': by / ms / 1e6 / PEAK, 'agents_per_s': B * N / (ms * 1e-3),
         'note': 'read state 16 B + shaping 8 B, write shaping 8 B per agent, reward 4 B per env'}
    res.append(r); print(json.dumps(r), flush=True)
    co = ops.reward_spec(L.REWARD_COHESION, B, N)
    rew2 = torch.zeros(B, N, device=dev)
    def coh():
        L.check(L.lib().swarm_scenario_reward(C.byref(co), L.ptr(state), None, L.ptr(rew2), None, L.stream_ptr(dev)))
    ms = timeit(coh)
    by = B * N * 20.0
    r = {'kernel': f'scenario_reward[cohesion N={N}]', 'ms': ms, 'algorithmic_bytes': by, 'GBps': by / ms / 1e6,
         'frac_of_measured_hbm_peak': by / ms / 1e6 / PEAK, 'agents_per_s': B * N / (ms * 1e-3),
         'note': 'read state 16 B, write reward 4 B per agent'}
    res.append(r); print(json.dumps(r), flush=True)
    del state, shaping, reward, rew2, centers
    torch.cuda.empty_cache()
if len(sys.argv) > 1:
    json.dump({'peak_hbm_gbs': PEAK, 'results': res}, open(sys.argv[1], 'w'), indent=1)
