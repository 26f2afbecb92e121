"""Rollout rate of the shipped three-layer Flocking checkpoints through swarm_rollout_stack: python scripts/stack_rate.py"""
import sys, json, os
import numpy as np, torch
sys.path.insert(0, '.')
import swarm_b200 as sb
from swarm_b200 import ops
dev = torch.device('cuda:0'); L = sb._lib
z = np.load('tests/golden/flocking_models.npz')
sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('0/')}
spec = ops.stack_spec(3, 8, 7)
w = ops.pack_stack_weights(sd, spec, dev)
res = {}
for B, N, gm, name in ((4096, 12, L.GRAPH_COMPLETE, "complete"), (4096, 12, L.GRAPH_KNN, "knn_k5"), (65536, 12, L.GRAPH_COMPLETE, "complete")):
    cfg = ops.make_config(L.SCENARIO_GOTO, B, N, gm, 5)
    g = torch.Generator().manual_seed(0)
    centers = (torch.tensor([0.9, -0.9]) + 0.1 * torch.randn(B, 2, generator=g)).to(dev)
    rs = ops.reward_spec(L.REWARD_FLOCKING, B, N)
    for rname, reward in (("world", None), ("flocking", rs)):
        state = ops.reset_grid(cfg, centers)
        shaping = torch.zeros(B, N, 2, device=dev)
        ops.scenario_reward(rs, state, shaping, reset=True)
        ops.rollout_stack(cfg, spec, w, state, 5, reward=reward, shaping=shaping)
        T = 50
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.rollout_stack(cfg, spec, w, state, T, reward=reward, shaping=shaping); b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / T
        res[f"{B}x{N}_{name}_{rname}"] = {"us_per_tick": us, "agent_steps_per_s": B * N / (us * 1e-6)}
        print(B, N, name, rname, res[f"{B}x{N}_{name}_{rname}"], flush=True)
    # forward alone
    state = ops.reset_grid(cfg, centers)
    ops.gatstack_forward(cfg, spec, w, state, want_q=False, want_actions=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): ops.gatstack_forward(cfg, spec, w, state, want_q=False, want_actions=True)
    b.record(); torch.cuda.synchronize()
    res[f"{B}x{N}_{name}_forward_us"] = a.elapsed_time(b) * 1e3 / 20
    print("forward us", res[f"{B}x{N}_{name}_forward_us"], flush=True)
json.dump(res, open('gpurun_out/r2_stack_rate.json', 'w'), indent=1)
