"""World step of large swarms, grid vs full sweep: python scripts/c4_step.py"""
import os, sys, json
import torch
sys.path.insert(0, '.')
import swarm_b200 as sb
from swarm_b200 import ops
dev = torch.device('cuda:0'); L = sb._lib
res = {}
for N, B in ((1024, 1024), (4096, 256), (256, 4096)):
    cfg = ops.make_config(L.SCENARIO_OBSTACLE_AVOIDANCE, B, N)
    g = torch.Generator().manual_seed(9)
    centers = (torch.tensor([0.6, -0.6]) + 0.1 * torch.randn(B, 2, generator=g)).to(dev)
    state = ops.reset_grid(cfg, centers)
    actions = torch.randint(0, 9, (B, N), generator=g).to(torch.int32).to(dev)
    for mode in ("grid", "sweep"):
        if mode == "sweep": os.environ["SWARM_STEP_GRID"] = "0"
        else: os.environ.pop("SWARM_STEP_GRID", None)
        out = torch.empty_like(state)
        ops.sim_step(cfg, state, actions, state_out=out, want_obs=False)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): ops.sim_step(cfg, state, actions, state_out=out, want_obs=False)
        b.record(); torch.cuda.synchronize()
        res[f"{N}x{B}_{mode}_ms"] = a.elapsed_time(b) / 10
    os.environ.pop("SWARM_STEP_GRID", None)
    print(N, B, {k: v for k, v in res.items() if k.startswith(f"{N}x{B}")}, flush=True)
json.dump(res, open('gpurun_out/r2_c4_step.json', 'w'), indent=1)
