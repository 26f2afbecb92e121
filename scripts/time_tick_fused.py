"""A/B of the train tick as two library calls (four launches) and as swarm_train_tick (three launches for G <= 64 CTAs),
100 ticks per CUDA-graph replay: python scripts/time_tick_fused.py [G]"""
import json, os, sys, torch
sys.path.insert(0, '.')
import numpy as np
import swarm_b200 as sb
from swarm_b200 import ops
dev = torch.device('cuda:0')
B, N = 4096, 12
out = {}
for G in ([int(a) for a in sys.argv[1:]] or [32, 4096]):
    res = {}
    for mode in (("one_call", "two_calls") if os.environ.get("REVERSE") else ("two_calls", "one_call")):
        cfg = ops.make_config(1, B, N)
        g = torch.Generator().manual_seed(0)
        centers = (torch.tensor([0.6, -0.6]) + 0.1 * torch.randn(B, 2, generator=g)).to(dev)
        state = ops.reset_grid(cfg, centers)
        models = np.load('tests/golden/models.npz')
        pre = 'ObstacleAvoidance/0/'
        w = sb.pack_weights({k[len(pre):]: torch.from_numpy(models[k]) for k in models.files if k.startswith(pre)}, dev)
        w_t = w.clone(); m = torch.zeros_like(w); v = torch.zeros_like(w)
        ring = ops.ReplayRing(1 << 20, N, dev)
        returns = torch.zeros(B, N, device=dev); hits = torch.zeros(B, dtype=torch.int32, device=dev)
        tt = ops.TrainTick(cfg, ring, graphs_per_update=G, update_target_every=200)
        tt.load_cursor(0, 0, 0.3)

        def tick():
            if mode == "one_call":
                tt.tick(w, w_t, m, v, state, returns, hits)
            else:
                tt.grad_phase(w, w_t, state, returns, hits)
                tt.apply_phase(w, w_t, m, v)
        for _ in range(3):
            tick()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph(); s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.graph(gr, stream=s):
            for _ in range(100):
                tick()
        best = 1e9
        for _ in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / 100 * 1e3)
        res[mode] = {"us_per_tick": best, "updates_per_s": 1e6 / best, "w_sum": float(w.double().sum())}
    res["same_weights"] = res["two_calls"]["w_sum"] == res["one_call"]["w_sum"]
    out["G%d" % G] = res
print(json.dumps(out, indent=1))
