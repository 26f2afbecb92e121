"""Warm per-kernel durations of the train tick parts (CUDA events over back-to-back launches inside a CUDA graph,
so launch gaps are minimal): python scripts/time_tick_parts.py [G]"""
import sys, torch
sys.path.insert(0, '.')
import numpy as np
import swarm_b200 as sb
from swarm_b200 import ops
dev = torch.device('cuda:0')
B, N = 4096, 12
G = int(sys.argv[1]) if len(sys.argv) > 1 else 32
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
ops.rollout(cfg, w, state, 50, epsilon=0.3, replay=ring)
gcfg = ops.clone_config(cfg, num_envs=G)
idx = torch.randint(0, len(ring), (G,), device=dev)
grad = torch.empty_like(w); loss = torch.empty(1, device=dev)

def graph_time(fn, reps=50):
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph(); s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(gr, stream=s):
        for _ in range(reps): fn()
    gr.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

ws = {}
print(f'G={G}')
print('rollout tick (eps .3, push)  %.1f us' % graph_time(lambda: ops.rollout(cfg, w, state, 1, epsilon=0.3, replay=ring, returns=returns, hits=hits)))
print('rollout tick greedy no push  %.1f us' % graph_time(lambda: ops.rollout(cfg, w, state, 1, returns=returns, hits=hits)))
import ctypes as C
L = sb._lib
wb = int(L.lib().swarm_dqn_workspace_bytes(C.byref(gcfg), G)); wsb = torch.empty(wb, dtype=torch.uint8, device=dev); rs = ring.struct()
def dg():
    L.check(L.lib().swarm_dqn_grad(C.byref(gcfg), L.ptr(w), L.ptr(w_t), C.byref(rs), L.ptr(idx), G, 0.99, 1.0 / (G * N), L.ptr(grad), L.ptr(loss), None, L.ptr(wsb), wb, L.stream_ptr(dev)))
print('dqn_grad + reduce            %.1f us' % graph_time(dg))
print('adam_clip                    %.1f us' % graph_time(lambda: ops.adam_clip_step(w, grad, m, v, 5)))
print('forward only (gatq_forward)  %.1f us' % graph_time(lambda: ops.gatq_forward(cfg, w, state, want_q=False)))
