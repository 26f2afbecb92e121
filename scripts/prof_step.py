import sys, ctypes as C, torch
sys.path.insert(0, '.')
import swarm_b200 as sb
from swarm_b200 import ops
L = sb._lib
dev = torch.device('cuda:0')
B, N = 1 << 22, 12
cfg = ops.make_config(1, B, N)
g = torch.Generator().manual_seed(0)
centers = (torch.tensor([0.6, -0.6]) + 0.1 * torch.randn(B, 2, generator=g)).to(dev)
state = ops.reset_grid(cfg, centers)
out_state = torch.empty_like(state)
actions = torch.randint(0, 9, (B, N), device=dev, dtype=torch.int32)
rewards = torch.empty(B, N, device=dev)
for _ in range(4):
    L.check(L.lib().swarm_sim_step(C.byref(cfg), L.ptr(state), L.ptr(actions), L.ptr(out_state), L.ptr(rewards), None, None, None, None, L.stream_ptr(dev)))
torch.cuda.synchronize()
print('ok')
