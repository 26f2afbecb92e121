import sys, numpy as np, torch, subprocess
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from oracle import swarm_oracle as so, batched_oracle as bo
from helpers import load_params
import test_gpu_parity as T
import torch.nn.functional as F
print(subprocess.run("lscpu | grep -i 'model name\\|^CPU(s)\\|Thread\\|Socket'", shell=True, capture_output=True, text=True).stdout)
print('capability', torch.backends.cpu.get_cpu_capability(), 'threads', torch.get_num_threads(), 'mkl', torch.backends.mkl.is_available(), 'mkldnn', torch.backends.mkldnn.is_available())
print(torch.__config__.parallel_info())
exp,scen,n,mode,k = "GoTo","go_to",5,'knn',5
pos, vel = T._random_states(scen, 200, n, seed=n, crowd=True)
p = load_params(exp, 0)
edges = bo.graph_edges(pos, mode, k); ei = bo.batch_edge_index(edges, n)
x = bo.node_features(pos, vel).reshape(-1,7)
p64 = {k_:v.double() for k_,v in p.items()}
def stages(x, p):
    W=p['conv1.lin.weight']
    h = F.linear(x, W)
    conv = so.gat_conv(x, ei, W, p['conv1.att_src'], p['conv1.att_dst'], p['conv1.bias'])
    t = torch.tanh(conv)
    l1 = F.linear(t, p['lin1.weight'], p['lin1.bias'])
    r = torch.relu(l1)
    q = F.linear(r, p['lin2.weight'], p['lin2.bias'])
    return dict(h=h, conv=conv, tanh=t, lin1=l1, q=q)
with torch.no_grad():
    s32 = stages(x, p); s64 = stages(x.double(), p64)
    for key in s32:
        a, b = s32[key].double(), s64[key]
        print(key, 'max abs err', f'{(a-b).abs().max().item():.3e}', 'rel to max', f'{((a-b).abs().max()/b.abs().max()).item():.3e}')
    # feed f64-exact inputs to each f32 stage separately
    t64 = s64['tanh'].float()
    l1 = F.linear(t64, p['lin1.weight'], p['lin1.bias'])
    print('lin1 alone (f32 on f64-rounded input): rel', f'{((l1.double()-s64["lin1"]).abs().max()/s64["lin1"].abs().max()).item():.3e}')
    l1b = (t64 @ p['lin1.weight'].t()) + p['lin1.bias']
    print('lin1 via matmul+add: rel', f'{((l1b.double()-s64["lin1"]).abs().max()/s64["lin1"].abs().max()).item():.3e}')
    r64 = torch.relu(s64['lin1']).float()
    q = F.linear(r64, p['lin2.weight'], p['lin2.bias'])
    print('lin2 alone: rel', f'{((q.double()-s64["q"]).abs().max()/s64["q"].abs().max()).item():.3e}')
    for nt in (1,):
        torch.set_num_threads(nt)
        s32b = stages(x, p)
        print('threads', nt, 'q rel', f'{((s32b["q"].double()-s64["q"]).abs().max()/s64["q"].abs().max()).item():.3e}')
    np.savez('gpurun_out/diag_q.npz', q32=s32['q'].numpy(), q64=s64['q'].numpy())
