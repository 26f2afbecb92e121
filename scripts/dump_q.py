import sys, numpy as np, torch
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import swarm_b200 as sb
from oracle import batched_oracle as bo
from helpers import load_params
import test_gpu_parity as T
dev = torch.device('cuda:0')
out = {}
for exp,scen,n,mode,k,crowd in (("GoTo","go_to",5,'knn',5,True),("GoTo","go_to",12,'knn',5,True),("ObstacleAvoidance","obstacle_avoidance",12,'complete',0,False)):
    pos, vel = T._random_states(scen, 200, n, seed=n, crowd=crowd)
    for m in (0,3,7):
        p = load_params(exp, m)
        gm = sb._lib.GRAPH_KNN if mode=='knn' else sb._lib.GRAPH_COMPLETE
        cfg = sb.ops.make_config(T._scen_id(sb, scen), 200, n, gm, max(k,1))
        q, act = sb.ops.gatq_forward(cfg, sb.pack_weights(p, dev), T._pack_state(pos, vel).to(dev))
        out[f'{exp}_{n}_{mode}_{m}'] = q.cpu().numpy()
np.savez('gpurun_out/q_dump.npz', **out)
print('saved')
