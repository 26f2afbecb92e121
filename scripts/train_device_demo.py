"""BASELINE config 5: end-to-end DQN training, env-sharded over the GPUs of one box, whole run on the device
(DQNTrainer.train_model_device: one CUDA-graph launch per episode; gradient exchange fused into clip + Adam).

  python scripts/train_device_demo.py [episodes]                                   # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
         scripts/train_device_demo.py [episodes]
"""
import json, os, sys, time
import torch
sys.path.insert(0, '.')
import swarm_b200 as sb
from swarm_b200 import parallel

episodes = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rank, world, local = parallel.init_from_env()
torch.cuda.set_device(local)
dev = f"cuda:{local}"
B, N, T, G = 4096, 12, 100, 32
sb.set_seed(0)
env = sb.make_env(sb.ObstacleAvoidanceScenario(), num_envs=B, device=dev, continuous_actions=False, max_steps=T,
                  dict_spaces=True, seed=0, n_agents=N, random=True)
trainer = sb.DQNTrainer(env, 0, "/tmp/swarm_models", "/tmp/swarm_stats", "ObstacleAvoidance", replay_capacity=1 << 20)
cfg = {"episodes": episodes, "epsilon": 0.99, "epsilon_decay": 0.05, "min_epsilon": 0.05, "graphs_per_update": G,
       "update_target_every": 200, "env_offset": rank * B}
torch.cuda.synchronize()
t0 = time.perf_counter()
stats = trainer.train_model_device(cfg)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
# every rank must hold the same weights
ok = True
if world > 1:
    ws = [torch.empty_like(trainer.w) for _ in range(world)]
    torch.distributed.all_gather(ws, trainer.w)
    ok = all(torch.equal(ws[0], w) for w in ws)
if rank == 0:
    print(json.dumps({"gpus": world, "episodes": episodes, "ticks": episodes * T, "wall_s": dt,
                      "updates_per_s": episodes * T / dt, "agent_steps_per_s": world * B * N * T * episodes / dt,
                      "weights_identical_across_ranks": ok,
                      "first_rows": stats[:2].tolist(), "last_rows": stats[-2:].tolist()}))
if world > 1:
    torch.distributed.destroy_process_group()
