"""world_size = 2 over gloo on the CPU: the env-sharded DQN update (local gradient scaled by the global
node count, summed over ranks, identical optimiser step) equals the single-process update on the full
batch; weights start from rank 0's broadcast and stay replicated."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _batch(seed, G, N):
    from oracle import batched_oracle as bo
    g = torch.Generator().manual_seed(seed)
    pos = torch.randn(G, N, 2, generator=g)
    vel = 0.2 * torch.randn(G, N, 2, generator=g)
    actions = torch.randint(0, 9, (G, N), generator=g)
    out = bo.step("go_to", pos, vel, actions)
    return pos, vel, actions, out["rewards"], out["pos"], out["vel"]


def _partial_grad(model, target, batch, loss_scale):
    """What swarm_dqn_grad returns for one shard: sum of squared TD errors * loss_scale, and its gradient."""
    from oracle import batched_oracle as bo
    pos, vel, actions, rewards, pos2, vel2 = batch
    G, N, _ = pos.shape
    x = bo.node_features(pos, vel).reshape(G * N, 7)
    x2 = bo.node_features(pos2, vel2).reshape(G * N, 7)
    ei = bo.batch_edge_index(bo.edges_complete(G, N), N)
    v = model(x, ei).gather(1, actions.reshape(-1, 1)).reshape(-1)
    y = rewards.reshape(-1) + 0.99 * target(x2, ei).max(dim=1)[0].detach()
    loss = ((v - y) ** 2).sum() * loss_scale
    model.zero_grad()
    loss.backward()
    return torch.cat([p.grad.reshape(-1) for p in model.parameters()]), loss.detach().reshape(1)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    from swarm_b200 import parallel
    from oracle.dqn_oracle import OracleGCN
    r, w, _ = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world) and parallel.world_size() == world
    G, N = 6, 5
    shard = parallel.shard_envs(world * G, rank, world)
    torch.manual_seed(100 + rank)                       # ranks start from different weights ...
    model, target = OracleGCN(), OracleGCN()
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    parallel.broadcast_weights(flat)                    # ... and adopt rank 0's
    off = 0
    for p in model.parameters():
        p.data.copy_(flat[off:off + p.numel()].reshape(p.shape))
        off += p.numel()
    tflat = torch.cat([p.detach().reshape(-1) for p in target.parameters()])
    parallel.broadcast_weights(tflat)
    off = 0
    for p in target.parameters():
        p.data.copy_(tflat[off:off + p.numel()].reshape(p.shape))
        off += p.numel()
    grad, loss = _partial_grad(model, target, _batch(7 + shard.rank, G, N), parallel.global_loss_scale(G, N))
    parallel.allreduce_gradient(grad, loss)
    # every rank must take the same "does this tick update?" decision: equal shards pass, unequal ones are refused on
    # EVERY rank (a one-sided error would leave the peers waiting in the exchange)
    parallel.require_equal_shards(64, 0, 0, 4096, 32)
    try:
        parallel.require_equal_shards(64 + rank, 0, 0, 4096, 32)
        refused = False
    except ValueError:
        refused = True
    ret[rank] = (flat.clone(), tflat.clone(), grad.clone(), loss.clone(), shard, refused)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_update_equals_full_batch_update():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    (w0, t0, g0, l0, s0, r0), (w1, t1, g1, l1, s1, r1) = ret[0], ret[1]
    assert r0 and r1, "unequal env shards must be refused on every rank"
    assert torch.equal(w0, w1) and torch.equal(t0, t1), "weights must be replicated after the broadcast"
    assert torch.equal(g0, g1) and torch.equal(l0, l1), "all ranks hold the same reduced gradient"
    assert (s0.offset, s0.count, s1.offset, s1.count) == (0, 6, 6, 6)
    # single-process reference on the concatenated batch
    torch.set_num_threads(1)
    from oracle.dqn_oracle import OracleGCN
    model, target = OracleGCN(), OracleGCN()
    for m_, flat in ((model, w0), (target, t0)):
        off = 0
        for p in m_.parameters():
            p.data.copy_(flat[off:off + p.numel()].reshape(p.shape))
            off += p.numel()
    b0, b1 = _batch(7, 6, 5), _batch(8, 6, 5)
    full = tuple(torch.cat([a, b]) for a, b in zip(b0, b1))
    grad, loss = _partial_grad(model, target, full, 1.0 / (12 * 5))
    assert torch.allclose(g0, grad, rtol=1e-5, atol=1e-6 * grad.abs().max().item())
    assert torch.allclose(l0, loss, rtol=1e-5)
    # and the mean-squared-error loss of the reference (train:122) is what the scaled sum computes
    from oracle import batched_oracle as bo
    pos, vel, actions, rewards, pos2, vel2 = full
    x = bo.node_features(pos, vel).reshape(-1, 7)
    x2 = bo.node_features(pos2, vel2).reshape(-1, 7)
    ei = bo.batch_edge_index(bo.edges_complete(12, 5), 5)
    v = model(x, ei).gather(1, actions.reshape(-1, 1))
    y = rewards.reshape(-1) + 0.99 * target(x2, ei).max(dim=1)[0].detach()
    assert torch.allclose(nn.MSELoss()(v, y.unsqueeze(1)).detach(), loss[0], rtol=1e-5)
