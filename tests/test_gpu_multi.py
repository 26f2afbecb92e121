"""Multi-GPU checks of the data-parallel train tick (skipped on single-GPU boxes; uses every visible GPU, 2 ... 8): the
gradient all-reduce fused into the clip + Adam kernel (one-shot push over NVLink peer memory) gives the same weights as
the NCCL all-reduce between the two phases -- bit for bit at 2 ranks, where any summation order is the same sum, within
float32 rounding above -- the ranks stay bit-identical, and the ticks replay from a CUDA graph.
Run with `gpurun --gpus 2` (or 4 / 8).  `bench.py --gpus N` repeats the bit-identity check on its DQN ticks."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    import swarm_b200 as sb
    from swarm_b200 import ops, parallel
    B, N, G = 256, 12, 32
    cfg = ops.make_config(sb._lib.SCENARIO_OBSTACLE_AVOIDANCE, B, N)
    models = np.load(os.path.join(ROOT, "tests", "golden", "models.npz"))
    pre = "ObstacleAvoidance/0/"
    w0 = sb.pack_weights({k[len(pre):]: torch.from_numpy(models[k]) for k in models.files if k.startswith(pre)}, dev)
    res = {}
    for mode in ("nccl", "peer", "peer_graph", "peer_one_call"):
        g = torch.Generator().manual_seed(rank)
        centers = (torch.tensor([0.6, -0.6]) + 0.1 * torch.randn(B, 2, generator=g)).to(dev)
        state = ops.reset_grid(cfg, centers)
        ring = ops.ReplayRing(4096, N, dev)
        w, w_t = w0.clone(), w0.clone()
        m, v = torch.zeros_like(w), torch.zeros_like(w)
        returns = torch.zeros(B, N, device=dev)
        hits = torch.zeros(B, dtype=torch.int32, device=dev)
        tt = ops.TrainTick(cfg, ring, graphs_per_update=G, update_target_every=4, loss_scale=parallel.global_loss_scale(G, N),
                           rng_seed=3, sample_seed=100 + rank, env_offset=rank * B)
        tt.load_cursor(0, 0, 0.3)
        if mode != "nccl":
            tt.peers = parallel.PeerExchange(dev)

        def tick():
            if mode == "peer_one_call":     # swarm_train_tick: reduction + exchange + clip + Adam as one cluster launch
                tt.tick(w, w_t, m, v, state, returns, hits)
                return
            tt.grad_phase(w, w_t, state, returns, hits)
            if mode == "nccl":
                dist.all_reduce(tt.grad_loss)
            tt.apply_phase(w, w_t, m, v)

        if mode == "peer_graph":
            for _ in range(3):
                tick()
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.graph(graph, stream=side):
                for _ in range(3):
                    tick()
            for _ in range(3):
                graph.replay()
        else:
            for _ in range(12):
                tick()
        torch.cuda.synchronize(dev)
        cur = tt.read_cursor()
        assert cur["tick"] == 12 and cur["opt_step"] == 12
        res[mode] = (w.clone(), w_t.clone(), tt.grad_loss.clone())
        gathered = [torch.empty_like(w) for _ in range(world)]
        dist.all_gather(gathered, w)
        assert all(torch.equal(gathered[0], x) for x in gathered), f"{mode}: ranks diverged"
    for other in ("peer_graph", "peer_one_call"):
        for x, y in zip(res["peer"], res[other]):
            assert torch.equal(x, y), f"peer vs {other} differ: {(x - y).abs().max().item():.3e}"
    for x, y in zip(res["nccl"], res["peer"]):
        if world == 2:
            assert torch.equal(x, y), f"nccl vs peer differ: {(x - y).abs().max().item():.3e}"
        else:       # NCCL's reduction order is its own; Adam's normalisation (an early step moves a weight by +-lr whatever
            #         the gradient's size) amplifies last-bit gradient differences, so this is a sanity bound only
            assert torch.allclose(x, y, rtol=1e-2, atol=5e-3), f"nccl vs peer differ: {(x - y).abs().max().item():.3e}"
    assert not torch.equal(res["nccl"][0], w0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        open(out, "w").write("ok")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least two GPUs")
def test_peer_allreduce_matches_nccl(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "done")
    world = min(torch.cuda.device_count(), 8)
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert open(out).read() == "ok"
