"""Host-side logic that needs no GPU: containers, weight packing, the nn.Module seam's state-dict / RNG
contract, env sharding."""
import random

import numpy as np
import pytest
import torch

from helpers import load_params


@pytest.fixture(scope="module")
def sb():
    import swarm_b200
    return swarm_b200


def test_gcn_state_dict_contract(sb):
    """Key set / shapes of the shipped .pth files and loading them."""
    model = sb.GCN(input_dim=7, hidden_dim=32, output_dim=9)
    sd = model.state_dict()
    assert list(sd.keys()) == ["conv1.att_src", "conv1.att_dst", "conv1.bias", "conv1.lin.weight", "lin1.weight",
                               "lin1.bias", "lin2.weight", "lin2.bias"]
    assert [tuple(v.shape) for v in sd.values()] == [(1, 1, 32), (1, 1, 32), (32,), (32, 7), (32, 32), (32,), (9, 32), (9,)]
    assert sum(p.numel() for p in model.parameters()) == 1673
    for exp in ("GoTo", "ObstacleAvoidance"):
        for seed in range(10):
            model.load_state_dict(load_params(exp, seed))          # strict
    model.eval()
    with pytest.raises(NotImplementedError):
        sb.GCN(7, 8, 9)                                             # e.g. the older Flocking checkpoints' width


def test_gcn_init_consumes_rng_like_the_reference(sb):
    """1 865 uniform draws in torch_geometric's order (conv1.lin twice): same parameters as the oracle's
    restatement and the same generator state afterwards (the golden evaluation recipe depends on it)."""
    from oracle.dqn_oracle import OracleGCN
    torch.manual_seed(6967)
    ours = sb.GCN(7, 32, 9)
    after_ours = torch.rand(3)
    torch.manual_seed(6967)
    ref = OracleGCN(7, 32, 9)
    after_ref = torch.rand(3)
    for (k1, v1), (k2, v2) in zip(ours.state_dict().items(), ref.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    assert torch.equal(after_ours, after_ref)
    assert torch.count_nonzero(ours.conv1.bias) == 0


def test_pack_unpack_weights_roundtrip(sb):
    params = load_params("ObstacleAvoidance", 3)
    w = sb.pack_weights(params)
    assert w.shape == (1673,) and w.dtype == torch.float32
    assert torch.equal(w[:224].reshape(32, 7), params["conv1.lin.weight"])
    assert torch.equal(w[1664:], params["lin2.bias"])
    back = sb.unpack_weights(w)
    for k, v in params.items():
        assert torch.equal(back[k], v)
    model = sb.GCN()
    model.load_state_dict(params)
    assert torch.equal(model.packed_weights().detach(), w)
    bad = dict(params)
    bad["lin1.weight"] = torch.zeros(8, 8)
    with pytest.raises(ValueError):
        sb.pack_weights(bad)


def test_data_and_batch_containers(sb):
    """Batch.from_data_list == torch_geometric's collation as the reference uses it (train:45)."""
    from oracle import swarm_oracle as so
    xs = [torch.randn(n, 7) for n in (5, 3, 12)]
    eis = [so.graph_complete(5), so.graph_complete(3), so.graph_complete(12)]
    b = sb.Batch.from_data_list([sb.Data(x=x, edge_index=e) for x, e in zip(xs, eis)])
    x_ref, ei_ref = so.batch_graphs(xs, eis)
    assert torch.equal(b.x, x_ref) and torch.equal(b.edge_index, ei_ref)
    assert b.num_graphs == 3 and b.num_nodes == 20 and b.ptr.tolist() == [0, 5, 8, 20]
    assert b.batch.tolist() == [0] * 5 + [1] * 3 + [2] * 12
    with pytest.raises(ValueError):
        sb.Batch.from_data_list([])


def test_env_sharding(sb):
    from swarm_b200 import parallel
    for total, world in ((4096, 8), (10, 3), (7, 7), (65536, 4)):
        shards = [parallel.shard_envs(total, r, world) for r in range(world)]
        assert sum(s.count for s in shards) == total
        assert shards[0].offset == 0
        for a, b in zip(shards, shards[1:]):
            assert a.offset + a.count == b.offset and 0 <= a.count - b.count <= 1
    with pytest.raises(ValueError):
        parallel.shard_envs(2, 0, 3)
    with pytest.raises(ValueError):
        parallel.shard_envs(8, 8, 8)


def test_replay_sample_indices_follow_python_random(sb):
    """GraphReplayBuffer.sample draws like random.sample(self.buffer, k): a function of (len, k) only."""
    buf = sb.GraphReplayBuffer(100)

    class _Ring:
        size = 57

        def __len__(self):
            return self.size
    buf.ring = _Ring()
    random.seed(5)
    ours = buf.sample_indices(32)
    random.seed(5)
    population = [object() for _ in range(57)]
    ref = random.sample(population, 32)
    assert [population.index(o) for o in ref] == ours


def test_module_name_shim_exposes_the_vmas_and_pyg_surface():
    """shim/: `vmas` / `torch_geometric` module names resolve to the swarm_b200 classes with vmas' constructor
    signatures (no GPU needed to import or to build entities; worlds themselves are CUDA-only)."""
    import importlib
    import os
    import sys
    import pytest
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.path.join(root, "shim"), root):
        if p not in sys.path:
            sys.path.insert(0, p)
    vmas = importlib.import_module("vmas")
    core = importlib.import_module("vmas.simulator.core")
    scenario = importlib.import_module("vmas.simulator.scenario")
    utils = importlib.import_module("vmas.simulator.utils")
    pyg_nn = importlib.import_module("torch_geometric.nn")
    pyg_data = importlib.import_module("torch_geometric.data")
    import swarm_b200 as sb
    assert vmas.make_env is sb.make_env and scenario.BaseScenario is sb.BaseScenario
    assert core.World is sb.World and core.Agent is sb.Agent and core.Landmark is sb.Landmark and core.Sphere is sb.Sphere
    assert pyg_nn.GATConv is sb.GATConv and pyg_data.Data is sb.Data and pyg_data.Batch is sb.Batch
    assert utils.Color.GREEN is sb.Color.GREEN
    # entity constructors as the reference scenario files call them (oa:29-56)
    goal = core.Landmark(name="goal", collide=False, color=utils.Color.BLACK)
    agent = core.Agent(name="agent0", collide=True, color=utils.Color.GREEN, render_action=True)
    assert goal.shape.radius == 0.05 and agent.shape.radius == 0.05 and agent.movable and not goal.movable
    agent.pos_rew = torch.zeros(1)                      # scenarios hang their own attributes on entities
    with pytest.raises(AssertionError):
        core.Sphere(radius=0.0)
    with pytest.raises(NotImplementedError):
        core.Agent(name="a", u_multiplier=2.0)
    with pytest.raises(sb.SwarmError, match="CUDA"):
        core.World(1, "cpu")
    # the built-in scenarios are BaseScenario subclasses with the reference's method set
    for cls in (sb.GoToPositionScenario, sb.ObstacleAvoidanceScenario):
        for m in ("make_world", "reset_world_at", "observation", "reward", "done", "info", "average_distance_to_goal",
                  "average_distance_to_obstacles", "obstacles_hits"):
            assert callable(getattr(cls, m)), (cls.__name__, m)


def test_flocking_and_cohesion_scenarios_have_no_cpu_fallback():
    """The two remaining reference scenarios are exported with the reference's kwargs; their worlds live on a CUDA
    device like every other world of this package."""
    import swarm_b200 as sb
    for cls in (sb.FlockingScenario, sb.CohesionScenario):
        sc = cls()
        with pytest.raises(sb.SwarmError):
            sc.env_make_world(2, "cpu", n_agents=3)
        assert sc.n_agents == 3 and sc.desired_distance == 0.15 and sc.min_collision_distance == 0.005
    assert len(sb.CohesionScenario._START) == 9                      # cohesion:46-56
