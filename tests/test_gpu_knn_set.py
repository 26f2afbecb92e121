"""Order-free kNN rows of the tensor-core rollout (csrc/tile_device.cuh tile_knn_small_set_np): the set of each row's k
neighbours from the ranks alone, boundary ties through the per-thread memo, the global memo table and the libstdc++
emulation.  The ordered row code (torch.topk's output order, tie-exact, pinned by the golden sweeps) is the reference:
both feed the same multiplicities to the same arithmetic, so whole rollouts must agree bit for bit."""
import pytest
import torch

from helpers import load_params

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("scenario,N,K,B", [("obstacle_avoidance", 12, 5, 700), ("go_to", 5, 5, 64), ("go_to", 8, 3, 300),
                                           ("obstacle_avoidance", 9, 5, 200), ("go_to", 16, 6, 100),
                                           ("obstacle_avoidance", 12, 11, 64), ("go_to", 12, 1, 64), ("go_to", 7, 7, 50)])
def test_set_path_equals_ordered_path(scenario, N, K, B, monkeypatch):
    import swarm_b200 as sb
    ops, L = sb.ops, sb._lib
    dev = _dev()
    scen = L.SCENARIO_GOTO if scenario == "go_to" else L.SCENARIO_OBSTACLE_AVOIDANCE
    w = sb.pack_weights(load_params("GoTo" if scenario == "go_to" else "ObstacleAvoidance", 0), dev)
    cfg = ops.make_config(scen, B, N, L.GRAPH_KNN, K)
    g = torch.Generator().manual_seed(N * 100 + K)
    base = torch.tensor([1.5, -1.5]) + torch.tensor([-0.6, 0.6]) if scenario == "go_to" else torch.tensor([0.6, -0.6])
    centers = (base + 0.1 * torch.randn(B, 2, generator=g)).to(dev)
    centers[: B // 4] = centers[0]                              # a quarter of the envs share one exact lattice
    state0 = ops.reset_grid(cfg, centers)
    T = 60

    def run(**kw):
        out = ops.rollout(cfg, w, state0.clone(), T, **kw)
        return out["state"], out["returns"], out["hits"]

    monkeypatch.setenv("SWARM_KNN_ORDERED", "1")
    ref = run(knn_memo=None)
    monkeypatch.delenv("SWARM_KNN_ORDERED")
    table = torch.zeros(1 << 12, dtype=torch.int64, device=dev)          # small: conflict evictions happen
    for label, kw in (("no table", dict(knn_memo=None)), ("cold table", dict(knn_memo=table)),
                      ("warm table", dict(knn_memo=table)), ("cached table", dict())):
        got = run(**kw)
        for a, b, what in zip(got, ref, ("state", "returns", "hits")):
            assert torch.equal(a, b), f"{label}: {what} differs from the ordered path"
    if N <= 12 and 1 < K < N:
        used = int((table != 0).sum())
        assert used > 0, "the lattice starts must produce boundary ties"
        e = table[table != 0]
        assert bool(((e & 0xFFFF).to(torch.int32).cpu().apply_(lambda v: bin(v).count("1")) == K).all())
    else:
        assert N > 12 or int((table != 0).sum()) == 0


def test_set_path_trace_edges_still_ordered():
    """An edge trace needs torch.topk's output order: the rollout falls back to the ordered rows and the trace equals
    the stand-alone graph builder's edge list."""
    import swarm_b200 as sb
    ops, L = sb.ops, sb._lib
    dev = _dev()
    B, N, K = 50, 12, 5
    cfg = ops.make_config(L.SCENARIO_OBSTACLE_AVOIDANCE, B, N, L.GRAPH_KNN, K)
    centers = (torch.tensor([0.6, -0.6]) + 0.1 * torch.randn(B, 2, generator=torch.Generator().manual_seed(1))).to(dev)
    state0 = ops.reset_grid(cfg, centers)
    w = sb.pack_weights(load_params("ObstacleAvoidance", 0), dev)
    edges0, _ = ops.graph_build(cfg, state0)
    out = ops.rollout(cfg, w, state0.clone(), 2, trace=dict(edges=True))
    assert torch.equal(out["trace_edges"][0], edges0)
    plain = ops.rollout(cfg, w, state0.clone(), 2)
    assert torch.equal(plain["state"], out["state"])


@pytest.mark.parametrize("N", [2, 3, 5, 6, 9, 12, 13, 16])
def test_set_path_fuzz_on_tie_heavy_states(N, monkeypatch):
    """Positions snapped to a coarse lattice (many exactly equal distances, co-located agents -- zero distances that tie
    with the row's own entry), every k from 1 to N: rollouts with the order-free rows equal the ordered rows bit for
    bit, with no table, a tiny (evicting) table and a table shared across different start states."""
    import swarm_b200 as sb
    ops, L = sb.ops, sb._lib
    dev = _dev()
    w = sb.pack_weights(load_params("GoTo", 1), dev)
    B, T = 400, 6
    g = torch.Generator().manual_seed(N)
    pos = (torch.randn(B, N, 2, generator=g) * 0.25 / 0.05).round() * 0.05 + torch.tensor([0.9, -0.9])
    vel = (torch.randn(B, N, 2, generator=g) * 0.2 / 0.1).round() * 0.1
    state0 = torch.cat([pos, vel], 2).contiguous().to(dev)
    for K in sorted({1, 2, N // 2, N - 1, N} - {0}):
        cfg = ops.make_config(L.SCENARIO_GOTO, B, N, L.GRAPH_KNN, K)
        monkeypatch.setenv("SWARM_KNN_ORDERED", "1")
        ref = ops.rollout(cfg, w, state0.clone(), T, knn_memo=None)
        monkeypatch.delenv("SWARM_KNN_ORDERED")
        table = torch.zeros(64, dtype=torch.int64, device=dev)
        for kw in (dict(knn_memo=None), dict(knn_memo=table), dict(knn_memo=table)):
            got = ops.rollout(cfg, w, state0.clone(), T, **kw)
            assert torch.equal(got["state"], ref["state"]) and torch.equal(got["returns"], ref["returns"]), (N, K, kw.keys())
        # the same table, different states (other patterns hash into the same 64 slots)
        other = state0.roll(1, dims=0).contiguous()
        monkeypatch.setenv("SWARM_KNN_ORDERED", "1")
        ref2 = ops.rollout(cfg, w, other.clone(), T, knn_memo=None)
        monkeypatch.delenv("SWARM_KNN_ORDERED")
        got2 = ops.rollout(cfg, w, other.clone(), T, knn_memo=table)
        assert torch.equal(got2["state"], ref2["state"]), (N, K, "shared table")
