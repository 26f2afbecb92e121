"""Pins the CPU oracle against the reference's own shipped artefacts (tests/golden/*.npz, converted from
/root/reference/data by tests/golden/make_golden.py).  This is a subset sized for the default CPU run; the
complete sweep (1 280 evaluation episodes, 20 training runs) is tests/golden/verify_full.py -- result recorded
in DESIGN.md: GoTo 640/640 and ObstacleAvoidance 636/640 episodes bit-identical, 17/20 first training rows
exact."""
import numpy as np
import pytest
import torch

from helpers import eval_centers, golden_eval, load_params, npz
from oracle import batched_oracle as bo
from oracle import swarm_oracle as so
from oracle.dqn_oracle import OracleGCN, OracleTrainer


def _run_eval(exp, model, n):
    scen = so.GOTO if exp == "go_to" else so.OBSTACLE_AVOIDANCE
    T = 50 if exp == "go_to" else 100
    torch.manual_seed(6967)
    w = so.OracleWorld(scen, n, random=True)
    w.reset()
    OracleGCN(7, 32, 9)              # the reference builds the model after the env: 1 865 RNG draws
    return so.run_evaluation(w, load_params(exp, model), 8, T, "knn", 5)


@pytest.mark.parametrize("exp,model,n", [("go_to", 0, 5), ("go_to", 7, 12), ("go_to", 3, 8),
                                          ("obstacle_avoidance", 0, 5), ("obstacle_avoidance", 2, 12),
                                          ("obstacle_avoidance", 9, 9)])
def test_evaluation_goldens_bit_exact(exp, model, n):
    """data/test_stats/{exp}/seed_{model}/agents_{n}: positions, mean goal distance, hits, result.csv."""
    out = _run_eval(exp, model, n)
    gold = golden_eval(exp, model, n)
    pos = np.stack([np.stack([np.array(out["pos_x"][e], dtype=np.float32), np.array(out["pos_y"][e], dtype=np.float32)], -1)
                    for e in range(8)])
    assert np.array_equal(pos, gold["pos"]), "per-tick positions differ from the reference's CSVs"
    assert np.array_equal(np.array(out["distance"], dtype=np.float32), gold["dist"])
    assert np.array_equal(np.array(out["hits"], dtype=np.float32), gold["hits"])
    res = gold["result"]
    assert np.allclose(np.array(out["reward"]), res[:, 0], rtol=1e-6, atol=0)
    assert np.array_equal(np.array(out["collisions"], dtype=np.float64), res[:, 1])
    assert np.array_equal(np.array(out["distance_end"], dtype=np.float32), res[:, 2].astype(np.float32))
    assert np.array_equal(np.array(out["distance_beginning"], dtype=np.float32), res[:, 3].astype(np.float32))


def test_known_answers_from_survey():
    """RNG recipe, start grid, kNN edge order, Q-values, first step."""
    c = eval_centers("go_to", 5, episodes=1)[0]
    assert c.tolist() == [1.0912246704101562, -1.1671851873397827]
    grid = so.generate_grid(c, 5)
    exp = torch.tensor([[0.941224694, -1.242185235], [1.091224670, -1.242185235], [1.241224647, -1.242185235],
                        [0.941224694, -1.092185140], [1.091224670, -1.092185140]])
    assert torch.equal(grid, exp)
    w = so.OracleWorld(so.GOTO, 5)
    w.reset(c)
    x = so.node_features(w.observations())
    ei = so.graph_knn(x, 5)
    assert ei.shape == (2, 51) and ei[:, :10].t().tolist() == [[0, 0], [0, 0], [0, 1], [1, 0], [0, 3], [3, 0], [0, 4], [4, 0], [0, 2], [2, 0]]
    q = so.gatq_forward(load_params("go_to", 0), x, ei)
    assert abs(q[0, 0].item() - (-480.535125732)) < 1e-3 and torch.argmax(q, dim=1).tolist() == [5] * 5
    # D.5 contact force
    f = so.constraint_force(torch.tensor([[0.0, 0.0]]), torch.tensor([[0.03, 0.04]]))
    assert torch.allclose(f, torch.tensor([[-3.0, -4.0]]), rtol=1e-6)
    f = so.constraint_force(torch.tensor([[0.0, 0.0]]), torch.tensor([[0.0600001, 0.08]]))
    assert f.abs().max().item() == 0.0


@pytest.mark.parametrize("scenario,mode", [(so.GOTO, "knn"), (so.OBSTACLE_AVOIDANCE, "complete"), (so.OBSTACLE_AVOIDANCE, "knn")])
def test_batched_oracle_equals_single_env_oracle(scenario, mode):
    """'B envs' == B independent copies of the num_envs = 1 reference, bit for bit (incl. contacts)."""
    exp = "GoTo" if scenario == so.GOTO else "ObstacleAvoidance"
    params = load_params(exp, 4)
    N, B, T = 7, 5, 30
    g = torch.Generator().manual_seed(3)
    centers = torch.stack([so.draw_center(scenario, True, g) for _ in range(B)])
    centers[0] = torch.tensor([-0.1, 0.1]) + 0.05        # one swarm on the obstacle / crowded
    pos, vel = bo.reset_grid(scenario, centers, N)
    tr = bo.rollout(scenario, params, pos, vel, T, mode, 5)
    for b in range(B):
        w = so.OracleWorld(scenario, N, random=True)
        w.reset(centers[b])
        for t in range(T):
            x = so.node_features(w.observations())
            ei = so.graph_knn(x, 5) if mode == "knn" else so.graph_complete(N)
            with torch.no_grad():
                q = so.gatq_forward(params, x, ei)
            r = w.step(torch.argmax(q, dim=1))
            P, V = w.state()
            assert torch.equal(ei, tr["edges"][t, b]) and torch.equal(q, tr["q"][t, b])
            assert torch.equal(P, tr["pos"][t, b]) and torch.equal(V, tr["vel"][t, b]) and torch.equal(r, tr["rewards"][t, b])
            assert int(w.obstacles_hits()) == int(((tr["flags"][t, b] & 2) != 0).sum())


@pytest.mark.parametrize("exp,seed", [("ObstacleAvoidance", 0), ("GoTo", 0)])
def test_training_goldens(exp, seed):
    """data/stats/experiment_{exp}-seed_{seed}.csv: `Loss` row r = mean loss of episode r (train:231 quirk)."""
    stats = npz("train_stats.npz")[f"{exp}/{seed}"]
    tr = OracleTrainer(exp, seed, n_agents=5, max_steps=100)
    tr.train(2)
    for e in range(2):
        assert abs(tr.episode_losses[e] - stats[e, 2]) <= 1e-6 * abs(stats[e, 2])
    if exp == "ObstacleAvoidance":
        assert abs(tr.episode_returns[0] - (-44.196868896)) < 1e-6
