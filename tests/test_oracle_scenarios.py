"""Known-answer checks of the Flocking / Cohesion oracle (oracle/scenario_rewards_oracle.py) against values worked out
independently in float64 from the scenario source (flocking_scenario.py:93-171, cohesion_scenario.py:44-85).  The
reference ships no golden vectors for these two scenarios, so the restatement is pinned twice: to the formulas (the
hand computations below) and -- at the end of this file -- to the outputs of the reference's own scenario files,
executed unmodified on the CPU stand-in of vmas (tests/golden/reference_runs.npz, made by
tests/golden/make_reference_runs.py)."""
import math

import numpy as np
import pytest
import torch

from oracle import scenario_rewards_oracle as sro
from oracle import swarm_oracle as so


def _spacing64(p, i, others, desired=0.15, factor=10.0):
    return factor * np.mean([(math.dist(p[i], q) - desired) ** 2 for q in others])


def test_flocking_reset_memory_sees_unplaced_agents_at_origin():
    n = 5
    orc = sro.FlockingOracle(n)
    center = torch.tensor([0.3, -0.2])
    orc.reset(center)
    grid = so.generate_grid(center, n).double().numpy()
    goal = np.array(sro.GOAL_POS)
    for i in range(n):
        want_goal = 10.0 * math.dist(grid[i], goal)
        others = [grid[j] if j < i else np.zeros(2) for j in range(n) if j != i]
        want_sp = _spacing64(grid, i, others)
        assert abs(float(orc.previous_distance_to_goal[i]) - want_goal) < 1e-5
        assert abs(float(orc.previous_distance_to_agents[i]) - want_sp) < 1e-5 * max(1.0, want_sp)
    # the last agent saw everybody placed; the first saw everybody else at the origin
    assert abs(float(orc.previous_distance_to_agents[0]) - 10.0 * (math.hypot(*grid[0]) - 0.15) ** 2) < 1e-5


def test_flocking_reset_consumes_two_plus_2n_normals():
    n = 7
    torch.manual_seed(11)
    sro.FlockingOracle(n).reset()
    after = torch.rand(1)
    torch.manual_seed(11)
    torch.normal(mean=torch.tensor([-0.6, 0.6]), std=torch.tensor([0.4, 0.4]))
    for _ in range(2 * n):
        torch.normal(mean=torch.tensor([0.0]), std=torch.tensor([0.1]))
    assert torch.equal(after, torch.rand(1))


def test_flocking_collective_reward_known_answer():
    """Three agents at hand-placed positions: one on the goal, two nearly touching."""
    n = 3
    orc = sro.FlockingOracle(n)
    orc.reset(torch.tensor([0.0, 0.0]))
    prev_goal = [float(v) for v in orc.previous_distance_to_goal]
    prev_sp = [float(v) for v in orc.previous_distance_to_agents]
    pos = np.array([[-0.79, 0.81], [0.2, 0.2], [0.2, 0.3045]])           # |p1 - p2| = 0.1045 -> gap 0.0045 <= 0.005
    orc.world.set_state(torch.tensor(pos, dtype=torch.float32), torch.zeros(n, 2))
    got = float(orc.reward())
    goal = np.array(sro.GOAL_POS)
    want = 0.0
    for i in range(n):
        d = math.dist(pos[i], goal)
        term = prev_goal[i] - 10.0 * d + (50.0 if d < 0.05 else 0.0)
        term += -sum(1 for j in range(n) if j != i and math.dist(pos[i], pos[j]) - 0.1 <= 0.005)
        term += prev_sp[i] - _spacing64(pos, i, [pos[j] for j in range(n) if j != i])
        want += term
    assert abs(got - want) < 1e-4 * abs(want), (got, want)
    # agent 0 is on the goal (+50), agents 1 and 2 see each other too close (-1 each)
    assert float(orc.distance_to_goal[0]) < 0.05
    # the memory moved to the new state: an immediate second call only returns the bonus and the penalties
    again = float(orc.reward())
    assert abs(again - (50.0 - 2.0)) < 1e-4


def test_cohesion_known_answers():
    orc = sro.CohesionOracle(9)
    orc.reset()
    r = orc.reward().double().numpy()
    table = np.array(sro.COHESION_START)
    for i in range(9):
        gaps = [math.dist(table[i], table[j]) - 0.1 for j in range(9) if j != i]
        mn, mx = min(gaps), max(gaps)
        want = (0.0 if mn > 0.15 else math.exp(-mn / 0.15)) + (0.0 if mn < 0.15 else -(mx - 0.15))
        assert abs(r[i] - want) < 1e-6, (i, r[i], want)
    # two agents closer than sigma: only the collision factor (a positive number, as the reference has it) applies
    orc2 = sro.CohesionOracle(2)
    orc2.reset()
    orc2.world.set_state(torch.tensor([[0.0, 0.0], [0.2, 0.0]]), torch.zeros(2, 2))
    r2 = orc2.reward()
    assert abs(float(r2[0]) - math.exp(-(0.1 / 0.15))) < 1e-6 and float(r2[0]) == float(r2[1])
    obs = orc2.observations()
    assert obs.shape == (2, 4)


def test_cohesion_more_than_nine_agents_raises():
    import pytest
    with pytest.raises(IndexError):
        sro.CohesionOracle(10)


# ---- the pin: the reference's own reward() source (tests/golden/reference_runs.npz) ---------------------------------
@pytest.mark.parametrize("n", [2, 5, 7, 8, 9, 12, 40])
def test_flocking_oracle_equals_the_reference_source(n):
    """FlockingScenario (flocking_scenario.py, executed unmodified on oracle/refstub by
    tests/golden/make_reference_runs.py) vs the restatement: shaping memory after reset, 60 ticks of positions and
    collective rewards incl. contacts, the -1 collision terms and the +50 on-goal bonus -- bit for bit."""
    import numpy as np
    from helpers import npz
    g, pre = npz("reference_runs.npz"), f"flocking/n{n}/"
    o = sro.FlockingOracle(n)
    w = o.world
    w.goal = torch.tensor(list(sro.GOAL_POS)).unsqueeze(0)
    pos0 = torch.from_numpy(g[pre + "pos0"])
    for i in range(n):
        w.pos[i], w.vel[i] = torch.zeros(1, 2), torch.zeros(1, 2)
    for i in range(n):                                  # reset_world_at's placement loop (flocking:93-121)
        w.pos[i] = pos0[i:i + 1].clone()
        o.previous_distance_to_goal[i] = torch.linalg.vector_norm(w.pos[i] - w.goal, dim=1) * o.pos_shaping_factor
        o.previous_distance_to_agents[i] = o._spacing(i)
    shaping = lambda: np.stack([[float(o.previous_distance_to_goal[i]), float(o.previous_distance_to_agents[i])]
                                for i in range(n)]).astype(np.float32)
    assert np.array_equal(shaping(), g[pre + "shaping0"])
    contact = False
    for t in range(g[pre + "actions"].shape[0]):
        r = o.step(torch.from_numpy(g[pre + "actions"][t].astype(np.int64)))
        pos = torch.cat(w.pos).numpy()
        assert np.array_equal(pos, g[pre + "pos"][t]) and np.array_equal(torch.cat(w.vel).numpy(), g[pre + "vel"][t])
        assert np.all(np.float32(r.item()) == g[pre + "rewards"][t]), f"tick {t}"
        assert np.array_equal(shaping(), g[pre + "shaping"][t])
        d = np.linalg.norm(pos[:, None] - pos[None], axis=-1) + 9 * np.eye(n)
        contact |= bool(d.min() <= 0.1)
    assert contact, "the fixture is meant to exercise contacts"


@pytest.mark.parametrize("n", [2, 5, 9])
def test_cohesion_oracle_equals_the_reference_source(n):
    """CohesionScenario (cohesion_scenario.py incl. its np.exp on a tensor, cohesion:80): fixed start table, 60 ticks of
    positions and per-agent rewards -- bit for bit; observations are 4 floats."""
    import numpy as np
    from helpers import npz
    g, pre = npz("reference_runs.npz"), f"cohesion/n{n}/"
    o = sro.CohesionOracle(n)
    o.reset()
    assert np.array_equal(torch.cat(o.world.pos).numpy(), g[pre + "pos0"])
    assert o.observations().shape[-1] == int(g[pre + "obs_dim"]) == 4
    branches = set()
    for t in range(g[pre + "actions"].shape[0]):
        r = o.step(torch.from_numpy(g[pre + "actions"][t].astype(np.int64)))
        assert np.array_equal(torch.cat(o.world.pos).numpy(), g[pre + "pos"][t])
        assert np.array_equal(r.numpy(), g[pre + "rewards"][t]), f"tick {t}"
        branches |= {"near" if v > 0 else "far" for v in r.tolist()}
    assert branches == {"near", "far"}, "both sigma branches must occur"
