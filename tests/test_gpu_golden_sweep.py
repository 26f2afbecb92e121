"""The complete golden sweep on the GPU: all 1 280 greedy evaluation episodes the reference ships
(data/test_stats/**, converted to tests/golden/eval_*.npz) through the fused CUDA rollout (kNN k = 5), judged at
the oracle's own level (tests/golden/verify_full.py is the CPU twin: 640 / 640 GoTo, 636 / 640 ObstacleAvoidance).

For every episode:
  * the greedy action stream must equal the oracle's (tests/golden/eval_oracle_trace.npz) up to the first tick at
    which the oracle itself calls the deciding action a near-tie -- top-2 Q gap <= 2e-5 of max|Q| (the float32 Q
    tolerance, both sides) -- a "flip"; nothing is asserted about an episode after its flip (it is a different,
    equally valid greedy trajectory from there on);
  * until the first contact force in the env the positions equal the reference's CSV values bit for bit; afterwards
    (the contact magnitude goes through log1p(exp(.)), SLEEF on the CPU vs CUDA libm) within 1e-4 absolute;
  * hit counts equal the reference's distances_episode CSV on every bit-identical episode.
The counts are asserted against what was measured on a B200 (recorded in DESIGN.md), so a regression that breaks even
one more episode fails.
"""
import numpy as np
import pytest
import torch

from helpers import eval_centers, golden_eval, load_params, npz

pytestmark = pytest.mark.gpu

BAND = 2e-5          # relative top-2 Q gap below which a greedy action is a float32 near-tie (2 x Q tolerance 1e-5)
# measured on B200 (round 2): bit-identical episodes / flips inside the band, per experiment
# go_to: the 10 others differ only under contact forces (<= 7e-6); obstacle_avoidance: 23 flips, all at oracle gaps <= 1.8e-7
EXPECT = {"go_to": dict(min_exact=630, max_flips=0), "obstacle_avoidance": dict(min_exact=610, max_flips=23)}


@pytest.mark.parametrize("exp", ["go_to", "obstacle_avoidance"])
def test_full_golden_sweep(exp):
    import swarm_b200 as sb
    dev = torch.device("cuda:0")
    trace = npz("eval_oracle_trace.npz")
    scen = sb._lib.SCENARIO_GOTO if exp == "go_to" else sb._lib.SCENARIO_OBSTACLE_AVOIDANCE
    T = 50 if exp == "go_to" else 100
    exact = flips = total = contact_eps = 0
    worst_contact_err = 0.0
    misses = []
    for n in range(5, 13):
        centers = eval_centers(exp, n).to(dev)
        cfg = sb.ops.make_config(scen, 8, n, sb._lib.GRAPH_KNN, 5)
        for m in range(10):
            gold = golden_eval(exp, m, n)
            key = f"{exp}/s{m}_n{n}"
            o_act = trace[f"{key}/actions"].astype(np.int64)            # [8, T, n]
            o_gap = trace[f"{key}/gap"].astype(np.float32)
            o_touch = trace[f"{key}/touch"]                             # [8, T]
            o_equal = trace[f"{key}/equal"]
            state = sb.ops.reset_grid(cfg, centers)
            out = sb.ops.rollout(cfg, sb.pack_weights(load_params(exp, m), dev), state, T,
                                 trace=dict(state=True, actions=True, flags=True, contact=True))
            pos = out["trace_state"][..., :2].cpu().permute(1, 0, 2, 3).numpy()          # [8, T, n, 2]
            act = out["trace_actions"].cpu().permute(1, 0, 2).numpy().astype(np.int64)    # [8, T, n]
            flg = out["trace_flags"].cpu().permute(1, 0, 2).numpy()
            con = out["trace_contact"].cpu().permute(1, 0, 2).numpy()
            hits = ((flg & 2) != 0).sum(axis=2).astype(np.float32)                        # [8, T]
            touch = ((con != 0) | ((flg & 1) != 0)).any(axis=2)                           # [8, T]
            for e in range(8):
                total += 1
                diff = act[e] != o_act[e]
                horizon = T                                   # ticks whose outcome is asserted
                if diff.any():
                    t0 = int(np.argmax(diff.any(axis=1)))
                    bad = diff[t0] & ~(o_gap[e, t0] <= BAND)
                    assert not bad.any(), (f"{key} episode {e}: greedy action differs from the oracle at tick {t0} outside "
                                           f"the near-tie band (gaps {o_gap[e, t0][diff[t0]]})")
                    flips += 1
                    horizon = t0
                    misses.append((key, e, t0, float(o_gap[e, t0][diff[t0]].max())))
                # contact masks are bit-exact wherever the action streams agree
                assert (touch[e, :horizon] == o_touch[e, :horizon]).all(), f"{key} episode {e}: contact ticks differ"
                ever = np.cumsum(touch[e]) > 0                # a contact force has acted at or before tick t
                gp = gold["pos"][e]
                if not o_equal[e]:
                    # one of the oracle's own four misses: the reference took the other branch of a near-tie; the golden
                    # positions are only binding up to the oracle's flip tick
                    horizon = min(horizon, int(np.argmax((pos[e] != gp).any(axis=(1, 2)))) if (pos[e] != gp).any() else T)
                free = ~ever
                free[horizon:] = False
                assert (pos[e][free] == gp[free]).all(), f"{key} episode {e}: contact-free prefix is not bit-identical"
                if horizon > 0:
                    err = float(np.abs(pos[e][:horizon] - gp[:horizon]).max())
                    assert err <= 1e-4, f"{key} episode {e}: position error {err:.2e} before tick {horizon}"
                    if ever[:horizon].any():
                        worst_contact_err = max(worst_contact_err, err)
                contact_eps += int(ever.any())
                if (pos[e] == gp).all():
                    exact += 1
                    assert (hits[e] == gold["hits"][e]).all(), f"{key} episode {e}: hit counts differ"
    print(f"\n{exp}: {exact}/{total} golden episodes bit-identical on the GPU; {flips} near-tie flips "
          f"{misses}; {contact_eps} episodes with contact forces, worst position error under contact "
          f"{worst_contact_err:.2e}")
    assert total == 640
    assert exact >= EXPECT[exp]["min_exact"], f"only {exact}/640 bit-identical (expected >= {EXPECT[exp]['min_exact']})"
    assert flips <= EXPECT[exp]["max_flips"], f"{flips} near-tie flips (expected <= {EXPECT[exp]['max_flips']})"
