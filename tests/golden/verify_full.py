"""Full oracle-vs-golden sweep (1 280 evaluation episodes + 20 training runs x 10 episodes).

    python tests/golden/verify_full.py [--procs 8] [--train-episodes 10]

The default CPU test-suite runs a subset (tests/test_oracle_golden.py); this script is the complete
pin and its summary is recorded in DESIGN.md.  It needs only tests/golden/*.npz (no /root/reference).
"""
import argparse
import multiprocessing as mp
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import swarm_oracle as so          # noqa: E402
from oracle.dqn_oracle import OracleGCN, OracleTrainer   # noqa: E402

EXP_NAME = {"go_to": "GoTo", "obstacle_avoidance": "ObstacleAvoidance"}


def load_params(exp: str, seed: int):
    models = np.load(os.path.join(GOLD, "models.npz"))
    pre = f"{EXP_NAME.get(exp, exp)}/{seed}/"
    return {k[len(pre):]: torch.from_numpy(models[k]) for k in models.files if k.startswith(pre)}


def eval_case(args):
    exp, m, n = args
    torch.set_num_threads(1)
    scen = so.GOTO if exp == "go_to" else so.OBSTACLE_AVOIDANCE
    T = 50 if exp == "go_to" else 100
    g = np.load(os.path.join(GOLD, f"eval_{exp}.npz"))
    # recipe: env seed -> construction-time reset draw -> GCN() constructor draws -> episodes
    torch.manual_seed(6967)
    w = so.OracleWorld(scen, n, random=True)
    w.reset()
    OracleGCN(7, 32, 9)
    out = so.run_evaluation(w, load_params(exp, m), 8, T, "knn", 5)
    key = f"s{m}_n{n}"
    pos = np.stack([np.stack([np.array(out["pos_x"][e], dtype=np.float32),
                              np.array(out["pos_y"][e], dtype=np.float32)], -1) for e in range(8)])
    gp = g[f"{key}/pos"]
    ep_equal = [bool((pos[e] == gp[e]).all()) for e in range(8)]
    first_bad = [int(np.argmax((pos[e] != gp[e]).any(axis=(1, 2)))) if not ep_equal[e] else -1 for e in range(8)]
    dist_eq = [bool((np.array(out["distance"][e], dtype=np.float32) == g[f"{key}/dist"][e]).all()) for e in range(8)]
    hits_eq = [bool((np.array(out["hits"][e], dtype=np.float32) == g[f"{key}/hits"][e]).all()) for e in range(8)]
    res = g[f"{key}/result"]
    rew_err = np.abs(np.array(out["reward"]) - res[:, 0]) / np.abs(res[:, 0])
    return exp, m, n, ep_equal, first_bad, dist_eq, hits_eq, rew_err.tolist()


def train_case(args):
    exp, seed, episodes = args
    torch.set_num_threads(1)
    stats = np.load(os.path.join(GOLD, "train_stats.npz"))[f"{exp}/{seed}"]
    tr = OracleTrainer(exp, seed, n_agents=5, max_steps=100)
    tr.train(episodes)
    # CSV row r: Reward = mean over episodes 10r..10r+9 of agent-0 return/N; Loss = mean loss of episode r
    loss_rel = [abs(tr.episode_losses[i] - stats[i, 2]) / abs(stats[i, 2]) for i in range(min(episodes, 100))]
    row0 = None
    if episodes >= 10:
        row0 = (sum(torch.tensor(v) for v in tr.episode_returns[:10]) / 10).item()
    return exp, seed, loss_rel, row0, float(stats[0, 1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--train-episodes", type=int, default=10)
    ap.add_argument("--skip-eval", action="store_true")
    a = ap.parse_args()
    with mp.get_context("fork").Pool(a.procs) as pool:
        if not a.skip_eval:
            cases = [(exp, m, n) for exp in ("go_to", "obstacle_avoidance") for m in range(10) for n in range(5, 13)]
            tot = {"go_to": [0, 0], "obstacle_avoidance": [0, 0]}
            for exp, m, n, ep_equal, first_bad, dist_eq, hits_eq, rew_err in pool.imap_unordered(eval_case, cases):
                tot[exp][0] += sum(ep_equal)
                tot[exp][1] += len(ep_equal)
                for e, ok in enumerate(ep_equal):
                    if not ok:
                        print(f"MISMATCH {exp} model {m} N {n} episode {e} first differing tick {first_bad[e]} "
                              f"dist_eq {dist_eq[e]} hits_eq {hits_eq[e]} reward rel err {rew_err[e]:.2e}", flush=True)
            for exp, (ok, n) in tot.items():
                print(f"eval {exp}: {ok}/{n} episodes bit-equal positions")
        tcases = [(exp, s, a.train_episodes) for exp in ("GoTo", "ObstacleAvoidance") for s in range(10)]
        exact = 0
        for exp, seed, loss_rel, row0, gold0 in pool.imap_unordered(train_case, tcases):
            exact += int(row0 == gold0)
            print(f"train {exp} seed {seed}: episode-0 loss rel err {loss_rel[0]:.2e}, max over {len(loss_rel)} "
                  f"episodes {max(loss_rel):.2e}; first-row reward {row0} vs golden {gold0} "
                  f"{'EXACT' if row0 == gold0 else 'differs'}", flush=True)
        print(f"train: {exact}/20 first-row rewards exact")


if __name__ == "__main__":
    main()
