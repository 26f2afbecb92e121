"""Per-tick greedy actions, top-2 Q gaps and contact ticks of the ORACLE on all 1 280 golden evaluation episodes.

    python tests/golden/make_oracle_trace.py [--procs 8]        ->  tests/golden/eval_oracle_trace.npz

The GPU sweep (tests/test_gpu_golden_sweep.py) compares the fused CUDA rollout with the reference's shipped
trajectories (eval_*.npz) episode by episode; where the two part ways it needs to know whether the oracle -- which is
pinned to those trajectories (tests/golden/verify_full.py: 1 276 / 1 280 bit-identical) -- calls the deciding greedy
action a near-tie.  Running the oracle inside the GPU test would take minutes, so its verdicts are precomputed here:
  s{m}_n{n}/actions  int8   [8, T, n]  oracle greedy action per episode, tick, agent
  s{m}_n{n}/gap      float16[8, T, n]  (Q_top1 - Q_top2) / max|Q| of that agent's row (clipped to 1e-2)
  s{m}_n{n}/touch    bool   [8, T]     a contact force (agent-agent or agent-obstacle) acted in the env at this tick
  s{m}_n{n}/equal    bool   [8]        the oracle's positions equal the golden ones for the whole episode
Needs only tests/golden/*.npz and oracle/ (no /root/reference).  Uses the batched oracle (proven bit-equal to the
single-env one, tests/test_oracle_golden.py) with the 8 episodes of a golden run as 8 envs.
"""
import argparse
import multiprocessing as mp
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def case(args):
    exp, m, n = args
    torch.set_num_threads(1)
    from helpers import eval_centers, golden_eval, load_params
    from oracle import batched_oracle as bo
    T = 50 if exp == "go_to" else 100
    pos, vel = bo.reset_grid(exp, eval_centers(exp, n), n)
    ref = bo.rollout(exp, load_params(exp, m), pos, vel, T, "knn", 5)
    q = ref["q"]                                                     # [T, 8, n, 9]
    top2 = torch.topk(q, 2, dim=-1).values
    gap = ((top2[..., 0] - top2[..., 1]) / q.abs().amax(dim=-1)).clamp(max=1e-2)
    touch = ((ref["contact"] != 0) | ((ref["flags"] & 1) != 0)).any(dim=2)          # [T, 8]
    gold = golden_eval(exp, m, n)
    equal = (ref["pos"].permute(1, 0, 2, 3).numpy() == gold["pos"]).all(axis=(1, 2, 3))
    key = f"{exp}/s{m}_n{n}"
    return {f"{key}/actions": ref["actions"].permute(1, 0, 2).numpy().astype(np.int8),
            f"{key}/gap": gap.permute(1, 0, 2).numpy().astype(np.float16),
            f"{key}/touch": touch.t().numpy(),
            f"{key}/equal": equal}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    a = ap.parse_args()
    cases = [(exp, m, n) for exp in ("go_to", "obstacle_avoidance") for m in range(10) for n in range(5, 13)]
    out = {}
    with mp.get_context("fork").Pool(a.procs) as pool:
        for i, res in enumerate(pool.imap_unordered(case, cases)):
            out.update(res)
    eq = {exp: sum(int(v.sum()) for k, v in out.items() if k.startswith(exp) and k.endswith("/equal"))
          for exp in ("go_to", "obstacle_avoidance")}
    print(f"oracle == golden episodes: {eq}")
    np.savez_compressed(os.path.join(HERE, "eval_oracle_trace.npz"), **out)


if __name__ == "__main__":
    main()
