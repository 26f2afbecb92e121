"""Builds the committed golden fixtures from the reference's shipped artefacts.

Run once in the build container (``/root/reference`` is not available on the GPU box):

    python tests/golden/make_golden.py [/root/reference]

Outputs (all under tests/golden/):
  models.npz        the 20 GoTo / ObstacleAvoidance state dicts of data/models/*.pth
                    (key ``{GoTo|ObstacleAvoidance}/{seed}/{param name}``)
  eval_go_to.npz, eval_obstacle_avoidance.npz
                    data/test_stats/{exp}/seed_{m}/agents_{n}/**: per-tick positions (float32, exact:
                    the CSVs hold repr(float(f32))), mean goal distance, hits, and result.csv rows
                    (key ``s{m}_n{n}/{pos|dist|hits|result}``)
  train_stats.npz   data/stats/experiment_{exp}-seed_{s}.csv (Episode, Reward, Loss as float64)
Only data files are converted; no reference source is copied.
"""
import csv
import os
import sys

import numpy as np
import torch

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def read_csv(path):
    with open(path, newline="") as f:
        rows = list(csv.reader(f))
    return rows[0], rows[1:]


def main():
    models = {}
    for exp in ("GoTo", "ObstacleAvoidance"):
        for seed in range(10):
            sd = torch.load(f"{REF}/data/models/experiment_{exp}-seed_{seed}.pth", map_location="cpu")
            for k, v in sd.items():
                models[f"{exp}/{seed}/{k}"] = v.numpy()
    np.savez_compressed(f"{HERE}/models.npz", **models)

    for exp in ("go_to", "obstacle_avoidance"):
        out = {}
        for m in range(10):
            for n in range(5, 13):
                d = f"{REF}/data/test_stats/{exp}/seed_{m}/agents_{n}"
                pos, dist, hits = [], [], []
                for e in range(8):
                    _, xs = read_csv(f"{d}/positions/positions_episode_{e}_x.csv")
                    _, ys = read_csv(f"{d}/positions/positions_episode_{e}_y.csv")
                    x = np.array([[float(v) for v in r[1:]] for r in xs], dtype=np.float64)
                    y = np.array([[float(v) for v in r[1:]] for r in ys], dtype=np.float64)
                    p = np.stack([x, y], axis=-1)
                    assert np.array_equal(p.astype(np.float32).astype(np.float64), p), "not exact f32"
                    pos.append(p.astype(np.float32))
                    _, ds = read_csv(f"{d}/data/distances_episode_{e}.csv")
                    dd = np.array([float(r[1]) for r in ds], dtype=np.float64)
                    assert np.array_equal(dd.astype(np.float32).astype(np.float64), dd)
                    dist.append(dd.astype(np.float32))
                    hits.append(np.array([float(r[2]) for r in ds], dtype=np.float32))
                _, rs = read_csv(f"{d}/result.csv")
                res = np.array([[float(v) for v in r[1:]] for r in rs], dtype=np.float64)
                key = f"s{m}_n{n}"
                out[f"{key}/pos"] = np.stack(pos)          # [8, T, n, 2]
                out[f"{key}/dist"] = np.stack(dist)        # [8, T]
                out[f"{key}/hits"] = np.stack(hits)        # [8, T]
                out[f"{key}/result"] = res                 # [8, 4] Reward, Collisions, Dist end, Dist beginning
        np.savez_compressed(f"{HERE}/eval_{exp}.npz", **out)

    stats = {}
    for exp in ("GoTo", "ObstacleAvoidance"):
        for seed in range(10):
            _, rows = read_csv(f"{REF}/data/stats/experiment_{exp}-seed_{seed}.csv")
            stats[f"{exp}/{seed}"] = np.array([[float(v) for v in r] for r in rows], dtype=np.float64)
    np.savez_compressed(f"{HERE}/train_stats.npz", **stats)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
