"""Shared helpers for the test-suite: golden fixtures and oracle-side recipes."""
import os
from typing import Dict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
EXP_NAME = {"go_to": "GoTo", "obstacle_avoidance": "ObstacleAvoidance"}

_cache: Dict[str, object] = {}


def npz(name: str):
    if name not in _cache:
        _cache[name] = np.load(os.path.join(GOLD, name))
    return _cache[name]


def load_params(exp: str, seed: int) -> Dict[str, torch.Tensor]:
    """Reference state dict of data/models/experiment_{exp}-seed_{seed}.pth."""
    models = npz("models.npz")
    pre = f"{EXP_NAME.get(exp, exp)}/{seed}/"
    return {k[len(pre):]: torch.from_numpy(models[k]).clone() for k in models.files if k.startswith(pre)}


def golden_eval(exp: str, model: int, n: int):
    g = npz(f"eval_{exp}.npz")
    key = f"s{model}_n{n}"
    return {k: g[f"{key}/{k}"] for k in ("pos", "dist", "hits", "result")}


def eval_centers(exp: str, n_agents: int, episodes: int = 8, seed: int = 6967) -> torch.Tensor:
    """Start centres of the golden evaluation episodes (SURVEY.md 8c recipe): env seed -> construction-time
    reset draw -> GCN() constructor draws -> one draw per episode."""
    from oracle import swarm_oracle as so
    from oracle.dqn_oracle import OracleGCN
    scen = so.GOTO if exp == "go_to" else so.OBSTACLE_AVOIDANCE
    torch.manual_seed(seed)
    so.draw_center(scen, True)
    OracleGCN(7, 32, 9)
    return torch.stack([so.draw_center(scen, True) for _ in range(episodes)])


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max(|b|, tiny) over all elements."""
    a = a.double()
    b = b.double()
    return float(((a - b).abs() / b.abs().clamp_min(1e-30)).max())
