"""csrc/knn_select.h (the tie-exact selection the kNN graph kernel runs per thread) against torch.topk on
the CPU -- the call the reference makes at src/simulation/simulator.py:19.  Grid start states make exact
distance ties the common case, so tie-heavy rows are the point of this test."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = tmp_path_factory.mktemp("knn") / "libknn_shim.so"
    src = os.path.join(ROOT, "tests", "host_shim", "knn_shim.cpp")
    inc = os.path.join(ROOT, "experiments-2025-acsos-marl-for-swarming-behaviors_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", inc, src, "-o", str(out)], check=True)
    lib = ctypes.CDLL(str(out))
    lib.swarm_host_topk_smallest.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.swarm_host_topk_small.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_void_p]
    lib.swarm_host_topk_small.restype = ctypes.c_int
    return lib


def _ours(lib, values: torch.Tensor, k: int) -> torch.Tensor:
    rows, n = values.shape
    v = values.contiguous()
    out = torch.empty(rows, k, dtype=torch.int32)
    lib.swarm_host_topk_smallest(v.data_ptr(), rows, n, k, out.data_ptr())
    return out.long()


def _grid_distance_rows(n: int, rng: torch.Generator, jitter: float) -> torch.Tensor:
    """distance rows of n agents on the scenarios' 0.15-spaced start grid (ties!), optionally perturbed."""
    cols = int(np.ceil(np.sqrt(n)))
    pts = torch.tensor([[(i % cols) * 0.15, (i // cols) * 0.15] for i in range(n)], dtype=torch.float32)
    pts = pts + torch.randn(2, generator=rng)
    if jitter:
        pts = pts + jitter * torch.randn(n, 2, generator=rng)
    return torch.linalg.norm(pts.unsqueeze(0) - pts.unsqueeze(1), dim=-1)      # [i, j] = ||p_j - p_i||


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 16, 20, 29, 32, 64, 100])
def test_matches_torch_topk_on_grids(shim, n):
    g = torch.Generator().manual_seed(n)
    rows = torch.cat([_grid_distance_rows(n, g, j) for j in (0.0, 0.0, 0.0, 1e-3, 0.05)])
    for k in sorted({1, min(2, n), min(3, n), min(5, n), min(10, n), min(17, n), n}):
        ref = torch.topk(rows, k, dim=-1, largest=False).indices
        assert torch.equal(_ours(shim, rows, k), ref), f"n={n} k={k}"


@pytest.mark.parametrize("n,k", [(5, 5), (12, 5), (12, 10), (32, 10), (32, 17), (40, 5), (128, 10), (128, 2), (1024, 10), (1024, 17)])
def test_matches_torch_topk_on_tie_heavy_random_rows(shim, n, k):
    g = torch.Generator().manual_seed(1000 + n + k)
    # few distinct values -> massive ties; mixes both torch branches (nth_element+sort, partial_sort for 64k <= n)
    rows = torch.randint(0, 7, (400, n), generator=g).float() * 0.25
    rows2 = torch.rand(200, n, generator=g)
    allrows = torch.cat([rows, rows2])
    ref = torch.topk(allrows, k, dim=-1, largest=False).indices
    assert torch.equal(_ours(shim, allrows, k), ref)


def test_nan_ordering_matches_torch(shim):
    rows = torch.tensor([[0.3, float("nan"), 0.1, 0.1, float("nan"), 0.2, 0.0, 5.0]] * 3)
    for k in (1, 3, 6, 8):
        ref = torch.topk(rows, k, dim=-1, largest=False).indices
        assert torch.equal(_ours(shim, rows, k), ref)


def _small(lib, values: torch.Tensor, k: int, mode: int):
    rows, n = values.shape
    v = values.contiguous()
    out = torch.empty(rows, k, dtype=torch.int32)
    emulated = lib.swarm_host_topk_small(v.data_ptr(), rows, n, k, mode, out.data_ptr())
    return out.long(), emulated


@pytest.mark.parametrize("n", list(range(1, 17)))
def test_small_register_path_matches_torch_topk(shim, n):
    """csrc/knn_small.h: ranks from pairwise comparisons, the tie-free shortcut, and the libstdc++ emulation on the
    nibble-packed rank pattern (what the env-tile kernels run for n <= 16)."""
    g = torch.Generator().manual_seed(77 + n)
    rows = torch.cat([_grid_distance_rows(n, g, j) for j in (0.0, 0.0, 0.0, 0.0, 1e-7, 1e-3, 0.05)] +
                     [torch.randint(0, 5, (300, n), generator=g).float() * 0.25, torch.rand(200, n, generator=g),
                      torch.randint(0, 2, (100, n), generator=g).float()])
    rows[-1, n // 2] = float("nan")
    rows[-2, :] = float("nan")
    rows[-3, 0] = float("inf")
    took_shortcut = False
    for k in range(1, n + 1):
        ref = torch.topk(rows, k, dim=-1, largest=False).indices
        got, emulated = _small(shim, rows, k, 0)
        assert torch.equal(got, ref), f"n={n} k={k} (kernel path)"
        took_shortcut |= emulated < rows.shape[0]
        got, emulated = _small(shim, rows, k, 1)
        assert emulated == rows.shape[0] and torch.equal(got, ref), f"n={n} k={k} (emulation on every row)"
        got, emulated = _small(shim, rows, k, 2)
        assert emulated == rows.shape[0] and torch.equal(got, ref), f"n={n} k={k} (word-level emulation on every row)"
    assert took_shortcut


@pytest.mark.parametrize("n,k", [(4, 2), (5, 3), (6, 3), (6, 5), (7, 4)])
def test_small_register_path_exhaustive_patterns(shim, n, k):
    """every order pattern with ties of n values drawn from n levels (n^n rows)"""
    grids = torch.cartesian_prod(*[torch.arange(n, dtype=torch.float32)] * n)
    ref = torch.topk(grids, k, dim=-1, largest=False).indices
    got, _ = _small(shim, grids, k, 0)
    assert torch.equal(got, ref)
    got, _ = _small(shim, grids, k, 1)
    assert torch.equal(got, ref)
    got, _ = _small(shim, grids, k, 2)
    assert torch.equal(got, ref)


@pytest.mark.parametrize("n", [8, 12, 16])
def test_word_level_emulation_equals_the_stepwise_one_on_adversarial_patterns(shim, n):
    """knn_small_topk_fast against knn_small_topk (and torch.topk) on patterns chosen to stress introselect: few
    distinct values, sorted / reversed / organ-pipe rows (bad median-of-three pivots, depth limit), all k."""
    g = torch.Generator().manual_seed(n)
    rows = [torch.randint(0, m, (400, n), generator=g).float() for m in (2, 3, 4, n)]
    base = torch.arange(n, dtype=torch.float32)
    organ = torch.cat([base[::2], base[1::2].flip(0)])
    rows += [base[None], base.flip(0)[None], organ[None], organ.flip(0)[None], (base // 2)[None], (base // 3).flip(0)[None]]
    rows += [torch.stack([torch.roll(organ, s) for s in range(n)]), torch.stack([torch.roll(base // 2, s) for s in range(n)])]
    rows = torch.cat(rows)
    for k in range(1, n + 1):
        ref = torch.topk(rows, k, dim=-1, largest=False).indices
        slow, _ = _small(shim, rows, k, 1)
        fast, _ = _small(shim, rows, k, 2)
        assert torch.equal(slow, ref) and torch.equal(fast, ref), f"n={n} k={k}"


@pytest.mark.parametrize("n,k", [(5, 5), (6, 3), (8, 3), (9, 5), (12, 5), (12, 10), (16, 6), (16, 15)])
def test_rank_below_k_decides_the_neighbour_set(n, k):
    """The order-free shortcut of the tensor-core rollout (csrc/tile_device.cuh tile_knn_small_set_np): with
    rank_j = #{l : d_l < d_j}, S = {j : rank_j < k} always contains torch.topk's k indices, so |S| = k means the SET is
    S whatever order the algorithm returns it in -- checked against torch.topk on tie-heavy rows (lattice distances)."""
    g = torch.Generator().manual_seed(n * 31 + k)
    decided = undecided = 0
    for trial in range(400):
        rows = _grid_distance_rows(n, g, jitter=0.0 if trial % 2 == 0 else 1e-3)
        for row in rows:
            rank = (row[None, :] < row[:, None]).sum(1)
            s = set(torch.nonzero(rank < k).flatten().tolist())
            top = set(torch.topk(row, k, largest=False).indices.tolist())
            assert top <= s
            if len(s) == k:
                assert top == s
                decided += 1
            else:
                undecided += 1
    assert decided > 0
    if (n, k) == (12, 5):
        assert undecided > 0, "the lattice rows must contain boundary ties"
