"""The summation order `torch_row_sum` (csrc/swarm_device.cuh) emulates: torch's CPU sum / mean over a contiguous float32
row (ATen SumKernel.cpp: 8-lane vectors, four interleaved vector accumulators, tail first; four interleaved scalar
accumulators for rows shorter than a vector).  The Flocking reward's
`.pow(2).mean(-1)` over the N - 1 partner distances (flocking_scenario.py:109-121,148-160) goes through it, so the CUDA
kernels must add in this order to stay bit-exact beyond 9 agents.  This test pins the order itself against torch for every
row length the kernels support (N <= 128)."""
import numpy as np
import torch


def row_sum_order(e: np.ndarray) -> np.float32:
    """Step-for-step twin of the device function (same variable names)."""
    m = len(e)
    f = np.float32
    if m < 8:
        p = [f(0)] * 4
        k = 0
        if m >= 4:
            p, k = [e[0], e[1], e[2], e[3]], 4
        for k in range(k, m):
            p[0] = f(p[0] + e[k])
        return f(f(f(p[0] + p[1]) + p[2]) + p[3])
    chunks, groups = m >> 3, (m >> 3) >> 2
    a = [[f(0)] * 8 for _ in range(4)]
    c = 0
    for _ in range(groups):
        for l in range(8):
            for r in range(4):
                a[r][l] = f(a[r][l] + e[(c + r) * 8 + l])
        c += 4
    while c < chunks:
        for l in range(8):
            a[0][l] = f(a[0][l] + e[c * 8 + l])
        c += 1
    if groups > 0:
        for l in range(8):
            a[0][l] = f(f(f(a[0][l] + a[1][l]) + a[2][l]) + a[3][l])
    fin = f(0)
    for k in range(chunks * 8, m):
        fin = f(fin + e[k])
    for l in range(8):
        fin = f(fin + a[0][l])
    return fin


def test_row_sum_order_equals_torch_cpu_sum():
    torch.set_num_threads(1)
    rng = np.random.default_rng(0)
    sequential_differs = 0
    for m in range(1, 128):
        for rep in range(20):
            e = (rng.random(m, dtype=np.float32) * np.float32(10.0 ** rng.integers(-3, 3))).astype(np.float32)
            want = torch.from_numpy(e).reshape(1, m).sum(-1).numpy()[0]
            assert row_sum_order(e) == want, f"m = {m}"
            mean = torch.from_numpy(e).reshape(1, m).mean(-1).numpy()[0]
            assert np.float32(want / np.float32(m)) == mean            # mean = sum / count, one rounding
            seq = np.float32(0)
            for v in e:
                seq = np.float32(seq + v)
            sequential_differs += int(seq != want)
    assert sequential_differs > 100, "a left-to-right sum would be indistinguishable: the test would prove nothing"
