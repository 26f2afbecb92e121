"""CPU oracle for the swarm hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU with torch float32 ops, the arithmetic that the
reference (davidedomini/experiments-2025-acsos-marl-for-swarming-behaviors) executes through its
un-vendored pip dependencies vmas==1.4.0 and torch_geometric==2.5.3 (requirements.txt:1-3) plus its
own scenario / graph / DQN code.  It is the *checker* for the CUDA path:

  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
    reference`` legs may import it;
  * nothing under the product package imports it, and the product fails loudly when the CUDA
    extension is missing -- there is no CPU fallback.

Parity status: **pinned**.  ``tests/test_oracle_golden.py`` checks this oracle against the reference's
own shipped artefacts (per-tick positions / distances / hits / result.csv of ``data/test_stats`` and
the ``Episode,Reward,Loss`` rows of ``data/stats``; committed in compact form under ``tests/golden``
by ``tests/golden/make_golden.py``).

``scenario_rewards_oracle.py`` (Flocking / Cohesion) is pinned against the reference's own scenario source executed
unmodified on ``oracle/refstub`` (CPU stand-ins of vmas / torch_geometric backed by this oracle; see that directory's
README and ``tests/golden/make_reference_runs.py``).
"""
