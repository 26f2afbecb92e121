#!/usr/bin/env python
"""Stage the reference's own Python files for the script-level drop-in tests.

    python scripts/stage_reference.py [--src /root/reference] [--dst tests/_refsrc]

Copies, byte for byte, `src/scenarios/*.py`, `src/simulation/simulator.py`, `src/training/train_gcn_dqn.py` and
`tests/test_*.py` of the reference into `tests/_refsrc/` (same relative layout).  The directory is git-ignored --
reference sources never enter this repository's history -- but it is NOT gpurun-ignored, so the staged copy travels
to the GPU box, where `/root/reference` does not exist; there `tests/test_gpu_reference_files.py` executes these
files unmodified on top of `shim/` (the `vmas` / `torch_geometric` module names backed by libswarm_b200.so).
`__graft_entry__.build()` calls this whenever `/root/reference` is present.  A MANIFEST with the sha256 of every
staged file is written next to them; the tests check it so a stale or edited copy is noticed.
"""
import argparse
import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["src/__init__.py", "src/scenarios/__init__.py", "src/scenarios/go_to_position_scenario.py",
         "src/scenarios/obstacle_avoidance_scenario.py", "src/scenarios/flocking_scenario.py",
         "src/scenarios/cohesion_scenario.py", "src/simulation/__init__.py", "src/simulation/simulator.py",
         "src/training/__init__.py", "src/training/train_gcn_dqn.py", "tests/__init__.py",
         "tests/test_go_to_position.py", "tests/test_obstacle_avoidance.py"]


def stage(src: str = "/root/reference", dst: str = os.path.join(ROOT, "tests", "_refsrc")) -> bool:
    if not os.path.isdir(src):
        return False
    manifest = {}
    for rel in FILES:
        s = os.path.join(src, rel)
        if not os.path.exists(s):
            continue
        d = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest[rel] = hashlib.sha256(open(d, "rb").read()).hexdigest()
    json.dump(manifest, open(os.path.join(dst, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--dst", default=os.path.join(ROOT, "tests", "_refsrc"))
    a = ap.parse_args()
    print("staged" if stage(a.src, a.dst) else f"{a.src} not present: nothing staged")
