"""Import alias: ``import swarm_b200`` loads the package directory
``experiments-2025-acsos-marl-for-swarming-behaviors_b200/`` (whose name is not a Python identifier)."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("experiments-2025-acsos-marl-for-swarming-behaviors_b200")
sys.modules[__name__] = _pkg
for _name, _mod in list(sys.modules.items()):
    if _name.startswith("experiments-2025-acsos-marl-for-swarming-behaviors_b200."):
        sys.modules["swarm_b200." + _name.split(".", 1)[1]] = _mod
