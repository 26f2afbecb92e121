"""`vmas` module name on top of swarm_b200 (see shim/README.md)."""
from swarm_b200 import make_env          # noqa: F401
from swarm_b200.env import Environment   # noqa: F401

__version__ = "1.4.0+swarm_b200"
