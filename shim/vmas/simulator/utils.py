from swarm_b200.scenarios import Color   # noqa: F401
