from swarm_b200.scenarios import BaseScenario   # noqa: F401
