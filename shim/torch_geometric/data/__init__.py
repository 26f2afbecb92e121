from swarm_b200.graph import Batch, Data   # noqa: F401
