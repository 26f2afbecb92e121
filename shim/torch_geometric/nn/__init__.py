from swarm_b200.gcn import GATConv   # noqa: F401
