"""`torch_geometric` module name on top of swarm_b200 (see shim/README.md)."""
__version__ = "2.5.3+swarm_b200"
