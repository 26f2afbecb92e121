"""CPU stand-in for the slice of torch_geometric==2.5.3 the reference uses (TEST INFRASTRUCTURE, see ../README.md)."""
__version__ = "2.5.3+refstub"
