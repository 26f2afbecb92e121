"""torch_geometric.nn.GATConv(in, out, add_self_loops=False, bias=True), heads = 1: the pinned restatement of
oracle/dqn_oracle.py (same parameter names, shapes, init order and RNG draws as PyG 2.5.3)."""
from oracle.dqn_oracle import OracleGATConv as _OracleGATConv


class GATConv(_OracleGATConv):
    def __init__(self, in_channels, out_channels, heads=1, add_self_loops=True, bias=True, **kwargs):
        assert heads == 1 and not add_self_loops and bias and not kwargs, "only the reference's configuration is restated"
        super().__init__(in_channels, out_channels)
