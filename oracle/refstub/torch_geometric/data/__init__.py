"""torch_geometric.data.Data / Batch (what train_gcn_dqn.py:45,109 and simulator.py:25 use)."""
import torch


class Data:
    def __init__(self, x=None, edge_index=None, **kwargs):
        self.x = x
        self.edge_index = edge_index
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        return self.x.shape[0]


class Batch(Data):
    @classmethod
    def from_data_list(cls, data_list):
        # concatenate x; add the running node offset to each graph's edge_index
        xs, eis, batch, off = [], [], [], 0
        for g, d in enumerate(data_list):
            xs.append(d.x)
            eis.append(d.edge_index + off)
            batch.append(torch.full((d.x.shape[0],), g, dtype=torch.int64))
            off += d.x.shape[0]
        out = cls(x=torch.cat(xs, dim=0), edge_index=torch.cat(eis, dim=1))
        out.batch = torch.cat(batch)
        out.num_graphs = len(data_list)
        return out
