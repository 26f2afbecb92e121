"""Graph containers and builders at the reference's Python seam.

``Data`` / ``Batch`` stand in for ``torch_geometric.data.{Data, Batch}`` exactly as far as the reference
uses them (src/training/train_gcn_dqn.py:18,45,109; src/simulation/simulator.py:2,25): attribute access
to ``x`` / ``edge_index`` and ``Batch.from_data_list``.  The two ``create_graph_from_observations``
variants of the reference are reproduced on the device by ``swarm_graph_build``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import _lib, ops


class Data:
    """Minimal torch_geometric.data.Data: node features ``x`` f32[n,7] and ``edge_index`` int64[2,E]."""

    def __init__(self, x: Optional[torch.Tensor] = None, edge_index: Optional[torch.Tensor] = None, **kwargs):
        self.x = x
        self.edge_index = edge_index
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_nodes(self) -> int:
        return 0 if self.x is None else self.x.shape[0]

    @property
    def num_edges(self) -> int:
        return 0 if self.edge_index is None else self.edge_index.shape[1]

    def to(self, device) -> "Data":
        out = self.__class__.__new__(self.__class__)
        out.__dict__.update(self.__dict__)
        out.x = None if self.x is None else self.x.to(device)
        out.edge_index = None if self.edge_index is None else self.edge_index.to(device)
        return out

    def __repr__(self) -> str:
        xs = None if self.x is None else list(self.x.shape)
        es = None if self.edge_index is None else list(self.edge_index.shape)
        return f"{self.__class__.__name__}(x={xs}, edge_index={es})"


class Batch(Data):
    """torch_geometric.data.Batch.from_data_list: concatenated x, edge_index offset by the running
    node count, ``batch`` = graph id per node, ``ptr`` = node offsets."""

    @classmethod
    def from_data_list(cls, data_list: Sequence[Data]) -> "Batch":
        if len(data_list) == 0:
            raise ValueError("from_data_list needs at least one graph")
        xs, eis, batch, ptrs = [], [], [], [0]
        off = 0
        for g, d in enumerate(data_list):
            xs.append(d.x)
            eis.append(d.edge_index + off)
            n = d.x.shape[0]
            batch.append(torch.full((n,), g, dtype=torch.long, device=d.x.device))
            off += n
            ptrs.append(off)
        out = cls(x=torch.cat(xs, dim=0), edge_index=torch.cat(eis, dim=1))
        out.batch = torch.cat(batch)
        out.ptr = torch.tensor(ptrs, dtype=torch.long, device=xs[0].device)
        out.num_graphs = len(data_list)
        return out


def stack_observations(observations: Dict[str, torch.Tensor]) -> torch.Tensor:
    """{'agent{i}': f32[B,6]} -> f32[B,N,6] (train:95-96, simulator:10-11 for B = 1)."""
    n = len(observations)
    return torch.stack([observations[f"agent{i}"] for i in range(n)], dim=1)


def node_features(obs: torch.Tensor) -> torch.Tensor:
    """[obs | float(agent id)]: f32[B,N,6] -> f32[B*N,7] (train:98-99)."""
    B, N, _ = obs.shape
    ids = torch.arange(N, device=obs.device, dtype=torch.float32).view(1, N, 1).expand(B, N, 1)
    return torch.cat([obs, ids], dim=2).reshape(B * N, 7)


def create_graph_from_observations(observations: Dict[str, torch.Tensor], num_agents: Optional[int] = None,
                                   mode: str = "complete", k: int = 10) -> Data:
    """Both reference graph builders behind one call.

    mode='complete': DQNTrainer.create_graph_from_observations (train:94-110) -- edges (i,j),(j,i) for
    i<j then one (0,0).  mode='knn': simulator.create_graph_from_observations (simulator:9-26) -- per
    node the k nearest (self included, torch.topk order) emit (i,a),(a,i); then (0,0).  With batched
    observations (B > 1) the result is the Batch of the B per-env graphs (node ids offset by b*N).
    """
    obs = stack_observations(observations)
    B, N, _ = obs.shape
    if num_agents is not None and num_agents != N:
        raise ValueError(f"num_agents={num_agents} does not match the {N} observations")
    if not obs.is_cuda:
        raise _lib.SwarmError("graph builders run on CUDA tensors only (no CPU fallback)")
    gm = {"complete": _lib.GRAPH_COMPLETE, "knn": _lib.GRAPH_KNN}[mode]
    cfg = ops.make_config(_lib.SCENARIO_GOTO, B, N, gm, k)
    state = obs[:, :, :4].contiguous()
    edges, _ = ops.graph_build(cfg, state)                       # int32[B,2,E], env-local ids
    offs = (torch.arange(B, device=obs.device, dtype=torch.int64) * N).view(B, 1, 1)
    edge_index = (edges.to(torch.int64) + offs).permute(1, 0, 2).reshape(2, -1).contiguous()
    return Data(x=node_features(obs), edge_index=edge_index)
