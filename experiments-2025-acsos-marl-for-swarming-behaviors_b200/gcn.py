"""The Q-network seam: ``GCN`` with the reference's nn.Module interface and state-dict key set.

Mirrors src/training/train_gcn_dqn.py:50-70 -- ``GCN(input_dim, hidden_dim, output_dim)`` holding
``conv1 = GATConv(input_dim, hidden_dim, add_self_loops=False, bias=True)``, ``lin1``, ``lin2`` and
``forward(data)`` reading ``data.x`` / ``data.edge_index``.  Parameter names, shapes and initialisation
(including the order and count of RNG draws) match torch_geometric 2.5.3 + torch so that
the shipped ``data/models/*.pth`` load and seeded runs consume the generator identically.  The forward and the
backward arithmetic (``loss.backward()`` through ``GCN.forward`` for any graph) run in libswarm_b200.so; there is no
PyTorch implementation behind it.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from . import _lib, ops

_FEAT, _HIDDEN, _ACTIONS = 7, 32, 9


def _glorot(t: torch.Tensor) -> None:
    # torch_geometric.nn.inits.glorot: U(-a, a), a = sqrt(6 / (fan_in + fan_out)) on the last two dims
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    t.data.uniform_(-a, a)


class _Projection(nn.Module):
    """torch_geometric Linear(bias=False, weight_initializer='glorot'): key ``weight`` [out, in]."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        self.reset_parameters()

    def reset_parameters(self) -> None:
        _glorot(self.weight)


def _compute_device(x: torch.Tensor) -> torch.device:
    """The CUDA device a layer call runs on: the tensor's own, or -- for host tensors of a script that hard-codes
    device='cpu' -- the B200 named by SWARM_DEVICE (``_lib.offload_device``; results return to the host)."""
    if x.is_cuda:
        return x.device
    off = _lib.offload_device()
    if off is None:
        raise _lib.SwarmError("GATConv / GCN run on CUDA tensors only (swarm_b200 has no CPU fallback); move the model "
                              "and the graph to a B200, or set SWARM_DEVICE=cuda to serve host tensors from one")
    return off


class _GatConvFunction(torch.autograd.Function):
    """Stand-alone GATConv layer through swarm_gatconv_forward_csr / swarm_gatconv_backward_csr (weights only)."""

    @staticmethod
    def forward(ctx, packed, x, edge_index):
        dev = _compute_device(x)
        ctx.homes = (packed.device, x.device)
        packed = packed.detach().to(dev).contiguous()
        x, edge_index = x.detach().to(dev).contiguous(), edge_index.to(dev).contiguous()
        row_ptr, src, perm = ops.csr_from_edges(edge_index, x.shape[0])
        out = ops.gatconv_forward_csr(packed, x, row_ptr, src)
        ctx.save_for_backward(packed, x, edge_index, row_ptr, src, perm)
        return out.to(ctx.homes[1])

    @staticmethod
    def backward(ctx, grad_out):
        packed, x, edge_index, row_ptr, src, perm = ctx.saved_tensors
        grad_w = ops.gatq_backward_csr(packed, x, edge_index, grad_out.to(x.device).contiguous(),
                                       by_target=(row_ptr, src, perm), conv_only=True)
        return grad_w.to(ctx.homes[0]), None, None


class _GatLayerFunction(torch.autograd.Function):
    """One GATConv layer of any width <= 64 (swarm_gat_layer_forward / swarm_gat_layer_backward): differentiable w.r.t.
    its four parameters and the node features, so layers stack under autograd."""

    @staticmethod
    def forward(ctx, x, lin_weight, att_src, att_dst, bias, edge_index, csr=None):
        dev = _compute_device(x)
        ctx.homes = (x.device, lin_weight.device)
        ctx.att_shape = att_src.shape
        x, lin_weight = x.detach().to(dev).contiguous(), lin_weight.detach().to(dev).contiguous()
        att_src, att_dst, bias = (t.detach().to(dev).reshape(-1).contiguous() for t in (att_src, att_dst, bias))
        edge_index = edge_index.to(dev).contiguous()
        row_ptr, src, perm = csr if csr is not None else ops.csr_from_edges(edge_index, x.shape[0])
        out = ops.gat_layer_forward(lin_weight, att_src, att_dst, bias, x, row_ptr, src)
        ctx.save_for_backward(x, lin_weight, att_src, att_dst, edge_index, row_ptr, src, perm)
        return out.to(ctx.homes[0])

    @staticmethod
    def backward(ctx, grad_out):
        x, lin_weight, att_src, att_dst, edge_index, row_ptr, src, perm = ctx.saved_tensors
        gw, gas, gad, gb, gx = ops.gat_layer_backward(lin_weight, att_src, att_dst, x, edge_index,
                                                      grad_out.to(x.device).contiguous(),
                                                      want_grad_x=ctx.needs_input_grad[0], by_target=(row_ptr, src, perm))
        hx, hw = ctx.homes
        return (gx.to(hx) if gx is not None else None), gw.to(hw), gas.view(ctx.att_shape).to(hw), \
            gad.view(ctx.att_shape).to(hw), gb.to(hw), None, None


class GATConv(nn.Module):
    """GATConv(in, out, heads=H, concat=True, add_self_loops=False, bias=True) (train:53 uses heads = 1).  Keys as in
    torch_geometric 2.5.3: ``att_src`` [1,H,out], ``att_dst`` [1,H,out], ``lin.weight`` [H*out,in], ``bias`` [H*out]
    (``concat=True``) or [out] (``concat=False``: mean over the heads, then the bias).  The projection is initialised
    twice, like torch_geometric (Linear.__init__ followed by GATConv.reset_parameters).  Inside ``swarm_b200.GCN`` it is a
    parameter container (the whole network is one fused call); called on its own -- ``conv(x, edge_index)`` as the
    reference's ``GCN.forward`` does (train:61) -- it runs the layer alone in the CUDA kernels: the (7 -> 32) first
    layer on data through the kernels of the fused network, any other width up to 64 (or an input that needs a
    gradient, i.e. a stacked layer) through the generic layer kernels, differentiable w.r.t. parameters and node
    features.  Heads are independent attention layers over slices of the projection, so a multi-head layer is H
    launches of the single-head layer kernels on one shared CSR grouping (forward and backward), concatenated or
    averaged as torch_geometric does."""

    def __init__(self, in_channels: int, out_channels: int, heads: int = 1, concat: bool = True,
                 add_self_loops: bool = False, bias: bool = True):
        super().__init__()
        if add_self_loops or not bias or heads < 1:
            raise NotImplementedError("the swarm_b200 kernels implement GATConv(add_self_loops=False, bias=True)")
        self.in_channels, self.out_channels, self.heads, self.concat = in_channels, out_channels, heads, concat
        self.lin = _Projection(in_channels, heads * out_channels)
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.empty(heads * out_channels if concat else out_channels))
        self.reset_parameters()

    def reset_parameters(self) -> None:
        self.lin.reset_parameters()
        _glorot(self.att_src)
        _glorot(self.att_dst)
        self.bias.data.zero_()

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        dev = _compute_device(x)                # raises for host tensors unless SWARM_DEVICE names the serving B200
        edge_index = edge_index.to(torch.int64).contiguous()
        if self.heads > 1:
            if self.in_channels > 64 or self.out_channels > 64:
                raise NotImplementedError("the swarm_b200 GAT layer kernels cover in_channels, out_channels <= 64")
            C_ = self.out_channels
            csr = ops.csr_from_edges(edge_index.to(dev), x.shape[0])          # one grouping for all heads
            zero = torch.zeros(C_, dtype=torch.float32, device=self.bias.device)
            outs = []
            for h in range(self.heads):
                b = self.bias[h * C_:(h + 1) * C_] if self.concat else zero
                outs.append(_GatLayerFunction.apply(x, self.lin.weight[h * C_:(h + 1) * C_], self.att_src[:, h:h + 1],
                                                    self.att_dst[:, h:h + 1], b, edge_index, csr))
            if self.concat:
                return torch.cat(outs, dim=1)
            return torch.stack(outs, dim=0).mean(dim=0) + self.bias.to(outs[0].device)
        if (self.in_channels, self.out_channels) != (_FEAT, _HIDDEN) or x.requires_grad:
            if self.in_channels > 64 or self.out_channels > 64:
                raise NotImplementedError("the swarm_b200 GAT layer kernels cover in_channels, out_channels <= 64")
            return _GatLayerFunction.apply(x, self.lin.weight, self.att_src, self.att_dst, self.bias, edge_index)
        head = torch.zeros(_lib.W_COUNT - 320, dtype=torch.float32, device=x.device)
        packed = torch.cat([self.lin.weight.reshape(-1), self.att_src.reshape(-1), self.att_dst.reshape(-1),
                            self.bias.reshape(-1), head])
        return _GatConvFunction.apply(packed, x.contiguous(), edge_index)


class _GatQFunction(torch.autograd.Function):
    """packed weights f32[1673], x f32[n,7], graph -> Q f32[n,9]; forward and backward run in the CUDA kernels
    (swarm_gatq_forward_csr / swarm_gatq_backward_csr).  Gradients flow to the weights only (x is data)."""

    @staticmethod
    def forward(ctx, packed, x, edge_index):
        dev = _compute_device(x)
        ctx.homes = (packed.device, x.device)
        packed = packed.detach().to(dev).contiguous()
        x, edge_index = x.detach().to(dev).contiguous(), edge_index.to(dev).contiguous()
        row_ptr, src, perm = ops.csr_from_edges(edge_index, x.shape[0])
        q = ops.gatq_forward_csr(packed, x, row_ptr, src)
        ctx.save_for_backward(packed, x, edge_index, row_ptr, src, perm)
        return q.to(ctx.homes[1])

    @staticmethod
    def backward(ctx, grad_q):
        packed, x, edge_index, row_ptr, src, perm = ctx.saved_tensors
        grad_w = ops.gatq_backward_csr(packed, x, edge_index, grad_q.to(x.device).contiguous(),
                                       by_target=(row_ptr, src, perm))
        return grad_w.to(ctx.homes[0]), None, None


class GCN(nn.Module):
    """train_gcn_dqn.py:50-70.  ``forward(data)`` accepts anything with ``x`` f32[n,7] and ``edge_index``
    int64[2,E] (a ``Data`` / ``Batch`` of this package or of torch_geometric)."""

    def __init__(self, input_dim: int = _FEAT, hidden_dim: int = _HIDDEN, output_dim: int = _ACTIONS):
        super().__init__()
        if (input_dim, hidden_dim, output_dim) != (_FEAT, _HIDDEN, _ACTIONS):
            raise NotImplementedError(
                f"the swarm_b200 kernels are specialised for GCN({_FEAT}, {_HIDDEN}, {_ACTIONS}) "
                f"(train_gcn_dqn.py:79-82), got ({input_dim}, {hidden_dim}, {output_dim})")
        self.conv1 = GATConv(input_dim, hidden_dim, add_self_loops=False, bias=True)
        self.lin1 = nn.Linear(hidden_dim, hidden_dim)
        self.lin2 = nn.Linear(hidden_dim, output_dim)

    def packed_weights(self) -> torch.Tensor:
        """float[1673] in the C ABI's SWARM_W_* order (differentiable w.r.t. the parameters)."""
        return torch.cat([self.conv1.lin.weight.reshape(-1), self.conv1.att_src.reshape(-1),
                          self.conv1.att_dst.reshape(-1), self.conv1.bias.reshape(-1),
                          self.lin1.weight.reshape(-1), self.lin1.bias.reshape(-1),
                          self.lin2.weight.reshape(-1), self.lin2.bias.reshape(-1)])

    def forward(self, data) -> torch.Tensor:
        x, edge_index = data.x, data.edge_index
        _compute_device(x)                      # raises for host tensors unless SWARM_DEVICE names the serving B200
        if edge_index.dtype != torch.int64:
            edge_index = edge_index.to(torch.int64)
        return _GatQFunction.apply(self.packed_weights(), x.contiguous(), edge_index.contiguous())


class StackedGCN(nn.Module):
    """The multi-layer form of ``GCN`` the reference keeps in comments (train_gcn_dqn.py:54-55, 64-67) and ships as
    data/models/experiment_Flocking-seed_*.pth: ``conv1 .. convL`` (single-head GATConv, one width), ``lin1``, ``lin2``;
    tanh after conv1, relu after the others unless ``activations`` says otherwise.  ``forward(data)`` stacks the generic
    layer kernels under torch autograd (any graph, trainable); greedy evaluation goes through ONE launch per forward
    (``ops.gatstack_forward`` / ``ops.rollout_stack``; ``Simulator`` picks that path for instances of this class)."""

    def __init__(self, input_dim: int = _FEAT, hidden_dim: int = 8, output_dim: int = _ACTIONS, n_layers: int = 3,
                 activations=None):
        super().__init__()
        if output_dim != _ACTIONS or input_dim not in (5, 7) or not 1 <= hidden_dim <= 32 or not 1 <= n_layers <= 4:
            raise NotImplementedError("StackedGCN: input_dim 7 or 5, 1 <= hidden_dim <= 32, 1 <= n_layers <= 4, 9 actions")
        self.n_layers = n_layers
        self.activations = list(activations) if activations else ["tanh"] + ["relu"] * (n_layers - 1)
        for l in range(1, n_layers + 1):
            setattr(self, f"conv{l}", GATConv(input_dim if l == 1 else hidden_dim, hidden_dim, add_self_loops=False, bias=True))
        self.lin1 = nn.Linear(hidden_dim, hidden_dim)
        self.lin2 = nn.Linear(hidden_dim, output_dim)

    @classmethod
    def from_state_dict(cls, sd, activations=None) -> "StackedGCN":
        """Shapes from the checkpoint itself (``torch.load`` of an experiment_Flocking-seed_*.pth)."""
        n_layers = sum(1 for k in sd if k.endswith(".lin.weight"))
        hidden, input_dim = sd["conv1.lin.weight"].shape
        model = cls(int(input_dim), int(hidden), int(sd["lin2.weight"].shape[0]), n_layers, activations)
        model.load_state_dict(sd)
        return model

    def stack_spec(self):
        from . import ops
        return ops.stack_spec(self.n_layers, self.lin1.in_features, self.conv1.lin.weight.shape[1], self.activations)

    def packed_stack_weights(self, device) -> torch.Tensor:
        from . import ops
        return ops.pack_stack_weights(self.state_dict(), self.stack_spec(), device)

    def forward(self, data) -> torch.Tensor:
        h, edge_index = data.x, data.edge_index
        for l in range(1, self.n_layers + 1):
            h = getattr(self, f"conv{l}")(h, edge_index)
            h = torch.tanh(h) if self.activations[l - 1] == "tanh" else torch.relu(h)
        return self.lin2(torch.relu(self.lin1(h)))
