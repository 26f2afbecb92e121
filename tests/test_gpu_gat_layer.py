"""The generic GATConv layer (swarm_gat_layer_forward / swarm_gat_layer_backward; any width <= 64, gradients w.r.t.
parameters and node features) against the oracle's restatement of torch_geometric's GATConv (oracle/swarm_oracle.py
``gat_conv``) under torch autograd, in float64 for the reference values.

Tolerance (float32 kernels, written here): forward 5e-6 of the largest output magnitude; each gradient tensor within
max(2e-5 x its largest entry, 20 x the error float32 autograd itself makes against float64)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _swarm():
    import swarm_b200 as sb
    return sb


def _dev():
    return torch.device("cuda:0")


def _graph(n, seed, extra_edges=4):
    """Random multigraph on n nodes: a ring, random extra edges, duplicates, self loops, one isolated target."""
    g = torch.Generator().manual_seed(seed)
    ring = torch.stack([torch.arange(n - 1), (torch.arange(n - 1) + 1) % (n - 1)])        # node n-1 has no in-edge
    rnd = torch.randint(0, n - 1, (2, extra_edges * n), generator=g)
    loops = torch.tensor([[0, 0, 3 % n], [0, 0, 3 % n]])
    out_only = torch.tensor([[n - 1], [1]])                                               # n-1 is a source only
    ei = torch.cat([ring, rnd, loops, ring[:, :3], out_only], dim=1)
    return ei[:, torch.randperm(ei.shape[1], generator=g)].contiguous()


def _oracle_layer(x, ei, w, a_s, a_d, b):
    from oracle import swarm_oracle as so
    return so.gat_conv(x, ei, w, a_s, a_d, b)


def _grad_check(got, ref32, ref64, name):
    scale = ref64.abs().max().item()
    err = (got.double() - ref64).abs().max().item()
    err32 = (ref32.double() - ref64).abs().max().item()
    assert err <= max(2e-5 * scale, 20 * err32, 1e-12), f"{name}: abs error {err:.3e} (float32 autograd {err32:.3e}, max {scale:.3e})"


@pytest.mark.parametrize("ci,co,n", [(7, 8, 50), (8, 8, 50), (7, 32, 64), (5, 3, 33), (64, 64, 40), (1, 1, 9), (13, 20, 300)])
def test_gat_layer_forward_backward_matches_oracle(ci, co, n):
    sb = _swarm()
    g = torch.Generator().manual_seed(ci * 100 + co)
    ei = _graph(n, seed=n + co)
    x = torch.randn(n, ci, generator=g)
    w = torch.randn(co, ci, generator=g) * 0.5
    a_s, a_d = torch.randn(1, 1, co, generator=g) * 0.5, torch.randn(1, 1, co, generator=g) * 0.5
    b = torch.randn(co, generator=g) * 0.1
    cot = torch.randn(n, co, generator=g)

    def run_oracle(dtype):
        leaves = [t.clone().to(dtype).requires_grad_(True) for t in (x, w, a_s, a_d, b)]
        out = _oracle_layer(leaves[0], ei, *leaves[1:])
        (out * cot.to(dtype)).sum().backward()
        return out.detach(), [t.grad for t in leaves]

    out64, g64 = run_oracle(torch.float64)
    out32, g32 = run_oracle(torch.float32)

    conv = sb.GATConv(ci, co).to(_dev())
    with torch.no_grad():
        conv.lin.weight.copy_(w)
        conv.att_src.copy_(a_s)
        conv.att_dst.copy_(a_d)
        conv.bias.copy_(b)
    xg = x.to(_dev()).requires_grad_(True)
    grads = []
    for _ in range(2):
        conv.zero_grad()
        xg.grad = None
        out = conv(xg, ei.to(_dev()))
        (out * cot.to(_dev())).sum().backward()
        grads.append([xg.grad.clone(), conv.lin.weight.grad.clone(), conv.att_src.grad.clone(), conv.att_dst.grad.clone(),
                      conv.bias.grad.clone()])
    err = (out.detach().cpu().double() - out64).abs().max().item()
    assert err <= 5e-6 * out64.abs().max().item(), f"forward: {err:.3e}"
    isolated = out.detach().cpu()[n - 1]
    assert torch.equal(isolated, b), "a node without in-edges receives the bias"
    for a, c in zip(grads[0], grads[1]):
        assert torch.equal(a, c), "gradients differ between two identical calls"
    for got, r32, r64, name in zip(grads[0], g32, g64, ("x", "lin.weight", "att_src", "att_dst", "bias")):
        _grad_check(got.cpu().reshape(r64.shape), r32, r64, name)


def test_three_layer_stack_on_the_shipped_flocking_weights():
    """The reference's data/models/experiment_Flocking-seed_*.pth hold conv1 (7 -> 8), conv2, conv3 (8 -> 8), lin1 (8 -> 8),
    lin2 (8 -> 9).  The class that produced them is not in the repository; the stack is assembled here with tanh between
    the layers (the activation of the shipped one-layer GCN, train:61-63) to exercise stacking: every layer gets gradients
    through the layers above it, compared with autograd through the oracle layers."""
    import torch.nn as nn
    sb = _swarm()
    z = np.load(os.path.join(ROOT, "tests", "golden", "flocking_models.npz"))
    sd = {k[len("4/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("4/")}
    assert sd["conv1.lin.weight"].shape == (8, 7) and sd["conv3.lin.weight"].shape == (8, 8) and sd["lin2.weight"].shape == (9, 8)

    class StackedGCN(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1, self.conv2, self.conv3 = sb.GATConv(7, 8), sb.GATConv(8, 8), sb.GATConv(8, 8)
            self.lin1, self.lin2 = nn.Linear(8, 8), nn.Linear(8, 9)

        def forward(self, data):
            h = torch.tanh(self.conv1(data.x, data.edge_index))
            h = torch.tanh(self.conv2(h, data.edge_index))
            h = torch.tanh(self.conv3(h, data.edge_index))
            return self.lin2(torch.relu(self.lin1(h)))

    model = StackedGCN()
    model.load_state_dict(sd)                          # same key set as the shipped files
    model = model.to(_dev())
    n = 60
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, 7, generator=g)
    ei = _graph(n, seed=21)
    cot = torch.randn(n, 9, generator=g)
    q = model(sb.Data(x=x.to(_dev()), edge_index=ei.to(_dev())))
    (q * cot.to(_dev())).sum().backward()

    def oracle(dtype):
        p = {k: v.clone().to(dtype).requires_grad_(True) for k, v in sd.items()}
        h = x.to(dtype)
        for l in ("conv1", "conv2", "conv3"):
            h = torch.tanh(_oracle_layer(h, ei, p[f"{l}.lin.weight"], p[f"{l}.att_src"], p[f"{l}.att_dst"], p[f"{l}.bias"]))
        h = torch.relu(h @ p["lin1.weight"].T + p["lin1.bias"])
        out = h @ p["lin2.weight"].T + p["lin2.bias"]
        (out * cot.to(dtype)).sum().backward()
        return out.detach(), {k: v.grad for k, v in p.items()}

    q64, g64 = oracle(torch.float64)
    _, g32 = oracle(torch.float32)
    assert (q.detach().cpu().double() - q64).abs().max().item() <= 1e-5 * q64.abs().max().item()
    for name, prm in model.named_parameters():
        _grad_check(prm.grad.cpu().reshape(g64[name].shape), g32[name], g64[name], name)


def test_gat_layer_argument_errors_and_empty_graph():
    sb = _swarm()
    conv = sb.GATConv(4, 6).to(_dev())
    x = torch.randn(5, 4, device=_dev())
    out = conv(x, torch.zeros(2, 0, dtype=torch.int64, device=_dev()))
    assert torch.equal(out, conv.bias.detach().expand(5, 6))
    xg = x.clone().requires_grad_(True)
    conv(xg, torch.zeros(2, 0, dtype=torch.int64, device=_dev())).sum().backward()
    assert float(xg.grad.abs().max()) == 0.0 and torch.allclose(conv.bias.grad, torch.full((6,), 5.0, device=_dev()))
    with pytest.raises(NotImplementedError):
        sb.GATConv(65, 8).to(_dev())(torch.randn(3, 65, device=_dev()), torch.zeros(2, 0, dtype=torch.int64, device=_dev()))
    with pytest.raises(NotImplementedError):
        sb.GATConv(8, 8, add_self_loops=True)
    with pytest.raises(sb.SwarmError):
        sb.GATConv(4, 6)(torch.randn(5, 4), torch.zeros(2, 0, dtype=torch.int64))          # CPU tensors: no fallback


@pytest.mark.parametrize("heads,concat,ci,co,n", [(3, True, 7, 8, 60), (2, False, 5, 16, 41), (4, True, 12, 4, 90)])
def test_multi_head_layer_matches_per_head_oracle(heads, concat, ci, co, n):
    """GATConv(heads = H): torch_geometric projects to H * C channels, runs one attention per head on its slice and
    concatenates (``concat=True``) or averages (``concat=False``) the heads before the bias.  Reference = the oracle's
    single-head layer applied per head under float64 autograd; ours = swarm_b200.GATConv (H launches of the layer kernels
    on one CSR grouping) with gradients for every parameter and the node features."""
    sb = _swarm()
    g = torch.Generator().manual_seed(heads * 1000 + ci * 10 + co)
    ei = _graph(n, seed=n + heads)
    x = torch.randn(n, ci, generator=g)
    conv = sb.GATConv(ci, co, heads=heads, concat=concat)
    assert conv.lin.weight.shape == (heads * co, ci) and conv.att_src.shape == (1, heads, co)
    assert conv.bias.shape == ((heads * co,) if concat else (co,))
    with torch.no_grad():
        conv.lin.weight.copy_(torch.randn(heads * co, ci, generator=g) * 0.5)
        conv.att_src.copy_(torch.randn(1, heads, co, generator=g) * 0.5)
        conv.att_dst.copy_(torch.randn(1, heads, co, generator=g) * 0.5)
        conv.bias.copy_(torch.randn(conv.bias.shape, generator=g) * 0.1)
    cot = torch.randn(n, heads * co if concat else co, generator=g)

    def run_oracle(dtype):
        xx = x.detach().clone().to(dtype).requires_grad_(True)
        ps = [p.detach().clone().to(dtype).requires_grad_(True) for p in (conv.lin.weight, conv.att_src, conv.att_dst, conv.bias)]
        w, a_s, a_d, b = ps
        outs = []
        for h in range(heads):
            bh = b[h * co:(h + 1) * co] if concat else torch.zeros(co, dtype=dtype)
            outs.append(_oracle_layer(xx, ei, w[h * co:(h + 1) * co], a_s[:, h:h + 1], a_d[:, h:h + 1], bh))
        out = torch.cat(outs, dim=1) if concat else torch.stack(outs).mean(dim=0) + b
        (out * cot.to(dtype)).sum().backward()
        return out.detach(), [xx.grad] + [p.grad for p in ps]

    ref64, g64 = run_oracle(torch.float64)
    _, g32 = run_oracle(torch.float32)
    conv = conv.to(_dev())
    xg = x.detach().clone().to(_dev()).requires_grad_(True)
    out = conv(xg, ei.to(_dev()))
    assert out.shape == ref64.shape
    assert (out.detach().cpu().double() - ref64).abs().max().item() <= 5e-6 * max(ref64.abs().max().item(), 1.0)
    (out * cot.to(_dev())).sum().backward()
    got = [xg.grad, conv.lin.weight.grad, conv.att_src.grad, conv.att_dst.grad, conv.bias.grad]
    for name, a, r32, r64 in zip(["x", "lin.weight", "att_src", "att_dst", "bias"], got, g32, g64):
        assert a is not None and r64 is not None, f"{name}: missing gradient (ours {a is not None}, oracle {r64 is not None})"
        assert a.shape == r64.shape
        _grad_check(a.cpu(), r32, r64, name)
