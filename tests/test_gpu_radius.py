"""Radius graph (EXTENSION -- the reference has none): the complete builder filtered by distance.
Checked against its oracle definition (oracle/batched_oracle.py::edges_radius) and against the reference-pinned
complete graph, which it must reproduce edge for edge (and Q for Q) when the radius is infinite."""
import pytest
import torch

from helpers import load_params
from test_gpu_parity import _pack_state, _random_states, _scen_id, Q_RTOL

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("n,B,radius", [(12, 200, 0.2), (5, 64, 0.16), (32, 40, 0.25), (9, 50, 0.0), (100, 6, 0.3)])
def test_radius_edges_bitexact(n, B, radius):
    import swarm_b200 as sb
    from oracle import batched_oracle as bo
    pos, vel = _random_states("obstacle_avoidance", B, n, seed=5 + n, crowd=True)
    if radius > 0:
        pos[:, 1] = pos[:, 0] + torch.tensor([radius, 0.0])    # pairs exactly at the boundary (<= is inclusive)
    ref = bo.edges_radius(pos, radius)
    cfg = sb.ops.make_config(sb._lib.SCENARIO_OBSTACLE_AVOIDANCE, B, n, sb._lib.GRAPH_RADIUS, graph_radius=radius)
    edges, counts = sb.ops.graph_build(cfg, _pack_state(pos, vel).to(_dev()))
    assert edges.shape == (B, 2, n * (n - 1) + 1)
    assert torch.equal(edges.cpu().long(), ref)
    assert torch.equal(counts.cpu().long(), (ref[:, 0] >= 0).sum(dim=1))
    if radius > 0:
        assert counts.max() > 1 and (counts.cpu() % 2 == 1).all()
    else:
        assert (counts == 1).all()


@pytest.mark.parametrize("exp,scenario,n", [("ObstacleAvoidance", "obstacle_avoidance", 12), ("GoTo", "go_to", 7)])
def test_infinite_radius_is_the_complete_graph(exp, scenario, n, monkeypatch):
    import swarm_b200 as sb
    ops, L = sb.ops, sb._lib
    B, T = 300, 20
    pos, vel = _random_states(scenario, B, n, seed=3, crowd=False)
    state = _pack_state(pos, vel).to(_dev())
    w = sb.pack_weights(load_params(exp, 2), _dev())
    cc = ops.make_config(_scen_id(sb, scenario), B, n, L.GRAPH_COMPLETE)
    cr = ops.make_config(_scen_id(sb, scenario), B, n, L.GRAPH_RADIUS, graph_radius=float("inf"))
    ec, _ = ops.graph_build(cc, state)
    er, counts = ops.graph_build(cr, state)
    assert torch.equal(ec, er) and (counts == n * (n - 1) + 1).all()
    # tensor-core path: the complete graph is summed without an edge list (node 0's self loop first instead of last),
    # so the two agree to rounding: Q within the Q tolerance, greedy actions equal wherever the top-2 gap is not a near-tie
    qc, ac = ops.gatq_forward(cc, w, state, want_actions=True)
    qr, ar = ops.gatq_forward(cr, w, state, want_actions=True)
    scale = qc.abs().amax(dim=-1, keepdim=True)
    assert ((qc - qr).abs() / scale).max().item() <= Q_RTOL
    top2 = qc.topk(2, dim=-1).values
    clear = ((top2[..., 0] - top2[..., 1]) / scale[..., 0]) > 2 * Q_RTOL
    assert torch.equal(ac[clear], ar[clear])
    # bit-faithful path (SWARM_TC=0, the reference's op order over the explicit edge list): Q for Q, state for state
    monkeypatch.setenv("SWARM_TC", "0")
    qc, ac = ops.gatq_forward(cc, w, state, want_actions=True)
    qr, ar = ops.gatq_forward(cr, w, state, want_actions=True)
    assert torch.equal(qc, qr) and torch.equal(ac, ar)
    oc = ops.rollout(cc, w, state.clone(), T)
    orr = ops.rollout(cr, w, state.clone(), T)
    assert torch.equal(oc["state"], orr["state"]) and torch.equal(oc["returns"], orr["returns"])


@pytest.mark.parametrize("exp,scenario,n,radius", [("ObstacleAvoidance", "obstacle_avoidance", 12, 0.22), ("GoTo", "go_to", 9, 0.17)])
def test_radius_forward_and_rollout_match_oracle(exp, scenario, n, radius):
    import swarm_b200 as sb
    from oracle import batched_oracle as bo
    ops, L = sb.ops, sb._lib
    B, T = 40, 12
    params = load_params(exp, 1)
    pos, vel = _random_states(scenario, B, n, seed=21, crowd=False)
    ref = bo.rollout(scenario, params, pos, vel, T, "radius", radius)
    cfg = ops.make_config(_scen_id(sb, scenario), B, n, L.GRAPH_RADIUS, graph_radius=radius)
    state = _pack_state(pos, vel).to(_dev())
    forced = ref["actions"].to(torch.int32).to(_dev()).contiguous()
    out = ops.rollout(cfg, sb.pack_weights(params, _dev()), state, T, forced_actions=forced,
                      trace=dict(state=True, q=True, rewards=True))
    touched = ((ref["contact"] != 0) | ((ref["flags"] & 1) != 0)).any(dim=2, keepdim=True)
    ever = (torch.cumsum(touched.long(), dim=0) > 0).expand(T, B, n)
    st = out["trace_state"].cpu()
    assert torch.equal(st[..., :2][~ever], ref["pos"][~ever])
    pre = torch.cat([torch.zeros(1, B, n, dtype=torch.bool), ever[:-1]], dim=0)
    scale = ref["q"].abs().amax(dim=-1, keepdim=True)
    err = ((out["trace_q"].cpu().double() - ref["q"].double()).abs() / scale.double())[~pre].max().item()
    assert err <= Q_RTOL, f"Q relative error {err:.3e}"
    degs = (bo.edges_radius(pos, radius)[:, 1] >= 0).sum(1)
    assert degs.min() < n * (n - 1) + 1, "the radius must actually prune edges in this test"
    with pytest.raises(sb.SwarmError):
        ops.rollout(cfg, sb.pack_weights(params, _dev()), state, 2, trace=dict(edges=True))


def test_radius_dqn_gradient_matches_autograd():
    import swarm_b200 as sb
    from test_gpu_dqn import _oracle_loss_and_grads, _ring_from, _transitions, GRAD_RTOL
    G, N, radius = 48, 10, 0.2
    pos, vel, actions, rewards, pos2, vel2 = _transitions("obstacle_avoidance", G, N, seed=13)
    loss_ref, grads, td_ref, online, target = _oracle_loss_and_grads("ObstacleAvoidance", 4, "obstacle_avoidance", pos, vel,
                                                                     actions, rewards, pos2, vel2, "radius", radius)
    ring = _ring_from(sb, pos, vel, actions, rewards, pos2, vel2)
    cfg = sb.ops.make_config(sb._lib.SCENARIO_OBSTACLE_AVOIDANCE, G, N, sb._lib.GRAPH_RADIUS, graph_radius=radius)
    w_on = sb.pack_weights(online.state_dict(), _dev())
    w_tg = sb.pack_weights(target.state_dict(), _dev())
    grad, loss, td = sb.ops.dqn_grad(cfg, w_on, w_tg, ring, None, G, want_td=True)
    assert abs(loss.item() - loss_ref) <= GRAD_RTOL * abs(loss_ref)
    got = sb.unpack_weights(grad.cpu())
    for name, gref in grads.items():
        g = got[name].reshape(gref.shape)
        bound = GRAD_RTOL * gref.abs().max().item() + 2e-6 * max(v.abs().max().item() for v in grads.values())
        assert (g - gref).abs().max().item() <= bound, name
