"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle only finishes small cases):

  * C2 (ObstacleAvoidance, 12 agents, 4 096 envs) and C3 (GoTo, 5..12 agents, 65 536 envs): envs are independent,
    so a full-size fused rollout must be bit-identical, env by env, to the same envs rolled out in small batches
    (a different grid / tile assignment), and a sample of its envs must match the CPU oracle;
  * a T-tick launch equals two T/2-tick launches (state resident in registers vs. written back in between);
  * the fused rollout with its own actions replayed equals T calls of the stand-alone world step (streaming kernel);
  * replay push -> gather over the whole ring is the identity; rollout-push equals explicit push;
  * DQN gradient: the gradient of a batch is the weighted sum of the gradients of its two halves;
  * kNN graphs at full size: every row holds its own node first (distance 0), rows are sorted by distance, the edge
    list is the symmetrised table and its sets equal torch's topk on the device where no distance tie exists.
"""
import numpy as np
import pytest
import torch

from helpers import load_params

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _setup(scen_name, B, N, mode="complete", k=5, seed=0):
    import swarm_b200 as sb
    from swarm_b200 import ops
    L = sb._lib
    scen = L.SCENARIO_OBSTACLE_AVOIDANCE if scen_name == "obstacle_avoidance" else L.SCENARIO_GOTO
    exp = "ObstacleAvoidance" if scen_name == "obstacle_avoidance" else "GoTo"
    cfg = ops.make_config(scen, B, N, L.GRAPH_KNN if mode == "knn" else L.GRAPH_COMPLETE, k)
    g = torch.Generator().manual_seed(seed)
    if scen_name == "obstacle_avoidance":
        centers = torch.tensor([0.6, -0.6]) + 0.1 * torch.randn(B, 2, generator=g)
    else:
        centers = torch.tensor([1.5, -1.5]) + torch.tensor([-0.6, 0.6]) + 0.4 * torch.randn(B, 2, generator=g)
    w = sb.pack_weights(load_params(exp, 0), _dev())
    return sb, ops, cfg, centers, w


@pytest.mark.parametrize("scen,B,N,mode,T", [("obstacle_avoidance", 4096, 12, "complete", 100),      # C2
                                             ("go_to", 65536, 12, "complete", 30),                    # C3
                                             ("go_to", 65536, 5, "knn", 30),
                                             ("go_to", 65536, 9, "knn", 20)])
def test_full_size_rollout_is_batch_invariant_and_matches_oracle(scen, B, N, mode, T):
    sb, ops, cfg, centers, w = _setup(scen, B, N, mode)
    dev = _dev()
    state = ops.reset_grid(cfg, centers.to(dev))
    out = ops.rollout(cfg, w, state, T)
    full_state, full_ret, full_hits = out["state"].clone(), out["returns"].clone(), out["hits"].clone()
    assert torch.isfinite(full_state).all()
    # the same envs in small batches (other tile / CTA assignment, other grid size): bit-identical
    for lo, n in ((0, 64), (B // 3, 37), (B - 129, 129)):
        sub = ops.clone_config(cfg, num_envs=n)
        st = ops.reset_grid(sub, centers[lo:lo + n].to(dev))
        o = ops.rollout(sub, w, st, T)
        assert torch.equal(o["state"], full_state[lo:lo + n]), f"envs {lo}..{lo + n} differ from the full-size launch"
        assert torch.equal(o["returns"], full_ret[lo:lo + n]) and torch.equal(o["hits"], full_hits[lo:lo + n])
    # a sample of envs against the CPU oracle over a short horizon
    from oracle import batched_oracle as bo, swarm_oracle as so
    idx = torch.arange(0, B, B // 16)[:16]
    T0 = 8
    pos, vel = bo.reset_grid(scen, centers[idx], N)
    ref = bo.rollout(scen, load_params("ObstacleAvoidance" if scen == "obstacle_avoidance" else "GoTo", 0), pos, vel, T0,
                     mode, 5)
    state = ops.reset_grid(cfg, centers.to(dev))
    o = ops.rollout(cfg, w, state, T0, trace=dict(actions=True))
    got_actions = o["trace_actions"][:, idx.to(dev)].cpu().long()
    same = (got_actions == ref["actions"]).all(dim=0).all(dim=1)                    # per sampled env
    # an env may only leave the oracle's action stream at a tick where the oracle itself has a float32 near-tie between
    # its two best actions (top-2 Q gap <= 2e-5 of max |Q|) -- the criterion of tests/test_gpu_golden_sweep.py
    for e in torch.nonzero(~same).flatten().tolist():
        diff = got_actions[:, e] != ref["actions"][:, e]                            # [T0, N]
        t0 = int(torch.nonzero(diff.any(dim=1))[0])
        q = ref["q"][t0, e]                                                         # [N, 9]
        top2 = q.topk(2, dim=-1).values
        gap = (top2[:, 0] - top2[:, 1]) / q.abs().amax(dim=-1)
        assert (gap[diff[t0]] <= 2e-5).all(), f"env {int(idx[e])}: action differs from the oracle at tick {t0} outside a near-tie"
    got_pos = o["state"][idx.to(dev)][..., :2].cpu()
    assert torch.allclose(got_pos[same], ref["pos"][-1][same], rtol=0, atol=1e-5)


@pytest.mark.parametrize("scen,B,N,mode", [("obstacle_avoidance", 4096, 12, "complete"), ("go_to", 16384, 7, "knn")])
def test_rollout_composes_in_time_and_equals_stepwise_world_steps(scen, B, N, mode):
    sb, ops, cfg, centers, w = _setup(scen, B, N, mode, seed=3)
    dev = _dev()
    T = 24
    s1 = ops.reset_grid(cfg, centers.to(dev))
    one = ops.rollout(cfg, w, s1, T, trace=dict(actions=True, rewards=True, flags=True))
    s2 = ops.reset_grid(cfg, centers.to(dev))
    a = ops.rollout(cfg, w, s2, T // 2)
    b = ops.rollout(cfg, w, s2, T // 2, returns=a["returns"], hits=a["hits"])
    assert torch.equal(one["state"], b["state"]) and torch.equal(one["hits"], b["hits"])
    torch.testing.assert_close(one["returns"], b["returns"], rtol=1e-5, atol=1e-5)   # (r1+..+rT) vs (r1+..)+(..+rT)
    # replay the recorded actions through the stand-alone (streaming) world step
    s3 = ops.reset_grid(cfg, centers.to(dev))
    hits = torch.zeros(B, dtype=torch.int64, device=dev)
    for t in range(T):
        o = ops.sim_step(cfg, s3, one["trace_actions"][t].contiguous(), state_out=s3, want_obs=False)
        assert torch.equal(o["rewards"], one["trace_rewards"][t]) and torch.equal(o["flags"], one["trace_flags"][t])
        hits += ((o["flags"] & sb._lib.FLAG_HIT) != 0).sum(dim=1)
    assert torch.equal(s3, one["state"]) and torch.equal(hits.int(), one["hits"])


def test_replay_ring_roundtrip_and_rollout_push_at_full_size():
    sb, ops, cfg, centers, w = _setup("obstacle_avoidance", 4096, 12)
    dev = _dev()
    B, N, T = cfg.num_envs, cfg.n_agents, 5
    ring = ops.ReplayRing(B * T, N, dev)
    state = ops.reset_grid(cfg, centers.to(dev))
    out = ops.rollout(cfg, w, state, T, epsilon=0.5, rng_seed=9, replay=ring, trace=dict(state=True, actions=True, rewards=True))
    assert len(ring) == B * T and ring.position == 0
    everything = ops.replay_gather(ring, torch.arange(B * T, device=dev))
    prev = torch.cat([ops.reset_grid(cfg, centers.to(dev)).unsqueeze(0), out["trace_state"][:-1]], dim=0)
    assert torch.equal(everything["next_state"].view(T, B, N, 4), out["trace_state"])
    assert torch.equal(everything["state"].view(T, B, N, 4), prev)
    assert torch.equal(everything["actions"].view(T, B, N), out["trace_actions"])
    assert torch.equal(everything["rewards"].view(T, B, N), out["trace_rewards"])
    # explicit push of the same transitions into a second ring gives the same bytes
    ring2 = ops.ReplayRing(B * T, N, dev)
    for t in range(T):
        ops.replay_push(cfg, ring2, prev[t].contiguous(), out["trace_actions"][t].contiguous(),
                        out["trace_rewards"][t].contiguous(), out["trace_state"][t].contiguous())
    for x, y in ((ring.state, ring2.state), (ring.next_state, ring2.next_state), (ring.actions, ring2.actions),
                 (ring.rewards, ring2.rewards)):
        assert torch.equal(x, y)
    # exploration actually happened and is a pure function of (seed, env, tick)
    greedy = ops.rollout(cfg, w, ops.reset_grid(cfg, centers.to(dev)), 1, trace=dict(actions=True))["trace_actions"][0]
    frac = (out["trace_actions"][0] != greedy).float().mean().item()
    assert 0.3 < frac < 0.6          # eps 0.5 x 8/9 of the random draws differ from the greedy action


def test_dqn_gradient_is_additive_over_the_batch():
    sb, ops, cfg, centers, w = _setup("obstacle_avoidance", 4096, 12)
    dev = _dev()
    B, N = cfg.num_envs, cfg.n_agents
    ring = ops.ReplayRing(B, N, dev)
    state = ops.reset_grid(cfg, centers.to(dev))
    ops.rollout(cfg, w, state, 30)                                  # move into a contact-rich region
    ops.rollout(cfg, w, state, 1, epsilon=0.5, rng_seed=1, replay=ring)
    w_t = sb.pack_weights(load_params("ObstacleAvoidance", 3), dev)
    G = 4096
    gcfg = ops.clone_config(cfg, num_envs=G)
    idx = torch.arange(G, device=dev)
    full, loss, _ = ops.dqn_grad(gcfg, w, w_t, ring, idx, G)
    h = G // 2
    hcfg = ops.clone_config(cfg, num_envs=h)
    g1, l1, _ = ops.dqn_grad(hcfg, w, w_t, ring, idx[:h].contiguous(), h, loss_scale=1.0 / (G * N))
    g2, l2, _ = ops.dqn_grad(hcfg, w, w_t, ring, idx[h:].contiguous(), h, loss_scale=1.0 / (G * N))
    scale = full.abs().max().item()
    assert (full - (g1 + g2)).abs().max().item() <= 2e-5 * scale
    assert abs(loss.item() - (l1.item() + l2.item())) <= 1e-5 * abs(loss.item())
    # and small-batch tiles (one graph per CTA, target / online side by side) agree with the packed tiles
    parts = torch.zeros_like(full)
    for c in range(0, 256, 32):
        gi, _, _ = ops.dqn_grad(ops.clone_config(cfg, num_envs=32), w, w_t, ring, idx[c:c + 32].contiguous(), 32,
                                loss_scale=1.0 / (256 * N))
        parts += gi
    g256, _, _ = ops.dqn_grad(ops.clone_config(cfg, num_envs=256), w, w_t, ring, idx[:256].contiguous(), 256)
    assert (g256 - parts).abs().max().item() <= 2e-5 * g256.abs().max().item()


@pytest.mark.parametrize("B,N,k", [(65536, 12, 5), (65536, 5, 5), (4096, 12, 10)])
def test_knn_graph_properties_at_full_size(B, N, k):
    sb, ops, cfg, centers, w = _setup("go_to", B, N, "knn", k)
    dev = _dev()
    state = ops.reset_grid(cfg, centers.to(dev))
    g = torch.Generator(device=dev).manual_seed(1)
    state[:, :, :2] += 0.02 * torch.randn(B, N, 2, device=dev, generator=g)          # tie-free with probability ~1
    edges, nbr = ops.graph_build(cfg, state, want_neighbours=True)
    E = ops.edges_per_env(cfg)
    assert edges.shape == (B, 2, E) and nbr.shape == (B, N, k)
    ids = torch.arange(N, device=dev).view(1, N).expand(B, N)
    assert torch.equal(nbr[:, :, 0].long(), ids), "the nearest neighbour of a node is itself (distance 0)"
    pos = state[:, :, :2]
    d = torch.sqrt(torch.addcmul((pos[:, :, None, 0] - pos[:, None, :, 0]) ** 2, pos[:, :, None, 1] - pos[:, None, :, 1],
                                 pos[:, :, None, 1] - pos[:, None, :, 1]))              # [B, i, j]
    dn = torch.gather(d, 2, nbr.long())
    assert (dn[:, :, 1:] >= dn[:, :, :-1]).all(), "rows are sorted by distance"
    ref = torch.topk(d, k, dim=2, largest=False).indices
    assert torch.equal(torch.sort(ref, dim=2).values, torch.sort(nbr.long(), dim=2).values), "neighbour sets differ"
    # edge list = for i, r: (i -> a), (a -> i); last (0 -> 0)   (simulator.py:20-24)
    src, dst = edges[:, 0, :-1].view(B, N, k, 2), edges[:, 1, :-1].view(B, N, k, 2)
    own = ids.view(B, N, 1).expand(B, N, k).int()
    assert torch.equal(src[..., 0], own) and torch.equal(dst[..., 0], nbr) and torch.equal(src[..., 1], nbr)
    assert torch.equal(dst[..., 1], own) and (edges[:, :, -1] == 0).all()


def test_c4_large_swarm_is_batch_invariant():
    """C4 (1 024 agents x 1 024 envs): kNN k = 10 edge lists, the world step and one tick of the large-swarm rollout
    for the full batch equal the same envs processed three at a time (and those are oracle-checked in
    test_gpu_large.py)."""
    import swarm_b200 as sb
    from swarm_b200 import ops
    dev = _dev()
    B, N, k = 1024, 1024, 10
    cfg = ops.make_config(sb._lib.SCENARIO_OBSTACLE_AVOIDANCE, B, N, sb._lib.GRAPH_KNN, k)
    g = torch.Generator().manual_seed(0)
    centers = (torch.tensor([0.6, -0.6]) + 0.1 * torch.randn(B, 2, generator=g)).to(dev)
    w = sb.pack_weights(load_params("ObstacleAvoidance", 0), dev)
    state = ops.reset_grid(cfg, centers)
    # squeeze the 4.65-wide grids so that contacts occur, jitter a little (most rows tie-free, some exact ties stay)
    ctr = state[:, :, :2].mean(dim=1, keepdim=True)
    gen = torch.Generator(device=dev).manual_seed(2)
    state[:, :, :2] = ctr + (state[:, :, :2] - ctr) * 0.6
    state[:, ::2, :2] += 0.01 * torch.randn(B, N // 2, 2, device=dev, generator=gen)
    edges, nbr = ops.graph_build(cfg, state, want_neighbours=True)
    actions = torch.randint(0, 9, (B, N), device=dev, dtype=torch.int32, generator=gen)
    stepped = ops.sim_step(cfg, state, actions, want_obs=False)
    rolled = ops.rollout_large(cfg, w, state.clone(), 1)
    sub = ops.clone_config(cfg, num_envs=3)
    for lo in (0, 511, B - 3):
        st = state[lo:lo + 3].contiguous()
        e3, n3 = ops.graph_build(sub, st, want_neighbours=True)
        assert torch.equal(e3, edges[lo:lo + 3]) and torch.equal(n3, nbr[lo:lo + 3])
        s3 = ops.sim_step(sub, st, actions[lo:lo + 3].contiguous(), want_obs=False)
        assert torch.equal(s3["state"], stepped["state"][lo:lo + 3]) and torch.equal(s3["rewards"], stepped["rewards"][lo:lo + 3])
        assert torch.equal(s3["flags"], stepped["flags"][lo:lo + 3])
        r3 = ops.rollout_large(sub, w, st.clone(), 1)
        assert torch.equal(r3["state"], rolled["state"][lo:lo + 3]) and torch.equal(r3["returns"], rolled["returns"][lo:lo + 3])
    assert (nbr >= 0).all() and (nbr < N).all()
