"""GPU parity of the DQN update path (replay ring, TD target + loss + hand-written backward, clip + Adam,
and the reference training loop) against the CPU oracle / torch autograd.

Bars: one-step gradients and loss within 1e-5 relative (per parameter tensor, relative to that tensor's
largest gradient magnitude); Adam-updated weights within 1e-6; replay copies bit-exact; the num_envs = 1
training loop reproduces the reference's shipped data/stats rows.
"""
import random

import numpy as np
import pytest
import torch
import torch.nn as nn

from helpers import load_params, npz

pytestmark = pytest.mark.gpu
GRAD_RTOL = 1e-5


def _dev():
    return torch.device("cuda:0")


def _transitions(scenario, G, N, seed):
    """G whole-swarm transitions (s, a, r, s') from oracle steps with random actions on crowded states."""
    from oracle import batched_oracle as bo
    from test_gpu_parity import _random_states
    pos, vel = _random_states(scenario, G, N, seed=seed, crowd=True)
    g = torch.Generator().manual_seed(seed)
    actions = torch.randint(0, 9, (G, N), generator=g)
    out = bo.step(scenario, pos, vel, actions)
    return pos, vel, actions, out["rewards"], out["pos"], out["vel"]


def _oracle_loss_and_grads(exp, model_seed, scenario, pos, vel, actions, rewards, pos2, vel2, mode, k, gamma=0.99,
                           dtype=torch.float32):
    from oracle import batched_oracle as bo
    from oracle.dqn_oracle import OracleGCN
    G, N, _ = pos.shape
    online = OracleGCN(7, 32, 9)
    online.load_state_dict(load_params(exp, model_seed))
    target = OracleGCN(7, 32, 9)
    target.load_state_dict(load_params(exp, (model_seed + 1) % 10))
    online, target = online.to(dtype), target.to(dtype)
    x = bo.node_features(pos, vel).reshape(G * N, 7).to(dtype)
    x2 = bo.node_features(pos2, vel2).reshape(G * N, 7).to(dtype)
    ei = bo.batch_edge_index(bo.graph_edges(pos, mode, k), N)
    ei2 = bo.batch_edge_index(bo.graph_edges(pos2, mode, k), N)
    values = online(x, ei).gather(1, actions.reshape(-1, 1))
    next_values = target(x2, ei2).max(dim=1)[0].detach()
    target_values = rewards.reshape(-1).to(dtype) + (gamma * next_values)
    loss = nn.MSELoss()(values, target_values.unsqueeze(1))
    loss.backward()
    grads = {k_: v.grad.detach().clone() for k_, v in online.named_parameters()}
    td = (values.detach().reshape(-1) - target_values).reshape(G, N)
    return loss.item(), grads, td, online.float(), target.float()


def _ring_from(sb, pos, vel, actions, rewards, pos2, vel2, capacity=None):
    G, N, _ = pos.shape
    ring = sb.ops.ReplayRing(capacity or G, N, _dev())
    cfg = sb.ops.make_config(sb._lib.SCENARIO_GOTO, G, N)
    sb.ops.replay_push(cfg, ring, torch.cat([pos, vel], 2).contiguous().to(_dev()), actions.to(torch.int32).to(_dev()),
                       rewards.contiguous().to(_dev()), torch.cat([pos2, vel2], 2).contiguous().to(_dev()))
    return ring


def test_replay_push_gather_bitexact():
    import swarm_b200 as sb
    pos, vel, actions, rewards, pos2, vel2 = _transitions("obstacle_avoidance", 40, 7, seed=1)
    ring = _ring_from(sb, pos, vel, actions, rewards, pos2, vel2, capacity=64)
    assert len(ring) == 40 and ring.position == 40
    # wrap-around: push again, slots 40..63 then 0..15
    cfg = sb.ops.make_config(sb._lib.SCENARIO_GOTO, 40, 7)
    sb.ops.replay_push(cfg, ring, torch.cat([pos2, vel2], 2).contiguous().to(_dev()), actions.to(torch.int32).to(_dev()),
                       rewards.contiguous().to(_dev()), torch.cat([pos, vel], 2).contiguous().to(_dev()))
    assert len(ring) == 64 and ring.position == 16
    idx = torch.tensor([16, 39, 40, 63, 0, 15, 39], dtype=torch.int64, device=_dev())
    b = sb.ops.replay_gather(ring, idx)
    s_first = torch.cat([pos, vel], 2)
    s_second = torch.cat([pos2, vel2], 2)
    exp_state = torch.stack([s_first[16], s_first[39], s_second[0], s_second[23], s_second[24], s_second[39], s_first[39]])
    assert torch.equal(b["state"].cpu(), exp_state)
    exp_act = torch.stack([actions[16], actions[39], actions[0], actions[23], actions[24], actions[39], actions[39]])
    assert torch.equal(b["actions"].cpu().long(), exp_act)
    exp_rew = torch.stack([rewards[16], rewards[39], rewards[0], rewards[23], rewards[24], rewards[39], rewards[39]])
    assert torch.equal(b["rewards"].cpu(), exp_rew)


@pytest.mark.parametrize("exp,scenario", [("GoTo", "go_to"), ("ObstacleAvoidance", "obstacle_avoidance")])
@pytest.mark.parametrize("G,N,mode,k", [(32, 5, "complete", 0), (32, 12, "complete", 0), (100, 12, "knn", 5), (7, 9, "knn", 9),
                                        (1900, 12, "complete", 0), (1000, 20, "knn", 6),    # sequential-pass tiles

                                        (300, 5, "complete", 0), (5, 32, "complete", 0)])
def test_dqn_grad_parity(exp, scenario, G, N, mode, k):
    import swarm_b200 as sb
    pos, vel, actions, rewards, pos2, vel2 = _transitions(scenario, G, N, seed=G + N)
    loss_ref, grads_ref, td_ref, online, target = _oracle_loss_and_grads(exp, 3, scenario, pos, vel, actions, rewards,
                                                                         pos2, vel2, mode, k)
    ring = _ring_from(sb, pos, vel, actions, rewards, pos2, vel2)
    gm = sb._lib.GRAPH_KNN if mode == "knn" else sb._lib.GRAPH_COMPLETE
    cfg = sb.ops.make_config(sb._lib.SCENARIO_GOTO if scenario == "go_to" else sb._lib.SCENARIO_OBSTACLE_AVOIDANCE, G, N,
                             gm, max(k, 1))
    w_on = sb.pack_weights(online.state_dict(), _dev())
    w_tg = sb.pack_weights(target.state_dict(), _dev())
    grad, loss, td = sb.ops.dqn_grad(cfg, w_on, w_tg, ring, None, G, want_td=True)
    assert abs(loss.item() - loss_ref) <= GRAD_RTOL * abs(loss_ref), f"loss {loss.item()} vs {loss_ref}"
    scale = td_ref.abs().max().item()
    assert (td.cpu() - td_ref).abs().max().item() <= 2e-5 * max(scale, 1.0), "TD errors differ"
    # gradients: against the float64 evaluation of the same graph (the float32 autograd of the oracle is itself
    # only accurate to ~5e-7 of the global gradient scale; d att_dst in particular is a cancellation residue --
    # the softmax logit gradients of a node sum to zero -- whose float32 value is mostly rounding noise).
    # Bound per tensor: 1e-5 of that tensor's largest gradient + 2e-6 of the largest gradient overall.
    _, grads64, _, _, _ = _oracle_loss_and_grads(exp, 3, scenario, pos, vel, actions, rewards, pos2, vel2, mode, k,
                                                 dtype=torch.float64)
    gmax = max(v.abs().max().item() for v in grads64.values())
    got = sb.unpack_weights(grad.cpu())
    for name, gref in grads64.items():
        ggpu = got[name].reshape(gref.shape).double()
        err = (ggpu - gref).abs().max().item()
        bound = GRAD_RTOL * gref.abs().max().item() + 2e-6 * gmax
        assert err <= bound, f"{name}: gradient abs error {err:.3e} > bound {bound:.3e} (|g|max {gref.abs().max().item():.3e})"
        err32 = (grads_ref[name].double() - gref).abs().max().item()
        assert err <= max(20 * err32, GRAD_RTOL * gref.abs().max().item()), f"{name}: far noisier than float32 autograd"
    # deterministic: a second call gives the same bits
    grad2, _, _ = sb.ops.dqn_grad(cfg, w_on, w_tg, ring, None, G)
    assert torch.equal(grad, grad2)
    # indices: a permuted / repeated gather equals the dense batch built the same way
    perm = torch.randperm(G, generator=torch.Generator().manual_seed(0))
    grad3, loss3, _ = sb.ops.dqn_grad(cfg, w_on, w_tg, ring, perm.to(_dev()), G)
    err = (grad3 - grad).abs().max().item() / grad.abs().max().item()
    assert err <= GRAD_RTOL and abs(loss3.item() - loss.item()) <= GRAD_RTOL * abs(loss.item())


def test_adam_clip_step_parity():
    import swarm_b200 as sb
    torch.manual_seed(0)
    params = load_params("ObstacleAvoidance", 5)
    from oracle.dqn_oracle import OracleGCN
    model = OracleGCN(7, 32, 9)
    model.load_state_dict(params)
    opt = torch.optim.Adam(model.parameters(), 0.001)
    w = sb.pack_weights(model.state_dict(), _dev())
    m = torch.zeros_like(w)
    v = torch.zeros_like(w)
    target = torch.zeros_like(w)
    norm_out = torch.zeros(1, device=_dev())
    g = torch.Generator().manual_seed(1)
    for step in range(1, 8):
        scale = [5.0, 0.01, 1.0, 300.0, 1e-4, 2.0, 0.5][step - 1]         # both clipped and un-clipped steps
        flat = torch.randn(1673, generator=g) * scale
        opt.zero_grad()
        parts = sb.unpack_weights(flat)
        for name, p_ in model.named_parameters():
            p_.grad = parts[name].reshape(p_.shape).clone()
        ref_norm = torch.nn.utils.clip_grad_norm_(model.parameters(), 1)
        opt.step()
        sb.ops.adam_clip_step(w, flat.to(_dev()), m, v, step, target=target if step == 4 else None, grad_norm=norm_out)
        ref = sb.pack_weights(model.state_dict())
        err = ((w.cpu() - ref).abs() / ref.abs().clamp_min(1e-3)).max().item()
        assert err <= 1e-6, f"step {step}: weights relative error {err:.3e}"
        assert abs(norm_out.item() - ref_norm.item()) <= 1e-6 * ref_norm.item()
        if step == 4:
            assert torch.equal(target, w)
    st = opt.state_dict()["state"]
    order = [3, 0, 1, 2, 4, 5, 6, 7]       # parameters(): att_src, att_dst, bias, lin.weight, ... -> packed order
    ref_m = torch.cat([st[i]["exp_avg"].reshape(-1) for i in order])
    ref_v = torch.cat([st[i]["exp_avg_sq"].reshape(-1) for i in order])
    assert (m.cpu() - ref_m).abs().max().item() <= 1e-6 * ref_m.abs().max().item()
    assert (v.cpu() - ref_v).abs().max().item() <= 1e-6 * ref_v.abs().max().item()


@pytest.mark.parametrize("exp,seed", [("ObstacleAvoidance", 0), ("GoTo", 0), ("ObstacleAvoidance", 4)])
def test_training_loop_reproduces_reference_stats(exp, seed):
    """The reference training loop (N = 5, num_envs = 1, Python `random` exploration + sampling) on the CUDA
    path against the shipped data/stats/experiment_{exp}-seed_{seed}.csv: per-episode mean loss (rows 0..) and
    the first 10-episode reward row."""
    import swarm_b200 as sb
    stats = npz("train_stats.npz")[f"{exp}/{seed}"]
    sb.set_seed(seed)
    scenario = sb.GoToPositionScenario() if exp == "GoTo" else sb.ObstacleAvoidanceScenario()
    env = sb.make_env(scenario=scenario, num_envs=1, device="cuda:0", continuous_actions=False, wrapper=None,
                      max_steps=100, dict_spaces=True, n_agents=5, seed=seed)
    trainer = sb.DQNTrainer(env, seed, "/tmp/swarm_models", "/tmp/swarm_stats", exp, replay_capacity=4096)
    trainer.train_model({"epsilon": 0.99, "epsilon_decay": 0.01, "min_epsilon": 0.05, "episodes": 10, "verbose": False,
                         "save": False})
    loss0 = trainer.episode_losses[0]
    assert abs(loss0 - stats[0, 2]) <= 1e-4 * abs(stats[0, 2]), f"episode-0 loss {loss0} vs golden {stats[0, 2]}"
    row0 = trainer.episode_rewards[0].item()
    print(f"{exp} seed {seed}: episode-0 loss {loss0} (golden {stats[0, 2]}), first-row reward {row0} (golden {stats[0, 1]})")
    assert abs(row0 - stats[0, 1]) <= 2e-2 * abs(stats[0, 1])


# ---- device-driven train tick (swarm_train_tick_grad / swarm_train_tick_apply) ----------------------------------
def _tick_setup(sb, B, N, G, capacity, scen_id, seed=3):
    from swarm_b200 import ops
    dev = _dev()
    cfg = ops.make_config(scen_id, B, N)
    g = torch.Generator().manual_seed(seed)
    centers = (torch.tensor([0.6, -0.6]) + 0.1 * torch.randn(B, 2, generator=g)).to(dev)
    state = ops.reset_grid(cfg, centers)
    w = sb.pack_weights(load_params("ObstacleAvoidance", 0), dev)
    ring = ops.ReplayRing(capacity, N, dev)
    return cfg, state, w, ring


@pytest.mark.parametrize("B,N,G,capacity", [(64, 12, 32, 1000), (16, 5, 32, 40)])
def test_train_tick_matches_composed_calls(B, N, G, capacity):
    """The device-driven tick == swarm_rollout(1 tick, push) + swarm_dqn_grad(its own index draw) + swarm_adam_clip_step,
    including the 'not enough samples' ticks, the ring wrap-around and the hard target sync."""
    import swarm_b200 as sb
    from swarm_b200 import ops
    dev = _dev()
    cfg, state, w, ring = _tick_setup(sb, B, N, G, capacity, sb._lib.SCENARIO_OBSTACLE_AVOIDANCE)
    every = 3
    tt = ops.TrainTick(cfg, ring, graphs_per_update=G, update_target_every=every, rng_seed=11, sample_seed=5, env_offset=7)
    tt.load_cursor(0, 0, 0.4)
    w_t = w.clone(); m = torch.zeros_like(w); v = torch.zeros_like(w)
    returns = torch.zeros(B, N, device=dev); hits = torch.zeros(B, dtype=torch.int32, device=dev)

    # composed reference on copies
    state_c, w_c, wt_c, m_c, v_c = state.clone(), w.clone(), w.clone(), torch.zeros_like(w), torch.zeros_like(w)
    ring_c = ops.ReplayRing(capacity, N, dev)
    ret_c = torch.zeros(B, N, device=dev); hits_c = torch.zeros(B, dtype=torch.int32, device=dev)
    gcfg = ops.clone_config(cfg, num_envs=G)
    opt = 0
    for tick in range(1, 9):
        tt.grad_phase(w, w_t, state, returns, hits)
        cur = tt.read_cursor()          # before apply: tick not advanced yet, `updating` set by the sampler
        assert cur["tick"] == tick - 1
        ops.rollout(cfg, w_c, state_c, 1, returns=ret_c, hits=hits_c, epsilon=0.4, rng_seed=11, rng_tick0=tick, env_offset=7,
                    replay=ring_c)
        assert torch.equal(state, state_c) and torch.equal(returns, ret_c) and torch.equal(hits, hits_c)
        for a, b in ((ring.state, ring_c.state), (ring.next_state, ring_c.next_state), (ring.actions, ring_c.actions),
                     (ring.rewards, ring_c.rewards)):
            assert torch.equal(a, b)
        enough = len(ring_c) >= G
        assert cur["updating"] == int(enough)
        if enough:
            idx = tt.indices.clone()
            assert int(idx.min()) >= 0 and int(idx.max()) < len(ring_c)
            grad_c, loss_c = ops.dqn_grad(gcfg, w_c, wt_c, ring_c, idx, G)[:2]
            assert torch.equal(tt.grad, grad_c.reshape(-1)) and torch.equal(tt.loss, loss_c.reshape(-1))
            opt += 1
            ops.adam_clip_step(w_c, grad_c, m_c, v_c, opt, 1e-3, (0.9, 0.999), 1e-8, 1.0,
                               target=wt_c if tick % every == 0 else None)
        tt.apply_phase(w, w_t, m, v)
        cur = tt.read_cursor()
        assert cur["tick"] == tick and cur["opt_step"] == opt
        assert ring.position == ring_c.position and ring.size == ring_c.size
        torch.testing.assert_close(w, w_c, rtol=1e-6, atol=1e-9)
        torch.testing.assert_close(w_t, wt_c, rtol=1e-6, atol=1e-9)
        w_c.copy_(w); wt_c.copy_(w_t); m_c.copy_(m); v_c.copy_(v)     # keep the two tracks on identical weights
    assert opt >= 5
    # sampled slots cover the ring roughly uniformly
    assert tt.indices.unique().numel() > G // 2


@pytest.mark.parametrize("B,N,G,capacity,graph", [(64, 12, 32, 1000, "complete"), (16, 5, 32, 40, "complete"),
                                                   (64, 12, 32, 1000, "knn"), (256, 12, 1024, 4096, "complete")])
def test_one_call_tick_equals_the_two_phases(B, N, G, capacity, graph):
    """swarm_train_tick (partials summed inside clip + Adam for <= 64 gradient CTAs, else the reduce launch) leaves the
    same bits as swarm_train_tick_grad + swarm_train_tick_apply: state, ring, cursor, gradient, loss, moments, weights,
    over 'not enough samples' ticks, updates, the ring wrap-around and target syncs."""
    import swarm_b200 as sb
    from swarm_b200 import ops
    dev = _dev()
    cfg, state_a, w_a, ring_a = _tick_setup(sb, B, N, G, capacity, sb._lib.SCENARIO_OBSTACLE_AVOIDANCE)
    if graph == "knn":
        cfg = ops.clone_config(cfg, graph_mode=sb._lib.GRAPH_KNN, knn_k=5)
    ring_b = ops.ReplayRing(capacity, N, dev)
    state_b, w_b = state_a.clone(), w_a.clone()

    def track(ring):
        tt = ops.TrainTick(cfg, ring, graphs_per_update=G, update_target_every=3, rng_seed=11, sample_seed=5, env_offset=7)
        tt.load_cursor(0, 0, 0.4)
        return tt
    tt_a, tt_b = track(ring_a), track(ring_b)
    bufs = lambda w: (w.clone(), torch.zeros_like(w), torch.zeros_like(w), torch.zeros(B, N, device=dev),
                      torch.zeros(B, dtype=torch.int32, device=dev))
    wt_a, m_a, v_a, ret_a, hits_a = bufs(w_a)
    wt_b, m_b, v_b, ret_b, hits_b = bufs(w_b)
    updates = 0
    for tick in range(1, 13):
        tt_a.grad_phase(w_a, wt_a, state_a, ret_a, hits_a)
        tt_a.apply_phase(w_a, wt_a, m_a, v_a)
        tt_b.tick(w_b, wt_b, m_b, v_b, state_b, ret_b, hits_b)
        ca, cb = tt_a.read_cursor(), tt_b.read_cursor()
        assert ca == cb and ca["tick"] == tick
        updates += ca["updating"]
        pairs = [(state_a, state_b), (ret_a, ret_b), (hits_a, hits_b), (w_a, w_b), (wt_a, wt_b), (m_a, m_b), (v_a, v_b),
                 (ring_a.state, ring_b.state), (ring_a.actions, ring_b.actions), (ring_a.rewards, ring_b.rewards),
                 (tt_a.indices, tt_b.indices)]
        if ca["updating"]:
            pairs.append((tt_a.grad_loss, tt_b.grad_loss))
        for x, y in pairs:
            assert torch.equal(x, y), tick
    assert updates >= 8 and ca["opt_step"] == updates


def test_train_model_batched_graph_equals_eager():
    """An episode replayed from the captured CUDA graph gives bit-identical weights to eager launches."""
    import swarm_b200 as sb
    out = []
    for use_graph in (False, True):
        random.seed(0); torch.manual_seed(0)
        env = sb.make_env(sb.ObstacleAvoidanceScenario(), num_envs=256, device="cuda", continuous_actions=False,
                          max_steps=12, dict_spaces=True, seed=0, n_agents=12, random=True)
        tr = sb.DQNTrainer(env, 0, "/tmp/none", "/tmp/none", "t", replay_capacity=4096)
        stats = tr.train_model_batched({"episodes": 4, "epsilon": 0.9, "epsilon_decay": 0.3, "min_epsilon": 0.05,
                                        "graphs_per_update": 64, "update_target_every": 5, "cuda_graph": use_graph})
        torch.cuda.synchronize()
        out.append((tr.w.clone(), tr.w_target.clone(), stats))
    assert out[0][2]["ticks"] == 48 and out[0][2]["opt_steps"] == 48
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])
    assert out[0][2] == out[1][2]


def test_reset_random_statistics_and_determinism():
    """Device-side reset: centres follow base + N(mean, std^2) per env and episode, are reproducible, differ between
    episodes / envs, and the agents sit on the exact start grid around them (bit-equal to swarm_reset_grid)."""
    import swarm_b200 as sb
    from swarm_b200 import ops
    dev = _dev()
    B, N = 20000, 7
    for scen, mean, std in ((sb._lib.SCENARIO_GOTO, (0.9, -0.9), 0.4), (sb._lib.SCENARIO_OBSTACLE_AVOIDANCE, (0.6, -0.6), 0.1)):
        cfg = ops.make_config(scen, B, N)
        spec = ops.reset_spec(scen, random=True, seed=5, env_offset=100)
        state = torch.empty(B, N, 4, device=dev)
        c0 = torch.empty(B, 2, device=dev)
        ops.reset_random(cfg, spec, state, episode=3, centers_out=c0)
        assert torch.equal(state, ops.reset_grid(cfg, c0))
        m, s = c0.mean(0).cpu(), c0.std(0).cpu()
        assert abs(m[0] - mean[0]) < 4 * std / B ** 0.5 + 1e-3 and abs(m[1] - mean[1]) < 4 * std / B ** 0.5 + 1e-3
        assert abs(s[0] - std) < 0.03 * std and abs(s[1] - std) < 0.03 * std
        assert abs(torch.corrcoef(c0.T.cpu())[0, 1]) < 0.03
        c1 = torch.empty(B, 2, device=dev)
        ops.reset_random(cfg, spec, state, episode=3, centers_out=c1)
        assert torch.equal(c0, c1)
        ops.reset_random(cfg, spec, state, episode=4, centers_out=c1)
        assert not torch.equal(c0, c1)
        # env shards: the draw depends on the global env index only
        half = ops.make_config(scen, B // 2, N)
        spec2 = ops.reset_spec(scen, random=True, seed=5, env_offset=100 + B // 2)
        c2 = torch.empty(B // 2, 2, device=dev)
        ops.reset_random(half, spec2, state[:B // 2].contiguous(), episode=3, centers_out=c2)
        assert torch.equal(c2, c0[B // 2:])
        shared = ops.reset_spec(scen, random=True, seed=5, shared_center=True)
        ops.reset_random(cfg, shared, state, episode=0, centers_out=c1)
        assert (c1 == c1[0]).all()
    fixed = ops.reset_spec(sb._lib.SCENARIO_OBSTACLE_AVOIDANCE, random=False)
    ops.reset_random(cfg, fixed, state, episode=9, centers_out=c1)
    assert (c1.cpu() == torch.tensor([0.6, -0.6])).all()


def test_train_model_device_graph_equals_eager_and_host_loop():
    """Whole-run-on-device training: graph replays == eager launches (bit-exact weights and statistics), and the
    statistics rows equal a host-side recomputation (returns mean, epsilon schedule of train:180)."""
    import swarm_b200 as sb
    cfgd = {"episodes": 5, "epsilon": 0.9, "epsilon_decay": 0.3, "min_epsilon": 0.05, "graphs_per_update": 64,
            "update_target_every": 7}
    out = []
    for use_graph in (False, True):
        random.seed(0); torch.manual_seed(0)
        env = sb.make_env(sb.ObstacleAvoidanceScenario(), num_envs=256, device="cuda", continuous_actions=False,
                          max_steps=10, dict_spaces=True, seed=0, n_agents=12, random=True)
        tr = sb.DQNTrainer(env, 0, "/tmp/none", "/tmp/none", "t", replay_capacity=2048)
        stats = tr.train_model_device(dict(cfgd, cuda_graph=use_graph))
        torch.cuda.synchronize()
        out.append((tr.w.clone(), tr.w_target.clone(), stats, tr.opt_step))
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][2], out[1][2])
    assert out[0][3] == 50
    stats = out[0][2]
    eps = [0.9] + [max(0.05, 0.9 * float(np.exp(-0.3 * e))) for e in range(4)]
    assert np.allclose(stats[:, 3].numpy(), np.array(eps, dtype=np.float32), rtol=1e-6)
    assert (stats[:, 0] < 0).all() and (stats[:, 2] > 0).all() and (stats[:, 1] >= 0).all()
