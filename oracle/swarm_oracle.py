"""Single-env CPU oracle (TEST INFRASTRUCTURE -- see oracle/__init__.py).

A torch-CPU float32 restatement, structured like the reference's ``num_envs = 1`` execution, of

  * the two scenarios                     src/scenarios/go_to_position_scenario.py:9-149,
                                          src/scenarios/obstacle_avoidance_scenario.py:9-181
  * the vmas==1.4.0 world step they run on (requirements.txt:3; call sites
    src/training/train_gcn_dqn.py:169 and src/simulation/simulator.py:68).  vmas is not vendored in
    the reference and is not installable here, so its published algorithm (``Environment.step`` /
    ``_set_action``, ``World.step``, ``_get_constraint_forces``, ``_integrate_state``,
    ``get_distance``) is restated below; SURVEY.md Appendix A is the spec.
  * the two graph builders                src/training/train_gcn_dqn.py:94-110 (complete),
                                          src/simulation/simulator.py:9-26 (symmetrised kNN)
  * the Q-network                         src/training/train_gcn_dqn.py:50-70 on torch_geometric==2.5.3
                                          ``GATConv`` (requirements.txt:2), restated from its published
                                          algorithm (GATConv.forward / utils.softmax / scatter).
  * the greedy evaluation loop            src/simulation/simulator.py:47-109

Every torch op below is one separately-rounded float32 op, exactly as the reference issues them;
that ordering is what the CUDA kernels are checked against.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

GOTO = "go_to"
OBSTACLE_AVOIDANCE = "obstacle_avoidance"

# vmas World defaults
DT = 0.1
DRAG = 0.25
COLLISION_FORCE = 100
CONTACT_MARGIN = 1e-3
SPHERE_RADIUS = 0.05
MIN_DIST = 1e-6
# scenario constants
GOAL_POS = (-0.8, 0.8)        # go_to:86, oa:97
OBSTACLE_POS = (-0.1, 0.1)    # oa:99
DESIRED_DISTANCE = 0.15       # go_to:17, oa:23
HIT_DISTANCE = 0.2            # oa:25 (min_collision_distance_count)
PENALTY_DISTANCE = 1          # oa:24 (min_collision_distance_reward)
OBSTACLE_WEIGHT = 2.5         # oa:136


# --------------------------------------------------------------------------------------------
# scenario reset
# --------------------------------------------------------------------------------------------
def generate_grid(center: torch.Tensor, num_points: int, distance: float = DESIRED_DISTANCE) -> torch.Tensor:
    """go_to:52-80 / oa:63-91.  ``center`` is f32[2]; offsets are Python doubles added to 0-d f32 tensors."""
    x_center, y_center = center
    num_cols = math.ceil(math.sqrt(num_points))
    num_rows = math.ceil(num_points / num_cols)
    grid = []
    for i in range(num_rows):
        for j in range(num_cols):
            x = x_center + (j - (num_cols - 1) / 2) * distance
            y = y_center + (i - (num_rows - 1) / 2) * distance
            grid.append([x, y])
            if len(grid) >= num_points:
                break
        if len(grid) >= num_points:
            break
    return torch.tensor(grid)


def draw_center(scenario: str, random: bool = True, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """One start-centre draw, consuming the torch CPU generator exactly like reset_world_at
    (go_to:84-88, oa:100-102)."""
    if scenario == GOTO:
        position_range = -torch.tensor([-1.5, 1.5])
        return position_range + torch.normal(mean=torch.tensor([-0.6, 0.6]), std=torch.tensor([0.4, 0.4]),
                                             generator=generator)
    if scenario == OBSTACLE_AVOIDANCE:
        delta = (torch.normal(mean=torch.tensor([0.0, 0.0]), std=torch.tensor([0.1, 0.1]), generator=generator)
                 if random else torch.tensor([0.0, 0.0]))
        return torch.tensor([0.6, -0.6]) + delta
    raise ValueError(f"unknown scenario {scenario!r}")


# --------------------------------------------------------------------------------------------
# world step (vmas 1.4.0 semantics, batch_dim = 1)
# --------------------------------------------------------------------------------------------
def decode_action(a: torch.Tensor) -> torch.Tensor:
    """vmas Environment._set_action for ``discrete_action_nvec = [3, 3]``: flat a in [0, 8] ->
    (a // 3, a % 3); index 0 -> u = 0, 1 -> -1, 2 -> +1 (u_range = u_multiplier = 1).
    ``a`` is int64[...]; returns f32[..., 2]."""
    comps = []
    flat = a
    for n_rest in (3, 1):
        idx = flat // n_rest
        flat = flat % n_rest
        idx = idx.clone()
        stay = idx == 0
        decrement = (idx > 0) & (idx <= 1)
        idx[stay] = 1
        idx[decrement] -= 1
        comps.append((idx / 2) * 2.0 - 1.0)
    return torch.stack(comps, dim=-1).to(torch.float32)


def constraint_force(pos_a: torch.Tensor, pos_b: torch.Tensor) -> torch.Tensor:
    """vmas World._get_constraint_forces for sphere-sphere pairs (force on a; b receives the negation).
    pos_* are f32[..., 2]."""
    dist_min = torch.tensor(SPHERE_RADIUS) + torch.tensor(SPHERE_RADIUS)
    delta_pos = pos_a - pos_b
    dist = torch.linalg.vector_norm(delta_pos, dim=-1)
    k = CONTACT_MARGIN
    penetration = torch.logaddexp(torch.tensor(0.0, dtype=torch.float32), (dist_min - dist) * 1 / k) * k
    force = (1 * COLLISION_FORCE * delta_pos
             / torch.where(dist > 0, dist, 1e-8).unsqueeze(-1)
             * penetration.unsqueeze(-1))
    force = torch.where((dist < MIN_DIST).unsqueeze(-1), 0.0, force)
    force = torch.where((dist > dist_min).unsqueeze(-1), 0.0, force)
    return force


def get_distance(pos_a: torch.Tensor, pos_b: torch.Tensor) -> torch.Tensor:
    """vmas World.get_distance for two spheres: centre distance minus both radii, subtracted one
    after the other."""
    return (torch.linalg.vector_norm(pos_a - pos_b, dim=-1) - SPHERE_RADIUS) - SPHERE_RADIUS


class OracleWorld:
    """One env (batch_dim = 1) of GoTo / ObstacleAvoidance, entity order = landmarks then agents."""

    def __init__(self, scenario: str, n_agents: int, random: bool = False, max_steps: Optional[int] = None):
        if scenario not in (GOTO, OBSTACLE_AVOIDANCE):
            raise ValueError(f"unknown scenario {scenario!r}")
        self.scenario = scenario
        self.n_agents = n_agents
        self.random = random
        self.max_steps = max_steps
        self.goal = torch.zeros(1, 2)
        self.obstacle = torch.zeros(1, 2) if scenario == OBSTACLE_AVOIDANCE else None
        self.pos = [torch.zeros(1, 2) for _ in range(n_agents)]
        self.vel = [torch.zeros(1, 2) for _ in range(n_agents)]
        self.distance_to_goal = [None] * n_agents
        self.steps = 0
        self.last_contacts: List[Tuple[int, int]] = []   # (a, b) entity pairs in contact at the last step; -1 = obstacle

    # -- reset ------------------------------------------------------------------------------
    def reset(self, center: Optional[torch.Tensor] = None) -> torch.Tensor:
        """vmas world.reset (zero pos/vel) + scenario.reset_world_at(None).  If ``center`` is None it is
        drawn from the global torch generator (reference behaviour)."""
        self.steps = 0
        self.goal = torch.tensor(list(GOAL_POS)).unsqueeze(0)
        if self.obstacle is not None:
            self.obstacle = torch.tensor(list(OBSTACLE_POS)).unsqueeze(0)
        if center is None:
            center = draw_center(self.scenario, self.random)
        grid = generate_grid(center, self.n_agents)
        for i in range(self.n_agents):
            self.pos[i] = grid[i].unsqueeze(0).clone()
            self.vel[i] = torch.zeros(1, 2)
        return self.observations()

    # -- step -------------------------------------------------------------------------------
    def step(self, actions: torch.Tensor) -> torch.Tensor:
        """``actions`` int64[N].  Returns rewards f32[N] (reward of agent i computed in agent order)."""
        n = self.n_agents
        u = decode_action(actions.to(torch.int64))
        forces = [torch.zeros(1, 2) + u[i].unsqueeze(0) for i in range(n)]

        # vmas _apply_vectorized_enviornment_force: pairs a < b over [landmarks..., agents...]
        pairs: List[Tuple[int, int]] = []
        if self.obstacle is not None:
            for i in range(n):
                if self._collides(self.obstacle, self.pos[i]):
                    pairs.append((-1, i))
        for i in range(n):
            for j in range(i + 1, n):
                if self._collides(self.pos[i], self.pos[j]):
                    pairs.append((i, j))
        self.last_contacts = pairs
        if pairs:
            pos_a = torch.stack([self.obstacle if a < 0 else self.pos[a] for a, _ in pairs], dim=-2)
            pos_b = torch.stack([self.pos[b] for _, b in pairs], dim=-2)
            force_a = constraint_force(pos_a, pos_b)
            force_b = -force_a
            for p, (a, b) in enumerate(pairs):
                if a >= 0:
                    forces[a] = forces[a] + force_a[:, p]
                forces[b] = forces[b] + force_b[:, p]

        # vmas _integrate_state (substeps = 1)
        for i in range(n):
            self.vel[i] = self.vel[i] * (1 - DRAG)
            accel = forces[i] / 1.0
            self.vel[i] = self.vel[i] + accel * DT
            self.pos[i] = self.pos[i] + self.vel[i] * DT
        self.steps += 1
        return self.rewards()

    @staticmethod
    def _collides(pos_a: torch.Tensor, pos_b: torch.Tensor) -> bool:
        return bool((torch.linalg.vector_norm(pos_a - pos_b, dim=-1) <= SPHERE_RADIUS + SPHERE_RADIUS).any())

    # -- scenario callbacks -----------------------------------------------------------------
    def rewards(self) -> torch.Tensor:
        n = self.n_agents
        out = []
        if self.scenario == GOTO:
            collective = 0                                          # go_to:108-115
            for i in range(n):
                self.distance_to_goal[i] = torch.linalg.vector_norm(self.pos[i] - self.goal, dim=-1)
                collective = collective + (-self.distance_to_goal[i])
            out = [collective.clone() for _ in range(n)]
        else:
            for i in range(n):                                      # oa:135-152
                self.distance_to_goal[i] = torch.linalg.vector_norm(self.pos[i] - self.goal, dim=-1)
                d_obs = get_distance(self.pos[i], self.obstacle)
                if d_obs <= PENALTY_DISTANCE:
                    avoid = -(PENALTY_DISTANCE - get_distance(self.pos[i], self.obstacle))
                else:
                    avoid = 0
                out.append((-self.distance_to_goal[i]) + OBSTACLE_WEIGHT * avoid)
        return torch.cat(out)

    def observations(self) -> torch.Tensor:
        """f32[N, 6] = [pos, vel, goal] per agent (go_to:124-132, oa:154-162)."""
        return torch.cat([torch.cat([self.pos[i], self.vel[i], self.goal], dim=-1) for i in range(self.n_agents)])

    def average_distance_to_goal(self) -> torch.Tensor:
        return torch.mean(torch.stack(self.distance_to_goal))       # go_to:134-135, oa:164-165

    def obstacles_hits(self) -> torch.Tensor:
        if self.scenario == GOTO:
            return torch.tensor(0.0)                                # go_to:140-141
        hits = torch.stack([get_distance(self.pos[i], self.obstacle) <= HIT_DISTANCE
                            for i in range(self.n_agents)])         # oa:170-173
        return torch.sum(hits)

    def state(self) -> Tuple[torch.Tensor, torch.Tensor]:
        return torch.cat(self.pos), torch.cat(self.vel)

    def set_state(self, pos: torch.Tensor, vel: torch.Tensor) -> None:
        for i in range(self.n_agents):
            self.pos[i] = pos[i].reshape(1, 2).clone().float()
            self.vel[i] = vel[i].reshape(1, 2).clone().float()


# --------------------------------------------------------------------------------------------
# graphs
# --------------------------------------------------------------------------------------------
def node_features(obs: torch.Tensor) -> torch.Tensor:
    """[obs | float(agent id)] -> f32[N, 7] (train:95-99, simulator:10-14)."""
    n = obs.shape[0]
    ids = torch.arange(n).float().unsqueeze(1)
    return torch.cat([obs, ids], dim=1)


def graph_complete(n: int) -> torch.Tensor:
    """train:101-108: (i,j),(j,i) for i<j, then one (0,0).  int64[2, n(n-1)+1], row 0 = source."""
    edge_index = []
    for i in range(n):
        for j in range(i + 1, n):
            edge_index.append([i, j])
            edge_index.append([j, i])
    edge_index.append([0, 0])
    return torch.tensor(edge_index, dtype=torch.long).t().contiguous()


def graph_knn(x: torch.Tensor, k: int) -> torch.Tensor:
    """simulator:15-24: per node i the k nearest (self included, torch.topk order) emit (i,a),(a,i);
    then one (0,0).  int64[2, 2kn+1]."""
    n = x.shape[0]
    edge_index = []
    for i in range(n):
        distance_to_i = torch.linalg.norm(x[:, :2] - x[i, :2], dim=1)
        _, nearest = torch.topk(distance_to_i, k, largest=False)
        for a in nearest:
            edge_index.append([i, a.item()])
            edge_index.append([a.item(), i])
    edge_index.append([0, 0])
    return torch.tensor(edge_index, dtype=torch.long).t().contiguous()


# --------------------------------------------------------------------------------------------
# GAT Q-network forward (torch_geometric 2.5.3 GATConv semantics)
# --------------------------------------------------------------------------------------------
def _scatter_sum(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    # torch_geometric.utils.scatter(reduce='sum'): broadcast the index, scatter_add_ into zeros
    size = [dim_size] + list(src.shape[1:])
    idx = index.view([-1] + [1] * (src.dim() - 1)).expand_as(src)
    return src.new_zeros(size).scatter_add_(0, idx, src)


def _scatter_max(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    size = [dim_size] + list(src.shape[1:])
    idx = index.view([-1] + [1] * (src.dim() - 1)).expand_as(src)
    return src.new_zeros(size).scatter_reduce_(0, idx, src, reduce="amax", include_self=False)


def segment_softmax(src: torch.Tensor, index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """torch_geometric.utils.softmax(src, index, num_nodes=N)."""
    src_max = _scatter_max(src.detach(), index, num_nodes)
    out = src - src_max.index_select(0, index)
    out = out.exp()
    out_sum = _scatter_sum(out, index, num_nodes) + 1e-16
    out_sum = out_sum.index_select(0, index)
    return out / out_sum


def gat_conv(x: torch.Tensor, edge_index: torch.Tensor, weight: torch.Tensor, att_src: torch.Tensor,
             att_dst: torch.Tensor, bias: torch.Tensor, negative_slope: float = 0.2) -> torch.Tensor:
    """GATConv(heads=1, concat=True, add_self_loops=False, bias=True).forward; flow source->target."""
    H, C = 1, weight.shape[0]
    h = F.linear(x, weight).view(-1, H, C)
    alpha_src = (h * att_src.view(1, H, C)).sum(dim=-1)
    alpha_dst = (h * att_dst.view(1, H, C)).sum(dim=-1)
    src, dst = edge_index[0], edge_index[1]
    alpha = alpha_src.index_select(0, src) + alpha_dst.index_select(0, dst)
    alpha = F.leaky_relu(alpha, negative_slope)
    alpha = segment_softmax(alpha, dst, x.shape[0])
    msg = alpha.unsqueeze(-1) * h.index_select(0, src)
    out = _scatter_sum(msg, dst, x.shape[0])
    out = out.view(-1, H * C)
    return out + bias


def gatq_forward(params: Dict[str, torch.Tensor], x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
    """GCN.forward (train:59-70) with a reference state dict."""
    h = gat_conv(x, edge_index, params["conv1.lin.weight"], params["conv1.att_src"], params["conv1.att_dst"],
                 params["conv1.bias"])
    h = torch.tanh(h)
    h = F.linear(h, params["lin1.weight"], params["lin1.bias"])
    h = torch.relu(h)
    return F.linear(h, params["lin2.weight"], params["lin2.bias"])


def batch_graphs(xs: List[torch.Tensor], eis: List[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """torch_geometric Batch.from_data_list: concatenate x; offset each edge_index by the node count."""
    off = 0
    out = []
    for x, ei in zip(xs, eis):
        out.append(ei + off)
        off += x.shape[0]
    return torch.cat(xs, dim=0), torch.cat(out, dim=1)


# --------------------------------------------------------------------------------------------
# greedy evaluation loop (simulator.py:47-109)
# --------------------------------------------------------------------------------------------
def run_evaluation(world: OracleWorld, params: Dict[str, torch.Tensor], episodes: int, max_steps: int,
                   graph_mode: str = "knn", k: int = 5, record_q: bool = False) -> Dict[str, list]:
    """Restates Simulator.run_simulation and returns what save_metrics_to_csv would write."""
    out = {"pos_x": [], "pos_y": [], "distance": [], "hits": [], "reward": [], "collisions": [],
           "distance_end": [], "distance_beginning": [], "actions": [], "q": []}
    n = world.n_agents
    for _ in range(episodes):
        obs = world.reset()
        total_reward = 0
        collisions = 0
        ep = {key: [] for key in ("pos_x", "pos_y", "distance", "hits", "actions", "q")}
        for i in range(max_steps):
            x = node_features(obs)
            ei = graph_knn(x, k) if graph_mode == "knn" else graph_complete(n)
            with torch.no_grad():
                q = gatq_forward(params, x, ei)
                actions = torch.argmax(q, dim=1)
            rewards = world.step(actions)
            obs = world.observations()
            if i == 0:
                out["distance_beginning"].append(world.average_distance_to_goal().item())
            total_reward += sum(rewards[j:j + 1] for j in range(n))
            collisions += world.obstacles_hits()
            ep["pos_x"].append([obs[j, 0].item() for j in range(n)])
            ep["pos_y"].append([obs[j, 1].item() for j in range(n)])
            ep["distance"].append(world.average_distance_to_goal().item())
            ep["hits"].append(world.obstacles_hits().item())
            ep["actions"].append(actions.tolist())
            if record_q:
                ep["q"].append(q.clone())
        for key in ep:
            out[key].append(ep[key])
        out["reward"].append((total_reward / max_steps).item())
        out["collisions"].append(collisions.item())
        out["distance_end"].append(world.average_distance_to_goal().item())
    return out
