"""DQN training oracle (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates src/training/train_gcn_dqn.py:25-204 (GraphReplayBuffer, GCN, DQNTrainer) on the single-env
oracle world, including the RNG consumption of the constructors so that the shipped
``data/stats/*.csv`` rows are reproducible from the seeds alone.
"""
from __future__ import annotations

import math
import random
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import swarm_oracle as so


def _glorot(t: torch.Tensor) -> None:
    # torch_geometric.nn.inits.glorot
    stdv = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    t.data.uniform_(-stdv, stdv)


class _PygLinear(nn.Module):
    """torch_geometric.nn.dense.linear.Linear(bias=False, weight_initializer='glorot')."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        self.reset_parameters()

    def reset_parameters(self) -> None:
        _glorot(self.weight)


class OracleGATConv(nn.Module):
    """GATConv(in, out, heads=1, add_self_loops=False, bias=True) of torch_geometric 2.5.3.  The
    projection is initialised twice (Linear.__init__, then GATConv.reset_parameters)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.lin = _PygLinear(in_channels, out_channels)
        self.att_src = nn.Parameter(torch.empty(1, 1, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, 1, out_channels))
        self.bias = nn.Parameter(torch.empty(out_channels))
        self.reset_parameters()

    def reset_parameters(self) -> None:
        self.lin.reset_parameters()
        _glorot(self.att_src)
        _glorot(self.att_dst)
        self.bias.data.fill_(0.0)

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        return so.gat_conv(x, edge_index, self.lin.weight, self.att_src, self.att_dst, self.bias)


class OracleGCN(nn.Module):
    """train:50-70."""

    def __init__(self, input_dim: int = 7, hidden_dim: int = 32, output_dim: int = 9):
        super().__init__()
        self.conv1 = OracleGATConv(input_dim, hidden_dim)
        self.lin1 = nn.Linear(hidden_dim, hidden_dim)
        self.lin2 = nn.Linear(hidden_dim, output_dim)

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        h = torch.tanh(self.conv1(x, edge_index))
        h = torch.relu(self.lin1(h))
        return self.lin2(h)


class OracleTrainer:
    """DQNTrainer (train:72-204) for N agents on the single-env oracle."""

    def __init__(self, experiment: str, seed: int, n_agents: int = 5, max_steps: int = 100):
        scenario = {"GoTo": so.GOTO, "ObstacleAvoidance": so.OBSTACLE_AVOIDANCE}[experiment]
        # make_env: GoTo.make_world reseeds torch with 1 (go_to:13,23); Environment.__init__ then
        # reseeds torch / numpy / random with ``seed`` and performs one reset.
        if scenario == so.GOTO:
            torch.manual_seed(1)
        torch.manual_seed(seed)
        np.random.seed(seed)
        random.seed(seed)
        self.world = so.OracleWorld(scenario, n_agents, random=False, max_steps=max_steps)
        self.world.reset()
        self.n = n_agents
        self.max_steps = max_steps
        self.model = OracleGCN(7, 32, 9)
        self.target_model = OracleGCN(7, 32, 9)
        self.target_model.load_state_dict(self.model.state_dict())
        self.optimizer = torch.optim.Adam(self.model.parameters(), 0.001)
        self.buffer: List[tuple] = []
        self.capacity = 1000000
        self.position = 0
        self.edge_index = so.graph_complete(n_agents)
        self.episode_losses: List[float] = []
        self.episode_returns: List[float] = []      # total_episode_reward[0] per episode
        self.last_grads: Optional[Dict[str, torch.Tensor]] = None

    def push(self, x, actions, rewards, x_next) -> None:
        if len(self.buffer) < self.capacity:
            self.buffer.append(None)
        self.buffer[self.position] = (x, actions, rewards, x_next)
        self.position = (self.position + 1) % self.capacity

    def train_step(self, batch_size: int, ticks: int, gamma: float = 0.99, update_target_every: int = 200) -> float:
        if len(self.buffer) < batch_size:
            return 0
        self.optimizer.zero_grad()
        sample = random.sample(self.buffer, batch_size)
        eis = [self.edge_index] * batch_size
        obs_x, obs_ei = so.batch_graphs([s[0] for s in sample], eis)
        actions = torch.cat([s[1] for s in sample])
        rewards = torch.cat([s[2] for s in sample])
        nxt_x, nxt_ei = so.batch_graphs([s[3] for s in sample], eis)
        values = self.model(obs_x, obs_ei).gather(1, actions.unsqueeze(1))
        next_values = self.target_model(nxt_x, nxt_ei).max(dim=1)[0].detach()
        target_values = rewards + (gamma * next_values)
        loss = nn.MSELoss()(values, target_values.unsqueeze(1))
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), 1)
        self.optimizer.step()
        if ticks % update_target_every == 0:
            self.target_model.load_state_dict(self.model.state_dict())
        return loss.item()

    def train(self, episodes: int, epsilon0: float = 0.99, decay: float = 0.01, min_eps: float = 0.05) -> None:
        ticks = 0
        epsilon = epsilon0
        for episode in range(episodes):
            obs = self.world.reset()
            episode_loss = 0
            total = torch.zeros(self.n)
            for _ in range(self.max_steps):
                ticks += 1
                x = so.node_features(obs)
                logits = self.model(x, self.edge_index).detach()
                if random.random() < epsilon:
                    actions = torch.tensor([random.randint(0, 8) for _ in range(self.n)])
                else:
                    actions = torch.argmax(logits, dim=1)
                rewards = self.world.step(actions)
                new_obs = self.world.observations()
                rewards_tensor = torch.tensor([rewards[i] for i in range(self.n)], dtype=torch.float)
                self.push(x, actions, rewards_tensor, so.node_features(new_obs))
                loss = self.train_step(32, ticks, update_target_every=200)
                episode_loss += loss
                total += rewards_tensor / self.n
                obs = new_obs
            epsilon = max(min_eps, epsilon0 * np.exp(-decay * episode))
            self.episode_losses.append(episode_loss / self.max_steps)
            self.episode_returns.append(total[0].item())
