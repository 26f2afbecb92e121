"""B200-native swarm hot path: batched VMAS-style world step (GoTo / ObstacleAvoidance), per-step agent
graph, GAT Q-network and DQN update, as hand-written sm_100a CUDA kernels behind the reference's own
Python seams (``BaseScenario`` / ``make_env`` / ``GCN``).

The directory name carries the reference repository's name and is not a valid Python identifier; import it
through the ``swarm_b200`` alias module at the repository root (``import swarm_b200``).
"""
from . import _build, _lib, ops                                   # noqa: F401
from ._lib import SwarmConfig, SwarmError, pack_weights, unpack_weights   # noqa: F401
from .dqn import DQNTrainer, GraphReplayBuffer, set_seed          # noqa: F401
from .env import Environment, make_env                            # noqa: F401
from .gcn import GCN, GATConv, StackedGCN                                     # noqa: F401
from .graph import Batch, Data, create_graph_from_observations    # noqa: F401
from .scenarios import (Agent, BaseScenario, CohesionScenario, Color, FlockingScenario,   # noqa: F401
                        GoToPositionScenario, Landmark, ObstacleAvoidanceScenario, Sphere, World)
from .simulator import Simulator                                   # noqa: F401

__all__ = ["make_env", "Environment", "GCN", "GATConv", "StackedGCN", "Data", "Batch", "create_graph_from_observations",
           "BaseScenario", "GoToPositionScenario", "ObstacleAvoidanceScenario", "FlockingScenario", "CohesionScenario", "World", "Agent", "Landmark", "Sphere", "Color", "SwarmConfig", "SwarmError",
           "pack_weights", "unpack_weights", "ops", "DQNTrainer", "GraphReplayBuffer", "set_seed", "Simulator"]
