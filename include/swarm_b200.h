/*
 * swarm_b200.h -- C ABI of the B200-native swarm hot path.
 *
 * The reference (davidedomini/experiments-2025-acsos-marl-for-swarming-behaviors) has no FFI boundary:
 * it is in-process Python on top of vmas==1.4.0 and torch_geometric==2.5.3.  These entry points are
 * what a binding for its hot path would call; each one names the reference code it replaces
 * (paths relative to the reference repository root).  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in `_host`;
 *     the library never allocates, frees or keeps caller memory, and has no global mutable state
 *     apart from the thread-local error string;
 *   - `stream` is a cudaStream_t passed as void*; all calls are asynchronous on it and never
 *     synchronise the host;
 *   - return value 0 = OK; negative = error (SWARM_ERR_*), message via swarm_last_error();
 *   - layouts: state   float[B*N][4] = (x, y, vx, vy) of agent i of env b at index b*N+i
 *              actions int32[B*N], rewards float[B*N]
 *              edges   int32[B][2][E] env-local node ids, row 0 = source, row 1 = target
 *              weights float[1673], packed in SWARM_W_* order (all row-major as in the state dict)
 */
#ifndef SWARM_B200_H
#define SWARM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWARM_ABI_VERSION 1

enum { SWARM_SCENARIO_GOTO = 0, SWARM_SCENARIO_OBSTACLE_AVOIDANCE = 1 };
enum { SWARM_GRAPH_COMPLETE = 0, SWARM_GRAPH_KNN = 1 };
enum {
  SWARM_OK = 0,
  SWARM_ERR_INVALID_ARG = -1,
  SWARM_ERR_UNSUPPORTED = -2,
  SWARM_ERR_CUDA = -3
};

/* packed Q-network weights: offsets (in floats) of the state-dict tensors of
 * src/training/train_gcn_dqn.py:50-57 (GATConv(7,32) + Linear(32,32) + Linear(32,9)) */
enum {
  SWARM_W_CONV_LIN = 0,       /* conv1.lin.weight [32][7]  */
  SWARM_W_ATT_SRC = 224,      /* conv1.att_src    [32]     */
  SWARM_W_ATT_DST = 256,      /* conv1.att_dst    [32]     */
  SWARM_W_CONV_BIAS = 288,    /* conv1.bias       [32]     */
  SWARM_W_LIN1 = 320,         /* lin1.weight      [32][32] */
  SWARM_W_LIN1_BIAS = 1344,   /* lin1.bias        [32]     */
  SWARM_W_LIN2 = 1376,        /* lin2.weight      [9][32]  */
  SWARM_W_LIN2_BIAS = 1664,   /* lin2.bias        [9]      */
  SWARM_W_COUNT = 1673
};
enum { SWARM_FEAT = 7, SWARM_HIDDEN = 32, SWARM_ACTIONS = 9, SWARM_OBS = 6 };

/* per-agent flag bits written by swarm_sim_step / swarm_rollout traces */
enum {
  SWARM_FLAG_OBSTACLE_CONTACT = 1, /* centre distance to the obstacle <= r_a + r_o (vmas collides()) */
  SWARM_FLAG_HIT = 2,              /* get_distance <= hit_distance (obstacle_avoidance_scenario.py:170-173) */
  SWARM_FLAG_PENALTY = 4           /* get_distance <= penalty_distance (obstacle_avoidance_scenario.py:149) */
};

/* World + scenario constants.  swarm_default_config() fills the values the reference runs with:
 * vmas World defaults (dt .1, drag .25, collision_force 100, contact_margin 1e-3, sphere radius .05)
 * and the scenario constants of src/scenarios/{go_to_position,obstacle_avoidance}_scenario.py. */
typedef struct SwarmConfig {
  double grid_spacing;     /* desired_distance (go_to:17 / oa:23); a Python double in generate_grid */
  int32_t num_envs;        /* B */
  int32_t n_agents;        /* N */
  int32_t scenario;        /* SWARM_SCENARIO_* */
  int32_t graph_mode;      /* SWARM_GRAPH_* */
  int32_t knn_k;           /* k of simulator.py:19 (10 shipped, 5 in the goldens) */
  float dt;
  float drag;
  float collision_force;
  float contact_margin;
  float agent_radius;
  float landmark_radius;
  float goal_x, goal_y;            /* go_to:86, oa:97 */
  float obstacle_x, obstacle_y;    /* oa:99 */
  float hit_distance;              /* oa:25 */
  float penalty_distance;          /* oa:24 */
  float obstacle_weight;           /* oa:136 */
} SwarmConfig;

/* optional per-tick traces of swarm_rollout (any pointer may be NULL); T = ticks */
typedef struct SwarmTrace {
  float* state;          /* [T][B*N][4] state AFTER each tick                              */
  int32_t* actions;      /* [T][B*N]    greedy / injected action used at each tick          */
  float* q;              /* [T][B*N][9] Q-values computed at each tick (pre-step state)     */
  float* rewards;        /* [T][B*N]                                                        */
  uint8_t* flags;        /* [T][B*N]    SWARM_FLAG_* after each tick                        */
  uint32_t* contact;     /* [T][B*N]    bit j = agent-agent contact with agent j (N <= 32)  */
  int32_t* edges;        /* [T][B][2][E] graph built at each tick                           */
  float* dist;           /* [T][B*N][2] (distance_to_goal, get_distance to the obstacle) after each tick */
} SwarmTrace;

int swarm_abi_version(void);
const char* swarm_last_error(void);
void swarm_default_config(SwarmConfig* cfg, int32_t scenario, int32_t num_envs, int32_t n_agents);
/* number of edges per env for cfg->graph_mode: N(N-1)+1 (train_gcn_dqn.py:101-108) or 2kN+1 (simulator.py:15-24) */
int64_t swarm_edges_per_env(const SwarmConfig* cfg);

/* generate_grid + reset_world_at (go_to:52-106, oa:63-133): agents on the row-major grid around
 * centers[b] (float[B][2]; the caller draws them -- the reference draws one torch.normal per reset),
 * velocities zero. */
int swarm_reset_grid(const SwarmConfig* cfg, const float* centers, float* state, void* stream);

/* One vmas Environment.step (call sites train_gcn_dqn.py:169, simulator.py:68): action decode,
 * contact forces (obstacle first, then agent pairs in entity order), drag + semi-implicit Euler, then
 * the scenario's reward()/observation() and the obstacle flags.  Optional outputs may be NULL:
 * rewards float[B*N], flags uint8[B*N], contact uint32[B*N] (N <= 32 only), obs float[B*N][6],
 * dist float[B*N][2] = (agent.distance_to_goal, world.get_distance(agent, obstacle)) -- the per-agent terms
 * behind average_distance_to_goal() / average_distance_to_obstacles() (oa:164-168).  state_out may alias
 * state_in. */
int swarm_sim_step(const SwarmConfig* cfg, const float* state_in, const int32_t* actions, float* state_out,
                   float* rewards, uint8_t* flags, uint32_t* contact, float* obs, float* dist, void* stream);

/* Graph builders: train_gcn_dqn.py:94-110 (complete) and simulator.py:9-26 (symmetrised kNN with
 * torch.topk tie order).  edges int32[B][2][E]; neighbours (optional, kNN only) int32[B*N][k] holds the
 * topk index row of every agent. */
int swarm_graph_build(const SwarmConfig* cfg, const float* state, int32_t* edges, int32_t* neighbours,
                      void* stream);

/* GCN.forward (train_gcn_dqn.py:59-70) on the per-env graph named by cfg->graph_mode, node features
 * [pos, vel, goal, agent id] built from state (train:95-99).  q float[B*N][9] and actions int32[B*N]
 * (argmax, first maximum wins; train:167, simulator:64) are optional. */
int swarm_gatq_forward(const SwarmConfig* cfg, const float* weights, const float* state, float* q,
                       int32_t* actions, void* stream);

/* GCN.forward on an arbitrary graph: x float[n][7]; edges grouped by target in CSR form, each group in
 * edge-list order (row_ptr int32[n+1], src int32[E]).  Used by the nn.Module seam (train_gcn_dqn.py:59-70)
 * for arbitrary Data/Batch inputs.  workspace: swarm_gatq_workspace_bytes(n) bytes of device memory. */
int64_t swarm_gatq_workspace_bytes(int32_t n_nodes);
int swarm_gatq_forward_csr(int32_t n_nodes, const float* weights, const float* x, const int32_t* row_ptr,
                           const int32_t* src, float* q, void* workspace, int64_t workspace_bytes, void* stream);

/* Stable grouping of an edge list by target (the order scatter-add sees): edge_src/edge_dst int64[E]
 * (torch_geometric edge_index rows) -> row_ptr int32[n+1], src int32[E], perm int32[E] (perm[p] = position
 * in the input list of grouped edge p).  workspace: swarm_csr_workspace_bytes(n, E) bytes. */
int64_t swarm_csr_workspace_bytes(int32_t n_nodes, int64_t n_edges);
int swarm_csr_from_edges(int32_t n_nodes, int64_t n_edges, const int64_t* edge_src, const int64_t* edge_dst,
                         int32_t* row_ptr, int32_t* src, int32_t* perm, void* workspace, int64_t workspace_bytes,
                         void* stream);

/* Fused greedy rollout: `ticks` iterations of [graph build -> GCN forward -> argmax -> env step]
 * (the inner loops of simulator.py:59-93 and train_gcn_dqn.py:153-178 without exploration) with the
 * state resident on chip.  state is updated in place; returns float[B*N] accumulates each agent's
 * rewards (+=), hits int32[B] accumulates obstacles_hits() per tick (+=); both optional.
 * forced_actions (optional) int32[ticks][B*N]: entries >= 0 override the greedy action. */
int swarm_rollout(const SwarmConfig* cfg, const float* weights, float* state, int32_t ticks,
                  const int32_t* forced_actions, float* returns, int32_t* hits, const SwarmTrace* trace,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SWARM_B200_H */
