// api.cu -- extern "C" entry points of libswarm_b200.so (see include/swarm_b200.h).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "tile_kernels.cuh"

namespace swarm {
cudaError_t launch_tile(int mode, const TileParams& p, cudaStream_t stream);
cudaError_t launch_sim_step(const TileParams& p, cudaStream_t stream);
cudaError_t launch_sim_step_large(const TileParams& p, cudaStream_t stream);
cudaError_t launch_graph_large(const SwarmConfig& c, const float* state, int32_t* edges, int32_t* nbr, int edges_per_env,
                               cudaStream_t stream);
cudaError_t launch_reset_grid(const SwarmConfig& c, int cols, int rows, const float* centers, float* state,
                              cudaStream_t stream);
long long gat_layer_workspace_bytes(int n, long long E, int ci, int co, bool backward);
cudaError_t launch_gat_layer_forward(int n, int ci, int co, const float* W, const float* a_src, const float* a_dst,
                                     const float* bias, const float* x, const int32_t* row_ptr, const int32_t* src,
                                     float* out, void* workspace, cudaStream_t stream);
cudaError_t launch_gat_layer_backward(int n, long long E, int ci, int co, const float* W, const float* a_src,
                                      const float* a_dst, const float* x, const int32_t* row_ptr, const int32_t* src,
                                      const int32_t* perm, const int32_t* row_ptr_s, const int32_t* tgt_s,
                                      const int32_t* perm_s, const float* grad_out, float* gW, float* gas, float* gad,
                                      float* gb, float* grad_x, void* workspace, cudaStream_t stream);
cudaError_t launch_scenario_reward(const SwarmRewardSpec& sp, const float* state, float* shaping, float* reward,
                                   float* terms, cudaStream_t stream);
cudaError_t launch_reset_random(const SwarmConfig& c, const SwarmResetSpec& sp, int cols, int rows, const SwarmTrainCtl* ctl,
                                long long episode, float* centers_out, float* state, cudaStream_t stream);
cudaError_t launch_episode_end(const SwarmConfig& c, SwarmTrainCtl* ctl, float* returns, int32_t* hits, const float* loss,
                               float* stats, long long max_episodes, double eps0, double decay, double min_eps,
                               cudaStream_t stream);
int stack_weight_count_host(const SwarmStackSpec& s);
cudaError_t launch_gatstack_forward(const SwarmConfig& c, const SwarmStackSpec& spec, const float* weights, const float* state,
                                    float* q, int32_t* actions, cudaStream_t stream);
cudaError_t launch_gatstack_rollout(const SwarmConfig& c, const SwarmStackSpec& spec, const float* weights, float* state,
                                    int ticks, const TileParams& step, const SwarmRewardSpec* flock, float* shaping,
                                    float* returns, int32_t* hits, cudaStream_t stream);
cudaError_t launch_stack_accumulate(long long total, int N, const float* rewards, int per_env, const uint8_t* flags,
                                    float* returns, int32_t* hits, cudaStream_t stream);
bool gatq_knn_large_fits(int N, int K);
bool gatq_large_x_fits(int N, bool radius);
cudaError_t launch_gatq_large_x(const SwarmConfig& c, const float* weights, const float* state, float* q, int32_t* actions,
                                cudaStream_t stream);
cudaError_t launch_radius_csr(const SwarmConfig& c, const float* state, int32_t* degree, const int32_t* row_ptr,
                              int32_t* src, cudaStream_t stream);
bool gatq_knn_large_x_fits(int N, int K);
cudaError_t launch_gatq_knn_large_x(const SwarmConfig& c, const float* weights, const float* state, const int32_t* nbr,
                                    float* q, int32_t* actions, cudaStream_t stream);
cudaError_t launch_gatq_knn_large(const SwarmConfig& c, const float* weights, const float* state, const int32_t* nbr,
                                  float* q, int32_t* actions, cudaStream_t stream);
long long csr_workspace_bytes(int n, long long E);
cudaError_t launch_csr_from_edges(int n, long long E, const int64_t* edge_src, const int64_t* edge_dst, int32_t* row_ptr,
                                  int32_t* src, int32_t* perm, void* workspace, long long workspace_bytes,
                                  cudaStream_t stream);
cudaError_t launch_gatq_csr(int n, const float* weights, const float* x, const int32_t* row_ptr, const int32_t* src,
                            float* q, int32_t* actions, float* rows, cudaStream_t stream, bool conv_only = false);
long long gatq_workspace_bytes(int n);
long long gatq_backward_workspace_bytes(int n, long long E);
cudaError_t launch_gatq_backward_csr(int n, long long E, const float* weights, const float* x, const int32_t* row_ptr,
                                     const int32_t* src, const int32_t* perm, const int32_t* row_ptr_s,
                                     const int32_t* tgt_s, const int32_t* perm_s, const float* grad_q, float* grad_w,
                                     void* workspace, cudaStream_t stream, bool conv_only = false);
cudaError_t launch_replay_push(const SwarmReplay& r, long long cursor, int B, int N, const float* state,
                               const int32_t* actions, const float* rewards, const float* next_state,
                               cudaStream_t stream);
cudaError_t launch_replay_gather(const SwarmReplay& r, const int64_t* indices, int G, int N, float* state,
                                 int32_t* actions, float* rewards, float* next_state, cudaStream_t stream);
long long dqn_workspace_bytes(const SwarmConfig& c, int n_graphs);
int dqn_smem_bytes(const SwarmConfig& c);
bool dqn_fuse_reduce(const SwarmConfig& c, int n_graphs);
cudaError_t launch_dqn_grad(const SwarmConfig& c, const float* w_online, const float* w_target, const SwarmReplay& batch,
                            const int64_t* indices, int n_graphs, float gamma, float loss_scale, float* grad, float* loss,
                            float* td, void* workspace, cudaStream_t stream, SwarmTrainCtl* ctl = nullptr,
                            int64_t* indices_out = nullptr, unsigned long long sample_seed = 0, int pushed_envs = 0,
                            bool skip_reduce = false);
cudaError_t launch_adam_clip(float* w, const float* grad, float* m, float* v, long long step, double lr, double beta1,
                             double beta2, double eps, double max_norm, float* target, float* grad_norm,
                             cudaStream_t stream, SwarmTrainCtl* ctl = nullptr, int num_envs = 0,
                             long long ring_capacity = 1, int update_target_every = 1,
                             const SwarmPeerExchange* peers = nullptr, float* grad_rw = nullptr,
                             const SwarmConfig* reduce_cfg = nullptr, const void* reduce_workspace = nullptr,
                             int reduce_graphs = 0, float loss_scale = 0.0f);

namespace {
thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return SWARM_OK;
  return fail(SWARM_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

constexpr int kMaxLargeAgents = 4096;

int validate(const SwarmConfig* cfg, bool need_graph, bool allow_large = false) {
  if (!cfg) return fail(SWARM_ERR_INVALID_ARG, "cfg is NULL");
  if (cfg->num_envs <= 0) return fail(SWARM_ERR_INVALID_ARG, "num_envs must be positive");
  if (cfg->n_agents <= 0) return fail(SWARM_ERR_INVALID_ARG, "n_agents must be positive");
  if (cfg->scenario != SWARM_SCENARIO_GOTO && cfg->scenario != SWARM_SCENARIO_OBSTACLE_AVOIDANCE)
    return fail(SWARM_ERR_INVALID_ARG, "unknown scenario id");
  if (cfg->n_agents > kTileThreads && !allow_large)
    return fail(SWARM_ERR_UNSUPPORTED, "n_agents > 128 is not supported by the fused env-tile kernels (use swarm_sim_step / "
                                       "swarm_graph_build / swarm_gatq_forward_csr for large swarms)");
  if (cfg->n_agents > kMaxLargeAgents) return fail(SWARM_ERR_UNSUPPORTED, "n_agents > 4096");
  if (need_graph) {
    if (cfg->graph_mode != SWARM_GRAPH_COMPLETE && cfg->graph_mode != SWARM_GRAPH_KNN &&
        cfg->graph_mode != SWARM_GRAPH_RADIUS)
      return fail(SWARM_ERR_INVALID_ARG, "unknown graph mode");
    if (cfg->graph_mode == SWARM_GRAPH_RADIUS) {
      if (!(cfg->graph_radius >= 0.0f)) return fail(SWARM_ERR_INVALID_ARG, "graph_radius must be >= 0");
      if (cfg->n_agents > kTileThreads && !allow_large)
        return fail(SWARM_ERR_UNSUPPORTED, "the padded radius edge block is implemented for n_agents <= 128 (large swarms: "
                                           "swarm_graph_build_radius_csr / swarm_gatq_forward_large / swarm_rollout_large)");
    }
    if (cfg->graph_mode == SWARM_GRAPH_KNN) {
      // torch.topk raises "selected index k out of range" for k > n (simulator.py:19 with n_agents < 10)
      if (cfg->knn_k <= 0 || cfg->knn_k > cfg->n_agents)
        return fail(SWARM_ERR_INVALID_ARG, "selected index k out of range");
    }
  }
  return SWARM_OK;
}

int fill_params(TileParams& p, const SwarmConfig* cfg, int mode) {
  std::memset(&p, 0, sizeof(p));
  p.cfg = *cfg;
  const int n = cfg->n_agents;
  p.epb = n <= kTileThreads ? kTileThreads / n : 1;
  p.ticks = 1;
  const bool knn = cfg->graph_mode == SWARM_GRAPH_KNN;
  p.maxdeg = knn ? (n + cfg->knn_k + 1) : n;
  p.edges_per_env = (int32_t)swarm_edges_per_env(cfg);
  p.one_minus_drag = (float)(1.0 - (double)cfg->drag);
  p.dmin_aa = cfg->agent_radius + cfg->agent_radius;
  p.dmin_ao = cfg->landmark_radius + cfg->agent_radius;
  p.qmax_aa = sq_threshold(p.dmin_aa);
  p.qmax_ao = sq_threshold(p.dmin_ao);
  p.qmax_r = cfg->graph_mode == SWARM_GRAPH_RADIUS ? sq_threshold(cfg->graph_radius) : 0.0f;
  // dense contractions on the tensor cores (tcgen05 3xTF32) unless SWARM_TC=0 selects the CUDA-core FFMA path
  const char* tc_env = std::getenv("SWARM_TC");
  p.use_tc = (mode == MODE_ROLLOUT || mode == MODE_FORWARD) && !(tc_env && tc_env[0] == '0');
  const char* ord_env = std::getenv("SWARM_KNN_ORDERED");
  p.knn_ordered = (ord_env && ord_env[0] == '1') ? 1 : 0;
  TileLayout L = tile_layout(mode, kTileThreads, n, cfg->knn_k, p.maxdeg, cfg->graph_mode, p.use_tc != 0);
  if (L.total > 227 * 1024 && p.use_tc) {
    p.use_tc = 0;
    L = tile_layout(mode, kTileThreads, n, cfg->knn_k, p.maxdeg, cfg->graph_mode, false);
  }
  if (L.total > 227 * 1024) return fail(SWARM_ERR_UNSUPPORTED, "shared-memory tile exceeds 227 KB for this (N, k)");
  return SWARM_OK;
}
}  // namespace
}  // namespace swarm

using namespace swarm;

extern "C" {

int swarm_abi_version(void) { return SWARM_ABI_VERSION; }

const char* swarm_last_error(void) { return g_last_error.c_str(); }

void swarm_default_config(SwarmConfig* cfg, int32_t scenario, int32_t num_envs, int32_t n_agents) {
  if (!cfg) return;
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->grid_spacing = 0.15;
  cfg->num_envs = num_envs;
  cfg->n_agents = n_agents;
  cfg->scenario = scenario;
  cfg->graph_mode = SWARM_GRAPH_COMPLETE;
  cfg->knn_k = 10;
  cfg->dt = 0.1f;
  cfg->drag = 0.25f;
  cfg->collision_force = 100.0f;
  cfg->contact_margin = 1e-3f;
  cfg->agent_radius = 0.05f;
  cfg->landmark_radius = 0.05f;
  cfg->goal_x = -0.8f;
  cfg->goal_y = 0.8f;
  cfg->obstacle_x = -0.1f;
  cfg->obstacle_y = 0.1f;
  cfg->hit_distance = 0.2f;
  cfg->penalty_distance = 1.0f;
  cfg->obstacle_weight = 2.5f;
  cfg->graph_radius = 0.35f;       // extension default: <= 20 neighbours on the 0.15-spaced start grid
}

int64_t swarm_edges_per_env(const SwarmConfig* cfg) {
  if (!cfg) return 0;
  const int64_t n = cfg->n_agents;
  if (cfg->graph_mode == SWARM_GRAPH_KNN) return 2 * (int64_t)cfg->knn_k * n + 1;
  return n * (n - 1) + 1;
}

int swarm_reset_grid(const SwarmConfig* cfg, const float* centers, float* state, void* stream) {
  if (int rc = validate(cfg, false, true)) return rc;
  if (!centers || !state) return fail(SWARM_ERR_INVALID_ARG, "centers/state is NULL");
  const int n = cfg->n_agents;
  const int cols = (int)std::ceil(std::sqrt((double)n));     // math.ceil(math.sqrt(num_points))
  const int rows = (int)std::ceil((double)n / (double)cols);  // math.ceil(num_points / num_cols)
  return check_cuda(launch_reset_grid(*cfg, cols, rows, centers, state, (cudaStream_t)stream), "swarm_reset_grid");
}

int swarm_sim_step(const SwarmConfig* cfg, const float* state_in, const int32_t* actions, float* state_out,
                   float* rewards, uint8_t* flags, uint32_t* contact, float* obs, float* dist, void* stream) {
  if (int rc = validate(cfg, false, true)) return rc;
  if (!state_in || !actions || !state_out) return fail(SWARM_ERR_INVALID_ARG, "state_in/actions/state_out is NULL");
  if (contact && cfg->n_agents > 32) return fail(SWARM_ERR_UNSUPPORTED, "contact masks need n_agents <= 32");
  TileParams p;
  if (int rc = fill_params(p, cfg, MODE_STEP)) return rc;
  p.state_in = state_in;
  p.state_out = state_out;
  p.actions_in = actions;
  p.rewards_out = rewards;
  p.flags_out = flags;
  p.contact_out = contact;
  p.obs_out = obs;
  p.dist_out = dist;
  // n_agents <= 64: lean streaming kernel (step_kernels.cu); <= 128: env-tile kernel; above: one env per CTA row
  // (large_kernels.cu) -- all three run the same device arithmetic
  if (cfg->n_agents <= 64) return check_cuda(launch_sim_step(p, (cudaStream_t)stream), "swarm_sim_step");
  if (cfg->n_agents > kTileThreads) return check_cuda(launch_sim_step_large(p, (cudaStream_t)stream), "swarm_sim_step");
  return check_cuda(launch_tile(MODE_STEP, p, (cudaStream_t)stream), "swarm_sim_step");
}

int swarm_graph_build(const SwarmConfig* cfg, const float* state, int32_t* edges, int32_t* neighbours, void* stream) {
  if (int rc = validate(cfg, true, true)) return rc;
  if (!state) return fail(SWARM_ERR_INVALID_ARG, "state is NULL");
  if (!edges && !neighbours) return fail(SWARM_ERR_INVALID_ARG, "no output requested");
  if (cfg->graph_mode == SWARM_GRAPH_RADIUS)
    return fail(SWARM_ERR_INVALID_ARG, "radius graphs have a per-env edge count: use swarm_graph_build_radius");
  if (cfg->n_agents > kTileThreads) {
    if (cfg->graph_mode == SWARM_GRAPH_KNN && (int64_t)cfg->knn_k * 64 > cfg->n_agents)
      return fail(SWARM_ERR_UNSUPPORTED,
                  "kNN for n_agents > 128 implements torch.topk's partial_sort branch only (needs 64 * k <= n_agents)");
    if (cfg->graph_mode == SWARM_GRAPH_KNN && cfg->knn_k > 64) return fail(SWARM_ERR_UNSUPPORTED, "knn_k > 64");
    return check_cuda(launch_graph_large(*cfg, state, edges, neighbours, (int)swarm_edges_per_env(cfg), (cudaStream_t)stream),
                      "swarm_graph_build");
  }
  TileParams p;
  if (int rc = fill_params(p, cfg, MODE_GRAPH)) return rc;
  p.state_in = state;
  p.edges_out = edges;
  p.nbr_out = neighbours;
  return check_cuda(launch_tile(MODE_GRAPH, p, (cudaStream_t)stream), "swarm_graph_build");
}

int swarm_graph_build_radius(const SwarmConfig* cfg, const float* state, int32_t* edges, int32_t* counts, void* stream) {
  if (int rc = validate(cfg, true)) return rc;
  if (cfg->graph_mode != SWARM_GRAPH_RADIUS) return fail(SWARM_ERR_INVALID_ARG, "cfg->graph_mode must be SWARM_GRAPH_RADIUS");
  if (!state || !edges) return fail(SWARM_ERR_INVALID_ARG, "state/edges is NULL");
  TileParams p;
  if (int rc = fill_params(p, cfg, MODE_GRAPH)) return rc;
  p.state_in = state;
  p.edges_out = edges;
  p.counts_out = counts;
  return check_cuda(launch_tile(MODE_GRAPH, p, (cudaStream_t)stream), "swarm_graph_build_radius");
}

int swarm_gatq_forward(const SwarmConfig* cfg, const float* weights, const float* state, float* q, int32_t* actions,
                       void* stream) {
  if (int rc = validate(cfg, true)) return rc;
  if (!weights || !state) return fail(SWARM_ERR_INVALID_ARG, "weights/state is NULL");
  if (!q && !actions) return fail(SWARM_ERR_INVALID_ARG, "no output requested");
  TileParams p;
  if (int rc = fill_params(p, cfg, MODE_FORWARD)) return rc;
  p.weights = weights;
  p.state_in = state;
  p.q_out = q;
  p.act_out = actions;
  return check_cuda(launch_tile(MODE_FORWARD, p, (cudaStream_t)stream), "swarm_gatq_forward");
}

int64_t swarm_gatq_workspace_bytes(int32_t n_nodes) { return n_nodes > 0 ? gatq_workspace_bytes(n_nodes) : 0; }

int swarm_gatq_forward_csr(int32_t n_nodes, const float* weights, const float* x, const int32_t* row_ptr,
                           const int32_t* src, float* q, int32_t* actions, void* workspace, int64_t workspace_bytes,
                           void* stream) {
  if (n_nodes < 0) return fail(SWARM_ERR_INVALID_ARG, "n_nodes must be >= 0");
  if (n_nodes == 0) return SWARM_OK;
  if (!weights || !x || !row_ptr || !workspace) return fail(SWARM_ERR_INVALID_ARG, "NULL argument");
  if (!q && !actions) return fail(SWARM_ERR_INVALID_ARG, "no output requested");
  if (workspace_bytes < gatq_workspace_bytes(n_nodes)) return fail(SWARM_ERR_INVALID_ARG, "workspace too small");
  float* rows = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  return check_cuda(launch_gatq_csr(n_nodes, weights, x, row_ptr, src, q, actions, rows, (cudaStream_t)stream),
                    "swarm_gatq_forward_csr");
}

int swarm_gatq_forward_knn_large(const SwarmConfig* cfg, const float* weights, const float* state,
                                 const int32_t* neighbours, float* q, int32_t* actions, void* stream) {
  if (int rc = validate(cfg, true, true)) return rc;
  if (cfg->graph_mode != SWARM_GRAPH_KNN) return fail(SWARM_ERR_INVALID_ARG, "cfg->graph_mode must be SWARM_GRAPH_KNN");
  if (!weights || !state || !neighbours) return fail(SWARM_ERR_INVALID_ARG, "weights/state/neighbours is NULL");
  if (!q && !actions) return fail(SWARM_ERR_INVALID_ARG, "no output requested");
  if (!gatq_knn_large_fits(cfg->n_agents, cfg->knn_k))
    return fail(SWARM_ERR_UNSUPPORTED, "the env does not fit shared memory: use swarm_graph_build + swarm_csr_from_edges + "
                                       "swarm_gatq_forward_csr");
  return check_cuda(launch_gatq_knn_large(*cfg, weights, state, neighbours, q, actions, (cudaStream_t)stream),
                    "swarm_gatq_forward_knn_large");
}

int swarm_graph_build_radius_csr(const SwarmConfig* cfg, const float* state, int32_t* degree, const int32_t* row_ptr,
                                 int32_t* src, void* stream) {
  if (int rc = validate(cfg, true, true)) return rc;
  if (cfg->graph_mode != SWARM_GRAPH_RADIUS) return fail(SWARM_ERR_INVALID_ARG, "cfg->graph_mode must be SWARM_GRAPH_RADIUS");
  if (!state) return fail(SWARM_ERR_INVALID_ARG, "state is NULL");
  if ((src == nullptr) == (degree == nullptr))
    return fail(SWARM_ERR_INVALID_ARG, "pass degree (count pass) or row_ptr + src (fill pass), not both");
  if (src && !row_ptr) return fail(SWARM_ERR_INVALID_ARG, "the fill pass needs row_ptr");
  if ((int64_t)cfg->num_envs * cfg->n_agents > 0x7fffffffLL) return fail(SWARM_ERR_UNSUPPORTED, "more than 2^31 - 1 nodes");
  return check_cuda(launch_radius_csr(*cfg, state, degree, row_ptr, src, (cudaStream_t)stream), "swarm_graph_build_radius_csr");
}

int swarm_gatq_forward_large(const SwarmConfig* cfg, const float* weights, const float* state, float* q, int32_t* actions,
                             void* stream) {
  if (int rc = validate(cfg, true, true)) return rc;
  if (cfg->graph_mode != SWARM_GRAPH_RADIUS && cfg->graph_mode != SWARM_GRAPH_COMPLETE)
    return fail(SWARM_ERR_INVALID_ARG, "cfg->graph_mode must be SWARM_GRAPH_RADIUS or SWARM_GRAPH_COMPLETE "
                                       "(kNN: swarm_gatq_forward_knn_large)");
  if (!weights || !state) return fail(SWARM_ERR_INVALID_ARG, "weights/state is NULL");
  if (!q && !actions) return fail(SWARM_ERR_INVALID_ARG, "no output requested");
  if (!gatq_large_x_fits(cfg->n_agents, cfg->graph_mode == SWARM_GRAPH_RADIUS))
    return fail(SWARM_ERR_UNSUPPORTED, "the env does not fit shared memory");
  return check_cuda(launch_gatq_large_x(*cfg, weights, state, q, actions, (cudaStream_t)stream), "swarm_gatq_forward_large");
}

int64_t swarm_rollout_large_workspace_bytes(const SwarmConfig* cfg) {
  if (!cfg || cfg->num_envs <= 0 || cfg->n_agents <= 0) return 0;
  const int64_t bn = (int64_t)cfg->num_envs * cfg->n_agents;
  const int64_t k = cfg->graph_mode == SWARM_GRAPH_KNN ? (cfg->knn_k > 0 ? cfg->knn_k : 0) : 0;
  return bn * k * 4 + bn * 4 + 512;
}

int swarm_rollout_large(const SwarmConfig* cfg, const float* weights, float* state, int32_t ticks, float* returns,
                        int32_t* hits, void* workspace, int64_t workspace_bytes, void* stream) {
  if (int rc = validate(cfg, true, true)) return rc;
  if (cfg->n_agents <= kTileThreads) return fail(SWARM_ERR_INVALID_ARG, "n_agents <= 128: use swarm_rollout");
  if (!weights || !state || !workspace) return fail(SWARM_ERR_INVALID_ARG, "weights/state/workspace is NULL");
  if (ticks < 0) return fail(SWARM_ERR_INVALID_ARG, "ticks must be >= 0");
  const bool knn = cfg->graph_mode == SWARM_GRAPH_KNN;
  if (knn && (int64_t)cfg->knn_k * 64 > cfg->n_agents)
    return fail(SWARM_ERR_UNSUPPORTED, "large-swarm kNN implements torch.topk's partial_sort branch (64 k <= n)");
  // attention in input space (no projected-feature tile: envs up to 4 096 agents fit) unless SWARM_TC=0 asks for the
  // bit-faithful forward of swarm_gatq_forward_knn_large (kNN only; the radius / complete graph of a large swarm has
  // the input-space forward alone -- their bit-faithful path is the CSR one)
  const char* tc_env = std::getenv("SWARM_TC");
  const bool xspace = !knn || !(tc_env && tc_env[0] == '0');
  const bool fits = !knn ? gatq_large_x_fits(cfg->n_agents, cfg->graph_mode == SWARM_GRAPH_RADIUS)
                         : (xspace ? gatq_knn_large_x_fits(cfg->n_agents, cfg->knn_k)
                                   : gatq_knn_large_fits(cfg->n_agents, cfg->knn_k));
  if (!fits)
    return fail(SWARM_ERR_UNSUPPORTED, "the env does not fit shared memory: step it with swarm_graph_build + "
                                       "swarm_csr_from_edges + swarm_gatq_forward_csr + swarm_sim_step");
  if (workspace_bytes < swarm_rollout_large_workspace_bytes(cfg)) return fail(SWARM_ERR_INVALID_ARG, "workspace too small");
  TileParams p;
  if (int rc = fill_params(p, cfg, MODE_STEP)) return rc;
  const int64_t bn = (int64_t)cfg->num_envs * cfg->n_agents;
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255;
  int32_t* nbr = reinterpret_cast<int32_t*>(base);
  int32_t* actions = nbr + (knn ? bn * cfg->knn_k : 0);
  p.state_in = state;
  p.state_out = state;           // in place: one CTA owns a whole env and stages its positions first
  p.actions_in = actions;
  p.returns = returns;
  p.hits = hits;
  cudaStream_t st = (cudaStream_t)stream;
  for (int t = 0; t < ticks; ++t) {
    // simulator.py:59-68 per tick: topk table -> GCN forward + argmax straight from the table -> Environment.step,
    // returns / hits accumulated by the step kernel: three launches, nothing but the state leaves the device
    // (radius / complete graph: the forward finds its own sources -- uniform-grid broad phase / every other agent --
    // so a tick is two launches)
    if (knn) {
      if (cudaError_t e = launch_graph_large(*cfg, state, nullptr, nbr, p.edges_per_env, st); e != cudaSuccess)
        return check_cuda(e, "swarm_rollout_large (graph)");
      if (cudaError_t e = xspace ? launch_gatq_knn_large_x(*cfg, weights, state, nbr, nullptr, actions, st)
                                 : launch_gatq_knn_large(*cfg, weights, state, nbr, nullptr, actions, st);
          e != cudaSuccess)
        return check_cuda(e, "swarm_rollout_large (forward)");
    } else if (cudaError_t e = launch_gatq_large_x(*cfg, weights, state, nullptr, actions, st); e != cudaSuccess) {
      return check_cuda(e, "swarm_rollout_large (forward)");
    }
    if (cudaError_t e = launch_sim_step_large(p, st); e != cudaSuccess) return check_cuda(e, "swarm_rollout_large (step)");
  }
  return SWARM_OK;
}

int swarm_gatconv_forward_csr(int32_t n_nodes, const float* weights, const float* x, const int32_t* row_ptr,
                              const int32_t* src, float* out, void* workspace, int64_t workspace_bytes, void* stream) {
  if (n_nodes < 0) return fail(SWARM_ERR_INVALID_ARG, "n_nodes must be >= 0");
  if (n_nodes == 0) return SWARM_OK;
  if (!weights || !x || !row_ptr || !out || !workspace) return fail(SWARM_ERR_INVALID_ARG, "NULL argument");
  if (workspace_bytes < gatq_workspace_bytes(n_nodes)) return fail(SWARM_ERR_INVALID_ARG, "workspace too small");
  float* rows = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  return check_cuda(launch_gatq_csr(n_nodes, weights, x, row_ptr, src, out, nullptr, rows, (cudaStream_t)stream, true),
                    "swarm_gatconv_forward_csr");
}

int swarm_gatconv_backward_csr(int32_t n_nodes, int64_t n_edges, const float* weights, const float* x,
                               const int32_t* row_ptr, const int32_t* src, const int32_t* perm, const int32_t* row_ptr_s,
                               const int32_t* tgt_s, const int32_t* perm_s, const float* grad_out, float* grad_weights,
                               void* workspace, int64_t workspace_bytes, void* stream) {
  if (n_nodes <= 0 || n_edges < 0) return fail(SWARM_ERR_INVALID_ARG, "n_nodes must be > 0 and n_edges >= 0");
  if (n_edges >= (1LL << 31)) return fail(SWARM_ERR_UNSUPPORTED, "more than 2^31 - 1 edges");
  if (!weights || !x || !row_ptr || !row_ptr_s || !grad_out || !grad_weights || !workspace)
    return fail(SWARM_ERR_INVALID_ARG, "NULL argument");
  if (n_edges > 0 && (!src || !perm || !tgt_s || !perm_s)) return fail(SWARM_ERR_INVALID_ARG, "NULL edge array");
  if (workspace_bytes < gatq_backward_workspace_bytes(n_nodes, n_edges))
    return fail(SWARM_ERR_INVALID_ARG, "workspace too small");
  return check_cuda(launch_gatq_backward_csr(n_nodes, n_edges, weights, x, row_ptr, src, perm, row_ptr_s, tgt_s, perm_s,
                                             grad_out, grad_weights, workspace, (cudaStream_t)stream, true),
                    "swarm_gatconv_backward_csr");
}

static int gat_layer_check_dims(int32_t n_nodes, int64_t n_edges, int32_t ci, int32_t co) {
  if (n_nodes < 0 || n_edges < 0) return fail(SWARM_ERR_INVALID_ARG, "n_nodes and n_edges must be >= 0");
  if (n_edges >= (1LL << 31)) return fail(SWARM_ERR_UNSUPPORTED, "more than 2^31 - 1 edges");
  if (ci < 1 || co < 1) return fail(SWARM_ERR_INVALID_ARG, "in_channels and out_channels must be >= 1");
  if (ci > 64 || co > 64) return fail(SWARM_ERR_UNSUPPORTED, "the GAT layer kernels cover in_channels, out_channels <= 64");
  if ((int64_t)n_nodes * 68 >= (1LL << 31)) return fail(SWARM_ERR_UNSUPPORTED, "too many nodes for one call");
  return SWARM_OK;
}

int64_t swarm_gat_layer_workspace_bytes(int32_t n_nodes, int64_t n_edges, int32_t in_channels, int32_t out_channels,
                                        int32_t backward) {
  if (gat_layer_check_dims(n_nodes, n_edges, in_channels, out_channels)) return -1;
  return gat_layer_workspace_bytes(n_nodes, n_edges, in_channels, out_channels, backward != 0);
}

int swarm_gat_layer_forward(int32_t n_nodes, int32_t in_channels, int32_t out_channels, const float* lin_weight,
                            const float* att_src, const float* att_dst, const float* bias, const float* x,
                            const int32_t* row_ptr, const int32_t* src, float* out, void* workspace,
                            int64_t workspace_bytes, void* stream) {
  if (int rc = gat_layer_check_dims(n_nodes, 0, in_channels, out_channels)) return rc;
  if (n_nodes == 0) return SWARM_OK;
  if (!lin_weight || !att_src || !att_dst || !bias || !x || !row_ptr || !out || !workspace)
    return fail(SWARM_ERR_INVALID_ARG, "NULL argument");
  if (workspace_bytes < gat_layer_workspace_bytes(n_nodes, 0, in_channels, out_channels, false))
    return fail(SWARM_ERR_INVALID_ARG, "workspace too small");
  return check_cuda(launch_gat_layer_forward(n_nodes, in_channels, out_channels, lin_weight, att_src, att_dst, bias, x,
                                             row_ptr, src, out, workspace, (cudaStream_t)stream),
                    "swarm_gat_layer_forward");
}

int swarm_gat_layer_backward(int32_t n_nodes, int64_t n_edges, int32_t in_channels, int32_t out_channels,
                             const float* lin_weight, const float* att_src, const float* att_dst, const float* x,
                             const int32_t* row_ptr, const int32_t* src, const int32_t* perm, const int32_t* row_ptr_s,
                             const int32_t* tgt_s, const int32_t* perm_s, const float* grad_out, float* grad_lin_weight,
                             float* grad_att_src, float* grad_att_dst, float* grad_bias, float* grad_x, void* workspace,
                             int64_t workspace_bytes, void* stream) {
  if (int rc = gat_layer_check_dims(n_nodes, n_edges, in_channels, out_channels)) return rc;
  if (n_nodes == 0) return fail(SWARM_ERR_INVALID_ARG, "n_nodes must be > 0");
  if (!lin_weight || !att_src || !att_dst || !x || !row_ptr || !row_ptr_s || !grad_out || !grad_lin_weight ||
      !grad_att_src || !grad_att_dst || !grad_bias || !workspace)
    return fail(SWARM_ERR_INVALID_ARG, "NULL argument");
  if (n_edges > 0 && (!src || !perm || !tgt_s || !perm_s)) return fail(SWARM_ERR_INVALID_ARG, "NULL edge array");
  if (workspace_bytes < gat_layer_workspace_bytes(n_nodes, n_edges, in_channels, out_channels, true))
    return fail(SWARM_ERR_INVALID_ARG, "workspace too small");
  return check_cuda(launch_gat_layer_backward(n_nodes, n_edges, in_channels, out_channels, lin_weight, att_src, att_dst, x,
                                              row_ptr, src, perm, row_ptr_s, tgt_s, perm_s, grad_out, grad_lin_weight,
                                              grad_att_src, grad_att_dst, grad_bias, grad_x, workspace,
                                              (cudaStream_t)stream),
                    "swarm_gat_layer_backward");
}

int64_t swarm_gatq_backward_workspace_bytes(int32_t n_nodes, int64_t n_edges) {
  if (n_nodes <= 0 || n_edges < 0) return 0;
  return gatq_backward_workspace_bytes(n_nodes, n_edges);
}

int swarm_gatq_backward_csr(int32_t n_nodes, int64_t n_edges, const float* weights, const float* x, const int32_t* row_ptr,
                            const int32_t* src, const int32_t* perm, const int32_t* row_ptr_s, const int32_t* tgt_s,
                            const int32_t* perm_s, const float* grad_q, float* grad_weights, void* workspace,
                            int64_t workspace_bytes, void* stream) {
  if (n_nodes <= 0 || n_edges < 0) return fail(SWARM_ERR_INVALID_ARG, "n_nodes must be > 0 and n_edges >= 0");
  if (n_edges >= (1LL << 31)) return fail(SWARM_ERR_UNSUPPORTED, "more than 2^31 - 1 edges");
  if (!weights || !x || !row_ptr || !row_ptr_s || !grad_q || !grad_weights || !workspace)
    return fail(SWARM_ERR_INVALID_ARG, "NULL argument");
  if (n_edges > 0 && (!src || !perm || !tgt_s || !perm_s)) return fail(SWARM_ERR_INVALID_ARG, "NULL edge array");
  if (workspace_bytes < gatq_backward_workspace_bytes(n_nodes, n_edges))
    return fail(SWARM_ERR_INVALID_ARG, "workspace too small");
  return check_cuda(launch_gatq_backward_csr(n_nodes, n_edges, weights, x, row_ptr, src, perm, row_ptr_s, tgt_s, perm_s,
                                             grad_q, grad_weights, workspace, (cudaStream_t)stream),
                    "swarm_gatq_backward_csr");
}

int64_t swarm_csr_workspace_bytes(int32_t n_nodes, int64_t n_edges) {
  if (n_nodes <= 0 || n_edges < 0) return 0;
  return csr_workspace_bytes(n_nodes, n_edges);
}

int swarm_csr_from_edges(int32_t n_nodes, int64_t n_edges, const int64_t* edge_src, const int64_t* edge_dst,
                         int32_t* row_ptr, int32_t* src, int32_t* perm, void* workspace, int64_t workspace_bytes,
                         void* stream) {
  if (n_nodes <= 0 || n_edges < 0) return fail(SWARM_ERR_INVALID_ARG, "n_nodes must be > 0 and n_edges >= 0");
  if (n_edges >= (1LL << 31)) return fail(SWARM_ERR_UNSUPPORTED, "more than 2^31 - 1 edges");
  if (!row_ptr || !workspace) return fail(SWARM_ERR_INVALID_ARG, "NULL argument");
  if (n_edges > 0 && (!edge_src || !edge_dst || !src || !perm)) return fail(SWARM_ERR_INVALID_ARG, "NULL edge array");
  if (workspace_bytes < csr_workspace_bytes(n_nodes, n_edges)) return fail(SWARM_ERR_INVALID_ARG, "workspace too small");
  return check_cuda(launch_csr_from_edges(n_nodes, n_edges, edge_src, edge_dst, row_ptr, src, perm, workspace,
                                          workspace_bytes, (cudaStream_t)stream),
                    "swarm_csr_from_edges");
}

// SwarmRolloutOptions / SwarmTrainHyper: the optional Flocking reward of the fused tick
static int attach_flocking(TileParams& p, const SwarmConfig* cfg, const SwarmRewardSpec* fs, float* shaping) {
  if (!fs) return SWARM_OK;
  if (fs->kind != SWARM_REWARD_FLOCKING) return fail(SWARM_ERR_INVALID_ARG, "`flocking` must be a Flocking reward spec");
  if (cfg->scenario != SWARM_SCENARIO_GOTO)
    return fail(SWARM_ERR_INVALID_ARG, "the Flocking reward runs on the GoTo world (cfg->scenario = SWARM_SCENARIO_GOTO)");
  if (cfg->n_agents < 2) return fail(SWARM_ERR_INVALID_ARG, "the scenario rewards need at least two agents");
  if (!shaping) return fail(SWARM_ERR_INVALID_ARG, "Flocking needs the shaping buffer");
  p.flock = *fs;
  p.shaping = reinterpret_cast<float2*>(shaping);
  p.use_flock = 1;
  return SWARM_OK;
}

int swarm_rollout(const SwarmConfig* cfg, const float* weights, float* state, int32_t ticks,
                  const SwarmRolloutOptions* opts, float* returns, int32_t* hits, const SwarmTrace* trace, void* stream) {
  if (int rc = validate(cfg, true)) return rc;
  if (!weights || !state) return fail(SWARM_ERR_INVALID_ARG, "weights/state is NULL");
  if (ticks < 0) return fail(SWARM_ERR_INVALID_ARG, "ticks must be >= 0");
  if (ticks == 0) return SWARM_OK;
  if (trace && trace->contact && cfg->n_agents > 32)
    return fail(SWARM_ERR_UNSUPPORTED, "contact masks need n_agents <= 32");
  if (trace && trace->edges && cfg->graph_mode == SWARM_GRAPH_RADIUS)
    return fail(SWARM_ERR_UNSUPPORTED, "per-tick edge traces are not available for radius graphs (variable edge count)");
  TileParams p;
  if (int rc = fill_params(p, cfg, MODE_ROLLOUT)) return rc;
  p.weights = weights;
  p.state_in = state;
  p.state_out = state;
  p.ticks = ticks;
  p.returns = returns;
  p.hits = hits;
  if (opts) {
    if (opts->epsilon < 0.0f || opts->epsilon > 1.0f) return fail(SWARM_ERR_INVALID_ARG, "epsilon must be in [0, 1]");
    p.actions_in = opts->forced_actions;
    p.epsilon = opts->epsilon;
    p.rng_seed = opts->rng_seed;
    p.rng_tick0 = opts->rng_tick0;
    p.env_offset = opts->env_offset;
    if (opts->replay) {
      const SwarmReplay& r = *opts->replay;
      if (!r.state || !r.next_state || !r.actions || !r.rewards || r.capacity <= 0)
        return fail(SWARM_ERR_INVALID_ARG, "replay ring has NULL arrays or no capacity");
      if ((int64_t)ticks * cfg->num_envs > r.capacity)
        return fail(SWARM_ERR_INVALID_ARG, "ticks * num_envs exceeds the replay capacity (slots would be overwritten inside one call)");
      p.replay = r;
      p.replay_cursor = opts->replay_cursor % r.capacity;
    }
    if (int rc = attach_flocking(p, cfg, opts->flocking, opts->flocking_shaping)) return rc;
    if (opts->knn_memo) {
      const int64_t n = opts->knn_memo_entries;
      if (n <= 0 || (n & (n - 1)) != 0 || n > (1LL << 32))
        return fail(SWARM_ERR_INVALID_ARG, "knn_memo_entries must be a power of two (<= 2^32)");
      p.knn_memo = reinterpret_cast<unsigned long long*>(opts->knn_memo);
      p.knn_memo_mask = (uint32_t)(n - 1);
    }
  }
  if (trace) p.trace = *trace;
  return check_cuda(launch_tile(MODE_ROLLOUT, p, (cudaStream_t)stream), "swarm_rollout");
}

int swarm_replay_push(const SwarmConfig* cfg, const SwarmReplay* replay, int64_t cursor, const float* state,
                      const int32_t* actions, const float* rewards, const float* next_state, void* stream) {
  if (int rc = validate(cfg, false)) return rc;
  if (!replay || !replay->state || !replay->next_state || !replay->actions || !replay->rewards || replay->capacity <= 0)
    return fail(SWARM_ERR_INVALID_ARG, "replay ring has NULL arrays or no capacity");
  if (!state || !actions || !rewards || !next_state) return fail(SWARM_ERR_INVALID_ARG, "NULL transition array");
  if (cfg->num_envs > replay->capacity) return fail(SWARM_ERR_INVALID_ARG, "num_envs exceeds the replay capacity");
  if (cursor < 0) return fail(SWARM_ERR_INVALID_ARG, "cursor must be >= 0");
  return check_cuda(launch_replay_push(*replay, cursor % replay->capacity, cfg->num_envs, cfg->n_agents, state, actions,
                                       rewards, next_state, (cudaStream_t)stream),
                    "swarm_replay_push");
}

int swarm_replay_gather(const SwarmConfig* cfg, const SwarmReplay* replay, const int64_t* indices, int32_t n_graphs,
                        float* state, int32_t* actions, float* rewards, float* next_state, void* stream) {
  if (!cfg || cfg->n_agents <= 0) return fail(SWARM_ERR_INVALID_ARG, "cfg is NULL or n_agents <= 0");
  if (!replay || !replay->state || !replay->next_state || !replay->actions || !replay->rewards)
    return fail(SWARM_ERR_INVALID_ARG, "replay ring has NULL arrays");
  if (n_graphs < 0) return fail(SWARM_ERR_INVALID_ARG, "n_graphs must be >= 0");
  if (n_graphs == 0) return SWARM_OK;
  if (!indices || !state || !actions || !rewards || !next_state) return fail(SWARM_ERR_INVALID_ARG, "NULL argument");
  return check_cuda(launch_replay_gather(*replay, indices, n_graphs, cfg->n_agents, state, actions, rewards, next_state,
                                         (cudaStream_t)stream),
                    "swarm_replay_gather");
}

static int validate_dqn(const SwarmConfig* cfg, int32_t n_graphs) {
  if (!cfg) return fail(SWARM_ERR_INVALID_ARG, "cfg is NULL");
  SwarmConfig c = *cfg;
  c.num_envs = n_graphs > 0 ? n_graphs : 1;
  if (int rc = validate(&c, true)) return rc;
  if (n_graphs <= 0) return fail(SWARM_ERR_INVALID_ARG, "n_graphs must be positive");
  if (dqn_smem_bytes(c) > 227 * 1024)
    return fail(SWARM_ERR_UNSUPPORTED, "the DQN gradient tile exceeds 227 KB of shared memory for this (N, k)");
  return SWARM_OK;
}

int64_t swarm_dqn_workspace_bytes(const SwarmConfig* cfg, int32_t n_graphs) {
  if (validate_dqn(cfg, n_graphs) != SWARM_OK) return 0;
  return dqn_workspace_bytes(*cfg, n_graphs);
}

int swarm_dqn_grad(const SwarmConfig* cfg, const float* online_weights, const float* target_weights,
                   const SwarmReplay* batch, const int64_t* indices, int32_t n_graphs, float gamma, float loss_scale,
                   float* grad, float* loss, float* td, void* workspace, int64_t workspace_bytes, void* stream) {
  if (int rc = validate_dqn(cfg, n_graphs)) return rc;
  if (!online_weights || !target_weights || !grad || !loss || !workspace) return fail(SWARM_ERR_INVALID_ARG, "NULL argument");
  if (!batch || !batch->state || !batch->next_state || !batch->actions || !batch->rewards)
    return fail(SWARM_ERR_INVALID_ARG, "batch has NULL arrays");
  if (workspace_bytes < dqn_workspace_bytes(*cfg, n_graphs)) return fail(SWARM_ERR_INVALID_ARG, "workspace too small");
  return check_cuda(launch_dqn_grad(*cfg, online_weights, target_weights, *batch, indices, n_graphs, gamma, loss_scale,
                                    grad, loss, td, workspace, (cudaStream_t)stream),
                    "swarm_dqn_grad");
}

int swarm_adam_clip_step(float* weights, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t step, double lr,
                         double beta1, double beta2, double eps, double max_norm, float* target_weights,
                         float* grad_norm, void* stream) {
  if (!weights || !grad || !exp_avg || !exp_avg_sq) return fail(SWARM_ERR_INVALID_ARG, "NULL argument");
  if (step < 1) return fail(SWARM_ERR_INVALID_ARG, "step is 1-based");
  if (!(lr >= 0.0) || !(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0) || !(eps >= 0.0))
    return fail(SWARM_ERR_INVALID_ARG, "invalid Adam hyper-parameter");     // torch.optim.Adam's ValueError checks
  return check_cuda(launch_adam_clip(weights, grad, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps, max_norm,
                                     target_weights, grad_norm, (cudaStream_t)stream),
                    "swarm_adam_clip_step");
}

static int validate_hyper(const SwarmTrainHyper* h) {
  if (!h) return fail(SWARM_ERR_INVALID_ARG, "hyper is NULL");
  if (h->graphs_per_update <= 0) return fail(SWARM_ERR_INVALID_ARG, "graphs_per_update must be positive");
  if (h->update_target_every <= 0) return fail(SWARM_ERR_INVALID_ARG, "update_target_every must be positive");
  if (!(h->lr >= 0.0) || !(h->beta1 >= 0.0 && h->beta1 < 1.0) || !(h->beta2 >= 0.0 && h->beta2 < 1.0) || !(h->eps >= 0.0))
    return fail(SWARM_ERR_INVALID_ARG, "invalid Adam hyper-parameter");
  return SWARM_OK;
}

// shared argument checks and rollout-tick parameters of swarm_train_tick_grad / swarm_train_tick
static int train_tick_rollout(const SwarmConfig* cfg, const SwarmTrainHyper* hyper, SwarmTrainCtl* ctl, const float* weights,
                              const float* target_weights, float* state, float* returns, int32_t* hits,
                              const SwarmReplay* ring, float* grad, float* loss, void* workspace, int64_t workspace_bytes,
                              cudaStream_t st, const char* what) {
  if (int rc = validate(cfg, true)) return rc;
  if (int rc = validate_hyper(hyper)) return rc;
  if (int rc = validate_dqn(cfg, hyper->graphs_per_update)) return rc;
  if (!ctl || !weights || !target_weights || !state || !grad || !loss || !workspace)
    return fail(SWARM_ERR_INVALID_ARG, "NULL argument");
  if (!ring || !ring->state || !ring->next_state || !ring->actions || !ring->rewards || ring->capacity <= 0)
    return fail(SWARM_ERR_INVALID_ARG, "replay ring has NULL arrays or no capacity");
  if (cfg->num_envs > ring->capacity) return fail(SWARM_ERR_INVALID_ARG, "num_envs exceeds the replay capacity");
  if (workspace_bytes < dqn_workspace_bytes(*cfg, hyper->graphs_per_update))
    return fail(SWARM_ERR_INVALID_ARG, "workspace too small");
  TileParams p;
  if (int rc = fill_params(p, cfg, MODE_ROLLOUT)) return rc;
  p.weights = weights;
  p.state_in = state;
  p.state_out = state;
  p.ticks = 1;
  p.returns = returns;
  p.hits = hits;
  p.rng_seed = hyper->rng_seed;
  p.env_offset = hyper->env_offset;
  p.replay = *ring;
  p.ctl = ctl;
  if (int rc = attach_flocking(p, cfg, hyper->flocking, hyper->flocking_shaping)) return rc;
  return check_cuda(launch_tile(MODE_ROLLOUT, p, st), what);
}

static int validate_peers(const SwarmPeerExchange* peers) {
  if (!peers) return SWARM_OK;
  if (peers->world_size < 1 || peers->world_size > SWARM_MAX_PEERS || peers->rank < 0 || peers->rank >= peers->world_size)
    return fail(SWARM_ERR_INVALID_ARG, "peer exchange: bad world_size / rank");
  for (int r = 0; r < peers->world_size; ++r)
    if (!peers->data[r]) return fail(SWARM_ERR_INVALID_ARG, "peer exchange: NULL peer buffer");
  return SWARM_OK;
}

int swarm_train_tick_grad(const SwarmConfig* cfg, const SwarmTrainHyper* hyper, SwarmTrainCtl* ctl, const float* weights,
                          const float* target_weights, float* state, float* returns, int32_t* hits,
                          const SwarmReplay* ring, int64_t* indices, float* grad, float* loss, void* workspace,
                          int64_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = train_tick_rollout(cfg, hyper, ctl, weights, target_weights, state, returns, hits, ring, grad, loss,
                                  workspace, workspace_bytes, st, "swarm_train_tick_grad(rollout)"))
    return rc;
  return check_cuda(launch_dqn_grad(*cfg, weights, target_weights, *ring, nullptr, hyper->graphs_per_update, hyper->gamma,
                                    hyper->loss_scale, grad, loss, nullptr, workspace, st, ctl, indices,
                                    hyper->sample_seed, cfg->num_envs),
                    "swarm_train_tick_grad(grad)");
}

int swarm_train_tick_apply(const SwarmConfig* cfg, const SwarmTrainHyper* hyper, SwarmTrainCtl* ctl, float* weights,
                           float* target_weights, float* exp_avg, float* exp_avg_sq, float* grad,
                           int64_t ring_capacity, const SwarmPeerExchange* peers, void* stream) {
  if (!cfg || cfg->num_envs <= 0) return fail(SWARM_ERR_INVALID_ARG, "cfg is NULL or num_envs <= 0");
  if (int rc = validate_hyper(hyper)) return rc;
  if (!ctl || !weights || !target_weights || !exp_avg || !exp_avg_sq || !grad) return fail(SWARM_ERR_INVALID_ARG, "NULL argument");
  if (ring_capacity < cfg->num_envs) return fail(SWARM_ERR_INVALID_ARG, "ring_capacity must be >= num_envs");
  if (int rc = validate_peers(peers)) return rc;
  return check_cuda(launch_adam_clip(weights, grad, exp_avg, exp_avg_sq, 1, hyper->lr, hyper->beta1, hyper->beta2,
                                     hyper->eps, hyper->max_norm, target_weights, nullptr, (cudaStream_t)stream, ctl,
                                     cfg->num_envs, ring_capacity, hyper->update_target_every, peers, grad),
                    "swarm_train_tick_apply");
}

int swarm_train_tick(const SwarmConfig* cfg, const SwarmTrainHyper* hyper, SwarmTrainCtl* ctl, float* weights,
                     float* target_weights, float* exp_avg, float* exp_avg_sq, float* state, float* returns,
                     int32_t* hits, const SwarmReplay* ring, int64_t* indices, float* grad, void* workspace,
                     int64_t workspace_bytes, const SwarmPeerExchange* peers, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!exp_avg || !exp_avg_sq) return fail(SWARM_ERR_INVALID_ARG, "NULL argument");
  if (int rc = validate_peers(peers)) return rc;
  if (int rc = train_tick_rollout(cfg, hyper, ctl, weights, target_weights, state, returns, hits, ring, grad,
                                  grad ? grad + SWARM_W_COUNT : nullptr, workspace, workspace_bytes, st,
                                  "swarm_train_tick(rollout)"))
    return rc;
  const int G = hyper->graphs_per_update;
  // gradient kernel on few CTAs: partial reduction (+ peer exchange) + clip + Adam as one cluster launch; a large update
  // batch keeps the reduce launch and the single-CTA clip + Adam kernel
  const bool fuse = dqn_fuse_reduce(*cfg, G);
  if (int rc = check_cuda(launch_dqn_grad(*cfg, weights, target_weights, *ring, nullptr, G, hyper->gamma, hyper->loss_scale,
                                          grad, grad + SWARM_W_COUNT, nullptr, workspace, st, ctl, indices,
                                          hyper->sample_seed, cfg->num_envs, fuse),
                          "swarm_train_tick(grad)"))
    return rc;
  return check_cuda(launch_adam_clip(weights, grad, exp_avg, exp_avg_sq, 1, hyper->lr, hyper->beta1, hyper->beta2,
                                     hyper->eps, hyper->max_norm, target_weights, nullptr, st, ctl, cfg->num_envs,
                                     ring->capacity, hyper->update_target_every, peers, grad, fuse ? cfg : nullptr,
                                     fuse ? workspace : nullptr, G, hyper->loss_scale),
                    "swarm_train_tick(apply)");
}

int swarm_reset_random(const SwarmConfig* cfg, const SwarmResetSpec* spec, const SwarmTrainCtl* ctl, int64_t episode,
                       float* centers_out, float* state, void* stream) {
  if (int rc = validate(cfg, false, true)) return rc;
  if (!spec || !state) return fail(SWARM_ERR_INVALID_ARG, "spec/state is NULL");
  if (!(spec->std_x >= 0.0f) || !(spec->std_y >= 0.0f)) return fail(SWARM_ERR_INVALID_ARG, "std must be >= 0");
  if (!ctl && episode < 0) return fail(SWARM_ERR_INVALID_ARG, "episode must be >= 0");
  const int n = cfg->n_agents;
  const int cols = (int)std::ceil(std::sqrt((double)n));
  const int rows = (int)std::ceil((double)n / (double)cols);
  return check_cuda(launch_reset_random(*cfg, *spec, cols, rows, ctl, episode, centers_out, state, (cudaStream_t)stream),
                    "swarm_reset_random");
}

void swarm_default_reward_spec(SwarmRewardSpec* spec, int32_t kind, int32_t num_envs, int32_t n_agents) {
  if (!spec) return;
  *spec = SwarmRewardSpec{};
  spec->kind = kind;
  spec->num_envs = num_envs;
  spec->n_agents = n_agents;
  spec->reset = 0;
  spec->env_index = -1;
  spec->goal_x = -0.8f;
  spec->goal_y = 0.8f;
  spec->goal_radius = 0.05f;
  spec->agent_radius = 0.05f;
  spec->pos_shaping = 10.0f;
  spec->dist_shaping = 10.0f;
  spec->desired_distance = 0.15f;
  spec->min_collision_distance = 0.005f;
  spec->collision_reward = -1.0f;
  spec->on_goal_bonus = 50.0f;
  spec->sigma = 0.15f;
}

int swarm_scenario_reward(const SwarmRewardSpec* spec, const float* state, float* shaping, float* reward, float* terms,
                          void* stream) {
  if (!spec) return fail(SWARM_ERR_INVALID_ARG, "spec is NULL");
  if (spec->kind != SWARM_REWARD_FLOCKING && spec->kind != SWARM_REWARD_COHESION)
    return fail(SWARM_ERR_INVALID_ARG, "unknown reward kind");
  if (spec->num_envs < 0) return fail(SWARM_ERR_INVALID_ARG, "num_envs must be >= 0");
  if (spec->n_agents < 2) return fail(SWARM_ERR_INVALID_ARG, "the scenario rewards need at least two agents");
  if (spec->n_agents > 128) return fail(SWARM_ERR_UNSUPPORTED, "scenario rewards are implemented for n_agents <= 128");
  if (spec->env_index < -1 || spec->env_index >= spec->num_envs)
    return fail(SWARM_ERR_INVALID_ARG, "env_index out of range");
  if ((int64_t)spec->num_envs * spec->n_agents + 128 * 148 * 16 >= (1LL << 31))
    return fail(SWARM_ERR_UNSUPPORTED, "num_envs * n_agents must stay below 2^31");
  if (spec->num_envs == 0) return SWARM_OK;
  if (!state) return fail(SWARM_ERR_INVALID_ARG, "state is NULL");
  if (spec->kind == SWARM_REWARD_FLOCKING) {
    if (!shaping) return fail(SWARM_ERR_INVALID_ARG, "Flocking needs the shaping buffer");
    if (!spec->reset && !reward) return fail(SWARM_ERR_INVALID_ARG, "reward is NULL");
  } else {
    if (spec->reset) return fail(SWARM_ERR_INVALID_ARG, "Cohesion keeps no shaping memory (reset must be 0)");
    if (!reward) return fail(SWARM_ERR_INVALID_ARG, "reward is NULL");
    if (!(spec->sigma > 0.0f)) return fail(SWARM_ERR_INVALID_ARG, "sigma must be > 0");
  }
  return check_cuda(launch_scenario_reward(*spec, state, shaping, reward, terms, (cudaStream_t)stream),
                    "swarm_scenario_reward");
}

namespace {
int validate_stack(const SwarmConfig* cfg, const SwarmStackSpec* spec) {
  if (int rc = validate(cfg, true)) return rc;
  if (!spec) return fail(SWARM_ERR_INVALID_ARG, "spec is NULL");
  if (spec->n_layers < 1 || spec->n_layers > 4) return fail(SWARM_ERR_INVALID_ARG, "n_layers must be 1 .. 4");
  if (spec->hidden < 1 || spec->hidden > 32) return fail(SWARM_ERR_UNSUPPORTED, "hidden must be 1 .. 32");
  if (spec->in_features != 7 && spec->in_features != 5) return fail(SWARM_ERR_INVALID_ARG, "in_features must be 7 or 5");
  for (int l = 0; l < spec->n_layers; ++l)
    if (spec->activation[l] != SWARM_ACT_TANH && spec->activation[l] != SWARM_ACT_RELU)
      return fail(SWARM_ERR_INVALID_ARG, "unknown activation");
  if (cfg->graph_mode == SWARM_GRAPH_KNN && cfg->n_agents > 16)
    return fail(SWARM_ERR_UNSUPPORTED, "stacked networks on the kNN graph are implemented for n_agents <= 16");
  return SWARM_OK;
}
}  // namespace

int64_t swarm_stack_weight_count(const SwarmStackSpec* spec) {
  if (!spec || spec->n_layers < 1 || spec->n_layers > 4 || spec->hidden < 1 || spec->hidden > 32) return 0;
  return stack_weight_count_host(*spec);
}

int swarm_gatstack_forward(const SwarmConfig* cfg, const SwarmStackSpec* spec, const float* weights, const float* state,
                           float* q, int32_t* actions, void* stream) {
  if (int rc = validate_stack(cfg, spec)) return rc;
  if (!weights || !state) return fail(SWARM_ERR_INVALID_ARG, "weights/state is NULL");
  if (!q && !actions) return fail(SWARM_ERR_INVALID_ARG, "no output requested");
  return check_cuda(launch_gatstack_forward(*cfg, *spec, weights, state, q, actions, (cudaStream_t)stream),
                    "swarm_gatstack_forward");
}

int64_t swarm_rollout_stack_workspace_bytes(const SwarmConfig* cfg) {
  if (!cfg || cfg->num_envs <= 0 || cfg->n_agents <= 0) return 0;
  const int64_t bn = (int64_t)cfg->num_envs * cfg->n_agents;
  return bn * 4 + bn * 4 + bn + 1024;             // actions, rewards, flags
}

int swarm_rollout_stack(const SwarmConfig* cfg, const SwarmStackSpec* spec, const float* weights, float* state,
                        int32_t ticks, const SwarmRewardSpec* reward, float* shaping, float* returns, int32_t* hits,
                        void* workspace, int64_t workspace_bytes, void* stream) {
  if (int rc = validate_stack(cfg, spec)) return rc;
  if (!weights || !state || !workspace) return fail(SWARM_ERR_INVALID_ARG, "weights/state/workspace is NULL");
  if (ticks < 0) return fail(SWARM_ERR_INVALID_ARG, "ticks must be >= 0");
  if (workspace_bytes < swarm_rollout_stack_workspace_bytes(cfg)) return fail(SWARM_ERR_INVALID_ARG, "workspace too small");
  SwarmRewardSpec rs;
  if (reward) {
    rs = *reward;
    if (rs.kind != SWARM_REWARD_FLOCKING && rs.kind != SWARM_REWARD_COHESION)
      return fail(SWARM_ERR_INVALID_ARG, "unknown reward kind");
    if (cfg->scenario != SWARM_SCENARIO_GOTO)
      return fail(SWARM_ERR_INVALID_ARG, "the Flocking / Cohesion rewards run on the GoTo world (cfg->scenario)");
    if (rs.num_envs != cfg->num_envs || rs.n_agents != cfg->n_agents)
      return fail(SWARM_ERR_INVALID_ARG, "reward spec and cfg disagree on num_envs / n_agents");
    if (rs.n_agents < 2) return fail(SWARM_ERR_INVALID_ARG, "the scenario rewards need at least two agents");
    if (rs.kind == SWARM_REWARD_FLOCKING && !shaping) return fail(SWARM_ERR_INVALID_ARG, "Flocking needs the shaping buffer");
    if (rs.kind == SWARM_REWARD_COHESION && !(rs.sigma > 0.0f)) return fail(SWARM_ERR_INVALID_ARG, "sigma must be > 0");
    rs.reset = 0;
    rs.env_index = -1;
  }
  TileParams p;
  if (int rc = fill_params(p, cfg, MODE_STEP)) return rc;
  const int64_t bn = (int64_t)cfg->num_envs * cfg->n_agents;
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255;
  int32_t* actions = reinterpret_cast<int32_t*>(base);
  float* rewards = reinterpret_cast<float*>(actions + bn);
  uint8_t* flags = reinterpret_cast<uint8_t*>(rewards + bn);
  p.state_in = state;
  p.state_out = state;
  p.actions_in = actions;
  p.rewards_out = rewards;
  p.flags_out = flags;
  cudaStream_t st = (cudaStream_t)stream;
  // the whole loop is ONE launch with the state in registers (stack_kernels.cu gatstack_kernel<HP, true>);
  // SWARM_STACK_FUSED=0 takes the launch sequence below (forward, step, reward kernel, totals per tick)
  const char* fused_env = std::getenv("SWARM_STACK_FUSED");
  if (!(fused_env && fused_env[0] == '0'))
    return check_cuda(launch_gatstack_rollout(*cfg, *spec, weights, state, ticks, p, reward ? &rs : nullptr, shaping, returns,
                                              hits, st),
                      "swarm_rollout_stack");
  for (int t = 0; t < ticks; ++t) {
    if (cudaError_t e = launch_gatstack_forward(*cfg, *spec, weights, state, nullptr, actions, st); e != cudaSuccess)
      return check_cuda(e, "swarm_rollout_stack (forward)");
    cudaError_t e = cfg->n_agents <= 64 ? launch_sim_step(p, st) : launch_tile(MODE_STEP, p, st);
    if (e != cudaSuccess) return check_cuda(e, "swarm_rollout_stack (step)");
    int per_env = 0;
    if (reward) {
      per_env = rs.kind == SWARM_REWARD_FLOCKING ? 1 : 0;
      if (cudaError_t e2 = launch_scenario_reward(rs, state, shaping, rewards, nullptr, st); e2 != cudaSuccess)
        return check_cuda(e2, "swarm_rollout_stack (reward)");
    }
    if (returns || hits)
      if (cudaError_t e3 = launch_stack_accumulate(bn, cfg->n_agents, rewards, per_env, flags, returns, hits, st);
          e3 != cudaSuccess)
        return check_cuda(e3, "swarm_rollout_stack (accumulate)");
  }
  return SWARM_OK;
}

int swarm_episode_end(const SwarmConfig* cfg, SwarmTrainCtl* ctl, float* returns, int32_t* hits, const float* loss,
                      float* stats, int64_t max_episodes, double epsilon0, double epsilon_decay, double min_epsilon,
                      void* stream) {
  if (int rc = validate(cfg, false, true)) return rc;
  if (!ctl || !returns) return fail(SWARM_ERR_INVALID_ARG, "ctl/returns is NULL");
  if (stats && max_episodes <= 0) return fail(SWARM_ERR_INVALID_ARG, "max_episodes must be positive when stats is given");
  return check_cuda(launch_episode_end(*cfg, ctl, returns, hits, loss, stats, max_episodes, epsilon0, epsilon_decay,
                                       min_epsilon, (cudaStream_t)stream),
                    "swarm_episode_end");
}

}  // extern "C"
