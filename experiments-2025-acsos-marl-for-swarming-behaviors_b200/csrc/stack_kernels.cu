// stack_kernels.cu -- multi-layer GAT Q-networks in ONE launch.
//
// The reference's GCN class carries two more GATConv layers in comments (train_gcn_dqn.py:54-55, 64-67) and ships ten
// checkpoints of that three-layer form (data/models/experiment_Flocking-seed_*.pth: conv1 7 -> 8, conv2 / conv3 8 -> 8,
// lin1 8 -> 8, lin2 8 -> 9).  Stacking the generic layer kernels (gatlayer_kernels.cu) costs one launch set per layer
// plus the edge-list / CSR glue; here a whole forward
//     node features -> per-env graph -> L x [GATConv -> activation] -> lin1 -> ReLU -> lin2 (-> argmax)
// is one env-tile kernel (thread = agent, CTA = floor(128 / N) envs, tile_device.cuh): the in-edge lists are built once
// in shared memory and reused by every layer, the layer outputs never leave the chip.  The input features are the
// reference's [pos, vel, goal, agent id] (train:95-99) or [pos, vel, agent id] for scenarios whose observation() is
// cat[pos, vel] (cohesion_scenario.py:87-94).  Per layer the arithmetic follows torch_geometric's GATConv like the
// one-layer kernels: h = x W^T, alpha = <h, att>, LeakyReLU(0.2) logits, max-subtracted softmax with the 1e-16
// denominator, messages alpha * h_src summed in edge-list order, + bias (alpha = weight x reciprocal of the sum and fused
// multiply-adds: float32-level different rounding than torch's separately rounded ops, like the tensor-core path).
#include <cstring>

#include "tile_device.cuh"

namespace swarm {

struct StackParams {
  SwarmConfig cfg;
  SwarmStackSpec spec;
  const float* weights;
  const float4* state;
  float* q_out;              // [B*N][9] or nullptr
  int32_t* act_out;          // [B*N] or nullptr
  int32_t epb, maxdeg;
  float qmax_r;
  // fused rollout (gatstack_rollout_kernel): `ticks` x [forward -> argmax -> world step -> reward], state in registers
  TileParams step;           // physics constants of the world step (cfg, contact thresholds, drag)
  float4* state_rw;          // [B*N] advanced in place
  float* returns;            // [B*N] += (optional)
  int32_t* hits;             // [B] += (optional)
  SwarmRewardSpec flock;     // use_flock 1: the Flocking collective reward, 2: Cohesion's per-agent reward, instead of the world's
  float2* shaping;           // [B*N] Flocking memory, read at launch, written back
  int32_t use_flock;
  int32_t ticks;
};

constexpr int kStackMaxLayers = 4;

// shared-memory weight block, everything padded to HP channels / 8 input features with zeros
template <int HP>
struct StackLayout {
  static constexpr int kIn0 = 8;                                        // padded input features
  static constexpr int kLayer0 = kIn0 * HP + 3 * HP;                    // WT[8][HP], att_s, att_d, bias
  static constexpr int kLayer = HP * HP + 3 * HP;                       // WT[HP][HP], att_s, att_d, bias
  static constexpr int kW2Cols = 12;
  __host__ __device__ static int layer_off(int l) { return l == 0 ? 0 : kLayer0 + (l - 1) * kLayer; }
  __host__ __device__ static int lin1_off(int L) { return layer_off(L); }
  __host__ __device__ static int lin2_off(int L) { return lin1_off(L) + HP * HP + HP; }
  __host__ __device__ static int total(int L) { return lin2_off(L) + HP * kW2Cols + kW2Cols; }
};

__host__ __device__ inline int stack_weight_count(const SwarmStackSpec& s) {
  const int H = s.hidden;
  int n = 0;
  for (int l = 0; l < s.n_layers; ++l) n += 3 * H + H * (l == 0 ? s.in_features : H);
  return n + H * H + H + 9 * H + 9;
}

// packed weights (state-dict order: per layer att_src, att_dst, bias, lin.weight[H][Cin]; lin1.weight[H][H], lin1.bias,
// lin2.weight[9][H], lin2.bias) -> padded k-major shared layout
template <int HP>
__device__ __forceinline__ void stack_stage_weights(const SwarmStackSpec& s, const float* __restrict__ g, float* __restrict__ sw,
                                                    int tid, int nthreads) {
  using SL = StackLayout<HP>;
  const int H = s.hidden, L = s.n_layers;
  for (int o = tid; o < SL::total(L); o += nthreads) sw[o] = 0.0f;
  __syncthreads();
  int goff = 0;
  for (int l = 0; l < L; ++l) {
    const int cin = l == 0 ? s.in_features : H;
    const int cinp = l == 0 ? SL::kIn0 : HP;
    float* base = sw + SL::layer_off(l);
    float* vec = base + cinp * HP;
    for (int o = tid; o < 3 * H; o += nthreads) vec[(o / H) * HP + (o % H)] = g[goff + o];
    goff += 3 * H;
    for (int o = tid; o < H * cin; o += nthreads) {
      const int c = o / cin, k = o - c * cin;
      base[k * HP + c] = g[goff + o];
    }
    goff += H * cin;
  }
  float* l1 = sw + SL::lin1_off(L);
  for (int o = tid; o < H * H; o += nthreads) {
    const int c = o / H, k = o - c * H;
    l1[k * HP + c] = g[goff + o];
  }
  goff += H * H;
  for (int o = tid; o < H; o += nthreads) l1[HP * HP + o] = g[goff + o];
  goff += H;
  float* l2 = sw + SL::lin2_off(L);
  for (int o = tid; o < 9 * H; o += nthreads) {
    const int a = o / H, k = o - a * H;
    l2[k * SL::kW2Cols + a] = g[goff + o];
  }
  goff += 9 * H;
  for (int o = tid; o < 9; o += nthreads) l2[HP * SL::kW2Cols + o] = g[goff + o];
}

struct StackSmem {
  int w, st, red, h, asrc, inl, wt, kv, total;
};
template <int HP>
__host__ __device__ inline StackSmem stack_smem(int L, int n, int k, int maxdeg, int graph_mode) {
  StackSmem s;
  int off = 0;
  s.w = off;    off = tile_align16(off + StackLayout<HP>::total(L) * 4);
  s.st = off;   off = tile_align16(off + 2 * kTileThreads * 16);
  s.red = off;  off = tile_align16(off + kTileThreads * 4);
  s.h = off;    off = tile_align16(off + kTileThreads * (HP + 4) * 4);
  s.asrc = off; off = tile_align16(off + kTileThreads * 4);
  s.inl = off;  off = tile_align16(off + maxdeg * kTileThreads);
  s.wt = off;   off = tile_align16(off + maxdeg * kTileThreads * 4);
  s.kv = off;   off = tile_align16(off + (graph_mode == SWARM_GRAPH_KNN ? kTileThreads * 4 : 0));
  s.total = off;
  return s;
}

// ROLLOUT = false: one forward (swarm_gatstack_forward).  ROLLOUT = true: the evaluation loop simulator.py:59-93 for a
// stacked network -- p.ticks x [forward -> argmax -> world step -> reward] with the agent's state, running return and
// Flocking memory in registers, exactly the tick of tile_kernels.cu's rollout (same world-step and reward device code,
// so it equals forward + swarm_sim_step + swarm_scenario_reward composed by hand bit for bit).
template <int HP, bool ROLLOUT>
__global__ void __launch_bounds__(kTileThreads) gatstack_kernel(const __grid_constant__ StackParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  using SL = StackLayout<HP>;
  constexpr int kRow = HP + 4;
  const SwarmConfig& c = p.cfg;
  const SwarmStackSpec& sp = p.spec;
  const int T = kTileThreads, N = c.n_agents, K = c.knn_k, L = sp.n_layers;
  const TileThread t = tile_thread(N, p.epb, c.num_envs);
  const int tid = t.tid;
  const StackSmem S = stack_smem<HP>(L, N, K, p.maxdeg, c.graph_mode);
  float* sw = reinterpret_cast<float*>(smem + S.w);
  float4* sst = reinterpret_cast<float4*>(smem + S.st);
  float* sh = reinterpret_cast<float*>(smem + S.h);
  TileGraphSmem g = {};
  g.sas = reinterpret_cast<float*>(smem + S.asrc);
  g.sin = smem + S.inl;
  g.skv = reinterpret_cast<float*>(smem + S.kv);

  float* sred = reinterpret_cast<float*>(smem + S.red);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  float ret0 = 0.0f, ret = 0.0f;
  float2 shp = make_float2(0.0f, 0.0f);
  int myhits = 0;
  if (t.active) {
    s = ROLLOUT ? p.state_rw[t.gidx] : p.state[t.gidx];
    if (ROLLOUT && p.returns) ret0 = p.returns[t.gidx];
    if (ROLLOUT && p.use_flock == 1) shp = p.shaping[t.gidx];
  }
  stack_stage_weights<HP>(sp, p.weights, sw, tid, T);
  uint64_t cache_rank = ~0ull, cache_nbr = 0;
  int deg = 0;
  const int n_ticks = ROLLOUT ? p.ticks : 1;
  for (int tick = 0; tick < n_ticks; ++tick) {
  const float4* pos = sst + (tick & 1) * T;
  sst[(tick & 1) * T + tid] = s;
  __syncthreads();

  // ---- graph: in-edge list of this node in edge-list order, once for all layers ------------------------------------
  if (c.graph_mode == SWARM_GRAPH_KNN) {
    const uint64_t nbr = tile_knn_small(t, pos, s, N, K, cache_rank, cache_nbr);
    deg = tile_in_edges_knn_small(g, t, N, K, nbr, reinterpret_cast<uint32_t*>(g.skv));
  } else if (c.graph_mode == SWARM_GRAPH_RADIUS) {
    if (t.active) deg = tile_in_edges_radius(g, t, pos, s, N, p.qmax_r);
  } else if (t.active && tick == 0) {
    deg = tile_in_edges_complete(g, t, N);
  }

  // ---- node features (train:95-99) ------------------------------------------------------------------------------
  float x[HP];
#pragma unroll
  for (int k = 0; k < HP; ++k) x[k] = 0.0f;
  x[0] = s.x; x[1] = s.y; x[2] = s.z; x[3] = s.w;
  if (sp.in_features == 7) {
    x[4] = c.goal_x; x[5] = c.goal_y; x[6] = (float)t.i;
  } else {
    x[4] = (float)t.i;
  }

  const uint8_t* sin = g.sin + tid;
  float* swt = reinterpret_cast<float*>(smem + S.wt) + tid;
  for (int l = 0; l < L; ++l) {
    const float* base = sw + SL::layer_off(l);
    const int cinp = l == 0 ? SL::kIn0 : HP;
    const float* vec = base + cinp * HP;
    // h = x W^T (sequential-k FFMA chains), alpha terms
    float h[HP];
#pragma unroll
    for (int cc = 0; cc < HP; ++cc) h[cc] = 0.0f;
    // (float4 weight loads + packed FFMA2: per element the same sequential-k chain of IEEE fused multiply-adds)
    if (l == 0) {
#pragma unroll
      for (int k = 0; k < SL::kIn0; ++k) {
#pragma unroll
        for (int cc = 0; cc < HP; cc += 4)
          fma4_packed(x[k], *reinterpret_cast<const float4*>(base + k * HP + cc), h[cc], h[cc + 1], h[cc + 2], h[cc + 3]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < HP; ++k) {
#pragma unroll
        for (int cc = 0; cc < HP; cc += 4)
          fma4_packed(x[k], *reinterpret_cast<const float4*>(base + k * HP + cc), h[cc], h[cc + 1], h[cc + 2], h[cc + 3]);
      }
    }
    float asrc = 0.0f, adst = 0.0f;
#pragma unroll
    for (int cc = 0; cc < HP; ++cc) {
      asrc = __fadd_rn(asrc, __fmul_rn(h[cc], vec[cc]));
      adst = __fadd_rn(adst, __fmul_rn(h[cc], vec[HP + cc]));
    }
    if (l > 0) __syncthreads();                       // the previous layer's rows are no longer read
#pragma unroll
    for (int cc = 0; cc < HP; cc += 4)
      *reinterpret_cast<float4*>(sh + tid * kRow + cc) = make_float4(h[cc], h[cc + 1], h[cc + 2], h[cc + 3]);
    g.sas[tid] = asrc;
    __syncthreads();
    // softmax over the in-edges (max, denominator, messages), sources read from the env's rows
    const float* rows = sh + t.envbase * kRow;
    const float* as_env = g.sas + t.envbase;
    float m = -INFINITY;
    for (int e = 0; e < deg; ++e) m = fmaxf(m, gat_logit(as_env[sin[e * T]], adst));
    // unnormalised weights once (kept in the per-thread scratch column), one reciprocal, messages as fused
    // multiply-adds in edge-list order
    float den = 0.0f;
    for (int e = 0; e < deg; ++e) {
      const float we = expf(__fsub_rn(gat_logit(as_env[sin[e * T]], adst), m));
      swt[e * T] = we;
      den = __fadd_rn(den, we);
    }
    const float inv = __fdiv_rn(1.0f, __fadd_rn(den, 1e-16f));
#pragma unroll
    for (int cc = 0; cc < HP; ++cc) x[cc] = 0.0f;
    for (int e = 0; e < deg; ++e) {
      const float a = __fmul_rn(swt[e * T], inv);
      const float* hj = rows + sin[e * T] * kRow;
#pragma unroll
      for (int cc = 0; cc < HP; cc += 4)
        fma4_packed(a, *reinterpret_cast<const float4*>(hj + cc), x[cc], x[cc + 1], x[cc + 2], x[cc + 3]);
    }
    const bool relu = sp.activation[l] == 1;
#pragma unroll
    for (int cc = 0; cc < HP; ++cc) {
      const float v = __fadd_rn(x[cc], vec[2 * HP + cc]);
      x[cc] = relu ? fmaxf(v, 0.0f) : tanhf(v);       // padded channels: weights and bias are 0, tanh(0) = relu(0) = 0
    }
  }

  // ---- head: lin1 -> ReLU -> lin2, argmax (first maximum wins) ----------------------------------------------------
  const float* l1 = sw + SL::lin1_off(L);
  float r[HP];
#pragma unroll
  for (int cc = 0; cc < HP; ++cc) r[cc] = 0.0f;
#pragma unroll
  for (int k = 0; k < HP; ++k) {
#pragma unroll
    for (int cc = 0; cc < HP; cc += 4)
      fma4_packed(x[k], *reinterpret_cast<const float4*>(l1 + k * HP + cc), r[cc], r[cc + 1], r[cc + 2], r[cc + 3]);
  }
#pragma unroll
  for (int cc = 0; cc < HP; ++cc) r[cc] = fmaxf(__fadd_rn(r[cc], l1[HP * HP + cc]), 0.0f);
  const float* l2 = sw + SL::lin2_off(L);
  float q[SL::kW2Cols];                                  // 9 actions + 3 zero-padded columns
#pragma unroll
  for (int a = 0; a < SL::kW2Cols; ++a) q[a] = 0.0f;
#pragma unroll
  for (int k = 0; k < HP; ++k) {
#pragma unroll
    for (int a = 0; a < SL::kW2Cols; a += 4)
      fma4_packed(r[k], *reinterpret_cast<const float4*>(l2 + k * SL::kW2Cols + a), q[a], q[a + 1], q[a + 2], q[a + 3]);
  }
  float best = 0.0f;
  int action = 0;
#pragma unroll
  for (int a = 0; a < 9; ++a) {
    q[a] = __fadd_rn(q[a], l2[HP * SL::kW2Cols + a]);
    if (a == 0 || q[a] > best) { best = q[a]; action = a; }
  }
  if (!ROLLOUT) {
    if (t.active) {
      if (p.q_out) {
#pragma unroll
        for (int a = 0; a < 9; ++a) p.q_out[t.gidx * 9 + a] = q[a];
      }
      if (p.act_out) p.act_out[t.gidx] = action;
    }
  } else {
    // ---- world step + reward: the tick of tile_kernels.cu ------------------------------------------------------------
    float reward = 0.0f;
    if (t.active) {
      uint8_t flags = 0;
      uint32_t cmask = 0;
      tile_world_step(p.step, t, pos, action, s, flags, cmask);
      const float dgoal = goal_distance(s.x, s.y, c);
      if (c.scenario == SWARM_SCENARIO_OBSTACLE_AVOIDANCE) {
        const float dobs = obstacle_distance(s.x, s.y, c);
        reward = oa_reward(dgoal, dobs, c, flags);
        myhits += (flags & SWARM_FLAG_HIT) ? 1 : 0;
      } else if (!p.use_flock) {
        sred[tid] = dgoal;     // GoTo's collective reward needs every agent's distance: finished below
      }
    }
    if (p.use_flock == 2) {
      // Cohesion (cohesion:66-85): smallest / largest surface distance to the other agents at their POST-step positions;
      // the op sequence of reward_kernels.cu (extreme squared distances, two exact square roots)
      const SwarmRewardSpec& fs = p.flock;
      float4* post = sst + ((tick + 1) & 1) * T;
      post[tid] = s;
      __syncthreads();
      if (t.active) {
        const float4* others = post + t.envbase;
        float lo = INFINITY, hi = 0.0f;
#pragma unroll 4
        for (int j = 0; j < N; ++j) {
          const float2 q = xy_of(others[j]);
          const float dx = __fsub_rn(s.x, q.x), dy = __fsub_rn(s.y, q.y);
          const float d2 = __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
          lo = fminf(lo, j == t.i ? INFINITY : d2);
          hi = fmaxf(hi, d2);
        }
        const float mn = __fsub_rn(__fsub_rn(__fsqrt_rn(lo), fs.agent_radius), fs.agent_radius);
        const float mx = __fsub_rn(__fsub_rn(__fsqrt_rn(hi), fs.agent_radius), fs.agent_radius);
        const float collision = mn > fs.sigma ? 0.0f : expf(-__fdiv_rn(mn, fs.sigma));    // cohesion:79-80
        const float cohesion = mn < fs.sigma ? 0.0f : -__fsub_rn(mx, fs.sigma);           // cohesion:82-83
        reward = __fadd_rn(collision, cohesion);
      }
    } else if (p.use_flock) {
      // Flocking (flocking:124-171): every agent's term from the POST-step positions of its env, then the collective sum
      // in agent order (the op sequence of reward_kernels.cu / the FLOCK tile kernel)
      const SwarmRewardSpec& fs = p.flock;
      float4* post = sst + ((tick + 1) & 1) * T;
      post[tid] = s;
      __syncthreads();
      float term = 0.0f;
      if (t.active) {
        const float d_goal = norm2(__fsub_rn(s.x, fs.goal_x), __fsub_rn(s.y, fs.goal_y));
        const float shaped_goal = __fmul_rn(d_goal, fs.pos_shaping);
        const float4* others = post + t.envbase;
        float sum;
        int close;
        flocking_partner_sweep(s.x, s.y, t.i, N, [&](int j) { return xy_of(others[j]); }, fs.desired_distance,
                               fs.agent_radius, fs.min_collision_distance, sum, close);
        const float spacing = __fmul_rn(__fdiv_rn(sum, (float)(N - 1)), fs.dist_shaping);
        const float pos_rew = __fsub_rn(shp.x, shaped_goal);
        float r = pos_rew;
        if (d_goal < fs.goal_radius) r = __fadd_rn(r, fs.on_goal_bonus);
        const float avoid = close ? __fmul_rn((float)close, fs.collision_reward) : 0.0f;
        const float dist_rew = __fsub_rn(shp.y, spacing);
        shp = make_float2(shaped_goal, spacing);
        term = __fadd_rn(__fadd_rn(r, avoid), dist_rew);
      }
      sred[tid] = term;
      __syncthreads();
      if (t.active)
        for (int a = 0; a < N; ++a) reward = __fadd_rn(reward, sred[t.envbase + a]);
    } else if (c.scenario == SWARM_SCENARIO_GOTO) {
      // collective reward (go_to:108-115): 0 + (-d_0) + (-d_1) + ... in agent order, same for all agents
      __syncthreads();
      if (t.active)
        for (int a = 0; a < N; ++a) reward = __fadd_rn(reward, -sred[t.envbase + a]);
    }
    ret = __fadd_rn(ret, reward);
  }
  }   // tick loop

  if (ROLLOUT) {
    if (t.active) {
      p.state_rw[t.gidx] = s;
      if (p.returns) p.returns[t.gidx] = __fadd_rn(ret0, ret);
      if (p.use_flock == 1) p.shaping[t.gidx] = shp;
    }
    if (p.hits) {
      __syncthreads();
      reinterpret_cast<int*>(sred)[tid] = myhits;
      __syncthreads();
      if (t.active && t.i == 0) {
        int tot = 0;
        for (int a = 0; a < N; ++a) tot += reinterpret_cast<int*>(sred)[t.envbase + a];
        p.hits[t.env] += tot;
      }
    }
  }
}

template <int HP>
static cudaError_t launch_stack_hp(const StackParams& p, cudaStream_t stream) {
  const SwarmConfig& c = p.cfg;
  const StackSmem S = stack_smem<HP>(p.spec.n_layers, c.n_agents, c.knn_k, p.maxdeg, c.graph_mode);
  if (S.total > 227 * 1024) return cudaErrorInvalidConfiguration;
  if (S.total > 48 * 1024) {
    cudaError_t err = p.ticks > 0
        ? cudaFuncSetAttribute(gatstack_kernel<HP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S.total)
        : cudaFuncSetAttribute(gatstack_kernel<HP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S.total);
    if (err != cudaSuccess) return err;
  }
  const int grid = (c.num_envs + p.epb - 1) / p.epb;
  if (p.ticks > 0) gatstack_kernel<HP, true><<<grid, kTileThreads, S.total, stream>>>(p);
  else gatstack_kernel<HP, false><<<grid, kTileThreads, S.total, stream>>>(p);
  return cudaGetLastError();
}

int stack_weight_count_host(const SwarmStackSpec& s) { return stack_weight_count(s); }

static cudaError_t launch_stack(const StackParams& p, cudaStream_t stream) {
  if (p.spec.hidden <= 8) return launch_stack_hp<8>(p, stream);
  if (p.spec.hidden <= 16) return launch_stack_hp<16>(p, stream);
  return launch_stack_hp<32>(p, stream);
}

static void stack_fill(StackParams& p, const SwarmConfig& c, const SwarmStackSpec& spec, const float* weights) {
  std::memset(&p, 0, sizeof(p));
  p.cfg = c;
  p.spec = spec;
  p.weights = weights;
  p.epb = kTileThreads / c.n_agents;
  p.maxdeg = c.graph_mode == SWARM_GRAPH_KNN ? (c.n_agents + c.knn_k + 1) : c.n_agents;
  p.qmax_r = c.graph_mode == SWARM_GRAPH_RADIUS ? sq_threshold(c.graph_radius) : 0.0f;
}

// the fused evaluation loop; `step` carries the physics constants (api.cu fill_params), flock == nullptr: world reward
cudaError_t launch_gatstack_rollout(const SwarmConfig& c, const SwarmStackSpec& spec, const float* weights, float* state,
                                    int ticks, const TileParams& step, const SwarmRewardSpec* flock, float* shaping,
                                    float* returns, int32_t* hits, cudaStream_t stream) {
  if (ticks <= 0) return cudaSuccess;
  StackParams p;
  stack_fill(p, c, spec, weights);
  p.step = step;
  p.state_rw = reinterpret_cast<float4*>(state);
  p.returns = returns;
  p.hits = hits;
  p.ticks = ticks;
  if (flock) {
    p.flock = *flock;
    p.shaping = reinterpret_cast<float2*>(shaping);
    p.use_flock = flock->kind == SWARM_REWARD_COHESION ? 2 : 1;
  }
  return launch_stack(p, stream);
}

cudaError_t launch_gatstack_forward(const SwarmConfig& c, const SwarmStackSpec& spec, const float* weights, const float* state,
                                    float* q, int32_t* actions, cudaStream_t stream) {
  StackParams p;
  std::memset(&p, 0, sizeof(p));
  p.cfg = c;
  p.spec = spec;
  p.weights = weights;
  p.state = reinterpret_cast<const float4*>(state);
  p.q_out = q;
  p.act_out = actions;
  p.epb = kTileThreads / c.n_agents;
  p.maxdeg = c.graph_mode == SWARM_GRAPH_KNN ? (c.n_agents + c.knn_k + 1) : c.n_agents;
  p.qmax_r = c.graph_mode == SWARM_GRAPH_RADIUS ? sq_threshold(c.graph_radius) : 0.0f;
  if (spec.hidden <= 8) return launch_stack_hp<8>(p, stream);
  if (spec.hidden <= 16) return launch_stack_hp<16>(p, stream);
  return launch_stack_hp<32>(p, stream);
}

// returns[b][i] += reward (per agent, or the env's collective reward); hits[b] += agents inside hit_distance this tick
__global__ void __launch_bounds__(256) stack_accumulate_kernel(long long total, int N, const float* __restrict__ rewards,
                                                               int per_env, const uint8_t* __restrict__ flags,
                                                               float* __restrict__ returns, int32_t* __restrict__ hits) {
  for (long long gI = (long long)blockIdx.x * blockDim.x + threadIdx.x; gI < total; gI += (long long)gridDim.x * blockDim.x) {
    const long long b = gI / N;
    if (returns) returns[gI] = __fadd_rn(returns[gI], per_env ? rewards[b] : rewards[gI]);
    if (hits && flags && (flags[gI] & SWARM_FLAG_HIT)) atomicAdd(&hits[b], 1);        // integer: order-independent
  }
}

cudaError_t launch_stack_accumulate(long long total, int N, const float* rewards, int per_env, const uint8_t* flags,
                                    float* returns, int32_t* hits, cudaStream_t stream) {
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  stack_accumulate_kernel<<<(int)(blocks < 1 ? 1 : blocks), 256, 0, stream>>>(total, N, rewards, per_env, flags, returns, hits);
  return cudaGetLastError();
}

}  // namespace swarm
