// tile_kernels.cu -- the env-tile kernel family for swarms of N <= 128 agents.
//
// One thread owns one agent; a CTA of 128 threads owns floor(128 / N) whole envs, so everything an
// agent needs from its swarm (partner positions, projected features, attention terms, kNN rows) is
// exchanged through shared memory and no env ever straddles a CTA.  The same device code is
// instantiated in four modes:
//   MODE_ROLLOUT  T x [graph -> GAT-Q -> argmax -> world step], state resident in registers
//                 (simulator.py:59-93 / train_gcn_dqn.py:153-178 inner loops)
//   MODE_FORWARD  graph -> GAT-Q (-> argmax) once                         (train_gcn_dqn.py:59-70)
//   MODE_STEP     one world step with given actions                       (vmas Environment.step)
//   MODE_GRAPH    edge list / neighbour table export                      (train:94-110, simulator:9-26)
// so the stand-alone kernels and the fused rollout are consistent by construction.
#include "gatq_device.cuh"
#include "knn_select.h"

namespace swarm {

namespace {

struct SmemPairs {
  float* v;
  uint8_t* x;
  int stride;
  __device__ __forceinline__ KnnPair get(int j) const {
    KnnPair p;
    p.v = v[j * stride];
    p.i = x[j * stride];
    return p;
  }
  __device__ __forceinline__ void set(int j, const KnnPair& p) {
    v[j * stride] = p.v;
    x[j * stride] = (uint8_t)p.i;
  }
};

}  // namespace

template <int MODE>
__global__ void __launch_bounds__(kTileThreads) tile_kernel(const __grid_constant__ TileParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr bool kQ = (MODE == MODE_ROLLOUT || MODE == MODE_FORWARD);
  constexpr bool kStep = (MODE == MODE_ROLLOUT || MODE == MODE_STEP);
  constexpr bool kGraphOut = (MODE == MODE_GRAPH || MODE == MODE_ROLLOUT);

  const SwarmConfig& c = p.cfg;
  const int T = kTileThreads;
  const int N = c.n_agents;
  const int K = c.knn_k;
  const bool knn = (c.graph_mode == SWARM_GRAPH_KNN) && (kQ || MODE == MODE_GRAPH);
  const int tid = threadIdx.x;
  const int el = tid / N;
  const int i = tid - el * N;
  const long long env = (long long)blockIdx.x * p.epb + el;
  const bool active = (el < p.epb) && (env < c.num_envs);
  const int envbase = el * N;
  const long long gidx = env * N + i;
  const long long BN = (long long)c.num_envs * N;

  const TileLayout L = tile_layout(MODE, T, N, K, p.maxdeg, c.graph_mode);
  float* sw = reinterpret_cast<float*>(smem + L.w);
  float4* sst = reinterpret_cast<float4*>(smem + L.st);
  float* sh = reinterpret_cast<float*>(smem + L.h);
  float* sas = reinterpret_cast<float*>(smem + L.asrc);
  float* swt = reinterpret_cast<float*>(smem + L.wt);
  uint8_t* sin = smem + L.inl;
  uint8_t* sdeg = smem + L.deg;
  float* skv = reinterpret_cast<float*>(smem + L.kv);
  uint8_t* ski = smem + L.ki;
  uint8_t* snbr = smem + L.nbr;
  float* sred = reinterpret_cast<float*>(smem + L.red);

  if (kQ) stage_weights(p.weights, sw, tid, T);

  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) s = reinterpret_cast<const float4*>(p.state_in)[gidx];

  int deg = 0;
  if (kQ && !knn && active) {
    // complete graph (train:101-108): sources into node d in edge-list order are 0..N-1 without d;
    // node 0 additionally receives the final (0,0) self loop.
    for (int j = 0; j < N; ++j)
      if (j != i) sin[(deg++) * T + tid] = (uint8_t)j;
    if (i == 0) sin[(deg++) * T + tid] = 0;
  }

  float ret = 0.0f;
  int myhits = 0;

  for (int tick = 0; tick < p.ticks; ++tick) {
    const int buf = tick & 1;
    sst[buf * T + tid] = s;
    __syncthreads();

    float q[9];
    int action = 0;

    // ------------------------------------------------------------------ graph (kNN) ----------
    if (knn) {
      if (active) {
        // simulator.py:17-19: distance_to_i = ||x[:, :2] - x[i, :2]||, topk(k, largest=False)
        SmemPairs row{skv + tid, ski + tid, T};
        for (int j = 0; j < N; ++j) {
          const float4 o = sst[buf * T + envbase + j];
          KnnPair pr;
          pr.v = norm2(__fsub_rn(o.x, s.x), __fsub_rn(o.y, s.y));
          pr.i = j;
          row.set(j, pr);
        }
        knn_topk_smallest(row, N, K);
        for (int r = 0; r < K; ++r) snbr[r * T + tid] = ski[r * T + tid];
      }
      __syncthreads();
      if (kQ && active) {
        // in-edges of node d = i in edge-list order: for each row ii, slot r with a = topk[ii][r]:
        // edge (ii -> a) then edge (a -> ii); finally (0 -> 0).
        deg = 0;
        for (int ii = 0; ii < N; ++ii) {
          for (int r = 0; r < K; ++r) {
            const int a = snbr[r * T + envbase + ii];
            if (a == i) sin[(deg++) * T + tid] = (uint8_t)ii;
            if (ii == i) sin[(deg++) * T + tid] = (uint8_t)a;
          }
        }
        if (i == 0) sin[(deg++) * T + tid] = 0;
      }
    }

    if (kGraphOut) {
      int32_t* eout = nullptr;
      if (MODE == MODE_GRAPH) eout = p.edges_out;
      else if (p.trace.edges) eout = p.trace.edges + (long long)tick * c.num_envs * 2 * p.edges_per_env;
      if (eout && active) {
        const int E = p.edges_per_env;
        int32_t* r0 = eout + env * 2 * E;
        int32_t* r1 = r0 + E;
        if (knn) {
          for (int r = 0; r < K; ++r) {
            const int a = snbr[r * T + tid];
            const int e = (i * K + r) * 2;
            r0[e] = i; r1[e] = a;
            r0[e + 1] = a; r1[e + 1] = i;
          }
        } else {
          for (int j = i + 1; j < N; ++j) {
            const int e = 2 * (i * N - (i * (i + 1)) / 2 + (j - i - 1));
            r0[e] = i; r1[e] = j;
            r0[e + 1] = j; r1[e + 1] = i;
          }
        }
        if (i == 0) { r0[E - 1] = 0; r1[E - 1] = 0; }
      }
      if (MODE == MODE_GRAPH && knn && p.nbr_out && active)
        for (int r = 0; r < K; ++r) p.nbr_out[gidx * K + r] = snbr[r * T + tid];
    }

    // ------------------------------------------------------------------ GAT-Q forward --------
    if (kQ) {
      float tdst = 0.0f;
      if (active) {
        // node features (train:95-99): [pos, vel, goal, agent id]
        const float x[7] = {s.x, s.y, s.z, s.w, c.goal_x, c.goal_y, (float)i};
        float h[32];
        float asrc;
        gat_project(x, sw, h, asrc, tdst);
        float4* hrow = reinterpret_cast<float4*>(sh + tid * kHPad);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) hrow[c4] = make_float4(h[4 * c4], h[4 * c4 + 1], h[4 * c4 + 2], h[4 * c4 + 3]);
        sas[tid] = asrc;
      }
      __syncthreads();

      if (active) {
        // edge softmax over the in-edges of this node, in edge-list order (torch_geometric.utils.softmax):
        // max, exp(z - max), sum + 1e-16, divide
        float m = -INFINITY;
        for (int e = 0; e < deg; ++e) {
          const float z = gat_logit(sas[envbase + sin[e * T + tid]], tdst);
          swt[e * T + tid] = z;
          m = fmaxf(m, z);
        }
        float den = 0.0f;
        for (int e = 0; e < deg; ++e) {
          const float w = expf(__fsub_rn(swt[e * T + tid], m));
          swt[e * T + tid] = w;
          den = __fadd_rn(den, w);
        }
        den = __fadd_rn(den, 1e-16f);
        float a1[32];
#pragma unroll
        for (int cc = 0; cc < 32; ++cc) a1[cc] = 0.0f;
        for (int e = 0; e < deg; ++e) {
          const int j = sin[e * T + tid];
          const float alpha = __fdiv_rn(swt[e * T + tid], den);
          gat_accumulate(a1, alpha, reinterpret_cast<const float4*>(sh + (envbase + j) * kHPad));
        }
        action = gat_head(a1, sw, q);
        if (MODE == MODE_FORWARD) {
          if (p.q_out) {
#pragma unroll
            for (int a = 0; a < 9; ++a) p.q_out[gidx * 9 + a] = q[a];
          }
          if (p.act_out) p.act_out[gidx] = action;
        } else if (p.trace.q) {
          float* tq = p.trace.q + ((long long)tick * BN + gidx) * 9;
#pragma unroll
          for (int a = 0; a < 9; ++a) tq[a] = q[a];
        }
      }
    }

    // ------------------------------------------------------------------ world step ------------
    if (kStep) {
      if (active) {
        if (MODE == MODE_STEP) {
          action = p.actions_in[gidx];
        } else if (p.actions_in) {
          const int fa = p.actions_in[(long long)tick * BN + gidx];
          if (fa >= 0) action = fa;
        }
        float fx, fy;
        decode_action(action, fx, fy);          // F = 0 + u
        uint8_t flags = 0;
        uint32_t cmask = 0;
        float gx, gy;
        if (c.scenario == SWARM_SCENARIO_OBSTACLE_AVOIDANCE) {
          // vmas entity order: the obstacle landmark precedes the agents
          const float dx = __fsub_rn(s.x, c.obstacle_x), dy = __fsub_rn(s.y, c.obstacle_y);
          if (__fmaf_rn(dy, dy, __fmul_rn(dx, dx)) <= p.qmax_ao) {
            if (contact_force(s.x, s.y, c.obstacle_x, c.obstacle_y, p.dmin_ao, c.collision_force,
                              c.contact_margin, gx, gy)) {
              fx = __fadd_rn(fx, gx);
              fy = __fadd_rn(fy, gy);
              flags |= SWARM_FLAG_OBSTACLE_CONTACT;
            }
          }
        }
        for (int j = 0; j < N; ++j) {
          if (j == i) continue;
          const float4 o = sst[buf * T + envbase + j];
          const float dx = __fsub_rn(s.x, o.x), dy = __fsub_rn(s.y, o.y);
          if (__fmaf_rn(dy, dy, __fmul_rn(dx, dx)) <= p.qmax_aa) {
            if (contact_force(s.x, s.y, o.x, o.y, p.dmin_aa, c.collision_force, c.contact_margin, gx, gy)) {
              fx = __fadd_rn(fx, gx);
              fy = __fadd_rn(fy, gy);
              if (j < 32) cmask |= (1u << j);
            }
          }
        }
        integrate(s, fx, fy, c.dt, p.one_minus_drag);

        // scenario.reward(agent) on the post-step state
        const float dgoal = goal_distance(s.x, s.y, c);
        float reward;
        float dobs = 0.0f;
        if (c.scenario == SWARM_SCENARIO_OBSTACLE_AVOIDANCE) {
          dobs = obstacle_distance(s.x, s.y, c);
          reward = oa_reward(dgoal, dobs, c, flags);
        } else {
          sred[tid] = dgoal;
          reward = 0.0f;
        }
        // (GoTo's collective reward needs every agent's distance: completed after the barrier below)
        if (c.scenario == SWARM_SCENARIO_OBSTACLE_AVOIDANCE) {
          ret = __fadd_rn(ret, reward);
          myhits += (flags & SWARM_FLAG_HIT) ? 1 : 0;
          if (MODE == MODE_STEP) {
            if (p.rewards_out) p.rewards_out[gidx] = reward;
          } else if (p.trace.rewards) {
            p.trace.rewards[(long long)tick * BN + gidx] = reward;
          }
        }
        if (MODE == MODE_STEP) {
          reinterpret_cast<float4*>(p.state_out)[gidx] = s;
          if (p.flags_out) p.flags_out[gidx] = flags;
          if (p.contact_out) p.contact_out[gidx] = cmask;
          if (p.obs_out) {
            float* o = p.obs_out + gidx * 6;
            o[0] = s.x; o[1] = s.y; o[2] = s.z; o[3] = s.w; o[4] = c.goal_x; o[5] = c.goal_y;
          }
          if (p.dist_out) reinterpret_cast<float2*>(p.dist_out)[gidx] = make_float2(dgoal, dobs);
        } else {
          const long long tb = (long long)tick * BN + gidx;
          if (p.trace.state) reinterpret_cast<float4*>(p.trace.state)[tb] = s;
          if (p.trace.actions) p.trace.actions[tb] = action;
          if (p.trace.flags) p.trace.flags[tb] = flags;
          if (p.trace.contact) p.trace.contact[tb] = cmask;
          if (p.trace.dist) reinterpret_cast<float2*>(p.trace.dist)[tb] = make_float2(dgoal, dobs);
        }
      }
      if (c.scenario == SWARM_SCENARIO_GOTO) {
        // collective reward (go_to:108-115): 0 + (-d_0) + (-d_1) + ... in agent order, same for all agents
        __syncthreads();
        if (active) {
          float reward = 0.0f;
          for (int a = 0; a < N; ++a) reward = __fadd_rn(reward, -sred[envbase + a]);
          ret = __fadd_rn(ret, reward);
          if (MODE == MODE_STEP) {
            if (p.rewards_out) p.rewards_out[gidx] = reward;
          } else if (p.trace.rewards) {
            p.trace.rewards[(long long)tick * BN + gidx] = reward;
          }
        }
      }
    }
  }

  if (MODE == MODE_ROLLOUT) {
    if (active) {
      reinterpret_cast<float4*>(p.state_out)[gidx] = s;
      if (p.returns) p.returns[gidx] = __fadd_rn(p.returns[gidx], ret);
    }
    if (p.hits) {
      __syncthreads();
      reinterpret_cast<int*>(sred)[tid] = myhits;
      __syncthreads();
      if (active && i == 0) {
        int tot = 0;
        for (int a = 0; a < N; ++a) tot += reinterpret_cast<int*>(sred)[envbase + a];
        p.hits[env] += tot;
      }
    }
  }
}

// explicit instantiations + launcher
cudaError_t launch_tile(int mode, const TileParams& p, cudaStream_t stream) {
  const SwarmConfig& c = p.cfg;
  const TileLayout L = tile_layout(mode, kTileThreads, c.n_agents, c.knn_k, p.maxdeg, c.graph_mode);
  const int grid = (c.num_envs + p.epb - 1) / p.epb;
  cudaError_t err = cudaSuccess;
#define SWARM_LAUNCH(M)                                                                              \
  do {                                                                                               \
    if (L.total > 48 * 1024) {                                                                       \
      err = cudaFuncSetAttribute(tile_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total); \
      if (err != cudaSuccess) return err;                                                            \
    }                                                                                                \
    tile_kernel<M><<<grid, kTileThreads, L.total, stream>>>(p);                                      \
  } while (0)
  switch (mode) {
    case MODE_ROLLOUT: SWARM_LAUNCH(MODE_ROLLOUT); break;
    case MODE_FORWARD: SWARM_LAUNCH(MODE_FORWARD); break;
    case MODE_STEP: SWARM_LAUNCH(MODE_STEP); break;
    case MODE_GRAPH: SWARM_LAUNCH(MODE_GRAPH); break;
    default: return cudaErrorInvalidValue;
  }
#undef SWARM_LAUNCH
  return cudaGetLastError();
}

}  // namespace swarm
