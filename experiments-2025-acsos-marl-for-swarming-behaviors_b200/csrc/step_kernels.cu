// step_kernels.cu -- the stand-alone world step (vmas Environment.step, train_gcn_dqn.py:169 / simulator.py:68)
// as a lean HBM-streaming kernel: 40 algorithmic bytes per agent-step (float4 state in, int32 action in, float4
// state out, float reward out) plus optional observation / distance / flag / mask outputs.
//
// 256 threads per CTA, thread = agent, floor(256/N) whole envs per CTA (N = 12: 252 of 256 lanes busy).  Partner
// positions are staged once in shared memory as float2; the pair loop is branch-light (the self pair is tested
// too: it yields a zero force and its mask bit is cleared afterwards) and unrolled for memory-level parallelism.
// Same arithmetic, in the same order, as the env-tile kernels (swarm_device.cuh), so results are bit-identical to
// MODE_STEP of tile_kernels.cu.
#include "tile_kernels.cuh"

namespace swarm {

constexpr int kStepThreads = 256;

struct StepParams {
  SwarmConfig cfg;
  const float4* state_in;
  const int32_t* actions;
  float4* state_out;
  float* rewards;
  uint8_t* flags;
  uint32_t* contact;
  float* obs;
  float2* dist;
  int32_t epb;
  float one_minus_drag, dmin_aa, dmin_ao, qmax_aa, qmax_ao;
};

__global__ void __launch_bounds__(kStepThreads) sim_step_kernel(const __grid_constant__ StepParams p) {
  __shared__ float2 spos[kStepThreads];
  __shared__ float sdg[kStepThreads];
  const SwarmConfig& c = p.cfg;
  const int N = c.n_agents;
  const int tid = threadIdx.x;
  const int el = tid / N;
  const int i = tid - el * N;
  const long long env = (long long)blockIdx.x * p.epb + el;
  const bool active = (el < p.epb) && (env < c.num_envs);
  const int envbase = el * N;
  const long long gidx = env * N + i;

  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  int action = 0;
  if (active) {
    s = p.state_in[gidx];
    action = p.actions[gidx];
  }
  spos[tid] = make_float2(s.x, s.y);
  __syncthreads();

  float reward = 0.0f;
  if (active) {
    float fx, fy, gx, gy;
    decode_action(action, fx, fy);
    uint8_t flags = 0;
    uint32_t cmask = 0;
    if (c.scenario == SWARM_SCENARIO_OBSTACLE_AVOIDANCE) {
      const float dx = __fsub_rn(s.x, c.obstacle_x), dy = __fsub_rn(s.y, c.obstacle_y);
      if (__fmaf_rn(dy, dy, __fmul_rn(dx, dx)) <= p.qmax_ao) {
        if (contact_force(s.x, s.y, c.obstacle_x, c.obstacle_y, p.dmin_ao, c.collision_force, c.contact_margin, gx, gy)) {
          fx = __fadd_rn(fx, gx);
          fy = __fadd_rn(fy, gy);
          flags |= SWARM_FLAG_OBSTACLE_CONTACT;
        }
      }
    }
    agent_contacts(spos + envbase, N, i, s.x, s.y, p.qmax_aa, p.dmin_aa, c.collision_force, c.contact_margin, fx, fy, cmask);
    integrate(s, fx, fy, c.dt, p.one_minus_drag);

    const float dgoal = goal_distance(s.x, s.y, c);
    float dobs = 0.0f;
    if (c.scenario == SWARM_SCENARIO_OBSTACLE_AVOIDANCE) {
      dobs = obstacle_distance(s.x, s.y, c);
      reward = oa_reward(dgoal, dobs, c, flags);
    } else {
      sdg[tid] = dgoal;
    }
    p.state_out[gidx] = s;
    if (p.flags) p.flags[gidx] = flags;
    if (p.contact) p.contact[gidx] = cmask;
    if (p.obs) {
      float2* o = reinterpret_cast<float2*>(p.obs + gidx * 6);
      o[0] = make_float2(s.x, s.y);
      o[1] = make_float2(s.z, s.w);
      o[2] = make_float2(c.goal_x, c.goal_y);
    }
    if (p.dist) p.dist[gidx] = make_float2(dgoal, dobs);
  }
  if (c.scenario == SWARM_SCENARIO_GOTO) {
    // collective reward (go_to:108-115): 0 + (-d_0) + (-d_1) + ... in agent order, same value for every agent
    __syncthreads();
    if (active)
      for (int a = 0; a < N; ++a) reward = __fadd_rn(reward, -sdg[envbase + a]);
  }
  if (active && p.rewards) p.rewards[gidx] = reward;
}

cudaError_t launch_sim_step(const TileParams& tp, cudaStream_t stream) {
  StepParams p;
  p.cfg = tp.cfg;
  p.state_in = reinterpret_cast<const float4*>(tp.state_in);
  p.actions = tp.actions_in;
  p.state_out = reinterpret_cast<float4*>(tp.state_out);
  p.rewards = tp.rewards_out;
  p.flags = tp.flags_out;
  p.contact = tp.contact_out;
  p.obs = tp.obs_out;
  p.dist = reinterpret_cast<float2*>(tp.dist_out);
  p.epb = kStepThreads / tp.cfg.n_agents;
  p.one_minus_drag = tp.one_minus_drag;
  p.dmin_aa = tp.dmin_aa;
  p.dmin_ao = tp.dmin_ao;
  p.qmax_aa = tp.qmax_aa;
  p.qmax_ao = tp.qmax_ao;
  const long long grid = ((long long)tp.cfg.num_envs + p.epb - 1) / p.epb;
  sim_step_kernel<<<(unsigned)grid, kStepThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace swarm
