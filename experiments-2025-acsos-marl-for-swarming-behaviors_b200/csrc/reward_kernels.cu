// reward_kernels.cu -- the reward functions of the two remaining reference scenarios, evaluated on the state that
// swarm_sim_step leaves in HBM (their physics is the GoTo world: colliding sphere agents, no colliding landmark):
//   Flocking  flocking_scenario.py:93-176  collective reward = sum over agents of (shaped goal progress [+ 50 on the
//             goal] + (-1) per other agent closer than 0.005 surface to surface + shaped spacing progress); keeps the
//             two `previous_*` shaping terms per agent between ticks, initialised the way reset_world_at does
//   Cohesion  cohesion_scenario.py:66-85   per-agent reward from the smallest / largest surface distance to the others
// Thread = agent, a CTA owns floor(128 / N) whole envs: positions go through shared memory once, each thread sweeps its
// N - 1 partners in agent order (the order the reference's list comprehensions visit them), and -- Flocking -- the first
// thread of every env adds the N per-agent terms in agent order, as the reference's `+=` loop does.
// The reference evaluates Python `if`s on tensors (flocking:141,169; cohesion:80,83) and therefore only runs with one
// env; here every env is treated as its own copy of that one-env computation.
#include <cstdint>

#include "swarm_device.cuh"

namespace swarm {

constexpr int kRewardThreads = 128;

template <int KIND>
__global__ void __launch_bounds__(kRewardThreads) scenario_reward_kernel(SwarmRewardSpec sp,
                                                                         const float4* __restrict__ state,
                                                                         float2* __restrict__ shaping,
                                                                         float* __restrict__ reward,
                                                                         float4* __restrict__ terms) {
  __shared__ float2 spos[kRewardThreads];
  __shared__ float sterm[kRewardThreads];
  const int N = sp.n_agents;
  const int epb = kRewardThreads / N;
  const int tid = threadIdx.x;
  const int le = tid / N;                       // env within the CTA
  const int i = tid - le * N;                   // agent within the env
  for (long long tile = blockIdx.x; tile * epb < sp.num_envs; tile += gridDim.x) {
    const long long env = tile * epb + le;
    const bool live = le < epb && env < sp.num_envs;
    const bool selected = live && (sp.env_index < 0 || env == sp.env_index);
    float2 p = make_float2(0.0f, 0.0f);
    if (live) {
      const float4 s = state[env * N + i];
      p = make_float2(s.x, s.y);
    }
    spos[tid] = p;
    __syncthreads();
    const float2* others = spos + le * N;
    float term = 0.0f;
    if (selected) {
      const long long g = env * N + i;
      if (KIND == SWARM_REWARD_FLOCKING) {
        const float d_goal = norm2(__fsub_rn(p.x, sp.goal_x), __fsub_rn(p.y, sp.goal_y));
        const float shaped_goal = __fmul_rn(d_goal, sp.pos_shaping);
        // spacing term ((|p_i - p_j| - desired)^2 over j != i).mean() * dist_shaping (flocking:109-121,148-160); at a
        // reset the agents after i have not been placed yet and still sit at the origin world.reset put them at
        float sum = 0.0f;
        int close = 0;
        for (int j = 0; j < N; ++j) {
          if (j == i) continue;
          float2 q = others[j];
          if (sp.reset && j > i) q = make_float2(0.0f, 0.0f);
          const float d = norm2(__fsub_rn(p.x, q.x), __fsub_rn(p.y, q.y));
          const float e = __fsub_rn(d, sp.desired_distance);
          sum = __fadd_rn(sum, __fmul_rn(e, e));
          // world.get_distance: centre distance minus both radii (flocking:166)
          const float gap = __fsub_rn(__fsub_rn(d, sp.agent_radius), sp.agent_radius);
          close += gap <= sp.min_collision_distance ? 1 : 0;
        }
        const float spacing = __fmul_rn(__fdiv_rn(sum, (float)(N - 1)), sp.dist_shaping);
        if (sp.reset) {
          shaping[g] = make_float2(shaped_goal, spacing);                       // flocking:101-121
        } else {
          const float2 prev = shaping[g];
          const float pos_rew = __fsub_rn(prev.x, shaped_goal);                 // flocking:136-137
          float r = pos_rew;
          if (d_goal < sp.goal_radius) r = __fadd_rn(r, sp.on_goal_bonus);      // flocking:141-142
          const float avoid = close ? __fmul_rn((float)close, sp.collision_reward) : 0.0f;   // flocking:163-168
          const float dist_rew = __fsub_rn(prev.y, spacing);                    // flocking:158
          shaping[g] = make_float2(shaped_goal, spacing);
          term = __fadd_rn(__fadd_rn(r, avoid), dist_rew);                      // flocking:129
          if (terms) terms[g] = make_float4(pos_rew, avoid, dist_rew, d_goal);
        }
      } else {
        // cohesion:66-85: distances = get_distance(agent, other) for every other agent; min / max of them
        float mn = INFINITY, mx = -INFINITY;
        for (int j = 0; j < N; ++j) {
          if (j == i) continue;
          const float2 q = others[j];
          const float d = norm2(__fsub_rn(p.x, q.x), __fsub_rn(p.y, q.y));
          const float gap = __fsub_rn(__fsub_rn(d, sp.agent_radius), sp.agent_radius);
          mn = fminf(mn, gap);
          mx = fmaxf(mx, gap);
        }
        const float collision = mn > sp.sigma ? 0.0f : expf(-__fdiv_rn(mn, sp.sigma));    // cohesion:79-80
        const float cohesion = mn < sp.sigma ? 0.0f : -__fsub_rn(mx, sp.sigma);           // cohesion:82-83
        reward[g] = __fadd_rn(collision, cohesion);
        if (terms) terms[g] = make_float4(collision, cohesion, mn, mx);
      }
    }
    if (KIND == SWARM_REWARD_FLOCKING && !sp.reset) {
      sterm[tid] = term;
      __syncthreads();
      if (selected && i == 0) {
        float c = 0.0f;                                   // self.collective_reward = 0; += per agent (flocking:125-129)
        for (int j = 0; j < N; ++j) c = __fadd_rn(c, sterm[le * N + j]);
        reward[env] = c;
      }
    }
    __syncthreads();
  }
}

cudaError_t launch_scenario_reward(const SwarmRewardSpec& sp, const float* state, float* shaping, float* reward,
                                   float* terms, cudaStream_t stream) {
  const int epb = kRewardThreads / sp.n_agents;
  long long tiles = ((long long)sp.num_envs + epb - 1) / epb;
  if (tiles > 148 * 16) tiles = 148 * 16;
  const float4* s = reinterpret_cast<const float4*>(state);
  float2* sh = reinterpret_cast<float2*>(shaping);
  float4* t = reinterpret_cast<float4*>(terms);
  if (sp.kind == SWARM_REWARD_FLOCKING)
    scenario_reward_kernel<SWARM_REWARD_FLOCKING><<<(int)tiles, kRewardThreads, 0, stream>>>(sp, s, sh, reward, t);
  else
    scenario_reward_kernel<SWARM_REWARD_COHESION><<<(int)tiles, kRewardThreads, 0, stream>>>(sp, s, sh, reward, t);
  return cudaGetLastError();
}

}  // namespace swarm
