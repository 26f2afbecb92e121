// reset_kernels.cu -- generate_grid + reset_world_at (go_to_position_scenario.py:52-106,
// obstacle_avoidance_scenario.py:63-133): one thread per agent, float4 state writes.
#include "swarm_device.cuh"

namespace swarm {

__global__ void __launch_bounds__(256) reset_grid_kernel(SwarmConfig c, int cols, int rows,
                                                         const float2* __restrict__ centers,
                                                         float4* __restrict__ state) {
  const long long total = (long long)c.num_envs * c.n_agents;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const long long env = g / c.n_agents;
    const int i = (int)(g - env * c.n_agents);
    const float2 ctr = centers[env];
    const float2 p = grid_position(ctr.x, ctr.y, i, cols, rows, c.grid_spacing);
    state[g] = make_float4(p.x, p.y, 0.0f, 0.0f);
  }
}

cudaError_t launch_reset_grid(const SwarmConfig& c, int cols, int rows, const float* centers, float* state,
                              cudaStream_t stream) {
  const long long total = (long long)c.num_envs * c.n_agents;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  reset_grid_kernel<<<(int)blocks, 256, 0, stream>>>(c, cols, rows, reinterpret_cast<const float2*>(centers),
                                                     reinterpret_cast<float4*>(state));
  return cudaGetLastError();
}

}  // namespace swarm
