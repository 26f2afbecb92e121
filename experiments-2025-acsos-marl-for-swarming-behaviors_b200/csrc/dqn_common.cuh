// dqn_common.cuh -- parameter block shared by the two DQN gradient kernels (dqn_kernels.cu: CUDA-core parity path,
// dqn_tc_kernels.cu: tensor-core path).
#ifndef SWARM_DQN_COMMON_CUH
#define SWARM_DQN_COMMON_CUH

#include "tile_device.cuh"

namespace swarm {

constexpr int kPartialStride = 1680;   // 1673 gradients + [1673] = sum of squared TD errors, padded
constexpr int kXPad = 8;

struct DqnParams {
  SwarmConfig cfg;
  const float* w_online;
  const float* w_target;
  SwarmReplay batch;
  const int64_t* indices;
  int32_t n_graphs;
  float gamma;
  float loss_scale;
  float* partials;
  float* td;
  int32_t epb, maxdeg;
  float qmax_r;               // radius graph threshold (see TileParams)
  int32_t parallel;           // 1: target pass and online pass of a transition run side by side on two thread groups
  // device-driven tick (swarm_train_tick_grad): slots are drawn here from the ring fill after this tick's push
  const SwarmTrainCtl* ctl;
  int64_t* indices_out;       // [n_graphs] the drawn slots (exported for tests / logging)
  unsigned long long sample_seed;
  int32_t pushed_envs;        // transitions pushed by this tick's rollout
};

// ring fill after this tick's push, and whether it allows an update (train:113-115)
__device__ __forceinline__ long long train_ring_size(const SwarmTrainCtl* ctl, int pushed, long long capacity) {
  const long long size = ctl->ring_size + pushed;
  return size < capacity ? size : capacity;
}

}  // namespace swarm
#endif
