// Host build of csrc/knn_select.h for the CPU test-suite: runs the device selection routine on the CPU
// so it can be compared with torch.topk without a GPU.
#include <cstdint>
#include <vector>

#include "knn_select.h"

namespace {
struct HostPairs {
  float* v;
  int32_t* x;
  swarm::KnnPair get(int j) const { return swarm::KnnPair{v[j], x[j]}; }
  void set(int j, const swarm::KnnPair& p) { v[j] = p.v; x[j] = p.i; }
};
}  // namespace

extern "C" void swarm_host_topk_smallest(const float* values, int rows, int n, int k, int32_t* out_idx) {
  std::vector<float> v(n);
  std::vector<int32_t> x(n);
  for (int r = 0; r < rows; ++r) {
    for (int j = 0; j < n; ++j) {
      v[j] = values[(long long)r * n + j];
      x[j] = j;
    }
    HostPairs a{v.data(), x.data()};
    swarm::knn_topk_smallest(a, n, k);
    for (int j = 0; j < k; ++j) out_idx[(long long)r * k + j] = x[j];
  }
}
