// Host build of csrc/knn_select.h for the CPU test-suite: runs the device selection routine on the CPU
// so it can be compared with torch.topk without a GPU.
#include <cstdint>
#include <vector>

#include "knn_select.h"
#include "knn_small.h"

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

// The register-resident path of small swarms (csrc/knn_small.h), n <= 16, non-negative values or NaN.
// mode 0: as the kernel runs it (rank answer when no tie can matter, emulation on the rank pattern otherwise);
// mode 1: the step-by-step emulation on the rank pattern for every row; mode 2: the word-level emulation
// (knn_small_topk_fast) for every row.  Returns the number of rows that took the emulation.
extern "C" int swarm_host_topk_small(const float* values, int rows, int n, int k, int mode, int32_t* out_idx) {
  int emulated = 0;
  for (int r = 0; r < rows; ++r) {
    uint32_t u[16];
    int rk[16];
    for (int j = 0; j < 16; ++j) u[j] = j < n ? swarm::knn_key_nonneg(values[(long long)r * n + j]) : swarm::kKnnPadKey;
    uint32_t present;
    const uint64_t rank = swarm::knn_small_ranks<16>(u, rk, present);
    uint64_t w;
    if (mode == 0 && swarm::knn_small_tie_free(present, n, k)) {
      w = swarm::knn_small_by_rank<16>(rk, k);
    } else {
      w = (mode == 2) ? swarm::knn_small_topk_fast(rank, n, k) : swarm::knn_small_topk(rank, n, k);
      ++emulated;
    }
    for (int j = 0; j < k; ++j) out_idx[(long long)r * k + j] = (int32_t)((w >> (4 * j)) & 15u);
  }
  return emulated;
}
