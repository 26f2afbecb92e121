// knn_small.h -- register-resident tie-exact "k smallest" for swarms of n <= 16 agents.
//
// torch.topk on the CPU (src/simulation/simulator.py:18-19) runs comparison-only algorithms (knn_select.h), so its
// result is a function of the ORDER PATTERN of the row alone: replace every distance by its rank
// r_j = #{l : d_l < d_j} (ties share a rank, NaN ranks above everything like in torch's comparator) and the algorithms
// take exactly the same decisions.  For n <= 16 a rank fits a nibble, so a whole row is one 64-bit word and the row
// never touches memory:
//   * ranks come from the n (n - 1) / 2 pairwise comparisons of the distances held in registers;
//   * the set of ranks present tells whether any tie can influence the result: ranks 0 .. K present  <=>  the K
//     smallest are K singletons followed by a strictly larger value, and then the answer is "index of rank 0, 1, ...,
//     K - 1" whatever the algorithm (the common case away from the regular start grids);
//   * otherwise the step-for-step libstdc++ emulation of knn_select.h runs on the nibble-packed (permutation, rank)
//     words -- shifts and masks instead of shared-memory traffic.
// __host__ __device__, so tests/test_knn_select.py runs the same code on the CPU against torch.topk.
#ifndef SWARM_KNN_SMALL_H
#define SWARM_KNN_SMALL_H

#include <stdint.h>

#include "knn_select.h"

namespace swarm {

struct KnnRankPair {
  int v;   // rank
  int i;   // index
};

SWARM_HD bool knn_less(const KnnRankPair& a, const KnnRankPair& b) { return a.v < b.v; }

// row of n <= 16 elements: nibble p of `perm` = index of the element at position p, nibble i of `rank` = rank of index i
struct KnnNibRow {
  uint64_t perm;
  uint64_t rank;
  SWARM_HD KnnRankPair get(int p) const {
    KnnRankPair e;
    e.i = (int)((perm >> (4 * p)) & 15u);
    e.v = (int)((rank >> (4 * e.i)) & 15u);
    return e;
  }
  SWARM_HD void set(int p, const KnnRankPair& e) {
    perm = (perm & ~((uint64_t)15u << (4 * p))) | ((uint64_t)(uint32_t)e.i << (4 * p));
  }
};

// comparison key of a NON-NEGATIVE distance (a 2-norm): the bit pattern is monotone; every NaN compares equal to every
// other NaN and above +inf, like torch's comparator  (!isnan(x) && isnan(y)) || x < y
SWARM_HD uint32_t knn_key_nonneg(float d) {
#if defined(__CUDA_ARCH__)
  const uint32_t b = __float_as_uint(d);
#else
  union { float f; uint32_t u; } c;
  c.f = d;
  const uint32_t b = c.u;
#endif
  return (d != d) ? 0x7FC00000u : b;
}
constexpr uint32_t kKnnPadKey = 0x7FFFFFFFu;     // padding beyond n: above every real key

// Ranks of the NP keys (entries j >= n must hold kKnnPadKey).  Returns the rank word; `present` gets bit r set for
// every rank r taken by some entry.
template <int NP>
SWARM_HD uint64_t knn_small_ranks(const uint32_t (&u)[NP], int (&r)[NP], uint32_t& present) {
  static_assert(NP <= 16, "a rank must fit a nibble");
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < NP; ++j) r[j] = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int l = 0; l < NP; ++l) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = l + 1; j < NP; ++j) {
      // keys are below 2^31, so the sign bit of the 32-bit difference is the comparison (one subtract + one
      // shift-and-add per direction; the `?:` form compiles to select chains twice as long)
      r[j] += (int)((u[l] - u[j]) >> 31);
      r[l] += (int)((u[j] - u[l]) >> 31);
    }
  }
  uint32_t lo = 0, hi = 0, pres = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < NP; ++j) {
    pres |= 1u << r[j];
    if (j < 8) lo |= (uint32_t)r[j] << (4 * j);
    else hi |= (uint32_t)r[j] << (4 * (j - 8));
  }
  present = pres;
  return ((uint64_t)hi << 32) | lo;
}

// true when no tie can influence the k smallest of the n real entries
SWARM_HD bool knn_small_tie_free(uint32_t present, int n, int k) {
  const int need = (k + 1 < n) ? (k + 1) : n;
  const uint32_t full = (need >= 32) ? 0xFFFFFFFFu : ((1u << need) - 1u);
  return (present & full) == full;
}

// tie-free answer: nibble s of the result = index of the entry with rank s, s < k
template <int NP>
SWARM_HD uint64_t knn_small_by_rank(const int (&r)[NP], int k) {
  uint32_t lo = 0, hi = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < NP; ++j) {
    const int s = r[j];
    const bool take = s < k;
    lo |= (take && s < 8) ? ((uint32_t)j << (4 * (s & 7))) : 0u;
    hi |= (take && s >= 8) ? ((uint32_t)j << (4 * (s & 7))) : 0u;
  }
  return ((uint64_t)hi << 32) | lo;
}

// torch.topk(row, k, largest=False) for the order pattern `rank` of n <= 16 entries: nibble s < k of the result is the
// index torch puts at output slot s.
SWARM_HD uint64_t knn_small_topk(uint64_t rank, int n, int k) {
  KnnNibRow row;
  row.perm = 0xFEDCBA9876543210ull;
  row.rank = rank;
  // knn_topk_smallest with k * 64 > n: nth_element, then std::sort of k - 1 <= 15 entries, which below libstdc++'s
  // threshold of 16 is a plain insertion sort
  if (k <= 0) return row.perm;
  if (k - 1 != n) knn_introselect(row, 0, k - 1, n, knn_lg(n) * 2);
  knn_insertion_sort(row, 0, k - 1);
  return row.perm;
}

}  // namespace swarm
#endif  // SWARM_KNN_SMALL_H
