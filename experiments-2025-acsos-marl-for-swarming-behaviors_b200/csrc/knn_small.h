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

// ---- the same algorithms, written for latency ----------------------------------------------------------------------
// knn_small_topk walks the generic templates of knn_select.h element by element: every get() is two dependent 64-bit
// variable shifts, every comparison a branch, and a warp executes the union of its lanes' paths -- measured 3-4 us per
// tick on the critical path of a CTA.  knn_small_topk_fast produces the same permutation with word-level operations:
//   * the row is (perm, rk): nibble p of `perm` = index at position p, nibble p of `rk` = its rank (kept in step);
//   * std::__unguarded_partition: the comparisons of all 16 positions against the pivot are two 16-bit masks (SWAR
//     byte compares on the rank word), the two scans become count-trailing / count-leading-zeros on them, and a swap
//     exchanges the two mask bits too (the element that arrives at `first` was not greater, the one at `last` not less);
//   * std::__insertion_sort is a STABLE sort (it only moves an element past strictly greater ones): short leftward scan
//     on the rank word, then one masked shift of the nibble field instead of an element-by-element move;
//   * the heap-select fallback of introselect (depth limit hit: rare) restarts with knn_small_topk.
// tests/test_knn_select.py checks it against torch.topk and against knn_small_topk (every n^n pattern for n <= 7).
SWARM_HD int knn_ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __ffs((int)x) - 1;
#else
  return __builtin_ctz(x);
#endif
}
SWARM_HD int knn_fls32(uint32_t x) {     // index of the highest set bit
#if defined(__CUDA_ARCH__)
  return 31 - __clz((int)x);
#else
  return 31 - __builtin_clz(x);
#endif
}
SWARM_HD int knn_nibble(uint64_t w, int p) { return (int)((w >> (4 * p)) & 15u); }
SWARM_HD void knn_swap_nibbles(uint64_t& w, int p, int q) {
  const uint64_t x = ((w >> (4 * p)) ^ (w >> (4 * q))) & 15u;
  w ^= (x << (4 * p)) | (x << (4 * q));
}
// bit7 of each byte of x (x & 0x8080...) -> 8 consecutive bits
SWARM_HD uint32_t knn_gather_msb(uint64_t x) { return (uint32_t)(((x >> 7) * 0x0102040810204080ull) >> 56); }
SWARM_HD uint32_t knn_spread8(uint32_t x) {     // bit b -> bit 2 b
  x = (x | (x << 4)) & 0x0F0Fu;
  x = (x | (x << 2)) & 0x3333u;
  x = (x | (x << 1)) & 0x5555u;
  return x;
}
// 16-bit mask of the positions whose rank is >= v (v in 0 .. 16)
SWARM_HD uint32_t knn_mask_ge(uint64_t rk, int v) {
  const uint64_t lo = 0x0F0F0F0F0F0F0F0Full, hi = 0x8080808080808080ull;
  const uint64_t vb = (uint64_t)(uint32_t)v * 0x0101010101010101ull;
  const uint64_t even = (((rk & lo) | hi) - vb) & hi;            // byte b: position 2 b
  const uint64_t odd = ((((rk >> 4) & lo) | hi) - vb) & hi;      // byte b: position 2 b + 1
  return knn_spread8(knn_gather_msb(even)) | (knn_spread8(knn_gather_msb(odd)) << 1);
}

// stable sort by rank of positions [a, b): std::__insertion_sort
SWARM_HD void knn_fast_insertion_sort(uint64_t& perm, uint64_t& rk, int a, int b) {
  for (int i = a + 1; i < b; ++i) {
    const int rv = knn_nibble(rk, i);
    int pos = i;
    while (pos > a && rv < knn_nibble(rk, pos - 1)) --pos;
    if (pos != i) {
      // nibbles [pos, i) move up by one, the element of position i lands at pos
      const uint64_t field = (~0ull >> (64 - 4 * (i - pos))) << (4 * pos);          // nibbles pos .. i - 1
      const uint64_t whole = field | (field << 4);                                   // nibbles pos .. i
      const uint64_t iv = (uint64_t)(uint32_t)knn_nibble(perm, i);
      perm = (perm & ~whole) | ((perm & field) << 4) | (iv << (4 * pos));
      rk = (rk & ~whole) | ((rk & field) << 4) | ((uint64_t)(uint32_t)rv << (4 * pos));
    }
  }
}

SWARM_HD uint64_t knn_small_topk_fast(uint64_t rank, int n, int k) {
  uint64_t perm = 0xFEDCBA9876543210ull, rk = rank;
  if (k <= 0) return perm;
  const int nth = k - 1;
  int first = 0, last = n, depth = knn_lg(n) * 2;
  while (last - first > 3) {
    if (depth == 0) return knn_small_topk(rank, n, k);            // heap-select branch: the step-by-step emulation
    --depth;
    const int mid = first + (last - first) / 2;
    // std::__move_median_to_first(first, first + 1, mid, last - 1)
    const int ra = knn_nibble(rk, first + 1), rb = knn_nibble(rk, mid), rc = knn_nibble(rk, last - 1);
    int m;
    if (ra < rb) m = (rb < rc) ? mid : ((ra < rc) ? last - 1 : first + 1);
    else m = (ra < rc) ? first + 1 : ((rb < rc) ? last - 1 : mid);
    knn_swap_nibbles(perm, first, m);
    knn_swap_nibbles(rk, first, m);
    // std::__unguarded_partition(first + 1, last, pivot = *first)
    const int pv = knn_nibble(rk, first);
    uint32_t not_lt = knn_mask_ge(rk, pv);                       // positions with rank >= pivot
    uint32_t not_gt = ~knn_mask_ge(rk, pv + 1) & 0xFFFFu;        // positions with rank <= pivot
    int f = first + 1, l = last, cut;
    while (true) {
      f += knn_ctz32(not_lt >> f);                               // while (*f < pivot) ++f
      l = knn_fls32(not_gt & ((1u << l) - 1u));                  // --l; while (pivot < *l) --l
      if (!(f < l)) { cut = f; break; }
      knn_swap_nibbles(perm, f, l);
      knn_swap_nibbles(rk, f, l);
      const uint32_t x1 = ((not_lt >> f) ^ (not_lt >> l)) & 1u;
      not_lt ^= (x1 << f) | (x1 << l);
      const uint32_t x2 = ((not_gt >> f) ^ (not_gt >> l)) & 1u;
      not_gt ^= (x2 << f) | (x2 << l);
      ++f;
    }
    if (cut <= nth) first = cut;
    else last = cut;
  }
  knn_fast_insertion_sort(perm, rk, first, last);
  knn_fast_insertion_sort(perm, rk, 0, k - 1);
  return perm;
}

}  // namespace swarm
#endif  // SWARM_KNN_SMALL_H
