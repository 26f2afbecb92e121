// knn_select.h -- tie-exact "k smallest" selection for the kNN graph builder.
//
// The reference builds its evaluation graph with torch.topk(distance_to_i, k, largest=False)
// (src/simulation/simulator.py:18-19).  On CPU, torch runs, per row, over (value, index) pairs:
//     k*64 <= n :  std::partial_sort(b, b+k, e)
//     otherwise :  std::nth_element(b, b+k-1, e); std::sort(b, b+k-1)
// with the comparator  (!isnan(x) && isnan(y)) || x < y  (libstdc++ 13 algorithms, unstable on ties).
// Start states are regular grids, so exact distance ties are the common case and the *order* of tied
// neighbours decides the edge list.  This header re-implements those algorithms step for step
// (introselect / median-of-three partition / insertion sort / heap select) so that a device thread
// produces the same index order.  It is __host__ __device__ so tests/test_knn_select.py can run it on
// the CPU against torch.topk without a GPU.
#ifndef SWARM_KNN_SELECT_H
#define SWARM_KNN_SELECT_H

#if defined(__CUDACC__)
#define SWARM_HD __host__ __device__ __forceinline__
#else
#define SWARM_HD inline
#endif

namespace swarm {

struct KnnPair {
  float v;
  int i;
};

SWARM_HD bool knn_less(const KnnPair& a, const KnnPair& b) {
  return ((a.v == a.v) && (b.v != b.v)) || (a.v < b.v);
}

// Arr must provide:  Pair get(int) const;  void set(int, const Pair&);  with knn_less(Pair, Pair) defined -- KnnPair
// (float value, index) for rows in memory, KnnRankPair (knn_small.h) for the register-resident rows of small swarms.
template <class Arr>
SWARM_HD void knn_swap(Arr& a, int p, int q) {
  auto t = a.get(p);
  a.set(p, a.get(q));
  a.set(q, t);
}

template <class Arr>
SWARM_HD void knn_move_median_to_first(Arr& a, int result, int ia, int ib, int ic) {
  auto A = a.get(ia), B = a.get(ib), C = a.get(ic);
  if (knn_less(A, B)) {
    if (knn_less(B, C)) knn_swap(a, result, ib);
    else if (knn_less(A, C)) knn_swap(a, result, ic);
    else knn_swap(a, result, ia);
  } else if (knn_less(A, C)) knn_swap(a, result, ia);
  else if (knn_less(B, C)) knn_swap(a, result, ic);
  else knn_swap(a, result, ib);
}

template <class Arr>
SWARM_HD int knn_unguarded_partition(Arr& a, int first, int last, int pivot) {
  while (true) {
    while (knn_less(a.get(first), a.get(pivot))) ++first;
    --last;
    while (knn_less(a.get(pivot), a.get(last))) --last;
    if (!(first < last)) return first;
    knn_swap(a, first, last);
    ++first;
  }
}

template <class Arr>
SWARM_HD int knn_partition_pivot(Arr& a, int first, int last) {
  int mid = first + (last - first) / 2;
  knn_move_median_to_first(a, first, first + 1, mid, last - 1);
  return knn_unguarded_partition(a, first + 1, last, first);
}

template <class Arr>
SWARM_HD void knn_unguarded_linear_insert(Arr& a, int last) {
  auto val = a.get(last);
  int next = last - 1;
  while (knn_less(val, a.get(next))) {
    a.set(last, a.get(next));
    last = next;
    --next;
  }
  a.set(last, val);
}

template <class Arr>
SWARM_HD void knn_insertion_sort(Arr& a, int first, int last) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    if (knn_less(a.get(i), a.get(first))) {
      auto val = a.get(i);
      for (int j = i; j > first; --j) a.set(j, a.get(j - 1));
      a.set(first, val);
    } else {
      knn_unguarded_linear_insert(a, i);
    }
  }
}

// ---- heap helpers (bits/stl_heap.h); indices are relative to `first` ----
template <class Arr, class Pair>
SWARM_HD void knn_push_heap(Arr& a, int first, int hole, int top, const Pair& value) {
  int parent = (hole - 1) / 2;
  while (hole > top && knn_less(a.get(first + parent), value)) {
    a.set(first + hole, a.get(first + parent));
    hole = parent;
    parent = (hole - 1) / 2;
  }
  a.set(first + hole, value);
}

template <class Arr, class Pair>
SWARM_HD void knn_adjust_heap(Arr& a, int first, int hole, int len, const Pair& value) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (knn_less(a.get(first + child), a.get(first + child - 1))) child--;
    a.set(first + hole, a.get(first + child));
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    a.set(first + hole, a.get(first + child - 1));
    hole = child - 1;
  }
  knn_push_heap(a, first, hole, top, value);
}

template <class Arr>
SWARM_HD void knn_pop_heap(Arr& a, int first, int last, int result) {
  auto value = a.get(result);
  a.set(result, a.get(first));
  knn_adjust_heap(a, first, 0, last - first, value);
}

template <class Arr>
SWARM_HD void knn_make_heap(Arr& a, int first, int last) {
  if (last - first < 2) return;
  const int len = last - first;
  int parent = (len - 2) / 2;
  while (true) {
    auto value = a.get(first + parent);
    knn_adjust_heap(a, first, parent, len, value);
    if (parent == 0) return;
    parent--;
  }
}

template <class Arr>
SWARM_HD void knn_heap_select(Arr& a, int first, int middle, int last) {
  knn_make_heap(a, first, middle);
  for (int i = middle; i < last; ++i)
    if (knn_less(a.get(i), a.get(first))) knn_pop_heap(a, first, middle, i);
}

template <class Arr>
SWARM_HD void knn_sort_heap(Arr& a, int first, int last) {
  while (last - first > 1) {
    --last;
    knn_pop_heap(a, first, last, last);
  }
}

SWARM_HD int knn_lg(int n) {
  int r = 0;
  while (n > 1) {
    n >>= 1;
    ++r;
  }
  return r;
}

template <class Arr>
SWARM_HD void knn_introselect(Arr& a, int first, int nth, int last, int depth_limit) {
  while (last - first > 3) {
    if (depth_limit == 0) {
      knn_heap_select(a, first, nth + 1, last);
      knn_swap(a, first, nth);
      return;
    }
    --depth_limit;
    int cut = knn_partition_pivot(a, first, last);
    if (cut <= nth) first = cut;
    else last = cut;
  }
  knn_insertion_sort(a, first, last);
}

// std::sort(first, last) -- introsort with an explicit stack (sub-ranges are disjoint, so the order in
// which they are partitioned does not change the result).
template <class Arr>
SWARM_HD void knn_sort(Arr& a, int first, int last) {
  if (first == last) return;
  const int kThreshold = 16;
  int stack_f[48], stack_l[48], stack_d[48];
  int sp = 0;
  stack_f[sp] = first; stack_l[sp] = last; stack_d[sp] = knn_lg(last - first) * 2; ++sp;
  while (sp > 0) {
    --sp;
    int f = stack_f[sp], l = stack_l[sp], d = stack_d[sp];
    while (l - f > kThreshold) {
      if (d == 0) {
        knn_heap_select(a, f, l, l);
        knn_sort_heap(a, f, l);
        break;
      }
      --d;
      int cut = knn_partition_pivot(a, f, l);
      stack_f[sp] = cut; stack_l[sp] = l; stack_d[sp] = d; ++sp;
      l = cut;
    }
  }
  if (last - first > kThreshold) {
    knn_insertion_sort(a, first, first + kThreshold);
    for (int i = first + kThreshold; i != last; ++i) knn_unguarded_linear_insert(a, i);
  } else {
    knn_insertion_sort(a, first, last);
  }
}

// torch.topk(values, k, largest=False, sorted=True) on CPU: after the call a.get(0..k-1) hold the result.
template <class Arr>
SWARM_HD void knn_topk_smallest(Arr& a, int n, int k) {
  if (k <= 0) return;
  if ((long long)k * 64 <= (long long)n) {
    knn_heap_select(a, 0, k, n);
    knn_sort_heap(a, 0, k);
  } else {
    if (k - 1 != n) knn_introselect(a, 0, k - 1, n, knn_lg(n) * 2);
    knn_sort(a, 0, k - 1);
  }
}

}  // namespace swarm
#endif  // SWARM_KNN_SELECT_H
