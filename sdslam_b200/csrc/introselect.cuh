// introselect.cuh -- the selection order of cv::KeyPointsFilter::retainBest, as a __host__ __device__ routine.
//
// The reference trims every cell and every level with retainBest(v, n) followed by resize(n)
// (/root/reference/src/ORBextractor.cc:586-588, 601-604).  retainBest is
//     std::nth_element(begin, begin + n - 1, end, response >);  partition the tail by response >= v[n-1]
// and the resize(n) that follows throws the partitioned tail away, so the surviving keypoints AND their
// order are exactly the first n elements as libstdc++'s nth_element leaves them.  FAST responses are small
// integers, so the cut almost always falls inside a run of ties: which tied keypoints survive is decided by
// the element moves of introselect (median-of-3 pivot, unguarded Hoare partition, insertion sort for <= 3
// elements, heap-select when the depth limit 2*floor(log2 n) runs out).  This file re-implements that
// algorithm move for move (GCC 13 bits/stl_algo.h, bits/stl_heap.h) on packed 32-bit entries whose low
// 8 bits are the response; tests/test_introselect.py checks it against the real std::nth_element.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SDORB_HD __host__ __device__ __forceinline__
#else
#define SDORB_HD inline
#endif

namespace sdorb {

// comp(a, b) of the reference: a.response > b.response
SDORB_HD bool resp_gt(uint32_t a, uint32_t b) { return (a & 0xFFu) > (b & 0xFFu); }

SDORB_HD void swap_u32(uint32_t* a, int i, int j) {
  const uint32_t t = a[i];
  a[i] = a[j];
  a[j] = t;
}

// std::__adjust_heap + std::__push_heap on a[first ...), heap indices relative to `first`
SDORB_HD void adjust_heap(uint32_t* a, int first, int hole, int len, uint32_t value) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (resp_gt(a[first + child], a[first + child - 1])) child--;
    a[first + hole] = a[first + child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    a[first + hole] = a[first + child - 1];
    hole = child - 1;
  }
  int parent = (hole - 1) / 2;
  while (hole > top && resp_gt(a[first + parent], value)) {
    a[first + hole] = a[first + parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  a[first + hole] = value;
}

// std::__heap_select(first, middle, last)
SDORB_HD void heap_select(uint32_t* a, int first, int middle, int last) {
  const int len = middle - first;
  if (len >= 2) {
    for (int parent = (len - 2) / 2;; --parent) {
      adjust_heap(a, first, parent, len, a[first + parent]);
      if (parent == 0) break;
    }
  }
  for (int i = middle; i < last; ++i)
    if (resp_gt(a[i], a[first])) {
      const uint32_t value = a[i];
      a[i] = a[first];
      adjust_heap(a, first, 0, len, value);
    }
}

// std::nth_element(a + first, a + nth, a + last, response >)
SDORB_HD void nth_element_resp(uint32_t* a, int first, int nth, int last) {
  if (first == last || nth == last) return;
  int n = last - first, lg = 0;
  while (n > 1) {
    n >>= 1;
    ++lg;
  }
  int depth = 2 * lg;
  while (last - first > 3) {
    if (depth == 0) {
#ifdef SDORB_INTROSELECT_TRACE
      SDORB_INTROSELECT_TRACE;
#endif
      heap_select(a, first, nth + 1, last);
      swap_u32(a, first, nth);
      return;
    }
    --depth;
    // __unguarded_partition_pivot: median of (first+1, mid, last-1) moved to first
    const int mid = first + (last - first) / 2;
    {
      const int x = first + 1, y = mid, z = last - 1;
      if (resp_gt(a[x], a[y])) {
        if (resp_gt(a[y], a[z]))
          swap_u32(a, first, y);
        else if (resp_gt(a[x], a[z]))
          swap_u32(a, first, z);
        else
          swap_u32(a, first, x);
      } else if (resp_gt(a[x], a[z]))
        swap_u32(a, first, x);
      else if (resp_gt(a[y], a[z]))
        swap_u32(a, first, z);
      else
        swap_u32(a, first, y);
    }
    // __unguarded_partition(first+1, last, pivot = *first)
    int lo = first + 1, hi = last;
    const uint32_t pivot_resp = a[first] & 0xFFu;  // the pivot slot itself is never moved by the loop
    for (;;) {
      while ((a[lo] & 0xFFu) > pivot_resp) ++lo;
      --hi;
      while (pivot_resp > (a[hi] & 0xFFu)) --hi;
      if (!(lo < hi)) break;
      swap_u32(a, lo, hi);
      ++lo;
    }
    const int cut = lo;
    if (cut <= nth)
      first = cut;
    else
      last = cut;
  }
  // __insertion_sort(first, last)
  for (int i = first + 1; i < last; ++i) {
    const uint32_t val = a[i];
    if (resp_gt(val, a[first])) {
      for (int k = i; k > first; --k) a[k] = a[k - 1];
      a[first] = val;
    } else {
      int k = i;
      while (resp_gt(val, a[k - 1])) {
        a[k] = a[k - 1];
        --k;
      }
      a[k] = val;
    }
  }
}

}  // namespace sdorb
