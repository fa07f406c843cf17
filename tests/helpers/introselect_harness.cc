// Test harness: the hand-written nth_element twin (sdslam_b200/csrc/introselect.cuh, compiled for the host)
// against the real libstdc++ std::nth_element with the reference's comparator (response >).
#include <algorithm>
#include <cstdint>
#include <vector>

static long g_heap_select_calls = 0;
#define SDORB_INTROSELECT_TRACE ++g_heap_select_calls
#include "introselect.cuh"

extern "C" {
// a: packed entries (low 8 bits = response); runs both and returns the index of the first difference, -1 if equal
int twin_vs_std(const uint32_t* in, int n, int nth, uint32_t* out_twin, uint32_t* out_std) {
  std::vector<uint32_t> a(in, in + n), b(in, in + n);
  sdorb::nth_element_resp(a.data(), 0, nth, n);
  std::nth_element(b.begin(), b.begin() + nth, b.end(), [](uint32_t x, uint32_t y) { return (x & 0xFFu) > (y & 0xFFu); });
  std::copy(a.begin(), a.end(), out_twin);
  std::copy(b.begin(), b.end(), out_std);
  for (int i = 0; i < n; ++i)
    if (a[i] != b[i]) return i;
  return -1;
}
// every response string of length n over {1..distinct}, every nth in nths[]: returns the number of mismatching runs
long exhaustive(int n, int distinct, const int* nths, int n_nths) {
  std::vector<uint32_t> in(n), a(n), b(n);
  std::vector<int> digit(n, 0);
  long bad = 0;
  for (;;) {
    for (int i = 0; i < n; ++i) in[i] = ((uint32_t)i << 8) | (uint32_t)(digit[i] + 1);
    for (int k = 0; k < n_nths; ++k) {
      a = in;
      b = in;
      sdorb::nth_element_resp(a.data(), 0, nths[k], n);
      std::nth_element(b.begin(), b.begin() + nths[k], b.end(), [](uint32_t x, uint32_t y) { return (x & 0xFFu) > (y & 0xFFu); });
      if (a != b) ++bad;
    }
    int i = 0;
    while (i < n && ++digit[i] == distinct) digit[i++] = 0;
    if (i == n) break;
  }
  return bad;
}
long heap_select_calls() { return g_heap_select_calls; }

// McIlroy's adversary ("A killer adversary for quicksort", 1999) run against the real std::nth_element with the
// reference's descending comparator: items are "gas" until a comparison between two gas items freezes one at the
// next solid value.  The frozen values (folded to <= 250 distinct responses) form an input on which introselect's
// partitions shave only a few elements per round, so the depth limit runs out and heap-select takes over.
void adversary_responses(int n, int nth, uint32_t* resp_out) {
  const int GAS = 251;
  std::vector<int> val(n, GAS), idx(n);
  int nsolid = 0, candidate = 0;
  for (int i = 0; i < n; ++i) idx[i] = i;
  auto comp = [&](int x, int y) {
    if (val[x] == GAS && val[y] == GAS && nsolid < 250) {
      if (x == candidate) val[x] = ++nsolid; else val[y] = ++nsolid;
    }
    if (val[x] == GAS) candidate = x; else if (val[y] == GAS) candidate = y;
    // descending order of "badness": gas sorts FIRST under response >, solids after it in reverse freeze order
    return val[x] > val[y];
  };
  std::nth_element(idx.begin(), idx.begin() + nth, idx.end(), comp);
  for (int i = 0; i < n; ++i) resp_out[i] = (uint32_t)val[i];
}
// retainBest(v, n) + resize(n) on packed entries with the twin: returns the kept count
int twin_retain_best(uint32_t* a, int count, int n) {
  if (n >= 0 && count > n) {
    if (n == 0) return 0;
    sdorb::nth_element_resp(a, 0, n - 1, count);
    return n;
  }
  return count;
}
}
