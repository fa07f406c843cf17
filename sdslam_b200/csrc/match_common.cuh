// match_common.cuh -- DescriptorDistance (/root/reference/src/ORBmatcher.cc:1459-1473) for one 256-bit pair, shared by the
// all-pairs matcher (kernels_match.cu) and the guided matchers (kernels_search.cu).
#pragma once
#include <stdint.h>

namespace sdorb {

#ifndef SDORB_MATCH_CSA
#define SDORB_MATCH_CSA 3  // carry-save adders in front of the popcounts (3 -> 5 POPC per pair, 4 -> 4 POPC)
#endif
// popcount of a 256-bit XOR.  POPC runs on the quarter-rate XU pipe (16 lanes / clk / SM) and is what bounds the
// matcher, so three carry-save adders (two LOP3 each, ALU pipe) fold seven of the eight words into one "ones" word and
// three "twos" words first: 5 POPC per pair instead of 8, same integer result.
template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
  return d;
}
__device__ __forceinline__ int hamming256(const uint32_t q[8], const uint4 lo, const uint4 hi) {
  const uint32_t x0 = q[0] ^ lo.x, x1 = q[1] ^ lo.y, x2 = q[2] ^ lo.z, x3 = q[3] ^ lo.w;
  const uint32_t x4 = q[4] ^ hi.x, x5 = q[5] ^ hi.y, x6 = q[6] ^ hi.z, x7 = q[7] ^ hi.w;
  // carry-save adder: sum = a ^ b ^ c (LUT 0x96), carry = majority(a, b, c) (LUT 0xE8)
  const uint32_t s1 = lop3<0x96>(x0, x1, x2), c1 = lop3<0xE8>(x0, x1, x2);
  const uint32_t s2 = lop3<0x96>(x3, x4, x5), c2 = lop3<0xE8>(x3, x4, x5);
  const uint32_t s3 = lop3<0x96>(s1, s2, x6), c3 = lop3<0xE8>(s1, s2, x6);
#if SDORB_MATCH_CSA == 4
  const uint32_t s4 = lop3<0x96>(c1, c2, c3), c4 = lop3<0xE8>(c1, c2, c3);
  return __popc(s3) + __popc(x7) + 2 * __popc(s4) + 4 * __popc(c4);
#else
  return __popc(s3) + __popc(x7) + 2 * (__popc(c1) + __popc(c2) + __popc(c3));
#endif
}

}  // namespace sdorb
