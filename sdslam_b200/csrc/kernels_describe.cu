// kernels_describe.cu -- IC_Angle orientation, rBRIEF descriptor and output assembly for sm_100a.
//
// Reference: /root/reference/src/ORBextractor.cc
//   IC_Angle            :78-102   integer moments m10/m01 over the radius-15 disc of the UNBLURRED level,
//                                 angle = cv::fastAtan2((float)m01, (float)m10)  (degrees, 7th-order polynomial, no FMA)
//   computeOrbDescriptor :105-143 a = cosf(angle*pi/180), b = sinf(..) (glibc sincosf), 512 rotated samples of the
//                                 BLURRED level at (cvRound(x*b + y*a), cvRound(x*a - y*b)), 256 comparisons
//   operator()          :669-676  pt *= mvScaleFactor[level] for level > 0, level-major output order
// One warp per output keypoint: lanes are the 31 columns of the disc for the moments, then the 32 descriptor bytes.
// The float chain is reproduced operation by operation with round-to-nearest intrinsics (no contraction); the
// sincosf is glibc 2.39's algorithm (ARM optimized-routines) evaluated in double, which is bit-identical to libm on
// every float in [0, 2*pi] (tests/test_oracle_primitives.py::test_sincosf_restated_exhaustive).
#include "kernels.cuh"

namespace sdorb {

// bit_pattern_31_ (src/ORBextractor.cc:146-404) as floats, bit-major: entry [bit * 32 + byte] = (x0, y0, x1, y1) of
// comparison 8 * byte + bit.  Lane = descriptor byte, so a warp reads 32 consecutive entries (512 B) through L1, which
// keeps the kernel free of shared memory and lets CTAs be two warps.
__device__ const float4 d_pattern_f[256] = {
#define P4(a, b, c, d) {(float)(a), (float)(b), (float)(c), (float)(d)},
#include "orb_pattern4.inc"
#undef P4
};

// cvRound(v) (cvtss2si, round half to even) for |v| < 2^22 without the conversion unit: adding 1.5 * 2^23 rounds to the
// nearest integer in the float's low mantissa bits with the same tie rule.
__device__ __forceinline__ int round_even_small(float v) { return __float_as_int(__fadd_rn(v, 12582912.f)) - 0x4B400000; }

// cv::fastAtan2 scalar path (OpenCV core mathfuncs_core: atan_f32), degrees in [0, 360)
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
  const float p1 = 0x1.ca44dep+5f, p3 = -0x1.2aaddcp+4f, p5 = 0x1.1d3f7ep+3f, p7 = -0x1.4515b2p+1f;
  const float eps = 0x1p-52f;  // (float)DBL_EPSILON
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = __fdiv_rn(ay, __fadd_rn(ax, eps));
    c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = __fdiv_rn(ax, __fadd_rn(ay, eps));
    c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

// glibc sincosf for 0 <= y < 120 (the extractor only produces y in [0, 2*pi])
__device__ __forceinline__ void sincosf_glibc(float y, float* sinp, float* cosp) {
  const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
  const double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10,
               C4 = 0x1.99343027bf8c3p-16, S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7,
               S3 = -0x1.994eb3774cf24p-13;
  const uint32_t top = (__float_as_uint(y) >> 20) & 0x7ff;
  double x = (double)y;
  int n = 0;
  double sgn = 1.0, flip = 1.0;  // flip = -1 selects the negated cosine table (quadrants 2,3)
  if (top < ((__float_as_uint(0x1.921FB6p-1f) >> 20) & 0x7ff)) {
    if (top < ((__float_as_uint(0x1p-12f) >> 20) & 0x7ff)) {
      *sinp = y;
      *cosp = 1.0f;
      return;
    }
  } else {
    const double r = __dmul_rn(x, hpi_inv);
    n = (__double2int_rz(r) + 0x800000) >> 24;
    x = __dsub_rn(x, __dmul_rn((double)n, hpi));
    sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    if (n & 2) flip = -1.0;
  }
  const double x2 = __dmul_rn(x, x);  // (x*s)^2 == x^2
  const double xs = __dmul_rn(x, sgn);
  const double c0 = C0 * flip, c1 = C1 * flip, c2k = C2 * flip, c3 = C3 * flip, c4 = C4 * flip;  // exact sign flips
  const double x4 = __dmul_rn(x2, x2);
  const double x3 = __dmul_rn(x2, xs);
  const double cc2 = __dadd_rn(c3, __dmul_rn(x2, c4));
  const double ss1 = __dadd_rn(S2, __dmul_rn(x2, S3));
  const double cc1 = __dadd_rn(c0, __dmul_rn(x2, c1));
  const double x5 = __dmul_rn(x3, x2);
  const double x6 = __dmul_rn(x4, x2);
  const double s = __dadd_rn(xs, __dmul_rn(x3, S1));
  const double c = __dadd_rn(cc1, __dmul_rn(x4, c2k));
  const float sv = (float)__dadd_rn(s, __dmul_rn(x5, ss1));
  const float cv = (float)__dadd_rn(c, __dmul_rn(x6, cc2));
  if (n & 1) {
    *sinp = cv;
    *cosp = sv;
  } else {
    *sinp = sv;
    *cosp = cv;
  }
}

constexpr int DESC_THREADS = 128;
constexpr int PATCH_ROWS = 37, PATCH_WORDS = 11, PATCH_LOAD_WORDS = 10;  // +-18 rows; 37 columns starting up to 3 px left of kx - 18

__global__ void __launch_bounds__(DESC_THREADS) describe_kernel(const FrameGeom* __restrict__ geom, BatchPlanes p,
                                                                SelectBuffers buf, const int* __restrict__ umax_tab,
                                                                float* __restrict__ kps_out, uint8_t* __restrict__ desc_out,
                                                                int32_t* __restrict__ counts_out, int capacity) {
  __shared__ uint32_t s_patch[DESC_THREADS / 32][PATCH_ROWS * PATCH_WORDS];
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int slot = blockIdx.x * (DESC_THREADS / 32) + (threadIdx.x >> 5);
  const int frame = blockIdx.y;
  const int nlevels = geom->nlevels;
  // locate the slot: exclusive prefix of the per-level counts
  const int my = lane < nlevels ? buf.sel_count[(int64_t)frame * nlevels + lane] : 0;
  int incl = my;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  if (slot == 0 && lane == 0) counts_out[frame] = total;
  if (slot >= total || slot >= capacity) return;
  const unsigned hit = __ballot_sync(0xffffffffu, slot < incl);
  const int level = __ffs(hit) - 1;
  const int idx = slot - (__shfl_sync(0xffffffffu, incl, level) - __shfl_sync(0xffffffffu, my, level));
  const LevelGeom& L = geom->lv[level];
  const uint32_t e = buf.sel[(int64_t)frame * geom->sel_total + L.sel_base + idx];
  const int kx = SDORB_ENTRY_X(e), ky = SDORB_ENTRY_Y(e);

  // ---- stage the blurred patch: the 512 descriptor samples lie within +-18 px of the keypoint.  37 rows x 11 aligned
  // words are fetched as coalesced row segments (74 sectors) instead of 512 scattered byte reads through L1; issued
  // here so that they complete under the intensity-centroid phase.
  uint32_t* patch = s_patch[threadIdx.x >> 5];
  const int xa = (kx - 18) & ~3;  // first staged column (word aligned); the patch row r holds image row ky - 18 + r
  {
    // 37 columns starting at most 3 px right of xa fit in 10 words: lane = (row % 3, word) for 30 lanes, three rows per step
    // (the patch rows keep a pitch of 11 words: odd, so the scattered byte reads below spread over the banks)
    const int pr = lane / PATCH_LOAD_WORDS, pc = lane - pr * PATCH_LOAD_WORDS;
    const uint8_t* bsrc = p.blur + L.plane_base * p.batch_cap + (int64_t)frame * L.plane_bytes + (int64_t)(ky - 18 + pr) * L.pitch + xa + 4 * pc;
    uint32_t* dst = patch + pr * PATCH_WORDS + pc;
    const int64_t step = 3 * (int64_t)L.pitch;
    const bool on = lane < 3 * PATCH_LOAD_WORDS;
#pragma unroll
    for (int j = 0; j < (PATCH_ROWS + 2) / 3; ++j) {
      if (on && 3 * j + pr < PATCH_ROWS) dst[3 * j * PATCH_WORDS] = *reinterpret_cast<const uint32_t*>(bsrc);
      bsrc += step;
    }
  }

  // ---- IC_Angle on the unblurred level
  int pitch;
  const uint8_t* img;
  if (level == 0) {
    pitch = p.img0_pitch;
    img = p.img0 + (int64_t)frame * p.img0_frame_stride;
  } else {
    pitch = L.pitch;
    img = p.pyr + L.plane_base * p.batch_cap + (int64_t)frame * L.plane_bytes;
  }
  const uint8_t* center = img + (int64_t)ky * pitch + kx;
  const int u = lane - SDORB_HALF_PATCH;
  int m10 = 0, m01 = 0;
  if (lane < 31) {
    // lane = column u of the disc; its rows are v = -vmax .. vmax with vmax = max{v : |u| <= umax[v]}
    // the disc is symmetric under u <-> v (the constructor's second loop makes umax so, src/ORBextractor.cc:449-456), hence
    // max{v : |u| <= umax[v]} = umax[|u|]
    const int au = u < 0 ? -u : u;
    const int vmax = umax_tab[au];
    const uint8_t* pc = center + u;
    int colsum = *pc;  // sum of the column (for m10), v-weighted difference (for m01)
#pragma unroll
    for (int v = 1; v <= SDORB_HALF_PATCH; ++v) {
      if (v <= vmax) {
        const int plus = pc[v * pitch], minus = pc[-v * pitch];  // 32-bit offsets (IMAD.WIDE per address measured 24 % slower)
        m01 += v * (plus - minus);
        colsum += plus + minus;
      }
    }
    m10 = u * colsum;
  }
  m10 = __reduce_add_sync(0xffffffffu, m10);
  m01 = __reduce_add_sync(0xffffffffu, m01);
  const float angle = fast_atan2_deg((float)m01, (float)m10);

  // ---- descriptor on the blurred level: lane i produces byte i
  float sn, cs;
  sincosf_glibc(__fmul_rn(angle, 0x1.1df46ap-6f), &sn, &cs);
  const float a = cs, b = sn;
  __syncwarp();
  const uint8_t* bc = reinterpret_cast<const uint8_t*>(patch) + 18 * (PATCH_WORDS * 4) + (kx - xa);  // the keypoint inside the staged patch
  constexpr int bp = PATCH_WORDS * 4;
  const float4* pat = d_pattern_f + lane;
  int val = 0;
#pragma unroll
  for (int bit = 0; bit < 8; ++bit) {
    const float4 pp = __ldg(pat + 32 * bit);
    const int r0 = round_even_small(__fmaf_rn(pp.x, b, __fmul_rn(pp.y, a))), c0 = round_even_small(__fmaf_rn(pp.x, a, -__fmul_rn(pp.y, b)));
    const int r1 = round_even_small(__fmaf_rn(pp.z, b, __fmul_rn(pp.w, a))), c1 = round_even_small(__fmaf_rn(pp.z, a, -__fmul_rn(pp.w, b)));
    const int t0 = bc[r0 * bp + c0], t1 = bc[r1 * bp + c1];
    val |= (t0 < t1) << bit;
  }
  desc_out[((int64_t)frame * capacity + slot) * 32 + lane] = (uint8_t)val;
  if (lane == 0) {
    float* o = kps_out + ((int64_t)frame * capacity + slot) * 7;
    float fx = (float)kx, fy = (float)ky;
    if (level != 0) {
      fx = __fmul_rn(fx, L.scale);
      fy = __fmul_rn(fy, L.scale);
    }
    o[0] = fx;
    o[1] = fy;
    o[2] = (float)L.scaled_patch_size;
    o[3] = angle;
    o[4] = (float)SDORB_ENTRY_SCORE(e);
    o[5] = __int_as_float(level);
    o[6] = __int_as_float(-1);
  }
}

void launch_describe(const FrameGeom* d_geom, const FrameGeom& g, const BatchPlanes& p, const SelectBuffers& b,
                     const int* d_umax, void* kps_out, uint8_t* desc_out, int32_t* counts_out, int capacity,
                     int nframes, cudaStream_t s) {
  const int slots = g.sel_total > 0 ? g.sel_total : 1;
  const int per_block = DESC_THREADS / 32;
  launch_pdl(describe_kernel, dim3((slots + per_block - 1) / per_block, nframes), dim3(DESC_THREADS), 0, s, d_geom, p, b, d_umax,
             (float*)kps_out, desc_out, counts_out, capacity);
}

}  // namespace sdorb
