// seqsum.h — the float value that `count` sequential additions of a constant reach: s = 0; s += incr (count
// times), IEEE binary32, round to nearest even.  PCL builds its FPFH / PFH histograms that way
// (`hist[bin] += hist_incr` once per vote), so an integer vote count plus this function reproduces its float
// histogram bit for bit - including the accumulated round-off, which for the tens of thousands of votes of a PFH
// bin is far above the parity tolerance.
//
// Inside one binade the additions are regular: s is a multiple of the binade's ulp u, so s + incr rounds to
// s + d with a constant d (incr rounded to a multiple of u; a tie alternates only on the first step, after which
// the mantissa parity is fixed).  The function therefore performs real float additions until three consecutive
// sums share a binade, measures d, jumps to the last sum of the binade with integer arithmetic and continues:
// O(number of binades) steps instead of O(count).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define PFX_HD __host__ __device__ inline
#else
#define PFX_HD inline
#endif

namespace pfx {

PFX_HD uint32_t seq_bits(float v) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(v);
#else
  uint32_t b;
  memcpy(&b, &v, 4);
  return b;
#endif
}
PFX_HD float seq_from_bits(uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(b);
#else
  float v;
  memcpy(&v, &b, 4);
  return v;
#endif
}
PFX_HD float seq_add(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b;  // no contraction / excess precision
  return r;
#endif
}

// incr >= 0 finite (or +inf / NaN, which simply propagate), count >= 0
PFX_HD float seq_float_sum(float incr, long long count) {
  float s = 0.f;
  long long left = count;
  while (left > 0) {
    const float s1 = seq_add(s, incr);
    --left;
    if (left == 0 || !(s1 < 3.0e38f) || s1 == s) return s1;  // done, overflow / NaN, or incr no longer registers
    const float s2 = seq_add(s1, incr);
    --left;
    const uint32_t b0 = seq_bits(s), b1 = seq_bits(s1), b2 = seq_bits(s2);
    const uint32_t e0 = b0 >> 23, e1 = b1 >> 23, e2 = b2 >> 23;  // sign bit is 0
    s = s2;
    if (left == 0 || e0 != e1 || e1 != e2 || e1 == 0) continue;  // binade boundary (or subnormal): step by step
    const uint32_t d = b2 - b1;  // mantissa step in ulps (same exponent: the bit patterns subtract exactly)
    if (d == 0) return s2;
    const uint32_t top = ((e2 + 1) << 23) - 1;  // last bit pattern of the binade
    long long m = (long long)((top - b2) / d);
    if (m > left) m = left;
    s = seq_from_bits(b2 + (uint32_t)m * d);
    left -= m;
  }
  return s;
}

}  // namespace pfx
