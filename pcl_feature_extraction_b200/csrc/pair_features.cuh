// pair_features.cuh — Darboux-frame pair features of two oriented points (pcl::computePairFeatures, used by
// FPFHEstimation and PFHEstimation), shared by fpfh.cu and pfh.cu.
#pragma once
#include "common.cuh"

namespace pfx {

// F1_BINS: number of bins the caller cuts f1's range [-pi, pi] into (11 for FPFH, 5 for PFH)
template <int F1_BINS>
__device__ __forceinline__ bool pair_features(float p1x, float p1y, float p1z, float4 n1, float p2x, float p2y,
                                              float p2z, float4 n2, float& f1, float& f2, float& f3) {
  float dx = p2x - p1x, dy = p2y - p1y, dz = p2z - p1z;
  float f4 = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  if (f4 == 0.0f) return false;
  float a1 = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(n1.x, dx), __fmul_rn(n1.y, dy)), __fmul_rn(n1.z, dz)), f4);
  float a2 = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(n2.x, dx), __fmul_rn(n2.y, dy)), __fmul_rn(n2.z, dz)), f4);
  // upstream: acos(|a1|) > acos(|a2|)  (false when either is NaN, i.e. |a| > 1)
  float b1 = fabsf(a1), b2 = fabsf(a2);
  bool swap = (b1 < b2) && (b2 <= 1.0f);
  float ux, uy, uz, wx_, wy_, wz_;
  if (swap) {
    ux = n2.x; uy = n2.y; uz = n2.z;
    wx_ = n1.x; wy_ = n1.y; wz_ = n1.z;
    dx = -dx; dy = -dy; dz = -dz;
    f3 = -a2;
  } else {
    ux = n1.x; uy = n1.y; uz = n1.z;
    wx_ = n2.x; wy_ = n2.y; wz_ = n2.z;
    f3 = a1;
  }
  float vx = __fsub_rn(__fmul_rn(dy, uz), __fmul_rn(dz, uy));
  float vy = __fsub_rn(__fmul_rn(dz, ux), __fmul_rn(dx, uz));
  float vz = __fsub_rn(__fmul_rn(dx, uy), __fmul_rn(dy, ux));
  float vn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz)));
  if (vn == 0.0f) return false;
  vx = __fdiv_rn(vx, vn); vy = __fdiv_rn(vy, vn); vz = __fdiv_rn(vz, vn);
  float wx = __fsub_rn(__fmul_rn(uy, vz), __fmul_rn(uz, vy));
  float wy = __fsub_rn(__fmul_rn(uz, vx), __fmul_rn(ux, vz));
  float wz = __fsub_rn(__fmul_rn(ux, vy), __fmul_rn(uy, vx));
  f2 = __fadd_rn(__fadd_rn(__fmul_rn(vx, wx_), __fmul_rn(vy, wy_)), __fmul_rn(vz, wz_));
  float sn = __fadd_rn(__fadd_rn(__fmul_rn(wx, wx_), __fmul_rn(wy, wy_)), __fmul_rn(wz, wz_));
  float cs = __fadd_rn(__fadd_rn(__fmul_rn(ux, wx_), __fmul_rn(uy, wy_)), __fmul_rn(uz, wz_));
  // f1 only selects one of F1_BINS bins (floor(F1_BINS (f1 + pi) / 2 pi)): the 9-term polynomial (3e-7 from atan2f) decides
  // it unless f1 lies within 2e-5 bins of an edge; those pairs (about 4 in 1e5) take the correctly rounded value -
  // atan2 in double, rounded to float - so the bin is the CPU's (computePairFeatures calls atan2f)
  f1 = fast_atan2f(sn, cs);
  {
    const float u = (f1 + 3.14159274f) * ((float)F1_BINS * 0.159154943f);  // F1_BINS / (2 pi)
    const float fr = u - floorf(u);
    if (fr < 2e-5f || fr > 1.0f - 2e-5f) f1 = (float)atan2((double)sn, (double)cs);
  }
  return true;
}

__device__ __forceinline__ int clamp_bin(double v) {
  int b = (int)floor(v);
  return min(max(b, 0), 10);
}
__device__ __forceinline__ int clamp_bin_n(double v, int nbins) {
  int b = (int)floor(v);
  return min(max(b, 0), nbins - 1);
}

}  // namespace pfx
