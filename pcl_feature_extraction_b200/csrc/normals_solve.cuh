// normals_solve.cuh — covariance moments -> normal + curvature (shared by normals.cu and knn_tile.cu).
#pragma once
#include "common.cuh"

namespace pfx {

struct Moments {
  double s[9];  // sum d (3), sum d d^T upper triangle (6); d = p - q formed exactly in double
  int n;
};

__device__ __forceinline__ void mom_add(Moments& m, float4 p, float4 q) {
  double dx = (double)p.x - (double)q.x, dy = (double)p.y - (double)q.y, dz = (double)p.z - (double)q.z;
  m.s[0] += dx; m.s[1] += dy; m.s[2] += dz;
  m.s[3] += dx * dx; m.s[4] += dx * dy; m.s[5] += dx * dz;
  m.s[6] += dy * dy; m.s[7] += dy * dz; m.s[8] += dz * dz;
  m.n += 1;
}

// covariance in double -> float Jacobi (cheap, ~1e-7) -> one double refinement step: Rayleigh
// quotient for l0, then the largest cross product of two rows of (C - l0 I) (pcl::eigen33's
// eigenvector construction) -> ~1e-12 of the double oracle unless the eigen-gap is ~1e-6 or less.
// Then flipNormalTowardsViewpoint and curvature = |l0 / trace|.
__device__ __forceinline__ float4 solve_normal_m9(const double* s, int n, float qx, float qy, float qz, float vx,
                                                  float vy, float vz) {
  const float nanv = __int_as_float(0x7fc00000);
  if (n == 0) return make_float4(nanv, nanv, nanv, nanv);
  double inv = 1.0 / (double)n;
  double mx = s[0] * inv, my = s[1] * inv, mz = s[2] * inv;
  double c[6];
  c[0] = s[3] * inv - mx * mx;
  c[1] = s[4] * inv - mx * my;
  c[2] = s[5] * inv - mx * mz;
  c[3] = s[6] * inv - my * my;
  c[4] = s[7] * inv - my * mz;
  c[5] = s[8] * inv - mz * mz;
  double tr = c[0] + c[3] + c[5];
  double sc = fmax(fmax(fabs(c[0]), fabs(c[1])), fmax(fmax(fabs(c[2]), fabs(c[3])), fmax(fabs(c[4]), fabs(c[5]))));
  double isc = (sc > 1e-300) ? 1.0 / sc : 1.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) c[i] *= isc;
  float a[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) a[i] = (float)c[i];
  float w[3], v[3][3];
  eig_sym3<float>(a, w, v, 8);
  double n0 = v[0][0], n1 = v[1][0], n2 = v[2][0];
  double cx = c[0] * n0 + c[1] * n1 + c[2] * n2;
  double cy = c[1] * n0 + c[3] * n1 + c[4] * n2;
  double cz = c[2] * n0 + c[4] * n1 + c[5] * n2;
  double l0 = (n0 * cx + n1 * cy + n2 * cz) / (n0 * n0 + n1 * n1 + n2 * n2);
  double r0[3] = {c[0] - l0, c[1], c[2]}, r1[3] = {c[1], c[3] - l0, c[4]}, r2[3] = {c[2], c[4], c[5] - l0};
  double e0[3] = {r0[1] * r1[2] - r0[2] * r1[1], r0[2] * r1[0] - r0[0] * r1[2], r0[0] * r1[1] - r0[1] * r1[0]};
  double e1[3] = {r0[1] * r2[2] - r0[2] * r2[1], r0[2] * r2[0] - r0[0] * r2[2], r0[0] * r2[1] - r0[1] * r2[0]};
  double e2[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
  double l_0 = e0[0] * e0[0] + e0[1] * e0[1] + e0[2] * e0[2];
  double l_1 = e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2];
  double l_2 = e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2];
  double bx = e0[0], by = e0[1], bz = e0[2], bl = l_0;
  if (l_1 > bl) { bx = e1[0]; by = e1[1]; bz = e1[2]; bl = l_1; }
  if (l_2 > bl) { bx = e2[0]; by = e2[1]; bz = e2[2]; bl = l_2; }
  if (bl > 1e-280) {  // otherwise (C - l0 I) has rank < 2: keep the Jacobi vector
    double il = rsqrt(bl);
    bx *= il; by *= il; bz *= il;
    if (bx * n0 + by * n1 + bz * n2 < 0) { bx = -bx; by = -by; bz = -bz; }
    n0 = bx; n1 = by; n2 = bz;
  }
  double trs = c[0] + c[3] + c[5];
  double curv = (tr != 0.0 && trs != 0.0) ? fabs(l0 / trs) : 0.0;
  double dp = ((double)vx - (double)qx) * n0 + ((double)vy - (double)qy) * n1 + ((double)vz - (double)qz) * n2;
  if (dp < 0) { n0 = -n0; n1 = -n1; n2 = -n2; }
  return make_float4((float)n0, (float)n1, (float)n2, (float)curv);
}

}  // namespace pfx
