// normals.cpp — CPU oracle: surface normals.  TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
// Restates pcl::NormalEstimationOMP<PointXYZRGB, Normal> as driven by reference tools.h:22-32
// (called from features.h:187); upstream features/impl/normal_3d_omp.hpp, features/normal_3d.h,
// common/impl/centroid.hpp, common/impl/eigen.hpp.  SURVEY.md A.2.
#include "oracle_common.hpp"
#include "pcl_oracle.h"

using namespace orc;

namespace {

// mode 0 — centred double covariance + Jacobi (the parity gate).
void normalDouble(const float* surf, const std::vector<Nbr>& nb, const float* qp, const float vp[3],
                  float out[4], float* gap) {
  const int m = (int)nb.size();
  double mu[3] = {0, 0, 0};
  for (const Nbr& b : nb)
    for (int a = 0; a < 3; ++a) mu[a] += (double)surf[3 * b.idx + a];
  for (int a = 0; a < 3; ++a) mu[a] /= m;
  double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (const Nbr& b : nb) {
    double d[3];
    for (int a = 0; a < 3; ++a) d[a] = (double)surf[3 * b.idx + a] - mu[a];
    for (int r = 0; r < 3; ++r)
      for (int c = r; c < 3; ++c) C[r][c] += d[r] * d[c];
  }
  for (int r = 0; r < 3; ++r)
    for (int c = r; c < 3; ++c) {
      C[r][c] /= m;
      C[c][r] = C[r][c];
    }
  double w[3], V[3][3];
  eigSym3(C, w, V);
  double nrm[3] = {V[0][0], V[1][0], V[2][0]};
  double tr = C[0][0] + C[1][1] + C[2][2];
  double curv = (tr != 0.0) ? std::fabs(w[0] / tr) : 0.0;
  // flipNormalTowardsViewpoint: (vp - p) . n < 0  =>  n = -n
  double dp = 0;
  for (int a = 0; a < 3; ++a) dp += ((double)vp[a] - (double)qp[a]) * nrm[a];
  if (dp < 0)
    for (int a = 0; a < 3; ++a) nrm[a] = -nrm[a];
  for (int a = 0; a < 3; ++a) out[a] = (float)nrm[a];
  out[3] = (float)curv;
  if (gap) *gap = (m >= 3 && w[2] > 0) ? (float)((w[1] - w[0]) / w[2]) : -1.f;
}

// ---- mode 1: PCL 1.7 float arithmetic (reporting only) ----
void computeRoots2f(float b, float c, float roots[3]) {
  roots[0] = 0.f;
  float d = b * b - 4.f * c;
  if (d < 0.f) d = 0.f;
  float sd = std::sqrt(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}
void computeRootsf(const float m[3][3], float roots[3]) {
  float c0 = m[0][0] * m[1][1] * m[2][2] + 2.f * m[0][1] * m[0][2] * m[1][2] -
             m[0][0] * m[1][2] * m[1][2] - m[1][1] * m[0][2] * m[0][2] - m[2][2] * m[0][1] * m[0][1];
  float c1 = m[0][0] * m[1][1] - m[0][1] * m[0][1] + m[0][0] * m[2][2] - m[0][2] * m[0][2] +
             m[1][1] * m[2][2] - m[1][2] * m[1][2];
  float c2 = m[0][0] + m[1][1] + m[2][2];
  if (std::fabs(c0) < std::numeric_limits<float>::epsilon()) {
    computeRoots2f(c2, c1, roots);
    return;
  }
  const float s_inv3 = 1.f / 3.f, s_sqrt3 = std::sqrt(3.f);
  float c2_over_3 = c2 * s_inv3;
  float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.f) a_over_3 = 0.f;
  float half_b = 0.5f * (c0 + c2_over_3 * (2.f * c2_over_3 * c2_over_3 - c1));
  float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.f) q = 0.f;
  float rho = std::sqrt(-a_over_3);
  float theta = std::atan2(std::sqrt(-q), half_b) * s_inv3;
  float ct = std::cos(theta), st = std::sin(theta);
  roots[0] = c2_over_3 + 2.f * rho * ct;
  roots[1] = c2_over_3 - rho * (ct + s_sqrt3 * st);
  roots[2] = c2_over_3 - rho * (ct - s_sqrt3 * st);
  if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
  if (roots[1] >= roots[2]) {
    std::swap(roots[1], roots[2]);
    if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
  }
  if (roots[0] <= 0.f) computeRoots2f(c2, c1, roots);
}

void normalPclFloat(const float* surf, const std::vector<Nbr>& nb, const float* qp,
                    const float vp[3], float out[4]) {
  // computeMeanAndCovarianceMatrix (1.7.x): single pass, float accumulators, absolute coordinates
  float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (const Nbr& b : nb) {
    const float* p = surf + 3 * b.idx;
    acc[0] += p[0] * p[0];
    acc[1] += p[0] * p[1];
    acc[2] += p[0] * p[2];
    acc[3] += p[1] * p[1];
    acc[4] += p[1] * p[2];
    acc[5] += p[2] * p[2];
    acc[6] += p[0];
    acc[7] += p[1];
    acc[8] += p[2];
  }
  float inv = 1.f / (float)nb.size();
  for (float& a : acc) a *= inv;
  float C[3][3];
  C[0][0] = acc[0] - acc[6] * acc[6];
  C[0][1] = acc[1] - acc[6] * acc[7];
  C[0][2] = acc[2] - acc[6] * acc[8];
  C[1][1] = acc[3] - acc[7] * acc[7];
  C[1][2] = acc[4] - acc[7] * acc[8];
  C[2][2] = acc[5] - acc[8] * acc[8];
  C[1][0] = C[0][1];
  C[2][0] = C[0][2];
  C[2][1] = C[1][2];
  // eigen33: scale, roots, largest cross product of rows of (C - l0 I)
  float scale = 0.f;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) scale = std::max(scale, std::fabs(C[r][c]));
  if (scale <= std::numeric_limits<float>::min()) scale = 1.f;
  float S[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) S[r][c] = C[r][c] / scale;
  float roots[3];
  computeRootsf(S, roots);
  float ev = roots[0] * scale;
  for (int r = 0; r < 3; ++r) S[r][r] -= roots[0];
  auto cross = [](const float a[3], const float b[3], float o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
  };
  float v1[3], v2[3], v3[3];
  cross(S[0], S[1], v1);
  cross(S[0], S[2], v2);
  cross(S[1], S[2], v3);
  auto sq = [](const float v[3]) { return v[0] * v[0] + v[1] * v[1] + v[2] * v[2]; };
  float l1 = sq(v1), l2 = sq(v2), l3 = sq(v3);
  const float* best = v3;
  float bl = l3;
  if (l1 >= l2 && l1 >= l3) {
    best = v1;
    bl = l1;
  } else if (l2 >= l1 && l2 >= l3) {
    best = v2;
    bl = l2;
  }
  float inv_l = 1.f / std::sqrt(bl);
  float nrm[3] = {best[0] * inv_l, best[1] * inv_l, best[2] * inv_l};
  float tr = C[0][0] + C[1][1] + C[2][2];
  float curv = (tr != 0.f) ? std::fabs(ev / tr) : 0.f;
  float dp = (vp[0] - qp[0]) * nrm[0] + (vp[1] - qp[1]) * nrm[1] + (vp[2] - qp[2]) * nrm[2];
  if (dp < 0)
    for (float& a : nrm) a = -a;
  out[0] = nrm[0];
  out[1] = nrm[1];
  out[2] = nrm[2];
  out[3] = curv;
}

}  // namespace

extern "C" int orc_normals(const float* surf, int n, const float* q, int nq, double radius, int k,
                           const float vp[3], int mode, float* normals4, int* n_nbrs,
                           float* eig_gap) {
  if ((radius > 0) == (k > 0)) return -1;  // Feature::initCompute: exactly one must be set
  Searcher s;
  s.init(surf, n, radius, k);
  const float nanv = std::numeric_limits<float>::quiet_NaN();
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 256)
    for (int i = 0; i < nq; ++i) {
      float* o = normals4 + 4 * (size_t)i;
      s.query(q + 3 * i, nb);
      if (n_nbrs) n_nbrs[i] = (int)nb.size();
      if (nb.empty()) {  // non-finite query or no neighbours -> NaN row
        o[0] = o[1] = o[2] = o[3] = nanv;
        if (eig_gap) eig_gap[i] = -1.f;
        continue;
      }
      if (mode == 0)
        normalDouble(surf, nb, q + 3 * i, vp, o, eig_gap ? eig_gap + i : nullptr);
      else {
        normalPclFloat(surf, nb, q + 3 * i, vp, o);
        if (eig_gap) eig_gap[i] = -1.f;
      }
    }
  }
  return 0;
}
