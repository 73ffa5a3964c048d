// pfh.cpp — CPU oracle: PFH125 and PrincipalCurvatures.  TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
// Restates pcl::PFHEstimation<PointXYZRGB, Normal, PFHSignature125> (reference evaluation.cpp:676-695) and
// pcl::PrincipalCurvaturesEstimation<PointXYZRGB, Normal, PrincipalCurvatures> (evaluation.cpp:696-715), both
// driven through Features<T>::compute (features.h:181-195: search surface = cloud, input = keypoints, normals of
// the cloud, radius search); upstream features/impl/pfh.hpp and features/impl/principal_curvatures.hpp.
//
// PFH: every unordered pair (i, j), j < i in the distance-sorted neighbour list, votes once:
// computePairFeatures(p_i, n_i, p_j, n_j) -> three of the four features binned 5 x 5 x 5,
// hist[b1 + 5 b2 + 25 b3] += 100 / (n (n - 1) / 2) (integer division, sequential float additions).
// Deviation shared with the FPFH oracle: a pair with a non-finite normal is skipped (upstream bins NaN through an
// undefined float -> int conversion).
//
// PrincipalCurvatures: normals of the neighbours projected onto the tangent plane of normal n_idx, covariance of
// the projections, eigen decomposition; principal direction = eigenvector of the largest eigenvalue, pc1 / pc2 =
// largest / middle eigenvalue / n.  n_idx is normals[query ordinal], exactly as upstream indexes it (which is the
// query's own normal when the queries are the surface).  Deviation: the sums run in double and the eigenvalues
// come from a Jacobi solve (upstream: float sums and the closed-form eigen33, whose error on these tiny matrices
// is of the order of the values themselves); the eigenvector is built like pcl::computeCorrespondingEigenVector
// (largest cross product of two rows of C - lambda I), which also fixes its sign.
#include "oracle_common.hpp"
#include "pcl_oracle.h"

using namespace orc;

namespace {

bool pairFeaturesPfh(const float* p1, const float* n1in, const float* p2, const float* n2in, float& f1, float& f2,
                     float& f3, float& f4) {
  float d[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  f4 = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (f4 == 0.0f) return false;
  float n1[3] = {n1in[0], n1in[1], n1in[2]}, n2[3] = {n2in[0], n2in[1], n2in[2]};
  float angle1 = (n1[0] * d[0] + n1[1] * d[1] + n1[2] * d[2]) / f4;
  float angle2 = (n2[0] * d[0] + n2[1] * d[1] + n2[2] * d[2]) / f4;
  if (std::acos((double)std::fabs(angle1)) > std::acos((double)std::fabs(angle2))) {
    for (int a = 0; a < 3; ++a) {
      std::swap(n1[a], n2[a]);
      d[a] = -d[a];
    }
    f3 = -angle2;
  } else {
    f3 = angle1;
  }
  float v[3] = {d[1] * n1[2] - d[2] * n1[1], d[2] * n1[0] - d[0] * n1[2], d[0] * n1[1] - d[1] * n1[0]};
  float vn = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if (vn == 0.0f) return false;
  for (float& x : v) x /= vn;
  float w[3] = {n1[1] * v[2] - n1[2] * v[1], n1[2] * v[0] - n1[0] * v[2], n1[0] * v[1] - n1[1] * v[0]};
  f2 = v[0] * n2[0] + v[1] * n2[1] + v[2] * n2[2];
  f1 = atan2f(w[0] * n2[0] + w[1] * n2[1] + w[2] * n2[2], n1[0] * n2[0] + n1[1] * n2[1] + n1[2] * n2[2]);
  return true;
}

inline bool finiteN(const float* nrm) { return std::isfinite(nrm[0]) && std::isfinite(nrm[1]) && std::isfinite(nrm[2]); }

}  // namespace

// out: nq x 125; counts (optional): nq x 125 integer votes (stage-wise parity)
extern "C" int orc_pfh125(const float* surf, const float* normals4, int n, const float* q, int nq, double radius, int k,
                          float* out125, int* counts125) {
  if ((radius > 0) == (k > 0)) return -1;
  Searcher s;
  s.init(surf, n, radius, k);
  const float nanv = std::numeric_limits<float>::quiet_NaN();
  const float d_pi = 1.0f / (2.0f * (float)M_PI);
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 4)
    for (int i = 0; i < nq; ++i) {
      float* H = out125 + 125 * (size_t)i;
      int* C = counts125 ? counts125 + 125 * (size_t)i : nullptr;
      if (C) std::fill(C, C + 125, 0);
      nb.clear();
      if (finite3(q + 3 * (size_t)i)) s.query(q + 3 * (size_t)i, nb);
      if (nb.empty()) {
        for (int b = 0; b < 125; ++b) H[b] = nanv;
        continue;
      }
      for (int b = 0; b < 125; ++b) H[b] = 0.f;
      const size_t m = nb.size();
      const float hist_incr = 100.0f / (float)(m * (m - 1) / 2);
      for (size_t a = 0; a < m; ++a)
        for (size_t b = 0; b < a; ++b) {
          const int ia = nb[a].idx, ib = nb[b].idx;
          if (!finiteN(normals4 + 4 * (size_t)ia) || !finiteN(normals4 + 4 * (size_t)ib)) continue;
          float f1, f2, f3, f4;
          if (!pairFeaturesPfh(surf + 3 * (size_t)ia, normals4 + 4 * (size_t)ia, surf + 3 * (size_t)ib,
                               normals4 + 4 * (size_t)ib, f1, f2, f3, f4))
            continue;
          int i1 = (int)std::floor(5 * (((double)f1 + M_PI) * (double)d_pi));
          i1 = std::min(std::max(i1, 0), 4);
          int i2 = (int)std::floor(5 * (((double)f2 + 1.0) * 0.5));
          i2 = std::min(std::max(i2, 0), 4);
          int i3 = (int)std::floor(5 * (((double)f3 + 1.0) * 0.5));
          i3 = std::min(std::max(i3, 0), 4);
          const int h = i1 + 5 * i2 + 25 * i3;
          H[h] += hist_incr;
          if (C) ++C[h];
        }
    }
  }
  return 0;
}

// out: nq x 5 (principal direction x, y, z, pc1, pc2); gap (optional, nq): (l2 - l1) / l2 of the covariance
extern "C" int orc_principal_curvatures(const float* surf, const float* normals4, int n, const float* q, int nq,
                                        double radius, int k, float* out5, float* gap) {
  if ((radius > 0) == (k > 0)) return -1;
  Searcher s;
  s.init(surf, n, radius, k);
  const float nanv = std::numeric_limits<float>::quiet_NaN();
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 64)
    for (int i = 0; i < nq; ++i) {
      float* O = out5 + 5 * (size_t)i;
      if (gap) gap[i] = -1.f;
      nb.clear();
      if (finite3(q + 3 * (size_t)i)) s.query(q + 3 * (size_t)i, nb);
      if (nb.empty() || i >= n) {
        for (int b = 0; b < 5; ++b) O[b] = nanv;
        continue;
      }
      const float* ni = normals4 + 4 * (size_t)i;  // upstream: normals.points[(*indices_)[idx]]
      const double nx = ni[0], ny = ni[1], nz = ni[2];
      const double M[3][3] = {{1 - nx * nx, -nx * ny, -nx * nz}, {-ny * nx, 1 - ny * ny, -ny * nz}, {-nz * nx, -nz * ny, 1 - nz * nz}};
      const size_t m = nb.size();
      std::vector<double> pr(3 * m);
      double c[3] = {0, 0, 0};
      for (size_t a = 0; a < m; ++a) {
        const float* nn = normals4 + 4 * (size_t)nb[a].idx;
        for (int r = 0; r < 3; ++r) {
          pr[3 * a + r] = M[r][0] * nn[0] + M[r][1] * nn[1] + M[r][2] * nn[2];
          c[r] += pr[3 * a + r];
        }
      }
      for (int r = 0; r < 3; ++r) c[r] /= (double)m;
      double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      for (size_t a = 0; a < m; ++a) {
        const double d[3] = {pr[3 * a] - c[0], pr[3 * a + 1] - c[1], pr[3 * a + 2] - c[2]};
        for (int r = 0; r < 3; ++r)
          for (int t = 0; t < 3; ++t) C[r][t] += d[r] * d[t];
      }
      bool fin = true;
      for (int r = 0; r < 3; ++r)
        for (int t = 0; t < 3; ++t) fin = fin && std::isfinite(C[r][t]);
      if (!fin) {
        for (int b = 0; b < 5; ++b) O[b] = nanv;
        continue;
      }
      double w[3], V[3][3];
      eigSym3(C, w, V);
      // computeCorrespondingEigenVector(C, w[2]): rows of the scaled (C - l I), largest cross product
      double scale = 0;
      for (int r = 0; r < 3; ++r)
        for (int t = 0; t < 3; ++t) scale = std::max(scale, std::fabs(C[r][t]));
      if (scale <= std::numeric_limits<double>::min()) scale = 1.0;
      double S[3][3];
      for (int r = 0; r < 3; ++r)
        for (int t = 0; t < 3; ++t) S[r][t] = C[r][t] / scale - (r == t ? w[2] / scale : 0.0);
      auto cross = [](const double* a, const double* b, double* o) {
        o[0] = a[1] * b[2] - a[2] * b[1];
        o[1] = a[2] * b[0] - a[0] * b[2];
        o[2] = a[0] * b[1] - a[1] * b[0];
      };
      double v1[3], v2[3], v3[3];
      cross(S[0], S[1], v1);
      cross(S[0], S[2], v2);
      cross(S[1], S[2], v3);
      const double l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2], l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2],
                   l3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
      const double* best = v3;
      double bl = l3;
      if (l1 >= l2 && l1 >= l3) { best = v1; bl = l1; }
      else if (l2 >= l1 && l2 >= l3) { best = v2; bl = l2; }
      const double inv = 1.0 / std::sqrt(bl);
      for (int r = 0; r < 3; ++r) O[r] = (float)(best[r] * inv);
      const double im = 1.0 / (double)m;
      O[3] = (float)(w[2] * im);
      O[4] = (float)(w[1] * im);
      if (gap) gap[i] = w[2] > 0 ? (float)((w[2] - w[1]) / w[2]) : 0.f;
    }
  }
  return 0;
}

// MomentInvariants (reference evaluation.cpp:555-574 -> MomentInvariantsEstimation::computePointMomentInvariants):
// central second moments of the neighbourhood about its centroid, j1 = trace, j2 = sum of principal 2x2 minors,
// j3 = determinant.  Sums in double (upstream: float, sequential); no normals involved.  out: nq x 3.
extern "C" int orc_moment_invariants(const float* surf, int n, const float* q, int nq, double radius, int k, float* out3) {
  if ((radius > 0) == (k > 0)) return -1;
  Searcher s;
  s.init(surf, n, radius, k);
  const float nanv = std::numeric_limits<float>::quiet_NaN();
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 64)
    for (int i = 0; i < nq; ++i) {
      float* O = out3 + 3 * (size_t)i;
      nb.clear();
      if (finite3(q + 3 * (size_t)i)) s.query(q + 3 * (size_t)i, nb);
      if (nb.empty()) {
        O[0] = O[1] = O[2] = nanv;
        continue;
      }
      double c[3] = {0, 0, 0};
      for (const Nbr& b : nb)
        for (int a = 0; a < 3; ++a) c[a] += surf[3 * (size_t)b.idx + a];
      for (int a = 0; a < 3; ++a) c[a] /= (double)nb.size();
      double m200 = 0, m020 = 0, m002 = 0, m110 = 0, m101 = 0, m011 = 0;
      for (const Nbr& b : nb) {
        const double x = surf[3 * (size_t)b.idx] - c[0], y = surf[3 * (size_t)b.idx + 1] - c[1], z = surf[3 * (size_t)b.idx + 2] - c[2];
        m200 += x * x; m020 += y * y; m002 += z * z;
        m110 += x * y; m101 += x * z; m011 += y * z;
      }
      O[0] = (float)(m200 + m020 + m002);
      O[1] = (float)(m200 * m020 + m200 * m002 + m020 * m002 - m110 * m110 - m101 * m101 - m011 * m011);
      O[2] = (float)(m200 * m020 * m002 + 2 * m110 * m101 * m011 - m002 * m110 * m110 - m020 * m101 * m101 - m200 * m011 * m011);
    }
  }
  return 0;
}
