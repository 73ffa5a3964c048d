// fpfh.cpp — CPU oracle: SPFH / FPFH33.  TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
// Restates pcl::FPFHEstimation<PointXYZRGB, Normal, FPFHSignature33> as instantiated at reference
// evaluation.cpp:597-602 and driven by features.h:181-195; upstream features/impl/fpfh.hpp and
// features/src/pfh.cpp (computePairFeatures).  SURVEY.md A.6.
#include "oracle_common.hpp"
#include "pcl_oracle.h"

using namespace orc;

namespace {

// pcl::computePairFeatures: float 4-vectors with w = 0.
bool pairFeatures(const float* p1, const float* n1in, const float* p2, const float* n2in, float& f1,
                  float& f2, float& f3, float& f4) {
  float d[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  f4 = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (f4 == 0.0f) {
    f1 = f2 = f3 = f4 = 0;
    return false;
  }
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
  if (vn == 0.0f) {
    f1 = f2 = f3 = f4 = 0;
    return false;
  }
  for (float& x : v) x /= vn;
  float w[3] = {n1[1] * v[2] - n1[2] * v[1], n1[2] * v[0] - n1[0] * v[2], n1[0] * v[1] - n1[1] * v[0]};
  f2 = v[0] * n2[0] + v[1] * n2[1] + v[2] * n2[2];
  f1 = atan2f(w[0] * n2[0] + w[1] * n2[1] + w[2] * n2[2], n1[0] * n2[0] + n1[1] * n2[1] + n1[2] * n2[2]);
  return true;
}

inline bool finiteNormal(const float* nrm) {
  return std::isfinite(nrm[0]) && std::isfinite(nrm[1]) && std::isfinite(nrm[2]);
}

// computePointSPFHSignature: 3 x 11 bins, increments 100/(n-1)
void spfhRow(const float* surf, const float* normals4, int p, const std::vector<Nbr>& nb, float* h) {
  for (int b = 0; b < 33; ++b) h[b] = 0.f;
  if (nb.size() < 2) return;
  const float d_pi = 1.0f / (2.0f * (float)M_PI);
  float hist_incr = 100.0f / (float)(nb.size() - 1);
  for (const Nbr& b : nb) {
    if (b.idx == p) continue;
    // documented deviation: a pair with a non-finite normal is skipped
    if (!finiteNormal(normals4 + 4 * (size_t)p) || !finiteNormal(normals4 + 4 * (size_t)b.idx)) continue;
    float f1, f2, f3, f4;
    if (!pairFeatures(surf + 3 * (size_t)p, normals4 + 4 * (size_t)p, surf + 3 * (size_t)b.idx,
                      normals4 + 4 * (size_t)b.idx, f1, f2, f3, f4))
      continue;
    int i1 = (int)std::floor(11 * (((double)f1 + M_PI) * (double)d_pi));
    i1 = std::min(std::max(i1, 0), 10);
    int i2 = (int)std::floor(11 * (((double)f2 + 1.0) * 0.5));
    i2 = std::min(std::max(i2, 0), 10);
    int i3 = (int)std::floor(11 * (((double)f3 + 1.0) * 0.5));
    i3 = std::min(std::max(i3, 0), 10);
    h[i1] += hist_incr;
    h[11 + i2] += hist_incr;
    h[22 + i3] += hist_incr;
  }
}

}  // namespace

extern "C" int orc_spfh(const float* surf, const float* normals4, int n, const int* pidx, int np,
                        double radius, int k, float* out33) {
  if ((radius > 0) == (k > 0)) return -1;
  Searcher s;
  s.init(surf, n, radius, k);
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 128)
    for (int i = 0; i < np; ++i) {
      int p = pidx[i];
      s.query(surf + 3 * (size_t)p, nb);
      spfhRow(surf, normals4, p, nb, out33 + 33 * (size_t)i);
    }
  }
  return 0;
}

extern "C" int orc_fpfh(const float* surf, const float* normals4, int n, const float* q, int nq,
                        double radius, int k, float* out33) {
  if ((radius > 0) == (k > 0)) return -1;
  Searcher s;
  s.init(surf, n, radius, k);
  // neighbourhoods of the queries
  std::vector<std::vector<Nbr>> qn(nq);
#pragma omp parallel for schedule(dynamic, 128)
  for (int i = 0; i < nq; ++i) s.query(q + 3 * (size_t)i, qn[i]);
  // SPFH set = union of the neighbourhoods (std::set<int> upstream); lookup[p] = row
  std::vector<int> lookup(n, -1);
  std::vector<int> members;
  for (int i = 0; i < nq; ++i)
    for (const Nbr& b : qn[i])
      if (lookup[b.idx] < 0) {
        lookup[b.idx] = 0;
        members.push_back(b.idx);
      }
  std::sort(members.begin(), members.end());
  for (size_t r = 0; r < members.size(); ++r) lookup[members[r]] = (int)r;
  std::vector<float> spfh(members.size() * 33);
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 128)
    for (int r = 0; r < (int)members.size(); ++r) {
      int p = members[r];
      s.query(surf + 3 * (size_t)p, nb);
      spfhRow(surf, normals4, p, nb, spfh.data() + 33 * (size_t)r);
    }
  }
  // weightPointSPFHSignature
  const float nanv = std::numeric_limits<float>::quiet_NaN();
#pragma omp parallel for schedule(dynamic, 128)
  for (int i = 0; i < nq; ++i) {
    float* F = out33 + 33 * (size_t)i;
    if (qn[i].empty()) {
      for (int b = 0; b < 33; ++b) F[b] = nanv;
      continue;
    }
    for (int b = 0; b < 33; ++b) F[b] = 0.f;
    double sum[3] = {0, 0, 0};
    for (const Nbr& b : qn[i]) {
      if (b.d2 == 0) continue;  // "minus the query point itself"
      float w = 1.0f / b.d2;
      const float* h = spfh.data() + 33 * (size_t)lookup[b.idx];
      for (int blk = 0; blk < 3; ++blk)
        for (int c = 0; c < 11; ++c) {
          float val = h[11 * blk + c] * w;
          sum[blk] += val;
          F[11 * blk + c] += val;
        }
    }
    for (int blk = 0; blk < 3; ++blk) {
      if (sum[blk] != 0) sum[blk] = 100.0 / sum[blk];
      float sc = (float)sum[blk];
      for (int c = 0; c < 11; ++c) F[11 * blk + c] *= sc;
    }
  }
  return 0;
}
