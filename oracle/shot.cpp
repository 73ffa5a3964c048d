// shot.cpp — CPU oracle: SHOT local reference frame + SHOT352.  TEST INFRASTRUCTURE ONLY.
// PARITY UNPINNED.  Restates pcl::SHOTEstimationOMP<PointXYZRGB, Normal, SHOT352> (+ its internal
// SHOTLocalReferenceFrameEstimationOMP) as instantiated at reference evaluation.cpp:770-775 and
// driven by features.h:181-195; upstream features/impl/shot.hpp, features/impl/shot_lrf.hpp.
// SURVEY.md A.9.  Neighbour order (matters only for the LRF sign-vote tie fallback) is defined as
// ascending (d2, index).
#include "oracle_common.hpp"
#include "pcl_oracle.h"

using namespace orc;

namespace {

const float kNaN = std::numeric_limits<float>::quiet_NaN();

// getLocalRF.  rf = x_axis, y_axis, z_axis.  Returns false (NaN frame) with < 5 valid neighbours.
bool localRF(const float* surf, const float* c, const std::vector<Nbr>& nb, double R, float rf[9],
             float* gap) {
  std::vector<double> vij;
  vij.reserve(nb.size() * 3);
  double M[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  double sum = 0;
  int valid = 0;
  for (const Nbr& b : nb) {
    const float* p = surf + 3 * (size_t)b.idx;
    if (p[0] == c[0] && p[1] == c[1] && p[2] == c[2]) continue;
    double v[3] = {(double)(p[0] - c[0]), (double)(p[1] - c[1]), (double)(p[2] - c[2])};
    double w = R - std::sqrt((double)b.d2);
    for (int r = 0; r < 3; ++r)
      for (int cc = 0; cc < 3; ++cc) M[r][cc] += w * (v[r] * v[cc]);
    sum += w;
    vij.insert(vij.end(), v, v + 3);
    ++valid;
  }
  if (gap) gap[0] = gap[1] = -1.f;
  if (valid < 5) {
    for (int i = 0; i < 9; ++i) rf[i] = kNaN;
    return false;
  }
  for (int r = 0; r < 3; ++r)
    for (int cc = 0; cc < 3; ++cc) M[r][cc] /= sum;
  double w[3], V[3][3];
  eigSym3(M, w, V);
  if (!std::isfinite(w[0]) || !std::isfinite(w[1]) || !std::isfinite(w[2])) {
    for (int i = 0; i < 9; ++i) rf[i] = kNaN;
    return false;
  }
  if (gap && w[2] > 0) {
    gap[0] = (float)((w[2] - w[1]) / w[2]);
    gap[1] = (float)((w[1] - w[0]) / w[2]);
  }
  double v1[3] = {V[0][2], V[1][2], V[2][2]};  // largest eigenvalue -> x
  double v3[3] = {V[0][0], V[1][0], V[2][0]};  // smallest -> z
  auto disambiguate = [&](double* ax) {
    int plus = 0;
    for (int e = 0; e < valid; ++e) {
      double dp = vij[3 * e] * ax[0] + vij[3 * e + 1] * ax[1] + vij[3 * e + 2] * ax[2];
      if (dp >= 0) ++plus;
    }
    plus = 2 * plus - valid;
    if (plus == 0) {
      const int points = 5;
      int med = valid / 2;
      for (int i = -points / 2; i <= points / 2; ++i) {
        int e = med - i;
        double dp = vij[3 * e] * ax[0] + vij[3 * e + 1] * ax[1] + vij[3 * e + 2] * ax[2];
        if (dp > 0) ++plus;
      }
      if (plus < points / 2 + 1)
        for (int a = 0; a < 3; ++a) ax[a] = -ax[a];
    } else if (plus < 0) {
      for (int a = 0; a < 3; ++a) ax[a] = -ax[a];
    }
  };
  disambiguate(v1);
  disambiguate(v3);
  float x[3] = {(float)v1[0], (float)v1[1], (float)v1[2]};
  float z[3] = {(float)v3[0], (float)v3[1], (float)v3[2]};
  float y[3] = {z[1] * x[2] - z[2] * x[1], z[2] * x[0] - z[0] * x[2], z[0] * x[1] - z[1] * x[0]};
  for (int a = 0; a < 3; ++a) {
    rf[a] = x[a];
    rf[3 + a] = y[a];
    rf[6 + a] = z[a];
  }
  return true;
}

// computePointSHOT: createBinDistanceShape + interpolateSingleChannel + normalizeHistogram
void shotRow(const float* surf, const float* normals4, const float* c, const std::vector<Nbr>& nb,
             double R, const float rf[9], float* shot) {
  const int nr_bins = 10, slots = nr_bins + 1, maxSectors = 32;
  const double r12 = R / 2, r14 = R / 4, r34 = 3 * R / 4;
  const double RAD45 = 0.78539816339744830961566084581988, RAD90 = 2 * RAD45, RAD135 = 3 * RAD45,
               RAD_PI_7_8 = 2.7488935718910690836548129603691;
  for (int i = 0; i < 352; ++i) shot[i] = 0.f;
  const float* fx = rf;
  const float* fy = rf + 3;
  const float* fz = rf + 6;
  for (const Nbr& b : nb) {
    const float* nrm = normals4 + 4 * (size_t)b.idx;
    if (!std::isfinite(nrm[0]) || !std::isfinite(nrm[1]) || !std::isfinite(nrm[2])) continue;
    double cosd = (double)(nrm[0] * fz[0] + nrm[1] * fz[1] + nrm[2] * fz[2]);
    if (cosd > 1.0) cosd = 1.0;
    if (cosd < -1.0) cosd = -1.0;
    double bd = ((1.0 + cosd) * nr_bins) / 2;

    const float* p = surf + 3 * (size_t)b.idx;
    float d[3] = {p[0] - c[0], p[1] - c[1], p[2] - c[2]};
    double dist = std::sqrt((double)b.d2);
    if (std::fabs(dist) < 1e-15) continue;
    double x = (double)(d[0] * fx[0] + d[1] * fx[1] + d[2] * fx[2]);
    double y = (double)(d[0] * fy[0] + d[1] * fy[1] + d[2] * fy[2]);
    double z = (double)(d[0] * fz[0] + d[1] * fz[1] + d[2] * fz[2]);
    if (std::fabs(y) < 1e-30) y = 0;
    if (std::fabs(x) < 1e-30) x = 0;
    if (std::fabs(z) < 1e-30) z = 0;
    int bit4 = ((y > 0) || ((y == 0.0) && (x < 0))) ? 1 : 0;
    int bit3 = ((x > 0) || ((x == 0.0) && (y > 0))) ? !bit4 : bit4;
    int di = ((bit4 << 3) + (bit3 << 2)) << 1;
    if ((x * y > 0) || (x == 0.0))
      di += (std::fabs(x) >= std::fabs(y)) ? 0 : 4;
    else
      di += (std::fabs(x) > std::fabs(y)) ? 4 : 0;
    di += z > 0 ? 1 : 0;
    di += (dist > r12) ? 2 : 0;
    int step = (int)std::floor(bd + 0.5);
    int vol = di * slots;
    bd -= step;
    double w = 1 - std::fabs(bd);
    if (bd > 0)
      shot[vol + ((step + 1) % nr_bins)] += (float)bd;
    else
      shot[vol + ((step - 1 + nr_bins) % nr_bins)] += -(float)bd;
    if (dist > r12) {
      double rd = (dist - r34) / r12;
      if (dist > r34)
        w += 1 - rd;
      else {
        w += 1 + rd;
        shot[(di - 2) * slots + step] -= (float)rd;
      }
    } else {
      double rd = (dist - r14) / r12;
      if (dist < r14)
        w += 1 + rd;
      else {
        w += 1 - rd;
        shot[(di + 2) * slots + step] += (float)rd;
      }
    }
    double ic = z / dist;
    if (ic < -1.0) ic = -1.0;
    if (ic > 1.0) ic = 1.0;
    double inc = std::acos(ic);
    if (inc > RAD90 || (std::fabs(inc - RAD90) < 1e-30 && z <= 0)) {
      double e = (inc - RAD135) / RAD90;
      if (inc > RAD135)
        w += 1 - e;
      else {
        w += 1 + e;
        shot[(di + 1) * slots + step] -= (float)e;
      }
    } else {
      double e = (inc - RAD45) / RAD90;
      if (inc < RAD45)
        w += 1 + e;
      else {
        w += 1 - e;
        shot[(di - 1) * slots + step] += (float)e;
      }
    }
    if (y != 0.0 || x != 0.0) {
      double az = std::atan2(y, x);
      int sel = di >> 2;
      double ad = (az - (-RAD_PI_7_8 + RAD45 * sel)) / RAD45;
      ad = std::max(-0.5, std::min(ad, 0.5));
      if (ad > 0) {
        w += 1 - ad;
        int ii = (di + 4) % maxSectors;
        shot[ii * slots + step] += (float)ad;
      } else {
        int ii = (di - 4 + maxSectors) % maxSectors;
        w += 1 + ad;
        shot[ii * slots + step] -= (float)ad;
      }
    }
    shot[vol + step] += (float)w;
  }
  double acc = 0;
  for (int j = 0; j < 352; ++j) acc += (double)(shot[j] * shot[j]);
  acc = std::sqrt(acc);
  for (int j = 0; j < 352; ++j) shot[j] /= (float)acc;
}

}  // namespace

extern "C" int orc_shot_lrf(const float* surf, int n, const float* q, int nq, double radius,
                            float* rf9, float* lrf_gap) {
  if (!(radius > 0)) return -1;
  Searcher s;
  s.init(surf, n, radius, 0);
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 128)
    for (int i = 0; i < nq; ++i) {
      s.query(q + 3 * (size_t)i, nb);
      localRF(surf, q + 3 * (size_t)i, nb, radius, rf9 + 9 * (size_t)i,
              lrf_gap ? lrf_gap + 2 * (size_t)i : nullptr);
    }
  }
  return 0;
}

extern "C" int orc_shot352(const float* surf, const float* normals4, int n, const float* q, int nq,
                           double radius, const float* lrf_in, float* out352, float* rf9) {
  if (!(radius > 0)) return -1;  // SHOT rejects k-search (SURVEY A.9 preconditions)
  Searcher s;
  s.init(surf, n, radius, 0);
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 64)
    for (int i = 0; i < nq; ++i) {
      float* o = out352 + 352 * (size_t)i;
      float* rf = rf9 + 9 * (size_t)i;
      const float* c = q + 3 * (size_t)i;
      s.query(c, nb);
      bool ok;
      if (lrf_in) {
        std::memcpy(rf, lrf_in + 9 * (size_t)i, 9 * sizeof(float));
        ok = std::isfinite(rf[0]) && std::isfinite(rf[3]) && std::isfinite(rf[6]);
      } else {
        ok = localRF(surf, c, nb, radius, rf, nullptr);
      }
      if (!ok || nb.empty() || !finite3(c) || nb.size() < 5) {
        for (int d = 0; d < 352; ++d) o[d] = kNaN;
        if (!ok || nb.empty() || !finite3(c))
          for (int d = 0; d < 9; ++d) rf[d] = kNaN;
        continue;
      }
      shotRow(surf, normals4, c, nb, radius, rf, o);
    }
  }
  return 0;
}
