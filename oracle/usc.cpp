// usc.cpp — CPU oracle: Unique Shape Context (USC, 12 x 11 x 15 = 1980 bins).  TEST INFRASTRUCTURE ONLY.
// PARITY UNPINNED.  Restates pcl::UniqueShapeContext<PointXYZRGB, ShapeContext1980, ReferenceFrame> (reference
// evaluation.cpp:344-371: setMinimalRadius(r / 10), setPointDensityRadius(r / 5), local radius left at PCL's 2.5,
// search radius r through features.h:181-195; upstream features/impl/usc.hpp):
//   frame    = SHOT local reference frame of the query at local_radius (orc_shot_lrf);
//   per neighbour within r (squared distance not "equal" to 0, i.e. > FLT_EPSILON): radius = sqrt(d2); azimuth phi
//   = angle of its tangent-plane projection against the x axis in [0, 360] degrees; elevation theta = angle against
//   the z axis in [0, 180]; bin (j, k, l) = first log-spaced radius shell / elevation / azimuth division that holds
//   it (0 when none does); weight = 1 / (points within density_radius of the NEIGHBOUR) / cbrt(volume of the bin);
//   desc[l * 11 * 15 + k * 15 + j] += weight (float, in the distance-sorted order of the neighbours).
// NaN frame -> NaN descriptor and a zero frame, as upstream.
// Definitions where upstream depends on a library version: Eigen 3.2's normalize() multiplies by 1 / norm.
#include "oracle_common.hpp"
#include "pcl_oracle.h"

using namespace orc;

namespace {

constexpr int AZ = 12, EL = 11, RB = 15;

struct UscTables {
  float radii[RB + 1], theta[EL + 1], phi[AZ + 1], vol[AZ * EL * RB];
  UscTables(double min_radius, double search_radius) {
    const float az_int = 360.0f / static_cast<float>(AZ), el_int = 180.0f / static_cast<float>(EL);
    for (int j = 0; j < RB + 1; ++j)
      radii[j] = static_cast<float>(std::exp(std::log(min_radius) + ((static_cast<float>(j) / static_cast<float>(RB)) *
                                                                      std::log(search_radius / min_radius))));
    for (int k = 0; k < EL + 1; ++k) theta[k] = static_cast<float>(k) * el_int;
    for (int l = 0; l < AZ + 1; ++l) phi[l] = static_cast<float>(l) * az_int;
    auto deg2rad = [](float a) { return a * 0.017453293f; };
    const float integr_phi = deg2rad(phi[1]) - deg2rad(phi[0]);
    const float e = 1.0f / 3.0f;
    for (int j = 0; j < RB; ++j) {
      const float integr_r = (radii[j + 1] * radii[j + 1] * radii[j + 1] / 3) - (radii[j] * radii[j] * radii[j] / 3);
      for (int k = 0; k < EL; ++k) {
        const float integr_theta = cosf(deg2rad(theta[k])) - cosf(deg2rad(theta[k + 1]));
        const float V = integr_phi * integr_theta * integr_r;
        for (int l = 0; l < AZ; ++l) vol[(l * EL * RB) + k * RB + j] = 1.0f / powf(V, e);
      }
    }
  }
};

}  // namespace

// out: nq x 1980, rf9: nq x 9 (lrf_in optional: frames given).  density (optional out, n): neighbour counts of the
// surface points at density_radius.
extern "C" int orc_usc1980(const float* surf, int n, const float* q, int nq, double search_radius, double min_radius,
                           double density_radius, double local_radius, const float* lrf_in, float* out1980, float* rf9,
                           int* density_out) {
  if (!(search_radius > 0) || !(min_radius > 0) || !(density_radius > 0) || !(local_radius > 0) || search_radius < min_radius)
    return -1;
  if (lrf_in)
    std::memcpy(rf9, lrf_in, (size_t)nq * 9 * sizeof(float));
  else if (orc_shot_lrf(surf, n, q, nq, local_radius, rf9, nullptr) != 0)
    return -1;
  std::vector<int> dens(std::max(n, 1));
  if (orc_radius_count(surf, n, surf, n, density_radius, dens.data()) != 0) return -1;
  if (density_out) std::memcpy(density_out, dens.data(), (size_t)n * sizeof(int));
  const UscTables T(min_radius, search_radius);
  Searcher s;
  s.init(surf, n, search_radius, 0);
  const float nanv = std::numeric_limits<float>::quiet_NaN();
  const float rad2deg = 57.29578f;
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 16)
    for (int i = 0; i < nq; ++i) {
      float* D = out1980 + 1980 * (size_t)i;
      float* rf = rf9 + 9 * (size_t)i;
      if (!std::isfinite(rf[0]) || !std::isfinite(rf[3]) || !std::isfinite(rf[6])) {
        for (int b = 0; b < 1980; ++b) D[b] = nanv;
        for (int b = 0; b < 9; ++b) rf[b] = 0.f;
        continue;
      }
      for (int b = 0; b < 1980; ++b) D[b] = 0.f;
      const float* o = q + 3 * (size_t)i;
      const float *xa = rf, *nz = rf + 6;
      nb.clear();
      if (finite3(o)) s.query(o, nb);
      for (const Nbr& b : nb) {
        if (std::fabs(b.d2 - 0.0f) <= std::numeric_limits<float>::epsilon()) continue;
        const float* p = surf + 3 * (size_t)b.idx;
        const float r = sqrtf(b.d2);
        // pcl::geometry::project + proj -= origin + normalize (Eigen 3.2: times 1 / norm)
        const float po[3] = {p[0] - o[0], p[1] - o[1], p[2] - o[2]};
        float lambda = nz[0] * po[0];
        lambda = lambda + nz[1] * po[1];
        lambda = lambda + nz[2] * po[2];
        float pr[3];
        for (int a = 0; a < 3; ++a) pr[a] = (p[a] - lambda * nz[a]) - o[a];
        float pn = pr[0] * pr[0];
        pn = pn + pr[1] * pr[1];
        pn = pn + pr[2] * pr[2];
        const float inv = 1.0f / std::sqrt(pn);
        for (int a = 0; a < 3; ++a) pr[a] = pr[a] * inv;
        const float cr[3] = {xa[1] * pr[2] - xa[2] * pr[1], xa[2] * pr[0] - xa[0] * pr[2], xa[0] * pr[1] - xa[1] * pr[0]};
        float cn = cr[0] * cr[0];
        cn = cn + cr[1] * cr[1];
        cn = cn + cr[2] * cr[2];
        float xd = xa[0] * pr[0];
        xd = xd + xa[1] * pr[1];
        xd = xd + xa[2] * pr[2];
        float phi = rad2deg * atan2f(std::sqrt(cn), xd);
        float cdn = cr[0] * nz[0];
        cdn = cdn + cr[1] * nz[1];
        cdn = cdn + cr[2] * nz[2];
        phi = cdn < 0.f ? (360.0f - phi) : phi;
        float no[3] = {po[0], po[1], po[2]};
        float nn = no[0] * no[0];
        nn = nn + no[1] * no[1];
        nn = nn + no[2] * no[2];
        const float ninv = 1.0f / std::sqrt(nn);
        for (int a = 0; a < 3; ++a) no[a] = no[a] * ninv;
        float th = nz[0] * no[0];
        th = th + nz[1] * no[1];
        th = th + nz[2] * no[2];
        th = rad2deg * acosf(std::min(1.0f, std::max(-1.0f, th)));
        int j = 0, k = 0, l = 0;
        for (int rad = 1; rad < RB + 1; ++rad)
          if (r <= T.radii[rad]) { j = rad - 1; break; }
        for (int ang = 1; ang < EL + 1; ++ang)
          if (th <= T.theta[ang]) { k = ang - 1; break; }
        for (int ang = 1; ang < AZ + 1; ++ang)
          if (phi <= T.phi[ang]) { l = ang - 1; break; }
        const float point_density = static_cast<float>(dens[b.idx]);
        const float w = (1.0f / point_density) * T.vol[(l * EL * RB) + (k * RB) + j];
        D[(l * EL * RB) + (k * RB) + j] += w;
      }
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ 3DSC
// pcl::ShapeContext3DEstimation<PointXYZRGB, Normal, ShapeContext1980> (reference evaluation.cpp:319-345:
// setMinimalRadius(r / 10), setPointDensityRadius(r / 5), search radius r through features.h:181-195; upstream
// features/impl/3dsc.hpp).  Same bins and weights as USC (which upstream derived from it); only the frame differs:
//   z = the NORMAL of the nearest surface point of the query (first of the (d2, index)-sorted neighbours),
//   x = a RANDOM vector (three uniform [0, 1) draws) made orthogonal to z by solving for one component
//       (the z component when |n_z| > FLT_EPSILON, else y, else x), normalised (Eigen 3.2: times 1 / norm),
//   y = z cross x.
// Upstream draws from a boost::mt19937 seeded with the wall clock: unpinnable.  Contract here: the three draws of
// query i are the top 24 bits of SplitMix64(seed + golden * (3 i + t + 1)), t = 0, 1, 2, as floats in [0, 1).
// The descriptor is computed once per query (PCL 1.7's ShapeContext1980 output); upstream zeroes rf afterwards
// ("3DSC does not define a repeatable local RF"), and so does this.  No neighbours / NaN normal -> NaN descriptor.
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

extern "C" int orc_sc3d_frames(const float* surf, const float* normals4, int n, const float* q, int nq, double search_radius,
                               unsigned long long seed, float* rf9) {
  if (!(search_radius > 0)) return -1;
  Searcher s;
  s.init(surf, n, search_radius, 0);
  const float nanv = std::numeric_limits<float>::quiet_NaN();
  const float eps = std::numeric_limits<float>::epsilon();
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 64)
    for (int i = 0; i < nq; ++i) {
      float* rf = rf9 + 9 * (size_t)i;
      for (int b = 0; b < 9; ++b) rf[b] = nanv;
      const float* o = q + 3 * (size_t)i;
      nb.clear();
      if (finite3(o)) s.query(o, nb);
      if (nb.empty()) continue;
      const float* nz = normals4 + 4 * (size_t)nb[0].idx;  // strict "<" over the sorted list keeps its first element
      if (!finite3(nz)) continue;
      float x[3];
      for (int t = 0; t < 3; ++t) {
        const uint64_t z = splitmix64(seed + 0x9E3779B97F4A7C15ull * (uint64_t)(3 * (uint64_t)i + t + 1));
        x[t] = (float)(z >> 40) * (1.0f / 16777216.0f);
      }
      if (std::fabs(nz[2]) > eps)
        x[2] = -(nz[0] * x[0] + nz[1] * x[1]) / nz[2];
      else if (std::fabs(nz[1]) > eps)
        x[1] = -(nz[0] * x[0] + nz[2] * x[2]) / nz[1];
      else if (std::fabs(nz[0]) > eps)
        x[0] = -(nz[1] * x[1] + nz[2] * x[2]) / nz[0];
      float xn = x[0] * x[0];
      xn = xn + x[1] * x[1];
      xn = xn + x[2] * x[2];
      const float inv = 1.0f / std::sqrt(xn);
      for (int a = 0; a < 3; ++a) x[a] = x[a] * inv;
      rf[0] = x[0]; rf[1] = x[1]; rf[2] = x[2];
      rf[3] = nz[1] * x[2] - nz[2] * x[1];
      rf[4] = nz[2] * x[0] - nz[0] * x[2];
      rf[5] = nz[0] * x[1] - nz[1] * x[0];
      rf[6] = nz[0]; rf[7] = nz[1]; rf[8] = nz[2];
    }
  }
  return 0;
}

// out: nq x 1980; frames_out (optional, nq x 9): the frames the descriptors were computed in (upstream returns zeros)
extern "C" int orc_sc3d1980(const float* surf, const float* normals4, int n, const float* q, int nq, double search_radius,
                            double min_radius, double density_radius, unsigned long long seed, float* out1980,
                            float* frames_out) {
  if (!(search_radius > 0) || !(min_radius > 0) || !(density_radius > 0) || search_radius < min_radius) return -1;
  std::vector<float> rf((size_t)std::max(nq, 1) * 9), rf2((size_t)std::max(nq, 1) * 9);
  if (orc_sc3d_frames(surf, normals4, n, q, nq, search_radius, seed, rf.data()) != 0) return -1;
  if (frames_out) std::memcpy(frames_out, rf.data(), (size_t)nq * 9 * sizeof(float));
  // the shared binning: USC with the frames given (local radius unused)
  return orc_usc1980(surf, n, q, nq, search_radius, min_radius, density_radius, 1.0, rf.data(), out1980, rf2.data(), nullptr);
}
