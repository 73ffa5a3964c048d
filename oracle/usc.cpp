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
