// spin.cpp — CPU oracle: spin images, 8-bin image width = 9 x 17 = 153 values.  TEST INFRASTRUCTURE ONLY.
// PARITY UNPINNED.  Restates pcl::SpinImageEstimation<PointXYZRGB, Normal, Histogram<153>> with its defaults (image
// width 8, support angle cosine 0, rectangular image, rotation axis = the query's normal), as driven by the reference
// at evaluation.cpp:515-554 (normals estimated on the KEYPOINT cloud, search surface = the full cloud, radius r);
// upstream features/impl/spin_image.hpp, computeSiForPoint:
//   bin = r / 8 / sqrt(2); per neighbour: direction = p - origin, beta = |direction| cos(direction, axis),
//   alpha = |direction| sqrt(1 - cos^2); outside the cylinder (|beta| >= 8 bin or alpha >= 8 bin) -> skipped;
//   bilinear vote into the (alpha, beta + 8 bins) cell of a 9 x 17 double matrix; matrix / its sum when the query has
//   more than one neighbour.
// A query with a non-finite normal gets a NaN row (upstream throws on the first |cos| > 1 check).
#include "oracle_common.hpp"
#include "pcl_oracle.h"

using namespace orc;

// qnormals: nq x 4 (the normals of the QUERIES); out: nq x 153
extern "C" int orc_spin_image153(const float* surf, int n, const float* q, const float* qnormals4, int nq, double radius,
                                 float* out153) {
  if (!(radius > 0)) return -1;
  const int W = 8;
  const double bin_size = radius / W / std::sqrt(2.0);
  Searcher s;
  s.init(surf, n, radius, 0);
  const float nanv = std::numeric_limits<float>::quiet_NaN();
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 32)
    for (int i = 0; i < nq; ++i) {
      float* O = out153 + 153 * (size_t)i;
      const float* o = q + 3 * (size_t)i;
      const float* ax = qnormals4 + 4 * (size_t)i;
      if (!finite3(o) || !std::isfinite(ax[0]) || !std::isfinite(ax[1]) || !std::isfinite(ax[2])) {
        for (int b = 0; b < 153; ++b) O[b] = nanv;
        continue;
      }
      double M[9][17];
      for (auto& r : M)
        for (double& v : r) v = 0.0;
      nb.clear();
      s.query(o, nb);
      for (const Nbr& b : nb) {
        const float* p = surf + 3 * (size_t)b.idx;
        const float d[3] = {p[0] - o[0], p[1] - o[1], p[2] - o[2]};
        float n2 = d[0] * d[0];
        n2 = n2 + d[1] * d[1];
        n2 = n2 + d[2] * d[2];
        const double dn = (double)std::sqrt(n2);  // Eigen: float norm, promoted
        if (std::fabs(dn) < 10 * std::numeric_limits<double>::epsilon()) continue;
        float dot = d[0] * ax[0];
        dot = dot + d[1] * ax[1];
        dot = dot + d[2] * ax[2];
        double c = (double)dot / dn;
        c = std::max(-1.0, std::min(1.0, c));
        double beta = dn * c;
        double alpha = dn * std::sqrt(1.0 - c * c);
        if (std::fabs(beta) >= bin_size * W || alpha >= bin_size * W) continue;
        int beta_bin = (int)std::floor(beta / bin_size) + W;
        int alpha_bin = (int)std::floor(alpha / bin_size);
        if (alpha_bin == W) {
          alpha_bin--;
          alpha = bin_size * (alpha_bin + 1) - std::numeric_limits<double>::epsilon();
        }
        if (beta_bin == 2 * W) {
          beta_bin--;
          beta = bin_size * (beta_bin - W + 1) - std::numeric_limits<double>::epsilon();
        }
        const double a = alpha / bin_size - (double)alpha_bin;
        const double bb = beta / bin_size - (double)(beta_bin - W);
        M[alpha_bin][beta_bin] += (1 - a) * (1 - bb);
        M[alpha_bin + 1][beta_bin] += a * (1 - bb);
        M[alpha_bin][beta_bin + 1] += (1 - a) * bb;
        M[alpha_bin + 1][beta_bin + 1] += a * bb;
      }
      double sum = 0;
      for (auto& r : M)
        for (double v : r) sum += v;
      const bool norm = nb.size() > 1;
      for (int r = 0; r < 9; ++r)
        for (int cc = 0; cc < 17; ++cc) O[r * 17 + cc] = (float)(norm ? M[r][cc] / sum : M[r][cc]);
    }
  }
  return 0;
}
