// keypoints.cpp — CPU oracle: ISS3D and Harris3D keypoints.  TEST INFRASTRUCTURE ONLY.
// PARITY UNPINNED.  Restates pcl::ISSKeypoint3D<PointXYZRGB,PointXYZRGB> as configured at reference
// keypoints.h:182-196 (upstream keypoints/impl/iss_3d.hpp) and
// pcl::HarrisKeypoint3D<PointXYZRGB,PointXYZI> as configured at keypoints.h:150-164 (upstream
// keypoints/impl/harris_3d.hpp) plus the reference's own snap keypoints.h:360-395.
// SURVEY.md A.4, A.5.  Output order = ascending index (upstream order is OpenMP-nondeterministic).
#include "oracle_common.hpp"
#include "pcl_oracle.h"

using namespace orc;

// ------------------------------------------------------------------------------------ ISS
extern "C" int orc_iss_saliency(const float* pts, int n, double salient_radius, int min_neighbors,
                                double gamma21, double gamma32, double* saliency) {
  Grid g;
  g.build(pts, n, salient_radius);
  float r2f = (float)(salient_radius * salient_radius);
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 128)
    for (int i = 0; i < n; ++i) {
      saliency[i] = 0.0;
      const float* c = pts + 3 * (size_t)i;
      if (!finite3(c)) continue;
      g.radius(c, salient_radius, r2f, nb);
      double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      if ((int)nb.size() >= min_neighbors) {  // getScatterMatrix: about the point, un-normalised
        for (const Nbr& b : nb) {
          double d[3];
          for (int a = 0; a < 3; ++a) d[a] = (double)pts[3 * (size_t)b.idx + a] - (double)c[a];
          for (int r = 0; r < 3; ++r)
            for (int cc = 0; cc < 3; ++cc) S[r][cc] += d[r] * d[cc];
        }
      }
      double w[3], V[3][3];
      eigSym3(S, w, V);
      double e1 = w[2], e2 = w[1], e3 = w[0];
      if (!std::isfinite(e1) || !std::isfinite(e2) || !std::isfinite(e3)) continue;
      if (e3 < 0) continue;  // upstream warns and skips (stale-scratch quirk not reproduced)
      if ((e2 / e1 < gamma21) && (e3 / e2 < gamma32)) saliency[i] = e3;
    }
  }
  return 0;
}

extern "C" int orc_iss_nms(const float* pts, int n, const double* saliency, double nonmax_radius,
                           int min_neighbors, int* kp_idx, int* n_kp) {
  Grid g;
  g.build(pts, n, nonmax_radius);
  float r2f = (float)(nonmax_radius * nonmax_radius);
  std::vector<char> is_max(n, 0);
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 128)
    for (int i = 0; i < n; ++i) {
      const float* c = pts + 3 * (size_t)i;
      if (!(saliency[i] > 0.0) || !finite3(c)) continue;
      g.radius(c, nonmax_radius, r2f, nb);
      if ((int)nb.size() < min_neighbors) continue;
      bool mx = true;
      for (const Nbr& b : nb)
        if (saliency[i] < saliency[b.idx]) mx = false;
      is_max[i] = mx;
    }
  }
  int m = 0;
  for (int i = 0; i < n; ++i)
    if (is_max[i]) kp_idx[m++] = i;
  *n_kp = m;
  return 0;
}

extern "C" int orc_iss(const float* pts, int n, double salient_radius, double nonmax_radius,
                       int min_neighbors, double gamma21, double gamma32, int* kp_idx, int* n_kp,
                       double* saliency) {
  std::vector<double> tmp;
  if (!saliency) {
    tmp.resize(n);
    saliency = tmp.data();
  }
  orc_iss_saliency(pts, n, salient_radius, min_neighbors, gamma21, gamma32, saliency);
  return orc_iss_nms(pts, n, saliency, nonmax_radius, min_neighbors, kp_idx, n_kp);
}

// --------------------------------------------------------------------------------- Harris3D
extern "C" int orc_harris_response(const float* pts, const float* normals4, int n, double radius,
                                   float* response) {
  Grid g;
  g.build(pts, n, radius);
  float r2f = (float)(radius * radius);
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 128)
    for (int i = 0; i < n; ++i) {
      response[i] = 0.f;
      const float* c = pts + 3 * (size_t)i;
      if (!finite3(c)) continue;
      g.radius(c, radius, r2f, nb);
      // calculateNormalCovar: mean of n n^T over neighbours with finite normal_x
      float xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0;
      unsigned count = 0;
      for (const Nbr& b : nb) {
        const float* nr = normals4 + 4 * (size_t)b.idx;
        if (!std::isfinite(nr[0])) continue;
        xx += nr[0] * nr[0];
        xy += nr[0] * nr[1];
        xz += nr[0] * nr[2];
        yy += nr[1] * nr[1];
        yz += nr[1] * nr[2];
        zz += nr[2] * nr[2];
        ++count;
      }
      if (count > 0) {
        float fc = (float)count;
        xx /= fc; xy /= fc; xz /= fc; yy /= fc; yz /= fc; zz /= fc;
      }
      float trace = xx + yy + zz;
      if (trace != 0) {
        float det = xx * yy * zz + 2.0f * xy * xz * yz - xz * xz * yy - xy * xy * zz - yz * yz * xx;
        response[i] = 0.04f + det - 0.04f * trace * trace;
      }
    }
  }
  return 0;
}

extern "C" int orc_harris_nms(const float* pts, const float* response, int n, double radius,
                              float threshold, int* kp_idx, int* n_kp) {
  Grid g;
  g.build(pts, n, radius);
  float r2f = (float)(radius * radius);
  std::vector<char> is_max(n, 0);
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 128)
    for (int i = 0; i < n; ++i) {
      const float* c = pts + 3 * (size_t)i;
      if (!finite3(c) || !std::isfinite(response[i]) || response[i] < threshold) continue;
      g.radius(c, radius, r2f, nb);
      bool mx = true;
      for (const Nbr& b : nb)
        if (response[i] < response[b.idx]) {
          mx = false;
          break;
        }
      is_max[i] = mx;
    }
  }
  int m = 0;
  for (int i = 0; i < n; ++i)
    if (is_max[i]) kp_idx[m++] = i;
  *n_kp = m;
  return 0;
}

extern "C" int orc_harris_refine(const float* pts, const float* normals4, int n, double radius,
                                 float* corners, int nc) {
  Grid g;
  g.build(pts, n, radius);
  float r2f = (float)(radius * radius);
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 16)
    for (int ci = 0; ci < nc; ++ci) {
      float* cr = corners + 3 * (size_t)ci;
      unsigned iterations = 0;
      float diff;
      do {
        float NNT[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, NNTp[3] = {0, 0, 0};
        float corner[3] = {cr[0], cr[1], cr[2]};
        g.radius(corner, radius, r2f, nb);
        for (const Nbr& b : nb) {
          const float* nr = normals4 + 4 * (size_t)b.idx;
          if (!std::isfinite(nr[0])) continue;
          const float* p = pts + 3 * (size_t)b.idx;
          float nnT[9];
          for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) nnT[3 * r + c] = nr[r] * nr[c];
          for (int e = 0; e < 9; ++e) NNT[e] += nnT[e];
          for (int r = 0; r < 3; ++r)
            NNTp[r] += nnT[3 * r] * p[0] + nnT[3 * r + 1] * p[1] + nnT[3 * r + 2] * p[2];
        }
        // invert3x3SymMatrix (column-major coeff(): symmetric, so layout is irrelevant)
        float a = NNT[0], bq = NNT[1], c = NNT[2], d = NNT[4], e = NNT[5], f = NNT[8];
        float fd_ee = d * f - e * e;
        float ce_bf = c * e - bq * f;
        float be_cd = bq * e - c * d;
        float det = a * fd_ee + bq * ce_bf + c * be_cd;
        if (det != 0) {
          float inv[9] = {fd_ee,  ce_bf,         be_cd,
                          ce_bf,  a * f - c * c, bq * c - a * e,
                          be_cd,  bq * c - a * e, a * d - bq * bq};
          for (float& x : inv) x /= det;
          for (int r = 0; r < 3; ++r)
            cr[r] = inv[3 * r] * NNTp[0] + inv[3 * r + 1] * NNTp[1] + inv[3 * r + 2] * NNTp[2];
        }
        float dx = cr[0] - corner[0], dy = cr[1] - corner[1], dz = cr[2] - corner[2];
        diff = dx * dx + dy * dy + dz * dz;
      } while (diff > 1e-6 && ++iterations < 10);
    }
  }
  return 0;
}

// keypoints.h:374-394: 1-NN of each (finite) corner in the cloud, kept when the SQUARED distance
// is < max_d2 (1e-4 in the reference).  Ties: lowest index.
extern "C" int orc_snap_to_cloud(const float* pts, int n, const float* q, int nq, float max_d2,
                                 int* snapped_idx) {
  Grid g;
  g.build(pts, n, autoEdge(pts, n, 2));
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 64)
    for (int i = 0; i < nq; ++i) {
      snapped_idx[i] = -1;
      if (!finite3(q + 3 * (size_t)i)) continue;
      g.knn(q + 3 * (size_t)i, 1, nb);
      if (!nb.empty() && nb[0].d2 < max_d2) snapped_idx[i] = nb[0].idx;
    }
  }
  return 0;
}
