// harris6d.cpp — CPU oracle: Harris 6D keypoints.  TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
// Restates pcl::HarrisKeypoint6D<PointXYZRGB, PointXYZI> as driven by reference keypoints.h:166-179
// (setNonMaxSupression(true), setThreshold(1e-6), radius left at PCL's 0.01, refine left on; it is in the active
// detector list, evaluation.cpp:63-65); upstream keypoints/impl/harris_6d.hpp and
// features/impl/intensity_gradient.hpp (PCL 1.7.x).
//
//   intensity  = 0.00390625 (0.114 b + 0.5870 g + 0.2989 r)   (double expression, rounded once to float)
//   normals    = NormalEstimation at the detector radius (orc_normals, mode 0)
//   gradient   = IntensityGradientEstimation at the same radius: centroid and mean intensity of the neighbourhood,
//                A = sum d d^T, b = sum d (I - mean) over the neighbours (float, list order), x = A^-1 b by a
//                column-pivoting Householder QR, gradient = (I - n n^T) x; fewer than 3 neighbours -> NaN
//   keep       = squared length > 200 ? gradient / length : 0      (upstream's magic number)
//   response   = 4th smallest eigenvalue of the 6x6 mean of (n, g)(n, g)^T over the neighbours with a finite
//                normal and gradient (float sums in list order, times float(1.0 / count))
//   keypoints  = response >= threshold and no neighbour with a larger response; refineCorners; snap (keypoints.h)
//
// Definitions where upstream's result depends on Eigen internals that cannot be pinned here (each is followed
// operation for operation by the CUDA kernels, harris6d.cu):
//   * the QR: Eigen 3.2's ColPivHouseholderQR written out for 3x3 floats - pivot = column of largest remaining
//     squared norm (first on ties), squared norms recomputed for the pivot and down-dated for the others, a pivot
//     below max_col_sqnorm * eps^2 / 3 * (3 - k) ends the factorisation (rank k, remaining unknowns 0), Householder
//     vectors as makeHouseholder builds them; every dot product sequential;
//   * centroid /= n as a multiplication by float(1) / n (Eigen 3.2's operator/=), mean_intensity /= n as a division;
//   * the 6x6 eigenvalues: cyclic Jacobi in double on the float matrix (sweeps over (p, q), p < q, in row order; stop
//     when the off-diagonal mass is <= 1e-18 of the diagonal's or after 60 sweeps), sorted ascending, [3] rounded to
//     float.  (Upstream: SelfAdjointEigenSolver<Matrix<float, 6, 6>>, tridiagonal QL in float.)
#include "oracle_common.hpp"
#include "pcl_oracle.h"

using namespace orc;

namespace {

// x = A^-1 b (A symmetric 3x3 given in full), Eigen 3.2 ColPivHouseholderQR<Matrix3f>::solve
void colPivQrSolve3(const float Ain[3][3], const float bin[3], float x[3]) {
  float qr[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) qr[r][c] = Ain[r][c];
  float hco[3] = {0, 0, 0};
  int transp[3] = {0, 1, 2};
  float colsq[3];
  // (a fixed-size column: Eigen unrolls the reduction as a + (b + c))
  for (int k = 0; k < 3; ++k) colsq[k] = qr[0][k] * qr[0][k] + (qr[1][k] * qr[1][k] + qr[2][k] * qr[2][k]);
  const float eps = std::numeric_limits<float>::epsilon();
  const float thr_helper = std::max(colsq[0], std::max(colsq[1], colsq[2])) * (eps * eps) / 3.0f;
  int nonzero = 3;
  for (int k = 0; k < 3; ++k) {
    int big = k;
    for (int c = k + 1; c < 3; ++c)
      if (colsq[c] > colsq[big]) big = c;
    float bsq = 0.f;
    for (int r = k; r < 3; ++r) bsq = (r == k) ? qr[r][big] * qr[r][big] : bsq + qr[r][big] * qr[r][big];
    colsq[big] = bsq;
    if (bsq < thr_helper * (float)(3 - k)) {
      nonzero = k;
      break;
    }
    transp[k] = big;
    if (k != big) {
      for (int r = 0; r < 3; ++r) std::swap(qr[r][k], qr[r][big]);
      std::swap(colsq[k], colsq[big]);
    }
    // makeHouseholderInPlace on qr[k..2][k]
    float tailsq = 0.f;
    for (int r = k + 1; r < 3; ++r) tailsq = (r == k + 1) ? qr[r][k] * qr[r][k] : tailsq + qr[r][k] * qr[r][k];
    const float c0 = qr[k][k];
    float tau, beta;
    if (k == 2 || tailsq == 0.f) {
      tau = 0.f;
      beta = c0;
      for (int r = k + 1; r < 3; ++r) qr[r][k] = 0.f;
    } else {
      beta = std::sqrt(c0 * c0 + tailsq);
      if (c0 >= 0.f) beta = -beta;
      const float den = c0 - beta;
      for (int r = k + 1; r < 3; ++r) qr[r][k] = qr[r][k] / den;
      tau = (beta - c0) / beta;
    }
    hco[k] = tau;
    qr[k][k] = beta;
    // applyHouseholderOnTheLeft to the trailing columns
    if (k < 2) {
      for (int c = k + 1; c < 3; ++c) {
        float tmp = 0.f;
        for (int r = k + 1; r < 3; ++r) tmp = (r == k + 1) ? qr[r][k] * qr[r][c] : tmp + qr[r][k] * qr[r][c];
        tmp = tmp + qr[k][c];
        qr[k][c] = qr[k][c] - tau * tmp;
        for (int r = k + 1; r < 3; ++r) qr[r][c] = qr[r][c] - (tau * qr[r][k]) * tmp;
      }
    } else {
      // a 1 x 0 block: nothing to apply
    }
    for (int c = k + 1; c < 3; ++c) colsq[c] = colsq[c] - qr[k][c] * qr[k][c];
  }
  // column permutation: identity with the transpositions applied on the right, k = 0 .. nonzero-1
  int perm[3] = {0, 1, 2};
  for (int k = 0; k < nonzero; ++k) std::swap(perm[k], perm[transp[k]]);
  x[0] = x[1] = x[2] = 0.f;
  if (nonzero == 0) return;
  // c = Q^T b : apply H_0, H_1, ... in turn
  float c[3] = {bin[0], bin[1], bin[2]};
  for (int k = 0; k < nonzero; ++k) {
    if (k == 2) {
      c[2] = c[2] * (1.0f - hco[2]);
      continue;
    }
    float tmp = 0.f;
    for (int r = k + 1; r < 3; ++r) tmp = (r == k + 1) ? qr[r][k] * c[r] : tmp + qr[r][k] * c[r];
    tmp = tmp + c[k];
    c[k] = c[k] - hco[k] * tmp;
    for (int r = k + 1; r < 3; ++r) c[r] = c[r] - (hco[k] * qr[r][k]) * tmp;
  }
  // back substitution on the leading nonzero x nonzero upper triangle
  for (int i = nonzero - 1; i >= 0; --i) {
    float s = c[i];
    for (int j = i + 1; j < nonzero; ++j) s = s - qr[i][j] * c[j];
    c[i] = s / qr[i][i];
  }
  for (int i = 0; i < nonzero; ++i) x[perm[i]] = c[i];
}

// eigenvalues (ascending) of a symmetric 6x6 given in full, cyclic Jacobi in double
void eigvalsSym6(double A[6][6], double w[6]) {
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0, dg = 0;
    for (int p = 0; p < 6; ++p) {
      dg += std::fabs(A[p][p]);
      for (int q = p + 1; q < 6; ++q) off += std::fabs(A[p][q]);
    }
    if (off <= 1e-300 || off <= 1e-18 * dg) break;
    for (int p = 0; p < 5; ++p)
      for (int q = p + 1; q < 6; ++q) {
        if (A[p][q] == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int r = 0; r < 6; ++r) {  // A <- A J
          const double arp = A[r][p], arq = A[r][q];
          A[r][p] = c * arp - s * arq;
          A[r][q] = s * arp + c * arq;
        }
        for (int r = 0; r < 6; ++r) {  // A <- J^T A
          const double apr = A[p][r], aqr = A[q][r];
          A[p][r] = c * apr - s * aqr;
          A[q][r] = s * apr + c * aqr;
        }
      }
  }
  for (int i = 0; i < 6; ++i) w[i] = A[i][i];
  for (int i = 1; i < 6; ++i) {  // insertion sort
    const double v = w[i];
    int j = i - 1;
    while (j >= 0 && w[j] > v) {
      w[j + 1] = w[j];
      --j;
    }
    w[j + 1] = v;
  }
}

}  // namespace

// rgb: packed 0x00RRGGBB per point.  normals4: n x 4 (NormalEstimation at `radius`).  gradients_out (optional): n x 3
// after the length rule; intensity_out (optional): n.
extern "C" int orc_harris6d_response(const float* pts, const uint32_t* rgb, const float* normals4, int n, double radius,
                                     float* response, float* gradients_out, float* intensity_out) {
  if (!(radius > 0)) return -1;
  Grid g;
  g.build(pts, n, radius);
  const float r2f = (float)(radius * radius);
  const float nanv = std::numeric_limits<float>::quiet_NaN();
  std::vector<float> inten(std::max(n, 1)), grad((size_t)std::max(n, 1) * 3);
  for (int i = 0; i < n; ++i) {
    const uint32_t c = rgb[i];
    const float r = (float)((c >> 16) & 255u), gch = (float)((c >> 8) & 255u), b = (float)(c & 255u);
    inten[i] = (float)(0.00390625 * (0.114 * b + 0.5870 * gch + 0.2989 * r));
  }
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 128)
    for (int i = 0; i < n; ++i) {
      float* G = grad.data() + 3 * (size_t)i;
      const float* q = pts + 3 * (size_t)i;
      g.radius(q, radius, r2f, nb);
      if (nb.empty()) {  // searchForNeighbors found nothing (non-finite point)
        G[0] = G[1] = G[2] = nanv;
        continue;
      }
      float cen[3] = {0, 0, 0}, mean_i = 0.f;
      for (const Nbr& b : nb) {
        const float* p = pts + 3 * (size_t)b.idx;
        cen[0] += p[0];
        cen[1] += p[1];
        cen[2] += p[2];
        mean_i += inten[b.idx];
      }
      const float fn = (float)nb.size();
      const float inv_n = 1.0f / fn;
      cen[0] *= inv_n;
      cen[1] *= inv_n;
      cen[2] *= inv_n;
      mean_i /= fn;
      if (nb.size() < 3) {
        G[0] = G[1] = G[2] = nanv;
        continue;
      }
      float A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, bv[3] = {0, 0, 0};
      for (const Nbr& b : nb) {
        const float* p = pts + 3 * (size_t)b.idx;
        if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2]) || !std::isfinite(inten[b.idx])) continue;
        const float dx = p[0] - cen[0], dy = p[1] - cen[1], dz = p[2] - cen[2];
        const float di = inten[b.idx] - mean_i;
        A[0][0] += dx * dx;
        A[0][1] += dx * dy;
        A[0][2] += dx * dz;
        A[1][1] += dy * dy;
        A[1][2] += dy * dz;
        A[2][2] += dz * dz;
        bv[0] += dx * di;
        bv[1] += dy * di;
        bv[2] += dz * di;
      }
      A[1][0] = A[0][1];
      A[2][0] = A[0][2];
      A[2][1] = A[1][2];
      float x[3];
      colPivQrSolve3(A, bv, x);
      const float* nr = normals4 + 4 * (size_t)i;
      // (I - n n^T) x, rows evaluated left to right
      for (int r = 0; r < 3; ++r) {
        float acc = 0.f;
        for (int c = 0; c < 3; ++c) {
          const float m = (r == c ? 1.0f : 0.0f) - nr[r] * nr[c];
          acc = (c == 0) ? m * x[c] : acc + m * x[c];
        }
        G[r] = acc;
      }
    }
#pragma omp for schedule(static)
    for (int i = 0; i < n; ++i) {
      float* G = grad.data() + 3 * (size_t)i;
      float len = G[0] * G[0] + G[1] * G[1] + G[2] * G[2];
      if ((double)len > 200.0) {
        len = (float)(1.0 / std::sqrt((double)len));
        G[0] *= len;
        G[1] *= len;
        G[2] *= len;
      } else {
        G[0] = G[1] = G[2] = 0.f;
      }
    }
#pragma omp for schedule(dynamic, 128)
    for (int i = 0; i < n; ++i) {
      response[i] = 0.f;
      const float* q = pts + 3 * (size_t)i;
      if (!finite3(q)) continue;
      g.radius(q, radius, r2f, nb);
      float co[21];
      for (float& v : co) v = 0.f;
      unsigned count = 0;
      for (const Nbr& b : nb) {
        const float* nr = normals4 + 4 * (size_t)b.idx;
        const float* gr = grad.data() + 3 * (size_t)b.idx;
        if (!std::isfinite(nr[0]) || !std::isfinite(gr[0])) continue;
        const float v[6] = {nr[0], nr[1], nr[2], gr[0], gr[1], gr[2]};
        int t = 0;
        for (int a = 0; a < 6; ++a)
          for (int c = a; c < 6; ++c) co[t++] += v[a] * v[c];
        ++count;
      }
      if (count > 0) {
        const float norm = (float)(1.0 / (double)(float)count);
        for (float& v : co) v *= norm;
      }
      // diagonal entries sit at 0, 6, 11, 15, 18, 20
      const float trace = co[0] + co[6] + co[11] + co[15] + co[18] + co[20];
      if (trace != 0) {
        double M[6][6];
        int t = 0;
        for (int a = 0; a < 6; ++a)
          for (int c = a; c < 6; ++c) {
            M[a][c] = (double)co[t];
            M[c][a] = (double)co[t];
            ++t;
          }
        double w[6];
        eigvalsSym6(M, w);
        response[i] = (float)w[3];
      }
    }
  }
  if (gradients_out) std::memcpy(gradients_out, grad.data(), (size_t)n * 3 * sizeof(float));
  if (intensity_out) std::memcpy(intensity_out, inten.data(), (size_t)n * sizeof(float));
  return 0;
}
