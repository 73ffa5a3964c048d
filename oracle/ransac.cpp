// ransac.cpp — CPU oracle: RANSAC correspondence rejection + rigid transform.  TEST INFRASTRUCTURE ONLY.
// PARITY UNPINNED.  Restates pcl::registration::CorrespondenceRejectorSampleConsensus<PointXYZRGB> as driven by
// the reference at features.h:282-297 (inlier threshold 0.015, 1000 iterations): RandomSampleConsensus over a
// SampleConsensusModelRegistration - 3 random correspondences -> rigid transform (SVD / Horn) -> inliers
// |T s_i - t_i|^2 < thr^2 -> keep the hypothesis with most inliers, adaptive iteration bound
// k = log(1 - 0.99) / log(1 - w^3), stop when iterations >= k or > max_iterations.
//
// Definitions where upstream cannot be pinned (each also in DESIGN.md):
//  * random samples: upstream shuffles an index vector with boost::mt19937 (fixed seed 12345); here hypothesis h
//    draws its three distinct correspondences from SplitMix64(seed + golden * (3 h + t + 1)) (see sample3);
//  * a degenerate sample (collinear or coincident points) counts as an iteration with no inliers (upstream
//    re-draws it, up to 10 x max_iterations times);
//  * the transform is fitted and applied in double (upstream: float Eigen / Umeyama);
//  * output correspondences keep their input order; the transform is the winning 3-point hypothesis (upstream
//    does the same unless setRefineModel(true), which the reference does not call).
// The rotation is found with Horn's quaternion method (largest eigenvector of the 4x4 profile matrix), an
// algorithm independent of the Kabsch / eigen(H^T H) construction the CUDA kernel uses.
#include "oracle_common.hpp"
#include "pcl_oracle.h"

namespace {

inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// three distinct indices in [0, n): draw from n, n-1, n-2 and skip over the ones already taken (ascending)
void sample3(uint64_t seed, int h, int n, int out[3]) {
  uint64_t z[3];
  for (int t = 0; t < 3; ++t) z[t] = splitmix64(seed + 0x9E3779B97F4A7C15ull * (uint64_t)(3 * (uint64_t)h + t + 1));
  int i0 = (int)(z[0] % (uint64_t)n);
  int i1 = (int)(z[1] % (uint64_t)(n - 1));
  if (i1 >= i0) ++i1;
  int i2 = (int)(z[2] % (uint64_t)(n - 2));
  int lo = std::min(i0, i1), hi = std::max(i0, i1);
  if (i2 >= lo) ++i2;
  if (i2 >= hi) ++i2;
  out[0] = i0; out[1] = i1; out[2] = i2;
}

// symmetric 4x4 Jacobi: eigenvector of the largest eigenvalue
void largest_eigvec4(double A[4][4], double v[4]) {
  double V[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) V[i][j] = i == j;
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0;
    for (int p = 0; p < 4; ++p)
      for (int q = p + 1; q < 4; ++q) off += std::fabs(A[p][q]);
    if (off == 0) break;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 4; ++q) {
        if (A[p][q] == 0) continue;
        double theta = (A[q][q] - A[p][p]) / (2 * A[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
        double c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < 4; ++k) {
          double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 4; ++k) {
          double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 4; ++k) {
          double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int best = 0;
  for (int i = 1; i < 4; ++i)
    if (A[i][i] > A[best][best]) best = i;
  for (int i = 0; i < 4; ++i) v[i] = V[i][best];
}

// rigid transform (rows of R, then t) that maps the three source points onto the three target points in the
// least-squares sense; false for degenerate (collinear / coincident) triples
bool fit3(const double s[3][3], const double t[3][3], double T[12]) {
  double cs[3] = {0, 0, 0}, ct[3] = {0, 0, 0};
  for (int i = 0; i < 3; ++i)
    for (int a = 0; a < 3; ++a) {
      cs[a] += s[i][a] / 3.0;
      ct[a] += t[i][a] / 3.0;
    }
  // degeneracy: the triangle areas of both triples against their edge lengths
  for (int side = 0; side < 2; ++side) {
    const double(*p)[3] = side ? t : s;
    double e1[3], e2[3];
    for (int a = 0; a < 3; ++a) {
      e1[a] = p[1][a] - p[0][a];
      e2[a] = p[2][a] - p[0][a];
    }
    double cr[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    double area2 = cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2];
    double l1 = e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2], l2 = e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2];
    if (!(area2 > 1e-12 * l1 * l2) || !(l1 > 0) || !(l2 > 0)) return false;
  }
  double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int i = 0; i < 3; ++i)
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) S[a][b] += (s[i][a] - cs[a]) * (t[i][b] - ct[b]);
  double N[4][4] = {
      {S[0][0] + S[1][1] + S[2][2], S[1][2] - S[2][1], S[2][0] - S[0][2], S[0][1] - S[1][0]},
      {S[1][2] - S[2][1], S[0][0] - S[1][1] - S[2][2], S[0][1] + S[1][0], S[2][0] + S[0][2]},
      {S[2][0] - S[0][2], S[0][1] + S[1][0], -S[0][0] + S[1][1] - S[2][2], S[1][2] + S[2][1]},
      {S[0][1] - S[1][0], S[2][0] + S[0][2], S[1][2] + S[2][1], -S[0][0] - S[1][1] + S[2][2]}};
  double q[4];
  largest_eigvec4(N, q);
  double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  double w = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
  double R[3][3] = {{1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)},
                    {2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)},
                    {2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)}};
  for (int a = 0; a < 3; ++a) {
    for (int b = 0; b < 3; ++b) T[4 * a + b] = R[a][b];
    T[4 * a + 3] = ct[a] - (R[a][0] * cs[0] + R[a][1] * cs[1] + R[a][2] * cs[2]);
  }
  return true;
}

}  // namespace

// src / tgt: keypoint clouds (n x 3); corr_q / corr_m: indices into them; out_keep: n_corr flags; T16 row-major 4x4
extern "C" int orc_ransac_reject(const float* src, int ns, const float* tgt, int nt, const int* corr_q, const int* corr_m,
                                 int n_corr, double threshold, int max_iterations, uint64_t seed, int* out_keep,
                                 float* T16, int* n_inliers, int* iterations, int* best_hypothesis) {
  (void)ns; (void)nt;
  for (int i = 0; i < 16; ++i) T16[i] = (i % 5 == 0) ? 1.f : 0.f;
  *n_inliers = n_corr;
  *iterations = 0;
  *best_hypothesis = -1;
  for (int i = 0; i < n_corr; ++i) out_keep[i] = 1;
  if (n_corr < 3) return 0;  // PCL: too few correspondences -> everything kept, identity
  const double thr2 = threshold * threshold;
  double k = 1.0;
  int best = -1, best_h = -1;
  double bestT[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  int h = 0;
  for (; (double)h < k; ++h) {
    int idx[3];
    sample3(seed, h, n_corr, idx);
    double s[3][3], t[3][3], T[12];
    for (int i = 0; i < 3; ++i)
      for (int a = 0; a < 3; ++a) {
        s[i][a] = src[3 * (size_t)corr_q[idx[i]] + a];
        t[i][a] = tgt[3 * (size_t)corr_m[idx[i]] + a];
      }
    int count = 0;
    if (fit3(s, t, T)) {
      for (int i = 0; i < n_corr; ++i) {
        const float* p = src + 3 * (size_t)corr_q[i];
        const float* g = tgt + 3 * (size_t)corr_m[i];
        double d2 = 0;
        for (int a = 0; a < 3; ++a) {
          double v = T[4 * a] * p[0] + T[4 * a + 1] * p[1] + T[4 * a + 2] * p[2] + T[4 * a + 3] - g[a];
          d2 += v * v;
        }
        if (d2 < thr2) ++count;
      }
    }
    if (count > best) {
      best = count;
      best_h = h;
      std::memcpy(bestT, T, sizeof(T));
      double w = (double)count / (double)n_corr;
      double p_no = 1.0 - w * w * w;
      p_no = std::max(std::numeric_limits<double>::epsilon(), p_no);
      p_no = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no);
      k = std::log(1.0 - 0.99) / std::log(p_no);
    }
    if (h + 1 > max_iterations) {
      ++h;
      break;
    }
  }
  *iterations = h;
  *best_hypothesis = best_h;
  if (best <= 0) {  // no hypothesis produced an inlier: nothing survives
    for (int i = 0; i < n_corr; ++i) out_keep[i] = 0;
    *n_inliers = 0;
    return 0;
  }
  int cnt = 0;
  for (int i = 0; i < n_corr; ++i) {
    const float* p = src + 3 * (size_t)corr_q[i];
    const float* g = tgt + 3 * (size_t)corr_m[i];
    double d2 = 0;
    for (int a = 0; a < 3; ++a) {
      double v = bestT[4 * a] * p[0] + bestT[4 * a + 1] * p[1] + bestT[4 * a + 2] * p[2] + bestT[4 * a + 3] - g[a];
      d2 += v * v;
    }
    out_keep[i] = d2 < thr2 ? 1 : 0;
    cnt += out_keep[i];
  }
  *n_inliers = cnt;
  for (int i = 0; i < 12; ++i) T16[i] = (float)bestT[i];
  T16[12] = T16[13] = T16[14] = 0.f;
  T16[15] = 1.f;
  return 0;
}
