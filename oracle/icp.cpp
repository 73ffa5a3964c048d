// icp.cpp — CPU oracle: point-to-point ICP.  TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
// Restates pcl::IterativeClosestPoint<PointXYZRGB, PointXYZRGB> (PCL 1.7.x registration/impl/icp.hpp) as driven by
// the reference at evaluation.cpp:863-885 (max correspondence distance 0.07, transformation epsilon 1e-6,
// euclidean fitness epsilon 1e-4, 100 iterations; the RANSAC outlier threshold it also sets is not read by
// PCL 1.7's ICP, whose rejector list is empty by default):
//
//   final = guess; cur = guess * source; prev_mse = DBL_MAX
//   repeat
//     correspondences = { (i, nn(cur_i)) : |cur_i - nn|^2 <= max_dist^2 }          CorrespondenceEstimation
//     fewer than 3 -> state NO_CORRESPONDENCES, not converged, stop
//     T = least-squares rigid transform cur -> target over them                     TransformationEstimationSVD (Umeyama)
//     cur = T * cur (in place, float); final = T * final; ++iterations
//     converged = DefaultConvergenceCriteria(iterations, T, correspondences)
//   until converged
//   fitness = mean over ALL source points of |final * s - nn|^2                     getFitnessScore()
//
// DefaultConvergenceCriteria::hasConverged, in this order: (1) iterations >= max_iterations -> ITERATIONS;
// (2) cos_angle = 0.5 (trace R - 1) >= 1 - eps_T and |t|^2 <= eps_T -> TRANSFORM; (3) mse = mean correspondence d2,
// |mse - prev| < 1e-12 -> ABS_MSE; |mse - prev| / prev < eps_fit -> REL_MSE; else prev = mse.
//
// Definitions where upstream cannot be pinned (each also in DESIGN.md):
//  * Eigen's umeyama sums the demeaned outer products in float in an unspecified (vectorised) order and runs a
//    float JacobiSVD; here the sums are sequential doubles and the rotation comes from Horn's quaternion method
//    in double (the same optimum: a proper rotation, reflection folded onto the weakest singular direction),
//    rounded to float when stored into the 4x4 - an algorithm independent of the CUDA kernel's Kabsch solve;
//  * a float 4x4 * point product is ((m0 x + m1 y) + m2 z) + m3, no FMA; 4x4 * 4x4 sums k = 0..3 in that order;
//  * nearest neighbour ties resolve to the lowest target index; non-finite source points never correspond.
#include "oracle_common.hpp"
#include "pcl_oracle.h"

namespace {

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

inline void xform(const float* M, const float* p, float* o) {
  for (int r = 0; r < 3; ++r) {
    float v = M[4 * r] * p[0];
    v = v + M[4 * r + 1] * p[1];
    v = v + M[4 * r + 2] * p[2];
    v = v + M[4 * r + 3];
    o[r] = v;
  }
}

inline void mul44(const float* A, const float* B, float* C) {
  float R[16];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float s = A[4 * i] * B[j];
      for (int k = 1; k < 4; ++k) s = s + A[4 * i + k] * B[4 * k + j];
      R[4 * i + j] = s;
    }
  std::memcpy(C, R, sizeof(R));
}

}  // namespace

// state: 0 not converged, 1 ITERATIONS, 2 TRANSFORM, 3 ABS_MSE, 4 REL_MSE, 5 NO_CORRESPONDENCES
extern "C" int orc_icp(const float* src, int ns, const float* tgt, int nt, double max_corr_dist, int max_iterations,
                       double transformation_epsilon, double euclidean_fitness_epsilon, const float* guess16,
                       float* T16, double* fitness, int* converged, int* iterations, int* state) {
  float fin[16];
  for (int i = 0; i < 16; ++i) fin[i] = guess16 ? guess16[i] : ((i % 5 == 0) ? 1.f : 0.f);
  bool guess_is_identity = true;
  for (int i = 0; i < 16; ++i) guess_is_identity = guess_is_identity && fin[i] == ((i % 5 == 0) ? 1.f : 0.f);
  std::vector<float> cur((size_t)ns * 3);
  for (int i = 0; i < ns; ++i) {
    if (guess_is_identity || !orc::finite3(src + 3 * (size_t)i)) std::memcpy(&cur[3 * (size_t)i], src + 3 * (size_t)i, 12);
    else xform(fin, src + 3 * (size_t)i, &cur[3 * (size_t)i]);
  }
  std::vector<int> nn(std::max(ns, 1));
  std::vector<float> d2(std::max(ns, 1));
  const double max_d2 = max_corr_dist * max_corr_dist;
  const double rot_thr = 1.0 - transformation_epsilon, tr_thr = transformation_epsilon;
  double prev_mse = std::numeric_limits<double>::max();
  int it = 0, st = 0;
  bool conv = false;
  do {
    orc_knn(tgt, nt, cur.data(), ns, 1, nn.data(), d2.data());
    // correspondences + Umeyama sums
    long long cnt = 0;
    double ms[3] = {0, 0, 0}, mt[3] = {0, 0, 0}, mse = 0;
    for (int i = 0; i < ns; ++i) {
      if (nn[i] < 0 || !orc::finite3(&cur[3 * (size_t)i]) || (double)d2[i] > max_d2) {
        nn[i] = -1;
        continue;
      }
      ++cnt;
      for (int a = 0; a < 3; ++a) {
        ms[a] += cur[3 * (size_t)i + a];
        mt[a] += tgt[3 * (size_t)nn[i] + a];
      }
      mse += d2[i];
    }
    if (cnt < 3) {
      st = 5;
      conv = false;
      break;
    }
    for (int a = 0; a < 3; ++a) {
      ms[a] /= (double)cnt;
      mt[a] /= (double)cnt;
    }
    mse /= (double)cnt;
    double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};  // sum (s - ms)(t - mt)^T
    for (int i = 0; i < ns; ++i) {
      if (nn[i] < 0) continue;
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
          S[a][b] += ((double)cur[3 * (size_t)i + a] - ms[a]) * ((double)tgt[3 * (size_t)nn[i] + b] - mt[b]);
    }
    double N[4][4] = {
        {S[0][0] + S[1][1] + S[2][2], S[1][2] - S[2][1], S[2][0] - S[0][2], S[0][1] - S[1][0]},
        {S[1][2] - S[2][1], S[0][0] - S[1][1] - S[2][2], S[0][1] + S[1][0], S[2][0] + S[0][2]},
        {S[2][0] - S[0][2], S[0][1] + S[1][0], -S[0][0] + S[1][1] - S[2][2], S[1][2] + S[2][1]},
        {S[0][1] - S[1][0], S[2][0] + S[0][2], S[1][2] + S[2][1], -S[0][0] - S[1][1] + S[2][2]}};
    double q[4];
    largest_eigvec4(N, q);
    double qn = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    double w = q[0] / qn, x = q[1] / qn, y = q[2] / qn, z = q[3] / qn;
    double R[3][3] = {{1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)},
                      {2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)},
                      {2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)}};
    float T[16] = {0};
    for (int a = 0; a < 3; ++a) {
      for (int b = 0; b < 3; ++b) T[4 * a + b] = (float)R[a][b];
      T[4 * a + 3] = (float)(mt[a] - (R[a][0] * ms[0] + R[a][1] * ms[1] + R[a][2] * ms[2]));
    }
    T[15] = 1.f;
    for (int i = 0; i < ns; ++i) {
      float o[3];
      if (!orc::finite3(&cur[3 * (size_t)i])) continue;
      xform(T, &cur[3 * (size_t)i], o);
      std::memcpy(&cur[3 * (size_t)i], o, 12);
    }
    mul44(T, fin, fin);
    ++it;
    // DefaultConvergenceCriteria::hasConverged
    st = 0;
    conv = false;
    if (it >= max_iterations) {
      st = 1;
      conv = true;
    } else {
      float tr = T[0] + T[5];
      tr = tr + T[10];
      tr = tr - 1.f;
      const double cos_angle = 0.5 * (double)tr;
      float t2 = T[3] * T[3];
      t2 = t2 + T[7] * T[7];
      t2 = t2 + T[11] * T[11];
      if (cos_angle >= rot_thr && (double)t2 <= tr_thr) {
        st = 2;
        conv = true;
      } else if (std::fabs(mse - prev_mse) < 1e-12) {
        st = 3;
        conv = true;
      } else if (std::fabs(mse - prev_mse) / prev_mse < euclidean_fitness_epsilon) {
        st = 4;
        conv = true;
      } else {
        prev_mse = mse;
      }
    }
  } while (!conv);
  std::memcpy(T16, fin, sizeof(fin));
  *converged = conv ? 1 : 0;
  *iterations = it;
  *state = st;
  // getFitnessScore(): every (finite) source point, no distance bound
  if (fitness) {
    std::vector<float> moved((size_t)ns * 3);
    for (int i = 0; i < ns; ++i) {
      if (orc::finite3(src + 3 * (size_t)i)) xform(fin, src + 3 * (size_t)i, &moved[3 * (size_t)i]);
      else std::memcpy(&moved[3 * (size_t)i], src + 3 * (size_t)i, 12);
    }
    orc_knn(tgt, nt, moved.data(), ns, 1, nn.data(), d2.data());
    double sum = 0;
    long long nr = 0;
    for (int i = 0; i < ns; ++i)
      if (nn[i] >= 0 && orc::finite3(&moved[3 * (size_t)i])) {
        sum += d2[i];
        ++nr;
      }
    *fitness = nr > 0 ? sum / (double)nr : std::numeric_limits<double>::max();
  }
  return 0;
}
