// ransac.cu — RANSAC correspondence rejection + rigid transform (the step right after the
// hot path; replaces pcl::registration::CorrespondenceRejectorSampleConsensus as driven by the reference at
// features.h:282-297: inlier threshold 0.015, 1000 iterations).
//
// PCL evaluates its hypotheses one after the other and stops when the iteration count reaches the adaptive bound
// k = log(1 - 0.99) / log(1 - w^3).  Here all max_iterations + 1 hypotheses are evaluated AT ONCE - one warp per
// hypothesis: lane 0 draws the three correspondences and fits the rigid transform (Kabsch: eigen decomposition of
// H^T H in double), the warp counts the inliers |T s_i - t_i|^2 < thr^2 - and the host then replays PCL's
// sequential rule over the 1001 counts, so the result is the one the sequential loop would have produced.
// The sampling contract (SplitMix64 per hypothesis, no re-draw of degenerate samples) is stated in DESIGN.md §3;
// upstream's boost::mt19937 index shuffle cannot be pinned.
#include <cmath>
#include <limits>

#include "internal.h"

namespace pfx {

__host__ __device__ inline uint64_t rs_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__device__ __forceinline__ void rs_sample3(uint64_t seed, int h, int n, int out[3]) {
  uint64_t z[3];
#pragma unroll
  for (int t = 0; t < 3; ++t) z[t] = rs_splitmix64(seed + 0x9E3779B97F4A7C15ull * (uint64_t)(3 * (uint64_t)h + t + 1));
  int i0 = (int)(z[0] % (uint64_t)n);
  int i1 = (int)(z[1] % (uint64_t)(n - 1));
  if (i1 >= i0) ++i1;
  int i2 = (int)(z[2] % (uint64_t)(n - 2));
  const int lo = min(i0, i1), hi = max(i0, i1);
  if (i2 >= lo) ++i2;
  if (i2 >= hi) ++i2;
  out[0] = i0; out[1] = i1; out[2] = i2;
}

// Kabsch on three point pairs: H = sum (s - cs)(t - ct)^T = U S V^T, R = V U^T.  V from eigen(H^T H); the three
// centred points span a plane, so the third singular value is 0: u2 = u0 x u1 and v2 = v0 x v1 complete two
// right-handed bases and R is a proper rotation by construction.
__device__ bool rs_fit3(const double s[3][3], const double t[3][3], double T[12]) {
  double cs[3] = {0, 0, 0}, ct[3] = {0, 0, 0};
  for (int i = 0; i < 3; ++i)
    for (int a = 0; a < 3; ++a) {
      cs[a] += s[i][a] / 3.0;
      ct[a] += t[i][a] / 3.0;
    }
  for (int side = 0; side < 2; ++side) {
    const double(*p)[3] = side ? t : s;
    double e1[3], e2[3];
    for (int a = 0; a < 3; ++a) {
      e1[a] = p[1][a] - p[0][a];
      e2[a] = p[2][a] - p[0][a];
    }
    const double cr[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    const double area2 = cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2];
    const double l1 = e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2], l2 = e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2];
    if (!(area2 > 1e-12 * l1 * l2) || !(l1 > 0) || !(l2 > 0)) return false;
  }
  double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int i = 0; i < 3; ++i)
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) H[a][b] += (s[i][a] - cs[a]) * (t[i][b] - ct[b]);
  // here H = sum s t^T, so R s ~ t  <=>  R = V U^T with H = U S V^T; eigen(H^T H) gives V
  double hth[6];
  {
    double M[3][3];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) M[a][b] = H[0][a] * H[0][b] + H[1][a] * H[1][b] + H[2][a] * H[2][b];
    hth[0] = M[0][0]; hth[1] = M[0][1]; hth[2] = M[0][2]; hth[3] = M[1][1]; hth[4] = M[1][2]; hth[5] = M[2][2];
  }
  double w[3], v[3][3];
  eig_sym3<double>(hth, w, v, 40);  // ascending eigenvalues, eigenvectors in columns
  double v0[3] = {v[0][2], v[1][2], v[2][2]}, v1[3] = {v[0][1], v[1][1], v[2][1]};
  const double s0 = sqrt(fmax(w[2], 0.0)), s1 = sqrt(fmax(w[1], 0.0));
  if (!(s1 > 1e-14 * s0) || !(s0 > 0)) return false;
  double u0[3], u1[3];
  for (int a = 0; a < 3; ++a) {
    u0[a] = (H[a][0] * v0[0] + H[a][1] * v0[1] + H[a][2] * v0[2]) / s0;
    u1[a] = (H[a][0] * v1[0] + H[a][1] * v1[1] + H[a][2] * v1[2]) / s1;
  }
  // re-orthonormalise u1 against u0 (they are orthogonal up to round-off)
  const double d01 = u0[0] * u1[0] + u0[1] * u1[1] + u0[2] * u1[2];
  for (int a = 0; a < 3; ++a) u1[a] -= d01 * u0[a];
  const double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
  for (int a = 0; a < 3; ++a) u1[a] /= n1;
  const double u2[3] = {u0[1] * u1[2] - u0[2] * u1[1], u0[2] * u1[0] - u0[0] * u1[2], u0[0] * u1[1] - u0[1] * u1[0]};
  const double v2[3] = {v0[1] * v1[2] - v0[2] * v1[1], v0[2] * v1[0] - v0[0] * v1[2], v0[0] * v1[1] - v0[1] * v1[0]};
  // with H = sum s t^T: columns of U live in source space, columns of V in target space -> R = V U^T
  for (int a = 0; a < 3; ++a) {
    for (int b = 0; b < 3; ++b) T[4 * a + b] = v0[a] * u0[b] + v1[a] * u1[b] + v2[a] * u2[b];
  }
  for (int a = 0; a < 3; ++a) T[4 * a + 3] = ct[a] - (T[4 * a] * cs[0] + T[4 * a + 1] * cs[1] + T[4 * a + 2] * cs[2]);
  return true;
}

__device__ __forceinline__ bool rs_is_inlier(const double* T, const float* p, const float* g, double thr2) {
  double d2 = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double v = T[4 * a] * (double)p[0] + T[4 * a + 1] * (double)p[1] + T[4 * a + 2] * (double)p[2] + T[4 * a + 3] - (double)g[a];
    d2 += v * v;
  }
  return d2 < thr2;
}

// one warp per hypothesis; counts[h] = inliers (0 for a degenerate sample), transforms[h] = 12 doubles
__global__ void __launch_bounds__(128)
ransac_hypotheses_kernel(const float* __restrict__ src, size_t stride_s, const float* __restrict__ tgt, size_t stride_t,
                         const pfx_correspondence* __restrict__ corr, int n_corr, double thr2, uint64_t seed, int n_hyp,
                         int* __restrict__ counts, double* __restrict__ transforms) {
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (h >= n_hyp) return;
  double T[12];
  int ok = 0;
  if (lane == 0) {
    int idx[3];
    rs_sample3(seed, h, n_corr, idx);
    double s[3][3], t[3][3];
    for (int i = 0; i < 3; ++i) {
      const pfx_correspondence c = corr[idx[i]];
      const float* p = reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(src) + (size_t)c.index_query * stride_s);
      const float* g = reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(tgt) + (size_t)c.index_match * stride_t);
      for (int a = 0; a < 3; ++a) {
        s[i][a] = p[a];
        t[i][a] = g[a];
      }
    }
    ok = rs_fit3(s, t, T) ? 1 : 0;
  }
  ok = __shfl_sync(FULL, ok, 0);
#pragma unroll
  for (int i = 0; i < 12; ++i) T[i] = __shfl_sync(FULL, T[i], 0);
  int cnt = 0;
  if (ok) {
    for (int i = lane; i < n_corr; i += 32) {
      const pfx_correspondence c = corr[i];
      const float* p = reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(src) + (size_t)c.index_query * stride_s);
      const float* g = reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(tgt) + (size_t)c.index_match * stride_t);
      cnt += rs_is_inlier(T, p, g, thr2) ? 1 : 0;
    }
    cnt = warp_sum(cnt);
  }
  if (lane == 0) {
    counts[h] = cnt;
    for (int i = 0; i < 12; ++i) transforms[(size_t)h * 12 + i] = ok ? T[i] : 0.0;
  }
}

__global__ void ransac_flags_kernel(const float* __restrict__ src, size_t stride_s, const float* __restrict__ tgt,
                                    size_t stride_t, const pfx_correspondence* __restrict__ corr, int n_corr, double thr2,
                                    const double* __restrict__ T, int* __restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_corr) return;
  const pfx_correspondence c = corr[i];
  const float* p = reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(src) + (size_t)c.index_query * stride_s);
  const float* g = reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(tgt) + (size_t)c.index_match * stride_t);
  flags[i] = rs_is_inlier(T, p, g, thr2) ? 1 : 0;
}

__global__ void ransac_gather_kernel(const pfx_correspondence* __restrict__ corr, const int* __restrict__ idx, int n,
                                     pfx_correspondence* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = corr[idx[i]];
}

// all pointers device; out_dev has room for n_corr entries.  Returns the kept count, the transform (row-major
// 4x4), the number of iterations PCL's loop would have run and the winning hypothesis.
int ransac_reject_run(Ctx* ctx, const float* src, size_t stride_s, const float* tgt, size_t stride_t,
                      const pfx_correspondence* corr, int n_corr, double threshold, int max_iterations, uint64_t seed,
                      pfx_correspondence* out_dev, int* n_out, float* T16_host, int* iterations, int* best_h) {
  for (int i = 0; i < 16; ++i) T16_host[i] = (i % 5 == 0) ? 1.f : 0.f;
  *iterations = 0;
  *best_h = -1;
  *n_out = n_corr;
  if (n_corr < 3) {  // PCL: too few correspondences -> everything kept, identity transform
    if (n_corr > 0)
      PFX_CUDA(cudaMemcpyAsync(out_dev, corr, (size_t)n_corr * sizeof(pfx_correspondence), cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
  }
  const int n_hyp = max_iterations + 1;
  const double thr2 = threshold * threshold;
  PFX_CUDA(ctx->tmp0.ensure((size_t)n_hyp * (sizeof(int) + 12 * sizeof(double)) + 64));
  double* dT = ctx->tmp0.as<double>();
  int* dcounts = reinterpret_cast<int*>(dT + (size_t)n_hyp * 12);
  PFX_LAUNCH(ctx, ransac_hypotheses_kernel, div_up(n_hyp, 4), 128, 0, src, stride_s, tgt, stride_t, corr, n_corr, thr2, seed,
             n_hyp, dcounts, dT);
  PFX_CUDA(cudaGetLastError());
  std::vector<int> counts(n_hyp);
  PFX_CUDA(cudaMemcpyAsync(counts.data(), dcounts, (size_t)n_hyp * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  // RandomSampleConsensus::computeModel's sequential rule over the precomputed counts
  double k = 1.0;
  int best = -1, bh = -1, h = 0;
  for (; (double)h < k; ++h) {
    const int c = counts[h];
    if (c > best) {
      best = c;
      bh = h;
      const double w = (double)c / (double)n_corr;
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
  *best_h = bh;
  if (best <= 0) {
    *n_out = 0;
    return 0;
  }
  double Th[12];
  PFX_CUDA(cudaMemcpyAsync(Th, dT + (size_t)bh * 12, sizeof(Th), cudaMemcpyDeviceToHost, ctx->stream));
  PFX_CUDA(ctx->tmp2.ensure((size_t)n_corr * sizeof(int)));
  PFX_CUDA(ctx->tmp3.ensure((size_t)n_corr * sizeof(int)));
  PFX_LAUNCH(ctx, ransac_flags_kernel, div_up(n_corr, 256), 256, 0, src, stride_s, tgt, stride_t, corr, n_corr, thr2,
             dT + (size_t)bh * 12, ctx->tmp2.as<int>());
  int cnt = 0;
  PFX_TRY(compact_flags(ctx, ctx->tmp2.as<int>(), n_corr, ctx->tmp3.as<int>(), &cnt));  // synchronises
  if (cnt > 0) PFX_LAUNCH(ctx, ransac_gather_kernel, div_up(cnt, 256), 256, 0, corr, ctx->tmp3.as<int>(), cnt, out_dev);
  PFX_CUDA(cudaGetLastError());
  *n_out = cnt;
  for (int i = 0; i < 12; ++i) T16_host[i] = (float)Th[i];
  T16_host[12] = T16_host[13] = T16_host[14] = 0.f;
  T16_host[15] = 1.f;
  return 0;
}

}  // namespace pfx
