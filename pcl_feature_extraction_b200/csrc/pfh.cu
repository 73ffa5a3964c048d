// pfh.cu — PFH125 and PrincipalCurvatures (descriptors of the reference's evaluation matrix
// that share the path's neighbourhood machinery; replace pcl::PFHEstimation::compute as instantiated at reference
// evaluation.cpp:676-695 and pcl::PrincipalCurvaturesEstimation::compute at evaluation.cpp:696-715, both driven
// through features.h:181-195).
//
// pfh_kernel: one block per query.  The neighbourhood (positions, normals, distance keys) is staged once in shared
// memory - from the 3x3x3 stencil of the radius grid, or from the query's kNN row - and the block then walks all
// n (n - 1) / 2 unordered pairs: pair features (pair_features.cuh), 5 x 5 x 5 bin, integer votes in per-warp
// shared-memory histograms (votes of a warp to the same bin are merged with __match_any_sync first).  PCL adds the
// constant 100 / (n (n - 1) / 2) per vote in float; seq_float_sum(incr, votes) reproduces that sequential sum
// bit for bit (seqsum.h), so the row is PCL's row whenever the votes are.  The pair's argument order follows
// PCL's loop (i later than j in its distance-sorted list): the neighbour with the larger (d2, index) key is i.
// Queries whose neighbourhood exceeds the staged capacity are redone by a second launch with a 6144-point stage.
//
// curvature_kernel: one warp per query: normals of the neighbours projected onto the tangent plane of the query's
// normal, covariance of the projections in double, Jacobi eigen solve, eigenvector of the largest eigenvalue
// built like pcl::computeCorrespondingEigenVector (which fixes its sign), pc1 / pc2 = eigenvalues / n.
#include "internal.h"
#include "pair_features.cuh"
#include "seqsum.h"

namespace pfx {

constexpr int PFH_THREADS = 256;
constexpr int PFH_WARPS = PFH_THREADS / 32;
constexpr int PFH_CAP_SMALL = 1024;
constexpr int PFH_CAP_LARGE = 6144;

struct PfhSmemHead {
  int hist[PFH_WARPS][125];
  int count;
  int pad[3];
};

// mode: 0 = all queries, 1 = only the queries listed in `work` (second pass, large stage)
template <bool DENSE, bool USE_LIST>
__global__ void __launch_bounds__(PFH_THREADS)
pfh_kernel(GridDev g, const float4* __restrict__ nrm, const float4* __restrict__ queries, int nq, float r2,
           const int* __restrict__ lists, int k, int cap, const int* __restrict__ work, int* __restrict__ overflow,
           int* __restrict__ overflow_count, float* __restrict__ out, size_t stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PfhSmemHead* H = reinterpret_cast<PfhSmemHead*>(smem_raw);
  float4* sp = reinterpret_cast<float4*>(smem_raw + sizeof(PfhSmemHead));  // x, y, z, d2
  float4* sn = sp + cap;                                                     // nx, ny, nz, original index bits
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int qi = work ? work[blockIdx.x] : blockIdx.x;
  if (qi >= nq) return;
  const GridParams P = *g.gp;
  const float4 q = DENSE ? g.pts[qi] : queries[qi];
  const int row = DENSE ? __float_as_int(q.w) : qi;
  float* o = out + (size_t)row * stride;
  for (int b = tid; b < PFH_WARPS * 125; b += PFH_THREADS) (&H->hist[0][0])[b] = 0;
  if (tid == 0) H->count = 0;
  __syncthreads();
  const bool ok = finite3(q.x, q.y, q.z) && (!DENSE || qi < P.n_valid);
  if (ok) {
    if (USE_LIST) {
      for (int t = tid; t < k; t += PFH_THREADS) {
        const int j = lists[(size_t)qi * k + t];
        if (j >= 0) {
          const float4 p = g.pts[j];
          const float4 nj = nrm[j];
          const int pos = atomicAdd(&H->count, 1);
          if (pos < cap) {
            sp[pos] = make_float4(p.x, p.y, p.z, dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z));
            sn[pos] = make_float4(nj.x, nj.y, nj.z, p.w);
          }
        }
      }
    } else {
      const CellBlock blk = DENSE ? stencil_of_point(g, qi, lane) : stencil_of_pos(g, q.x, q.y, q.z, lane);
      for (int base = wid * 32; base < blk.total; base += PFH_THREADS) {
        const int t = base + lane;
        const bool valid = t < blk.total;
        const int j = block_candidate(blk, valid ? t : 0);
        if (valid) {
          const float4 p = g.pts[j];
          const float d2 = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);
          if (d2 < r2) {
            const float4 nj = nrm[j];
            const int pos = atomicAdd(&H->count, 1);
            if (pos < cap) {
              sp[pos] = make_float4(p.x, p.y, p.z, d2);
              sn[pos] = make_float4(nj.x, nj.y, nj.z, p.w);
            }
          }
        }
      }
    }
  }
  __syncthreads();
  const int n = H->count;
  if (n > cap) {  // redo with the large stage (or report, when this already is the large stage)
    if (tid == 0) {
      const int slot = atomicAdd(overflow_count, 1);
      if (overflow) overflow[slot] = qi;
    }
    return;
  }
  if (n == 0) {  // PCL: no neighbours -> NaN row
    for (int b = tid; b < 125; b += PFH_THREADS) o[b] = __int_as_float(0x7fc00000);
    return;
  }
  const double d_pi = (double)(1.0f / (2.0f * 3.14159265358979323846f));
  const long long npairs = (long long)n * (n - 1) / 2;
  for (long long pbase = 0; pbase < npairs; pbase += PFH_THREADS) {
    const long long p = pbase + tid;
    int bin = -1;
    if (p < npairs) {
      // p -> (a, b), 0 <= b < a < n, p = a (a - 1) / 2 + b
      int a = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)p)) * 0.5f);
      while ((long long)a * (a - 1) / 2 > p) --a;
      while ((long long)(a + 1) * a / 2 <= p) ++a;
      const int b = (int)(p - (long long)a * (a - 1) / 2);
      float4 pa = sp[a], na = sn[a], pb = sp[b], nb = sn[b];
      // PCL's (i, j): i comes later in its ascending (d2, index) list
      const bool a_later = (pa.w > pb.w) || (pa.w == pb.w && __float_as_int(na.w) > __float_as_int(nb.w));
      if (!a_later) {
        float4 t = pa; pa = pb; pb = t;
        t = na; na = nb; nb = t;
      }
      float f1, f2, f3;
      if (finite3(na.x, na.y, na.z) && finite3(nb.x, nb.y, nb.z) &&
          pair_features<5>(pa.x, pa.y, pa.z, na, pb.x, pb.y, pb.z, nb, f1, f2, f3)) {
        const int b1 = clamp_bin_n(5 * (((double)f1 + 3.14159265358979323846) * d_pi), 5);
        const int b2 = clamp_bin_n(5 * (((double)f2 + 1.0) * 0.5), 5);
        const int b3 = clamp_bin_n(5 * (((double)f3 + 1.0) * 0.5), 5);
        bin = b1 + 5 * b2 + 25 * b3;
      }
    }
    // merge the warp's votes per bin, one shared-memory add per distinct bin
    const unsigned active = __ballot_sync(FULL, bin >= 0);
    if (bin >= 0) {
      const unsigned peers = __match_any_sync(active, bin);
      if (lane == __ffs(peers) - 1) H->hist[wid][bin] += __popc(peers);
    }
    __syncwarp();
  }
  __syncthreads();
  const float incr = __fdiv_rn(100.0f, (float)npairs);  // n = 1: 100 / 0 = inf, no votes -> zeros (as upstream)
  for (int b = tid; b < 125; b += PFH_THREADS) {
    int c = 0;
#pragma unroll
    for (int w = 0; w < PFH_WARPS; ++w) c += H->hist[w][b];
    o[b] = (c > 0) ? seq_float_sum(incr, c) : 0.f;
  }
}

static size_t pfh_smem(int cap) { return sizeof(PfhSmemHead) + (size_t)cap * 2 * sizeof(float4); }

template <bool DENSE, bool USE_LIST>
static int pfh_launch(Ctx* ctx, Grid* g, const float4* nrm, int nq, float r2, const int* lists, int k, float* out_dev,
                      size_t stride_floats) {
  PFX_CUDA(ctx->tmp1.ensure(((size_t)nq + 16) * sizeof(int)));
  int* ov_count = ctx->tmp1.as<int>();
  int* ov_list = ov_count + 16;
  PFX_CUDA(cudaMemsetAsync(ov_count, 0, 16 * sizeof(int), ctx->stream));
  auto kern = pfh_kernel<DENSE, USE_LIST>;
  PFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pfh_smem(PFH_CAP_LARGE)));
  const float4* qry = DENSE ? nullptr : ctx->qry.as<float4>();
  PFX_LAUNCH(ctx, kern, nq, PFH_THREADS, pfh_smem(PFH_CAP_SMALL), g->view(), nrm, qry, nq, r2, lists, k, PFH_CAP_SMALL,
             (const int*)nullptr, ov_list, ov_count, out_dev, stride_floats);
  int nov = 0;
  PFX_CUDA(cudaMemcpyAsync(&nov, ov_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  if (nov > 0) {
    PFX_CUDA(cudaMemsetAsync(ov_count, 0, sizeof(int), ctx->stream));
    PFX_LAUNCH(ctx, kern, nov, PFH_THREADS, pfh_smem(PFH_CAP_LARGE), g->view(), nrm, qry, nq, r2, lists, k, PFH_CAP_LARGE,
               (const int*)ov_list, (int*)nullptr, ov_count, out_dev, stride_floats);
    int still = 0;
    PFX_CUDA(cudaMemcpyAsync(&still, ov_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PFX_CUDA(cudaStreamSynchronize(ctx->stream));
    if (still > 0) return ctx->fail(PFX_E_CAPACITY, "pfx_pfh125: a neighbourhood exceeds 6144 points (18.9 M pairs)");
  }
  PFX_CUDA(cudaGetLastError());
  return 0;
}

// out_dev: rows of 125 floats at stride_floats, caller query order
int pfh_compute(Ctx* ctx, Grid* g, double radius, int k, float* out_dev, size_t stride_floats) {
  const int nq = (int)ctx->num_queries();
  if (nq == 0 || ctx->n == 0) return 0;
  const float4* nrm = nullptr;
  PFX_TRY(normals_sorted_for_grid(ctx, g, &nrm));
  const float r2 = (float)(radius * radius);
  if (k > 0) {
    PFX_TRY(knn_lists(ctx, g, k, false));
    if (ctx->q_is_surface) return pfh_launch<true, true>(ctx, g, nrm, nq, r2, ctx->knn_idx.as<int>(), k, out_dev, stride_floats);
    return pfh_launch<false, true>(ctx, g, nrm, nq, r2, ctx->knn_idx.as<int>(), k, out_dev, stride_floats);
  }
  if (ctx->q_is_surface) return pfh_launch<true, false>(ctx, g, nrm, nq, r2, nullptr, 0, out_dev, stride_floats);
  return pfh_launch<false, false>(ctx, g, nrm, nq, r2, nullptr, 0, out_dev, stride_floats);
}

// ------------------------------------------------------------------------------------ PrincipalCurvatures
constexpr int PC_WPB = 8;

template <bool DENSE, bool USE_LIST>
__global__ void __launch_bounds__(PC_WPB * 32)
curvature_kernel(GridDev g, const float4* __restrict__ nrm, const float4* __restrict__ nrm_orig, int n_surf,
                 const float4* __restrict__ queries, int nq, float r2, const int* __restrict__ lists, int k,
                 float* __restrict__ out, size_t stride) {
  const int lane = threadIdx.x & 31;
  const int qi = blockIdx.x * PC_WPB + (threadIdx.x >> 5);
  if (qi >= nq) return;
  const GridParams P = *g.gp;
  const float4 q = DENSE ? g.pts[qi] : queries[qi];
  const int row = DENSE ? __float_as_int(q.w) : qi;
  float* o = out + (size_t)row * stride;
  const float NaN = __int_as_float(0x7fc00000);
  const bool ok = finite3(q.x, q.y, q.z) && (!DENSE || qi < P.n_valid) && row < n_surf;
  // upstream indexes the normals with the query's ordinal: normals.points[(*indices_)[idx]]
  const float4 nq4 = ok ? nrm_orig[row] : make_float4(0.f, 0.f, 1.f, 0.f);
  const double nx = nq4.x, ny = nq4.y, nz = nq4.z;
  // moments of the projected normals M n_j, M = I - n n^T
  double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  int cnt = 0;
  auto add = [&](const float4& nj) {
    const double ax = nj.x, ay = nj.y, az = nj.z;
    const double px = (1.0 - nx * nx) * ax + (-nx * ny) * ay + (-nx * nz) * az;
    const double py = (-ny * nx) * ax + (1.0 - ny * ny) * ay + (-ny * nz) * az;
    const double pz = (-nz * nx) * ax + (-nz * ny) * ay + (1.0 - nz * nz) * az;
    s[0] += px; s[1] += py; s[2] += pz;
    s[3] += px * px; s[4] += px * py; s[5] += px * pz;
    s[6] += py * py; s[7] += py * pz; s[8] += pz * pz;
    ++cnt;
  };
  if (ok) {
    if (USE_LIST) {
      for (int t = lane; t < k; t += 32) {
        const int j = lists[(size_t)qi * k + t];
        if (j >= 0) add(nrm[j]);
      }
    } else {
      const CellBlock blk = DENSE ? stencil_of_point(g, qi, lane) : stencil_of_pos(g, q.x, q.y, q.z, lane);
      for (int base = 0; base < blk.total; base += 32) {
        const int t = base + lane;
        const bool valid = t < blk.total;
        const int j = block_candidate(blk, valid ? t : 0);
        if (valid) {
          const float4 p = g.pts[j];
          if (dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z) < r2) add(nrm[j]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) s[i] = warp_sum(s[i]);
  cnt = warp_sum(cnt);
  if (lane != 0) return;
  bool fin = ok && cnt > 0;
  for (int i = 0; i < 9; ++i) fin = fin && isfinite(s[i]);
  if (!fin) {
    for (int i = 0; i < 5; ++i) o[i] = NaN;
    return;
  }
  const double m = (double)cnt, cx = s[0] / m, cy = s[1] / m, cz = s[2] / m;
  // sum (p - c)(p - c)^T = sum p p^T - m c c^T
  const double C[6] = {s[3] - m * cx * cx, s[4] - m * cx * cy, s[5] - m * cx * cz,
                       s[6] - m * cy * cy, s[7] - m * cy * cz, s[8] - m * cz * cz};
  double w[3], v[3][3];
  eig_sym3<double>(C, w, v, 60);
  // pcl::computeCorrespondingEigenVector(C, w[2]): rows of the scaled C - l I, largest cross product
  double scale = fmax(fmax(fabs(C[0]), fabs(C[1])), fmax(fmax(fabs(C[2]), fabs(C[3])), fmax(fabs(C[4]), fabs(C[5]))));
  if (scale <= 2.2250738585072014e-308) scale = 1.0;
  const double l = w[2] / scale;
  const double r0[3] = {C[0] / scale - l, C[1] / scale, C[2] / scale};
  const double r1[3] = {C[1] / scale, C[3] / scale - l, C[4] / scale};
  const double r2v[3] = {C[2] / scale, C[4] / scale, C[5] / scale - l};
  const double v1[3] = {r0[1] * r1[2] - r0[2] * r1[1], r0[2] * r1[0] - r0[0] * r1[2], r0[0] * r1[1] - r0[1] * r1[0]};
  const double v2[3] = {r0[1] * r2v[2] - r0[2] * r2v[1], r0[2] * r2v[0] - r0[0] * r2v[2], r0[0] * r2v[1] - r0[1] * r2v[0]};
  const double v3[3] = {r1[1] * r2v[2] - r1[2] * r2v[1], r1[2] * r2v[0] - r1[0] * r2v[2], r1[0] * r2v[1] - r1[1] * r2v[0]};
  const double l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2], l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2],
               l3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
  const double* best = v3;
  double bl = l3;
  if (l1 >= l2 && l1 >= l3) { best = v1; bl = l1; }
  else if (l2 >= l1 && l2 >= l3) { best = v2; bl = l2; }
  const double inv = 1.0 / sqrt(bl);
  o[0] = (float)(best[0] * inv);
  o[1] = (float)(best[1] * inv);
  o[2] = (float)(best[2] * inv);
  o[3] = (float)(w[2] / m);
  o[4] = (float)(w[1] / m);
}

// out_dev: rows of 5 floats (principal direction, pc1, pc2) at stride_floats, caller query order
int curvature_compute(Ctx* ctx, Grid* g, double radius, int k, float* out_dev, size_t stride_floats) {
  const int nq = (int)ctx->num_queries();
  if (nq == 0 || ctx->n == 0) return 0;
  const float4* nrm = nullptr;
  PFX_TRY(normals_sorted_for_grid(ctx, g, &nrm));
  const float r2 = (float)(radius * radius);
  const float4* nrm_orig = ctx->normals.as<float4>();
  const int blocks = div_up(nq, PC_WPB);
  if (k > 0) {
    PFX_TRY(knn_lists(ctx, g, k, false));
    if (ctx->q_is_surface)
      PFX_LAUNCH(ctx, (curvature_kernel<true, true>), blocks, PC_WPB * 32, 0, g->view(), nrm, nrm_orig, (int)ctx->n, nullptr, nq,
                 r2, ctx->knn_idx.as<int>(), k, out_dev, stride_floats);
    else
      PFX_LAUNCH(ctx, (curvature_kernel<false, true>), blocks, PC_WPB * 32, 0, g->view(), nrm, nrm_orig, (int)ctx->n,
                 ctx->qry.as<float4>(), nq, r2, ctx->knn_idx.as<int>(), k, out_dev, stride_floats);
  } else {
    if (ctx->q_is_surface)
      PFX_LAUNCH(ctx, (curvature_kernel<true, false>), blocks, PC_WPB * 32, 0, g->view(), nrm, nrm_orig, (int)ctx->n, nullptr,
                 nq, r2, nullptr, 0, out_dev, stride_floats);
    else
      PFX_LAUNCH(ctx, (curvature_kernel<false, false>), blocks, PC_WPB * 32, 0, g->view(), nrm, nrm_orig, (int)ctx->n,
                 ctx->qry.as<float4>(), nq, r2, nullptr, 0, out_dev, stride_floats);
  }
  PFX_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------ MomentInvariants
// pcl::MomentInvariantsEstimation (reference evaluation.cpp:555-574): central second moments of the neighbourhood
// about its centroid -> j1 (trace), j2 (sum of the principal 2x2 minors), j3 (determinant).  One warp per query,
// moments about the query point in double, shifted to the centroid in closed form.
template <bool DENSE, bool USE_LIST>
__global__ void __launch_bounds__(PC_WPB * 32)
moments_kernel(GridDev g, const float4* __restrict__ queries, int nq, float r2, const int* __restrict__ lists, int k,
               float* __restrict__ out, size_t stride) {
  const int lane = threadIdx.x & 31;
  const int qi = blockIdx.x * PC_WPB + (threadIdx.x >> 5);
  if (qi >= nq) return;
  const GridParams P = *g.gp;
  const float4 q = DENSE ? g.pts[qi] : queries[qi];
  const int row = DENSE ? __float_as_int(q.w) : qi;
  float* o = out + (size_t)row * stride;
  const bool ok = finite3(q.x, q.y, q.z) && (!DENSE || qi < P.n_valid);
  double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  int cnt = 0;
  auto add = [&](const float4& p) {
    const double x = (double)p.x - (double)q.x, y = (double)p.y - (double)q.y, z = (double)p.z - (double)q.z;
    s[0] += x; s[1] += y; s[2] += z;
    s[3] += x * x; s[4] += y * y; s[5] += z * z;
    s[6] += x * y; s[7] += x * z; s[8] += y * z;
    ++cnt;
  };
  if (ok) {
    if (USE_LIST) {
      for (int t = lane; t < k; t += 32) {
        const int j = lists[(size_t)qi * k + t];
        if (j >= 0) add(g.pts[j]);
      }
    } else {
      const CellBlock blk = DENSE ? stencil_of_point(g, qi, lane) : stencil_of_pos(g, q.x, q.y, q.z, lane);
      for (int base = 0; base < blk.total; base += 32) {
        const int t = base + lane;
        const bool valid = t < blk.total;
        const int j = block_candidate(blk, valid ? t : 0);
        if (valid) {
          const float4 p = g.pts[j];
          if (dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z) < r2) add(p);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) s[i] = warp_sum(s[i]);
  cnt = warp_sum(cnt);
  if (lane != 0) return;
  if (!ok || cnt == 0) {
    o[0] = o[1] = o[2] = __int_as_float(0x7fc00000);
    return;
  }
  const double m = (double)cnt, cx = s[0] / m, cy = s[1] / m, cz = s[2] / m;
  const double m200 = s[3] - m * cx * cx, m020 = s[4] - m * cy * cy, m002 = s[5] - m * cz * cz;
  const double m110 = s[6] - m * cx * cy, m101 = s[7] - m * cx * cz, m011 = s[8] - m * cy * cz;
  o[0] = (float)(m200 + m020 + m002);
  o[1] = (float)(m200 * m020 + m200 * m002 + m020 * m002 - m110 * m110 - m101 * m101 - m011 * m011);
  o[2] = (float)(m200 * m020 * m002 + 2 * m110 * m101 * m011 - m002 * m110 * m110 - m020 * m101 * m101 - m200 * m011 * m011);
}

// out_dev: rows of 3 floats (j1, j2, j3) at stride_floats, caller query order
int moments_compute(Ctx* ctx, Grid* g, double radius, int k, float* out_dev, size_t stride_floats) {
  const int nq = (int)ctx->num_queries();
  if (nq == 0 || ctx->n == 0) return 0;
  const float r2 = (float)(radius * radius);
  const int blocks = div_up(nq, PC_WPB);
  if (k > 0) {
    PFX_TRY(knn_lists(ctx, g, k, false));
    if (ctx->q_is_surface)
      PFX_LAUNCH(ctx, (moments_kernel<true, true>), blocks, PC_WPB * 32, 0, g->view(), nullptr, nq, r2, ctx->knn_idx.as<int>(), k,
                 out_dev, stride_floats);
    else
      PFX_LAUNCH(ctx, (moments_kernel<false, true>), blocks, PC_WPB * 32, 0, g->view(), ctx->qry.as<float4>(), nq, r2,
                 ctx->knn_idx.as<int>(), k, out_dev, stride_floats);
  } else {
    if (ctx->q_is_surface)
      PFX_LAUNCH(ctx, (moments_kernel<true, false>), blocks, PC_WPB * 32, 0, g->view(), nullptr, nq, r2, nullptr, 0, out_dev,
                 stride_floats);
    else
      PFX_LAUNCH(ctx, (moments_kernel<false, false>), blocks, PC_WPB * 32, 0, g->view(), ctx->qry.as<float4>(), nq, r2, nullptr, 0,
                 out_dev, stride_floats);
  }
  PFX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pfx
