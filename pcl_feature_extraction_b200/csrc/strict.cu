// strict.cu — "reference order" variants of the stages whose float results feed INDEX outputs
// (pfx_set_parity_mode(ctx, PFX_PARITY_STRICT)).
//
// The fast kernels accumulate in whatever order the voxel hash hands them neighbours and solve the 3x3 eigen
// problems with their own solver: their normals / responses / FPFH rows agree with the CPU path to the stated
// tolerances, but the reference pipeline then takes DISCRETE decisions on those floats - Harris3D thresholds a
// response of ~1e-6 and keeps local maxima (keypoints.h:154-162), reciprocal matching keeps an argmin
// (features.h:240-250) - and on a degenerate neighbourhood (two points, collinear points) the normal is whatever
// the solver's rounding makes of a rank-deficient matrix.  End-to-end index parity therefore needs the same
// arithmetic in the same order.  The kernels here walk the neighbours in PCL's list order (ascending (d2, index),
// what a sorted kd-tree search returns: SURVEY.md A.1) with one thread or one lane doing the sums sequentially,
// every operation rounded separately (this file is compiled with -fmad=false), and solve the eigen problem with
// the same cyclic Jacobi sweep order as the CPU restatement; IEEE add / mul / div / sqrt are correctly rounded on
// both machines, so the results are bit-identical.  Cost: one sorted list per query; meant for the reference's
// own workloads (keypoint pipelines on clouds of 1e4 .. 1e5 points), not for the dense 1M-point path.
//
//   strict_normals_kernel      NormalEstimation (tools.h:26-31): centred covariance in double, Jacobi, flip
//   harris_response_strict     HarrisKeypoint3D::responseHarris (keypoints.h:154-159): float sums in list order
//   harris_refine_strict       HarrisKeypoint3D::refineCorners: float sums in list order, <= 10 iterations
//   fpfh_sorted_kernel         FPFHEstimation::weightPointSPFHSignature (evaluation.cpp:597-602): float sums in
//                              list order, per-block double sums
#include "internal.h"

namespace pfx {

// ------------------------------------------------------------------------------------------- lists
// Sorted neighbour lists of the current queries in CALLER order, neighbours as ORIGINAL indices:
// radius search: ctx->st_off[nq + 1] offsets into st_idx / st_d2; k-search: rows of k entries (-1 = none).
int strict_lists_build(Ctx* ctx, Grid* g, double radius, int k) {
  const size_t nq = ctx->num_queries();
  ctx->st_k = k;
  ctx->st_total = 0;
  if (nq == 0) return 0;
  if (k > 0) {
    PFX_TRY(knn_lists(ctx, g, k, true));
    PFX_CUDA(ctx->st_idx.ensure(nq * k * sizeof(int)));
    PFX_CUDA(ctx->st_d2.ensure(nq * k * sizeof(float)));
    PFX_TRY(knn_export(ctx, k, ctx->st_idx.as<int>(), ctx->st_d2.as<float>(), PFX_DEVICE));
    ctx->st_total = (long long)nq * k;
    return 0;
  }
  PFX_CUDA(ctx->st_cnt.ensure(nq * sizeof(int)));
  PFX_CUDA(ctx->st_off.ensure((nq + 1) * sizeof(long long)));
  PFX_TRY(radius_count(ctx, g, radius, ctx->st_cnt.as<int>()));
  PFX_TRY(scan_exclusive_i64(ctx, ctx->st_cnt.as<int>(), ctx->st_off.as<long long>(), (int)nq, ctx->scanbuf));
  long long total = 0;
  PFX_CUDA(cudaMemcpyAsync(&total, ctx->st_off.as<long long>() + nq, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->st_total = total;
  PFX_CUDA(ctx->st_idx.ensure(std::max<size_t>((size_t)total, 1) * sizeof(int)));
  PFX_CUDA(ctx->st_d2.ensure(std::max<size_t>((size_t)total, 1) * sizeof(float)));
  if (total > 0) PFX_TRY(radius_fill(ctx, g, radius, 1, ctx->st_off.as<long long>(), ctx->st_idx.as<int>(), ctx->st_d2.as<float>()));
  return 0;
}

struct ListView {
  const long long* off;  // radius search (null in k-search)
  const int* idx;
  const float* d2;
  int k;
  __device__ __forceinline__ void row(int i, long long& b, int& n) const {
    if (off) {
      b = off[i];
      n = (int)(off[i + 1] - b);
    } else {
      b = (long long)i * k;
      n = 0;
      while (n < k && idx[b + n] >= 0) ++n;
    }
  }
};

static ListView list_view(const Ctx* ctx) {
  ListView v;
  v.off = ctx->st_k > 0 ? nullptr : ctx->st_off.as<long long>();
  v.idx = ctx->st_idx.as<int>();
  v.d2 = ctx->st_d2.as<float>();
  v.k = ctx->st_k;
  return v;
}

// ------------------------------------------------------------------------------------------- Jacobi
// One rotation of the cyclic sweep on the FULL 3x3 matrix (A <- A J, A <- J^T A, V <- V J), written for
// compile-time (P, Q) so that every element stays in a register.
template <int P, int Q>
__device__ __forceinline__ void strict_rot(double (&A)[3][3], double (&V)[3][3]) {
  if (A[P][Q] == 0.0) return;
  const double theta = (A[Q][Q] - A[P][P]) / (2.0 * A[P][Q]);
  const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
  const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double arp = A[r][P], arq = A[r][Q];
    A[r][P] = c * arp - s * arq;
    A[r][Q] = s * arp + c * arq;
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double apr = A[P][r], aqr = A[Q][r];
    A[P][r] = c * apr - s * aqr;
    A[Q][r] = s * apr + c * aqr;
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double vrp = V[r][P], vrq = V[r][Q];
    V[r][P] = c * vrp - s * vrq;
    V[r][Q] = s * vrp + c * vrq;
  }
}

// eigenvalues ascending in w, eigenvectors in the columns of V: cyclic Jacobi, sweeps (0,1) (0,2) (1,2), at most
// 60, stop when the off-diagonal mass is <= 1e-18 of the diagonal's; columns ordered by a stable insertion sort
__device__ __forceinline__ void strict_eig3(const double (&C)[3][3], double (&w)[3], double (&V)[3][3]) {
  double A[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      A[i][j] = C[i][j];
      V[i][j] = (i == j) ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 60; ++sweep) {
    const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    const double dg = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
    if (off <= 1e-300 || off <= 1e-18 * dg) break;
    strict_rot<0, 1>(A, V);
    strict_rot<0, 2>(A, V);
    strict_rot<1, 2>(A, V);
  }
  double d0 = A[0][0], d1 = A[1][1], d2 = A[2][2];
  double c0[3] = {V[0][0], V[1][0], V[2][0]}, c1[3] = {V[0][1], V[1][1], V[2][1]}, c2[3] = {V[0][2], V[1][2], V[2][2]};
#define PFX_SWAP_COL(da, ca, db, cb) \
  {                                  \
    double td = da;                  \
    da = db;                         \
    db = td;                         \
    _Pragma("unroll") for (int r = 0; r < 3; ++r) { \
      double tv = ca[r];             \
      ca[r] = cb[r];                 \
      cb[r] = tv;                    \
    }                                \
  }
  // stable insertion sort of three (an element moves only past strictly larger ones)
  if (d1 < d0) PFX_SWAP_COL(d0, c0, d1, c1)
  if (d2 < d1) {
    PFX_SWAP_COL(d1, c1, d2, c2)
    if (d1 < d0) PFX_SWAP_COL(d0, c0, d1, c1)
  }
#undef PFX_SWAP_COL
  w[0] = d0;
  w[1] = d1;
  w[2] = d2;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    V[r][0] = c0[r];
    V[r][1] = c1[r];
    V[r][2] = c2[r];
  }
}

// ------------------------------------------------------------------------------------------- normals
// one thread per query; surf = float4 points in ORIGINAL order; rows in caller order
__global__ void __launch_bounds__(128)
strict_normals_kernel(ListView L, const float4* __restrict__ surf, const float4* __restrict__ queries, int nq, float vx,
                      float vy, float vz, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  long long b;
  int m;
  L.row(i, b, m);
  const float nanv = __int_as_float(0x7fc00000);
  const float4 q = queries[i];
  if (m == 0 || !finite3(q.x, q.y, q.z)) {  // non-finite query or no neighbours -> NaN row
    out[i] = make_float4(nanv, nanv, nanv, nanv);
    return;
  }
  double mu[3] = {0, 0, 0};
  for (int t = 0; t < m; ++t) {
    const float4 p = surf[L.idx[b + t]];
    mu[0] += (double)p.x;
    mu[1] += (double)p.y;
    mu[2] += (double)p.z;
  }
  mu[0] /= m;
  mu[1] /= m;
  mu[2] /= m;
  double c00 = 0, c01 = 0, c02 = 0, c11 = 0, c12 = 0, c22 = 0;
  for (int t = 0; t < m; ++t) {
    const float4 p = surf[L.idx[b + t]];
    const double d0 = (double)p.x - mu[0], d1 = (double)p.y - mu[1], d2 = (double)p.z - mu[2];
    c00 += d0 * d0; c01 += d0 * d1; c02 += d0 * d2;
    c11 += d1 * d1; c12 += d1 * d2; c22 += d2 * d2;
  }
  c00 /= m; c01 /= m; c02 /= m; c11 /= m; c12 /= m; c22 /= m;
  const double C[3][3] = {{c00, c01, c02}, {c01, c11, c12}, {c02, c12, c22}};
  double w[3], V[3][3];
  strict_eig3(C, w, V);
  double n0 = V[0][0], n1 = V[1][0], n2 = V[2][0];
  const double tr = c00 + c11 + c22;
  const double curv = (tr != 0.0) ? fabs(w[0] / tr) : 0.0;
  // flipNormalTowardsViewpoint: (vp - p) . n < 0  =>  n = -n
  double dp = 0;
  dp += ((double)vx - (double)q.x) * n0;
  dp += ((double)vy - (double)q.y) * n1;
  dp += ((double)vz - (double)q.z) * n2;
  if (dp < 0) {
    n0 = -n0;
    n1 = -n1;
    n2 = -n2;
  }
  out[i] = make_float4((float)n0, (float)n1, (float)n2, (float)curv);
}

// dense: out rows = surface normals in original order (ctx->normals); keypoint queries: out_query_order
int strict_normals(Ctx* ctx, Grid* g, double radius, int k, float4* out_query_order) {
  const int nq = (int)ctx->num_queries();
  const bool dense = ctx->q_is_surface;
  if (dense) PFX_CUDA(ctx->normals.ensure(std::max<size_t>(ctx->n, 1) * sizeof(float4)));
  if (nq == 0) {
    if (dense) {
      ctx->have_normals = true;
      ctx->normals_version++;
    }
    return 0;
  }
  if (!dense && !out_query_order) return 0;
  PFX_TRY(strict_lists_build(ctx, g, radius, k));
  float4* out = dense ? ctx->normals.as<float4>() : out_query_order;
  const float4* qry = dense ? ctx->surf.as<float4>() : ctx->qry.as<float4>();
  PFX_LAUNCH(ctx, strict_normals_kernel, div_up(nq, 128), 128, 0, list_view(ctx), ctx->surf.as<float4>(), qry, nq,
             ctx->vp[0], ctx->vp[1], ctx->vp[2], out);
  PFX_CUDA(cudaGetLastError());
  if (dense) {
    ctx->have_normals = true;
    ctx->normals_version++;
    ctx->normals_sorted_for = nullptr;  // the sorted copy of any grid is rebuilt on demand
    if (out_query_order)
      PFX_CUDA(cudaMemcpyAsync(out_query_order, out, (size_t)nq * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------- Harris3D
// responseHarris: mean of n n^T over the neighbours with a finite normal, float sums in list order;
// response = 0.04 + det - 0.04 tr^2.  One thread per surface point; lists = the surface's own (dense) lists.
__global__ void __launch_bounds__(128)
harris_response_strict_kernel(ListView L, const float4* __restrict__ surf, const float4* __restrict__ nrm, int n,
                              float* __restrict__ resp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float r = 0.f;
  const float4 q = surf[i];
  if (finite3(q.x, q.y, q.z)) {
    long long b;
    int m;
    L.row(i, b, m);
    float xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0;
    unsigned count = 0;
    for (int t = 0; t < m; ++t) {
      const float4 nr = nrm[L.idx[b + t]];
      if (!isfinite(nr.x)) continue;
      xx += nr.x * nr.x; xy += nr.x * nr.y; xz += nr.x * nr.z;
      yy += nr.y * nr.y; yz += nr.y * nr.z; zz += nr.z * nr.z;
      ++count;
    }
    if (count > 0) {
      const float fc = (float)count;
      xx /= fc; xy /= fc; xz /= fc; yy /= fc; yz /= fc; zz /= fc;
    }
    const float trace = xx + yy + zz;
    if (trace != 0) {
      const float det = xx * yy * zz + 2.0f * xy * xz * yz - xz * xz * yy - xy * xy * zz - yz * yz * xx;
      r = 0.04f + det - 0.04f * trace * trace;
    }
  }
  resp[i] = r;
}

int harris_response_strict(Ctx* ctx, Grid* g, double radius, float* resp_dev_orig) {
  const int n = (int)ctx->n;
  if (n == 0) return 0;
  const bool saved = ctx->q_is_surface;
  ctx->q_is_surface = true;
  int rc = strict_lists_build(ctx, g, radius, 0);
  ctx->q_is_surface = saved;
  if (rc) return rc;
  PFX_LAUNCH(ctx, harris_response_strict_kernel, div_up(n, 128), 128, 0, list_view(ctx), ctx->surf.as<float4>(),
             ctx->normals.as<float4>(), n, resp_dev_orig);
  PFX_CUDA(cudaGetLastError());
  return 0;
}

// refineCorners: one warp per corner.  Every iteration the warp collects the neighbours of the current corner
// position, ranks them by (d2, index) in shared memory, and lane 0 adds n n^T and (n n^T) p in that order.
constexpr int RCAP = 512;  // neighbours of a corner the warp can rank (r = 1 cm: ~20)
constexpr int RWPB = 4;

__global__ void __launch_bounds__(RWPB * 32)
harris_refine_strict_kernel(GridDev g, const float4* __restrict__ nrm_orig, float r2, float* __restrict__ corners, int nc,
                            int* __restrict__ overflow) {
  __shared__ unsigned long long skeys[RWPB][RCAP];
  __shared__ int sorder[RWPB][RCAP];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int ci = blockIdx.x * RWPB + wid;
  if (ci >= nc) return;
  unsigned long long* keys = skeys[wid];
  int* order = sorder[wid];
  const unsigned lt = (1u << lane) - 1u;
  float cr0 = corners[3 * ci], cr1 = corners[3 * ci + 1], cr2 = corners[3 * ci + 2];
  unsigned iterations = 0;
  for (;;) {
    const float c0 = cr0, c1 = cr1, c2 = cr2;
    int n = 0;
    if (finite3(c0, c1, c2)) {
      CellBlock blk = stencil_of_pos(g, c0, c1, c2, lane);
      for (int base = 0; base < blk.total; base += 32) {
        const int c = base + lane;
        bool in = c < blk.total;
        const int j = block_candidate(blk, in ? c : 0);
        unsigned long long key = 0;
        if (in) {
          const float4 p = g.pts[j];
          const float d2 = dist2_flann(c0, c1, c2, p.x, p.y, p.z);
          in = d2 < r2;
          key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)__float_as_int(p.w);
        }
        const unsigned m = __ballot_sync(FULL, in);
        const int pos = n + __popc(m & lt);
        if (in && pos < RCAP) keys[pos] = key;
        n += __popc(m);
      }
    }
    __syncwarp();
    if (n > RCAP) {  // cannot rank that many: reported, the corner stays where it is
      if (lane == 0) atomicAdd(overflow, 1);
      break;
    }
    for (int a = lane; a < n; a += 32) {
      const unsigned long long ka = keys[a];
      int rank = 0;
      for (int b = 0; b < n; ++b) rank += (keys[b] < ka) ? 1 : 0;
      order[rank] = (int)(unsigned)(ka & 0xffffffffull);  // keys are distinct (distinct indices)
    }
    __syncwarp();
    float diff = 0.f;
    if (lane == 0) {
      float NNT[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, NNTp[3] = {0, 0, 0};
      for (int t = 0; t < n; ++t) {
        const int o = order[t];
        const float4 nr4 = nrm_orig[o];
        if (!isfinite(nr4.x)) continue;
        const float4 p4 = g.pts[g.inv_perm[o]];
        const float nr[3] = {nr4.x, nr4.y, nr4.z};
        const float p[3] = {p4.x, p4.y, p4.z};
        float nnT[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) nnT[3 * r + c] = nr[r] * nr[c];
#pragma unroll
        for (int e = 0; e < 9; ++e) NNT[e] += nnT[e];
#pragma unroll
        for (int r = 0; r < 3; ++r) NNTp[r] += nnT[3 * r] * p[0] + nnT[3 * r + 1] * p[1] + nnT[3 * r + 2] * p[2];
      }
      // invert3x3SymMatrix
      const float a = NNT[0], bq = NNT[1], c = NNT[2], d = NNT[4], e = NNT[5], f = NNT[8];
      const float fd_ee = d * f - e * e;
      const float ce_bf = c * e - bq * f;
      const float be_cd = bq * e - c * d;
      const float det = a * fd_ee + bq * ce_bf + c * be_cd;
      if (det != 0) {
        float inv[9] = {fd_ee, ce_bf, be_cd, ce_bf, a * f - c * c, bq * c - a * e, be_cd, bq * c - a * e, a * d - bq * bq};
#pragma unroll
        for (int t = 0; t < 9; ++t) inv[t] /= det;
        cr0 = inv[0] * NNTp[0] + inv[1] * NNTp[1] + inv[2] * NNTp[2];
        cr1 = inv[3] * NNTp[0] + inv[4] * NNTp[1] + inv[5] * NNTp[2];
        cr2 = inv[6] * NNTp[0] + inv[7] * NNTp[1] + inv[8] * NNTp[2];
      }
      const float dx = cr0 - c0, dy = cr1 - c1, dz = cr2 - c2;
      diff = dx * dx + dy * dy + dz * dz;
    }
    cr0 = __shfl_sync(FULL, cr0, 0);
    cr1 = __shfl_sync(FULL, cr1, 0);
    cr2 = __shfl_sync(FULL, cr2, 0);
    diff = __shfl_sync(FULL, diff, 0);
    __syncwarp();
    // do { ... } while (diff > 1e-6 && ++iterations < 10)   (float against the double constant, as upstream)
    if (!((double)diff > 1e-6 && ++iterations < 10)) break;
  }
  if (lane == 0) {
    corners[3 * ci] = cr0;
    corners[3 * ci + 1] = cr1;
    corners[3 * ci + 2] = cr2;
  }
}

int harris_refine_strict(Ctx* ctx, Grid* g, double radius, float* corners_dev, int nc) {
  if (nc == 0) return 0;
  PFX_CUDA(ctx->small.ensure(256));
  int* ov = ctx->small.as<int>() + 40;
  PFX_CUDA(cudaMemsetAsync(ov, 0, sizeof(int), ctx->stream));
  PFX_LAUNCH(ctx, harris_refine_strict_kernel, div_up(nc, RWPB), RWPB * 32, 0, g->view(), ctx->normals.as<float4>(),
             (float)(radius * radius), corners_dev, nc, ov);
  PFX_CUDA(cudaGetLastError());
  int h = 0;
  PFX_CUDA(cudaMemcpyAsync(&h, ov, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h) return ctx->fail(PFX_E_CAPACITY, "strict Harris refinement: a corner has more than 512 neighbours");
  return 0;
}

// ------------------------------------------------------------------------------------------- FPFH
// weightPointSPFHSignature in list order: one warp per query, lanes are bins (lane 0 also bin 32); spfh = 33-float
// rows in the sorted order of grid g.  F += h * w in float; the three per-block double sums take the added values
// neighbour by neighbour, bin by bin, exactly as the CPU loop does.
__global__ void __launch_bounds__(128)
fpfh_sorted_kernel(ListView L, GridDev g, const float* __restrict__ spfh, const float4* __restrict__ queries, int nq,
                   float* __restrict__ out, size_t stride) {
  const int lane = threadIdx.x & 31;
  const int qi = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (qi >= nq) return;
  float* o = out + (size_t)qi * stride;
  long long b;
  int m;
  L.row(qi, b, m);
  const float4 q = queries[qi];
  const float nanv = __int_as_float(0x7fc00000);
  if (m == 0 || !finite3(q.x, q.y, q.z)) {
    o[lane] = nanv;
    if (lane == 0) o[32] = nanv;
    return;
  }
  float F0 = 0.f, F1 = 0.f;
  double sum0 = 0.0, sum1 = 0.0, sum2 = 0.0;  // every lane keeps all three block sums (identical values)
  for (int t = 0; t < m; ++t) {
    const float d2 = L.d2[b + t];
    if (d2 == 0.f) continue;  // "minus the query point itself"
    const float w = 1.0f / d2;
    const float* rr = spfh + (size_t)g.inv_perm[L.idx[b + t]] * 33;
    const float val = rr[lane] * w;
    const float val1 = (lane == 0) ? rr[32] * w : 0.f;
    F0 += val;
    F1 += val1;
    // the CPU adds the 33 values of a neighbour to its three double sums one after the other
#pragma unroll
    for (int c = 0; c < 11; ++c) sum0 += (double)__shfl_sync(FULL, val, c);
#pragma unroll
    for (int c = 11; c < 22; ++c) sum1 += (double)__shfl_sync(FULL, val, c);
#pragma unroll
    for (int c = 22; c < 32; ++c) sum2 += (double)__shfl_sync(FULL, val, c);
    sum2 += (double)__shfl_sync(FULL, val1, 0);
  }
  if (sum0 != 0.0) sum0 = 100.0 / sum0;
  if (sum1 != 0.0) sum1 = 100.0 / sum1;
  if (sum2 != 0.0) sum2 = 100.0 / sum2;
  const float sc = (float)((lane < 11) ? sum0 : (lane < 22 ? sum1 : sum2));
  o[lane] = F0 * sc;
  if (lane == 0) o[32] = F1 * (float)sum2;
}

// radius search only; rows in caller query order at `stride_floats`
int fpfh_sorted(Ctx* ctx, Grid* g, double radius, const float* spfh_sorted_rows, float* out_dev, size_t stride_floats) {
  const int nq = (int)ctx->num_queries();
  if (nq == 0) return 0;
  PFX_TRY(strict_lists_build(ctx, g, radius, 0));
  const float4* qry = ctx->q_is_surface ? ctx->surf.as<float4>() : ctx->qry.as<float4>();
  PFX_LAUNCH(ctx, fpfh_sorted_kernel, div_up(nq, 4), 128, 0, list_view(ctx), g->view(), spfh_sorted_rows, qry, nq,
             out_dev, stride_floats);
  PFX_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------- Harris 6D
// HarrisKeypoint6D (reference keypoints.h:166-179): intensity from the colours, IntensityGradientEstimation, the
// 6x6 covariance of (normal, normalised gradient), response = its 4th smallest eigenvalue.  Everything here walks the
// sorted lists like the kernels above; the definitions that replace unpinnable Eigen internals (the 3x3
// column-pivoting Householder QR written out, the 6x6 cyclic Jacobi in double) are stated in DESIGN.md section 3 and
// followed operation for operation on both sides.
__global__ void harris6d_intensity_kernel(const unsigned* __restrict__ rgb, int n, float* __restrict__ inten) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned c = rgb[i];
  const float r = (float)((c >> 16) & 255u), g = (float)((c >> 8) & 255u), b = (float)(c & 255u);
  // grayscale = 0.00390625 * (0.114 b + 0.5870 g + 0.2989 r): a double expression rounded once
  inten[i] = (float)(0.00390625 * (0.114 * (double)b + 0.5870 * (double)g + 0.2989 * (double)r));
}

// x = A^-1 b by Eigen 3.2's ColPivHouseholderQR<Matrix3f> written out (see oracle header for the rules)
__device__ __forceinline__ void colpiv_qr_solve3(const float (&Ain)[3][3], const float (&bin)[3], float (&x)[3]) {
  float qr[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) qr[r][c] = Ain[r][c];
  float hco[3] = {0.f, 0.f, 0.f};
  int transp[3] = {0, 1, 2};
  float colsq[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) colsq[k] = qr[0][k] * qr[0][k] + (qr[1][k] * qr[1][k] + qr[2][k] * qr[2][k]);
  const float eps = 1.1920929e-07f;
  const float thr_helper = fmaxf(colsq[0], fmaxf(colsq[1], colsq[2])) * (eps * eps) / 3.0f;
  int nonzero = 3;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (nonzero != 3) continue;  // (the factorisation ended at an earlier k)
    int big = k;
#pragma unroll
    for (int c = k + 1; c < 3; ++c)
      if (colsq[c] > colsq[big]) big = c;
    // column `big` (runtime index): select without dynamic register indexing
    float colv[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) colv[r] = (big == 0) ? qr[r][0] : ((big == 1) ? qr[r][1] : qr[r][2]);
    float bsq = 0.f;
#pragma unroll
    for (int r = k; r < 3; ++r) bsq = (r == k) ? colv[r] * colv[r] : bsq + colv[r] * colv[r];
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (c == big) colsq[c] = bsq;
    if (bsq < thr_helper * (float)(3 - k)) {
      nonzero = k;
      continue;
    }
    transp[k] = big;
    if (k != big) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
        if (c == big) {
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const float t = qr[r][k];
            qr[r][k] = qr[r][c];
            qr[r][c] = t;
          }
          const float t = colsq[k];
          colsq[k] = colsq[c];
          colsq[c] = t;
        }
    }
    float tailsq = 0.f;
#pragma unroll
    for (int r = k + 1; r < 3; ++r) tailsq = (r == k + 1) ? qr[r][k] * qr[r][k] : tailsq + qr[r][k] * qr[r][k];
    const float c0 = qr[k][k];
    float tau, beta;
    if (k == 2 || tailsq == 0.f) {
      tau = 0.f;
      beta = c0;
#pragma unroll
      for (int r = k + 1; r < 3; ++r) qr[r][k] = 0.f;
    } else {
      beta = sqrtf(c0 * c0 + tailsq);
      if (c0 >= 0.f) beta = -beta;
      const float den = c0 - beta;
#pragma unroll
      for (int r = k + 1; r < 3; ++r) qr[r][k] = qr[r][k] / den;
      tau = (beta - c0) / beta;
    }
    hco[k] = tau;
    qr[k][k] = beta;
    if (k < 2) {
#pragma unroll
      for (int c = k + 1; c < 3; ++c) {
        float tmp = 0.f;
#pragma unroll
        for (int r = k + 1; r < 3; ++r) tmp = (r == k + 1) ? qr[r][k] * qr[r][c] : tmp + qr[r][k] * qr[r][c];
        tmp = tmp + qr[k][c];
        qr[k][c] = qr[k][c] - tau * tmp;
#pragma unroll
        for (int r = k + 1; r < 3; ++r) qr[r][c] = qr[r][c] - (tau * qr[r][k]) * tmp;
      }
    }
#pragma unroll
    for (int c = k + 1; c < 3; ++c) colsq[c] = colsq[c] - qr[k][c] * qr[k][c];
  }
  int perm[3] = {0, 1, 2};
#pragma unroll
  for (int k = 0; k < 3; ++k)
    if (k < nonzero) {
      // swap perm[k] and perm[transp[k]]
      const int tk = transp[k];
      const int a = perm[k];
      const int b = (tk == 0) ? perm[0] : ((tk == 1) ? perm[1] : perm[2]);
      perm[k] = b;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        if (c == tk && c != k) perm[c] = a;
    }
  x[0] = x[1] = x[2] = 0.f;
  if (nonzero == 0) return;
  float c[3] = {bin[0], bin[1], bin[2]};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (k >= nonzero) continue;
    if (k == 2) {
      c[2] = c[2] * (1.0f - hco[2]);
      continue;
    }
    float tmp = 0.f;
#pragma unroll
    for (int r = k + 1; r < 3; ++r) tmp = (r == k + 1) ? qr[r][k] * c[r] : tmp + qr[r][k] * c[r];
    tmp = tmp + c[k];
    c[k] = c[k] - hco[k] * tmp;
#pragma unroll
    for (int r = k + 1; r < 3; ++r) c[r] = c[r] - (hco[k] * qr[r][k]) * tmp;
  }
#pragma unroll
  for (int i = 2; i >= 0; --i) {
    if (i >= nonzero) continue;
    float sacc = c[i];
#pragma unroll
    for (int j = i + 1; j < 3; ++j)
      if (j < nonzero) sacc = sacc - qr[i][j] * c[j];
    c[i] = sacc / qr[i][i];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
    if (i < nonzero) {
#pragma unroll
      for (int t = 0; t < 3; ++t)
        if (perm[i] == t) x[t] = c[i];
    }
}

// IntensityGradientEstimation::computeFeature + the length rule of HarrisKeypoint6D::detectKeypoints; one thread per
// surface point, lists = the surface's own sorted lists, everything in ORIGINAL order
__global__ void __launch_bounds__(128)
harris6d_gradient_kernel(ListView L, const float4* __restrict__ surf, const float* __restrict__ inten,
                         const float4* __restrict__ nrm, int n, float* __restrict__ grad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float nanv = __int_as_float(0x7fc00000);
  long long b0;
  int m;
  L.row(i, b0, m);
  const float4 q = surf[i];
  float G[3] = {nanv, nanv, nanv};
  if (m > 0 && finite3(q.x, q.y, q.z)) {
    float cen[3] = {0.f, 0.f, 0.f}, mean_i = 0.f;
    for (int t = 0; t < m; ++t) {
      const int j = L.idx[b0 + t];
      const float4 p = surf[j];
      cen[0] += p.x;
      cen[1] += p.y;
      cen[2] += p.z;
      mean_i += inten[j];
    }
    const float fn = (float)m;
    const float inv_n = 1.0f / fn;
    cen[0] *= inv_n;
    cen[1] *= inv_n;
    cen[2] *= inv_n;
    mean_i /= fn;
    if (m >= 3) {
      float A[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}}, bv[3] = {0.f, 0.f, 0.f};
      for (int t = 0; t < m; ++t) {
        const int j = L.idx[b0 + t];
        const float4 p = surf[j];
        const float it = inten[j];
        if (!isfinite(p.x) || !isfinite(p.y) || !isfinite(p.z) || !isfinite(it)) continue;
        const float dx = p.x - cen[0], dy = p.y - cen[1], dz = p.z - cen[2];
        const float di = it - mean_i;
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
      colpiv_qr_solve3(A, bv, x);
      const float4 nr4 = nrm[i];
      const float nr[3] = {nr4.x, nr4.y, nr4.z};
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float mm = (r == c ? 1.0f : 0.0f) - nr[r] * nr[c];
          acc = (c == 0) ? mm * x[c] : acc + mm * x[c];
        }
        G[r] = acc;
      }
    }
  }
  // keep the direction of gradients with squared length > 200 (upstream's magic number), zero the rest (NaN included)
  float len = G[0] * G[0] + G[1] * G[1] + G[2] * G[2];
  if ((double)len > 200.0) {
    len = (float)(1.0 / sqrt((double)len));
    G[0] *= len;
    G[1] *= len;
    G[2] *= len;
  } else {
    G[0] = G[1] = G[2] = 0.f;
  }
  grad[3 * (size_t)i] = G[0];
  grad[3 * (size_t)i + 1] = G[1];
  grad[3 * (size_t)i + 2] = G[2];
}

// responseTomasi: 6x6 covariance of (normal, gradient) over the list, 4th smallest eigenvalue by cyclic Jacobi in double
__global__ void __launch_bounds__(64)
harris6d_response_kernel(ListView L, const float4* __restrict__ surf, const float4* __restrict__ nrm,
                         const float* __restrict__ grad, int n, float* __restrict__ resp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float r = 0.f;
  const float4 q = surf[i];
  if (finite3(q.x, q.y, q.z)) {
    long long b0;
    int m;
    L.row(i, b0, m);
    float co[21];
#pragma unroll
    for (int t = 0; t < 21; ++t) co[t] = 0.f;
    unsigned count = 0;
    for (int t = 0; t < m; ++t) {
      const int j = L.idx[b0 + t];
      const float4 nr = nrm[j];
      const float g0 = grad[3 * (size_t)j], g1 = grad[3 * (size_t)j + 1], g2 = grad[3 * (size_t)j + 2];
      if (!isfinite(nr.x) || !isfinite(g0)) continue;
      const float v[6] = {nr.x, nr.y, nr.z, g0, g1, g2};
      int e = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int c = a; c < 6; ++c) co[e++] += v[a] * v[c];
      ++count;
    }
    if (count > 0) {
      const float norm = (float)(1.0 / (double)(float)count);
#pragma unroll
      for (int t = 0; t < 21; ++t) co[t] *= norm;
    }
    const float trace = co[0] + co[6] + co[11] + co[15] + co[18] + co[20];
    if (trace != 0) {
      double A[6][6];
      {
        int e = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
          for (int c = a; c < 6; ++c) {
            A[a][c] = (double)co[e];
            A[c][a] = (double)co[e];
            ++e;
          }
      }
      for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0, dg = 0;
#pragma unroll
        for (int p = 0; p < 6; ++p) {
          dg += fabs(A[p][p]);
#pragma unroll
          for (int qq = p + 1; qq < 6; ++qq) off += fabs(A[p][qq]);
        }
        if (off <= 1e-300 || off <= 1e-18 * dg) break;
#pragma unroll
        for (int p = 0; p < 5; ++p)
#pragma unroll
          for (int qq = p + 1; qq < 6; ++qq) {
            if (A[p][qq] == 0.0) continue;
            const double theta = (A[qq][qq] - A[p][p]) / (2.0 * A[p][qq]);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
            for (int rr = 0; rr < 6; ++rr) {
              const double arp = A[rr][p], arq = A[rr][qq];
              A[rr][p] = c * arp - s * arq;
              A[rr][qq] = s * arp + c * arq;
            }
#pragma unroll
            for (int rr = 0; rr < 6; ++rr) {
              const double apr = A[p][rr], aqr = A[qq][rr];
              A[p][rr] = c * apr - s * aqr;
              A[qq][rr] = s * apr + c * aqr;
            }
          }
      }
      double w[6];
#pragma unroll
      for (int t = 0; t < 6; ++t) w[t] = A[t][t];
      // ascending order; only the value of rank 3 is needed: count how many are smaller (ties by position, as a
      // stable sort would place them)
      double w3 = 0.0;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        int rank = 0;
#pragma unroll
        for (int c = 0; c < 6; ++c) rank += (w[c] < w[a] || (w[c] == w[a] && c < a)) ? 1 : 0;
        if (rank == 3) w3 = w[a];
      }
      r = (float)w3;
    }
  }
  resp[i] = r;
}

// response of every surface point (original order).  Needs: surface colours (ctx->surf_rgb) and normals at `radius`.
int harris6d_response(Ctx* ctx, Grid* g, double radius, float* resp_dev_orig, float* grad_out_dev) {
  const int n = (int)ctx->n;
  if (n == 0) return 0;
  const bool saved = ctx->q_is_surface;
  ctx->q_is_surface = true;
  int rc = strict_lists_build(ctx, g, radius, 0);
  ctx->q_is_surface = saved;
  if (rc) return rc;
  PFX_CUDA(ctx->h6_inten.ensure((size_t)n * sizeof(float)));
  PFX_CUDA(ctx->h6_grad.ensure((size_t)n * 3 * sizeof(float)));
  PFX_LAUNCH(ctx, harris6d_intensity_kernel, div_up(n, 256), 256, 0, ctx->surf_rgb.as<unsigned>(), n, ctx->h6_inten.as<float>());
  PFX_LAUNCH(ctx, harris6d_gradient_kernel, div_up(n, 128), 128, 0, list_view(ctx), ctx->surf.as<float4>(),
             ctx->h6_inten.as<float>(), ctx->normals.as<float4>(), n, ctx->h6_grad.as<float>());
  PFX_LAUNCH(ctx, harris6d_response_kernel, div_up(n, 64), 64, 0, list_view(ctx), ctx->surf.as<float4>(),
             ctx->normals.as<float4>(), ctx->h6_grad.as<float>(), n, resp_dev_orig);
  if (grad_out_dev)
    PFX_CUDA(cudaMemcpyAsync(grad_out_dev, ctx->h6_grad.p, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  PFX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pfx
