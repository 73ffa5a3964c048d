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

}  // namespace pfx
