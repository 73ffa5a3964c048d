// normals.cu — surface normals + curvature (replaces pcl::NormalEstimationOMP::compute as driven by
// reference tools.h:22-32 / features.h:187 / keypoints.h:298-308).
//
// Per query: neighbours (kNN list or fused radius scan of the cell stencil) -> 3x3 covariance of the
// QUERY-CENTRED coordinates, accumulated in double (12 DP ops per neighbour next to a 16-byte gather;
// PCL's own absolute-coordinate float sums are off by ~3e-3 from the exact answer) -> symmetric eigen solve
// (float Jacobi on the scaled matrix + one double refinement step) -> smallest eigenvector, flipped toward the viewpoint,
// curvature = l0 / trace.  A warp accumulates one query at a time and parks the reduced moments in
// lane t; after 32 queries every lane solves its own 3x3 problem, so the eigen solve runs at full
// SIMT width and the float4 results are written as one coalesced 512-byte row.
#include "internal.h"

namespace pfx {

constexpr int NWPB = 8;

struct Moments {
  double s[9];  // sum d (3), sum d d^T upper triangle (6); d = p - q formed exactly in double
  int n;
};

__device__ __forceinline__ void mom_add(Moments& m, float4 p, float4 q) {
  double dx = (double)p.x - (double)q.x, dy = (double)p.y - (double)q.y, dz = (double)p.z - (double)q.z;
  m.s[0] += dx; m.s[1] += dy; m.s[2] += dz;
  m.s[3] += dx * dx; m.s[4] += dx * dy; m.s[5] += dx * dz;
  m.s[6] += dy * dy; m.s[7] += dy * dz; m.s[8] += dz * dz;
  m.n += 1;
}

// covariance in double -> float Jacobi (cheap, ~1e-7) -> one double refinement step: Rayleigh
// quotient for l0, then the largest cross product of two rows of (C - l0 I) (pcl::eigen33's
// eigenvector construction) -> ~1e-12 of the double oracle unless the eigen-gap is ~1e-6 or less.
__device__ __forceinline__ float4 solve_normal(const Moments& m, float qx, float qy, float qz, float vx,
                                               float vy, float vz) {
  const float nanv = __int_as_float(0x7fc00000);
  if (m.n == 0) return make_float4(nanv, nanv, nanv, nanv);
  double inv = 1.0 / (double)m.n;
  double mx = m.s[0] * inv, my = m.s[1] * inv, mz = m.s[2] * inv;
  double c[6];
  c[0] = m.s[3] * inv - mx * mx;
  c[1] = m.s[4] * inv - mx * my;
  c[2] = m.s[5] * inv - mx * mz;
  c[3] = m.s[6] * inv - my * my;
  c[4] = m.s[7] * inv - my * mz;
  c[5] = m.s[8] * inv - mz * mz;
  double tr = c[0] + c[3] + c[5];
  double sc = fmax(fmax(fabs(c[0]), fabs(c[1])), fmax(fmax(fabs(c[2]), fabs(c[3])), fmax(fabs(c[4]), fabs(c[5]))));
  double isc = (sc > 1e-300) ? 1.0 / sc : 1.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) c[i] *= isc;
  float a[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) a[i] = (float)c[i];
  float w[3], v[3][3];
  eig_sym3<float>(a, w, v, 8);
  double n0 = v[0][0], n1 = v[1][0], n2 = v[2][0];
  // Rayleigh quotient in double
  double cx = c[0] * n0 + c[1] * n1 + c[2] * n2;
  double cy = c[1] * n0 + c[3] * n1 + c[4] * n2;
  double cz = c[2] * n0 + c[4] * n1 + c[5] * n2;
  double l0 = (n0 * cx + n1 * cy + n2 * cz) / (n0 * n0 + n1 * n1 + n2 * n2);
  // rows of (C - l0 I)
  double r0[3] = {c[0] - l0, c[1], c[2]}, r1[3] = {c[1], c[3] - l0, c[4]}, r2[3] = {c[2], c[4], c[5] - l0};
  double e0[3] = {r0[1] * r1[2] - r0[2] * r1[1], r0[2] * r1[0] - r0[0] * r1[2], r0[0] * r1[1] - r0[1] * r1[0]};
  double e1[3] = {r0[1] * r2[2] - r0[2] * r2[1], r0[2] * r2[0] - r0[0] * r2[2], r0[0] * r2[1] - r0[1] * r2[0]};
  double e2[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
  double l_0 = e0[0] * e0[0] + e0[1] * e0[1] + e0[2] * e0[2];
  double l_1 = e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2];
  double l_2 = e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2];
  double bx = e0[0], by = e0[1], bz = e0[2], bl = l_0;
  if (l_1 > bl) { bx = e1[0]; by = e1[1]; bz = e1[2]; bl = l_1; }
  if (l_2 > bl) { bx = e2[0]; by = e2[1]; bz = e2[2]; bl = l_2; }
  if (bl > 1e-280) {  // otherwise (C - l0 I) has rank < 2: keep the Jacobi vector
    double il = rsqrt(bl);
    bx *= il; by *= il; bz *= il;
    if (bx * n0 + by * n1 + bz * n2 < 0) { bx = -bx; by = -by; bz = -bz; }
    n0 = bx; n1 = by; n2 = bz;
  }
  double trs = c[0] + c[3] + c[5];
  double curv = (tr != 0.0 && trs != 0.0) ? fabs(l0 / trs) : 0.0;
  // flipNormalTowardsViewpoint
  double dp = ((double)vx - (double)qx) * n0 + ((double)vy - (double)qy) * n1 + ((double)vz - (double)qz) * n2;
  if (dp < 0) { n0 = -n0; n1 = -n1; n2 = -n2; }
  return make_float4((float)n0, (float)n1, (float)n2, (float)curv);
}

// rows: DENSE -> row = sorted surface position; out_rows[row] always written (when non-null);
// out_orig (DENSE only) receives the same rows scattered to the original order.
template <bool DENSE, bool USE_LIST>
__global__ void __launch_bounds__(NWPB * 32)
normals_kernel(GridDev g, const float4* __restrict__ queries, int nq, float r2,
               const int* __restrict__ lists, int k, float vx, float vy, float vz,
               float4* __restrict__ out_rows, float4* __restrict__ out_orig) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int qbase = (blockIdx.x * NWPB + wid) * 32;
  if (qbase >= nq) return;
  const int n_valid = g.gp->n_valid;
  Moments mine;
#pragma unroll
  for (int i = 0; i < 9; ++i) mine.s[i] = 0.0;
  mine.n = 0;
  float4 myq = make_float4(0.f, 0.f, 0.f, 0.f);
  const int qend = min(32, nq - qbase);
  for (int t = 0; t < qend; ++t) {
    const int qi = qbase + t;
    float4 q = DENSE ? g.pts[qi] : queries[qi];
    Moments m;
#pragma unroll
    for (int i = 0; i < 9; ++i) m.s[i] = 0.0;
    m.n = 0;
    bool ok = finite3(q.x, q.y, q.z) && (!DENSE || qi < n_valid);
    if (ok) {
      if (USE_LIST) {
        for (int c = lane; c < k; c += 32) {
          int j = lists[(size_t)qi * k + c];
          if (j >= 0) {
            float4 p = g.pts[j];
            mom_add(m, p, q);
          }
        }
      } else {
        CellBlock blk = DENSE ? stencil_of_point(g, qi, lane) : stencil_of_pos(g, q.x, q.y, q.z, lane);
        for (int base = 0; base < blk.total; base += 32) {
          int c = base + lane;
          bool valid = c < blk.total;
          int j = block_candidate(blk, valid ? c : 0);
          if (valid) {
            float4 p = g.pts[j];
            float d2 = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);
            if (d2 < r2) mom_add(m, p, q);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 9; ++i) m.s[i] = warp_sum(m.s[i]);
      m.n = warp_sum(m.n);
    }
    if (lane == t) {
      mine = m;
      myq = q;
    }
  }
  if (lane < qend) {
    const int qi = qbase + lane;
    float4 r = solve_normal(mine, myq.x, myq.y, myq.z, vx, vy, vz);
    if (out_rows) out_rows[qi] = r;
    if (DENSE && out_orig) out_orig[__float_as_int(myq.w)] = r;
  }
}

// out_query_order: device float4 rows in caller order (may be null).  Dense runs also refresh the
// context's surface normals (original order) and the sorted copy for grid g.
int normals_compute(Ctx* ctx, Grid* g, double radius, int k, float4* out_query_order) {
  const int nq = (int)ctx->num_queries();
  const bool dense = ctx->q_is_surface;
  float r2 = (float)(radius * radius);
  const int* lists = nullptr;
  if (k > 0) {
    PFX_TRY(knn_lists(ctx, g, k, false));
    lists = ctx->knn_idx.as<int>();
  }
  if (dense) {
    PFX_CUDA(ctx->normals.ensure(std::max<size_t>(ctx->n, 1) * sizeof(float4)));
    PFX_CUDA(ctx->normals_sorted.ensure(std::max<size_t>(ctx->n, 1) * sizeof(float4)));
  }
  if (nq == 0) return 0;
  const int blocks = div_up(nq, NWPB * 32);
  if (dense) {
    float4* sorted = ctx->normals_sorted.as<float4>();
    float4* orig = ctx->normals.as<float4>();
    if (k > 0)
      PFX_LAUNCH(ctx, (normals_kernel<true, true>), blocks, NWPB * 32, 0, g->view(), nullptr, nq, r2, lists, k,
                 ctx->vp[0], ctx->vp[1], ctx->vp[2], sorted, orig);
    else
      PFX_LAUNCH(ctx, (normals_kernel<true, false>), blocks, NWPB * 32, 0, g->view(), nullptr, nq, r2, nullptr,
                 0, ctx->vp[0], ctx->vp[1], ctx->vp[2], sorted, orig);
    ctx->have_normals = true;
    ctx->normals_version++;
    ctx->normals_sorted_for = g;
    ctx->normals_sorted_version = ctx->normals_version;
    if (out_query_order)
      PFX_CUDA(cudaMemcpyAsync(out_query_order, orig, (size_t)nq * sizeof(float4), cudaMemcpyDeviceToDevice,
                               ctx->stream));
  } else {
    if (!out_query_order) return 0;
    if (k > 0)
      PFX_LAUNCH(ctx, (normals_kernel<false, true>), blocks, NWPB * 32, 0, g->view(), ctx->qry.as<float4>(), nq,
                 r2, lists, k, ctx->vp[0], ctx->vp[1], ctx->vp[2], out_query_order, nullptr);
    else
      PFX_LAUNCH(ctx, (normals_kernel<false, false>), blocks, NWPB * 32, 0, g->view(), ctx->qry.as<float4>(), nq,
                 r2, nullptr, 0, ctx->vp[0], ctx->vp[1], ctx->vp[2], out_query_order, nullptr);
  }
  PFX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pfx
