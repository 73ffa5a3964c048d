// normals.cu — surface normals + curvature (replaces pcl::NormalEstimationOMP::compute as driven by
// reference tools.h:22-32 / features.h:187 / keypoints.h:298-308).
//
// Per query: neighbours (kNN list or fused radius scan of the cell stencil) -> 3x3 covariance of the
// QUERY-CENTRED coordinates in float32 (the centring is what keeps float32 at ~1e-7 of the double
// answer; PCL's own absolute-coordinate float sums are off by ~3e-3) -> symmetric eigen solve
// (cyclic Jacobi on the scaled matrix) -> smallest eigenvector, flipped toward the viewpoint,
// curvature = l0 / trace.  A warp accumulates one query at a time and parks the reduced moments in
// lane t; after 32 queries every lane solves its own 3x3 problem, so the eigen solve runs at full
// SIMT width and the float4 results are written as one coalesced 512-byte row.
#include "internal.h"

namespace pfx {

constexpr int NWPB = 8;

struct Moments {
  float s[9];  // sum d (3), sum d d^T upper triangle (6)
  int n;
};

__device__ __forceinline__ void mom_add(Moments& m, float dx, float dy, float dz) {
  m.s[0] += dx; m.s[1] += dy; m.s[2] += dz;
  m.s[3] += dx * dx; m.s[4] += dx * dy; m.s[5] += dx * dz;
  m.s[6] += dy * dy; m.s[7] += dy * dz; m.s[8] += dz * dz;
  m.n += 1;
}

__device__ __forceinline__ float4 solve_normal(const Moments& m, float qx, float qy, float qz, float vx,
                                               float vy, float vz) {
  const float nanv = __int_as_float(0x7fc00000);
  if (m.n == 0) return make_float4(nanv, nanv, nanv, nanv);
  float inv = 1.0f / (float)m.n;
  float mx = m.s[0] * inv, my = m.s[1] * inv, mz = m.s[2] * inv;
  float c[6];
  c[0] = m.s[3] * inv - mx * mx;
  c[1] = m.s[4] * inv - mx * my;
  c[2] = m.s[5] * inv - mx * mz;
  c[3] = m.s[6] * inv - my * my;
  c[4] = m.s[7] * inv - my * mz;
  c[5] = m.s[8] * inv - mz * mz;
  float tr = c[0] + c[3] + c[5];
  float sc = fmaxf(fmaxf(fabsf(c[0]), fabsf(c[1])), fmaxf(fmaxf(fabsf(c[2]), fabsf(c[3])), fmaxf(fabsf(c[4]), fabsf(c[5]))));
  float isc = (sc > 1e-37f) ? 1.0f / sc : 1.0f;
  float a[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) a[i] = c[i] * isc;
  float w[3], v[3][3];
  eig_sym3<float>(a, w, v, 8);
  float nx = v[0][0], ny = v[1][0], nz = v[2][0];
  float trs = a[0] + a[3] + a[5];
  float curv = (tr != 0.f && trs != 0.f) ? fabsf(w[0] / trs) : 0.f;
  // flipNormalTowardsViewpoint
  float dp = (vx - qx) * nx + (vy - qy) * ny + (vz - qz) * nz;
  if (dp < 0.f) {
    nx = -nx; ny = -ny; nz = -nz;
  }
  return make_float4(nx, ny, nz, curv);
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
  for (int i = 0; i < 9; ++i) mine.s[i] = 0.f;
  mine.n = 0;
  float4 myq = make_float4(0.f, 0.f, 0.f, 0.f);
  const int qend = min(32, nq - qbase);
  for (int t = 0; t < qend; ++t) {
    const int qi = qbase + t;
    float4 q = DENSE ? g.pts[qi] : queries[qi];
    Moments m;
#pragma unroll
    for (int i = 0; i < 9; ++i) m.s[i] = 0.f;
    m.n = 0;
    bool ok = finite3(q.x, q.y, q.z) && (!DENSE || qi < n_valid);
    if (ok) {
      if (USE_LIST) {
        for (int c = lane; c < k; c += 32) {
          int j = lists[(size_t)qi * k + c];
          if (j >= 0) {
            float4 p = g.pts[j];
            mom_add(m, p.x - q.x, p.y - q.y, p.z - q.z);
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
            if (d2 < r2) mom_add(m, p.x - q.x, p.y - q.y, p.z - q.z);
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
