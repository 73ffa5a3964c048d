// normals.cu — surface normals + curvature (replaces pcl::NormalEstimationOMP::compute as driven by
// reference tools.h:22-32 / features.h:187 / keypoints.h:298-308).
//
// Per query: neighbours (kNN list or fused radius scan of the cell stencil) -> 3x3 covariance of the
// QUERY-CENTRED coordinates, accumulated in double (12 DP ops per neighbour next to a 16-byte gather;
// PCL's own absolute-coordinate float sums are ~3e-3 from the exact answer) -> float Jacobi on the
// scaled matrix + one double refinement step -> smallest eigenvector, flipped toward the viewpoint,
// curvature = l0 / trace (normals_solve.cuh).
//
// Dense kNN normals take the cell-tile path (knn_tile.cu: selection and moments in one pass over
// shared memory).  This file holds the generic kernel: one warp per query, the reduced moments of
// query t parked in lane t, and after 32 queries every lane solves its own 3x3 problem so the eigen
// solve runs at full SIMT width.  With a work list it runs persistently over the queries the tile path handed back.
#include "internal.h"
#include "normals_solve.cuh"

namespace pfx {

constexpr int NWPB = 8;

// rows: DENSE -> row = sorted surface position; out_rows[row] written when non-null;
// out_orig (DENSE only) receives the same rows scattered to the original order.
template <bool DENSE, bool USE_LIST>
__global__ void __launch_bounds__(NWPB * 32)
normals_kernel(GridDev g, const float4* __restrict__ queries, int nq, float r2, const int* __restrict__ lists, int k,
               float vx, float vy, float vz, float4* __restrict__ out_rows, float4* __restrict__ out_orig,
               const int* __restrict__ worklist, const int* __restrict__ wl_count) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int limit = worklist ? *wl_count : nq;
  const int n_valid = g.gp->n_valid;
  // queries per warp before the lanes solve: 32 fills the SIMT width of the eigen solve; a work list (a few hundred
  // queries the tile path handed back) is spread four to a warp instead, so that it does not end on a few busy warps
  const int qpw = worklist ? 4 : 32;
  for (int qbase = (blockIdx.x * NWPB + wid) * qpw; qbase < limit; qbase += gridDim.x * NWPB * qpw) {
    const int qend = min(qpw, limit - qbase);
    Moments mine;
#pragma unroll
    for (int i = 0; i < 9; ++i) mine.s[i] = 0.0;
    mine.n = 0;
    float4 myq = make_float4(0.f, 0.f, 0.f, 0.f);
    int myqi = 0;
    for (int t = 0; t < qend; ++t) {
      const int qi = worklist ? worklist[qbase + t] : qbase + t;
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
            if (j >= 0) mom_add(m, g.pts[j], q);
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
        myqi = qi;
      }
    }
    if (lane < qend) {
      float4 r = solve_normal_m9(mine.s, mine.n, myq.x, myq.y, myq.z, vx, vy, vz);
      if (out_rows) out_rows[myqi] = r;
      if (DENSE && out_orig) out_orig[__float_as_int(myq.w)] = r;
    }
  }
}

// out_query_order: device float4 rows in caller order (may be null).  Dense runs also refresh the
// context's surface normals (original order) and the sorted copy for grid g.
int normals_compute(Ctx* ctx, Grid* g, double radius, int k, float4* out_query_order) {
  const int nq = (int)ctx->num_queries();
  const bool dense = ctx->q_is_surface;
  float r2 = (float)(radius * radius);
  if (dense) {
    PFX_CUDA(ctx->normals.ensure(std::max<size_t>(ctx->n, 1) * sizeof(float4)));
    PFX_CUDA(ctx->normals_sorted.ensure(std::max<size_t>(ctx->n, 1) * sizeof(float4)));
  }
  if (nq == 0) {
    if (dense) {  // an empty surface has (zero) normals: later stages see their precondition met
      ctx->have_normals = true;
      ctx->normals_version++;
    }
    return 0;
  }
  const int blocks = div_up(nq, NWPB * 32);
  if (dense) {
    float4* sorted = ctx->normals_sorted.as<float4>();
    float4* orig = ctx->normals.as<float4>();
    if (k > 0) {
      // tile path: selection + moments fused; the generic kernel only redoes flagged queries
      PFX_TRY(knn_tile_lists(ctx, g, k, true));
      PFX_LAUNCH(ctx, (normals_kernel<true, true>), ctx->sm_count, NWPB * 32, 0, g->view(), nullptr, nq, r2,
                 ctx->knn_idx.as<int>(), k, ctx->vp[0], ctx->vp[1], ctx->vp[2], sorted, orig,
                 ctx->worklist.as<int>() + 16, ctx->worklist.as<int>());
    } else {
      PFX_LAUNCH(ctx, (normals_kernel<true, false>), blocks, NWPB * 32, 0, g->view(), nullptr, nq, r2, nullptr, 0,
                 ctx->vp[0], ctx->vp[1], ctx->vp[2], sorted, orig, nullptr, nullptr);
    }
    ctx->have_normals = true;
    ctx->normals_version++;
    ctx->normals_sorted_for = g;
    ctx->normals_sorted_version = ctx->normals_version;
    if (out_query_order)
      PFX_CUDA(cudaMemcpyAsync(out_query_order, orig, (size_t)nq * sizeof(float4), cudaMemcpyDeviceToDevice,
                               ctx->stream));
  } else {
    if (!out_query_order) return 0;
    const int* lists = nullptr;
    if (k > 0) {
      PFX_TRY(knn_lists(ctx, g, k, false));
      lists = ctx->knn_idx.as<int>();
      PFX_LAUNCH(ctx, (normals_kernel<false, true>), blocks, NWPB * 32, 0, g->view(), ctx->qry.as<float4>(), nq, r2,
                 lists, k, ctx->vp[0], ctx->vp[1], ctx->vp[2], out_query_order, nullptr, nullptr, nullptr);
    } else {
      PFX_LAUNCH(ctx, (normals_kernel<false, false>), blocks, NWPB * 32, 0, g->view(), ctx->qry.as<float4>(), nq, r2,
                 nullptr, 0, ctx->vp[0], ctx->vp[1], ctx->vp[2], out_query_order, nullptr, nullptr, nullptr);
    }
  }
  PFX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pfx
