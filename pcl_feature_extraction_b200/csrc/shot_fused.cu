// shot_fused.cu — SHOT local reference frame + SHOT352 in ONE kernel for neighbourhoods of <= NCAP
// points (replaces SHOTEstimationOMP::compute incl. its internal SHOTLocalReferenceFrameEstimation,
// reference evaluation.cpp:770-775).
//
// A warp owns 32 queries.  Phases A-C run one query per LANE (no shuffles, no warp reductions): A walks the
// 27 cells of the query's stencil once, writes the neighbours that pass d2 < r^2 into a [position][lane]
// table in shared memory and accumulates the (R - d)-weighted scatter matrix in double; B solves the 32
// eigen problems together (float Jacobi + one double refinement step per axis); C takes the sign votes
// from the cached list (vote ties: the warp ranks that query's list for PCL's median fallback).  Phase D
// runs one query per WARP, one neighbour per lane: the 352-bin histogram (int32 fixed point in shared
// memory, order-independent => bit-reproducible), L2 normalisation and the coalesced 1444-byte row.
// Neighbouring queries read the same cells at the same time, so phase A's loads are warp broadcasts or
// L1 hits.  Queries with more than NCAP neighbours go to the generic kernels (shot.cu) over a work list.
#include "internal.h"
#include "shot_common.cuh"

namespace pfx {

constexpr int NCAP = 48;
constexpr int FS_WPB = 4;
constexpr int NPAD = 33;  // row pitch of the neighbour table: [position][query lane], conflict-free both ways

struct FusedSmem {
  int nbr[NCAP][NPAD];
  int hist[352];
  unsigned long long keys[NCAP];
};

// cell id of stencil slot l (0..26) of a query: adjacency row for surface points, hash probe otherwise
template <bool DENSE>
__device__ __forceinline__ int stencil_cell(const GridDev& g, const GridParams& P, const int* __restrict__ adj, int cx,
                                            int cy, int cz, int l) {
  if (DENSE) return adj[l];
  const int x2 = cx + l % 3 - 1, y2 = cy + (l / 3) % 3 - 1, z2 = cz + l / 9 - 1;
  if (x2 < 0 || x2 >= P.nx || y2 < 0 || y2 >= P.ny || z2 < 0 || z2 >= P.nz) return -1;
  return hash_lookup(g, morton3(x2, y2, z2));
}

template <bool DENSE>
__global__ void __launch_bounds__(FS_WPB * 32)
shot_fused_kernel(GridDev g, const float4* __restrict__ queries, int nq, const float4* __restrict__ nrm, float r2,
                  double R, float* __restrict__ out, size_t stride, int* __restrict__ wl_count, int* __restrict__ wl) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  FusedSmem* S = reinterpret_cast<FusedSmem*>(smem_raw) + wid;
  const GridParams P = *g.gp;
  const int n_valid = P.n_valid;
  const float nanv = __int_as_float(0x7fc00000);
  const float Rf = (float)R;
  const float t12 = shot_d2_threshold(R / 2);  // d2 > t12  <=>  sqrt((double)d2) > R / 2
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < 11; ++i) S->hist[lane + 32 * i] = 0;
  __syncwarp();
  for (int qbase = (blockIdx.x * FS_WPB + wid) * 32; qbase < nq; qbase += gridDim.x * FS_WPB * 32) {
    // ---------------- phase A (lane = query): one pass over the 3x3x3 stencil builds the neighbour list
    // in shared memory and the (R - d)-weighted scatter matrix of getLocalRF in double
    const int qi = qbase + lane;
    const bool have_q = qi < nq;
    const float4 q = have_q ? (DENSE ? g.pts[qi] : queries[qi]) : make_float4(0.f, 0.f, 0.f, 0.f);
    const bool q_ok = have_q && finite3(q.x, q.y, q.z) && (!DENSE || qi < n_valid) && n_valid > 0;
    double m6[6] = {0, 0, 0, 0, 0, 0};
    double msw = 0.0;
    int n_all = 0, n_val = 0;
    if (q_ok) {
      const int* adj = nullptr;
      int cx = 0, cy = 0, cz = 0;
      if (DENSE) {
        adj = g.cell_nbr + (size_t)g.pt_cell[qi] * 27;
      } else {
        cx = cell_coord(q.x, P.ox, P.inv_e, P.nx);
        cy = cell_coord(q.y, P.oy, P.inv_e, P.ny);
        cz = cell_coord(q.z, P.oz, P.inv_e, P.nz);
      }
      for (int l = 0; l < 27; ++l) {
        const int c = stencil_cell<DENSE>(g, P, adj, cx, cy, cz, l);
        if (c < 0) continue;
        const int j1 = g.cell_start[c + 1];
        for (int j = g.cell_start[c]; j < j1; ++j) {
          const float4 p = g.pts[j];
          const float d2 = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);
          if (d2 < r2) {
            if (n_all < NCAP) S->nbr[n_all][lane] = j;
            ++n_all;
          }
        }
      }
      // scatter matrix from the compact list (uniform trip counts: no lanes idle on rejected candidates)
      const int n_list = min(n_all, NCAP);
      for (int c = 0; c < n_list; ++c) {
        const float4 p = g.pts[S->nbr[c][lane]];
        if (!(p.x == q.x && p.y == q.y && p.z == q.z)) {
          const float d2 = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);
          const double vx = (double)__fsub_rn(p.x, q.x), vy = (double)__fsub_rn(p.y, q.y),
                       vz = (double)__fsub_rn(p.z, q.z);
          // sqrt((double)d2): float rsqrt seed + one Newton step in double (error ~1e-14 relative)
          const double xd = (double)d2, rs = (d2 > 0.f) ? (double)rsqrtf(d2) : 0.0;
          double sq = xd * rs;
          sq = fma(0.5 * rs, fma(-sq, sq, xd), sq);
          const double w = R - sq;
          m6[0] += w * (vx * vx); m6[1] += w * (vx * vy); m6[2] += w * (vx * vz);
          m6[3] += w * (vy * vy); m6[4] += w * (vy * vz); m6[5] += w * (vz * vz);
          msw += w;
          ++n_val;
        }
      }
    }
    const bool overflow = n_all > NCAP;  // handed to the generic kernels
    // ---------------- phase B: every lane solves its own 3x3 eigen problem
    double x[3] = {0, 0, 0}, z[3] = {0, 0, 0};
    bool good = false;
    if (q_ok && !overflow && n_val >= 5) {
      double a[6];
      const double inv = 1.0 / msw;
#pragma unroll
      for (int i = 0; i < 6; ++i) a[i] = m6[i] * inv;
      good = eig_extreme_refined(a, x, z);
    }
    // ---------------- phase C (lane = query): sign votes over the cached list
    int votex = 1, votez = 1;
    if (good) {
      int px = 0, pz = 0;
      for (int c = 0; c < n_all; ++c) {
        const float4 p = g.pts[S->nbr[c][lane]];
        if (!(p.x == q.x && p.y == q.y && p.z == q.z)) {
          const double vx = (double)__fsub_rn(p.x, q.x), vy = (double)__fsub_rn(p.y, q.y),
                       vz = (double)__fsub_rn(p.z, q.z);
          if (vx * x[0] + vy * x[1] + vz * x[2] >= 0) ++px;
          if (vx * z[0] + vy * z[1] + vz * z[2] >= 0) ++pz;
        }
      }
      votex = 2 * px - n_val;
      votez = 2 * pz - n_val;
    }
    bool fx = votex < 0, fz = votez < 0;
    // vote ties: PCL's fallback needs the neighbours' ranks in (d2, index) order -> the warp ranks the list of
    // each tied query together
    unsigned ties = __ballot_sync(FULL, good && (votex == 0 || votez == 0));
    while (ties) {
      const int t = __ffs(ties) - 1;
      ties &= ties - 1;
      const int n_t = __shfl_sync(FULL, n_all, t);
      const float qx = __shfl_sync(FULL, q.x, t), qy = __shfl_sync(FULL, q.y, t), qz = __shfl_sync(FULL, q.z, t);
      double xs[3], zs[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        xs[i] = __shfl_sync(FULL, x[i], t);
        zs[i] = __shfl_sync(FULL, z[i], t);
      }
      int cntv = 0;
      for (int c0 = 0; c0 < n_t; c0 += 32) {
        const int c = c0 + lane;
        bool v = false;
        unsigned long long key = 0;
        if (c < n_t) {
          const float4 p = g.pts[S->nbr[c][t]];
          v = !(p.x == qx && p.y == qy && p.z == qz);
          const float d2 = dist2_flann(qx, qy, qz, p.x, p.y, p.z);
          key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)__float_as_int(p.w);
        }
        const unsigned m = __ballot_sync(FULL, v);
        if (v) S->keys[cntv + __popc(m & lt)] = key;
        cntv += __popc(m);
      }
      __syncwarp();
      const int med = cntv / 2;
      int plx = 0, plz = 0;
      for (int a = lane; a < cntv; a += 32) {
        const unsigned long long ka = S->keys[a];
        int rank = 0;
        for (int b = 0; b < cntv; ++b) rank += (S->keys[b] < ka) ? 1 : 0;
        if (rank >= med - 2 && rank <= med + 2) {
          const float4 p = g.pts[g.inv_perm[(int)(unsigned)(ka & 0xffffffffull)]];
          const double vx = (double)__fsub_rn(p.x, qx), vy = (double)__fsub_rn(p.y, qy), vz = (double)__fsub_rn(p.z, qz);
          if (vx * xs[0] + vy * xs[1] + vz * xs[2] > 0) ++plx;
          if (vx * zs[0] + vy * zs[1] + vz * zs[2] > 0) ++plz;
        }
      }
      plx = warp_sum(plx);
      plz = warp_sum(plz);
      if (lane == t) {
        if (votex == 0) fx = plx < 3;
        if (votez == 0) fz = plz < 3;
      }
      __syncwarp();
    }
    float rf[9];
    {
      if (fx) { x[0] = -x[0]; x[1] = -x[1]; x[2] = -x[2]; }
      if (fz) { z[0] = -z[0]; z[1] = -z[1]; z[2] = -z[2]; }
      lrf_to_float9(x, z, rf);
    }
    // ---------------- phase D (warp per query, lane = neighbour): 352-bin histogram, normalise, write the row
    const int qend = min(32, nq - qbase);
    for (int t = 0; t < qend; ++t) {
      const int n_t = __shfl_sync(FULL, n_all, t);
      const bool ok_t = __shfl_sync(FULL, (int)q_ok, t);
      const bool of_t = __shfl_sync(FULL, (int)overflow, t);
      const bool gd = __shfl_sync(FULL, (int)good, t);
      const float4 qt = make_float4(__shfl_sync(FULL, q.x, t), __shfl_sync(FULL, q.y, t), __shfl_sync(FULL, q.z, t),
                                    __shfl_sync(FULL, q.w, t));
      const size_t row = DENSE ? (size_t)__float_as_int(qt.w) : (size_t)(qbase + t);
      float* o = out + row * stride;
      if (of_t) {
        if (lane == 0) wl[atomicAdd(wl_count, 1)] = qbase + t;
        continue;
      }
      if (!ok_t || n_t <= 0 || !gd) {  // non-finite query, no neighbours, or NaN frame: all-NaN row
        for (int c = lane; c < 361; c += 32) o[c] = nanv;
        continue;
      }
      float rft[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) rft[i] = __shfl_sync(FULL, rf[i], t);
      if (lane < 9) {
        float v = rft[0];
#pragma unroll
        for (int i = 1; i < 9; ++i) v = (lane == i) ? rft[i] : v;
        o[352 + lane] = v;
      }
      if (n_t < 5) {  // computePointSHOT: too few neighbours -> NaN descriptor, frame kept
        for (int c = lane; c < 352; c += 32) o[c] = nanv;
        continue;
      }
      const float scale = shot_scale(n_t);  // (the histogram is zero here: cleared below as it is read out)
      for (int c = lane; c < n_t; c += 32) {
        const int j = S->nbr[c][t];
        const float4 p = g.pts[j];
        const float d2 = dist2_flann(qt.x, qt.y, qt.z, p.x, p.y, p.z);
        shot_accumulate_neighbor_f(S->hist, scale, qt, p, d2, nrm[j], rft, Rf, t12);
      }
      __syncwarp();
      float acc = 0.f;
      float hv[11];
#pragma unroll
      for (int i = 0; i < 11; ++i) {
        hv[i] = (float)S->hist[lane + 32 * i];  // fixed point: the common scale cancels in the normalisation
        S->hist[lane + 32 * i] = 0;
        acc = fmaf(hv[i], hv[i], acc);
      }
      acc = warp_sum(acc);
      const float inv_n = rsqrtf(acc);
#pragma unroll
      for (int i = 0; i < 11; ++i) o[lane + 32 * i] = hv[i] * inv_n;
      __syncwarp();
    }
    __syncwarp();
  }
}

int shot_lrf_worklist(Ctx* ctx, Grid* g, double radius, float* rf9_dev, const int* wl, const int* wl_count);
int shot_worklist(Ctx* ctx, Grid* g, double radius, const float* rf9_dev, float* out_dev, size_t stride_floats,
                  const int* wl, const int* wl_count);

// Fused LRF + SHOT352 for the current queries; neighbourhoods larger than NCAP go through the
// generic kernels (shot.cu) over a device-side work list.
int shot_fused_compute(Ctx* ctx, Grid* g, double radius, float* out_dev, size_t stride_floats) {
  const int nq = (int)ctx->num_queries();
  if (nq == 0) return 0;
  const float r2 = (float)(radius * radius);
  const float4* nrm = nullptr;
  PFX_TRY(normals_sorted_for_grid(ctx, g, &nrm));
  if (!ctx->q_is_surface && nq < ctx->sm_count * 512) {
    // keypoint queries (the reference's use: evaluation.cpp:770-775 at r = 5 cm, hundreds of neighbours each):
    // too few queries for the lane-per-query kernel; the generic kernels give every query its own warp
    PFX_CUDA(ctx->tmp2.ensure((size_t)nq * 9 * sizeof(float)));
    PFX_TRY(shot_lrf_compute(ctx, g, radius, ctx->tmp2.as<float>(), nullptr));
    return shot_compute(ctx, g, radius, ctx->tmp2.as<float>(), out_dev, stride_floats);
  }
  PFX_CUDA(ctx->worklist2.ensure(((size_t)nq + 16) * sizeof(int)));
  int* wl_count = ctx->worklist2.as<int>();
  int* wl = wl_count + 16;
  PFX_CUDA(cudaMemsetAsync(wl_count, 0, 16 * sizeof(int), ctx->stream));
  const size_t smem = sizeof(FusedSmem) * FS_WPB;
  if (!ctx->smem_attr_shot_fused) {
    PFX_CUDA(cudaFuncSetAttribute(shot_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PFX_CUDA(cudaFuncSetAttribute(shot_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctx->smem_attr_shot_fused = true;
  }
  const int blocks = std::min(div_up(nq, FS_WPB * 32), ctx->sm_count * 16);
  if (ctx->q_is_surface)
    PFX_LAUNCH(ctx, shot_fused_kernel<true>, blocks, FS_WPB * 32, smem, g->view(), nullptr, nq, nrm, r2, radius, out_dev,
               stride_floats, wl_count, wl);
  else
    PFX_LAUNCH(ctx, shot_fused_kernel<false>, blocks, FS_WPB * 32, smem, g->view(), ctx->qry.as<float4>(), nq, nrm, r2,
               radius, out_dev, stride_floats, wl_count, wl);
  PFX_CUDA(cudaGetLastError());
  // neighbourhoods beyond NCAP: frames then descriptors with the generic kernels, work-list driven
  PFX_CUDA(ctx->tmp2.ensure((size_t)nq * 9 * sizeof(float)));
  PFX_TRY(shot_lrf_worklist(ctx, g, radius, ctx->tmp2.as<float>(), wl, wl_count));
  PFX_TRY(shot_worklist(ctx, g, radius, ctx->tmp2.as<float>(), out_dev, stride_floats, wl, wl_count));
  return 0;
}

}  // namespace pfx
