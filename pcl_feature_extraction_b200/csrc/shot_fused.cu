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
//
// ROWS variant: when the dense k-search of the same surface is resident (normals / FPFH ran with k <= 32), phase A
// does not search at all.  A k-NN row that holds an entry with d2 >= r^2 contains EVERY point nearer than r (the
// set is closed under "nearer than its farthest member"), so the radius neighbourhood is the row filtered by
// d2 < r^2: the warp reads the 32 rows of its queries as one contiguous block and compacts them with a ballot.  Rows
// without such an entry (32 or more points inside the radius, ~3 % of the 2^20-point sheet) go to a work list that
// the stencil variant walks on the k-search grid itself, so no second voxel hash is built for the radius.
#include "internal.h"
#include "shot_common.cuh"

namespace pfx {

constexpr int NCAP = 48;       // neighbours a query may have in the stencil variants
constexpr int NCAP_ROWS = 32;  // ... and in the rows variant (k <= 32): a smaller table, more warps per SM
constexpr int FS_WPB = 4;
#ifndef PFX_SHOT_A_UNROLL
#define PFX_SHOT_A_UNROLL 8
#endif
constexpr int SHOT_A_UNROLL = PFX_SHOT_A_UNROLL;  // rows of phase A (rows variant) in flight per warp
#ifndef PFX_SHOT_BC_UNROLL
#define PFX_SHOT_BC_UNROLL 1
#endif
constexpr int SHOT_BC_UNROLL = PFX_SHOT_BC_UNROLL;  // neighbours of the lane-per-query loops (scatter matrix, votes) in flight
constexpr int NPAD = 33;  // row pitch of the neighbour table: [position][query lane], conflict-free both ways

template <int CAP>
struct FusedSmem {
  int nbr[CAP][NPAD];
  int hist[352];
  unsigned long long keys[CAP];
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

// ROWS: neighbour sets from the resident k-search rows (rows_idx / rows_d2, k entries per sorted query).
// qmap / qcount: optional device-side list of the queries to process (sorted positions when DENSE).
// need = radius (1 + 1e-3): the stencil variant walks 3x3x3 cells, which covers the radius only on a grid whose edge
// is at least that; on a finer grid it hands every query to the generic kernels (m-ring stencils).
// COOP (work-list passes): phase A scans each query's stencil with the whole warp (coalesced candidates, ballot
// compaction) instead of one lane per query - a short list of scattered queries is latency-bound, not issue-bound.
template <bool DENSE, bool ROWS, bool COOP>
__global__ void __launch_bounds__(FS_WPB * 32, ROWS ? 8 : 6)
shot_fused_kernel(GridDev g, const float4* __restrict__ queries, int nq, const float4* __restrict__ nrm, float r2,
                  double R, float* __restrict__ out, size_t stride, int* __restrict__ wl_count, int* __restrict__ wl,
                  const int* __restrict__ rows_idx, const float* __restrict__ rows_d2, int k,
                  const int* __restrict__ qmap, const int* __restrict__ qcount, float need,
                  int qpw /* queries per warp: 32, fewer for a short work list (more warps in flight) */,
                  int* __restrict__ ticket /* zeroed: batches of qpw queries are handed out dynamically */) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr int CAP = ROWS ? NCAP_ROWS : NCAP;
  FusedSmem<CAP>* S = reinterpret_cast<FusedSmem<CAP>*>(smem_raw) + wid;
  const GridParams P = *g.gp;
  const int n_valid = P.n_valid;
  const float nanv = __int_as_float(0x7fc00000);
  const float Rf = (float)R;
  const float t12 = shot_d2_threshold(R / 2);  // d2 > t12  <=>  sqrt((double)d2) > R / 2
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < 11; ++i) S->hist[lane + 32 * i] = 0;
  __syncwarp();
  const int limit = qmap ? min(*qcount, nq) : nq;
  const bool covers = ROWS || P.edge >= need;
  if (COOP) {
    // a work list is short and its queries are scattered: the pass is bound by the chain of dependent loads of a
    // query, not by issue slots, so the list is spread over every resident warp (at least four queries each)
    const int slots = gridDim.x * FS_WPB;
    qpw = max(4, min(qpw, (limit + slots - 1) / slots));
  }
  for (;;) {
    int qbase = 0;
    if (lane == 0) qbase = atomicAdd(ticket, 1) * qpw;
    qbase = __shfl_sync(FULL, qbase, 0);
    if (qbase >= limit) break;
    // ---------------- phase A (lane = query): one pass over the 3x3x3 stencil builds the neighbour list
    // in shared memory and the (R - d)-weighted scatter matrix of getLocalRF in double
    const bool have_q = lane < qpw && qbase + lane < limit;
    const int qi = have_q ? (qmap ? qmap[qbase + lane] : qbase + lane) : 0;
    const float4 q = have_q ? (DENSE ? g.pts[qi] : queries[qi]) : make_float4(0.f, 0.f, 0.f, 0.f);
    const bool q_ok = have_q && finite3(q.x, q.y, q.z) && (!DENSE || qi < n_valid) && n_valid > 0;
    double m6[6] = {0, 0, 0, 0, 0, 0};
    double msw = 0.0;
    int n_all = 0, n_val = 0;
    bool handoff = q_ok && !covers;  // not answered here: goes to the work list
    if (ROWS) {
      // the rows of the warp's 32 queries are one contiguous block: query t's row is read by the whole warp (lane =
      // slot), filtered by d2 < r2 and compacted into column t of the neighbour table
      const int qend_a = min(qpw, limit - qbase);
      if (k == 32) {
        // k = 32: one slot per lane, two coalesced 128-byte loads per query
        const int* ri = rows_idx + (size_t)qbase * 32 + lane;
        const float* rd = rows_d2 + (size_t)qbase * 32 + lane;
#pragma unroll SHOT_A_UNROLL
        for (int t = 0; t < qend_a; ++t) {
          const int j = ri[t * 32];
          const float d2 = rd[t * 32];
          const bool in = j >= 0 && d2 < r2;
          const unsigned m = __ballot_sync(FULL, in);
          if (in) S->nbr[__popc(m & lt)][t] = j;
          if (lane == t) {
            n_all = __popc(m);
            handoff = q_ok && m == FULL;  // no entry at or beyond the radius: the set is open
          }
        }
      } else {
        for (int t = 0; t < qend_a; ++t) {
          int cnt = 0;
          bool closed = false;
          for (int s0 = 0; s0 < k; s0 += 32) {
            const int s = s0 + lane;
            int j = -1;
            float d2 = CUDART_INF_F;
            if (s < k) {
              j = rows_idx[(size_t)(qbase + t) * k + s];
              d2 = rows_d2[(size_t)(qbase + t) * k + s];
            }
            const bool in = j >= 0 && d2 < r2;
            const unsigned m = __ballot_sync(FULL, in);
            if (in) S->nbr[cnt + __popc(m & lt)][t] = j;
            cnt += __popc(m);
            // an entry at or beyond the radius (or an unfilled slot: the cloud has fewer than k points) closes the set
            closed |= __ballot_sync(FULL, s < k && !in) != 0u;
          }
          if (lane == t) {
            n_all = cnt;
            handoff = q_ok && !closed;
          }
        }
      }
      __syncwarp();
    } else if (COOP) {
      const int qend_a = min(qpw, limit - qbase);
      for (int t = 0; t < qend_a; ++t) {
        if (!__shfl_sync(FULL, (int)(q_ok && covers), t)) continue;
        const int qi_t = __shfl_sync(FULL, qi, t);
        const float qx = __shfl_sync(FULL, q.x, t), qy = __shfl_sync(FULL, q.y, t), qz = __shfl_sync(FULL, q.z, t);
        const CellBlock blk = DENSE ? stencil_of_point(g, qi_t, lane) : stencil_of_pos(g, qx, qy, qz, lane);
        int cnt = 0;
        for (int base = 0; base < blk.total; base += 32) {  // same candidate order as the lane-serial walk below
          const int c = base + lane;
          const bool valid = c < blk.total;
          const int j = block_candidate(blk, valid ? c : 0);
          bool in = false;
          if (valid) {
            const float4 p = g.pts[j];
            in = dist2_flann(qx, qy, qz, p.x, p.y, p.z) < r2;
          }
          const unsigned m = __ballot_sync(FULL, in);
          const int pos = cnt + __popc(m & lt);
          if (in && pos < CAP) S->nbr[pos][t] = j;
          cnt += __popc(m);
        }
        if (lane == t) n_all = cnt;
      }
      __syncwarp();
    } else if (q_ok && covers) {
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
            if (n_all < CAP) S->nbr[n_all][lane] = j;
            ++n_all;
          }
        }
      }
    }
    {
      // scatter matrix from the compact list (uniform trip counts: no lanes idle on rejected candidates)
      const int n_list = (q_ok && !handoff) ? min(n_all, CAP) : 0;
#pragma unroll SHOT_BC_UNROLL
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
    const bool overflow = handoff || n_all > CAP;  // handed to the work list (ROWS: the stencil variant; else the generic kernels)
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
#pragma unroll SHOT_BC_UNROLL
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
    const int qend = min(qpw, limit - qbase);
    for (int t = 0; t < qend; ++t) {
      const int n_t = __shfl_sync(FULL, n_all, t);
      const bool ok_t = __shfl_sync(FULL, (int)q_ok, t);
      const bool of_t = __shfl_sync(FULL, (int)overflow, t);
      const bool gd = __shfl_sync(FULL, (int)good, t);
      const float4 qt = make_float4(__shfl_sync(FULL, q.x, t), __shfl_sync(FULL, q.y, t), __shfl_sync(FULL, q.z, t),
                                    __shfl_sync(FULL, q.w, t));
      const int qi_t = __shfl_sync(FULL, qi, t);
      const size_t row = DENSE ? (size_t)__float_as_int(qt.w) : (size_t)qi_t;
      float* o = out + row * stride;
      if (of_t) {
        if (lane == 0) wl[atomicAdd(wl_count, 1)] = qi_t;
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
// generic kernels (shot.cu) over a device-side work list.  g == nullptr: the caller has no grid for the radius yet -
// the dense k-search rows serve when they are resident (shot_rows_available), else the radius grid is fetched here.
bool shot_rows_available(Ctx* ctx, double radius) {
  if (!(ctx->shot_from_rows && ctx->q_is_surface && ctx->knn_dense && ctx->knn_grid &&
        ctx->knn_grid->surf_version == ctx->surf_version && ctx->knn_sversion == ctx->surf_version &&
        ctx->knn_k >= 8 && ctx->knn_k <= 32 && ctx->n >= (size_t)ctx->sm_count * 512))
    return false;
  // The rows pay off while most of them close (the k-th neighbour lies beyond the radius).  What the previous call with
  // this (radius, k) handed to the stencil pass is read back asynchronously and never waited for: when more than a
  // quarter of the rows stayed open, the radius is too large for this k and a radius grid serves better.
  Ctx::RowsStat& st = ctx->rows_stat;
  if (st.host && st.radius == radius && st.k == ctx->knn_k && st.n > 0) {
    if (st.pending && cudaEventQuery(st.ev) == cudaSuccess) {
      st.open_frac = (double)st.host[0] / (double)st.n;
      st.pending = false;
    } else if (st.pending) {
      (void)cudaGetLastError();  // cudaErrorNotReady is not an error
    }
    if (st.open_frac > 0.25) return false;
  }
  return true;
}

int shot_fused_compute(Ctx* ctx, Grid* g, double radius, float* out_dev, size_t stride_floats) {
  const int nq = (int)ctx->num_queries();
  if (nq == 0) return 0;
  const float r2 = (float)(radius * radius);
  const float need = (float)(radius * (1.0 + 1e-3));
  const bool rows = (g == nullptr) && shot_rows_available(ctx, radius);
  if (rows) g = const_cast<Grid*>(ctx->knn_grid);
  if (!g) PFX_TRY(grid_for_radius(ctx, radius, &g));
  const float4* nrm = nullptr;
  PFX_TRY(normals_sorted_for_grid(ctx, g, &nrm));
  if (!ctx->q_is_surface && nq < ctx->sm_count * 512) {
    // keypoint queries (the reference's use: evaluation.cpp:770-775 at r = 5 cm, hundreds of neighbours each):
    // too few queries for the lane-per-query kernel; the generic kernels give every query its own warp
    PFX_CUDA(ctx->tmp2.ensure((size_t)nq * 9 * sizeof(float)));
    PFX_TRY(shot_lrf_compute(ctx, g, radius, ctx->tmp2.as<float>(), nullptr));
    return shot_compute(ctx, g, radius, ctx->tmp2.as<float>(), out_dev, stride_floats);
  }
  // two device-side work lists: A = rows the k-search could not close (ROWS only), B = neighbourhoods beyond NCAP
  PFX_CUDA(ctx->worklist2.ensure(2 * ((size_t)nq + 16) * sizeof(int)));
  int* wlA_count = ctx->worklist2.as<int>();
  int* wlB_count = wlA_count + 8;
  int* wlA = wlA_count + 16;
  int* wlB = wlA + nq + 16;
  PFX_CUDA(cudaMemsetAsync(wlA_count, 0, 16 * sizeof(int), ctx->stream));
  const size_t smem = sizeof(FusedSmem<NCAP>) * FS_WPB, smem_rows = sizeof(FusedSmem<NCAP_ROWS>) * FS_WPB;
  if (!ctx->smem_attr_shot_fused) {
    PFX_CUDA(cudaFuncSetAttribute(shot_fused_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PFX_CUDA(cudaFuncSetAttribute(shot_fused_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PFX_CUDA(cudaFuncSetAttribute(shot_fused_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
    PFX_CUDA(cudaFuncSetAttribute(shot_fused_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctx->smem_attr_shot_fused = true;
  }
  if (!ctx->shot_fused_blocks_per_sm[0]) {
    int b = 0;
    PFX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, shot_fused_kernel<true, false, false>, FS_WPB * 32, smem));
    ctx->shot_fused_blocks_per_sm[0] = std::max(1, b);
    PFX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, shot_fused_kernel<true, true, false>, FS_WPB * 32, smem_rows));
    ctx->shot_fused_blocks_per_sm[1] = std::max(1, b);
  }
  // persistent over a ticket: one wave of resident blocks
  const int blocks = std::min(div_up(nq, FS_WPB * 32), ctx->sm_count * ctx->shot_fused_blocks_per_sm[rows ? 1 : 0]);
  int* ticket0 = wlA_count + 4;
  int* ticket1 = wlA_count + 12;
  if (rows) {
    PFX_LAUNCH(ctx, (shot_fused_kernel<true, true, false>), blocks, FS_WPB * 32, smem_rows, g->view(), nullptr, nq, nrm, r2, radius,
               out_dev, stride_floats, wlA_count, wlA, ctx->knn_idx.as<int>(), ctx->knn_d2.as<float>(), ctx->knn_k,
               nullptr, nullptr, need, 32, ticket0);
    // the rows the k-search could not close: stencil walk on the k-search grid (a persistent launch over list A,
    // 16 queries per warp: the list is short, more warps in flight hide its latency)
    PFX_LAUNCH(ctx, (shot_fused_kernel<true, false, true>), ctx->sm_count * 8, FS_WPB * 32, smem, g->view(), nullptr, nq, nrm, r2,
               radius, out_dev, stride_floats, wlB_count, wlB, nullptr, nullptr, 0, wlA, wlA_count, need, 16, ticket1);
    Ctx::RowsStat& st = ctx->rows_stat;
    if (!st.host) {
      PFX_CUDA(cudaMallocHost(reinterpret_cast<void**>(&st.host), 16 * sizeof(int)));
      PFX_CUDA(cudaEventCreateWithFlags(&st.ev, cudaEventDisableTiming));
    }
    if (!(st.radius == radius && st.k == ctx->knn_k)) st.open_frac = 0.0;  // another configuration: no history yet
    if (!st.pending) {  // (a read-back still in flight keeps its slot)
      PFX_CUDA(cudaMemcpyAsync(st.host, wlA_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      PFX_CUDA(cudaEventRecord(st.ev, ctx->stream));
      st.pending = true;
      st.n = (size_t)nq;
    }
    st.radius = radius;
    st.k = ctx->knn_k;
  } else if (ctx->q_is_surface) {
    PFX_LAUNCH(ctx, (shot_fused_kernel<true, false, false>), blocks, FS_WPB * 32, smem, g->view(), nullptr, nq, nrm, r2, radius,
               out_dev, stride_floats, wlB_count, wlB, nullptr, nullptr, 0, nullptr, nullptr, need, 32, ticket0);
  } else {
    PFX_LAUNCH(ctx, (shot_fused_kernel<false, false, false>), blocks, FS_WPB * 32, smem, g->view(), ctx->qry.as<float4>(), nq, nrm,
               r2, radius, out_dev, stride_floats, wlB_count, wlB, nullptr, nullptr, 0, nullptr, nullptr, need, 32, ticket0);
  }
  PFX_CUDA(cudaGetLastError());
  // neighbourhoods beyond NCAP (or every query of list A when the grid's cells are finer than the radius): frames then
  // descriptors with the generic kernels, work-list driven
  PFX_CUDA(ctx->tmp2.ensure((size_t)nq * 9 * sizeof(float)));
  PFX_TRY(shot_lrf_worklist(ctx, g, radius, ctx->tmp2.as<float>(), wlB, wlB_count));
  PFX_TRY(shot_worklist(ctx, g, radius, ctx->tmp2.as<float>(), out_dev, stride_floats, wlB, wlB_count));
  return 0;
}

}  // namespace pfx
