// knn_tile.cu — dense kNN neighbour SETS (+ fused normals) on cell tiles.
//
// One warp per occupied cell: the <= TCAP candidate points of the 3x3x3 stencil are staged in shared
// memory once; queries of the cell are processed four at a time, eight lanes per query.  A lane keeps
// the squared distances of its share of the candidates in registers; the k nearest are found by
// SELECTION, not sorting: a bracketing search on the distance threshold tau (interpolation while the
// bracket is wide - count(d2 <= tau) is nearly linear in tau on a surface - then bisection) until
// exactly k candidates satisfy d2 <= tau.  Each step is one compare-and-count pass over registers.
// Rows are written as global sorted positions + d2 (same layout as the generic kernel, search.cu),
// unsorted inside a row.  With NORMALS the covariance moments of the selected neighbours are
// accumulated in the same pass (NormalEstimation with the same k: tools.h:26-31 / features.h:187) and
// solved 32 queries at a time (normals_solve.cuh).
//
// Anything the tile path cannot certify - a stencil with more than TCAP points, fewer than k
// candidates, a k-th distance that reaches beyond the 3x3x3 block, exact distance ties across the
// k-th rank (the set then depends on the index tie-break) - is flagged per query and redone by the
// generic ring-expanding kernel, which implements the full ascending (d2, index) rule.
#include "internal.h"
#include "normals_solve.cuh"
#include "tile.cuh"

namespace pfx {

constexpr int TCAP = 256;
constexpr int TR = TCAP / 8;
constexpr int TWPB = 8;

constexpr int XCAP = 16;  // bracket candidates ranked directly

struct KnnTileSmem {
  float4 pts[TCAP];
  int gidx[TCAP];
  float xs[4][XCAP];
  TileTab tab;
};

__device__ __forceinline__ int group_sum8(int v) {
  v += __shfl_xor_sync(FULL, v, 1);
  v += __shfl_xor_sync(FULL, v, 2);
  v += __shfl_xor_sync(FULL, v, 4);
  return v;
}
__device__ __forceinline__ double group_sum8(double v) {
  v += __shfl_xor_sync(FULL, v, 1);
  v += __shfl_xor_sync(FULL, v, 2);
  v += __shfl_xor_sync(FULL, v, 4);
  return v;
}
__device__ __forceinline__ int group_excl_scan8(int v, int sl) {
  int inc = v;
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    int t = __shfl_up_sync(FULL, inc, o, 8);
    if (sl >= o) inc += t;
  }
  return inc - v;
}

template <bool NORMALS>
__global__ void __launch_bounds__(TWPB * 32)
knn_tile_kernel(GridDev g, int k, int* __restrict__ out_idx, float* __restrict__ out_d2,
                unsigned char* __restrict__ qflag, int* __restrict__ wl_count, int* __restrict__ wl,
                double* __restrict__ mom_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int sub = lane >> 3, sl = lane & 7;
  KnnTileSmem* S = reinterpret_cast<KnnTileSmem*>(smem_raw) + wid;
  const GridParams P = *g.gp;
  const float INF = CUDART_INF_F;
  // non-finite points sit after n_valid in sorted order and belong to no cell: the generic kernels
  // give them their rows
  if (blockIdx.x == 0)
    for (int i = P.n_valid + threadIdx.x; i < g.n; i += blockDim.x) {
      qflag[i] = 1;
      wl[atomicAdd(wl_count, 1)] = i;
    }
  // cells are handed out by a ticket (wl_count[8], zeroed with the work-list counter): stencils differ by a factor of
  // two in size, and a static split leaves the last blocks running alone
  for (;;) {
    int cell = 0;
    if (lane == 0) cell = atomicAdd(wl_count + 8, 1);
    cell = __shfl_sync(FULL, cell, 0);
    if (cell >= P.ncells) break;
    __syncwarp();
    const int M = tile_setup(g, cell, lane, &S->tab);
    const int q0 = S->tab.start[13];
    const int nqc = S->tab.prefix[14] - S->tab.prefix[13];
    const int own = S->tab.prefix[13];
    if (M > TCAP || M < k) {  // generic kernel: bigger stencil / ring expansion
      for (int t = lane; t < nqc; t += 32) {
        qflag[q0 + t] = 1;
        wl[atomicAdd(wl_count, 1)] = q0 + t;
      }
      continue;
    }
    for (int t = lane; t < M; t += 32) {
      int j = tile_global_index(&S->tab, t);
      S->gidx[t] = j;
      S->pts[t] = g.pts[j];
    }
    // the tile is padded to a multiple of 32 candidates with points at infinity: their distances come out infinite
    // by themselves, and the distance loop needs neither an index clamp nor a select per candidate
    if (M + lane < ((M + 31) & ~31)) S->pts[M + lane] = make_float4(INF, 0.f, 0.f, 0.f);
    __syncwarp();
    int cx, cy, cz;
    {
      float4 p0 = S->pts[own];
      cx = cell_coord(p0.x, P.ox, P.inv_e, P.nx);
      cy = cell_coord(p0.y, P.oy, P.inv_e, P.ny);
      cz = cell_coord(p0.z, P.oz, P.inv_e, P.nz);
    }
    const float tau0 = 2.865f * P.edge * P.edge * (float)k / (float)M;
    {
      const int cend = nqc;
      for (int qb = 0; qb < cend; qb += 4) {
        const bool active = (qb + sub) < cend;
        const float4 q = S->pts[own + min(qb + sub, nqc - 1)];
        float d2[TR];
#pragma unroll
        for (int r = 0; r < TR; ++r) {
          if ((r & 3) == 0 && r * 8 >= M) break;
          const float4 p = S->pts[r * 8 + sl];
          d2[r] = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);
        }
        // ---- find tau with count(d2 <= tau) == k.  The first sweep counts at two thresholds around the density
        // estimate (0.8 / 1.25 tau0), which usually brackets rank k at once; interpolation narrows a wide bracket
        // (the count is nearly linear in tau on a surface); once at most XCAP candidates lie between the bracket ends
        // the (k - clo)-th smallest of them is picked by rank counting.
        float lo = -1.f, hi = INF, tau = tau0;
        int clo = 0, chi = M;
        bool done = !active, fail = false;
        {
          const float ta = 0.8f * tau0, tb = 1.25f * tau0;
          int ca = 0, cb = 0;
#pragma unroll
          for (int r = 0; r < TR; ++r) {
            if ((r & 3) == 0 && r * 8 >= M) break;
            ca += (d2[r] <= ta) ? 1 : 0;
            cb += (d2[r] <= tb) ? 1 : 0;
          }
          ca = group_sum8(ca);
          cb = group_sum8(cb);
          if (cb < k) { lo = tb; clo = cb; }
          else if (ca >= k) { hi = ta; chi = ca; }
          else { lo = ta; clo = ca; hi = tb; chi = cb; }
          if (!done && ca == k) { tau = ta; done = true; }
          if (!done && cb == k) { tau = tb; done = true; }
        }
        // The four queries of a pass narrow their brackets TOGETHER (a compare-and-count pass over registers per step,
        // idle for the queries that are already there) until every one of them is exact or holds at most XCAP
        // candidates between its bracket ends; the rank extraction then runs ONCE for the whole warp.  (Extracting as
        // soon as any query was ready ran the extraction block two to four times per pass: the queries of a warp reach
        // their brackets at different steps.)
        for (int it = 0; it < 40; ++it) {
          const bool ready = done || fail || (hi != INF && (chi - clo) <= XCAP);
          if (__all_sync(FULL, ready)) break;
          const bool want_c = !ready;
          if (want_c) {
            float t;
            if (hi == INF) {
              t = lo * fmaxf(1.25f, ((float)k + 1.f) / ((float)clo + 0.5f));
            } else if (lo < 0.f) {
              t = hi * ((float)k / ((float)chi + 0.5f));
            } else if ((it & 3) != 3) {
              t = lo + (hi - lo) * (((float)(k - clo) + 0.5f) / (float)(chi - clo + 1));
            } else {
              t = 0.5f * lo + 0.5f * hi;
            }
            const float lo_next = (lo < 0.f) ? 0.f : __uint_as_float(__float_as_uint(lo) + 1u);
            if (!(t > lo)) t = lo_next;
            if (!(t < hi)) t = __uint_as_float(__float_as_uint(hi) - 1u);
            if (!(t > lo) || !(t < hi)) fail = true;  // adjacent floats around many equal distances
            tau = t;
          }
          int c = 0;
#pragma unroll
          for (int r = 0; r < TR; ++r) {
            if ((r & 3) == 0 && r * 8 >= M) break;
            c += (d2[r] <= tau) ? 1 : 0;
          }
          c = group_sum8(c);
          if (want_c && !fail) {
            if (c == k) done = true;
            else if (c < k) { lo = tau; clo = c; }
            else { hi = tau; chi = c; }
          }
        }
        {
          // ---- rank extraction: the candidates between the bracket ends go to shared memory and the (k - clo)-th
          // smallest of them is picked by rank counting - one step, however close the k-th and (k+1)-th distances are
          const bool want_x = !(done || fail) && hi != INF && (chi - clo) <= XCAP;
          if (__any_sync(FULL, want_x)) {
            unsigned mask = 0;
#pragma unroll
            for (int r = 0; r < TR; ++r) {
              if ((r & 3) == 0 && r * 8 >= M) break;
              if (d2[r] > lo && d2[r] <= hi) mask |= 1u << r;
            }
            if (!want_x) mask = 0;
            int pos = group_excl_scan8(__popc(mask), sl);
            float* xs = S->xs[sub];
            while (mask) {
              const int c = (__ffs(mask) - 1) * 8 + sl;
              mask &= mask - 1;
              const float4 p = S->pts[c];
              xs[pos++] = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);  // the same value as the d2 register
            }
            __syncwarp();
            const int nb = want_x ? chi - clo : 0, need = k - clo;
            const float v0 = (sl < nb) ? xs[sl] : INF, v1 = (sl + 8 < nb) ? xs[sl + 8] : INF;
            // strict ranks only: the value of rank need-1 (0-based) is the LARGEST value with fewer than `need` values
            // below it; whether a run of equal values straddles rank k (the set then depends on the index tie-break,
            // which only the generic kernel implements) is told by counting the copies of that one value afterwards
            int r0 = 0, r1 = 0;
            const int nbmax = __reduce_max_sync(FULL, nb);
            for (int j = 0; j < nbmax; ++j) {
              const float x = (j < nb) ? xs[j] : INF;
              r0 += (x < v0) ? 1 : 0;
              r1 += (x < v1) ? 1 : 0;
            }
            const float NINF = -CUDART_INF_F;
            float cand = NINF;
            int rc = 0;
            if (v0 != INF && r0 < need) { cand = v0; rc = r0; }
            if (v1 != INF && r1 < need && v1 >= cand) { cand = v1; rc = r1; }   // (v1 >= v0 is not implied: slots are unordered)
            float gmax = fmaxf(cand, __shfl_xor_sync(FULL, cand, 1));
            gmax = fmaxf(gmax, __shfl_xor_sync(FULL, gmax, 2));
            gmax = fmaxf(gmax, __shfl_xor_sync(FULL, gmax, 4));
            const int copies = group_sum8(((v0 == gmax) ? 1 : 0) + ((v1 == gmax) ? 1 : 0));
            // the lanes that hold the value know its strict rank
            int bad = (cand == gmax && gmax != NINF && rc + copies != need) ? 1 : 0;
            bad = group_sum8(bad);
            cand = (gmax == NINF) ? INF : gmax;
            if (want_x) {
              if (cand == INF || bad) fail = true;
              else { tau = cand; done = true; }
            }
            __syncwarp();
          }
        }
        if (!done) fail = true;
        // ---- certificate: the k-th distance must lie inside the scanned 3x3x3 block
        if (active && !fail) {
          const float ux = __fmul_rn(__fsub_rn(q.x, P.ox), P.inv_e), uy = __fmul_rn(__fsub_rn(q.y, P.oy), P.inv_e),
                      uz = __fmul_rn(__fsub_rn(q.z, P.oz), P.inv_e);
          float safe = INF;
          if (cx - 1 > 0) safe = fminf(safe, ux - (float)(cx - 1));
          if (cx + 1 < P.nx - 1) safe = fminf(safe, (float)(cx + 2) - ux);
          if (cy - 1 > 0) safe = fminf(safe, uy - (float)(cy - 1));
          if (cy + 1 < P.ny - 1) safe = fminf(safe, (float)(cy + 2) - uy);
          if (cz - 1 > 0) safe = fminf(safe, uz - (float)(cz - 1));
          if (cz + 1 < P.nz - 1) safe = fminf(safe, (float)(cz + 2) - uz);
          if (safe != INF) {
            safe = (safe - 1e-3f) * P.edge;
            if (!(safe > 0.f && tau < safe * safe)) fail = true;  // tau >= k-th distance
          }
        }
        // ---- emit the set (+ moments)
        const bool good = active && !fail;
        unsigned selmask = 0;  // bit r: my r-th candidate belongs to the set
#pragma unroll
        for (int r = 0; r < TR; ++r) {
          if ((r & 3) == 0 && r * 8 >= M) break;
          if (d2[r] <= tau) selmask |= 1u << r;
        }
        if (!good) selmask = 0;
        int pos = group_excl_scan8(__popc(selmask), sl);
        const int qi = q0 + qb + sub;
        double m9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        while (selmask) {
          const int c = (__ffs(selmask) - 1) * 8 + sl;
          selmask &= selmask - 1;
          const float4 p = S->pts[c];
          out_idx[(size_t)qi * k + pos] = S->gidx[c];
          out_d2[(size_t)qi * k + pos] = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);
          ++pos;
          if (NORMALS) {
            double dx = (double)p.x - (double)q.x, dy = (double)p.y - (double)q.y, dz = (double)p.z - (double)q.z;
            m9[0] += dx; m9[1] += dy; m9[2] += dz;
            m9[3] += dx * dx; m9[4] += dx * dy; m9[5] += dx * dz;
            m9[6] += dy * dy; m9[7] += dy * dz; m9[8] += dz * dz;
          }
        }
        if (active && !good && sl == 0) {
          qflag[qi] = 1;
          wl[atomicAdd(wl_count, 1)] = qi;
        }
        if (NORMALS) {
          // the 3x3 eigen problems are solved by normals_from_moments_kernel at full SIMT width
          // (a cell holds ~10 queries: solving here would leave 2/3 of the lanes idle)
#pragma unroll
          for (int i = 0; i < 9; ++i) m9[i] = group_sum8(m9[i]);
          if (good && sl == 0) {
            // nine planes of n doubles: the four queries of a pass are consecutive, so each store is one sector,
            // and normals_from_moments_kernel reads every plane coalesced
            double* dst = mom_out + qi;
#pragma unroll
            for (int i = 0; i < 9; ++i) dst[(size_t)i * g.n] = m9[i];
          }
        }
      }
    }
  }
}

// one thread per query: moments -> normal (rows flagged for the generic path are skipped)
__global__ void normals_from_moments_kernel(GridDev g, const double* __restrict__ mom, int k,
                                            const unsigned char* __restrict__ qflag, float vx, float vy, float vz,
                                            float4* __restrict__ nrm_sorted, float4* __restrict__ nrm_orig) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.gp->n_valid || qflag[i]) return;
  double s[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) s[t] = mom[(size_t)t * g.n + i];
  const float4 q = g.pts[i];
  const float4 r = solve_normal_m9(s, k, q.x, q.y, q.z, vx, vy, vz);
  nrm_sorted[i] = r;
  nrm_orig[__float_as_int(q.w)] = r;
}

// Dense kNN sets (+ normals) of the surface: fills ctx->knn_idx / knn_d2 (rows in sorted query
// order, global sorted positions, UNSORTED inside a row except for the rows redone generically) and
// ctx->qflag (which queries were redone).
int knn_tile_lists(Ctx* ctx, Grid* g, int k, bool with_normals) {
  const int n = (int)ctx->n;
  if (ctx->knn_grid == g && ctx->knn_k == k && ctx->knn_sversion == ctx->surf_version && ctx->knn_dense &&
      (!with_normals || ctx->tile_has_normals))
    return 0;
  PFX_CUDA(ctx->qflag.ensure((size_t)std::max(n, 1)));
  PFX_CUDA(ctx->knn_idx.ensure((size_t)std::max(n, 1) * k * sizeof(int)));
  PFX_CUDA(ctx->knn_d2.ensure((size_t)std::max(n, 1) * k * sizeof(float)));
  if (with_normals) {
    PFX_CUDA(ctx->normals.ensure(std::max<size_t>(n, 1) * sizeof(float4)));
    PFX_CUDA(ctx->normals_sorted.ensure(std::max<size_t>(n, 1) * sizeof(float4)));
  }
  if (n == 0) return 0;
  PFX_CUDA(ctx->worklist.ensure(((size_t)n + 16) * sizeof(int)));
  PFX_CUDA(cudaMemsetAsync(ctx->qflag.p, 0, (size_t)n, ctx->stream));
  PFX_CUDA(cudaMemsetAsync(ctx->worklist.p, 0, 16 * sizeof(int), ctx->stream));
  int* wl_count = ctx->worklist.as<int>();
  int* wl = wl_count + 16;
  const size_t smem = sizeof(KnnTileSmem) * TWPB;
  if (!ctx->smem_attr_knn_tile) {
    PFX_CUDA(cudaFuncSetAttribute(knn_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PFX_CUDA(cudaFuncSetAttribute(knn_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctx->smem_attr_knn_tile = true;
  }
  // one wave of resident blocks (the kernel is persistent over a cell ticket)
  if (!ctx->knn_tile_blocks_per_sm) {
    int b0 = 0, b1 = 0;
    PFX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b0, knn_tile_kernel<true>, TWPB * 32, smem));
    PFX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, knn_tile_kernel<false>, TWPB * 32, smem));
    ctx->knn_tile_blocks_per_sm = std::max(1, std::min(b0, b1));
  }
  const int blocks = ctx->sm_count * ctx->knn_tile_blocks_per_sm;
  if (with_normals) {
    PFX_CUDA(ctx->tmp4.ensure((size_t)n * 9 * sizeof(double)));
    PFX_LAUNCH(ctx, knn_tile_kernel<true>, blocks, TWPB * 32, smem, g->view(), k, ctx->knn_idx.as<int>(),
               ctx->knn_d2.as<float>(), ctx->qflag.as<unsigned char>(), wl_count, wl, ctx->tmp4.as<double>());
    PFX_LAUNCH(ctx, normals_from_moments_kernel, div_up(n, 128), 128, 0, g->view(), ctx->tmp4.as<double>(), k,
               ctx->qflag.as<unsigned char>(), ctx->vp[0], ctx->vp[1], ctx->vp[2], ctx->normals_sorted.as<float4>(),
               ctx->normals.as<float4>());
  } else {
    PFX_LAUNCH(ctx, knn_tile_kernel<false>, blocks, TWPB * 32, smem, g->view(), k, ctx->knn_idx.as<int>(),
               ctx->knn_d2.as<float>(), ctx->qflag.as<unsigned char>(), wl_count, wl, nullptr);
  }
  PFX_CUDA(cudaGetLastError());
  // generic completion of the queries handed back (persistent kernel over the device-side work list)
  PFX_TRY(knn_run_worklist(ctx, g, k, ctx->knn_idx.as<int>(), ctx->knn_d2.as<float>(), wl, wl_count));
  ctx->knn_grid = g;
  ctx->knn_k = k;
  ctx->knn_sversion = ctx->surf_version;
  ctx->knn_qversion = ctx->qry_version;
  ctx->knn_dense = true;
  ctx->knn_sorted = false;
  ctx->tile_has_normals = with_normals;
  return 0;
}

}  // namespace pfx
