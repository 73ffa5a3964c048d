// search.cu — exact kNN and radius search over the voxel hash (replaces
// pcl::KdTreeFLANN::nearestKSearch / radiusSearch; reference call sites features.h:192-193,
// tools.h:29-30, keypoints.h:371-386, 408-417).
//
// One warp per query.  Candidates come from the 3x3x3 cell stencil (adjacency table for surface
// queries, hash probes otherwise), flattened over the warp so that lanes stay busy even when cells
// hold few points.  kNN keeps the running k best as a warp-distributed sorted list of 64-bit keys
// (d2 bits << 32 | original index), merged with bitonic networks; rings of cells are added until
// the k-th distance is provably inside the scanned block.  Tie-break: ascending (d2, index).
#include "internal.h"

namespace pfx {

constexpr int WPB = 8;  // warps per block
constexpr unsigned long long KMAX = 0xffffffffffffffffull;

__device__ __forceinline__ unsigned long long umin64(unsigned long long a, unsigned long long b) {
  return a < b ? a : b;
}
__device__ __forceinline__ unsigned long long umax64(unsigned long long a, unsigned long long b) {
  return a < b ? b : a;
}

__device__ __forceinline__ unsigned long long bitonic_sort32(unsigned long long key, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      unsigned long long other = __shfl_xor_sync(FULL, key, j);
      bool up = ((lane & k) == 0);
      bool lower = ((lane & j) == 0);
      key = (lower == up) ? umin64(key, other) : umax64(key, other);
    }
  }
  return key;
}

__device__ __forceinline__ unsigned long long bitonic_merge32(unsigned long long key, int lane) {
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    unsigned long long other = __shfl_xor_sync(FULL, key, j);
    key = ((lane & j) == 0) ? umin64(key, other) : umax64(key, other);
  }
  return key;
}

struct KnnState {
  unsigned long long best;  // lane i holds the i-th smallest key so far
  unsigned long long thr;   // key of rank k-1 (warp-uniform)
  int nb;                   // keys waiting in the shared buffer (warp-uniform)
};

__device__ __forceinline__ void knn_flush(KnnState& s, unsigned long long* buf, int lane, int k) {
  unsigned long long b = (lane < s.nb) ? buf[lane] : KMAX;
  b = bitonic_sort32(b, lane);
  unsigned long long rev = __shfl_sync(FULL, b, 31 - lane);
  s.best = bitonic_merge32(umin64(s.best, rev), lane);
  s.thr = __shfl_sync(FULL, s.best, k - 1);
  s.nb = 0;
  __syncwarp();
}

__device__ __forceinline__ void knn_push(KnnState& s, unsigned long long* buf, int lane, int k,
                                         unsigned long long key) {
  const unsigned lt = (1u << lane) - 1u;
  bool acc = key < s.thr;
  unsigned m = __ballot_sync(FULL, acc);
  if (m == 0) return;
  int cnt = __popc(m);
  int pos = s.nb + __popc(m & lt);
  if (acc && pos < 32) buf[pos] = key;
  __syncwarp();
  if (s.nb + cnt >= 32) {
    s.nb = 32;
    knn_flush(s, buf, lane, k);
    bool left = acc && pos >= 32 && key < s.thr;
    unsigned m2 = __ballot_sync(FULL, left);
    if (left) buf[__popc(m2 & lt)] = key;
    s.nb = __popc(m2);
    __syncwarp();
  } else {
    s.nb += cnt;
  }
}

__device__ __forceinline__ void knn_scan_block(const GridDev& g, const CellBlock& blk, float qx,
                                               float qy, float qz, KnnState& s,
                                               unsigned long long* buf, int lane, int k) {
  for (int base = 0; base < blk.total; base += 32) {
    int t = base + lane;
    bool valid = t < blk.total;
    int j = block_candidate(blk, valid ? t : 0);
    unsigned long long key = KMAX;
    if (valid) {
      float4 p = g.pts[j];
      float d2 = dist2_flann(qx, qy, qz, p.x, p.y, p.z);
      key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)__float_as_int(p.w);
    }
    knn_push(s, buf, lane, k, key);
  }
}

// rows: DENSE -> row i = sorted surface point i; else row q = query q (caller order).
// out_idx holds SORTED positions of the neighbours (-1 padding), out_d2 the squared distances.
template <bool DENSE>
__device__ __forceinline__ void knn_one_query(const GridDev& g, const float4* __restrict__ queries, int qi, int k,
                                              int* __restrict__ out_idx, float* __restrict__ out_d2,
                                              unsigned long long* buf, int lane) {
  const GridParams P = *g.gp;
  float4 q = DENSE ? g.pts[qi] : queries[qi];
  KnnState s;
  s.best = KMAX;
  s.thr = KMAX;
  s.nb = 0;
  bool ok = finite3(q.x, q.y, q.z) && (!DENSE || qi < P.n_valid) && P.n_valid > 0;
  if (ok) {
    const int cx = cell_coord(q.x, P.ox, P.inv_e, P.nx), cy = cell_coord(q.y, P.oy, P.inv_e, P.ny),
              cz = cell_coord(q.z, P.oz, P.inv_e, P.nz);
    const float ux = __fmul_rn(__fsub_rn(q.x, P.ox), P.inv_e), uy = __fmul_rn(__fsub_rn(q.y, P.oy), P.inv_e),
                uz = __fmul_rn(__fsub_rn(q.z, P.oz), P.inv_e);
    CellBlock blk = DENSE ? stencil_of_point(g, qi, lane) : stencil_of_pos(g, q.x, q.y, q.z, lane);
    knn_scan_block(g, blk, q.x, q.y, q.z, s, buf, lane, k);
    for (int R = 1;; ++R) {
      if (s.nb) knn_flush(s, buf, lane, k);
      // distance (in cells) from q to the nearest face of the scanned cube that has cells beyond it
      float safe = CUDART_INF_F;
      if (cx - R > 0) safe = fminf(safe, ux - (float)(cx - R));
      if (cx + R < P.nx - 1) safe = fminf(safe, (float)(cx + R + 1) - ux);
      if (cy - R > 0) safe = fminf(safe, uy - (float)(cy - R));
      if (cy + R < P.ny - 1) safe = fminf(safe, (float)(cy + R + 1) - uy);
      if (cz - R > 0) safe = fminf(safe, uz - (float)(cz - R));
      if (cz + R < P.nz - 1) safe = fminf(safe, (float)(cz + R + 1) - uz);
      if (safe == CUDART_INF_F) break;  // whole grid scanned
      safe = (safe - 1e-3f) * P.edge;
      if (s.thr != KMAX && safe > 0.f) {
        float kd2 = __uint_as_float((unsigned)(s.thr >> 32));
        if (kd2 < safe * safe) break;
      }
      const int R2 = R + 1;
      if (R2 > 6) {  // pathological density: exact scan of everything not yet visited
        for (int base = 0; base < P.n_valid; base += 32) {
          int j = base + lane;
          unsigned long long key = KMAX;
          if (j < P.n_valid) {
            float4 p = g.pts[j];
            int px = cell_coord(p.x, P.ox, P.inv_e, P.nx), py = cell_coord(p.y, P.oy, P.inv_e, P.ny),
                pz = cell_coord(p.z, P.oz, P.inv_e, P.nz);
            if (max(abs(px - cx), max(abs(py - cy), abs(pz - cz))) > R) {
              float d2 = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);
              key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)__float_as_int(p.w);
            }
          }
          knn_push(s, buf, lane, k, key);
        }
        if (s.nb) knn_flush(s, buf, lane, k);
        break;
      }
      // shell at Chebyshev distance R2, clipped to the grid, 32 cells at a time
      const int x0 = max(cx - R2, 0), x1 = min(cx + R2, P.nx - 1);
      const int y0 = max(cy - R2, 0), y1 = min(cy + R2, P.ny - 1);
      const int z0 = max(cz - R2, 0), z1 = min(cz + R2, P.nz - 1);
      const int wx = x1 - x0 + 1, wy = y1 - y0 + 1, wz = z1 - z0 + 1;
      const int ncube = wx * wy * wz;
      for (int cb = 0; cb < ncube; cb += 32) {
        int t = cb + lane;
        int c = -1;
        if (t < ncube) {
          int x = x0 + t % wx, y = y0 + (t / wx) % wy, z = z0 + t / (wx * wy);
          if (max(abs(x - cx), max(abs(y - cy), abs(z - cz))) == R2) c = hash_lookup(g, morton3(x, y, z));
        }
        CellBlock sb = make_block(g, c, lane);
        if (sb.total) knn_scan_block(g, sb, q.x, q.y, q.z, s, buf, lane, k);
      }
    }
  }
  if (lane < k) {
    unsigned long long key = s.best;
    int j = -1;
    float d2 = CUDART_INF_F;
    if (key != KMAX) {
      j = g.inv_perm[(int)(unsigned)(key & 0xffffffffull)];
      d2 = __uint_as_float((unsigned)(key >> 32));
    }
    out_idx[(size_t)qi * k + lane] = j;
    out_d2[(size_t)qi * k + lane] = d2;
  }
}

template <bool DENSE>
__global__ void __launch_bounds__(WPB * 32)
knn_kernel(GridDev g, const float4* __restrict__ queries, int nq, int k, int* __restrict__ out_idx,
           float* __restrict__ out_d2) {
  __shared__ unsigned long long sbuf[WPB][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int qi = blockIdx.x * WPB + wid;
  if (qi >= nq) return;
  knn_one_query<DENSE>(g, queries, qi, k, out_idx, out_d2, sbuf[wid], lane);
}

// persistent variant over a device-side work list (the queries the cell-tile path handed back)
__global__ void __launch_bounds__(WPB * 32)
knn_worklist_kernel(GridDev g, const int* __restrict__ worklist, const int* __restrict__ wl_count, int k,
                    int* __restrict__ out_idx, float* __restrict__ out_d2) {
  __shared__ unsigned long long sbuf[WPB][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int count = *wl_count;
  for (int w = blockIdx.x * WPB + wid; w < count; w += gridDim.x * WPB) {
    knn_one_query<true>(g, nullptr, worklist[w], k, out_idx, out_d2, sbuf[wid], lane);
    __syncwarp();
  }
}

// kNN of `nq` rows: q_dev == nullptr -> rows are the sorted surface points (dense), else float4
// positions.  idx receives SORTED positions (-1 padding).
int knn_run(Ctx* ctx, Grid* g, const float4* q_dev, int nq, int k, int* idx_dev, float* d2_dev) {
  if (k < 1 || k > 32) return ctx->fail(PFX_E_INVALID, "k must be in [1, 32]");
  if (nq <= 0) return 0;
  if (!q_dev)
    PFX_LAUNCH(ctx, knn_kernel<true>, div_up(nq, WPB), WPB * 32, 0, g->view(), nullptr, nq, k, idx_dev, d2_dev);
  else
    PFX_LAUNCH(ctx, knn_kernel<false>, div_up(nq, WPB), WPB * 32, 0, g->view(), q_dev, nq, k, idx_dev, d2_dev);
  PFX_CUDA(cudaGetLastError());
  return 0;
}

// dense rows of the queries the tile path handed back (device-side work list, persistent grid)
int knn_run_worklist(Ctx* ctx, Grid* g, int k, int* idx_dev, float* d2_dev, const int* worklist, const int* wl_count) {
  PFX_LAUNCH(ctx, knn_worklist_kernel, ctx->sm_count * 2, WPB * 32, 0, g->view(), worklist, wl_count, k, idx_dev, d2_dev);
  PFX_CUDA(cudaGetLastError());
  return 0;
}

// Fills ctx->knn_idx / knn_d2 for the current queries (cached per grid, k, query version).
int knn_lists(Ctx* ctx, Grid* g, int k, bool need_sorted) {
  if (k < 1 || k > 32) return ctx->fail(PFX_E_INVALID, "k must be in [1, 32]");
  const bool dense = ctx->q_is_surface;
  if (ctx->knn_grid == g && ctx->knn_k == k && ctx->knn_sversion == ctx->surf_version &&
      ctx->knn_dense == dense && (dense || ctx->knn_qversion == ctx->qry_version) &&
      (!need_sorted || ctx->knn_sorted))
    return 0;
  const int nq = (int)ctx->num_queries();
  PFX_CUDA(ctx->knn_idx.ensure((size_t)std::max(nq, 1) * k * sizeof(int)));
  PFX_CUDA(ctx->knn_d2.ensure((size_t)std::max(nq, 1) * k * sizeof(float)));
  PFX_TRY(knn_run(ctx, g, dense ? nullptr : ctx->qry.as<float4>(), nq, k, ctx->knn_idx.as<int>(),
                  ctx->knn_d2.as<float>()));
  ctx->knn_grid = g;
  ctx->knn_k = k;
  ctx->knn_sversion = ctx->surf_version;
  ctx->knn_qversion = ctx->qry_version;
  ctx->knn_dense = dense;
  ctx->knn_sorted = true;
  ctx->tile_has_normals = false;
  return 0;
}

// rows back to caller order, neighbour ids back to original indices
__global__ void knn_export_kernel(GridDev g, const int* __restrict__ lidx, const float* __restrict__ ld2,
                                  int nrows, int k, int dense, int* __restrict__ oidx,
                                  float* __restrict__ od2) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)nrows * k) return;
  int row = (int)(t / k), c = (int)(t % k);
  int j = lidx[t];
  int orow = dense ? __float_as_int(g.pts[row].w) : row;
  oidx[(size_t)orow * k + c] = (j >= 0) ? __float_as_int(g.pts[j].w) : -1;
  od2[(size_t)orow * k + c] = ld2[t];
}

int knn_export(Ctx* ctx, int k, int32_t* idx_dev, float* d2_dev, int) {
  const int nq = (int)ctx->num_queries();
  if (nq == 0) return 0;
  long long tot = (long long)nq * k;
  PFX_LAUNCH(ctx, knn_export_kernel, div_up(tot, 256), 256, 0, ctx->knn_grid->view(), ctx->knn_idx.as<int>(),
             ctx->knn_d2.as<float>(), nq, k, ctx->knn_dense ? 1 : 0, idx_dev, d2_dev);
  PFX_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------- radius
// counts / lists are in CALLER query order (row = original surface index when the queries are the
// surface), neighbour ids are original indices.
template <bool DENSE, bool FILL>
__global__ void __launch_bounds__(WPB * 32)
radius_kernel(GridDev g, const float4* __restrict__ queries, int nq, float r2, int* __restrict__ counts,
              const long long* __restrict__ offsets, int* __restrict__ oidx, float* __restrict__ od2) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int qi = blockIdx.x * WPB + wid;
  if (qi >= nq) return;
  const GridParams P = *g.gp;
  float4 q = DENSE ? g.pts[qi] : queries[qi];
  const int row = DENSE ? __float_as_int(q.w) : qi;
  int total = 0;
  if (finite3(q.x, q.y, q.z) && (!DENSE || qi < P.n_valid)) {
    CellBlock blk = DENSE ? stencil_of_point(g, qi, lane) : stencil_of_pos(g, q.x, q.y, q.z, lane);
    long long off = FILL ? offsets[row] : 0;
    const unsigned lt = (1u << lane) - 1u;
    for (int base = 0; base < blk.total; base += 32) {
      int t = base + lane;
      bool valid = t < blk.total;
      int j = block_candidate(blk, valid ? t : 0);
      bool in = false;
      float d2 = 0.f;
      float4 p;
      if (valid) {
        p = g.pts[j];
        d2 = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);
        in = d2 < r2;
      }
      unsigned m = __ballot_sync(FULL, in);
      if (FILL && in) {
        long long dst = off + total + __popc(m & lt);
        oidx[dst] = __float_as_int(p.w);
        od2[dst] = d2;
      }
      total += __popc(m);
    }
  }
  if (!FILL && lane == 0) counts[row] = total;
}

int radius_count(Ctx* ctx, Grid* g, double radius, int* counts_dev) {
  const int nq = (int)ctx->num_queries();
  if (nq == 0) return 0;
  float r2 = (float)(radius * radius);
  if (ctx->q_is_surface) {
    // rows of non-finite surface points are not visited by the kernel's sorted range
    PFX_CUDA(cudaMemsetAsync(counts_dev, 0, (size_t)nq * sizeof(int), ctx->stream));
    PFX_LAUNCH(ctx, (radius_kernel<true, false>), div_up(nq, WPB), WPB * 32, 0, g->view(), nullptr, nq, r2,
               counts_dev, nullptr, nullptr, nullptr);
  } else {
    PFX_LAUNCH(ctx, (radius_kernel<false, false>), div_up(nq, WPB), WPB * 32, 0, g->view(),
               ctx->qry.as<float4>(), nq, r2, counts_dev, nullptr, nullptr, nullptr);
  }
  PFX_CUDA(cudaGetLastError());
  return 0;
}

// per-row sort by (d2, index): one block per row, bitonic in shared memory (rows <= 4096)
constexpr int SORT_CAP = 4096;
__global__ void __launch_bounds__(256)
row_sort_kernel(const long long* __restrict__ offsets, int* __restrict__ idx, float* __restrict__ d2,
                int* __restrict__ overflow) {
  __shared__ unsigned long long sk[SORT_CAP];
  const long long o0 = offsets[blockIdx.x], o1 = offsets[blockIdx.x + 1];
  const int m = (int)(o1 - o0);
  if (m <= 1) return;
  if (m > SORT_CAP) {
    if (threadIdx.x == 0) atomicAdd(overflow, 1);
    return;
  }
  int m2 = 1;
  while (m2 < m) m2 <<= 1;
  for (int i = threadIdx.x; i < m2; i += blockDim.x)
    sk[i] = (i < m) ? (((unsigned long long)__float_as_uint(d2[o0 + i]) << 32) | (unsigned)idx[o0 + i]) : KMAX;
  __syncthreads();
  for (int k = 2; k <= m2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < m2; i += blockDim.x) {
        int p = i ^ j;
        if (p > i) {
          unsigned long long a = sk[i], b = sk[p];
          bool up = ((i & k) == 0);
          if ((a > b) == up) {
            sk[i] = b;
            sk[p] = a;
          }
        }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    idx[o0 + i] = (int)(unsigned)(sk[i] & 0xffffffffull);
    d2[o0 + i] = __uint_as_float((unsigned)(sk[i] >> 32));
  }
}

int radius_fill(Ctx* ctx, Grid* g, double radius, int sorted, const long long* offsets_dev, int* idx_dev,
                float* d2_dev) {
  const int nq = (int)ctx->num_queries();
  if (nq == 0) return 0;
  float r2 = (float)(radius * radius);
  if (ctx->q_is_surface)
    PFX_LAUNCH(ctx, (radius_kernel<true, true>), div_up(nq, WPB), WPB * 32, 0, g->view(), nullptr, nq, r2,
               nullptr, offsets_dev, idx_dev, d2_dev);
  else
    PFX_LAUNCH(ctx, (radius_kernel<false, true>), div_up(nq, WPB), WPB * 32, 0, g->view(),
               ctx->qry.as<float4>(), nq, r2, nullptr, offsets_dev, idx_dev, d2_dev);
  if (sorted) {
    PFX_CUDA(ctx->small.ensure(256));
    PFX_CUDA(cudaMemsetAsync(ctx->small.p, 0, sizeof(int), ctx->stream));
    PFX_LAUNCH(ctx, row_sort_kernel, nq, 256, 0, offsets_dev, idx_dev, d2_dev, ctx->small.as<int>());
    int ov = 0;
    PFX_CUDA(cudaMemcpyAsync(&ov, ctx->small.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PFX_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ov) return ctx->fail(PFX_E_CAPACITY, "sorted radius search: a row exceeds 4096 neighbours");
  }
  PFX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pfx
