// shot.cu — SHOT local reference frame + SHOT352 (replaces pcl::SHOTEstimationOMP::compute and its
// internal SHOTLocalReferenceFrameEstimationOMP as instantiated at reference evaluation.cpp:770-775
// and driven through features.h:181-195).
//
// K12 lrf_kernel: one warp per query, 32 queries per warp.  Pass 1 accumulates the (R - d) weighted
// scatter matrix in double, the reduced matrix of query t is parked in lane t and all lanes solve
// their 3x3 eigen problem together (double Jacobi); pass 2 re-scans the neighbourhood for the sign
// votes.  Vote ties (2*count == n) use PCL's median-neighbour fallback, which needs the neighbours'
// rank in (d2, index) order: lrf_tie_kernel handles exactly those queries, one block each.
// K13 shot_kernel: one warp per query, the 352-bin histogram lives in shared memory as int32 fixed
// point (power-of-two scale chosen from the neighbour count) so that accumulation order cannot
// change the result: the output is bit-reproducible run to run and across shardings.  All DISCRETE
// decisions (volume index, cosine step, radial / elevation branch) replicate the CPU arithmetic
// exactly (float dots, double compares); only the continuous interpolation weights use float
// acosf / atan2f.
#include "internal.h"

namespace pfx {

constexpr int SWPB = 4;

struct LrfAcc {
  double m[6];  // xx xy xz yy yz zz
  double sw;
  int nvalid;
};

__device__ __forceinline__ void lrf_scan_pass1(const GridDev& g, const CellBlock& blk, float4 q, float r2,
                                               double R, int lane, LrfAcc& a) {
  for (int base = 0; base < blk.total; base += 32) {
    int c = base + lane;
    bool valid = c < blk.total;
    int j = block_candidate(blk, valid ? c : 0);
    if (valid) {
      float4 p = g.pts[j];
      float d2 = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);
      if (d2 < r2 && !(p.x == q.x && p.y == q.y && p.z == q.z)) {
        double vx = (double)__fsub_rn(p.x, q.x), vy = (double)__fsub_rn(p.y, q.y), vz = (double)__fsub_rn(p.z, q.z);
        double w = R - sqrt((double)d2);
        a.m[0] += w * (vx * vx); a.m[1] += w * (vx * vy); a.m[2] += w * (vx * vz);
        a.m[3] += w * (vy * vy); a.m[4] += w * (vy * vz); a.m[5] += w * (vz * vz);
        a.sw += w;
        a.nvalid += 1;
      }
    }
  }
}

// --- vote-tie fallback helpers -------------------------------------------------------------
constexpr int WTIE_CAP = 512;  // valid neighbours a warp can rank in shared memory

// collect the keys (d2 bits << 32 | original index) of the valid neighbours into `keys`
__device__ __forceinline__ int lrf_collect_keys(const GridDev& g, const CellBlock& blk, float4 q, float r2,
                                                int lane, unsigned long long* keys, int cap, int n) {
  const unsigned lt = (1u << lane) - 1u;
  for (int base = 0; base < blk.total; base += 32) {
    int c = base + lane;
    bool valid = c < blk.total;
    int j = block_candidate(blk, valid ? c : 0);
    unsigned long long key = 0;
    if (valid) {
      float4 p = g.pts[j];
      float d2 = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);
      valid = d2 < r2 && !(p.x == q.x && p.y == q.y && p.z == q.z);
      key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)__float_as_int(p.w);
    }
    unsigned m = __ballot_sync(FULL, valid);
    int pos = n + __popc(m & lt);
    if (valid && pos < cap) keys[pos] = key;
    n += __popc(m);
  }
  return n;
}

// PCL: among the valid neighbours in (d2, index) order, ranks n/2-2 .. n/2+2; count strictly
// positive projections on each axis
__device__ __forceinline__ void lrf_median_votes(const GridDev& g, float4 q, const unsigned long long* keys,
                                                 int n, const double xs[3], const double zs[3], int lane,
                                                 int& plus_x, int& plus_z) {
  const int med = n / 2;
  int px = 0, pz = 0;
  for (int a = lane; a < n; a += 32) {
    unsigned long long ka = keys[a];
    int rank = 0;
    for (int b = 0; b < n; ++b) rank += (keys[b] < ka) ? 1 : 0;
    if (rank >= med - 2 && rank <= med + 2) {
      float4 p = g.pts[g.inv_perm[(int)(unsigned)(ka & 0xffffffffull)]];
      double vx = (double)__fsub_rn(p.x, q.x), vy = (double)__fsub_rn(p.y, q.y), vz = (double)__fsub_rn(p.z, q.z);
      if (vx * xs[0] + vy * xs[1] + vz * xs[2] > 0) ++px;
      if (vx * zs[0] + vy * zs[1] + vz * zs[2] > 0) ++pz;
    }
  }
  plus_x = warp_sum(px);
  plus_z = warp_sum(pz);
}

__device__ __forceinline__ void lrf_store(float* o, const double x[3], const double z[3]) {
  float fx[3] = {(float)x[0], (float)x[1], (float)x[2]};
  float fz[3] = {(float)z[0], (float)z[1], (float)z[2]};
  o[0] = fx[0]; o[1] = fx[1]; o[2] = fx[2];
  o[6] = fz[0]; o[7] = fz[1]; o[8] = fz[2];
  // y = z x x in float (rf.row(1) = rf.row(2).cross(rf.row(0)))
  o[3] = __fsub_rn(__fmul_rn(fz[1], fx[2]), __fmul_rn(fz[2], fx[1]));
  o[4] = __fsub_rn(__fmul_rn(fz[2], fx[0]), __fmul_rn(fz[0], fx[2]));
  o[5] = __fsub_rn(__fmul_rn(fz[0], fx[1]), __fmul_rn(fz[1], fx[0]));
}

struct TieItem {  // a query whose tie must be resolved by the block-level kernel (n > WTIE_CAP)
  int qi, tflag;
  double x[3], z[3];
};

// rf9 rows in caller query order
template <bool DENSE>
__global__ void __launch_bounds__(SWPB * 32)
lrf_kernel(GridDev g, const float4* __restrict__ queries, int nq, float r2, double R, float* __restrict__ rf9,
           TieItem* __restrict__ tie_list, int* __restrict__ tie_count, int tie_cap, const int* __restrict__ qmap,
           const int* __restrict__ qcount, int qpw /* queries per warp, 1..32 */, float need) {
  __shared__ unsigned long long skeys[SWPB][WTIE_CAP];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int qbase = (blockIdx.x * SWPB + wid) * qpw;
  const int limit = qmap ? min(*qcount, nq) : nq;  // qmap: work list of query numbers (optional)
  if (qbase >= limit) return;
  const int n_valid = g.gp->n_valid;
  const int rings = stencil_rings(*g.gp, need), nch = stencil_chunks(rings);  // 1 / 1 on a grid built for the radius
  const int qend = min(qpw, limit - qbase);
  LrfAcc mine;
#pragma unroll
  for (int i = 0; i < 6; ++i) mine.m[i] = 0.0;
  mine.sw = 0.0;
  mine.nvalid = 0;
  for (int t = 0; t < qend; ++t) {
    const int qi = qmap ? qmap[qbase + t] : qbase + t;
    float4 q = DENSE ? g.pts[qi] : queries[qi];
    LrfAcc a;
#pragma unroll
    for (int i = 0; i < 6; ++i) a.m[i] = 0.0;
    a.sw = 0.0;
    a.nvalid = 0;
    if (finite3(q.x, q.y, q.z) && (!DENSE || qi < n_valid)) {
      for (int ch = 0; ch < nch; ++ch) {
        CellBlock blk = stencil_block<DENSE>(g, qi, q, rings, ch, lane);
        lrf_scan_pass1(g, blk, q, r2, R, lane, a);
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) a.m[i] = warp_sum(a.m[i]);
      a.sw = warp_sum(a.sw);
      a.nvalid = warp_sum(a.nvalid);
    }
    if (lane == t) mine = a;
  }
  // every lane solves its own query
  double x[3] = {0, 0, 0}, z[3] = {0, 0, 0};
  bool good = false;
  if (lane < qend && mine.nvalid >= 5) {
    double a[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) a[i] = mine.m[i] / mine.sw;
    double w[3], v[3][3];
    eig_sym3<double>(a, w, v, 12);
    good = isfinite(w[0]) && isfinite(w[1]) && isfinite(w[2]);
    x[0] = v[0][2]; x[1] = v[1][2]; x[2] = v[2][2];
    z[0] = v[0][0]; z[1] = v[1][0]; z[2] = v[2][0];
  }
  // pass 2: sign votes (+ PCL's median fallback on ties)
  bool flip_x = false, flip_z = false;
  for (int t = 0; t < qend; ++t) {
    const int qi = qmap ? qmap[qbase + t] : qbase + t;
    bool gd = __shfl_sync(FULL, (int)good, t);
    if (!gd) continue;
    const int nv = __shfl_sync(FULL, mine.nvalid, t);
    double xs[3], zs[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      xs[i] = __shfl_sync(FULL, x[i], t);
      zs[i] = __shfl_sync(FULL, z[i], t);
    }
    float4 q = DENSE ? g.pts[qi] : queries[qi];
    int px = 0, pz = 0;
    for (int ch = 0; ch < nch; ++ch) {
      CellBlock blk = stencil_block<DENSE>(g, qi, q, rings, ch, lane);
      for (int base = 0; base < blk.total; base += 32) {
        int c = base + lane;
        bool valid = c < blk.total;
        int j = block_candidate(blk, valid ? c : 0);
        if (valid) {
          float4 p = g.pts[j];
          float d2 = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);
          if (d2 < r2 && !(p.x == q.x && p.y == q.y && p.z == q.z)) {
            double vx = (double)__fsub_rn(p.x, q.x), vy = (double)__fsub_rn(p.y, q.y), vz = (double)__fsub_rn(p.z, q.z);
            if (vx * xs[0] + vy * xs[1] + vz * xs[2] >= 0) ++px;
            if (vx * zs[0] + vy * zs[1] + vz * zs[2] >= 0) ++pz;
          }
        }
      }
    }
    px = warp_sum(px);
    pz = warp_sum(pz);
    int votex = 2 * px - nv, votez = 2 * pz - nv;
    bool fx = votex < 0, fz = votez < 0;
    if (votex == 0 || votez == 0) {
      if (nv <= WTIE_CAP) {
        int n = 0;
        for (int ch = 0; ch < nch; ++ch) {
          CellBlock blk = stencil_block<DENSE>(g, qi, q, rings, ch, lane);
          n = lrf_collect_keys(g, blk, q, r2, lane, skeys[wid], WTIE_CAP, n);
        }
        __syncwarp();
        int plx, plz;
        lrf_median_votes(g, q, skeys[wid], n, xs, zs, lane, plx, plz);
        if (votex == 0) fx = plx < 3;
        if (votez == 0) fz = plz < 3;
        __syncwarp();
      } else if (lane == 0) {  // too many neighbours for the warp buffer: block-level kernel
        int slot = atomicAdd(tie_count, 1);
        if (slot < tie_cap) {
          TieItem it;
          it.qi = qi;
          it.tflag = (votex == 0 ? 1 : 0) | (votez == 0 ? 2 : 0);
          for (int i = 0; i < 3; ++i) {
            it.x[i] = fx ? -xs[i] : xs[i];
            it.z[i] = fz ? -zs[i] : zs[i];
          }
          tie_list[slot] = it;
        }
      }
    }
    if (lane == t) {
      flip_x = fx;
      flip_z = fz;
    }
  }
  if (lane < qend) {
    const int qi = qmap ? qmap[qbase + lane] : qbase + lane;
    float4 q = DENSE ? g.pts[qi] : queries[qi];
    const size_t row = DENSE ? (size_t)__float_as_int(q.w) : (size_t)qi;
    float* o = rf9 + row * 9;
    const float nanv = __int_as_float(0x7fc00000);
    if (!good) {
#pragma unroll
      for (int i = 0; i < 9; ++i) o[i] = nanv;
    } else {
      if (flip_x) { x[0] = -x[0]; x[1] = -x[1]; x[2] = -x[2]; }
      if (flip_z) { z[0] = -z[0]; z[1] = -z[1]; z[2] = -z[2]; }
      lrf_store(o, x, z);
    }
  }
}

// Block-level tie resolution for neighbourhoods larger than the warp buffer (persistent grid over
// the device-side work list; usually empty).
constexpr int TIE_CAP = 6000;
template <bool DENSE>
__global__ void __launch_bounds__(128)
lrf_tie_kernel(GridDev g, const float4* __restrict__ queries, float r2, float* __restrict__ rf9,
               const TieItem* __restrict__ tie_list, const int* __restrict__ tie_count, int tie_cap,
               int* __restrict__ overflow, float need) {
  __shared__ unsigned long long keys[TIE_CAP];
  __shared__ int cnt, plus_x, plus_z;
  const int total = min(*tie_count, tie_cap);
  if (*tie_count > tie_cap && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(overflow, 1);
  const GridParams P = *g.gp;
  const int rings = stencil_rings(P, need), side = 2 * rings + 1;
  for (int e = blockIdx.x; e < total; e += gridDim.x) {
    const TieItem it = tie_list[e];
    const int qi = it.qi;
    float4 q = DENSE ? g.pts[qi] : queries[qi];
    const size_t row = DENSE ? (size_t)__float_as_int(q.w) : (size_t)qi;
    __syncthreads();
    if (threadIdx.x == 0) {
      cnt = 0;
      plus_x = 0;
      plus_z = 0;
    }
    __syncthreads();
    int cx = cell_coord(q.x, P.ox, P.inv_e, P.nx), cy = cell_coord(q.y, P.oy, P.inv_e, P.ny),
        cz = cell_coord(q.z, P.oz, P.inv_e, P.nz);
    for (int l = 0; l < side * side * side; ++l) {
      int x2 = cx + l % side - rings, y2 = cy + (l / side) % side - rings, z2 = cz + l / (side * side) - rings;
      if (x2 < 0 || x2 >= P.nx || y2 < 0 || y2 >= P.ny || z2 < 0 || z2 >= P.nz) continue;
      int c = hash_lookup(g, morton3(x2, y2, z2));
      if (c < 0) continue;
      int s = g.cell_start[c], en = g.cell_start[c + 1];
      for (int j = s + threadIdx.x; j < en; j += blockDim.x) {
        float4 p = g.pts[j];
        float d2 = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);
        if (d2 < r2 && !(p.x == q.x && p.y == q.y && p.z == q.z)) {
          int slot = atomicAdd(&cnt, 1);
          if (slot < TIE_CAP)
            keys[slot] = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)__float_as_int(p.w);
        }
      }
    }
    __syncthreads();
    const int n = cnt;
    if (n > TIE_CAP) {
      if (threadIdx.x == 0) atomicAdd(overflow, 1);
      continue;
    }
    const int med = n / 2;
    for (int a = threadIdx.x; a < n; a += blockDim.x) {
      unsigned long long ka = keys[a];
      int rank = 0;
      for (int b = 0; b < n; ++b) rank += (keys[b] < ka) ? 1 : 0;
      if (rank >= med - 2 && rank <= med + 2) {
        float4 p = g.pts[g.inv_perm[(int)(unsigned)(ka & 0xffffffffull)]];
        double vx = (double)__fsub_rn(p.x, q.x), vy = (double)__fsub_rn(p.y, q.y), vz = (double)__fsub_rn(p.z, q.z);
        if (vx * it.x[0] + vy * it.x[1] + vz * it.x[2] > 0) atomicAdd(&plus_x, 1);
        if (vx * it.z[0] + vy * it.z[1] + vz * it.z[2] > 0) atomicAdd(&plus_z, 1);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double x[3] = {it.x[0], it.x[1], it.x[2]}, z[3] = {it.z[0], it.z[1], it.z[2]};
      if ((it.tflag & 1) && plus_x < 3) { x[0] = -x[0]; x[1] = -x[1]; x[2] = -x[2]; }
      if ((it.tflag & 2) && plus_z < 3) { z[0] = -z[0]; z[1] = -z[1]; z[2] = -z[2]; }
      lrf_store(rf9 + row * 9, x, z);
    }
  }
}

static int lrf_run(Ctx* ctx, Grid* g, double radius, float* rf9_dev, const int* qmap, const int* qcount);

int shot_lrf_compute(Ctx* ctx, Grid* g, double radius, float* rf9_dev, int*) {
  return lrf_run(ctx, g, radius, rf9_dev, nullptr, nullptr);
}

// frames of the queries on a device-side work list only (rows of the others are left untouched)
int shot_lrf_worklist(Ctx* ctx, Grid* g, double radius, float* rf9_dev, const int* wl, const int* wl_count) {
  return lrf_run(ctx, g, radius, rf9_dev, wl, wl_count);
}

static int lrf_run(Ctx* ctx, Grid* g, double radius, float* rf9_dev, const int* qmap, const int* qcount) {
  const int nq = (int)ctx->num_queries();
  if (nq == 0) return 0;
  const bool dense = ctx->q_is_surface;
  const float r2 = (float)(radius * radius);
  const float need = (float)(radius * (1.0 + 1e-3));  // the edge of a radius grid: one ring of cells on such a grid
  const int tie_cap = 65536;
  PFX_CUDA(ctx->tmp1.ensure((size_t)tie_cap * sizeof(TieItem)));
  PFX_CUDA(ctx->small.ensure(256));
  TieItem* tl = ctx->tmp1.as<TieItem>();
  int* flags = ctx->small.as<int>() + 32;  // [0] tie count, [1] overflow
  PFX_CUDA(cudaMemsetAsync(flags, 0, 2 * sizeof(int), ctx->stream));
  // Queries per warp: 32 keeps every lane busy in the eigen solve, but a few thousand keypoint queries with
  // hundreds of neighbours each need the warps for the neighbourhood scans instead.
  const int qpw = std::max(1, std::min(32, nq / (ctx->sm_count * 16)));
  const int blocks = div_up(nq, SWPB * qpw);
  if (dense) {
    PFX_LAUNCH(ctx, lrf_kernel<true>, blocks, SWPB * 32, 0, g->view(), nullptr, nq, r2, radius, rf9_dev, tl, flags,
               tie_cap, qmap, qcount, qpw, need);
    PFX_LAUNCH(ctx, lrf_tie_kernel<true>, ctx->sm_count, 128, 0, g->view(), nullptr, r2, rf9_dev, tl, flags, tie_cap,
               flags + 1, need);
  } else {
    PFX_LAUNCH(ctx, lrf_kernel<false>, blocks, SWPB * 32, 0, g->view(), ctx->qry.as<float4>(), nq, r2, radius,
               rf9_dev, tl, flags, tie_cap, qmap, qcount, qpw, need);
    PFX_LAUNCH(ctx, lrf_tie_kernel<false>, ctx->sm_count, 128, 0, g->view(), ctx->qry.as<float4>(), r2, rf9_dev, tl,
               flags, tie_cap, flags + 1, need);
  }
  PFX_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------- SHOT352
__device__ __forceinline__ void shot_add(int* h, int slot, double v, float scale) {
  atomicAdd(&h[slot], __double2int_rn(v * (double)scale));
}

template <bool DENSE>
__device__ __forceinline__ void shot_one_query(const GridDev& g, const float4* __restrict__ queries, int qi,
                                               const float4* __restrict__ nrm, float r2, double R,
                                               const float* __restrict__ rf9, float* __restrict__ out, size_t stride,
                                               int* h, int lane, float need) {
  const int n_valid = g.gp->n_valid;
  const int rings = stencil_rings(*g.gp, need), nch = stencil_chunks(rings);  // 1 / 1 on a grid built for the radius
  for (int c = lane; c < 352; c += 32) h[c] = 0;
  __syncwarp();
  float4 q = DENSE ? g.pts[qi] : queries[qi];
  const size_t row = DENSE ? (size_t)__float_as_int(q.w) : (size_t)qi;
  float* o = out + row * stride;
  const float nanv = __int_as_float(0x7fc00000);
  float rf[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) rf[i] = rf9[row * 9 + i];
  bool ok = finite3(q.x, q.y, q.z) && (!DENSE || qi < n_valid) && isfinite(rf[0]) && isfinite(rf[3]) &&
            isfinite(rf[6]);
  int n_nb = 0;
  if (ok) {
    // neighbour count first: it fixes the fixed-point scale (and the < 5 rule)
    for (int ch = 0; ch < nch; ++ch) {
      const CellBlock blk = stencil_block<DENSE>(g, qi, q, rings, ch, lane);
      for (int base = 0; base < blk.total; base += 32) {
        int c = base + lane;
        bool valid = c < blk.total;
        int j = block_candidate(blk, valid ? c : 0);
        if (valid) {
          float4 p = g.pts[j];
          valid = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z) < r2;
        }
        n_nb += __popc(__ballot_sync(FULL, valid));
      }
    }
  }
  if (!ok || n_nb == 0) {
    for (int c = lane; c < 352; c += 32) o[c] = nanv;
    if (lane < 9) o[352 + lane] = nanv;
    return;
  }
  if (lane < 9) o[352 + lane] = rf[lane];
  if (n_nb < 5) {  // computePointSHOT: too few neighbours -> NaN descriptor, frame kept
    for (int c = lane; c < 352; c += 32) o[c] = nanv;
    return;
  }
  // every bin receives at most 4.5 per neighbour
  int bits = 32 - __clz(5 * n_nb + 8);
  const float scale = exp2f((float)(30 - bits));
  const double r12 = R / 2, r14 = R / 4, r34 = 3 * R / 4;
  const float RAD45 = 0.78539816339744830962f, RAD90 = 1.57079632679489661923f, RAD135 = 2.35619449019234492885f,
              RAD_PI_7_8 = 2.7488935718910690836f;
  for (int ch = 0; ch < nch; ++ch) {
  const CellBlock blk = stencil_block<DENSE>(g, qi, q, rings, ch, lane);
  for (int base = 0; base < blk.total; base += 32) {
    int c = base + lane;
    bool valid = c < blk.total;
    int j = block_candidate(blk, valid ? c : 0);
    if (!valid) continue;
    float4 p = g.pts[j];
    float d2 = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);
    if (!(d2 < r2)) continue;
    float4 nj = nrm[j];
    if (!finite3(nj.x, nj.y, nj.z)) continue;
    double cosd = (double)__fadd_rn(__fadd_rn(__fmul_rn(nj.x, rf[6]), __fmul_rn(nj.y, rf[7])), __fmul_rn(nj.z, rf[8]));
    cosd = fmin(1.0, fmax(-1.0, cosd));
    double bd = ((1.0 + cosd) * 10) / 2;
    float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y), dz = __fsub_rn(p.z, q.z);
    double dist = sqrt((double)d2);
    if (fabs(dist) < 1e-15) continue;
    double x = (double)__fadd_rn(__fadd_rn(__fmul_rn(dx, rf[0]), __fmul_rn(dy, rf[1])), __fmul_rn(dz, rf[2]));
    double y = (double)__fadd_rn(__fadd_rn(__fmul_rn(dx, rf[3]), __fmul_rn(dy, rf[4])), __fmul_rn(dz, rf[5]));
    double z = (double)__fadd_rn(__fadd_rn(__fmul_rn(dx, rf[6]), __fmul_rn(dy, rf[7])), __fmul_rn(dz, rf[8]));
    if (fabs(y) < 1e-30) y = 0;
    if (fabs(x) < 1e-30) x = 0;
    if (fabs(z) < 1e-30) z = 0;
    int bit4 = ((y > 0) || ((y == 0.0) && (x < 0))) ? 1 : 0;
    int bit3 = ((x > 0) || ((x == 0.0) && (y > 0))) ? !bit4 : bit4;
    int di = ((bit4 << 3) + (bit3 << 2)) << 1;
    if ((x * y > 0) || (x == 0.0))
      di += (fabs(x) >= fabs(y)) ? 0 : 4;
    else
      di += (fabs(x) > fabs(y)) ? 4 : 0;
    di += z > 0 ? 1 : 0;
    di += (dist > r12) ? 2 : 0;
    int step = (int)floor(bd + 0.5);
    int vol = di * 11;
    bd -= step;
    double w = 1 - fabs(bd);
    if (bd > 0)
      shot_add(h, vol + ((step + 1) % 10), bd, scale);
    else
      shot_add(h, vol + ((step - 1 + 10) % 10), -bd, scale);
    if (dist > r12) {
      double rd = (dist - r34) / r12;
      if (dist > r34)
        w += 1 - rd;
      else {
        w += 1 + rd;
        shot_add(h, (di - 2) * 11 + step, -rd, scale);
      }
    } else {
      double rd = (dist - r14) / r12;
      if (dist < r14)
        w += 1 + rd;
      else {
        w += 1 - rd;
        shot_add(h, (di + 2) * 11 + step, rd, scale);
      }
    }
    float ic = (float)fmin(1.0, fmax(-1.0, z / dist));
    float inc = acosf(ic);
    if (z <= 0) {  // == (inc > 90deg || (|inc - 90deg| < 1e-30 && z <= 0)) in exact arithmetic
      float e = (inc - RAD135) / RAD90;
      if (inc > RAD135)
        w += 1 - e;
      else {
        w += 1 + e;
        shot_add(h, (di + 1) * 11 + step, -e, scale);
      }
    } else {
      float e = (inc - RAD45) / RAD90;
      if (inc < RAD45)
        w += 1 + e;
      else {
        w += 1 - e;
        shot_add(h, (di - 1) * 11 + step, e, scale);
      }
    }
    if (y != 0.0 || x != 0.0) {
      float az = atan2f((float)y, (float)x);
      int sel = di >> 2;
      float ad = (az - (-RAD_PI_7_8 + RAD45 * sel)) / RAD45;
      ad = fmaxf(-0.5f, fminf(ad, 0.5f));
      if (ad > 0) {
        w += 1 - ad;
        shot_add(h, ((di + 4) % 32) * 11 + step, ad, scale);
      } else {
        w += 1 + ad;
        shot_add(h, ((di - 4 + 32) % 32) * 11 + step, -ad, scale);
      }
    }
    shot_add(h, vol + step, w, scale);
  }
  }
  __syncwarp();
  // normalise: h / sqrt(sum h^2)
  const float inv_scale = 1.0f / scale;
  double acc = 0.0;
  for (int c = lane; c < 352; c += 32) {
    float v = (float)h[c] * inv_scale;
    acc += (double)__fmul_rn(v, v);
  }
  acc = warp_sum(acc);
  float nrmv = (float)sqrt(acc);
  for (int c = lane; c < 352; c += 32) o[c] = __fdiv_rn((float)h[c] * inv_scale, nrmv);
}

template <bool DENSE>
__global__ void __launch_bounds__(SWPB * 32)
shot_kernel(GridDev g, const float4* __restrict__ queries, int nq, const float4* __restrict__ nrm, float r2,
            double R, const float* __restrict__ rf9, float* __restrict__ out, size_t stride, float need) {
  __shared__ int hist[SWPB][352];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int qi = blockIdx.x * SWPB + wid;
  if (qi >= nq) return;
  shot_one_query<DENSE>(g, queries, qi, nrm, r2, R, rf9, out, stride, hist[wid], lane, need);
}

// persistent variant over a device-side work list (queries the fused kernel handed back)
template <bool DENSE>
__global__ void __launch_bounds__(SWPB * 32)
shot_worklist_kernel(GridDev g, const float4* __restrict__ queries, const int* __restrict__ wl,
                     const int* __restrict__ wl_count, const float4* __restrict__ nrm, float r2, double R,
                     const float* __restrict__ rf9, float* __restrict__ out, size_t stride, float need) {
  __shared__ int hist[SWPB][352];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int count = *wl_count;
  for (int w = blockIdx.x * SWPB + wid; w < count; w += gridDim.x * SWPB) {
    shot_one_query<DENSE>(g, queries, wl[w], nrm, r2, R, rf9, out, stride, hist[wid], lane, need);
    __syncwarp();
  }
}

int shot_worklist(Ctx* ctx, Grid* g, double radius, const float* rf9_dev, float* out_dev, size_t stride_floats,
                  const int* wl, const int* wl_count) {
  const float r2 = (float)(radius * radius);
  const float need = (float)(radius * (1.0 + 1e-3));
  const float4* nrm = nullptr;
  PFX_TRY(normals_sorted_for_grid(ctx, g, &nrm));
  if (ctx->q_is_surface)
    PFX_LAUNCH(ctx, shot_worklist_kernel<true>, ctx->sm_count * 4, SWPB * 32, 0, g->view(), nullptr, wl, wl_count, nrm,
               r2, radius, rf9_dev, out_dev, stride_floats, need);
  else
    PFX_LAUNCH(ctx, shot_worklist_kernel<false>, ctx->sm_count * 4, SWPB * 32, 0, g->view(), ctx->qry.as<float4>(), wl,
               wl_count, nrm, r2, radius, rf9_dev, out_dev, stride_floats, need);
  PFX_CUDA(cudaGetLastError());
  return 0;
}

int shot_compute(Ctx* ctx, Grid* g, double radius, const float* rf9_dev, float* out_dev, size_t stride_floats) {
  const int nq = (int)ctx->num_queries();
  if (nq == 0) return 0;
  const float r2 = (float)(radius * radius);
  const float need = (float)(radius * (1.0 + 1e-3));
  const float4* nrm = nullptr;
  PFX_TRY(normals_sorted_for_grid(ctx, g, &nrm));
  const int blocks = div_up(nq, SWPB);
  if (ctx->q_is_surface)
    PFX_LAUNCH(ctx, shot_kernel<true>, blocks, SWPB * 32, 0, g->view(), nullptr, nq, nrm, r2, radius, rf9_dev,
               out_dev, stride_floats, need);
  else
    PFX_LAUNCH(ctx, shot_kernel<false>, blocks, SWPB * 32, 0, g->view(), ctx->qry.as<float4>(), nq, nrm, r2, radius,
               rf9_dev, out_dev, stride_floats, need);
  PFX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pfx
