// fpfh.cu — SPFH + FPFH33 (replaces pcl::FPFHEstimation::compute as instantiated at reference
// evaluation.cpp:597-602 and driven through features.h:181-195).
//
// SPFH rows.  PCL adds the constant hist_incr = 100/(n-1) once per pair, so a row is fully described
// by its 33 integer hit counts and n.  In k-search n is the same for every point, hence rows are stored
// as 36 BYTES of counts (4x less HBM/L2/shared traffic than 33 floats) and a 33-entry table
// T[c] = (((incr + incr) + incr) ...) c times reproduces PCL's sequential float sum bit for bit.
// In radius search n varies per point: rows are 33 floats with the additions replayed per bin.
//
// Kernels (one warp per point / query)
//  spfh_kernel       Darboux pair features of the neighbours, one per lane; the 3 x 11 histogram is
//                    counted with __match_any_sync (integer hit counts in shared memory, no atomics).
//  fpfh_list_kernel  k-search: three queries per warp, nine lanes per query (lane c = word c of the 36-byte
//                    count rows); the 1/d2-weighted sum is taken in fixed point, independent of list order.
//  fpfh_kernel       radius search: lanes are bins, 33-float rows gathered during the stencil scan.
#include "internal.h"
#include "pair_features.cuh"

namespace pfx {

constexpr int FWPB = 8;
constexpr int SROW = 36;  // bytes per count row

// the three bins of one pair (b1 in 0..10, b2 in 11..21, b3 in 22..32), or false when the pair fails
__device__ __forceinline__ bool pair_bins(float4 q, float4 nq, float4 p, float4 nj, int& b1, int& b2, int& b3) {
  const double d_pi = (double)(1.0f / (2.0f * 3.14159265358979323846f));
  float f1, f2, f3;
  if (!finite3(nj.x, nj.y, nj.z) || !pair_features<11>(q.x, q.y, q.z, nq, p.x, p.y, p.z, nj, f1, f2, f3)) return false;
  b1 = clamp_bin(11 * (((double)f1 + 3.14159265358979323846) * d_pi));
  b2 = 11 + clamp_bin(11 * (((double)f2 + 1.0) * 0.5));
  b3 = 22 + clamp_bin(11 * (((double)f3 + 1.0) * 0.5));
  return true;
}

// The same three bins with a third of the arithmetic: reciprocal square roots instead of the two IEEE square roots and
// five IEEE divisions, float bin coordinates instead of double.  Every approximation is a few ulp, i.e. at most 1e-4
// of a bin; a pair whose result could differ from the reference arithmetic - a bin coordinate within 2.5e-4 of an
// edge, the source / target swap decided by less than 1e-5, an ill-conditioned atan2 (|(sn, cs)| < 0.03), anything
// non-finite - is recomputed by pair_bins, so the bins are ALWAYS those of the exact path (about 1 pair in 800 is).
__device__ __forceinline__ bool pair_bins_fast(float4 q, float4 nq, float4 p, float4 nj, int& b1, int& b2, int& b3) {
  float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
  const float s4 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
  if (!(s4 > 1e-30f)) return pair_bins(q, nq, p, nj, b1, b2, b3);  // coincident (or denormal distance): exact path decides
  const float inv4 = rsqrtf(s4);
  const float a1 = fmaf(nq.z, dz, fmaf(nq.y, dy, nq.x * dx)) * inv4;
  const float a2 = fmaf(nj.z, dz, fmaf(nj.y, dy, nj.x * dx)) * inv4;
  const float m1 = fabsf(a1), m2 = fabsf(a2);
  const bool swap = (m1 < m2) && (m2 <= 1.0f);
  bool risky = fabsf(m1 - m2) < 1e-5f || fabsf(m2 - 1.0f) < 1e-5f;
  float ux, uy, uz, tx, ty, tz, f3;
  if (swap) {
    ux = nj.x; uy = nj.y; uz = nj.z;
    tx = nq.x; ty = nq.y; tz = nq.z;
    dx = -dx; dy = -dy; dz = -dz;
    f3 = -a2;
  } else {
    ux = nq.x; uy = nq.y; uz = nq.z;
    tx = nj.x; ty = nj.y; tz = nj.z;
    f3 = a1;
  }
  float vx = dy * uz - dz * uy, vy = dz * ux - dx * uz, vz = dx * uy - dy * ux;
  const float vs = fmaf(vz, vz, fmaf(vy, vy, vx * vx));
  const float vinv = rsqrtf(vs);
  vx *= vinv; vy *= vinv; vz *= vinv;
  const float wx = uy * vz - uz * vy, wy = uz * vx - ux * vz, wz = ux * vy - uy * vx;
  const float f2 = fmaf(vz, tz, fmaf(vy, ty, vx * tx));
  const float sn = fmaf(wz, tz, fmaf(wy, ty, wx * tx));
  const float cs = fmaf(uz, tz, fmaf(uy, ty, ux * tx));
  const float f1 = fast_atan2f(sn, cs);
  const float u1 = (f1 + 3.14159274f) * (11.0f * 0.159154943f);
  const float u2 = (f2 + 1.0f) * 5.5f, u3 = (f3 + 1.0f) * 5.5f;
  const float e1 = fabsf(u1 - rintf(u1)), e2 = fabsf(u2 - rintf(u2)), e3 = fabsf(u3 - rintf(u3));
  // (a NaN / Inf normal makes f2 and sn non-finite; vs == 0 makes vinv infinite: all caught by the last test)
  risky = risky || fminf(e1, fminf(e2, e3)) < 2.5e-4f || fmaf(sn, sn, cs * cs) < 1e-3f || !(fabsf(f2) + fabsf(sn) < 1e30f);
  if (risky) return pair_bins(q, nq, p, nj, b1, b2, b3);
  b1 = min(max((int)floorf(u1), 0), 10);
  b2 = 11 + min(max((int)floorf(u2), 0), 10);
  b3 = 22 + min(max((int)floorf(u3), 0), 10);
  return true;
}

// T[c] = c sequential float additions of incr (PCL's hist += hist_incr), c = 0..32
__device__ __forceinline__ void build_incr_table(float* T, int n_nb, int lane) {
  float incr = (n_nb > 1) ? __fdiv_rn(100.0f, (float)(n_nb - 1)) : 0.f;
  if (lane == 0) {
    float v = 0.f;
    T[0] = 0.f;
    for (int c = 1; c <= 32; ++c) {
      v = __fadd_rn(v, incr);
      T[c] = v;
    }
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------ generic SPFH
// rows in the sorted order of grid g.  COUNTS (list mode): 36-byte count rows; else 33 floats.
// flags / only (optional): restrict to marked points.
template <bool USE_LIST>
__global__ void __launch_bounds__(FWPB * 32, 6)
spfh_kernel(GridDev g, const float4* __restrict__ nrm, float r2, const int* __restrict__ lists, int k,
            const int* __restrict__ flags, const unsigned char* __restrict__ only, float* __restrict__ spfh,
            unsigned char* __restrict__ rows8) {
  __shared__ int hist[FWPB][36];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int i = blockIdx.x * FWPB + wid;
  const int n_valid = g.gp->n_valid;
  if (i >= n_valid) return;
  if (flags && !flags[i]) return;
  if (only && !only[i]) return;
  int* h = hist[wid];
  h[lane] = 0;
  if (lane < 4) h[32 + lane] = 0;
  __syncwarp();
  const float4 q = g.pts[i];
  const float4 nq = nrm[i];
  const bool nq_ok = finite3(nq.x, nq.y, nq.z);
  int n_nb = 0;

  auto consume = [&](int j, bool valid) {
    int b1 = -1, b2 = -1, b3 = -1;
    if (valid && j != i && nq_ok) {
      if (!pair_bins_fast(q, nq, g.pts[j], nrm[j], b1, b2, b3)) b1 = b2 = b3 = -1;
    }
    unsigned m1 = __match_any_sync(FULL, b1);
    unsigned m2 = __match_any_sync(FULL, b2);
    unsigned m3 = __match_any_sync(FULL, b3);
    if (b1 >= 0) {  // the three sub-histograms occupy disjoint slots; one leader per distinct bin
      if (lane == __ffs(m1) - 1) h[b1] += __popc(m1);
      if (lane == __ffs(m2) - 1) h[b2] += __popc(m2);
      if (lane == __ffs(m3) - 1) h[b3] += __popc(m3);
    }
    __syncwarp();
  };

  if (USE_LIST) {
    for (int c0 = 0; c0 < k; c0 += 32) {
      int c = c0 + lane;
      int j = (c < k) ? lists[(size_t)i * k + c] : -1;
      bool valid = j >= 0;
      n_nb += __popc(__ballot_sync(FULL, valid));
      consume(j, valid);
    }
    // count row (the table T turns counts into PCL's float sums in the consumer)
    if (lane < 9) {
      unsigned w = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        int bin = lane * 4 + b;
        unsigned c = (bin < 33) ? (unsigned)h[bin] : 0u;
        w |= (c & 255u) << (8 * b);
      }
      reinterpret_cast<unsigned*>(rows8 + (size_t)i * SROW)[lane] = w;
    }
  } else {
    CellBlock blk = stencil_of_point(g, i, lane);
    for (int base = 0; base < blk.total; base += 32) {
      int c = base + lane;
      bool valid = c < blk.total;
      int j = block_candidate(blk, valid ? c : 0);
      if (valid) {
        float4 p = g.pts[j];
        valid = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z) < r2;
      }
      unsigned m = __ballot_sync(FULL, valid);
      if (m == 0) continue;
      n_nb += __popc(m);
      consume(j, valid);
    }
    // PCL: hist_incr = 100 / (n - 1), added once per hit (sequential float sum)
    float incr = (n_nb > 1) ? __fdiv_rn(100.0f, (float)(n_nb - 1)) : 0.f;
    float v0 = 0.f, v1 = 0.f;
    int c0 = h[lane], c1 = (lane == 0) ? h[32] : 0;
    for (int t = 0; t < c0; ++t) v0 = __fadd_rn(v0, incr);
    for (int t = 0; t < c1; ++t) v1 = __fadd_rn(v1, incr);
    spfh[(size_t)i * 33 + lane] = v0;
    if (lane == 0) spfh[(size_t)i * 33 + 32] = v1;
  }
}

// ------------------------------------------------------------------------------ SPFH from kNN rows, k <= 32
// The k-search rows of every surface point (one list entry per lane).  A warp describes SPFH_NP consecutive points
// and keeps two loads ahead of its arithmetic: while point i is binned, the neighbour points / normals of point i + 1
// are in flight and so is the list entry of point i + 2.  (One point per warp left the warp idle through two
// dependent global latencies - list entry, then the gathers - and a block lived for a single point.)
#ifndef PFX_SPFH_NP
#define PFX_SPFH_NP 8
#endif
#ifndef PFX_SPFH_MINB
#define PFX_SPFH_MINB 4
#endif
constexpr int SPFH_NP = PFX_SPFH_NP;
__global__ void __launch_bounds__(FWPB * 32, PFX_SPFH_MINB)
spfh_list32_kernel(GridDev g, const float4* __restrict__ nrm, const int* __restrict__ lists, int k,
                   unsigned char* __restrict__ rows8) {
  __shared__ __align__(16) int hist[FWPB][36];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int n_valid = g.gp->n_valid;
  const int i0 = (blockIdx.x * FWPB + wid) * SPFH_NP;
  if (i0 >= n_valid) return;
  const int iend = min(i0 + SPFH_NP, n_valid);
  int* h = hist[wid];
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  int j_cur = (lane < k) ? lists[(size_t)i0 * k + lane] : -1;
  int j_nxt = (i0 + 1 < iend && lane < k) ? lists[(size_t)(i0 + 1) * k + lane] : -1;
  float4 p_cur = zero4, n_cur = zero4;
  if (j_cur >= 0) {
    p_cur = g.pts[j_cur];
    n_cur = nrm[j_cur];
  }
  float4 q = g.pts[i0], nq = nrm[i0];
  for (int i = i0; i < iend; ++i) {
    // ---- loads of the points ahead
    float4 p_nxt = zero4, n_nxt = zero4, q_nxt = zero4, nq_nxt = zero4;
    if (i + 1 < iend) {
      if (j_nxt >= 0) {
        p_nxt = g.pts[j_nxt];
        n_nxt = nrm[j_nxt];
      }
      q_nxt = g.pts[i + 1];
      nq_nxt = nrm[i + 1];
    }
    const int j_nn = (i + 2 < iend && lane < k) ? lists[(size_t)(i + 2) * k + lane] : -1;
    // ---- point i
    h[lane] = 0;
    if (lane < 4) h[32 + lane] = 0;
    __syncwarp();
    int b1 = -1, b2 = -1, b3 = -1;
    if (j_cur >= 0 && j_cur != i && finite3(nq.x, nq.y, nq.z)) {
      if (!pair_bins_fast(q, nq, p_cur, n_cur, b1, b2, b3)) b1 = b2 = b3 = -1;
    }
    const unsigned m1 = __match_any_sync(FULL, b1);
    const unsigned m2 = __match_any_sync(FULL, b2);
    const unsigned m3 = __match_any_sync(FULL, b3);
    if (b1 >= 0) {  // the three sub-histograms occupy disjoint slots; one leader per distinct bin
      if (lane == __ffs(m1) - 1) h[b1] = __popc(m1);
      if (lane == __ffs(m2) - 1) h[b2] = __popc(m2);
      if (lane == __ffs(m3) - 1) h[b3] = __popc(m3);
    }
    __syncwarp();
    if (lane < 9) {
      const int4 c4 = (lane < 8) ? *reinterpret_cast<const int4*>(h + 4 * lane) : make_int4(h[32], 0, 0, 0);
      const unsigned w = (unsigned)(c4.x & 255) | ((unsigned)(c4.y & 255) << 8) | ((unsigned)(c4.z & 255) << 16) |
                         ((unsigned)(c4.w & 255) << 24);
      reinterpret_cast<unsigned*>(rows8 + (size_t)i * SROW)[lane] = w;
    }
    __syncwarp();
    j_cur = j_nxt;
    j_nxt = j_nn;
    p_cur = p_nxt;
    n_cur = n_nxt;
    q = q_nxt;
    nq = nq_nxt;
  }
}

// mark the union of the queries' neighbourhoods (radius search, keypoint queries)
__global__ void __launch_bounds__(FWPB * 32)
mark_kernel(GridDev g, const float4* __restrict__ queries, int nq, float r2, int* __restrict__ flags) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int qi = blockIdx.x * FWPB + wid;
  if (qi >= nq) return;
  float4 q = queries[qi];
  if (!finite3(q.x, q.y, q.z)) return;
  CellBlock blk = stencil_of_pos(g, q.x, q.y, q.z, lane);
  for (int base = 0; base < blk.total; base += 32) {
    int c = base + lane;
    bool valid = c < blk.total;
    int j = block_candidate(blk, valid ? c : 0);
    if (valid) {
      float4 p = g.pts[j];
      if (dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z) < r2) flags[j] = 1;
    }
  }
}

// per 11-bin block: scale to sum 100 (double sums of the added float values) and store the row
__device__ __forceinline__ void fpfh_finish(float F0, float F1, double S0, double S1, int lane, float* o) {
  int blk_id = (lane < 11) ? 0 : (lane < 22 ? 1 : 2);
  double s_b0 = warp_sum(blk_id == 0 ? S0 : 0.0);
  double s_b1 = warp_sum(blk_id == 1 ? S0 : 0.0);
  double s_b2 = warp_sum(blk_id == 2 ? S0 : 0.0) + __shfl_sync(FULL, S1, 0);
  double sb = (blk_id == 0) ? s_b0 : (blk_id == 1 ? s_b1 : s_b2);
  if (sb != 0.0) sb = 100.0 / sb;
  o[lane] = __fmul_rn(F0, (float)sb);
  if (lane == 0) {
    double s2 = s_b2;
    if (s2 != 0.0) s2 = 100.0 / s2;
    o[32] = __fmul_rn(F1, (float)s2);
  }
}

// ------------------------------------------------------------------------------ FPFH, radius search
// out row = original index of the query (dense) or query number; row stride in floats.  One warp per
// query, lanes are bins, neighbours' 33-float SPFH rows are gathered while the stencil is scanned.
template <bool DENSE>
__global__ void __launch_bounds__(FWPB * 32)
fpfh_kernel(GridDev g, const float4* __restrict__ queries, int nq, float r2, const float* __restrict__ spfh,
            float* __restrict__ out, size_t stride) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int qi = blockIdx.x * FWPB + wid;
  if (qi >= nq) return;
  const int n_valid = g.gp->n_valid;
  float4 q = DENSE ? g.pts[qi] : queries[qi];
  const size_t row = DENSE ? (size_t)__float_as_int(q.w) : (size_t)qi;
  float* o = out + row * stride;
  const float nanv = __int_as_float(0x7fc00000);
  float F0 = 0.f, F1 = 0.f;  // bin `lane`, and bin 32 on lane 0
  double S0 = 0.0, S1 = 0.0;
  int n_nb = 0;
  bool ok = finite3(q.x, q.y, q.z) && (!DENSE || qi < n_valid);
  if (ok) {
    CellBlock blk = DENSE ? stencil_of_point(g, qi, lane) : stencil_of_pos(g, q.x, q.y, q.z, lane);
    for (int base = 0; base < blk.total; base += 32) {
      int c = base + lane;
      bool valid = c < blk.total;
      int j = block_candidate(blk, valid ? c : 0);
      float d2 = 0.f;
      if (valid) {
        float4 p = g.pts[j];
        d2 = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);
        valid = d2 < r2;
      }
      n_nb += __popc(__ballot_sync(FULL, valid));
      bool use = valid && (d2 != 0.f);  // "minus the query point itself": dists == 0 skipped
      float w = use ? __fdiv_rn(1.0f, d2) : 0.f;
      unsigned m = __ballot_sync(FULL, use);
      while (m) {  // neighbours with non-zero weight, in lane order
        int s = __ffs(m) - 1;
        m &= m - 1;
        int js = __shfl_sync(FULL, j, s);
        float ws = __shfl_sync(FULL, w, s);
        const float* rr = spfh + (size_t)js * 33;
        float val = __fmul_rn(rr[lane], ws);
        F0 = __fadd_rn(F0, val);
        S0 += (double)val;
        if (lane == 0) {
          float val1 = __fmul_rn(rr[32], ws);
          F1 = __fadd_rn(F1, val1);
          S1 += (double)val1;
        }
      }
    }
  }
  if (!ok || n_nb == 0) {  // PCL: NaN row, is_dense = false
    o[lane] = nanv;
    if (lane == 0) o[32] = nanv;
    return;
  }
  fpfh_finish(F0, F1, S0, S1, lane, o);
}

// ------------------------------------------------------------------------------ FPFH from kNN rows
// k-search.  A warp describes THREE queries at a time: nine lanes per query, lane c of a group owning 32-bit word c
// of the 36-byte count rows (bins 4c .. 4c+3) for ALL neighbours of its query, so nothing is folded across lanes
// until the three block sums.  The weighted sum is taken in FIXED POINT and is therefore independent of the order
// of the neighbour list (the cell-tile k-search emits unsorted sets whose order depends on the voxel grid: a cloud
// described whole and the same cloud described slab by slab on several GPUs must give the same bits):
//   w_s = 1 / d2_s (PCL's weight), q_s = rint(w_s / max_s w_s * 2^24), W = sum q_s (exact integers),
//   i_s = q_s >> sh (rounded) with sh chosen from W so that sum i_s ~ 2^wbits, acc[bin] = sum_s count_s[bin] * i_s.
// A bin receives at most k - 1 votes per neighbour row, so acc < (k - 1) * 2^wbits < 2^32 with
// wbits = 31 - bits(k - 1).  The counts are weighted as they are: PCL's SPFH value of a count c is the float sum of
// c increments 100 / (k - 1), the same increment for every point of a k-search, and FPFH's rescaling of each 11-bin
// block to 100 divides it out.  Quantisation: 2^-26 of the block's weight per neighbour (k = 32): < 3e-5 in PCL's
// percent units, typically 2e-6 - closer to the exact sum than PCL's own float accumulation.
// reductions over the nine lanes of a group (lane c9 of group grp = lane grp * 9 + c9), every lane of the warp
// taking part in the shuffles; the result lands in all nine lanes
__device__ __forceinline__ unsigned group9_sum(unsigned v, int c9, int grp) {
#pragma unroll
  for (int d = 1; d < 16; d <<= 1) {
    const unsigned t = __shfl_down_sync(FULL, v, d);
    if (c9 + d < 9) v += t;
  }
  return __shfl_sync(FULL, v, grp * 9);
}
__device__ __forceinline__ float group9_max(float v, int c9, int grp) {
#pragma unroll
  for (int d = 1; d < 16; d <<= 1) {
    const float t = __shfl_down_sync(FULL, v, d);
    if (c9 + d < 9) v = fmaxf(v, t);
  }
  return __shfl_sync(FULL, v, grp * 9);
}

#ifndef PFX_FL_NB
#define PFX_FL_NB 2
#endif
#ifndef PFX_FL_MINB
#define PFX_FL_MINB 4
#endif
constexpr int FL_NB = PFX_FL_NB;  // batches of three queries per warp
template <bool DENSE, bool K32>
__global__ void __launch_bounds__(FWPB * 32, PFX_FL_MINB)
fpfh_list_kernel(GridDev g, const float4* __restrict__ queries, int nq, const int* __restrict__ lists,
                 const float* __restrict__ ld2, int k_rt, int wbits, const unsigned char* __restrict__ rows8,
                 float* __restrict__ out, size_t stride) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int grp = lane / 9, c9 = lane - grp * 9;
  const int k = K32 ? 32 : k_rt;
  const int n_valid = g.gp->n_valid;
  const unsigned* rows32 = reinterpret_cast<const unsigned*>(rows8);
  // A warp describes FL_NB batches of three queries.  K32: the list entries of the next batch (and its query points)
  // are loaded while the rows of the current one are gathered, so a batch waits for one global latency, not two.
  const int qwarp = (blockIdx.x * FWPB + wid) * (3 * FL_NB);
  int pj[4] = {-1, -1, -1, -1};
  float pd[4] = {0.f, 0.f, 0.f, 0.f};
  float4 pq = make_float4(0.f, 0.f, 0.f, 0.f);
  auto prefetch = [&](int qn) {  // list entries c9 + 9 t and the point of query qn (K32 only)
    const bool mem = grp < 3 && qn < nq;
    const int qsn = mem ? qn : 0;
    pq = DENSE ? g.pts[qsn] : queries[qsn];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int sidx = c9 + 9 * t;
      pj[t] = -1;
      pd[t] = 0.f;
      if (mem && sidx < 32) {
        pj[t] = lists[(size_t)qsn * 32 + sidx];
        pd[t] = ld2[(size_t)qsn * 32 + sidx];
      }
    }
  };
  if (K32) prefetch(qwarp + grp);
  for (int batch = 0; batch < FL_NB; ++batch) {
  const int qi = qwarp + batch * 3 + grp;
  // every shuffle below names all 32 lanes (a compile-time full mask: no collective-sync sequences); lanes 27..31
  // and the groups past the last query idle through them
  const bool member = grp < 3 && qi < nq;
  if (__all_sync(FULL, !member)) return;
  const int qs = member ? qi : 0;
  const float4 q = K32 ? pq : (DENSE ? g.pts[qs] : queries[qs]);
  const size_t row = DENSE ? (size_t)__float_as_int(q.w) : (size_t)qs;
  float* o = out + row * stride;
  const bool ok = member && finite3(q.x, q.y, q.z) && (!DENSE || qi < n_valid);
  const int* lj = lists + (size_t)qs * k;
  const float* ldd = ld2 + (size_t)qs * k;
  unsigned a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  int n_nb = 0;

  if (K32) {
    // lane c holds neighbours c, c + 9, c + 18, c + 27 of its query
    int myj[4];
    float myw[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      myj[t] = ok ? pj[t] : -1;
      myw[t] = 0.f;
      // "minus the query point itself": dists == 0 skipped.  The weights only enter through their 24-bit quantised
      // ratios to the largest one: the reciprocal unit (1 ulp, deterministic) is as good as the IEEE reciprocal
      if (myj[t] >= 0 && pd[t] != 0.f) myw[t] = __fdividef(1.0f, pd[t]);
      n_nb += (myj[t] >= 0) ? 1 : 0;
    }
    if (batch + 1 < FL_NB) prefetch(qi + 3);
    n_nb = (int)group9_sum((unsigned)n_nb, c9, grp);
    const float wl = fmaxf(fmaxf(myw[0], myw[1]), fmaxf(myw[2], myw[3]));
    const float wmax = group9_max(wl, c9, grp);
    const float inv = (wmax > 0.f) ? __fdividef(16777216.0f, wmax) : 0.f;
    unsigned qv[4], wsum = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      qv[t] = (unsigned)__float2int_rn(myw[t] * inv);
      wsum += qv[t];
    }
    wsum = group9_sum(wsum, c9, grp);
    const int sh = max(0, (32 - __clz((int)wsum)) - wbits);
    const unsigned half = (1u << sh) >> 1;
#pragma unroll
    for (int t = 0; t < 4; ++t) qv[t] = (qv[t] + half) >> sh;
#pragma unroll
    for (int sidx = 0; sidx < 32; ++sidx) {
      const int src = grp * 9 + (sidx % 9), t = sidx / 9;
      const int j = __shfl_sync(FULL, myj[t], src);
      const unsigned wi = __shfl_sync(FULL, qv[t], src);
      const unsigned cw = (wi != 0u && member) ? rows32[(size_t)j * 9 + c9] : 0u;
      a0 += __byte_perm(cw, 0u, 0x4440) * wi;  // one PRMT per zero-extended byte
      a1 += __byte_perm(cw, 0u, 0x4441) * wi;
      a2 += __byte_perm(cw, 0u, 0x4442) * wi;
      a3 += __byte_perm(cw, 0u, 0x4443) * wi;
    }
  } else {
    // any k: every lane of the group walks the whole list (broadcast loads), three passes
    float wmax = 0.f;
    if (ok)
      for (int sidx = 0; sidx < k; ++sidx) {
        const int j = lj[sidx];
        const float d2 = ldd[sidx];
        n_nb += (j >= 0) ? 1 : 0;
        if (j >= 0 && d2 != 0.f) wmax = fmaxf(wmax, __frcp_rn(d2));
      }
    const float inv = (wmax > 0.f) ? __frcp_rn(wmax) * 16777216.0f : 0.f;
    unsigned long long wsum64 = 0;
    if (ok)
      for (int sidx = 0; sidx < k; ++sidx) {
        const int j = lj[sidx];
        const float d2 = ldd[sidx];
        if (j >= 0 && d2 != 0.f) wsum64 += (unsigned)__float2int_rn(__frcp_rn(d2) * inv);
      }
    const int sh = max(0, (64 - __clzll((long long)wsum64)) - wbits);
    const unsigned half = (1u << sh) >> 1;
    if (ok)
      for (int sidx = 0; sidx < k; ++sidx) {
        const int j = lj[sidx];
        const float d2 = ldd[sidx];
        if (j < 0 || d2 == 0.f) continue;
        const unsigned wi = ((unsigned)__float2int_rn(__frcp_rn(d2) * inv) + half) >> sh;
        const unsigned cw = rows32[(size_t)j * 9 + c9];
        a0 += __byte_perm(cw, 0u, 0x4440) * wi;  // one PRMT per zero-extended byte
        a1 += __byte_perm(cw, 0u, 0x4441) * wi;
        a2 += __byte_perm(cw, 0u, 0x4442) * wi;
        a3 += __byte_perm(cw, 0u, 0x4443) * wi;
      }
  }
  const bool nan_row = !ok || n_nb == 0;  // PCL: NaN row, is_dense = false
  // per-block sums (bins 0..10 | 11..21 | 22..32); lane c owns bins 4c .. 4c+3
  const unsigned v[4] = {a0, a1, a2, a3};
  unsigned s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const int bin = c9 * 4 + b;
    if (bin < 11) s0 += v[b];
    else if (bin < 22) s1 += v[b];
    else if (bin < 33) s2 += v[b];
  }
  s0 = group9_sum(s0, c9, grp);
  s1 = group9_sum(s1, c9, grp);
  s2 = group9_sum(s2, c9, grp);
  const float k0 = s0 ? __fdiv_rn(100.0f, (float)s0) : 0.f, k1 = s1 ? __fdiv_rn(100.0f, (float)s1) : 0.f,
              k2 = s2 ? __fdiv_rn(100.0f, (float)s2) : 0.f;
  const float nanv = __int_as_float(0x7fc00000);
  if (member) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int bin = c9 * 4 + b;
      if (bin < 33) o[bin] = nan_row ? nanv : __fmul_rn((float)v[b], bin < 11 ? k0 : (bin < 22 ? k1 : k2));
    }
  }
  }  // batch
}

// k = 32: FOUR queries per warp, EIGHT lanes per query - 32 list entries are exactly four per lane, and no lane idles
// (the nine-lane layout above leaves 5 of 32 lanes and 4 of 36 list slots unused).  Lane c of a group owns words
// 0..7 of the 36-byte count rows (bins 4c .. 4c+3); the ninth word holds bin 32 alone and is read by lane 0 of the
// group with a second load.  Same fixed-point sum as fpfh_list_kernel (integer arithmetic: identical results).
#ifndef PFX_FL8
#define PFX_FL8 1
#endif
__device__ __forceinline__ unsigned group8_sum(unsigned v) {
  v += __shfl_xor_sync(FULL, v, 1);
  v += __shfl_xor_sync(FULL, v, 2);
  v += __shfl_xor_sync(FULL, v, 4);
  return v;
}
__device__ __forceinline__ float group8_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(FULL, v, 1));
  v = fmaxf(v, __shfl_xor_sync(FULL, v, 2));
  v = fmaxf(v, __shfl_xor_sync(FULL, v, 4));
  return v;
}

template <bool DENSE>
__global__ void __launch_bounds__(FWPB * 32, PFX_FL_MINB)
fpfh_list32_kernel(GridDev g, const float4* __restrict__ queries, int nq, const int* __restrict__ lists,
                   const float* __restrict__ ld2, int wbits, const unsigned char* __restrict__ rows8,
                   float* __restrict__ out, size_t stride) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int grp = lane >> 3, c8 = lane & 7;
  const int n_valid = g.gp->n_valid;
  const unsigned* rows32 = reinterpret_cast<const unsigned*>(rows8);
  const int qwarp = (blockIdx.x * FWPB + wid) * (4 * FL_NB);
  int pj[4];
  float pd[4];
  float4 pq;
  auto prefetch = [&](int qn) {  // list entries c8 + 8 t and the point of query qn
    const bool mem = qn < nq;
    const int qsn = mem ? qn : 0;
    pq = DENSE ? g.pts[qsn] : queries[qsn];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      pj[t] = mem ? lists[(size_t)qsn * 32 + c8 + 8 * t] : -1;
      pd[t] = mem ? ld2[(size_t)qsn * 32 + c8 + 8 * t] : 0.f;
    }
  };
  prefetch(qwarp + grp);
  for (int batch = 0; batch < FL_NB; ++batch) {
    const int qi = qwarp + batch * 4 + grp;
    const bool member = qi < nq;
    if (__all_sync(FULL, !member)) return;
    const float4 q = pq;
    const size_t row = DENSE ? (size_t)__float_as_int(q.w) : (size_t)(member ? qi : 0);
    float* o = out + row * stride;
    const bool ok = member && finite3(q.x, q.y, q.z) && (!DENSE || qi < n_valid);
    int myj[4];
    float myw[4];
    int n_nb = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      myj[t] = ok ? pj[t] : -1;
      myw[t] = 0.f;
      // "minus the query point itself": dists == 0 skipped (weights enter through their 24-bit quantised ratios only)
      if (myj[t] >= 0 && pd[t] != 0.f) myw[t] = __fdividef(1.0f, pd[t]);
      n_nb += (myj[t] >= 0) ? 1 : 0;
    }
    if (batch + 1 < FL_NB) prefetch(qi + 4);
    n_nb = (int)group8_sum((unsigned)n_nb);
    const float wmax = group8_max(fmaxf(fmaxf(myw[0], myw[1]), fmaxf(myw[2], myw[3])));
    const float inv = (wmax > 0.f) ? __fdividef(16777216.0f, wmax) : 0.f;
    unsigned qv[4], wsum = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      qv[t] = (unsigned)__float2int_rn(myw[t] * inv);
      wsum += qv[t];
    }
    wsum = group8_sum(wsum);
    const int sh = max(0, (32 - __clz((int)wsum)) - wbits);
    const unsigned half = (1u << sh) >> 1;
#pragma unroll
    for (int t = 0; t < 4; ++t) qv[t] = (qv[t] + half) >> sh;
    unsigned a0 = 0, a1 = 0, a2 = 0, a3 = 0, a32 = 0;
#pragma unroll
    for (int sidx = 0; sidx < 32; ++sidx) {
      const int src = (lane & 24) | (sidx & 7), t = sidx >> 3;
      const int j = __shfl_sync(FULL, myj[t], src);
      const unsigned wi = __shfl_sync(FULL, qv[t], src);
      const bool use = wi != 0u && member;
      const unsigned cw = use ? rows32[(size_t)j * 9 + c8] : 0u;
      a0 += __byte_perm(cw, 0u, 0x4440) * wi;  // one PRMT per zero-extended byte
      a1 += __byte_perm(cw, 0u, 0x4441) * wi;
      a2 += __byte_perm(cw, 0u, 0x4442) * wi;
      a3 += __byte_perm(cw, 0u, 0x4443) * wi;
      if (c8 == 0 && use) a32 += (rows32[(size_t)j * 9 + 8] & 255u) * wi;
    }
    const bool nan_row = !ok || n_nb == 0;  // PCL: NaN row, is_dense = false
    // per-block sums (bins 0..10 | 11..21 | 22..32)
    const unsigned v[4] = {a0, a1, a2, a3};
    unsigned s0 = 0, s1 = 0, s2 = a32;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int bin = c8 * 4 + b;
      if (bin < 11) s0 += v[b];
      else if (bin < 22) s1 += v[b];
      else s2 += v[b];
    }
    s0 = group8_sum(s0);
    s1 = group8_sum(s1);
    s2 = group8_sum(s2);
    const float k0 = s0 ? __fdiv_rn(100.0f, (float)s0) : 0.f, k1 = s1 ? __fdiv_rn(100.0f, (float)s1) : 0.f,
                k2 = s2 ? __fdiv_rn(100.0f, (float)s2) : 0.f;
    const float nanv = __int_as_float(0x7fc00000);
    if (member) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int bin = c8 * 4 + b;
        o[bin] = nan_row ? nanv : __fmul_rn((float)v[b], bin < 11 ? k0 : (bin < 22 ? k1 : k2));
      }
      if (c8 == 0) o[32] = nan_row ? nanv : __fmul_rn((float)a32, k2);
    }
  }
}

__global__ void spfh_export_kernel(GridDev g, const float* __restrict__ spfh, const unsigned char* __restrict__ rows8,
                                   int n, int n_nb, float* __restrict__ out) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)n * 33) return;
  int pos = (int)(t / 33), b = (int)(t % 33);
  int o = __float_as_int(g.pts[pos].w);
  float v = 0.f;
  if (pos < g.gp->n_valid) {
    if (rows8) {
      int nn = min(n_nb, g.gp->n_valid);
      float incr = (nn > 1) ? __fdiv_rn(100.0f, (float)(nn - 1)) : 0.f;
      int c = rows8[(size_t)pos * SROW + b];
      for (int i = 0; i < c; ++i) v = __fadd_rn(v, incr);
    } else {
      v = spfh[t];
    }
  }
  out[(size_t)o * 33 + b] = v;
}

// out_dev: rows of 33 floats at `stride_floats` (caller query order); spfh_out_dev (optional):
// n x 33 SPFH rows in ORIGINAL surface order (stage-wise parity).
int fpfh_compute(Ctx* ctx, Grid* g, double radius, int k, float* out_dev, size_t stride_floats, float* spfh_out_dev) {
  const int n = (int)ctx->n;
  const int nq = (int)ctx->num_queries();
  const bool dense = ctx->q_is_surface;
  const float r2 = (float)(radius * radius);
  const float4* nrm = nullptr;
  PFX_TRY(normals_sorted_for_grid(ctx, g, &nrm));
  if (n == 0) return 0;

  if (k > 0) {
    // ---- k-search: 36-byte count rows for every surface point, then the row gather
    DevBuf& rows8 = ctx->tmp0;
    PFX_CUDA(rows8.ensure((size_t)n * SROW));
    if (!dense) {
      // query rows first (kept in tmp2 / tmp3): the kNN cache is about to hold the dense surface rows
      PFX_TRY(knn_lists(ctx, g, k, false));
      PFX_CUDA(ctx->tmp2.ensure((size_t)std::max(nq, 1) * k * sizeof(int)));
      PFX_CUDA(ctx->tmp3.ensure((size_t)std::max(nq, 1) * k * sizeof(float)));
      if (nq > 0) {
        PFX_CUDA(cudaMemcpyAsync(ctx->tmp2.p, ctx->knn_idx.p, (size_t)nq * k * sizeof(int), cudaMemcpyDeviceToDevice,
                                 ctx->stream));
        PFX_CUDA(cudaMemcpyAsync(ctx->tmp3.p, ctx->knn_d2.p, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToDevice,
                                 ctx->stream));
      }
      ctx->knn_grid = nullptr;
    }
    PFX_TRY(knn_tile_lists(ctx, g, k, false));
    if (k <= 32)
      PFX_LAUNCH(ctx, spfh_list32_kernel, div_up(n, FWPB * SPFH_NP), FWPB * 32, 0, g->view(), nrm, ctx->knn_idx.as<int>(), k,
                 rows8.as<unsigned char>());
    else
      PFX_LAUNCH(ctx, spfh_kernel<true>, div_up(n, FWPB), FWPB * 32, 0, g->view(), nrm, r2, ctx->knn_idx.as<int>(), k,
                 nullptr, nullptr, nullptr, rows8.as<unsigned char>());
    if (spfh_out_dev)
      PFX_LAUNCH(ctx, spfh_export_kernel, div_up((long long)n * 33, 256), 256, 0, g->view(), nullptr,
                 rows8.as<unsigned char>(), n, k, spfh_out_dev);
    if (out_dev && nq > 0) {
      const int blocks = div_up(nq, FWPB * 3 * FL_NB);  // three queries per warp and batch
      int kb = 0;  // bits of k - 1: a bin holds at most k - 1 votes of a neighbour row
      while ((1 << kb) <= std::max(k - 1, 1)) ++kb;
      const int wbits = 31 - kb;
      if (dense) {
        if (k == 32 && PFX_FL8)
          PFX_LAUNCH(ctx, fpfh_list32_kernel<true>, div_up(nq, FWPB * 4 * FL_NB), FWPB * 32, 0, g->view(), nullptr, nq,
                     ctx->knn_idx.as<int>(), ctx->knn_d2.as<float>(), wbits, rows8.as<unsigned char>(), out_dev, stride_floats);
        else if (k == 32)
          PFX_LAUNCH(ctx, (fpfh_list_kernel<true, true>), blocks, FWPB * 32, 0, g->view(), nullptr, nq, ctx->knn_idx.as<int>(),
                     ctx->knn_d2.as<float>(), k, wbits, rows8.as<unsigned char>(), out_dev, stride_floats);
        else
          PFX_LAUNCH(ctx, (fpfh_list_kernel<true, false>), blocks, FWPB * 32, 0, g->view(), nullptr, nq, ctx->knn_idx.as<int>(),
                     ctx->knn_d2.as<float>(), k, wbits, rows8.as<unsigned char>(), out_dev, stride_floats);
      } else {
        PFX_LAUNCH(ctx, (fpfh_list_kernel<false, false>), blocks, FWPB * 32, 0, g->view(), ctx->qry.as<float4>(), nq,
                   ctx->tmp2.as<int>(), ctx->tmp3.as<float>(), k, wbits, rows8.as<unsigned char>(), out_dev, stride_floats);
      }
    }
    PFX_CUDA(cudaGetLastError());
    return 0;
  }

  // ---- radius search: float rows, fused stencil scans
  DevBuf& spfh = ctx->tmp0;
  PFX_CUDA(spfh.ensure((size_t)n * 33 * sizeof(float)));
  int* flags = nullptr;
  if (!dense && !spfh_out_dev) {
    PFX_CUDA(ctx->tmp1.ensure((size_t)n * sizeof(int)));
    flags = ctx->tmp1.as<int>();
    PFX_CUDA(cudaMemsetAsync(flags, 0, (size_t)n * sizeof(int), ctx->stream));
    if (nq > 0)
      PFX_LAUNCH(ctx, mark_kernel, div_up(nq, FWPB), FWPB * 32, 0, g->view(), ctx->qry.as<float4>(), nq, r2, flags);
  }
  PFX_LAUNCH(ctx, spfh_kernel<false>, div_up(n, FWPB), FWPB * 32, 0, g->view(), nrm, r2, nullptr, 0, flags, nullptr,
             spfh.as<float>(), nullptr);
  if (spfh_out_dev)
    PFX_LAUNCH(ctx, spfh_export_kernel, div_up((long long)n * 33, 256), 256, 0, g->view(), spfh.as<float>(), nullptr, n,
               0, spfh_out_dev);
  if (out_dev && nq > 0 && ctx->parity_mode == PFX_PARITY_STRICT) {
    // reference-order weighting (strict.cu): rows come out in caller order, which is the surface's original order
    // for dense queries
    PFX_TRY(fpfh_sorted(ctx, g, radius, spfh.as<float>(), out_dev, stride_floats));
  } else if (out_dev && nq > 0) {
    const int blocks = div_up(nq, FWPB);
    if (dense)
      PFX_LAUNCH(ctx, fpfh_kernel<true>, blocks, FWPB * 32, 0, g->view(), nullptr, nq, r2, spfh.as<float>(), out_dev,
                 stride_floats);
    else
      PFX_LAUNCH(ctx, fpfh_kernel<false>, blocks, FWPB * 32, 0, g->view(), ctx->qry.as<float4>(), nq, r2,
                 spfh.as<float>(), out_dev, stride_floats);
  }
  PFX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pfx
