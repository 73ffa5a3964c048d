// shot_fused.cu — SHOT local reference frame + SHOT352 in ONE kernel for neighbourhoods of <= NCAP
// points (replaces SHOTEstimationOMP::compute incl. its internal SHOTLocalReferenceFrameEstimation,
// reference evaluation.cpp:770-775).
//
// A warp owns 32 queries.  Phase A scans the cell stencil of each query ONCE and compacts the
// neighbours that pass d2 < r^2 into a per-query list in shared memory; the (R - d)-weighted scatter
// matrix is accumulated from that dense list (all lanes busy) and parked in lane t.  Phase B solves
// the 32 3x3 eigen problems together (double Jacobi, one per lane).  Phase C, per query: sign votes
// and PCL's median tie fallback from the cached list, then the 352-bin histogram (int32 fixed point
// in shared memory), L2 normalisation and the 1444-byte row.  The generic kernels in shot.cu scan the
// stencil four times per query with ~20 % of the lanes doing the per-neighbour work; this one scans
// once and runs the heavy code on compact lists.  Queries with more than NCAP neighbours are put on
// a work list for the generic kernels.
#include "internal.h"
#include "shot_common.cuh"

namespace pfx {

constexpr int NCAP = 64;
constexpr int FS_WPB = 4;

struct FusedSmem {
  int nbr[32][NCAP];
  int hist[352];
  unsigned long long keys[NCAP];
};

template <bool DENSE>
__global__ void __launch_bounds__(FS_WPB * 32)
shot_fused_kernel(GridDev g, const float4* __restrict__ queries, int nq, const float4* __restrict__ nrm, float r2,
                  double R, float* __restrict__ out, size_t stride, int* __restrict__ wl_count, int* __restrict__ wl) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  FusedSmem* S = reinterpret_cast<FusedSmem*>(smem_raw) + wid;
  const int n_valid = g.gp->n_valid;
  const float nanv = __int_as_float(0x7fc00000);
  const unsigned lt = (1u << lane) - 1u;
  for (int qbase = (blockIdx.x * FS_WPB + wid) * 32; qbase < nq; qbase += gridDim.x * FS_WPB * 32) {
    const int qend = min(32, nq - qbase);
    // ---------------- phase A: neighbour lists + weighted scatter
    double m6[6] = {0, 0, 0, 0, 0, 0};
    double msw = 0.0;
    int mvalid = 0, mn = 0;  // valid (p != q) neighbours, all neighbours; -1: overflow / no query
    for (int t = 0; t < qend; ++t) {
      const int qi = qbase + t;
      const float4 q = DENSE ? g.pts[qi] : queries[qi];
      int n_t = -1;
      double a6[6] = {0, 0, 0, 0, 0, 0};
      double asw = 0.0;
      int avalid = 0;
      if (finite3(q.x, q.y, q.z) && (!DENSE || qi < n_valid)) {
        CellBlock blk = DENSE ? stencil_of_point(g, qi, lane) : stencil_of_pos(g, q.x, q.y, q.z, lane);
        n_t = 0;
        for (int base = 0; base < blk.total; base += 32) {
          int c = base + lane;
          bool valid = c < blk.total;
          int j = block_candidate(blk, valid ? c : 0);
          if (valid) {
            float4 p = g.pts[j];
            valid = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z) < r2;
          }
          unsigned m = __ballot_sync(FULL, valid);
          int pos = n_t + __popc(m & lt);
          if (valid && pos < NCAP) S->nbr[t][pos] = j;
          n_t += __popc(m);
        }
        __syncwarp();
        if (n_t > NCAP) {
          n_t = -2;  // generic kernels
        } else {
          for (int c = lane; c < n_t; c += 32) {
            float4 p = g.pts[S->nbr[t][c]];
            if (!(p.x == q.x && p.y == q.y && p.z == q.z)) {
              float d2 = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);
              double vx = (double)__fsub_rn(p.x, q.x), vy = (double)__fsub_rn(p.y, q.y), vz = (double)__fsub_rn(p.z, q.z);
              double w = R - sqrt((double)d2);
              a6[0] += w * (vx * vx); a6[1] += w * (vx * vy); a6[2] += w * (vx * vz);
              a6[3] += w * (vy * vy); a6[4] += w * (vy * vz); a6[5] += w * (vz * vz);
              asw += w;
              avalid += 1;
            }
          }
#pragma unroll
          for (int i = 0; i < 6; ++i) a6[i] = warp_sum(a6[i]);
          asw = warp_sum(asw);
          avalid = warp_sum(avalid);
        }
      }
      if (lane == t) {
#pragma unroll
        for (int i = 0; i < 6; ++i) m6[i] = a6[i];
        msw = asw;
        mvalid = avalid;
        mn = n_t;
      }
    }
    if (lane >= qend) mn = -1;
    // ---------------- phase B: every lane solves its own query
    double x[3] = {0, 0, 0}, z[3] = {0, 0, 0};
    bool good = false;
    if (mn >= 0 && mvalid >= 5) {
      double a[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) a[i] = m6[i] / msw;
      double w[3], v[3][3];
      eig_sym3<double>(a, w, v, 12);
      good = isfinite(w[0]) && isfinite(w[1]) && isfinite(w[2]);
      x[0] = v[0][2]; x[1] = v[1][2]; x[2] = v[2][2];
      z[0] = v[0][0]; z[1] = v[1][0]; z[2] = v[2][0];
    }
    // ---------------- phase C: votes, frame, histogram, row
    for (int t = 0; t < qend; ++t) {
      const int qi = qbase + t;
      const int n_t = __shfl_sync(FULL, mn, t);
      const float4 q = DENSE ? g.pts[qi] : queries[qi];
      const size_t row = DENSE ? (size_t)__float_as_int(q.w) : (size_t)qi;
      float* o = out + row * stride;
      if (n_t == -2) {  // hand over to the generic kernels
        if (lane == 0) wl[atomicAdd(wl_count, 1)] = qi;
        continue;
      }
      const bool gd = __shfl_sync(FULL, (int)good, t);
      if (n_t <= 0 || !gd) {  // non-finite query, no neighbours, or NaN frame: all-NaN row
        for (int c = lane; c < 361; c += 32) o[c] = nanv;
        continue;
      }
      const int nv = __shfl_sync(FULL, mvalid, t);
      double xs[3], zs[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        xs[i] = __shfl_sync(FULL, x[i], t);
        zs[i] = __shfl_sync(FULL, z[i], t);
      }
      // sign votes over the cached list
      int px = 0, pz = 0;
      for (int c = lane; c < n_t; c += 32) {
        float4 p = g.pts[S->nbr[t][c]];
        if (!(p.x == q.x && p.y == q.y && p.z == q.z)) {
          double vx = (double)__fsub_rn(p.x, q.x), vy = (double)__fsub_rn(p.y, q.y), vz = (double)__fsub_rn(p.z, q.z);
          if (vx * xs[0] + vy * xs[1] + vz * xs[2] >= 0) ++px;
          if (vx * zs[0] + vy * zs[1] + vz * zs[2] >= 0) ++pz;
        }
      }
      px = warp_sum(px);
      pz = warp_sum(pz);
      const int votex = 2 * px - nv, votez = 2 * pz - nv;
      bool fx = votex < 0, fz = votez < 0;
      if (votex == 0 || votez == 0) {
        // PCL's fallback: valid neighbours in (d2, index) order, ranks nv/2-2 .. nv/2+2, strictly positive
        int cntv = 0;
        for (int c0 = 0; c0 < n_t; c0 += 32) {
          int c = c0 + lane;
          bool v = false;
          unsigned long long key = 0;
          if (c < n_t) {
            float4 p = g.pts[S->nbr[t][c]];
            v = !(p.x == q.x && p.y == q.y && p.z == q.z);
            float d2 = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);
            key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)__float_as_int(p.w);
          }
          unsigned m = __ballot_sync(FULL, v);
          if (v) S->keys[cntv + __popc(m & lt)] = key;
          cntv += __popc(m);
        }
        __syncwarp();
        const int med = cntv / 2;
        int plx = 0, plz = 0;
        for (int a = lane; a < cntv; a += 32) {
          unsigned long long ka = S->keys[a];
          int rank = 0;
          for (int b = 0; b < cntv; ++b) rank += (S->keys[b] < ka) ? 1 : 0;
          if (rank >= med - 2 && rank <= med + 2) {
            float4 p = g.pts[g.inv_perm[(int)(unsigned)(ka & 0xffffffffull)]];
            double vx = (double)__fsub_rn(p.x, q.x), vy = (double)__fsub_rn(p.y, q.y), vz = (double)__fsub_rn(p.z, q.z);
            if (vx * xs[0] + vy * xs[1] + vz * xs[2] > 0) ++plx;
            if (vx * zs[0] + vy * zs[1] + vz * zs[2] > 0) ++plz;
          }
        }
        plx = warp_sum(plx);
        plz = warp_sum(plz);
        if (votex == 0) fx = plx < 3;
        if (votez == 0) fz = plz < 3;
        __syncwarp();
      }
      if (fx) { xs[0] = -xs[0]; xs[1] = -xs[1]; xs[2] = -xs[2]; }
      if (fz) { zs[0] = -zs[0]; zs[1] = -zs[1]; zs[2] = -zs[2]; }
      float rf[9];
      lrf_to_float9(xs, zs, rf);
      if (lane < 9) o[352 + lane] = rf[lane];
      if (n_t < 5) {  // computePointSHOT: too few neighbours -> NaN descriptor, frame kept
        for (int c = lane; c < 352; c += 32) o[c] = nanv;
        continue;
      }
      // histogram
      for (int c = lane; c < 352; c += 32) S->hist[c] = 0;
      __syncwarp();
      const float scale = shot_scale(n_t);
      for (int c = lane; c < n_t; c += 32) {
        const int j = S->nbr[t][c];
        const float4 p = g.pts[j];
        const float d2 = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);
        shot_accumulate_neighbor(S->hist, scale, q, p, d2, nrm[j], rf, R);
      }
      __syncwarp();
      const float inv_scale = 1.0f / scale;
      double acc = 0.0;
      float hv[11];
#pragma unroll
      for (int i = 0; i < 11; ++i) {
        hv[i] = (float)S->hist[lane + 32 * i] * inv_scale;
        acc += (double)__fmul_rn(hv[i], hv[i]);
      }
      acc = warp_sum(acc);
      const float nrmv = (float)sqrt(acc);
#pragma unroll
      for (int i = 0; i < 11; ++i) o[lane + 32 * i] = __fdiv_rn(hv[i], nrmv);
      __syncwarp();
    }
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
  PFX_CUDA(ctx->worklist2.ensure(((size_t)nq + 16) * sizeof(int)));
  int* wl_count = ctx->worklist2.as<int>();
  int* wl = wl_count + 16;
  PFX_CUDA(cudaMemsetAsync(wl_count, 0, 16 * sizeof(int), ctx->stream));
  const size_t smem = sizeof(FusedSmem) * FS_WPB;
  static bool attr_set = false;
  if (!attr_set) {
    PFX_CUDA(cudaFuncSetAttribute(shot_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PFX_CUDA(cudaFuncSetAttribute(shot_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
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
