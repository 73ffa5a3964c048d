// fpfh.cu — SPFH + FPFH33 (replaces pcl::FPFHEstimation::compute as instantiated at reference
// evaluation.cpp:597-602 and driven through features.h:181-195).
//
// K10 spfh_kernel: one warp per surface point; lanes take neighbours, compute the Darboux pair
// features and bin them; the 3 x 11 histogram is counted with __match_any_sync (integer hit counts
// in shared memory, no atomics), then each bin replays PCL's `hist += 100/(n-1)` float additions so
// the row is bit-identical to the sequential CPU sum.
// K11 fpfh_kernel: one warp per query; lanes are histogram bins, neighbours' SPFH rows are gathered
// with weight 1/d2, each 11-bin block is rescaled to sum 100.
// Neighbourhoods come either from the cached kNN lists or from a fused radius scan of the stencil.
#include "internal.h"

namespace pfx {

constexpr int FWPB = 8;

__device__ __forceinline__ bool pair_features(float p1x, float p1y, float p1z, float4 n1, float p2x,
                                              float p2y, float p2z, float4 n2, float& f1, float& f2,
                                              float& f3) {
  float dx = p2x - p1x, dy = p2y - p1y, dz = p2z - p1z;
  float f4 = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  if (f4 == 0.0f) return false;
  float a1 = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(n1.x, dx), __fmul_rn(n1.y, dy)), __fmul_rn(n1.z, dz)), f4);
  float a2 = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(n2.x, dx), __fmul_rn(n2.y, dy)), __fmul_rn(n2.z, dz)), f4);
  // upstream: acos(|a1|) > acos(|a2|)  (false when either is NaN, i.e. |a| > 1)
  float b1 = fabsf(a1), b2 = fabsf(a2);
  bool swap = (b1 < b2) && (b2 <= 1.0f);
  float ux, uy, uz, wx_, wy_, wz_;
  if (swap) {
    ux = n2.x; uy = n2.y; uz = n2.z;
    wx_ = n1.x; wy_ = n1.y; wz_ = n1.z;
    dx = -dx; dy = -dy; dz = -dz;
    f3 = -a2;
  } else {
    ux = n1.x; uy = n1.y; uz = n1.z;
    wx_ = n2.x; wy_ = n2.y; wz_ = n2.z;
    f3 = a1;
  }
  // v = d x u
  float vx = __fsub_rn(__fmul_rn(dy, uz), __fmul_rn(dz, uy));
  float vy = __fsub_rn(__fmul_rn(dz, ux), __fmul_rn(dx, uz));
  float vz = __fsub_rn(__fmul_rn(dx, uy), __fmul_rn(dy, ux));
  float vn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz)));
  if (vn == 0.0f) return false;
  vx = __fdiv_rn(vx, vn); vy = __fdiv_rn(vy, vn); vz = __fdiv_rn(vz, vn);
  // w = u x v
  float wx = __fsub_rn(__fmul_rn(uy, vz), __fmul_rn(uz, vy));
  float wy = __fsub_rn(__fmul_rn(uz, vx), __fmul_rn(ux, vz));
  float wz = __fsub_rn(__fmul_rn(ux, vy), __fmul_rn(uy, vx));
  f2 = __fadd_rn(__fadd_rn(__fmul_rn(vx, wx_), __fmul_rn(vy, wy_)), __fmul_rn(vz, wz_));
  float sn = __fadd_rn(__fadd_rn(__fmul_rn(wx, wx_), __fmul_rn(wy, wy_)), __fmul_rn(wz, wz_));
  float cs = __fadd_rn(__fadd_rn(__fmul_rn(ux, wx_), __fmul_rn(uy, wy_)), __fmul_rn(uz, wz_));
  f1 = atan2f(sn, cs);
  return true;
}

__device__ __forceinline__ int clamp_bin(double v) {
  int b = (int)floor(v);
  return min(max(b, 0), 10);
}

// SPFH rows are stored in the sorted order of grid g: spfh[pos * 33 + bin].
// flags (optional): only points with flags[pos] != 0 are computed.
template <bool USE_LIST>
__global__ void __launch_bounds__(FWPB * 32)
spfh_kernel(GridDev g, const float4* __restrict__ nrm, float r2, const int* __restrict__ lists, int k,
            const int* __restrict__ flags, float* __restrict__ spfh) {
  __shared__ int hist[FWPB][36];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int i = blockIdx.x * FWPB + wid;
  const int n_valid = g.gp->n_valid;
  if (i >= n_valid) return;
  if (flags && !flags[i]) return;
  int* h = hist[wid];
  h[lane] = 0;
  if (lane < 4) h[32 + lane] = 0;
  __syncwarp();
  const float4 q = g.pts[i];
  const float4 nq = nrm[i];
  const bool nq_ok = finite3(nq.x, nq.y, nq.z);
  const double d_pi = (double)(1.0f / (2.0f * 3.14159265358979323846f));
  int n_nb = 0;

  auto consume = [&](int j, bool valid) {
    int b1 = -1, b2 = -1, b3 = -1;
    if (valid && j != i && nq_ok) {
      float4 p = g.pts[j];
      float4 nj = nrm[j];
      float f1, f2, f3;
      if (finite3(nj.x, nj.y, nj.z) && pair_features(q.x, q.y, q.z, nq, p.x, p.y, p.z, nj, f1, f2, f3)) {
        b1 = clamp_bin(11 * (((double)f1 + 3.14159265358979323846) * d_pi));
        b2 = 11 + clamp_bin(11 * (((double)f2 + 1.0) * 0.5));
        b3 = 22 + clamp_bin(11 * (((double)f3 + 1.0) * 0.5));
      }
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

// mark the union of the queries' neighbourhoods (sparse mode): flags[sorted pos] = 1
template <bool USE_LIST>
__global__ void __launch_bounds__(FWPB * 32)
mark_kernel(GridDev g, const float4* __restrict__ queries, int nq, float r2, const int* __restrict__ lists,
            int k, int* __restrict__ flags) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int qi = blockIdx.x * FWPB + wid;
  if (qi >= nq) return;
  if (USE_LIST) {
    for (int c = lane; c < k; c += 32) {
      int j = lists[(size_t)qi * k + c];
      if (j >= 0) flags[j] = 1;
    }
  } else {
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
}

// out row = original index of the query (dense) or query number; row stride in floats.
template <bool DENSE, bool USE_LIST>
__global__ void __launch_bounds__(FWPB * 32)
fpfh_kernel(GridDev g, const float4* __restrict__ queries, int nq, float r2, const int* __restrict__ lists,
            const float* __restrict__ ld2, int k, const float* __restrict__ spfh, float* __restrict__ out,
            size_t stride) {
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

  auto gather = [&](int j, float w, unsigned m) {
    // m: lanes that hold a neighbour with non-zero weight, processed in lane order
    while (m) {
      int s = __ffs(m) - 1;
      m &= m - 1;
      int js = __shfl_sync(FULL, j, s);
      float ws = __shfl_sync(FULL, w, s);
      const float* r = spfh + (size_t)js * 33;
      float val = __fmul_rn(r[lane], ws);
      F0 = __fadd_rn(F0, val);
      S0 += (double)val;
      if (lane == 0) {
        float val1 = __fmul_rn(r[32], ws);
        F1 = __fadd_rn(F1, val1);
        S1 += (double)val1;
      }
    }
  };

  if (ok) {
    if (USE_LIST) {
      for (int c0 = 0; c0 < k; c0 += 32) {
        int c = c0 + lane;
        int j = (c < k) ? lists[(size_t)qi * k + c] : -1;
        float d2 = (j >= 0) ? ld2[(size_t)qi * k + c] : 0.f;
        n_nb += __popc(__ballot_sync(FULL, j >= 0));
        bool use = (j >= 0) && (d2 != 0.f);  // "minus the query point itself": dists == 0 skipped
        float w = use ? __fdiv_rn(1.0f, d2) : 0.f;
        gather(j, w, __ballot_sync(FULL, use));
      }
    } else {
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
        bool use = valid && (d2 != 0.f);
        float w = use ? __fdiv_rn(1.0f, d2) : 0.f;
        gather(j, w, __ballot_sync(FULL, use));
      }
    }
  }
  if (!ok || n_nb == 0) {  // PCL: NaN row, is_dense = false
    o[lane] = nanv;
    if (lane == 0) o[32] = nanv;
    return;
  }
  // per 11-bin block: scale to sum 100 (double sum of the added float values)
  // block of lane: 0..10 -> 0, 11..21 -> 1, 22..31 -> 2 (+ bin 32)
  int blk_id = (lane < 11) ? 0 : (lane < 22 ? 1 : 2);
  double s_b0 = warp_sum(blk_id == 0 ? S0 : 0.0);
  double s_b1 = warp_sum(blk_id == 1 ? S0 : 0.0);
  double s_b2 = warp_sum(blk_id == 2 ? S0 : 0.0) + __shfl_sync(FULL, S1, 0);
  double sb = (blk_id == 0) ? s_b0 : (blk_id == 1 ? s_b1 : s_b2);
  if (sb != 0.0) sb = 100.0 / sb;
  float sc = (float)sb;
  o[lane] = __fmul_rn(F0, sc);
  if (lane == 0) {
    double s2 = s_b2;
    if (s2 != 0.0) s2 = 100.0 / s2;
    o[32] = __fmul_rn(F1, (float)s2);
  }
}

// out_dev: rows of 33 floats at `stride_floats` (caller query order); spfh_out_dev (optional):
// n x 33 SPFH rows in ORIGINAL surface order (stage-wise parity).
int fpfh_compute(Ctx* ctx, Grid* g, double radius, int k, float* out_dev, size_t stride_floats,
                 float* spfh_out_dev);

__global__ void spfh_export_kernel(GridDev g, const float* __restrict__ spfh, int n, float* __restrict__ out) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)n * 33) return;
  int pos = (int)(t / 33), b = (int)(t % 33);
  int o = __float_as_int(g.pts[pos].w);
  out[(size_t)o * 33 + b] = (pos < g.gp->n_valid) ? spfh[t] : 0.f;
}

int fpfh_compute(Ctx* ctx, Grid* g, double radius, int k, float* out_dev, size_t stride_floats,
                 float* spfh_out_dev) {
  const int n = (int)ctx->n;
  const int nq = (int)ctx->num_queries();
  const bool dense = ctx->q_is_surface;
  const float r2 = (float)(radius * radius);
  const float4* nrm = nullptr;
  PFX_TRY(normals_sorted_for_grid(ctx, g, &nrm));
  DevBuf& spfh = ctx->tmp0;
  PFX_CUDA(spfh.ensure((size_t)std::max(n, 1) * 33 * sizeof(float)));
  if (n == 0) return 0;

  // SPFH needs the neighbourhood of every involved SURFACE point: kNN lists are dense rows
  const int* dense_lists = nullptr;
  int* flags = nullptr;
  if (!dense && !spfh_out_dev) {
    PFX_CUDA(ctx->tmp1.ensure((size_t)n * sizeof(int)));
    flags = ctx->tmp1.as<int>();
    PFX_CUDA(cudaMemsetAsync(flags, 0, (size_t)n * sizeof(int), ctx->stream));
  }
  if (k > 0) {
    if (!dense) {
      // query lists first (used for marking and for the final gather), kept in tmp2 / tmp3
      PFX_TRY(knn_lists(ctx, g, k, false));
      PFX_CUDA(ctx->tmp2.ensure((size_t)std::max(nq, 1) * k * sizeof(int)));
      PFX_CUDA(ctx->tmp3.ensure((size_t)std::max(nq, 1) * k * sizeof(float)));
      if (nq > 0) {
        PFX_CUDA(cudaMemcpyAsync(ctx->tmp2.p, ctx->knn_idx.p, (size_t)nq * k * sizeof(int),
                                 cudaMemcpyDeviceToDevice, ctx->stream));
        PFX_CUDA(cudaMemcpyAsync(ctx->tmp3.p, ctx->knn_d2.p, (size_t)nq * k * sizeof(float),
                                 cudaMemcpyDeviceToDevice, ctx->stream));
        if (flags)
          PFX_LAUNCH(ctx, mark_kernel<true>, div_up(nq, FWPB), FWPB * 32, 0, g->view(), nullptr, nq, r2,
                     ctx->tmp2.as<int>(), k, flags);
      }
      // now the dense lists of the surface
      bool saved = ctx->q_is_surface;
      ctx->q_is_surface = true;
      int rc = knn_lists(ctx, g, k, false);
      ctx->q_is_surface = saved;
      if (rc) return rc;
      dense_lists = ctx->knn_idx.as<int>();
    } else {
      PFX_TRY(knn_lists(ctx, g, k, false));
      dense_lists = ctx->knn_idx.as<int>();
    }
    PFX_LAUNCH(ctx, spfh_kernel<true>, div_up(n, FWPB), FWPB * 32, 0, g->view(), nrm, r2, dense_lists, k, flags,
               spfh.as<float>());
  } else {
    if (!dense && flags && nq > 0)
      PFX_LAUNCH(ctx, mark_kernel<false>, div_up(nq, FWPB), FWPB * 32, 0, g->view(), ctx->qry.as<float4>(), nq,
                 r2, nullptr, 0, flags);
    PFX_LAUNCH(ctx, spfh_kernel<false>, div_up(n, FWPB), FWPB * 32, 0, g->view(), nrm, r2, nullptr, 0, flags,
               spfh.as<float>());
  }
  if (spfh_out_dev)
    PFX_LAUNCH(ctx, spfh_export_kernel, div_up((long long)n * 33, 256), 256, 0, g->view(), spfh.as<float>(), n,
               spfh_out_dev);
  if (out_dev && nq > 0) {
    const int blocks = div_up(nq, FWPB);
    if (dense) {
      if (k > 0)
        PFX_LAUNCH(ctx, (fpfh_kernel<true, true>), blocks, FWPB * 32, 0, g->view(), nullptr, nq, r2, dense_lists,
                   ctx->knn_d2.as<float>(), k, spfh.as<float>(), out_dev, stride_floats);
      else
        PFX_LAUNCH(ctx, (fpfh_kernel<true, false>), blocks, FWPB * 32, 0, g->view(), nullptr, nq, r2, nullptr,
                   nullptr, 0, spfh.as<float>(), out_dev, stride_floats);
    } else {
      if (k > 0)
        PFX_LAUNCH(ctx, (fpfh_kernel<false, true>), blocks, FWPB * 32, 0, g->view(), ctx->qry.as<float4>(), nq, r2,
                   ctx->tmp2.as<int>(), ctx->tmp3.as<float>(), k, spfh.as<float>(), out_dev, stride_floats);
      else
        PFX_LAUNCH(ctx, (fpfh_kernel<false, false>), blocks, FWPB * 32, 0, g->view(), ctx->qry.as<float4>(), nq,
                   r2, nullptr, nullptr, 0, spfh.as<float>(), out_dev, stride_floats);
    }
  }
  PFX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pfx
