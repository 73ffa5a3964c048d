// keypoints.cu — cloud resolution, ISS3D and Harris3D keypoints (replaces
// Keypoints::computeCloudResolution keypoints.h:401-428, pcl::ISSKeypoint3D as configured at
// keypoints.h:182-196, pcl::HarrisKeypoint3D as configured at keypoints.h:150-164 and the
// reference's snap Keypoints::getKeypointsCloud keypoints.h:360-395).
//
// All per-point stages are one warp per point over the fused radius scan of the cell stencil; the
// 3x3 eigen problems of ISS are solved 32 at a time (one per lane) in double like the CPU path.
// Non-max suppression is a pure comparison against neighbours' values, so keypoint indices are
// order-free; the output list is compacted in ascending ORIGINAL index.
#include "internal.h"

namespace pfx {

constexpr int KWPB = 4;

// ------------------------------------------------------------------------------ cloud resolution
__global__ void resolution_partial_kernel(GridDev g, const float* __restrict__ d2, int n,
                                          double* __restrict__ psum, int* __restrict__ pcnt) {
  // row i = sorted point i; d2 rows hold (self, nearest other)
  __shared__ double ssum[8];
  __shared__ int scnt[8];
  double s = 0.0;
  int c = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = g.pts[i];
    float dd = d2[2 * (size_t)i + 1];
    if (isfinite(p.x) && i < g.gp->n_valid && isfinite(dd)) {
      s += (double)sqrtf(dd);
      ++c;
    }
  }
  s = warp_sum(s);
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) {
    ssum[threadIdx.x >> 5] = s;
    scnt[threadIdx.x >> 5] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    int tc = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      t += ssum[w];
      tc += scnt[w];
    }
    psum[blockIdx.x] = t;
    pcnt[blockIdx.x] = tc;
  }
}

__global__ void resolution_final_kernel(const double* psum, const int* pcnt, int nb, double* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {  // fixed order: deterministic
    double t = 0;
    long long c = 0;
    for (int i = 0; i < nb; ++i) {
      t += psum[i];
      c += pcnt[i];
    }
    out[0] = (c > 0) ? t / (double)c : 0.0;
  }
}

int cloud_resolution(Ctx* ctx, double* res) {
  const int n = (int)ctx->n;
  *res = 0.0;
  if (n == 0) return 0;
  Grid* g = nullptr;
  PFX_TRY(grid_get(ctx, 0.0, 2, &g));
  PFX_CUDA(ctx->tmp2.ensure((size_t)n * 2 * sizeof(int)));
  PFX_CUDA(ctx->tmp3.ensure((size_t)n * 2 * sizeof(float)));
  PFX_TRY(knn_run(ctx, g, nullptr, n, 2, ctx->tmp2.as<int>(), ctx->tmp3.as<float>()));
  const int nb = 256;
  PFX_CUDA(ctx->small.ensure(nb * (sizeof(double) + sizeof(int)) + 64));
  double* psum = ctx->small.as<double>();
  int* pcnt = reinterpret_cast<int*>(psum + nb + 1);
  PFX_LAUNCH(ctx, resolution_partial_kernel, nb, 256, 0, g->view(), ctx->tmp3.as<float>(), n, psum, pcnt);
  PFX_LAUNCH(ctx, resolution_final_kernel, 1, 32, 0, psum, pcnt, nb, psum + nb);
  PFX_CUDA(cudaMemcpyAsync(res, psum + nb, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ---------------------------------------------------------------------------------------- ISS
// third_eigen_value per ORIGINAL index (0 where the ratio tests fail)
__global__ void __launch_bounds__(KWPB * 32)
iss_saliency_kernel(GridDev g, int n, float r2, int min_nb, double g21, double g32, double* __restrict__ sal) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int qbase = (blockIdx.x * KWPB + wid) * 32;
  if (qbase >= n) return;
  const int n_valid = g.gp->n_valid;
  const int qend = min(32, n - qbase);
  double mine[6] = {0, 0, 0, 0, 0, 0};
  int mycnt = 0;
  for (int t = 0; t < qend; ++t) {
    const int qi = qbase + t;
    float4 q = g.pts[qi];
    double a[6] = {0, 0, 0, 0, 0, 0};
    int cnt = 0;
    if (qi < n_valid) {
      CellBlock blk = stencil_of_point(g, qi, lane);
      for (int base = 0; base < blk.total; base += 32) {
        int c = base + lane;
        bool valid = c < blk.total;
        int j = block_candidate(blk, valid ? c : 0);
        if (valid) {
          float4 p = g.pts[j];
          if (dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z) < r2) {
            double dx = (double)p.x - (double)q.x, dy = (double)p.y - (double)q.y, dz = (double)p.z - (double)q.z;
            a[0] += dx * dx; a[1] += dx * dy; a[2] += dx * dz;
            a[3] += dy * dy; a[4] += dy * dz; a[5] += dz * dz;
            ++cnt;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) a[i] = warp_sum(a[i]);
      cnt = warp_sum(cnt);
    }
    if (lane == t) {
#pragma unroll
      for (int i = 0; i < 6; ++i) mine[i] = a[i];
      mycnt = cnt;
    }
  }
  if (lane < qend) {
    const int qi = qbase + lane;
    double out = 0.0;
    if (qi < n_valid) {
      if (mycnt < min_nb) {
#pragma unroll
        for (int i = 0; i < 6; ++i) mine[i] = 0.0;  // getScatterMatrix returns the zero matrix
      }
      double w[3], v[3][3];
      eig_sym3<double>(mine, w, v, 12);
      double e1 = w[2], e2 = w[1], e3 = w[0];
      if (isfinite(e1) && isfinite(e2) && isfinite(e3) && !(e3 < 0)) {
        if ((e2 / e1 < g21) && (e3 / e2 < g32)) out = e3;
      }
    }
    sal[__float_as_int(g.pts[qi].w)] = out;
  }
}

// generic NMS: value array in ORIGINAL order (double or float); flags[orig] = 1 for keypoints
template <typename T, bool ISS>
__global__ void __launch_bounds__(KWPB * 32)
nms_kernel(GridDev g, int n, float r2, const T* __restrict__ val, int min_nb, float thr, int* __restrict__ flags) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int qi = blockIdx.x * KWPB + wid;
  if (qi >= n) return;
  float4 q = g.pts[qi];
  const int orig = __float_as_int(q.w);
  int res = 0;
  if (qi < g.gp->n_valid) {
    T mv = val[orig];
    bool cand = ISS ? (mv > T(0)) : (isfinite((float)mv) && !((float)mv < thr));
    if (cand) {
      CellBlock blk = stencil_of_point(g, qi, lane);
      int cnt = 0;
      bool beaten = false;
      for (int base = 0; base < blk.total; base += 32) {
        int c = base + lane;
        bool valid = c < blk.total;
        int j = block_candidate(blk, valid ? c : 0);
        if (valid) {
          float4 p = g.pts[j];
          if (dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z) < r2) {
            ++cnt;
            if (mv < val[__float_as_int(p.w)]) beaten = true;
          }
        }
      }
      cnt = warp_sum(cnt);
      beaten = __any_sync(FULL, beaten);
      res = (!beaten && (!ISS || cnt >= min_nb)) ? 1 : 0;
    }
  }
  if (lane == 0) flags[orig] = res;
}

int iss_saliency(Ctx* ctx, Grid* g, double radius, int min_nb, double g21, double g32, double* sal_dev_orig) {
  const int n = (int)ctx->n;
  if (n == 0) return 0;
  PFX_LAUNCH(ctx, iss_saliency_kernel, div_up(n, KWPB * 32), KWPB * 32, 0, g->view(), n, (float)(radius * radius),
             min_nb, g21, g32, sal_dev_orig);
  PFX_CUDA(cudaGetLastError());
  return 0;
}

int iss_nms(Ctx* ctx, Grid* g, const double* sal_dev_orig, double radius, int min_nb, int* flags_dev) {
  const int n = (int)ctx->n;
  if (n == 0) return 0;
  PFX_LAUNCH(ctx, (nms_kernel<double, true>), div_up(n, KWPB), KWPB * 32, 0, g->view(), n, (float)(radius * radius),
             sal_dev_orig, min_nb, 0.f, flags_dev);
  PFX_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------ Harris3D
__global__ void __launch_bounds__(KWPB * 32)
harris_response_kernel(GridDev g, const float4* __restrict__ nrm, int n, float r2, float* __restrict__ resp) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int qi = blockIdx.x * KWPB + wid;
  if (qi >= n) return;
  float4 q = g.pts[qi];
  float r = 0.f;
  if (qi < g.gp->n_valid) {
    CellBlock blk = stencil_of_point(g, qi, lane);
    float xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0;
    int cnt = 0;
    for (int base = 0; base < blk.total; base += 32) {
      int c = base + lane;
      bool valid = c < blk.total;
      int j = block_candidate(blk, valid ? c : 0);
      if (valid) {
        float4 p = g.pts[j];
        if (dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z) < r2) {
          float4 m = nrm[j];
          if (isfinite(m.x)) {
            xx += m.x * m.x; xy += m.x * m.y; xz += m.x * m.z;
            yy += m.y * m.y; yz += m.y * m.z; zz += m.z * m.z;
            ++cnt;
          }
        }
      }
    }
    xx = warp_sum(xx); xy = warp_sum(xy); xz = warp_sum(xz);
    yy = warp_sum(yy); yz = warp_sum(yz); zz = warp_sum(zz);
    cnt = warp_sum(cnt);
    if (cnt > 0) {
      float fc = (float)cnt;
      xx = __fdiv_rn(xx, fc); xy = __fdiv_rn(xy, fc); xz = __fdiv_rn(xz, fc);
      yy = __fdiv_rn(yy, fc); yz = __fdiv_rn(yz, fc); zz = __fdiv_rn(zz, fc);
    }
    float trace = __fadd_rn(__fadd_rn(xx, yy), zz);
    if (trace != 0.f) {
      // det = xx*yy*zz + 2*xy*xz*yz - xz*xz*yy - xy*xy*zz - yz*yz*xx  (left to right, no FMA)
      float det = __fmul_rn(__fmul_rn(xx, yy), zz);
      det = __fadd_rn(det, __fmul_rn(__fmul_rn(__fmul_rn(2.0f, xy), xz), yz));
      det = __fsub_rn(det, __fmul_rn(__fmul_rn(xz, xz), yy));
      det = __fsub_rn(det, __fmul_rn(__fmul_rn(xy, xy), zz));
      det = __fsub_rn(det, __fmul_rn(__fmul_rn(yz, yz), xx));
      r = __fsub_rn(__fadd_rn(0.04f, det), __fmul_rn(__fmul_rn(0.04f, trace), trace));
    }
  }
  if (lane == 0) resp[__float_as_int(q.w)] = r;
}

int harris_response(Ctx* ctx, Grid* g, double radius, float* resp_dev_orig) {
  const int n = (int)ctx->n;
  if (n == 0) return 0;
  const float4* nrm = nullptr;
  PFX_TRY(normals_sorted_for_grid(ctx, g, &nrm));
  PFX_LAUNCH(ctx, harris_response_kernel, div_up(n, KWPB), KWPB * 32, 0, g->view(), nrm, n,
             (float)(radius * radius), resp_dev_orig);
  PFX_CUDA(cudaGetLastError());
  return 0;
}

int harris_nms(Ctx* ctx, Grid* g, const float* resp_dev_orig, double radius, float thr, int* flags_dev) {
  const int n = (int)ctx->n;
  if (n == 0) return 0;
  PFX_LAUNCH(ctx, (nms_kernel<float, false>), div_up(n, KWPB), KWPB * 32, 0, g->view(), n, (float)(radius * radius),
             resp_dev_orig, 0, thr, flags_dev);
  PFX_CUDA(cudaGetLastError());
  return 0;
}

// refineCorners: one warp per corner, <= 10 iterations of corner = (sum n n^T)^-1 (sum n n^T p)
__global__ void __launch_bounds__(KWPB * 32)
harris_refine_kernel(GridDev g, const float4* __restrict__ nrm, float r2, float* __restrict__ corners, int nc) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int ci = blockIdx.x * KWPB + wid;
  if (ci >= nc) return;
  float cx = corners[3 * ci], cy = corners[3 * ci + 1], cz = corners[3 * ci + 2];
  for (int it = 0; it < 10; ++it) {
    float N[6] = {0, 0, 0, 0, 0, 0}, Np[3] = {0, 0, 0};
    CellBlock blk = stencil_of_pos(g, cx, cy, cz, lane);
    for (int base = 0; base < blk.total; base += 32) {
      int c = base + lane;
      bool valid = c < blk.total;
      int j = block_candidate(blk, valid ? c : 0);
      if (valid) {
        float4 p = g.pts[j];
        if (dist2_flann(cx, cy, cz, p.x, p.y, p.z) < r2) {
          float4 m = nrm[j];
          if (isfinite(m.x)) {
            float a = m.x * m.x, b = m.x * m.y, c2 = m.x * m.z, d = m.y * m.y, e = m.y * m.z, f = m.z * m.z;
            N[0] += a; N[1] += b; N[2] += c2; N[3] += d; N[4] += e; N[5] += f;
            Np[0] += a * p.x + b * p.y + c2 * p.z;
            Np[1] += b * p.x + d * p.y + e * p.z;
            Np[2] += c2 * p.x + e * p.y + f * p.z;
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) N[i] = warp_sum(N[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i) Np[i] = warp_sum(Np[i]);
    // invert3x3SymMatrix
    float a = N[0], b = N[1], c = N[2], d = N[3], e = N[4], f = N[5];
    float fd_ee = d * f - e * e, ce_bf = c * e - b * f, be_cd = b * e - c * d;
    float det = a * fd_ee + b * ce_bf + c * be_cd;
    float nx = cx, ny = cy, nz = cz;
    if (det != 0.f) {
      float i00 = fd_ee / det, i01 = ce_bf / det, i02 = be_cd / det;
      float i11 = (a * f - c * c) / det, i12 = (b * c - a * e) / det, i22 = (a * d - b * b) / det;
      nx = i00 * Np[0] + i01 * Np[1] + i02 * Np[2];
      ny = i01 * Np[0] + i11 * Np[1] + i12 * Np[2];
      nz = i02 * Np[0] + i12 * Np[1] + i22 * Np[2];
    }
    float dx = nx - cx, dy = ny - cy, dz = nz - cz;
    float diff = dx * dx + dy * dy + dz * dz;
    cx = nx; cy = ny; cz = nz;
    if (!(diff > 1e-6f)) break;
  }
  if (lane == 0) {
    corners[3 * ci] = cx;
    corners[3 * ci + 1] = cy;
    corners[3 * ci + 2] = cz;
  }
}

int harris_refine(Ctx* ctx, Grid* g, double radius, float* corners_dev, int nc) {
  if (nc == 0) return 0;
  const float4* nrm = nullptr;
  PFX_TRY(normals_sorted_for_grid(ctx, g, &nrm));
  PFX_LAUNCH(ctx, harris_refine_kernel, div_up(nc, KWPB), KWPB * 32, 0, g->view(), nrm, (float)(radius * radius),
             corners_dev, nc);
  PFX_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------ snap + compaction
__global__ void pack_queries_kernel(const float* __restrict__ xyz, int n, float4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], __int_as_float(i));
}

__global__ void snap_final_kernel(GridDev g, const int* __restrict__ idx, const float* __restrict__ d2, int nq,
                                  float max_d2, int* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  int j = idx[i];
  out[i] = (j >= 0 && d2[i] < max_d2) ? __float_as_int(g.pts[j].w) : -1;
}

// keypoints.h:374-394: nearest cloud point of each corner, kept when the squared distance < max_d2
int snap_to_cloud(Ctx* ctx, const float* q_dev, int nq, float max_d2, int* out_dev) {
  if (nq == 0) return 0;
  Grid* g = nullptr;
  PFX_TRY(grid_get(ctx, 0.0, 2, &g));
  PFX_CUDA(ctx->tmp2.ensure((size_t)nq * sizeof(int)));
  PFX_CUDA(ctx->tmp3.ensure((size_t)nq * sizeof(float)));
  PFX_CUDA(ctx->tmp4.ensure((size_t)nq * sizeof(float4)));
  PFX_LAUNCH(ctx, pack_queries_kernel, div_up(nq, 256), 256, 0, q_dev, nq, ctx->tmp4.as<float4>());
  PFX_TRY(knn_run(ctx, g, ctx->tmp4.as<float4>(), nq, 1, ctx->tmp2.as<int>(), ctx->tmp3.as<float>()));
  PFX_LAUNCH(ctx, snap_final_kernel, div_up(nq, 256), 256, 0, g->view(), ctx->tmp2.as<int>(), ctx->tmp3.as<float>(),
             nq, max_d2, out_dev);
  PFX_CUDA(cudaGetLastError());
  return 0;
}

__global__ void compact_kernel(const int* __restrict__ flags, const int* __restrict__ pos, int n,
                               int* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flags[i]) out[pos[i]] = i;
}

// ascending-index list of the set flags; the count comes back to the host (one small D2H)
int compact_flags(Ctx* ctx, const int* flags_dev, int n, int* idx_out_dev, int* count_host) {
  *count_host = 0;
  if (n == 0) return 0;
  PFX_CUDA(ctx->tmp4.ensure((size_t)n * sizeof(int)));
  PFX_CUDA(ctx->small.ensure(256));
  int* total = ctx->small.as<int>() + 16;
  DevBuf& bsum = ctx->scanbuf;
  PFX_TRY(scan_exclusive_i32(ctx, flags_dev, ctx->tmp4.as<int>(), n, total, bsum));
  PFX_LAUNCH(ctx, compact_kernel, div_up(n, 256), 256, 0, flags_dev, ctx->tmp4.as<int>(), n, idx_out_dev);
  PFX_CUDA(cudaMemcpyAsync(count_host, total, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // namespace pfx
