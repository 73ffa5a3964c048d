// usc.cu — Unique Shape Context, 12 x 11 x 15 = 1980 bins (replaces
// pcl::UniqueShapeContext<PointXYZRGB, ShapeContext1980, ReferenceFrame>::compute as instantiated at reference
// evaluation.cpp:344-371: setMinimalRadius(r / 10), setPointDensityRadius(r / 5), PCL's default local radius 2.5,
// search radius r through features.h:181-195).
//
// Three device stages: (1) the SHOT local reference frame of every query at the local radius (shot.cu);
// (2) the local point density = radius count of EVERY surface point at the density radius, one dense pass
// (search.cu) instead of upstream's radius search per (query, neighbour) pair; (3) usc_kernel, one block per
// query: neighbours from the 3x3x3 stencil of the search-radius grid, polar coordinates in the frame with the CPU's
// float operation order (no FMA), bin = first log-spaced shell / elevation / azimuth division that holds the
// neighbour, weight = 1 / density / cbrt(bin volume) accumulated as 64-bit fixed point (2^-32 units) in shared
// memory - order-independent, hence bit-reproducible - and converted to float once.  The division and volume
// tables are computed on the host with upstream's float / libm expressions and uploaded.
#include <cmath>

#include "internal.h"

namespace pfx {

constexpr int USC_AZ = 12, USC_EL = 11, USC_RB = 15, USC_LEN = USC_AZ * USC_EL * USC_RB;
constexpr int USC_THREADS = 128;

struct UscTab {
  float radii[USC_RB + 1], theta[USC_EL + 1], phi[USC_AZ + 1], vol[USC_LEN];
};

__global__ void __launch_bounds__(USC_THREADS)
usc_kernel(GridDev g, const float4* __restrict__ queries, int nq, int dense, float r2, const float* __restrict__ rf9,
           const int* __restrict__ density_orig, const UscTab* __restrict__ T, float* __restrict__ out, size_t stride) {
  __shared__ unsigned long long bins[USC_LEN];
  __shared__ UscTab tab;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int qi = blockIdx.x;
  if (qi >= nq) return;
  const GridParams P = *g.gp;
  const float4 q = dense ? g.pts[qi] : queries[qi];
  const size_t row = dense ? (size_t)__float_as_int(q.w) : (size_t)qi;
  float* o = out + row * stride;
  for (int b = tid; b < USC_LEN; b += USC_THREADS) bins[b] = 0ull;
  for (int b = tid; b < (int)(sizeof(UscTab) / 4); b += USC_THREADS)
    reinterpret_cast<float*>(&tab)[b] = reinterpret_cast<const float*>(T)[b];
  float rf[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) rf[i] = rf9[row * 9 + i];
  const bool ok = finite3(q.x, q.y, q.z) && (!dense || qi < P.n_valid);
  const bool frame_ok = ok && isfinite(rf[0]) && isfinite(rf[3]) && isfinite(rf[6]);
  if (!frame_ok) {  // upstream: NaN descriptor, frame zeroed
    for (int b = tid; b < USC_LEN; b += USC_THREADS) o[b] = __int_as_float(0x7fc00000);
    if (tid < 9) o[USC_LEN + tid] = 0.f;
    return;
  }
  __syncthreads();
  const float RAD2DEG = 57.29578f;
  const CellBlock blk = dense ? stencil_of_point(g, qi, lane) : stencil_of_pos(g, q.x, q.y, q.z, lane);
  for (int base = wid * 32; base < blk.total; base += USC_THREADS) {
    const int t = base + lane;
    const bool valid = t < blk.total;
    const int jn = block_candidate(blk, valid ? t : 0);
    if (!valid) continue;
    const float4 p = g.pts[jn];
    const float d2 = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);
    if (!(d2 < r2)) continue;
    if (fabsf(d2) <= 1.1920929e-07f) continue;  // pcl::utils::equal(d2, 0): the query itself and near-coincident points
    const float r = __fsqrt_rn(d2);
    const float pox = __fsub_rn(p.x, q.x), poy = __fsub_rn(p.y, q.y), poz = __fsub_rn(p.z, q.z);
    const float lambda = __fadd_rn(__fadd_rn(__fmul_rn(rf[6], pox), __fmul_rn(rf[7], poy)), __fmul_rn(rf[8], poz));
    float prx = __fsub_rn(__fsub_rn(p.x, __fmul_rn(lambda, rf[6])), q.x);
    float pry = __fsub_rn(__fsub_rn(p.y, __fmul_rn(lambda, rf[7])), q.y);
    float prz = __fsub_rn(__fsub_rn(p.z, __fmul_rn(lambda, rf[8])), q.z);
    const float pn = __fadd_rn(__fadd_rn(__fmul_rn(prx, prx), __fmul_rn(pry, pry)), __fmul_rn(prz, prz));
    const float inv = __fdiv_rn(1.0f, __fsqrt_rn(pn));
    prx = __fmul_rn(prx, inv); pry = __fmul_rn(pry, inv); prz = __fmul_rn(prz, inv);
    const float cx = __fsub_rn(__fmul_rn(rf[1], prz), __fmul_rn(rf[2], pry));
    const float cy = __fsub_rn(__fmul_rn(rf[2], prx), __fmul_rn(rf[0], prz));
    const float cz = __fsub_rn(__fmul_rn(rf[0], pry), __fmul_rn(rf[1], prx));
    const float cn = __fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz));
    const float xd = __fadd_rn(__fadd_rn(__fmul_rn(rf[0], prx), __fmul_rn(rf[1], pry)), __fmul_rn(rf[2], prz));
    float phi = __fmul_rn(RAD2DEG, atan2f(__fsqrt_rn(cn), xd));
    const float cdn = __fadd_rn(__fadd_rn(__fmul_rn(cx, rf[6]), __fmul_rn(cy, rf[7])), __fmul_rn(cz, rf[8]));
    phi = cdn < 0.f ? __fsub_rn(360.0f, phi) : phi;
    const float nn = __fadd_rn(__fadd_rn(__fmul_rn(pox, pox), __fmul_rn(poy, poy)), __fmul_rn(poz, poz));
    const float ninv = __fdiv_rn(1.0f, __fsqrt_rn(nn));
    const float nox = __fmul_rn(pox, ninv), noy = __fmul_rn(poy, ninv), noz = __fmul_rn(poz, ninv);
    float th = __fadd_rn(__fadd_rn(__fmul_rn(rf[6], nox), __fmul_rn(rf[7], noy)), __fmul_rn(rf[8], noz));
    th = __fmul_rn(RAD2DEG, acosf(fminf(1.0f, fmaxf(-1.0f, th))));
    int j = 0, k = 0, l = 0;
    for (int rad = 1; rad < USC_RB + 1; ++rad)
      if (r <= tab.radii[rad]) { j = rad - 1; break; }
    for (int ang = 1; ang < USC_EL + 1; ++ang)
      if (th <= tab.theta[ang]) { k = ang - 1; break; }
    for (int ang = 1; ang < USC_AZ + 1; ++ang)
      if (phi <= tab.phi[ang]) { l = ang - 1; break; }
    const int bin = (l * USC_EL * USC_RB) + (k * USC_RB) + j;
    const float dens = (float)density_orig[__float_as_int(p.w)];
    const float w = __fmul_rn(__fdiv_rn(1.0f, dens), tab.vol[bin]);
    if (w > 0.f && w < 4.0e9f) atomicAdd(&bins[bin], __double2ull_rn((double)w * 4294967296.0));
    else if (!(w == 0.f)) atomicAdd(&bins[bin], 0xFFFFFFFFFFFFull << 16);  // inf / NaN weight: saturate (upstream logs an error)
  }
  __syncthreads();
  for (int b = tid; b < USC_LEN; b += USC_THREADS) o[b] = (float)((double)bins[b] * (1.0 / 4294967296.0));
  if (tid < 9) o[USC_LEN + tid] = rf[tid];
}

static void usc_tables(UscTab& T, double min_radius, double search_radius) {
  // UniqueShapeContext::initCompute, same float / double expressions and libm calls
  const float az_int = 360.0f / static_cast<float>(USC_AZ), el_int = 180.0f / static_cast<float>(USC_EL);
  for (int j = 0; j < USC_RB + 1; ++j)
    T.radii[j] = static_cast<float>(std::exp(std::log(min_radius) + ((static_cast<float>(j) / static_cast<float>(USC_RB)) *
                                                                      std::log(search_radius / min_radius))));
  for (int k = 0; k < USC_EL + 1; ++k) T.theta[k] = static_cast<float>(k) * el_int;
  for (int l = 0; l < USC_AZ + 1; ++l) T.phi[l] = static_cast<float>(l) * az_int;
  auto deg2rad = [](float a) { return a * 0.017453293f; };
  const float integr_phi = deg2rad(T.phi[1]) - deg2rad(T.phi[0]);
  const float e = 1.0f / 3.0f;
  for (int j = 0; j < USC_RB; ++j) {
    const float integr_r = (T.radii[j + 1] * T.radii[j + 1] * T.radii[j + 1] / 3) - (T.radii[j] * T.radii[j] * T.radii[j] / 3);
    for (int k = 0; k < USC_EL; ++k) {
      const float integr_theta = cosf(deg2rad(T.theta[k])) - cosf(deg2rad(T.theta[k + 1]));
      const float V = integr_phi * integr_theta * integr_r;
      for (int l = 0; l < USC_AZ; ++l) T.vol[(l * USC_EL * USC_RB) + k * USC_RB + j] = 1.0f / powf(V, e);
    }
  }
}

// ------------------------------------------------------------------------------------------------ 3DSC frames
// ShapeContext3DEstimation::computePoint (reference evaluation.cpp:319-345): z = the normal of the query's nearest
// surface point, x = a random vector made orthogonal to z, y = z cross x.  One warp per query: the nearest point by
// (d2, index) over the 3x3x3 stencil, then lane 0 builds the frame in the CPU's float operation order.  The random
// vector follows the library's seeded contract (upstream's wall-clock-seeded mt19937 cannot be pinned): the three
// draws of query i are the top 24 bits of SplitMix64(seed + golden * (3 i + t + 1)) as floats in [0, 1).
__device__ __forceinline__ uint64_t sc_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void __launch_bounds__(128)
sc3d_frames_kernel(GridDev g, const float4* __restrict__ queries, int nq, int dense, float r2,
                   const float4* __restrict__ nrm_sorted, unsigned long long seed, float* __restrict__ rf9) {
  const int lane = threadIdx.x & 31;
  const int qi = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (qi >= nq) return;
  const float4 q = dense ? g.pts[qi] : queries[qi];
  const int row = dense ? __float_as_int(q.w) : qi;
  const float nanv = __int_as_float(0x7fc00000);
  unsigned long long best = ~0ull;  // (d2 bits << 32) | original index
  int best_j = -1;
  if (finite3(q.x, q.y, q.z) && (!dense || qi < g.gp->n_valid)) {
    const CellBlock blk = dense ? stencil_of_point(g, qi, lane) : stencil_of_pos(g, q.x, q.y, q.z, lane);
    for (int base = 0; base < blk.total; base += 32) {
      const int t = base + lane;
      const bool valid = t < blk.total;
      const int j = block_candidate(blk, valid ? t : 0);
      if (!valid) continue;
      const float4 p = g.pts[j];
      const float d2 = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);
      if (!(d2 < r2)) continue;
      const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)__float_as_int(p.w);
      if (key < best) {
        best = key;
        best_j = j;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long ob = __shfl_xor_sync(FULL, best, o);
    const int oj = __shfl_xor_sync(FULL, best_j, o);
    if (ob < best) {
      best = ob;
      best_j = oj;
    }
  }
  if (lane != 0) return;
  float* rf = rf9 + (size_t)row * 9;
  float4 nz = make_float4(nanv, nanv, nanv, 0.f);
  if (best_j >= 0) nz = nrm_sorted[best_j];
  if (best_j < 0 || !finite3(nz.x, nz.y, nz.z)) {
    for (int b = 0; b < 9; ++b) rf[b] = nanv;
    return;
  }
  float x[3];
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const uint64_t z = sc_splitmix64(seed + 0x9E3779B97F4A7C15ull * (uint64_t)(3 * (uint64_t)row + t + 1));
    x[t] = __fmul_rn((float)(z >> 40), 1.0f / 16777216.0f);
  }
  const float eps = 1.1920929e-07f;
  if (fabsf(nz.z) > eps)
    x[2] = __fdiv_rn(-__fadd_rn(__fmul_rn(nz.x, x[0]), __fmul_rn(nz.y, x[1])), nz.z);
  else if (fabsf(nz.y) > eps)
    x[1] = __fdiv_rn(-__fadd_rn(__fmul_rn(nz.x, x[0]), __fmul_rn(nz.z, x[2])), nz.y);
  else if (fabsf(nz.x) > eps)
    x[0] = __fdiv_rn(-__fadd_rn(__fmul_rn(nz.y, x[1]), __fmul_rn(nz.z, x[2])), nz.x);
  const float xn = __fadd_rn(__fadd_rn(__fmul_rn(x[0], x[0]), __fmul_rn(x[1], x[1])), __fmul_rn(x[2], x[2]));
  const float inv = __fdiv_rn(1.0f, __fsqrt_rn(xn));
  x[0] = __fmul_rn(x[0], inv); x[1] = __fmul_rn(x[1], inv); x[2] = __fmul_rn(x[2], inv);
  rf[0] = x[0]; rf[1] = x[1]; rf[2] = x[2];
  rf[3] = __fsub_rn(__fmul_rn(nz.y, x[2]), __fmul_rn(nz.z, x[1]));
  rf[4] = __fsub_rn(__fmul_rn(nz.z, x[0]), __fmul_rn(nz.x, x[2]));
  rf[5] = __fsub_rn(__fmul_rn(nz.x, x[1]), __fmul_rn(nz.y, x[0]));
  rf[6] = nz.x; rf[7] = nz.y; rf[8] = nz.z;
}

__global__ void sc3d_zero_rf_kernel(float* __restrict__ out, size_t stride, int nq) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nq * 9) out[(size_t)(t / 9) * stride + USC_LEN + (t % 9)] = 0.f;
}

int usc_compute(Ctx* ctx, double search_radius, double min_radius, double density_radius, double local_radius,
                const float* lrf_dev, float* out_dev, size_t stride_floats);

// 3DSC rows (1980 + 9 floats, the frame zeroed as upstream does); frames_out_dev (optional, nq x 9): the frames used
int sc3d_compute(Ctx* ctx, double search_radius, double min_radius, double density_radius, unsigned long long seed,
                 float* out_dev, size_t stride_floats, float* frames_out_dev) {
  const int nq = (int)ctx->num_queries();
  if (nq == 0) return 0;
  Grid* g = nullptr;
  PFX_TRY(grid_get(ctx, search_radius, 0, &g));
  const float4* nrm = nullptr;
  PFX_TRY(normals_sorted_for_grid(ctx, g, &nrm));
  PFX_CUDA(ctx->tmp4.ensure((size_t)nq * 9 * sizeof(float)));
  float* frames = ctx->tmp4.as<float>();
  PFX_LAUNCH(ctx, sc3d_frames_kernel, div_up(nq, 4), 128, 0, g->view(), ctx->q_is_surface ? nullptr : ctx->qry.as<float4>(), nq,
             ctx->q_is_surface ? 1 : 0, (float)(search_radius * search_radius), nrm, seed, frames);
  if (frames_out_dev)
    PFX_CUDA(cudaMemcpyAsync(frames_out_dev, frames, (size_t)nq * 9 * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  PFX_TRY(usc_compute(ctx, search_radius, min_radius, density_radius, 1.0, frames, out_dev, stride_floats));
  PFX_LAUNCH(ctx, sc3d_zero_rf_kernel, div_up(nq * 9, 256), 256, 0, out_dev, stride_floats, nq);
  PFX_CUDA(cudaGetLastError());
  return 0;
}

// out_dev: rows of 1980 + 9 floats at stride_floats, caller query order.  lrf_dev: frames given by the caller (nq x 9)
// or null (SHOT frames at local_radius).
int usc_compute(Ctx* ctx, double search_radius, double min_radius, double density_radius, double local_radius,
                const float* lrf_dev, float* out_dev, size_t stride_floats) {
  const int nq = (int)ctx->num_queries();
  const int n = (int)ctx->n;
  if (nq == 0) return 0;
  // (1) frames
  PFX_CUDA(ctx->tmp2.ensure((size_t)nq * 9 * sizeof(float)));
  float* drf = ctx->tmp2.as<float>();
  if (lrf_dev) {
    PFX_CUDA(cudaMemcpyAsync(drf, lrf_dev, (size_t)nq * 9 * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  } else {
    Grid* gl = nullptr;
    PFX_TRY(grid_get(ctx, local_radius, 0, &gl));
    PFX_TRY(shot_lrf_compute(ctx, gl, local_radius, drf, nullptr));
  }
  // (2) density of every surface point
  PFX_CUDA(ctx->tmp3.ensure((size_t)std::max(n, 1) * sizeof(int)));
  int* dens = ctx->tmp3.as<int>();
  {
    Grid* gd = nullptr;
    PFX_TRY(grid_get(ctx, density_radius, 0, &gd));
    const bool saved = ctx->q_is_surface;
    ctx->q_is_surface = true;
    const int rc = radius_count(ctx, gd, density_radius, dens);
    ctx->q_is_surface = saved;
    if (rc != 0) return rc;
  }
  // (3) tables + descriptor
  UscTab host;
  usc_tables(host, min_radius, search_radius);
  PFX_CUDA(ctx->usc_tab.ensure(sizeof(UscTab)));
  PFX_CUDA(cudaMemcpyAsync(ctx->usc_tab.p, &host, sizeof(UscTab), cudaMemcpyHostToDevice, ctx->stream));
  PFX_CUDA(cudaStreamSynchronize(ctx->stream));  // `host` is a stack object
  Grid* g = nullptr;
  PFX_TRY(grid_get(ctx, search_radius, 0, &g));
  const float r2 = (float)(search_radius * search_radius);
  PFX_LAUNCH(ctx, usc_kernel, nq, USC_THREADS, 0, g->view(), ctx->q_is_surface ? nullptr : ctx->qry.as<float4>(), nq,
             ctx->q_is_surface ? 1 : 0, r2, drf, dens, ctx->usc_tab.as<UscTab>(), out_dev, stride_floats);
  PFX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pfx
