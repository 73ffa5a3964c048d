// spin.cu — spin images, 9 x 17 = 153 values (replaces
// pcl::SpinImageEstimation<PointXYZRGB, Normal, Histogram<153>>::compute with its defaults - image width 8,
// support angle cosine 0, rectangular image, rotation axis = the query's normal - as driven by the reference at
// evaluation.cpp:515-554: the normals are those of the QUERY cloud, the search surface is the full cloud).
//
// One warp per query, one neighbour per lane from the 3x3x3 stencil of the radius grid: cylindrical coordinates
// (alpha, beta) about the query's normal with upstream's mixed precision (float difference / norm / dot, double
// from there on), the cylinder test, the bilinear vote.  The 153 cells are 64-bit fixed point (2^-40 units) in
// shared memory - order-independent, bit-reproducible - and the image is divided by its sum in double when the
// query has more than one neighbour.
#include "internal.h"

namespace pfx {

constexpr int SPIN_W = 8, SPIN_ROWS = SPIN_W + 1, SPIN_COLS = 2 * SPIN_W + 1, SPIN_LEN = SPIN_ROWS * SPIN_COLS;
constexpr int SPIN_WPB = 4;

template <bool DENSE>
__global__ void __launch_bounds__(SPIN_WPB * 32)
spin_kernel(GridDev g, const float4* __restrict__ queries, int nq, const float* __restrict__ qnormals, size_t nstride,
            float r2, double radius, float* __restrict__ out, size_t stride) {
  __shared__ unsigned long long cells[SPIN_WPB][SPIN_LEN];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int qi = blockIdx.x * SPIN_WPB + wid;
  if (qi >= nq) return;
  unsigned long long* M = cells[wid];
  const GridParams P = *g.gp;
  const float4 q = DENSE ? g.pts[qi] : queries[qi];
  const size_t row = DENSE ? (size_t)__float_as_int(q.w) : (size_t)qi;
  float* o = out + row * stride;
  const float ax = qnormals[row * nstride], ay = qnormals[row * nstride + 1], az = qnormals[row * nstride + 2];
  const bool ok = finite3(q.x, q.y, q.z) && (!DENSE || qi < P.n_valid) && finite3(ax, ay, az);
  if (!ok) {
    for (int b = lane; b < SPIN_LEN; b += 32) o[b] = __int_as_float(0x7fc00000);
    return;
  }
  for (int b = lane; b < SPIN_LEN; b += 32) M[b] = 0ull;
  __syncwarp();
  const double bin = radius / SPIN_W / sqrt(2.0);
  const double lim = bin * SPIN_W;
  const double SCALE = 1099511627776.0;  // 2^40
  int n_nb = 0;
  const CellBlock blk = DENSE ? stencil_of_point(g, qi, lane) : stencil_of_pos(g, q.x, q.y, q.z, lane);
  for (int base = 0; base < blk.total; base += 32) {
    const int t = base + lane;
    const bool valid = t < blk.total;
    const int j = block_candidate(blk, valid ? t : 0);
    bool in = false;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
      p = g.pts[j];
      in = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z) < r2;
    }
    n_nb += __popc(__ballot_sync(FULL, in));
    if (!in) continue;
    const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y), dz = __fsub_rn(p.z, q.z);
    const double dn = (double)__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    if (fabs(dn) < 10 * 2.220446049250313e-16) continue;  // the point itself
    const float dot = __fadd_rn(__fadd_rn(__fmul_rn(dx, ax), __fmul_rn(dy, ay)), __fmul_rn(dz, az));
    double c = (double)dot / dn;
    c = fmax(-1.0, fmin(1.0, c));
    double beta = dn * c;
    double alpha = dn * sqrt(1.0 - c * c);
    if (fabs(beta) >= lim || alpha >= lim) continue;  // outside the cylinder
    int beta_bin = (int)floor(beta / bin) + SPIN_W;
    int alpha_bin = (int)floor(alpha / bin);
    if (alpha_bin == SPIN_W) {
      alpha_bin--;
      alpha = bin * (alpha_bin + 1) - 2.220446049250313e-16;
    }
    if (beta_bin == 2 * SPIN_W) {
      beta_bin--;
      beta = bin * (beta_bin - SPIN_W + 1) - 2.220446049250313e-16;
    }
    const double a = alpha / bin - (double)alpha_bin;
    const double b = beta / bin - (double)(beta_bin - SPIN_W);
    atomicAdd(&M[alpha_bin * SPIN_COLS + beta_bin], __double2ull_rn((1 - a) * (1 - b) * SCALE));
    atomicAdd(&M[(alpha_bin + 1) * SPIN_COLS + beta_bin], __double2ull_rn(a * (1 - b) * SCALE));
    atomicAdd(&M[alpha_bin * SPIN_COLS + beta_bin + 1], __double2ull_rn((1 - a) * b * SCALE));
    atomicAdd(&M[(alpha_bin + 1) * SPIN_COLS + beta_bin + 1], __double2ull_rn(a * b * SCALE));
  }
  __syncwarp();
  unsigned long long part = 0ull;
  for (int b = lane; b < SPIN_LEN; b += 32) part += M[b];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(FULL, part, off);
  const double sum = (double)part;
  const bool norm = n_nb > 1;  // (a sum of 0 with several neighbours gives NaN cells, as upstream's 0 / 0)
  for (int b = lane; b < SPIN_LEN; b += 32) o[b] = (float)(norm ? (double)M[b] / sum : (double)M[b] / SCALE);
}

// qnormals_dev: the normals of the queries (x, y, z first, nstride floats apart); out_dev rows of 153 floats
int spin_compute(Ctx* ctx, Grid* g, double radius, const float* qnormals_dev, size_t nstride_floats, float* out_dev,
                 size_t stride_floats) {
  const int nq = (int)ctx->num_queries();
  if (nq == 0) return 0;
  const float r2 = (float)(radius * radius);
  const int blocks = div_up(nq, SPIN_WPB);
  if (ctx->q_is_surface)
    PFX_LAUNCH(ctx, spin_kernel<true>, blocks, SPIN_WPB * 32, 0, g->view(), nullptr, nq, qnormals_dev, nstride_floats, r2,
               radius, out_dev, stride_floats);
  else
    PFX_LAUNCH(ctx, spin_kernel<false>, blocks, SPIN_WPB * 32, 0, g->view(), ctx->qry.as<float4>(), nq, qnormals_dev,
               nstride_floats, r2, radius, out_dev, stride_floats);
  PFX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pfx
