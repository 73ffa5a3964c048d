// icp.cu — point-to-point ICP against the context's cloud (replaces
// pcl::IterativeClosestPoint<PointXYZRGB, PointXYZRGB> as driven by the reference at evaluation.cpp:863-885:
// max correspondence distance 0.07, transformation epsilon 1e-6, euclidean fitness epsilon 1e-4, 100 iterations).
//
// The whole loop stays on the device.  One iteration = four launches on the context's stream:
//   icp_nn_kernel      one warp per source point: nearest target point through the voxel hash, rings of cells
//                      added until the best distance (or the correspondence bound) is inside the scanned cube;
//   icp_accum_kernel   Umeyama moments of the correspondences (count, sum s, sum t, sum t s^T, sum d2) as doubles,
//                      fixed thread -> element mapping and a fixed reduction tree: bit-reproducible run to run;
//   icp_update_kernel  one thread: Kabsch rotation from eigen(H^T H) in double, T and final = T * final in float,
//                      PCL's DefaultConvergenceCriteria (iterations / transform / absolute / relative MSE);
//   icp_move_kernel    cur = T * cur in place (float, ((m0 x + m1 y) + m2 z) + m3, no FMA), like PCL's
//                      transformCloud on input_transformed.
// Every kernel returns at once when the state says "done", so the host enqueues iterations in chunks and reads
// the flag back once per chunk instead of once per iteration.  Fitness = getFitnessScore(): exact 1-NN of every
// source point moved by the final transform (no distance bound), mean of the squared distances.
#include <cmath>
#include <limits>

#include "internal.h"

namespace pfx {

struct IcpState {
  float T[16];    // transform of the last iteration (row-major)
  float fin[16];  // accumulated transform
  double prev_mse;
  double mse;
  long long cnt;
  int iterations, converged, state, done;
};

constexpr int ICP_NMOM = 17;  // cnt, s(3), t(3), t s^T (9), d2
constexpr int ICP_ACC_THREADS = 256;

__device__ __forceinline__ float4 icp_xform(const float* __restrict__ M, float4 p) {
  float4 o;
  o.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(M[0], p.x), __fmul_rn(M[1], p.y)), __fmul_rn(M[2], p.z)), M[3]);
  o.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(M[4], p.x), __fmul_rn(M[5], p.y)), __fmul_rn(M[6], p.z)), M[7]);
  o.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(M[8], p.x), __fmul_rn(M[9], p.y)), __fmul_rn(M[10], p.z)), M[11]);
  o.w = p.w;
  return o;
}

__global__ void icp_seed_kernel(const float* __restrict__ src, size_t stride, int n, const IcpState* __restrict__ st,
                                int apply, float4* __restrict__ cur) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = src + (size_t)i * stride;
  float4 v = make_float4(p[0], p[1], p[2], 0.f);
  if (apply && finite3(v.x, v.y, v.z)) v = icp_xform(st->fin, v);
  cur[i] = v;
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(FULL, v, o);
    v = t < v ? t : v;
  }
  return v;
}

__device__ __forceinline__ unsigned long long icp_scan_block(const GridDev& g, const CellBlock& blk, float4 q,
                                                             unsigned long long best, int lane) {
  for (int base = 0; base < blk.total; base += 32) {
    const int t = base + lane;
    const bool valid = t < blk.total;
    const int j = block_candidate(blk, valid ? t : 0);
    if (valid) {
      const float4 p = g.pts[j];
      const float d2 = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);
      const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)__float_as_int(p.w);
      best = key < best ? key : best;
    }
  }
  return best;
}

// nearest target point of q (ties: lowest original index) among the points with d2 <= bound_d2 (bound_d2 = +inf:
// unbounded).  Returns the packed (d2 bits, original index) key, ~0 when there is none.  Warp-synchronous.
__device__ unsigned long long icp_nearest(const GridDev& g, float4 q, float bound_d2, int lane) {
  const GridParams P = *g.gp;
  const unsigned long long NONE = 0xffffffffffffffffull;
  if (!finite3(q.x, q.y, q.z) || P.n_valid <= 0) return NONE;
  const int cx = cell_coord(q.x, P.ox, P.inv_e, P.nx), cy = cell_coord(q.y, P.oy, P.inv_e, P.ny),
            cz = cell_coord(q.z, P.oz, P.inv_e, P.nz);
  const float ux = __fmul_rn(__fsub_rn(q.x, P.ox), P.inv_e), uy = __fmul_rn(__fsub_rn(q.y, P.oy), P.inv_e),
              uz = __fmul_rn(__fsub_rn(q.z, P.oz), P.inv_e);
  // a bounded search ends after bound / edge rings; only an unbounded one may have to fall back to a full scan
  const int ring_cap = (bound_d2 < CUDART_INF_F) ? 64 : 8;
  unsigned long long best = NONE;
  CellBlock blk = stencil_of_pos(g, q.x, q.y, q.z, lane);
  best = icp_scan_block(g, blk, q, best, lane);
  for (int R = 1;; ++R) {
    best = warp_min_u64(best);
    float safe = CUDART_INF_F;  // distance (in cells) to the nearest face of the scanned cube with cells beyond it
    if (cx - R > 0) safe = fminf(safe, ux - (float)(cx - R));
    if (cx + R < P.nx - 1) safe = fminf(safe, (float)(cx + R + 1) - ux);
    if (cy - R > 0) safe = fminf(safe, uy - (float)(cy - R));
    if (cy + R < P.ny - 1) safe = fminf(safe, (float)(cy + R + 1) - uy);
    if (cz - R > 0) safe = fminf(safe, uz - (float)(cz - R));
    if (cz + R < P.nz - 1) safe = fminf(safe, (float)(cz + R + 1) - uz);
    if (safe == CUDART_INF_F) break;  // whole grid scanned
    safe = (safe - 1e-3f) * P.edge;
    if (safe > 0.f) {
      const float s2 = safe * safe;
      if (best != NONE && __uint_as_float((unsigned)(best >> 32)) < s2) break;
      if (bound_d2 < s2) break;  // nothing beyond the cube can be a correspondence
    }
    const int R2 = R + 1;
    if (R2 > ring_cap) {  // far outside the occupied cells with no bound to stop at: scan everything not yet visited
      for (int base = 0; base < P.n_valid; base += 32) {
        const int j = base + lane;
        if (j < P.n_valid) {
          const float4 p = g.pts[j];
          const int px = cell_coord(p.x, P.ox, P.inv_e, P.nx), py = cell_coord(p.y, P.oy, P.inv_e, P.ny),
                    pz = cell_coord(p.z, P.oz, P.inv_e, P.nz);
          if (max(abs(px - cx), max(abs(py - cy), abs(pz - cz))) > R) {
            const float d2 = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);
            const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)__float_as_int(p.w);
            best = key < best ? key : best;
          }
        }
      }
      best = warp_min_u64(best);
      break;
    }
    const int x0 = max(cx - R2, 0), x1 = min(cx + R2, P.nx - 1);
    const int y0 = max(cy - R2, 0), y1 = min(cy + R2, P.ny - 1);
    const int z0 = max(cz - R2, 0), z1 = min(cz + R2, P.nz - 1);
    const int wx = x1 - x0 + 1, wy = y1 - y0 + 1, wz = z1 - z0 + 1;
    const int ncube = wx * wy * wz;
    for (int cb = 0; cb < ncube; cb += 32) {
      const int t = cb + lane;
      int c = -1;
      if (t < ncube) {
        const int x = x0 + t % wx, y = y0 + (t / wx) % wy, z = z0 + t / (wx * wy);
        if (max(abs(x - cx), max(abs(y - cy), abs(z - cz))) == R2) c = hash_lookup(g, morton3(x, y, z));
      }
      const CellBlock sb = make_block(g, c, lane);
      if (sb.total) best = icp_scan_block(g, sb, q, best, lane);
    }
  }
  if (best != NONE && !(__uint_as_float((unsigned)(best >> 32)) <= bound_d2)) best = NONE;
  return best;
}

// nn_idx: ORIGINAL target index (-1: no correspondence), nn_d2: squared distance
__global__ void __launch_bounds__(256)
icp_nn_kernel(GridDev g, const float4* __restrict__ cur, int n, float bound_d2, double max_d2,
              const IcpState* __restrict__ st, int* __restrict__ nn_idx, float* __restrict__ nn_d2) {
  if (st && st->done) return;
  const int lane = threadIdx.x & 31;
  const int nwarp = (gridDim.x * blockDim.x) >> 5;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nwarp) {
    const unsigned long long key = icp_nearest(g, cur[i], bound_d2, lane);
    if (lane == 0) {
      int j = -1;
      float d2 = CUDART_INF_F;
      if (key != 0xffffffffffffffffull) {
        d2 = __uint_as_float((unsigned)(key >> 32));
        // PCL: "if (distance[0] > max_dist_sqr) continue" with the float distance promoted to double
        if (!((double)d2 > max_d2)) j = (int)(unsigned)(key & 0xffffffffull);
      }
      nn_idx[i] = j;
      nn_d2[i] = d2;
    }
    __syncwarp();
  }
}

// moments of the correspondences; thread t of block b takes elements b * T + t, + gridDim * T, ... and the block
// reduces with a fixed tree, so the partials (and their sum in icp_update_kernel) do not depend on scheduling
__global__ void __launch_bounds__(ICP_ACC_THREADS)
icp_accum_kernel(const float4* __restrict__ cur, int n, const int* __restrict__ nn_idx, const float* __restrict__ nn_d2,
                 const float4* __restrict__ tgt_orig, const IcpState* __restrict__ st, double* __restrict__ partials) {
  if (st && st->done) return;
  __shared__ double sh[ICP_ACC_THREADS / 32][ICP_NMOM];
  double m[ICP_NMOM];
#pragma unroll
  for (int k = 0; k < ICP_NMOM; ++k) m[k] = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int j = nn_idx[i];
    if (j < 0) continue;
    const float4 s = cur[i];
    const float4 t = tgt_orig[j];
    const double sx = s.x, sy = s.y, sz = s.z, tx = t.x, ty = t.y, tz = t.z;
    m[0] += 1.0;
    m[1] += sx; m[2] += sy; m[3] += sz;
    m[4] += tx; m[5] += ty; m[6] += tz;
    m[7] += tx * sx; m[8] += tx * sy; m[9] += tx * sz;
    m[10] += ty * sx; m[11] += ty * sy; m[12] += ty * sz;
    m[13] += tz * sx; m[14] += tz * sy; m[15] += tz * sz;
    m[16] += (double)nn_d2[i];
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < ICP_NMOM; ++k) {
    const double v = warp_sum(m[k]);
    if (lane == 0) sh[wid][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < ICP_NMOM) {
    double v = 0.0;
    for (int w = 0; w < ICP_ACC_THREADS / 32; ++w) v += sh[w][threadIdx.x];
    partials[(size_t)blockIdx.x * ICP_NMOM + threadIdx.x] = v;
  }
}

// rotation R (row-major) of the least-squares rigid transform from the cross-covariance H = sum (s - ms)(t - mt)^T:
// H = U S V^T, R = V diag(1, 1, det) U^T.  V from eigen(H^T H); u0, u1 = H v / sigma; the third pair is completed
// by cross products, which folds a reflection onto the weakest singular direction (Umeyama's S matrix).
__device__ bool icp_kabsch(const double H[3][3], double R[3][3]) {
  double hth[6];
  {
    double M[3][3];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) M[a][b] = H[0][a] * H[0][b] + H[1][a] * H[1][b] + H[2][a] * H[2][b];
    hth[0] = M[0][0]; hth[1] = M[0][1]; hth[2] = M[0][2]; hth[3] = M[1][1]; hth[4] = M[1][2]; hth[5] = M[2][2];
  }
  double w[3], v[3][3];
  eig_sym3<double>(hth, w, v, 60);
  const double v0[3] = {v[0][2], v[1][2], v[2][2]}, v1[3] = {v[0][1], v[1][1], v[2][1]};
  const double s0 = sqrt(fmax(w[2], 0.0)), s1 = sqrt(fmax(w[1], 0.0));
  if (!(s0 > 0) || !(s1 > 1e-14 * s0)) return false;
  double u0[3], u1[3];
  for (int a = 0; a < 3; ++a) {
    u0[a] = (H[a][0] * v0[0] + H[a][1] * v0[1] + H[a][2] * v0[2]) / s0;
    u1[a] = (H[a][0] * v1[0] + H[a][1] * v1[1] + H[a][2] * v1[2]) / s1;
  }
  const double n0 = sqrt(u0[0] * u0[0] + u0[1] * u0[1] + u0[2] * u0[2]);
  for (int a = 0; a < 3; ++a) u0[a] /= n0;
  const double d01 = u0[0] * u1[0] + u0[1] * u1[1] + u0[2] * u1[2];
  for (int a = 0; a < 3; ++a) u1[a] -= d01 * u0[a];
  const double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
  if (!(n1 > 0)) return false;
  for (int a = 0; a < 3; ++a) u1[a] /= n1;
  const double u2[3] = {u0[1] * u1[2] - u0[2] * u1[1], u0[2] * u1[0] - u0[0] * u1[2], u0[0] * u1[1] - u0[1] * u1[0]};
  const double v2[3] = {v0[1] * v1[2] - v0[2] * v1[1], v0[2] * v1[0] - v0[0] * v1[2], v0[0] * v1[1] - v0[1] * v1[0]};
  // columns of U live in source space, columns of V in target space: R = V U^T maps source onto target
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) R[a][b] = v0[a] * u0[b] + v1[a] * u1[b] + v2[a] * u2[b];
  return true;
}

__global__ void icp_update_kernel(const double* __restrict__ partials, int nblocks, int max_iterations, double rot_thr,
                                  double tr_thr, double rel_mse_thr, IcpState* __restrict__ st) {
  if (threadIdx.x != 0 || blockIdx.x != 0 || st->done) return;
  double m[ICP_NMOM];
  for (int k = 0; k < ICP_NMOM; ++k) m[k] = 0.0;
  for (int b = 0; b < nblocks; ++b)
    for (int k = 0; k < ICP_NMOM; ++k) m[k] += partials[(size_t)b * ICP_NMOM + k];
  const long long cnt = (long long)m[0];
  st->cnt = cnt;
  if (cnt < 3) {  // PCL: "Not enough correspondences found" -> NO_CORRESPONDENCES, not converged
    st->state = 5;
    st->converged = 0;
    st->done = 1;
    return;
  }
  const double inv = 1.0 / (double)cnt;
  const double ms[3] = {m[1] * inv, m[2] * inv, m[3] * inv}, mt[3] = {m[4] * inv, m[5] * inv, m[6] * inv};
  double H[3][3];  // sum (s - ms)(t - mt)^T = sum s t^T - cnt ms mt^T; m[7 + 3 b + a] = sum t_b s_a
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) H[a][b] = m[7 + 3 * b + a] - (double)cnt * ms[a] * mt[b];
  double R[3][3];
  if (!icp_kabsch(H, R)) {  // degenerate correspondences (all sources or all targets on one line): identity step
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) R[a][b] = a == b;
  }
  float T[16];
  for (int a = 0; a < 3; ++a) {
    for (int b = 0; b < 3; ++b) T[4 * a + b] = (float)R[a][b];
    T[4 * a + 3] = (float)(mt[a] - (R[a][0] * ms[0] + R[a][1] * ms[1] + R[a][2] * ms[2]));
  }
  T[12] = T[13] = T[14] = 0.f;
  T[15] = 1.f;
  float F[16];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float s = __fmul_rn(T[4 * i], st->fin[j]);
      for (int k = 1; k < 4; ++k) s = __fadd_rn(s, __fmul_rn(T[4 * i + k], st->fin[4 * k + j]));
      F[4 * i + j] = s;
    }
  for (int i = 0; i < 16; ++i) {
    st->T[i] = T[i];
    st->fin[i] = F[i];
  }
  const int it = ++st->iterations;
  const double mse = m[16] * inv;
  st->mse = mse;
  // pcl::registration::DefaultConvergenceCriteria::hasConverged
  int state = 0;
  if (it >= max_iterations) {
    state = 1;
  } else {
    const float tr = __fsub_rn(__fadd_rn(__fadd_rn(T[0], T[5]), T[10]), 1.f);
    const double cos_angle = 0.5 * (double)tr;
    const float t2 = __fadd_rn(__fadd_rn(__fmul_rn(T[3], T[3]), __fmul_rn(T[7], T[7])), __fmul_rn(T[11], T[11]));
    if (cos_angle >= rot_thr && (double)t2 <= tr_thr) state = 2;
    else if (fabs(mse - st->prev_mse) < 1e-12) state = 3;
    else if (fabs(mse - st->prev_mse) / st->prev_mse < rel_mse_thr) state = 4;
    else st->prev_mse = mse;
  }
  st->state = state;
  st->converged = state != 0;
  st->done = state != 0;
}

// cur = T * cur for iteration `iteration` (1-based).  It also runs when that iteration converged (PCL transforms
// before it tests); once the loop has stopped, st->iterations no longer advances and later launches do nothing
__global__ void icp_move_kernel(float4* __restrict__ cur, int n, const IcpState* __restrict__ st, int iteration) {
  if (st->iterations != iteration) return;  // the loop stopped before this iteration (or it found no correspondences)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = cur[i];
  if (finite3(p.x, p.y, p.z)) cur[i] = icp_xform(st->T, p);
}

__global__ void icp_final_kernel(const float* __restrict__ src, size_t stride, int n, const IcpState* __restrict__ st,
                                 float4* __restrict__ moved, float* __restrict__ aligned, size_t astride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = src + (size_t)i * stride;
  float4 v = make_float4(p[0], p[1], p[2], 0.f);
  if (finite3(v.x, v.y, v.z)) v = icp_xform(st->fin, v);
  moved[i] = v;
  if (aligned) {
    aligned[(size_t)i * astride] = v.x;
    aligned[(size_t)i * astride + 1] = v.y;
    aligned[(size_t)i * astride + 2] = v.z;
  }
}

// mean squared NN distance over the finite source points (fixed order: one block)
__global__ void __launch_bounds__(1024)
icp_fitness_kernel(const float4* __restrict__ moved, int n, const int* __restrict__ nn_idx, const float* __restrict__ nn_d2,
                   double* __restrict__ out /* sum, count */) {
  __shared__ double sh[2][32];
  double s = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if (nn_idx[i] >= 0) {
      s += (double)nn_d2[i];
      c += 1.0;
    }
  s = warp_sum(s);
  c = warp_sum(c);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    sh[0][wid] = s;
    sh[1][wid] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0, tc = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      ts += sh[0][w];
      tc += sh[1][w];
    }
    out[0] = ts;
    out[1] = tc;
  }
}

// src_dev: n rows of `stride` floats (xyz first) on the device.  The target is the context's cloud.
int icp_align_run(Ctx* ctx, const float* src_dev, int n, size_t stride, const pfx_icp_params* prm, const float* guess16,
                  pfx_icp_result* res, float* aligned_dev, size_t aligned_stride) {
  if (ctx->surf_version == 0 || ctx->n == 0) return ctx->fail(PFX_E_PRECOND, "pfx_icp_align: no target cloud set");
  if (!(prm->max_correspondence_distance > 0)) return ctx->fail(PFX_E_INVALID, "pfx_icp_align: max_correspondence_distance <= 0");
  const int max_it = prm->max_iterations;
  IcpState h;
  std::memset(&h, 0, sizeof(h));
  bool identity = true;
  for (int i = 0; i < 16; ++i) {
    const float e = (i % 5 == 0) ? 1.f : 0.f;
    h.fin[i] = guess16 ? guess16[i] : e;
    h.T[i] = e;
    identity = identity && h.fin[i] == e;
  }
  h.prev_mse = std::numeric_limits<double>::max();
  res->converged = 0;
  res->iterations = 0;
  res->state = 0;
  res->fitness = std::numeric_limits<double>::max();
  res->correspondences = 0;
  std::memcpy(res->transform, h.fin, sizeof(h.fin));
  if (n <= 0) {
    res->state = 5;
    return 0;
  }
  const int nacc = std::max(1, std::min(ctx->sm_count * 2, div_up(n, ICP_ACC_THREADS)));
  PFX_CUDA(ctx->icp_state.ensure(sizeof(IcpState)));
  PFX_CUDA(ctx->icp_cur.ensure((size_t)n * sizeof(float4)));
  PFX_CUDA(ctx->icp_nn.ensure((size_t)n * (sizeof(int) + sizeof(float))));
  PFX_CUDA(ctx->icp_partials.ensure(((size_t)nacc * ICP_NMOM + 2) * sizeof(double)));
  IcpState* st = ctx->icp_state.as<IcpState>();
  float4* cur = ctx->icp_cur.as<float4>();
  int* nn_idx = ctx->icp_nn.as<int>();
  float* nn_d2 = reinterpret_cast<float*>(nn_idx + n);
  double* partials = ctx->icp_partials.as<double>();
  PFX_CUDA(cudaMemcpyAsync(st, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
  PFX_LAUNCH(ctx, icp_seed_kernel, div_up(n, 256), 256, 0, src_dev, stride, n, st, identity ? 0 : 1, cur);

  // the density-adapted grid of the kNN searches (a few points per cell): the first 3x3x3 stencil answers points
  // that are already close, and max_distance / edge rings prove that a far point has no correspondence
  const double max_d = prm->max_correspondence_distance;
  Grid* g = nullptr;
  int rc = grid_get(ctx, 0.0, 8, &g);
  if (rc != 0) return rc;
  const double max_d2 = max_d * max_d;
  const float bound_d2 = nextafterf((float)max_d2, std::numeric_limits<float>::infinity());  // float bound that keeps every d2 <= max_d2 (double)
  const int nn_blocks = std::max(1, std::min(div_up(n, 8), ctx->sm_count * 8));
  const int CHUNK = 4;
  const int run_it = std::max(max_it, 1);  // PCL's do-while runs one iteration before it looks at the limit
  int it = 0;
  while (it < run_it) {
    const int upto = std::min(run_it, it + CHUNK);
    for (; it < upto; ++it) {
      PFX_LAUNCH(ctx, icp_nn_kernel, nn_blocks, 256, 0, g->view(), cur, n, bound_d2, max_d2, st, nn_idx, nn_d2);
      PFX_LAUNCH(ctx, icp_accum_kernel, nacc, ICP_ACC_THREADS, 0, cur, n, nn_idx, nn_d2, ctx->surf.as<float4>(), st,
                 partials);
      PFX_LAUNCH(ctx, icp_update_kernel, 1, 32, 0, partials, nacc, max_it, 1.0 - prm->transformation_epsilon,
                 prm->transformation_epsilon, prm->euclidean_fitness_epsilon, st);
      PFX_LAUNCH(ctx, icp_move_kernel, div_up(n, 256), 256, 0, cur, n, st, it + 1);
    }
    PFX_CUDA(cudaMemcpyAsync(&h, st, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    PFX_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h.done) break;
  }
  res->converged = h.converged;
  res->iterations = h.iterations;
  res->state = h.state;
  res->correspondences = (int)h.cnt;
  std::memcpy(res->transform, h.fin, sizeof(h.fin));

  // getFitnessScore(): exact, unbounded 1-NN of every moved source point
  PFX_LAUNCH(ctx, icp_final_kernel, div_up(n, 256), 256, 0, src_dev, stride, n, st, cur, aligned_dev, aligned_stride);
  PFX_LAUNCH(ctx, icp_nn_kernel, nn_blocks, 256, 0, g->view(), cur, n, std::numeric_limits<float>::infinity(), std::numeric_limits<double>::infinity(),
             (const IcpState*)nullptr, nn_idx, nn_d2);
  double* fit = partials + (size_t)nacc * ICP_NMOM;
  PFX_LAUNCH(ctx, icp_fitness_kernel, 1, 1024, 0, cur, n, nn_idx, nn_d2, fit);
  double hf[2] = {0, 0};
  PFX_CUDA(cudaMemcpyAsync(hf, fit, sizeof(hf), cudaMemcpyDeviceToHost, ctx->stream));
  PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  if (hf[1] > 0) res->fitness = hf[0] / hf[1];
  return 0;
}

}  // namespace pfx
