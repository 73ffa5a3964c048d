// common.cuh — shared device/host helpers of the sm_100a feature-extraction kernels.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <string>

namespace pfx {

constexpr unsigned FULL = 0xffffffffu;
constexpr uint32_t KEY_INVALID = 0xffffffffu;  // Morton key of a non-finite point (sorts last)
constexpr int MAX_AXIS_BITS = 10;              // 3 x 10-bit Morton code in a 32-bit key

// ---------------------------------------------------------------- host-side utilities
struct DevBuf {  // grow-only device allocation (no cudaMalloc on the steady-state path)
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

// ---------------------------------------------------------------- grid (voxel hash) description
// Lives in device memory so that the build never has to synchronise with the host.
struct GridParams {
  float ox, oy, oz;   // origin = bbox min of the finite points
  float edge, inv_e;  // cell edge (>= search radius) and its reciprocal
  int nx, ny, nz;     // cells per axis (<= 1024)
  int n_valid;        // finite points (they come first in sorted order)
  int ncells;         // occupied cells
  float mnx, mny, mnz, mxx, mxy, mxz;
};

struct GridDev {
  const GridParams* gp;
  const float4* pts;        // sorted xyz, w = original index bits
  const int* cell_start;    // [ncells + 1]
  const uint32_t* cell_key; // [ncells]
  const uint32_t* hkeys;    // hash table of cell keys (linear probing)
  const int* hvals;         // cell id
  uint32_t hmask;
  const int* pt_cell;       // [n] cell id of each sorted point
  const int* cell_nbr;      // [ncells * 27] cell ids of the 3x3x3 stencil (-1 = empty)
  const int* inv_perm;      // original index -> sorted position
  int n;
};

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ uint32_t part1by2(uint32_t x) {
  x &= 0x3ffu;
  x = (x | (x << 16)) & 0x030000ffu;
  x = (x | (x << 8)) & 0x0300f00fu;
  x = (x | (x << 4)) & 0x030c30c3u;
  x = (x | (x << 2)) & 0x09249249u;
  return x;
}
__device__ __forceinline__ uint32_t morton3(int cx, int cy, int cz) {
  return part1by2((uint32_t)cx) | (part1by2((uint32_t)cy) << 1) | (part1by2((uint32_t)cz) << 2);
}

__device__ __forceinline__ bool finite3(float x, float y, float z) {
  return isfinite(x) && isfinite(y) && isfinite(z);
}

// cell coordinate along one axis: monotone in x, clamped into the grid
__device__ __forceinline__ int cell_coord(float x, float o, float inv_e, int n) {
  float u = __fmul_rn(__fsub_rn(x, o), inv_e);
  int c = (int)floorf(u);
  return min(max(c, 0), n - 1);
}

// FLANN L2_Simple<float> in 3-D: ((dx*dx + dy*dy) + dz*dz), every op rounded, no FMA
// (reference search call sites features.h:192-193, tools.h:29-30).
__device__ __forceinline__ float dist2_flann(float ax, float ay, float az, float bx, float by,
                                             float bz) {
  float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  float s = __fmul_rn(dx, dx);
  s = __fadd_rn(s, __fmul_rn(dy, dy));
  s = __fadd_rn(s, __fmul_rn(dz, dz));
  return s;
}

// atan2 for continuous quantities (interpolation weights, histogram coordinates): octant reduction + the 9-term odd minimax polynomial of Abramowitz &
// Stegun 4.4.49 (|error| <= 2e-8 on [0, 1]); about half the instructions of atan2f
// (measured max abs error 3e-7).  (0, 0) -> 0.
__device__ __forceinline__ float fast_atan2f(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float t = (mx > 0.f) ? __fdividef(mn, mx) : 0.f;
  const float s = t * t;
  float r = 0.0028662257f;
  r = fmaf(r, s, -0.0161657367f);
  r = fmaf(r, s, 0.0429096138f);
  r = fmaf(r, s, -0.0752896400f);
  r = fmaf(r, s, 0.1065626393f);
  r = fmaf(r, s, -0.1420889944f);
  r = fmaf(r, s, 0.1999355085f);
  r = fmaf(r, s, -0.3333314528f);
  r = fmaf(r * s, t, t);
  if (ay > ax) r = 1.57079632679489661923f - r;
  if (x < 0.f) r = 3.14159265358979323846f - r;
  return copysignf(r, y);
}

__device__ __forceinline__ uint32_t hash_key(uint32_t k) { return k * 0x9E3779B1u; }

__device__ __forceinline__ int hash_lookup(const GridDev& g, uint32_t key) {
  uint32_t h = (hash_key(key) >> 7) & g.hmask;
  for (;;) {
    uint32_t k = g.hkeys[h];
    if (k == key) return g.hvals[h];
    if (k == KEY_INVALID) return -1;
    h = (h + 1) & g.hmask;
  }
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(FULL, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// A warp's view of a block of grid cells: lane l owns cell l of the block (start, count) and the
// exclusive prefix of the counts; candidates are addressed by a flat ordinal t in [0, total).
struct CellBlock {
  int start, cnt, prefix, total;
};

__device__ __forceinline__ CellBlock make_block(const GridDev& g, int cell_id, int lane) {
  CellBlock b;
  b.start = 0;
  b.cnt = 0;
  if (cell_id >= 0) {
    b.start = g.cell_start[cell_id];
    b.cnt = g.cell_start[cell_id + 1] - b.start;
  }
  int inc = warp_incl_scan(b.cnt, lane);
  b.prefix = inc - b.cnt;
  b.total = __shfl_sync(FULL, inc, 31);
  return b;
}

// sorted-point index of flat candidate t (t < total); warp-synchronous (all lanes must call)
__device__ __forceinline__ int block_candidate(const CellBlock& b, int t) {
  int lo = 0;  // largest lane with prefix <= t among lanes 0..31 (prefix is non-decreasing)
#pragma unroll
  for (int step = 16; step > 0; step >>= 1) {
    int p = __shfl_sync(FULL, b.prefix, lo + step);
    if (p <= t) lo += step;
  }
  int s = __shfl_sync(FULL, b.start, lo);
  int p = __shfl_sync(FULL, b.prefix, lo);
  return s + (t - p);
}

// 3x3x3 stencil of a surface point (sorted index i): lanes 0..26 read the adjacency table
__device__ __forceinline__ CellBlock stencil_of_point(const GridDev& g, int i, int lane) {
  int cid = g.pt_cell[i];
  int c = (lane < 27) ? g.cell_nbr[cid * 27 + lane] : -1;
  return make_block(g, c, lane);
}

// 3x3x3 stencil of an arbitrary position: lanes 0..26 probe the hash
__device__ __forceinline__ CellBlock stencil_of_pos(const GridDev& g, float x, float y, float z,
                                                    int lane) {
  const GridParams& P = *g.gp;
  int cx = cell_coord(x, P.ox, P.inv_e, P.nx), cy = cell_coord(y, P.oy, P.inv_e, P.ny),
      cz = cell_coord(z, P.oz, P.inv_e, P.nz);
  int c = -1;
  if (lane < 27) {
    int dx = lane % 3 - 1, dy = (lane / 3) % 3 - 1, dz = lane / 9 - 1;
    int x2 = cx + dx, y2 = cy + dy, z2 = cz + dz;
    if (x2 >= 0 && x2 < P.nx && y2 >= 0 && y2 < P.ny && z2 >= 0 && z2 < P.nz)
      c = hash_lookup(g, morton3(x2, y2, z2));
  }
  return make_block(g, c, lane);
}

// ---- searches on a grid whose cells are SMALLER than the radius (a k-search grid serving a radius stage): m rings of
// cells around the query's cell, (2m+1)^3 cells visited in chunks of 32.  `need` = radius (1 + 1e-3), the margin a radius
// grid's own edge carries, so m == 1 exactly when the grid was built for the radius (or a larger one).
__device__ __forceinline__ int stencil_rings(const GridParams& P, float need) {
  int m = max(1, (int)ceilf(need * P.inv_e));
  while ((float)m * P.edge < need) ++m;
  return m;
}
__device__ __forceinline__ int stencil_chunks(int m) {
  return m == 1 ? 1 : ((2 * m + 1) * (2 * m + 1) * (2 * m + 1) + 31) / 32;
}
// chunk ch of the (2m+1)^3 stencil around position (x, y, z): lane l probes cell 32 ch + l
__device__ __forceinline__ CellBlock stencil_chunk(const GridDev& g, float x, float y, float z, int m, int ch, int lane) {
  const GridParams& P = *g.gp;
  const int cx = cell_coord(x, P.ox, P.inv_e, P.nx), cy = cell_coord(y, P.oy, P.inv_e, P.ny),
            cz = cell_coord(z, P.oz, P.inv_e, P.nz);
  const int side = 2 * m + 1, l = ch * 32 + lane;
  int c = -1;
  if (l < side * side * side) {
    const int x2 = cx + l % side - m, y2 = cy + (l / side) % side - m, z2 = cz + l / (side * side) - m;
    if (x2 >= 0 && x2 < P.nx && y2 >= 0 && y2 < P.ny && z2 >= 0 && z2 < P.nz) c = hash_lookup(g, morton3(x2, y2, z2));
  }
  return make_block(g, c, lane);
}
// the stencil of a query as a sequence of CellBlocks: one block (adjacency row or 27 probes) when m == 1
template <bool DENSE>
__device__ __forceinline__ CellBlock stencil_block(const GridDev& g, int qi, float4 q, int m, int ch, int lane) {
  if (m == 1) return DENSE ? stencil_of_point(g, qi, lane) : stencil_of_pos(g, q.x, q.y, q.z, lane);
  return stencil_chunk(g, q.x, q.y, q.z, m, ch, lane);
}

// ---- symmetric 3x3 eigen decomposition by cyclic Jacobi (T = float or double).
// a = (xx, xy, xz, yy, yz, zz).  Eigenvalues ascending in w, eigenvectors in the columns v[r][c].
template <typename T>
__device__ __forceinline__ void jacobi_rot(T& app, T& aqq, T& apq, T& arp, T& arq, T (&v)[3][3],
                                           int p, int q) {
  if (apq == T(0)) return;
  T theta = (aqq - app) / (T(2) * apq);
  T t = (theta >= T(0) ? T(1) : T(-1)) / (fabs(theta) + sqrt(theta * theta + T(1)));
  T c = T(1) / sqrt(t * t + T(1)), s = t * c;
  T tau = s / (T(1) + c);
  T h = t * apq;
  app -= h;
  aqq += h;
  apq = T(0);
  T g1 = arp, g2 = arq;  // the third row/col element pair (r != p, q)
  arp = g1 - s * (g2 + g1 * tau);
  arq = g2 + s * (g1 - g2 * tau);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    T vp = v[r][p], vq = v[r][q];
    v[r][p] = vp - s * (vq + vp * tau);
    v[r][q] = vq + s * (vp - vq * tau);
  }
}

template <typename T>
__device__ __forceinline__ void eig_sym3(const T a_in[6], T w[3], T (&v)[3][3], int sweeps) {
  T xx = a_in[0], xy = a_in[1], xz = a_in[2], yy = a_in[3], yz = a_in[4], zz = a_in[5];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) v[i][j] = (i == j) ? T(1) : T(0);
  for (int s = 0; s < sweeps; ++s) {
    T off = fabs(xy) + fabs(xz) + fabs(yz);
    if (off == T(0)) break;
    // float: rotations by less than 1e-9 of the diagonal change nothing at 24 bits - stop there (the float solves are
    // refined in double by their callers); a surface neighbourhood is there after 3-4 of the 8 sweeps
    if (sizeof(T) == 4 && off <= T(1e-9) * (fabs(xx) + fabs(yy) + fabs(zz))) break;
    jacobi_rot<T>(xx, yy, xy, xz, yz, v, 0, 1);  // (p,q) = (0,1); third index 2: a[2][0], a[2][1]
    jacobi_rot<T>(xx, zz, xz, xy, yz, v, 0, 2);  // (0,2); third index 1: a[1][0], a[1][2]
    jacobi_rot<T>(yy, zz, yz, xy, xz, v, 1, 2);  // (1,2); third index 0: a[0][1], a[0][2]
  }
  w[0] = xx;
  w[1] = yy;
  w[2] = zz;
  // sort ascending (3-element network), permuting columns
#define PFX_SWAPC(i, j)                        \
  if (w[i] > w[j]) {                           \
    T tw = w[i];                               \
    w[i] = w[j];                               \
    w[j] = tw;                                 \
    _Pragma("unroll") for (int r = 0; r < 3; ++r) { \
      T tv = v[r][i];                          \
      v[r][i] = v[r][j];                       \
      v[r][j] = tv;                            \
    }                                          \
  }
  PFX_SWAPC(0, 1)
  PFX_SWAPC(1, 2)
  PFX_SWAPC(0, 1)
#undef PFX_SWAPC
}

}  // namespace pfx
