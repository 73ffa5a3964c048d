// narf.cu — range image, range-image borders, NARF keypoints and Narf36 descriptors on the GPU
// (replaces, for the reference call sites keypoints.h:204-224, tools.h:65-76 and evaluation.cpp:629-637:
// pcl::RangeImage(Planar)::createFromPointCloud*, pcl::RangeImageBorderExtractor, pcl::NarfKeypoint and
// pcl::NarfDescriptor / pcl::Narf).
//
// The images are small (<= 640 x 480) and every stage is a stencil over the previous stage's image, so the
// work is latency-bound: one kernel per stage, one thread (or one warp) per pixel, nothing to tile.
//   K7  ri_project_kernel      z-buffer with two atomicMin images (direct hits / floor-ceil splats); with
//                              noise_level = 0 the sequential z-buffer of PCL is order-independent
//       ri_finish_kernel       combine, crop (spherical), re-derive xyz from pixel centre + range
//   K8  nb_surface_kernel      local surface: 5x5 neighbours sorted by distance, PCA of the closest ones
//       nb_score_kernel / nb_smooth_kernel / nb_shadow_kernel (2 passes) / nb_classify_kernel /
//       nb_direction_kernel / nb_dir_average_kernel / nb_change_kernel
//       nk_interest_kernel     one WARP per pixel: the region PCL grows pixel by pixel is the 8-connected
//                              component of "eligible" pixels around the seed; it is found by iterated 3x3
//                              dilation of a bit mask held in shared memory, then the 18-bin direction
//                              histogram and the interest value
//       nk_candidates_kernel + nk_select_kernel   3x3 non-maximum suppression, sort by interest, greedy
//                              minimum-distance selection (one block; candidates are few)
//   K9  narf36_kernel          one warp per keypoint: normal-aligned pose, triangle-rasterised 10x10 patch,
//                              20x20 blur, 36 beams, rotation candidates
// Compiled with -fmad=false (see the Makefile): the oracle's float arithmetic has no FMA contraction, and
// the discrete outputs (border traits, keypoint pixels) depend on comparisons of these floats.
#include "internal.h"

namespace pfx {

constexpr float NB_PI = 3.14159265358979323846f;
__device__ __forceinline__ float nb_deg2rad(float d) { return d * (NB_PI / 180.0f); }
__device__ __forceinline__ float nb_norm_angle(float a) {
  if (a >= -NB_PI && a <= NB_PI) return a;
  if (a < -NB_PI) return a + 2 * NB_PI;
  return a - 2 * NB_PI;
}

struct F3 {
  float x, y, z;
};
__device__ __forceinline__ F3 f3(float x, float y, float z) { return F3{x, y, z}; }
__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ F3 operator*(float s, F3 a) { return f3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float fdot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ F3 fcross(F3 a, F3 b) {
  return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float fnorm(F3 a) { return sqrtf(fdot(a, a)); }
__device__ __forceinline__ F3 fnormalized(F3 a) {
  float n = fnorm(a);
  return n > 0 ? (1.0f / n) * a : a;
}

// device view of a range image
struct RiDev {
  int w, h, planar;
  float cx, cy, fx, fy, ares;
  int offx, offy;
  const float4* px;  // x, y, z, range (unobserved: NaN xyz, range -inf), SENSOR frame
  // sensor pose (world <- sensor): rows of R and t; identity unless pfx_range_image_set_pose was called
  int has_pose;
  float R[9], t[3];
  __device__ __forceinline__ F3 to_sensor(F3 p) const {  // R^T (p - t)
    if (!has_pose) return p;
    const F3 d = f3(p.x - t[0], p.y - t[1], p.z - t[2]);
    return f3(R[0] * d.x + R[3] * d.y + R[6] * d.z, R[1] * d.x + R[4] * d.y + R[7] * d.z, R[2] * d.x + R[5] * d.y + R[8] * d.z);
  }
  __device__ __forceinline__ F3 rot_to_world(F3 v) const {  // R v
    if (!has_pose) return v;
    return f3(R[0] * v.x + R[1] * v.y + R[2] * v.z, R[3] * v.x + R[4] * v.y + R[5] * v.z, R[6] * v.x + R[7] * v.y + R[8] * v.z);
  }
  __device__ __forceinline__ F3 to_world(F3 p) const {  // R p + t
    if (!has_pose) return p;
    const F3 r = rot_to_world(p);
    return f3(r.x + t[0], r.y + t[1], r.z + t[2]);
  }
  // the world's y axis (the "up" of pcl::Narf's upright frame) seen from the sensor frame: R^T e_y
  __device__ __forceinline__ F3 world_up() const { return has_pose ? f3(R[3], R[4], R[5]) : f3(0.0f, 1.0f, 0.0f); }
  __device__ __forceinline__ bool in_image(int x, int y) const { return x >= 0 && x < w && y >= 0 && y < h; }
  __device__ __forceinline__ float range(int x, int y) const { return in_image(x, y) ? px[y * w + x].w : -CUDART_INF_F; }
  __device__ __forceinline__ bool valid(int x, int y) const { return in_image(x, y) && isfinite(px[y * w + x].w); }
  __device__ __forceinline__ F3 pt(int x, int y) const {
    float4 p = px[y * w + x];
    return f3(p.x, p.y, p.z);
  }
  __device__ __forceinline__ void project(F3 p, float& ix, float& iy, float& r) const {
    r = fnorm(p);
    if (planar) {
      ix = cx + fx * p.x / p.z;
      iy = cy + fy * p.y / p.z;
    } else {
      float ax = atan2f(p.x, p.z), ay = asinf(p.y / r);
      ix = (ax * cosf(ay) + NB_PI) / ares - (float)offx;
      iy = (ay + 0.5f * NB_PI) / ares - (float)offy;
    }
  }
  __device__ __forceinline__ void project_int(F3 p, int& ix, int& iy, float& r) const {
    float fx_, fy_;
    project(p, fx_, fy_, r);
    ix = (int)lrintf(fx_);
    iy = (int)lrintf(fy_);
  }
  __device__ __forceinline__ F3 point3d(float ix, float iy, float r) const {
    if (planar) {
      float dx = (ix - cx) / fx, dy = (iy - cy) / fy;
      float z = r / sqrtf(dx * dx + dy * dy + 1.0f);
      return f3(dx * z, dy * z, z);
    }
    float ay = (iy + (float)offy) * ares - 0.5f * NB_PI;
    float cay = cosf(ay);
    float ax = cay == 0.0f ? 0.0f : ((ix + (float)offx) * ares - NB_PI) / cay;
    return f3(r * sinf(ax) * cay, r * sinf(ay), r * cosf(ax) * cay);
  }
};

// pcl::VectorAverage<float, 3> as a weighted mean / covariance in double (about a reference point for
// conditioning); doPCA = symmetric eigen decomposition
struct VAcc {
  double sw, s[3], ss[6];
  int n;
  __device__ __forceinline__ void clear() {
    sw = 0;
    n = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) s[i] = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) ss[i] = 0;
  }
  __device__ __forceinline__ void add(F3 p, F3 ref, float w = 1.0f) {
    if (w == 0.0f) return;
    ++n;
    double dw = (double)w, x = (double)p.x - (double)ref.x, y = (double)p.y - (double)ref.y, z = (double)p.z - (double)ref.z;
    sw += dw;
    s[0] += dw * x; s[1] += dw * y; s[2] += dw * z;
    ss[0] += dw * x * x; ss[1] += dw * x * y; ss[2] += dw * x * z;
    ss[3] += dw * y * y; ss[4] += dw * y * z; ss[5] += dw * z * z;
  }
  __device__ __forceinline__ F3 mean(F3 ref) const {
    double inv = 1.0 / sw;
    return f3((float)((double)ref.x + s[0] * inv), (float)((double)ref.y + s[1] * inv), (float)((double)ref.z + s[2] * inv));
  }
  // eigenvalues ascending (ev), e1 = eigenvector of the smallest, e3 of the largest
  __device__ __forceinline__ void pca(float ev[3], F3& e1, F3& e3) const {
    double inv = 1.0 / sw, mx = s[0] * inv, my = s[1] * inv, mz = s[2] * inv;
    double c[6] = {ss[0] * inv - mx * mx, ss[1] * inv - mx * my, ss[2] * inv - mx * mz,
                   ss[3] * inv - my * my, ss[4] * inv - my * mz, ss[5] * inv - mz * mz};
    double w[3], v[3][3];
    eig_sym3<double>(c, w, v, 30);
#pragma unroll
    for (int i = 0; i < 3; ++i) ev[i] = (float)w[i];
    e1 = f3((float)v[0][0], (float)v[1][0], (float)v[2][0]);
    e3 = f3((float)v[0][2], (float)v[1][2], (float)v[2][2]);
  }
};

// ------------------------------------------------------------------------------------ K7 range image
struct RiBuild {
  int top, right, bottom, left;
};

__global__ void ri_init_kernel(unsigned* direct, unsigned* splat, int np, RiBuild* bb, int w, int h) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < np) {
    direct[i] = 0x7f800000u;
    splat[i] = 0x7f800000u;
  }
  if (i == 0) {
    bb->top = h; bb->right = -1; bb->bottom = -1; bb->left = w;
  }
}

__global__ void ri_project_kernel(const float4* __restrict__ pts, int n, RiDev ri, float min_range, unsigned* direct,
                                  unsigned* splat, RiBuild* bb) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pts[i];
  if (!finite3(p.x, p.y, p.z)) return;
  float fx_, fy_, r;
  ri.project(ri.to_sensor(f3(p.x, p.y, p.z)), fx_, fy_, r);  // to_range_image_system = (sensor pose)^-1
  if (!isfinite(fx_) || !isfinite(fy_)) return;
  int x = (int)lrintf(fx_), y = (int)lrintf(fy_);
  if (r < min_range || !ri.in_image(x, y)) return;
  int fxl = (int)lrintf(floorf(fx_)), fyl = (int)lrintf(floorf(fy_)), cxl = (int)lrintf(ceilf(fx_)),
      cyl = (int)lrintf(ceilf(fy_));
  const unsigned rb = __float_as_uint(r);  // r >= 0: float order == unsigned order
  int top = y, bottom = y, left = x, right = x;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int X = (k & 1) ? cxl : fxl, Y = (k & 2) ? cyl : fyl;
    if ((X == x && Y == y) || !ri.in_image(X, Y)) continue;
    atomicMin(&splat[Y * ri.w + X], rb);
    top = min(top, Y); bottom = max(bottom, Y); left = min(left, X); right = max(right, X);
  }
  atomicMin(&direct[y * ri.w + x], rb);
  atomicMin(&bb->top, top);
  atomicMax(&bb->bottom, bottom);
  atomicMin(&bb->left, left);
  atomicMax(&bb->right, right);
}

// dst image (possibly a crop of the projection image: dst(x, y) = src(x + left, y + top)); xyz re-derived
__global__ void ri_finish_kernel(const unsigned* __restrict__ direct, const unsigned* __restrict__ splat, int src_w,
                                 int src_h, int left, int top, RiDev dst, float4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dst.w * dst.h) return;
  int y = i / dst.w, x = i - y * dst.w;
  int sx = x + left, sy = y + top;
  float r = -CUDART_INF_F;
  if (sx >= 0 && sx < src_w && sy >= 0 && sy < src_h) {
    unsigned d = direct[sy * src_w + sx], s = splat[sy * src_w + sx];
    if (d != 0x7f800000u) r = __uint_as_float(d);
    else if (s != 0x7f800000u) r = __uint_as_float(s);
  }
  const float nanv = __int_as_float(0x7fc00000);
  float4 o = make_float4(nanv, nanv, nanv, r);
  if (isfinite(r)) {
    F3 v = dst.point3d((float)x, (float)y, r);
    o.x = v.x; o.y = v.y; o.z = v.z;
  }
  out[i] = o;
}

// ------------------------------------------------------------------------------------ K8 borders
enum {
  T_OBSTACLE = 1 << 0, T_SHADOW = 1 << 1, T_VEIL = 1 << 2,
  T_OBST_TOP = 1 << 4, T_OBST_RIGHT = 1 << 5, T_OBST_BOTTOM = 1 << 6, T_OBST_LEFT = 1 << 7,
  T_SHAD_TOP = 1 << 8, T_SHAD_RIGHT = 1 << 9, T_SHAD_BOTTOM = 1 << 10, T_SHAD_LEFT = 1 << 11,
  T_VEIL_TOP = 1 << 12, T_VEIL_RIGHT = 1 << 13, T_VEIL_BOTTOM = 1 << 14, T_VEIL_LEFT = 1 << 15
};
constexpr int PR_BORDERS = 3, PR_PLANE = 2, PR_DIR = 2, PR_CURV = 2;
constexpr float MIN_PROB = 0.8f;

// per pixel: (normal_no_jumps.xyz, max_neighbor_distance_squared); w < 0 = no local surface
__global__ void nb_surface_kernel(RiDev ri, float4* __restrict__ surf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ri.w * ri.h) return;
  int y = i / ri.w, x = i - y * ri.w;
  float4 res = make_float4(0.f, 0.f, 0.f, -1.f);
  if (ri.valid(x, y)) {
    F3 p = ri.pt(x, y);
    // 5x5 neighbours ordered by (distance^2, linear index): insertion sort of <= 25 keys
    unsigned long long key[25];
    int cnt = 0;
    for (int y2 = y - PR_PLANE; y2 <= y + PR_PLANE; ++y2)
      for (int x2 = x - PR_PLANE; x2 <= x + PR_PLANE; ++x2) {
        if (!ri.valid(x2, y2)) continue;
        F3 d = ri.pt(x2, y2) - p;
        unsigned long long k = ((unsigned long long)__float_as_uint(fdot(d, d)) << 32) | (unsigned)(y2 * ri.w + x2);
        int j = cnt++;
        while (j > 0 && key[j - 1] > k) {
          key[j] = key[j - 1];
          --j;
        }
        key[j] = k;
      }
    if (cnt >= 3) {
      const int n_closest = min(cnt, (PR_PLANE + 1) * (PR_PLANE + 1));
      const float max_nd2 = __uint_as_float((unsigned)(key[n_closest - 1] >> 32));
      const float max_d2 = max_nd2 * 4.0f;
      VAcc va;
      va.clear();
      for (int j = 0; j < cnt; ++j) {
        if (__uint_as_float((unsigned)(key[j] >> 32)) > max_d2) break;
        int li = (int)(unsigned)(key[j] & 0xffffffffull);
        va.add(ri.pt(li % ri.w, li / ri.w), p);
      }
      if (va.n >= 3) {
        float ev[3];
        F3 nrm, e3;
        va.pca(ev, nrm, e3);
        F3 view = fnormalized(f3(0.f, 0.f, 0.f) - p);
        if (fdot(nrm, view) < 0) nrm = -1.0f * nrm;
        res = make_float4(nrm.x, nrm.y, nrm.z, max_nd2);
      }
    }
  }
  surf[i] = res;
}

// RangeImage::get1dPointAverage
__device__ __forceinline__ float4 nb_point_average_1d(const RiDev& ri, int x, int y, int dx, int dy, int npts) {
  const float nanv = __int_as_float(0x7fc00000);
  float wsum = 1.0f;
  float4 out = ri.in_image(x, y) ? ri.px[y * ri.w + x] : make_float4(nanv, nanv, nanv, -CUDART_INF_F);
  const float r0 = out.w;
  if (isinf(r0)) {
    if (r0 > 0) return out;
    wsum = 0;
    out = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  int x2 = x, y2 = y;
  for (int step = 1; step < npts; ++step) {
    x2 += dx; y2 += dy;
    if (!ri.valid(x2, y2)) continue;
    float4 p = ri.px[y2 * ri.w + x2];
    out.x += p.x; out.y += p.y; out.z += p.z; out.w += p.w;
    wsum += 1.0f;
  }
  if (wsum <= 0) return make_float4(nanv, nanv, nanv, -CUDART_INF_F);
  float f = 1.0f / wsum;
  out.x *= f; out.y *= f; out.z *= f; out.w *= f;
  return out;
}

__device__ __forceinline__ float nb_change_score(const RiDev& ri, float max_nd2, int x, int y, int ox, int oy, int pr) {
  float4 p = ri.px[y * ri.w + x];
  float4 nb = nb_point_average_1d(ri, x + ox, y + oy, ox, oy, pr);
  if (isinf(nb.w)) return nb.w < 0 ? 0.0f : 1.0f;
  float dx = nb.x - p.x, dy = nb.y - p.y, dz = nb.z - p.z;
  float d2 = dx * dx + dy * dy + dz * dz;
  if (d2 <= max_nd2) return 0.0f;
  float ret = 1.0f - sqrtf(max_nd2 / d2);
  if (nb.w < p.w) ret = -ret;
  return ret;
}

__constant__ int c_ox[4] = {-1, 1, 0, 0};
__constant__ int c_oy[4] = {0, 0, -1, 1};

// scores: [4][np] (left, right, top, bottom)
__global__ void nb_score_kernel(RiDev ri, const float4* __restrict__ surf, float* __restrict__ scores) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int np = ri.w * ri.h;
  if (i >= np) return;
  int y = i / ri.w, x = i - y * ri.w;
  float4 s = surf[i];
  bool ok = ri.valid(x, y) && s.w >= 0.f;
#pragma unroll
  for (int d = 0; d < 4; ++d) scores[d * np + i] = ok ? nb_change_score(ri, s.w, x, y, c_ox[d], c_oy[d], PR_BORDERS) : 0.0f;
}

__global__ void nb_smooth_kernel(RiDev ri, const float* __restrict__ in, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int np = ri.w * ri.h;
  if (i >= np) return;
  int y = i / ri.w, x = i - y * ri.w;
  for (int d = 0; d < 4; ++d) {
    const float* sc = in + d * np;
    float bs = sc[i];
    float res = bs;
    if (!(bs + 0.5f * (1.0f - bs) < MIN_PROB)) {
      float avg = 0, ws = 0;
      for (int y2 = y - 1; y2 <= y + 1; ++y2)
        for (int x2 = x - 1; x2 <= x + 1; ++x2) {
          if (!ri.in_image(x2, y2) || (x2 == x && y2 == y)) continue;
          avg += sc[y2 * ri.w + x2];
          ws += 1.0f;
        }
      avg /= ws;
      if (!(avg * bs < 0.0f)) res = bs + 0.5f * avg * (1.0f - fabsf(bs));
    }
    out[d * np + i] = res;
  }
}

// one direction d, reading the opposite direction's scores `other` (see the oracle for why two passes of
// this kernel reproduce upstream's scan-order dependent in-place update)
__global__ void nb_shadow_kernel(RiDev ri, float* __restrict__ scores, int d, int other, int* __restrict__ shadow) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int np = ri.w * ri.h;
  if (i >= np) return;
  int y = i / ri.w, x = i - y * ri.w;
  int sidx_out = -1;
  if (ri.valid(x, y)) {
    float* mine = scores + d * np;
    const float* oth = scores + other * np;
    float bs = mine[i];
    if (!(bs < MIN_PROB)) {
      const int ox = c_ox[d], oy = c_oy[d];
      const float rn = ri.range(x + ox, y + oy);
      if (bs == 1.0f && isinf(rn) && rn > 0) {
        sidx_out = (y + oy) * ri.w + x + ox;
      } else {
        float best = -0.5f * MIN_PROB;
        int sidx = -1;
        for (int nd = 1; nd <= PR_BORDERS; ++nd) {
          int nx = x + nd * ox, ny = y + nd * oy;
          if (!ri.in_image(nx, ny)) continue;
          float v = oth[ny * ri.w + nx];
          if (v < best) {
            sidx = ny * ri.w + nx;
            best = v;
          }
        }
        bool keep = false;
        if (sidx >= 0) {
          bs *= fmaxf(0.9f, 1.0f - powf(1.0f + best, 3.0f));
          if (bs >= MIN_PROB) {
            keep = true;
            sidx_out = sidx;
          }
        }
        mine[i] = keep ? bs : 0.0f;
      }
    }
  }
  shadow[d * np + i] = sidx_out;
}

__global__ void nb_classify_kernel(RiDev ri, const float* __restrict__ scores, const int* __restrict__ shadow,
                                   int* __restrict__ traits) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int np = ri.w * ri.h;
  if (i >= np) return;
  int y = i / ri.w, x = i - y * ri.w;
  const int obst_bit[4] = {T_OBST_LEFT, T_OBST_RIGHT, T_OBST_TOP, T_OBST_BOTTOM};
  const int shad_bit[4] = {T_SHAD_RIGHT, T_SHAD_LEFT, T_SHAD_BOTTOM, T_SHAD_TOP};
  const int veil_bit[4] = {T_VEIL_RIGHT, T_VEIL_LEFT, T_VEIL_BOTTOM, T_VEIL_TOP};
  for (int d = 0; d < 4; ++d) {
    const int sidx = shadow[d * np + i];
    if (sidx < 0) continue;
    const float* s = scores + d * np;
    const int ox = c_ox[d], oy = c_oy[d];
    const float bs = s[i];
    bool is_max = true;
    {
      int nx = x - ox, ny = y - oy;
      if (ri.in_image(nx, ny) && s[ny * ri.w + nx] > bs) is_max = false;
      for (int nd = 1; nd <= PR_BORDERS && is_max; ++nd) {
        nx = x + nd * ox; ny = y + nd * oy;
        if (!ri.in_image(nx, ny)) continue;
        int ni = ny * ri.w + nx;
        if (ni == sidx) break;
        if (s[ni] > bs) is_max = false;
      }
    }
    if (!is_max) continue;
    atomicOr(&traits[i], T_OBSTACLE | obst_bit[d]);
    atomicOr(&traits[sidx], T_SHADOW | shad_bit[d]);
    const int stepi = ox + oy * ri.w;
    for (int k = i + stepi; k != sidx; k += stepi) atomicOr(&traits[k], T_VEIL | veil_bit[d]);
  }
}

// dir: (x, y, z, has_direction)
__global__ void nb_direction_kernel(RiDev ri, const int* __restrict__ traits, float4* __restrict__ dir) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ri.w * ri.h) return;
  int y = i / ri.w, x = i - y * ri.w;
  float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
  const int t = traits[i];
  if (t & T_OBSTACLE) {
    out.w = 1.0f;
    int dx = 0, dy = 0;
    if (t & T_OBST_LEFT) dx -= 1;
    if (t & T_OBST_RIGHT) dx += 1;
    if (t & T_OBST_TOP) dy -= 1;
    if (t & T_OBST_BOTTOM) dy += 1;
    if ((dx != 0 || dy != 0) && ri.in_image(x + dx, y + dy)) {
      F3 nbp = ri.point3d((float)(x + dx), (float)(y + dy), ri.px[i].w);
      F3 dv = fnormalized(nbp - ri.pt(x, y));
      out.x = dv.x; out.y = dv.y; out.z = dv.z;
    }
  }
  dir[i] = out;
}

__global__ void nb_dir_average_kernel(RiDev ri, const float4* __restrict__ surf, const float4* __restrict__ dir,
                                      float4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ri.w * ri.h) return;
  int y = i / ri.w, x = i - y * ri.w;
  float4 res = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 me = dir[i];
  if (me.w != 0.f) {
    const float min_cos = cosf(nb_deg2rad(120.0f));
    F3 acc = f3(me.x, me.y, me.z);
    float ws = 1.0f;
    const float max_nd2 = surf[i].w;
    for (int y2 = max(0, y - PR_DIR); y2 <= min(y + PR_DIR, ri.h - 1); ++y2)
      for (int x2 = max(0, x - PR_DIR); x2 <= min(x + PR_DIR, ri.w - 1); ++x2) {
        int i2 = y2 * ri.w + x2;
        float4 o = dir[i2];
        if (o.w == 0.f || i2 == i) continue;
        if (fdot(f3(o.x, o.y, o.z), f3(me.x, me.y, me.z)) < min_cos) continue;
        float between = nb_change_score(ri, max_nd2, x, y, x2 - x, y2 - y, 1);
        if (fabsf(between) >= 0.95f * MIN_PROB) continue;
        acc = acc + f3(o.x, o.y, o.z);
        ws += 1.0f;
      }
    if ((int)lrintf(ws) >= PR_DIR + 1) {
      F3 a = fnormalized(acc);
      res = make_float4(a.x, a.y, a.z, 1.0f);
    }
  }
  out[i] = res;
}

// change: (direction xyz, score)
__global__ void nb_change_kernel(RiDev ri, const float4* __restrict__ surf, const int* __restrict__ traits,
                                 const float4* __restrict__ dir, float4* __restrict__ change) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ri.w * ri.h) return;
  int y = i / ri.w, x = i - y * ri.w;
  float4 res = make_float4(0.f, 0.f, 0.f, 0.f);
  const int t = traits[i];
  if (!(t & (T_VEIL | T_SHADOW))) {
    const float4 d = dir[i];
    if (d.w != 0.f) {
      res = make_float4(d.x, d.y, d.z, 1.0f);
    } else if (ri.valid(x, y) && surf[i].w >= 0.f) {
      VAcc va;
      va.clear();
      const F3 zero = f3(0.f, 0.f, 0.f);
      for (int y2 = y - PR_CURV; y2 <= y + PR_CURV; ++y2)
        for (int x2 = x - PR_CURV; x2 <= x + PR_CURV; ++x2) {
          if (!ri.valid(x2, y2)) continue;
          int i2 = y2 * ri.w + x2;
          if (traits[i2] & (T_VEIL | T_SHADOW)) continue;
          float4 s2 = surf[i2];
          if (s2.w < 0.f) continue;
          va.add(f3(s2.x, s2.y, s2.z), zero);
        }
      if (va.n >= 3) {
        float ev[3];
        F3 e1, e3;
        va.pca(ev, e1, e3);
        float mag = sqrtf(ev[2]);
        if (isfinite(mag)) res = make_float4(e3.x, e3.y, e3.z, mag);
      }
    }
  }
  change[i] = res;
}

// ------------------------------------------------------------------------------------ K8 interest image
constexpr int NK_W = 64;                    // largest window half-size (pixels) a seed may need
constexpr int NK_ROWS = 2 * NK_W + 1;       // 129
constexpr int NK_WORDS = (NK_ROWS + 31) / 32;  // 5
constexpr int NK_WPB = 4;
constexpr int NK_HB = 18;

struct NkSmem {
  unsigned elig[NK_ROWS][NK_WORDS];
  unsigned reach[NK_ROWS][NK_WORDS];
  int hist[NK_HB];
};

// One warp per seed pixel.  overflow: set when a seed's region touches the edge of the largest window.
__global__ void __launch_bounds__(NK_WPB * 32)
nk_interest_kernel(RiDev ri, const int* __restrict__ traits, const float4* __restrict__ change, float support_size,
                   float optimal_distance, float min_change_score, float* __restrict__ interest,
                   int* __restrict__ overflow) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  NkSmem* S = reinterpret_cast<NkSmem*>(smem_raw) + wid;
  const int np = ri.w * ri.h;
  const float search_radius = 0.5f * support_size, radius_sq = search_radius * search_radius,
              radius_recip = 1.0f / search_radius;
  for (int index = blockIdx.x * NK_WPB + wid; index < np; index += gridDim.x * NK_WPB) {
    const int y = index / ri.w, x = index - y * ri.w;
    __syncwarp();
    if (!ri.valid(x, y) || (traits[index] & (T_SHADOW | T_VEIL))) {
      if (lane == 0) interest[index] = 0.0f;
      continue;
    }
    const float4 pp = ri.px[index];
    const F3 p = f3(pp.x, pp.y, pp.z);
    // window half-size: a pixel farther than this cannot lie inside the search sphere (plus the 2-pixel rule).
    // The estimate is generous; if the grown region still touches the window's edge it is redone at the
    // largest size, and if it touches that too the call fails (overflow) instead of truncating silently.
    int W;
    {
      float ratio = fminf(search_radius / fmaxf(pp.w, 1e-6f), 0.999f);
      float ang = asinf(ratio);
      float pix;
      if (ri.planar) {
        float ca = fmaxf(pp.z / pp.w, 0.2f);  // cos of the off-axis angle: pixels stretch like 1 / cos^2
        pix = fmaxf(ri.fx, ri.fy) * tanf(ang) / (ca * ca);
      } else {
        const float ay = (y + ri.offy) * ri.ares - 0.5f * NB_PI;
        const float axc = (x + ri.offx) * ri.ares - NB_PI;  // = angle_x * cos(angle_y)
        pix = ang / ri.ares * (1.0f + fabsf(axc) * fabsf(tanf(ay)));
      }
      W = min(NK_W, max(3, (int)ceilf(1.25f * pix) + 3));
    }
    // ---- quick exit: no pixel with a usable surface-change score in the largest plausible window
    {
      const int rows = 2 * W + 1, x0 = x - W, y0 = y - W;
      bool any = false;
      for (int t = lane; t < rows * rows; t += 32) {
        int wy = t / rows, wx = t - wy * rows;
        int X = x0 + wx, Y = y0 + wy;
        if (ri.in_image(X, Y) && change[Y * ri.w + X].w >= min_change_score) any = true;
      }
      if (!__any_sync(FULL, any)) {
        if (lane == 0) interest[index] = 0.0f;
        continue;
      }
    }
    int rows, words, x0, y0;
    for (;;) {
      rows = 2 * W + 1;
      words = (rows + 31) >> 5;
      x0 = x - W;
      y0 = y - W;
      // ---- eligibility mask of the window
      for (int wy = 0; wy < rows; ++wy) {
        const int Y = y0 + wy;
        for (int k = 0; k < words; ++k) {
          const int wx = k * 32 + lane, X = x0 + wx;
          bool e = false;
          if (wx < rows && ri.valid(X, Y) && !(traits[Y * ri.w + X] & (T_SHADOW | T_VEIL))) {
            const int pixd = max(abs(X - x), abs(Y - y));
            if (pixd <= 2) e = true;
            else {
              F3 d = ri.pt(X, Y) - p;
              e = !(fdot(d, d) > radius_sq);
            }
          }
          const unsigned m = __ballot_sync(FULL, e);
          if (lane == 0) {
            S->elig[wy][k] = m;
            S->reach[wy][k] = 0u;
          }
        }
      }
      if (lane == 0) S->reach[W][W >> 5] = 1u << (W & 31);
      __syncwarp();
      // ---- 8-connected component of the seed inside the eligible set: iterate 3x3 dilation to a fixed point
      for (int iter = 0; iter < 4 * NK_ROWS; ++iter) {
        bool changed = false;
        for (int wy = lane; wy < rows; wy += 32) {
          unsigned v[NK_WORDS];
#pragma unroll
          for (int k = 0; k < NK_WORDS; ++k) {
            unsigned a = 0;
            if (k < words) {
              a = S->reach[wy][k];
              if (wy > 0) a |= S->reach[wy - 1][k];
              if (wy + 1 < rows) a |= S->reach[wy + 1][k];
            }
            v[k] = a;
          }
#pragma unroll
          for (int k = 0; k < NK_WORDS; ++k) {
            if (k >= words) break;
            unsigned hdil = v[k] | (v[k] << 1) | (v[k] >> 1);
            if (k > 0) hdil |= v[k - 1] >> 31;
            if (k + 1 < NK_WORDS) hdil |= v[k + 1] << 31;
            const unsigned old = S->reach[wy][k];
            const unsigned nw = old | (hdil & S->elig[wy][k]);
            if (nw != old) {
              S->reach[wy][k] = nw;
              changed = true;
            }
          }
        }
        __syncwarp();
        if (!__any_sync(FULL, changed)) break;
      }
      // a region that reaches the window's edge may continue outside it
      bool edge = false;
      for (int wy = lane; wy < rows; wy += 32) {
        if (wy == 0 || wy == rows - 1)
          for (int k = 0; k < words; ++k) edge |= S->reach[wy][k] != 0u;
        edge |= (S->reach[wy][0] & 1u) != 0u;
        edge |= ((S->reach[wy][(rows - 1) >> 5] >> ((rows - 1) & 31)) & 1u) != 0u;
      }
      // (the image border is a legitimate edge: only count window edges that lie inside the image)
      if (__any_sync(FULL, edge)) {
        bool inside = false;
        for (int wy = lane; wy < rows; wy += 32) {
          const int Y = y0 + wy;
          if (Y < 0 || Y >= ri.h) continue;
          for (int k = 0; k < words; ++k) {
            unsigned m = S->reach[wy][k];
            while (m) {
              const int b = __ffs(m) - 1;
              m &= m - 1;
              const int wx = k * 32 + b, X = x0 + wx;
              const bool on_edge = wy == 0 || wy == rows - 1 || wx == 0 || wx == rows - 1;
              if (on_edge && ((wy == 0 && Y > 0) || (wy == rows - 1 && Y < ri.h - 1) || (wx == 0 && X > 0) ||
                              (wx == rows - 1 && X < ri.w - 1)))
                inside = true;
            }
          }
        }
        if (__any_sync(FULL, inside)) {
          if (W < NK_W) {
            W = NK_W;
            __syncwarp();
            continue;
          }
          if (lane == 0) atomicAdd(overflow, 1);
        }
      }
      break;
    }
    if (lane < NK_HB) S->hist[lane] = 0;
    __syncwarp();
    // ---- histogram of surface-change directions over the region
    F3 rot0, rot1;
    {
      F3 zdir = fnormalized(p);
      F3 ydir = f3(0.0f, -1.0f, 0.0f);
      rot0 = fnormalized(fcross(ydir, zdir));
      rot1 = fnormalized(fcross(zdir, rot0));
    }
    float negative_score = 1.0f;
    for (int wy = 0; wy < rows; ++wy) {
      const int Y = y0 + wy;
      for (int k = 0; k < words; ++k) {
        const unsigned m = S->reach[wy][k];
        if (m == 0u || !((m >> lane) & 1u)) continue;
        const int X = x0 + k * 32 + lane;
        const int i2 = Y * ri.w + X;
        const float4 ch = change[i2];
        const float s = ch.w;
        if (s < min_change_score) continue;
        const float pixd = (float)max(abs(X - x), abs(Y - y));
        F3 dd = ri.pt(X, Y) - p;
        const float d2 = fdot(dd, dd);
        const float dist = sqrtf(d2), df = radius_recip * dist;
        float neg = 1.0f - 0.5f * s * fmaxf(1.0f - df / optimal_distance, 0.0f);
        neg = neg * neg;
        const float pos = pixd < 2.0f ? s : s * (1.0f - df);
        const F3 dir = f3(ch.x, ch.y, ch.z);
        const float rx = fdot(rot0, dir), ry = fdot(rot1, dir);
        const float rn = sqrtf(rx * rx + ry * ry);
        float c = rn > 0 ? rx / rn : 1.0f;
        c = fminf(1.0f, fmaxf(-1.0f, c));
        const float angle = 0.5f * nb_norm_angle(2.0f * acosf(c));
        int cell = min(NK_HB - 1, (int)lrintf(floorf((angle + nb_deg2rad(90.0f)) / nb_deg2rad(180.0f) * NK_HB)));
        cell = max(cell, 0);
        if (pos > 0.f) atomicMax(&S->hist[cell], __float_as_int(pos));
        negative_score = fminf(negative_score, neg);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) negative_score = fminf(negative_score, __shfl_xor_sync(FULL, negative_score, o));
    __syncwarp();
    float acv = 0.0f;
    for (int pair = lane; pair < NK_HB * NK_HB; pair += 32) {
      const int a = pair / NK_HB, b = pair - a * NK_HB;
      if (b <= a) continue;
      const float ha = __int_as_float(S->hist[a]), hb = __int_as_float(S->hist[b]);
      if (ha == 0.0f || hb == 0.0f) continue;
      float nd = 2.0f * (float)(b - a) / (float)NK_HB;
      nd = nd <= 1.0f ? nd : 2.0f - nd;
      acv = fmaxf(ha * hb * nd, acv);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acv = fmaxf(acv, __shfl_xor_sync(FULL, acv, o));
    if (lane == 0) interest[index] = negative_score * sqrtf(acv);
  }
}

// candidates: valid pixels with interest >= min_interest that are 3x3 maxima; key = (interest desc, index asc)
__global__ void nk_candidates_kernel(RiDev ri, const float* __restrict__ interest, float min_interest,
                                     unsigned long long* __restrict__ keys, int* __restrict__ count, int cap) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ri.w * ri.h) return;
  int y = i / ri.w, x = i - y * ri.w;
  const float v = interest[i];
  if (!ri.valid(x, y) || v < min_interest) return;
  for (int y2 = y - 1; y2 <= y + 1; ++y2)
    for (int x2 = x - 1; x2 <= x + 1; ++x2) {
      if (!ri.in_image(x2, y2)) continue;
      if (interest[y2 * ri.w + x2] > v) return;
    }
  int pos = atomicAdd(count, 1);
  if (pos < cap) keys[pos] = ((unsigned long long)(~__float_as_uint(v)) << 32) | (unsigned)i;  // v > 0
}

constexpr int NK_CAND_CAP = 4096;

// one block: sort the candidates (bitonic, shared memory), greedy minimum-distance selection, mark pixels
__global__ void __launch_bounds__(1024)
nk_select_kernel(RiDev ri, unsigned long long* __restrict__ keys, const int* __restrict__ count, float min_d2,
                 unsigned char* __restrict__ is_kp) {
  __shared__ unsigned long long sk[NK_CAND_CAP];
  __shared__ float kx[NK_CAND_CAP / 4], ky[NK_CAND_CAP / 4], kz[NK_CAND_CAP / 4];  // kept points (<= 2048)
  __shared__ int s_close, s_nkept;
  const int n = min(*count, NK_CAND_CAP);
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) sk[i] = i < n ? keys[i] : 0xffffffffffffffffull;
  __syncthreads();
  for (int k = 2; k <= n2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n2; i += blockDim.x) {
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
  if (threadIdx.x == 0) s_nkept = 0;
  __syncthreads();
  for (int c = 0; c < n; ++c) {
    const int idx = (int)(unsigned)(sk[c] & 0xffffffffull);
    const float4 pp = ri.px[idx];
    if (threadIdx.x == 0) s_close = 0;
    __syncthreads();
    const int nk = s_nkept;
    bool close = false;
    for (int k = threadIdx.x; k < nk; k += blockDim.x) {
      float dx = pp.x - kx[k], dy = pp.y - ky[k], dz = pp.z - kz[k];
      if (dx * dx + dy * dy + dz * dz < min_d2) close = true;
    }
    if (close) s_close = 1;
    __syncthreads();
    if (!s_close && threadIdx.x == 0 && nk < NK_CAND_CAP / 4) {
      kx[nk] = pp.x; ky[nk] = pp.y; kz[nk] = pp.z;
      s_nkept = nk + 1;
      int ix, iy;
      float r;
      ri.project_int(f3(pp.x, pp.y, pp.z), ix, iy, r);
      if (ri.valid(ix, iy)) is_kp[iy * ri.w + ix] = 1;
    }
    __syncthreads();
  }
}

__global__ void nk_flags_kernel(const unsigned char* __restrict__ is_kp, int n, int* __restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = is_kp[i] ? 1 : 0;
}

__global__ void nk_gather_kernel(RiDev ri, const int* __restrict__ kp, int n, const float* __restrict__ interest,
                                 float* __restrict__ xyz, float* __restrict__ val) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = ri.px[kp[i]];
  if (xyz) {
    const F3 w = ri.to_world(f3(p.x, p.y, p.z));
    xyz[3 * i] = w.x; xyz[3 * i + 1] = w.y; xyz[3 * i + 2] = w.z;
  }
  if (val) val[i] = interest[kp[i]];
}

// ------------------------------------------------------------------------------------ K9 Narf36
constexpr int N36_PS = 10, N36_NPS = 20, N36_DS = 36, N36_MAXROT = 8;

struct Pose {  // p' = R p + t, rows of R
  F3 r0, r1, r2, t;
  __device__ __forceinline__ F3 apply(F3 p) const { return f3(fdot(r0, p) + t.x, fdot(r1, p) + t.y, fdot(r2, p) + t.z); }
  __device__ __forceinline__ F3 apply_inv(F3 p) const {
    F3 q = p - t;
    return f3(r0.x * q.x + r1.x * q.y + r2.x * q.z, r0.y * q.x + r1.y * q.y + r2.y * q.z,
              r0.z * q.x + r1.z * q.y + r2.z * q.z);
  }
};

struct N36Smem {
  int patch[N36_PS * N36_PS];      // ordered-int encoding of the float heights (atomicMin)
  float patchf[N36_PS * N36_PS];
  unsigned char bg[N36_PS * N36_PS];
  float integ[N36_NPS * N36_NPS];
  float blurred[N36_NPS * N36_NPS];
  float desc[N36_DS];
  float score[N36_DS];
  float rots[N36_MAXROT];
  int nrot;
};

__device__ __forceinline__ int f2ordi(float f) {
  int b = __float_as_int(f);
  return b >= 0 ? b : (b ^ 0x7fffffff);
}
__device__ __forceinline__ float ordi2f(int b) { return __int_as_float(b >= 0 ? b : (b ^ 0x7fffffff)); }

// (x, y) of step i on the square ring of `radius` around (cx, cy), walking as PCL does
__device__ __forceinline__ void ring_pos(int cx, int cy, int radius, int i, int& x, int& y) {
  if (radius == 0) {
    x = cx; y = cy;
    return;
  }
  // start at (cx - radius - 1, cy - radius); steps 0..2r move +x, then +y, then -x, then -y
  const int r2 = 2 * radius;
  if (i <= r2) { x = cx - radius + i; y = cy - radius; }
  else if (i <= 2 * r2) { x = cx + radius; y = cy - radius + (i - r2); }
  else if (i <= 3 * r2) { x = cx + radius - (i - 2 * r2); y = cy + radius; }
  else { x = cx - radius; y = cy + radius - (i - 3 * r2); }
}

__device__ __forceinline__ void n36_descriptor(const float* patch, float world, float rotation, float* desc, int lane) {
  const int ps = N36_NPS;
  const float w_first = 2.0f;
  const int nbeam = (int)lrintf(ceilf(0.5f * (float)ps));
  const float wf = -2.0f * (w_first - 1.0f) / ((w_first + 1.0f) * (float)(nbeam - 1)), wo = 2.0f * w_first / (w_first + 1.0f);
  const float astep = nb_deg2rad(360.0f) / (float)N36_DS;
  const float cell = world / (float)ps, cf = 1.0f / cell, coff = 0.5f * (world - cell), max_dist = 0.5f * world,
              bpf = (max_dist - 0.5f * cell) / (float)nbeam;
  for (int k = lane; k < N36_DS; k += 32) {
    const float angle = (float)k * astep + rotation, fx_ = sinf(angle) * bpf, fy_ = -cosf(angle) * bpf;
    float cur = 0.0f, prev = 0.0f;
    for (int b = 0; b <= nbeam; ++b) {
      const float bx = fx_ * (float)b, by = fy_ * (float)b;
      int cx_ = (int)lrintf(cf * (bx + coff)), cy_ = (int)lrintf(cf * (by + coff));
      cx_ = min(max(cx_, 0), ps - 1);
      cy_ = min(max(cy_, 0), ps - 1);
      float v = patch[cy_ * ps + cx_];
      if (!isfinite(v)) v = v > 0 ? max_dist : -CUDART_INF_F;
      if (b > 0) cur += (wf * (float)(b - 1) + wo) * (v - prev);
      prev = v;
    }
    desc[k] = atan2f(cur, max_dist) / nb_deg2rad(180.0f);
  }
}

// out rows: [kp][N36_MAXROT][42]; counts[kp] = descriptors produced for keypoint kp
__global__ void __launch_bounds__(128)
narf36_kernel(RiDev ri, const int* __restrict__ kp_px, int n_kp, float support, int rotation_invariant,
              float* __restrict__ out, int* __restrict__ counts) {
  __shared__ N36Smem sm[4];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  N36Smem* S = &sm[wid];
  const int k = blockIdx.x * 4 + wid;
  if (k >= n_kp) return;
  const int idx = kp_px[k];
  const int py = idx / ri.w, pxx = idx - py * ri.w;
  bool ok = idx >= 0 && idx < ri.w * ri.h && ri.valid(pxx, py);
  Pose T;
  T.r0 = T.r1 = T.r2 = T.t = f3(0.f, 0.f, 0.f);
  const float max_dist = 0.5f * support;
  if (ok) {
    // ---- getNormalBasedUprightTransformation: weighted PCA over square rings until a ring has no point in range
    const F3 pos = ri.point3d((float)pxx, (float)py, ri.px[idx].w);
    int cx, cy;
    float rr;
    ri.project_int(pos, cx, cy, rr);
    VAcc va;
    va.clear();
    const float md2 = max_dist * max_dist, mdr = 1.0f / max_dist;
    for (int radius = 1; radius <= ri.w + ri.h; ++radius) {
      bool any = false;
      for (int i = lane; i < 8 * radius; i += 32) {
        int x2, y2;
        ring_pos(cx, cy, radius, i, x2, y2);
        if (!ri.valid(x2, y2)) continue;
        F3 nb = ri.pt(x2, y2);
        F3 d = nb - pos;
        float d2 = fdot(d, d);
        if (d2 > md2) continue;
        any = true;
        va.add(nb, pos, sqrtf(d2) * mdr);
      }
      if (!__any_sync(FULL, any)) break;
    }
    va.sw = warp_sum(va.sw);
    va.n = warp_sum(va.n);
#pragma unroll
    for (int i = 0; i < 3; ++i) va.s[i] = warp_sum(va.s[i]);
#pragma unroll
    for (int i = 0; i < 6; ++i) va.ss[i] = warp_sum(va.ss[i]);
    if (va.n <= 10) {
      ok = false;  // upstream falls back to getNormalForClosestNeighbors here; the oracle rejects as well
    } else {
      float ev[3];
      F3 normal, e3;
      va.pca(ev, normal, e3);
      F3 mean = va.mean(pos);
      if (fdot(normal, fnormalized(mean)) < 0.0f) normal = -1.0f * normal;
      F3 on_plane = (fdot(normal, mean) - fdot(normal, pos)) * normal + pos;
      F3 ydir = ri.world_up();
      T.r0 = fnormalized(fcross(ydir, normal));
      T.r1 = fnormalized(fcross(normal, T.r0));
      T.r2 = fnormalized(normal);
      T.t = -1.0f * f3(fdot(T.r0, on_plane), fdot(T.r1, on_plane), fdot(T.r2, on_plane));
    }
  }
  if (!ok) {
    if (lane == 0) counts[k] = 0;
    return;
  }
  // ---- getInterpolatedSurfaceProjection: rasterise the triangles of the range image into the 10x10 patch.
  // A cell keeps the minimum of the values written to it: atomicMin on order-preserving ints, unwritten = +inf.
  const int ps = N36_PS;
  const float cell = support / (float)ps;
  const float w2c = 1.0f / cell, w2c_off = 0.5f * (float)ps - 0.5f;
  const float c2w = cell, c2w_off = -max_dist + 0.5f * cell;
  const int POS_INF_ORD = f2ordi(CUDART_INF_F);
  for (int c = lane; c < ps * ps; c += 32) S->patch[c] = POS_INF_ORD;
  __syncwarp();
  {
    const F3 position = T.apply_inv(f3(0.f, 0.f, 0.f));
    int mx, my;
    float rr;
    ri.project_int(position, mx, my, rr);
    const int min_search_radius = 2;
    for (int radius = 0; radius <= ri.w + ri.h; ++radius) {
      bool any = radius < min_search_radius;
      const int steps = radius == 0 ? 1 : 8 * radius;
      for (int i = lane; i < steps; i += 32) {
        int x, y;
        ring_pos(mx, my, radius, i, x, y);
        if (!ri.valid(x, y) || !ri.valid(x + 1, y + 1)) continue;
        const F3 p1 = T.apply(ri.pt(x, y));
        if (fabsf(p1.z) > max_dist) continue;
        const F3 p2 = T.apply(ri.pt(x + 1, y + 1));
        if (fabsf(p2.z) > max_dist) continue;
        for (int tri = 0; tri <= 1; ++tri) {
          F3 p3;
          if (tri == 0) {
            if (!ri.valid(x, y + 1)) continue;
            p3 = ri.pt(x, y + 1);
          } else {
            if (!ri.valid(x + 1, y)) continue;
            p3 = ri.pt(x + 1, y);
          }
          p3 = T.apply(p3);
          if (fabsf(p3.z) > max_dist) continue;
          if ((p1.x < -max_dist && p2.x < -max_dist && p3.x < -max_dist) ||
              (p1.x > max_dist && p2.x > max_dist && p3.x > max_dist) ||
              (p1.y < -max_dist && p2.y < -max_dist && p3.y < -max_dist) ||
              (p1.y > max_dist && p2.y > max_dist && p3.y > max_dist))
            continue;
          any = true;
          const float c1x = w2c * p1.x + w2c_off, c1y = w2c * p1.y + w2c_off, c1z = p1.z;
          const float c2x = w2c * p2.x + w2c_off, c2y = w2c * p2.y + w2c_off, c2z = p2.z;
          const float c3x = w2c * p3.x + w2c_off, c3y = w2c * p3.y + w2c_off, c3z = p3.z;
          const int minx = max(0, (int)lrintf(ceilf(fminf(c1x, fminf(c2x, c3x))))),
                    maxx = min(ps - 1, (int)lrintf(floorf(fmaxf(c1x, fmaxf(c2x, c3x))))),
                    miny = max(0, (int)lrintf(ceilf(fminf(c1y, fminf(c2y, c3y))))),
                    maxy = min(ps - 1, (int)lrintf(floorf(fmaxf(c1y, fmaxf(c2y, c3y)))));
          if (maxx < minx || maxy < miny) continue;
          const float v0x = c3x - c1x, v0y = c3y - c1y, v1x = c2x - c1x, v1y = c2y - c1y;
          const float d00 = v0x * v0x + v0y * v0y, d01 = v0x * v1x + v0y * v1y, d11 = v1x * v1x + v1y * v1y;
          const float inv = 1.0f / (d00 * d11 - d01 * d01);
          for (int cx_ = minx; cx_ <= maxx; ++cx_)
            for (int cy_ = miny; cy_ <= maxy; ++cy_) {
              const float v2x = (float)cx_ - c1x, v2y = (float)cy_ - c1y;
              const float d02 = v0x * v2x + v0y * v2y, d12 = v1x * v2x + v1y * v2y;
              const float u = (d11 * d02 - d01 * d12) * inv, v = (d00 * d12 - d01 * d02) * inv;
              if (!((u > -0.01f) && (v >= -0.01f) && (u + v <= 1.01f))) continue;
              const float nv = c1z + u * (c3z - c1z) + v * (c2z - c1z);
              if (nv == nv) atomicMin(&S->patch[cy_ * ps + cx_], f2ordi(nv));
            }
        }
      }
      if (!__any_sync(FULL, any)) break;
    }
  }
  __syncwarp();
  for (int c = lane; c < ps * ps; c += 32) {
    const int o = S->patch[c];
    S->patchf[c] = (o == POS_INF_ORD) ? -CUDART_INF_F : ordi2f(o);
  }
  __syncwarp();
  // max-range cells: an unwritten cell next to a finite cell whose extrapolation lies in front of much farther
  // (or unobserved-far) range readings is background (+inf), and so are the unwritten cells around it
  for (int c = lane; c < ps * ps; c += 32) {
    const int cy_ = c / ps, cx_ = c - cy_ * ps;
    bool is_bg = false;
    if (isinf(S->patchf[c])) {
      for (int y2 = cy_ - 1; y2 <= cy_ + 1 && !is_bg; ++y2)
        for (int x2 = cx_ - 1; x2 <= cx_ + 1; ++x2) {
          if (x2 < 0 || x2 >= ps || y2 < 0 || y2 >= ps || (x2 == cx_ && y2 == cy_)) continue;
          const float nv = S->patchf[y2 * ps + x2];
          if (!isfinite(nv)) continue;
          const float px_ = (float)cx_ + 0.6f * (float)(cx_ - x2), py_ = (float)cy_ + 0.6f * (float)(cy_ - y2);
          F3 fake = f3(c2w * px_ + c2w_off, c2w * py_ + c2w_off, nv);
          fake = T.apply_inv(fake);
          int ix, iy;
          float r;
          ri.project_int(fake, ix, iy, r);
          if (ri.range(ix, iy) - r > max_dist) {
            is_bg = true;
            break;
          }
        }
    }
    S->bg[c] = is_bg ? 1 : 0;
  }
  __syncwarp();
  for (int c = lane; c < ps * ps; c += 32) {
    const int cy_ = c / ps, cx_ = c - cy_ * ps;
    float v = S->patchf[c];
    if (isinf(v)) {
      bool any = false;
      for (int y2 = cy_ - 1; y2 <= cy_ + 1; ++y2)
        for (int x2 = cx_ - 1; x2 <= cx_ + 1; ++x2) {
          if (x2 < 0 || x2 >= ps || y2 < 0 || y2 >= ps) continue;
          if (S->bg[y2 * ps + x2]) any = true;
        }
      if (any) v = CUDART_INF_F;
    }
    // getBlurredSurfacePatch replaces every infinite cell by half the patch size, so fold that in here
    S->integ[c] = isinf(v) ? 0.5f * support : v;  // (integ reused as the cleaned 10x10 patch for a moment)
  }
  __syncwarp();
  // ---- getBlurredSurfacePatch: 20x20 integral image built in upstream's sequential order, then box filter
  if (lane == 0) {
    float clean[N36_PS * N36_PS];
    for (int c = 0; c < ps * ps; ++c) clean[c] = S->integ[c];
    const int nps = N36_NPS;
    const float n2o = (float)ps / (float)nps;
    for (int y = 0; y < nps; ++y)
      for (int x = 0; x < nps; ++x) {
        const int ox = (int)lrintf(floorf(n2o * (float)x)), oy = (int)lrintf(floorf(n2o * (float)y));
        const float v = clean[oy * ps + ox];
        float l = 0, tl = 0, t = 0;
        if (x > 0) {
          l = S->integ[y * nps + x - 1];
          if (y > 0) tl = S->integ[(y - 1) * nps + x - 1];
        }
        if (y > 0) t = S->integ[(y - 1) * nps + x];
        S->integ[y * nps + x] = v + l + t - tl;
      }
  }
  __syncwarp();
  {
    const int nps = N36_NPS, br = 1;
    for (int c = lane; c < nps * nps; c += 32) {
      const int y = c / nps, x = c - y * nps;
      const int top = max(-1, y - br - 1), right = min(nps - 1, x + br), bottom = min(nps - 1, y + br),
                left = max(-1, x - br - 1);
      const float nf = 1.0f / (float)((right - left) * (bottom - top));
      float tlv = 0, trv = 0, brv = S->integ[bottom * nps + right], blv = 0;
      if (left >= 0) {
        blv = S->integ[bottom * nps + left];
        if (top >= 0) tlv = S->integ[top * nps + left];
      }
      if (top >= 0) trv = S->integ[top * nps + right];
      S->blurred[c] = nf * (brv + tlv - blv - trv);
    }
  }
  __syncwarp();
  // ---- descriptor at rotation 0, rotation candidates
  n36_descriptor(S->blurred, support, 0.0f, S->desc, lane);
  __syncwarp();
  if (rotation_invariant) {
    const float s1 = nb_deg2rad(360.0f) / (float)N36_DS, sn = 1.0f / (float)N36_DS;
    for (int st = lane; st < N36_DS; st += 32) {
      const float angle = (float)st * s1;
      float score = 0.0f;
      for (int d = 0; d < N36_DS; ++d) {
        const float a2 = (float)d * s1;
        const float dw = 1.0f - fabsf(nb_norm_angle(angle - a2)) / nb_deg2rad(180.0f);
        score += S->desc[d] * dw * dw;
      }
      S->score[st] = sn * score + 0.5f;
    }
    __syncwarp();
    if (lane == 0) {
      // multimap semantics: ascending score, ties in insertion (= ascending angle) order; take from the back
      float mn = S->score[0], mx = S->score[0];
      for (int i = 1; i < N36_DS; ++i) {
        mn = fminf(mn, S->score[i]);
        mx = fmaxf(mx, S->score[i]);
      }
      const float thr = mx - 0.2f * (mx - mn);
      unsigned long long alive = 0;
      for (int i = 0; i < N36_DS; ++i)
        if (S->score[i] > thr) alive |= 1ull << i;
      const float min_dist = nb_deg2rad(70.0f);
      int nrot = 0;
      while (alive && nrot < N36_MAXROT) {
        int best = -1;
        for (int i = 0; i < N36_DS; ++i)
          if (((alive >> i) & 1ull) && (best < 0 || S->score[i] >= S->score[best])) best = i;
        const float rot = (float)best * s1;
        S->rots[nrot++] = rot;
        alive &= ~(1ull << best);
        for (int i = 0; i < N36_DS; ++i)
          if (((alive >> i) & 1ull) && nb_norm_angle((float)i * s1 - rot) < min_dist) alive &= ~(1ull << i);
      }
      S->nrot = nrot;
    }
  } else if (lane == 0) {
    S->rots[0] = 0.0f;
    S->nrot = 1;
  }
  __syncwarp();
  const int nrot = S->nrot;
  for (int r = 0; r < nrot; ++r) {
    const float rot = S->rots[r];
    float* o = out + ((size_t)k * N36_MAXROT + r) * 42;
    Pose Tr = T;
    if (rotation_invariant) {
      const float c = cosf(-rot), s = sinf(-rot);
      Tr.r0 = c * T.r0 - s * T.r1;
      Tr.r1 = s * T.r0 + c * T.r1;
      Tr.t = f3(c * T.t.x - s * T.t.y, s * T.t.x + c * T.t.y, T.t.z);
      n36_descriptor(S->blurred, support, rot, o + 6, lane);
    } else {
      for (int d = lane; d < N36_DS; d += 32) o[6 + d] = S->desc[d];
    }
    if (lane == 0) {
      // the feature frame in WORLD coordinates: position R p + t, axes R r_i
      const F3 p = ri.to_world(Tr.apply_inv(f3(0.f, 0.f, 0.f)));
      const F3 w0 = ri.rot_to_world(Tr.r0), w1 = ri.rot_to_world(Tr.r1), w2 = ri.rot_to_world(Tr.r2);
      o[0] = p.x; o[1] = p.y; o[2] = p.z;
      o[3] = atan2f(w1.z, w2.z);
      o[4] = asinf(-w0.z);
      o[5] = atan2f(w0.y, w0.x);
    }
    __syncwarp();
  }
  if (lane == 0) counts[k] = nrot;
}

// dense rows from the per-keypoint slots: offsets = exclusive scan of counts
__global__ void narf36_compact_kernel(const float* __restrict__ slots, const int* __restrict__ counts,
                                      const int* __restrict__ offsets, int n_kp, unsigned char* __restrict__ out,
                                      size_t stride, int cap) {
  const int k = blockIdx.x;
  if (k >= n_kp) return;
  const int c = counts[k], o = offsets[k];
  for (int t = threadIdx.x; t < c * 42; t += blockDim.x) {
    const int r = t / 42, e = t - r * 42;
    if (o + r < cap) reinterpret_cast<float*>(out + (size_t)(o + r) * stride)[e] = slots[((size_t)k * N36_MAXROT + r) * 42 + e];
  }
}

// ================================================================================== host
static RiDev ri_view(const Ctx* ctx) {
  RiDev r;
  r.w = ctx->ri.width; r.h = ctx->ri.height; r.planar = ctx->ri.planar;
  r.cx = ctx->ri.cx; r.cy = ctx->ri.cy; r.fx = ctx->ri.fx; r.fy = ctx->ri.fy;
  r.ares = ctx->ri.ang_res; r.offx = ctx->ri.off_x; r.offy = ctx->ri.off_y;
  r.px = ctx->ri_img.as<float4>();
  r.has_pose = ctx->ri_has_pose ? 1 : 0;
  for (int i = 0; i < 9; ++i) r.R[i] = ctx->ri_R[i];
  for (int i = 0; i < 3; ++i) r.t[i] = ctx->ri_t[i];
  return r;
}

int range_image_build(Ctx* ctx, const pfx_range_image_desc* want, float max_angle_w, float max_angle_h, float min_range,
                      int border) {
  const int n = (int)ctx->n;
  pfx_range_image_desc d = *want;
  if (!d.planar) {
    const float recip = 1.0f / d.ang_res;
    d.width = (int)lrintf(floorf(max_angle_w * recip));
    d.height = (int)lrintf(floorf(max_angle_h * recip));
    const int full_w = (int)lrintf(floorf((float)(2.0 * 3.14159265358979323846) * recip)),
              full_h = (int)lrintf(floorf((float)3.14159265358979323846 * recip));
    d.off_x = (full_w - d.width) / 2;
    d.off_y = (full_h - d.height) / 2;
  }
  if (d.width <= 0 || d.height <= 0 || (long long)d.width * d.height > (1ll << 26))
    return ctx->fail(PFX_E_INVALID, "range image: bad image size");
  const int np = d.width * d.height;
  PFX_CUDA(ctx->tmp0.ensure((size_t)np * 2 * sizeof(unsigned)));
  PFX_CUDA(ctx->small.ensure(256));
  unsigned* direct = ctx->tmp0.as<unsigned>();
  unsigned* splat = direct + np;
  RiBuild* bb = reinterpret_cast<RiBuild*>(ctx->small.as<char>() + 128);
  ctx->ri = d;
  RiDev proj = ri_view(ctx);
  proj.px = nullptr;
  PFX_LAUNCH(ctx, ri_init_kernel, div_up(np, 256), 256, 0, direct, splat, np, bb, d.width, d.height);
  if (n > 0)
    PFX_LAUNCH(ctx, ri_project_kernel, div_up(n, 256), 256, 0, ctx->surf.as<float4>(), n, proj, min_range, direct, splat, bb);
  int left = 0, top = 0;
  pfx_range_image_desc out = d;
  if (!d.planar) {  // crop to the bounding box of the written pixels (+ border)
    RiBuild h;
    PFX_CUDA(cudaMemcpyAsync(&h, bb, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    PFX_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h.right < h.left || h.bottom < h.top) {
      out.width = out.height = 0;
      ctx->ri = out;
      ctx->ri_valid = true;
      ctx->ri_stage = 0;
      return 0;
    }
    left = h.left - border;
    top = h.top - border;
    out.width = (h.right + border) - left + 1;
    out.height = (h.bottom + border) - top + 1;
    out.off_x = d.off_x + left;
    out.off_y = d.off_y + top;
  }
  const int onp = out.width * out.height;
  PFX_CUDA(ctx->ri_img.ensure((size_t)onp * sizeof(float4)));
  ctx->ri = out;
  RiDev dst = ri_view(ctx);
  PFX_LAUNCH(ctx, ri_finish_kernel, div_up(onp, 256), 256, 0, direct, splat, d.width, d.height, left, top, dst,
             ctx->ri_img.as<float4>());
  PFX_CUDA(cudaGetLastError());
  ctx->ri_valid = true;
  ctx->ri_stage = 0;
  return 0;
}

// border extraction (stage 1) and interest image (stage 2, depends on the support size) of the current image
int narf_prepare(Ctx* ctx, int stage, float support_size) {
  if (!ctx->ri_valid) return ctx->fail(PFX_E_STATE, "no range image (pfx_range_image_planar / _spherical / _set)");
  const RiDev ri = ri_view(ctx);
  const int np = ri.w * ri.h;
  if (np == 0) return 0;
  const int B = div_up(np, 128), T = 128;
  if (ctx->ri_stage < 1) {
    PFX_CUDA(ctx->nb_surf.ensure((size_t)np * sizeof(float4)));
    PFX_CUDA(ctx->nb_scores.ensure((size_t)np * 8 * sizeof(float)));
    PFX_CUDA(ctx->nb_shadow.ensure((size_t)np * 4 * sizeof(int)));
    PFX_CUDA(ctx->nb_traits.ensure((size_t)np * sizeof(int)));
    PFX_CUDA(ctx->nb_dir.ensure((size_t)np * 2 * sizeof(float4)));
    PFX_CUDA(ctx->nb_change.ensure((size_t)np * sizeof(float4)));
    float4* surf = ctx->nb_surf.as<float4>();
    float* sc_raw = ctx->nb_scores.as<float>() + (size_t)4 * np;
    float* sc = ctx->nb_scores.as<float>();
    int* shadow = ctx->nb_shadow.as<int>();
    int* traits = ctx->nb_traits.as<int>();
    float4* dir0 = ctx->nb_dir.as<float4>() + np;
    float4* dir = ctx->nb_dir.as<float4>();
    float4* ch = ctx->nb_change.as<float4>();
    PFX_LAUNCH(ctx, nb_surface_kernel, B, T, 0, ri, surf);
    PFX_LAUNCH(ctx, nb_score_kernel, B, T, 0, ri, surf, sc_raw);
    PFX_LAUNCH(ctx, nb_smooth_kernel, B, T, 0, ri, sc_raw, sc);
    PFX_LAUNCH(ctx, nb_shadow_kernel, B, T, 0, ri, sc, 1, 0, shadow);  // right reads the original left scores
    PFX_LAUNCH(ctx, nb_shadow_kernel, B, T, 0, ri, sc, 3, 2, shadow);  // bottom reads the original top scores
    PFX_LAUNCH(ctx, nb_shadow_kernel, B, T, 0, ri, sc, 0, 1, shadow);  // left reads the updated right scores
    PFX_LAUNCH(ctx, nb_shadow_kernel, B, T, 0, ri, sc, 2, 3, shadow);  // top reads the updated bottom scores
    PFX_CUDA(cudaMemsetAsync(traits, 0, (size_t)np * sizeof(int), ctx->stream));
    PFX_LAUNCH(ctx, nb_classify_kernel, B, T, 0, ri, sc, shadow, traits);
    PFX_LAUNCH(ctx, nb_direction_kernel, B, T, 0, ri, traits, dir0);
    PFX_LAUNCH(ctx, nb_dir_average_kernel, B, T, 0, ri, surf, dir0, dir);
    PFX_LAUNCH(ctx, nb_change_kernel, B, T, 0, ri, surf, traits, dir, ch);
    PFX_CUDA(cudaGetLastError());
    ctx->ri_stage = 1;
  }
  if (stage >= 2 && (ctx->ri_stage < 2 || ctx->ri_support != support_size)) {
    PFX_CUDA(ctx->nk_interest.ensure((size_t)np * sizeof(float)));
    PFX_CUDA(ctx->small.ensure(256));
    int* overflow = ctx->small.as<int>() + 48;
    PFX_CUDA(cudaMemsetAsync(overflow, 0, sizeof(int), ctx->stream));
    const size_t smem = sizeof(NkSmem) * NK_WPB;
    if (!ctx->smem_attr_narf) {
      PFX_CUDA(cudaFuncSetAttribute(nk_interest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      ctx->smem_attr_narf = true;
    }
    PFX_LAUNCH(ctx, nk_interest_kernel, std::min(div_up(np, NK_WPB), ctx->sm_count * 16), NK_WPB * 32, smem, ri,
               ctx->nb_traits.as<int>(), ctx->nb_change.as<float4>(), support_size, 0.25f, 0.2f,
               ctx->nk_interest.as<float>(), overflow);
    PFX_CUDA(cudaGetLastError());
    int ov = 0;
    PFX_CUDA(cudaMemcpyAsync(&ov, overflow, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PFX_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ov) return ctx->fail(PFX_E_CAPACITY, "NARF interest image: a support region exceeds the 129 x 129 pixel window");
    ctx->ri_stage = 2;
    ctx->ri_support = support_size;
  }
  return 0;
}

int narf_keypoints(Ctx* ctx, float support_size, int** kp_dev, int* n_kp) {
  PFX_TRY(narf_prepare(ctx, 2, support_size));
  const RiDev ri = ri_view(ctx);
  const int np = ri.w * ri.h;
  *n_kp = 0;
  *kp_dev = nullptr;
  if (np == 0) return 0;
  PFX_CUDA(ctx->tmp1.ensure((size_t)NK_CAND_CAP * sizeof(unsigned long long)));
  PFX_CUDA(ctx->tmp2.ensure((size_t)np + (size_t)np * sizeof(int)));
  PFX_CUDA(ctx->tmp3.ensure((size_t)np * sizeof(int)));
  PFX_CUDA(ctx->small.ensure(256));
  int* count = ctx->small.as<int>() + 52;
  unsigned char* is_kp = ctx->tmp2.as<unsigned char>();
  int* flags = reinterpret_cast<int*>(ctx->tmp2.as<unsigned char>() + (((size_t)np + 15) & ~(size_t)15));
  PFX_CUDA(ctx->tmp2.ensure((((size_t)np + 15) & ~(size_t)15) + (size_t)np * sizeof(int)));
  is_kp = ctx->tmp2.as<unsigned char>();
  flags = reinterpret_cast<int*>(is_kp + (((size_t)np + 15) & ~(size_t)15));
  PFX_CUDA(cudaMemsetAsync(count, 0, sizeof(int), ctx->stream));
  PFX_CUDA(cudaMemsetAsync(is_kp, 0, (size_t)np, ctx->stream));
  PFX_LAUNCH(ctx, nk_candidates_kernel, div_up(np, 256), 256, 0, ri, ctx->nk_interest.as<float>(), 0.45f,
             ctx->tmp1.as<unsigned long long>(), count, NK_CAND_CAP);
  int hc = 0;
  PFX_CUDA(cudaMemcpyAsync(&hc, count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  if (hc > NK_CAND_CAP) return ctx->fail(PFX_E_CAPACITY, "NARF keypoints: more than 4096 interest maxima");
  const float md = 0.25f * support_size;
  PFX_LAUNCH(ctx, nk_select_kernel, 1, 1024, 0, ri, ctx->tmp1.as<unsigned long long>(), count, md * md, is_kp);
  PFX_LAUNCH(ctx, nk_flags_kernel, div_up(np, 256), 256, 0, is_kp, np, flags);
  PFX_CUDA(cudaGetLastError());
  int cnt = 0;
  PFX_TRY(compact_flags(ctx, flags, np, ctx->tmp3.as<int>(), &cnt));
  *kp_dev = ctx->tmp3.as<int>();
  *n_kp = cnt;
  return 0;
}

// the image in WORLD coordinates (what pcl::RangeImage::points holds) from the sensor-frame image, and back:
// xyz of an image pixel is a function of (pixel, range) alone, so the way back re-derives it exactly
__global__ void ri_to_world_kernel(RiDev ri, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ri.w * ri.h) return;
  const float4 p = ri.px[i];
  float4 o = p;
  if (isfinite(p.w)) {
    const F3 w = ri.to_world(f3(p.x, p.y, p.z));
    o.x = w.x; o.y = w.y; o.z = w.z;
  }
  out[i] = o;
}
__global__ void ri_from_ranges_kernel(RiDev ri, float4* __restrict__ img) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ri.w * ri.h) return;
  const float r = img[i].w;
  if (!isfinite(r)) return;
  const int y = i / ri.w, x = i - y * ri.w;
  const F3 v = ri.point3d((float)x, (float)y, r);
  img[i] = make_float4(v.x, v.y, v.z, r);
}

int range_image_export_world(Ctx* ctx, float4* out_dev) {
  const RiDev ri = ri_view(ctx);
  const int np = ri.w * ri.h;
  if (np > 0) PFX_LAUNCH(ctx, ri_to_world_kernel, div_up(np, 256), 256, 0, ri, out_dev);
  PFX_CUDA(cudaGetLastError());
  return 0;
}
int range_image_import_world(Ctx* ctx) {  // ctx->ri_img holds a world-frame image: make it the sensor-frame one
  RiDev ri = ri_view(ctx);
  const int np = ri.w * ri.h;
  if (np > 0) PFX_LAUNCH(ctx, ri_from_ranges_kernel, div_up(np, 256), 256, 0, ri, ctx->ri_img.as<float4>());
  PFX_CUDA(cudaGetLastError());
  return 0;
}

// 3-D points / interest values of the keypoints, delivered to host or device buffers
int narf_keypoint_attrs(Ctx* ctx, const int* kp_dev, int n, float* xyz, float* val, int mem) {
  const RiDev ri = ri_view(ctx);
  float* dx = xyz;
  float* dv = val;
  if (mem == PFX_HOST) {
    PFX_CUDA(ctx->out_stage.ensure((size_t)n * 4 * sizeof(float)));
    dx = ctx->out_stage.as<float>();
    dv = dx + (size_t)3 * n;
  }
  PFX_LAUNCH(ctx, nk_gather_kernel, div_up(n, 256), 256, 0, ri, kp_dev, n, ctx->nk_interest.as<float>(),
             xyz ? dx : nullptr, val ? dv : nullptr);
  PFX_CUDA(cudaGetLastError());
  if (mem == PFX_HOST) {
    if (xyz) PFX_CUDA(cudaMemcpyAsync(xyz, dx, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (val) PFX_CUDA(cudaMemcpyAsync(val, dv, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

int narf36_compute(Ctx* ctx, const int* kp_dev, int n_kp, float support_size, int rotation_invariant,
                   unsigned char* out_dev, size_t stride, int cap, int* n_out) {
  if (!ctx->ri_valid) return ctx->fail(PFX_E_STATE, "no range image (pfx_range_image_planar / _spherical / _set)");
  *n_out = 0;
  if (n_kp == 0 || ctx->ri.width * ctx->ri.height == 0) return 0;
  const RiDev ri = ri_view(ctx);
  PFX_CUDA(ctx->tmp0.ensure((size_t)n_kp * N36_MAXROT * 42 * sizeof(float)));
  PFX_CUDA(ctx->tmp1.ensure((size_t)(n_kp + 1) * 2 * sizeof(int)));
  int* counts = ctx->tmp1.as<int>();
  int* offsets = counts + n_kp + 1;
  PFX_CUDA(ctx->small.ensure(256));
  int* total = ctx->small.as<int>() + 56;
  PFX_LAUNCH(ctx, narf36_kernel, div_up(n_kp, 4), 128, 0, ri, kp_dev, n_kp, support_size, rotation_invariant,
             ctx->tmp0.as<float>(), counts);
  PFX_TRY(scan_exclusive_i32(ctx, counts, offsets, n_kp, total, ctx->scanbuf));
  int ht = 0;
  PFX_CUDA(cudaMemcpyAsync(&ht, total, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  *n_out = ht;
  if (ht > cap) return ctx->fail(PFX_E_CAPACITY, "pfx_narf36: output buffer too small");
  PFX_LAUNCH(ctx, narf36_compact_kernel, n_kp, 64, 0, ctx->tmp0.as<float>(), counts, offsets, n_kp, out_dev, stride, cap);
  PFX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pfx
