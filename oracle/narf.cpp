// narf.cpp — CPU oracle: range image, range-image border extraction, NARF keypoints, Narf36.
// TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (structural restatement; SURVEY.md A.7 / A.8 rate the
// confidence medium for the range image and LOW for border / interest constants).
//
// Restates, for the reference call sites keypoints.h:204-224, tools.h:65-76 and evaluation.cpp:629-637:
//   pcl::RangeImage / RangeImagePlanar      createFromPointCloud*, doZBuffer, cropImage,
//                                           recalculate3DPointPositions, getSurfaceInformation,
//                                           get1dPointAverage, getNormalBasedUprightTransformation,
//                                           getInterpolatedSurfaceProjection, getRangeDifference
//   pcl::RangeImageBorderExtractor          local surface structure, border scores, score smoothing,
//                                           shadow borders, border classification, border directions,
//                                           surface changes (+ blur)
//   pcl::NarfKeypoint                       interest image (the COMPLETE variant), non-maximum
//                                           suppression, greedy minimum-distance selection
//   pcl::Narf / NarfDescriptor              surface patch, blur, 36-beam descriptor, rotation invariance
//
// Definitions where upstream is order- or implementation-dependent (each also in DESIGN.md):
//  * sensor pose = identity (sensor at the origin, CAMERA_FRAME), which is what both reference call sites
//    produce for the bundled clouds (VIEWPOINT 0 0 0 1 0 0 0);
//  * exact libm trigonometry instead of upstream's lookup tables in the spherical projection;
//  * noise_level = 0 (both call sites), which makes the z-buffer order-independent;
//  * the interest image is upstream's calculateCompleteInterestImage; upstream's default
//    (calculate_sparse_interest_image) is an approximation of it with a scan-order dependent region growing;
//  * candidates of equal interest are ordered by ascending pixel index (upstream: unstable std::sort);
//  * rotation candidates of equal score are ordered by ascending angle (upstream: multimap insertion order).
#include "oracle_common.hpp"
#include "pcl_oracle.h"

namespace {

const float INF = std::numeric_limits<float>::infinity();
const float kPI = 3.14159265358979323846f;

inline float deg2rad(float d) { return d * (kPI / 180.0f); }
inline float normAngle(float a) {
  if (a >= -kPI && a <= kPI) return a;
  if (a < -kPI) return a + 2 * kPI;
  return a - 2 * kPI;
}

struct V3 {
  float x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float norm(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalized(V3 a) {
  float n = norm(a);
  return n > 0 ? (1.0f / n) * a : a;
}

// symmetric 3x3 eigen decomposition (double Jacobi): ascending eigenvalues, eigenvectors in columns
void eig3(const double A[3][3], double w[3], double V[3][3]) {
  double a[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      a[i][j] = A[i][j];
      V[i][j] = i == j;
    }
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = std::fabs(a[0][1]) + std::fabs(a[0][2]) + std::fabs(a[1][2]);
    if (off == 0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0) continue;
        double theta = (a[q][q] - a[p][p]) / (2 * a[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
        double c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < 3; ++k) {
          double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int o[3] = {0, 1, 2};
  double d[3] = {a[0][0], a[1][1], a[2][2]};
  std::sort(o, o + 3, [&](int i, int j) { return d[i] < d[j]; });
  double Vs[3][3];
  for (int c = 0; c < 3; ++c) {
    w[c] = d[o[c]];
    for (int r = 0; r < 3; ++r) Vs[r][c] = V[r][o[c]];
  }
  std::memcpy(V, Vs, sizeof(Vs));
}

// pcl::VectorAverage<float,3>: weighted running mean / covariance; doPCA = eigen decomposition
struct VecAvg {
  double wsum = 0;
  int n = 0;
  double m[3] = {0, 0, 0};
  double c[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  void add(V3 p, float weight = 1.0f) {
    if (weight == 0.0f) return;
    ++n;
    wsum += weight;
    double alpha = weight / wsum;
    double d[3] = {p.x - m[0], p.y - m[1], p.z - m[2]};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) c[i][j] = (1.0 - alpha) * (c[i][j] + alpha * d[i] * d[j]);
    for (int i = 0; i < 3; ++i) m[i] += alpha * d[i];
  }
  V3 mean() const { return {(float)m[0], (float)m[1], (float)m[2]}; }
  // eigenvalues ascending; e1 = eigenvector of the smallest, e3 of the largest
  void pca(float ev[3], V3& e1, V3& e2, V3& e3) const {
    double w[3], V[3][3];
    eig3(c, w, V);
    for (int i = 0; i < 3; ++i) ev[i] = (float)w[i];
    e1 = {(float)V[0][0], (float)V[1][0], (float)V[2][0]};
    e2 = {(float)V[0][1], (float)V[1][1], (float)V[2][1]};
    e3 = {(float)V[0][2], (float)V[1][2], (float)V[2][2]};
  }
};

// ------------------------------------------------------------------------------ range image
struct RI {
  int w = 0, h = 0;
  bool planar = true;
  float cx = 0, cy = 0, fx = 1, fy = 1;
  float ares = 0;
  int offx = 0, offy = 0;
  std::vector<float> px;  // h*w*4: x, y, z, range (unobserved: NaN xyz, range -inf)
  const float* P(int x, int y) const { return &px[4 * ((size_t)y * w + x)]; }
  bool inImage(int x, int y) const { return x >= 0 && x < w && y >= 0 && y < h; }
  float range(int x, int y) const { return inImage(x, y) ? P(x, y)[3] : -INF; }
  bool valid(int x, int y) const { return inImage(x, y) && std::isfinite(P(x, y)[3]); }
  bool maxRange(int x, int y) const {
    float r = range(x, y);
    return std::isinf(r) && r > 0;
  }
  V3 pt(int x, int y) const {
    const float* p = P(x, y);
    return {p[0], p[1], p[2]};
  }
  void project(V3 p, float& ix, float& iy, float& r) const {
    r = norm(p);
    if (planar) {
      ix = cx + fx * p.x / p.z;
      iy = cy + fy * p.y / p.z;
    } else {
      float ax = std::atan2(p.x, p.z), ay = std::asin(p.y / r);
      ix = (ax * std::cos(ay) + kPI) / ares - (float)offx;
      iy = (ay + 0.5f * kPI) / ares - (float)offy;
    }
  }
  void projectInt(V3 p, int& ix, int& iy, float& r) const {
    float fx_, fy_;
    project(p, fx_, fy_, r);
    ix = (int)std::lrint(fx_);
    iy = (int)std::lrint(fy_);
  }
  V3 point3d(float ix, float iy, float r) const {
    if (planar) {
      float dx = (ix - cx) / fx, dy = (iy - cy) / fy;
      float z = r / std::sqrt(dx * dx + dy * dy + 1.0f);
      return {dx * z, dy * z, z};
    }
    float ay = (iy + (float)offy) * ares - 0.5f * kPI;
    float cay = std::cos(ay);
    float ax = cay == 0.0f ? 0.0f : ((ix + (float)offx) * ares - kPI) / cay;
    return {r * std::sin(ax) * cay, r * std::sin(ay), r * std::cos(ax) * cay};
  }
};

// doZBuffer with noise_level = 0: a pixel with >= 1 direct hit holds the minimum direct range, a pixel with
// only "splat" hits (floor/ceil neighbours of a projection) holds the minimum splat range.
void zbuffer(RI& ri, const float* pts, int n, float min_range, int& top, int& right, int& bottom, int& left) {
  size_t np = (size_t)ri.w * ri.h;
  std::vector<float> direct(np, INF), splat(np, INF);
  std::vector<char> hasd(np, 0), hass(np, 0);
  top = ri.h; right = -1; bottom = -1; left = ri.w;
  for (int i = 0; i < n; ++i) {
    const float* p = pts + 3 * (size_t)i;
    if (!orc::finite3(p)) continue;
    float fx_, fy_, r;
    ri.project({p[0], p[1], p[2]}, fx_, fy_, r);
    if (!std::isfinite(fx_) || !std::isfinite(fy_)) continue;
    int x = (int)std::lrint(fx_), y = (int)std::lrint(fy_);
    if (r < min_range || !ri.inImage(x, y)) continue;
    int fxl = (int)std::lrint(std::floor(fx_)), fyl = (int)std::lrint(std::floor(fy_)),
        cxl = (int)std::lrint(std::ceil(fx_)), cyl = (int)std::lrint(std::ceil(fy_));
    int nx[4] = {fxl, cxl, fxl, cxl}, ny[4] = {fyl, fyl, cyl, cyl};
    for (int k = 0; k < 4; ++k) {
      int X = nx[k], Y = ny[k];
      if ((X == x && Y == y) || !ri.inImage(X, Y)) continue;
      size_t q = (size_t)Y * ri.w + X;
      hass[q] = 1;
      splat[q] = std::min(splat[q], r);
      top = std::min(top, Y); right = std::max(right, X); bottom = std::max(bottom, Y); left = std::min(left, X);
    }
    size_t q = (size_t)y * ri.w + x;
    hasd[q] = 1;
    direct[q] = std::min(direct[q], r);
    top = std::min(top, y); right = std::max(right, x); bottom = std::max(bottom, y); left = std::min(left, x);
  }
  ri.px.assign(np * 4, std::numeric_limits<float>::quiet_NaN());
  for (size_t q = 0; q < np; ++q) ri.px[4 * q + 3] = hasd[q] ? direct[q] : (hass[q] ? splat[q] : -INF);
}

void recalc3d(RI& ri) {
  for (int y = 0; y < ri.h; ++y)
    for (int x = 0; x < ri.w; ++x) {
      float* p = &ri.px[4 * ((size_t)y * ri.w + x)];
      if (!std::isfinite(p[3])) {
        p[0] = p[1] = p[2] = std::numeric_limits<float>::quiet_NaN();
        continue;
      }
      V3 v = ri.point3d((float)x, (float)y, p[3]);
      p[0] = v.x; p[1] = v.y; p[2] = v.z;
    }
}

RI from_desc(const float* img, const orc_ri_desc* d) {
  RI ri;
  ri.w = d->width; ri.h = d->height; ri.planar = d->planar != 0;
  ri.cx = d->cx; ri.cy = d->cy; ri.fx = d->fx; ri.fy = d->fy;
  ri.ares = d->ang_res; ri.offx = d->off_x; ri.offy = d->off_y;
  ri.px.assign(img, img + (size_t)4 * ri.w * ri.h);
  return ri;
}

// ------------------------------------------------------------------------------ border extractor
enum {
  T_OBSTACLE = 1 << 0, T_SHADOW = 1 << 1, T_VEIL = 1 << 2,
  T_OBST_TOP = 1 << 4, T_OBST_RIGHT = 1 << 5, T_OBST_BOTTOM = 1 << 6, T_OBST_LEFT = 1 << 7,
  T_SHAD_TOP = 1 << 8, T_SHAD_RIGHT = 1 << 9, T_SHAD_BOTTOM = 1 << 10, T_SHAD_LEFT = 1 << 11,
  T_VEIL_TOP = 1 << 12, T_VEIL_RIGHT = 1 << 13, T_VEIL_BOTTOM = 1 << 14, T_VEIL_LEFT = 1 << 15
};

struct Surf {  // LocalSurface
  bool ok = false;
  V3 normal_nj{0, 0, 0};  // normal_no_jumps (closest neighbours only)
  float max_nd2 = 0;      // max_neighbor_distance_squared
};

struct Borders {
  const RI* ri = nullptr;
  int pr_borders = 3, pr_plane = 2, pr_dir = 2, pr_curv = 2;
  float min_prob = 0.8f;
  std::vector<Surf> surf;
  std::vector<float> sc[4];  // border scores: 0 left, 1 right, 2 top, 3 bottom
  std::vector<int> shadow[4];
  std::vector<int> traits;
  std::vector<char> has_dir;
  std::vector<V3> dir;
  std::vector<float> sc_score;
  std::vector<V3> sc_dir;
};

// RangeImage::getSurfaceInformation (closest-neighbour part)
bool surface_info(const RI& ri, int x, int y, int radius, int n_closest, int step, Surf& s) {
  V3 p = ri.pt(x, y);
  std::vector<std::pair<float, int>> nb;  // (distance^2, linear index); ties by index
  for (int y2 = y - radius; y2 <= y + radius; y2 += step)
    for (int x2 = x - radius; x2 <= x + radius; x2 += step) {
      if (!ri.valid(x2, y2)) continue;
      V3 d = ri.pt(x2, y2) - p;
      nb.push_back({dot(d, d), y2 * ri.w + x2});
    }
  if ((int)nb.size() < 3) return false;
  std::sort(nb.begin(), nb.end());
  n_closest = std::min((int)nb.size(), n_closest);
  s.max_nd2 = nb[n_closest - 1].first;
  float max_d2 = s.max_nd2 * 4.0f;
  VecAvg va;
  for (auto& e : nb) {
    if (e.first > max_d2) break;
    va.add(ri.pt(e.second % ri.w, e.second / ri.w));
  }
  if (va.n < 3) return false;
  float ev[3];
  V3 e2, e3;
  va.pca(ev, s.normal_nj, e2, e3);
  V3 view = normalized(V3{0, 0, 0} - p);
  if (dot(s.normal_nj, view) < 0) s.normal_nj = -1.0f * s.normal_nj;
  return true;
}

// RangeImage::get1dPointAverage: returns (x, y, z, range)
void point_average_1d(const RI& ri, int x, int y, int dx, int dy, int npts, float out[4]) {
  float wsum = 1.0f;
  float r0 = ri.range(x, y);
  if (ri.inImage(x, y)) {
    const float* p = ri.P(x, y);
    out[0] = p[0]; out[1] = p[1]; out[2] = p[2]; out[3] = p[3];
  } else {
    out[0] = out[1] = out[2] = std::numeric_limits<float>::quiet_NaN();
    out[3] = -INF;
  }
  if (std::isinf(r0)) {
    if (r0 > 0) return;
    wsum = 0;
    out[0] = out[1] = out[2] = out[3] = 0;
  }
  int x2 = x, y2 = y;
  for (int step = 1; step < npts; ++step) {
    x2 += dx; y2 += dy;
    if (!ri.valid(x2, y2)) continue;
    const float* p = ri.P(x2, y2);
    for (int k = 0; k < 4; ++k) out[k] += p[k];
    wsum += 1.0f;
  }
  if (wsum <= 0) {
    out[0] = out[1] = out[2] = std::numeric_limits<float>::quiet_NaN();
    out[3] = -INF;
    return;
  }
  float f = 1.0f / wsum;
  for (int k = 0; k < 4; ++k) out[k] *= f;
}

float neighbor_change_score(const RI& ri, const Surf& s, int x, int y, int ox, int oy, int pr) {
  const float* p = ri.P(x, y);
  float nb[4];
  point_average_1d(ri, x + ox, y + oy, ox, oy, pr, nb);
  if (std::isinf(nb[3])) return nb[3] < 0 ? 0.0f : 1.0f;
  float dx = nb[0] - p[0], dy = nb[1] - p[1], dz = nb[2] - p[2];
  float d2 = dx * dx + dy * dy + dz * dz;
  if (d2 <= s.max_nd2) return 0.0f;
  float ret = 1.0f - std::sqrt(s.max_nd2 / d2);
  if (nb[3] < p[3]) ret = -ret;
  return ret;
}

void extract_borders(const RI& ri, Borders& B) {
  B.ri = &ri;
  const int w = ri.w, h = ri.h;
  const size_t np = (size_t)w * h;
  // local surface structure
  B.surf.assign(np, Surf());
  const int step = std::max(1, B.pr_plane / 2);
  const int n_closest = (B.pr_plane / step + 1) * (B.pr_plane / step + 1);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      if (!ri.valid(x, y)) continue;
      Surf s;
      s.ok = surface_info(ri, x, y, B.pr_plane, n_closest, step, s);
      B.surf[(size_t)y * w + x] = s;
    }
  // border scores
  const int ox[4] = {-1, 1, 0, 0}, oy[4] = {0, 0, -1, 1};
  for (int d = 0; d < 4; ++d) B.sc[d].assign(np, 0.0f);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      size_t i = (size_t)y * w + x;
      if (!ri.valid(x, y) || !B.surf[i].ok) continue;
      for (int d = 0; d < 4; ++d) B.sc[d][i] = neighbor_change_score(ri, B.surf[i], x, y, ox[d], oy[d], B.pr_borders);
    }
  // updateScoresAccordingToNeighborValues
  for (int d = 0; d < 4; ++d) {
    std::vector<float> ns(np);
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        size_t i = (size_t)y * w + x;
        float bs = B.sc[d][i];
        ns[i] = bs;
        if (bs + 0.5f * (1.0f - bs) < B.min_prob) continue;
        float avg = 0, ws = 0;
        for (int y2 = y - 1; y2 <= y + 1; ++y2)
          for (int x2 = x - 1; x2 <= x + 1; ++x2) {
            if (!ri.inImage(x2, y2) || (x2 == x && y2 == y)) continue;
            avg += B.sc[d][(size_t)y2 * w + x2];
            ws += 1.0f;
          }
        avg /= ws;
        if (avg * bs < 0.0f) continue;
        ns[i] = bs + 0.5f * avg * (1.0f - std::fabs(bs));
      }
    B.sc[d].swap(ns);
  }
  // shadow borders.  Upstream mutates the scores in scan order; the sequence resolves to two passes:
  // right / bottom read the ORIGINAL left / top scores, left / top read the UPDATED right / bottom scores.
  for (int d = 0; d < 4; ++d) B.shadow[d].assign(np, -1);
  auto shadow_pass = [&](int d, int other) {
    std::vector<float>& mine = B.sc[d];
    const std::vector<float>& oth = B.sc[other];
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        size_t i = (size_t)y * w + x;
        if (!ri.valid(x, y)) continue;
        float& bs = mine[i];
        if (bs < B.min_prob) continue;
        if (bs == 1.0f && ri.maxRange(x + ox[d], y + oy[d])) {
          B.shadow[d][i] = (y + oy[d]) * w + x + ox[d];
          continue;
        }
        float best = -0.5f * B.min_prob;
        int sidx = -1;
        for (int nd = 1; nd <= B.pr_borders; ++nd) {
          int nx = x + nd * ox[d], ny = y + nd * oy[d];
          if (!ri.inImage(nx, ny)) continue;
          float v = oth[(size_t)ny * w + nx];
          if (v < best) {
            sidx = ny * w + nx;
            best = v;
          }
        }
        if (sidx >= 0) {
          bs *= std::max(0.9f, 1.0f - std::pow(1.0f + best, 3.0f));
          if (bs >= B.min_prob) {
            B.shadow[d][i] = sidx;
            continue;
          }
        }
        bs = 0.0f;
      }
  };
  shadow_pass(1, 0);  // pass 1: right (reads original left), bottom (reads original top)
  shadow_pass(3, 2);
  shadow_pass(0, 1);  // pass 2: left (reads updated right), top (reads updated bottom)
  shadow_pass(2, 3);
  // classifyBorders
  B.traits.assign(np, 0);
  auto is_max = [&](int x, int y, int d, int sidx) {
    const std::vector<float>& s = B.sc[d];
    float bs = s[(size_t)y * w + x];
    int nx = x - ox[d], ny = y - oy[d];
    if (ri.inImage(nx, ny) && s[(size_t)ny * w + nx] > bs) return false;
    for (int nd = 1; nd <= B.pr_borders; ++nd) {
      nx = x + nd * ox[d]; ny = y + nd * oy[d];
      if (!ri.inImage(nx, ny)) continue;
      int ni = ny * w + nx;
      if (ni == sidx) return true;
      if (s[ni] > bs) return false;
    }
    return true;
  };
  const int obst_bit[4] = {T_OBST_LEFT, T_OBST_RIGHT, T_OBST_TOP, T_OBST_BOTTOM};
  const int shad_bit[4] = {T_SHAD_RIGHT, T_SHAD_LEFT, T_SHAD_BOTTOM, T_SHAD_TOP};
  const int veil_bit[4] = {T_VEIL_RIGHT, T_VEIL_LEFT, T_VEIL_BOTTOM, T_VEIL_TOP};
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      int i = y * w + x;
      for (int d = 0; d < 4; ++d) {
        int sidx = B.shadow[d][i];
        if (sidx < 0 || !is_max(x, y, d, sidx)) continue;
        B.traits[i] |= T_OBSTACLE | obst_bit[d];
        B.traits[sidx] |= T_SHADOW | shad_bit[d];
        int stepi = ox[d] + oy[d] * w;
        for (int k = i + stepi; k != sidx; k += stepi) B.traits[k] |= T_VEIL | veil_bit[d];
      }
    }
  // border directions
  B.has_dir.assign(np, 0);
  B.dir.assign(np, V3{0, 0, 0});
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      int i = y * w + x;
      int t = B.traits[i];
      if (!(t & T_OBSTACLE)) continue;
      B.has_dir[i] = 1;
      int dx = 0, dy = 0;
      if (t & T_OBST_LEFT) dx -= 1;
      if (t & T_OBST_RIGHT) dx += 1;
      if (t & T_OBST_TOP) dy -= 1;
      if (t & T_OBST_BOTTOM) dy += 1;
      if (dx == 0 && dy == 0) continue;
      if (!ri.inImage(x + dx, y + dy)) continue;
      V3 nbp = ri.point3d((float)(x + dx), (float)(y + dy), ri.P(x, y)[3]);
      B.dir[i] = normalized(nbp - ri.pt(x, y));
    }
  {
    std::vector<char> hd2(np, 0);
    std::vector<V3> d2(np, V3{0, 0, 0});
    const float min_cos = std::cos(deg2rad(120.0f));
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        int i = y * w + x;
        if (!B.has_dir[i]) continue;
        V3 acc = B.dir[i];
        float ws = 1.0f;
        for (int y2 = std::max(0, y - B.pr_dir); y2 <= std::min(y + B.pr_dir, h - 1); ++y2)
          for (int x2 = std::max(0, x - B.pr_dir); x2 <= std::min(x + B.pr_dir, w - 1); ++x2) {
            int i2 = y2 * w + x2;
            if (!B.has_dir[i2] || i2 == i) continue;
            if (dot(B.dir[i2], B.dir[i]) < min_cos) continue;
            float between = neighbor_change_score(ri, B.surf[i], x, y, x2 - x, y2 - y, 1);
            if (std::fabs(between) >= 0.95f * B.min_prob) continue;
            acc = acc + B.dir[i2];
            ws += 1.0f;
          }
        if ((int)std::lrint(ws) < B.pr_dir + 1) continue;
        hd2[i] = 1;
        d2[i] = normalized(acc);
      }
    B.has_dir.swap(hd2);
    B.dir.swap(d2);
  }
  // surface changes
  B.sc_score.assign(np, 0.0f);
  B.sc_dir.assign(np, V3{0, 0, 0});
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      int i = y * w + x;
      int t = B.traits[i];
      if (t & (T_VEIL | T_SHADOW)) continue;
      if (B.has_dir[i]) {
        B.sc_score[i] = 1.0f;
        B.sc_dir[i] = B.dir[i];
        continue;
      }
      if (!ri.valid(x, y) || !B.surf[i].ok) continue;
      VecAvg va;
      for (int y2 = y - B.pr_curv; y2 <= y + B.pr_curv; ++y2)
        for (int x2 = x - B.pr_curv; x2 <= x + B.pr_curv; ++x2) {
          if (!ri.valid(x2, y2)) continue;
          int i2 = y2 * w + x2;
          if (B.traits[i2] & (T_VEIL | T_SHADOW)) continue;
          if (!B.surf[i2].ok) continue;
          va.add(B.surf[i2].normal_nj);
        }
      if (va.n < 3) continue;
      float ev[3];
      V3 e1, e2, e3;
      va.pca(ev, e1, e2, e3);
      float mag = std::sqrt(ev[2]);
      if (!std::isfinite(mag)) continue;
      B.sc_score[i] = mag;
      B.sc_dir[i] = e3;
    }
  // (upstream's blurSurfaceChanges() exists but its call at the end of calculateSurfaceChanges is commented
  // out in 1.7.x, so the scores are used unblurred)
}

// ------------------------------------------------------------------------------ NARF keypoints
struct NarfKpParams {
  float support_size = 0.2f;
  float min_distance_between_interest_points = 0.25f;
  float optimal_distance_to_high_surface_change = 0.25f;
  float min_interest_value = 0.45f;
  float min_surface_change_score = 0.2f;
};

// rows of the rotation that takes the viewing direction to +z with "up" = -y
void rotation_to_viewer(V3 point, V3 rows[3]) {
  V3 zdir = normalized(point);
  V3 ydir{0.0f, -1.0f, 0.0f};
  rows[0] = normalized(cross(ydir, zdir));
  rows[1] = normalized(cross(zdir, rows[0]));
  rows[2] = zdir;
}

void interest_image(const RI& ri, const Borders& B, const NarfKpParams& P, std::vector<float>& out) {
  const int w = ri.w, h = ri.h;
  const size_t np = (size_t)w * h;
  out.assign(np, 0.0f);
  const float search_radius = 0.5f * P.support_size, radius_sq = search_radius * search_radius,
              radius_recip = 1.0f / search_radius;
  const int HB = 18;
#pragma omp parallel
  {
    std::vector<char> touched(np, 0);
    std::vector<int> queue;
#pragma omp for schedule(dynamic, 64)
    for (long long index = 0; index < (long long)np; ++index) {
      int y = (int)(index / w), x = (int)(index - (long long)y * w);
      if (!ri.valid(x, y)) continue;
      if (B.traits[index] & (T_SHADOW | T_VEIL)) continue;
      V3 p = ri.pt(x, y);
      V3 rot[3];
      rotation_to_viewer(p, rot);
      float hist[HB];
      for (int k = 0; k < HB; ++k) hist[k] = 0.0f;
      float negative_score = 1.0f;
      queue.clear();
      queue.push_back((int)index);
      touched[index] = 1;
      for (size_t qi = 0; qi < queue.size(); ++qi) {
        int i2 = queue[qi];
        int y2 = i2 / w, x2 = i2 - y2 * w;
        if (!ri.valid(x2, y2)) continue;
        if (B.traits[i2] & (T_SHADOW | T_VEIL)) continue;
        V3 p2 = ri.pt(x2, y2);
        float pixd = (float)std::max(std::abs(x2 - x), std::abs(y2 - y));
        V3 dd = p2 - p;
        float d2 = dot(dd, dd);
        if (pixd > 2.0f && d2 > radius_sq) continue;
        for (int y3 = y2 - 1; y3 <= y2 + 1; ++y3)
          for (int x3 = x2 - 1; x3 <= x2 + 1; ++x3) {
            if (!ri.inImage(x3, y3)) continue;
            int i3 = y3 * w + x3;
            if (!touched[i3]) {
              queue.push_back(i3);
              touched[i3] = 1;
            }
          }
        float s = B.sc_score[i2];
        if (s < P.min_surface_change_score) continue;
        V3 dir = B.sc_dir[i2];
        float dist = std::sqrt(d2), df = radius_recip * dist;
        float neg = 1.0f - 0.5f * s * std::max(1.0f - df / P.optimal_distance_to_high_surface_change, 0.0f);
        neg = neg * neg;
        float pos = pixd < 2.0f ? s : s * (1.0f - df);
        float rx = dot(rot[0], dir), ry = dot(rot[1], dir);
        float rn = std::sqrt(rx * rx + ry * ry);
        float c = rn > 0 ? rx / rn : 1.0f;
        c = std::min(1.0f, std::max(-1.0f, c));
        float angle = 0.5f * normAngle(2.0f * std::acos(c));
        int cell = std::min(HB - 1, (int)std::lrint(std::floor((angle + deg2rad(90.0f)) / deg2rad(180.0f) * HB)));
        cell = std::max(cell, 0);
        hist[cell] = std::max(hist[cell], pos);
        negative_score = std::min(negative_score, neg);
      }
      for (int qi : queue) touched[qi] = 0;
      float acv = 0.0f;
      for (int a = 0; a < HB - 1; ++a) {
        if (hist[a] == 0.0f) continue;
        for (int b = a + 1; b < HB; ++b) {
          if (hist[b] == 0.0f) continue;
          float nd = 2.0f * (float)(b - a) / (float)HB;
          nd = nd <= 1.0f ? nd : 2.0f - nd;
          acv = std::max(hist[a] * hist[b] * nd, acv);
        }
      }
      out[index] = negative_score * std::sqrt(acv);
    }
  }
}

// ------------------------------------------------------------------------------ Narf36
struct Pose {  // rigid transform: p' = R p + t (rows of R)
  V3 r[3];
  V3 t;
  V3 apply(V3 p) const { return {dot(r[0], p) + t.x, dot(r[1], p) + t.y, dot(r[2], p) + t.z}; }
  V3 apply_inv(V3 p) const {
    V3 q = p - t;
    return {r[0].x * q.x + r[1].x * q.y + r[2].x * q.z, r[0].y * q.x + r[1].y * q.y + r[2].y * q.z,
            r[0].z * q.x + r[1].z * q.y + r[2].z * q.z};
  }
};

bool upright_transformation(const RI& ri, V3 point, float max_dist, Pose& T) {
  int x, y;
  float r;
  ri.projectInt(point, x, y, r);
  VecAvg va;
  const float md2 = max_dist * max_dist, mdr = 1.0f / max_dist;
  bool still = true;
  for (int radius = 1; still; ++radius) {
    int x2 = x - radius - 1, y2 = y - radius;
    still = false;
    for (int i = 0; i < 8 * radius; ++i) {
      if (i <= 2 * radius) ++x2;
      else if (i <= 4 * radius) ++y2;
      else if (i <= 6 * radius) --x2;
      else --y2;
      if (!ri.valid(x2, y2)) continue;
      V3 nb = ri.pt(x2, y2);
      V3 d = nb - point;
      float d2 = dot(d, d);
      if (d2 > md2) continue;
      still = true;
      va.add(nb, std::sqrt(d2) * mdr);
    }
    if (radius > ri.w + ri.h) break;
  }
  if (va.n <= 10) return false;  // upstream falls back to getNormalForClosestNeighbors; we reject (rare)
  float ev[3];
  V3 normal, e2, e3;
  va.pca(ev, normal, e2, e3);
  V3 mean = va.mean();
  if (dot(normal, normalized(mean)) < 0.0f) normal = -1.0f * normal;
  V3 on_plane = (dot(normal, mean) - dot(normal, point)) * normal + point;
  V3 ydir{0.0f, 1.0f, 0.0f};
  T.r[0] = normalized(cross(ydir, normal));
  T.r[1] = normalized(cross(normal, T.r[0]));
  T.r[2] = normalized(normal);
  V3 tr{dot(T.r[0], on_plane), dot(T.r[1], on_plane), dot(T.r[2], on_plane)};
  T.t = -1.0f * tr;
  return true;
}

float range_difference(const RI& ri, V3 p) {
  int x, y;
  float r;
  ri.projectInt(p, x, y, r);
  return ri.range(x, y) - r;
}

void surface_patch(const RI& ri, const Pose& T, int ps, float world, std::vector<float>& patch) {
  const float max_dist = 0.5f * world, cell = world / (float)ps;
  const float w2c = 1.0f / cell, w2c_off = 0.5f * (float)ps - 0.5f;
  const float c2w = cell, c2w_off = -max_dist + 0.5f * cell;
  patch.assign((size_t)ps * ps, -INF);
  V3 position = T.apply_inv(V3{0, 0, 0});
  int mx, my;
  float rr;
  ri.projectInt(position, mx, my, rr);
  const int min_search_radius = 2;
  bool still = true;
  for (int radius = 0; still; ++radius) {
    int x = mx - radius - 1, y = my - radius;
    still = radius < min_search_radius;
    for (int i = 0; i < 8 * radius || (radius == 0 && i == 0); ++i) {
      if (i <= 2 * radius) ++x;
      else if (i <= 4 * radius) ++y;
      else if (i <= 6 * radius) --x;
      else --y;
      if (!ri.valid(x, y) || !ri.valid(x + 1, y + 1)) continue;
      V3 p1 = T.apply(ri.pt(x, y));
      if (std::fabs(p1.z) > max_dist) continue;
      V3 p2 = T.apply(ri.pt(x + 1, y + 1));
      if (std::fabs(p2.z) > max_dist) continue;
      for (int tri = 0; tri <= 1; ++tri) {
        V3 p3;
        if (tri == 0) {
          if (!ri.valid(x, y + 1)) continue;
          p3 = ri.pt(x, y + 1);
        } else {
          if (!ri.valid(x + 1, y)) continue;
          p3 = ri.pt(x + 1, y);
        }
        p3 = T.apply(p3);
        if (std::fabs(p3.z) > max_dist) continue;
        if ((p1.x < -max_dist && p2.x < -max_dist && p3.x < -max_dist) ||
            (p1.x > max_dist && p2.x > max_dist && p3.x > max_dist) ||
            (p1.y < -max_dist && p2.y < -max_dist && p3.y < -max_dist) ||
            (p1.y > max_dist && p2.y > max_dist && p3.y > max_dist))
          continue;
        still = true;
        float c1x = w2c * p1.x + w2c_off, c1y = w2c * p1.y + w2c_off, c1z = p1.z;
        float c2x = w2c * p2.x + w2c_off, c2y = w2c * p2.y + w2c_off, c2z = p2.z;
        float c3x = w2c * p3.x + w2c_off, c3y = w2c * p3.y + w2c_off, c3z = p3.z;
        int minx = std::max(0, (int)std::lrint(std::ceil(std::min(c1x, std::min(c2x, c3x))))),
            maxx = std::min(ps - 1, (int)std::lrint(std::floor(std::max(c1x, std::max(c2x, c3x))))),
            miny = std::max(0, (int)std::lrint(std::ceil(std::min(c1y, std::min(c2y, c3y))))),
            maxy = std::min(ps - 1, (int)std::lrint(std::floor(std::max(c1y, std::max(c2y, c3y)))));
        if (maxx < minx || maxy < miny) continue;
        float v0x = c3x - c1x, v0y = c3y - c1y, v1x = c2x - c1x, v1y = c2y - c1y;
        float d00 = v0x * v0x + v0y * v0y, d01 = v0x * v1x + v0y * v1y, d11 = v1x * v1x + v1y * v1y;
        float inv = 1.0f / (d00 * d11 - d01 * d01);
        for (int cx_ = minx; cx_ <= maxx; ++cx_)
          for (int cy_ = miny; cy_ <= maxy; ++cy_) {
            float v2x = (float)cx_ - c1x, v2y = (float)cy_ - c1y;
            float d02 = v0x * v2x + v0y * v2y, d12 = v1x * v2x + v1y * v2y;
            float u = (d11 * d02 - d01 * d12) * inv, v = (d00 * d12 - d01 * d02) * inv;
            if (!((u > -0.01f) && (v >= -0.01f) && (u + v <= 1.01f))) continue;
            float nv = c1z + u * (c3z - c1z) + v * (c2z - c1z);
            float& val = patch[(size_t)cy_ * ps + cx_];
            val = std::isinf(val) ? nv : std::min(val, nv);
          }
      }
    }
    if (radius > ri.w + ri.h) break;
  }
  // max-range cells: a pure function of the finite cells (see the header comment of this file)
  std::vector<char> bg((size_t)ps * ps, 0);
  for (int cy_ = 0; cy_ < ps; ++cy_)
    for (int cx_ = 0; cx_ < ps; ++cx_) {
      if (!std::isinf(patch[(size_t)cy_ * ps + cx_])) continue;
      bool is_bg = false;
      for (int y2 = cy_ - 1; y2 <= cy_ + 1 && !is_bg; ++y2)
        for (int x2 = cx_ - 1; x2 <= cx_ + 1; ++x2) {
          if (x2 < 0 || x2 >= ps || y2 < 0 || y2 >= ps || (x2 == cx_ && y2 == cy_)) continue;
          float nv = patch[(size_t)y2 * ps + x2];
          if (!std::isfinite(nv)) continue;
          float px_ = (float)cx_ + 0.6f * (float)(cx_ - x2), py_ = (float)cy_ + 0.6f * (float)(cy_ - y2);
          V3 fake{c2w * px_ + c2w_off, c2w * py_ + c2w_off, nv};
          fake = T.apply_inv(fake);
          if (range_difference(ri, fake) > max_dist) {
            is_bg = true;
            break;
          }
        }
      bg[(size_t)cy_ * ps + cx_] = is_bg;
    }
  for (int cy_ = 0; cy_ < ps; ++cy_)
    for (int cx_ = 0; cx_ < ps; ++cx_) {
      float& v = patch[(size_t)cy_ * ps + cx_];
      if (!std::isinf(v)) continue;
      bool any = false;
      for (int y2 = cy_ - 1; y2 <= cy_ + 1; ++y2)
        for (int x2 = cx_ - 1; x2 <= cx_ + 1; ++x2) {
          if (x2 < 0 || x2 >= ps || y2 < 0 || y2 >= ps) continue;
          if (std::isinf(patch[(size_t)y2 * ps + x2]) && bg[(size_t)y2 * ps + x2]) any = true;
        }
      if (any) v = INF;  // (a +inf written here is never read as "finite" by another cell)
    }
}

void blurred_patch(const std::vector<float>& patch, int ps, float world, int nps, int br, std::vector<float>& out) {
  float n2o = (float)ps / (float)nps;
  std::vector<float> integ((size_t)nps * nps);
  for (int y = 0; y < nps; ++y)
    for (int x = 0; x < nps; ++x) {
      int ox = (int)std::lrint(std::floor(n2o * (float)x)), oy = (int)std::lrint(std::floor(n2o * (float)y));
      float v = patch[(size_t)oy * ps + ox];
      if (std::isinf(v)) v = 0.5f * world;
      float l = 0, tl = 0, t = 0;
      if (x > 0) {
        l = integ[(size_t)y * nps + x - 1];
        if (y > 0) tl = integ[(size_t)(y - 1) * nps + x - 1];
      }
      if (y > 0) t = integ[(size_t)(y - 1) * nps + x];
      integ[(size_t)y * nps + x] = v + l + t - tl;
    }
  out.assign((size_t)nps * nps, 0.0f);
  for (int y = 0; y < nps; ++y)
    for (int x = 0; x < nps; ++x) {
      int top = std::max(-1, y - br - 1), right = std::min(nps - 1, x + br), bottom = std::min(nps - 1, y + br),
          left = std::max(-1, x - br - 1);
      float nf = 1.0f / (float)((right - left) * (bottom - top));
      float tlv = 0, trv = 0, brv = integ[(size_t)bottom * nps + right], blv = 0;
      if (left >= 0) {
        blv = integ[(size_t)bottom * nps + left];
        if (top >= 0) tlv = integ[(size_t)top * nps + left];
      }
      if (top >= 0) trv = integ[(size_t)top * nps + right];
      out[(size_t)y * nps + x] = nf * (brv + tlv - blv - trv);
    }
}

void extract_descriptor(const std::vector<float>& patch, int ps, float world, float rotation, float* desc, int dsize) {
  const float w_first = 2.0f;
  const int nbeam = (int)std::lrint(std::ceil(0.5f * (float)ps));
  const float wf = -2.0f * (w_first - 1.0f) / ((w_first + 1.0f) * (float)(nbeam - 1)), wo = 2.0f * w_first / (w_first + 1.0f);
  const float astep = deg2rad(360.0f) / (float)dsize;
  const float cell = world / (float)ps, cf = 1.0f / cell, coff = 0.5f * (world - cell), max_dist = 0.5f * world,
              bpf = (max_dist - 0.5f * cell) / (float)nbeam;
  std::vector<float> bv(nbeam + 1);
  for (int k = 0; k < dsize; ++k) {
    float angle = (float)k * astep + rotation, fx_ = std::sin(angle) * bpf, fy_ = -std::cos(angle) * bpf;
    for (int b = 0; b <= nbeam; ++b) {
      float bx = fx_ * (float)b, by = fy_ * (float)b;
      int cx_ = (int)std::lrint(cf * (bx + coff)), cy_ = (int)std::lrint(cf * (by + coff));
      cx_ = std::min(std::max(cx_, 0), ps - 1);
      cy_ = std::min(std::max(cy_, 0), ps - 1);
      float v = patch[(size_t)cy_ * ps + cx_];
      if (!std::isfinite(v)) v = v > 0 ? max_dist : -INF;
      bv[b] = v;
    }
    float cur = 0.0f;
    for (int b = 0; b < nbeam; ++b) cur += (wf * (float)b + wo) * (bv[b + 1] - bv[b]);
    desc[k] = std::atan2(cur, max_dist) / deg2rad(180.0f);
  }
}

void get_rotations(const float* desc, int dsize, std::vector<float>& rotations) {
  const int steps = std::max(dsize, 36);
  const float min_dist = deg2rad(70.0f), s1 = deg2rad(360.0f) / (float)steps, s2 = deg2rad(360.0f) / (float)dsize,
              sn = 1.0f / (float)dsize;
  std::vector<std::pair<float, float>> so;  // (score, angle), ascending score then ascending angle
  for (int st = 0; st < steps; ++st) {
    float angle = (float)st * s1, score = 0.0f;
    for (int k = 0; k < dsize; ++k) {
      float a2 = (float)k * s2;
      float dw = 1.0f - std::fabs(normAngle(angle - a2)) / deg2rad(180.0f);
      score += desc[k] * dw * dw;
    }
    so.push_back({sn * score + 0.5f, angle});
  }
  std::stable_sort(so.begin(), so.end(), [](const std::pair<float, float>& a, const std::pair<float, float>& b) { return a.first < b.first; });
  float mn = so.front().first, mx = so.back().first;
  float thr = mx - 0.2f * (mx - mn);
  std::vector<std::pair<float, float>> rem;
  for (auto& e : so)
    if (e.first > thr) rem.push_back(e);
  while (!rem.empty()) {
    float rot = rem.back().second;
    rotations.push_back(rot);
    rem.pop_back();
    std::vector<std::pair<float, float>> keep;
    for (auto& e : rem)
      if (!(normAngle(e.second - rot) < min_dist)) keep.push_back(e);
    rem.swap(keep);
  }
}

}  // namespace

// ================================================================================== C interface
extern "C" int orc_range_image_planar(const float* pts, int n, int width, int height, float cx, float cy, float fx,
                                      float fy, float min_range, float* img) {
  RI ri;
  ri.w = width; ri.h = height; ri.planar = true;
  ri.cx = cx; ri.cy = cy; ri.fx = fx; ri.fy = fy;
  int t, r, b, l;
  zbuffer(ri, pts, n, min_range, t, r, b, l);
  recalc3d(ri);
  std::memcpy(img, ri.px.data(), ri.px.size() * sizeof(float));
  return 0;
}

extern "C" int orc_range_image_spherical(const float* pts, int n, float ang_res, float max_angle_w, float max_angle_h,
                                         float min_range, int border, float* img, int cap_px, int* out_w, int* out_h,
                                         int* off_x, int* off_y) {
  RI ri;
  ri.planar = false;
  ri.ares = ang_res;
  const float recip = 1.0f / ang_res;
  ri.w = (int)std::lrint(std::floor(max_angle_w * recip));
  ri.h = (int)std::lrint(std::floor(max_angle_h * recip));
  int full_w = (int)std::lrint(std::floor(deg2rad(360.0f) * recip)), full_h = (int)std::lrint(std::floor(deg2rad(180.0f) * recip));
  ri.offx = (full_w - ri.w) / 2;
  ri.offy = (full_h - ri.h) / 2;
  int top, right, bottom, left;
  zbuffer(ri, pts, n, min_range, top, right, bottom, left);
  if (right < left || bottom < top) {
    *out_w = *out_h = 0;
    *off_x = ri.offx; *off_y = ri.offy;
    return 0;
  }
  // cropImage
  top -= border; right += border; bottom += border; left -= border;
  RI cr = ri;
  cr.w = right - left + 1;
  cr.h = bottom - top + 1;
  cr.offx = ri.offx + left;
  cr.offy = ri.offy + top;
  cr.px.assign((size_t)4 * cr.w * cr.h, std::numeric_limits<float>::quiet_NaN());
  for (int y = 0; y < cr.h; ++y)
    for (int x = 0; x < cr.w; ++x) {
      int ox = x + left, oy = y + top;
      cr.px[4 * ((size_t)y * cr.w + x) + 3] = ri.inImage(ox, oy) ? ri.P(ox, oy)[3] : -INF;
    }
  recalc3d(cr);
  *out_w = cr.w; *out_h = cr.h; *off_x = cr.offx; *off_y = cr.offy;
  if ((long long)cr.w * cr.h > cap_px) return 1;
  std::memcpy(img, cr.px.data(), cr.px.size() * sizeof(float));
  return 0;
}

// stage outputs of the border extractor (all optional): traits h*w, border scores 4*h*w (left, right, top,
// bottom; after smoothing and shadow-border evaluation), surface-change score h*w and direction h*w*3
extern "C" int orc_narf_borders(const float* img, const orc_ri_desc* d, int* traits, float* border_scores,
                                float* sc_score, float* sc_dir) {
  RI ri = from_desc(img, d);
  Borders B;
  extract_borders(ri, B);
  const size_t np = (size_t)ri.w * ri.h;
  if (traits) std::memcpy(traits, B.traits.data(), np * sizeof(int));
  if (border_scores)
    for (int k = 0; k < 4; ++k) std::memcpy(border_scores + k * np, B.sc[k].data(), np * sizeof(float));
  if (sc_score) std::memcpy(sc_score, B.sc_score.data(), np * sizeof(float));
  if (sc_dir)
    for (size_t i = 0; i < np; ++i) {
      sc_dir[3 * i] = B.sc_dir[i].x; sc_dir[3 * i + 1] = B.sc_dir[i].y; sc_dir[3 * i + 2] = B.sc_dir[i].z;
    }
  return 0;
}

extern "C" int orc_narf_keypoints(const float* img, const orc_ri_desc* d, float support_size, int* kp_px,
                                  float* kp_interest, int cap, int* n_kp, float* interest_out) {
  RI ri = from_desc(img, d);
  Borders B;
  extract_borders(ri, B);
  NarfKpParams P;
  P.support_size = support_size;
  std::vector<float> interest;
  interest_image(ri, B, P, interest);
  const int w = ri.w, h = ri.h;
  if (interest_out) std::memcpy(interest_out, interest.data(), interest.size() * sizeof(float));
  struct Cand {
    float v;
    int idx;
  };
  std::vector<Cand> cands;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      int i = y * w + x;
      float v = interest[i];
      if (!ri.valid(x, y) || v < P.min_interest_value) continue;
      bool is_max = true;
      for (int y2 = y - 1; y2 <= y + 1 && is_max; ++y2)
        for (int x2 = x - 1; x2 <= x + 1; ++x2) {
          if (!ri.inImage(x2, y2)) continue;
          if (interest[y2 * w + x2] > v) {
            is_max = false;
            break;
          }
        }
      if (is_max) cands.push_back({v, i});
    }
  std::sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) { return a.v > b.v || (a.v == b.v && a.idx < b.idx); });
  const float min_d2 = (P.min_distance_between_interest_points * P.support_size) * (P.min_distance_between_interest_points * P.support_size);
  std::vector<Cand> kept;
  std::vector<char> is_kp((size_t)w * h, 0);
  for (const Cand& c : cands) {
    V3 p = ri.pt(c.idx % w, c.idx / w);
    bool close = false;
    for (const Cand& k : kept) {
      V3 dd = p - ri.pt(k.idx % w, k.idx / w);
      if (dot(dd, dd) < min_d2) {
        close = true;
        break;
      }
    }
    if (close) continue;
    kept.push_back(c);
    int ix, iy;
    float r;
    ri.projectInt(p, ix, iy, r);
    if (ri.valid(ix, iy)) is_kp[(size_t)iy * w + ix] = 1;
  }
  int cnt = 0;
  for (int i = 0; i < w * h; ++i) {
    if (!is_kp[i]) continue;
    if (cnt < cap) {
      kp_px[cnt] = i;
      if (kp_interest) kp_interest[cnt] = interest[i];
    }
    ++cnt;
  }
  *n_kp = cnt;
  return cnt > cap ? 1 : 0;
}

// rows of 42 floats: x, y, z, roll, pitch, yaw, descriptor[36] (pcl::Narf36)
extern "C" int orc_narf36(const float* img, const orc_ri_desc* d, const int* kp_px, int n_kp, float support_size,
                          int rotation_invariant, float* out, int cap, int* n_out) {
  RI ri = from_desc(img, d);
  const int PS = 10, DS = 36;
  int cnt = 0;
  for (int k = 0; k < n_kp; ++k) {
    int idx = kp_px[k];
    int y = idx / ri.w, x = idx - y * ri.w;
    if (!ri.valid(x, y)) continue;
    V3 pos = ri.point3d((float)x, (float)y, ri.P(x, y)[3]);
    Pose T;
    if (!upright_transformation(ri, pos, 0.5f * support_size, T)) continue;
    std::vector<float> patch, blurred;
    surface_patch(ri, T, PS, support_size, patch);
    blurred_patch(patch, PS, support_size, 2 * PS, 1, blurred);
    float desc[DS];
    extract_descriptor(blurred, 2 * PS, support_size, 0.0f, desc, DS);
    std::vector<float> rots;
    if (rotation_invariant) get_rotations(desc, DS, rots);
    else rots.push_back(0.0f);
    for (float rot : rots) {
      Pose Tr = T;
      float dd[DS];
      if (rotation_invariant) {
        // transformation = AngleAxis(-rot, z) * transformation
        float c = std::cos(-rot), s = std::sin(-rot);
        V3 r0 = T.r[0], r1 = T.r[1];
        Tr.r[0] = c * r0 - s * r1;
        Tr.r[1] = s * r0 + c * r1;
        Tr.t = {c * T.t.x - s * T.t.y, s * T.t.x + c * T.t.y, T.t.z};
        extract_descriptor(blurred, 2 * PS, support_size, rot, dd, DS);
      } else {
        std::memcpy(dd, desc, sizeof(dd));
      }
      if (cnt < cap) {
        float* o = out + (size_t)42 * cnt;
        V3 p = Tr.apply_inv(V3{0, 0, 0});
        o[0] = p.x; o[1] = p.y; o[2] = p.z;
        // inverse transformation = (R^T, position): m(i, j) = R^T(i, j) = r[j].component(i)
        float m21 = Tr.r[1].z, m22 = Tr.r[2].z, m20 = Tr.r[0].z, m10 = Tr.r[0].y, m00 = Tr.r[0].x;
        o[3] = std::atan2(m21, m22);
        o[4] = std::asin(-m20);
        o[5] = std::atan2(m10, m00);
        std::memcpy(o + 6, dd, sizeof(dd));
      }
      ++cnt;
    }
  }
  *n_out = cnt;
  return cnt > cap ? 1 : 0;
}
