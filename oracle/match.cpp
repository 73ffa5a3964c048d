// match.cpp — CPU oracle: descriptor matching + VoxelGrid.  TEST INFRASTRUCTURE ONLY.
// PARITY UNPINNED.  Matching restates reference features.h:224-273 (two exact 1-NN passes with
// pcl::KdTreeFLANN<FeatureT>, L2_Simple<float>: sequential float sum over dimensions, no FMA) and
// the reciprocity filter :240-250.  Ties -> lowest index; NaN target rows are not indexed; NaN
// query rows never match.  VoxelGrid restates upstream filters/impl/voxel_grid.hpp
// (centroid per voxel, ascending voxel id) for config C1's ingest.
#include "oracle_common.hpp"
#include "pcl_oracle.h"

namespace {
inline bool rowFinite(const float* r, int dim) {
  for (int d = 0; d < dim; ++d)
    if (!std::isfinite(r[d])) return false;
  return true;
}
}  // namespace

extern "C" int orc_match_nn(const float* a, int na, const float* b, int nb, int dim, int* nn_idx,
                            float* nn_d2) {
  std::vector<char> bok(nb);
  for (int j = 0; j < nb; ++j) bok[j] = rowFinite(b + (size_t)j * dim, dim);
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = 0; i < na; ++i) {
    const float* ra = a + (size_t)i * dim;
    int best = -1;
    float bd = std::numeric_limits<float>::infinity();
    if (rowFinite(ra, dim)) {
      for (int j = 0; j < nb; ++j) {
        if (!bok[j]) continue;
        const float* rb = b + (size_t)j * dim;
        float s = 0.f;
        for (int d = 0; d < dim; ++d) {
          float df = ra[d] - rb[d];
          s += df * df;
        }
        if (s < bd || best < 0) {
          bd = s;
          best = j;
        }
      }
    }
    nn_idx[i] = best;
    if (nn_d2) nn_d2[i] = bd;
  }
  return 0;
}

extern "C" int orc_match_reciprocal(const float* a, int na, const float* b, int nb, int dim,
                                    int* q_idx, int* m_idx, float* dist, int* n_out) {
  std::vector<int> s2t(na), t2s(nb);
  std::vector<float> ds(na);
  orc_match_nn(a, na, b, nb, dim, s2t.data(), ds.data());
  orc_match_nn(b, nb, a, na, dim, t2s.data(), nullptr);
  int m = 0;
  for (int i = 0; i < na; ++i) {
    if (s2t[i] < 0) continue;
    if (t2s[s2t[i]] == i) {
      q_idx[m] = i;
      m_idx[m] = s2t[i];
      if (dist) dist[m] = ds[i];
      ++m;
    }
  }
  *n_out = m;
  return 0;
}

extern "C" int orc_voxel_grid(const float* pts, int n, float leaf, float* out_xyz, int cap,
                              int* n_out) {
  float inv = 1.0f / leaf;
  float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
  for (int i = 0; i < n; ++i) {
    if (!orc::finite3(pts + 3 * (size_t)i)) continue;
    for (int a = 0; a < 3; ++a) {
      mn[a] = std::min(mn[a], pts[3 * (size_t)i + a]);
      mx[a] = std::max(mx[a], pts[3 * (size_t)i + a]);
    }
  }
  long long minb[3], divb[3];
  for (int a = 0; a < 3; ++a) {
    minb[a] = (long long)std::floor(mn[a] * inv);
    long long maxb = (long long)std::floor(mx[a] * inv);
    divb[a] = maxb - minb[a] + 1;
  }
  if (divb[0] * divb[1] * divb[2] > (long long)std::numeric_limits<int32_t>::max()) return -2;
  std::vector<std::pair<int, int>> vi;  // (voxel id, point)
  vi.reserve(n);
  for (int i = 0; i < n; ++i) {
    const float* p = pts + 3 * (size_t)i;
    if (!orc::finite3(p)) continue;
    int ijk0 = (int)(std::floor(p[0] * inv) - (float)minb[0]);
    int ijk1 = (int)(std::floor(p[1] * inv) - (float)minb[1]);
    int ijk2 = (int)(std::floor(p[2] * inv) - (float)minb[2]);
    int id = ijk0 + ijk1 * (int)divb[0] + ijk2 * (int)(divb[0] * divb[1]);
    vi.push_back({id, i});
  }
  std::sort(vi.begin(), vi.end());
  int m = 0;
  for (size_t s = 0; s < vi.size();) {
    size_t e = s;
    float c[3] = {0, 0, 0};
    while (e < vi.size() && vi[e].first == vi[s].first) {
      for (int a = 0; a < 3; ++a) c[a] += pts[3 * (size_t)vi[e].second + a];
      ++e;
    }
    if (m >= cap) return -3;
    float cnt = (float)(e - s);
    for (int a = 0; a < 3; ++a) out_xyz[3 * (size_t)m + a] = c[a] / cnt;
    ++m;
    s = e;
  }
  *n_out = m;
  return 0;
}
