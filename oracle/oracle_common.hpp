// oracle_common.hpp — shared helpers of the CPU oracle (TEST INFRASTRUCTURE ONLY; see pcl_oracle.h).
// PARITY UNPINNED: restated PCL 1.7.x / FLANN 1.8 semantics, SURVEY.md Appendix A.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <utility>
#include <vector>

namespace orc {

// FLANN L2_Simple<float> in 3-D as used by pcl::KdTreeFLANN: each op rounded to
// float, accumulation order x, y, z, no FMA (this file is compiled with -ffp-contract=off).
static inline float dist2f(const float* a, const float* b) {
  float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  float s = dx * dx;
  s = s + dy * dy;
  s = s + dz * dz;
  return s;
}

static inline bool finite3(const float* p) {
  return std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]);
}

struct Nbr {
  float d2;
  int idx;
  bool operator<(const Nbr& o) const { return d2 < o.d2 || (d2 == o.d2 && idx < o.idx); }
};

// Uniform grid over the finite surface points.  Only an accelerator: results are identical to the
// brute-force scan (tests/test_oracle_search.py checks that), because cells are over-covered.
struct Grid {
  const float* pts = nullptr;
  int n = 0;
  double edge = 0, inv = 0;
  double mn[3] = {0, 0, 0};
  int dim[3] = {1, 1, 1};
  std::vector<int> cell_start;  // ncell+1
  std::vector<int> order;       // point indices grouped by cell, ascending index inside a cell

  inline void cellOf(const float* p, int c[3]) const {
    for (int a = 0; a < 3; ++a) {
      int v = (int)std::floor(((double)p[a] - mn[a]) * inv);
      c[a] = std::min(std::max(v, 0), dim[a] - 1);
    }
  }
  inline size_t lin(int x, int y, int z) const { return ((size_t)z * dim[1] + y) * dim[0] + x; }

  void build(const float* p, int n_, double edge_) {
    pts = p;
    n = n_;
    double mx[3] = {-1e300, -1e300, -1e300};
    mn[0] = mn[1] = mn[2] = 1e300;
    int nf = 0;
    for (int i = 0; i < n; ++i) {
      if (!finite3(p + 3 * i)) continue;
      ++nf;
      for (int a = 0; a < 3; ++a) {
        mn[a] = std::min(mn[a], (double)p[3 * i + a]);
        mx[a] = std::max(mx[a], (double)p[3 * i + a]);
      }
    }
    if (nf == 0) {
      mn[0] = mn[1] = mn[2] = 0;
      mx[0] = mx[1] = mx[2] = 0;
    }
    edge = edge_;
    // keep the table bounded: at most ~64M cells
    for (;;) {
      double cells = 1;
      for (int a = 0; a < 3; ++a) cells *= std::floor((mx[a] - mn[a]) / edge) + 1;
      if (cells <= 6.4e7) break;
      edge *= 1.5;
    }
    inv = 1.0 / edge;
    for (int a = 0; a < 3; ++a) dim[a] = (int)std::floor((mx[a] - mn[a]) * inv) + 1;
    size_t nc = (size_t)dim[0] * dim[1] * dim[2];
    cell_start.assign(nc + 1, 0);
    std::vector<size_t> cid(n);
    for (int i = 0; i < n; ++i) {
      if (!finite3(p + 3 * i)) {
        cid[i] = (size_t)-1;
        continue;
      }
      int c[3];
      cellOf(p + 3 * i, c);
      cid[i] = lin(c[0], c[1], c[2]);
      cell_start[cid[i] + 1]++;
    }
    for (size_t c = 0; c < nc; ++c) cell_start[c + 1] += cell_start[c];
    order.resize(cell_start[nc]);
    std::vector<int> fill(cell_start.begin(), cell_start.end() - 1);
    for (int i = 0; i < n; ++i)
      if (cid[i] != (size_t)-1) order[fill[cid[i]]++] = i;
  }

  // all points with d2 < r2f, sorted ascending (d2, idx)
  void radius(const float* q, double r, float r2f, std::vector<Nbr>& out) const {
    out.clear();
    if (!finite3(q)) return;
    int lo[3], hi[3];
    double pad = r * (1.0 + 1e-6) + 1e-9;
    for (int a = 0; a < 3; ++a) {
      lo[a] = (int)std::floor(((double)q[a] - pad - mn[a]) * inv);
      hi[a] = (int)std::floor(((double)q[a] + pad - mn[a]) * inv);
      if (hi[a] < 0 || lo[a] > dim[a] - 1) return;
      lo[a] = std::max(lo[a], 0);
      hi[a] = std::min(hi[a], dim[a] - 1);
    }
    for (int z = lo[2]; z <= hi[2]; ++z)
      for (int y = lo[1]; y <= hi[1]; ++y) {
        size_t c0 = lin(lo[0], y, z), c1 = lin(hi[0], y, z);
        for (int t = cell_start[c0]; t < cell_start[c1 + 1]; ++t) {
          int j = order[t];
          float d2 = dist2f(q, pts + 3 * j);
          if (d2 < r2f) out.push_back({d2, j});
        }
      }
    std::sort(out.begin(), out.end());
  }

  // k smallest by (d2, idx), ascending
  void knn(const float* q, int k, std::vector<Nbr>& out) const {
    out.clear();
    if (!finite3(q) || k <= 0) return;
    int c[3];
    cellOf(q, c);
    std::vector<Nbr> heap;  // max-heap of size <= k
    int maxR = std::max(dim[0], std::max(dim[1], dim[2]));
    for (int R = 0; R <= maxR; ++R) {
      // visit the shell at Chebyshev distance R
      for (int z = c[2] - R; z <= c[2] + R; ++z) {
        if (z < 0 || z >= dim[2]) continue;
        for (int y = c[1] - R; y <= c[1] + R; ++y) {
          if (y < 0 || y >= dim[1]) continue;
          bool inner = (std::abs(z - c[2]) < R) && (std::abs(y - c[1]) < R);
          for (int x = c[0] - R; x <= c[0] + R; inner ? x += std::max(2 * R, 1) : ++x) {
            if (x < 0 || x >= dim[0]) continue;
            size_t ci = lin(x, y, z);
            for (int t = cell_start[ci]; t < cell_start[ci + 1]; ++t) {
              int j = order[t];
              Nbr nb{dist2f(q, pts + 3 * j), j};
              if ((int)heap.size() < k) {
                heap.push_back(nb);
                std::push_heap(heap.begin(), heap.end());
              } else if (nb < heap.front()) {
                std::pop_heap(heap.begin(), heap.end());
                heap.back() = nb;
                std::push_heap(heap.begin(), heap.end());
              }
            }
          }
        }
      }
      if ((int)heap.size() == k) {
        // distance from q to the nearest face of the scanned block that still has cells beyond it
        double safe = 1e300;
        for (int a = 0; a < 3; ++a) {
          double u = ((double)q[a] - mn[a]) * inv;
          if (c[a] - R > 0) safe = std::min(safe, (u - (c[a] - R)) * edge);
          if (c[a] + R < dim[a] - 1) safe = std::min(safe, ((c[a] + R + 1) - u) * edge);
        }
        safe *= (1.0 - 1e-6);
        if (safe > 0 && (double)heap.front().d2 < safe * safe) break;
        if (safe == 1e300) break;  // whole grid scanned
      }
      bool all = true;
      for (int a = 0; a < 3; ++a)
        if (c[a] - R > 0 || c[a] + R < dim[a] - 1) all = false;
      if (all) break;
    }
    std::sort(heap.begin(), heap.end());
    out = heap;
  }
};

// cell edge giving a few tens of points per 27-cell block for kNN on surface-like data
static inline double autoEdge(const float* p, int n, int k) {
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  int nf = 0;
  for (int i = 0; i < n; ++i) {
    if (!finite3(p + 3 * i)) continue;
    ++nf;
    for (int a = 0; a < 3; ++a) {
      mn[a] = std::min(mn[a], (double)p[3 * i + a]);
      mx[a] = std::max(mx[a], (double)p[3 * i + a]);
    }
  }
  if (nf < 2) return 1.0;
  double ext[3] = {mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]};
  std::sort(ext, ext + 3);
  // assume a 2-manifold spanning the two largest extents
  double area = std::max(ext[2] * ext[1], 1e-12);
  double e = std::sqrt(area * std::max(k, 2) / (double)nf);
  return std::max(e, 1e-6);
}

// ---- symmetric 3x3 eigen decomposition, cyclic Jacobi in double.  Eigenvalues ascending,
// eigenvectors in the columns of V (V[r][c]).  Stands in for Eigen::SelfAdjointEigenSolver<Matrix3d>
// (ISS, SHOT LRF) and, in double, for pcl::eigen33 (normals gate).
static inline void eigSym3(const double A_in[3][3], double w[3], double V[3][3]) {
  double A[3][3];
  std::memcpy(A, A_in, sizeof(A));
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = (i == j);
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = std::fabs(A[0][1]) + std::fabs(A[0][2]) + std::fabs(A[1][2]);
    double dg = std::fabs(A[0][0]) + std::fabs(A[1][1]) + std::fabs(A[2][2]);
    if (off <= 1e-300 || off <= 1e-18 * dg) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (A[p][q] == 0.0) continue;
        double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int r = 0; r < 3; ++r) {  // A <- A J
          double arp = A[r][p], arq = A[r][q];
          A[r][p] = c * arp - s * arq;
          A[r][q] = s * arp + c * arq;
        }
        for (int r = 0; r < 3; ++r) {  // A <- J^T A
          double apr = A[p][r], aqr = A[q][r];
          A[p][r] = c * apr - s * aqr;
          A[q][r] = s * apr + c * aqr;
        }
        for (int r = 0; r < 3; ++r) {
          double vrp = V[r][p], vrq = V[r][q];
          V[r][p] = c * vrp - s * vrq;
          V[r][q] = s * vrp + c * vrq;
        }
      }
  }
  int id[3] = {0, 1, 2};
  double d[3] = {A[0][0], A[1][1], A[2][2]};
  std::sort(id, id + 3, [&](int a, int b) { return d[a] < d[b]; });
  double Vc[3][3];
  std::memcpy(Vc, V, sizeof(Vc));
  for (int c = 0; c < 3; ++c) {
    w[c] = d[id[c]];
    for (int r = 0; r < 3; ++r) V[r][c] = Vc[r][id[c]];
  }
}

// neighbourhood provider used by all per-point stages: radius (sorted) or k (sorted)
struct Searcher {
  Grid g;
  double radius = 0;
  int k = 0;
  float r2f = 0;
  void init(const float* surf, int n, double radius_, int k_) {
    radius = radius_;
    k = k_;
    if (radius > 0) {
      r2f = (float)(radius * radius);
      g.build(surf, n, radius);
    } else {
      g.build(surf, n, autoEdge(surf, n, k));
    }
  }
  void query(const float* q, std::vector<Nbr>& out) const {
    if (radius > 0)
      g.radius(q, radius, r2f, out);
    else
      g.knn(q, k, out);
  }
};

}  // namespace orc
