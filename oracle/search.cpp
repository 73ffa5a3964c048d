// search.cpp — CPU oracle: neighbour search + cloud resolution.  TEST INFRASTRUCTURE ONLY.
// PARITY UNPINNED (see pcl_oracle.h).  Restates the results of pcl::KdTreeFLANN /
// pcl::search::KdTree (FLANN KDTreeSingleIndex, L2_Simple<float>, exact, sorted) as used at
// reference features.h:192-193, tools.h:29-30, keypoints.h:186-187,371-372,408-409.
#include <omp.h>

#include "oracle_common.hpp"
#include "pcl_oracle.h"

using namespace orc;

extern "C" int orc_num_threads(void) { return omp_get_max_threads(); }
extern "C" void orc_set_num_threads(int n) { omp_set_num_threads(n > 0 ? n : omp_get_num_procs()); }

extern "C" int orc_radius_count(const float* surf, int n, const float* q, int nq, double radius,
                                int* counts) {
  Grid g;
  g.build(surf, n, radius);
  float r2f = (float)(radius * radius);  // product in double, then cast
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 256)
    for (int i = 0; i < nq; ++i) {
      g.radius(q + 3 * i, radius, r2f, nb);
      counts[i] = (int)nb.size();
    }
  }
  return 0;
}

extern "C" int orc_radius_search(const float* surf, int n, const float* q, int nq, double radius,
                                 const int64_t* offsets, int* idx, float* d2) {
  Grid g;
  g.build(surf, n, radius);
  float r2f = (float)(radius * radius);
  int bad = 0;
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 256)
    for (int i = 0; i < nq; ++i) {
      g.radius(q + 3 * i, radius, r2f, nb);
      if ((int64_t)nb.size() != offsets[i + 1] - offsets[i]) {
#pragma omp atomic
        bad++;
        continue;
      }
      for (size_t t = 0; t < nb.size(); ++t) {
        idx[offsets[i] + t] = nb[t].idx;
        d2[offsets[i] + t] = nb[t].d2;
      }
    }
  }
  return bad ? -1 : 0;
}

// O(N) scan per query: the ground truth the grid versions are validated against.
extern "C" int orc_radius_search_brute(const float* surf, int n, const float* q, int nq,
                                       double radius, const int64_t* offsets, int* idx, float* d2) {
  float r2f = (float)(radius * radius);
  int bad = 0;
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 16)
    for (int i = 0; i < nq; ++i) {
      nb.clear();
      if (finite3(q + 3 * i))
        for (int j = 0; j < n; ++j) {
          if (!finite3(surf + 3 * j)) continue;
          float dd = dist2f(q + 3 * i, surf + 3 * j);
          if (dd < r2f) nb.push_back({dd, j});
        }
      std::sort(nb.begin(), nb.end());
      if ((int64_t)nb.size() != offsets[i + 1] - offsets[i]) {
#pragma omp atomic
        bad++;
        continue;
      }
      for (size_t t = 0; t < nb.size(); ++t) {
        idx[offsets[i] + t] = nb[t].idx;
        d2[offsets[i] + t] = nb[t].d2;
      }
    }
  }
  return bad ? -1 : 0;
}

static void writeKnnRow(const std::vector<Nbr>& nb, int k, int* idx, float* d2) {
  for (int t = 0; t < k; ++t) {
    if (t < (int)nb.size()) {
      idx[t] = nb[t].idx;
      d2[t] = nb[t].d2;
    } else {
      idx[t] = -1;
      d2[t] = std::numeric_limits<float>::infinity();
    }
  }
}

// Tie-break (documented deviation): ascending (d2, index).
extern "C" int orc_knn(const float* surf, int n, const float* q, int nq, int k, int* idx,
                       float* d2) {
  Grid g;
  g.build(surf, n, autoEdge(surf, n, k));
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 256)
    for (int i = 0; i < nq; ++i) {
      g.knn(q + 3 * i, k, nb);
      writeKnnRow(nb, k, idx + (size_t)i * k, d2 + (size_t)i * k);
    }
  }
  return 0;
}

extern "C" int orc_knn_brute(const float* surf, int n, const float* q, int nq, int k, int* idx,
                             float* d2) {
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 16)
    for (int i = 0; i < nq; ++i) {
      nb.clear();
      if (finite3(q + 3 * i))
        for (int j = 0; j < n; ++j)
          if (finite3(surf + 3 * j)) nb.push_back({dist2f(q + 3 * i, surf + 3 * j), j});
      size_t kk = std::min((size_t)k, nb.size());
      std::partial_sort(nb.begin(), nb.begin() + kk, nb.end());
      nb.resize(kk);
      writeKnnRow(nb, k, idx + (size_t)i * k, d2 + (size_t)i * k);
    }
  }
  return 0;
}

// keypoints.h:401-428: mean over finite points of sqrt(d2 to the 2nd nearest neighbour), double
// accumulator in index order (the sum is order-dependent in the last bits; we keep index order).
extern "C" int orc_cloud_resolution(const float* pts, int n, double* res) {
  Grid g;
  g.build(pts, n, autoEdge(pts, n, 2));
  std::vector<float> s(n, -1.f);
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
      if (!std::isfinite(pts[3 * i])) continue;  // the reference tests x only (keypoints.h:413)
      g.knn(pts + 3 * i, 2, nb);
      if (nb.size() == 2) s[i] = std::sqrt(nb[1].d2);
    }
  }
  double sum = 0;
  int cnt = 0;
  for (int i = 0; i < n; ++i)
    if (s[i] >= 0) {
      sum += (double)s[i];
      ++cnt;
    }
  *res = cnt ? sum / cnt : 0.0;
  return 0;
}
