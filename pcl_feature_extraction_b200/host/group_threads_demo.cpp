// group_threads_demo.cpp — the multi-GPU group API of include/pfx_b200.h driven from ONE C++ process with one thread
// per GPU (the reference is a single process; a host that wants several GPUs for one cloud needs no launcher):
//
//   group_threads_demo <n_gpus> [side = 256]
//
// Every thread creates a context on its device, joins the group with the id made by thread 0, holds an interleaved part
// of a synthetic sheet, calls pfx_slab_distribute (device-resident all-to-all of slab + halo points over NCCL), runs the
// dense normals + FPFH33 stages on its owned + halo points, and thread 0 compares the gathered owned rows with a
// single-GPU run of the whole cloud.  Prints one line: "group_threads_demo n_gpus=N points=P rows_bit_identical=0|1".
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/pfx_b200.h"

static std::vector<float> sheet(int side) {
  std::vector<float> p((size_t)side * side * 3);
  unsigned long long s = 88172645463325252ull;
  auto rnd = [&]() {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    return (double)(s >> 11) * (1.0 / 9007199254740992.0);
  };
  const double h = 0.004;
  for (int i = 0; i < side; ++i)
    for (int j = 0; j < side; ++j) {
      const double x = (i + 0.5 + 0.8 * (rnd() - 0.5)) * h, y = (j + 0.5 + 0.8 * (rnd() - 0.5)) * h;
      const double z = 1.5 + 0.10 * std::sin(6.2831853 * x / 0.9) * std::cos(6.2831853 * y / 1.3) + 0.0005 * (2 * rnd() - 1);
      float* o = &p[((size_t)i * side + j) * 3];
      o[0] = (float)x; o[1] = (float)y; o[2] = (float)z;
    }
  return p;
}

int main(int argc, char** argv) {
  const int world = argc > 1 ? std::atoi(argv[1]) : 2;
  const int side = argc > 2 ? std::atoi(argv[2]) : 256;
  const std::vector<float> pts = sheet(side);
  const int n = side * side, K = 16;
  unsigned char id[PFX_GROUP_ID_BYTES];
  if (pfx_group_unique_id(id) != 0) { std::fprintf(stderr, "NCCL not available\n"); return 3; }
  std::vector<std::vector<float>> rows(world);       // owned FPFH rows per rank
  std::vector<std::vector<int>> gids(world);         // their global ids
  std::atomic<int> failed{0};
  auto worker = [&](int rank) {
    pfx_ctx* ctx = nullptr;
    if (pfx_create(rank, &ctx) != 0) { failed++; return; }
    auto chk = [&](int rc, const char* what) {
      if (rc != 0) { std::fprintf(stderr, "rank %d: %s failed (%d): %s\n", rank, what, rc, pfx_last_error(ctx)); failed++; }
      return rc == 0;
    };
    if (!chk(pfx_group_join(ctx, rank, world, id), "pfx_group_join")) return;
    std::vector<float> part;
    std::vector<int> ids;
    for (int i = rank; i < n; i += world) { ids.push_back(i); part.insert(part.end(), &pts[3 * (size_t)i], &pts[3 * (size_t)i] + 3); }
    pfx_set_viewpoint(ctx, 0, 0, 0);
    size_t n_owned = 0, n_local = 0;
    const double halo = 3 * 0.045;  // three k-th neighbour distances of this sheet at k = 16 (generous)
    if (!chk(pfx_slab_distribute(ctx, part.data(), ids.size(), 12, ids.data(), PFX_HOST, halo, &n_owned, &n_local), "pfx_slab_distribute")) return;
    std::vector<float> f(n_local * 33);
    if (!chk(pfx_normals(ctx, 0.0, K, nullptr, 16, 3, PFX_HOST), "pfx_normals")) return;
    if (!chk(pfx_fpfh(ctx, 0.0, K, f.data(), 132, PFX_HOST), "pfx_fpfh")) return;
    std::vector<int> own(n_owned), gid(n_local);
    chk(pfx_slab_owned_rows(ctx, own.data(), PFX_HOST), "pfx_slab_owned_rows");
    chk(pfx_slab_global_ids(ctx, gid.data(), PFX_HOST), "pfx_slab_global_ids");
    for (size_t t = 0; t < n_owned; ++t) {
      gids[rank].push_back(gid[own[t]]);
      rows[rank].insert(rows[rank].end(), &f[(size_t)own[t] * 33], &f[(size_t)own[t] * 33] + 33);
    }
    double tot[1] = {(double)n_owned};
    chk(pfx_group_allreduce(ctx, tot, 1, 0), "pfx_group_allreduce");
    if ((long long)tot[0] != n) { std::fprintf(stderr, "rank %d: owned points over all ranks %.0f != %d\n", rank, tot[0], n); failed++; }
    pfx_group_leave(ctx);
    pfx_destroy(ctx);
  };
  std::vector<std::thread> th;
  for (int r = 0; r < world; ++r) th.emplace_back(worker, r);
  for (auto& t : th) t.join();
  if (failed) return 1;
  // single-GPU run of the whole cloud
  pfx_ctx* ctx = nullptr;
  if (pfx_create(0, &ctx) != 0) return 1;
  pfx_set_viewpoint(ctx, 0, 0, 0);
  std::vector<float> ref((size_t)n * 33);
  int rc = pfx_set_surface(ctx, pts.data(), n, 12, PFX_HOST);
  rc |= pfx_normals(ctx, 0.0, K, nullptr, 16, 3, PFX_HOST);
  rc |= pfx_fpfh(ctx, 0.0, K, ref.data(), 132, PFX_HOST);
  if (rc != 0) { std::fprintf(stderr, "single-GPU run failed: %s\n", pfx_last_error(ctx)); return 1; }
  pfx_destroy(ctx);
  long long seen = 0, same = 0;
  for (int r = 0; r < world; ++r)
    for (size_t t = 0; t < gids[r].size(); ++t) {
      ++seen;
      same += std::memcmp(&rows[r][t * 33], &ref[(size_t)gids[r][t] * 33], 132) == 0;
    }
  std::printf("group_threads_demo n_gpus=%d points=%d rows_seen=%lld rows_bit_identical=%d\n", world, n, seen,
              (seen == n && same == n) ? 1 : 0);
  return (seen == n && same == n) ? 0 : 1;
}
