// group.cu — multi-GPU plumbing behind the C ABI: a context joins a GROUP (one NCCL communicator, one rank per GPU;
// ranks may be processes or threads), and the three exchanges the sharded path needs run on device memory:
//
//   pfx_slab_distribute   one cloud, held in arbitrary parts by the ranks, is cut into `world` slabs along its longest
//                         axis (equal-count cuts from an all-reduced histogram) and every rank receives the points of
//                         its slab plus a halo of the given width: one device-side routing pass for all peers
//                         (an atomic per warp and peer), the count matrix by ncclAllGather, the payload by grouped ncclSend / ncclRecv
//                         (NVLink P2P), then a sort by global id.  The surface of the context becomes the owned +
//                         halo points in ascending global-id order (so that index tie-breaks resolve as on one
//                         GPU); the unchanged dense stages run on it and pfx_slab_owned_rows lists the rows that
//                         are this rank's results.  (SURVEY.md section 8e partitioning 2.)
//   pfx_match_ring        exact 1-NN with BOTH descriptor sets sharded: target blocks rotate around the ring on a
//                         second stream (ncclSend / ncclRecv) while the current block is matched (tcgen05 engine or
//                         exact scan); each rank keeps the packed (d2 bits << 32 | global index) minimum of its own
//                         queries - the single-GPU "smallest distance, then lowest index" rule.
//   pfx_group_allreduce   small host-side reductions (cloud resolution = sum / count over ranks; support radius of a
//                         k-search chain = max over ranks).
//
// The reference has nothing distributed; the partitioning follows its per-point stages
// (features.h:181-195: every descriptor is a function of a bounded neighbourhood).  NCCL is bound at run time
// (dlopen, preferring a copy the process already carries - PyTorch bundles its own) so that the library has no
// load-order interplay with the host's NCCL and still loads on a box without one.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <mutex>

#include "internal.h"

namespace pfx {

// ------------------------------------------------------------------------------------------- NCCL binding
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the host process' own copy first
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      api.err = std::string("NCCL not found: ") + (dlerror() ? dlerror() : "dlopen failed");
      return;
    }
    api.handle = h;
#define PFX_NCCL_SYM(field, name)                                   \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name)); \
  if (!api.field) api.err = std::string("NCCL symbol missing: ") + name;
    PFX_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    PFX_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    PFX_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    PFX_NCCL_SYM(GroupStart, "ncclGroupStart")
    PFX_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    PFX_NCCL_SYM(Send, "ncclSend")
    PFX_NCCL_SYM(Recv, "ncclRecv")
    PFX_NCCL_SYM(AllReduce, "ncclAllReduce")
    PFX_NCCL_SYM(AllGather, "ncclAllGather")
    PFX_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef PFX_NCCL_SYM
  });
  return &api;
}

struct Group {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  cudaStream_t comm_stream = nullptr;  // ring rotation under the match
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
};

#define PFX_NCCL(call)                                                                             \
  do {                                                                                             \
    ncclResult_t r__ = (call);                                                                     \
    if (r__ != ncclSuccess)                                                                        \
      return ctx->fail(PFX_E_STATE, std::string("NCCL error at " #call ": ") + nccl_api()->GetErrorString(r__)); \
  } while (0)

static int need_group(Ctx* ctx, const char* who) {
  if (!ctx->group) return ctx->fail(PFX_E_STATE, std::string(who) + ": the context has not joined a group (pfx_group_join)");
  return 0;
}

// ------------------------------------------------------------------------------------------- kernels
__device__ __forceinline__ uint32_t ord_bits(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord_float(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// part -> float4 (x, y, z, global id bits); bbox of the finite points as ordered uints (min xyz, max xyz)
__global__ void slab_ingest_kernel(const unsigned char* __restrict__ src, size_t stride, int n, const int* __restrict__ gids,
                                   int gid_offset, float4* __restrict__ out, uint32_t* __restrict__ bbox) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
  if (i < n) {
    const float* p = reinterpret_cast<const float*>(src + (size_t)i * stride);
    const float x = p[0], y = p[1], z = p[2];
    out[i] = make_float4(x, y, z, __int_as_float(gids ? gids[i] : gid_offset + i));
    if (finite3(x, y, z)) {
      mn[0] = mx[0] = ord_bits(x);
      mn[1] = mx[1] = ord_bits(y);
      mn[2] = mx[2] = ord_bits(z);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    mn[a] = __reduce_min_sync(FULL, mn[a]);
    mx[a] = __reduce_max_sync(FULL, mx[a]);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (mn[a] != 0xffffffffu) atomicMin(&bbox[a], mn[a]);
      if (mx[a] != 0u) atomicMax(&bbox[3 + a], mx[a]);
    }
  }
}

__global__ void slab_bbox_to_float_kernel(const uint32_t* __restrict__ bbox, float* __restrict__ mn3, float* __restrict__ mx3) {
  const int a = threadIdx.x;
  if (a < 3) {
    mn3[a] = (bbox[a] == 0xffffffffu) ? CUDART_INF_F : ord_float(bbox[a]);
    mx3[a] = (bbox[3 + a] == 0u) ? -CUDART_INF_F : ord_float(bbox[3 + a]);
  }
}

constexpr int SLAB_BINS = 8192;

__device__ __forceinline__ float coord_of(float4 p, int axis) { return axis == 0 ? p.x : (axis == 1 ? p.y : p.z); }

__global__ void slab_hist_kernel(const float4* __restrict__ pts, int n, int axis, float lo, float inv_w,
                                 unsigned* __restrict__ hist) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = pts[i];
    if (!finite3(p.x, p.y, p.z)) continue;
    int b = (int)floorf((coord_of(p, axis) - lo) * inv_w);
    b = min(max(b, 0), SLAB_BINS - 1);
    atomicAdd(&hist[b], 1u);
  }
}

// All peers in one pass.  A point goes to every rank whose slab, widened by the halo, holds it (non-finite points: to
// rank 0 only).  The order inside a peer's block does not matter - the receiver sorts by global id - so positions
// come from one atomic per (warp, peer) instead of a stable compaction per peer (two scans and three launches for
// each of the `world` peers, twice: for the counts and for the rows).
struct SlabCuts {
  float c[34];  // c[p] .. c[p + 1]: slab of rank p
  int world;
};
template <bool SCATTER>
__global__ void slab_route_kernel(const float4* __restrict__ pts, int n, int axis, SlabCuts cuts, float halo,
                                  int* __restrict__ counters /* COUNT: totals per peer; SCATTER: running write positions */,
                                  float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool in = i < n;
  const float4 p = in ? pts[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  const bool fin = in && finite3(p.x, p.y, p.z);
  const float c = coord_of(p, axis);
  for (int peer = 0; peer < cuts.world; ++peer) {
    const bool need = in && (fin ? (c >= cuts.c[peer] - halo && c < cuts.c[peer + 1] + halo) : peer == 0);
    const unsigned m = __ballot_sync(0xffffffffu, need);
    if (m == 0u) continue;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&counters[peer], __popc(m));
    if (SCATTER) {
      base = __shfl_sync(0xffffffffu, base, leader);
      if (need) out[base + __popc(m & ((1u << lane) - 1u))] = p;
    }
  }
}

// sort keys of the received rows: the global id (unique), value = position in the receive buffer
__global__ void slab_keys_kernel(const float4* __restrict__ rows, int n, uint32_t* __restrict__ keys, int* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = (uint32_t)__float_as_int(rows[i].w);
  vals[i] = i;
}

// local point j = the received row of j-th smallest global id: surface (w = local index), global ids, and the flag
// "inside my slab" (non-finite points belong to slab 0)
__global__ void slab_commit_kernel(const float4* __restrict__ rows, const int* __restrict__ order, int n, int axis, float lo,
                                   float hi, int i_am_first, float4* __restrict__ surf, int* __restrict__ gid,
                                   int* __restrict__ own_flag) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float4 p = rows[order[j]];
  surf[j] = make_float4(p.x, p.y, p.z, __int_as_float(j));
  gid[j] = __float_as_int(p.w);
  int own;
  if (!finite3(p.x, p.y, p.z)) {
    own = i_am_first;
  } else {
    const float c = coord_of(p, axis);
    own = (c >= lo && c < hi) ? 1 : 0;
  }
  own_flag[j] = own;
}

__global__ void slab_own_list_kernel(const int* __restrict__ flags, const int* __restrict__ pos, int n, int* __restrict__ list) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n && flags[j]) list[pos[j]] = j;
}

// ------------------------------------------------------------------------------------------- ring match
__global__ void ring_init_kernel(unsigned long long* __restrict__ best, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) best[i] = ~0ull;
}
// best = min(best, (d2 bits << 32) | (idx + offset)); d2 >= 0 so its bit pattern orders like its value
__global__ void ring_merge_kernel(unsigned long long* __restrict__ best, const int* __restrict__ idx,
                                  const float* __restrict__ d2, int n, int offset) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = idx[i];
  if (j < 0) return;
  const float d = d2[i];
  if (d != d) return;
  const unsigned long long key = ((unsigned long long)__float_as_uint(fabsf(d)) << 32) | (unsigned)(j + offset);
  if (key < best[i]) best[i] = key;
}
__global__ void ring_unpack_kernel(const unsigned long long* __restrict__ best, int n, int* __restrict__ idx,
                                   float* __restrict__ d2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long k = best[i];
  if (k == ~0ull) {
    idx[i] = -1;
    d2[i] = CUDART_INF_F;
  } else {
    idx[i] = (int)(unsigned)(k & 0xffffffffull);
    d2[i] = __uint_as_float((unsigned)(k >> 32));
  }
}

int match_dispatch_dev(Ctx* ctx, const float* a, int na, int lda, const float* b, int nb, int ldb, int dim, int* idx,
                       float* d2);  // capi.cu

void group_release(Ctx* ctx) {
  Group* g = ctx->group;
  if (!g) return;
  if (g->comm) nccl_api()->CommDestroy(g->comm);
  if (g->comm_stream) cudaStreamDestroy(g->comm_stream);
  if (g->ev_a) cudaEventDestroy(g->ev_a);
  if (g->ev_b) cudaEventDestroy(g->ev_b);
  delete g;
  ctx->group = nullptr;
}

}  // namespace pfx

using namespace pfx;

// ================================================================================== group membership
extern "C" int pfx_group_unique_id(void* id128) {
  if (!id128) return PFX_E_INVALID;
  NcclApi* api = nccl_api();
  if (!api->err.empty()) return PFX_E_STATE;
  static_assert(sizeof(ncclUniqueId) == PFX_GROUP_ID_BYTES, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (api->GetUniqueId(&id) != ncclSuccess) return PFX_E_STATE;
  memcpy(id128, &id, sizeof(id));
  return 0;
}

extern "C" int pfx_group_join(pfx_ctx* ctx, int rank, int world, const void* id128) {
  if (!ctx) return PFX_E_INVALID;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return ctx->fail(PFX_E_STATE, "pfx_group_join: cudaSetDevice failed");
  if (!id128 || world < 1 || rank < 0 || rank >= world) return ctx->fail(PFX_E_INVALID, "pfx_group_join: bad rank / world / id");
  if (ctx->group) return ctx->fail(PFX_E_STATE, "pfx_group_join: the context already belongs to a group");
  NcclApi* api = nccl_api();
  if (!api->err.empty()) return ctx->fail(PFX_E_STATE, api->err);
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  Group* g = new Group();
  g->rank = rank;
  g->world = world;
  ncclResult_t r = api->CommInitRank(&g->comm, world, id, rank);
  if (r != ncclSuccess) {
    delete g;
    return ctx->fail(PFX_E_STATE, std::string("pfx_group_join: ncclCommInitRank: ") + api->GetErrorString(r));
  }
  if (cudaStreamCreateWithFlags(&g->comm_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&g->ev_a, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&g->ev_b, cudaEventDisableTiming) != cudaSuccess) {
    ctx->group = g;
    group_release(ctx);
    return ctx->fail(PFX_E_STATE, "pfx_group_join: could not create the communication stream");
  }
  ctx->group = g;
  return 0;
}

extern "C" int pfx_group_leave(pfx_ctx* ctx) {
  if (!ctx) return PFX_E_INVALID;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  group_release(ctx);
  ctx->slab_active = false;
  return 0;
}

extern "C" int pfx_group_info(const pfx_ctx* ctx, int* rank, int* world) {
  if (!ctx) return PFX_E_INVALID;
  if (rank) *rank = ctx->group ? ctx->group->rank : 0;
  if (world) *world = ctx->group ? ctx->group->world : 1;
  return 0;
}

// op: 0 sum, 1 max, 2 min; vals: n host doubles, reduced in place over the ranks
extern "C" int pfx_group_allreduce(pfx_ctx* ctx, double* vals, int n, int op) {
  if (!ctx) return PFX_E_INVALID;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return ctx->fail(PFX_E_STATE, "cudaSetDevice failed");
  if (!vals || n < 1 || n > 4096 || op < 0 || op > 2) return ctx->fail(PFX_E_INVALID, "pfx_group_allreduce: bad arguments");
  if (!ctx->group || ctx->group->world == 1) return 0;
  PFX_CUDA(ctx->small.ensure(256));
  DevBuf& buf = ctx->grp_tmp;
  PFX_CUDA(buf.ensure((size_t)n * sizeof(double)));
  PFX_CUDA(cudaMemcpyAsync(buf.p, vals, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  const ncclRedOp_t ops[3] = {ncclSum, ncclMax, ncclMin};
  PFX_NCCL(nccl_api()->AllReduce(buf.p, buf.p, (size_t)n, ncclDouble, ops[op], ctx->group->comm, ctx->stream));
  PFX_CUDA(cudaMemcpyAsync(vals, buf.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ================================================================================== slab distribution
extern "C" int pfx_slab_distribute(pfx_ctx* ctx, const void* part, size_t n_part, size_t stride, const int32_t* global_ids,
                                   int mem, double halo, size_t* n_owned_out, size_t* n_local_out) {
  if (!ctx) return PFX_E_INVALID;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return ctx->fail(PFX_E_STATE, "cudaSetDevice failed");
  PFX_TRY(need_group(ctx, "pfx_slab_distribute"));
  if ((n_part && !part) || stride < 12 || (stride & 3) || n_part > 0x7fffffffull || !(halo >= 0) ||
      (mem != PFX_HOST && mem != PFX_DEVICE))
    return ctx->fail(PFX_E_INVALID, "pfx_slab_distribute: bad pointer / stride / size / halo");
  Group* G = ctx->group;
  NcclApi* api = nccl_api();
  const int world = G->world, rank = G->rank;
  const int n = (int)n_part;
  cudaStream_t st = ctx->stream;
  PFX_TRY(grid_wait_pending(ctx));

  // ---- ingest: part -> (x, y, z, global id), local bbox
  const unsigned char* src = static_cast<const unsigned char*>(part);
  const int* dgid = global_ids;
  if (mem == PFX_HOST) {
    PFX_CUDA(ctx->stage.ensure(std::max<size_t>((size_t)n * stride, 16)));
    if (n) PFX_CUDA(cudaMemcpyAsync(ctx->stage.p, part, (size_t)n * stride, cudaMemcpyHostToDevice, st));
    src = ctx->stage.as<unsigned char>();
    if (global_ids) {
      PFX_CUDA(ctx->stage2.ensure(std::max<size_t>((size_t)n * sizeof(int), 16)));
      if (n) PFX_CUDA(cudaMemcpyAsync(ctx->stage2.p, global_ids, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
      dgid = ctx->stage2.as<int>();
    }
  }
  // small device scratch: [0..5] bbox as ordered uints, [8..13] bbox floats, [16..16+world) counts, then the matrix
  PFX_CUDA(ctx->grp_tmp.ensure(4096 + (size_t)SLAB_BINS * 4 + (size_t)world * world * 4 + 1024));
  uint32_t* d_bbox = ctx->grp_tmp.as<uint32_t>();
  float* d_bbf = reinterpret_cast<float*>(d_bbox + 8);
  int* d_counts = reinterpret_cast<int*>(d_bbox + 16);
  int* d_matrix = d_counts + 64;
  unsigned* d_hist = reinterpret_cast<unsigned*>(ctx->grp_tmp.as<char>() + 4096 + (size_t)world * world * 4 + 512);
  if (world > 32) return ctx->fail(PFX_E_INVALID, "pfx_slab_distribute: at most 32 ranks");
  {
    uint32_t init[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, 0u, 0u};
    PFX_CUDA(cudaMemcpyAsync(d_bbox, init, sizeof(init), cudaMemcpyHostToDevice, st));
  }
  // global ids default to rank-offset numbering: exclusive prefix of the part sizes
  int gid_offset = 0;
  std::vector<int> sizes(world, 0);
  {
    PFX_CUDA(cudaMemcpyAsync(d_counts, &n, sizeof(int), cudaMemcpyHostToDevice, st));
    PFX_NCCL(api->AllGather(d_counts, d_matrix, 1, ncclInt32, G->comm, st));
    PFX_CUDA(cudaMemcpyAsync(sizes.data(), d_matrix, (size_t)world * sizeof(int), cudaMemcpyDeviceToHost, st));
    PFX_CUDA(cudaStreamSynchronize(st));
    for (int r = 0; r < rank; ++r) gid_offset += sizes[r];
  }
  long long n_total = 0;
  for (int r = 0; r < world; ++r) n_total += sizes[r];
  DevBuf& rows = ctx->slab_rows;  // this rank's part as float4 rows
  PFX_CUDA(rows.ensure(std::max<size_t>(n, 1) * sizeof(float4)));
  PFX_LAUNCH(ctx, slab_ingest_kernel, std::max(1, div_up(n, 256)), 256, 0, src, stride, n, dgid, gid_offset, rows.as<float4>(), d_bbox);
  PFX_LAUNCH(ctx, slab_bbox_to_float_kernel, 1, 32, 0, d_bbox, d_bbf, d_bbf + 3);
  PFX_NCCL(api->AllReduce(d_bbf, d_bbf, 3, ncclFloat, ncclMin, G->comm, st));
  PFX_NCCL(api->AllReduce(d_bbf + 3, d_bbf + 3, 3, ncclFloat, ncclMax, G->comm, st));
  float bb[6];
  PFX_CUDA(cudaMemcpyAsync(bb, d_bbf, sizeof(bb), cudaMemcpyDeviceToHost, st));
  PFX_CUDA(cudaStreamSynchronize(st));
  int axis = 0;
  for (int a = 1; a < 3; ++a)
    if (bb[3 + a] - bb[a] > bb[3 + axis] - bb[axis]) axis = a;
  const float lo = bb[axis], hi = bb[3 + axis];
  const float width = (hi > lo) ? (hi - lo) / (float)SLAB_BINS : 1.f;

  // ---- cuts: equal-count quantiles of the all-reduced histogram along the longest axis (bin edges)
  std::vector<float> cuts(world + 1);
  cuts[0] = -INFINITY;
  cuts[world] = INFINITY;
  if (world > 1) {
    PFX_CUDA(cudaMemsetAsync(d_hist, 0, SLAB_BINS * sizeof(unsigned), st));
    if (n) PFX_LAUNCH(ctx, slab_hist_kernel, std::min(ctx->sm_count * 4, div_up(n, 256)), 256, 0, rows.as<float4>(), n, axis, lo, 1.0f / width, d_hist);
    PFX_NCCL(api->AllReduce(d_hist, d_hist, SLAB_BINS, ncclUint32, ncclSum, G->comm, st));
    std::vector<unsigned> hist(SLAB_BINS);
    PFX_CUDA(cudaMemcpyAsync(hist.data(), d_hist, SLAB_BINS * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    PFX_CUDA(cudaStreamSynchronize(st));
    unsigned long long tot = 0, run = 0;
    for (unsigned v : hist) tot += v;
    int b = 0;
    for (int r = 1; r < world; ++r) {
      const unsigned long long want = tot * (unsigned long long)r / (unsigned long long)world;
      while (b < SLAB_BINS && run + hist[b] <= want) run += hist[b++];
      cuts[r] = lo + (float)b * width;  // the same float expression on every rank
      if (cuts[r] < cuts[r - 1]) cuts[r] = cuts[r - 1];
    }
  }

  // ---- pack per peer (stable compaction), counts
  DevBuf& pack = ctx->slab_pack;
  std::vector<int> send_count(world, 0), send_off(world + 1, 0);
  PFX_CUDA(ctx->tmp1.ensure(std::max<size_t>(n, 1) * sizeof(int)));
  PFX_CUDA(ctx->tmp2.ensure(std::max<size_t>(n, 1) * sizeof(int)));
  int* flags = ctx->tmp1.as<int>();
  int* pos = ctx->tmp2.as<int>();
  // pass 1: counts per peer
  SlabCuts sc;
  sc.world = world;
  for (int p = 0; p <= world; ++p) sc.c[p] = cuts[p];
  PFX_CUDA(cudaMemsetAsync(d_counts, 0, (size_t)world * sizeof(int), st));
  if (n > 0)
    PFX_LAUNCH(ctx, slab_route_kernel<false>, div_up(n, 256), 256, 0, rows.as<float4>(), n, axis, sc, (float)halo, d_counts, nullptr);
  // ---- count matrix: row r = what rank r sends to each rank (my own row comes back with it: one host round trip)
  std::vector<int> matrix((size_t)world * world, 0);
  PFX_NCCL(api->AllGather(d_counts, d_matrix, (size_t)world, ncclInt32, G->comm, st));
  PFX_CUDA(cudaMemcpyAsync(matrix.data(), d_matrix, matrix.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
  PFX_CUDA(cudaStreamSynchronize(st));
  for (int p = 0; p < world; ++p) send_count[p] = matrix[(size_t)rank * world + p];
  for (int p = 0; p < world; ++p) send_off[p + 1] = send_off[p] + send_count[p];
  PFX_CUDA(pack.ensure(std::max<size_t>(send_off[world], 1) * sizeof(float4)));
  // pass 2: the rows, each peer's block starting at its offset
  if (n > 0) {
    int* d_cursor = d_counts + 32;  // (d_counts has room for 64 ints before the matrix; world <= 32)
    PFX_CUDA(cudaMemcpyAsync(d_cursor, send_off.data(), (size_t)world * sizeof(int), cudaMemcpyHostToDevice, st));
    PFX_LAUNCH(ctx, slab_route_kernel<true>, div_up(n, 256), 256, 0, rows.as<float4>(), n, axis, sc, (float)halo, d_cursor,
               pack.as<float4>());
  }

  std::vector<int> recv_off(world + 1, 0);
  for (int p = 0; p < world; ++p) recv_off[p + 1] = recv_off[p] + matrix[(size_t)p * world + rank];
  const int n_local = recv_off[world];

  // ---- payload: grouped send / recv, device to device (NVLink P2P), my own share by a device copy
  DevBuf& recv = ctx->slab_recv;
  PFX_CUDA(recv.ensure(std::max<size_t>(n_local, 1) * sizeof(float4)));
  PFX_NCCL(api->GroupStart());
  for (int p = 0; p < world; ++p) {
    if (p == rank) continue;
    if (send_count[p])
      PFX_NCCL(api->Send(pack.as<float4>() + send_off[p], (size_t)send_count[p] * 4, ncclFloat, p, G->comm, st));
    const int rc = matrix[(size_t)p * world + rank];
    if (rc) PFX_NCCL(api->Recv(recv.as<float4>() + recv_off[p], (size_t)rc * 4, ncclFloat, p, G->comm, st));
  }
  PFX_NCCL(api->GroupEnd());
  if (send_count[rank])
    PFX_CUDA(cudaMemcpyAsync(recv.as<float4>() + recv_off[rank], pack.as<float4>() + send_off[rank],
                             (size_t)send_count[rank] * sizeof(float4), cudaMemcpyDeviceToDevice, st));

  // ---- local order = ascending GLOBAL id, owned and halo points alike.  Every (d2, index) tie-break of the stages
  // (the k-th neighbour of a k-search, the median fallback of the SHOT frame) then resolves exactly as it does when
  // the whole cloud is described on one GPU, where the index IS the global id: sharded rows equal single-GPU rows bit
  // for bit.  The rows of this rank's own points are listed in ctx->slab_own (ascending).
  PFX_CUDA(ctx->surf.ensure(std::max<size_t>(n_local, 1) * sizeof(float4)));
  PFX_CUDA(ctx->slab_gid.ensure(std::max<size_t>(n_local, 1) * sizeof(int)));
  PFX_CUDA(ctx->slab_own.ensure(std::max<size_t>(n_local, 1) * sizeof(int)));
  int n_owned = 0;
  if (n_local > 0) {
    uint32_t* skeys = nullptr;
    int* svals = nullptr;
    PFX_TRY(sort_pairs_scratch(ctx, n_local, &skeys, &svals));
    PFX_LAUNCH(ctx, slab_keys_kernel, div_up(n_local, 256), 256, 0, recv.as<float4>(), n_local, skeys, svals);
    PFX_TRY(sort_pairs_scratch_run(ctx, n_local));
    PFX_CUDA(ctx->tmp1.ensure((size_t)n_local * sizeof(int)));
    PFX_CUDA(ctx->tmp2.ensure((size_t)n_local * sizeof(int)));
    flags = ctx->tmp1.as<int>();
    pos = ctx->tmp2.as<int>();
    int* d_total = d_counts + 40;
    PFX_LAUNCH(ctx, slab_commit_kernel, div_up(n_local, 256), 256, 0, recv.as<float4>(), svals, n_local, axis, cuts[rank],
               cuts[rank + 1], rank == 0 ? 1 : 0, ctx->surf.as<float4>(), ctx->slab_gid.as<int>(), flags);
    PFX_TRY(scan_exclusive_i32(ctx, flags, pos, n_local, d_total, ctx->scanbuf));
    PFX_LAUNCH(ctx, slab_own_list_kernel, div_up(n_local, 256), 256, 0, flags, pos, n_local, ctx->slab_own.as<int>());
    PFX_CUDA(cudaMemcpyAsync(&n_owned, d_total, sizeof(int), cudaMemcpyDeviceToHost, st));
    PFX_CUDA(cudaStreamSynchronize(st));
  }
  PFX_CUDA(cudaGetLastError());
  // ---- the context's surface is now owned + halo
  ctx->n = (size_t)n_local;
  ctx->surf_version = ++ctx->tick + (1ull << 32);
  ctx->have_normals = false;
  ctx->normals_sorted_for = nullptr;
  ctx->knn_grid = nullptr;
  ctx->q_is_surface = true;
  ctx->nq = 0;
  ctx->slab_active = true;
  ctx->slab_axis = axis;
  ctx->slab_owned = (size_t)n_owned;
  ctx->slab_total = n_total;
  ctx->slab_lo = cuts[rank];
  ctx->slab_hi = cuts[rank + 1];
  if (n_owned_out) *n_owned_out = (size_t)n_owned;
  if (n_local_out) *n_local_out = (size_t)n_local;
  return 0;
}

// local row numbers (ascending) of this rank's own points, int32 [n_owned]
extern "C" int pfx_slab_owned_rows(pfx_ctx* ctx, int32_t* out, int mem) {
  if (!ctx) return PFX_E_INVALID;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return ctx->fail(PFX_E_STATE, "cudaSetDevice failed");
  if (!ctx->slab_active) return ctx->fail(PFX_E_STATE, "pfx_slab_owned_rows: no slab surface (pfx_slab_distribute)");
  if (!out || (mem != PFX_HOST && mem != PFX_DEVICE)) return ctx->fail(PFX_E_INVALID, "pfx_slab_owned_rows: bad arguments");
  if (ctx->slab_owned == 0) return 0;
  PFX_CUDA(cudaMemcpyAsync(out, ctx->slab_own.p, ctx->slab_owned * sizeof(int),
                           mem == PFX_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, ctx->stream));
  if (mem == PFX_HOST) PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// global ids of the local surface points (ascending), int32 [n_local]
extern "C" int pfx_slab_global_ids(pfx_ctx* ctx, int32_t* out, int mem) {
  if (!ctx) return PFX_E_INVALID;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return ctx->fail(PFX_E_STATE, "cudaSetDevice failed");
  if (!ctx->slab_active) return ctx->fail(PFX_E_STATE, "pfx_slab_global_ids: no slab surface (pfx_slab_distribute)");
  if (!out || (mem != PFX_HOST && mem != PFX_DEVICE)) return ctx->fail(PFX_E_INVALID, "pfx_slab_global_ids: bad arguments");
  if (ctx->n == 0) return 0;
  PFX_CUDA(cudaMemcpyAsync(out, ctx->slab_gid.p, ctx->n * sizeof(int),
                           mem == PFX_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, ctx->stream));
  if (mem == PFX_HOST) PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// info[0] axis, [1] n_owned, [2] n_local, [3] total points over all ranks, [4] slab lower bound, [5] slab upper bound
extern "C" int pfx_slab_info(const pfx_ctx* ctx, double* info6) {
  if (!ctx || !info6) return PFX_E_INVALID;
  if (!ctx->slab_active) return PFX_E_STATE;
  info6[0] = ctx->slab_axis;
  info6[1] = (double)ctx->slab_owned;
  info6[2] = (double)ctx->n;
  info6[3] = (double)ctx->slab_total;
  info6[4] = ctx->slab_lo;
  info6[5] = ctx->slab_hi;
  return 0;
}

// ================================================================================== sharded matching
// a: this rank's na query rows; b: this rank's nb target rows, whose global row numbers start at b_offset.
// Every rank calls with its own shards; nn_idx (GLOBAL target row, -1 = none) and nn_d2 are sized na.  Device memory.
extern "C" int pfx_match_ring(pfx_ctx* ctx, const float* a, size_t na, size_t stride_a, const float* b, size_t nb,
                              size_t stride_b, int dim, int b_offset, int32_t* nn_idx, float* nn_d2, int mem) {
  if (!ctx) return PFX_E_INVALID;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return ctx->fail(PFX_E_STATE, "cudaSetDevice failed");
  PFX_TRY(need_group(ctx, "pfx_match_ring"));
  if (mem != PFX_DEVICE) return ctx->fail(PFX_E_INVALID, "pfx_match_ring: device buffers only (PFX_DEVICE)");
  if (dim <= 0 || (na && (!a || !nn_idx || !nn_d2)) || (nb && !b) || stride_a < (size_t)dim * 4 || stride_b < (size_t)dim * 4 ||
      (stride_a & 3) || (stride_b & 3) || na > 0x7fffffffull || nb > 0x7fffffffull || b_offset < 0)
    return ctx->fail(PFX_E_INVALID, "pfx_match_ring: bad arguments");
  Group* G = ctx->group;
  NcclApi* api = nccl_api();
  const int world = G->world, rank = G->rank;
  cudaStream_t st = ctx->stream;
  // block sizes and offsets of every rank
  PFX_CUDA(ctx->grp_tmp.ensure(4096));
  int* d_pair = ctx->grp_tmp.as<int>();
  int mine[2] = {(int)nb, b_offset};
  std::vector<int> all(2 * (size_t)world);
  PFX_CUDA(cudaMemcpyAsync(d_pair, mine, sizeof(mine), cudaMemcpyHostToDevice, st));
  PFX_NCCL(api->AllGather(d_pair, d_pair + 2, 2, ncclInt32, G->comm, st));
  PFX_CUDA(cudaMemcpyAsync(all.data(), d_pair + 2, all.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
  PFX_CUDA(cudaStreamSynchronize(st));
  int max_nb = 0;
  for (int r = 0; r < world; ++r) max_nb = std::max(max_nb, all[2 * r]);
  // two rotating block buffers (dense rows of dim floats) + this rank's block packed densely
  const size_t blk = (size_t)std::max(max_nb, 1) * dim;
  PFX_CUDA(ctx->ring_buf[0].ensure(blk * sizeof(float)));
  PFX_CUDA(ctx->ring_buf[1].ensure(blk * sizeof(float)));
  PFX_CUDA(ctx->ring_best.ensure(std::max<size_t>(na, 1) * sizeof(unsigned long long)));
  PFX_CUDA(ctx->ring_res.ensure(std::max<size_t>(na, 1) * (sizeof(int) + sizeof(float))));
  unsigned long long* best = ctx->ring_best.as<unsigned long long>();
  int* ridx = ctx->ring_res.as<int>();
  float* rd2 = reinterpret_cast<float*>(ridx + std::max<size_t>(na, 1));
  if (nb) PFX_CUDA(cudaMemcpy2DAsync(ctx->ring_buf[0].p, (size_t)dim * 4, b, stride_b, (size_t)dim * 4, nb, cudaMemcpyDeviceToDevice, st));
  if (na) PFX_LAUNCH(ctx, ring_init_kernel, div_up((long long)na, 256), 256, 0, best, (int)na);
  const int left = (rank + world - 1) % world, right = (rank + 1) % world;
  int cur = 0;
  for (int step = 0; step < world; ++step) {
    const int owner = (rank + step) % world;  // whose block sits in buffer `cur`
    const int cur_n = all[2 * owner], cur_off = all[2 * owner + 1];
    if (step < world - 1) {
      // the block moves on to the left neighbour, the next one arrives from the right, on the communication stream,
      // while this stream matches the current block (ncclSend / ncclRecv over NVLink)
      const int next_owner = (rank + step + 1) % world;
      const int next_n = all[2 * next_owner];
      PFX_CUDA(cudaEventRecord(G->ev_a, st));  // buffer `cur` is complete, buffer `cur ^ 1` is no longer being read
      PFX_CUDA(cudaStreamWaitEvent(G->comm_stream, G->ev_a, 0));
      PFX_NCCL(api->GroupStart());
      if (cur_n) PFX_NCCL(api->Send(ctx->ring_buf[cur].p, (size_t)cur_n * dim, ncclFloat, left, G->comm, G->comm_stream));
      if (next_n) PFX_NCCL(api->Recv(ctx->ring_buf[cur ^ 1].p, (size_t)next_n * dim, ncclFloat, right, G->comm, G->comm_stream));
      PFX_NCCL(api->GroupEnd());
      PFX_CUDA(cudaEventRecord(G->ev_b, G->comm_stream));
    }
    if (na && cur_n) {
      PFX_TRY(match_dispatch_dev(ctx, a, (int)na, (int)(stride_a / 4), ctx->ring_buf[cur].as<float>(), cur_n, dim, dim, ridx, rd2));
      PFX_LAUNCH(ctx, ring_merge_kernel, div_up((long long)na, 256), 256, 0, best, ridx, rd2, (int)na, cur_off);
    }
    if (step < world - 1) {
      PFX_CUDA(cudaStreamWaitEvent(st, G->ev_b, 0));
      cur ^= 1;
    }
  }
  if (na) PFX_LAUNCH(ctx, ring_unpack_kernel, div_up((long long)na, 256), 256, 0, best, (int)na, nn_idx, nn_d2);
  PFX_CUDA(cudaGetLastError());
  return 0;
}
