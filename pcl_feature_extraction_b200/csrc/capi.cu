// capi.cu — the C ABI declared in include/pfx_b200.h: context, host<->device staging, layout
// conversion (strided AoS records <-> the float4 SoA rows the kernels use) and stage sequencing.
#include <cstring>

#include <cstdlib>

#include "internal.h"
#include "seqsum.h"

using namespace pfx;

namespace pfx {

// ------------------------------------------------------------------------------ layout kernels
__global__ void aos_to_float4_kernel(const unsigned char* __restrict__ src, size_t stride, int n,
                                     float4* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = reinterpret_cast<const float*>(src + (size_t)i * stride);
  dst[i] = make_float4(p[0], p[1], p[2], __int_as_float(i));
}

__global__ void normals_in_kernel(const unsigned char* __restrict__ src, size_t stride, int curv_off, int n,
                                  float4* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = reinterpret_cast<const float*>(src + (size_t)i * stride);
  dst[i] = make_float4(p[0], p[1], p[2], p[curv_off]);
}

__global__ void normals_out_kernel(const float4* __restrict__ src, int n, unsigned char* __restrict__ dst,
                                   size_t stride, int curv_off) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* p = reinterpret_cast<float*>(dst + (size_t)i * stride);
  float4 v = src[i];
  p[0] = v.x;
  p[1] = v.y;
  p[2] = v.z;
  p[curv_off] = v.w;
}

__global__ void permute_rows_kernel(const float4* __restrict__ src_orig, const float4* __restrict__ sorted_pts,
                                    int n, float4* __restrict__ dst_sorted) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst_sorted[i] = src_orig[__float_as_int(sorted_pts[i].w)];
}

__global__ void gather_xyz_kernel(const float4* __restrict__ surf, const int* __restrict__ idx, int m,
                                  float* __restrict__ xyz) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  float4 p = surf[idx[i]];
  xyz[3 * i] = p.x;
  xyz[3 * i + 1] = p.y;
  xyz[3 * i + 2] = p.z;
}

// surface normals in the sorted order of grid g (cached)
int normals_sorted_for_grid(Ctx* ctx, Grid* g, const float4** out) {
  if (!ctx->have_normals) return ctx->fail(PFX_E_STATE, "surface normals are not set (setInputNormals)");
  if (ctx->normals_sorted_for == g && ctx->normals_sorted_version == ctx->normals_version &&
      g->surf_version == ctx->surf_version) {
    *out = ctx->normals_sorted.as<float4>();
    return 0;
  }
  const int n = (int)ctx->n;
  PFX_CUDA(ctx->normals_sorted.ensure(std::max<size_t>(n, 1) * sizeof(float4)));
  if (n > 0)
    PFX_LAUNCH(ctx, permute_rows_kernel, div_up(n, 256), 256, 0, ctx->normals.as<float4>(), g->pts.as<float4>(), n,
               ctx->normals_sorted.as<float4>());
  PFX_CUDA(cudaGetLastError());
  ctx->normals_sorted_for = g;
  ctx->normals_sorted_version = ctx->normals_version;
  *out = ctx->normals_sorted.as<float4>();
  return 0;
}

static int upload_records(Ctx* ctx, const void* src, size_t n, size_t stride, int mem, DevBuf& stage,
                          const unsigned char** dev_src) {
  if (mem == PFX_DEVICE) {
    *dev_src = static_cast<const unsigned char*>(src);
    return 0;
  }
  PFX_CUDA(stage.ensure(std::max<size_t>(n * stride, 16)));
  if (n) PFX_CUDA(cudaMemcpyAsync(stage.p, src, n * stride, cudaMemcpyHostToDevice, ctx->stream));
  *dev_src = stage.as<unsigned char>();
  return 0;
}

// copy a device result to the caller (host: synchronous at return)
static int deliver(Ctx* ctx, void* dst, const void* dev_src, size_t bytes, int mem) {
  if (bytes == 0) return 0;
  if (mem == PFX_DEVICE) {
    if (dst != dev_src)
      PFX_CUDA(cudaMemcpyAsync(dst, dev_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
  }
  PFX_CUDA(cudaMemcpyAsync(dst, dev_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ---- PFX_HOST_ASYNC plumbing
static int async_init(Ctx* ctx) {
  if (ctx->copy_stream) return 0;
  PFX_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (int s = 0; s < 2; ++s) {
    PFX_CUDA(cudaEventCreateWithFlags(&ctx->ev_ready[s], cudaEventDisableTiming));
    PFX_CUDA(cudaEventCreateWithFlags(&ctx->ev_copied[s], cudaEventDisableTiming));
  }
  return 0;
}
// staging buffer of slot s, safe to overwrite: the compute stream first waits for the slot's previous copy
static int async_stage_acquire(Ctx* ctx, int s, size_t bytes, void** out) {
  PFX_TRY(async_init(ctx));
  if (ctx->copy_pending[s]) {
    if (bytes > ctx->async_stage[s].cap) PFX_CUDA(cudaEventSynchronize(ctx->ev_copied[s]));  // about to be reallocated
    PFX_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[s], 0));
  }
  PFX_CUDA(ctx->async_stage[s].ensure(std::max<size_t>(bytes, 16)));
  *out = ctx->async_stage[s].p;
  return 0;
}
// enqueue the device-to-host copy of slot s behind everything issued so far on the compute stream
static int async_deliver(Ctx* ctx, int s, void* dst, size_t bytes) {
  PFX_CUDA(cudaEventRecord(ctx->ev_ready[s], ctx->stream));
  PFX_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_ready[s], 0));
  if (bytes) PFX_CUDA(cudaMemcpyAsync(dst, ctx->async_stage[s].p, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
  PFX_CUDA(cudaEventRecord(ctx->ev_copied[s], ctx->copy_stream));
  ctx->copy_pending[s] = true;
  return 0;
}

// fingerprint of a host buffer of n records: every record up to 262144 of them (the reference's clouds: an edit in
// place cannot go unnoticed), 8192 evenly spaced ones beyond
static Ctx::HostFp host_fingerprint(const void* p, size_t n, size_t stride, size_t rec_bytes) {
  Ctx::HostFp f;
  f.ptr = p;
  f.n = n;
  f.stride = stride;
  uint64_t h = 0x9E3779B97F4A7C15ull ^ (uint64_t)n ^ ((uint64_t)stride << 40);
  const unsigned char* b = static_cast<const unsigned char*>(p);
  const size_t samples = n <= 262144 ? n : 8192;
  const size_t words = rec_bytes / 4;
  for (size_t s = 0; s < samples; ++s) {
    const size_t i = (samples == n) ? s : (size_t)(((unsigned __int128)s * (n - 1)) / (samples - 1));
    const uint32_t* w = reinterpret_cast<const uint32_t*>(b + i * stride);
    for (size_t t = 0; t < words; ++t) {
      h ^= w[t];
      h *= 0xBF58476D1CE4E5B9ull;
      h ^= h >> 29;
    }
  }
  f.hash = h;
  f.valid = true;
  return f;
}
static bool same_fp(const Ctx::HostFp& a, const Ctx::HostFp& b) {
  return a.valid && b.valid && a.ptr == b.ptr && a.n == b.n && a.stride == b.stride && a.hash == b.hash;
}

static int check_ctx(pfx_ctx* ctx) {
  if (!ctx) return PFX_E_INVALID;
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e != cudaSuccess) return ctx->fail_cuda(e, "cudaSetDevice", __FILE__, __LINE__);
  return 0;
}

static int check_search_params(Ctx* ctx, double radius, int k, const char* who) {
  // pcl::Feature::initCompute: exactly one of radius / k
  if (ctx->surf_version == 0) return ctx->fail(PFX_E_PRECOND, std::string(who) + ": no surface set");
  if (radius > 0 && k > 0) return ctx->fail(PFX_E_PRECOND, std::string(who) + ": both radius and K defined");
  if (!(radius > 0) && k <= 0) return ctx->fail(PFX_E_PRECOND, std::string(who) + ": neither radius nor K defined");
  if (radius < 0 || k < 0) return ctx->fail(PFX_E_INVALID, std::string(who) + ": negative search parameter");
  return 0;
}

}  // namespace pfx

// ================================================================================== context
extern "C" int pfx_version(void) { return 100; }

extern "C" int pfx_create(int device, pfx_ctx** out) {
  if (!out) return PFX_E_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess) return (int)e;
  if (device < 0 || device >= count) return PFX_E_INVALID;
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return (int)e;
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return (int)e;
  if (prop.major < 10) return (int)cudaErrorNoKernelImageForDevice;  // sm_100a only, no fallback
  pfx_ctx* c = new pfx_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  if (const char* e = getenv("PFX_SHOT_ROWS")) c->shot_from_rows = atoi(e) != 0;
  if (const char* e = getenv("PFX_TC_PAIR")) c->tc_pair = atoi(e);
  *out = c;
  return 0;
}

extern "C" int pfx_destroy(pfx_ctx* ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->aux_stream) {
    cudaStreamSynchronize(ctx->aux_stream);
    cudaStreamDestroy(ctx->aux_stream);
    cudaEventDestroy(ctx->ev_surface);
  }
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
    for (int s = 0; s < 2; ++s) {
      cudaEventDestroy(ctx->ev_ready[s]);
      cudaEventDestroy(ctx->ev_copied[s]);
      ctx->async_stage[s].release();
    }
  }
  group_release(ctx);
  grid_free_all(ctx);
  match_tc_release(ctx);
  for (DevBuf* b : {&ctx->surf, &ctx->normals, &ctx->normals_sorted, &ctx->qry, &ctx->knn_idx, &ctx->knn_d2,
                    &ctx->stage, &ctx->stage2, &ctx->tmp0, &ctx->tmp1, &ctx->tmp2, &ctx->tmp3, &ctx->tmp4,
                    &ctx->small, &ctx->scanbuf, &ctx->match_flags, &ctx->match_best, &ctx->out_stage, &ctx->qflag,
                    &ctx->worklist, &ctx->worklist2, &ctx->ri_img, &ctx->nb_surf, &ctx->nb_scores, &ctx->nb_shadow,
                    &ctx->nb_traits, &ctx->nb_dir, &ctx->nb_change, &ctx->nk_interest, &ctx->icp_state, &ctx->icp_cur,
                    &ctx->icp_nn, &ctx->icp_partials, &ctx->usc_tab, &ctx->lab_tab, &ctx->surf_lab, &ctx->qry_lab,
                    &ctx->st_cnt, &ctx->st_off, &ctx->st_idx, &ctx->st_d2, &ctx->grp_tmp, &ctx->slab_rows, &ctx->slab_pack,
                    &ctx->slab_recv, &ctx->slab_gid, &ctx->ring_buf[0], &ctx->ring_buf[1], &ctx->ring_best, &ctx->ring_res})
    b->release();
  ctx->vg_scratch.release();
  for (Ctx::ProfRec& r : ctx->prof_recs) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->rows_stat.host) {
    cudaFreeHost(ctx->rows_stat.host);
    cudaEventDestroy(ctx->rows_stat.ev);
  }
  delete ctx;
  return 0;
}

extern "C" const char* pfx_last_error(const pfx_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int pfx_set_stream(pfx_ctx* ctx, void* s) {
  if (!ctx) return PFX_E_INVALID;
  ctx->stream = static_cast<cudaStream_t>(s);
  return 0;
}

extern "C" int pfx_prepare_radius(pfx_ctx* ctx, double radius) {
  PFX_TRY(check_ctx(ctx));
  if (ctx->surf_version == 0) return ctx->fail(PFX_E_PRECOND, "pfx_prepare_radius: no surface set");
  if (!(radius > 0)) return ctx->fail(PFX_E_INVALID, "pfx_prepare_radius: radius must be positive");
  return grid_prepare_async(ctx, radius);
}

extern "C" int pfx_sync(pfx_ctx* ctx) {
  PFX_TRY(check_ctx(ctx));
  if (ctx->aux_stream) PFX_CUDA(cudaStreamSynchronize(ctx->aux_stream));
  PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->copy_stream) {
    PFX_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    ctx->copy_pending[0] = ctx->copy_pending[1] = false;
  }
  return 0;
}

extern "C" uint64_t pfx_launch_count(const pfx_ctx* ctx) { return ctx ? ctx->launches : 0; }

// Diagnostics of the voxel hash used by the last call: out[0] cell edge, out[1..3] cells per axis,
// out[4] occupied cells, out[5] finite points, out[6] queries the cell-tile path handed to the generic
// kernels (-1 when no tile pass ran on that grid), out[7] reserved.
extern "C" int pfx_grid_info(pfx_ctx* ctx, double* out8) {
  PFX_TRY(check_ctx(ctx));
  if (!out8) return ctx->fail(PFX_E_INVALID, "pfx_grid_info: null output");
  for (int i = 0; i < 8; ++i) out8[i] = 0;
  Grid* g = ctx->last_grid;
  if (!g || g->surf_version != ctx->surf_version) return ctx->fail(PFX_E_STATE, "pfx_grid_info: no grid built yet");
  GridParams P;
  PFX_CUDA(cudaMemcpyAsync(&P, g->params.p, sizeof(P), cudaMemcpyDeviceToHost, ctx->stream));
  PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  out8[0] = P.edge; out8[1] = P.nx; out8[2] = P.ny; out8[3] = P.nz; out8[4] = P.ncells; out8[5] = P.n_valid;
  out8[6] = -1;
  if (ctx->knn_grid == g && ctx->knn_sversion == ctx->surf_version && ctx->knn_dense && !ctx->knn_sorted && ctx->n > 0) {
    std::vector<unsigned char> f(ctx->n);
    PFX_CUDA(cudaMemcpy(f.data(), ctx->qflag.p, ctx->n, cudaMemcpyDeviceToHost));
    long long c = 0;
    for (unsigned char v : f) c += v ? 1 : 0;
    out8[6] = (double)c;
  }
  return 0;
}

// target points per occupied cell of kNN grids, as a fraction of k (default 0.4)
extern "C" int pfx_set_knn_occupancy(pfx_ctx* ctx, float fraction_of_k) {
  if (!ctx || !(fraction_of_k > 0.f)) return PFX_E_INVALID;
  ctx->knn_occupancy = fraction_of_k;
  return 0;
}

extern "C" int pfx_profile_begin(pfx_ctx* ctx, const char* filter) {
  if (!ctx) return PFX_E_INVALID;
  ctx->prof_filter = filter ? filter : "";
  ctx->prof_used = 0;
  ctx->prof_on = true;
  return 0;
}

// Stops profiling, waits for the stream and writes one line per kernel: "<name>\t<launches>\t<total ms>\n"
extern "C" int pfx_profile_end(pfx_ctx* ctx, char* buf, size_t buflen) {
  PFX_TRY(check_ctx(ctx));
  ctx->prof_on = false;
  if (ctx->aux_stream) PFX_CUDA(cudaStreamSynchronize(ctx->aux_stream));
  PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<std::pair<std::string, std::pair<int, double>>> agg;
  for (size_t i = 0; i < ctx->prof_used; ++i) {
    float ms = 0.f;
    PFX_CUDA(cudaEventElapsedTime(&ms, ctx->prof_recs[i].e0, ctx->prof_recs[i].e1));
    std::string nm = ctx->prof_recs[i].name;
    bool found = false;
    for (auto& a : agg)
      if (a.first == nm) {
        a.second.first++;
        a.second.second += ms;
        found = true;
        break;
      }
    if (!found) agg.push_back({nm, {1, (double)ms}});
  }
  ctx->prof_used = 0;
  std::string out;
  for (auto& a : agg) {
    char line[256];
    snprintf(line, sizeof(line), "%s\t%d\t%.6f\n", a.first.c_str(), a.second.first, a.second.second);
    out += line;
  }
  if (buf && buflen) {
    size_t m = std::min(buflen - 1, out.size());
    memcpy(buf, out.data(), m);
    buf[m] = 0;
  }
  return 0;
}
extern "C" size_t pfx_num_surface(const pfx_ctx* ctx) { return ctx ? ctx->n : 0; }

namespace pfx {
__global__ void surface_out_kernel(const float4* __restrict__ surf, int n, unsigned char* __restrict__ out, size_t stride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = surf[i];
  float* o = reinterpret_cast<float*>(out + (size_t)i * stride);
  o[0] = p.x;
  o[1] = p.y;
  o[2] = p.z;
  if (stride >= 16) o[3] = 0.f;
}
}  // namespace pfx

// the points of the current surface in its own order (after pfx_slab_distribute: owned points first, then the halo)
extern "C" int pfx_get_surface(pfx_ctx* ctx, void* out, size_t stride, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (ctx->surf_version == 0) return ctx->fail(PFX_E_PRECOND, "pfx_get_surface: no surface set");
  if (!out || stride < 12 || (stride & 3) || (mem != PFX_HOST && mem != PFX_DEVICE))
    return ctx->fail(PFX_E_INVALID, "pfx_get_surface: bad output / stride / mem");
  const size_t n = ctx->n;
  if (n == 0) return 0;
  unsigned char* dout = static_cast<unsigned char*>(out);
  if (mem == PFX_HOST) {
    PFX_CUDA(ctx->out_stage.ensure(n * stride));
    dout = ctx->out_stage.as<unsigned char>();
    if (stride > 16) PFX_CUDA(cudaMemsetAsync(dout, 0, n * stride, ctx->stream));
  }
  PFX_LAUNCH(ctx, surface_out_kernel, div_up((long long)n, 256), 256, 0, ctx->surf.as<float4>(), (int)n, dout, stride);
  PFX_CUDA(cudaGetLastError());
  if (mem == PFX_HOST) return deliver(ctx, out, dout, n * stride, mem);
  return 0;
}
extern "C" size_t pfx_num_queries(const pfx_ctx* ctx) { return ctx ? ctx->num_queries() : 0; }

// ================================================================================== inputs
extern "C" int pfx_set_surface(pfx_ctx* ctx, const void* pts, size_t n, size_t stride, int mem) {
  PFX_TRY(check_ctx(ctx));
  if ((n && !pts) || stride < 12 || (stride & 3) || n > 0x7fffffffull)
    return ctx->fail(PFX_E_INVALID, "pfx_set_surface: bad pointer / stride / size");
  Ctx::HostFp fp;
  if (mem == PFX_HOST && ctx->reuse && n > 0) {
    // the same host cloud announced again (the reference does so for every Feature object, features.h:190-193):
    // the device copy, its voxel hashes, kNN lists and normals stay
    fp = host_fingerprint(pts, n, stride, 12);
    if (ctx->surf_version != 0 && !ctx->slab_active && same_fp(fp, ctx->surf_fp)) {
      // the visible state is that of a fresh surface (queries = the surface, no input normals); the normals stay
      // RESIDENT and come back into force when pfx_normals asks for the same ones or pfx_set_surface_normals hands in
      // the buffer they were delivered to
      ctx->stat_surface_reused++;
      ctx->q_is_surface = true;
      ctx->nq = 0;
      ctx->qry_version++;
      ctx->have_normals = false;
      ctx->surf_lab_version = 0;  // colours are cheap to hand in again: a fresh surface has none
      ctx->surf_rgb_version = 0;
      return 0;
    }
  }
  ctx->surf_fp = fp;
  ctx->nrm_fp.valid = false;
  ctx->stat_surface_uploads++;
  const unsigned char* src = nullptr;
  PFX_TRY(grid_wait_pending(ctx));  // a build in flight on the auxiliary stream still reads the old surface
  PFX_TRY(upload_records(ctx, pts, n, stride, mem, ctx->stage, &src));
  PFX_CUDA(ctx->surf.ensure(std::max<size_t>(n, 1) * sizeof(float4)));
  if (n) PFX_LAUNCH(ctx, aos_to_float4_kernel, div_up((long long)n, 256), 256, 0, src, stride, (int)n, ctx->surf.as<float4>());
  PFX_CUDA(cudaGetLastError());
  ctx->n = n;
  ctx->surf_version = ++ctx->tick + (1ull << 32);
  ctx->have_normals = false;
  ctx->normals_sorted_for = nullptr;
  ctx->knn_grid = nullptr;
  ctx->q_is_surface = true;
  ctx->nq = 0;
  ctx->slab_active = false;
  return 0;
}

extern "C" int pfx_set_queries(pfx_ctx* ctx, const void* pts, size_t n, size_t stride, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (n == 0 || !pts) {
    ctx->q_is_surface = true;
    ctx->nq = 0;
    ctx->qry_version++;
    return 0;
  }
  if (stride < 12 || (stride & 3) || n > 0x7fffffffull) return ctx->fail(PFX_E_INVALID, "pfx_set_queries: bad stride / size");
  const unsigned char* src = nullptr;
  PFX_TRY(upload_records(ctx, pts, n, stride, mem, ctx->stage, &src));
  PFX_CUDA(ctx->qry.ensure(n * sizeof(float4)));
  PFX_LAUNCH(ctx, aos_to_float4_kernel, div_up((long long)n, 256), 256, 0, src, stride, (int)n, ctx->qry.as<float4>());
  PFX_CUDA(cudaGetLastError());
  ctx->nq = n;
  ctx->q_is_surface = false;
  ctx->qry_version++;
  return 0;
}

extern "C" int pfx_set_surface_normals(pfx_ctx* ctx, const void* normals, size_t n, size_t stride, int curv_off,
                                       int mem) {
  PFX_TRY(check_ctx(ctx));
  if (ctx->surf_version == 0) return ctx->fail(PFX_E_PRECOND, "pfx_set_surface_normals: no surface set");
  // FeatureFromNormals::initCompute: normals.size() must equal surface.size()
  if (n != ctx->n) return ctx->fail(PFX_E_PRECOND, "pfx_set_surface_normals: the number of normals differs from the surface size");
  if (stride < 16 || (stride & 3) || curv_off < 3 || (size_t)(curv_off + 1) * 4 > stride)
    return ctx->fail(PFX_E_INVALID, "pfx_set_surface_normals: bad stride / curvature offset");
  Ctx::HostFp fp;
  if (mem == PFX_HOST && ctx->reuse && n > 0) {
    // normals the library itself delivered into this very buffer (pfx_normals), or uploaded from it before, for this
    // surface: the device copy is current (features.h:187-188 hands the normals of the whole cloud back for every
    // descriptor type)
    fp = host_fingerprint(normals, n, stride, ((size_t)curv_off + 1) * 4);
    if (ctx->nrm_fp_surf == ctx->surf_version && ctx->nrm_fp_version == ctx->normals_version && same_fp(fp, ctx->nrm_fp)) {
      ctx->stat_normals_upload_skipped++;
      ctx->have_normals = true;
      return 0;
    }
  }
  ctx->stat_normals_uploads++;
  const unsigned char* src = nullptr;
  PFX_TRY(upload_records(ctx, normals, n, stride, mem, ctx->stage, &src));
  PFX_CUDA(ctx->normals.ensure(std::max<size_t>(n, 1) * sizeof(float4)));
  if (n) PFX_LAUNCH(ctx, normals_in_kernel, div_up((long long)n, 256), 256, 0, src, stride, curv_off, (int)n, ctx->normals.as<float4>());
  PFX_CUDA(cudaGetLastError());
  ctx->have_normals = true;
  ctx->normals_version++;
  ctx->nrm_fp = fp;
  ctx->nrm_fp_surf = ctx->surf_version;
  ctx->nrm_fp_version = ctx->normals_version;
  return 0;
}

extern "C" int pfx_set_viewpoint(pfx_ctx* ctx, float vx, float vy, float vz) {
  if (!ctx) return PFX_E_INVALID;
  ctx->vp[0] = vx;
  ctx->vp[1] = vy;
  ctx->vp[2] = vz;
  return 0;
}

// ================================================================================== search
extern "C" int pfx_knn(pfx_ctx* ctx, int k, int32_t* idx, float* d2, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (ctx->surf_version == 0) return ctx->fail(PFX_E_PRECOND, "pfx_knn: no surface set");
  if (k < 1 || k > 32) return ctx->fail(PFX_E_INVALID, "pfx_knn: k must be in [1, 32]");
  if (!idx || !d2) return ctx->fail(PFX_E_INVALID, "pfx_knn: null output");
  const size_t nq = ctx->num_queries();
  if (nq == 0) return 0;
  Grid* g = nullptr;
  PFX_TRY(grid_get(ctx, 0.0, k, &g));
  PFX_TRY(knn_lists(ctx, g, k, true));
  int32_t* didx = idx;
  float* dd2 = d2;
  if (mem == PFX_HOST) {
    PFX_CUDA(ctx->out_stage.ensure(nq * k * (sizeof(int) + sizeof(float))));
    didx = ctx->out_stage.as<int32_t>();
    dd2 = reinterpret_cast<float*>(didx + nq * k);
  }
  PFX_TRY(knn_export(ctx, k, didx, dd2, mem));
  if (mem == PFX_HOST) {
    PFX_CUDA(cudaMemcpyAsync(idx, didx, nq * k * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PFX_TRY(deliver(ctx, d2, dd2, nq * k * sizeof(float), mem));
  }
  return 0;
}

extern "C" int pfx_radius_count(pfx_ctx* ctx, double radius, int32_t* counts, int64_t* total, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (ctx->surf_version == 0) return ctx->fail(PFX_E_PRECOND, "pfx_radius_count: no surface set");
  if (!(radius > 0)) return ctx->fail(PFX_E_INVALID, "pfx_radius_count: radius must be > 0");
  const size_t nq = ctx->num_queries();
  if (total) *total = 0;
  if (nq == 0) return 0;
  Grid* g = nullptr;
  PFX_TRY(grid_get(ctx, radius, 0, &g));
  int* dcounts = counts;
  if (mem == PFX_HOST || !counts) {
    PFX_CUDA(ctx->out_stage.ensure(nq * sizeof(int)));
    dcounts = ctx->out_stage.as<int>();
  }
  PFX_TRY(radius_count(ctx, g, radius, dcounts));
  if (total) {
    PFX_CUDA(ctx->tmp4.ensure((nq + 1) * sizeof(long long)));
    PFX_TRY(scan_exclusive_i64(ctx, dcounts, ctx->tmp4.as<long long>(), (int)nq, ctx->scanbuf));
    long long t = 0;
    PFX_CUDA(cudaMemcpyAsync(&t, ctx->tmp4.as<long long>() + nq, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    PFX_CUDA(cudaStreamSynchronize(ctx->stream));
    *total = t;
  }
  if (mem == PFX_HOST && counts) PFX_TRY(deliver(ctx, counts, dcounts, nq * sizeof(int), mem));
  return 0;
}

extern "C" int pfx_radius_search(pfx_ctx* ctx, double radius, int sorted, const int64_t* offsets, int32_t* idx,
                                 float* d2, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (ctx->surf_version == 0) return ctx->fail(PFX_E_PRECOND, "pfx_radius_search: no surface set");
  if (!(radius > 0) || !offsets) return ctx->fail(PFX_E_INVALID, "pfx_radius_search: bad arguments");
  const size_t nq = ctx->num_queries();
  if (nq == 0) return 0;
  Grid* g = nullptr;
  PFX_TRY(grid_get(ctx, radius, 0, &g));
  const long long* doff = reinterpret_cast<const long long*>(offsets);
  int* didx = idx;
  float* dd2 = d2;
  long long total = 0;
  if (mem == PFX_HOST) {
    total = offsets[nq];
    PFX_CUDA(ctx->tmp4.ensure((nq + 1) * sizeof(long long)));
    PFX_CUDA(cudaMemcpyAsync(ctx->tmp4.p, offsets, (nq + 1) * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    doff = ctx->tmp4.as<long long>();
    PFX_CUDA(ctx->out_stage.ensure(std::max<size_t>((size_t)total, 1) * (sizeof(int) + sizeof(float))));
    didx = ctx->out_stage.as<int>();
    dd2 = reinterpret_cast<float*>(didx + total);
  }
  PFX_TRY(radius_fill(ctx, g, radius, sorted, doff, didx, dd2));
  if (mem == PFX_HOST && total > 0) {
    PFX_CUDA(cudaMemcpyAsync(idx, didx, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PFX_TRY(deliver(ctx, d2, dd2, (size_t)total * sizeof(float), mem));
  }
  return 0;
}

// ================================================================================== normals
extern "C" int pfx_normals(pfx_ctx* ctx, double radius, int k, void* out, size_t stride, int curv_off, int mem) {
  PFX_TRY(check_ctx(ctx));
  PFX_TRY(check_search_params(ctx, radius, k, "pfx_normals"));
  if (k > 32) return ctx->fail(PFX_E_INVALID, "pfx_normals: k must be <= 32");
  if (out && (stride < 16 || (stride & 3) || curv_off < 3 || (size_t)(curv_off + 1) * 4 > stride))
    return ctx->fail(PFX_E_INVALID, "pfx_normals: bad stride / curvature offset");
  const size_t nq = ctx->num_queries();
  Grid* g = nullptr;
  PFX_TRY(grid_get(ctx, radius > 0 ? radius : 0.0, k, &g));
  float4* rows = nullptr;
  if (out || !ctx->q_is_surface) {
    PFX_CUDA(ctx->tmp0.ensure(std::max<size_t>(nq, 1) * sizeof(float4)));
    rows = ctx->tmp0.as<float4>();
  }
  if (!out && !ctx->q_is_surface) return 0;
  // dense normals computed before with the same parameters for this surface are still resident: the reference
  // recomputes the normals of the whole cloud for every descriptor type (features.h:186-188)
  const bool dense = ctx->q_is_surface;
  const Ctx::NrmKey& K = ctx->nrm_key;
  const bool cached = dense && ctx->reuse && K.surf == ctx->surf_version && K.version == ctx->normals_version && K.version != 0 &&
                      K.radius == radius && K.k == k && K.parity == ctx->parity_mode && K.vp[0] == ctx->vp[0] &&
                      K.vp[1] == ctx->vp[1] && K.vp[2] == ctx->vp[2];
  if (cached) {
    ctx->stat_normals_reused++;
    ctx->have_normals = true;
    rows = ctx->normals.as<float4>();  // original order = caller order of dense queries
  } else {
    if (ctx->parity_mode == PFX_PARITY_STRICT) PFX_TRY(strict_normals(ctx, g, radius, k, rows));
    else PFX_TRY(normals_compute(ctx, g, radius, k, rows));
    if (dense) {
      ctx->stat_normals_passes++;
      ctx->nrm_key.surf = ctx->surf_version;
      ctx->nrm_key.version = ctx->normals_version;
      ctx->nrm_key.radius = radius;
      ctx->nrm_key.k = k;
      ctx->nrm_key.parity = ctx->parity_mode;
      for (int a = 0; a < 3; ++a) ctx->nrm_key.vp[a] = ctx->vp[a];
      ctx->nrm_fp.valid = false;
    }
  }
  if (!out || nq == 0) return 0;
  struct RecordFp {  // host delivery of dense normals: remember the buffer, a later setInputNormals of it is a no-op
    Ctx* c; const void* out; size_t n, stride; int curv_off; bool on;
    ~RecordFp() {
      if (!on) return;
      c->nrm_fp = host_fingerprint(out, n, stride, ((size_t)curv_off + 1) * 4);
      c->nrm_fp_surf = c->surf_version;
      c->nrm_fp_version = c->normals_version;
    }
  } record{ctx, out, nq, stride, curv_off, dense && mem == PFX_HOST && ctx->reuse};
  if (mem == PFX_DEVICE) {
    if (stride == 16 && curv_off == 3) return deliver(ctx, out, rows, nq * sizeof(float4), mem);
    PFX_LAUNCH(ctx, normals_out_kernel, div_up((long long)nq, 256), 256, 0, rows, (int)nq, static_cast<unsigned char*>(out), stride, curv_off);
    PFX_CUDA(cudaGetLastError());
    return 0;
  }
  if (stride == 16 && curv_off == 3) return deliver(ctx, out, rows, nq * sizeof(float4), mem);
  PFX_CUDA(ctx->out_stage.ensure(nq * stride));
  PFX_CUDA(cudaMemsetAsync(ctx->out_stage.p, 0, nq * stride, ctx->stream));
  PFX_LAUNCH(ctx, normals_out_kernel, div_up((long long)nq, 256), 256, 0, rows, (int)nq, ctx->out_stage.as<unsigned char>(), stride, curv_off);
  return deliver(ctx, out, ctx->out_stage.p, nq * stride, mem);
}

// ================================================================================== keypoints
extern "C" int pfx_cloud_resolution(pfx_ctx* ctx, double* resolution) {
  PFX_TRY(check_ctx(ctx));
  if (ctx->surf_version == 0) return ctx->fail(PFX_E_PRECOND, "pfx_cloud_resolution: no surface set");
  if (!resolution) return ctx->fail(PFX_E_INVALID, "null output");
  return cloud_resolution(ctx, resolution);
}

static int emit_keypoints(Ctx* ctx, const int* flags_dev, int32_t* kp_idx, size_t cap, size_t* n_kp, int mem,
                          int** idx_dev_out) {
  const int n = (int)ctx->n;
  PFX_CUDA(ctx->tmp3.ensure(std::max<size_t>(n, 1) * sizeof(int)));
  int cnt = 0;
  PFX_TRY(compact_flags(ctx, flags_dev, n, ctx->tmp3.as<int>(), &cnt));
  if (n_kp) *n_kp = (size_t)cnt;
  if (idx_dev_out) *idx_dev_out = ctx->tmp3.as<int>();
  if (kp_idx) {
    if ((size_t)cnt > cap) return ctx->fail(PFX_E_CAPACITY, "keypoint buffer too small");
    PFX_TRY(deliver(ctx, kp_idx, ctx->tmp3.p, (size_t)cnt * sizeof(int), mem));
  }
  return 0;
}

extern "C" int pfx_iss_nms(pfx_ctx* ctx, const double* saliency, double nonmax_radius, int min_neighbors,
                           int32_t* kp_idx, size_t cap, size_t* n_kp, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (ctx->surf_version == 0) return ctx->fail(PFX_E_PRECOND, "pfx_iss_nms: no surface set");
  if (!(nonmax_radius > 0) || !saliency) return ctx->fail(PFX_E_INVALID, "pfx_iss_nms: bad arguments");
  const size_t n = ctx->n;
  if (n_kp) *n_kp = 0;
  if (n == 0) return 0;
  const double* dsal = saliency;
  if (mem == PFX_HOST) {
    PFX_CUDA(ctx->tmp1.ensure(n * sizeof(double)));
    PFX_CUDA(cudaMemcpyAsync(ctx->tmp1.p, saliency, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    dsal = ctx->tmp1.as<double>();
  }
  Grid* g = nullptr;
  PFX_TRY(grid_get(ctx, nonmax_radius, 0, &g));
  PFX_CUDA(ctx->tmp2.ensure(n * sizeof(int)));
  PFX_CUDA(cudaMemsetAsync(ctx->tmp2.p, 0, n * sizeof(int), ctx->stream));
  PFX_TRY(iss_nms(ctx, g, dsal, nonmax_radius, min_neighbors, ctx->tmp2.as<int>()));
  return emit_keypoints(ctx, ctx->tmp2.as<int>(), kp_idx, cap, n_kp, mem, nullptr);
}

extern "C" int pfx_iss(pfx_ctx* ctx, double salient_radius, double nonmax_radius, int min_neighbors, double gamma21,
                       double gamma32, int32_t* kp_idx, size_t cap, size_t* n_kp, double* saliency, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (ctx->surf_version == 0) return ctx->fail(PFX_E_PRECOND, "pfx_iss: no surface set");
  // ISSKeypoint3D::initCompute: salient radius and non-max radius must be strictly positive
  if (!(salient_radius > 0)) return ctx->fail(PFX_E_PRECOND, "pfx_iss: the salient radius must be strictly positive");
  if (!(nonmax_radius > 0)) return ctx->fail(PFX_E_PRECOND, "pfx_iss: the non maxima radius must be strictly positive");
  if (min_neighbors < 0) return ctx->fail(PFX_E_PRECOND, "pfx_iss: negative minimum neighbours");
  const size_t n = ctx->n;
  if (n_kp) *n_kp = 0;
  if (n == 0) return 0;
  PFX_CUDA(ctx->tmp1.ensure(n * sizeof(double)));
  double* dsal = ctx->tmp1.as<double>();
  PFX_CUDA(cudaMemsetAsync(dsal, 0, n * sizeof(double), ctx->stream));
  Grid* gs = nullptr;
  PFX_TRY(grid_get(ctx, salient_radius, 0, &gs));
  PFX_TRY(iss_saliency(ctx, gs, salient_radius, min_neighbors, gamma21, gamma32, dsal));
  Grid* gn = nullptr;
  PFX_TRY(grid_get(ctx, nonmax_radius, 0, &gn));
  PFX_CUDA(ctx->tmp2.ensure(n * sizeof(int)));
  PFX_CUDA(cudaMemsetAsync(ctx->tmp2.p, 0, n * sizeof(int), ctx->stream));
  PFX_TRY(iss_nms(ctx, gn, dsal, nonmax_radius, min_neighbors, ctx->tmp2.as<int>()));
  if (saliency) {
    if (mem == PFX_DEVICE) PFX_TRY(deliver(ctx, saliency, dsal, n * sizeof(double), mem));
    else PFX_CUDA(cudaMemcpyAsync(saliency, dsal, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  return emit_keypoints(ctx, ctx->tmp2.as<int>(), kp_idx, cap, n_kp, mem, nullptr);
}

extern "C" int pfx_harris_nms(pfx_ctx* ctx, const float* response, double radius, float threshold, int32_t* kp_idx,
                              size_t cap, size_t* n_kp, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (ctx->surf_version == 0) return ctx->fail(PFX_E_PRECOND, "pfx_harris_nms: no surface set");
  if (!(radius > 0) || !response) return ctx->fail(PFX_E_INVALID, "pfx_harris_nms: bad arguments");
  const size_t n = ctx->n;
  if (n_kp) *n_kp = 0;
  if (n == 0) return 0;
  const float* dresp = response;
  if (mem == PFX_HOST) {
    PFX_CUDA(ctx->tmp1.ensure(n * sizeof(float)));
    PFX_CUDA(cudaMemcpyAsync(ctx->tmp1.p, response, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    dresp = ctx->tmp1.as<float>();
  }
  Grid* g = nullptr;
  PFX_TRY(grid_get(ctx, radius, 0, &g));
  PFX_CUDA(ctx->tmp2.ensure(n * sizeof(int)));
  PFX_CUDA(cudaMemsetAsync(ctx->tmp2.p, 0, n * sizeof(int), ctx->stream));
  PFX_TRY(harris_nms(ctx, g, dresp, radius, threshold, ctx->tmp2.as<int>()));
  return emit_keypoints(ctx, ctx->tmp2.as<int>(), kp_idx, cap, n_kp, mem, nullptr);
}

static int harris3d_impl(pfx_ctx* ctx, double radius, float threshold, int nonmax, int refine, float snap_max_d2,
                         float* response, int32_t* kp_idx, float* kp_xyz, int32_t* snapped_idx, size_t cap,
                         size_t* n_kp, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (ctx->surf_version == 0) return ctx->fail(PFX_E_PRECOND, "pfx_harris3d: no surface set");
  if (!(radius > 0)) return ctx->fail(PFX_E_PRECOND, "pfx_harris3d: radius must be > 0");
  const size_t n = ctx->n;
  if (n_kp) *n_kp = 0;
  if (n == 0) return 0;
  Grid* g = nullptr;
  PFX_TRY(grid_get(ctx, radius, 0, &g));
  // HarrisKeypoint3D::initCompute: no normals given -> NormalEstimation on the surface at the same radius
  if (!ctx->have_normals) {
    bool saved = ctx->q_is_surface;
    ctx->q_is_surface = true;
    int rc = ctx->parity_mode == PFX_PARITY_STRICT ? strict_normals(ctx, g, radius, 0, nullptr)
                                                   : normals_compute(ctx, g, radius, 0, nullptr);
    ctx->q_is_surface = saved;
    if (rc) return rc;
  }
  PFX_CUDA(ctx->tmp1.ensure(n * sizeof(float)));
  float* dresp = ctx->tmp1.as<float>();
  if (ctx->parity_mode == PFX_PARITY_STRICT) PFX_TRY(harris_response_strict(ctx, g, radius, dresp));
  else PFX_TRY(harris_response(ctx, g, radius, dresp));
  if (response) {
    if (mem == PFX_DEVICE) PFX_TRY(deliver(ctx, response, dresp, n * sizeof(float), mem));
    else PFX_CUDA(cudaMemcpyAsync(response, dresp, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (!nonmax) {  // PCL: output = the response cloud itself (every point)
    if (n_kp) *n_kp = n;
    PFX_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
  }
  PFX_CUDA(ctx->tmp2.ensure(n * sizeof(int)));
  PFX_CUDA(cudaMemsetAsync(ctx->tmp2.p, 0, n * sizeof(int), ctx->stream));
  PFX_TRY(harris_nms(ctx, g, dresp, radius, threshold, ctx->tmp2.as<int>()));
  int* didx = nullptr;
  size_t cnt = 0;
  PFX_TRY(emit_keypoints(ctx, ctx->tmp2.as<int>(), kp_idx, cap, &cnt, mem, &didx));
  if (n_kp) *n_kp = cnt;
  if (cnt == 0 || (!kp_xyz && !snapped_idx)) return 0;
  if (cnt > cap) return ctx->fail(PFX_E_CAPACITY, "keypoint buffer too small");
  // corner positions (refined) and the snap back onto the cloud
  PFX_CUDA(ctx->stage2.ensure(cnt * 3 * sizeof(float) + cnt * sizeof(int)));
  float* dxyz = ctx->stage2.as<float>();
  int* dsnap = reinterpret_cast<int*>(dxyz + cnt * 3);
  PFX_LAUNCH(ctx, gather_xyz_kernel, div_up((long long)cnt, 256), 256, 0, ctx->surf.as<float4>(), didx, (int)cnt, dxyz);
  if (refine) {
    if (ctx->parity_mode == PFX_PARITY_STRICT) PFX_TRY(harris_refine_strict(ctx, g, radius, dxyz, (int)cnt));
    else PFX_TRY(harris_refine(ctx, g, radius, dxyz, (int)cnt));
  }
  if (kp_xyz) {
    if (mem == PFX_DEVICE) PFX_TRY(deliver(ctx, kp_xyz, dxyz, cnt * 3 * sizeof(float), mem));
    else PFX_CUDA(cudaMemcpyAsync(kp_xyz, dxyz, cnt * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (snapped_idx) {
    PFX_TRY(snap_to_cloud(ctx, dxyz, (int)cnt, snap_max_d2, dsnap));
    PFX_TRY(deliver(ctx, snapped_idx, dsnap, cnt * sizeof(int), mem));
  }
  PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

static int harris6d_impl(pfx_ctx* ctx, double radius, float threshold, int nonmax, int refine, float snap_max_d2,
                         float* response, int32_t* kp_idx, float* kp_xyz, int32_t* snapped_idx, size_t cap,
                         size_t* n_kp, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (ctx->surf_version == 0) return ctx->fail(PFX_E_PRECOND, "pfx_harris6d: no surface set");
  if (!(radius > 0)) return ctx->fail(PFX_E_PRECOND, "pfx_harris6d: radius must be > 0");
  const size_t n = ctx->n;
  if (n_kp) *n_kp = 0;
  if (n == 0) return 0;
  Grid* g = nullptr;
  PFX_TRY(grid_get(ctx, radius, 0, &g));
  // HarrisKeypoint3D::initCompute: no normals given -> NormalEstimation on the surface at the same radius
  if (!ctx->have_normals) {
    bool saved = ctx->q_is_surface;
    ctx->q_is_surface = true;
    int rc = strict_normals(ctx, g, radius, 0, nullptr);
    ctx->q_is_surface = saved;
    if (rc) return rc;
  }
  PFX_CUDA(ctx->tmp1.ensure(n * sizeof(float)));
  float* dresp = ctx->tmp1.as<float>();
  PFX_TRY(harris6d_response(ctx, g, radius, dresp, nullptr));
  if (response) {
    if (mem == PFX_DEVICE) PFX_TRY(deliver(ctx, response, dresp, n * sizeof(float), mem));
    else PFX_CUDA(cudaMemcpyAsync(response, dresp, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (!nonmax) {  // PCL: output = the response cloud itself (every point)
    if (n_kp) *n_kp = n;
    PFX_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
  }
  PFX_CUDA(ctx->tmp2.ensure(n * sizeof(int)));
  PFX_CUDA(cudaMemsetAsync(ctx->tmp2.p, 0, n * sizeof(int), ctx->stream));
  PFX_TRY(harris_nms(ctx, g, dresp, radius, threshold, ctx->tmp2.as<int>()));
  int* didx = nullptr;
  size_t cnt = 0;
  PFX_TRY(emit_keypoints(ctx, ctx->tmp2.as<int>(), kp_idx, cap, &cnt, mem, &didx));
  if (n_kp) *n_kp = cnt;
  if (cnt == 0 || (!kp_xyz && !snapped_idx)) return 0;
  if (cnt > cap) return ctx->fail(PFX_E_CAPACITY, "keypoint buffer too small");
  // corner positions (refined) and the snap back onto the cloud
  PFX_CUDA(ctx->stage2.ensure(cnt * 3 * sizeof(float) + cnt * sizeof(int)));
  float* dxyz = ctx->stage2.as<float>();
  int* dsnap = reinterpret_cast<int*>(dxyz + cnt * 3);
  PFX_LAUNCH(ctx, gather_xyz_kernel, div_up((long long)cnt, 256), 256, 0, ctx->surf.as<float4>(), didx, (int)cnt, dxyz);
  if (refine) {
    PFX_TRY(harris_refine_strict(ctx, g, radius, dxyz, (int)cnt));
  }
  if (kp_xyz) {
    if (mem == PFX_DEVICE) PFX_TRY(deliver(ctx, kp_xyz, dxyz, cnt * 3 * sizeof(float), mem));
    else PFX_CUDA(cudaMemcpyAsync(kp_xyz, dxyz, cnt * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (snapped_idx) {
    PFX_TRY(snap_to_cloud(ctx, dxyz, (int)cnt, snap_max_d2, dsnap));
    PFX_TRY(deliver(ctx, snapped_idx, dsnap, cnt * sizeof(int), mem));
  }
  PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int pfx_harris3d(pfx_ctx* ctx, double radius, float threshold, int nonmax, int refine, float snap_max_d2,
                            float* response, int32_t* kp_idx, float* kp_xyz, int32_t* snapped_idx, size_t cap,
                            size_t* n_kp, int mem) {
  if (!ctx) return PFX_E_INVALID;
  // normals estimated internally (HarrisKeypoint3D::initCompute) stay private to this call
  const bool internal_normals = !ctx->have_normals;
  int rc = harris3d_impl(ctx, radius, threshold, nonmax, refine, snap_max_d2, response, kp_idx, kp_xyz, snapped_idx,
                         cap, n_kp, mem);
  if (internal_normals) {
    ctx->have_normals = false;
    ctx->normals_sorted_for = nullptr;
  }
  return rc;
}

// HarrisKeypoint6D (keypoints.h:166-179): the surface's colours must have been set (pfx_set_surface_colors)
extern "C" int pfx_harris6d(pfx_ctx* ctx, double radius, float threshold, int nonmax, int refine, float snap_max_d2,
                            float* response, int32_t* kp_idx, float* kp_xyz, int32_t* snapped_idx, size_t cap,
                            size_t* n_kp, int mem) {
  if (!ctx) return PFX_E_INVALID;
  if (ctx->surf_version != 0 && ctx->surf_rgb_version != ctx->surf_version)
    return ctx->fail(PFX_E_STATE, "pfx_harris6d: the surface has no colours (pfx_set_surface_colors)");
  // the detector always estimates its own normals at its radius and keeps them private to the call; they are computed
  // in reference order (strict.cu) whatever the parity mode: the response is an eigenvalue of their covariance
  const bool had = ctx->have_normals;
  DevBuf keep;
  if (had) {  // park the caller's normals
    keep = ctx->normals;
    ctx->normals = DevBuf();
  }
  ctx->have_normals = false;
  int rc = harris6d_impl(ctx, radius, threshold, nonmax, refine, snap_max_d2, response, kp_idx, kp_xyz, snapped_idx,
                         cap, n_kp, mem);
  if (had) {
    ctx->normals.release();
    ctx->normals = keep;
    ctx->have_normals = true;
    ctx->normals_version++;
  } else {
    ctx->have_normals = false;
  }
  ctx->normals_sorted_for = nullptr;
  return rc;
}

// ================================================================================== descriptors
extern "C" int pfx_fpfh(pfx_ctx* ctx, double radius, int k, float* out, size_t stride, int mem) {
  PFX_TRY(check_ctx(ctx));
  PFX_TRY(check_search_params(ctx, radius, k, "pfx_fpfh"));
  if (k > 32) return ctx->fail(PFX_E_INVALID, "pfx_fpfh: k must be <= 32");
  if (!ctx->have_normals) return ctx->fail(PFX_E_STATE, "pfx_fpfh: no input normals (setInputNormals)");
  if (!out || stride < 132 || (stride & 3)) return ctx->fail(PFX_E_INVALID, "pfx_fpfh: bad output / stride");
  const size_t nq = ctx->num_queries();
  Grid* g = nullptr;
  PFX_TRY(grid_get(ctx, radius > 0 ? radius : 0.0, k, &g));
  float* dout = out;
  if (mem == PFX_HOST_ASYNC) {
    void* st = nullptr;
    PFX_TRY(async_stage_acquire(ctx, 0, nq * stride, &st));
    dout = static_cast<float*>(st);
    if (stride != 132) PFX_CUDA(cudaMemsetAsync(dout, 0, nq * stride, ctx->stream));
  } else if (mem == PFX_HOST) {
    PFX_CUDA(ctx->out_stage.ensure(std::max<size_t>(nq * stride, 16)));
    dout = ctx->out_stage.as<float>();
    if (stride != 132) PFX_CUDA(cudaMemsetAsync(dout, 0, nq * stride, ctx->stream));
  }
  PFX_TRY(fpfh_compute(ctx, g, radius, k, dout, stride / 4, nullptr));
  if (mem == PFX_HOST_ASYNC) return async_deliver(ctx, 0, out, nq * stride);
  if (mem == PFX_HOST) return deliver(ctx, out, dout, nq * stride, mem);
  return 0;
}

// shared body of the per-query descriptor calls that write fixed-size float rows
template <typename F>
static int rows_call(pfx_ctx* ctx, double radius, int k, float* out, size_t stride, int mem, size_t row_bytes,
                     const char* what, bool needs_normals, F&& compute) {
  PFX_TRY(check_ctx(ctx));
  PFX_TRY(check_search_params(ctx, radius, k, what));
  if (k > 32) return ctx->fail(PFX_E_INVALID, std::string(what) + ": k must be <= 32");
  if (needs_normals && !ctx->have_normals) return ctx->fail(PFX_E_STATE, std::string(what) + ": no input normals (setInputNormals)");
  if (!out || stride < row_bytes || (stride & 3) || (mem != PFX_HOST && mem != PFX_DEVICE))
    return ctx->fail(PFX_E_INVALID, std::string(what) + ": bad output / stride / mem");
  const size_t nq = ctx->num_queries();
  Grid* g = nullptr;
  PFX_TRY(grid_get(ctx, radius > 0 ? radius : 0.0, k, &g));
  float* dout = out;
  if (mem == PFX_HOST) {
    PFX_CUDA(ctx->out_stage.ensure(std::max<size_t>(nq * stride, 16)));
    dout = ctx->out_stage.as<float>();
    if (stride != row_bytes) PFX_CUDA(cudaMemsetAsync(dout, 0, nq * stride, ctx->stream));
  }
  PFX_TRY(compute(g, dout));
  if (mem == PFX_HOST) return deliver(ctx, out, dout, nq * stride, mem);
  return 0;
}

extern "C" int pfx_pfh125(pfx_ctx* ctx, double radius, int k, float* out, size_t stride, int mem) {
  return rows_call(ctx, radius, k, out, stride, mem, 500, "pfx_pfh125", true,
                   [&](Grid* g, float* dout) { return pfh_compute(ctx, g, radius, k, dout, stride / 4); });
}

extern "C" int pfx_principal_curvatures(pfx_ctx* ctx, double radius, int k, float* out, size_t stride, int mem) {
  return rows_call(ctx, radius, k, out, stride, mem, 20, "pfx_principal_curvatures", true,
                   [&](Grid* g, float* dout) { return curvature_compute(ctx, g, radius, k, dout, stride / 4); });
}

extern "C" int pfx_moment_invariants(pfx_ctx* ctx, double radius, int k, float* out, size_t stride, int mem) {
  return rows_call(ctx, radius, k, out, stride, mem, 12, "pfx_moment_invariants", false,
                   [&](Grid* g, float* dout) { return moments_compute(ctx, g, radius, k, dout, stride / 4); });
}

extern "C" float pfx_seq_float_sum(float incr, long long count) { return seq_float_sum(incr, count < 0 ? 0 : count); }

extern "C" int pfx_spfh(pfx_ctx* ctx, double radius, int k, float* out, int mem) {
  PFX_TRY(check_ctx(ctx));
  PFX_TRY(check_search_params(ctx, radius, k, "pfx_spfh"));
  if (k > 32) return ctx->fail(PFX_E_INVALID, "pfx_spfh: k must be <= 32");
  if (!ctx->have_normals) return ctx->fail(PFX_E_STATE, "pfx_spfh: no input normals");
  if (!out) return ctx->fail(PFX_E_INVALID, "pfx_spfh: null output");
  const size_t n = ctx->n;
  if (n == 0) return 0;
  Grid* g = nullptr;
  PFX_TRY(grid_get(ctx, radius > 0 ? radius : 0.0, k, &g));
  float* dout = out;
  if (mem == PFX_HOST) {
    PFX_CUDA(ctx->out_stage.ensure(n * 33 * sizeof(float)));
    dout = ctx->out_stage.as<float>();
  }
  bool saved = ctx->q_is_surface;
  ctx->q_is_surface = true;
  int rc = fpfh_compute(ctx, g, radius, k, nullptr, 33, dout);
  ctx->q_is_surface = saved;
  if (rc) return rc;
  if (mem == PFX_HOST) return deliver(ctx, out, dout, n * 33 * sizeof(float), mem);
  return 0;
}

extern "C" int pfx_shot_lrf(pfx_ctx* ctx, double radius, float* rf9, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (ctx->surf_version == 0) return ctx->fail(PFX_E_PRECOND, "pfx_shot_lrf: no surface set");
  if (!(radius > 0)) return ctx->fail(PFX_E_PRECOND, "pfx_shot_lrf: radius must be > 0");
  if (!rf9) return ctx->fail(PFX_E_INVALID, "pfx_shot_lrf: null output");
  const size_t nq = ctx->num_queries();
  if (nq == 0) return 0;
  Grid* g = nullptr;
  PFX_TRY(grid_get(ctx, radius, 0, &g));
  float* d = rf9;
  if (mem == PFX_HOST) {
    PFX_CUDA(ctx->out_stage.ensure(nq * 9 * sizeof(float)));
    d = ctx->out_stage.as<float>();
  }
  PFX_TRY(shot_lrf_compute(ctx, g, radius, d, nullptr));
  if (mem == PFX_HOST) return deliver(ctx, rf9, d, nq * 9 * sizeof(float), mem);
  return 0;
}

extern "C" int pfx_shot352(pfx_ctx* ctx, double radius, const float* lrf_in, float* out, size_t stride, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (ctx->surf_version == 0) return ctx->fail(PFX_E_PRECOND, "pfx_shot352: no surface set");
  // SHOTEstimation::initCompute rejects k-search; the radius must be set
  if (!(radius > 0)) return ctx->fail(PFX_E_PRECOND, "pfx_shot352: SHOT needs a radius search (setRadiusSearch)");
  if (!ctx->have_normals) return ctx->fail(PFX_E_STATE, "pfx_shot352: no input normals (setInputNormals)");
  if (!out || stride < 1444 || (stride & 3)) return ctx->fail(PFX_E_INVALID, "pfx_shot352: bad output / stride");
  const size_t nq = ctx->num_queries();
  if (nq == 0) return 0;
  Grid* g = nullptr;
  // frames estimated here + the dense k-search rows of this surface resident: no radius index at all (shot_fused.cu)
  if (lrf_in || !shot_rows_available(ctx, radius)) PFX_TRY(grid_for_radius(ctx, radius, &g));
  float* dout = out;
  if (mem == PFX_HOST_ASYNC) {
    if (lrf_in) return ctx->fail(PFX_E_INVALID, "pfx_shot352: PFX_HOST_ASYNC output with caller-supplied frames is not supported");
    void* st = nullptr;
    PFX_TRY(async_stage_acquire(ctx, 1, nq * stride, &st));
    dout = static_cast<float*>(st);
    if (stride != 1444) PFX_CUDA(cudaMemsetAsync(dout, 0, nq * stride, ctx->stream));
  } else if (mem == PFX_HOST) {
    PFX_CUDA(ctx->out_stage.ensure(nq * stride));
    dout = ctx->out_stage.as<float>();
    if (stride != 1444) PFX_CUDA(cudaMemsetAsync(dout, 0, nq * stride, ctx->stream));
  }
  if (lrf_in) {  // caller-supplied frames (setInputReferenceFrames)
    PFX_CUDA(ctx->tmp2.ensure(nq * 9 * sizeof(float)));
    float* drf = ctx->tmp2.as<float>();
    PFX_CUDA(cudaMemcpyAsync(drf, lrf_in, nq * 9 * sizeof(float),
                             mem == PFX_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, ctx->stream));
    PFX_TRY(shot_compute(ctx, g, radius, drf, dout, stride / 4));
  } else {  // frames estimated at the same radius: one fused kernel
    PFX_TRY(shot_fused_compute(ctx, g, radius, dout, stride / 4));
  }
  if (mem == PFX_HOST_ASYNC) return async_deliver(ctx, 1, out, nq * stride);
  if (mem == PFX_HOST) return deliver(ctx, out, dout, nq * stride, mem);
  return 0;
}

namespace pfx {
__global__ void rgb_words_kernel(const unsigned char* __restrict__ src, size_t stride, int n, unsigned* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = *reinterpret_cast<const unsigned*>(src + (size_t)i * stride) & 0x00ffffffu;
}
}  // namespace pfx

// SHOT1344: colours of the surface / the queries, then the descriptor
extern "C" int pfx_set_surface_colors(pfx_ctx* ctx, const void* rgb, size_t n, size_t stride, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (ctx->surf_version == 0) return ctx->fail(PFX_E_PRECOND, "pfx_set_surface_colors: no surface set");
  if (n != ctx->n) return ctx->fail(PFX_E_PRECOND, "pfx_set_surface_colors: the number of colours differs from the surface size");
  if ((n && !rgb) || stride < 4 || (stride & 3) || (mem != PFX_HOST && mem != PFX_DEVICE))
    return ctx->fail(PFX_E_INVALID, "pfx_set_surface_colors: bad pointer / stride / mem");
  const unsigned char* src = nullptr;
  PFX_TRY(upload_records(ctx, rgb, n, stride, mem, ctx->stage, &src));
  PFX_TRY(colors_to_lab(ctx, src, stride, (int)n, ctx->surf_lab));
  ctx->surf_lab_version = ctx->surf_version;
  // the packed words themselves (Harris 6D derives its intensity from them)
  PFX_CUDA(ctx->surf_rgb.ensure(std::max<size_t>(n, 1) * sizeof(unsigned)));
  if (n) PFX_LAUNCH(ctx, rgb_words_kernel, div_up((long long)n, 256), 256, 0, src, stride, (int)n, ctx->surf_rgb.as<unsigned>());
  PFX_CUDA(cudaGetLastError());
  ctx->surf_rgb_version = ctx->surf_version;
  return 0;
}

extern "C" int pfx_set_query_colors(pfx_ctx* ctx, const void* rgb, size_t n, size_t stride, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (ctx->q_is_surface) return ctx->fail(PFX_E_PRECOND, "pfx_set_query_colors: the queries are the surface (pfx_set_surface_colors)");
  if (n != ctx->nq) return ctx->fail(PFX_E_PRECOND, "pfx_set_query_colors: the number of colours differs from the number of queries");
  if ((n && !rgb) || stride < 4 || (stride & 3) || (mem != PFX_HOST && mem != PFX_DEVICE))
    return ctx->fail(PFX_E_INVALID, "pfx_set_query_colors: bad pointer / stride / mem");
  const unsigned char* src = nullptr;
  PFX_TRY(upload_records(ctx, rgb, n, stride, mem, ctx->stage, &src));
  PFX_TRY(colors_to_lab(ctx, src, stride, (int)n, ctx->qry_lab));
  ctx->qry_lab_version = ctx->qry_version;
  return 0;
}

extern "C" int pfx_shot1344(pfx_ctx* ctx, double radius, const float* lrf_in, float* out, size_t stride, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (ctx->surf_version == 0) return ctx->fail(PFX_E_PRECOND, "pfx_shot1344: no surface set");
  if (!(radius > 0)) return ctx->fail(PFX_E_PRECOND, "pfx_shot1344: SHOT needs a radius search (setRadiusSearch)");
  if (!ctx->have_normals) return ctx->fail(PFX_E_STATE, "pfx_shot1344: no input normals (setInputNormals)");
  if (ctx->surf_lab_version != ctx->surf_version) return ctx->fail(PFX_E_STATE, "pfx_shot1344: no surface colours (pfx_set_surface_colors)");
  if (!ctx->q_is_surface && ctx->qry_lab_version != ctx->qry_version)
    return ctx->fail(PFX_E_STATE, "pfx_shot1344: no query colours (pfx_set_query_colors)");
  if (!out || stride < 5412 || (stride & 3) || (mem != PFX_HOST && mem != PFX_DEVICE))
    return ctx->fail(PFX_E_INVALID, "pfx_shot1344: bad output / stride / mem");
  const size_t nq = ctx->num_queries();
  if (nq == 0) return 0;
  Grid* g = nullptr;
  PFX_TRY(grid_get(ctx, radius, 0, &g));
  float* dout = out;
  if (mem == PFX_HOST) {
    PFX_CUDA(ctx->out_stage.ensure(nq * stride));
    dout = ctx->out_stage.as<float>();
    if (stride != 5412) PFX_CUDA(cudaMemsetAsync(dout, 0, nq * stride, ctx->stream));
  }
  PFX_CUDA(ctx->tmp2.ensure(nq * 9 * sizeof(float)));
  float* drf = ctx->tmp2.as<float>();
  if (lrf_in)
    PFX_CUDA(cudaMemcpyAsync(drf, lrf_in, nq * 9 * sizeof(float),
                             mem == PFX_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, ctx->stream));
  else
    PFX_TRY(shot_lrf_compute(ctx, g, radius, drf, nullptr));
  PFX_TRY(shot_color_compute(ctx, g, radius, drf, dout, stride / 4));
  if (mem == PFX_HOST) return deliver(ctx, out, dout, nq * stride, mem);
  return 0;
}

extern "C" int pfx_spin_image153(pfx_ctx* ctx, double radius, const void* query_normals, size_t n_normals,
                                 size_t stride_normals, float* out, size_t stride, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (ctx->surf_version == 0) return ctx->fail(PFX_E_PRECOND, "pfx_spin_image153: no surface set");
  // SpinImageEstimation::initCompute: radius search only, normals of the input cloud required
  if (!(radius > 0)) return ctx->fail(PFX_E_PRECOND, "pfx_spin_image153: needs a radius search (setRadiusSearch)");
  const size_t nq = ctx->num_queries();
  if (!query_normals) return ctx->fail(PFX_E_PRECOND, "pfx_spin_image153: no input normals (setInputNormals)");
  if (n_normals != nq)
    return ctx->fail(PFX_E_PRECOND, "pfx_spin_image153: the number of normals differs from the number of input points");
  if (!out || stride < 612 || (stride & 3) || stride_normals < 12 || (stride_normals & 3) || (mem != PFX_HOST && mem != PFX_DEVICE))
    return ctx->fail(PFX_E_INVALID, "pfx_spin_image153: bad output / stride / mem");
  if (nq == 0) return 0;
  Grid* g = nullptr;
  PFX_TRY(grid_get(ctx, radius, 0, &g));
  const unsigned char* dn = nullptr;
  PFX_TRY(upload_records(ctx, query_normals, nq, stride_normals, mem, ctx->stage2, &dn));
  float* dout = out;
  if (mem == PFX_HOST) {
    PFX_CUDA(ctx->out_stage.ensure(nq * stride));
    dout = ctx->out_stage.as<float>();
    if (stride != 612) PFX_CUDA(cudaMemsetAsync(dout, 0, nq * stride, ctx->stream));
  }
  PFX_TRY(spin_compute(ctx, g, radius, reinterpret_cast<const float*>(dn), stride_normals / 4, dout, stride / 4));
  if (mem == PFX_HOST) return deliver(ctx, out, dout, nq * stride, mem);
  return 0;
}

extern "C" int pfx_usc1980(pfx_ctx* ctx, double search_radius, double min_radius, double density_radius, double local_radius,
                           const float* lrf_in, float* out, size_t stride, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (ctx->surf_version == 0) return ctx->fail(PFX_E_PRECOND, "pfx_usc1980: no surface set");
  if (!(search_radius > 0)) return ctx->fail(PFX_E_PRECOND, "pfx_usc1980: needs a radius search (setRadiusSearch)");
  // UniqueShapeContext::initCompute: "search_radius_ must be GREATER than min_radius_"
  if (!(min_radius > 0) || !(density_radius > 0) || !(local_radius > 0) || search_radius < min_radius)
    return ctx->fail(PFX_E_PRECOND, "pfx_usc1980: radii must be positive and search_radius >= min_radius");
  if (!out || stride < 7956 || (stride & 3) || (mem != PFX_HOST && mem != PFX_DEVICE))
    return ctx->fail(PFX_E_INVALID, "pfx_usc1980: bad output / stride / mem");
  const size_t nq = ctx->num_queries();
  if (nq == 0) return 0;
  float* dout = out;
  const float* dlrf = lrf_in;
  if (mem == PFX_HOST) {
    PFX_CUDA(ctx->out_stage.ensure(nq * stride));
    dout = ctx->out_stage.as<float>();
    if (stride != 7956) PFX_CUDA(cudaMemsetAsync(dout, 0, nq * stride, ctx->stream));
    if (lrf_in) {
      PFX_CUDA(ctx->tmp4.ensure(nq * 9 * sizeof(float)));
      PFX_CUDA(cudaMemcpyAsync(ctx->tmp4.p, lrf_in, nq * 9 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
      dlrf = ctx->tmp4.as<float>();
    }
  }
  PFX_TRY(usc_compute(ctx, search_radius, min_radius, density_radius, local_radius, dlrf, dout, stride / 4));
  if (mem == PFX_HOST) return deliver(ctx, out, dout, nq * stride, mem);
  return 0;
}

// 3DSC: the reference's first active descriptor (evaluation.cpp:67, :319-345)
extern "C" int pfx_sc3d1980(pfx_ctx* ctx, double search_radius, double min_radius, double density_radius, uint64_t seed,
                            float* out, size_t stride, float* frames_out, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (ctx->surf_version == 0) return ctx->fail(PFX_E_PRECOND, "pfx_sc3d1980: no surface set");
  if (!(search_radius > 0)) return ctx->fail(PFX_E_PRECOND, "pfx_sc3d1980: needs a radius search (setRadiusSearch)");
  // ShapeContext3DEstimation::initCompute: "search_radius_ must be GREATER than min_radius_"
  if (!(min_radius > 0) || !(density_radius > 0) || search_radius < min_radius)
    return ctx->fail(PFX_E_PRECOND, "pfx_sc3d1980: radii must be positive and search_radius >= min_radius");
  if (!ctx->have_normals) return ctx->fail(PFX_E_STATE, "pfx_sc3d1980: no input normals (setInputNormals)");
  if (!out || stride < 7956 || (stride & 3) || (mem != PFX_HOST && mem != PFX_DEVICE))
    return ctx->fail(PFX_E_INVALID, "pfx_sc3d1980: bad output / stride / mem");
  const size_t nq = ctx->num_queries();
  if (nq == 0) return 0;
  float* dout = out;
  float* dfr = frames_out;
  if (mem == PFX_HOST) {
    PFX_CUDA(ctx->out_stage.ensure(nq * stride));
    dout = ctx->out_stage.as<float>();
    if (stride != 7956) PFX_CUDA(cudaMemsetAsync(dout, 0, nq * stride, ctx->stream));
    if (frames_out) {
      PFX_CUDA(ctx->stage2.ensure(nq * 9 * sizeof(float)));
      dfr = ctx->stage2.as<float>();
    }
  }
  PFX_TRY(sc3d_compute(ctx, search_radius, min_radius, density_radius, seed, dout, stride / 4, dfr));
  if (mem == PFX_HOST) {
    if (frames_out) PFX_CUDA(cudaMemcpyAsync(frames_out, dfr, nq * 9 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    return deliver(ctx, out, dout, nq * stride, mem);
  }
  return 0;
}

// ================================================================================== matching
namespace pfx {
__global__ void reciprocal_kernel(const int* __restrict__ s2t, const float* __restrict__ sd2,
                                  const int* __restrict__ t2s, int na, int reciprocal, float max_d2,
                                  int* __restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= na) return;
  int j = s2t[i];
  bool keep = j >= 0;
  if (keep && reciprocal) keep = (t2s[j] == i);
  if (keep && max_d2 >= 0.f) keep = sd2[i] <= max_d2;
  flags[i] = keep ? 1 : 0;
}
__global__ void corr_emit_kernel(const int* __restrict__ flags, const int* __restrict__ pos,
                                 const int* __restrict__ s2t, const float* __restrict__ sd2, int na,
                                 pfx_correspondence* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= na || !flags[i]) return;
  pfx_correspondence c;
  c.index_query = i;
  c.index_match = s2t[i];
  c.distance = sd2[i];
  out[pos[i]] = c;
}
}  // namespace pfx

static int match_upload(Ctx* ctx, const float* m, size_t rows, size_t stride, int dim, int mem, DevBuf& buf,
                        const float** dev, int* ld) {
  if (stride < (size_t)dim * 4 || (stride & 3)) return ctx->fail(PFX_E_INVALID, "pfx_match: bad row stride");
  *ld = (int)(stride / 4);
  if (mem == PFX_DEVICE) {
    *dev = m;
    return 0;
  }
  PFX_CUDA(buf.ensure(std::max<size_t>(rows * stride, 16)));
  if (rows) PFX_CUDA(cudaMemcpyAsync(buf.p, m, rows * stride, cudaMemcpyHostToDevice, ctx->stream));
  *dev = buf.as<float>();
  return 0;
}

// engine: 1 = tcgen05 candidate GEMM + exact fp32 rescore (match_tc.cu), 0 = exact fp32 all-pairs scan.
// Both return bit-identical results; auto picks the tensor cores once the all-pairs work is large.
static bool match_use_tc(const Ctx* ctx, size_t na, size_t nb, int dim) {
  // a descriptor whose A tile does not fit in shared memory (SHOT1344, USC1980) goes to the exact scan whatever
  // engine was asked for: both engines return the same bits
  if (nb == 0 || !match_tc_fits(dim)) return false;
  if (ctx->match_engine >= 0) return ctx->match_engine == 1;
  return (double)na * (double)nb * (double)dim >= 1.5e9;
}

namespace pfx {
// device rows in, device results out: the engine choice of pfx_match_nn for the ring matcher (group.cu)
int match_dispatch_dev(Ctx* ctx, const float* a, int na, int lda, const float* b, int nb, int ldb, int dim, int* idx,
                       float* d2) {
  if (match_use_tc(ctx, na, nb, dim)) return match_nn_tc(ctx, a, na, lda, b, nb, ldb, dim, idx, d2);
  return match_nn_exact(ctx, a, na, lda, b, nb, ldb, dim, idx, d2);
}
}  // namespace pfx

static int match_dispatch(Ctx* ctx, const float* a, int na, int lda, const float* b, int nb, int ldb, int dim,
                          int* idx, float* d2) {
  if (match_use_tc(ctx, na, nb, dim)) return match_nn_tc(ctx, a, na, lda, b, nb, ldb, dim, idx, d2);
  return match_nn_exact(ctx, a, na, lda, b, nb, ldb, dim, idx, d2);
}

// statistics of the tensor-core matcher since the context was created: out[0] 1-NN passes run on the
// tensor cores, out[1] query rows they processed, out[2] rows whose certificate failed (redone exactly)
extern "C" int pfx_match_info(pfx_ctx* ctx, double* out4) {
  if (!ctx || !out4) return PFX_E_INVALID;
  out4[0] = (double)ctx->match_tc_calls;
  out4[1] = (double)ctx->match_rows;
  out4[2] = (double)ctx->match_redo;
  out4[3] = 0;
  return 0;
}

extern "C" int pfx_set_reuse(pfx_ctx* ctx, int enable) {
  if (!ctx) return PFX_E_INVALID;
  ctx->reuse = enable != 0;
  ctx->surf_fp.valid = false;
  ctx->nrm_fp.valid = false;
  return 0;
}

// out6: surface uploads, surface announcements answered from the resident copy, dense normals passes, pfx_normals
// calls answered from the resident normals, normals uploads, normals uploads skipped
extern "C" int pfx_reuse_info(const pfx_ctx* ctx, uint64_t* out6) {
  if (!ctx || !out6) return PFX_E_INVALID;
  out6[0] = ctx->stat_surface_uploads;
  out6[1] = ctx->stat_surface_reused;
  out6[2] = ctx->stat_normals_passes;
  out6[3] = ctx->stat_normals_reused;
  out6[4] = ctx->stat_normals_uploads;
  out6[5] = ctx->stat_normals_upload_skipped;
  return 0;
}

extern "C" int pfx_set_parity_mode(pfx_ctx* ctx, int mode) {
  if (!ctx || (mode != PFX_PARITY_FAST && mode != PFX_PARITY_STRICT)) return PFX_E_INVALID;
  if (ctx->parity_mode != mode) {  // normals of the other mode are not reused
    ctx->parity_mode = mode;
  }
  return 0;
}

extern "C" int pfx_set_match_engine(pfx_ctx* ctx, int engine) {
  if (!ctx || engine < -1 || engine > 1) return PFX_E_INVALID;
  ctx->match_engine = engine;
  return 0;
}

extern "C" int pfx_match_nn(pfx_ctx* ctx, const float* a, size_t na, size_t stride_a, const float* b, size_t nb,
                            size_t stride_b, int dim, int32_t* nn_idx, float* nn_d2, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (dim <= 0 || !nn_idx || (na && !a) || (nb && !b)) return ctx->fail(PFX_E_INVALID, "pfx_match_nn: bad arguments");
  if (na == 0) return 0;
  const float *da, *db;
  int lda, ldb;
  PFX_TRY(match_upload(ctx, a, na, stride_a, dim, mem, ctx->stage, &da, &lda));
  PFX_TRY(match_upload(ctx, b, nb, stride_b, dim, mem, ctx->stage2, &db, &ldb));
  int* didx = nn_idx;
  float* dd2 = nn_d2;
  if (mem == PFX_HOST) {
    PFX_CUDA(ctx->out_stage.ensure(na * (sizeof(int) + sizeof(float))));
    didx = ctx->out_stage.as<int>();
    dd2 = reinterpret_cast<float*>(didx + na);
  } else if (!dd2) {
    PFX_CUDA(ctx->out_stage.ensure(na * sizeof(float)));
    dd2 = ctx->out_stage.as<float>();
  }
  PFX_TRY(match_dispatch(ctx, da, (int)na, lda, db, (int)nb, ldb, dim, didx, dd2));
  if (mem == PFX_HOST) {
    if (nn_d2) PFX_CUDA(cudaMemcpyAsync(nn_d2, dd2, na * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    return deliver(ctx, nn_idx, didx, na * sizeof(int), mem);
  }
  return 0;
}

extern "C" int pfx_match(pfx_ctx* ctx, const float* a, size_t na, size_t stride_a, const float* b, size_t nb,
                         size_t stride_b, int dim, int reciprocal, float max_dist2, pfx_correspondence* out,
                         size_t cap, size_t* n_out, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (dim <= 0 || !out || !n_out || (na && !a) || (nb && !b)) return ctx->fail(PFX_E_INVALID, "pfx_match: bad arguments");
  *n_out = 0;
  if (na == 0 || nb == 0) return 0;
  const float *da, *db;
  int lda, ldb;
  PFX_TRY(match_upload(ctx, a, na, stride_a, dim, mem, ctx->stage, &da, &lda));
  PFX_TRY(match_upload(ctx, b, nb, stride_b, dim, mem, ctx->stage2, &db, &ldb));
  PFX_CUDA(ctx->tmp0.ensure(na * (sizeof(int) + sizeof(float))));
  PFX_CUDA(ctx->tmp1.ensure(nb * (sizeof(int) + sizeof(float))));
  int* s2t = ctx->tmp0.as<int>();
  float* sd2 = reinterpret_cast<float*>(s2t + na);
  int* t2s = ctx->tmp1.as<int>();
  float* td2 = reinterpret_cast<float*>(t2s + nb);
  if (match_use_tc(ctx, na, nb, dim)) {
    PFX_TRY(match_pair_tc(ctx, da, (int)na, lda, db, (int)nb, ldb, dim, s2t, sd2, reciprocal ? t2s : nullptr, td2));
  } else {
    PFX_TRY(match_nn_exact(ctx, da, (int)na, lda, db, (int)nb, ldb, dim, s2t, sd2));
    if (reciprocal) PFX_TRY(match_nn_exact(ctx, db, (int)nb, ldb, da, (int)na, lda, dim, t2s, td2));
  }
  PFX_CUDA(ctx->tmp2.ensure(na * sizeof(int)));
  PFX_CUDA(ctx->tmp3.ensure(na * sizeof(int)));
  int* flags = ctx->tmp2.as<int>();
  int* pos = ctx->tmp3.as<int>();
  PFX_LAUNCH(ctx, reciprocal_kernel, div_up((long long)na, 256), 256, 0, s2t, sd2, t2s, (int)na, reciprocal, max_dist2, flags);
  PFX_CUDA(ctx->small.ensure(256));
  int* total = ctx->small.as<int>() + 16;
  PFX_TRY(scan_exclusive_i32(ctx, flags, pos, (int)na, total, ctx->scanbuf));
  int cnt = 0;
  PFX_CUDA(cudaMemcpyAsync(&cnt, total, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  *n_out = (size_t)cnt;
  if ((size_t)cnt > cap) return ctx->fail(PFX_E_CAPACITY, "pfx_match: correspondence buffer too small");
  pfx_correspondence* dout = out;
  if (mem == PFX_HOST) {
    PFX_CUDA(ctx->out_stage.ensure(std::max<size_t>(cnt, 1) * sizeof(pfx_correspondence)));
    dout = ctx->out_stage.as<pfx_correspondence>();
  }
  PFX_LAUNCH(ctx, corr_emit_kernel, div_up((long long)na, 256), 256, 0, flags, pos, s2t, sd2, (int)na, dout);
  PFX_CUDA(cudaGetLastError());
  if (mem == PFX_HOST) return deliver(ctx, out, dout, (size_t)cnt * sizeof(pfx_correspondence), mem);
  return 0;
}

// ================================================================================== RANSAC rejection
// device-resident correspondences (e.g. the output of pfx_match with PFX_DEVICE): count the out-of-range indices
__global__ void corr_validate_kernel(const pfx_correspondence* __restrict__ corr, int n, int n_src, int n_tgt,
                                     int* __restrict__ bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const pfx_correspondence c = corr[i];
  if (c.index_query < 0 || c.index_query >= n_src || c.index_match < 0 || c.index_match >= n_tgt) atomicAdd(bad, 1);
}

extern "C" int pfx_ransac_reject(pfx_ctx* ctx, const void* src, size_t n_src, size_t stride_src, const void* tgt,
                                 size_t n_tgt, size_t stride_tgt, const pfx_correspondence* corr, size_t n_corr,
                                 double inlier_threshold, int max_iterations, uint64_t seed, pfx_correspondence* out,
                                 size_t cap, size_t* n_out, float* transform16, int* iterations_out,
                                 int* best_hypothesis_out, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (!n_out || !transform16 || (n_corr && (!corr || !src || !tgt)) || stride_src < 12 || stride_tgt < 12 ||
      (stride_src & 3) || (stride_tgt & 3) || !(inlier_threshold > 0) || max_iterations < 1 || n_corr > 0x7fffffffull)
    return ctx->fail(PFX_E_INVALID, "pfx_ransac_reject: bad arguments");
  if (cap < n_corr || (n_corr && !out)) return ctx->fail(PFX_E_CAPACITY, "pfx_ransac_reject: output capacity must be >= n_corr");
  *n_out = 0;
  const float* dsrc = static_cast<const float*>(src);
  const float* dtgt = static_cast<const float*>(tgt);
  const pfx_correspondence* dcorr = corr;
  pfx_correspondence* dout = out;
  if (mem == PFX_HOST) {
    // correspondences index the clouds: validate on the host before anything is dereferenced on the device
    for (size_t i = 0; i < n_corr; ++i)
      if (corr[i].index_query < 0 || (size_t)corr[i].index_query >= n_src || corr[i].index_match < 0 ||
          (size_t)corr[i].index_match >= n_tgt)
        return ctx->fail(PFX_E_INVALID, "pfx_ransac_reject: correspondence index out of range");
    const size_t bs = n_src * stride_src, bt = n_tgt * stride_tgt, bc = n_corr * sizeof(pfx_correspondence);
    PFX_CUDA(ctx->stage.ensure(std::max<size_t>(bs, 16)));
    PFX_CUDA(ctx->stage2.ensure(std::max<size_t>(bt, 16)));
    PFX_CUDA(ctx->tmp1.ensure(std::max<size_t>(2 * bc, 16)));
    if (bs) PFX_CUDA(cudaMemcpyAsync(ctx->stage.p, src, bs, cudaMemcpyHostToDevice, ctx->stream));
    if (bt) PFX_CUDA(cudaMemcpyAsync(ctx->stage2.p, tgt, bt, cudaMemcpyHostToDevice, ctx->stream));
    if (bc) PFX_CUDA(cudaMemcpyAsync(ctx->tmp1.p, corr, bc, cudaMemcpyHostToDevice, ctx->stream));
    dsrc = ctx->stage.as<float>();
    dtgt = ctx->stage2.as<float>();
    dcorr = ctx->tmp1.as<pfx_correspondence>();
    dout = ctx->tmp1.as<pfx_correspondence>() + n_corr;
  } else if (n_corr) {
    // same validation for device buffers, before the hypothesis kernels dereference the indices
    PFX_CUDA(ctx->small.ensure(256));
    int* bad = ctx->small.as<int>() + 32;
    PFX_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), ctx->stream));
    PFX_LAUNCH(ctx, corr_validate_kernel, div_up((long long)n_corr, 256), 256, 0, dcorr, (int)n_corr,
               (int)std::min<size_t>(n_src, 0x7fffffff), (int)std::min<size_t>(n_tgt, 0x7fffffff), bad);
    int nbad = 0;
    PFX_CUDA(cudaMemcpyAsync(&nbad, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PFX_CUDA(cudaStreamSynchronize(ctx->stream));
    if (nbad) return ctx->fail(PFX_E_INVALID, "pfx_ransac_reject: correspondence index out of range");
  }
  int cnt = 0, iters = 0, bh = -1;
  PFX_TRY(ransac_reject_run(ctx, dsrc, stride_src, dtgt, stride_tgt, dcorr, (int)n_corr, inlier_threshold, max_iterations,
                            seed, dout, &cnt, transform16, &iters, &bh));
  *n_out = (size_t)cnt;
  if (iterations_out) *iterations_out = iters;
  if (best_hypothesis_out) *best_hypothesis_out = bh;
  if (mem == PFX_HOST) return deliver(ctx, out, dout, (size_t)cnt * sizeof(pfx_correspondence), mem);
  PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ================================================================================== ICP
extern "C" int pfx_icp_align(pfx_ctx* ctx, const void* src, size_t n_src, size_t stride_src, const pfx_icp_params* params,
                             const float* guess16, pfx_icp_result* result, void* aligned, size_t stride_aligned, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (!params || !result || (n_src && !src) || stride_src < 12 || (stride_src & 3) || n_src > 0x7fffffffull ||
      (aligned && (stride_aligned < 12 || (stride_aligned & 3))) || (mem != PFX_HOST && mem != PFX_DEVICE))
    return ctx->fail(PFX_E_INVALID, "pfx_icp_align: bad arguments");
  const float* dsrc = static_cast<const float*>(src);
  float* dal = static_cast<float*>(aligned);
  if (mem == PFX_HOST) {
    const size_t bs = n_src * stride_src;
    PFX_CUDA(ctx->stage.ensure(std::max<size_t>(bs, 16)));
    if (bs) PFX_CUDA(cudaMemcpyAsync(ctx->stage.p, src, bs, cudaMemcpyHostToDevice, ctx->stream));
    dsrc = ctx->stage.as<float>();
    if (aligned) {
      PFX_CUDA(ctx->out_stage.ensure(std::max<size_t>(n_src * 12, 16)));
      dal = ctx->out_stage.as<float>();
    }
  }
  PFX_TRY(icp_align_run(ctx, dsrc, (int)n_src, stride_src / 4, params, guess16, result, dal,
                        mem == PFX_HOST ? 3 : stride_aligned / 4));
  if (mem == PFX_HOST && aligned && n_src) {
    PFX_CUDA(cudaMemcpy2DAsync(aligned, stride_aligned, dal, 12, 12, n_src, cudaMemcpyDeviceToHost, ctx->stream));
    PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

// ================================================================================== range image / NARF
extern "C" int pfx_range_image_planar(pfx_ctx* ctx, int width, int height, float cx, float cy, float fx, float fy,
                                      float min_range, pfx_range_image_desc* desc_out) {
  PFX_TRY(check_ctx(ctx));
  if (ctx->surf_version == 0) return ctx->fail(PFX_E_PRECOND, "pfx_range_image_planar: no surface set");
  if (width <= 0 || height <= 0 || !(fx > 0) || !(fy > 0)) return ctx->fail(PFX_E_INVALID, "pfx_range_image_planar: bad geometry");
  pfx_range_image_desc d = {};
  d.width = width; d.height = height; d.planar = 1;
  d.cx = cx; d.cy = cy; d.fx = fx; d.fy = fy;
  PFX_TRY(range_image_build(ctx, &d, 0.f, 0.f, min_range, 0));
  if (desc_out) *desc_out = ctx->ri;
  return 0;
}

extern "C" int pfx_range_image_spherical(pfx_ctx* ctx, float ang_res, float max_angle_width, float max_angle_height,
                                         float min_range, int border, pfx_range_image_desc* desc_out) {
  PFX_TRY(check_ctx(ctx));
  if (ctx->surf_version == 0) return ctx->fail(PFX_E_PRECOND, "pfx_range_image_spherical: no surface set");
  if (!(ang_res > 0) || !(max_angle_width > 0) || !(max_angle_height > 0) || border < 0)
    return ctx->fail(PFX_E_INVALID, "pfx_range_image_spherical: bad geometry");
  pfx_range_image_desc d = {};
  d.planar = 0;
  d.ang_res = ang_res;
  d.fx = d.fy = 1.f;
  PFX_TRY(range_image_build(ctx, &d, max_angle_width, max_angle_height, min_range, border));
  if (desc_out) *desc_out = ctx->ri;
  return 0;
}

extern "C" int pfx_range_image_set(pfx_ctx* ctx, const pfx_range_image_desc* desc, const float* img, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (!desc || desc->width < 0 || desc->height < 0 || (desc->width * desc->height > 0 && !img))
    return ctx->fail(PFX_E_INVALID, "pfx_range_image_set: bad arguments");
  const size_t bytes = (size_t)desc->width * desc->height * sizeof(float4);
  PFX_CUDA(ctx->ri_img.ensure(std::max<size_t>(bytes, 16)));
  if (bytes)
    PFX_CUDA(cudaMemcpyAsync(ctx->ri_img.p, img, bytes, mem == PFX_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice,
                             ctx->stream));
  ctx->ri = *desc;
  ctx->ri_valid = true;
  ctx->ri_stage = 0;
  // with a sensor pose the caller's image is in world coordinates; the library works in the sensor frame, where the
  // xyz of a pixel follows from (pixel, range) exactly
  if (ctx->ri_has_pose) PFX_TRY(range_image_import_world(ctx));
  return 0;
}

// Sensor pose of the range image (keypoints.h:207-210: translation(sensor_origin_) * rotation(sensor_orientation_)):
// pose16 = row-major 4x4, world <- sensor; NULL = identity.  Applies to the images built or set AFTER the call.
extern "C" int pfx_range_image_set_pose(pfx_ctx* ctx, const float* pose16) {
  PFX_TRY(check_ctx(ctx));
  bool identity = true;
  if (pose16) {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) {
        const float v = pose16[4 * r + c];
        if (!std::isfinite(v)) return ctx->fail(PFX_E_INVALID, "pfx_range_image_set_pose: non-finite pose");
        if (v != (r == c ? 1.f : 0.f)) identity = false;
      }
    // a rigid pose: R R^T = I to float accuracy
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        double d = 0;
        for (int k = 0; k < 3; ++k) d += (double)pose16[4 * a + k] * pose16[4 * b + k];
        if (std::fabs(d - (a == b ? 1.0 : 0.0)) > 1e-4)
          return ctx->fail(PFX_E_INVALID, "pfx_range_image_set_pose: the rotation part is not orthonormal");
      }
  }
  ctx->ri_has_pose = !identity;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) ctx->ri_R[3 * r + c] = identity ? (r == c ? 1.f : 0.f) : pose16[4 * r + c];
    ctx->ri_t[r] = identity ? 0.f : pose16[4 * r + 3];
  }
  ctx->ri_valid = false;  // an image built under another pose is no longer current
  ctx->ri_stage = 0;
  return 0;
}

extern "C" int pfx_range_image_get(pfx_ctx* ctx, pfx_range_image_desc* desc_out, float* img, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (!ctx->ri_valid) return ctx->fail(PFX_E_STATE, "pfx_range_image_get: no range image");
  if (desc_out) *desc_out = ctx->ri;
  if (!img) return 0;
  const size_t bytes = (size_t)ctx->ri.width * ctx->ri.height * sizeof(float4);
  if (!ctx->ri_has_pose || bytes == 0) return deliver(ctx, img, ctx->ri_img.p, bytes, mem);
  // pcl::RangeImage::points are WORLD coordinates
  PFX_CUDA(ctx->tmp1.ensure(bytes));
  PFX_TRY(range_image_export_world(ctx, ctx->tmp1.as<float4>()));
  return deliver(ctx, img, ctx->tmp1.p, bytes, mem);
}

namespace pfx {
__global__ void narf_export_kernel(const float4* __restrict__ change, int np, float* __restrict__ score,
                                   float* __restrict__ dir) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  float4 c = change[i];
  if (score) score[i] = c.w;
  if (dir) {
    dir[3 * i] = c.x; dir[3 * i + 1] = c.y; dir[3 * i + 2] = c.z;
  }
}
}  // namespace pfx

extern "C" int pfx_narf_borders(pfx_ctx* ctx, int32_t* traits, float* border_scores, float* change_score,
                                float* change_dir, int mem) {
  PFX_TRY(check_ctx(ctx));
  PFX_TRY(narf_prepare(ctx, 1, 0.f));
  const size_t np = (size_t)ctx->ri.width * ctx->ri.height;
  if (np == 0) return 0;
  if (traits) PFX_TRY(deliver(ctx, traits, ctx->nb_traits.p, np * sizeof(int), mem));
  if (border_scores) PFX_TRY(deliver(ctx, border_scores, ctx->nb_scores.p, np * 4 * sizeof(float), mem));
  if (change_score || change_dir) {
    float* ds = change_score;
    float* dd = change_dir;
    if (mem == PFX_HOST) {
      PFX_CUDA(ctx->out_stage.ensure(np * 4 * sizeof(float)));
      ds = ctx->out_stage.as<float>();
      dd = ds + np;
    }
    PFX_LAUNCH(ctx, narf_export_kernel, div_up((long long)np, 256), 256, 0, ctx->nb_change.as<float4>(), (int)np,
               change_score ? ds : nullptr, change_dir ? dd : nullptr);
    PFX_CUDA(cudaGetLastError());
    if (mem == PFX_HOST) {
      if (change_score) PFX_TRY(deliver(ctx, change_score, ds, np * sizeof(float), mem));
      if (change_dir) PFX_TRY(deliver(ctx, change_dir, dd, np * 3 * sizeof(float), mem));
    }
  }
  return 0;
}

extern "C" int pfx_narf_keypoints(pfx_ctx* ctx, float support_size, int32_t* kp_px, float* kp_xyz, float* kp_interest,
                                  size_t cap, size_t* n_kp, float* interest_image, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (!(support_size > 0)) return ctx->fail(PFX_E_PRECOND, "pfx_narf_keypoints: support_size must be > 0");
  if (n_kp) *n_kp = 0;
  int* dkp = nullptr;
  int cnt = 0;
  PFX_TRY(narf_keypoints(ctx, support_size, &dkp, &cnt));
  if (n_kp) *n_kp = (size_t)cnt;
  const size_t np = (size_t)ctx->ri.width * ctx->ri.height;
  if (interest_image && np) PFX_TRY(deliver(ctx, interest_image, ctx->nk_interest.p, np * sizeof(float), mem));
  if (cnt == 0) return 0;
  if ((size_t)cnt > cap && (kp_px || kp_xyz || kp_interest)) return ctx->fail(PFX_E_CAPACITY, "pfx_narf_keypoints: keypoint buffer too small");
  if (kp_px) PFX_TRY(deliver(ctx, kp_px, dkp, (size_t)cnt * sizeof(int), mem));
  if (kp_xyz || kp_interest) PFX_TRY(narf_keypoint_attrs(ctx, dkp, cnt, kp_xyz, kp_interest, mem));
  return 0;
}

extern "C" int pfx_narf36(pfx_ctx* ctx, const int32_t* kp_px, size_t n_kp, float support_size, int rotation_invariant,
                          void* out, size_t stride, size_t cap, size_t* n_out, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (!(support_size > 0)) return ctx->fail(PFX_E_PRECOND, "pfx_narf36: support_size must be > 0");
  if (!n_out || (n_kp && !kp_px) || stride < 168 || (stride & 3) || (cap && !out))
    return ctx->fail(PFX_E_INVALID, "pfx_narf36: bad arguments");
  *n_out = 0;
  if (n_kp == 0) return 0;
  const int* dkp = kp_px;
  if (mem == PFX_HOST) {
    PFX_CUDA(ctx->stage.ensure(n_kp * sizeof(int)));
    PFX_CUDA(cudaMemcpyAsync(ctx->stage.p, kp_px, n_kp * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    dkp = ctx->stage.as<int>();
  }
  unsigned char* dout = static_cast<unsigned char*>(out);
  if (mem == PFX_HOST) {
    PFX_CUDA(ctx->out_stage.ensure(std::max<size_t>(cap * stride, 16)));
    dout = ctx->out_stage.as<unsigned char>();
    if (stride != 168 && cap) PFX_CUDA(cudaMemsetAsync(dout, 0, cap * stride, ctx->stream));
  }
  int cnt = 0;
  PFX_TRY(narf36_compute(ctx, dkp, (int)n_kp, support_size, rotation_invariant, dout, stride, (int)cap, &cnt));
  *n_out = (size_t)cnt;
  if (mem == PFX_HOST && cnt > 0) return deliver(ctx, out, dout, (size_t)cnt * stride, mem);
  return 0;
}

// ================================================================================== ingest
extern "C" int pfx_voxel_grid(pfx_ctx* ctx, float leaf, float* out_xyz, size_t cap, size_t* n_out, int mem) {
  PFX_TRY(check_ctx(ctx));
  if (ctx->surf_version == 0) return ctx->fail(PFX_E_PRECOND, "pfx_voxel_grid: no surface set");
  if (!(leaf > 0) || !n_out || (cap && !out_xyz)) return ctx->fail(PFX_E_INVALID, "pfx_voxel_grid: bad arguments");
  float* dout = out_xyz;
  if (mem == PFX_HOST) {
    PFX_CUDA(ctx->out_stage.ensure(std::max<size_t>(cap * 3 * sizeof(float), 16)));
    dout = ctx->out_stage.as<float>();
  }
  PFX_TRY(voxel_grid_run(ctx, leaf, dout, cap, n_out));
  if (*n_out > cap) return ctx->fail(PFX_E_CAPACITY, "pfx_voxel_grid: output buffer too small");
  if (mem == PFX_HOST) return deliver(ctx, out_xyz, dout, *n_out * 3 * sizeof(float), mem);
  return 0;
}
