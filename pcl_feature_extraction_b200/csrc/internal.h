// internal.h — host-side structures shared by the .cu files (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pfx_b200.h"
#include "common.cuh"

namespace pfx {

#define PFX_CUDA(call)                                                            \
  do {                                                                            \
    cudaError_t e__ = (call);                                                     \
    if (e__ != cudaSuccess) return ctx->fail_cuda(e__, #call, __FILE__, __LINE__); \
  } while (0)

#define PFX_TRY(call)          \
  do {                         \
    int rc__ = (call);         \
    if (rc__ != 0) return rc__; \
  } while (0)

// kernel launch + bookkeeping; every launch of ours goes through this
#define PFX_LAUNCH(ctx, kernel, grid, block, smem, ...)                      \
  do {                                                                       \
    bool pr__ = (ctx)->prof_on && (ctx)->prof_begin(#kernel);                \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);         \
    (ctx)->launches++;                                                       \
    if (pr__) (ctx)->prof_end();                                             \
  } while (0)

inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

struct Ctx;
struct Group;

// One voxel hash over the current surface (sorted copy of the points + cell table).
struct Grid {
  // identity
  uint64_t surf_version = 0;
  double radius = -1;  // radius grids: cell edge = radius * (1 + 1e-3)
  int knn_k = 0;       // kNN grids: cell edge estimated on the device from the point density
  uint64_t last_use = 0;
  // built ahead of use on the context's auxiliary stream (pfx_prepare_radius): `ready` is recorded behind the
  // build; the first consumer makes the main stream wait for it
  cudaEvent_t ready = nullptr;
  bool pending = false;
  // read-back of the device-chosen parameters (pinned) and the event recorded behind the build
  void* host_params = nullptr;
  cudaEvent_t built = nullptr;
  int n = 0;
  uint32_t hmask = 0;
  DevBuf params, pts, inv_perm, keys, vals, keys2, vals2, ghist, cell_start, cell_key, hkeys, hvals,
      pt_cell, cell_nbr, bsum, misc;
  GridDev view() const {
    GridDev g;
    g.gp = params.as<GridParams>();
    g.pts = pts.as<float4>();
    g.cell_start = cell_start.as<int>();
    g.cell_key = cell_key.as<uint32_t>();
    g.hkeys = hkeys.as<uint32_t>();
    g.hvals = hvals.as<int>();
    g.hmask = hmask;
    g.pt_cell = pt_cell.as<int>();
    g.cell_nbr = cell_nbr.as<int>();
    g.inv_perm = inv_perm.as<int>();
    g.n = n;
    return g;
  }
  void release() {
    for (DevBuf* b : {&params, &pts, &inv_perm, &keys, &vals, &keys2, &vals2, &ghist, &cell_start,
                      &cell_key, &hkeys, &hvals, &pt_cell, &cell_nbr, &bsum, &misc})
      b->release();
    if (host_params) cudaFreeHost(host_params);
    if (built) cudaEventDestroy(built);
    host_params = nullptr;
    built = nullptr;
  }
};

// one matrix prepared for the tensor-core matcher (match_tc.cu): bf16 core-matrix tiles + row statistics
struct TcOperand {
  DevBuf tiles, norm, err, maxima;
  int n = 0, npad = 0, dpad = 0;
  bool tf32 = false;  // element type of the tiles: tf32 (4 bytes) or bf16 (2 bytes)
};

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  uint64_t launches = 0;
  uint64_t tick = 0;
  int sm_count = 148;

  // surface (original order): float4 (x, y, z, index bits)
  size_t n = 0;
  uint64_t surf_version = 0;
  DevBuf surf;
  float vp[3] = {0, 0, 0};
  // surface normals, original order, float4 (nx, ny, nz, curvature)
  DevBuf normals;
  bool have_normals = false;
  // normals in the sorted order of one grid (cache, avoids the re-permute on the dense path)
  DevBuf normals_sorted;
  const Grid* normals_sorted_for = nullptr;
  uint64_t normals_version = 0, normals_sorted_version = 0;

  // queries: float4 in caller order; q_is_surface => the surface itself
  size_t nq = 0;
  bool q_is_surface = true;
  DevBuf qry;
  uint64_t qry_version = 0;

  std::vector<Grid*> grids;

  // kNN list cache (sorted-position indices + d2), valid for (grid, k, query version)
  DevBuf knn_idx, knn_d2;
  const Grid* knn_grid = nullptr;
  int knn_k = 0;
  uint64_t knn_qversion = 0, knn_sversion = 0;
  bool knn_dense = false;

  bool knn_sorted = false;  // rows ascending (d2, index)? (the cell-tile path writes unsorted sets)
  // qflag[i] != 0: dense query i was handed from the cell-tile path to the generic kernels
  DevBuf worklist2;  // work list of the fused SHOT kernel (same layout)
  DevBuf qflag, worklist;  // worklist: [0] = count, [16..] = sorted positions of the flagged queries
  bool tile_has_normals = false;

  // asynchronous result delivery (PFX_HOST_ASYNC): a copy stream and one staging slot per descriptor type;
  // ev_ready[s] = slot s has been produced (compute stream), ev_copied[s] = slot s has been copied out
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  bool copy_pending[2] = {false, false};
  DevBuf async_stage[2];

  // kernels that need > 48 KB of dynamic shared memory get the attribute once per context (= per device;
  // cudaFuncSetAttribute is a per-device setting, and a process may hold contexts on several devices)
  int knn_tile_blocks_per_sm = 0, shot_fused_blocks_per_sm[2] = {0, 0};  // resident blocks per SM (occupancy API, cached)
  bool smem_attr_knn_tile = false, smem_attr_shot_fused = false, smem_attr_match_tc = false, smem_attr_narf = false;
  Grid vg_scratch;  // sort buffers of pfx_voxel_grid

  // scratch
  DevBuf stage, stage2, tmp0, tmp1, tmp2, tmp3, tmp4, small, scanbuf, match_flags, match_best, out_stage;
  void* pinned = nullptr;
  size_t pinned_cap = 0;

  // range image + NARF state (narf.cu)
  pfx_range_image_desc ri = {};
  // sensor pose of the range image (pfx_range_image_set_pose): world <- sensor, rotation rows + translation.  The
  // image and every NARF stage live in the SENSOR frame; inputs / outputs are mapped from / to the world frame
  float ri_R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, ri_t[3] = {0, 0, 0};
  bool ri_has_pose = false;
  DevBuf ri_img, nb_surf, nb_scores, nb_shadow, nb_traits, nb_dir, nb_change, nk_interest;
  bool ri_valid = false;
  int ri_stage = 0;  // 0 image only, 1 borders extracted, 2 interest image for ri_support
  float ri_support = 0.f;

  cudaStream_t aux_stream = nullptr;  // index builds ahead of use (pfx_prepare_radius)
  cudaEvent_t ev_surface = nullptr;
  DevBuf icp_state, icp_cur, icp_nn, icp_partials;  // icp.cu
  // colours (shot_color.cu): normalised CIELab of the surface points / queries, valid for these versions
  DevBuf usc_tab;  // usc.cu
  DevBuf lab_tab, surf_lab, qry_lab;
  uint64_t surf_lab_version = 0, qry_lab_version = 0;

  // reuse of HOST inputs that are announced again unchanged (the reference re-submits the same cloud, and recomputes
  // its normals, for every descriptor type: features.h:186-193): fingerprint = pointer, size, stride and a hash of
  // the records
  struct HostFp {
    const void* ptr = nullptr;
    size_t n = 0, stride = 0;
    uint64_t hash = 0;
    bool valid = false;
  };
  bool reuse = true;
  HostFp surf_fp, nrm_fp;
  uint64_t nrm_fp_surf = 0, nrm_fp_version = 0;  // surface / normals version the normals fingerprint belongs to
  // what the resident dense normals were computed with (pfx_normals); valid while normals_version == nrm_key_version
  struct NrmKey {
    uint64_t surf = 0, version = 0;
    double radius = 0;
    int k = 0, parity = 0;
    float vp[3] = {0, 0, 0};
  } nrm_key;
  uint64_t stat_surface_uploads = 0, stat_surface_reused = 0, stat_normals_passes = 0, stat_normals_reused = 0,
           stat_normals_uploads = 0, stat_normals_upload_skipped = 0;

  // parity mode (pfx_set_parity_mode): 0 fast kernels (tolerance contract), 1 strict = reference-order arithmetic
  // for the stages whose floats feed index outputs (strict.cu)
  int parity_mode = 0;
  DevBuf st_cnt, st_off, st_idx, st_d2;  // sorted neighbour lists of strict.cu
  DevBuf surf_rgb, h6_inten, h6_grad;    // Harris 6D: packed 0x00RRGGBB of the surface points, intensity, gradients
  uint64_t surf_rgb_version = 0;
  int st_k = 0;
  long long st_total = 0;

  // multi-GPU (group.cu): NCCL group membership, slab surface, ring-match buffers
  struct Group* group = nullptr;
  DevBuf grp_tmp, slab_rows, slab_pack, slab_recv, slab_gid, slab_own, ring_buf[2], ring_best, ring_res;
  bool slab_active = false;
  int slab_axis = 0;
  size_t slab_owned = 0;
  long long slab_total = 0;
  double slab_lo = 0, slab_hi = 0;

  int match_engine = -1;  // -1 auto, 0 exact fp32 scan, 1 tcgen05 candidates + fp32 rescore
  TcOperand tc_ops[2];
  DevBuf tc_cand_d, tc_cand_j, tc_redo, tc_rows, tc_res;
  long long match_rows = 0, match_redo = 0, match_tc_calls = 0;  // statistics of the tensor-core matcher
  float knn_occupancy = 0.4f;  // target points per occupied cell of a kNN grid, as a fraction of k
  // dense SHOT takes its radius neighbourhoods from the resident k-search rows when it can (shot_fused.cu);
  // PFX_SHOT_ROWS=0 in the environment keeps the stencil walk on a radius grid (A/B measurements)
  bool shot_from_rows = true;
  int tc_pair = 1;  // tensor-core matcher on CTA pairs (cta_group::2): 1 on (default), 0 off (PFX_TC_PAIR=0)
  struct RowsStat {  // share of the k-search rows the last rows-based SHOT call could not close (asynchronous read-back)
    int* host = nullptr;
    cudaEvent_t ev = nullptr;
    bool pending = false;
    double radius = 0, open_frac = 0;
    int k = 0;
    size_t n = 0;
  } rows_stat;
  Grid* last_grid = nullptr;

  // optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline leg)
  struct ProfRec {
    const char* name;
    cudaEvent_t e0, e1;
  };
  bool prof_on = false;
  std::string prof_filter;
  std::vector<ProfRec> prof_recs;
  size_t prof_used = 0;
  bool prof_begin(const char* name) {
    if (!prof_filter.empty() && !strstr(name, prof_filter.c_str())) return false;
    if (prof_used == prof_recs.size()) {
      ProfRec r;
      r.name = name;
      if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return false;
      prof_recs.push_back(r);
    }
    prof_recs[prof_used].name = name;
    cudaEventRecord(prof_recs[prof_used].e0, stream);
    return true;
  }
  void prof_end() {
    cudaEventRecord(prof_recs[prof_used].e1, stream);
    prof_used++;
  }

  int fail(int code, const std::string& msg) {
    err = msg;
    return code;
  }
  int fail_cuda(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file,
             line, what);
    err = buf;
    return (int)e;
  }
  size_t num_queries() const { return q_is_surface ? n : nq; }
};

// ---- grid.cu
int grid_get(Ctx* ctx, double radius, int knn_k, Grid** out);
int sort_pairs_scratch(Ctx* ctx, int n, uint32_t** keys, int** vals);  // scratch (key, value) buffers ...
int sort_pairs_scratch_run(Ctx* ctx, int n);                           // ... sorted ascending by key in place
int grid_for_radius(Ctx* ctx, double radius, Grid** out);  // exact, or any grid of the surface whose edge covers the radius
int grid_prepare_async(Ctx* ctx, double radius);  // build the radius grid on the auxiliary stream
int grid_wait_pending(Ctx* ctx);                  // main stream waits for every build in flight
void grid_free_all(Ctx* ctx);
int voxel_grid_run(Ctx* ctx, float leaf, float* out_dev, size_t cap, size_t* n_out);
// device-wide exclusive scan of int32 -> int32 / int64 (total written to *total_dev when non-null)
int scan_exclusive_i32(Ctx* ctx, const int* in, int* out, int n, int* total_dev, DevBuf& bsum);
int scan_exclusive_i64(Ctx* ctx, const int* in, long long* out, int n, DevBuf& bsum);

// ---- search.cu
int knn_run(Ctx* ctx, Grid* g, const float4* q_dev, int nq, int k, int* idx_dev, float* d2_dev);
int knn_lists(Ctx* ctx, Grid* g, int k, bool need_sorted_ids);  // fills ctx->knn_idx / knn_d2
int knn_run_worklist(Ctx* ctx, Grid* g, int k, int* idx_dev, float* d2_dev, const int* worklist, const int* wl_count);
// ---- knn_tile.cu
int knn_tile_lists(Ctx* ctx, Grid* g, int k, bool with_normals);
int knn_export(Ctx* ctx, int k, int32_t* idx, float* d2, int mem);
int radius_count(Ctx* ctx, Grid* g, double radius, int* counts_dev);
int radius_fill(Ctx* ctx, Grid* g, double radius, int sorted, const long long* offsets_dev,
                int* idx_dev, float* d2_dev);

// ---- normals.cu
int normals_compute(Ctx* ctx, Grid* g, double radius, int k, float4* out_query_order);

// ---- fpfh.cu
int fpfh_compute(Ctx* ctx, Grid* g, double radius, int k, float* out_dev, size_t stride_floats,
                 float* spfh_out_dev);

// ---- shot.cu
int shot_lrf_compute(Ctx* ctx, Grid* g, double radius, float* rf9_dev, int* nvalid_dev);
int shot_compute(Ctx* ctx, Grid* g, double radius, const float* rf9_dev, float* out_dev,
                 size_t stride_floats);
// ---- shot_fused.cu
int shot_fused_compute(Ctx* ctx, Grid* g, double radius, float* out_dev, size_t stride_floats);  // g may be null
bool shot_rows_available(Ctx* ctx, double radius);

// ---- keypoints.cu
int cloud_resolution(Ctx* ctx, double* res);
int iss_saliency(Ctx* ctx, Grid* g, double radius, int min_nb, double g21, double g32,
                 double* sal_dev_orig);
int iss_nms(Ctx* ctx, Grid* g, const double* sal_dev_orig, double radius, int min_nb, int* flags_dev);
int harris_response(Ctx* ctx, Grid* g, double radius, float* resp_dev_orig);
int harris_nms(Ctx* ctx, Grid* g, const float* resp_dev_orig, double radius, float thr, int* flags_dev);
int harris_refine(Ctx* ctx, Grid* g, double radius, float* corners_dev, int nc);
int snap_to_cloud(Ctx* ctx, const float* q_dev, int nq, float max_d2, int* out_dev);
int compact_flags(Ctx* ctx, const int* flags_dev, int n, int* idx_out_dev, int* count_host);

// ---- match.cu
int match_nn_exact(Ctx* ctx, const float* a, int na, int lda, const float* b, int nb, int ldb, int dim,
                   int* nn_idx, float* nn_d2);
// ---- match_tc.cu
int match_nn_tc(Ctx* ctx, const float* a, int na, int lda, const float* b, int nb, int ldb, int dim,
                int* nn_idx, float* nn_d2);
int match_pair_tc(Ctx* ctx, const float* a, int na, int lda, const float* b, int nb, int ldb, int dim, int* s2t,
                  float* sd2, int* t2s, float* td2);
void match_tc_release(Ctx* ctx);
bool match_tc_fits(int dim);  // the resident A tile + two ring stages fit in shared memory

// ---- narf.cu
int range_image_build(Ctx* ctx, const pfx_range_image_desc* want, float max_angle_w, float max_angle_h, float min_range,
                      int border);
int narf_prepare(Ctx* ctx, int stage, float support_size);
int range_image_export_world(Ctx* ctx, float4* out_dev);
int range_image_import_world(Ctx* ctx);
int narf_keypoints(Ctx* ctx, float support_size, int** kp_dev, int* n_kp);
int narf_keypoint_attrs(Ctx* ctx, const int* kp_dev, int n, float* xyz, float* val, int mem);
int narf36_compute(Ctx* ctx, const int* kp_dev, int n_kp, float support_size, int rotation_invariant,
                   unsigned char* out_dev, size_t stride, int cap, int* n_out);

// ---- ransac.cu
int ransac_reject_run(Ctx* ctx, const float* src, size_t stride_s, const float* tgt, size_t stride_t,
                      const pfx_correspondence* corr, int n_corr, double threshold, int max_iterations, uint64_t seed,
                      pfx_correspondence* out_dev, int* n_out, float* T16_host, int* iterations, int* best_h);

// ---- strict.cu (reference-order arithmetic, pfx_set_parity_mode)
int strict_lists_build(Ctx* ctx, Grid* g, double radius, int k);
int strict_normals(Ctx* ctx, Grid* g, double radius, int k, float4* out_query_order);
int harris_response_strict(Ctx* ctx, Grid* g, double radius, float* resp_dev_orig);
int harris_refine_strict(Ctx* ctx, Grid* g, double radius, float* corners_dev, int nc);
int harris6d_response(Ctx* ctx, Grid* g, double radius, float* resp_dev_orig, float* grad_out_dev);
int fpfh_sorted(Ctx* ctx, Grid* g, double radius, const float* spfh_sorted_rows, float* out_dev, size_t stride_floats);

// ---- group.cu
void group_release(Ctx* ctx);

// ---- helpers (capi.cu)
int pfh_compute(Ctx* ctx, Grid* g, double radius, int k, float* out_dev, size_t stride_floats);
int moments_compute(Ctx* ctx, Grid* g, double radius, int k, float* out_dev, size_t stride_floats);
int curvature_compute(Ctx* ctx, Grid* g, double radius, int k, float* out_dev, size_t stride_floats);
int colors_to_lab(Ctx* ctx, const unsigned char* rgb_dev, size_t stride_bytes, int n, DevBuf& lab);
int shot_color_compute(Ctx* ctx, Grid* g, double radius, const float* rf9_dev, float* out_dev, size_t stride_floats);
int usc_compute(Ctx* ctx, double search_radius, double min_radius, double density_radius, double local_radius,
                const float* lrf_dev, float* out_dev, size_t stride_floats);
int sc3d_compute(Ctx* ctx, double search_radius, double min_radius, double density_radius, unsigned long long seed,
                 float* out_dev, size_t stride_floats, float* frames_out_dev);
int spin_compute(Ctx* ctx, Grid* g, double radius, const float* qnormals_dev, size_t nstride_floats, float* out_dev,
                 size_t stride_floats);
int icp_align_run(Ctx* ctx, const float* src_dev, int n, size_t stride_floats, const pfx_icp_params* prm,
                  const float* guess16, pfx_icp_result* res, float* aligned_dev, size_t aligned_stride_floats);
int normals_sorted_for_grid(Ctx* ctx, Grid* g, const float4** out);
int ensure_pinned(Ctx* ctx, size_t bytes);

}  // namespace pfx

struct pfx_ctx : public pfx::Ctx {};
