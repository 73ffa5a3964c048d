/*
 * pfx_b200.h — C ABI of the B200-native point-cloud feature-extraction hot path.
 *
 * This is the drop-in boundary for the path that the reference (srv/pcl_feature_extraction)
 * drives through PCL: NormalEstimation -> ISS / Harris3D / NARF keypoints -> FPFH33 / SHOT352 /
 * Narf36 descriptors -> descriptor correspondence matching.  Every entry point names the reference
 * call site it replaces (paths are under /root/reference: include/pcl_feature_extraction/{features,keypoints,tools}.h and
 * src/evaluation.cpp).  The header-only C++ shim in pcl_feature_extraction_b200/host/pcl_compat.hpp
 * puts the pcl::Feature-style classes (setInputCloud / setSearchSurface / setRadiusSearch /
 * setKSearch / compute) on top of these calls; INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - plain pointers and sizes only; `mem` says where a buffer lives (PFX_HOST or PFX_DEVICE).
 *    Device buffers must be on the context's device.
 *  - point buffers are arrays of records whose first three floats are x, y, z; `stride` is the
 *    record size in bytes (16 for PointXYZ, 32 for PointXYZRGB / PointXYZI, 12 for packed xyz).
 *  - normal buffers are records of (nx, ny, nz, pad, curvature, ...) like pcl::Normal (stride 32)
 *    or packed (nx, ny, nz, curvature) float4 (stride 16): `curv_off` is the float offset of the
 *    curvature inside a record (4 for pcl::Normal, 3 for float4).
 *  - all indices are 32-bit and refer to the ORIGINAL order of the surface / query buffers.
 *  - return value: 0 = ok; < 0 = PFX_E_* (exactly where PCL's initCompute() would print an error
 *    and return an empty cloud); > 0 = cudaError_t.  No exceptions cross this boundary.
 *  - a context is bound to one device and may be used by one host thread at a time.  Calls are
 *    synchronous at return when any output is PFX_HOST; with PFX_DEVICE outputs they are enqueued
 *    on the context's stream (pfx_set_stream) and complete in stream order (pfx_sync to wait).
 *  - there is NO CPU fallback: every function fails with a CUDA error when no sm_100 device exists.
 */
#ifndef PFX_B200_H
#define PFX_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct pfx_ctx pfx_ctx;

/* PFX_HOST_ASYNC (accepted for the OUTPUT of pfx_fpfh and pfx_shot352): like PFX_HOST, but the device-to-host
 * copy is enqueued on the context's copy stream behind the producing kernels and the call returns at once;
 * the rows are valid after pfx_sync().  Lets the PCIe transfer of one result overlap the compute of the next
 * call / the next cloud (the buffer should be page-locked, otherwise the copy is staged by the driver). */
enum { PFX_HOST = 0, PFX_DEVICE = 1, PFX_HOST_ASYNC = 2 };
enum {
  PFX_OK = 0,
  PFX_E_INVALID = -1,  /* bad argument */
  PFX_E_PRECOND = -2,  /* PCL initCompute() precondition (no surface, both/neither radius and k, ...) */
  PFX_E_CAPACITY = -3, /* caller buffer too small */
  PFX_E_STATE = -4     /* required input (normals, range image, ...) not set */
};

/* pcl::Correspondence {int index_query; int index_match; float distance} (features.h:245-249) */
typedef struct { int32_t index_query; int32_t index_match; float distance; } pfx_correspondence;

/* ------------------------------------------------------------------ context */
int pfx_version(void);
int pfx_create(int device, pfx_ctx** ctx);
int pfx_destroy(pfx_ctx* ctx);
const char* pfx_last_error(const pfx_ctx* ctx);
int pfx_set_stream(pfx_ctx* ctx, void* cuda_stream); /* cudaStream_t; NULL = default stream */
int pfx_sync(pfx_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t pfx_launch_count(const pfx_ctx* ctx);
/* per-kernel device timing with CUDA events on the context's stream (replaces the reference's
 * ros::WallTime stop-watches, evaluation.cpp:278-283,596-603).  filter: substring of the kernel
 * name to time, NULL/"" = every kernel.  pfx_profile_end waits for the stream and writes one line
 * per kernel "<name>\t<launches>\t<total ms>\n" into buf. */
/* diagnostics of the voxel hash used by the last call: out8 = {cell edge, nx, ny, nz, occupied cells,
 * finite points, queries handed from the cell-tile path to the generic kernels (-1: no tile pass), 0} */
int pfx_grid_info(pfx_ctx* ctx, double* out8);
/* tuning: target points per occupied cell of kNN grids as a fraction of k (default 0.4) */
int pfx_set_knn_occupancy(pfx_ctx* ctx, float fraction_of_k);
int pfx_profile_begin(pfx_ctx* ctx, const char* filter);
int pfx_profile_end(pfx_ctx* ctx, char* buf, size_t buflen);

/* ------------------------------------------------------------------ inputs
 * pfx_set_surface   <- Feature::setSearchSurface + search::KdTree::setInputCloud
 *                      (features.h:190,192-193; tools.h:27,29-30; keypoints.h:185-187)
 * pfx_set_queries   <- Feature::setInputCloud(keypoints) (features.h:191); n = 0 => queries = surface
 * pfx_set_surface_normals <- FeatureFromNormals::setInputNormals (features.h:188)
 * pfx_set_viewpoint <- cloud.sensor_origin_ used by flipNormalTowardsViewpoint (default 0,0,0) */
int pfx_set_surface(pfx_ctx* ctx, const void* pts, size_t n, size_t stride, int mem);
/* Reuse of unchanged HOST inputs (default on).  The reference announces the same cloud to every Feature object,
 * builds a fresh kd-tree for it and recomputes the normals of the whole cloud per descriptor type
 * (features.h:186-193, inside the loops of evaluation.cpp:272,302).  Here a PFX_HOST cloud whose fingerprint
 * (pointer, size, stride, hash of every record up to 262144 points, of 8192 evenly spaced ones beyond) equals the
 * resident surface's is NOT uploaded
 * again: its voxel hashes, kNN lists and normals stay.  pfx_normals with the parameters of the resident dense normals
 * only delivers them; pfx_set_surface_normals of the very buffer pfx_normals filled (or of one uploaded before) is a
 * no-op.  A cloud of more than 262144 points edited IN PLACE between two calls must be announced with reuse switched
 * off (pfx_set_reuse(ctx, 0)) or after a call with n = 0.  pfx_reuse_info out6: surface uploads, surface announcements answered from the resident
 * copy, dense normals passes, pfx_normals calls answered from the resident normals, normals uploads, normals
 * uploads skipped. */
int pfx_set_reuse(pfx_ctx* ctx, int enable);
int pfx_reuse_info(const pfx_ctx* ctx, uint64_t* out6);
int pfx_set_queries(pfx_ctx* ctx, const void* pts, size_t n, size_t stride, int mem);
/* Optional hint: build the search index of `radius` now, on an auxiliary stream, behind the surface upload.  A
 * later radius stage with the same radius (pfx_shot352, pfx_fpfh, pfx_radius_*, ...) finds it ready instead of
 * building it in line; everything enqueued in between overlaps the build.  Results are unaffected. */
int pfx_prepare_radius(pfx_ctx* ctx, double radius);
int pfx_set_surface_normals(pfx_ctx* ctx, const void* normals, size_t n, size_t stride,
                            int curv_off, int mem);
int pfx_set_viewpoint(pfx_ctx* ctx, float vx, float vy, float vz);
size_t pfx_num_surface(const pfx_ctx* ctx);
/* the points of the current surface in its own order, rows of `stride` bytes with xyz first (after
 * pfx_slab_distribute: this rank's owned points, then the halo points) */
int pfx_get_surface(pfx_ctx* ctx, void* out, size_t stride, int mem);
size_t pfx_num_queries(const pfx_ctx* ctx);

/* ------------------------------------------------------------------ neighbour search
 * replaces pcl::KdTreeFLANN / pcl::search::KdTree radiusSearch / nearestKSearch
 * (features.h:192-193, keypoints.h:371-386, 408-417).  d2 = ((dx*dx+dy*dy)+dz*dz) in float, no FMA;
 * radius membership d2 < (float)(radius*radius); kNN order and tie-break ascending (d2, index). */
int pfx_knn(pfx_ctx* ctx, int k, int32_t* idx, float* d2, int mem); /* nq x k, padded -1 / +inf */
int pfx_radius_count(pfx_ctx* ctx, double radius, int32_t* counts, int64_t* total, int mem);
/* offsets: nq+1 exclusive prefix of counts (int64, same `mem`); sorted != 0 => ascending (d2, index) */
int pfx_radius_search(pfx_ctx* ctx, double radius, int sorted, const int64_t* offsets,
                      int32_t* idx, float* d2, int mem);

/* ------------------------------------------------------------------ normals
 * pfx_normals <- NormalEstimationOMP::compute (tools.h:26-31; features.h:187; keypoints.h:302-308).
 * Exactly one of radius / k non-zero.  Output row i belongs to query i; NaN row when the query is
 * non-finite or has no neighbours.  out may be NULL.  When the queries are the surface the result
 * also becomes the surface's normals (the setInputNormals that follows at features.h:188), so the
 * pair costs no host round trip. */
int pfx_normals(pfx_ctx* ctx, double radius, int k, void* out, size_t stride, int curv_off, int mem);

/* ------------------------------------------------------------------ keypoints
 * pfx_cloud_resolution <- Keypoints::computeCloudResolution (keypoints.h:401-428) */
int pfx_cloud_resolution(pfx_ctx* ctx, double* resolution);
/* pfx_iss <- ISSKeypoint3D::compute as configured at keypoints.h:184-194.  kp_idx: ascending
 * surface indices (capacity cap); saliency (optional, n doubles, `mem`): third_eigen_value. */
int pfx_iss(pfx_ctx* ctx, double salient_radius, double nonmax_radius, int min_neighbors,
            double gamma21, double gamma32, int32_t* kp_idx, size_t cap, size_t* n_kp,
            double* saliency, int mem);
/* NMS stage alone on caller-supplied saliency (stage-wise parity) */
int pfx_iss_nms(pfx_ctx* ctx, const double* saliency, double nonmax_radius, int min_neighbors,
                int32_t* kp_idx, size_t cap, size_t* n_kp, int mem);
/* pfx_harris3d <- HarrisKeypoint3D::compute as configured at keypoints.h:154-159 (radius default
 * 0.01, internal NormalEstimation at the same radius when no normals were set) followed by the
 * reference's snap Keypoints::getKeypointsCloud keypoints.h:360-395 (snap_max_d2 = 1e-4).
 * Outputs (each optional): response n floats; kp_idx surface index of every NMS maximum (ascending);
 * kp_xyz refined corner positions (cap x 3); snapped_idx cloud index after the snap or -1. */
int pfx_harris3d(pfx_ctx* ctx, double radius, float threshold, int nonmax, int refine,
                 float snap_max_d2, float* response, int32_t* kp_idx, float* kp_xyz,
                 int32_t* snapped_idx, size_t cap, size_t* n_kp, int mem);
int pfx_harris_nms(pfx_ctx* ctx, const float* response, double radius, float threshold,
                   int32_t* kp_idx, size_t cap, size_t* n_kp, int mem);
/* pfx_harris6d <- HarrisKeypoint6D<PointXYZRGB, PointXYZI>::compute (keypoints.h:166-179; in the reference's active
 * detector list, evaluation.cpp:63-65): same arguments and outputs as pfx_harris3d.  response = the 4th smallest
 * eigenvalue of the 6x6 covariance of (normal, normalised intensity gradient) over the neighbourhood; the normals are
 * estimated internally at `radius`, the intensity comes from the colours given with pfx_set_surface_colors
 * (0.00390625 (0.114 b + 0.5870 g + 0.2989 r)), the gradient is IntensityGradientEstimation's at `radius`.  All of
 * it runs in reference order (strict.cu); where upstream leans on Eigen internals that cannot be pinned (its float
 * column-pivoting QR, its 6x6 tridiagonal eigen solver) the library follows the written-out rules of DESIGN.md
 * section 3, which the CPU restatement follows too.  PFX_E_STATE without colours. */
int pfx_harris6d(pfx_ctx* ctx, double radius, float threshold, int nonmax, int refine, float snap_max_d2, float* response,
                 int32_t* kp_idx, float* kp_xyz, int32_t* snapped_idx, size_t cap, size_t* n_kp, int mem);

/* ------------------------------------------------------------------ descriptors
 * pfx_fpfh <- FPFHEstimation::compute (evaluation.cpp:597-602 via features.h:190-195): SPFH over the
 * union of the queries' neighbourhoods, then the 1/d2-weighted gather.  out: nq rows of 33 floats at
 * `stride` bytes (132 for FPFHSignature33).  Requires surface normals. */
int pfx_fpfh(pfx_ctx* ctx, double radius, int k, float* out, size_t stride, int mem);
/* SPFH rows of all surface points (stage-wise parity): n x 33 floats, packed */
int pfx_spfh(pfx_ctx* ctx, double radius, int k, float* out, int mem);
/* pfx_shot352 <- SHOTEstimationOMP::compute incl. its internal SHOTLocalReferenceFrameEstimation
 * (evaluation.cpp:770-775).  Radius search only (k-search is rejected like PCL does).  out: nq rows
 * at `stride` bytes holding descriptor[352] then rf[9] (1444 for pcl::SHOT352).  lrf_in (optional,
 * nq x 9 packed, `mem`): use these frames instead of estimating them. */
int pfx_shot352(pfx_ctx* ctx, double radius, const float* lrf_in, float* out, size_t stride, int mem);
int pfx_shot_lrf(pfx_ctx* ctx, double radius, float* rf9, int mem); /* nq x 9 */

/* ------------------------------------------------------------------ matching
 * pfx_match <- Features<T>::findCorrespondences / getCorrespondences (features.h:224-273) and
 * pcl::registration::CorrespondenceEstimation::determine[Reciprocal]Correspondences.
 * a: na x dim (source), b: nb x dim (target), row strides in bytes.  Distances are the sequential
 * float sum of squares of FLANN's L2_Simple; ties -> lowest index; NaN rows never match.
 * reciprocal != 0 keeps i only when nn_b(nn_a(i)) == i.  max_dist2 < 0 = no limit, else keep
 * d2 <= max_dist2.  out capacity cap (<= na needed). */
int pfx_match(pfx_ctx* ctx, const float* a, size_t na, size_t stride_a, const float* b, size_t nb,
              size_t stride_b, int dim, int reciprocal, float max_dist2, pfx_correspondence* out,
              size_t cap, size_t* n_out, int mem);
/* one-directional exact 1-NN (nn_idx, nn_d2 sized na) */
int pfx_match_nn(pfx_ctx* ctx, const float* a, size_t na, size_t stride_a, const float* b, size_t nb,
                 size_t stride_b, int dim, int32_t* nn_idx, float* nn_d2, int mem);
/* engine: 0 = exact fp32 scan, 1 = tcgen05 bf16 candidate GEMM + fp32 rescore with certificate
 * (falls back to the exact scan per row when the certificate fails), -1 = automatic */
int pfx_set_match_engine(pfx_ctx* ctx, int engine);
/* statistics of the tensor-core matcher: out4[0] 1-NN passes run on the tensor cores, out4[1] query rows
 * they processed, out4[2] rows whose exactness certificate failed and were redone by the exact scan */
int pfx_match_info(pfx_ctx* ctx, double* out4);

/* ------------------------------------------------------------------ multi-GPU: groups, slabs, sharded matching
 * The reference is a single CPU process; its per-point stages (features.h:181-195) shard over
 * GPUs either cloud by cloud (no communication) or, for ONE large cloud, as spatial slabs with a halo.  A context
 * joins a GROUP = one NCCL communicator with one rank per GPU; ranks may be processes (one per GPU, the id
 * broadcast by the launcher) or threads of one process (one context per device, pfx_group_join from each thread).
 * Collective calls (pfx_slab_distribute, pfx_match_ring, pfx_group_allreduce) must be made by every rank.
 *
 * pfx_group_unique_id   fills 128 bytes (ncclUniqueId) on one rank; hand them to all ranks
 * pfx_group_join        ncclCommInitRank for this context's device
 * pfx_slab_distribute   the ranks hold arbitrary disjoint parts of one cloud (records of `stride` bytes, xyz first;
 *                       global_ids optional int32 per point, unique, default = position in the concatenation of the
 *                       parts in rank order).  The cloud is cut into `world` slabs along its longest axis at
 *                       equal-count cuts; on return the surface of this context = the n_owned points of its slab and
 *                       every point within `halo` of it (n_local rows in all), moved device to device by grouped
 *                       ncclSend / ncclRecv and held in ASCENDING GLOBAL-ID order, so that every (distance, index)
 *                       tie-break of the stages resolves as it does on one GPU and sharded rows equal single-GPU rows
 *                       bit for bit.  With halo >= the support of the stage chain
 *                       (normals r_n; FPFH 2 r_f + r_n; SHOT r_s + r_n; k-searches: multiples of the largest k-th
 *                       neighbour distance) the dense stages give every owned point the rows it gets on one GPU.
 * pfx_slab_owned_rows   int32 [n_owned]: the local rows (ascending) that belong to this rank's slab
 * pfx_slab_global_ids   int32 [n_local]: the global id of every local surface point (ascending)
 * pfx_match_ring        exact 1-NN with both descriptor sets sharded: every rank passes its query rows `a` and its
 *                       target block `b` (first global row b_offset); target blocks rotate around the ring under the
 *                       match; nn_idx = GLOBAL target row (-1 none), ties -> lowest global row, as pfx_match_nn.
 * pfx_group_allreduce   n host doubles reduced in place over the ranks: op 0 sum, 1 max, 2 min */
#define PFX_GROUP_ID_BYTES 128
int pfx_group_unique_id(void* id128);
int pfx_group_join(pfx_ctx* ctx, int rank, int world, const void* id128);
int pfx_group_leave(pfx_ctx* ctx);
int pfx_group_info(const pfx_ctx* ctx, int* rank, int* world);
int pfx_group_allreduce(pfx_ctx* ctx, double* vals, int n, int op);
int pfx_slab_distribute(pfx_ctx* ctx, const void* part, size_t n_part, size_t stride, const int32_t* global_ids, int mem,
                        double halo, size_t* n_owned, size_t* n_local);
int pfx_slab_owned_rows(pfx_ctx* ctx, int32_t* out, int mem);
int pfx_slab_global_ids(pfx_ctx* ctx, int32_t* out, int mem);
/* info6: [0] axis, [1] n_owned, [2] n_local, [3] points over all ranks, [4] / [5] lower / upper bound of the slab */
int pfx_slab_info(const pfx_ctx* ctx, double* info6);
int pfx_match_ring(pfx_ctx* ctx, const float* a, size_t na, size_t stride_a, const float* b, size_t nb, size_t stride_b,
                   int dim, int b_offset, int32_t* nn_idx, float* nn_d2, int mem);

/* ------------------------------------------------------------------ parity mode
 * PFX_PARITY_FAST (default): the throughput kernels; floats within the tolerances of DESIGN.md section 4.
 * PFX_PARITY_STRICT: the stages whose floats feed INDEX outputs of the reference pipeline run in reference-order
 * arithmetic (strict.cu) - neighbours in ascending (d2, index) order, sequential sums, the CPU restatement's
 * eigen solver, no FMA contraction - so that NormalEstimation (tools.h:26-31), HarrisKeypoint3D response /
 * refinement / snap (keypoints.h:154-162, :360-395) and radius-search FPFH at keypoints (evaluation.cpp:597-602)
 * are bit-identical to the CPU path and the keypoint / correspondence indices of Features::findCorrespondences
 * (features.h:240-250) agree end to end, degenerate neighbourhoods included.  Meant for the reference's own
 * keypoint pipelines (1e4 .. 1e5 points); k-search FPFH, SHOT and the dense 1M-point path keep the fast kernels. */
enum { PFX_PARITY_FAST = 0, PFX_PARITY_STRICT = 1 };
int pfx_set_parity_mode(pfx_ctx* ctx, int mode);

/* ------------------------------------------------------------------ PFH125, PrincipalCurvatures (next rows)
 * pfx_pfh125 <- PFHEstimation<PointXYZRGB, Normal, PFHSignature125>::compute (evaluation.cpp:676-695 through
 * features.h:181-195): rows of 125 floats (pcl::PFHSignature125, 500 B) for the current queries; same
 * preconditions and search parameters as pfx_fpfh.  A row is NaN when the query has no neighbours.
 * pfx_principal_curvatures <- PrincipalCurvaturesEstimation<PointXYZRGB, Normal, PrincipalCurvatures>::compute
 * (evaluation.cpp:696-715): rows of 5 floats (pcl::PrincipalCurvatures: principal_curvature_x/y/z, pc1, pc2).
 * The tangent plane of query i is that of surface normal i - upstream indexes the normals with the query's
 * ordinal - which is the query's own normal when the queries are the surface. */
int pfx_pfh125(pfx_ctx* ctx, double radius, int k, float* out, size_t stride, int mem);
int pfx_principal_curvatures(pfx_ctx* ctx, double radius, int k, float* out, size_t stride, int mem);
/* pfx_moment_invariants <- MomentInvariantsEstimation<PointXYZRGB, MomentInvariants>::compute (evaluation.cpp:555-574):
 * rows of 3 floats (pcl::MomentInvariants: j1, j2, j3) from the central second moments of each neighbourhood;
 * needs no normals. */
int pfx_moment_invariants(pfx_ctx* ctx, double radius, int k, float* out, size_t stride, int mem);
/* the float PCL's `hist[bin] += incr` holds after `count` additions (IEEE binary32, sequential): turns integer vote
 * counts (pfx_spfh count rows, PFH votes) into PCL's histogram values bit for bit.  Pure host function. */
float pfx_seq_float_sum(float incr, long long count);

/* ------------------------------------------------------------------ SHOT1344: shape + colour (next row)
 * pfx_shot1344 <- SHOTColorEstimation<PointXYZRGB, Normal, SHOT1344>::compute (evaluation.cpp:786-805 through
 * features.h:181-195).  Colours are the packed 0x00RRGGBB words of pcl::PointXYZRGB (pass &points[0].rgba with
 * stride sizeof(PointXYZRGB)); pfx_set_surface_colors after pfx_set_surface (one per surface point),
 * pfx_set_query_colors after pfx_set_queries (the reference colour of a descriptor is its query point's).
 * Rows = pcl::SHOT1344: descriptor[1344] (352 shape slots, then 32 x 31 colour slots) + rf[9], 5412 B.
 * lrf_in as for pfx_shot352. */
int pfx_set_surface_colors(pfx_ctx* ctx, const void* rgb, size_t n, size_t stride, int mem);
int pfx_set_query_colors(pfx_ctx* ctx, const void* rgb, size_t n, size_t stride, int mem);
int pfx_shot1344(pfx_ctx* ctx, double radius, const float* lrf_in, float* out, size_t stride, int mem);

/* ------------------------------------------------------------------ Unique Shape Context (next row)
 * pfx_usc1980 <- UniqueShapeContext<PointXYZRGB, ShapeContext1980>::compute (evaluation.cpp:344-371, which sets
 * setMinimalRadius(r / 10), setPointDensityRadius(r / 5) and leaves PCL's local radius 2.5 for the frames).
 * Rows = pcl::ShapeContext1980: descriptor[1980] (12 azimuth x 11 elevation x 15 log-spaced radius bins) + rf[9],
 * 7956 B.  lrf_in (optional): frames of the queries; otherwise SHOT frames at local_radius.  Needs no normals. */
int pfx_usc1980(pfx_ctx* ctx, double search_radius, double min_radius, double density_radius, double local_radius,
                const float* lrf_in, float* out, size_t stride, int mem);

/* ------------------------------------------------------------------ 3D Shape Context (next row)
 * pfx_sc3d1980 <- ShapeContext3DEstimation<PointXYZRGB, Normal, ShapeContext1980>::compute (evaluation.cpp:319-345,
 * the first entry of the reference's descriptor list, which sets setMinimalRadius(r / 10) and
 * setPointDensityRadius(r / 5)).  Same rows, bins and weights as pfx_usc1980; the frame of a query is the NORMAL
 * of its nearest surface point plus a RANDOM tangent direction.  Upstream seeds a boost::mt19937 from the wall
 * clock; here the three uniform [0, 1) draws of query i are the top 24 bits of SplitMix64(seed + golden (3 i + t +
 * 1)), so a run is reproducible and a CPU restatement can follow it.  rf[9] of every row is zero, as upstream
 * ("3DSC does not define a repeatable local RF"); frames_out (optional, nq x 9 floats) receives the frames that
 * were used.  Needs surface normals. */
int pfx_sc3d1980(pfx_ctx* ctx, double search_radius, double min_radius, double density_radius, uint64_t seed, float* out,
                 size_t stride, float* frames_out, int mem);

/* ------------------------------------------------------------------ spin images (next row)
 * pfx_spin_image153 <- SpinImageEstimation<PointXYZRGB, Normal, Histogram<153>>::compute with its defaults
 * (evaluation.cpp:515-554): rows of 153 floats (9 alpha rows x 17 beta columns, pcl::Histogram<153>, 612 B).
 * query_normals: one normal per QUERY (setInputNormals of the input cloud; records with nx, ny, nz first) - the
 * rotation axis of its spin image; the surface needs no normals. */
int pfx_spin_image153(pfx_ctx* ctx, double radius, const void* query_normals, size_t n_normals, size_t stride_normals,
                      float* out, size_t stride, int mem);

/* ------------------------------------------------------------------ RANSAC correspondence rejection (next row)
 * pfx_ransac_reject <- Features<T>::filterCorrespondences (features.h:282-297):
 * CorrespondenceRejectorSampleConsensus with setInlierThreshold(0.015), setMaximumIterations(1000).
 * src / tgt: the keypoint clouds the correspondences index (records with x, y, z first; strides in bytes).
 * out: the surviving correspondences in input order (cap >= n_corr); transform16: row-major 4x4 of the winning
 * 3-point hypothesis (getBestTransformation).  n_corr < 3 keeps everything with the identity, like PCL.
 * Random samples follow the documented SplitMix64 contract with `seed` (PCL's fixed-seed mt19937 shuffle cannot be
 * pinned); iterations_out / best_hypothesis_out (optional) report what PCL's sequential loop would have run. */
int pfx_ransac_reject(pfx_ctx* ctx, const void* src, size_t n_src, size_t stride_src, const void* tgt, size_t n_tgt,
                      size_t stride_tgt, const pfx_correspondence* corr, size_t n_corr, double inlier_threshold,
                      int max_iterations, uint64_t seed, pfx_correspondence* out, size_t cap, size_t* n_out,
                      float* transform16, int* iterations_out, int* best_hypothesis_out, int mem);

/* ------------------------------------------------------------------ ICP (next row)
 * pfx_icp_align <- Evaluation::icpAlign (evaluation.cpp:863-885): pcl::IterativeClosestPoint<PointXYZRGB,
 * PointXYZRGB> with setMaxCorrespondenceDistance(0.07), setTransformationEpsilon(1e-6),
 * setEuclideanFitnessEpsilon(1e-4), setMaximumIterations(100); setInputTarget = the context's cloud
 * (pfx_set_cloud), setInputSource = src (records with x, y, z first; stride in bytes).
 * guess16 (optional, always HOST memory): row-major 4x4 initial transform (align(output, guess)).  result: getFinalTransformation
 * (row-major), getFitnessScore(), hasConverged(), the iteration count and the convergence state
 * (1 ITERATIONS, 2 TRANSFORM, 3 ABS_MSE, 4 REL_MSE, 5 NO_CORRESPONDENCES; 0 = not converged).
 * aligned (optional): the source cloud moved by the final transform, n_src rows of x, y, z at stride_aligned. */
typedef struct {
  double max_correspondence_distance;
  int max_iterations;
  double transformation_epsilon;
  double euclidean_fitness_epsilon;
} pfx_icp_params;
typedef struct {
  float transform[16];
  double fitness;
  int converged, iterations, state, correspondences;
} pfx_icp_result;
int pfx_icp_align(pfx_ctx* ctx, const void* src, size_t n_src, size_t stride_src, const pfx_icp_params* params,
                  const float* guess16, pfx_icp_result* result, void* aligned, size_t stride_aligned, int mem);

/* ------------------------------------------------------------------ range image, NARF keypoints, Narf36
 * pfx_range_image_planar <- RangeImagePlanar::createFromPointCloudWithFixedSize (keypoints.h:204-216,
 * tools.h:65-76); pfx_range_image_spherical <- RangeImage::createFromPointCloud (config C3).  Both project
 * the current surface through the sensor pose set with pfx_range_image_set_pose (default: the sensor at the origin
 * of the cloud frame, what the reference passes for the bundled clouds), CAMERA_FRAME, noise_level 0.  The image (height x width pixels of
 * x, y, z, range = pcl::PointWithRange without padding; unobserved: NaN xyz, range -inf) stays in the
 * context; desc_out receives its geometry (the spherical image is cropped to the observed box + border). */
typedef struct {
  int32_t width, height, planar;
  float cx, cy, fx, fy; /* planar */
  float ang_res;        /* spherical: radians per pixel, both axes */
  int32_t off_x, off_y; /* spherical: offset of the image inside the full 360 x 180 degree grid */
} pfx_range_image_desc;
int pfx_range_image_planar(pfx_ctx* ctx, int width, int height, float cx, float cy, float fx, float fy,
                           float min_range, pfx_range_image_desc* desc_out);
int pfx_range_image_spherical(pfx_ctx* ctx, float ang_res, float max_angle_width, float max_angle_height,
                              float min_range, int border, pfx_range_image_desc* desc_out);
/* Sensor pose (keypoints.h:207-210 builds it as translation(sensor_origin_) * rotation(sensor_orientation_)):
 * pose16 = row-major 4x4, world <- sensor, NULL = identity (what the bundled clouds carry).  It applies to the
 * images built or set after the call: points are projected with its inverse (PCL's to_range_image_system), the
 * image handed out by pfx_range_image_get, the keypoint positions and the Narf36 poses are in world coordinates,
 * Narf36's upright frame uses the world's y axis.  Internally the image lives in the sensor frame, so the border
 * and interest stages are those of the identity pose. */
int pfx_range_image_set_pose(pfx_ctx* ctx, const float* pose16);
/* use a caller-supplied image (height*width*4 floats) / read the current one back */
int pfx_range_image_set(pfx_ctx* ctx, const pfx_range_image_desc* desc, const float* img, int mem);
int pfx_range_image_get(pfx_ctx* ctx, pfx_range_image_desc* desc_out, float* img, int mem);
/* stage outputs of RangeImageBorderExtractor, each optional: traits (h*w int32 bit sets: 1 obstacle border,
 * 2 shadow border, 4 veil point, then per-direction bits), border scores (4*h*w: left, right, top,
 * bottom), surface-change score (h*w) and direction (h*w*3) */
int pfx_narf_borders(pfx_ctx* ctx, int32_t* traits, float* border_scores, float* change_score,
                     float* change_dir, int mem);
/* pfx_narf_keypoints <- NarfKeypoint::compute (keypoints.h:218-224) with PCL's default parameters except
 * support_size.  kp_px: range-image pixel indices y * width + x, ascending; kp_xyz (optional, 3 floats
 * each) the pixels' 3-D points; interest_image (optional, h*w). */
int pfx_narf_keypoints(pfx_ctx* ctx, float support_size, int32_t* kp_px, float* kp_xyz, float* kp_interest,
                       size_t cap, size_t* n_kp, float* interest_image, int mem);
/* pfx_narf36 <- NarfDescriptor::compute (evaluation.cpp:629-637): one row per (keypoint, dominant
 * rotation) when rotation_invariant, rows in keypoint order; row = pcl::Narf36 {x, y, z, roll, pitch, yaw,
 * descriptor[36]} (168 bytes) at `stride` bytes.  Keypoints whose pixel is unobserved or whose
 * neighbourhood cannot support a normal produce no row. */
int pfx_narf36(pfx_ctx* ctx, const int32_t* kp_px, size_t n_kp, float support_size, int rotation_invariant,
               void* out, size_t stride, size_t cap, size_t* n_out, int mem);

/* ------------------------------------------------------------------ ingest
 * pfx_voxel_grid <- pcl::VoxelGrid centroid filter (config C1 ingest; not in the reference code).
 * Operates on the current surface; out: cap x 3 packed floats, ascending voxel id. */
int pfx_voxel_grid(pfx_ctx* ctx, float leaf, float* out_xyz, size_t cap, size_t* n_out, int mem);

#ifdef __cplusplus
}
#endif
#endif
