/*
 * pcl_oracle.h — C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.
 *
 * PARITY UNPINNED: the arithmetic of the reference's hot path lives in PCL 1.7.x
 * (+ FLANN 1.8, Eigen 3.2), which is neither vendored under /root/reference nor
 * installed here, and the reference ships no tests / golden vectors.  This is a
 * restatement of the published PCL algorithms, anchored on
 * the reference's call sites cited per function, closed-form known answers and a
 * brute-force neighbour search.
 *
 * Conventions: points are N x 3 contiguous float32; normals are N x 4 float32
 * (nx, ny, nz, curvature); indices int32; all functions return 0 on success.
 */
#ifndef PCL_ORACLE_H
#define PCL_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- neighbour search: restates KdTreeFLANN<PointT, L2_Simple<float>> results
 * (reference call sites features.h:192-193, tools.h:29-30, keypoints.h:186,371,408).
 * d2 = ((dx*dx + dy*dy) + dz*dz) in float, no FMA; radius test d2 < (float)(r*r);
 * order ascending (d2, index). */
int orc_radius_count(const float* surf, int n, const float* q, int nq, double radius, int* counts);
/* offsets: nq+1 (exclusive prefix of counts, int64); idx/d2 sized offsets[nq] */
int orc_radius_search(const float* surf, int n, const float* q, int nq, double radius,
                      const int64_t* offsets, int* idx, float* d2);
int orc_radius_search_brute(const float* surf, int n, const float* q, int nq, double radius,
                            const int64_t* offsets, int* idx, float* d2);
/* kNN: idx/d2 are nq x k; rows padded with -1 / +inf when n < k */
int orc_knn(const float* surf, int n, const float* q, int nq, int k, int* idx, float* d2);
int orc_knn_brute(const float* surf, int n, const float* q, int nq, int k, int* idx, float* d2);

/* ---- normals: NormalEstimationOMP (tools.h:26-31, features.h:187).
 * mode 0: centred double (the parity gate); mode 1: PCL-1.7 float single-pass +
 * closed-form eigen33 (reporting only).  Exactly one of radius / k non-zero.
 * eig_gap (optional, nq): (l1-l0)/l2 of the centred double covariance, -1 if n<3. */
int orc_normals(const float* surf, int n, const float* q, int nq, double radius, int k,
                const float vp[3], int mode, float* normals4, int* n_nbrs, float* eig_gap);

/* ---- cloud resolution (keypoints.h:401-428) */
int orc_cloud_resolution(const float* pts, int n, double* res);

/* ---- ISS keypoints (keypoints.h:184-194 -> ISSKeypoint3D).
 * orc_iss_saliency: third_eigen_value[] (n doubles; 0 where the ratio tests fail).
 * orc_iss_nms: keypoints (ascending index) from a given saliency array.
 * orc_iss = both. */
int orc_iss_saliency(const float* pts, int n, double salient_radius, int min_neighbors,
                     double gamma21, double gamma32, double* saliency);
int orc_iss_nms(const float* pts, int n, const double* saliency, double nonmax_radius,
                int min_neighbors, int* kp_idx, int* n_kp);
int orc_iss(const float* pts, int n, double salient_radius, double nonmax_radius, int min_neighbors,
            double gamma21, double gamma32, int* kp_idx, int* n_kp, double* saliency);

/* ---- Harris3D (keypoints.h:154-162 -> HarrisKeypoint3D, + snap keypoints.h:360-395).
 * normals4: n x 4 (NaN rows allowed).  Stages are exposed separately so that index parity can be
 * checked stage by stage on identical inputs. */
int orc_harris_response(const float* pts, const float* normals4, int n, double radius,
                        float* response);
int orc_harris_nms(const float* pts, const float* response, int n, double radius, float threshold,
                   int* kp_idx, int* n_kp);
/* corners: nc x 3 in/out (refineCorners, <= 10 iterations) */
int orc_harris_refine(const float* pts, const float* normals4, int n, double radius,
                      float* corners, int nc);
/* 1-NN snap: snapped_idx[i] = nearest cloud index if d2 < max_d2 else -1 */
int orc_snap_to_cloud(const float* pts, int n, const float* q, int nq, float max_d2,
                      int* snapped_idx);

/* ---- FPFH (evaluation.cpp:597-602 -> FPFHEstimation).  out: nq x 33. */
int orc_fpfh(const float* surf, const float* normals4, int n, const float* q, int nq,
             double radius, int k, float* out33);
/* SPFH of surface points listed in pidx (np of them): out np x 33 */
int orc_spfh(const float* surf, const float* normals4, int n, const int* pidx, int np,
             double radius, int k, float* out33);

/* ---- PFH125 (evaluation.cpp:676-695 -> PFHEstimation) and PrincipalCurvatures (evaluation.cpp:696-715).
 * out125: nq x 125 (NaN rows for queries without neighbours); counts125 (optional): the integer votes.
 * out5: nq x 5 (principal direction, pc1, pc2); the tangent plane of query i is that of normals4[i] (upstream's
 * indexing); gap (optional): (l2 - l1) / l2, -1 for NaN rows. */
int orc_pfh125(const float* surf, const float* normals4, int n, const float* q, int nq, double radius, int k,
               float* out125, int* counts125);
int orc_principal_curvatures(const float* surf, const float* normals4, int n, const float* q, int nq,
                             double radius, int k, float* out5, float* gap);

/* ---- MomentInvariants (evaluation.cpp:555-574): out nq x 3 (j1, j2, j3) */
int orc_moment_invariants(const float* surf, int n, const float* q, int nq, double radius, int k, float* out3);

/* ---- SHOT (evaluation.cpp:770-775 -> SHOTEstimationOMP + SHOTLocalReferenceFrameEstimation).
 * rf: nq x 9 (x_axis, y_axis, z_axis). lrf_in (optional): use these frames instead.
 * lrf_gap (optional, nq x 2): relative eigen gaps (l2-l1)/l2 and (l1-l0)/l2 of the LRF matrix. */
int orc_shot_lrf(const float* surf, int n, const float* q, int nq, double radius, float* rf9,
                 float* lrf_gap);
int orc_shot352(const float* surf, const float* normals4, int n, const float* q, int nq,
                double radius, const float* lrf_in, float* out352, float* rf9);

/* ---- SHOT1344 = shape + colour (evaluation.cpp:786-805 -> SHOTColorEstimation).  rgb / qrgb: packed 0x00RRGGBB
 * of the surface points / queries; out: nq x 1344; lab_out (optional): n x 3 (L/100, a/120, b/120). */
int orc_shot1344(const float* surf, const uint32_t* rgb, const float* normals4, int n, const float* q,
                 const uint32_t* qrgb, int nq, double radius, const float* lrf_in, float* out1344, float* rf9,
                 float* lab_out);

/* ---- Unique Shape Context (evaluation.cpp:344-371 -> UniqueShapeContext<PointXYZRGB, ShapeContext1980>):
 * out nq x 1980, rf9 nq x 9 (SHOT frames at local_radius unless lrf_in is given); density_out (optional, n). */
int orc_usc1980(const float* surf, int n, const float* q, int nq, double search_radius, double min_radius,
                double density_radius, double local_radius, const float* lrf_in, float* out1980, float* rf9,
                int* density_out);

/* ---- Harris 6D (keypoints.h:166-179 -> HarrisKeypoint6D): response = 4th smallest eigenvalue of the 6x6 covariance
 * of (normal, normalised intensity gradient); rgb packed 0x00RRGGBB; gradients_out (n x 3) / intensity_out (n)
 * optional.  NMS / refinement / snap are Harris3D's (orc_harris_nms, orc_harris_refine, orc_snap_to_cloud). */
int orc_harris6d_response(const float* pts, const uint32_t* rgb, const float* normals4, int n, double radius,
                          float* response, float* gradients_out, float* intensity_out);

/* 3DSC (evaluation.cpp:319-345): frames = nearest neighbour's normal + a seeded random tangent direction (SplitMix64
 * contract, usc.cpp); descriptor = USC's bins in that frame.  frames_out optional. */
int orc_sc3d_frames(const float* surf, const float* normals4, int n, const float* q, int nq, double search_radius,
                    unsigned long long seed, float* rf9);
int orc_sc3d1980(const float* surf, const float* normals4, int n, const float* q, int nq, double search_radius,
                 double min_radius, double density_radius, unsigned long long seed, float* out1980, float* frames_out);

/* ---- spin images (evaluation.cpp:515-554 -> SpinImageEstimation<PointXYZRGB, Normal, Histogram<153>>, defaults):
 * qnormals4: the normals of the QUERIES (nq x 4); out nq x 153 (9 alpha rows x 17 beta columns). */
int orc_spin_image153(const float* surf, int n, const float* q, const float* qnormals4, int nq, double radius,
                      float* out153);

/* ---- matching (features.h:224-273): exact L2 1-NN with sequential float sum.
 * nn_idx: na (argmin over b; -1 for NaN query rows / empty b), nn_d2: na */
int orc_match_nn(const float* a, int na, const float* b, int nb, int dim, int* nn_idx, float* nn_d2);
/* reciprocal correspondences in ascending query order; out arrays sized na */
int orc_match_reciprocal(const float* a, int na, const float* b, int nb, int dim,
                         int* q_idx, int* m_idx, float* dist, int* n_out);

/* ---- VoxelGrid centroid filter (config C1 ingest). out_xyz cap x 3 */
int orc_voxel_grid(const float* pts, int n, float leaf, float* out_xyz, int cap, int* n_out);

/* ---- Range image (keypoints.h:204-216, tools.h:65-76). img: h*w*4 floats (x,y,z,range) */
int orc_range_image_planar(const float* pts, int n, int width, int height, float cx, float cy,
                           float fx, float fy, float min_range, float* img);
/* spherical image with crop: out dims and offsets returned; img sized max_w*max_h*4 */
int orc_range_image_spherical(const float* pts, int n, float ang_res, float max_angle_w,
                              float max_angle_h, float min_range, int border, float* img,
                              int cap_px, int* out_w, int* out_h, int* off_x, int* off_y);

/* ---- NARF keypoints + Narf36 (keypoints.h:218-224, evaluation.cpp:629-637).
 * planar=1: planar projection parameters (cx,cy,fx,fy); planar=0: spherical with
 * ang_res and image offsets. */
typedef struct {
  int width, height, planar;
  float cx, cy, fx, fy;      /* planar */
  float ang_res; int off_x, off_y; /* spherical */
} orc_ri_desc;
/* stage outputs of RangeImageBorderExtractor (each optional): traits h*w (bit set, see narf.cpp),
 * border scores 4*h*w (left, right, top, bottom), surface-change score h*w and direction h*w*3 */
int orc_narf_borders(const float* img, const orc_ri_desc* d, int* traits, float* border_scores,
                     float* sc_score, float* sc_dir);
int orc_narf_keypoints(const float* img, const orc_ri_desc* d, float support_size,
                       int* kp_px, float* kp_interest, int cap, int* n_kp,
                       float* interest_image /* optional h*w */);
int orc_narf36(const float* img, const orc_ri_desc* d, const int* kp_px, int n_kp,
               float support_size, int rotation_invariant, float* out /* cap x 42 */,
               int cap, int* n_out);

/* ---- RANSAC correspondence rejection + rigid transform (features.h:282-297 ->
 * CorrespondenceRejectorSampleConsensus, threshold 0.015, 1000 iterations).  keep: n_corr flags; T16: row-major
 * 4x4 of the winning 3-point hypothesis; the sampling contract is defined in ransac.cpp. */
int orc_ransac_reject(const float* src, int ns, const float* tgt, int nt, const int* corr_q, const int* corr_m,
                      int n_corr, double threshold, int max_iterations, uint64_t seed, int* keep, float* T16,
                      int* n_inliers, int* iterations, int* best_hypothesis);

/* ---- ICP (evaluation.cpp:863-885 -> IterativeClosestPoint, max correspondence distance 0.07, epsilons 1e-6 /
 * 1e-4, 100 iterations).  guess16 optional (row-major 4x4).  state: 1 ITERATIONS, 2 TRANSFORM, 3 ABS_MSE,
 * 4 REL_MSE, 5 NO_CORRESPONDENCES.  fitness = getFitnessScore() (mean squared NN distance of all source points). */
int orc_icp(const float* src, int ns, const float* tgt, int nt, double max_corr_dist, int max_iterations,
            double transformation_epsilon, double euclidean_fitness_epsilon, const float* guess16,
            float* T16, double* fitness, int* converged, int* iterations, int* state);

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
