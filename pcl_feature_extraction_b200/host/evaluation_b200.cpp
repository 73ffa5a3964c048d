// evaluation_b200.cpp — ROS-free driver with the sequencing of the reference's evaluate() loop
// (src/evaluation.cpp:272-852): for every keypoint detector x descriptor of the B200 path, detect
// keypoints on source and target, describe them, match reciprocally, and print one CSV row with the
// reference's column meaning (evaluation.cpp:190-206), preceded by the direct ICP of the two clouds
// (evaluation.cpp:246-266) and, per detector, the ICP of the keypoint clouds (evaluation.cpp:289-293).
//
//   evaluation_b200 <source.pcd> <target.pcd> [feat_radius=0.08] [normal_radius=0.05] [dump_dir] [strict=1]
//
// With dump_dir, raw results are written (kp indices are implied by the keypoint clouds) so that
// tests/test_host_shim.py can compare this C++ path with the Python-bound C ABI and the oracle.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "feature_pipeline.hpp"

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <typename T>
static void dump(const std::string& dir, const std::string& name, const std::vector<T>& v) {
  if (dir.empty()) return;
  std::ofstream f(dir + "/" + name, std::ios::binary);
  f.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
}

// Evaluation::icpAlign (evaluation.cpp:863-885), same settings
static void icpAlign(const PointCloudRGB::Ptr& src, const PointCloudRGB::Ptr& tgt, pcl::Matrix4f& output, double& score,
                     bool& convergence) {
  PointCloudRGB::Ptr aligned(new PointCloudRGB);
  pcl::IterativeClosestPoint<PointRGB, PointRGB> icp;
  icp.setMaxCorrespondenceDistance(0.07);
  icp.setRANSACOutlierRejectionThreshold(0.005);
  icp.setTransformationEpsilon(0.000001);
  icp.setEuclideanFitnessEpsilon(0.0001);
  icp.setMaximumIterations(100);
  icp.setInputSource(src);
  icp.setInputTarget(tgt);
  icp.align(*aligned);
  output = icp.getFinalTransformation();
  score = icp.getFitnessScore();
  convergence = icp.hasConverged();
}

static void print_icp(const char* what, const pcl::Matrix4f& tf, double runtime, double score, bool convergence,
                      const std::string& dump_dir, const std::string& dump_name) {
  std::printf("# %s: (%g | %g | %g), runtime %.6f, score %.9g, convergence %d\n", what, tf(0, 3), tf(1, 3), tf(2, 3), runtime,
              score, convergence ? 1 : 0);
  std::vector<float> v(tf.m, tf.m + 16);
  v.push_back((float)score);
  v.push_back(convergence ? 1.f : 0.f);
  dump(dump_dir, dump_name, v);
}

template <typename FeatureT>
static void run_descriptor(const std::string& kp_type, const std::string& desc_type,
                           typename pcl::Feature<PointRGB, FeatureT>::Ptr extractor, const PointCloudRGB::Ptr& source,
                           const PointCloudRGB::Ptr& target, const PointCloudRGB::Ptr& skp, const PointCloudRGB::Ptr& tkp,
                           double feat_r, double normal_r, double kp_runtime, const std::string& dump_dir) {
  typename pcl::PointCloud<FeatureT>::Ptr sf(new pcl::PointCloud<FeatureT>), tf(new pcl::PointCloud<FeatureT>);
  double t0 = now_s();
  Features<FeatureT> feat(extractor, feat_r, normal_r);
  feat.compute(source, skp, sf);
  feat.compute(target, tkp, tf);
  double desc_runtime = now_s() - t0;
  t0 = now_s();
  pcl::CorrespondencesPtr corr(new pcl::Correspondences), filtered(new pcl::Correspondences);
  feat.findCorrespondences(sf, tf, corr);
  float ransac_tf[16];
  feat.filterCorrespondences(skp, tkp, corr, filtered, ransac_tf);  // evaluation.cpp:610 (the "Ransac rejector" column)
  double corr_runtime = now_s() - t0;
  std::printf("%s, %s, %zu, %zu, %zu, %zu, %zu, %zu, %zu, %zu, %.6f, %.6f, %.6f\n", kp_type.c_str(), desc_type.c_str(),
              source->size(), target->size(), skp->size(), tkp->size(), sf->size(), tf->size(), corr->size(),
              filtered->size(), kp_runtime, desc_runtime, corr_runtime);
  dump(dump_dir, kp_type + "_" + desc_type + "_filtered.bin", *filtered);
  dump(dump_dir, kp_type + "_" + desc_type + "_tf.bin", std::vector<float>(ransac_tf, ransac_tf + 16));
  dump(dump_dir, kp_type + "_" + desc_type + "_src.bin", sf->points);
  dump(dump_dir, kp_type + "_" + desc_type + "_tgt.bin", tf->points);
  dump(dump_dir, kp_type + "_" + desc_type + "_corr.bin", *corr);
}

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s source.pcd target.pcd [feat_radius] [normal_radius] [dump_dir]\n", argv[0]);
    return 2;
  }
  const double feat_r = argc > 3 ? std::atof(argv[3]) : 0.08;    // evaluation.cpp:167
  const double normal_r = argc > 4 ? std::atof(argv[4]) : 0.05;  // evaluation.cpp:168
  const std::string dump_dir = argc > 5 ? argv[5] : "";
  PointCloudRGB::Ptr source(new PointCloudRGB), target(new PointCloudRGB);
  if (loadPCDFile(argv[1], *source) != 0 || loadPCDFile(argv[2], *target) != 0) {
    std::fprintf(stderr, "cannot read the input clouds\n");
    return 1;
  }
  if (!pcl::b200::ctx()) return 3;
  // the reference's pipeline takes index decisions on floats: run it in reference-order arithmetic unless told not to
  pcl::b200::set_parity_strict(!(argc > 6 && std::atoi(argv[6]) == 0));
  std::printf("Keypoint name, Descriptor name, Source cloud size, Target cloud size, Source keypoints size, "
              "Target keypoints size, Source features size, Target features size, Correspondences, "
              "Filtered correspondences, Keypoints runtime, Features runtime, Correspondences runtime\n");
  {  // Step 1: direct ICP (evaluation.cpp:246-266)
    pcl::Matrix4f icp_tf;
    double score = 0;
    bool convergence = false;
    double t0 = now_s();
    icpAlign(source, target, icp_tf, score, convergence);
    print_icp("direct ICP", icp_tf, now_s() - t0, score, convergence, dump_dir, "direct_icp.bin");
  }
  const std::string keypoints_list[] = {KP_HARRIS_3D, KP_HARRIS_6D, KP_ISS};  // the active list of evaluation.cpp:63-65
  for (const std::string& kp_type : keypoints_list) {
    PointCloudRGB::Ptr skp(new PointCloudRGB), tkp(new PointCloudRGB);
    double t0 = now_s();
    Keypoints kp(kp_type, normal_r);
    kp.compute(source, skp);
    kp.compute(target, tkp);
    double kp_runtime = now_s() - t0;
    if (skp->points.empty() || tkp->points.empty()) continue;  // evaluation.cpp:286-287
    {  // ICP of the keypoint clouds (evaluation.cpp:289-293)
      pcl::Matrix4f icp_kp_tf;
      double score = 0;
      bool convergence = false;
      double t1 = now_s();
      icpAlign(skp, tkp, icp_kp_tf, score, convergence);
      print_icp((kp_type + " keypoints ICP").c_str(), icp_kp_tf, now_s() - t1, score, convergence, dump_dir,
                kp_type + "_icp_kp.bin");
    }
    dump(dump_dir, kp_type + "_src_kp.bin", skp->points);
    dump(dump_dir, kp_type + "_tgt_kp.bin", tkp->points);
    {  // evaluation.cpp:319-345, the first entry of the descriptor list
      pcl::ShapeContext3DEstimation<PointRGB, pcl::Normal, pcl::ShapeContext1980>::Ptr sc(
          new pcl::ShapeContext3DEstimation<PointRGB, pcl::Normal, pcl::ShapeContext1980>);
      sc->setMinimalRadius(feat_r / 10.0);
      sc->setPointDensityRadius(feat_r / 5.0);
      pcl::Feature<PointRGB, pcl::ShapeContext1980>::Ptr ex(sc);
      run_descriptor<pcl::ShapeContext1980>(kp_type, DESC_SHAPE_CONTEXT, ex, source, target, skp, tkp, feat_r, normal_r, kp_runtime,
                                            dump_dir);
    }
    {
      pcl::Feature<PointRGB, pcl::FPFHSignature33>::Ptr ex(new pcl::FPFHEstimation<PointRGB, pcl::Normal, pcl::FPFHSignature33>);
      run_descriptor<pcl::FPFHSignature33>(kp_type, DESC_FPFH, ex, source, target, skp, tkp, feat_r, normal_r, kp_runtime, dump_dir);
    }
    {
      pcl::Feature<PointRGB, pcl::SHOT352>::Ptr ex(new pcl::SHOTEstimationOMP<PointRGB, pcl::Normal, pcl::SHOT352>);
      run_descriptor<pcl::SHOT352>(kp_type, DESC_SHOT, ex, source, target, skp, tkp, feat_r, normal_r, kp_runtime, dump_dir);
    }
    {  // evaluation.cpp:786-805
      pcl::Feature<PointRGB, pcl::SHOT1344>::Ptr ex(new pcl::SHOTColorEstimationOMP<PointRGB, pcl::Normal, pcl::SHOT1344>);
      run_descriptor<pcl::SHOT1344>(kp_type, DESC_SHOT_COLOR, ex, source, target, skp, tkp, feat_r, normal_r, kp_runtime, dump_dir);
    }
    {  // evaluation.cpp:515-554: the normals are estimated on the keypoint clouds, the surface is the full cloud
      typedef pcl::Histogram<153> Spin;
      double t1 = now_s();
      pcl::PointCloud<Spin>::Ptr sf(new pcl::PointCloud<Spin>), tf(new pcl::PointCloud<Spin>);
      pcl::PointCloud<pcl::Normal>::Ptr sn(new pcl::PointCloud<pcl::Normal>), tn(new pcl::PointCloud<pcl::Normal>);
      Tools::estimateNormals(skp, sn, normal_r);
      Tools::estimateNormals(tkp, tn, normal_r);
      pcl::SpinImageEstimation<PointRGB, pcl::Normal, Spin> ex;
      pcl::search::KdTree<PointRGB>::Ptr kdtree(new pcl::search::KdTree<PointRGB>);
      ex.setInputNormals(sn);
      ex.setSearchSurface(source);
      ex.setInputCloud(skp);
      ex.setSearchMethod(kdtree);
      ex.setRadiusSearch(feat_r);
      ex.compute(*sf);
      ex.setInputNormals(tn);
      ex.setSearchSurface(target);
      ex.setInputCloud(tkp);
      ex.setSearchMethod(kdtree);
      ex.setRadiusSearch(feat_r);
      ex.compute(*tf);
      double desc_runtime = now_s() - t1;
      t1 = now_s();
      Features<Spin> feat;
      pcl::CorrespondencesPtr corr(new pcl::Correspondences), filtered(new pcl::Correspondences);
      feat.findCorrespondences(sf, tf, corr);
      float ransac_tf[16];
      feat.filterCorrespondences(skp, tkp, corr, filtered, ransac_tf);
      std::printf("%s, %s, %zu, %zu, %zu, %zu, %zu, %zu, %zu, %zu, %.6f, %.6f, %.6f\n", kp_type.c_str(), DESC_SPIN_IMAGE.c_str(),
                  source->size(), target->size(), skp->size(), tkp->size(), sf->size(), tf->size(), corr->size(),
                  filtered->size(), kp_runtime, desc_runtime, now_s() - t1);
      dump(dump_dir, kp_type + "_SpinImage_src.bin", sf->points);
      dump(dump_dir, kp_type + "_SpinImage_src_normals.bin", sn->points);
    }
    {  // evaluation.cpp:344-371
      pcl::UniqueShapeContext<PointRGB, pcl::ShapeContext1980>::Ptr usc(new pcl::UniqueShapeContext<PointRGB, pcl::ShapeContext1980>);
      usc->setMinimalRadius(feat_r / 10.0);
      usc->setPointDensityRadius(feat_r / 5.0);
      pcl::Feature<PointRGB, pcl::ShapeContext1980>::Ptr ex(usc);
      run_descriptor<pcl::ShapeContext1980>(kp_type, DESC_USC, ex, source, target, skp, tkp, feat_r, normal_r, kp_runtime, dump_dir);
    }
    {  // evaluation.cpp:555-574
      pcl::Feature<PointRGB, pcl::MomentInvariants>::Ptr ex(new pcl::MomentInvariantsEstimation<PointRGB, pcl::MomentInvariants>);
      run_descriptor<pcl::MomentInvariants>(kp_type, DESC_MOMENT_INV, ex, source, target, skp, tkp, feat_r, normal_r, kp_runtime,
                                            dump_dir);
    }
    {  // evaluation.cpp:676-695
      pcl::Feature<PointRGB, pcl::PFHSignature125>::Ptr ex(new pcl::PFHEstimation<PointRGB, pcl::Normal, pcl::PFHSignature125>);
      run_descriptor<pcl::PFHSignature125>(kp_type, DESC_PFH, ex, source, target, skp, tkp, feat_r, normal_r, kp_runtime, dump_dir);
    }
    {  // evaluation.cpp:696-715
      pcl::Feature<PointRGB, pcl::PrincipalCurvatures>::Ptr ex(
          new pcl::PrincipalCurvaturesEstimation<PointRGB, pcl::Normal, pcl::PrincipalCurvatures>);
      run_descriptor<pcl::PrincipalCurvatures>(kp_type, DESC_PPAL_CURV, ex, source, target, skp, tkp, feat_r, normal_r, kp_runtime,
                                               dump_dir);
    }
  }
  {
    // NARF keypoints + Narf36 (keypoints.h:199-231, evaluation.cpp:613-648).  Each cloud is described on ITS
    // OWN range image (the reference builds the target descriptors on the source image, evaluation.cpp:630).
    double t0 = now_s();
    PointCloudRGB::Ptr skp(new PointCloudRGB), tkp(new PointCloudRGB);
    Keypoints skd(KP_NARF, normal_r), tkd(KP_NARF, normal_r);
    skd.compute(source, skp);
    tkd.compute(target, tkp);
    double kp_runtime = now_s() - t0;
    if (!skp->points.empty() && !tkp->points.empty()) {
      t0 = now_s();
      pcl::PointCloud<pcl::Narf36>::Ptr sf(new pcl::PointCloud<pcl::Narf36>), tf(new pcl::PointCloud<pcl::Narf36>);
      pcl::RangeImagePlanar sri, tri;
      Tools::convertToRangeImage(source, sri);
      Tools::convertToRangeImage(target, tri);
      pcl::NarfDescriptor sd(&sri, &skd.getNarfPixelIndices()), td(&tri, &tkd.getNarfPixelIndices());
      sd.getParameters().support_size = 0.2f;
      sd.getParameters().rotation_invariant = true;
      td.getParameters().support_size = 0.2f;
      td.getParameters().rotation_invariant = true;
      sd.compute(*sf);
      td.compute(*tf);
      double desc_runtime = now_s() - t0;
      t0 = now_s();
      // only descriptor[36] is the point representation (not the pose): match on packed copies
      std::vector<float> a(sf->size() * 36), b(tf->size() * 36);
      for (size_t i = 0; i < sf->size(); ++i) std::memcpy(&a[36 * i], sf->points[i].descriptor, 144);
      for (size_t i = 0; i < tf->size(); ++i) std::memcpy(&b[36 * i], tf->points[i].descriptor, 144);
      pcl::Correspondences corr(sf->size());
      size_t nc = 0;
      if (!sf->points.empty() && !tf->points.empty())
        pcl::b200::ok(pfx_match(pcl::b200::ctx(), a.data(), sf->size(), 144, b.data(), tf->size(), 144, 36, 1, -1.f,
                                reinterpret_cast<pfx_correspondence*>(corr.data()), corr.size(), &nc, PFX_HOST), "match");
      corr.resize(nc);
      double corr_runtime = now_s() - t0;
      std::printf("%s, %s, %zu, %zu, %zu, %zu, %zu, %zu, %zu, %zu, %.6f, %.6f, %.6f\n", KP_NARF.c_str(), DESC_NARF.c_str(),
                  source->size(), target->size(), skp->size(), tkp->size(), sf->size(), tf->size(), corr.size(), (size_t)0,
                  kp_runtime, desc_runtime, corr_runtime);
      dump(dump_dir, "Narf_src_px.bin", skd.getNarfPixelIndices());
      dump(dump_dir, "Narf_NARF_src.bin", sf->points);
      dump(dump_dir, "Narf_NARF_corr.bin", corr);
    }
  }
  {
    // what the reference's redundancy (the same cloud and its normals re-submitted per descriptor type,
    // features.h:186-193) cost here: uploads and normals passes actually performed vs answered from resident state
    unsigned long long r[6];
    pcl::b200::reuse_totals(r);
    std::printf("# reuse: surface uploads %llu (reused %llu), dense normals passes %llu (reused %llu), normals uploads %llu "
                "(skipped %llu)\n", r[0], r[1], r[2], r[3], r[4], r[5]);
    std::vector<unsigned long long> v(r, r + 6);
    dump(dump_dir, "reuse.bin", v);
  }
  return 0;  // the thread's contexts are released with its pool
}
