// feature_pipeline.hpp — the reference's own wrapper interface for this path (class Tools
// tools.h:14-103, class Keypoints keypoints.h:54-85, template class Features features.h:106-148:
// same names, argument meaning and error behaviour), written against pcl_compat.hpp so that the
// sequencing of src/evaluation.cpp:272-852 runs on the B200 library.  Differences that the GPU
// path makes on purpose are marked "batched:" — a per-point host loop in the reference becomes one
// C-ABI call here.
#pragma once
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "pcl_compat.hpp"

typedef pcl::PointXYZRGB PointRGB;
typedef pcl::PointCloud<pcl::PointXYZ> PointCloudXYZ;
typedef pcl::PointCloud<PointRGB> PointCloudRGB;
typedef pcl::PointCloud<pcl::PointXYZI> PointCloudXYZI;

static const std::string KP_HARRIS_3D = "Harris3D";
static const std::string KP_HARRIS_6D = "Harris6D";
static const std::string KP_ISS = "Iss";
static const std::string KP_NARF = "Narf";
static const std::string DESC_NARF = "NARF";
static const std::string DESC_FPFH = "FPFH";
static const std::string DESC_SHOT = "SHOT";
static const std::string DESC_SHOT_COLOR = "SHOTColor";
static const std::string DESC_SPIN_IMAGE = "SpinImage";
static const std::string DESC_USC = "USC";
static const std::string DESC_SHAPE_CONTEXT = "ShapeContext";
static const std::string DESC_MOMENT_INV = "MomentInvariants";
static const std::string DESC_PFH = "PFH";
static const std::string DESC_PPAL_CURV = "PrincipalCurvatures";

class Tools {
 public:
  // tools.h:22-32
  static void estimateNormals(const PointCloudRGB::Ptr& cloud, pcl::PointCloud<pcl::Normal>::Ptr& normals,
                              double radius_search) {
    pcl::NormalEstimationOMP<PointRGB, pcl::Normal> ne;
    ne.setInputCloud(cloud);
    ne.setRadiusSearch(radius_search);
    pcl::search::KdTree<PointRGB>::Ptr tree(new pcl::search::KdTree<PointRGB>);
    ne.setSearchMethod(tree);
    ne.compute(*normals);
  }

  // tools.h:61-77
  static void convertToRangeImage(const PointCloudRGB::Ptr& cloud, pcl::RangeImagePlanar& range_image) {
    int image_size_x = 640, image_size_y = 480;
    float center_x = (640.0f / 2.0f), center_y = (480.0f / 2.0f);
    float focal_length_x = 525.0f;
    float noise_level = 0.0f, minimum_range = 0.0f;
    // keypoints.h:207-210: translation(sensor_origin_) * rotation(sensor_orientation_) (identity for the bundled clouds)
    const pcl::Affine3f sensor_pose = pcl::poseFromOriginAndOrientation(cloud->sensor_origin_, cloud->sensor_orientation_);
    range_image.createFromPointCloudWithFixedSize(*cloud, image_size_x, image_size_y, center_x, center_y, focal_length_x,
                                                  focal_length_x, sensor_pose, pcl::RangeImage::CAMERA_FRAME,
                                                  noise_level, minimum_range);
  }
};

class Keypoints {
 public:
  explicit Keypoints(const std::string kp_type) : kp_type_(kp_type), normal_radius_search_(0.05) {}
  Keypoints(const std::string kp_type, double normal_radius_search)
      : kp_type_(kp_type), normal_radius_search_(normal_radius_search) {}

  // keypoints.h:102-291 (Harris3D, Harris6D, ISS and NARF branches)
  void compute(const PointCloudRGB::Ptr& cloud, PointCloudRGB::Ptr& cloud_keypoints) {
    if (kp_type_ == KP_HARRIS_3D) {
      pcl::HarrisKeypoint3D<PointRGB, pcl::PointXYZI> harris3d;
      PointCloudXYZI::Ptr keypoints(new PointCloudXYZI);
      harris3d.setNonMaxSupression(true);
      harris3d.setInputCloud(cloud);
      harris3d.setThreshold(1e-6f);
      harris3d.compute(*keypoints);
      // batched: getKeypointsCloud (keypoints.h:360-395) = 1-NN snap with squared gate 1e-4,
      // already done on the device by pfx_harris3d
      cloud_keypoints.reset(new PointCloudRGB);
      for (int s : harris3d.getSnappedIndices())
        if (s >= 0) cloud_keypoints->push_back(cloud->points[s]);
      return;
    }
    if (kp_type_ == KP_HARRIS_6D) {  // keypoints.h:166-179
      pcl::HarrisKeypoint6D<PointRGB, pcl::PointXYZI> harris6d;
      PointCloudXYZI::Ptr keypoints(new PointCloudXYZI);
      harris6d.setNonMaxSupression(true);
      harris6d.setInputCloud(cloud);
      harris6d.setThreshold(1e-6f);
      harris6d.compute(*keypoints);
      cloud_keypoints.reset(new PointCloudRGB);
      for (int s : harris6d.getSnappedIndices())
        if (s >= 0) cloud_keypoints->push_back(cloud->points[s]);
      return;
    }
    if (kp_type_ == KP_ISS) {
      pcl::ISSKeypoint3D<PointRGB, PointRGB> detector;
      detector.setInputCloud(cloud);
      pcl::search::KdTree<PointRGB>::Ptr tree(new pcl::search::KdTree<PointRGB>);
      detector.setSearchMethod(tree);
      double resolution = computeCloudResolution(cloud);
      detector.setSalientRadius(6 * resolution);
      detector.setNonMaxRadius(4 * resolution);
      detector.setMinNeighbors(5);
      detector.setThreshold21(0.975);
      detector.setThreshold32(0.975);
      detector.compute(*cloud_keypoints);
      return;
    }
    if (kp_type_ == KP_NARF) {
      // keypoints.h:199-231.  The reference indexes the UNORGANISED cloud with range-image pixel indices
      // (keypoints.h:228-229, out of bounds for 640 x 480 = 307 200 > cloud size); here a keypoint is the
      // range-image point of its pixel, and the pixel indices are kept for the descriptor stage.
      pcl::RangeImagePlanar range_image;
      Tools::convertToRangeImage(cloud, range_image);
      pcl::PointCloud<int> keypoints;
      pcl::RangeImageBorderExtractor border_extractor;
      pcl::NarfKeypoint detector(&border_extractor);
      detector.setRangeImage(&range_image);
      detector.getParameters().support_size = 0.2f;
      detector.compute(keypoints);
      cloud_keypoints.reset(new PointCloudRGB);
      narf_pixel_indices_.assign(keypoints.points.begin(), keypoints.points.end());
      for (int px : keypoints.points) {
        const pcl::PointWithRange& p = range_image.getPoint(px);
        PointRGB q;
        q.x = p.x; q.y = p.y; q.z = p.z;
        cloud_keypoints->push_back(q);
      }
      return;
    }
    std::fprintf(stderr, "[Keypoints::compute] keypoint type %s is outside the B200 path\n", kp_type_.c_str());
  }

  // keypoints.h:401-428.  batched: one device-side 2-NN pass + reduction instead of N tree queries
  double computeCloudResolution(const PointCloudRGB::Ptr& cloud) {
    pfx_ctx* c = pcl::b200::use(cloud.get());  // the context bound to this cloud
    double res = 0.0;
    if (!c) return res;
    if (!pcl::b200::ok(pfx_set_surface(c, cloud->points.data(), cloud->size(), sizeof(PointRGB), PFX_HOST), "Keypoints")) return 0.0;
    pcl::b200::ok(pfx_cloud_resolution(c, &res), "Keypoints");
    return res;
  }

  // range-image pixel indices of the last NARF detection (what NarfDescriptor needs; the reference recovers
  // indices with the O(K N) Tools::getIndices, tools.h:91-102, which yields cloud indices instead)
  const std::vector<int>& getNarfPixelIndices() const { return narf_pixel_indices_; }

 private:
  std::string kp_type_;
  double normal_radius_search_;
  std::vector<int> narf_pixel_indices_;
};

template <typename FeatureType>
class Features {
 public:
  Features() : feat_radius_search_(0.08), normal_radius_search_(0.05) {}
  explicit Features(typename pcl::Feature<PointRGB, FeatureType>::Ptr feature_extractor)
      : feature_extractor_(feature_extractor), feat_radius_search_(0.08), normal_radius_search_(0.05) {}
  Features(typename pcl::Feature<PointRGB, FeatureType>::Ptr feature_extractor, const double feat_radius_search,
           const double normal_radius_search)
      : feature_extractor_(feature_extractor), feat_radius_search_(feat_radius_search),
        normal_radius_search_(normal_radius_search) {}

  // features.h:175-196
  void compute(const PointCloudRGB::Ptr cloud, const PointCloudRGB::Ptr keypoints,
               typename pcl::PointCloud<FeatureType>::Ptr& descriptors) {
    auto from_normals =
        std::dynamic_pointer_cast<pcl::FeatureFromNormals<PointRGB, pcl::Normal, FeatureType>>(feature_extractor_);
    if (from_normals) {
      pcl::PointCloud<pcl::Normal>::Ptr normals(new pcl::PointCloud<pcl::Normal>);
      Tools::estimateNormals(cloud, normals, normal_radius_search_);
      from_normals->setInputNormals(normals);
    }
    feature_extractor_->setSearchSurface(cloud);
    feature_extractor_->setInputCloud(keypoints);
    pcl::search::KdTree<PointRGB>::Ptr kdtree(new pcl::search::KdTree<PointRGB>);
    feature_extractor_->setSearchMethod(kdtree);
    feature_extractor_->setRadiusSearch(feat_radius_search_);
    feature_extractor_->compute(*descriptors);
  }

  // features.h:224-251 (the two boost::threads become two device passes inside one pfx_match call)
  void findCorrespondences(typename pcl::PointCloud<FeatureType>::Ptr source,
                           typename pcl::PointCloud<FeatureType>::Ptr target, pcl::CorrespondencesPtr& correspondences) {
    pcl::registration::CorrespondenceEstimation<FeatureType> est;
    est.setInputSource(source);
    est.setInputTarget(target);
    est.determineReciprocalCorrespondences(*correspondences);
  }

  // features.h:253-273, kept for callers that want one direction only
  void getCorrespondences(typename pcl::PointCloud<FeatureType>::Ptr source,
                          typename pcl::PointCloud<FeatureType>::Ptr target, std::vector<int>& source2target) {
    const int k = 1;
    std::vector<int> k_indices(k);
    std::vector<float> k_dist(k);
    pcl::KdTreeFLANN<FeatureType> descriptor_kdtree;
    descriptor_kdtree.setInputCloud(target);
    source2target.assign(source->size(), -1);
    for (size_t i = 0; i < source->size(); ++i)
      if (descriptor_kdtree.nearestKSearch(*source, (int)i, k, k_indices, k_dist) > 0) source2target[i] = k_indices[0];
  }

  // features.h:282-297.  transformation: row-major 4x4 (Eigen::Matrix4f in the reference)
  void filterCorrespondences(const PointCloudRGB::Ptr source, const PointCloudRGB::Ptr target,
                             pcl::CorrespondencesPtr correspondences, pcl::CorrespondencesPtr& filtered_correspondences,
                             float transformation[16]) {
    pcl::registration::CorrespondenceRejectorSampleConsensus<PointRGB> rejector;
    rejector.setInputSource(source);
    rejector.setInputTarget(target);
    rejector.setInputCorrespondences(correspondences);
    rejector.setInlierThreshold(0.015);
    rejector.setMaximumIterations(1000);
    rejector.getCorrespondences(*filtered_correspondences);
    std::memcpy(transformation, rejector.getBestTransformation(), 16 * sizeof(float));
  }

  void setFeatureRadiusSearch(double r) { feat_radius_search_ = r; }
  void setNormalRadiusSearch(double r) { normal_radius_search_ = r; }

 private:
  typename pcl::Feature<PointRGB, FeatureType>::Ptr feature_extractor_;
  double feat_radius_search_;
  double normal_radius_search_;
};

// ---- PCD v0.7 binary reader for `FIELDS x y z rgb` (pcl::io::loadPCDFile at evaluation.cpp:226,231)
inline int loadPCDFile(const std::string& path, PointCloudRGB& cloud) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return -1;
  std::string line;
  size_t n = 0;
  bool binary = false;
  std::vector<std::string> fields;
  while (std::getline(f, line)) {
    std::istringstream ss(line);
    std::string key;
    ss >> key;
    if (key == "FIELDS") {
      std::string t;
      while (ss >> t) fields.push_back(t);
    } else if (key == "POINTS") {
      ss >> n;
    } else if (key == "VIEWPOINT") {
      ss >> cloud.sensor_origin_[0] >> cloud.sensor_origin_[1] >> cloud.sensor_origin_[2];
      float q[4];  // qw qx qy qz
      if (ss >> q[0] >> q[1] >> q[2] >> q[3]) std::memcpy(cloud.sensor_orientation_, q, sizeof(q));
    } else if (key == "DATA") {
      std::string mode;
      ss >> mode;
      binary = (mode == "binary");
      break;
    }
  }
  if (!binary || fields.size() < 3) return -2;
  const size_t rec = 4 * fields.size();
  std::vector<char> raw(n * rec);
  f.read(raw.data(), (std::streamsize)raw.size());
  if ((size_t)f.gcount() != raw.size()) return -3;
  cloud.points.resize(n);
  for (size_t i = 0; i < n; ++i) {
    const float* r = reinterpret_cast<const float*>(raw.data() + i * rec);
    cloud.points[i].x = r[0];
    cloud.points[i].y = r[1];
    cloud.points[i].z = r[2];
    if (fields.size() > 3) std::memcpy(&cloud.points[i].rgba, r + 3, 4);
  }
  cloud.width = (uint32_t)n;
  cloud.height = 1;
  cloud.is_dense = true;
  return 0;
}
