// pcl_compat.hpp — header-only C++ shim that puts the pcl::Feature / pcl::Keypoint style API used by
// the reference (include/pcl_feature_extraction/{features,keypoints,tools}.h) on top of the C ABI in
// include/pfx_b200.h.  PCL's headers are not available here, so the POD point types are defined with
// PCL's memory layout (pcl/impl/point_types.hpp): PointXYZRGB / PointXYZI / Normal 32 B,
// FPFHSignature33 132 B, SHOT352 1444 B, Correspondence 12 B (static_asserts below).
//
// Semantics kept from pcl::Feature::compute / initCompute (features/impl/feature.hpp): no exception
// ever leaves compute(); when a precondition fails the error is printed and the output cloud is
// emptied (width = height = 0); no surface => surface = input; exactly one of radius / k; rows of
// non-finite queries or empty neighbourhoods are NaN and clear is_dense.
#pragma once
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "../../include/pfx_b200.h"

namespace pcl {

// ------------------------------------------------------------------------------- point types
struct alignas(16) PointXYZ { float x = 0, y = 0, z = 0, pad_ = 1.0f; };
struct alignas(16) PointXYZRGB {
  float x = 0, y = 0, z = 0, pad_ = 1.0f;
  union { float rgb; uint32_t rgba; struct { uint8_t b, g, r, a; }; };
  float pad2_[3];
  PointXYZRGB() : rgba(0xff000000u), pad2_{0, 0, 0} {}
};
struct alignas(16) PointXYZI {
  float x = 0, y = 0, z = 0, pad_ = 1.0f;
  float intensity = 0;
  float pad2_[3] = {0, 0, 0};
};
struct alignas(16) Normal {
  float normal_x = 0, normal_y = 0, normal_z = 0, pad_ = 0;
  float curvature = 0;
  float pad2_[3] = {0, 0, 0};
};
struct FPFHSignature33 { float histogram[33]; static int descriptorSize() { return 33; } };
struct PFHSignature125 { float histogram[125]; static int descriptorSize() { return 125; } };
struct PrincipalCurvatures {
  float principal_curvature[3];  // principal_curvature_x / _y / _z
  float pc1, pc2;
  static int descriptorSize() { return 5; }
};
struct SHOT352 { float descriptor[352]; float rf[9]; static int descriptorSize() { return 352; } };
template <int N> struct Histogram { float histogram[N]; static int descriptorSize() { return N; } };
struct ShapeContext1980 { float descriptor[1980]; float rf[9]; static int descriptorSize() { return 1980; } };
struct MomentInvariants { float j1, j2, j3; static int descriptorSize() { return 3; } };
struct SHOT1344 { float descriptor[1344]; float rf[9]; static int descriptorSize() { return 1344; } };
struct ReferenceFrame { float x_axis[3], y_axis[3], z_axis[3]; };
// pcl::Narf36: 168 bytes; the representation used for matching is the 36 descriptor floats only
struct Narf36 { float x, y, z, roll, pitch, yaw; float descriptor[36]; static int descriptorSize() { return 36; } };
struct alignas(16) PointWithRange { float x = 0, y = 0, z = 0, pad_ = 1.0f; float range = 0; float pad2_[3] = {0, 0, 0}; };
struct Correspondence {
  int index_query = 0, index_match = -1;
  float distance = std::numeric_limits<float>::max();
};
static_assert(sizeof(PointXYZ) == 16 && sizeof(PointXYZRGB) == 32 && sizeof(PointXYZI) == 32, "PCL layout");
static_assert(sizeof(Normal) == 32 && sizeof(FPFHSignature33) == 132 && sizeof(SHOT352) == 1444, "PCL layout");
static_assert(sizeof(ReferenceFrame) == 36 && sizeof(Correspondence) == 12, "PCL layout");
static_assert(sizeof(Correspondence) == sizeof(pfx_correspondence), "ABI layout");
static_assert(sizeof(Narf36) == 168 && sizeof(PointWithRange) == 32, "PCL layout");
static_assert(sizeof(PFHSignature125) == 500 && sizeof(PrincipalCurvatures) == 20, "PCL layout");
static_assert(sizeof(SHOT1344) == 5412 && offsetof(PointXYZRGB, rgba) == 16, "PCL layout");
static_assert(sizeof(ShapeContext1980) == 7956, "PCL layout");

typedef std::vector<Correspondence> Correspondences;
typedef std::shared_ptr<Correspondences> CorrespondencesPtr;
typedef std::shared_ptr<const Correspondences> CorrespondencesConstPtr;
struct PointIndices { std::vector<int> indices; };
typedef std::shared_ptr<PointIndices> PointIndicesPtr;
typedef std::shared_ptr<const PointIndices> PointIndicesConstPtr;

template <typename PointT>
struct PointCloud {
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  std::vector<PointT> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  float sensor_origin_[4] = {0, 0, 0, 0};
  float sensor_orientation_[4] = {1, 0, 0, 0};  // w x y z
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); width = height = 0; }
  void resize(size_t n) { points.resize(n); width = (uint32_t)n; height = 1; }
  void push_back(const PointT& p) { points.push_back(p); width = (uint32_t)points.size(); height = 1; }
  PointT& operator[](size_t i) { return points[i]; }
  const PointT& operator[](size_t i) const { return points[i]; }
};

template <typename T> inline bool isFinite(const T& p) { return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z); }

// ------------------------------------------------------------------------------- context
namespace b200 {
// Contexts.  A pfx_ctx holds ONE resident surface with its voxel hashes and normals; the reference alternates between
// its source and target cloud inside the detector x descriptor loops (evaluation.cpp:272,302) and re-announces the
// same cloud to every Feature object (features.h:186-193).  The shim therefore keeps a small pool of contexts per
// (thread, device) and binds each CLOUD OBJECT to one of them: use(cloud) makes that context current, so a cloud that
// comes back finds its device copy, hashes and normals (the C ABI recognises an unchanged host cloud, pfx_set_reuse).
// PCL objects are not thread-safe; the pools are thread-local, so every thread works on its own contexts,
// on the device it chose with set_device() (default 0).
constexpr int POOL = 4;
struct Pool {
  int device = 0;
  pfx_ctx* ctx[POOL] = {nullptr, nullptr, nullptr, nullptr};
  const void* key[POOL] = {nullptr, nullptr, nullptr, nullptr};
  unsigned long long used[POOL] = {0, 0, 0, 0};
  unsigned long long tick = 0;
  int current = 0;
  ~Pool() {
    for (pfx_ctx*& c : ctx) {
      if (c) pfx_destroy(c);
      c = nullptr;
    }
  }
};
inline Pool& pool() { static thread_local Pool p; return p; }
// the device of this thread's contexts; changing it releases the contexts of the previous device
inline void set_device(int d) {
  Pool& p = pool();
  if (p.device == d) return;
  for (int i = 0; i < POOL; ++i) {
    if (p.ctx[i]) pfx_destroy(p.ctx[i]);
    p.ctx[i] = nullptr;
    p.key[i] = nullptr;
    p.used[i] = 0;
  }
  p.device = d;
  p.current = 0;
}
inline bool& strict_flag() { static thread_local bool s = false; return s; }
inline pfx_ctx* make_ctx(Pool& p, int slot) {
  if (!p.ctx[slot]) {
    int rc = pfx_create(p.device, &p.ctx[slot]);
    if (rc == 0 && strict_flag()) pfx_set_parity_mode(p.ctx[slot], PFX_PARITY_STRICT);
    if (rc != 0) {
      std::fprintf(stderr, "[pcl::b200] pfx_create(%d) failed with %d: an sm_100 GPU is required (no CPU fallback)\n",
                   p.device, rc);
      p.ctx[slot] = nullptr;
    }
  }
  return p.ctx[slot];
}
// PFX_PARITY_STRICT (reference-order arithmetic, include/pfx_b200.h) for this thread's contexts, present and future
inline void set_parity_strict(bool strict) {
  strict_flag() = strict;
  Pool& p = pool();
  for (int i = 0; i < POOL; ++i)
    if (p.ctx[i]) pfx_set_parity_mode(p.ctx[i], strict ? PFX_PARITY_STRICT : PFX_PARITY_FAST);
}
// the current context (the one bound to the cloud announced last)
inline pfx_ctx* ctx() {
  Pool& p = pool();
  return make_ctx(p, p.current);
}
// bind `cloud` (its address is the key) to a context and make it current: the context that already holds it, else a
// free one, else the least recently used
inline pfx_ctx* use(const void* cloud) {
  Pool& p = pool();
  int slot = -1;
  for (int i = 0; i < POOL; ++i)
    if (p.key[i] == cloud && p.ctx[i]) slot = i;
  if (slot < 0)
    for (int i = 0; i < POOL && slot < 0; ++i)
      if (!p.key[i]) slot = i;
  if (slot < 0) {
    slot = 0;
    for (int i = 1; i < POOL; ++i)
      if (p.used[i] < p.used[slot]) slot = i;
  }
  p.key[slot] = cloud;
  p.used[slot] = ++p.tick;
  p.current = slot;
  return make_ctx(p, slot);
}
// make a context obtained earlier current again (objects that own device state: RangeImage)
inline void use_ctx(pfx_ctx* c) {
  Pool& p = pool();
  for (int i = 0; i < POOL; ++i)
    if (p.ctx[i] == c && c) {
      p.current = i;
      p.used[i] = ++p.tick;
    }
}
inline bool ok(int rc, const char* who) {
  if (rc == 0) return true;
  pfx_ctx* c = pool().ctx[pool().current];
  std::fprintf(stderr, "[pcl::%s] %s (code %d)\n", who, c ? pfx_last_error(c) : "no context", rc);
  return false;
}
// counters of the current thread's contexts (pfx_reuse_info summed): what the reference's redundancy costs here
inline void reuse_totals(unsigned long long out6[6]) {
  for (int k = 0; k < 6; ++k) out6[k] = 0;
  Pool& p = pool();
  for (int i = 0; i < POOL; ++i)
    if (p.ctx[i]) {
      uint64_t v[6];
      if (pfx_reuse_info(p.ctx[i], v) == 0)
        for (int k = 0; k < 6; ++k) out6[k] += v[k];
    }
}
}  // namespace b200

// ------------------------------------------------------------------------------- search objects
// The reference creates search::KdTree objects only to hand them to setSearchMethod (features.h:192,
// tools.h:29, keypoints.h:186); the voxel hash behind pfx_set_surface replaces them, so these are tags.
namespace search {
template <typename PointT>
struct KdTree {
  typedef std::shared_ptr<KdTree<PointT>> Ptr;
  typename PointCloud<PointT>::ConstPtr input_;
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr& c) { input_ = c; }
  // batched form of the per-point loop at keypoints.h:411-424
  int nearestKSearchAll(int k, std::vector<int>& idx, std::vector<float>& d2) const {
    pfx_ctx* c = b200::use(input_.get());  // the context bound to this cloud
    if (!c || !input_) return 0;
    idx.assign(input_->size() * k, -1);
    d2.assign(input_->size() * k, 0.f);
    if (!b200::ok(pfx_set_surface(c, input_->points.data(), input_->size(), sizeof(PointT), PFX_HOST), "KdTree")) return 0;
    if (!b200::ok(pfx_knn(c, k, idx.data(), d2.data(), PFX_HOST), "KdTree")) return 0;
    return k;
  }
};
}  // namespace search

// KdTreeFLANN over DESCRIPTORS as used by features.h:253-273: setInputCloud(target) then
// nearestKSearch(source, i, 1, ...) for every i.  The first call for a given source cloud runs ONE
// batched exact 1-NN on the GPU (pfx_match_nn) and the loop then reads the cached answers.
template <typename FeatureT>
class KdTreeFLANN {
 public:
  void setInputCloud(const typename PointCloud<FeatureT>::ConstPtr& t) { target_ = t; cached_src_ = nullptr; }
  int nearestKSearch(const PointCloud<FeatureT>& src, int index, int k, std::vector<int>& k_indices,
                     std::vector<float>& k_sqr_distances) {
    if (k != 1 || !target_) return 0;
    if (cached_src_ != &src || cached_n_ != src.size()) {
      pfx_ctx* c = b200::ctx();
      if (!c) return 0;
      idx_.assign(src.size(), -1);
      d2_.assign(src.size(), 0.f);
      const int dim = FeatureT::descriptorSize();
      int rc = pfx_match_nn(c, reinterpret_cast<const float*>(src.points.data()), src.size(), sizeof(FeatureT),
                            reinterpret_cast<const float*>(target_->points.data()), target_->size(), sizeof(FeatureT),
                            dim, idx_.data(), d2_.data(), PFX_HOST);
      if (!b200::ok(rc, "KdTreeFLANN")) return 0;
      cached_src_ = &src;
      cached_n_ = src.size();
    }
    k_indices.resize(1);
    k_sqr_distances.resize(1);
    k_indices[0] = idx_[index];
    k_sqr_distances[0] = d2_[index];
    return idx_[index] >= 0 ? 1 : 0;
  }

 private:
  typename PointCloud<FeatureT>::ConstPtr target_;
  const PointCloud<FeatureT>* cached_src_ = nullptr;
  size_t cached_n_ = 0;
  std::vector<int> idx_;
  std::vector<float> d2_;
};

// ------------------------------------------------------------------------------- Feature base
template <typename PointInT, typename PointOutT>
class Feature {
 public:
  typedef std::shared_ptr<Feature<PointInT, PointOutT>> Ptr;
  typedef PointCloud<PointInT> PointCloudIn;
  typedef PointCloud<PointOutT> PointCloudOut;
  virtual ~Feature() {}
  void setInputCloud(const typename PointCloudIn::ConstPtr& c) { input_ = c; }
  void setSearchSurface(const typename PointCloudIn::ConstPtr& c) { surface_ = c; }
  template <typename Tree> void setSearchMethod(const Tree&) {}
  void setRadiusSearch(double r) { search_radius_ = r; }
  void setKSearch(int k) { k_ = k; }
  double getRadiusSearch() const { return search_radius_; }
  int getKSearch() const { return k_; }
  void setNumberOfThreads(unsigned) {}  // OMP variants: the GPU path has no thread knob

  void compute(PointCloudOut& output) {
    if (!initCompute()) {
      output.width = output.height = 0;
      output.points.clear();
      return;
    }
    output.points.resize(input_->size());
    if (input_->width * input_->height == input_->size() && input_->size() > 0) {
      output.width = input_->width;
      output.height = input_->height;
    } else {
      output.width = (uint32_t)input_->size();
      output.height = 1;
    }
    output.is_dense = input_->is_dense;
    if (!computeFeature(output)) {
      output.width = output.height = 0;
      output.points.clear();
    }
  }

 protected:
  virtual const char* name() const = 0;
  virtual bool computeFeature(PointCloudOut& output) = 0;
  virtual bool initCompute() {
    if (!input_) {
      std::fprintf(stderr, "[pcl::%s::compute] no input dataset given!\n", name());
      return false;
    }
    if (!surface_) surface_ = input_;  // "fake surface"
    if (search_radius_ != 0.0 && k_ != 0) {
      std::fprintf(stderr, "[pcl::%s::compute] Both radius (%f) and K (%d) defined! Set one of them to zero first.\n",
                   name(), search_radius_, k_);
      return false;
    }
    if (search_radius_ == 0.0 && k_ == 0) {
      std::fprintf(stderr, "[pcl::%s::compute] Neither radius nor K defined!\n", name());
      return false;
    }
    return b200::ctx() != nullptr;
  }
  // surface + queries to the device; dense when input and surface are the same cloud
  bool upload() {
    pfx_ctx* c = b200::use(surface_.get());  // the context bound to this cloud
    if (!b200::ok(pfx_set_surface(c, surface_->points.data(), surface_->size(), sizeof(PointInT), PFX_HOST), name())) return false;
    pfx_set_viewpoint(c, surface_->sensor_origin_[0], surface_->sensor_origin_[1], surface_->sensor_origin_[2]);
    if (input_.get() == surface_.get()) return b200::ok(pfx_set_queries(c, nullptr, 0, 0, PFX_HOST), name());
    return b200::ok(pfx_set_queries(c, input_->points.data(), input_->size(), sizeof(PointInT), PFX_HOST), name());
  }
  typename PointCloudIn::ConstPtr input_, surface_;
  double search_radius_ = 0.0;
  int k_ = 0;
};

template <typename PointInT, typename PointNT, typename PointOutT>
class FeatureFromNormals : public Feature<PointInT, PointOutT> {
 public:
  typedef std::shared_ptr<FeatureFromNormals<PointInT, PointNT, PointOutT>> Ptr;
  void setInputNormals(const typename PointCloud<PointNT>::ConstPtr& n) { normals_ = n; }

 protected:
  bool initCompute() override {
    if (!Feature<PointInT, PointOutT>::initCompute()) return false;
    if (!normals_) {
      std::fprintf(stderr, "[pcl::%s::initCompute] No input dataset containing normals was given!\n", this->name());
      return false;
    }
    if (normals_->size() != this->surface_->size()) {
      std::fprintf(stderr, "[pcl::%s::initCompute] The number of points in the surface differs from the number of normals!\n", this->name());
      return false;
    }
    return true;
  }
  bool uploadWithNormals() {
    if (!this->upload()) return false;
    return b200::ok(pfx_set_surface_normals(b200::ctx(), normals_->points.data(), normals_->size(), sizeof(PointNT), 4, PFX_HOST),
                    this->name());
  }
  typename PointCloud<PointNT>::ConstPtr normals_;
};

// ------------------------------------------------------------------------------- NormalEstimation
template <typename PointInT, typename PointOutT>
class NormalEstimation : public Feature<PointInT, PointOutT> {
 public:
  void setViewPoint(float x, float y, float z) { vp_[0] = x; vp_[1] = y; vp_[2] = z; use_origin_ = false; }

 protected:
  const char* name() const override { return "NormalEstimation"; }
  bool computeFeature(PointCloud<PointOutT>& output) override {
    if (!this->upload()) return false;
    pfx_ctx* c = b200::ctx();
    if (!use_origin_) pfx_set_viewpoint(c, vp_[0], vp_[1], vp_[2]);
    for (auto& p : output.points) p = PointOutT();
    int rc = pfx_normals(c, this->search_radius_, this->k_, output.points.data(), sizeof(PointOutT), 4, PFX_HOST);
    if (!b200::ok(rc, name())) return false;
    for (const auto& p : output.points)
      if (!std::isfinite(p.normal_x)) { output.is_dense = false; break; }
    return true;
  }
  float vp_[3] = {0, 0, 0};
  bool use_origin_ = true;
};
template <typename PointInT, typename PointOutT>
class NormalEstimationOMP : public NormalEstimation<PointInT, PointOutT> {
 public:
  explicit NormalEstimationOMP(unsigned = 0) {}
};

// ------------------------------------------------------------------------------- FPFH
template <typename PointInT, typename PointNT, typename PointOutT = FPFHSignature33>
class FPFHEstimation : public FeatureFromNormals<PointInT, PointNT, PointOutT> {
 protected:
  const char* name() const override { return "FPFHEstimation"; }
  bool computeFeature(PointCloud<PointOutT>& output) override {
    if (!this->uploadWithNormals()) return false;
    int rc = pfx_fpfh(b200::ctx(), this->search_radius_, this->k_, reinterpret_cast<float*>(output.points.data()),
                      sizeof(PointOutT), PFX_HOST);
    if (!b200::ok(rc, name())) return false;
    for (const auto& p : output.points)
      if (!std::isfinite(p.histogram[0])) { output.is_dense = false; break; }
    return true;
  }
};
template <typename PointInT, typename PointNT, typename PointOutT = FPFHSignature33>
class FPFHEstimationOMP : public FPFHEstimation<PointInT, PointNT, PointOutT> {};

// ------------------------------------------------------------------------------- spin images
// pcl::SpinImageEstimation with its defaults (image width 8 -> Histogram<153>); the normals are those of the INPUT
// cloud (one per query), as the reference sets them at evaluation.cpp:521-529
template <typename PointInT, typename PointNT, typename PointOutT = Histogram<153>>
class SpinImageEstimation : public Feature<PointInT, PointOutT> {
 public:
  explicit SpinImageEstimation(unsigned int image_width = 8, double support_angle_cos = 0.0, unsigned int min_pts_neighb = 0)
      : image_width_(image_width), support_angle_cos_(support_angle_cos), min_pts_neighb_(min_pts_neighb) {}
  void setInputNormals(const typename PointCloud<PointNT>::ConstPtr& n) { input_normals_ = n; }

 protected:
  const char* name() const override { return "SpinImageEstimation"; }
  bool initCompute() override {
    if (!Feature<PointInT, PointOutT>::initCompute()) return false;
    if (!input_normals_) {
      std::fprintf(stderr, "[pcl::%s::initCompute] No input dataset containing normals was given!\n", name());
      return false;
    }
    if (input_normals_->size() != this->input_->size()) {
      std::fprintf(stderr, "[pcl::%s::initCompute] The number of points in the input dataset differs from the number of points in the dataset containing the normals!\n", name());
      return false;
    }
    if (this->k_ != 0) {
      std::fprintf(stderr, "[pcl::%s::initCompute] K-nearest neighbor search for spin images not implemented. Used a neighborhood radius search instead\n", name());
      return false;
    }
    if (image_width_ != 8 || support_angle_cos_ != 0.0 || PointOutT::descriptorSize() != 153) {
      std::fprintf(stderr, "[pcl::%s::initCompute] only the default 8-bin, full-support spin image (Histogram<153>) is on the B200 path\n", name());
      return false;
    }
    return true;
  }
  bool computeFeature(PointCloud<PointOutT>& output) override {
    if (!this->upload()) return false;
    int rc = pfx_spin_image153(b200::ctx(), this->search_radius_, input_normals_->points.data(), input_normals_->size(),
                               sizeof(PointNT), reinterpret_cast<float*>(output.points.data()), sizeof(PointOutT), PFX_HOST);
    if (!b200::ok(rc, name())) return false;
    for (const auto& p : output.points)
      if (!std::isfinite(p.histogram[0])) { output.is_dense = false; break; }
    return true;
  }
  typename PointCloud<PointNT>::ConstPtr input_normals_;
  unsigned int image_width_;
  double support_angle_cos_;
  unsigned int min_pts_neighb_;
};

// ------------------------------------------------------------------------------- Unique Shape Context (no normals)
template <typename PointInT, typename PointOutT = ShapeContext1980, typename PointRFT = ReferenceFrame>
class UniqueShapeContext : public Feature<PointInT, PointOutT> {
 public:
  typedef std::shared_ptr<UniqueShapeContext<PointInT, PointOutT, PointRFT>> Ptr;
  void setMinimalRadius(double r) { min_radius_ = r; }
  void setPointDensityRadius(double r) { point_density_radius_ = r; }
  void setLocalRadius(double r) { local_radius_ = r; }
  void setInputReferenceFrames(const typename PointCloud<PointRFT>::ConstPtr& f) { frames_ = f; }

 protected:
  const char* name() const override { return "UniqueShapeContext"; }
  bool initCompute() override {
    if (!Feature<PointInT, PointOutT>::initCompute()) return false;
    if (this->search_radius_ < min_radius_) {
      std::fprintf(stderr, "[pcl::%s::initCompute] search_radius_ must be GREATER than min_radius_.\n", name());
      return false;
    }
    if (frames_ && frames_->size() != this->input_->size()) {
      std::fprintf(stderr, "[pcl::%s::initCompute] The number of reference frames differs from the number of input points!\n", name());
      return false;
    }
    return true;
  }
  bool computeFeature(PointCloud<PointOutT>& output) override {
    if (!this->upload()) return false;
    const float* lrf_in = frames_ ? reinterpret_cast<const float*>(frames_->points.data()) : nullptr;
    int rc = pfx_usc1980(b200::ctx(), this->search_radius_, min_radius_, point_density_radius_, local_radius_, lrf_in,
                         reinterpret_cast<float*>(output.points.data()), sizeof(PointOutT), PFX_HOST);
    if (!b200::ok(rc, name())) return false;
    for (const auto& p : output.points)
      if (!std::isfinite(p.descriptor[0])) { output.is_dense = false; break; }
    return true;
  }
  typename PointCloud<PointRFT>::ConstPtr frames_;
  double min_radius_ = 0.1, point_density_radius_ = 0.2, local_radius_ = 2.5;  // PCL's defaults
};

// ------------------------------------------------------------------------------- 3D Shape Context
// pcl::ShapeContext3DEstimation<PointXYZRGB, Normal, ShapeContext1980> as configured at evaluation.cpp:319-345
// (setMinimalRadius, setPointDensityRadius; radius through Features<T>::compute).  The random tangent direction of
// every frame follows the library's seeded contract (pfx_sc3d1980); setSeed() is an addition of this shim.
template <typename PointInT, typename PointNT, typename PointOutT = ShapeContext1980>
class ShapeContext3DEstimation : public FeatureFromNormals<PointInT, PointNT, PointOutT> {
 public:
  typedef std::shared_ptr<ShapeContext3DEstimation<PointInT, PointNT, PointOutT>> Ptr;
  void setMinimalRadius(double r) { min_radius_ = r; }
  double getMinimalRadius() const { return min_radius_; }
  void setPointDensityRadius(double r) { point_density_radius_ = r; }
  double getPointDensityRadius() const { return point_density_radius_; }
  void setSeed(unsigned long long seed) { seed_ = seed; }

 protected:
  const char* name() const override { return "ShapeContext3DEstimation"; }
  bool initCompute() override {
    if (!FeatureFromNormals<PointInT, PointNT, PointOutT>::initCompute()) return false;
    if (this->search_radius_ < min_radius_) {
      std::fprintf(stderr, "[pcl::%s::initCompute] search_radius_ must be GREATER than min_radius_.\n", name());
      return false;
    }
    return true;
  }
  bool computeFeature(PointCloud<PointOutT>& output) override {
    if (!this->uploadWithNormals()) return false;
    int rc = pfx_sc3d1980(b200::ctx(), this->search_radius_, min_radius_, point_density_radius_, seed_,
                          reinterpret_cast<float*>(output.points.data()), sizeof(PointOutT), nullptr, PFX_HOST);
    if (!b200::ok(rc, name())) return false;
    for (const auto& p : output.points)
      if (!std::isfinite(p.descriptor[0])) { output.is_dense = false; break; }
    return true;
  }
  double min_radius_ = 0.1, point_density_radius_ = 0.2;  // PCL's defaults
  unsigned long long seed_ = 12345;
};

// ------------------------------------------------------------------------------- MomentInvariants (no normals)
template <typename PointInT, typename PointOutT = MomentInvariants>
class MomentInvariantsEstimation : public Feature<PointInT, PointOutT> {
 protected:
  const char* name() const override { return "MomentInvariantsEstimation"; }
  bool computeFeature(PointCloud<PointOutT>& output) override {
    if (!this->upload()) return false;
    int rc = pfx_moment_invariants(b200::ctx(), this->search_radius_, this->k_, reinterpret_cast<float*>(output.points.data()),
                                   sizeof(PointOutT), PFX_HOST);
    if (!b200::ok(rc, name())) return false;
    for (const auto& p : output.points)
      if (!std::isfinite(p.j1)) { output.is_dense = false; break; }
    return true;
  }
};

// ------------------------------------------------------------------------------- PFH, PrincipalCurvatures
template <typename PointInT, typename PointNT, typename PointOutT = PFHSignature125>
class PFHEstimation : public FeatureFromNormals<PointInT, PointNT, PointOutT> {
 protected:
  const char* name() const override { return "PFHEstimation"; }
  bool computeFeature(PointCloud<PointOutT>& output) override {
    if (!this->uploadWithNormals()) return false;
    int rc = pfx_pfh125(b200::ctx(), this->search_radius_, this->k_, reinterpret_cast<float*>(output.points.data()),
                        sizeof(PointOutT), PFX_HOST);
    if (!b200::ok(rc, name())) return false;
    for (const auto& p : output.points)
      if (!std::isfinite(p.histogram[0])) { output.is_dense = false; break; }
    return true;
  }
};
template <typename PointInT, typename PointNT, typename PointOutT = PrincipalCurvatures>
class PrincipalCurvaturesEstimation : public FeatureFromNormals<PointInT, PointNT, PointOutT> {
 protected:
  const char* name() const override { return "PrincipalCurvaturesEstimation"; }
  bool computeFeature(PointCloud<PointOutT>& output) override {
    if (!this->uploadWithNormals()) return false;
    int rc = pfx_principal_curvatures(b200::ctx(), this->search_radius_, this->k_,
                                      reinterpret_cast<float*>(output.points.data()), sizeof(PointOutT), PFX_HOST);
    if (!b200::ok(rc, name())) return false;
    for (const auto& p : output.points)
      if (!std::isfinite(p.pc1)) { output.is_dense = false; break; }
    return true;
  }
};

// ------------------------------------------------------------------------------- SHOT
template <typename PointInT, typename PointNT, typename PointOutT = SHOT352, typename PointRFT = ReferenceFrame>
class SHOTEstimation : public FeatureFromNormals<PointInT, PointNT, PointOutT> {
 public:
  void setInputReferenceFrames(const typename PointCloud<PointRFT>::ConstPtr& f) { frames_ = f; }
  void setLRFRadius(float r) { lrf_radius_ = r; }

 protected:
  const char* name() const override { return "SHOTEstimation"; }
  bool initCompute() override {
    if (!FeatureFromNormals<PointInT, PointNT, PointOutT>::initCompute()) return false;
    if (this->k_ != 0) {  // SHOT cannot work with k-search
      std::fprintf(stderr, "[pcl::%s::initCompute] Error! Search method set to k-neighborhood. Call setKSearch(0) and setRadiusSearch( radius ) to use this class.\n", name());
      return false;
    }
    if (frames_ && frames_->size() != this->input_->size()) {
      std::fprintf(stderr, "[pcl::%s::initCompute] The number of reference frames differs from the number of input points!\n", name());
      return false;
    }
    return true;
  }
  bool computeFeature(PointCloud<PointOutT>& output) override {
    if (!this->uploadWithNormals()) return false;
    pfx_ctx* c = b200::ctx();
    std::vector<float> lrf;
    const float* lrf_in = nullptr;
    if (frames_) {
      lrf_in = reinterpret_cast<const float*>(frames_->points.data());
    } else if (lrf_radius_ > 0 && (double)lrf_radius_ != this->search_radius_) {
      lrf.resize(output.points.size() * 9);
      if (!b200::ok(pfx_shot_lrf(c, lrf_radius_, lrf.data(), PFX_HOST), name())) return false;
      lrf_in = lrf.data();
    }
    int rc = pfx_shot352(c, this->search_radius_, lrf_in, reinterpret_cast<float*>(output.points.data()),
                         sizeof(PointOutT), PFX_HOST);
    if (!b200::ok(rc, name())) return false;
    for (const auto& p : output.points)
      if (!std::isfinite(p.descriptor[0])) { output.is_dense = false; break; }
    return true;
  }
  typename PointCloud<PointRFT>::ConstPtr frames_;
  float lrf_radius_ = 0;
};
template <typename PointInT, typename PointNT, typename PointOutT = SHOT352, typename PointRFT = ReferenceFrame>
class SHOTEstimationOMP : public SHOTEstimation<PointInT, PointNT, PointOutT, PointRFT> {};

// SHOT1344: shape + colour; PointInT carries PCL's packed rgba word (PointXYZRGB)
template <typename PointInT, typename PointNT, typename PointOutT = SHOT1344, typename PointRFT = ReferenceFrame>
class SHOTColorEstimation : public FeatureFromNormals<PointInT, PointNT, PointOutT> {
 public:
  void setInputReferenceFrames(const typename PointCloud<PointRFT>::ConstPtr& f) { frames_ = f; }

 protected:
  const char* name() const override { return "SHOTColorEstimation"; }
  bool initCompute() override {
    if (!FeatureFromNormals<PointInT, PointNT, PointOutT>::initCompute()) return false;
    if (this->k_ != 0) {
      std::fprintf(stderr, "[pcl::%s::initCompute] Error! Search method set to k-neighborhood. Call setKSearch(0) and setRadiusSearch( radius ) to use this class.\n", name());
      return false;
    }
    if (frames_ && frames_->size() != this->input_->size()) {
      std::fprintf(stderr, "[pcl::%s::initCompute] The number of reference frames differs from the number of input points!\n", name());
      return false;
    }
    return true;
  }
  bool computeFeature(PointCloud<PointOutT>& output) override {
    if (!this->uploadWithNormals()) return false;
    pfx_ctx* c = b200::ctx();
    const bool dense = this->input_.get() == this->surface_.get();
    if (!b200::ok(pfx_set_surface_colors(c, &this->surface_->points.data()->rgba, this->surface_->size(), sizeof(PointInT), PFX_HOST), name()))
      return false;
    if (!dense && !b200::ok(pfx_set_query_colors(c, &this->input_->points.data()->rgba, this->input_->size(), sizeof(PointInT), PFX_HOST), name()))
      return false;
    const float* lrf_in = frames_ ? reinterpret_cast<const float*>(frames_->points.data()) : nullptr;
    int rc = pfx_shot1344(c, this->search_radius_, lrf_in, reinterpret_cast<float*>(output.points.data()), sizeof(PointOutT), PFX_HOST);
    if (!b200::ok(rc, name())) return false;
    for (const auto& p : output.points)
      if (!std::isfinite(p.descriptor[0])) { output.is_dense = false; break; }
    return true;
  }
  typename PointCloud<PointRFT>::ConstPtr frames_;
};
template <typename PointInT, typename PointNT, typename PointOutT = SHOT1344, typename PointRFT = ReferenceFrame>
class SHOTColorEstimationOMP : public SHOTColorEstimation<PointInT, PointNT, PointOutT, PointRFT> {};

// ------------------------------------------------------------------------------- keypoints
template <typename PointInT, typename PointOutT>
class Keypoint {
 public:
  virtual ~Keypoint() {}
  void setInputCloud(const typename PointCloud<PointInT>::ConstPtr& c) { input_ = c; }
  void setSearchSurface(const typename PointCloud<PointInT>::ConstPtr& c) { surface_ = c; }
  template <typename Tree> void setSearchMethod(const Tree&) {}
  void setRadiusSearch(double r) { search_radius_ = r; }
  void setKSearch(int k) { k_ = k; }
  PointIndicesConstPtr getKeypointsIndices() const { return keypoints_indices_; }
  void compute(PointCloud<PointOutT>& output) {
    keypoints_indices_.reset(new PointIndices);
    output.points.clear();
    output.width = output.height = 0;
    if (!input_) {
      std::fprintf(stderr, "[pcl::Keypoint::compute] no input dataset given!\n");
      return;
    }
    if (!b200::ctx()) return;
    detectKeypoints(output);
    output.width = (uint32_t)output.points.size();
    output.height = 1;
    output.is_dense = true;
  }

 protected:
  virtual void detectKeypoints(PointCloud<PointOutT>& output) = 0;
  typename PointCloud<PointInT>::ConstPtr input_, surface_;
  double search_radius_ = 0.0;
  int k_ = 0;
  PointIndicesPtr keypoints_indices_;
};

template <typename PointInT, typename PointOutT, typename NormalT = Normal>
class ISSKeypoint3D : public Keypoint<PointInT, PointOutT> {
 public:
  explicit ISSKeypoint3D(double salient_radius = 0.0001) : salient_radius_(salient_radius) {}
  void setSalientRadius(double r) { salient_radius_ = r; }
  void setNonMaxRadius(double r) { non_max_radius_ = r; }
  void setMinNeighbors(int n) { min_neighbors_ = n; }
  void setThreshold21(double g) { gamma_21_ = g; }
  void setThreshold32(double g) { gamma_32_ = g; }
  void setNormalRadius(double) {}
  void setBorderRadius(double r) { border_radius_ = r; }
  void setNumberOfThreads(unsigned) {}

 protected:
  void detectKeypoints(PointCloud<PointOutT>& output) override {
    pfx_ctx* c = b200::use(this->input_.get());  // the context bound to this cloud
    if (border_radius_ > 0) {
      std::fprintf(stderr, "[pcl::ISSKeypoint3D] border estimation (border_radius > 0) is outside the reference's path\n");
      return;
    }
    const auto& in = *this->input_;
    if (!b200::ok(pfx_set_surface(c, in.points.data(), in.size(), sizeof(PointInT), PFX_HOST), "ISSKeypoint3D")) return;
    std::vector<int> idx(in.size());
    size_t n = 0;
    int rc = pfx_iss(c, salient_radius_, non_max_radius_, min_neighbors_, gamma_21_, gamma_32_, idx.data(), idx.size(), &n,
                     nullptr, PFX_HOST);
    if (!b200::ok(rc, "ISSKeypoint3D")) return;
    output.points.resize(n);
    this->keypoints_indices_->indices.assign(idx.begin(), idx.begin() + n);
    for (size_t i = 0; i < n; ++i) {  // upstream copies xyz only; other fields stay default-constructed
      output.points[i] = PointOutT();
      output.points[i].x = in.points[idx[i]].x;
      output.points[i].y = in.points[idx[i]].y;
      output.points[i].z = in.points[idx[i]].z;
    }
  }
  double salient_radius_, non_max_radius_ = 0.0, gamma_21_ = 0.975, gamma_32_ = 0.975, border_radius_ = 0.0;
  int min_neighbors_ = 5;
};

template <typename PointInT, typename PointOutT, typename NormalT = Normal>
class HarrisKeypoint3D : public Keypoint<PointInT, PointOutT> {
 public:
  enum ResponseMethod { HARRIS = 1 };
  explicit HarrisKeypoint3D(ResponseMethod = HARRIS, float radius = 0.01f, float threshold = 0.0f)
      : threshold_(threshold) { this->search_radius_ = radius; }
  void setRadius(float r) { this->search_radius_ = r; }
  void setThreshold(float t) { threshold_ = t; }
  void setNonMaxSupression(bool b) { nonmax_ = b; }
  void setRefine(bool b) { refine_ = b; }
  void setNormals(const typename PointCloud<NormalT>::ConstPtr& n) { normals_ = n; }
  void setNumberOfThreads(unsigned) {}
  // cloud indices of the corners after the reference's 1 cm snap (keypoints.h:360-395), -1 = dropped
  const std::vector<int>& getSnappedIndices() const { return snapped_; }

 protected:
  void detectKeypoints(PointCloud<PointOutT>& output) override {
    pfx_ctx* c = b200::use(this->input_.get());  // the context bound to this cloud
    const auto& in = *this->input_;
    if (!b200::ok(pfx_set_surface(c, in.points.data(), in.size(), sizeof(PointInT), PFX_HOST), "HarrisKeypoint3D")) return;
    pfx_set_viewpoint(c, in.sensor_origin_[0], in.sensor_origin_[1], in.sensor_origin_[2]);
    if (normals_ && !b200::ok(pfx_set_surface_normals(c, normals_->points.data(), normals_->size(), sizeof(NormalT), 4, PFX_HOST),
                              "HarrisKeypoint3D")) return;
    const size_t cap = in.size();
    std::vector<float> resp(cap), xyz(cap * 3);
    std::vector<int> idx(cap);
    snapped_.assign(cap, -1);
    size_t n = 0;
    int rc = pfx_harris3d(c, this->search_radius_, threshold_, nonmax_ ? 1 : 0, refine_ ? 1 : 0, 1e-4f, resp.data(),
                          idx.data(), xyz.data(), snapped_.data(), cap, &n, PFX_HOST);
    if (!b200::ok(rc, "HarrisKeypoint3D")) return;
    if (!nonmax_) {
      output.points.resize(in.size());
      for (size_t i = 0; i < in.size(); ++i) {
        output.points[i].x = in.points[i].x; output.points[i].y = in.points[i].y; output.points[i].z = in.points[i].z;
        output.points[i].intensity = resp[i];
      }
      snapped_.clear();
      return;
    }
    output.points.resize(n);
    snapped_.resize(n);
    this->keypoints_indices_->indices.assign(idx.begin(), idx.begin() + n);
    for (size_t i = 0; i < n; ++i) {
      output.points[i].x = xyz[3 * i]; output.points[i].y = xyz[3 * i + 1]; output.points[i].z = xyz[3 * i + 2];
      output.points[i].intensity = resp[idx[i]];
    }
  }
  float threshold_;
  bool nonmax_ = true, refine_ = true;
  typename PointCloud<NormalT>::ConstPtr normals_;
  std::vector<int> snapped_;
};

// pcl::HarrisKeypoint6D<PointXYZRGB, PointXYZI> as driven at keypoints.h:166-179 (setNonMaxSupression, setThreshold,
// radius left at 0.01, refinement left on).  The colours of the input cloud supply the intensity.
template <typename PointInT, typename PointOutT, typename NormalT = Normal>
class HarrisKeypoint6D : public Keypoint<PointInT, PointOutT> {
 public:
  explicit HarrisKeypoint6D(float radius = 0.01f, float threshold = 0.0f) : threshold_(threshold) { this->search_radius_ = radius; }
  void setRadius(float r) { this->search_radius_ = r; }
  void setThreshold(float t) { threshold_ = t; }
  void setNonMaxSupression(bool b) { nonmax_ = b; }
  void setRefine(bool b) { refine_ = b; }
  void setNumberOfThreads(unsigned) {}
  const std::vector<int>& getSnappedIndices() const { return snapped_; }

 protected:
  void detectKeypoints(PointCloud<PointOutT>& output) override {
    pfx_ctx* c = b200::use(this->input_.get());
    const auto& in = *this->input_;
    if (!c) return;
    if (!b200::ok(pfx_set_surface(c, in.points.data(), in.size(), sizeof(PointInT), PFX_HOST), "HarrisKeypoint6D")) return;
    pfx_set_viewpoint(c, in.sensor_origin_[0], in.sensor_origin_[1], in.sensor_origin_[2]);
    if (!b200::ok(pfx_set_surface_colors(c, &in.points.data()->rgba, in.size(), sizeof(PointInT), PFX_HOST), "HarrisKeypoint6D")) return;
    const size_t cap = in.size();
    std::vector<float> resp(cap), xyz(cap * 3);
    std::vector<int> idx(cap);
    snapped_.assign(cap, -1);
    size_t n = 0;
    int rc = pfx_harris6d(c, this->search_radius_, threshold_, nonmax_ ? 1 : 0, refine_ ? 1 : 0, 1e-4f, resp.data(),
                          idx.data(), xyz.data(), snapped_.data(), cap, &n, PFX_HOST);
    if (!b200::ok(rc, "HarrisKeypoint6D")) return;
    if (!nonmax_) {
      output.points.resize(in.size());
      for (size_t i = 0; i < in.size(); ++i) {
        output.points[i].x = in.points[i].x; output.points[i].y = in.points[i].y; output.points[i].z = in.points[i].z;
        output.points[i].intensity = resp[i];
      }
      snapped_.clear();
      return;
    }
    output.points.resize(n);
    snapped_.resize(n);
    this->keypoints_indices_->indices.assign(idx.begin(), idx.begin() + n);
    for (size_t i = 0; i < n; ++i) {
      output.points[i].x = xyz[3 * i]; output.points[i].y = xyz[3 * i + 1]; output.points[i].z = xyz[3 * i + 2];
      output.points[i].intensity = resp[idx[i]];
    }
  }
  float threshold_;
  bool nonmax_ = true, refine_ = true;
  std::vector<int> snapped_;
};

// ------------------------------------------------------------------------------- filters
// pcl::VoxelGrid (config C1 ingest): setInputCloud / setLeafSize / filter.  Centroids of xyz in ascending
// voxel-index order; colour is not carried (downsample_all_data for rgb is outside the path).
template <typename PointT>
class VoxelGrid {
 public:
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr& c) { input_ = c; }
  void setLeafSize(float lx, float ly, float lz) { leaf_ = lx; uniform_ = (lx == ly && ly == lz); }
  void filter(PointCloud<PointT>& output) {
    output.points.clear();
    output.width = output.height = 0;
    pfx_ctx* c = b200::use(input_.get());  // the context bound to this cloud
    if (!c || !input_) return;
    if (!uniform_) {
      std::fprintf(stderr, "[pcl::VoxelGrid] only cubic leaves are implemented\n");
      return;
    }
    if (!b200::ok(pfx_set_surface(c, input_->points.data(), input_->size(), sizeof(PointT), PFX_HOST), "VoxelGrid")) return;
    std::vector<float> xyz(3 * std::max<size_t>(input_->size(), 1));
    size_t n = 0;
    int rc = pfx_voxel_grid(c, leaf_, xyz.data(), input_->size(), &n, PFX_HOST);
    if (rc == PFX_E_PRECOND) {  // PCL: warns and returns the input unchanged when the index would overflow
      std::fprintf(stderr, "[pcl::VoxelGrid::applyFilter] Leaf size is too small for the input dataset. Integer indices would overflow.\n");
      output = *input_;
      return;
    }
    if (!b200::ok(rc, "VoxelGrid")) return;
    output.points.resize(n);
    for (size_t i = 0; i < n; ++i) {
      output.points[i].x = xyz[3 * i];
      output.points[i].y = xyz[3 * i + 1];
      output.points[i].z = xyz[3 * i + 2];
    }
    output.width = (uint32_t)n;
    output.height = 1;
    output.is_dense = true;
    output.sensor_origin_ = input_->sensor_origin_;
  }

 private:
  typename PointCloud<PointT>::ConstPtr input_;
  float leaf_ = 0.01f;
  bool uniform_ = true;
};

// ------------------------------------------------------------------------------- range image / NARF
// The subset of pcl::RangeImage(Planar), RangeImageBorderExtractor, NarfKeypoint and NarfDescriptor that
// the reference drives (keypoints.h:204-224, tools.h:65-76, evaluation.cpp:629-637).  The image itself
// lives in the library's context; these objects hold its geometry and re-install it before use, so several
// RangeImage objects may coexist like in the reference.
// row-major 4x4 rigid transform, the shim's stand-in for Eigen::Affine3f (sensor pose: world <- sensor)
typedef std::array<float, 16> Affine3f;
inline Affine3f identityPose() { return Affine3f{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; }
// Eigen::Affine3f(Eigen::Translation3f(origin)) * Eigen::Affine3f(orientation) as built at keypoints.h:207-210;
// origin = cloud.sensor_origin_ (x y z), q = cloud.sensor_orientation_ (w x y z)
inline Affine3f poseFromOriginAndOrientation(const float origin[4], const float q[4]) {
  const float w = q[0], x = q[1], y = q[2], z = q[3];
  return Affine3f{1 - 2 * (y * y + z * z), 2 * (x * y - z * w),     2 * (x * z + y * w),     origin[0],
                  2 * (x * y + z * w),     1 - 2 * (x * x + z * z), 2 * (y * z - x * w),     origin[1],
                  2 * (x * z - y * w),     2 * (y * z + x * w),     1 - 2 * (x * x + y * y), origin[2],
                  0, 0, 0, 1};
}

class RangeImage {
 public:
  enum CoordinateFrame { CAMERA_FRAME = 0, LASER_FRAME = 1 };
  virtual ~RangeImage() {}
  uint32_t width = 0, height = 0;
  std::vector<PointWithRange> points;  // world coordinates, as in PCL
  const Affine3f& getSensorPose() const { return pose_; }
  // RangeImage::createFromPointCloud (config C3): spherical projection, cropped.  sensor pose: identity.
  template <typename CloudT>
  void createFromPointCloud(const CloudT& cloud, float angular_resolution, float max_angle_width, float max_angle_height,
                            const Affine3f* sensor_pose = nullptr,
                            CoordinateFrame frame = CAMERA_FRAME, float noise_level = 0.0f, float min_range = 0.0f,
                            int border_size = 0) {
    pfx_ctx* c = b200::use(&cloud);  // the context bound to this cloud
    valid_ = false;
    pose_ = sensor_pose ? *sensor_pose : identityPose();
    if (c && !b200::ok(pfx_range_image_set_pose(c, pose_.data()), "RangeImage")) return;
    if (!c || frame != CAMERA_FRAME || noise_level != 0.0f) {
      std::fprintf(stderr, "[pcl::RangeImage] only CAMERA_FRAME with noise_level 0 is implemented\n");
      return;
    }
    if (!b200::ok(pfx_set_surface(c, cloud.points.data(), cloud.size(), sizeof(cloud.points[0]), PFX_HOST), "RangeImage")) return;
    if (!b200::ok(pfx_range_image_spherical(c, angular_resolution, max_angle_width, max_angle_height, min_range, border_size, &desc_), "RangeImage")) return;
    fetch(c);
  }
  bool isValid(int index) const { return index >= 0 && (size_t)index < points.size() && std::isfinite(points[index].range); }
  const PointWithRange& getPoint(int index) const { return points[index]; }
  // make this image the current one of the context (no-op cost when it already is)
  bool install() const {
    pfx_ctx* c = b200::ctx();
    if (!c || !valid_) return false;
    if (!b200::ok(pfx_range_image_set_pose(c, pose_.data()), "RangeImage")) return false;
    return b200::ok(pfx_range_image_set(c, &desc_, raw_.data(), PFX_HOST), "RangeImage");
  }
  const pfx_range_image_desc& desc() const { return desc_; }

 protected:
  void fetch(pfx_ctx* c) {
    raw_.assign((size_t)4 * desc_.width * desc_.height, 0.f);
    if (!b200::ok(pfx_range_image_get(c, &desc_, raw_.data(), PFX_HOST), "RangeImage")) return;
    width = (uint32_t)desc_.width;
    height = (uint32_t)desc_.height;
    points.resize((size_t)width * height);
    for (size_t i = 0; i < points.size(); ++i) {
      points[i].x = raw_[4 * i]; points[i].y = raw_[4 * i + 1]; points[i].z = raw_[4 * i + 2]; points[i].range = raw_[4 * i + 3];
    }
    valid_ = true;
  }
  static Affine3f poseOf(const Affine3f& p) { return p; }
  template <typename T> static Affine3f poseOf(const T&) { return identityPose(); }  // (any other type: identity)
  pfx_range_image_desc desc_ = {};
  std::vector<float> raw_;
  Affine3f pose_ = identityPose();
  bool valid_ = false;
};

class RangeImagePlanar : public RangeImage {
 public:
  // keypoints.h:212-216, tools.h:72-76
  template <typename CloudT, typename PoseT>
  void createFromPointCloudWithFixedSize(const CloudT& cloud, int di_width, int di_height, float di_center_x,
                                         float di_center_y, float di_focal_length_x, float di_focal_length_y,
                                         const PoseT& sensor_pose, CoordinateFrame frame = CAMERA_FRAME,
                                         float noise_level = 0.0f, float min_range = 0.0f) {
    pfx_ctx* c = b200::use(&cloud);  // the context bound to this cloud
    valid_ = false;
    pose_ = poseOf(sensor_pose);  // pcl::Affine3f (identity for the bundled clouds)
    if (c && !b200::ok(pfx_range_image_set_pose(c, pose_.data()), "RangeImagePlanar")) return;
    if (!c || frame != CAMERA_FRAME || noise_level != 0.0f) {
      std::fprintf(stderr, "[pcl::RangeImagePlanar] only CAMERA_FRAME with noise_level 0 is implemented\n");
      return;
    }
    if (!b200::ok(pfx_set_surface(c, cloud.points.data(), cloud.size(), sizeof(cloud.points[0]), PFX_HOST), "RangeImagePlanar")) return;
    if (!b200::ok(pfx_range_image_planar(c, di_width, di_height, di_center_x, di_center_y, di_focal_length_x, di_focal_length_y, min_range, &desc_), "RangeImagePlanar")) return;
    fetch(c);
  }
};

class RangeImageBorderExtractor {
 public:
  explicit RangeImageBorderExtractor(const RangeImage* range_image = nullptr) : range_image_(range_image) {}
  void setRangeImage(const RangeImage* range_image) { range_image_ = range_image; }
  const RangeImage* getRangeImagePtr() const { return range_image_; }

 private:
  const RangeImage* range_image_;
};

class NarfKeypoint {
 public:
  struct Parameters { float support_size = -1.0f; };
  explicit NarfKeypoint(RangeImageBorderExtractor* border_extractor = nullptr, float support_size = -1.0f)
      : border_extractor_(border_extractor) { parameters_.support_size = support_size; }
  void setRangeImageBorderExtractor(RangeImageBorderExtractor* b) { border_extractor_ = b; }
  void setRangeImage(const RangeImage* range_image) { if (border_extractor_) border_extractor_->setRangeImage(range_image); }
  Parameters& getParameters() { return parameters_; }
  // output: range-image pixel indices (y * width + x), ascending
  void compute(PointCloud<int>& output) {
    output.points.clear();
    output.width = output.height = 0;
    pfx_ctx* c = b200::ctx();
    const RangeImage* ri = border_extractor_ ? border_extractor_->getRangeImagePtr() : nullptr;
    if (!c || !ri) {
      std::fprintf(stderr, "[pcl::NarfKeypoint::detectKeypoints] RangeImageBorderExtractor member is NULL or has no range image\n");
      return;
    }
    if (!ri->install()) return;
    std::vector<int32_t> kp((size_t)ri->width * ri->height + 1);
    size_t n = 0;
    if (!b200::ok(pfx_narf_keypoints(c, parameters_.support_size, kp.data(), nullptr, nullptr, kp.size(), &n, nullptr, PFX_HOST), "NarfKeypoint")) return;
    output.points.assign(kp.begin(), kp.begin() + n);
    output.width = (uint32_t)n;
    output.height = 1;
  }

 private:
  RangeImageBorderExtractor* border_extractor_;
  Parameters parameters_;
};

class NarfDescriptor {
 public:
  struct Parameters { float support_size = -1.0f; bool rotation_invariant = true; };
  explicit NarfDescriptor(const RangeImage* range_image = nullptr, const std::vector<int>* indices = nullptr)
      : range_image_(range_image), indices_(indices) {}
  void setRangeImage(const RangeImage* range_image, const std::vector<int>* indices = nullptr) { range_image_ = range_image; indices_ = indices; }
  Parameters& getParameters() { return parameters_; }
  void compute(PointCloud<Narf36>& output) {
    output.points.clear();
    output.width = output.height = 0;
    pfx_ctx* c = b200::ctx();
    if (!c || !range_image_) {
      std::fprintf(stderr, "[pcl::NarfDescriptor::computeFeature] no range image given\n");
      return;
    }
    if (parameters_.support_size <= 0.0f) {
      std::fprintf(stderr, "[pcl::NarfDescriptor::computeFeature] support size is not set!\n");
      return;
    }
    if (!indices_ || indices_->empty() || !range_image_->install()) return;
    output.points.resize(indices_->size() * 8);
    size_t n = 0;
    if (!b200::ok(pfx_narf36(c, indices_->data(), indices_->size(), parameters_.support_size, parameters_.rotation_invariant ? 1 : 0,
                             output.points.data(), sizeof(Narf36), output.points.size(), &n, PFX_HOST), "NarfDescriptor")) n = 0;
    output.points.resize(n);
    output.width = (uint32_t)n;
    output.height = 1;
  }

 private:
  const RangeImage* range_image_;
  const std::vector<int>* indices_;
  Parameters parameters_;
};

// ------------------------------------------------------------------------------- registration
namespace registration {
// pcl::registration::CorrespondenceEstimation (included at evaluation.cpp:20; named by north_star)
template <typename FeatureT>
class CorrespondenceEstimation {
 public:
  void setInputSource(const typename PointCloud<FeatureT>::ConstPtr& s) { source_ = s; }
  void setInputTarget(const typename PointCloud<FeatureT>::ConstPtr& t) { target_ = t; }
  void determineCorrespondences(Correspondences& out, double max_distance = std::numeric_limits<double>::max()) { run(out, max_distance, 0); }
  void determineReciprocalCorrespondences(Correspondences& out, double max_distance = std::numeric_limits<double>::max()) { run(out, max_distance, 1); }

 private:
  void run(Correspondences& out, double max_distance, int reciprocal) {
    out.clear();
    pfx_ctx* c = b200::ctx();
    if (!c || !source_ || !target_) return;
    out.resize(source_->size());
    size_t n = 0;
    float md2 = (max_distance >= 1e18) ? -1.f : (float)(max_distance * max_distance);
    int rc = pfx_match(c, reinterpret_cast<const float*>(source_->points.data()), source_->size(), sizeof(FeatureT),
                       reinterpret_cast<const float*>(target_->points.data()), target_->size(), sizeof(FeatureT),
                       FeatureT::descriptorSize(), reciprocal, md2, reinterpret_cast<pfx_correspondence*>(out.data()),
                       out.size(), &n, PFX_HOST);
    if (!b200::ok(rc, "CorrespondenceEstimation")) n = 0;
    out.resize(n);
  }
  typename PointCloud<FeatureT>::ConstPtr source_, target_;
};

// pcl::registration::CorrespondenceRejectorSampleConsensus as driven at features.h:289-296
template <typename PointT>
class CorrespondenceRejectorSampleConsensus {
 public:
  void setInputSource(const typename PointCloud<PointT>::ConstPtr& c) { source_ = c; }
  void setInputTarget(const typename PointCloud<PointT>::ConstPtr& c) { target_ = c; }
  void setInputCorrespondences(const CorrespondencesConstPtr& c) { input_ = c; }
  void setInlierThreshold(double t) { inlier_threshold_ = t; }
  void setMaximumIterations(int n) { max_iterations_ = n; }
  void setSeed(uint64_t s) { seed_ = s; }  // PCL seeds its mt19937 with a constant; the contract here is explicit
  void getCorrespondences(Correspondences& out) {
    out.clear();
    for (int i = 0; i < 16; ++i) best_transformation_[i] = (i % 5 == 0) ? 1.f : 0.f;
    pfx_ctx* c = b200::ctx();
    if (!c || !source_ || !target_ || !input_) return;
    out.resize(input_->size());
    size_t n = 0;
    int rc = pfx_ransac_reject(c, source_->points.data(), source_->size(), sizeof(PointT), target_->points.data(),
                               target_->size(), sizeof(PointT), reinterpret_cast<const pfx_correspondence*>(input_->data()),
                               input_->size(), inlier_threshold_, max_iterations_, seed_,
                               reinterpret_cast<pfx_correspondence*>(out.data()), out.size(), &n, best_transformation_, nullptr,
                               nullptr, PFX_HOST);
    if (!b200::ok(rc, "CorrespondenceRejectorSampleConsensus")) n = 0;
    out.resize(n);
  }
  // row-major 4x4 (Eigen::Matrix4f in PCL)
  const float* getBestTransformation() const { return best_transformation_; }

 private:
  typename PointCloud<PointT>::ConstPtr source_, target_;
  CorrespondencesConstPtr input_;
  double inlier_threshold_ = 0.05;
  int max_iterations_ = 1000;
  uint64_t seed_ = 12345u;
  float best_transformation_[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
};
}  // namespace registration

// Eigen::Matrix4f stand-in for the transforms the registration classes hand back: m(row, col), row-major storage
struct Matrix4f {
  float m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  float operator()(int r, int c) const { return m[4 * r + c]; }
  float& operator()(int r, int c) { return m[4 * r + c]; }
  static Matrix4f Identity() { return Matrix4f(); }
};

// pcl::transformPointCloud(in, out, Matrix4f) as used at evaluation.cpp:257 (xyz moved, the other fields copied)
template <typename PointT>
inline void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Matrix4f& t) {
  if (&in != &out) out = in;
  for (auto& p : out.points) {
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    const float x = p.x, y = p.y, z = p.z;
    p.x = ((t(0, 0) * x + t(0, 1) * y) + t(0, 2) * z) + t(0, 3);
    p.y = ((t(1, 0) * x + t(1, 1) * y) + t(1, 2) * z) + t(1, 3);
    p.z = ((t(2, 0) * x + t(2, 1) * y) + t(2, 2) * z) + t(2, 3);
  }
}

// pcl::IterativeClosestPoint as driven by Evaluation::icpAlign (evaluation.cpp:863-885): the whole loop runs on
// the device (pfx_icp_align); align() fills the moved source cloud.
template <typename PointSource, typename PointTarget>
class IterativeClosestPoint {
 public:
  void setMaxCorrespondenceDistance(double d) { prm_.max_correspondence_distance = d; }
  void setRANSACOutlierRejectionThreshold(double t) { ransac_threshold_ = t; }  // stored; PCL 1.7's ICP never reads it
  void setTransformationEpsilon(double e) { prm_.transformation_epsilon = e; }
  void setEuclideanFitnessEpsilon(double e) { prm_.euclidean_fitness_epsilon = e; }
  void setMaximumIterations(int n) { prm_.max_iterations = n; }
  void setInputSource(const typename PointCloud<PointSource>::ConstPtr& c) { source_ = c; }
  void setInputTarget(const typename PointCloud<PointTarget>::ConstPtr& c) { target_ = c; }
  void align(PointCloud<PointSource>& output) { align(output, Matrix4f::Identity()); }
  void align(PointCloud<PointSource>& output, const Matrix4f& guess) {
    res_ = pfx_icp_result();
    final_ = guess;
    res_.fitness = std::numeric_limits<double>::max();
    pfx_ctx* c = b200::use(target_.get());  // the context bound to this cloud
    if (!c || !source_ || !target_) return;
    output = *source_;
    if (!b200::ok(pfx_set_surface(c, target_->points.data(), target_->size(), sizeof(PointTarget), PFX_HOST),
                  "IterativeClosestPoint"))
      return;
    int rc = pfx_icp_align(c, source_->points.data(), source_->size(), sizeof(PointSource), &prm_, guess.m, &res_,
                           output.points.data(), sizeof(PointSource), PFX_HOST);
    if (!b200::ok(rc, "IterativeClosestPoint")) return;
    std::memcpy(final_.m, res_.transform, sizeof(final_.m));
  }
  Matrix4f getFinalTransformation() const { return final_; }
  double getFitnessScore() const { return res_.fitness; }
  bool hasConverged() const { return res_.converged != 0; }
  int getNumberOfIterations() const { return res_.iterations; }

 private:
  typename PointCloud<PointSource>::ConstPtr source_;
  typename PointCloud<PointTarget>::ConstPtr target_;
  // PCL's defaults: corr distance sqrt(DBL_MAX), 10 iterations, epsilons 0 / -DBL_MAX
  pfx_icp_params prm_ = {1.3407807929942596e154, 10, 0.0, -std::numeric_limits<double>::max()};
  double ransac_threshold_ = 0.05;
  pfx_icp_result res_ = pfx_icp_result();
  Matrix4f final_;
};

}  // namespace pcl
