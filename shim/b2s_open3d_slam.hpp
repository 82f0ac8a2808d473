// b2s_open3d_slam.hpp -- the C++ subclasses a maintainer adds to open3d_slam to run the hot path on a B200 through the
// C ABI of include/b2s.h.  They implement the reference's own abstract interfaces
//     o3d_slam::CloudRegistration        (include/open3d_slam/CloudRegistration.hpp:19-27)
//     o3d_slam::ScanToMapRegistration    (include/open3d_slam/ScanToMapRegistration.hpp:29-38)
// and are selected from the reference's factories (src/CloudRegistration.cpp:85-100, src/ScanToMapRegistration.cpp:91-103)
// by one extra enum value each.  Host data stays in the reference's own layout
// (std::vector<Eigen::Vector3d>, 24-byte stride), which is exactly what the ABI takes.
#pragma once
#ifdef B2S_SHIM_STANDALONE_CHECK
#include "open3d_slam/interfaces.hpp"   // stand-in declarations (shim/stubs), type-check only
#else
#include "open3d_slam/CloudRegistration.hpp"
#include "open3d_slam/ScanToMapRegistration.hpp"
#include "open3d_slam/Submap.hpp"
#endif
#include <memory>
#include <mutex>
#include "b2s.h"

namespace o3d_slam {

// one engine handle per host thread (the reference calls registerClouds from up to three threads, SlamWrapper.cpp:228-231)
b2s_handle* b2sThreadHandle(const b2s_config& cfg);
b2s_config b2sConfigFrom(const IcpParameters& icp, const ScanProcessingParameters* scan, const MapBuilderParameters* mapBuilder);
[[noreturn]] void b2sThrow(int32_t code);   // status -> std::runtime_error, the reference's failure mode (assert.hpp:12-63)

class RegistrationIcpPointToPlaneB200 : public CloudRegistration {
 public:
  explicit RegistrationIcpPointToPlaneB200(const CloudRegistrationParameters& p);
  RegistrationResult registerClouds(const PointCloud& source, const PointCloud& target, const Transform& init) const final;
  void estimateNormalsOrCovariancesIfNeeded(PointCloud* cloud) const final;

 private:
  b2s_config cfg_;
};

// RegistrationIcpPointToPoint (src/CloudRegistration.cpp:69-82) on the device: same ICP loop, Eigen::umeyama updates
class RegistrationIcpPointToPointB200 : public CloudRegistration {
 public:
  explicit RegistrationIcpPointToPointB200(const CloudRegistrationParameters& p);
  RegistrationResult registerClouds(const PointCloud& source, const PointCloud& target, const Transform& init) const final;

 private:
  b2s_config cfg_;
};

// RegistrationIcpGeneralized (src/CloudRegistration.cpp:15-38) on the device: covariances from the normals the reference's own
// estimateNormalsOrCovariancesIfNeeded leaves on the clouds ([O3D] InitializePointCloudForGeneralizedICP, normals branch)
class RegistrationIcpGeneralizedB200 : public CloudRegistration {
 public:
  explicit RegistrationIcpGeneralizedB200(const CloudRegistrationParameters& p);
  RegistrationResult registerClouds(const PointCloud& source, const PointCloud& target, const Transform& init) const final;
  void estimateNormalsOrCovariancesIfNeeded(PointCloud* cloud) const final;

 private:
  b2s_config cfg_;
};

// Submap::carve for the sparse map (src/Submap.cpp:109-123): removes the carved points from *map in place.  cropperPose is
// the pose mapBuilderCropper_ currently holds (the previous insertion); the caller keeps the every-N-scans schedule.
void carveB200(const PointCloud& rawScan, const Transform& mapToRangeSensor, const Transform& cropperPose, const MapBuilderParameters& p,
               PointCloud* map);

// The map side of o3d_slam::Submap, device-resident: what a maintainer puts behind Submap's own methods (one member,
// `std::unique_ptr<SubmapB200> device_`).  Same method names and argument meaning as the methods it replaces:
//     Submap::insertScan            src/Submap.cpp:39-75     (transform, carve every N insertions, append, voxelize within the cropper)
//     Submap::insertScanDenseMap    src/Submap.cpp:77-92
//     Submap::transform             src/Submap.cpp:94-107
//     Submap::getMapPointCloud      src/Submap.cpp:184-186   (download on demand, cached until the next change)
//     Submap::isEmpty               src/Submap.cpp:221-223
// The b2s_submap lives on the handle of the thread that created the SubmapB200 (the mapping thread).
class SubmapB200 {
 public:
  SubmapB200(const MapperParameters& p, size_t capacityPoints = 2000000);
  ~SubmapB200();
  SubmapB200(const SubmapB200&) = delete;
  SubmapB200& operator=(const SubmapB200&) = delete;
  bool insertScan(const PointCloud& rawScan, const PointCloud& preProcessedScan, const Transform& mapToRangeSensor, bool isPerformCarving);
  bool insertScanDenseMap(const PointCloud& rawScan, const Transform& mapToRangeSensor, bool isPerformCarving);
  void transform(const Transform& T);
  const PointCloud& getMapPointCloud() const;
  bool isEmpty() const;
  void setMapPointCloud(const PointCloud& cloud);           // initial map (SlamWrapper::setInitialMap)
  b2s_submap* handle() const { return sm_; }
  b2s_handle* engine() const { return h_; }

 private:
  b2s_config cfg_;
  MapBuilderParameters mapBuilder_;
  MapBuilderParameters denseMapBuilder_;
  b2s_handle* h_ = nullptr;
  b2s_submap* sm_ = nullptr;
  size_t nScansInsertedMap_ = 0, nScansInsertedDenseMap_ = 0;
  Transform cropperPose_ = Transform::Identity();           // mapBuilderCropper_'s pose: set after every insertion (Submap.cpp:71)
  mutable PointCloud cache_;
  mutable bool cacheValid_ = false;
};

class ScanToMapIcpB200 : public ScanToMapRegistration {
 public:
  explicit ScanToMapIcpB200(const MapperParameters& p);
  // device-resident variant: no upload of the map, the patch crop and the index build run on the resident cloud
  RegistrationResult scanToMapRegistration(const PointCloud& scan, const SubmapB200& activeSubmap, const Transform& mapToRangeSensor,
                                           const Transform& initialGuess) const;
  ProcessedScans processForScanMatchingAndMerging(const PointCloud& in, const Transform& mapToRangeSensor) const final;
  RegistrationResult scanToMapRegistration(const PointCloud& scan, const Submap& activeSubmap, const Transform& mapToRangeSensor,
                                           const Transform& initialGuess) const final;
  bool isMergeScanValid(const PointCloud& in) const final { return in.HasNormals(); }
  void prepareInitialMap(PointCloud* map) const final;

 private:
  b2s_config cfg_;
};

}  // namespace o3d_slam
