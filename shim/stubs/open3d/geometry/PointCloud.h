// STAND-IN for open3d::geometry::PointCloud (Open3D is absent in this image): only the members the shim touches.
#pragma once
#include <Eigen/Dense>
#include <vector>
namespace open3d { namespace geometry {
class PointCloud {
 public:
  std::vector<Eigen::Vector3d> points_, normals_, colors_;
  bool HasNormals() const { return !points_.empty() && normals_.size() == points_.size(); }
  bool HasCovariances() const { return false; }
  bool IsEmpty() const { return points_.empty(); }
};
}}  // namespace open3d::geometry
