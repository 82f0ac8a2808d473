// STAND-IN for open3d::pipelines::registration::RegistrationResult: the three fields the callers read.
#pragma once
#include <Eigen/Dense>
namespace open3d { namespace pipelines { namespace registration {
class RegistrationResult {
 public:
  Eigen::Matrix4d transformation_ = Eigen::Matrix4d::Identity();
  double fitness_ = 0.0;
  double inlier_rmse_ = 0.0;
};
}}}  // namespace
