/*
 * include/utils.h -- drop-in replacement of the reference's include/utils.h (reference include/utils.h:1-19).
 * Host-side helpers the callers use next to `class ICET` (odometry.cpp:86, scanMatcher.cpp:73,
 * simpleMapMaker.cpp:141 call utils::R; icet_cpp_demo.cpp:25-26 calls utils::loadPointCloudCSV).
 * The registration itself never calls these: its geometry lives in the CUDA kernels.
 */
#ifndef UTILS_H
#define UTILS_H

#include <Eigen/Dense>
#include <string>

namespace utils {

// "ouster": CSV with two header rows, integer millimetres in columns 8-10 (reference src/utils.cpp:19-61);
// anything else: tab-separated x y z in metres (reference src/utils.cpp:64-88).
Eigen::MatrixXf loadPointCloudCSV(std::string filename, std::string datasetType = "csv");

// r, theta in [0, 2 pi), phi = acos(z / r); NaN -> 1000 (reference src/utils.cpp:93-119)
Eigen::MatrixXf cartesianToSpherical(const Eigen::MatrixXf& cartesianPoints);

// reference src/utils.cpp:121-142
Eigen::MatrixXf sphericalToCartesian(const Eigen::MatrixXf& sphericalPoints);

// body-frame xyz Euler angles -> rotation matrix (reference src/utils.cpp:144-152)
Eigen::Matrix3f R(float phi, float theta, float psi);

}  // namespace utils

#endif
