/*
 * include/icet.h -- drop-in replacement of the reference's include/icet.h (class ICET, reference
 * include/icet.h:36-116).  Same constructor signature, same public data members the callers read
 * (icet_cpp_demo.cpp:39,48-57; odometry.cpp:77-79; scanMatcher.cpp:64; simpleMapMaker.cpp:120-122).
 * The constructor does all the work, exactly like the reference (src/icet.cpp:29-63) -- but on the GPU,
 * through the C ABI of include/icet_b200.h (hand-written sm_100a kernels, no CPU fallback).
 *
 * Differences a caller can observe:
 *   - the CPU-only helper methods (fitScan1, fitCells1, findCluster, ... include/icet.h:44-68) and the 4-thread
 *     ThreadPool member (:113) do not exist: their work happens inside the kernels; the per-point scratch members no
 *     caller reads (points1Spherical, points2Spherical, points2_OG, pointIndices1/2, include/icet.h:79-82, :95-96) are
 *     not materialised on the host (two more N x 3 downloads and 1800 index lists per registration);
 *   - DEFAULT ROW ORDER: the voxels of scan 1 are clustered in ascending range order -- what the reference's comments
 *     intend -- whereas an unmodified reference build walks them in the order its broken permutation loop leaves
 *     (src/icet.cpp:78-83) and finds several times fewer clusters (89 instead of 336 Gaussians on the bundled pair), so
 *     X and pred_stds DIFFER from an upstream build unless ICET::shippedRowOrder = true (bit-identical voxels then);
 *   - `points2` is returned in the caller's row order (the reference leaves it in its internal radial order);
 *   - new member `Q`: the 6x6 error-bound covariance the reference computes and drops (src/icet.cpp:410-411);
 *   - a failed registration throws std::runtime_error (scanMatcher.cpp:98-104 already catches std::exception).
 */
#ifndef ICET_H
#define ICET_H

#include <Eigen/Dense>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "icet_b200.h"
#include "utils.h"

using CovarianceMatrix = Eigen::Matrix<float, 3, 3>;
using CovarianceMap = std::map<int, std::map<int, CovarianceMatrix>>;
using MeanMap = std::map<int, std::map<int, Eigen::Vector3f>>;

class ICET {
 public:
  ICET(Eigen::MatrixXf& scan1, Eigen::MatrixXf& scan2, int runlen, Eigen::VectorXf X0, int num_bins_phi,
       int num_bins_theta, int n = 25, float thresh = 0.1, float buff = 0.1);
  ~ICET();

  // algorithm params (reference include/icet.h:71-76)
  int rl;
  int numBinsPhi;
  int numBinsTheta;
  int n;
  float thresh;
  float buff;

  // reference include/icet.h:78-87
  Eigen::MatrixXf points1;
  Eigen::MatrixXf points2;        // scan 2 as transformed by the last iteration (src/icet.cpp:375-378)
  Eigen::MatrixXf clusterBounds;  // (numBinsPhi*numBinsTheta) x 6, row = numBinsTheta*phi + theta
  Eigen::MatrixXf testPoints;     // (6*numBinsPhi*numBinsTheta) x 3: the 2-sigma test points of the axes found extended
                                  // (src/icet.cpp:214-232); rows the reference never writes are 0 here
  Eigen::MatrixXf HTWH_i;         // 6 x 6 H^T W H of the last iteration (src/icet.cpp:383, :401)
  Eigen::MatrixXf HTWdz_i;        // 6 x 1 H^T W dz of the last iteration (:385, :402)
  Eigen::VectorXf pred_stds;

  // per-voxel scan-1 state, keyed [theta][phi] like the reference (include/icet.h:89-94)
  CovarianceMap sigma1;
  MeanMap mu1;
  CovarianceMap L;
  CovarianceMap U;
  // scan-2 Gaussians of the LAST iteration for the voxels that contributed.  (The reference declares these maps but
  // never fills them -- the assignments at src/icet.cpp:308-310 are commented out; here they are filled.)
  CovarianceMap sigma2;
  MeanMap mu2;

  Eigen::VectorXf X;   // solution vector [x y z phi theta psi]
  Eigen::VectorXf dx;  // the last update of X (src/icet.cpp:430-433)
  Eigen::MatrixXf Q;   // 6x6 pinv(H^T W H) of the last iteration (not in the reference)

  // for viz (reference include/icet.h:102-107); ellipsoid2* stay empty like in the reference (src/icet.cpp:49-61)
  std::vector<Eigen::Vector3f> ellipsoid1Means;
  std::vector<Eigen::Matrix3f> ellipsoid1Covariances;
  std::vector<float> ellipsoid1Alphas;
  std::vector<Eigen::Vector3f> ellipsoid2Means;
  std::vector<Eigen::Matrix3f> ellipsoid2Covariances;
  std::vector<float> ellipsoid2Alphas;

  // diagnostics of the GPU path
  int status;       // ICET_B200_OK or ICET_B200_COND_OVERFLOW
  int voxelsUsed;   // voxels that contributed to H^T W H in the last iteration
  int axesDropped;  // solution axes dropped by the condition check (src/icet.cpp:469-486)

  // When true (default, faithful drop-in) the constructor also downloads the members only the visualisation
  // reads (points2, clusterBounds, sigma1/mu1/L/U, ellipsoid1*).  Odometry-style callers that read only X and
  // pred_stds can set it to false to skip those device-to-host copies.
  static bool fillVisualization;
  // CUDA device used by the calling thread's context (default 0).  Set before the first construction.
  static int device;
  // false (default): findCluster sees each cell's points in ascending range order, which is what the reference's
  // comments intend (src/icet.cpp:71).  true: in the row order the reference's permutation loop really leaves
  // (src/icet.cpp:78-83 is not a valid permutation application) -- reproduces an unmodified reference build
  // (several times fewer voxels get a Gaussian), at the price of one host round trip (ICET_B200_FLAG_SHIPPED_ORDER).
  static bool shippedRowOrder;
};

#endif
