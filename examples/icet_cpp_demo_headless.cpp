// examples/icet_cpp_demo_headless.cpp -- the reference's src/icet_cpp_demo.cpp without the OpenGL part
// (its visualization.h is not in the reference repository): load two clouds, run
//     ICET it(scan1, scan2, run_length, X0, numBinsPhi, numBinsTheta);        (icet_cpp_demo.cpp:31-38)
// print the solution, the 1-sigma bounds and the timing, plus the sizes of the members the demo hands to its
// visualisation (icet_cpp_demo.cpp:48-57).
//
// usage: icet_cpp_demo_headless scan1 scan2 [ouster|txt|f32] [x0_x] [shipped]
//   "shipped": ICET::shippedRowOrder = true -- cluster in the row order an unmodified reference build ends up with
//   "f32": raw float32 file holding the x | y | z planes (column-major N x 3), the tests' exchange format.
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "icet.h"
#include "utils.h"

static Eigen::MatrixXf load_f32(const std::string& path) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) throw std::runtime_error("cannot open " + path);
  const std::streamsize bytes = f.tellg();
  f.seekg(0);
  const long n = (long)(bytes / 12);
  Eigen::MatrixXf m(n, 3);
  f.read(reinterpret_cast<char*>(m.data()), n * 12);
  return m;
}

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s scan1 scan2 [ouster|txt|f32] [x0_x]\n", argv[0]);
    return 2;
  }
  const std::string type = argc > 3 ? argv[3] : "ouster";
  int run_length = 7;
  int numBinsPhi = 24;
  int numBinsTheta = 75;
  Eigen::VectorXf X0;
  X0.resize(6);
  X0 << (argc > 4 ? std::stof(argv[4]) : 1.f), 0., 0., 0., 0., 0.;  // the demo's initial estimate

  if (argc > 5 && std::string(argv[5]) == "shipped") ICET::shippedRowOrder = true;
  try {
    Eigen::MatrixXf scan1 = type == "f32" ? load_f32(argv[1]) : utils::loadPointCloudCSV(argv[1], type);
    Eigen::MatrixXf scan2 = type == "f32" ? load_f32(argv[2]) : utils::loadPointCloudCSV(argv[2], type);
    ICET warm(scan1, scan2, run_length, X0, numBinsPhi, numBinsTheta);  // context creation + first launch
    auto before = std::chrono::steady_clock::now();
    ICET it(scan1, scan2, run_length, X0, numBinsPhi, numBinsTheta);
    auto ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - before).count();
    Eigen::VectorXf X = it.X;
    std::cout << "soln: " << std::endl << X << std::endl;
    std::cout << "1-sigma error bounds:" << std::endl << it.pred_stds << std::endl;
    std::cout << "Took: " << ms << " ms to register scans using ICET" << std::endl;
    Eigen::Matrix3f rot = utils::R(X[3], X[4], X[5]);
    std::cout << "R(0,0) " << rot(0, 0) << " points1 " << it.points1.rows() << " points2 " << it.points2.rows()
              << " clusterBounds " << it.clusterBounds.rows() << "x" << it.clusterBounds.cols() << " ellipsoids "
              << it.ellipsoid1Means.size() << " used " << it.voxelsUsed << std::endl;
    std::printf("X_JSON [%.9g, %.9g, %.9g, %.9g, %.9g, %.9g]\n", X[0], X[1], X[2], X[3], X[4], X[5]);
    std::printf("P2_JSON [%.9g, %.9g, %.9g]\n", it.points2(7, 0), it.points2(7, 1), it.points2(7, 2));
  } catch (const std::exception& e) {
    std::fprintf(stderr, "ICET failed: %s\n", e.what());
    return 1;
  }
  return 0;
}
