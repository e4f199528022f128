// examples/nodes_headless.cpp -- replays a recorded sequence through the ROS-free node classes
// (include/icet_nodes.h): what `rosbag play` + odometry_node / map_maker_node do in the reference
// (src/odometry.cpp, src/simpleMapMaker.cpp), printing what they would publish.
//
// usage: nodes_headless odometry|map|match points_per_scan file.f32 [map_size downsample]
//   file.f32: consecutive scans, each the x | y | z planes of points_per_scan float32 values.
#include <cstdio>
#include <fstream>
#include <string>

#include "icet_nodes.h"

int main(int argc, char** argv) {
  if (argc < 4) {
    std::fprintf(stderr, "usage: %s odometry|map points_per_scan file.f32 [map_size downsample]\n", argv[0]);
    return 2;
  }
  const std::string mode = argv[1];
  const long n = std::stol(argv[2]);
  std::ifstream f(argv[3], std::ios::binary | std::ios::ate);
  if (!f) { std::fprintf(stderr, "cannot open %s\n", argv[3]); return 2; }
  const long nscans = (long)f.tellg() / (12 * n);
  f.seekg(0);
  try {
    OdometryNode* odo = nullptr;
    MapMakerNode* mm = nullptr;
    ScanMatcherNode* sm = nullptr;
    if (mode == "match") sm = new ScanMatcherNode((int)n);
    else if (mode == "map") mm = new MapMakerNode((int)n, argc > 4 ? std::stoi(argv[4]) : 600000, argc > 5 ? std::stoi(argv[5]) : 2000);
    else odo = new OdometryNode((int)n);
    if (sm) { delete odo; odo = nullptr; }
    for (long k = 0; k < nscans; k++) {
      Eigen::MatrixXf cloud(n, 3);
      f.read(reinterpret_cast<char*>(cloud.data()), n * 12);
      NodeOutput o;
      const bool got = sm ? sm->pointcloudCallback(cloud, &o) : (mm ? mm->pointcloudCallback(cloud, &o) : odo->pointcloudCallback(cloud, &o));
      if (!got) continue;
      std::printf("POSE %ld X [%.9g, %.9g, %.9g, %.9g, %.9g, %.9g] pos [%.9g, %.9g, %.9g] quat [%.9g, %.9g, %.9g, %.9g] "
                  "points %d guarded %d\n", k, o.X[0], o.X[1], o.X[2], o.X[3], o.X[4], o.X[5], o.position[0],
                  o.position[1], o.position[2], o.orientation[0], o.orientation[1], o.orientation[2], o.orientation[3],
                  o.points, (int)o.guarded);
    }
    if (mm) {
      Eigen::MatrixXf m = mm->getQueue();
      double sx = 0, sy = 0, sz = 0;
      for (long i = 0; i < m.rows(); i++) { sx += m(i, 0); sy += m(i, 1); sz += m(i, 2); }
      std::printf("MAP rows %ld sum [%.9g, %.9g, %.9g] sample0 %d\n", (long)m.rows(), sx, sy, sz,
                  mm->lastSample.empty() ? -1 : mm->lastSample[0]);
    }
    if (sm) {
      const Eigen::MatrixXf& a = sm->scan2_in_scan1_frame;
      std::printf("ALIGNED rows %ld first [%.9g, %.9g, %.9g] trail %ld\n", (long)a.rows(), a(0, 0), a(0, 1), a(0, 2), (long)sm->snailTrail.rows());
    }
    delete sm;
    delete mm;
    delete odo;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "node failed: %s\n", e.what());
    return 1;
  }
  return 0;
}
