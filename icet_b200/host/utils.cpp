// icet_b200/host/utils.cpp -- host helpers of include/utils.h (replaces reference src/utils.cpp for the callers;
// written from the behaviour documented there, no csv.hpp dependency).
#include "utils.h"

#include <cmath>
#include <fstream>
#include <iostream>
#include <sstream>
#include <vector>

namespace utils {

namespace {
std::vector<std::string> split(const std::string& line, char delim) {
  std::vector<std::string> out;
  std::string cur;
  bool quoted = false;
  for (char ch : line) {
    if (ch == '"') quoted = !quoted;
    else if (ch == delim && !quoted) { out.push_back(cur); cur.clear(); }
    else if (ch != '\r') cur.push_back(ch);
  }
  out.push_back(cur);
  return out;
}
}  // namespace

// reference src/utils.cpp:12-91
Eigen::MatrixXf loadPointCloudCSV(std::string filename, std::string datasetType) {
  std::ifstream file(filename);
  if (!file.is_open()) std::cerr << "Error: Could not open the CSV file." << std::endl;
  std::vector<float> xyz;
  std::string line;
  if (datasetType == "ouster") {
    // Integer millimetres in columns 8, 9, 10.  FOUR lines are consumed before the first point, like in the reference:
    // csv-parser with header_row(1) drops lines 0 and 1, then the two read_row() calls "skip the first / second row"
    // (src/utils.cpp:20-30) -- i.e. the first two data rows are dropped as well (checked against the reference's own
    // loader compiled from its sources: tests/test_host.py::test_utils_loader_formats).
    for (int k = 0; k < 4; k++) std::getline(file, line);
    while (std::getline(file, line)) {
      if (line.empty()) continue;
      std::vector<std::string> f = split(line, ',');
      if (f.size() < 11) continue;
      for (int k = 8; k <= 10; k++) xyz.push_back(static_cast<float>(std::stoi(f[k])) / 1000);
    }
  } else {
    // tab-separated x y z; csv-parser takes line 0 as the header, so the reference never sees the first point (:64-78)
    std::getline(file, line);
    while (std::getline(file, line)) {
      if (line.empty()) continue;
      std::vector<std::string> f = split(line, '\t');
      if (f.size() < 3) continue;
      for (int k = 0; k < 3; k++) xyz.push_back(std::stof(f[k]));
    }
  }
  const int rows = (int)(xyz.size() / 3);
  Eigen::MatrixXf m(rows, 3);
  for (int i = 0; i < rows; i++)
    for (int k = 0; k < 3; k++) m(i, k) = xyz[(size_t)3 * i + k];
  return m;
}

// reference src/utils.cpp:93-119
Eigen::MatrixXf cartesianToSpherical(const Eigen::MatrixXf& c) {
  Eigen::MatrixXf s(c.rows(), 3);
  for (int i = 0; i < c.rows(); ++i) {
    const float x = c(i, 0), y = c(i, 1), z = c(i, 2);
    float r = std::sqrt(x * x + y * y + z * z);
    float th = std::atan2(y, x);
    if (th < 0.0) th += 2.0 * M_PI;
    float ph = std::acos(z / r);
    s(i, 0) = std::isnan(r) ? 1000.0f : r;
    s(i, 1) = std::isnan(th) ? 1000.0f : th;
    s(i, 2) = std::isnan(ph) ? 1000.0f : ph;
  }
  return s;
}

// reference src/utils.cpp:121-142
Eigen::MatrixXf sphericalToCartesian(const Eigen::MatrixXf& s) {
  Eigen::MatrixXf c(s.rows(), 3);
  for (int i = 0; i < s.rows(); ++i) {
    const float r = s(i, 0), th = s(i, 1), ph = s(i, 2);
    c(i, 0) = r * std::sin(ph) * std::cos(th);
    c(i, 1) = r * std::sin(ph) * std::sin(th);
    c(i, 2) = r * std::cos(ph);
  }
  return c;
}

// reference src/utils.cpp:144-152
Eigen::Matrix3f R(float phi, float theta, float psi) {
  using std::cos;
  using std::sin;
  Eigen::Matrix3f m;
  m(0, 0) = cos(theta) * cos(psi);
  m(0, 1) = sin(psi) * cos(phi) + sin(phi) * sin(theta) * cos(psi);
  m(0, 2) = sin(phi) * sin(psi) - sin(theta) * cos(phi) * cos(psi);
  m(1, 0) = -sin(psi) * cos(theta);
  m(1, 1) = cos(phi) * cos(psi) - sin(phi) * sin(theta) * sin(psi);
  m(1, 2) = sin(phi) * cos(psi) + sin(theta) * sin(psi) * cos(phi);
  m(2, 0) = sin(theta);
  m(2, 1) = -sin(phi) * cos(theta);
  m(2, 2) = cos(phi) * cos(theta);
  return m;
}

}  // namespace utils
