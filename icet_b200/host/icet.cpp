// icet_b200/host/icet.cpp -- host side of the drop-in `class ICET` (include/icet.h): the constructor forwards to
// the C ABI (include/icet_b200.h) and fills the public members the reference's callers read.
// Replaces reference src/icet.cpp:29-66 (constructor / destructor); everything else of src/icet.cpp runs in
// the CUDA kernels of icet_b200/csrc/icet_b200.cu.
#include "icet.h"

#include <cstdint>
#include <cstring>
#include <memory>

bool ICET::fillVisualization = true;
int ICET::device = 0;
bool ICET::shippedRowOrder = false;

namespace {

struct CtxHolder {  // one context per host thread ("run multiple ICETs at once", reference include/icet.h:43)
  icet_b200_ctx* ctx = nullptr;
  ~CtxHolder() {
    if (ctx) icet_b200_destroy(ctx);
  }
};

icet_b200_ctx* thread_ctx() {
  thread_local CtxHolder h;
  if (!h.ctx) {
    if (icet_b200_create(ICET::device, &h.ctx) != 0)
      throw std::runtime_error(std::string("ICET: ") + icet_b200_last_error());
  }
  return h.ctx;
}

void check(int rc) {
  if (rc < 0) throw std::runtime_error(std::string("ICET: ") + icet_b200_last_error());
}

}  // namespace

ICET::ICET(Eigen::MatrixXf& scan1, Eigen::MatrixXf& scan2, int runlen, Eigen::VectorXf X0, int num_bins_phi,
           int num_bins_theta, int n_, float thresh_, float buff_)
    : rl(runlen), numBinsPhi(num_bins_phi), numBinsTheta(num_bins_theta), n(n_), thresh(thresh_), buff(buff_),
      points1(scan1), points2(scan2), X(X0), status(0), voxelsUsed(0), axesDropped(0) {
  if (scan1.cols() != 3 || scan2.cols() != 3) throw std::runtime_error("ICET: clouds must be N x 3");
  if (X0.size() != 6) throw std::runtime_error("ICET: X0 must have 6 entries");
  icet_b200_ctx* ctx = thread_ctx();
  icet_b200_params p;
  std::memset(&p, 0, sizeof(p));
  p.runlen = runlen;
  p.bins_phi = num_bins_phi;
  p.bins_theta = num_bins_theta;
  p.n = n_;
  p.thresh = thresh_;
  p.buff = buff_;
  p.flags = shippedRowOrder ? ICET_B200_FLAG_SHIPPED_ORDER : 0;
  float x0[6];
  for (int k = 0; k < 6; k++) x0[k] = X0[k];
  icet_b200_result res;
  check(icet_b200_set_dump(ctx, fillVisualization ? 1 : 0));
  // Eigen::MatrixXf is column-major: data() is exactly the x | y | z plane layout the C ABI takes
  const int n1 = (int)scan1.rows(), n2 = (int)scan2.rows();
  check(icet_b200_register(ctx, &p, scan1.data(), n1, n1, scan2.data(), n2, n2, x0, &res));

  pred_stds = Eigen::VectorXf(6);
  Q = Eigen::MatrixXf(6, 6);
  for (int k = 0; k < 6; k++) {
    X[k] = res.X[k];
    pred_stds[k] = res.pred_stds[k];
    for (int j = 0; j < 6; j++) Q(k, j) = res.Q[k * 6 + j];
  }
  status = res.status;
  voxelsUsed = res.n_used;
  axesDropped = res.n_dropped;

  const int ncell = num_bins_phi * num_bins_theta;
  clusterBounds = Eigen::MatrixXf::Zero(ncell, 6);
  testPoints = Eigen::MatrixXf::Zero((long)6 * ncell, 3);
  HTWH_i = Eigen::MatrixXf::Zero(6, 6);
  HTWdz_i = Eigen::MatrixXf::Zero(6, 1);
  dx = Eigen::VectorXf(6);
  dx.setZero();
  if (!fillVisualization) return;

  // members only the visualisation reads
  if (n2 > 0) check(icet_b200_get_points2(ctx, points2.data(), n2));
  std::vector<float> bounds((size_t)ncell * 6), mu((size_t)ncell * 3), sig((size_t)ncell * 9), vec((size_t)ncell * 9);
  std::vector<uint8_t> has(ncell), lm((size_t)ncell * 3);
  const size_t rlz = (size_t)(runlen > 0 ? runlen : 1);
  std::vector<float> tp((size_t)ncell * 18), hh(rlz * 36), hz(rlz * 6), xit(rlz * 6), mu2v(rlz * ncell * 3), sg2v(rlz * ncell * 9);
  std::vector<uint8_t> used(rlz * ncell);
  icet_b200_voxel_dump d;
  std::memset(&d, 0, sizeof(d));
  d.testPoints = tp.data();
  d.HTWH = hh.data();
  d.HTWdz = hz.data();
  d.Xit = xit.data();
  d.mu2 = mu2v.data();
  d.sigma2 = sg2v.data();
  d.used2 = used.data();
  d.bounds = bounds.data();
  d.mu1 = mu.data();
  d.sigma1 = sig.data();
  d.evec1 = vec.data();
  d.has1 = has.data();
  d.lmask = lm.data();
  check(icet_b200_get_dump(ctx, &d));
  for (int c = 0; c < ncell; c++)
    for (int k = 0; k < 6; k++) clusterBounds(c, k) = bounds[(size_t)c * 6 + k];
  for (long r = 0; r < (long)6 * ncell; r++)
    for (int k = 0; k < 3; k++) testPoints(r, k) = tp[(size_t)r * 3 + k];
  if (runlen > 0) {
    const size_t last = (size_t)runlen - 1;
    for (int i = 0; i < 6; i++) {
      HTWdz_i(i, 0) = hz[last * 6 + i];
      dx[i] = xit[last * 6 + i] - (runlen > 1 ? xit[(last - 1) * 6 + i] : x0[i]);
      for (int j = 0; j < 6; j++) HTWH_i(i, j) = hh[last * 36 + 6 * i + j];
    }
    for (int phi = 0; phi < num_bins_phi; phi++)
      for (int theta = 0; theta < num_bins_theta; theta++) {
        const size_t c = (size_t)num_bins_theta * phi + theta;
        if (!used[last * ncell + c]) continue;
        Eigen::Vector3f m2;
        CovarianceMatrix S2;
        for (int i = 0; i < 3; i++) {
          m2[i] = mu2v[(last * ncell + c) * 3 + i];
          for (int j = 0; j < 3; j++) S2(i, j) = sg2v[(last * ncell + c) * 9 + 3 * i + j];
        }
        mu2[theta][phi] = m2;
        sigma2[theta][phi] = S2;
      }
  }
  // the reference visits cells phi-major, theta-minor (src/icet.cpp:95-101): same order for the viz vectors
  for (int phi = 0; phi < num_bins_phi; phi++)
    for (int theta = 0; theta < num_bins_theta; theta++) {
      const int c = num_bins_theta * phi + theta;
      if (!has[c]) continue;
      Eigen::Vector3f m;
      CovarianceMatrix S, Um, Lm;
      for (int i = 0; i < 3; i++) {
        m[i] = mu[(size_t)c * 3 + i];
        for (int j = 0; j < 3; j++) {
          S(i, j) = sig[(size_t)c * 9 + 3 * i + j];
          Um(i, j) = vec[(size_t)c * 9 + 3 * j + i];  // U = eigenvectors.transpose() (src/icet.cpp:184)
          Lm(i, j) = (i == j && lm[(size_t)c * 3 + i]) ? 1.f : 0.f;
        }
      }
      sigma1[theta][phi] = S;
      mu1[theta][phi] = m;
      U[theta][phi] = Um;
      L[theta][phi] = Lm;
      ellipsoid1Means.push_back(m);
      ellipsoid1Covariances.push_back(S);
      ellipsoid1Alphas.push_back(0.3f);
    }
}

ICET::~ICET() {}
