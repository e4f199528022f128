/*
 * oracle/icet_oracle.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU oracle: a plain C++17 restatement (no Eigen, no reference sources) of the
 * reference registration path, i.e. everything that runs inside
 *     ICET::ICET(...)                      reference src/icet.cpp:29-63
 * with the helper routines of src/utils.cpp:93-152.  Each function below cites
 * the reference file:line it follows.
 *
 * PARITY UNPINNED.  The reference has no tests or golden vectors for this path
 * and cannot be built in this image (Eigen3 headers absent, no network), so the
 * oracle cannot be checked against reference outputs.  The Eigen routines the
 * reference calls are restated from the published Eigen 3.3.7 algorithms
 * (Ubuntu 20.04 / ROS Noetic, the reference's documented platform):
 *   - SelfAdjointEigenSolver<Matrix3f>::compute   (fixed 3x3 tridiagonalisation +
 *     implicit symmetric QR with Wilkinson shift + selection sort)      -> eig3()
 *   - SelfAdjointEigenSolver<MatrixXf>::compute   (Householder tridiagonalisation
 *     + the same QR)                                                    -> eigsym()
 *   - CompleteOrthogonalDecomposition<MatrixXf>::pseudoInverse()  (column-pivoted
 *     Householder QR, rank by |R_ii| > eps*min(m,n)*max|R_ii|, Z reflectors for
 *     the rank-deficient case)                                          -> cod_pinv()
 * Build with -ffp-contract=off so that results do not depend on the host's FMA
 * support (oracle/Makefile).
 */
#include "icet_oracle.h"
#include "eigen_algos.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <thread>
#include <vector>

namespace {

using namespace eigen_restated;

// ---------------------------------------------------------------------------------------------
// point-wise geometry, fp32 (src/utils.cpp:93-152)
// ---------------------------------------------------------------------------------------------
// utils::cartesianToSpherical, src/utils.cpp:93-119
inline void c2s(float x, float y, float z, float& r, float& th, float& ph) {
  float s = x * x;
  s = s + y * y;
  s = s + z * z;
  r = std::sqrt(s);                                    // rowwise().norm()            :98
  th = std::atan2(y, x);                               //                             :104
  if (th < 0.0) th = (float)((double)th + 2.0 * M_PI); // theta(i) += 2.0 * M_PI      :105-107
  ph = std::acos(z / r);                               //                             :108
  if (std::isnan(r)) r = 1000.0f;                      // isNaN().select(1000.0, .)   :116
  if (std::isnan(th)) th = 1000.0f;
  if (std::isnan(ph)) ph = 1000.0f;
}
// utils::sphericalToCartesian, src/utils.cpp:121-142
inline void s2c(float r, float th, float ph, float& x, float& y, float& z) {
  x = r * std::sin(ph) * std::cos(th);
  y = r * std::sin(ph) * std::sin(th);
  z = r * std::cos(ph);
}
// utils::R, src/utils.cpp:144-152 (row-major 3x3)
inline void rotR(float phi, float theta, float psi, float R[9]) {
  using std::cos;
  using std::sin;
  R[0] = cos(theta) * cos(psi);
  R[1] = sin(psi) * cos(phi) + sin(phi) * sin(theta) * cos(psi);
  R[2] = sin(phi) * sin(psi) - sin(theta) * cos(phi) * cos(psi);
  R[3] = -sin(psi) * cos(theta);
  R[4] = cos(phi) * cos(psi) - sin(phi) * sin(theta) * sin(psi);
  R[5] = sin(phi) * cos(psi) + sin(theta) * sin(psi) * cos(phi);
  R[6] = sin(theta);
  R[7] = -sin(phi) * cos(theta);
  R[8] = cos(phi) * cos(theta);
}
// ICET::get_H, src/icet.cpp:494-532 (fp32 trig on fp32 angles)
template <class T>
Mx<T> get_H(const T mu[3], const float angs[3]) {
  using std::cos;
  using std::sin;
  float phi = angs[0], theta = angs[1], psi = angs[2];
  Mx<T> H(3, 6);
  H(0, 0) = T(-1);
  H(1, 1) = T(-1);
  H(2, 2) = T(-1);
  float Jx[9] = {0.f, (-sin(psi) * sin(phi) + cos(phi) * sin(theta) * cos(psi)),
                 (cos(phi) * sin(psi) + sin(theta) * sin(phi) * cos(psi)),
                 0.f, (-sin(phi) * cos(psi) - cos(phi) * sin(theta) * sin(psi)),
                 (cos(phi) * cos(psi) - sin(theta) * sin(psi) * sin(phi)),
                 0.f, (-cos(phi) * cos(theta)), (-sin(phi) * cos(theta))};
  float Jy[9] = {(-sin(theta) * cos(psi)), (cos(theta) * sin(phi) * cos(psi)),
                 (-cos(theta) * cos(phi) * cos(psi)),
                 (sin(psi) * sin(theta)), (-cos(theta) * sin(phi) * sin(psi)),
                 (cos(theta) * sin(psi) * cos(phi)),
                 (cos(theta)), (sin(phi) * sin(theta)), (-sin(theta) * cos(phi))};
  float Jz[9] = {(-cos(theta) * sin(psi)), (cos(psi) * cos(phi) - sin(phi) * sin(theta) * sin(psi)),
                 (cos(psi) * sin(phi) + sin(theta) * cos(phi) * sin(psi)),
                 (-cos(psi) * cos(theta)), (-sin(psi) * cos(phi) - sin(phi) * sin(theta) * cos(psi)),
                 (-sin(phi) * sin(psi) + sin(theta) * cos(psi) * cos(phi)),
                 0.f, 0.f, 0.f};
  for (int i = 0; i < 3; i++) {
    T sx = T(Jx[3 * i]) * mu[0], sy = T(Jy[3 * i]) * mu[0], sz = T(Jz[3 * i]) * mu[0];
    for (int k = 1; k < 3; k++) {
      sx = sx + T(Jx[3 * i + k]) * mu[k];
      sy = sy + T(Jy[3 * i + k]) * mu[k];
      sz = sz + T(Jz[3 * i + k]) * mu[k];
    }
    H(i, 3) = sx;
    H(i, 4) = sy;
    H(i, 5) = sz;
  }
  return H;
}

// bin indices, ICET::sortSphericalCoordinates src/icet.cpp:545-546 (double math on fp32 angles)
inline void bin_of(float th, float ph, int nT, int nP, int& bt, int& bp) {
  bt = static_cast<int>((th / (2 * M_PI)) * nT) % nT;
  bp = static_cast<int>((ph / M_PI) * nP) % nP;
}

struct Cloud {  // spherical coordinates, SoA (= the reference's column-major N x 3)
  std::vector<float> r, th, ph;
  void resize(size_t n) { r.resize(n); th.resize(n); ph.resize(n); }
  size_t size() const { return r.size(); }
};

// the radial "sort" of src/icet.cpp:72-83 (scan 1) and :264-274 (scan 2)
void radial_order(Cloud& s, std::vector<int>& orig, int mode) {
  const int N = (int)s.size();
  std::vector<int> index(N);
  std::iota(index.begin(), index.end(), 0);
  if (mode == ORACLE_ORDER_REF_SHIPPED) {
    // std::sort(std::execution::par, ...) -- libstdc++'s serial PSTL backend (no TBB here) is std::sort
    std::sort(index.begin(), index.end(), [&](int a, int b) { return s.r[a] < s.r[b]; });
    for (int i = 0; i < N; i++) {
      if (index[i] != i) {  // NOT a valid permutation application; kept bug-for-bug (:78-83)
        int j = index[i];
        std::swap(s.r[i], s.r[j]);
        std::swap(s.th[i], s.th[j]);
        std::swap(s.ph[i], s.ph[j]);
        std::swap(orig[i], orig[j]);
        std::swap(index[i], index[j]);
      }
    }
  } else {
    std::stable_sort(index.begin(), index.end(), [&](int a, int b) { return s.r[a] < s.r[b]; });
    Cloud t;
    t.resize(N);
    std::vector<int> o(N);
    for (int i = 0; i < N; i++) {
      t.r[i] = s.r[index[i]];
      t.th[i] = s.th[index[i]];
      t.ph[i] = s.ph[index[i]];
      o[i] = orig[index[i]];
    }
    s = std::move(t);
    orig = std::move(o);
  }
}

// ICET::findCluster, src/icet.cpp:557-607 -- walks the cell's points in stored order
std::pair<float, float> find_cluster(const std::vector<float>& r, int n, float thresh, float buff) {
  const int numPoints = (int)r.size();
  float innerDistance = 0.0f, outerDistance = 0.0f;
  int start = 0, len = 0;  // localPoints == r[start .. start+len)
  for (int i = 0; i < numPoints; i++) {
    if (len > 0 && std::abs(r[start + len - 1] - r[i]) <= thresh) {
      len++;
    } else {
      if (len >= n) {
        innerDistance = r[start] - buff;
        outerDistance = r[start + len - 1] + buff;
        return {innerDistance, outerDistance};
      } else {
        start = i;
        len = 1;
      }
    }
  }
  if (len >= n) {
    if (r[start] != 0) {
      innerDistance = r[start] - buff;
      outerDistance = r[start + len - 1] + buff;
      return {innerDistance, outerDistance};
    } else {
      return {0.0f, 0.0f};
    }
  }
  return {innerDistance, outerDistance};
}

// ---------------------------------------------------------------------------------------------
// the registration object
// ---------------------------------------------------------------------------------------------
template <class T>
struct Voxel1 {  // sigma1 / mu1 / U / L map entries of include/icet.h:89-94
  bool has = false;
  T mu[3];
  T sigma[9];
  float V[9];   // eigenvectors (columns); the reference stores U = V^T and always uses U^T
  float Ld[3];  // diagonal of L
};

template <class T>
struct Icet {
  oracle_params P;
  int nT, nP, ncell;
  Cloud sph1, sph2;
  std::vector<int> orig1, orig2;
  std::vector<float> ogx, ogy, ogz;  // points2_OG
  std::vector<std::vector<std::vector<int>>> idx1, idx2;  // [theta][phi]
  std::vector<float> bounds;  // ncell x 6
  std::vector<Voxel1<T>> vox;
  float X[6];
  float pred_stds[6];
  float Q[36];
  int status = 0;
  oracle_out* out;

  // ICET::sortSphericalCoordinates src/icet.cpp:534-554 (takes its matrix BY VALUE -> copy)
  std::vector<std::vector<std::vector<int>>> sort_spherical(Cloud s) {
    std::vector<std::vector<std::vector<int>>> pi(nT, std::vector<std::vector<int>>(nP));
    for (int i = 0; i < (int)s.size(); ++i) {
      int bt, bp;
      bin_of(s.th[i], s.ph[i], nT, nP, bt, bp);
      pi[bt][bp].push_back(i);
    }
    return pi;
  }

  // gather + ICET::filterPointsInsideCluster src/icet.cpp:609-652, then
  // sphericalToCartesian + mean + covariance (src/icet.cpp:159-162 / :303-306).
  // Returns the number of surviving rows; fills mean/cov when rows satisfy `enough`.
  int filter_stats(const Cloud& s, const std::vector<int>& ids, const float* lim, T mean[3],
                   T cov[9], bool (*enough)(int rows, int n), std::vector<int>* kept = nullptr) {
    std::vector<float> sel_r(ids.size()), sel_t(ids.size()), sel_p(ids.size());
    for (size_t i = 0; i < ids.size(); ++i) {  // selectedPoints gather :120-123 / :293-296
      sel_r[i] = s.r[ids[i]];
      sel_t[i] = s.th[ids[i]];
      sel_p[i] = s.ph[ids[i]];
    }
    std::vector<float> fr, ft, fp;
    fr.reserve(ids.size()); ft.reserve(ids.size()); fp.reserve(ids.size());
    for (size_t j = 0; j < ids.size(); ++j) {
      float azim = sel_t[j], elev = sel_p[j], r = sel_r[j];
      if (azim >= lim[0] && azim <= lim[1] && elev >= lim[2] && elev <= lim[3] && r >= lim[4] &&
          r <= lim[5]) {
        fr.push_back(r); ft.push_back(azim); fp.push_back(elev);
        if (kept) kept->push_back(ids[j]);
      }
    }
    const int rows = (int)fr.size();
    if (!enough(rows, P.n)) return rows;
    std::vector<float> cx(rows), cy(rows), cz(rows);
    for (int i = 0; i < rows; i++) s2c(fr[i], ft[i], fp[i], cx[i], cy[i], cz[i]);
    T sx = T(0), sy = T(0), sz = T(0);
    for (int i = 0; i < rows; i++) { sx = sx + T(cx[i]); sy = sy + T(cy[i]); sz = sz + T(cz[i]); }
    mean[0] = sx / T(rows); mean[1] = sy / T(rows); mean[2] = sz / T(rows);
    T c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < rows; i++) {
      T d[3] = {T(cx[i]) - mean[0], T(cy[i]) - mean[1], T(cz[i]) - mean[2]};
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) c[3 * a + b] = c[3 * a + b] + d[a] * d[b];
    }
    for (int a = 0; a < 9; a++) cov[a] = c[a] / T(rows - 1);
    return rows;
  }

  // ICET::fitCells1, src/icet.cpp:109-252
  void fit_cells1(const std::vector<int>& indices, int theta, int phi) {
    const int cell = nT * phi + theta;
    float* row = &bounds[6 * cell];
    float azimMin = (static_cast<float>(theta) / nT) * (2 * M_PI);
    float azimMax = (static_cast<float>(theta + 1) / nT) * (2 * M_PI);
    float elevMin = (static_cast<float>(phi) / nP) * (M_PI);
    float elevMax = (static_cast<float>(phi + 1) / nP) * (M_PI);
    if ((int)indices.size() >= P.n) {
      std::vector<float> rr(indices.size());
      for (size_t i = 0; i < indices.size(); ++i) rr[i] = sph1.r[indices[i]];
      auto cd = find_cluster(rr, P.n, P.thresh, P.buff);
      float inner = cd.first, outer = cd.second;
      row[0] = azimMin; row[1] = azimMax; row[2] = elevMin; row[3] = elevMax;
      row[4] = inner; row[5] = outer;
      T mean[3], cov[9];
      // `outerDistance > 0.1 && filteredPoints.size() >= n` with size() = 3*rows   (:158)
      int rows = filter_stats(sph1, indices, row, mean, cov,
                              [](int rws, int n) { return 3 * rws >= n; });
      if (out->nin1) out->nin1[cell] = rows;
      if (outer > 0.1 && 3 * rows >= P.n) {
        Voxel1<T>& v = vox[cell];
        v.has = true;
        for (int i = 0; i < 3; i++) v.mu[i] = mean[i];
        for (int i = 0; i < 9; i++) v.sigma[i] = cov[i];
        // SelfAdjointEigenSolver<Matrix3f> :181-184  (fp32 in both precision modes)
        float A[9], ev[3];
        for (int i = 0; i < 9; i++) A[i] = (float)cov[i];
        eig3<float>(A, P.eigen_flavor, ev, v.V);
        if (out->evec1_in && out->evec1_in_mask && out->evec1_in_mask[cell])  // test-harness injection
          for (int i = 0; i < 9; i++) v.V[i] = out->evec1_in[9 * cell + i];
        // sigma points :187-202: rotated = 2*sqrt(diag(ev)) * U^T = rows of (2 sqrt(ev_k)) V.row(k)
        float mu[3] = {(float)mean[0], (float)mean[1], (float)mean[2]};
        float sp[6][3];
        for (int k = 0; k < 3; k++) {
          float al = (float)(2.0 * (double)std::sqrt(ev[k]));  // 2.0 * axislen.array().sqrt()
          for (int c = 0; c < 3; c++) {
            // axislen * U^T : entry (k,c) = sum_j axislen(k,j) * V(j,c), only j=k is non-zero
            float rot = 0.f;
            for (int j = 0; j < 3; j++) {
              float a = (j == k) ? al : 0.f;
              rot = (j == 0) ? a * v.V[3 * j + c] : rot + a * v.V[3 * j + c];
            }
            sp[2 * k][c] = mu[c] + rot;
            sp[2 * k + 1][c] = mu[c] - rot;
          }
        }
        // c2s + ICET::testSigmaPoints :654-696 (breaks at the first point with r > outer)
        bool inside[6] = {false, false, false, false, false, false};
        for (int j = 0; j < 6; j++) {
          float r, th, ph;
          c2s(sp[j][0], sp[j][1], sp[j][2], r, th, ph);
          if (th >= row[0] && th <= row[1] && ph >= row[2] && ph <= row[3] && r >= row[4] &&
              r <= row[5])
            inside[j] = true;
          if (r > row[5]) break;
        }
        for (int k = 0; k < 3; k++) v.Ld[k] = (inside[2 * k] || inside[2 * k + 1]) ? 1.f : 0.f;
        if (out->has1) out->has1[cell] = 1;
        if (out->mu1) for (int i = 0; i < 3; i++) out->mu1[3 * cell + i] = (float)mean[i];
        if (out->sigma1) for (int i = 0; i < 9; i++) out->sigma1[9 * cell + i] = (float)cov[i];
        if (out->eval1) for (int i = 0; i < 3; i++) out->eval1[3 * cell + i] = ev[i];
        if (out->evec1) for (int i = 0; i < 9; i++) out->evec1[9 * cell + i] = v.V[i];
        if (out->lmask) for (int i = 0; i < 3; i++) out->lmask[3 * cell + i] = (uint8_t)v.Ld[i];
      }
    } else {
      row[0] = azimMin; row[1] = azimMax; row[2] = elevMin; row[3] = elevMax;
      row[4] = 0.f; row[5] = 0.f;
    }
  }

  // ICET::fitScan1, src/icet.cpp:68-107
  void fit_scan1(const float* scan1, int n1, int ld1) {
    sph1.resize(n1);
    orig1.resize(n1);
    std::iota(orig1.begin(), orig1.end(), 0);
    for (int i = 0; i < n1; i++)
      c2s(scan1[i], scan1[ld1 + i], scan1[2 * ld1 + i], sph1.r[i], sph1.th[i], sph1.ph[i]);
    if (out->sph1)
      for (int i = 0; i < n1; i++) {
        out->sph1[i] = sph1.r[i]; out->sph1[n1 + i] = sph1.th[i]; out->sph1[2 * n1 + i] = sph1.ph[i];
      }
    radial_order(sph1, orig1, P.order_mode);
    idx1 = sort_spherical(sph1);
    if (out->cell1 || out->cnt1)
      for (int t = 0; t < nT; t++)
        for (int p = 0; p < nP; p++) {
          if (out->cnt1) out->cnt1[nT * p + t] = (int)idx1[t][p].size();
          if (out->cell1)
            for (int i : idx1[t][p]) out->cell1[orig1[i]] = nT * p + t;
        }
    for (int phi = 0; phi < nP; phi++)
      for (int theta = 0; theta < nT; theta++) fit_cells1(idx1[theta][phi], theta, phi);
  }

  // ICET::prepScan2, src/icet.cpp:254-277
  void prep_scan2(const float* scan2, int n2, int ld2) {
    sph2.resize(n2);
    orig2.resize(n2);
    std::iota(orig2.begin(), orig2.end(), 0);
    for (int i = 0; i < n2; i++)
      c2s(scan2[i], scan2[ld2 + i], scan2[2 * ld2 + i], sph2.r[i], sph2.th[i], sph2.ph[i]);
    radial_order(sph2, orig2, P.order_mode);
    ogx.resize(n2); ogy.resize(n2); ogz.resize(n2);
    for (int i = 0; i < n2; i++) s2c(sph2.r[i], sph2.th[i], sph2.ph[i], ogx[i], ogy[i], ogz[i]);
    if (out->perm2) for (int i = 0; i < n2; i++) out->perm2[i] = orig2[i];
  }

  // ICET::fitCells2, src/icet.cpp:279-344.  Returns true if the voxel contributed.
  bool fit_cells2(const std::vector<int>& i1, const std::vector<int>& i2, int theta, int phi,
                  Mx<T>& HTWH_j, Mx<T>& HTWdz_j, int it) {
    const int cell = nT * phi + theta;
    HTWH_j = Mx<T>(6, 6);
    HTWdz_j = Mx<T>(6, 1);
    const float* row = &bounds[6 * cell];
    if (!((int)i2.size() > P.n && (int)i1.size() > P.n && row[5] > 1)) return false;
    T mean[3], cov[9];
    std::vector<int> kept;
    const bool want_kept = P.stats2_mode == ORACLE_STATS2_MOMENTS || out->in2;
    int rows = filter_stats(sph2, i2, row, mean, cov, [](int rws, int n) { return rws > n; },
                            want_kept ? &kept : nullptr);
    if (out->nin2) out->nin2[(size_t)it * ncell + cell] = rows;
    if (out->in2)
      for (int i : kept) out->in2[(size_t)it * ogx.size() + orig2[i]] = 1;
    if (!(rows > P.n)) return false;
    if (P.stats2_mode == ORACLE_STATS2_MOMENTS) {
      // Diagnostic twin of the product's incremental loop: SAME membership (the fp32 per-point pipeline above), but
      // mean / covariance from the exact moments of the untransformed members points2_OG, carried through the
      // transform analytically in double:  mu2 = (mean_OG + t) R,  Sigma2 = R^T Cov_OG R  (src/icet.cpp:375-378).
      double m[3] = {0, 0, 0};
      for (int i : kept) { m[0] += ogx[i]; m[1] += ogy[i]; m[2] += ogz[i]; }
      for (int a = 0; a < 3; a++) m[a] /= (double)rows;
      double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int i : kept) {
        const double d[3] = {ogx[i] - m[0], ogy[i] - m[1], ogz[i] - m[2]};
        for (int a = 0; a < 3; a++)
          for (int b = 0; b < 3; b++) c[3 * a + b] += d[a] * d[b];
      }
      for (int a = 0; a < 9; a++) c[a] /= (double)(rows - 1);
      float Rf[9];
      rotR(X[3], X[4], X[5], Rf);
      const double po[3] = {m[0] + (double)X[0], m[1] + (double)X[1], m[2] + (double)X[2]};
      for (int i = 0; i < 3; i++) mean[i] = T(po[0] * Rf[i] + po[1] * Rf[3 + i] + po[2] * Rf[6 + i]);
      double t9[9];
      for (int a = 0; a < 3; a++)
        for (int j = 0; j < 3; j++)
          t9[3 * a + j] = c[3 * a] * Rf[j] + c[3 * a + 1] * Rf[3 + j] + c[3 * a + 2] * Rf[6 + j];
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
          cov[3 * i + j] = T(Rf[i] * t9[j] + Rf[3 + i] * t9[3 + j] + Rf[6 + i] * t9[6 + j]);
    }
    const Voxel1<T>& v = vox[cell];
    // The reference reads sigma1/L/U/mu1 through std::map::operator[] (:315-336); when scan 1
    // fitted no Gaussian here that is a read of default-constructed (uninitialised) matrices.
    // Defined as "skip the voxel" (SURVEY.md H9); never triggered on the bundled fixtures.
    if (!v.has) return false;
    if (out->mu2) for (int i = 0; i < 3; i++) out->mu2[((size_t)it * ncell + cell) * 3 + i] = (float)mean[i];
    if (out->sigma2) for (int i = 0; i < 9; i++) out->sigma2[((size_t)it * ncell + cell) * 9 + i] = (float)cov[i];
    Mx<T> Rn(3, 3), L(3, 3), Ut(3, 3) /* = U.transpose() = V */, U(3, 3);
    const T d1 = T(i1.size() - 1), d2 = T(i2.size() - 1);
    for (int i = 0; i < 9; i++) Rn.a[i] = v.sigma[i] / d1 + cov[i] / d2;  // :315
    for (int i = 0; i < 3; i++) L(i, i) = T(v.Ld[i]);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) { Ut(i, j) = T(v.V[3 * i + j]); U(j, i) = T(v.V[3 * i + j]); }
    Mx<T> LUt = mul(L, Ut);
    Rn = mul(mul(mul(LUt, Rn), U), transpose(L));  // :317
    Mx<T> W = cod_pinv(Rn);                        // :320-321
    Mx<T> H_j = get_H<T>(mean, &X[3]);             // :324-326
    Mx<T> H_z = mul(LUt, H_j);                     // :329
    Mx<T> HzT = transpose(H_z);
    Mx<T> HzTW = mul(HzT, W);
    HTWH_j = mul(HzTW, H_z);                       // :332
    Mx<T> m1(3, 1), m2(3, 1);
    for (int i = 0; i < 3; i++) { m1(i, 0) = v.mu[i]; m2(i, 0) = mean[i]; }
    Mx<T> z1 = mul(LUt, m1), z2 = mul(LUt, m2), dz(3, 1);
    for (int i = 0; i < 3; i++) dz(i, 0) = z2(i, 0) - z1(i, 0);
    HTWdz_j = mul(HzTW, dz);                       // :338
    return true;
  }

  // ICET::fitScan2, src/icet.cpp:372-436 and ICET::checkCondition :443-492
  void fit_scan2(int it) {
    const int n2 = (int)ogx.size();
    if (out->X_in)  // test harness: start this iteration from the given iterate (see icet_oracle.h)
      for (int k = 0; k < 6; k++) X[k] = out->X_in[(size_t)it * 6 + k];
    float R[9];
    rotR(X[3], X[4], X[5], R);
    std::vector<float> px(n2), py(n2), pz(n2);
    for (int i = 0; i < n2; i++) {  // (points2_OG.rowwise() + trans) * rot_mat   :377-378
      float ax = ogx[i] + X[0], ay = ogy[i] + X[1], az = ogz[i] + X[2];
      float t;
      t = ax * R[0]; t = t + ay * R[3]; t = t + az * R[6]; px[i] = t;
      t = ax * R[1]; t = t + ay * R[4]; t = t + az * R[7]; py[i] = t;
      t = ax * R[2]; t = t + ay * R[5]; t = t + az * R[8]; pz[i] = t;
    }
    if (out->points2_final && it == P.runlen - 1)
      for (int i = 0; i < n2; i++) {
        out->points2_final[i] = px[i]; out->points2_final[n2 + i] = py[i];
        out->points2_final[2 * n2 + i] = pz[i];
      }
    for (int i = 0; i < n2; i++) c2s(px[i], py[i], pz[i], sph2.r[i], sph2.th[i], sph2.ph[i]);  // :387
    idx2 = sort_spherical(sph2);                                                                // :388
    if (out->sph2)
      for (int i = 0; i < n2; i++) {
        float* o = out->sph2 + (size_t)it * 3 * n2;
        o[orig2[i]] = sph2.r[i]; o[n2 + orig2[i]] = sph2.th[i]; o[2 * n2 + orig2[i]] = sph2.ph[i];
      }
    if (out->cell2 || out->cnt2)
      for (int t = 0; t < nT; t++)
        for (int p = 0; p < nP; p++) {
          if (out->cnt2) out->cnt2[(size_t)it * ncell + nT * p + t] = (int)idx2[t][p].size();
          if (out->cell2)
            for (int i : idx2[t][p]) out->cell2[(size_t)it * n2 + orig2[i]] = nT * p + t;
        }
    Mx<T> HTWH(6, 6), HTWdz(6, 1);
    for (int phi = 0; phi < nP; phi++)
      for (int theta = 0; theta < nT; theta++) {
        Mx<T> Hj, dj;
        bool used = fit_cells2(idx1[theta][phi], idx2[theta][phi], theta, phi, Hj, dj, it);
        if (out->used2) out->used2[(size_t)it * ncell + nT * phi + theta] = used ? 1 : 0;
        for (int i = 0; i < 36; i++) HTWH.a[i] = HTWH.a[i] + Hj.a[i];   // :401
        for (int i = 0; i < 6; i++) HTWdz.a[i] = HTWdz.a[i] + dj.a[i];  // :402
      }
    // noise matrix :410-417
    Mx<T> noise = cod_pinv(HTWH);
    for (int k = 0; k < 6; k++) pred_stds[k] = (float)std::sqrt(std::abs(noise(k, k)));
    for (int i = 0; i < 36; i++) Q[i] = (float)noise.a[i];
    // checkCondition :443-492
    T ev[6], U2[36];
    eigsym<T>(HTWH.a, 6, P.eigen_flavor, ev, U2);
    const float cutoff = 1e6f;
    T condition = ev[5] / ev[0];
    const T cond0 = condition;
    int eyecount = 1;
    int dropped = 0;
    while (std::abs(condition) > cutoff) {
      if (eyecount > 5) {  // eigenvalues(6): the reference would fail an Eigen assert here
        status = 1;
        break;
      }
      // pred_stds += U2.transpose().row(eyecount-1)   (a SIGNED eigenvector, :479)
      for (int k = 0; k < 6; k++) pred_stds[k] = (float)(T(pred_stds[k]) + U2[k * 6 + (eyecount - 1)]);
      dropped++;
      condition = ev[5] / ev[eyecount];
      eyecount++;
    }
    // L2: identity with the top `dropped` rows removed; lam = diag(ev)
    const int keep = 6 - dropped;
    Mx<T> L2(keep, 6), lam(6, 6), U2m(6, 6);
    for (int i = 0; i < keep; i++) L2(i, dropped + i) = T(1);
    for (int i = 0; i < 6; i++) lam(i, i) = ev[i];
    for (int i = 0; i < 36; i++) U2m.a[i] = U2[i];
    Mx<T> U2t = transpose(U2m);
    Mx<T> innards = mul(mul(L2, lam), U2t);   // :427
    Mx<T> inv = cod_pinv(innards);            // :428-429
    Mx<T> dxm = mul(mul(mul(inv, L2), U2t), HTWdz);  // :430
    float dx[6];
    for (int k = 0; k < 6; k++) {
      dx[k] = (float)dxm(k, 0);
      X[k] = (float)(T(X[k]) + dxm(k, 0));   // X += dx :433
    }
    if (out->HTWH) for (int i = 0; i < 36; i++) out->HTWH[(size_t)it * 36 + i] = (float)HTWH.a[i];
    if (out->HTWdz) for (int i = 0; i < 6; i++) out->HTWdz[(size_t)it * 6 + i] = (float)HTWdz.a[i];
    if (out->dx) for (int i = 0; i < 6; i++) out->dx[(size_t)it * 6 + i] = dx[i];
    if (out->Xit) for (int i = 0; i < 6; i++) out->Xit[(size_t)it * 6 + i] = X[i];
    if (out->Qit) for (int i = 0; i < 36; i++) out->Qit[(size_t)it * 36 + i] = Q[i];
    if (out->stds_it) for (int i = 0; i < 6; i++) out->stds_it[(size_t)it * 6 + i] = pred_stds[i];
    if (out->cond_it) out->cond_it[it] = (float)cond0;
    if (out->trunc_it) out->trunc_it[it] = dropped;
  }

  // ICET::ICET, src/icet.cpp:29-63
  int run(const oracle_params* p, const float* scan1, int n1, int ld1, const float* scan2, int n2,
          int ld2, const float x0[6], oracle_out* o) {
    P = *p;
    out = o;
    nT = P.bins_theta; nP = P.bins_phi; ncell = nT * nP;
    for (int i = 0; i < 6; i++) { X[i] = x0[i]; pred_stds[i] = 0.f; }
    for (int i = 0; i < 36; i++) Q[i] = 0.f;
    bounds.assign((size_t)ncell * 6, 0.f);
    vox.assign(ncell, Voxel1<T>());
    if (o->nin1) std::fill(o->nin1, o->nin1 + ncell, -1);
    if (o->has1) std::memset(o->has1, 0, ncell);
    if (o->nin2) std::fill(o->nin2, o->nin2 + (size_t)P.runlen * ncell, -1);
    if (o->in2) std::memset(o->in2, 0, (size_t)P.runlen * n2);
    fit_scan1(scan1, n1, ld1);
    if (o->bounds) std::memcpy(o->bounds, bounds.data(), sizeof(float) * 6 * ncell);
    prep_scan2(scan2, n2, ld2);
    for (int it = 0; it < P.runlen; it++) fit_scan2(it);
    for (int i = 0; i < 6; i++) { o->X[i] = X[i]; o->pred_stds[i] = pred_stds[i]; }
    for (int i = 0; i < 36; i++) o->Q[i] = Q[i];
    o->status = status;
    return 0;
  }
};

}  // namespace

extern "C" {

int icet_oracle_run(const oracle_params* p, const float* scan1, int32_t n1, int32_t ld1,
                    const float* scan2, int32_t n2, int32_t ld2, const float x0[6],
                    oracle_out* out) {
  if (!p || !scan1 || !scan2 || !out || n1 < 0 || n2 < 0 || p->bins_phi <= 0 || p->bins_theta <= 0)
    return -1;
  if (p->precise) {
    Icet<double> it;
    return it.run(p, scan1, n1, ld1, scan2, n2, ld2, x0, out);
  }
  Icet<float> it;
  return it.run(p, scan1, n1, ld1, scan2, n2, ld2, x0, out);
}

double icet_oracle_run_sequence(const oracle_params* p, const float* scans, int32_t n,
                                int32_t npairs, int32_t nthreads, float* results) {
  if (nthreads < 1) nthreads = 1;
  auto t0 = std::chrono::steady_clock::now();
  auto worker = [&](int tid) {
    const float x0[6] = {0, 0, 0, 0, 0, 0};
    for (int i = tid; i < npairs; i += nthreads) {
      oracle_out o;
      std::memset(&o, 0, sizeof(o));
      icet_oracle_run(p, scans + (size_t)i * 3 * n, n, n, scans + (size_t)(i + 1) * 3 * n, n, n, x0,
                      &o);
      if (results) {
        float* r = results + (size_t)i * 48;
        std::memcpy(r, o.X, 24);
        std::memcpy(r + 6, o.pred_stds, 24);
        std::memcpy(r + 12, o.Q, 144);
      }
    }
  };
  if (nthreads == 1) {
    worker(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back(worker, t);
    for (auto& t : th) t.join();
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

void icet_oracle_eig3(const float a[9], int32_t flavor, float evals[3], float evecs[9]) {
  eig3<float>(a, flavor, evals, evecs);
}
void icet_oracle_eigsym(const float* a, int32_t n, int32_t flavor, float* evals, float* evecs) {
  eigsym<float>(a, n, flavor, evals, evecs);
}
void icet_oracle_pinv(const float* a, int32_t rows, int32_t cols, float* out, int32_t* rank) {
  Mx<float> A(rows, cols);
  for (int i = 0; i < rows * cols; i++) A.a[i] = a[i];
  int rk = 0;
  Mx<float> P = cod_pinv(A, &rk);
  for (int i = 0; i < rows * cols; i++) out[i] = P.a[i];
  if (rank) *rank = rk;
}
void icet_oracle_c2s(const float* xyz, int32_t n, int32_t ld, float* sph) {
  for (int i = 0; i < n; i++) c2s(xyz[i], xyz[ld + i], xyz[2 * ld + i], sph[i], sph[n + i], sph[2 * n + i]);
}
void icet_oracle_bins(const float* sph, int32_t n, int32_t bins_phi, int32_t bins_theta,
                      int32_t* cell) {
  for (int i = 0; i < n; i++) {
    int bt, bp;
    bin_of(sph[n + i], sph[2 * n + i], bins_theta, bins_phi, bt, bp);
    cell[i] = bins_theta * bp + bt;
  }
}

}  // extern "C"
