/*
 * oracle/ref/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * C entry point around the REFERENCE's own `class ICET` (reference include/icet.h:36-116), compiled together with the
 * reference's unmodified src/icet.cpp, src/utils.cpp and src/ThreadPool.cpp where they lie under /root/reference
 * (oracle/Makefile `ref`).  Constructs the object exactly like icet_cpp_demo.cpp:31-38 / odometry.cpp:73-76 do and
 * copies its public members out, so that tools/pin_against_ref.py can diff them against the oracle restatement and
 * bench.py can time the reference's CPU path.  No reference source is copied into this repository.
 */
#include <cstdint>
#include <cstring>

#include "icet.h"  // the reference's header (-I/root/reference/include)

extern "C" {

struct ref_out {
  float X[6];
  float pred_stds[6];
  int32_t n_ellipsoids;      /* ellipsoid1Means.size() = voxels with a scan-1 Gaussian               */
  float* clusterBounds;      /* [ncell*6] row-major (clusterBounds.row(k))                            */
  int32_t* cnt1;             /* [ncell] pointIndices1[theta][phi].size(), index nTheta*phi + theta    */
  int32_t* cnt2;             /* [ncell] pointIndices2 after the LAST iteration                        */
  uint8_t* has1;             /* [ncell] sigma1 has an entry [theta][phi]                              */
  float* mu1;                /* [ncell*3]                                                             */
  float* sigma1;             /* [ncell*9] row-major                                                   */
  float* U;                  /* [ncell*9] row-major (U = eigenvectors^T, src/icet.cpp:184)            */
  float* L;                  /* [ncell*9] row-major                                                   */
  float* points2;            /* [3*n2] column planes: public member points2 (internal row order)      */
  float* HTWH;               /* [36] row-major, last iteration                                        */
  float* HTWdz;              /* [6]                                                                   */
  float* testPoints;         /* [ncell*6*3] row-major; rows the reference never writes are returned as written by
                                the shim's zero-initialised resize()                                    */
};

/* scan1 / scan2: column-major N x 3 planes (Eigen::MatrixXf::data()).  Returns 0, or 1 when the constructor threw. */
int icet_ref_run(const float* scan1, int32_t n1, const float* scan2, int32_t n2, int32_t runlen, const float x0[6],
                 int32_t bins_phi, int32_t bins_theta, int32_t n, float thresh, float buff, ref_out* out) {
  Eigen::MatrixXf s1(n1, 3), s2(n2, 3);
  std::memcpy(s1.data(), scan1, sizeof(float) * 3 * (size_t)n1);
  std::memcpy(s2.data(), scan2, sizeof(float) * 3 * (size_t)n2);
  Eigen::VectorXf X0(6);
  for (int k = 0; k < 6; k++) X0[k] = x0[k];
  try {
    ICET it(s1, s2, runlen, X0, bins_phi, bins_theta, n, thresh, buff);
    for (int k = 0; k < 6; k++) { out->X[k] = it.X[k]; out->pred_stds[k] = it.pred_stds[k]; }
    out->n_ellipsoids = (int32_t)it.ellipsoid1Means.size();
    const int ncell = bins_phi * bins_theta;
    for (int phi = 0; phi < bins_phi; phi++)
      for (int theta = 0; theta < bins_theta; theta++) {
        const int c = bins_theta * phi + theta;
        if (out->clusterBounds)
          for (int k = 0; k < 6; k++) out->clusterBounds[6 * c + k] = it.clusterBounds(c, k);
        if (out->cnt1) out->cnt1[c] = (int32_t)it.pointIndices1[theta][phi].size();
        if (out->cnt2) out->cnt2[c] = runlen > 0 ? (int32_t)it.pointIndices2[theta][phi].size() : 0;
        bool has = false;
        auto a = it.sigma1.find(theta);
        if (a != it.sigma1.end()) has = a->second.find(phi) != a->second.end();
        if (out->has1) out->has1[c] = has ? 1 : 0;
        if (has) {
          for (int i = 0; i < 3; i++) {
            if (out->mu1) out->mu1[3 * c + i] = it.mu1[theta][phi][i];
            for (int j = 0; j < 3; j++) {
              if (out->sigma1) out->sigma1[9 * c + 3 * i + j] = it.sigma1[theta][phi](i, j);
              if (out->U) out->U[9 * c + 3 * i + j] = it.U[theta][phi](i, j);
              if (out->L) out->L[9 * c + 3 * i + j] = it.L[theta][phi](i, j);
            }
          }
        }
      }
    if (out->points2)
      for (int j = 0; j < 3; j++)
        for (int i = 0; i < n2; i++) out->points2[(size_t)j * n2 + i] = it.points2(i, j);
    if (out->HTWH && runlen > 0)
      for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) out->HTWH[6 * i + j] = it.HTWH_i(i, j);
    if (out->HTWdz && runlen > 0)
      for (int i = 0; i < 6; i++) out->HTWdz[i] = it.HTWdz_i(i, 0);
    if (out->testPoints)
      for (int i = 0; i < 6 * ncell; i++)
        for (int j = 0; j < 3; j++) out->testPoints[3 * i + j] = it.testPoints(i, j);
  } catch (const std::exception&) {
    return 1;
  }
  return 0;
}

/* Registers npairs consecutive pairs (scans[i], scans[i+1]) of equally sized clouds stored back to back, X0 = 0, on
 * nthreads host threads (one ICET object per pair, like the reference's callers); results = [npairs*12] (X | pred_stds).
 * Returns the elapsed wall time in seconds. */
double icet_ref_run_sequence(const float* scans, int32_t n, int32_t npairs, int32_t nthreads, int32_t runlen,
                             int32_t bins_phi, int32_t bins_theta, int32_t nmin, float thresh, float buff, float* results) {
  if (nthreads < 1) nthreads = 1;
  auto t0 = std::chrono::steady_clock::now();
  auto worker = [&](int tid) {
    for (int i = tid; i < npairs; i += nthreads) {
      Eigen::MatrixXf s1(n, 3), s2(n, 3);
      std::memcpy(s1.data(), scans + (size_t)i * 3 * n, sizeof(float) * 3 * (size_t)n);
      std::memcpy(s2.data(), scans + (size_t)(i + 1) * 3 * n, sizeof(float) * 3 * (size_t)n);
      Eigen::VectorXf X0(6);
      X0.setZero();
      ICET it(s1, s2, runlen, X0, bins_phi, bins_theta, nmin, thresh, buff);
      if (results)
        for (int k = 0; k < 6; k++) { results[(size_t)i * 12 + k] = it.X[k]; results[(size_t)i * 12 + 6 + k] = it.pred_stds[k]; }
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

/* utils::loadPointCloudCSV of the reference (src/utils.cpp:12-91): returns the number of rows; copies up to `cap` rows
 * (row-major x y z) into out. */
int icet_ref_load_csv(const char* filename, const char* dataset_type, float* out, int32_t cap) {
  Eigen::MatrixXf m = utils::loadPointCloudCSV(filename, dataset_type);
  for (int i = 0; i < m.rows() && i < cap; i++)
    for (int j = 0; j < 3; j++) out[3 * i + j] = m(i, j);
  return (int)m.rows();
}

}  // extern "C"
