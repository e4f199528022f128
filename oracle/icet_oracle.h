/*
 * oracle/icet_oracle.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * C interface of the CPU oracle: a plain C++17 restatement (no Eigen) of the
 * reference registration path  ICET::ICET  (reference src/icet.cpp:29-63 and
 * everything it calls, src/utils.cpp:93-152).
 *
 * PARITY UNPINNED: the reference ships no tests / golden vectors for this path
 *  and cannot be compiled in this image (Eigen3 absent), so
 * the oracle is anchored on the reference's call sites and on a restatement of
 * the Eigen 3.3.7 algorithms it calls (SelfAdjointEigenSolver,
 * CompleteOrthogonalDecomposition).  See DESIGN.md "Oracle".
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may load this library.
 */
#ifndef ICET_ORACLE_H
#define ICET_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* order_mode: how the "radial sort" of src/icet.cpp:72-83 / :264-274 is realised */
enum {
  ORACLE_ORDER_SORTED = 0,      /* true ascending-range order (stable) -- what the comments intend  */
  ORACLE_ORDER_REF_SHIPPED = 1  /* std::sort + the shipped (broken) in-place permutation loop       */
};

/* eigen_flavor: deflation test of the tridiagonal QR (Eigen 3.3.x vs 3.4.x) */
enum { ORACLE_EIGEN_337 = 0, ORACLE_EIGEN_340 = 1 };

typedef struct {
  int32_t runlen;      /* ICET ctor arg `runlen`          (include/icet.h:38)  */
  int32_t bins_phi;    /* `num_bins_phi`   (elevation bins)                    */
  int32_t bins_theta;  /* `num_bins_theta` (azimuth bins)                      */
  int32_t n;           /* min cluster size            (default 25)             */
  float thresh;        /* radial jump threshold       (default 0.1)            */
  float buff;          /* radial buffer               (default 0.1)            */
  int32_t order_mode;  /* ORACLE_ORDER_*                                       */
  int32_t eigen_flavor;/* ORACLE_EIGEN_*                                       */
  int32_t precise;     /* 0 = fp32 everywhere like the reference;
                          1 = diagnostic twin: same fp32 point geometry and fp32
                              3x3 eigen solver, but statistics / per-voxel algebra /
                              6x6 solve accumulated in double                  */
  int32_t stats2_mode; /* ORACLE_STATS2_*                                      */
} oracle_params;

/* stats2_mode: how fitCells2 forms the scan-2 mean / covariance of a voxel (membership is the per-point fp32 pipeline
 * of the reference in both modes) */
enum {
  ORACLE_STATS2_REFERENCE = 0, /* sphericalToCartesian of the filtered points, mean, covariance (src/icet.cpp:303-306) */
  ORACLE_STATS2_MOMENTS = 1    /* diagnostic twin of the product's incremental loop: exact moments of the members'
                                  points2_OG, transformed analytically in double ((mean + t) R, R^T Cov R)           */
};

/* All dump pointers are optional (NULL = skip).  ncell = bins_phi*bins_theta, cell
 * index = bins_theta*phi + theta  (the clusterBounds row index, src/icet.cpp:149).
 * Per-point dumps are in the ORIGINAL input order of the scan. */
typedef struct {
  /* results (always written) */
  float X[6];
  float pred_stds[6];
  float Q[36];              /* pinv(HTWH) of the last iteration (noise_mat, src/icet.cpp:411) row-major */
  int32_t status;           /* 0 ok, 1 = reference would have asserted (eyecount overflow)            */

  /* scan 1 */
  float* sph1;              /* [3*n1] r | theta | phi planes                                           */
  int32_t* cell1;           /* [n1]                                                                    */
  int32_t* cnt1;            /* [ncell] points in the angular bin (pointIndices1[theta][phi].size())    */
  float* bounds;            /* [ncell*6] clusterBounds rows                                            */
  int32_t* nin1;            /* [ncell] rows surviving filterPointsInsideCluster (-1: not evaluated)    */
  uint8_t* has1;            /* [ncell] 1 if a Gaussian (mu1/sigma1/U/L) was fitted                     */
  float* mu1;               /* [ncell*3]                                                               */
  float* sigma1;            /* [ncell*9] row-major                                                     */
  float* eval1;             /* [ncell*3] ascending eigenvalues                                         */
  float* evec1;             /* [ncell*9] row-major V (columns = eigenvectors)                          */
  uint8_t* lmask;           /* [ncell*3] diagonal of L                                                 */

  /* scan 2, per iteration (leading dimension = iteration) */
  float* sph2;              /* [runlen*3*n2]                                                           */
  int32_t* cell2;           /* [runlen*n2]                                                             */
  int32_t* cnt2;            /* [runlen*ncell]                                                          */
  int32_t* nin2;            /* [runlen*ncell] (-1: voxel gated off before the filter)                  */
  uint8_t* used2;           /* [runlen*ncell] voxel contributed to HTWH                                */
  float* mu2;               /* [runlen*ncell*3]                                                        */
  float* sigma2;            /* [runlen*ncell*9]                                                        */
  float* HTWH;              /* [runlen*36]                                                             */
  float* HTWdz;             /* [runlen*6]                                                              */
  float* dx;                /* [runlen*6]                                                              */
  float* Xit;               /* [runlen*6]  X after each iteration                                      */
  float* Qit;               /* [runlen*36]                                                             */
  float* stds_it;           /* [runlen*6]  pred_stds after each iteration (incl. the :479 inflation)   */
  float* cond_it;           /* [runlen]    lambda_max/lambda_min of HTWH                               */
  int32_t* trunc_it;        /* [runlen]    number of solution axes dropped by checkCondition           */
  uint8_t* in2;             /* [runlen*n2] 1 if the point survived filterPointsInsideCluster of a voxel that
                               passed the gates of fitCells2 (:290) in that iteration (original order)    */
  float* points2_final;     /* [3*n2] public member points2 (src/icet.cpp:377-378 of the last
                               iteration), in the oracle's permuted row order, column-major           */
  int32_t* perm2;           /* [n2] original index of each permuted scan-2 row                         */

  /* INPUT (optional, test harness only): eigenvector injection.  The signs Eigen's QR gives the 3x3
   * eigenvectors flip under 1e-6-relative perturbations of the covariance for some voxels,
   * so two correct implementations can disagree there.  For cells with evec1_in_mask[cell] != 0 the
   * oracle uses evec1_in[cell*9..] (row-major V) instead of its own eigenvectors; eigenvalues stay.  */
  const float* evec1_in;
  const uint8_t* evec1_in_mask;
  /* INPUT (optional, test harness only): iterate alignment.  When given, iteration `it` STARTS from X_in[it*6 .. +6]
   * instead of the oracle's own X (its update is still computed and reported in dx / Xit).  Lets a test compare
   * per-iteration classes and statistics of another implementation at that implementation's own iterates: after a large
   * first step two correct implementations differ by ~1e-5 m in X, which moves many more boundary points than libm. */
  const float* X_in;
} oracle_out;

/* Clouds are column-major N x 3 (x-plane | y-plane | z-plane) with leading
 * dimension ld (>= n), i.e. Eigen::MatrixXf::data(). Returns 0 on success. */
int icet_oracle_run(const oracle_params* p, const float* scan1, int32_t n1, int32_t ld1,
                    const float* scan2, int32_t n2, int32_t ld2, const float x0[6],
                    oracle_out* out);

/* Registers `npairs` independent pairs (scans[i], scans[i+1]) of a sequence of
 * npairs+1 equally sized clouds stored back to back (each 3*n floats), X0 = 0,
 * on `nthreads` host threads (pairs are distributed round-robin).  results =
 * [npairs*48] (X 6 | pred_stds 6 | Q 36).  Returns elapsed seconds (wall). */
double icet_oracle_run_sequence(const oracle_params* p, const float* scans, int32_t n,
                                int32_t npairs, int32_t nthreads, float* results);

/* Stand-alone pieces exported for unit tests. */
void icet_oracle_eig3(const float a[9], int32_t flavor, float evals[3], float evecs[9]);
void icet_oracle_eigsym(const float* a, int32_t n, int32_t flavor, float* evals, float* evecs);
void icet_oracle_pinv(const float* a, int32_t rows, int32_t cols, float* out /* cols x rows */,
                      int32_t* rank);
void icet_oracle_c2s(const float* xyz, int32_t n, int32_t ld, float* sph /* 3*n planes */);
void icet_oracle_bins(const float* sph, int32_t n, int32_t bins_phi, int32_t bins_theta,
                      int32_t* cell);

#ifdef __cplusplus
}
#endif
#endif
