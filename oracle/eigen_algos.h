/*
 * oracle/eigen_algos.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The Eigen 3.3.7 routines the reference calls on its hot path, restated from the published algorithms (Eigen is
 * absent from this image and is not vendored or pinned by the reference; ROS Noetic => 3.3.7):
 *   - SelfAdjointEigenSolver<Matrix3f>::compute   src/icet.cpp:181-184   -> eig3()
 *   - SelfAdjointEigenSolver<MatrixXf>::compute   src/icet.cpp:455-458   -> eigsym()
 *   - CompleteOrthogonalDecomposition<MatrixXf>::pseudoInverse()  src/icet.cpp:320-321, :410-411, :428-429 -> cod_pinv()
 * Shared by the oracle (oracle/icet_oracle.cpp) and by the Eigen-API shim (oracle/eigen_shim/Eigen/Dense) that lets
 * the reference's own sources compile here (oracle/Makefile `ref`).
 */
#ifndef ICET_ORACLE_EIGEN_ALGOS_H
#define ICET_ORACLE_EIGEN_ALGOS_H

#include <algorithm>
#include <cmath>
#include <limits>

namespace eigen_restated {

// ---------------------------------------------------------------------------------------------
// small dense helpers (runtime dims, up to 6x6), scalar type T in {float, double}
// ---------------------------------------------------------------------------------------------
template <class T>
struct Mx {
  int r = 0, c = 0;
  T a[36];
  Mx() {}
  Mx(int r_, int c_) : r(r_), c(c_) {
    for (int i = 0; i < 36; i++) a[i] = T(0);
  }
  T& operator()(int i, int j) { return a[i * c + j]; }
  const T& operator()(int i, int j) const { return a[i * c + j]; }
};

// Coefficient-based product: every entry is a left-to-right dot product (what Eigen's
// lazy product does for the small matrices of src/icet.cpp:315-338).
template <class T>
Mx<T> mul(const Mx<T>& A, const Mx<T>& B) {
  Mx<T> C(A.r, B.c);
  for (int i = 0; i < A.r; i++)
    for (int j = 0; j < B.c; j++) {
      T s = A(i, 0) * B(0, j);
      for (int k = 1; k < A.c; k++) s = s + A(i, k) * B(k, j);
      C(i, j) = s;
    }
  return C;
}
template <class T>
Mx<T> transpose(const Mx<T>& A) {
  Mx<T> C(A.c, A.r);
  for (int i = 0; i < A.r; i++)
    for (int j = 0; j < A.c; j++) C(j, i) = A(i, j);
  return C;
}

// ---------------------------------------------------------------------------------------------
// Householder pieces (Eigen/src/Householder/Householder.h)
// ---------------------------------------------------------------------------------------------
// makeHouseholder on the vector v[0..len): returns tau, beta and overwrites v[1..] with the
// essential part (v[0] is left untouched; callers store beta there).
template <class T>
void make_householder(T* v, int len, int stride, T& tau, T& beta) {
  T tailSqNorm = T(0);
  for (int i = 1; i < len; i++) {
    T t = v[i * stride];
    tailSqNorm = (i == 1) ? t * t : tailSqNorm + t * t;
  }
  T c0 = v[0];
  const T tol = std::numeric_limits<T>::min();
  if (len == 1 || tailSqNorm <= tol) {
    tau = T(0);
    beta = c0;
    for (int i = 1; i < len; i++) v[i * stride] = T(0);
  } else {
    beta = std::sqrt(c0 * c0 + tailSqNorm);
    if (c0 >= T(0)) beta = -beta;
    for (int i = 1; i < len; i++) v[i * stride] = v[i * stride] / (c0 - beta);
    tau = (beta - c0) / beta;
  }
}

// ---------------------------------------------------------------------------------------------
// CompleteOrthogonalDecomposition<MatrixXf>::pseudoInverse()
// (Eigen/src/QR/ColPivHouseholderQR.h computeInPlace + CompleteOrthogonalDecomposition.h
//  computeInPlace / _solve_impl applied to the identity).  Used by the reference at
//  src/icet.cpp:320-321 (3x3), :410-411 (6x6), :428-429 (k x 6).
// ---------------------------------------------------------------------------------------------
template <class T>
Mx<T> cod_pinv(const Mx<T>& A, int* rank_out = nullptr) {
  const int rows = A.r, cols = A.c, size = std::min(rows, cols);
  // rank decisions always use the reference's fp32 epsilon, also in the double twin
  const T eps_rank = (T)std::numeric_limits<float>::epsilon();
  Mx<T> qr = A;
  T hCoeffs[6], colNormsUpdated[6], colNormsDirect[6];
  int transp[6];
  auto colnorm = [&](int j, int from) {
    T s = T(0);
    bool first = true;
    for (int i = from; i < rows; i++) {
      T t = qr(i, j);
      s = first ? t * t : s + t * t;
      first = false;
    }
    return std::sqrt(s);
  };
  T maxnorm = T(0);
  for (int k = 0; k < cols; k++) {
    colNormsDirect[k] = colnorm(k, 0);
    colNormsUpdated[k] = colNormsDirect[k];
    if (k == 0 || colNormsUpdated[k] > maxnorm) maxnorm = colNormsUpdated[k];
  }
  T th = maxnorm * eps_rank;
  const T threshold_helper = (th * th) / T(rows);
  const T norm_downdate_threshold = std::sqrt(std::numeric_limits<T>::epsilon());
  int nonzero_pivots = size;
  T maxpivot = T(0);
  for (int k = 0; k < size; k++) {
    int biggest = k;
    T big = colNormsUpdated[k];
    for (int j = k + 1; j < cols; j++)
      if (colNormsUpdated[j] > big) {
        big = colNormsUpdated[j];
        biggest = j;
      }
    T biggest_sq = big * big;
    if (nonzero_pivots == size && biggest_sq < threshold_helper * T(rows - k)) nonzero_pivots = k;
    transp[k] = biggest;
    if (k != biggest) {
      for (int i = 0; i < rows; i++) std::swap(qr(i, k), qr(i, biggest));
      std::swap(colNormsUpdated[k], colNormsUpdated[biggest]);
      std::swap(colNormsDirect[k], colNormsDirect[biggest]);
    }
    T beta;
    make_householder(&qr(k, k), rows - k, cols, hCoeffs[k], beta);
    qr(k, k) = beta;
    if (std::abs(beta) > maxpivot) maxpivot = std::abs(beta);
    // applyHouseholderOnTheLeft on bottomRightCorner(rows-k, cols-k-1)
    if (rows - k == 1) {
      for (int j = k + 1; j < cols; j++) qr(k, j) = qr(k, j) * (T(1) - hCoeffs[k]);
    } else if (hCoeffs[k] != T(0)) {
      for (int j = k + 1; j < cols; j++) {
        T t = T(0);
        bool first = true;
        for (int i = k + 1; i < rows; i++) {
          T pr = qr(i, k) * qr(i, j);
          t = first ? pr : t + pr;
          first = false;
        }
        t = t + qr(k, j);

        qr(k, j) = qr(k, j) - hCoeffs[k] * t;
        for (int i = k + 1; i < rows; i++) qr(i, j) = qr(i, j) - (hCoeffs[k] * qr(i, k)) * t;
      }
    }
    for (int j = k + 1; j < cols; j++) {
      if (colNormsUpdated[j] != T(0)) {
        T t = std::abs(qr(k, j)) / colNormsUpdated[j];
        t = (T(1) + t) * (T(1) - t);
        t = t < T(0) ? T(0) : t;
        T ratio = colNormsUpdated[j] / colNormsDirect[j];
        T t2 = t * (ratio * ratio);
        if (t2 <= norm_downdate_threshold) {
          colNormsDirect[j] = colnorm(j, k + 1);
          colNormsUpdated[j] = colNormsDirect[j];
        } else {
          colNormsUpdated[j] = colNormsUpdated[j] * std::sqrt(t);
        }
      }
    }
  }
  // rank(): threshold = eps * diagonalSize
  int rank = 0;
  {
    const T premult = std::abs(maxpivot) * (eps_rank * T(size));
    for (int i = 0; i < nonzero_pivots; i++)
      if (std::abs(qr(i, i)) > premult) rank++;
  }
  if (rank_out) *rank_out = rank;
  Mx<T> pinv(cols, rows);
  if (rank == 0) return pinv;

  // COD: reduce [R11 R12] to [T11 0] with reflectors applied from the right
  T zCoeffs[6];
  if (rank < cols) {
    for (int k = rank - 1; k >= 0; --k) {
      if (k != rank - 1)
        for (int i = 0; i <= k; i++) std::swap(qr(i, k), qr(i, rank - 1));
      // row k, entries [rank-1, rank, ..., cols-1]  (length cols-rank+1)
      T beta;
      make_householder(&qr(k, rank - 1), cols - rank + 1, 1, zCoeffs[k], beta);
      qr(k, rank - 1) = beta;
      if (k > 0 && zCoeffs[k] != T(0)) {
        // applyHouseholderOnTheRight to topRightCorner(k, cols-rank+1)
        for (int i = 0; i < k; i++) {
          T t = T(0);
          bool first = true;
          for (int j = rank; j < cols; j++) {
            T pr = qr(i, j) * qr(k, j);
            t = first ? pr : t + pr;
            first = false;
          }
          t = t + qr(i, rank - 1);
          qr(i, rank - 1) = qr(i, rank - 1) - zCoeffs[k] * t;
          for (int j = rank; j < cols; j++) qr(i, j) = qr(i, j) - (zCoeffs[k] * t) * qr(k, j);
        }
      }
      if (k != rank - 1)
        for (int i = 0; i <= k; i++) std::swap(qr(i, k), qr(i, rank - 1));
    }
  }
  // c = Q^T * I  (apply the first `rank` reflectors, k = 0 first)
  Mx<T> c(rows, rows);
  for (int i = 0; i < rows; i++) c(i, i) = T(1);
  for (int k = 0; k < rank; k++) {
    if (rows - k == 1) {
      for (int j = 0; j < rows; j++) c(k, j) = c(k, j) * (T(1) - hCoeffs[k]);
    } else if (hCoeffs[k] != T(0)) {
      for (int j = 0; j < rows; j++) {
        T t = T(0);
        bool first = true;
        for (int i = k + 1; i < rows; i++) {
          T pr = qr(i, k) * c(i, j);
          t = first ? pr : t + pr;
          first = false;
        }
        t = t + c(k, j);
        c(k, j) = c(k, j) - hCoeffs[k] * t;
        for (int i = k + 1; i < rows; i++) c(i, j) = c(i, j) - (hCoeffs[k] * qr(i, k)) * t;
      }
    }
  }
  // y(0:rank) = T11^{-1} c(0:rank)   (upper triangular back substitution)
  Mx<T> y(cols, rows);
  for (int j = 0; j < rows; j++) {
    for (int i = rank - 1; i >= 0; --i) {
      T s = c(i, j);
      for (int k = i + 1; k < rank; k++) s = s - qr(i, k) * y(k, j);
      y(i, j) = s / qr(i, i);
    }
  }
  if (rank < cols) {
    // applyZAdjointOnTheLeftInPlace
    for (int k = 0; k < rank; k++) {
      if (k != rank - 1)
        for (int j = 0; j < rows; j++) std::swap(y(k, j), y(rank - 1, j));
      // middleRows(rank-1, cols-rank+1).applyHouseholderOnTheLeft(essential = qr.row(k).tail(cols-rank))
      if (zCoeffs[k] != T(0)) {
        for (int j = 0; j < rows; j++) {
          T t = T(0);
          bool first = true;
          for (int i = rank; i < cols; i++) {
            T pr = qr(k, i) * y(i, j);
            t = first ? pr : t + pr;
            first = false;
          }
          t = t + y(rank - 1, j);
          y(rank - 1, j) = y(rank - 1, j) - zCoeffs[k] * t;
          for (int i = rank; i < cols; i++) y(i, j) = y(i, j) - (zCoeffs[k] * qr(k, i)) * t;
        }
      }
      if (k != rank - 1)
        for (int j = 0; j < rows; j++) std::swap(y(k, j), y(rank - 1, j));
    }
  }
  // undo the column permutation: x = P * y
  int perm[6];
  for (int i = 0; i < cols; i++) perm[i] = i;
  for (int k = 0; k < size; k++) std::swap(perm[k], perm[transp[k]]);
  for (int k = 0; k < cols; k++)
    for (int j = 0; j < rows; j++) pinv(perm[k], j) = y(k, j);
  return pinv;
}

// ---------------------------------------------------------------------------------------------
// SelfAdjointEigenSolver (Eigen/src/Eigenvalues/SelfAdjointEigenSolver.h, Tridiagonalization.h,
// Jacobi/Jacobi.h).  fp32 like the reference: src/icet.cpp:181-184 (3x3), :455-458 (6x6).
// ---------------------------------------------------------------------------------------------
template <class T>
inline T eig_hypot(T x, T y) {  // numext::hypot (positive_real_hypot)
  x = std::abs(x);
  y = std::abs(y);
  T p = std::max(x, y);
  if (p == T(0)) return T(0);
  T qp = std::min(y, x) / p;
  return p * std::sqrt(T(1) + qp * qp);
}

template <class T>
inline void make_givens(T p, T q, T& c, T& s) {  // JacobiRotation::makeGivens (real case)
  if (q == T(0)) {
    c = p < T(0) ? T(-1) : T(1);
    s = T(0);
  } else if (p == T(0)) {
    c = T(0);
    s = q < T(0) ? T(1) : T(-1);
  } else if (std::abs(p) > std::abs(q)) {
    T t = q / p;
    T u = std::sqrt(T(1) + t * t);
    if (p < T(0)) u = -u;
    c = T(1) / u;
    s = -t * c;
  } else {
    T t = p / q;
    T u = std::sqrt(T(1) + t * t);
    if (q < T(0)) u = -u;
    s = -T(1) / u;
    c = -t * s;
  }
}

// tridiagonal_qr_step; Q is n x n, Q(i,j) = q[i*n + j]
template <class T>
void tridiagonal_qr_step(T* diag, T* subdiag, int start, int end, T* q, int n, int flavor) {
  T td = (diag[end - 1] - diag[end]) * T(0.5);
  T e = subdiag[end - 1];
  T mu = diag[end];
  if (td == T(0)) {
    mu -= std::abs(e);
  } else if (e != T(0) || flavor == 0 /* 3.3.x rule */) {
    T e2 = e * e;
    T h = eig_hypot(td, e);
    if (e2 == T(0))
      mu -= (e / (td + (td > T(0) ? T(1) : T(-1)))) * (e / h);
    else
      mu -= e2 / (td + (td > T(0) ? h : -h));
  }
  T x = diag[start] - mu;
  T z = subdiag[start];
  for (int k = start; k < end && (flavor == 0 /* 3.3.x rule */ || z != T(0)); ++k) {
    T c, s;
    make_givens(x, z, c, s);
    T sdk = s * diag[k] + c * subdiag[k];
    T dkp1 = s * subdiag[k] + c * diag[k + 1];
    diag[k] = c * (c * diag[k] - s * subdiag[k]) - s * (c * subdiag[k] - s * diag[k + 1]);
    diag[k + 1] = s * sdk + c * dkp1;
    subdiag[k] = c * sdk - s * dkp1;
    if (k > start) subdiag[k - 1] = c * subdiag[k - 1] - s * z;
    x = subdiag[k];
    if (k < end - 1) {
      z = -s * subdiag[k + 1];
      subdiag[k + 1] = c * subdiag[k + 1];
    }
    // Q = Q * G  (applyOnTheRight(k, k+1, rot)); skipped by Eigen when the rotation is identity
    if (!(c == T(1) && s == T(0))) {
      for (int i = 0; i < n; i++) {
        T xi = q[i * n + k], yi = q[i * n + k + 1];
        q[i * n + k] = c * xi - s * yi;
        q[i * n + k + 1] = s * xi + c * yi;
      }
    }
  }
}

// computeFromTridiagonal_impl + the eigenvalue selection sort
template <class T>
void compute_from_tridiagonal(T* diag, T* subdiag, T* q, int n, int flavor) {
  int end = n - 1, start = 0, iter = 0;
  const int maxIterations = 30;
  const T considerAsZero = std::numeric_limits<T>::min();
  const T precision = T(2) * std::numeric_limits<T>::epsilon();
  const T precision_inv = T(1) / std::numeric_limits<T>::epsilon();
  while (end > 0) {
    for (int i = start; i < end; ++i) {
      if (flavor == 0 /* 3.3.x rule */) {
        // isMuchSmallerThan(|e_i|, |d_i|+|d_i+1|, 2 eps)  ||  |e_i| <= min
        if (std::abs(subdiag[i]) <= (std::abs(diag[i]) + std::abs(diag[i + 1])) * precision ||
            std::abs(subdiag[i]) <= considerAsZero)
          subdiag[i] = T(0);
      } else {
        if (std::abs(subdiag[i]) < considerAsZero) {
          subdiag[i] = T(0);
        } else {
          const T scaled = precision_inv * subdiag[i];
          if (scaled * scaled <= (std::abs(diag[i]) + std::abs(diag[i + 1]))) subdiag[i] = T(0);
        }
      }
    }
    while (end > 0 && subdiag[end - 1] == T(0)) end--;
    if (end <= 0) break;
    iter++;
    if (iter > maxIterations * n) break;
    start = end - 1;
    while (start > 0 && subdiag[start - 1] != T(0)) start--;
    tridiagonal_qr_step(diag, subdiag, start, end, q, n, flavor);
  }
  if (iter <= maxIterations * n) {
    for (int i = 0; i < n - 1; ++i) {
      int k = 0;
      T m = diag[i];
      for (int j = 1; j < n - i; j++)
        if (diag[i + j] < m) {
          m = diag[i + j];
          k = j;
        }
      if (k > 0) {
        std::swap(diag[i], diag[k + i]);
        for (int r = 0; r < n; r++) std::swap(q[r * n + i], q[r * n + k + i]);
      }
    }
  }
}

// SelfAdjointEigenSolver<Matrix3f>::compute (iterative path, NOT computeDirect)
template <class T>
void eig3(const T A[9], int flavor, T evals[3], T V[9]) {
  // mat = lower triangle of A, scaled into [-1, 1]
  T m00 = A[0], m10 = A[3], m11 = A[4], m20 = A[6], m21 = A[7], m22 = A[8];
  T scale = std::max({std::abs(m00), std::abs(m10), std::abs(m11), std::abs(m20), std::abs(m21),
                      std::abs(m22)});
  if (scale == T(0)) scale = T(1);
  m00 /= scale; m10 /= scale; m11 /= scale; m20 /= scale; m21 /= scale; m22 /= scale;
  T diag[3], sub[2];
  // tridiagonalization_inplace_selector<MatrixType,3,false>
  const T tol = std::numeric_limits<T>::min();
  diag[0] = m00;
  T v1norm2 = m20 * m20;
  if (v1norm2 <= tol) {
    diag[1] = m11;
    diag[2] = m22;
    sub[0] = m10;
    sub[1] = m21;
    for (int i = 0; i < 9; i++) V[i] = T(0);
    V[0] = V[4] = V[8] = T(1);
  } else {
    T beta = std::sqrt(m10 * m10 + v1norm2);
    T invBeta = T(1) / beta;
    T m01 = m10 * invBeta;
    T m02 = m20 * invBeta;
    T q = T(2) * m01 * m21 + m02 * (m22 - m11);
    diag[1] = m11 + m02 * q;
    diag[2] = m22 - m02 * q;
    sub[0] = beta;
    sub[1] = m21 - m01 * q;
    V[0] = 1; V[1] = 0;   V[2] = 0;
    V[3] = 0; V[4] = m01; V[5] = m02;
    V[6] = 0; V[7] = m02; V[8] = -m01;
  }
  compute_from_tridiagonal(diag, sub, V, 3, flavor);
  for (int i = 0; i < 3; i++) evals[i] = diag[i] * scale;
}

// SelfAdjointEigenSolver<MatrixXf>::compute for a dynamic n x n (n <= 6) matrix:
// Householder tridiagonalisation (Tridiagonalization.h tridiagonalization_inplace) + QR.
template <class T>
void eigsym(const T* Ain, int n, int flavor, T* evals, T* V) {
  T A[36];
  // lower triangle, scaled
  T scale = T(0);
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) scale = std::max(scale, std::abs(Ain[i * n + j]));
  if (scale == T(0)) scale = T(1);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) A[i * n + j] = (j <= i) ? Ain[i * n + j] / scale : T(0);
  if (n == 1) {
    evals[0] = Ain[0];
    V[0] = T(1);
    return;
  }
  T hCoeffs[6];
  for (int i = 0; i < n - 1; ++i) {
    int rem = n - i - 1;
    T beta, h;
    make_householder(&A[(i + 1) * n + i], rem, n, h, beta);
    A[(i + 1) * n + i] = T(1);
    // v = A.col(i).tail(rem); p = h * (A22.selfadjointView<Lower>() * v)
    T p[6], v[6];
    for (int r = 0; r < rem; r++) v[r] = A[(i + 1 + r) * n + i];
    for (int r = 0; r < rem; r++) {
      T s = T(0);
      for (int c = 0; c < rem; c++) {
        int rr = i + 1 + r, cc = i + 1 + c;
        T a = (cc <= rr) ? A[rr * n + cc] : A[cc * n + rr];
        s = s + a * (h * v[c]);
      }
      p[r] = s;
    }
    T dot = T(0);
    for (int r = 0; r < rem; r++) dot = dot + p[r] * v[r];
    T alpha = h * T(-0.5) * dot;
    for (int r = 0; r < rem; r++) p[r] = p[r] + alpha * v[r];
    // rankUpdate(v, p, -1): A22 -= v p^T + p v^T  (lower triangle)
    for (int r = 0; r < rem; r++)
      for (int c = 0; c <= r; c++)
        A[(i + 1 + r) * n + (i + 1 + c)] =
            A[(i + 1 + r) * n + (i + 1 + c)] - (v[r] * p[c] + p[r] * v[c]);
    A[(i + 1) * n + i] = beta;
    hCoeffs[i] = h;
  }
  T diag[6], sub[5];
  for (int i = 0; i < n; i++) diag[i] = A[i * n + i];
  for (int i = 0; i < n - 1; i++) sub[i] = A[(i + 1) * n + i];
  // Q = H_0 H_1 ... H_{n-2}  (HouseholderSequence, shift 1) evaluated to dense
  for (int i = 0; i < n * n; i++) V[i] = T(0);
  for (int i = 0; i < n; i++) V[i * n + i] = T(1);
  for (int k = n - 2; k >= 0; --k) {
    int corner = n - k - 1;  // rows/cols k+1 .. n-1
    int o = k + 1;
    // applyHouseholderOnTheLeft(essential = A.col(k).tail(corner-1), tau = hCoeffs[k])
    if (corner == 1) {
      V[o * n + o] = V[o * n + o] * (T(1) - hCoeffs[k]);
    } else if (hCoeffs[k] != T(0)) {
      for (int j = o; j < n; j++) {
        T t = T(0);
        for (int r = 1; r < corner; r++) t = t + A[(o + r) * n + k] * V[(o + r) * n + j];
        t = t + V[o * n + j];
        V[o * n + j] = V[o * n + j] - hCoeffs[k] * t;
        for (int r = 1; r < corner; r++)
          V[(o + r) * n + j] = V[(o + r) * n + j] - (hCoeffs[k] * A[(o + r) * n + k]) * t;
      }
    }
  }
  compute_from_tridiagonal(diag, sub, V, n, flavor);
  for (int i = 0; i < n; i++) evals[i] = diag[i] * scale;
}

}  // namespace eigen_restated

#endif
