// TEST INFRASTRUCTURE — see taichi/util.h in this directory.  svd / polar_decomp of the stand-in core.
#pragma once
#include <taichi/util.h>

namespace taichi {
namespace stub_detail {
// Eigen-decomposition of a symmetric n x n matrix (cyclic Jacobi, double): A = Q diag(w) Q^T.
template <int n>
inline void jacobi_eig(double A[n][n], double Q[n][n], double w[n]) {
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Q[i][j] = i == j;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) off += A[p][q] * A[p][q];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) {
        if (A[p][q] == 0) continue;
        double theta = (A[q][q] - A[p][p]) / (2 * A[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
        double c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < n; k++) {  // A <- A J
          double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {  // A <- J^T A
          double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          double qkp = Q[k][p], qkq = Q[k][q];
          Q[k][p] = c * qkp - s * qkq;
          Q[k][q] = s * qkp + c * qkq;
        }
      }
  }
  for (int i = 0; i < n; i++) w[i] = A[i][i];
}

template <int n>
inline double det(const double M[n][n]);
template <>
inline double det<2>(const double M[2][2]) { return M[0][0] * M[1][1] - M[0][1] * M[1][0]; }
template <>
inline double det<3>(const double M[3][3]) {
  return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
         M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
}

// A = U diag(s) V^T with U, V proper rotations, |s| descending, s[n-1] < 0 iff det A < 0.
template <int n>
inline void svd_rot(const double A[n][n], double U[n][n], double s[n], double V[n][n]) {
  double AtA[n][n], w[n];
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { AtA[i][j] = 0; for (int k = 0; k < n; k++) AtA[i][j] += A[k][i] * A[k][j]; }
  jacobi_eig<n>(AtA, V, w);
  int order[n];
  for (int i = 0; i < n; i++) order[i] = i;
  std::sort(order, order + n, [&](int a, int b) { return w[a] > w[b]; });
  double Vs[n][n];
  for (int j = 0; j < n; j++) { s[j] = std::sqrt(std::max(w[order[j]], 0.0)); for (int i = 0; i < n; i++) Vs[i][j] = V[i][order[j]]; }
  if (det<n>(Vs) < 0) for (int i = 0; i < n; i++) Vs[i][n - 1] = -Vs[i][n - 1];
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) V[i][j] = Vs[i][j];
  // U columns: A v_j / s_j, Gram-Schmidt completed for (near-)zero singular values
  for (int j = 0; j < n; j++) {
    double u[n];
    for (int i = 0; i < n; i++) { u[i] = 0; for (int k = 0; k < n; k++) u[i] += A[i][k] * V[k][j]; }
    for (int p = 0; p < j; p++) { double d = 0; for (int i = 0; i < n; i++) d += u[i] * U[i][p]; for (int i = 0; i < n; i++) u[i] -= d * U[i][p]; }
    double len = 0;
    for (int i = 0; i < n; i++) len += u[i] * u[i];
    len = std::sqrt(len);
    if (len < 1e-150) {  // rank deficient: any unit vector orthogonal to the previous columns
      for (int e = 0; e < n && len < 1e-150; e++) {
        for (int i = 0; i < n; i++) u[i] = i == e;
        for (int p = 0; p < j; p++) { double d = U[e][p]; for (int i = 0; i < n; i++) u[i] -= d * U[i][p]; }
        len = 0;
        for (int i = 0; i < n; i++) len += u[i] * u[i];
        len = std::sqrt(len);
      }
    }
    for (int i = 0; i < n; i++) U[i][j] = u[i] / len;
  }
  if (det<n>(U) < 0) {  // make U a rotation; the sign moves to the smallest singular value
    for (int i = 0; i < n; i++) U[i][n - 1] = -U[i][n - 1];
    s[n - 1] = -s[n - 1];
  }
}
}  // namespace stub_detail

template <int n, class T>
inline void svd(const MatrixND<n, T> &A, MatrixND<n, T> &U, MatrixND<n, T> &S, MatrixND<n, T> &V) {
  double a[n][n], u[n][n], s[n], v[n][n];
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) a[i][j] = A[j][i];  // a[row][col]
  stub_detail::svd_rot<n>(a, u, s, v);
  S = MatrixND<n, T>(T(0));
  for (int i = 0; i < n; i++) {
    S[i][i] = (T)s[i];
    for (int j = 0; j < n; j++) { U[j][i] = (T)u[i][j]; V[j][i] = (T)v[i][j]; }
  }
}

template <int n, class T>
inline void polar_decomp(const MatrixND<n, T> &A, MatrixND<n, T> &R, MatrixND<n, T> &S) {
  double a[n][n], u[n][n], s[n], v[n][n];
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) a[i][j] = A[j][i];
  stub_detail::svd_rot<n>(a, u, s, v);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      double r = 0, sm = 0;
      for (int k = 0; k < n; k++) { r += u[i][k] * v[j][k]; sm += v[i][k] * s[k] * v[j][k]; }
      R[j][i] = (T)r;
      S[j][i] = (T)sm;
    }
}
}  // namespace taichi
