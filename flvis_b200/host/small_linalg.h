// Tiny dense linear algebra for the host layer (the reference leans on Eigen / OpenCV for these): cyclic Jacobi
// eigen-decomposition of small symmetric matrices, null vectors, 3x3 helpers.  Header-only, no dependencies.
#pragma once
#include <cmath>
#include <vector>

namespace flv {

// eigen-decomposition of a symmetric n x n matrix A (row-major, destroyed): A V = V diag(w); V row-major columns
inline void jacobi_eigen(std::vector<double>& A, int n, std::vector<double>& w, std::vector<double>& V) {
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0, diag = 0;
    for (int i = 0; i < n; ++i) {
      diag += A[(size_t)i * n + i] * A[(size_t)i * n + i];
      for (int j = i + 1; j < n; ++j) off += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    }
    if (off <= 1e-30 * (diag + 1e-300)) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[(size_t)p * n + q];
        if (std::fabs(apq) < 1e-300) continue;
        const double theta = (A[(size_t)q * n + q] - A[(size_t)p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
          A[(size_t)k * n + p] = c * akp - s * akq; A[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
          A[(size_t)p * n + k] = c * apk - s * aqk; A[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
          V[(size_t)k * n + p] = c * vkp - s * vkq; V[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  w.resize(n);
  for (int i = 0; i < n; ++i) w[i] = A[(size_t)i * n + i];
}

// unit vector minimising |M x| for an m x n matrix M (row-major): eigenvector of M^T M with the smallest eigenvalue
inline std::vector<double> null_vector(const std::vector<double>& M, int m, int n) {
  std::vector<double> G((size_t)n * n, 0.0), w, V;
  for (int r = 0; r < m; ++r)
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) G[(size_t)i * n + j] += M[(size_t)r * n + i] * M[(size_t)r * n + j];
  jacobi_eigen(G, n, w, V);
  int best = 0;
  for (int i = 1; i < n; ++i) if (w[i] < w[best]) best = i;
  std::vector<double> x(n);
  for (int i = 0; i < n; ++i) x[i] = V[(size_t)i * n + best];
  return x;
}

inline void mat3_mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
inline double mat3_det(const double* m) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

// SVD of a 3x3 matrix through the eigen-decomposition of M^T M:  M = U diag(s) V^T, s descending
inline void svd3(const double* M, double* U, double* s, double* V) {
  std::vector<double> G(9, 0.0), w, Vv;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) G[3 * i + j] += M[3 * k + i] * M[3 * k + j];
  jacobi_eigen(G, 3, w, Vv);
  int idx[3] = {0, 1, 2};
  for (int a = 0; a < 3; ++a)
    for (int b = a + 1; b < 3; ++b)
      if (w[idx[b]] > w[idx[a]]) { int t = idx[a]; idx[a] = idx[b]; idx[b] = t; }
  for (int c = 0; c < 3; ++c) {
    s[c] = std::sqrt(std::fmax(w[idx[c]], 0.0));
    for (int r = 0; r < 3; ++r) V[3 * r + c] = Vv[3 * r + idx[c]];
  }
  for (int c = 0; c < 3; ++c) {
    double u[3] = {0, 0, 0};
    for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) u[r] += M[3 * r + k] * V[3 * k + c];
    double n = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    if (n < 1e-300) { u[0] = u[1] = u[2] = 0; u[c] = 1; n = 1; }
    for (int r = 0; r < 3; ++r) U[3 * r + c] = u[r] / n;
  }
}

}  // namespace flv
