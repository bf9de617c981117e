// procrustes.h -- the 3x3 orthogonal-Procrustes solve of utils/eval_utils.py:28-38 (reference), fp64, host + device.
// Shared by metrics.cu (one thread per body runs it on the device) and tests/host_procrustes.cpp, which compiles this very
// header with g++ so that the CPU test-suite checks the arithmetic against numpy's SVD without a GPU.
#pragma once
#include <math.h>
#ifdef __CUDACC__
#define STRAPS_HD __host__ __device__
#else
#define STRAPS_HD
#endif

namespace straps {

// eigen-decomposition of a symmetric 3x3 matrix (cyclic Jacobi, fp64): A = V diag(w) V^T, columns of V orthonormal
STRAPS_HD inline void jacobi_sym3(double A[3][3], double V[3][3], double w[3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {
    const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    const double diag = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
    if (off <= 1e-30 * diag || off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (A[p][q] == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {          // A <- A J
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {          // A <- J^T A
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < 3; ++i) w[i] = A[i][i];
}

// The rotation maximising trace(R K) with det R = +1 (K = X1 X2^T, 3x3), as utils/eval_utils.py:28-38 builds it from
// U, s, V^T = svd(K): R = V Z U^T with Z = diag(1, 1, sign det(U V^T)).  With u1, u2 / v1, v2 the singular vectors of the two
// largest singular values, R = v1 u1^T + v2 u2^T + (v1 x v2)(u1 x u2)^T -- the cross products carry exactly the sign Z adds.
STRAPS_HD inline void procrustes_rotation(const double K[3][3], double R[3][3]) {
  double KtK[3][3], E[3][3], w[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) KtK[i][j] = K[0][i] * K[0][j] + K[1][i] * K[1][j] + K[2][i] * K[2][j];
  jacobi_sym3(KtK, E, w);            // K = U S E^T: right singular vectors = eigenvectors of K^T K
  int i0 = 0, i1 = 1, i2 = 2;        // order the eigenvalues: w[i0] >= w[i1] >= w[i2]
  if (w[i0] < w[i1]) { int t = i0; i0 = i1; i1 = t; }
  if (w[i0] < w[i2]) { int t = i0; i0 = i2; i2 = t; }
  if (w[i1] < w[i2]) { int t = i1; i1 = i2; i2 = t; }
  double e1[3] = {E[0][i0], E[1][i0], E[2][i0]}, e2[3] = {E[0][i1], E[1][i1], E[2][i1]};
  double u1[3], u2[3];
  for (int i = 0; i < 3; ++i) {
    u1[i] = K[i][0] * e1[0] + K[i][1] * e1[1] + K[i][2] * e1[2];
    u2[i] = K[i][0] * e2[0] + K[i][1] * e2[1] + K[i][2] * e2[2];
  }
  double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
  for (int i = 0; i < 3; ++i) u1[i] /= n1;
  const double d = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];      // re-orthogonalise (exactly orthogonal in theory)
  for (int i = 0; i < 3; ++i) u2[i] -= d * u1[i];
  double n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
  for (int i = 0; i < 3; ++i) u2[i] /= n2;
  const double u3[3] = {u1[1] * u2[2] - u1[2] * u2[1], u1[2] * u2[0] - u1[0] * u2[2], u1[0] * u2[1] - u1[1] * u2[0]};
  const double e3[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
  // svd(K) = U s Vh with K = X1 X2^T: the reference's "V" = Vh^T = our E, its "U" = our u; R = V Z U^T
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i][j] = e1[i] * u1[j] + e2[i] * u2[j] + e3[i] * u3[j];
}

}  // namespace straps
