// Host build of csrc/procrustes.h for the CPU test-suite (tests/test_metrics_cpu.py compiles it with g++).
// It checks the SAME source the device kernel compiles -- not a second implementation.
#include "procrustes.h"

extern "C" void host_procrustes_rotation(const double* K9, double* R9) {
  double K[3][3], R[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) K[i][j] = K9[i * 3 + j];
  straps::procrustes_rotation(K, R);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R9[i * 3 + j] = R[i][j];
}

extern "C" void host_jacobi_sym3(const double* A9, double* V9, double* w3) {
  double A[3][3], V[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[i][j] = A9[i * 3 + j];
  straps::jacobi_sym3(A, V, w3);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V9[i * 3 + j] = V[i][j];
}
