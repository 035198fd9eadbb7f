// dense_chol.h — dense FP64 Cholesky solve of the reduced camera system (dense_chol.cu).
#pragma once
#include <cuda_runtime.h>

namespace ppsfm {

// Leading dimension (multiple of 64) for an n x n system with its right-hand side as row n.
int chol_ld(int n);
// Device scratch (in doubles) chol_solve_bordered needs for an n x n system.
size_t chol_work_doubles(int n);
// A: ld x ld row-major; rows [0,n) lower triangle of the SPD matrix, row n = rhs^T, rest zero.
// On return x (n doubles, device) holds the solution and A the factor.  work: chol_work_doubles(n) doubles
// of device scratch.  *status (device int) becomes 1 if a non-positive pivot was met.
// Asynchronous; returns the number of launches.
int chol_solve_bordered(double* A, int n, int ld, double* x, double* work, int* status,
                        cudaStream_t s);

}  // namespace ppsfm
