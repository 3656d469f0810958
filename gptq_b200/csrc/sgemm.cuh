// Internal fp32 GEMM launcher shared by the Cholesky chain and the lazy block update.
#pragma once
#include "common.cuh"

namespace gq {

enum SgemmMode {          // bit mask
    SG_FULL = 0,
    SG_UPPER_TILES = 1,   // C square: skip tiles that lie strictly below the diagonal
    SG_A_UPPER = 2,       // op(A) is upper triangular (op(A)[m][k] == 0 for k < m): skip leading k
    SG_B_UPPER = 4,       // op(B) is upper triangular (op(B)[k][n] == 0 for k > n): skip trailing k
};

// C[M,N] <- alpha * op(A) op(B) + beta * C   (row-major, fp32, IEEE FMA accumulation)
int sgemm_launch(int transa, int transb, int M, int N, int Kd, float alpha, const float* A, int64_t lda,
                 const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int mode, cudaStream_t st);

}  // namespace gq
