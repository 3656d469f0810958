// Internal fp32 GEMM launcher shared by the Cholesky chain and the lazy block update.
#pragma once
#include "common.cuh"

namespace gq {

enum SgemmMode {          // bit mask
    SG_FULL = 0,
    SG_UPPER_TILES = 1,   // C square: skip tiles that lie strictly below the diagonal
    SG_A_UPPER = 2,       // op(A) is upper triangular (op(A)[m][k] == 0 for k < m): skip leading k
    SG_B_UPPER = 4,       // op(B) is upper triangular (op(B)[k][n] == 0 for k > n): skip trailing k
    SG_NO_PDL = 8,        // the preceding stream operation is not one of our kernels: launch fully serialized
};

// C[M,N] <- alpha * op(A) op(B) + beta * C   (row-major, fp32, IEEE FMA accumulation)
int sgemm_launch(int transa, int transb, int M, int N, int Kd, float alpha, const float* A, int64_t lda,
                 const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int mode, cudaStream_t st);

}  // namespace gq

namespace gq {
// Tensor-core (3xTF32) variant with the same contract (gemm_tc.cu); needs lda/ldb % 4 == 0 and 16-byte aligned A, B.
bool gemm_tc_supported(const float* A, int64_t lda, const float* B, int64_t ldb, const float* C, int64_t ldc);
int gemm_tc_launch(int transa, int transb, int M, int N, int Kd, float alpha, const float* A, int64_t lda,
                   const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int mode, cudaStream_t st);
// Small-tile CUDA-core variant (sgemm_small.cu) for products whose M and N are at most a couple of hundred: the low
// levels of the chain recursion, where the tensor-core kernel's fixed cost dominates.
int sgemm_small_launch(int transa, int transb, int M, int N, int Kd, float alpha, const float* A, int64_t lda,
                       const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int mode, cudaStream_t st);
// Backend selection used by the chain and the lazy update: 1 = tensor cores when supported (default), 0 = CUDA cores.
extern int g_gemm_backend;
int gemm_fp32(int transa, int transb, int M, int N, int Kd, float alpha, const float* A, int64_t lda,
              const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int mode, cudaStream_t st);
}  // namespace gq
