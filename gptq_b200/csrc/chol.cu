// (2) Inverse-Cholesky chain in fp32 on device.
//
// The reference (gptq.py:177-179) computes  L = chol(H);  Hinv = cholesky_inverse(L);
// U = chol(Hinv, upper)  and afterwards only reads U's upper triangle.  U is the unique
// upper-triangular matrix with positive diagonal such that H^-1 = U^T U, i.e. U = R^-1 where
// H = R R^T with R upper triangular (the "UL" / backward Cholesky factorisation).  We compute
// exactly that: one blocked backward Cholesky (1/3 K^3 flops) and one blocked triangular inverse
// (1/3 K^3), half the work of the three LAPACK calls and one fewer ill-conditioned inversion.
// SURVEY Appendix B measured 0 code mismatches for this route against the reference.
//
// Block size 128.  Diagonal blocks are factored AND inverted by one CTA in shared memory
// (4 lanes per matrix row, warp shuffles for the dot-product reductions); panels and trailing
// updates are fp32 GEMMs (sgemm.cu).
#include "sgemm.cuh"

namespace gq {

constexpr int CB = 128;            // block size
constexpr int CP = CB + 1;         // smem pitch (conflict-free row-per-thread access)

// A_jj (upper triangle read) -> R_jj in place (strict lower zeroed), Dinv = R_jj^-1 (dense 128x128,
// strict lower zero).  512 threads: thread (i = tid/4, q = tid%4) owns a quarter of row i's dots.
__global__ void __launch_bounds__(512)
chol_diag_kernel(float* __restrict__ Ablk, int64_t lda, float* __restrict__ Dinv, int32_t* __restrict__ info, int base_index) {
    extern __shared__ float sm[];
    float* S = sm;                 // [128][129]  A -> R
    float* X = sm + CB * CP;       // [128][129]  R^-1
    __shared__ float s_piv;
    const int tid = threadIdx.x;
    const int i = tid >> 2, q = tid & 3;

    for (int idx = tid; idx < CB * CB; idx += 512) {
        int r = idx >> 7, c = idx & 127;
        S[r * CP + c] = (c >= r) ? Ablk[(int64_t)r * lda + c] : 0.f;
        X[r * CP + c] = 0.f;
    }
    __syncthreads();

    // ---- backward Cholesky, left-looking: column j uses columns k > j.
    for (int j = CB - 1; j >= 0; --j) {
        float s = 0.f;
        if (i <= j) {
            const float* ri = S + i * CP;
            const float* rj = S + j * CP;
            for (int k = j + 1 + q; k < CB; k += 4) s = fmaf(ri[k], rj[k], s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        float v = 0.f;
        if (i <= j) v = S[i * CP + j] - s;
        if (i == j && q == 0) {
            if (!(v > 0.f)) { if (atomicCAS(info, 0, base_index + j + 1) == 0) {} }
            s_piv = sqrtf(v);
        }
        __syncthreads();
        if (i <= j && q == 0) S[i * CP + j] = (i == j) ? s_piv : __fdiv_rn(v, s_piv);
        __syncthreads();
    }

    // ---- X = R^-1, row i independent: X[i][c] = -(sum_{k=i}^{c-1} X[i][k] R[k][c]) / R[c][c]
    {
        float* xi = X + i * CP;
        if (q == 0) xi[i] = __fdiv_rn(1.f, S[i * CP + i]);
        __syncwarp();
        for (int c = i + 1; c < CB; ++c) {
            float s = 0.f;
            for (int k = i + q; k < c; k += 4) s = fmaf(xi[k], S[k * CP + c], s);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (q == 0) xi[c] = __fdiv_rn(-s, S[c * CP + c]);
            __syncwarp();
        }
    }
    __syncthreads();
    for (int idx = tid; idx < CB * CB; idx += 512) {
        int r = idx >> 7, c = idx & 127;
        Ablk[(int64_t)r * lda + c] = S[r * CP + c];
        Dinv[r * CB + c] = X[r * CP + c];
    }
}

// zero the strict lower triangle of the off-diagonal blocks (diagonal blocks are handled by chol_diag)
__global__ void zero_lower_kernel(float* __restrict__ A, int Kp, int64_t lda) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int i = blockIdx.y;
    if (j < Kp && (j / CB) < (i / CB)) A[(int64_t)i * lda + j] = 0.f;
}

}  // namespace gq

using namespace gq;

extern "C" size_t gptq_b200_hinv_upper_workspace_bytes(int Kp) {
    if (Kp <= 0) return 0;
    size_t nb = (size_t)(Kp / CB);
    return (nb * CB * CB + (size_t)Kp * CB) * sizeof(float);
}

extern "C" int gptq_b200_hinv_upper(float* A, int Kp, int64_t lda, int32_t* info, void* ws, size_t ws_bytes, void* stream) {
    GQ_ARG(A, 1, "null"); GQ_ARG(Kp > 0 && Kp % CB == 0, 2, "Kp must be a positive multiple of 128");
    GQ_ARG(lda >= Kp, 3, "lda<Kp"); GQ_ARG(info, 4, "null"); GQ_ARG(ws, 5, "null");
    GQ_ARG(ws_bytes >= gptq_b200_hinv_upper_workspace_bytes(Kp), 6, "workspace too small");
    cudaStream_t st = as_stream(stream);
    const int nb = Kp / CB;
    float* Dinv = static_cast<float*>(ws);
    float* P = Dinv + (size_t)nb * CB * CB;
    const size_t smem = 2 * CB * CP * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
        GQ_CUDA(cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    GQ_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), st));

    // ---- phase 1: A = R R^T, R upper, from the last block column to the first
    for (int j = nb - 1; j >= 0; --j) {
        const int c0 = j * CB;
        float* Ajj = A + (int64_t)c0 * lda + c0;
        chol_diag_kernel<<<1, 512, smem, st>>>(Ajj, lda, Dinv + (size_t)j * CB * CB, info, c0);
        GQ_LAUNCHED();
        if (j == 0) break;
        // R[0:c0, j] = A[0:c0, j] * R_jj^-T
        int rc = sgemm_launch(0, 1, c0, CB, CB, 1.f, A + c0, lda, Dinv + (size_t)j * CB * CB, CB, 0.f, P, CB, SG_FULL, st);
        if (rc) return rc;
        // A[0:c0, 0:c0] -= R[0:c0, j] R[0:c0, j]^T   (upper tiles)
        rc = sgemm_launch(0, 1, c0, c0, CB, -1.f, P, CB, P, CB, 1.f, A, lda, SG_UPPER_TILES, st);
        if (rc) return rc;
        GQ_CUDA(cudaMemcpy2DAsync(A + c0, lda * sizeof(float), P, CB * sizeof(float), CB * sizeof(float), c0,
                                  cudaMemcpyDeviceToDevice, st));
    }
    zero_lower_kernel<<<dim3(cdiv(Kp, 256), Kp), 256, 0, st>>>(A, Kp, lda);
    GQ_LAUNCHED();

    // ---- phase 2: U = R^-1 in place, block column by block column
    for (int j = 0; j < nb; ++j) {
        const int c0 = j * CB;
        const float* Dj = Dinv + (size_t)j * CB * CB;
        if (j > 0) {
            // T = U[0:c0, 0:c0] * R[0:c0, j]
            int rc = sgemm_launch(0, 0, c0, CB, c0, 1.f, A, lda, A + c0, lda, 0.f, P, CB, SG_A_UPPER, st);
            if (rc) return rc;
            // U[0:c0, j] = -T * R_jj^-1
            rc = sgemm_launch(0, 0, c0, CB, CB, -1.f, P, CB, Dj, CB, 0.f, A + c0, lda, SG_FULL, st);
            if (rc) return rc;
        }
        GQ_CUDA(cudaMemcpy2DAsync(A + (int64_t)c0 * lda + c0, lda * sizeof(float), Dj, CB * sizeof(float),
                                  CB * sizeof(float), CB, cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}
