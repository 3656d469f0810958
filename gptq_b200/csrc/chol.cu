// (2) Inverse-Cholesky chain in fp32 on device.
//
// The reference (gptq.py:177-179) computes  L = chol(H);  Hinv = cholesky_inverse(L);
// U = chol(Hinv, upper)  and afterwards only reads U's upper triangle.  U is the unique
// upper-triangular matrix with positive diagonal such that H^-1 = U^T U, i.e. U = R^-1 where
// H = R R^T with R upper triangular (the "UL" / backward Cholesky factorisation).  We compute
// exactly that: one blocked backward Cholesky (1/3 K^3 flops) and one blocked triangular inverse
// (1/3 K^3), half the work of the three LAPACK calls and one fewer ill-conditioned inversion.
// Both are organised as ONE recursion over block ranges so that all the flops sit in large
// square GEMMs (see hinv_rec below) instead of 128-wide panels.
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
chol_diag_kernel(float* __restrict__ Ablk, int64_t lda, float* __restrict__ Dinv, int64_t ldd, int32_t* __restrict__ info, int base_index) {
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
        // warp-uniform trip count (a warp holds rows i0..i0+7): lanes whose row starts later idle
        const int i0 = (tid >> 5) * 8;
        for (int c = i0 + 1; c < CB; ++c) {
            const bool active = c > i;
            float s = 0.f;
            if (active)
                for (int k = i + q; k < c; k += 4) s = fmaf(xi[k], S[k * CP + c], s);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (active && q == 0) xi[c] = __fdiv_rn(-s, S[c * CP + c]);
            __syncwarp();
        }
    }
    __syncthreads();
    for (int idx = tid; idx < CB * CB; idx += 512) {
        int r = idx >> 7, c = idx & 127;
        Ablk[(int64_t)r * lda + c] = S[r * CP + c];
        Dinv[(int64_t)r * ldd + c] = X[r * CP + c];
    }
}

// ------------------------------------------------------------------ recursive driver
// f(a,b) on block range [a,b): on entry A[a:b,a:b] (upper triangle) carries every Schur update from
// blocks >= b; on exit V[a:b,a:b] = R[a:b,a:b]^-1 (upper).  With the split a < m < b
// (1 = [a,m), 2 = [m,b)):
//     f(m,b)                                   V22 = R22^-1
//     Tt  = V22 * A12^T      (= R12^T)         stored in A's unused lower mirror A[m:b, a:m]
//     A11 -= Tt^T * Tt       (upper tiles)     Schur complement
//     f(a,m)                                   V11 = R11^-1
//     S   = V11 * Tt^T       (= V11 R12)       scratch
//     V12 = -S * V22
// Every product is a square-ish fp32 GEMM with triangular-operand k-range skipping.
struct HinvCtx {
    float* A; int64_t lda; float* V; int64_t ldv; float* S; int32_t* info; cudaStream_t st; size_t smem; int rc;
};

static void hinv_rec(HinvCtx& c, int a, int b) {
    if (c.rc) return;
    if (b - a == 1) {
        const int c0 = a * CB;
        chol_diag_kernel<<<1, 512, c.smem, c.st>>>(c.A + (int64_t)c0 * c.lda + c0, c.lda, c.V + (int64_t)c0 * c.ldv + c0,
                                                   c.ldv, c.info, c0);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return;
    }
    const int m = a + (b - a) / 2;
    const int n1 = (m - a) * CB, n2 = (b - m) * CB;
    const int64_t ra = (int64_t)a * CB, rm = (int64_t)m * CB;
    float* A11 = c.A + ra * c.lda + ra;
    float* A12 = c.A + ra * c.lda + rm;
    float* Tt = c.A + rm * c.lda + ra;               // lower mirror, [n2 x n1]
    float* V11 = c.V + ra * c.ldv + ra;
    float* V12 = c.V + ra * c.ldv + rm;
    float* V22 = c.V + rm * c.ldv + rm;
    hinv_rec(c, m, b);
    if (c.rc) return;
    // Tt[n2 x n1] = V22[n2 x n2] * A12^T
    c.rc = sgemm_launch(0, 1, n2, n1, n2, 1.f, V22, c.ldv, A12, c.lda, 0.f, Tt, c.lda, SG_A_UPPER, c.st);
    if (c.rc) return;
    // A11 -= Tt^T Tt
    c.rc = sgemm_launch(1, 0, n1, n1, n2, -1.f, Tt, c.lda, Tt, c.lda, 1.f, A11, c.lda, SG_UPPER_TILES, c.st);
    if (c.rc) return;
    hinv_rec(c, a, m);
    if (c.rc) return;
    // S[n1 x n2] = V11 * Tt^T
    c.rc = sgemm_launch(0, 1, n1, n2, n1, 1.f, V11, c.ldv, Tt, c.lda, 0.f, c.S, n2, SG_A_UPPER, c.st);
    if (c.rc) return;
    // V12 = -S * V22
    c.rc = sgemm_launch(0, 0, n1, n2, n2, -1.f, c.S, n2, V22, c.ldv, 0.f, V12, c.ldv, SG_B_UPPER, c.st);
}

}  // namespace gq

using namespace gq;

extern "C" size_t gptq_b200_hinv_upper_workspace_bytes(int Kp) {
    if (Kp <= 0) return 0;
    size_t nb = (size_t)(Kp / CB);
    size_t h1 = (nb / 2) * CB, h2 = (nb - nb / 2) * CB;
    size_t s = h1 * h2;                                  // largest n1 x n2 scratch (top-level split)
    if (s < (size_t)CB * CB) s = (size_t)CB * CB;
    return s * sizeof(float);
}

extern "C" int gptq_b200_hinv_upper(float* A, int Kp, int64_t lda, float* U, int64_t ldu, int32_t* info, void* ws,
                                    size_t ws_bytes, void* stream) {
    GQ_ARG(A, 1, "null"); GQ_ARG(Kp > 0 && Kp % CB == 0, 2, "Kp must be a positive multiple of 128");
    GQ_ARG(lda >= Kp && lda % 4 == 0, 3, "lda must be >= Kp and a multiple of 4");
    GQ_ARG(U && U != A, 4, "U must be a separate matrix"); GQ_ARG(ldu >= Kp && ldu % 4 == 0, 5, "ldu must be >= Kp and a multiple of 4");
    GQ_ARG(info, 6, "null"); GQ_ARG(ws, 7, "null");
    GQ_ARG(ws_bytes >= gptq_b200_hinv_upper_workspace_bytes(Kp), 8, "workspace too small");
    cudaStream_t st = as_stream(stream);
    const size_t smem = 2 * CB * CP * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
        GQ_CUDA(cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    GQ_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), st));
    GQ_CUDA(cudaMemset2DAsync(U, ldu * sizeof(float), 0, (size_t)Kp * sizeof(float), Kp, st));
    HinvCtx c{A, lda, U, ldu, static_cast<float*>(ws), info, st, smem, 0};
    hinv_rec(c, 0, Kp / CB);
    if (c.rc) return c.rc;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err((int)e, "hinv_upper: launch failed: %s", cudaGetErrorString(e));
    return 0;
}
