// Small-tile fp32 GEMM on the CUDA cores for the LOW levels of the inverse-Cholesky recursion (reference
// gptq.py:177-179): products of 128- and 256-wide blocks.  There the tensor-core kernel (gemm_tc.cu) has one or two
// 128x256 tiles to work on and its fixed cost (TMEM allocation, TMA round trip, 3xTF32 split, drain) dominates:
// 15.7 us for 128^3.  32x32 tiles spread the same product over 16-64 SMs, each doing a few hundred exact IEEE fp32
// FMAs per thread from shared memory.  Same contract as sgemm_launch / gemm_tc_launch (alpha/beta, all four
// transpose combinations, triangular k-range skipping), plus programmatic dependent launch: operands and C are
// produced by the preceding kernels of the stream and are read through L2 (ld.global.cg) after griddepcontrol.wait.
#include "sgemm.cuh"
#include "ptx.cuh"

namespace gq {

constexpr int SS_T = 32;            // tile edge (M, N and k-chunk)
constexpr int SS_P = SS_T + 1;      // shared-memory pitch

// One 32x32 chunk of op(X) into registers: thread t holds 4 consecutive elements along the CONTIGUOUS direction.
//   contiguous along k (ROWMAJOR_K = true):  row = t / 8 (m or n index), k4 = (t % 8) * 4
//   contiguous along mn (false):             k = t / 8, mn4 = (t % 8) * 4
template <bool ROWMAJOR_K>
__device__ __forceinline__ float4 ss_gload(const float* __restrict__ X, int64_t ld, int mn0, int k0, int mn_extent,
                                           int k_extent, bool vec, int tid) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ROWMAJOR_K) {
        const int r = mn0 + (tid >> 3), k = k0 + (tid & 7) * 4;
        if (r < mn_extent) {
            const float* p = X + (int64_t)r * ld + k;
            if (vec && k + 3 < k_extent) t = __ldcg(reinterpret_cast<const float4*>(p));
            else {
                if (k + 0 < k_extent) t.x = __ldcg(p + 0);
                if (k + 1 < k_extent) t.y = __ldcg(p + 1);
                if (k + 2 < k_extent) t.z = __ldcg(p + 2);
                if (k + 3 < k_extent) t.w = __ldcg(p + 3);
            }
        }
    } else {
        const int k = k0 + (tid >> 3), c = mn0 + (tid & 7) * 4;
        if (k < k_extent) {
            const float* p = X + (int64_t)k * ld + c;
            if (vec && c + 3 < mn_extent) t = __ldcg(reinterpret_cast<const float4*>(p));
            else {
                if (c + 0 < mn_extent) t.x = __ldcg(p + 0);
                if (c + 1 < mn_extent) t.y = __ldcg(p + 1);
                if (c + 2 < mn_extent) t.z = __ldcg(p + 2);
                if (c + 3 < mn_extent) t.w = __ldcg(p + 3);
            }
        }
    }
    return t;
}

// registers -> shared tile S[k][mn]
template <bool ROWMAJOR_K>
__device__ __forceinline__ void ss_sstore(float (*S)[SS_P], float4 t, int tid) {
    if (ROWMAJOR_K) {
        const int r = tid >> 3, k = (tid & 7) * 4;
        S[k + 0][r] = t.x; S[k + 1][r] = t.y; S[k + 2][r] = t.z; S[k + 3][r] = t.w;
    } else {
        const int k = tid >> 3, c = (tid & 7) * 4;
        S[k][c + 0] = t.x; S[k][c + 1] = t.y; S[k][c + 2] = t.z; S[k][c + 3] = t.w;
    }
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
sgemm_small_kernel(int M, int N, int Kd, float alpha, const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                   int64_t ldb, float beta, float* __restrict__ C, int64_t ldc, int mode, int vec) {
    __shared__ float As[2][SS_T][SS_P];          // [k][m]
    __shared__ float Bs[2][SS_T][SS_P];          // [k][n]
    const int m0 = blockIdx.y * SS_T, n0 = blockIdx.x * SS_T;
    if ((mode & SG_UPPER_TILES) && n0 + SS_T - 1 < m0) return;      // tile strictly below the diagonal
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;      // outputs rows 2ty, 2ty+1; columns 2tx, 2tx+1
    int kbeg = 0, kend = Kd;
    if (mode & SG_A_UPPER) kbeg = (m0 / SS_T) * SS_T;               // op(A)[m][k] == 0 for k < m0 <= m
    if (mode & SG_B_UPPER) kend = min(Kd, n0 + SS_T);               // op(B)[k][n] == 0 for k > n
    const int nk = kend > kbeg ? (kend - kbeg + SS_T - 1) / SS_T : 0;

    ptx::pdl_launch_dependents();
    ptx::pdl_wait();

    // op(A)[m][k]: !TA -> A[m][k] (k contiguous);  TA -> A[k][m] (m contiguous).  op(B)[k][n]: !TB -> B[k][n]; TB -> B[n][k].
    auto loadA = [&](int kt) { return ss_gload<!TA>(A, lda, m0, kbeg + kt * SS_T, M, Kd, vec != 0, tid); };
    auto loadB = [&](int kt) { return ss_gload<TB>(B, ldb, n0, kbeg + kt * SS_T, N, Kd, vec != 0, tid); };

    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    if (nk > 0) {
        float4 ra = loadA(0), rb = loadB(0);
        ss_sstore<!TA>(As[0], ra, tid);
        ss_sstore<TB>(Bs[0], rb, tid);
        __syncthreads();
        for (int kt = 0; kt < nk; ++kt) {
            const int b = kt & 1;
            if (kt + 1 < nk) { ra = loadA(kt + 1); rb = loadB(kt + 1); }
#pragma unroll
            for (int k = 0; k < SS_T; ++k) {
                const float a0 = As[b][k][2 * ty], a1 = As[b][k][2 * ty + 1];
                const float b0 = Bs[b][k][2 * tx], b1 = Bs[b][k][2 * tx + 1];
                acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
                acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
            }
            if (kt + 1 < nk) {
                ss_sstore<!TA>(As[b ^ 1], ra, tid);
                ss_sstore<TB>(Bs[b ^ 1], rb, tid);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int gm = m0 + 2 * ty + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int gn = n0 + 2 * tx + j;
            if (gn >= N) continue;
            float* o = C + (int64_t)gm * ldc + gn;
            *o = (beta != 0.f) ? fmaf(alpha, acc[i][j], beta * __ldcg(o)) : alpha * acc[i][j];
        }
    }
}

int sgemm_small_launch(int transa, int transb, int M, int N, int Kd, float alpha, const float* A, int64_t lda, const float* B,
                       int64_t ldb, float beta, float* C, int64_t ldc, int mode, cudaStream_t st) {
    if (M <= 0 || N <= 0) return 0;
    const int vec = (lda % 4 == 0) && (ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
    const dim3 grid(cdiv(N, SS_T), cdiv(M, SS_T));
    const bool pdl = !(mode & SG_NO_PDL);
    cudaError_t e;
    if (!transa && !transb) e = launch_pdl(pdl, sgemm_small_kernel<false, false>, grid, dim3(256), 0, st, M, N, Kd, alpha, A, lda, B, ldb, beta, C, ldc, mode, vec);
    else if (!transa && transb) e = launch_pdl(pdl, sgemm_small_kernel<false, true>, grid, dim3(256), 0, st, M, N, Kd, alpha, A, lda, B, ldb, beta, C, ldc, mode, vec);
    else if (transa && !transb) e = launch_pdl(pdl, sgemm_small_kernel<true, false>, grid, dim3(256), 0, st, M, N, Kd, alpha, A, lda, B, ldb, beta, C, ldc, mode, vec);
    else e = launch_pdl(pdl, sgemm_small_kernel<true, true>, grid, dim3(256), 0, st, M, N, Kd, alpha, A, lda, B, ldb, beta, C, ldc, mode, vec);
    if (e != cudaSuccess) return set_err((int)e, "sgemm_small: launch failed: %s", cudaGetErrorString(e));
    GQ_LAUNCHED();
    return 0;
}

}  // namespace gq
