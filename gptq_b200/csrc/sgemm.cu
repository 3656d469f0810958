// fp32 GEMM on the CUDA cores: 128x128x16 tiles, 8x8 register micro-tiles, register-staged
// double buffering.  Exact IEEE fp32 FMA accumulation -- the "fp32-accurate" fallback used by the
// inverse-Cholesky chain (reference gptq.py:177-179 runs LAPACK in fp32) and by the lazy block
// update W[:, i2:] -= Err1 @ Hinv[i1:i2, i2:] (gptq.py:276; cuBLAS sgemm with TF32 off in the
// reference, gptq.py:18-19).
#include "sgemm.cuh"

namespace gq {

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 16, SG_PAD = 4;

template <bool TA, bool TB, bool VEC>
__global__ void __launch_bounds__(256)
sgemm_kernel(int M, int N, int Kd, float alpha, const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
             int64_t ldb, float beta, float* __restrict__ C, int64_t ldc, int mode) {
    __shared__ __align__(16) float As[2][SG_BK][SG_BM + SG_PAD];
    __shared__ __align__(16) float Bs[2][SG_BK][SG_BN + SG_PAD];

    const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
    if ((mode & SG_UPPER_TILES) && n0 + SG_BN - 1 < m0) return;     // tile strictly below the diagonal
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;

    int kbeg = 0, kend = Kd;
    if (mode & SG_A_UPPER) kbeg = (m0 / SG_BK) * SG_BK;             // op(A)[m][k] == 0 for k < m0 <= m
    if (mode & SG_B_UPPER) kend = min(Kd, n0 + SG_BN);              // op(B)[k][n] == 0 for k > n
    const int nk = kend > kbeg ? (kend - kbeg + SG_BK - 1) / SG_BK : 0;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 ra[2], rb[2];

    // global -> registers for k-tile kt
    auto gload = [&](int kt) {
        const int k0 = kbeg + kt * SG_BK;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            int v = tid + it * 256;
            // ---- A
            if (!TA) {                      // A[m][k], k contiguous: 128 rows x 4 float4
                int m = v >> 2, k4 = (v & 3) * 4;
                int gm = m0 + m, gk = k0 + k4;
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gm < M) {
                    const float* p = A + (int64_t)gm * lda + gk;
                    if (VEC && gk + 3 < Kd) t = *reinterpret_cast<const float4*>(p);
                    else {
                        if (gk + 0 < Kd) t.x = p[0];
                        if (gk + 1 < Kd) t.y = p[1];
                        if (gk + 2 < Kd) t.z = p[2];
                        if (gk + 3 < Kd) t.w = p[3];
                    }
                }
                ra[it] = t;
            } else {                        // A[k][m], m contiguous: 16 rows x 32 float4
                int k = v >> 5, m4 = (v & 31) * 4;
                int gk = k0 + k, gm = m0 + m4;
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gk < Kd) {
                    const float* p = A + (int64_t)gk * lda + gm;
                    if (VEC && gm + 3 < M) t = *reinterpret_cast<const float4*>(p);
                    else {
                        if (gm + 0 < M) t.x = p[0];
                        if (gm + 1 < M) t.y = p[1];
                        if (gm + 2 < M) t.z = p[2];
                        if (gm + 3 < M) t.w = p[3];
                    }
                }
                ra[it] = t;
            }
            // ---- B
            if (TB) {                       // B[n][k], k contiguous
                int n = v >> 2, k4 = (v & 3) * 4;
                int gn = n0 + n, gk = k0 + k4;
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gn < N) {
                    const float* p = B + (int64_t)gn * ldb + gk;
                    if (VEC && gk + 3 < Kd) t = *reinterpret_cast<const float4*>(p);
                    else {
                        if (gk + 0 < Kd) t.x = p[0];
                        if (gk + 1 < Kd) t.y = p[1];
                        if (gk + 2 < Kd) t.z = p[2];
                        if (gk + 3 < Kd) t.w = p[3];
                    }
                }
                rb[it] = t;
            } else {                        // B[k][n], n contiguous
                int k = v >> 5, n4 = (v & 31) * 4;
                int gk = k0 + k, gn = n0 + n4;
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gk < Kd) {
                    const float* p = B + (int64_t)gk * ldb + gn;
                    if (VEC && gn + 3 < N) t = *reinterpret_cast<const float4*>(p);
                    else {
                        if (gn + 0 < N) t.x = p[0];
                        if (gn + 1 < N) t.y = p[1];
                        if (gn + 2 < N) t.z = p[2];
                        if (gn + 3 < N) t.w = p[3];
                    }
                }
                rb[it] = t;
            }
        }
    };
    // registers -> smem buffer b
    auto sstore = [&](int b) {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            int v = tid + it * 256;
            if (!TA) {
                int m = v >> 2, k4 = (v & 3) * 4;
                As[b][k4 + 0][m] = ra[it].x; As[b][k4 + 1][m] = ra[it].y;
                As[b][k4 + 2][m] = ra[it].z; As[b][k4 + 3][m] = ra[it].w;
            } else {
                int k = v >> 5, m4 = (v & 31) * 4;
                *reinterpret_cast<float4*>(&As[b][k][m4]) = ra[it];
            }
            if (TB) {
                int n = v >> 2, k4 = (v & 3) * 4;
                Bs[b][k4 + 0][n] = rb[it].x; Bs[b][k4 + 1][n] = rb[it].y;
                Bs[b][k4 + 2][n] = rb[it].z; Bs[b][k4 + 3][n] = rb[it].w;
            } else {
                int k = v >> 5, n4 = (v & 31) * 4;
                *reinterpret_cast<float4*>(&Bs[b][k][n4]) = rb[it];
            }
        }
    };

    if (nk > 0) {
        gload(0);
        sstore(0);
    }
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int b = kt & 1;
        if (kt + 1 < nk) gload(kt + 1);
#pragma unroll
        for (int k = 0; k < SG_BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[b][k][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[b][k][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[b][k][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[b][k][64 + tx * 4]);
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nk) sstore(b ^ 1);
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (gm >= M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            int gn = n0 + jh * 64 + tx * 4;
            float* p = C + (int64_t)gm * ldc + gn;
            float v[4] = {acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
            if (VEC && gn + 3 < N) {
                float4 o;
                if (beta != 0.f) {
                    float4 c = *reinterpret_cast<const float4*>(p);
                    o.x = fmaf(alpha, v[0], beta * c.x); o.y = fmaf(alpha, v[1], beta * c.y);
                    o.z = fmaf(alpha, v[2], beta * c.z); o.w = fmaf(alpha, v[3], beta * c.w);
                } else {
                    o.x = alpha * v[0]; o.y = alpha * v[1]; o.z = alpha * v[2]; o.w = alpha * v[3];
                }
                *reinterpret_cast<float4*>(p) = o;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (gn + j < N) p[j] = (beta != 0.f) ? fmaf(alpha, v[j], beta * p[j]) : alpha * v[j];
            }
        }
    }
}

int sgemm_launch(int transa, int transb, int M, int N, int Kd, float alpha, const float* A, int64_t lda, const float* B,
                 int64_t ldb, float beta, float* C, int64_t ldc, int mode, cudaStream_t st) {
    if (M <= 0 || N <= 0) return 0;
    bool vec = (lda % 4 == 0) && (ldb % 4 == 0) && (ldc % 4 == 0) && ((((uintptr_t)A) & 15) == 0) &&
               ((((uintptr_t)B) & 15) == 0) && ((((uintptr_t)C) & 15) == 0);
    dim3 grid(cdiv(N, SG_BN), cdiv(M, SG_BM));
#define GQ_SG(TA_, TB_)                                                                                              \
    do {                                                                                                             \
        if (vec) sgemm_kernel<TA_, TB_, true><<<grid, 256, 0, st>>>(M, N, Kd, alpha, A, lda, B, ldb, beta, C, ldc, mode); \
        else sgemm_kernel<TA_, TB_, false><<<grid, 256, 0, st>>>(M, N, Kd, alpha, A, lda, B, ldb, beta, C, ldc, mode);    \
    } while (0)
    if (!transa && !transb) GQ_SG(false, false);
    else if (!transa && transb) GQ_SG(false, true);
    else if (transa && !transb) GQ_SG(true, false);
    else GQ_SG(true, true);
#undef GQ_SG
    GQ_LAUNCHED();
    return 0;
}

}  // namespace gq

using namespace gq;

extern "C" int gptq_b200_sgemm(int transa, int transb, int M, int N, int Kd, float alpha, const float* A, int64_t lda,
                               const float* B, int64_t ldb, float beta, float* C, int64_t ldc, void* stream) {
    GQ_ARG(M >= 0, 3, "M<0"); GQ_ARG(N >= 0, 4, "N<0"); GQ_ARG(Kd >= 0, 5, "K<0");
    GQ_ARG(A, 7, "null"); GQ_ARG(B, 9, "null"); GQ_ARG(C, 12, "null");
    GQ_ARG(lda >= (transa ? M : Kd), 8, "lda too small"); GQ_ARG(ldb >= (transb ? Kd : N), 10, "ldb too small");
    GQ_ARG(ldc >= N, 13, "ldc<N");
    return sgemm_launch(transa, transb, M, N, Kd, alpha, A, lda, B, ldb, beta, C, ldc, SG_FULL, as_stream(stream));
}
