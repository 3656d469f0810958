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
// Block size 128.  Diagonal blocks are factored AND inverted by one CTA with one matrix row per
// 8x8 register tile per thread (256 threads); everything else is fp32 GEMMs (sgemm.cu).
#include "sgemm.cuh"
#include "ptx.cuh"
#include <stdlib.h>
#include <mutex>
#include <vector>

namespace gq {

constexpr int CB = 128;            // block size

// A_jj (upper triangle read) -> Dinv = R_jj^-1 (dense 128x128, strict lower zero), A_jj = R_jj R_jj^T.
//
// 256 threads as a 16x16 grid, tid = tx*16 + ty; thread (ty,tx) owns the 8x8 cyclic sub-matrix
//     rows a*16+ty, cols b*16+tx  (a,b = 0..7)
// of the 128x128 block IN REGISTERS, so each of the 128 sequential steps is one rank-1 update of a
// register tile.  The 16 owners of a matrix column are one half-warp, so the pivot reaches them by
// shuffle and a step needs ONE block barrier (the broadcast column is double-buffered in smem).
// The step loop is (outer block index, statically unrolled) x (inner 16, rolled): every register
// index is static and tiles outside the still-active triangle are skipped at compile time.
//   phase 1: right-looking backward Cholesky  A = R R^T  (columns 127 .. 0)
//   phase 2: X = R^-1 by forward column elimination of  X R = I  (columns 0 .. 127), R read from smem
constexpr int CT = CB + 1;         // smem pitch

// 1/sqrt(v): MUFU.RSQ + one Newton step (error < 1 ulp; nothing downstream is bit-matched to LAPACK)
__device__ __forceinline__ float inv_sqrt(float v) {
    float y = rsqrtf(v);
    return y * fmaf(-0.5f * v, y * y, 1.5f);
}

template <int JB>
__device__ __forceinline__ void chol_phase1_block(float (&r)[8][8], float* colb2, float* rd, int32_t* info, int base_index,
                                                  int tx, int ty) {
#pragma unroll 1
    for (int jt = 15; jt >= 0; --jt) {
        const int j = JB * 16 + jt;
        float* colb = colb2 + (jt & 1) * CB;
        if ((int)(threadIdx.x >> 5) == (jt >> 1)) {          // the warp holding column j (warp-uniform)
            const float v = __shfl_sync(0xffffffffu, r[JB][JB], ((jt & 1) << 4) | jt);     // pivot (j,j)
            const float rinv = inv_sqrt(v);
            if (tx == jt) {
                if (ty == jt) {
                    if (!(v > 0.f)) atomicCAS(info, 0, base_index + j + 1);
                    rd[j] = rinv;
                }
#pragma unroll
                for (int a = 0; a <= JB; ++a) {
                    const int row = a * 16 + ty;
                    float c = r[a][JB] * rinv;                 // row == j: v * rsqrt(v) = sqrt(v)
                    if (row > j) c = 0.f;
                    r[a][JB] = c;
                    colb[row] = c;
                }
            }
        }
        __syncthreads();
        float cr[JB + 1], cc[JB + 1];
#pragma unroll
        for (int a = 0; a <= JB; ++a) cr[a] = colb[a * 16 + ty];
#pragma unroll
        for (int b = 0; b <= JB; ++b) cc[b] = colb[b * 16 + tx];
#pragma unroll
        for (int a = 0; a <= JB; ++a)
#pragma unroll
            for (int b = 0; b <= JB; ++b) {
                const bool act = (a < JB || ty < jt) && (b < JB || tx < jt);     // row < j && col < j
                if (act) r[a][b] = fmaf(-cr[a], cc[b], r[a][b]);
            }
    }
}

template <int CBI>
__device__ __forceinline__ void chol_phase2_block(float (&x)[8][8], float* colb2, const float* S, const float* rd, int tx, int ty) {
#pragma unroll 1
    for (int ct = 0; ct < 16; ++ct) {
        const int c = CBI * 16 + ct;
        float* colb = colb2 + (ct & 1) * CB;
        if (tx == ct) {                                  // owners of column c of X: rows a*16+ty, a <= CBI
            const float dinv = rd[c];                    // 1 / R[c][c] from phase 1
#pragma unroll
            for (int a = 0; a <= CBI; ++a) {
                const float v = x[a][CBI] * dinv;
                x[a][CBI] = v;
                colb[a * 16 + ty] = v;
            }
        }
        __syncthreads();
        float xc[CBI + 1], rr[8];
#pragma unroll
        for (int a = 0; a <= CBI; ++a) xc[a] = colb[a * 16 + ty];
#pragma unroll
        for (int b = CBI; b < 8; ++b) rr[b] = S[c * CT + b * 16 + tx];          // row c of R
#pragma unroll
        for (int a = 0; a <= CBI; ++a)
#pragma unroll
            for (int b = CBI; b < 8; ++b) {
                const bool act = (a < CBI || ty <= ct) && (b > CBI || tx > ct);    // row <= c && col > c
                if (act) x[a][b] = fmaf(-xc[a], rr[b], x[a][b]);
            }
    }
}

__global__ void __launch_bounds__(256, 1)
chol_diag_kernel(const float* __restrict__ Ablk, int64_t lda, float* __restrict__ Dinv, int64_t ldd,
                 int32_t* __restrict__ info, int base_index) {
    extern __shared__ float S[];                         // [128][129]: R for phase 2
    __shared__ float colb[2 * CB];
    __shared__ float rd[CB];                             // 1 / R[j][j]
    const int tid = threadIdx.x;
    const int tx = tid >> 4, ty = tid & 15;
    ptx::pdl_launch_dependents();
    ptx::pdl_wait();                                     // A_jj carries the preceding GEMM's Schur update

    float r[8][8];                                       // 64 independent loads per thread: one memory latency
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int row = a * 16 + ty, col = b * 16 + tx;
            r[a][b] = (col >= row) ? __ldcg(&Ablk[(int64_t)row * lda + col]) : 0.f;   // L2: written by the kernel just before (PDL)
        }

    chol_phase1_block<7>(r, colb, rd, info, base_index, tx, ty);
    chol_phase1_block<6>(r, colb, rd, info, base_index, tx, ty);
    chol_phase1_block<5>(r, colb, rd, info, base_index, tx, ty);
    chol_phase1_block<4>(r, colb, rd, info, base_index, tx, ty);
    chol_phase1_block<3>(r, colb, rd, info, base_index, tx, ty);
    chol_phase1_block<2>(r, colb, rd, info, base_index, tx, ty);
    chol_phase1_block<1>(r, colb, rd, info, base_index, tx, ty);
    chol_phase1_block<0>(r, colb, rd, info, base_index, tx, ty);

    __syncthreads();                                     // all phase-1 reads of colb done; S free
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) S[(a * 16 + ty) * CT + b * 16 + tx] = r[a][b];       // publish R (lower = 0)
    float x[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) x[a][b] = (a == b && tx == ty) ? 1.f : 0.f;
    __syncthreads();
    chol_phase2_block<0>(x, colb, S, rd, tx, ty);
    chol_phase2_block<1>(x, colb, S, rd, tx, ty);
    chol_phase2_block<2>(x, colb, S, rd, tx, ty);
    chol_phase2_block<3>(x, colb, S, rd, tx, ty);
    chol_phase2_block<4>(x, colb, S, rd, tx, ty);
    chol_phase2_block<5>(x, colb, S, rd, tx, ty);
    chol_phase2_block<6>(x, colb, S, rd, tx, ty);
    chol_phase2_block<7>(x, colb, S, rd, tx, ty);
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int row = a * 16 + ty, col = b * 16 + tx;
            Dinv[(int64_t)row * ldd + col] = (col >= row) ? x[a][b] : 0.f;
        }
}

// ------------------------------------------------------------------ diagonal block, fused factor + inverse
// chol_diag_kernel spends 256 barrier-separated steps: 128 for the backward Cholesky A = R R^T and 128 for X = R^-1.
// The inverse can ride along with the factorisation.  With T = diag(R) * R^-1 (so V = R^-1 has rows V[i] = T[i] / R[i][i]):
//     T[i][:] = e_i - sum_{k > i} R[i][k] * V[k][:]
// and row k of V is final as soon as the steps k' > k are done -- the same moment column k of R is.  So step k
// (k = 127 .. 0) publishes column k of R AND row k of T, and every thread applies two rank-1 updates to its register tiles:
//     A[i][c] -= R[i][k] R[c][k]     (i, c < k)            the Schur update of the factorisation
//     T[i][c] -= R[i][k] V[k][c]     (i < k, c >= k)       the elimination of the inverse
// The two updates touch complementary column ranges, i.e. one full 8x8 tile of FMAs per thread and step -- the same
// arithmetic as before, in 128 barrier steps instead of 256.
template <int NB>
__device__ __forceinline__ void diag3_publish(float (&r)[8][8], float (&t)[8][8], float* colb, float* trow, float* rd,
                                              int32_t* info, int base_index, int kt, int tx, int ty, int warp) {
    const int k = NB * 16 + kt;
    if (warp == (kt >> 1)) {                                   // the warp holding column k of A (warp-uniform)
        const float v = __shfl_sync(0xffffffffu, r[NB][NB], ((kt & 1) << 4) | kt);       // pivot (k,k)
        const float rinv = inv_sqrt(v);
        if (tx == kt) {
            if (ty == kt) {
                if (!(v > 0.f)) atomicCAS(info, 0, base_index + k + 1);
                rd[k] = rinv;
            }
#pragma unroll
            for (int a = 0; a <= NB; ++a) {
                const int row = a * 16 + ty;
                float c = r[a][NB] * rinv;                     // row == k: v * rsqrt(v) = sqrt(v)
                if (row > k) c = 0.f;
                colb[row] = c;
            }
        }
    }
    if (ty == kt) {                                            // the owners of row k of T (all warps): columns b*16+tx
#pragma unroll
        for (int b = NB; b < 8; ++b) trow[b * 16 + tx] = t[NB][b];
    }
}

template <int KB>
__device__ __forceinline__ void diag3_block(float (&r)[8][8], float (&t)[8][8], float* colb2, float* trow2, float* rd,
                                            int32_t* info, int base_index, int tx, int ty, int warp) {
#pragma unroll 1
    for (int kt = 15; kt >= 0; --kt) {
        const int k = KB * 16 + kt;
        const float* colb = colb2 + (k & 1) * CB;
        const float* trow = trow2 + (k & 1) * CB;
        __syncthreads();                                       // column k of R, row k of T, rd[k] visible
        float cr[KB + 1], cc[KB + 1], vr[8];
        const float rinv = rd[k];
#pragma unroll
        for (int a = 0; a <= KB; ++a) cr[a] = colb[a * 16 + ty];
#pragma unroll
        for (int b = 0; b <= KB; ++b) cc[b] = colb[b * 16 + tx];
#pragma unroll
        for (int b = KB; b < 8; ++b) vr[b] = trow[b * 16 + tx] * rinv;           // V[k][c], c >= k (zero left of k)
#pragma unroll
        for (int a = 0; a <= KB; ++a) {
            const bool rowact = (a < KB || ty < kt);                            // row < k
#pragma unroll
            for (int b = 0; b <= KB; ++b) {
                const bool act = rowact && (b < KB || tx < kt);                 // col < k
                if (act) r[a][b] = fmaf(-cr[a], cc[b], r[a][b]);
            }
#pragma unroll
            for (int b = KB; b < 8; ++b) {
                const bool act = rowact && (b > KB || tx >= kt);                // col >= k
                if (act) t[a][b] = fmaf(-cr[a], vr[b], t[a][b]);
            }
        }
        // publish step k-1 into the other buffers (their readers finished before the barrier above)
        float* ncol = colb2 + ((k - 1) & 1) * CB;
        float* nrow = trow2 + ((k - 1) & 1) * CB;
        if (kt >= 1) diag3_publish<KB>(r, t, ncol, nrow, rd, info, base_index, kt - 1, tx, ty, warp);
        else if (KB >= 1) diag3_publish<(KB >= 1 ? KB - 1 : 0)>(r, t, ncol, nrow, rd, info, base_index, 15, tx, ty, warp);
    }
}

__global__ void __launch_bounds__(256, 1)
chol_diag3_kernel(const float* __restrict__ Ablk, int64_t lda, float* __restrict__ Dinv, int64_t ldd,
                  int32_t* __restrict__ info, int base_index) {
    __shared__ float colb[2 * CB];
    __shared__ float trow[2 * CB];
    __shared__ float rd[CB];                             // 1 / R[k][k]
    const int tid = threadIdx.x;
    const int tx = tid >> 4, ty = tid & 15, warp = tid >> 5;
    ptx::pdl_launch_dependents();
    ptx::pdl_wait();                                     // A_jj carries the preceding GEMM's Schur update

    float r[8][8], t[8][8];                              // 64 independent loads per thread: one memory latency
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int row = a * 16 + ty, col = b * 16 + tx;
            r[a][b] = (col >= row) ? __ldcg(&Ablk[(int64_t)row * lda + col]) : 0.f;   // L2: written by the kernel just before (PDL)
            t[a][b] = (a == b && tx == ty) ? 1.f : 0.f;
        }
    for (int i = tid; i < 2 * CB; i += 256) { trow[i] = 0.f; }
    __syncthreads();
    diag3_publish<7>(r, t, colb + CB, trow + CB, rd, info, base_index, 15, tx, ty, warp);    // step 127 (odd -> buffer 1)
    diag3_block<7>(r, t, colb, trow, rd, info, base_index, tx, ty, warp);
    diag3_block<6>(r, t, colb, trow, rd, info, base_index, tx, ty, warp);
    diag3_block<5>(r, t, colb, trow, rd, info, base_index, tx, ty, warp);
    diag3_block<4>(r, t, colb, trow, rd, info, base_index, tx, ty, warp);
    diag3_block<3>(r, t, colb, trow, rd, info, base_index, tx, ty, warp);
    diag3_block<2>(r, t, colb, trow, rd, info, base_index, tx, ty, warp);
    diag3_block<1>(r, t, colb, trow, rd, info, base_index, tx, ty, warp);
    diag3_block<0>(r, t, colb, trow, rd, info, base_index, tx, ty, warp);
    __syncthreads();                                     // every rd[] entry visible
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int row = a * 16 + ty;
        const float ri = rd[row];
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int col = b * 16 + tx;
            Dinv[(int64_t)row * ldd + col] = (col >= row) ? t[a][b] * ri : 0.f;       // V = diag(R)^-1 T
        }
    }
}

// ------------------------------------------------------------------ fused kernel with a dedicated pivot warp
// In chol_diag3_kernel the warp that owns the next column first applies its whole share of the rank-1 updates and only
// then runs the pivot chain (shuffle, rsqrt, scale, store) -- with every other warp waiting at the barrier.  Here a ninth
// warp owns no tile and does nothing but that chain: right after the barrier the owners bring their column up to date
// (8 FMAs), hand it over raw through shared memory (named barrier 1: bar.arrive by the owner warp, bar.sync by the pivot
// warp) and go on with their updates, while the pivot warp finishes and publishes the column for the next step.
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// tile warps: send column (NB*16 + kt) of A, as it stands, to the pivot warp; publish row (NB*16 + kt) of T
template <int NB>
__device__ __forceinline__ void diag4_handover(float (&r)[8][8], float (&t)[8][8], float* rawc, float* trow, int kt, int tx,
                                               int ty, int warp) {
    if (warp == (kt >> 1)) {                                   // the warp holding that column (warp-uniform)
        if (tx == kt) {
#pragma unroll
            for (int a = 0; a <= NB; ++a) rawc[a * 16 + ty] = r[a][NB];
        }
        __syncwarp();
        named_arrive(1, 64);
    }
    if (ty == kt) {
#pragma unroll
        for (int b = NB; b < 8; ++b) trow[b * 16 + tx] = t[NB][b];
    }
}

// pivot warp: raw column k -> scaled column of R in colb, 1/R[k][k] in rd
__device__ __forceinline__ void diag4_pivot(const float* rawc, float* colb, float* rd, int32_t* info, int base_index, int k,
                                            int lane) {
    named_sync(1, 64);
    const float v = rawc[k];
    const float rinv = inv_sqrt(v);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int row = q * 32 + lane;
        if (row <= (k | 15)) {                                 // rows of the blocks the owners wrote
            float c = rawc[row] * rinv;                        // row == k: v * rsqrt(v) = sqrt(v)
            if (row > k) c = 0.f;
            colb[row] = c;
        }
    }
    if (lane == 0) {
        rd[k] = rinv;
        if (!(v > 0.f)) atomicCAS(info, 0, base_index + k + 1);
    }
}

// One step's rank-1 updates with NO per-element predicates (the predicated form was issue-bound: ~250 instructions per
// thread and step).  The masks are separable and almost free: the published column is already zero below its diagonal,
// the published row of T is zero left of its diagonal, so only the diagonal entry itself has to be masked out of the two
// multiplier vectors.  kt is uniform over the CTA, so the kt >= 1 / kt == 0 paths are plain branches over static code.
template <int KB>
__device__ __forceinline__ void diag4_block(float (&r)[8][8], float (&t)[8][8], float* colb2, float* trow2, float* rawc,
                                            float* rd, int tx, int ty, int warp) {
    constexpr int NBL = KB >= 1 ? KB - 1 : 0;                  // block of step k-1 when kt == 0
#pragma unroll 1
    for (int kt = 15; kt >= 0; --kt) {
        const int k = KB * 16 + kt;
        const float* colb = colb2 + (k & 1) * CB;
        const float* trow = trow2 + (k & 1) * CB;
        float* nrow = trow2 + ((k - 1) & 1) * CB;
        __syncthreads();                                       // column k of R, row k of T, rd[k] visible
        float cr[KB + 1], cc[KB + 1], vr[8];
        const float rinv = rd[k];
#pragma unroll
        for (int a = 0; a <= KB; ++a) cr[a] = colb[a * 16 + ty];              // R[i][k]; 0 for i > k
#pragma unroll
        for (int b = 0; b <= KB; ++b) cc[b] = colb[b * 16 + tx];              // R[c][k]; 0 for c > k
        cr[KB] = (ty < kt) ? cr[KB] : 0.f;                                     // drop i == k
        cc[KB] = (tx < kt) ? cc[KB] : 0.f;                                     // drop c == k
#pragma unroll
        for (int b = KB; b < 8; ++b) vr[b] = trow[b * 16 + tx] * rinv;         // V[k][c]; 0 for c < k
        if (kt >= 1) {
            // part 1: column block KB of A and row block KB of T hold column / row k-1 -> hand them over first
#pragma unroll
            for (int a = 0; a <= KB; ++a) r[a][KB] = fmaf(-cr[a], cc[KB], r[a][KB]);
#pragma unroll
            for (int b = KB; b < 8; ++b) t[KB][b] = fmaf(-cr[KB], vr[b], t[KB][b]);
            diag4_handover<KB>(r, t, rawc, nrow, kt - 1, tx, ty, warp);
#pragma unroll
            for (int a = 0; a <= KB; ++a) {
#pragma unroll
                for (int b = 0; b < KB; ++b) r[a][b] = fmaf(-cr[a], cc[b], r[a][b]);
                if (a < KB) {
#pragma unroll
                    for (int b = KB; b < 8; ++b) t[a][b] = fmaf(-cr[a], vr[b], t[a][b]);
                }
            }
        } else if (KB >= 1) {
            // k = KB*16: rows and columns < k are exactly the blocks <= NBL
#pragma unroll
            for (int a = 0; a <= NBL; ++a) r[a][NBL] = fmaf(-cr[a], cc[NBL], r[a][NBL]);
#pragma unroll
            for (int b = KB; b < 8; ++b) t[NBL][b] = fmaf(-cr[NBL], vr[b], t[NBL][b]);
            diag4_handover<NBL>(r, t, rawc, nrow, 15, tx, ty, warp);
#pragma unroll
            for (int a = 0; a <= NBL; ++a) {
#pragma unroll
                for (int b = 0; b < NBL; ++b) r[a][b] = fmaf(-cr[a], cc[b], r[a][b]);
                if (a < NBL) {
#pragma unroll
                    for (int b = KB; b < 8; ++b) t[a][b] = fmaf(-cr[a], vr[b], t[a][b]);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(288, 1)
chol_diag4_kernel(const float* __restrict__ Ablk, int64_t lda, float* __restrict__ Dinv, int64_t ldd,
                  int32_t* __restrict__ info, int base_index) {
    __shared__ float colb[2 * CB];
    __shared__ float trow[2 * CB];
    __shared__ float rawc[CB];
    __shared__ float rd[CB];                             // 1 / R[k][k]
    const int tid = threadIdx.x;
    const int tx = (tid >> 4) & 15, ty = tid & 15, warp = tid >> 5, lane = tid & 31;
    ptx::pdl_launch_dependents();
    ptx::pdl_wait();                                     // A_jj carries the preceding GEMM's Schur update
    if (warp == 8) {
        // ===== pivot warp: column k-1 is published while the tile warps work on step k
        diag4_pivot(rawc, colb + CB, rd, info, base_index, 127, lane);
#pragma unroll 1
        for (int k = 127; k >= 0; --k) {
            __syncthreads();
            if (k >= 1) diag4_pivot(rawc, colb + ((k - 1) & 1) * CB, rd, info, base_index, k - 1, lane);
        }
        __syncthreads();
        return;
    }
    float r[8][8], t[8][8];                              // 64 independent loads per thread: one memory latency
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int row = a * 16 + ty, col = b * 16 + tx;
            r[a][b] = (col >= row) ? __ldcg(&Ablk[(int64_t)row * lda + col]) : 0.f;   // L2: written by the kernel just before (PDL)
            t[a][b] = (a == b && tx == ty) ? 1.f : 0.f;
        }
    diag4_handover<7>(r, t, rawc, trow + CB, 15, tx, ty, warp);                  // step 127 (odd -> buffer 1)
    diag4_block<7>(r, t, colb, trow, rawc, rd, tx, ty, warp);
    diag4_block<6>(r, t, colb, trow, rawc, rd, tx, ty, warp);
    diag4_block<5>(r, t, colb, trow, rawc, rd, tx, ty, warp);
    diag4_block<4>(r, t, colb, trow, rawc, rd, tx, ty, warp);
    diag4_block<3>(r, t, colb, trow, rawc, rd, tx, ty, warp);
    diag4_block<2>(r, t, colb, trow, rawc, rd, tx, ty, warp);
    diag4_block<1>(r, t, colb, trow, rawc, rd, tx, ty, warp);
    diag4_block<0>(r, t, colb, trow, rawc, rd, tx, ty, warp);
    __syncthreads();                                     // every rd[] entry visible
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int row = a * 16 + ty;
        const float ri = rd[row];
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int col = b * 16 + tx;
            Dinv[(int64_t)row * ldd + col] = (col >= row) ? t[a][b] * ri : 0.f;       // V = diag(R)^-1 T
        }
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
static int diag_kernel_from_env() {
    const char* e = getenv("GPTQ_B200_DIAG_KERNEL");
    if (e && (e[0] == '1' || e[0] == '3' || e[0] == '4')) return e[0] - '0';
    return 4;
}
// 4 = fused factor + inverse with a dedicated pivot warp (default, 48.5 us), 3 = fused without it (52 us), 1 = two-phase
// register-tile kernel (58 us).  A recursive shared-memory kernel with a one-warp 32x32 base case was measured at 64 us
// and removed (profiles/r01_kernels_ncu_full.md).
static int g_diag_kernel = diag_kernel_from_env();

struct HinvCtx {
    float* A; int64_t lda; float* V; int64_t ldv; float* S; int32_t* info; cudaStream_t st; size_t smem; int rc; bool first;
};

static void hinv_rec(HinvCtx& c, int a, int b) {
    if (c.rc) return;
    if (b - a == 1) {
        const int c0 = a * CB;
        const bool pdl_ok = !c.first;      // the very first kernel follows the memsets of info / U
        c.first = false;
        cudaError_t e = g_diag_kernel == 4
            ? launch_pdl(pdl_ok, chol_diag4_kernel, dim3(1), dim3(288), (size_t)0, c.st, c.A + (int64_t)c0 * c.lda + c0, c.lda,
                         c.V + (int64_t)c0 * c.ldv + c0, c.ldv, c.info, c0)
            : g_diag_kernel == 3
            ? launch_pdl(pdl_ok, chol_diag3_kernel, dim3(1), dim3(256), (size_t)0, c.st, c.A + (int64_t)c0 * c.lda + c0, c.lda,
                         c.V + (int64_t)c0 * c.ldv + c0, c.ldv, c.info, c0)
            : launch_pdl(pdl_ok, chol_diag_kernel, dim3(1), dim3(256), c.smem, c.st, c.A + (int64_t)c0 * c.lda + c0, c.lda,
                         c.V + (int64_t)c0 * c.ldv + c0, c.ldv, c.info, c0);
        if (e != cudaSuccess) { c.rc = set_err((int)e, "hinv_upper: diagonal-block launch failed: %s", cudaGetErrorString(e)); return; }
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
    c.rc = gemm_fp32(0, 1, n2, n1, n2, 1.f, V22, c.ldv, A12, c.lda, 0.f, Tt, c.lda, SG_A_UPPER, c.st);
    if (c.rc) return;
    // A11 -= Tt^T Tt
    c.rc = gemm_fp32(1, 0, n1, n1, n2, -1.f, Tt, c.lda, Tt, c.lda, 1.f, A11, c.lda, SG_UPPER_TILES, c.st);
    if (c.rc) return;
    hinv_rec(c, a, m);
    if (c.rc) return;
    // S[n1 x n2] = V11 * Tt^T
    c.rc = gemm_fp32(0, 1, n1, n2, n1, 1.f, V11, c.ldv, Tt, c.lda, 0.f, c.S, n2, SG_A_UPPER, c.st);
    if (c.rc) return;
    // V12 = -S * V22
    c.rc = gemm_fp32(0, 0, n1, n2, n2, -1.f, c.S, n2, V22, c.ldv, 0.f, V12, c.ldv, SG_B_UPPER, c.st);
}

// ------------------------------------------------------------------ right-looking driver with look-ahead
// The recursion above keeps every flop in a large-k GEMM but runs strictly in sequence: while one SM factors a diagonal
// block (48.5 us) or a 128..256-wide GEMM of the lowest levels runs, the other SMs idle -- 2/3 of the chain's time at
// K = 4096 and 11008.  This driver keeps the recursion for PANELS of `pb` blocks and overlaps the rest on three
// internal streams (forked from / joined into the caller's stream with events; the call stays asynchronous and ordered
// on the caller's stream).  Panels p = [e0, e1) are processed from the last to the first; q = [f0, f1), f1 = e0, is the next:
//
//   critical stream   hinv_rec(p)                              V_pp = R_pp^-1 and R inside the panel
//                     wait trail(p+1)
//                     Tt_p[:, q]  = V_pp A[q, p]^T             look-ahead solve: only the columns of the next panel
//                     record rec(p)
//                     A[q, q]    -= Tt_p[:, q]^T Tt_p[:, q]    the one block the next hinv_rec needs
//   trailing stream   wait rec(p)
//                     Tt_p[:, 0:f0] = V_pp A[0:f0, p]^T        rest of the panel solve
//                     A[0:f0, 0:f1] -= Tt_p[:, 0:f0]^T Tt_p[:, 0:f1]   (upper tiles)    rank-|p| trailing update
//                     record trail(p)
//   inverse stream    wait rec(p)
//                     V[p, e1:] = -V_pp S_p                    row block p of U, with the S_p issued one panel earlier
//                     S_q = R[q, e0:] V[e0:, e0:]              for the NEXT panel: needs only finished panels
//
// Every (block, source panel) update is applied exactly once and in a fixed order (trail(p+1) before the look-ahead
// update of p, enforced by the event), so the result does not depend on how the streams interleave.
// Tt_p lives in A's lower mirror A[p, 0:e0] like in the recursion.
struct ChainStreams {
    cudaStream_t crit = nullptr, trail = nullptr, inv = nullptr;
    cudaEvent_t fork = nullptr, join_c = nullptr, join_t = nullptr, join_i = nullptr;
    std::vector<cudaEvent_t> rec, trl;
};
static std::mutex g_chain_mu;                       // one look-ahead chain is being ISSUED at a time (they share the event pool)
static ChainStreams g_chain_streams[64];

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
static int g_chain_lookahead = env_int("GPTQ_B200_CHAIN_LOOKAHEAD", 1);     // 0: the plain recursion
static int g_chain_panel = env_int("GPTQ_B200_CHAIN_PANEL", 0);             // blocks per panel; 0 = by size (panel_blocks)
static int g_chain_min_blocks = env_int("GPTQ_B200_CHAIN_MIN_BLOCKS", 8);   // fewer blocks: the recursion as it is
// TIMING EXPERIMENTS ONLY (results are wrong when set): bit 0 = issue nothing on the inverse stream, bit 1 = nothing on the
// trailing stream -- what is left is the critical stream alone (tools/chain_lab.py).
static int g_chain_skip = env_int("GPTQ_B200_CHAIN_SKIP", 0);
// GPTQ_B200_CHAIN_TRACE=1: timing events around every phase of every panel; the call then SYNCHRONISES and prints one line
// per panel to stderr (microseconds since the fork).  Diagnostics only.
static int g_chain_trace = env_int("GPTQ_B200_CHAIN_TRACE", 0);
struct ChainTrace {
    bool on;
    cudaEvent_t t0 = nullptr;
    std::vector<std::vector<cudaEvent_t>> ev;          // [panel][slot]
    enum { REC0, REC1, SOLVE0, UPD1, TR0, TR1, INV0, INV1, NSLOT };
    void init(int np) {
        if (!on) return;
        cudaEventCreate(&t0);
        ev.assign(np, std::vector<cudaEvent_t>(NSLOT, nullptr));
    }
    void mark(int panel, int slot, cudaStream_t st) {
        if (!on) return;
        cudaEventCreate(&ev[panel][slot]);
        cudaEventRecord(ev[panel][slot], st);
    }
    void dump(cudaStream_t st) {
        if (!on) return;
        cudaStreamSynchronize(st);
        static const char* names[NSLOT] = {"rec0", "rec1", "solve0", "upd1", "trail0", "trail1", "inv0", "inv1"};
        for (int i = (int)ev.size() - 1; i >= 0; --i) {
            fprintf(stderr, "chain_trace panel %d", i);
            for (int k = 0; k < NSLOT; ++k) {
                float ms = -1.f;
                if (ev[i][k]) { cudaEventElapsedTime(&ms, t0, ev[i][k]); cudaEventDestroy(ev[i][k]); }
                fprintf(stderr, " %s=%.1f", names[k], ms < 0 ? -1.0 : ms * 1e3);
            }
            fprintf(stderr, "\n");
        }
        cudaEventDestroy(t0);
    }
};

// Measured on B200 (tools/chain_sweep.py, profiles/r02_chain_lookahead.md): up to K = 4096 the chain is bound by the critical
// stream, whose cost per block is lowest with one block per panel (diag + two 128^3 products); from K ~ 8192 on it is bound
// by the throughput of the rank-|panel| trailing updates and the |panel|-row inverse products, which need k = M >= 1024 to
// run the tensor-core GEMM near its rate; beyond ~24k columns the diagonal blocks are a few
// per cent of the chain and the recursion's large square products win (K = 36864: 237 ms against 238-267).
static int g_chain_max_blocks = env_int("GPTQ_B200_CHAIN_MAX_BLOCKS", 192);
static int panel_blocks(int nb) {
    if (g_chain_panel > 0) return g_chain_panel;
    // (K = 16384 measured 24.0 ms with 4096-wide panels against 26.3 with 1024-wide ones and 29.2 for the recursion; the wider
    // panels are left to GPTQ_B200_CHAIN_PANEL=32: the full parity suite of this round ran with 8)
    return nb <= 36 ? 1 : nb <= 56 ? 2 : 8;
}
static bool use_lookahead(int nb) {
    if (!g_chain_lookahead || nb < g_chain_min_blocks || nb <= 2 * panel_blocks(nb)) return false;
    return g_chain_panel > 0 || nb <= g_chain_max_blocks;       // an explicit panel width (tests, sweeps) overrides the size limit
}

static size_t rec_scratch_floats(size_t nb) {
    size_t s = (nb / 2) * CB * ((nb - nb / 2) * CB);
    return s < (size_t)CB * CB ? (size_t)CB * CB : s;
}

static int chain_streams(ChainStreams** out, int npanels) {
    const int dev = cur_device();
    if (dev < 0 || dev >= 64) return set_err(-1, "hinv_upper: device ordinal %d out of range", dev);
    ChainStreams& cs = g_chain_streams[dev];
    if (!cs.crit) {
        int least = 0, greatest = 0;
        GQ_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));      // numerically lower = higher priority
        const int mid = greatest < least ? greatest + 1 : least;
        const bool inv_first = env_int("GPTQ_B200_CHAIN_INV_PRIO", 0) != 0;
        GQ_CUDA(cudaStreamCreateWithPriority(&cs.crit, cudaStreamNonBlocking, greatest));
        GQ_CUDA(cudaStreamCreateWithPriority(&cs.trail, cudaStreamNonBlocking, inv_first ? least : mid));
        GQ_CUDA(cudaStreamCreateWithPriority(&cs.inv, cudaStreamNonBlocking, inv_first ? mid : least));
        for (cudaEvent_t* e : {&cs.fork, &cs.join_c, &cs.join_t, &cs.join_i}) GQ_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    }
    while ((int)cs.rec.size() < npanels) {
        cudaEvent_t a, b;
        GQ_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        GQ_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        cs.rec.push_back(a); cs.trl.push_back(b);
    }
    *out = &cs;
    return 0;
}

#define GQ_RC(expr) do { int rc__ = (expr); if (rc__) return rc__; } while (0)

static int hinv_lookahead(float* A, int nb, int64_t lda, float* V, int64_t ldv, int32_t* info, float* ws, size_t smem,
                          cudaStream_t st) {
    const int pb = panel_blocks(nb);
    const int np = (nb + pb - 1) / pb;
    const int Kp = nb * CB;
    std::lock_guard<std::mutex> lock(g_chain_mu);
    ChainStreams* cs = nullptr;
    GQ_RC(chain_streams(&cs, np));
    float* S_inv = ws;                                       // [panel rows x columns right of the panel], inverse stream
    float* S_rec = ws + (size_t)pb * CB * Kp;                // scratch of the in-panel recursion, critical stream
    ChainTrace tr{g_chain_trace != 0};
    tr.init(np);
    if (tr.on) cudaEventRecord(tr.t0, st);
    GQ_CUDA(cudaEventRecord(cs->fork, st));
    GQ_CUDA(cudaStreamWaitEvent(cs->crit, cs->fork, 0));
    HinvCtx c{A, lda, V, ldv, S_rec, info, cs->crit, smem, 0, true};
    // panel i (i = np-1 .. 0) covers blocks [max(0, nb - (np - i) pb), nb - (np - 1 - i) pb): the short panel is the first one
    for (int i = np - 1; i >= 0; --i) {
        const int b1 = nb - (np - 1 - i) * pb, b0 = (b1 - pb > 0) ? b1 - pb : 0;
        const int64_t e0 = (int64_t)b0 * CB, e1 = (int64_t)b1 * CB;
        const int P = (int)(e1 - e0);
        const bool last = (i == np - 1);
        tr.mark(i, ChainTrace::REC0, cs->crit);
        hinv_rec(c, b0, b1);
        if (c.rc) return c.rc;
        tr.mark(i, ChainTrace::REC1, cs->crit);
        const float* Vpp = V + e0 * ldv + e0;
        float* Tt = A + e0 * lda;                            // Tt_p, [P x e0]
        if (i > 0) {
            const int a0 = (b0 - pb > 0) ? b0 - pb : 0;
            const int64_t f0 = (int64_t)a0 * CB;
            const int Q = (int)(e0 - f0);
            int nopdl = 0;
            if (!last) { GQ_CUDA(cudaStreamWaitEvent(cs->crit, cs->trl[i + 1], 0)); nopdl = SG_NO_PDL; }
            tr.mark(i, ChainTrace::SOLVE0, cs->crit);
            GQ_RC(gemm_fp32(0, 1, P, Q, P, 1.f, Vpp, ldv, A + f0 * lda + e0, lda, 0.f, Tt + f0, lda, SG_A_UPPER | nopdl, cs->crit));
            GQ_CUDA(cudaEventRecord(cs->rec[i], cs->crit));
            GQ_RC(gemm_fp32(1, 0, Q, Q, P, -1.f, Tt + f0, lda, Tt + f0, lda, 1.f, A + f0 * lda + f0, lda, SG_UPPER_TILES, cs->crit));
            tr.mark(i, ChainTrace::UPD1, cs->crit);
            // trailing stream
            GQ_CUDA(cudaStreamWaitEvent(cs->trail, cs->rec[i], 0));
            tr.mark(i, ChainTrace::TR0, cs->trail);
            if (f0 > 0 && !(g_chain_skip & 2)) {
                GQ_RC(gemm_fp32(0, 1, P, (int)f0, P, 1.f, Vpp, ldv, A + e0, lda, 0.f, Tt, lda, SG_A_UPPER | SG_NO_PDL, cs->trail));
                GQ_RC(gemm_fp32(1, 0, (int)f0, (int)e0, P, -1.f, Tt, lda, Tt, lda, 1.f, A, lda, SG_UPPER_TILES, cs->trail));
            }
            tr.mark(i, ChainTrace::TR1, cs->trail);
            GQ_CUDA(cudaEventRecord(cs->trl[i], cs->trail));
        } else {
            GQ_CUDA(cudaEventRecord(cs->rec[i], cs->crit));
        }
        // inverse stream: row block i of U right of its diagonal block,  V[p, e1:] = -V_pp S_p  with  S_p = R[p, e1:] V[e1:, e1:].
        // S_p needs nothing of panel p itself (its rows of R were written by the solves of the later panels, V[e1:, e1:] is
        // complete once the row block of panel i+1 is), so it was issued one iteration earlier and ran while panel p was
        // being factored; what is left after the last panel is one |p| x |p| product per row block.
        if (!(g_chain_skip & 1)) {
            GQ_CUDA(cudaStreamWaitEvent(cs->inv, cs->rec[i], 0));
            tr.mark(i, ChainTrace::INV0, cs->inv);
            int nopdl = SG_NO_PDL;
            if (!last) {
                const int nl = (int)(Kp - e1);
                GQ_RC(gemm_fp32(0, 0, P, nl, P, -1.f, Vpp, ldv, S_inv, nl, 0.f, V + e0 * ldv + e1, ldv, SG_A_UPPER | nopdl, cs->inv));
                nopdl = 0;
            }
            if (i > 0) {                                     // S of the NEXT panel q = [f0, e0): columns right of it = [e0, Kp)
                const int a0 = (b0 - pb > 0) ? b0 - pb : 0;
                const int64_t f0 = (int64_t)a0 * CB;
                const int Q = (int)(e0 - f0), nl = (int)(Kp - e0);
                GQ_RC(gemm_fp32(1, 0, Q, nl, nl, 1.f, A + e0 * lda + f0, lda, V + e0 * ldv + e0, ldv, 0.f, S_inv, nl,
                                SG_B_UPPER | nopdl, cs->inv));
            }
            tr.mark(i, ChainTrace::INV1, cs->inv);
        }
    }
    GQ_CUDA(cudaEventRecord(cs->join_c, cs->crit));
    GQ_CUDA(cudaEventRecord(cs->join_t, cs->trail));
    GQ_CUDA(cudaEventRecord(cs->join_i, cs->inv));
    GQ_CUDA(cudaStreamWaitEvent(st, cs->join_c, 0));
    GQ_CUDA(cudaStreamWaitEvent(st, cs->join_t, 0));
    GQ_CUDA(cudaStreamWaitEvent(st, cs->join_i, 0));
    tr.dump(st);
    return 0;
}

}  // namespace gq

using namespace gq;

extern "C" size_t gptq_b200_hinv_upper_workspace_bytes(int Kp) {
    if (Kp <= 0) return 0;
    const size_t nb = (size_t)(Kp / CB);
    size_t s = rec_scratch_floats(nb);                   // the recursion: largest n1 x n2 scratch (top-level split)
    if (use_lookahead((int)nb)) {
        const size_t pb = (size_t)panel_blocks((int)nb);
        const size_t la = pb * CB * Kp + rec_scratch_floats(pb);
        if (la > s) s = la;
    }
    return s * sizeof(float);
}

extern "C" int gptq_b200_set_chain_schedule(int lookahead, int panel_blocks) {
    GQ_ARG(panel_blocks <= 64, 2, "panel_blocks must be 0 (by size) .. 64 (negative: keep)");
    if (lookahead >= 0) g_chain_lookahead = lookahead ? 1 : 0;
    if (panel_blocks >= 0) g_chain_panel = panel_blocks;
    return 0;
}

extern "C" int gptq_b200_hinv_upper(float* A, int Kp, int64_t lda, float* U, int64_t ldu, int32_t* info, void* ws,
                                    size_t ws_bytes, void* stream) {
    GQ_ARG(A, 1, "null"); GQ_ARG(Kp > 0 && Kp % CB == 0, 2, "Kp must be a positive multiple of 128");
    GQ_ARG(lda >= Kp && lda % 4 == 0, 3, "lda must be >= Kp and a multiple of 4");
    GQ_ARG(U && U != A, 4, "U must be a separate matrix"); GQ_ARG(ldu >= Kp && ldu % 4 == 0, 5, "ldu must be >= Kp and a multiple of 4");
    GQ_ARG(info, 6, "null"); GQ_ARG(ws, 7, "null");
    GQ_ARG(ws_bytes >= gptq_b200_hinv_upper_workspace_bytes(Kp), 8, "workspace too small");
    cudaStream_t st = as_stream(stream);
    const size_t smem = (size_t)CB * CT * sizeof(float);
    GQ_ONCE_PER_DEVICE(GQ_CUDA(cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
    GQ_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), st));
    GQ_CUDA(cudaMemset2DAsync(U, ldu * sizeof(float), 0, (size_t)Kp * sizeof(float), Kp, st));
    if (use_lookahead(Kp / CB)) {
        int rc = hinv_lookahead(A, Kp / CB, lda, U, ldu, info, static_cast<float*>(ws), smem, st);
        if (rc) return rc;
    } else {
        HinvCtx c{A, lda, U, ldu, static_cast<float*>(ws), info, st, smem, 0, true};
        hinv_rec(c, 0, Kp / CB);
        if (c.rc) return c.rc;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err((int)e, "hinv_upper: launch failed: %s", cudaGetErrorString(e));
    return 0;
}
