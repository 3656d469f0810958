// fasterquant prologue / epilogue data movement: dead columns, damp, act-order permutation,
// gathers, un-permute + cast.  All HBM-bound byte movers; see include/gptq_b200.h for the
// reference lines each one replaces.
#include "common.cuh"

namespace gq {

// ------------------------------------------------------------------ dead columns (gptq.py:143)
__global__ void dead_columns_kernel(const float* __restrict__ H, int K, int64_t ldh, uint8_t* __restrict__ dead) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K) dead[i] = (H[(int64_t)i * ldh + i] == 0.f) ? 1 : 0;
}

// One CTA: fix dead diagonal entries, export diag, damp = percdamp * mean(diag) (gptq.py:144,174).
// Deterministic: fixed-shape tree reduction in fp32 pairs accumulated in double.
__global__ void __launch_bounds__(1024)
hessian_prepare_kernel(float* __restrict__ H, int K, int64_t ldh, float percdamp, float* __restrict__ diag_out,
                       float* __restrict__ damp_out, int32_t* __restrict__ ndead_out) {
    __shared__ double ssum[32];
    __shared__ int scnt[32];
    double acc = 0.0;
    int cnt = 0;
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        float d = H[(int64_t)i * ldh + i];
        if (d == 0.f) { d = 1.f; H[(int64_t)i * ldh + i] = 1.f; ++cnt; }
        diag_out[i] = d;
        acc += (double)d;
    }
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0) { ssum[threadIdx.x >> 5] = acc; scnt[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0; int c = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t += ssum[w]; c += scnt[w]; }
        float mean = (float)(t / (double)K);
        *damp_out = __fmul_rn(percdamp, mean);
        if (ndead_out) *ndead_out = c;
    }
}

// ------------------------------------------------------------------ W.float() + W[:,dead]=0
template <typename T>
__global__ void weights_prepare_kernel(const T* __restrict__ w, int R, int K, int64_t ldw, const uint8_t* __restrict__ dead,
                                       float* __restrict__ out, int64_t ldo) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= K) return;
    bool dz = dead && dead[c];
    for (int r = blockIdx.y; r < R; r += gridDim.y) {
        float v = to_f32<T>(w[(int64_t)r * ldw + c]);
        out[(int64_t)r * ldo + c] = dz ? 0.f : v;
    }
}

// ------------------------------------------------------------------ argsort (gptq.py:166)
// Rank-by-counting: rank(i) = #{j : v[j] > v[i] or (v[j]==v[i] and j<i)}.  O(n^2) compares,
// n <= 64K, fully parallel and deterministic (ties broken by ascending index).
__global__ void __launch_bounds__(256)
argsort_desc_kernel(const float* __restrict__ v, int n, int32_t* __restrict__ perm, int32_t* __restrict__ invperm) {
    __shared__ float tile[1024];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float vi = (i < n) ? v[i] : 0.f;
    const bool ni = vi != vi;
    int rank = 0;
    for (int base = 0; base < n; base += 1024) {
        int m = min(1024, n - base);
        __syncthreads();
        for (int t = threadIdx.x; t < m; t += blockDim.x) tile[t] = v[base + t];
        __syncthreads();
        if (i < n) {
            for (int t = 0; t < m; ++t) {
                float vj = tile[t];
                int j = base + t;
                // total order (NaN first, like torch.argsort descending) so perm is a permutation for ANY input
                const bool nj = vj != vj;
                rank += ni ? (nj && j < i) : (nj || (vj > vi) || (vj == vi && j < i));
            }
        }
    }
    if (i < n) { perm[rank] = i; invperm[i] = rank; }
}

// ------------------------------------------------------------------ W[:, perm]
// One CTA per row: the gather W[r, perm[c]] then reads from a row that stays in L1 (a CTA covering only 256 columns
// would pull the whole row through L2 for 1 KB of output).
__global__ void __launch_bounds__(256)
permute_cols_kernel(const float* __restrict__ W, int R, int K, int64_t ldw, const int32_t* __restrict__ perm,
                    float* __restrict__ out, int64_t ldo) {
    for (int r = blockIdx.x; r < R; r += gridDim.x) {
        const float* src = W + (int64_t)r * ldw;
        float* dst = out + (int64_t)r * ldo;
        for (int c = threadIdx.x; c < K; c += blockDim.x) dst[c] = src[perm ? perm[c] : c];
    }
}

// ------------------------------------------------------------------ H[perm][:,perm] + damp, padded
__global__ void __launch_bounds__(256)
hessian_gather_kernel(const float* __restrict__ H, int K, int64_t ldh, const int32_t* __restrict__ perm,
                      const float* __restrict__ damp, float* __restrict__ A, int Kp, int64_t lda) {
    // one CTA per output row: row perm[i] of H stays in L1 while it is gathered by perm[j]
    for (int i = blockIdx.x; i < Kp; i += gridDim.x) {
        float* dst = A + (int64_t)i * lda;
        if (i < K) {
            const float* src = H + (int64_t)(perm ? perm[i] : i) * ldh;
            for (int j = threadIdx.x; j < Kp; j += blockDim.x) {
                float v = 0.f;
                if (j < K) {
                    v = src[perm ? perm[j] : j];
                    if (i == j && damp) v = __fadd_rn(v, *damp);
                }
                dst[j] = v;
            }
        } else {
            for (int j = threadIdx.x; j < Kp; j += blockDim.x) dst[j] = (i == j) ? 1.f : 0.f;
        }
    }
}

// ------------------------------------------------------------------ mirror lower -> upper
__global__ void symmetrize_kernel(float* __restrict__ H, int K, int64_t ldh) {
    // 32x32 tiles through smem: tile (bi,bj) with bj > bi reads the lower tile (bj,bi) transposed.
    __shared__ float t[32][33];
    int bi = blockIdx.y, bj = blockIdx.x;
    if (bj < bi) return;
    int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        int gi = bj * 32 + r, gj = bi * 32 + tx;     // lower tile element (gi, gj)
        t[r][tx] = (gi < K && gj < K) ? H[(int64_t)gi * ldh + gj] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        int gi = bi * 32 + r, gj = bj * 32 + tx;     // upper tile element (gi, gj) = lower (gj, gi)
        if (gi < K && gj < K && gj > gi) H[(int64_t)gi * ldh + gj] = t[tx][r];
    }
}

// ------------------------------------------------------------------ lower triangle <-> packed rows (multi-GPU exchange #1)
// packed[r (r + 1) / 2 + c] = S[r][c] for c <= r: what the all-reduce of the partial Hessians has to carry (SURVEY 5 / 8e:
// the lower triangle suffices, half the bytes of the K x K square).  One CTA per row, grid-stride; coalesced on both sides.
__global__ void __launch_bounds__(256)
tril_pack_kernel(const float* __restrict__ S, int K, int64_t lds, float* __restrict__ packed, int unpack) {
    for (int r = blockIdx.x; r < K; r += gridDim.x) {
        const int64_t off = (int64_t)r * (r + 1) / 2;
        float* p = packed + off;
        float* row = const_cast<float*>(S) + (int64_t)r * lds;
        if (unpack) { for (int c = threadIdx.x; c <= r; c += blockDim.x) row[c] = p[c]; }
        else        { for (int c = threadIdx.x; c <= r; c += blockDim.x) p[c] = row[c]; }
    }
}

// ------------------------------------------------------------------ staging copy
template <typename T>
__global__ void stage_kernel(const T* __restrict__ src, int64_t tokens, int K, int64_t lds, T* __restrict__ dst, int64_t ldd) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ldd) return;
    for (int64_t t = blockIdx.y; t < tokens; t += gridDim.y)
        dst[t * ldd + c] = (c < K) ? src[t * lds + c] : from_f32<T>(0.f);
}

// 16-byte vectorised variant (K, both pitches multiples of 8 elements, 16-byte aligned bases): grid-stride over rows.
__global__ void __launch_bounds__(256)
stage_vec_kernel(const uint4* __restrict__ src, int64_t tokens, int kv, int64_t lds_v, uint4* __restrict__ dst, int64_t ldd_v) {
    // ldd_v >= kv: the pad vectors [kv, ldd_v) of every row are zeroed
    for (int64_t t = blockIdx.y; t < tokens; t += gridDim.y) {
        const uint4* s = src + t * lds_v;
        uint4* d = dst + t * ldd_v;
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ldd_v; c += gridDim.x * blockDim.x)
            d[c] = (c < kv) ? s[c] : make_uint4(0u, 0u, 0u, 0u);
    }
}

// Dense case (no padding on either side): one flat grid-stride pass, 4 independent 16-byte loads per thread in flight.
__global__ void __launch_bounds__(256)
copy_flat_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t nvec) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        const uint4 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < nvec; i += stride) dst[i] = src[i];
}

__global__ void __launch_bounds__(256)
compare_flat_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, int64_t nvec, int32_t* __restrict__ flag) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool diff = false;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        const uint4 x0 = a[i], x1 = a[i + stride], x2 = a[i + 2 * stride], x3 = a[i + 3 * stride];
        const uint4 y0 = b[i], y1 = b[i + stride], y2 = b[i + 2 * stride], y3 = b[i + 3 * stride];
        diff |= (x0.x != y0.x) | (x0.y != y0.y) | (x0.z != y0.z) | (x0.w != y0.w) | (x1.x != y1.x) | (x1.y != y1.y) |
                (x1.z != y1.z) | (x1.w != y1.w) | (x2.x != y2.x) | (x2.y != y2.y) | (x2.z != y2.z) | (x2.w != y2.w) |
                (x3.x != y3.x) | (x3.y != y3.y) | (x3.z != y3.z) | (x3.w != y3.w);
    }
    for (; i < nvec; i += stride) {
        const uint4 x = a[i], y = b[i];
        diff |= (x.x != y.x) | (x.y != y.y) | (x.z != y.z) | (x.w != y.w);
    }
    if (__syncthreads_or(diff) && threadIdx.x == 0) atomicOr(flag, 1);
}

// ------------------------------------------------------------------ exact (bitwise) comparison of two 16-bit batches
// flag |= 1 when any element differs.  Used to verify that a batch handed to a second GPTQ object is the very
// batch already staged by the first one before their Hessian accumulators are shared (gptq.py:38-65 computes
// the same H for q/k/v three times; we compute it once, but only for provably identical inputs).
__global__ void compare_rows_kernel(const uint16_t* __restrict__ a, int64_t lda, const uint16_t* __restrict__ b, int64_t ldb,
                                    int64_t tokens, int K, int vec, int32_t* __restrict__ flag) {
    bool diff = false;
    if (vec) {                                   // K % 8 == 0, both pitches % 8 == 0, both bases 16-byte aligned
        const int kv = K >> 3;
        for (int64_t t = blockIdx.y; t < tokens; t += gridDim.y) {
            const uint4* pa = reinterpret_cast<const uint4*>(a + t * lda);
            const uint4* pb = reinterpret_cast<const uint4*>(b + t * ldb);
            for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < kv; c += gridDim.x * blockDim.x) {
                const uint4 x = pa[c], y = pb[c];
                diff |= (x.x != y.x) | (x.y != y.y) | (x.z != y.z) | (x.w != y.w);
            }
        }
    } else {
        for (int64_t t = blockIdx.y; t < tokens; t += gridDim.y)
            for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < K; c += gridDim.x * blockDim.x)
                diff |= a[t * lda + c] != b[t * ldb + c];
    }
    if (__syncthreads_or(diff) && threadIdx.x == 0) atomicOr(flag, 1);
}

// ------------------------------------------------------------------ Q[:, invperm] + cast
template <typename T>
__global__ void __launch_bounds__(256)
finalize_kernel(const float* __restrict__ Q, int64_t ldq, const uint8_t* __restrict__ codes, int64_t ldc, int R,
                int K, const int32_t* __restrict__ invperm, T* __restrict__ wout, int64_t ldwo,
                uint8_t* __restrict__ cout, int64_t ldco) {
    // one CTA per row (see permute_cols_kernel)
    for (int r = blockIdx.x; r < R; r += gridDim.x) {
        const float* q = Q + (int64_t)r * ldq;
        const uint8_t* cd = codes ? codes + (int64_t)r * ldc : nullptr;
        for (int c = threadIdx.x; c < K; c += blockDim.x) {
            const int src = invperm ? invperm[c] : c;
            if (wout) wout[(int64_t)r * ldwo + c] = from_f32<T>(q[src]);
            if (cout) cout[(int64_t)r * ldco + c] = cd[src];
        }
    }
}

// ------------------------------------------------------------------ codes -> dequantised weights
// w[r][c] = T(scale[r][g(c)] * (code[r][c] - zero[r][g(c)]))  -- quant.py:10 followed by the cast of gptq.py:305: exactly what
// the column loop wrote for that weight, so a rank that received only the codes of another rank's rows (1 byte per weight
// instead of 3) rebuilds bit-identical layer weights.  One CTA per row; 16 codes per thread per step.
template <typename T>
__global__ void __launch_bounds__(256)
dequant_codes_kernel(const uint8_t* __restrict__ codes, int64_t ldc, int R, int K, const float* __restrict__ scale,
                     const float* __restrict__ zero, int ng, const int32_t* __restrict__ gidx, T* __restrict__ wout, int64_t ldwo) {
    for (int r = blockIdx.x; r < R; r += gridDim.x) {
        const uint8_t* cd = codes + (int64_t)r * ldc;
        const float* sr = scale + (int64_t)r * ng;
        const float* zr = zero + (int64_t)r * ng;
        T* wo = wout + (int64_t)r * ldwo;
        for (int c = threadIdx.x; c < K; c += blockDim.x) {
            const int g = (ng > 1) ? gidx[c] : 0;
            wo[c] = from_f32<T>(dequant((float)cd[c], sr[g], zr[g]));
        }
    }
}

// ------------------------------------------------------------------ quantize (quant.py:6-10)
__global__ void quantize_kernel(const float* __restrict__ x, int R, int C, int64_t ldx, const float* __restrict__ scale,
                                const float* __restrict__ zero, int64_t s_row, int64_t s_col, int maxq,
                                float* __restrict__ out, int64_t ldo) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    for (int r = blockIdx.y; r < R; r += gridDim.y) {
        float s = scale[(int64_t)r * s_row + (int64_t)c * s_col];
        float z = zero[(int64_t)r * s_row + (int64_t)c * s_col];
        float w = x[(int64_t)r * ldx + c];
        float q;
        if (maxq < 0) { float code; q = quant_trits(w, s, z, &code); }
        else q = dequant(quant_code(w, s, z, (float)maxq), s, z);
        out[(int64_t)r * ldo + c] = q;
    }
}

static inline dim3 grid_cols_rows(int cols, int rows, int threads) {
    int gy = rows < 1 ? 1 : (rows > 4096 ? 4096 : rows);
    return dim3(cdiv(cols, threads), gy);
}

}  // namespace gq

using namespace gq;

extern "C" int gptq_b200_dead_columns(const float* H, int K, int64_t ldh, uint8_t* dead, void* stream) {
    GQ_ARG(H, 1, "null"); GQ_ARG(K >= 0, 2, "K<0"); GQ_ARG(ldh >= K, 3, "ldh<K"); GQ_ARG(dead, 4, "null");
    if (K == 0) return 0;
    dead_columns_kernel<<<cdiv(K, 256), 256, 0, as_stream(stream)>>>(H, K, ldh, dead);
    GQ_LAUNCHED();
    return 0;
}

extern "C" int gptq_b200_hessian_prepare(float* H, int K, int64_t ldh, float percdamp, float* diag_out,
                                         float* damp_out, int32_t* ndead_out, void* stream) {
    GQ_ARG(H, 1, "null"); GQ_ARG(K > 0, 2, "K<=0"); GQ_ARG(ldh >= K, 3, "ldh<K");
    GQ_ARG(diag_out, 5, "null"); GQ_ARG(damp_out, 6, "null");
    hessian_prepare_kernel<<<1, 1024, 0, as_stream(stream)>>>(H, K, ldh, percdamp, diag_out, damp_out, ndead_out);
    GQ_LAUNCHED();
    return 0;
}

extern "C" int gptq_b200_weights_prepare(const void* weight, int dtype, int R, int K, int64_t ldw, const uint8_t* dead,
                                         float* W32, int64_t ldw32, void* stream) {
    GQ_ARG(weight, 1, "null"); GQ_ARG(dtype >= 0 && dtype <= 2, 2, "dtype"); GQ_ARG(R >= 0, 3, "R<0");
    GQ_ARG(K >= 0, 4, "K<0"); GQ_ARG(ldw >= K, 5, "ldw<K"); GQ_ARG(W32, 7, "null"); GQ_ARG(ldw32 >= K, 8, "ldw32<K");
    if (R == 0 || K == 0) return 0;
    dim3 g = grid_cols_rows(K, R, 256);
    cudaStream_t st = as_stream(stream);
    if (dtype == GPTQ_B200_F16) weights_prepare_kernel<__half><<<g, 256, 0, st>>>((const __half*)weight, R, K, ldw, dead, W32, ldw32);
    else if (dtype == GPTQ_B200_BF16) weights_prepare_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)weight, R, K, ldw, dead, W32, ldw32);
    else weights_prepare_kernel<float><<<g, 256, 0, st>>>((const float*)weight, R, K, ldw, dead, W32, ldw32);
    GQ_LAUNCHED();
    return 0;
}

extern "C" int gptq_b200_argsort_desc(const float* v, int n, int32_t* perm, int32_t* invperm, void* stream) {
    GQ_ARG(v, 1, "null"); GQ_ARG(n >= 0, 2, "n<0"); GQ_ARG(perm, 3, "null"); GQ_ARG(invperm, 4, "null");
    if (n == 0) return 0;
    argsort_desc_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(v, n, perm, invperm);
    GQ_LAUNCHED();
    return 0;
}

extern "C" int gptq_b200_permute_cols(const float* W, int R, int K, int64_t ldw, const int32_t* perm, float* out,
                                      int64_t ldo, void* stream) {
    GQ_ARG(W, 1, "null"); GQ_ARG(R >= 0, 2, "R<0"); GQ_ARG(K >= 0, 3, "K<0"); GQ_ARG(ldw >= K, 4, "ldw<K");
    GQ_ARG(out, 6, "null"); GQ_ARG(ldo >= K, 7, "ldo<K");
    if (R == 0 || K == 0) return 0;
    permute_cols_kernel<<<R < 65535 ? R : 65535, 256, 0, as_stream(stream)>>>(W, R, K, ldw, perm, out, ldo);
    GQ_LAUNCHED();
    return 0;
}

extern "C" int gptq_b200_hessian_gather(const float* H, int K, int64_t ldh, const int32_t* perm, const float* damp,
                                        float* A, int Kp, int64_t lda, void* stream) {
    GQ_ARG(H, 1, "null"); GQ_ARG(K > 0, 2, "K<=0"); GQ_ARG(ldh >= K, 3, "ldh<K"); GQ_ARG(A, 6, "null");
    GQ_ARG(Kp >= K && Kp % 128 == 0, 7, "Kp must be >= K and a multiple of 128"); GQ_ARG(lda >= Kp, 8, "lda<Kp");
    hessian_gather_kernel<<<Kp < 65535 ? Kp : 65535, 256, 0, as_stream(stream)>>>(H, K, ldh, perm, damp, A, Kp, lda);
    GQ_LAUNCHED();
    return 0;
}

extern "C" int gptq_b200_symmetrize(float* H, int K, int64_t ldh, void* stream) {
    GQ_ARG(H, 1, "null"); GQ_ARG(K >= 0, 2, "K<0"); GQ_ARG(ldh >= K, 3, "ldh<K");
    if (K == 0) return 0;
    int nb = cdiv(K, 32);
    symmetrize_kernel<<<dim3(nb, nb), dim3(32, 8), 0, as_stream(stream)>>>(H, K, ldh);
    GQ_LAUNCHED();
    return 0;
}

extern "C" int gptq_b200_tril_pack(const float* S, int K, int64_t lds, float* packed, void* stream) {
    GQ_ARG(S, 1, "null"); GQ_ARG(K >= 0, 2, "K<0"); GQ_ARG(lds >= K, 3, "lds<K"); GQ_ARG(packed, 4, "null");
    if (K == 0) return 0;
    tril_pack_kernel<<<K < 4 * num_sms() ? K : 4 * num_sms(), 256, 0, as_stream(stream)>>>(S, K, lds, packed, 0);
    GQ_LAUNCHED();
    return 0;
}

extern "C" int gptq_b200_tril_unpack(const float* packed, int K, float* S, int64_t lds, void* stream) {
    GQ_ARG(packed, 1, "null"); GQ_ARG(K >= 0, 2, "K<0"); GQ_ARG(S, 3, "null"); GQ_ARG(lds >= K, 4, "lds<K");
    if (K == 0) return 0;
    tril_pack_kernel<<<K < 4 * num_sms() ? K : 4 * num_sms(), 256, 0, as_stream(stream)>>>(S, K, lds, const_cast<float*>(packed), 1);
    GQ_LAUNCHED();
    return 0;
}

extern "C" int gptq_b200_stage_input(const void* src, int dtype, int64_t tokens, int K, int64_t ld_src, void* dst,
                                     int64_t ld_dst, void* stream) {
    GQ_ARG(src, 1, "null"); GQ_ARG(dtype == GPTQ_B200_F16 || dtype == GPTQ_B200_BF16, 2, "dtype must be f16/bf16");
    GQ_ARG(tokens >= 0, 3, "tokens<0"); GQ_ARG(K > 0, 4, "K<=0"); GQ_ARG(ld_src >= K, 5, "ld_src<K");
    GQ_ARG(dst, 6, "null"); GQ_ARG(ld_dst >= K && ld_dst % 8 == 0, 7, "ld_dst must be >= K and a multiple of 8");
    if (tokens == 0) return 0;
    cudaStream_t st = as_stream(stream);
    // Always an SM copy, never cudaMemcpyAsync: a device-to-device memcpy is executed by a copy engine and queues behind
    // whatever host-to-device transfers the caller has in flight on other streams (measured: it serialised the whole
    // Hessian phase behind the activation upload).
    if (K % 8 == 0 && ld_src % 8 == 0 && ld_dst % 8 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        const int kv = K / 8;
        const int64_t ldd_v = ld_dst / 8;
        if (ld_src == K && ld_dst == K) {
            const int64_t nvec = tokens * kv;
            int64_t nb = (nvec + 1023) / 1024; if (nb > num_sms() * 8) nb = num_sms() * 8;
            copy_flat_kernel<<<(int)nb, 256, 0, st>>>((const uint4*)src, (uint4*)dst, nvec);
            GQ_LAUNCHED();
            return 0;
        }
        int gx = cdiv(ldd_v, 256); if (gx > 4) gx = 4;
        dim3 g(gx, (unsigned)(tokens > 1184 ? 1184 : tokens));
        stage_vec_kernel<<<g, 256, 0, st>>>((const uint4*)src, tokens, kv, ld_src / 8, (uint4*)dst, ldd_v);
        GQ_LAUNCHED();
        return 0;
    }
    dim3 g(cdiv(ld_dst, 256), (unsigned)(tokens > 4096 ? 4096 : tokens));
    if (dtype == GPTQ_B200_F16) stage_kernel<__half><<<g, 256, 0, st>>>((const __half*)src, tokens, K, ld_src, (__half*)dst, ld_dst);
    else stage_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)src, tokens, K, ld_src, (__nv_bfloat16*)dst, ld_dst);
    GQ_LAUNCHED();
    return 0;
}

extern "C" int gptq_b200_compare_rows(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t tokens, int K,
                                      int32_t* flag, void* stream) {
    GQ_ARG(a, 1, "null"); GQ_ARG(lda >= K, 2, "lda<K"); GQ_ARG(b, 3, "null"); GQ_ARG(ldb >= K, 4, "ldb<K");
    GQ_ARG(tokens >= 0, 5, "tokens<0"); GQ_ARG(K > 0, 6, "K<=0"); GQ_ARG(flag, 7, "null");
    if (tokens == 0) return 0;
    const int vec = (K % 8 == 0) && (lda % 8 == 0) && (ldb % 8 == 0) && ((reinterpret_cast<uintptr_t>(a) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(b) & 15) == 0);
    if (vec && lda == K && ldb == K) {
        const int64_t nvec = tokens * (K / 8);
        int64_t nb = (nvec + 1023) / 1024; if (nb > num_sms() * 8) nb = num_sms() * 8;
        compare_flat_kernel<<<(int)nb, 256, 0, as_stream(stream)>>>((const uint4*)a, (const uint4*)b, nvec, flag);
        GQ_LAUNCHED();
        return 0;
    }
    const int per_row = vec ? K / 8 : K;
    dim3 g(cdiv(per_row, 256) > 8 ? 8 : cdiv(per_row, 256), (unsigned)(tokens > 2048 ? 2048 : tokens));
    compare_rows_kernel<<<g, 256, 0, as_stream(stream)>>>((const uint16_t*)a, lda, (const uint16_t*)b, ldb, tokens, K, vec, flag);
    GQ_LAUNCHED();
    return 0;
}

extern "C" int gptq_b200_finalize(const float* Q, int64_t ldq, const uint8_t* codes, int64_t ldc, int R, int K,
                                  const int32_t* invperm, void* weight_out, int dtype, int64_t ldwo, uint8_t* codes_out,
                                  int64_t ldco, void* stream) {
    GQ_ARG(Q, 1, "null"); GQ_ARG(ldq >= K, 2, "ldq<K"); GQ_ARG(R >= 0, 5, "R<0"); GQ_ARG(K >= 0, 6, "K<0");
    GQ_ARG(dtype >= 0 && dtype <= 2, 9, "dtype");
    GQ_ARG(!codes_out || codes, 3, "codes_out needs codes");
    if (R == 0 || K == 0) return 0;
    const int g = R < 65535 ? R : 65535;
    cudaStream_t st = as_stream(stream);
    if (dtype == GPTQ_B200_F16) finalize_kernel<__half><<<g, 256, 0, st>>>(Q, ldq, codes, ldc, R, K, invperm, (__half*)weight_out, ldwo, codes_out, ldco);
    else if (dtype == GPTQ_B200_BF16) finalize_kernel<__nv_bfloat16><<<g, 256, 0, st>>>(Q, ldq, codes, ldc, R, K, invperm, (__nv_bfloat16*)weight_out, ldwo, codes_out, ldco);
    else finalize_kernel<float><<<g, 256, 0, st>>>(Q, ldq, codes, ldc, R, K, invperm, (float*)weight_out, ldwo, codes_out, ldco);
    GQ_LAUNCHED();
    return 0;
}

extern "C" int gptq_b200_dequant_codes(const uint8_t* codes, int64_t ldc, int R, int K, const float* scale, const float* zero,
                                       int ngroups, const int32_t* g_idx, void* weight_out, int dtype, int64_t ldwo, void* stream) {
    GQ_ARG(codes, 1, "null"); GQ_ARG(ldc >= K, 2, "ldc<K"); GQ_ARG(R >= 0, 3, "R<0"); GQ_ARG(K >= 0, 4, "K<0");
    GQ_ARG(scale && zero, 5, "null"); GQ_ARG(ngroups >= 1, 7, "ngroups<1"); GQ_ARG(ngroups == 1 || g_idx, 8, "g_idx required when ngroups > 1");
    GQ_ARG(weight_out, 9, "null"); GQ_ARG(dtype >= 0 && dtype <= 2, 10, "dtype"); GQ_ARG(ldwo >= K, 11, "ldwo<K");
    if (R == 0 || K == 0) return 0;
    const int g = R < 65535 ? R : 65535;
    cudaStream_t st = as_stream(stream);
    if (dtype == GPTQ_B200_F16) dequant_codes_kernel<__half><<<g, 256, 0, st>>>(codes, ldc, R, K, scale, zero, ngroups, g_idx, (__half*)weight_out, ldwo);
    else if (dtype == GPTQ_B200_BF16) dequant_codes_kernel<__nv_bfloat16><<<g, 256, 0, st>>>(codes, ldc, R, K, scale, zero, ngroups, g_idx, (__nv_bfloat16*)weight_out, ldwo);
    else dequant_codes_kernel<float><<<g, 256, 0, st>>>(codes, ldc, R, K, scale, zero, ngroups, g_idx, (float*)weight_out, ldwo);
    GQ_LAUNCHED();
    return 0;
}

extern "C" int gptq_b200_quantize(const float* x, int R, int C, int64_t ldx, const float* scale, const float* zero,
                                  int64_t s_row, int64_t s_col, int maxq, float* out, int64_t ldo, void* stream) {
    GQ_ARG(x, 1, "null"); GQ_ARG(R >= 0, 2, "R<0"); GQ_ARG(C >= 0, 3, "C<0"); GQ_ARG(ldx >= C, 4, "ldx<C");
    GQ_ARG(scale && zero, 5, "null"); GQ_ARG(out, 10, "null"); GQ_ARG(ldo >= C, 11, "ldo<C");
    if (R == 0 || C == 0) return 0;
    quantize_kernel<<<grid_cols_rows(C, R, 256), 256, 0, as_stream(stream)>>>(x, R, C, ldx, scale, zero, s_row, s_col, maxq, out, ldo);
    GQ_LAUNCHED();
    return 0;
}
