// Packed-weight matvec for the layouts pack.cu emits (SURVEY 8f "next" #1):
//     y[col] += sum_k (scale[col,g(k)] * code[k,col] - zs[col,g(k)]) * x[k]
// Reference: quant_cuda_kernel.cu:88-165 (VecQuant3MatMulKernel, fp32) and :167-244 (half2 LUT variant),
// bound through quant_cuda.cpp:15-29 and used by Quant3Linear.forward (quant.py:189-203): single token,
// y pre-loaded with the bias, zs = zero*scale (quant.py:153).  The reference ships 3 bits only; the same
// kernel serves the 2/4/8-bit layouts.
//
// HBM-bound (the packed matrix is read exactly once, coalesced along the output dimension).  A CTA owns
// 128 output columns x 256 input channels: x slab in shared memory, each thread unpacks 32 codes per
// `bits` words and accumulates sum(code*x) in fp32; scale / zero are applied per group of channels
// (once per 32-channel stream when the stream lies inside one group), then one atomicAdd per (thread, slab) like the
// reference.  All packed words of a (column, slab) are loaded before the first is used.
#include "common.cuh"

namespace gq {

constexpr int VQ_COLS = 128;       // output columns per CTA (= threads)
constexpr int VQ_SLAB = 256;       // input channels per CTA

// 32 codes of one (column, 32-channel) stream from its BITS packed words (LSB-first bit stream, pack.cu).
template <int BITS>
__device__ __forceinline__ void unpack32(const uint32_t (&wd)[BITS], uint32_t (&c)[32]) {
    if constexpr (BITS == 3) {
        const uint64_t lo = (uint64_t)wd[0] | ((uint64_t)wd[1] << 32), hi = (uint64_t)wd[1] | ((uint64_t)wd[2] << 32);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int off = 3 * j;
            const uint64_t src = (off <= 61) ? lo : hi;          // code 21 straddles words 1/2: read it from the upper pair
            const int sh = (off <= 61) ? off : off - 32;
            c[j] = (uint32_t)((src >> sh) & 7u);
        }
    } else {
        constexpr int PER = 32 / BITS;
#pragma unroll
        for (int w = 0; w < BITS; ++w) {
#pragma unroll
            for (int j = 0; j < PER; ++j) c[w * PER + j] = (wd[w] >> (BITS * j)) & ((1u << BITS) - 1u);
        }
    }
}

template <int BITS, typename TX>
__global__ void __launch_bounds__(VQ_COLS)
vecquant_kernel(const TX* __restrict__ x, const int32_t* __restrict__ qweight, int R, int K, const float* __restrict__ scales,
                const float* __restrict__ zs, int ng, const int32_t* __restrict__ g_idx, int groupsize, float* __restrict__ y) {
    constexpr int NG32 = VQ_SLAB / 32;           // 32-channel streams per slab
    __shared__ float xs[VQ_SLAB];
    __shared__ int gs_[VQ_SLAB];
    __shared__ int uni[NG32];                    // 1 when the 32 channels of a stream share one group
    const int col = blockIdx.x * VQ_COLS + threadIdx.x;
    const int k0 = blockIdx.y * VQ_SLAB;
    const int nwords = min(VQ_SLAB, K - k0) / 32;
    // every packed word of this (column, slab) is requested before anything is consumed: one memory round trip per CTA
    uint32_t wq[NG32][BITS];
    if (col < R) {
#pragma unroll
        for (int g = 0; g < NG32; ++g)
#pragma unroll
            for (int w = 0; w < BITS; ++w)
                wq[g][w] = (g < nwords) ? (uint32_t)__ldg(qweight + ((int64_t)((k0 / 32 + g) * BITS + w)) * R + col) : 0u;
    }
    for (int i = threadIdx.x; i < VQ_SLAB; i += VQ_COLS) {
        const int k = k0 + i;
        xs[i] = (k < K) ? to_f32<TX>(x[k]) : 0.f;
        gs_[i] = (ng > 1 && k < K) ? (g_idx ? g_idx[k] : k / groupsize) : 0;
    }
    __syncthreads();
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int b = warp; b < NG32; b += VQ_COLS / 32) {
            const bool same = gs_[b * 32 + lane] == gs_[b * 32];
            const bool all = __all_sync(0xffffffffu, same);
            if (lane == 0) uni[b] = all ? 1 : 0;
        }
    }
    __syncthreads();
    if (col >= R) return;
    float res = 0.f;
#pragma unroll
    for (int g = 0; g < NG32; ++g) {
        if (g >= nwords) break;
        uint32_t c[32];
        unpack32<BITS>(wq[g], c);
        const float* xg = xs + g * 32;
        const int* gg = gs_ + g * 32;
        // one (scale, zero) pair for the 32 channels: per-row scales, or a group that covers the whole stream (group sizes
        // that are multiples of 32 in channel order: everything except act-order with dynamic groups)
        const int g0 = gg[0];
        if (ng == 1 || uni[g]) {
            float acc = 0.f, sx = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) { acc = fmaf((float)c[j], xg[j], acc); sx += xg[j]; }
            res += scales[(int64_t)col * ng + g0] * acc - zs[(int64_t)col * ng + g0] * sx;
        } else {                             // scattered groups: one table lookup per channel (static register indexing)
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int64_t o = (int64_t)col * ng + gg[j];
                res = fmaf(fmaf(scales[o], (float)c[j], -zs[o]), xg[j], res);
            }
        }
    }
    atomicAdd(&y[col], res);
}

template <typename TX>
static int launch_vq(int bits, const TX* x, const int32_t* qw, int R, int K, const float* scales, const float* zs, int ng,
                     const int32_t* g_idx, int groupsize, float* y, cudaStream_t st) {
    dim3 grid(cdiv(R, VQ_COLS), cdiv(K, VQ_SLAB));
    switch (bits) {
        case 2: vecquant_kernel<2, TX><<<grid, VQ_COLS, 0, st>>>(x, qw, R, K, scales, zs, ng, g_idx, groupsize, y); break;
        case 3: vecquant_kernel<3, TX><<<grid, VQ_COLS, 0, st>>>(x, qw, R, K, scales, zs, ng, g_idx, groupsize, y); break;
        case 4: vecquant_kernel<4, TX><<<grid, VQ_COLS, 0, st>>>(x, qw, R, K, scales, zs, ng, g_idx, groupsize, y); break;
        case 8: vecquant_kernel<8, TX><<<grid, VQ_COLS, 0, st>>>(x, qw, R, K, scales, zs, ng, g_idx, groupsize, y); break;
        default: return set_err(-6, "vecquant: bits must be 2, 3, 4 or 8 (got %d)", bits);
    }
    GQ_LAUNCHED();
    return 0;
}

}  // namespace gq

using namespace gq;

extern "C" int gptq_b200_vecquant_matvec(const void* x, int dtype, const int32_t* qweight, int R, int K, int bits,
                                         const float* scales, const float* zeros_scaled, int ngroups, const int32_t* g_idx,
                                         int groupsize, float* y, void* stream) {
    GQ_ARG(x, 1, "null"); GQ_ARG(dtype >= 0 && dtype <= 2, 2, "dtype"); GQ_ARG(qweight, 3, "null");
    GQ_ARG(R >= 0, 4, "R<0"); GQ_ARG(K >= 0 && K % 32 == 0, 5, "K must be a multiple of 32");
    GQ_ARG(scales && zeros_scaled, 7, "null"); GQ_ARG(ngroups >= 1, 9, "ngroups<1");
    GQ_ARG(ngroups == 1 || g_idx || groupsize > 0, 10, "grouped scales need g_idx or groupsize");
    GQ_ARG(y, 12, "null");
    if (R == 0 || K == 0) return 0;
    cudaStream_t st = as_stream(stream);
    if (dtype == GPTQ_B200_F16) return launch_vq<__half>(bits, (const __half*)x, qweight, R, K, scales, zeros_scaled, ngroups, g_idx, groupsize, y, st);
    if (dtype == GPTQ_B200_BF16) return launch_vq<__nv_bfloat16>(bits, (const __nv_bfloat16*)x, qweight, R, K, scales, zeros_scaled, ngroups, g_idx, groupsize, y, st);
    return launch_vq<float>(bits, (const float*)x, qweight, R, K, scales, zeros_scaled, ngroups, g_idx, groupsize, y, st);
}
