// Quantizer.find_params for weight layouts (reference quant.py:37-118): per-row, per-group
// min/max -> (scale, zero); optional symmetric grid, trits and the MSE shrink search.
//
// One warp per row walks the requested groups; lanes read the row slice with coalesced
// (float4 when aligned) loads and combine with shuffles.  HBM-bound: 4*R*ncols bytes in,
// 8*R*ngroups bytes out.
#include "common.cuh"

namespace gq {

struct FPArgs {
    const float* W; int R; int K; int64_t ldw; int col0; int ncols; int gs; int maxq; int sym; int mse;
    float norm; int grid; float maxshrink; float* scale; float* zero; int64_t ldsz; int gout0;
};

__device__ __forceinline__ void warp_minmax(float& mn, float& mx) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
}

// quant.py:57-77 on the reduced (min, max) of one row slice.
__device__ __forceinline__ void params_from_minmax(float xmin, float xmax, int maxq, int sym, float& scale, float& zero,
                                                   float& oxmin, float& oxmax) {
    xmin = fminf(xmin, 0.f);
    xmax = fmaxf(xmax, 0.f);
    if (sym) {
        xmax = fmaxf(fabsf(xmin), xmax);
        if (xmin < 0.f) xmin = -xmax;
    }
    if (xmin == 0.f && xmax == 0.f) { xmin = -1.f; xmax = 1.f; }
    if (maxq < 0) { scale = xmax; zero = xmin; }
    else {
        scale = __fdiv_rn(__fsub_rn(xmax, xmin), (float)maxq);
        zero = sym ? (float)(maxq + 1) * 0.5f : rintf(__fdiv_rn(-xmin, scale));
    }
    oxmin = xmin; oxmax = xmax;
}

__global__ void __launch_bounds__(256)
find_params_kernel(FPArgs a) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= a.R) return;
    const float* wr = a.W + (int64_t)row * a.ldw;
    const int gs = a.gs > 0 ? a.gs : a.ncols;
    const int ngr = (a.ncols + gs - 1) / gs;
    for (int g = 0; g < ngr; ++g) {
        const int c0 = a.col0 + g * gs;
        const int c1 = min(c0 + gs, a.K);            // python slice W[:, c0:c0+gs] clips at K
        float mn = INFINITY, mx = -INFINITY;
        const float* p = wr + c0;
        const int n = c1 - c0;
        if (((((uintptr_t)p) & 15) == 0) && (n % 4 == 0)) {
            const float4* p4 = reinterpret_cast<const float4*>(p);
            for (int i = lane; i < n / 4; i += 32) {
                float4 v = p4[i];
                mn = fminf(fminf(mn, v.x), fminf(v.y, fminf(v.z, v.w)));
                mx = fmaxf(fmaxf(mx, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
            }
        } else {
            for (int i = lane; i < n; i += 32) { float v = p[i]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
        }
        warp_minmax(mn, mx);
        float scale, zero, xmin, xmax;
        params_from_minmax(mn, mx, a.maxq, a.sym, scale, zero, xmin, xmax);
        if (a.mse && a.maxq > 0) {
            // quant.py:79-96: shrink search, error = sum |q(x) - x|^norm, first strict improvement wins.
            float best = INFINITY;
            const float fmaxq = (float)a.maxq;
            const int steps = (int)(a.maxshrink * (float)a.grid);
            for (int i = 0; i < steps; ++i) {
                float pfac = (float)(1.0 - (double)i / (double)a.grid);
                float xmin1 = __fmul_rn(pfac, xmin), xmax1 = __fmul_rn(pfac, xmax);
                float scale1 = __fdiv_rn(__fsub_rn(xmax1, xmin1), fmaxq);
                float zero1 = a.sym ? zero : rintf(__fdiv_rn(-xmin1, scale1));
                float err = 0.f;
                for (int j = lane; j < n; j += 32) {
                    float x = p[j];
                    float q = dequant(quant_code(x, scale1, zero1, fmaxq), scale1, zero1);
                    err += powf(fabsf(__fsub_rn(q, x)), a.norm);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) err += __shfl_xor_sync(0xffffffffu, err, o);
                if (err < best) { best = err; scale = scale1; zero = zero1; }
            }
        }
        if (lane == 0) {
            a.scale[(int64_t)row * a.ldsz + a.gout0 + g] = scale;
            a.zero[(int64_t)row * a.ldsz + a.gout0 + g] = zero;
        }
    }
}

}  // namespace gq

using namespace gq;

extern "C" int gptq_b200_find_params(const float* W, int R, int K, int64_t ldw, int col0, int ncols, int groupsize,
                                     int maxq, int sym, int mse, float norm, int grid, float maxshrink, float* scale,
                                     float* zero, int64_t ldsz, int gout0, void* stream) {
    GQ_ARG(W, 1, "null"); GQ_ARG(R >= 0, 2, "R<0"); GQ_ARG(K > 0, 3, "K<=0"); GQ_ARG(ldw >= K, 4, "ldw<K");
    GQ_ARG(col0 >= 0 && col0 < K, 5, "col0 out of range"); GQ_ARG(ncols > 0 && col0 + ncols <= K, 6, "ncols out of range");
    GQ_ARG(maxq != 0, 8, "maxq == 0 (Quantizer not configured)");
    GQ_ARG(scale && zero, 14, "null"); GQ_ARG(ldsz >= 1, 16, "ldsz<1"); GQ_ARG(gout0 >= 0, 17, "gout0<0");
    if (R == 0) return 0;
    FPArgs a{W, R, K, ldw, col0, ncols, groupsize, maxq, sym, mse, norm, grid, maxshrink, scale, zero, ldsz, gout0};
    find_params_kernel<<<cdiv(R, 8), 256, 0, as_stream(stream)>>>(a);
    GQ_LAUNCHED();
    return 0;
}
