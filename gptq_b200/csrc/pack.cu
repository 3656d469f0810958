// (4) Integer packing into the reference's quant_cuda layouts.
//
// Layout (reference quant.py:161-186 for 3 bits, zeroShot/models/quant.py:184-185 for 4 bits):
// qweight is [K*bits/32, R] int32.  For every 32 consecutive input channels k and every output
// row r the codes form an LSB-first bit stream of 32*bits bits, cut into `bits` little-endian
// 32-bit words stored at qweight[(k/32)*bits + w][r].  For bits in {2,4,8} this is exactly
// qweight[i / (32/bits)] |= code[i] << bits*(i % (32/bits)); for bits=3 codes 10 and 21 straddle
// two words, as in the reference's unrolled loop and in quant_cuda_kernel.cu:116-162.
//
// HBM-bound transpose-and-pack: a CTA stages a [64 rows x 128 channels] byte tile in shared
// memory with 16-byte coalesced loads along K, then each thread builds the words of one
// (row, 32-channel) stream and stores them coalesced along R.
#include "common.cuh"

namespace gq {

constexpr int PK_ROWS = 64;      // output rows (R) per CTA
constexpr int PK_COLS = 128;     // input channels (K) per CTA
constexpr int PK_PITCH = 33;     // smem row pitch in 32-bit words (32 + 1: conflict-free column reads)

template <int BITS>
__device__ __forceinline__ void emit_stream(const uint32_t* __restrict__ srow, int32_t* __restrict__ out, int64_t stride) {
    // srow: 8 words = 32 codes (one byte each).  out[w*stride], w < BITS.
    uint64_t acc = 0;
    int nb = 0, w = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        uint32_t c = (srow[j >> 2] >> (8 * (j & 3))) & ((1u << BITS) - 1u);
        acc |= (uint64_t)c << nb;
        nb += BITS;
        if (nb >= 32) {
            out[(int64_t)w * stride] = (int32_t)(uint32_t)acc;
            acc >>= 32;
            nb -= 32;
            ++w;
        }
    }
}

// SRC: 0 = uint8 codes; 1 = weights of type T recovered through quant.py:153,158.
template <int BITS, typename T, bool FROM_WEIGHTS>
__global__ void __launch_bounds__(256)
pack_kernel(const void* __restrict__ src_, int R, int K, int64_t ld, const float* __restrict__ scale,
            const float* __restrict__ zero, int ngroups, const int32_t* __restrict__ g_idx,
            int32_t* __restrict__ qweight) {
    __shared__ uint32_t tile[PK_ROWS * PK_PITCH];
    const int r0 = blockIdx.y * PK_ROWS;
    const int k0 = blockIdx.x * PK_COLS;
    const int tid = threadIdx.x;

    if constexpr (!FROM_WEIGHTS) {
        const uint8_t* src = static_cast<const uint8_t*>(src_);
        // 64 rows x 128 B = 512 x 16 B; 2 per thread.
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            int v = tid + it * 256;
            int r = v >> 3, c16 = v & 7;
            int gr = r0 + r, gk = k0 + c16 * 16;
            uint4 val = make_uint4(0, 0, 0, 0);
            if (gr < R && gk < K) {
                const uint8_t* p = src + (int64_t)gr * ld + gk;
                if ((((uintptr_t)p) & 15) == 0 && gk + 16 <= K) {
                    val = *reinterpret_cast<const uint4*>(p);
                } else {
                    uint32_t w[4] = {0, 0, 0, 0};
                    for (int b = 0; b < 16 && gk + b < K; ++b) w[b >> 2] |= (uint32_t)p[b] << (8 * (b & 3));
                    val = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            uint32_t* d = &tile[r * PK_PITCH + c16 * 4];
            d[0] = val.x; d[1] = val.y; d[2] = val.z; d[3] = val.w;
        }
    } else {
        const T* src = static_cast<const T*>(src_);
        // each thread converts 8 consecutive channels of one row per step (one 16-byte load for 16-bit weights):
        // 64 x 128 / 8 = 1024 octets, 4 per thread.  The (scale, zero) pair is reloaded only when the group changes --
        // groups are runs of consecutive channels except for act-order with dynamic groups.
        const bool vec16 = (sizeof(T) == 2) && ((ld & 7) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int v = tid + it * 256;
            const int r = v >> 4, q8 = v & 15;
            const int gr = r0 + r, gk = k0 + q8 * 8;
            uint32_t word[2] = {0u, 0u};
            if (gr < R && gk < K) {
                float w[8];
                const T* p = src + (int64_t)gr * ld + gk;
                bool loaded = false;
                if constexpr (sizeof(T) == 2) {
                    if (vec16 && gk + 8 <= K) {
                        const uint4 raw = *reinterpret_cast<const uint4*>(p);
                        const T* h = reinterpret_cast<const T*>(&raw);
#pragma unroll
                        for (int b = 0; b < 8; ++b) w[b] = to_f32<T>(h[b]);
                        loaded = true;
                    }
                }
                if (!loaded) {
#pragma unroll
                    for (int b = 0; b < 8; ++b) w[b] = (gk + b < K) ? to_f32<T>(p[b]) : 0.f;
                }
                int gcur = -1;
                float sc = 1.f, zs = 0.f;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const int k = gk + b;
                    if (k < K) {
                        const int g = g_idx ? g_idx[k] : 0;
                        if (g != gcur) {
                            gcur = g;
                            sc = scale[(int64_t)gr * ngroups + g];
                            zs = __fmul_rn(zero[(int64_t)gr * ngroups + g], sc);      // quant.py:153: zeros * scales
                        }
                        const int ci = (int)rintf(__fdiv_rn(__fadd_rn(w[b], zs), sc)); // quant.py:158: round((W + zeros) / scales)
                        word[b >> 2] |= ((uint32_t)ci & 0xFFu) << (8 * (b & 3));
                    }
                }
            }
            tile[r * PK_PITCH + q8 * 2] = word[0];
            tile[r * PK_PITCH + q8 * 2 + 1] = word[1];
        }
    }
    __syncthreads();

    const int r = tid & (PK_ROWS - 1);
    const int g = tid >> 6;                 // 0..3: which 32-channel stream of the tile
    const int gr = r0 + r;
    const int gk = k0 + g * 32;
    if (gr < R && gk < K) {
        int32_t* out = qweight + ((int64_t)(gk / 32) * BITS) * R + gr;
        emit_stream<BITS>(&tile[r * PK_PITCH + g * 8], out, R);
    }
}

template <int BITS>
__global__ void __launch_bounds__(256)
unpack_kernel(const int32_t* __restrict__ qweight, int R, int K, uint8_t* __restrict__ codes, int64_t ldc) {
    // one thread per (r, 32-channel stream); r fastest for coalesced reads
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int ng = K / 32;
    if (idx >= (int64_t)R * ng) return;
    int r = (int)(idx % R), g = (int)(idx / R);
    uint64_t acc = 0;
    int nb = 0, w = 0;
    const int32_t* in = qweight + ((int64_t)g * BITS) * R + r;
    uint8_t* out = codes + (int64_t)r * ldc + g * 32;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if (nb < BITS) {
            acc |= (uint64_t)(uint32_t)in[(int64_t)w * R] << nb;
            nb += 32;
            ++w;
        }
        out[j] = (uint8_t)(acc & ((1u << BITS) - 1u));
        acc >>= BITS;
        nb -= BITS;
    }
}

template <typename T, bool FW>
static int launch_pack(int bits, const void* src, int R, int K, int64_t ld, const float* scale, const float* zero,
                       int ngroups, const int32_t* g_idx, int32_t* qweight, cudaStream_t st) {
    dim3 grid(cdiv(K, PK_COLS), cdiv(R, PK_ROWS));
    switch (bits) {
        case 2: pack_kernel<2, T, FW><<<grid, 256, 0, st>>>(src, R, K, ld, scale, zero, ngroups, g_idx, qweight); break;
        case 3: pack_kernel<3, T, FW><<<grid, 256, 0, st>>>(src, R, K, ld, scale, zero, ngroups, g_idx, qweight); break;
        case 4: pack_kernel<4, T, FW><<<grid, 256, 0, st>>>(src, R, K, ld, scale, zero, ngroups, g_idx, qweight); break;
        case 8: pack_kernel<8, T, FW><<<grid, 256, 0, st>>>(src, R, K, ld, scale, zero, ngroups, g_idx, qweight); break;
        default: return set_err(-5, "pack: bits must be 2, 3, 4 or 8 (got %d)", bits);
    }
    GQ_LAUNCHED();
    return 0;
}

}  // namespace gq

using namespace gq;

extern "C" int gptq_b200_pack_codes(const uint8_t* codes, int R, int K, int64_t ldc, int bits,
                                    int32_t* qweight, void* stream) {
    GQ_ARG(codes != nullptr, 1, "null");
    GQ_ARG(R >= 0, 2, "R < 0");
    GQ_ARG(K >= 0 && K % 32 == 0, 3, "K must be a multiple of 32");
    GQ_ARG(ldc >= K, 4, "ldc < K");
    GQ_ARG(qweight != nullptr, 6, "null");
    if (R == 0 || K == 0) return 0;
    return launch_pack<uint8_t, false>(bits, codes, R, K, ldc, nullptr, nullptr, 1, nullptr, qweight, as_stream(stream));
}

extern "C" int gptq_b200_pack_weights(const void* W, int dtype, int R, int K, int64_t ldw,
                                      const float* scale, const float* zero, int ngroups,
                                      const int32_t* g_idx, int bits, int32_t* qweight, void* stream) {
    GQ_ARG(W != nullptr, 1, "null");
    GQ_ARG(dtype >= 0 && dtype <= 2, 2, "dtype");
    GQ_ARG(K >= 0 && K % 32 == 0, 4, "K must be a multiple of 32");
    GQ_ARG(ldw >= K, 5, "ldw < K");
    GQ_ARG(scale && zero, 6, "null scale/zero");
    GQ_ARG(ngroups >= 1, 8, "ngroups < 1");
    GQ_ARG(qweight != nullptr, 11, "null");
    if (R == 0 || K == 0) return 0;
    cudaStream_t st = as_stream(stream);
    if (dtype == GPTQ_B200_F16) return launch_pack<__half, true>(bits, W, R, K, ldw, scale, zero, ngroups, g_idx, qweight, st);
    if (dtype == GPTQ_B200_BF16) return launch_pack<__nv_bfloat16, true>(bits, W, R, K, ldw, scale, zero, ngroups, g_idx, qweight, st);
    return launch_pack<float, true>(bits, W, R, K, ldw, scale, zero, ngroups, g_idx, qweight, st);
}

extern "C" int gptq_b200_unpack_codes(const int32_t* qweight, int R, int K, int bits,
                                      uint8_t* codes, int64_t ldc, void* stream) {
    GQ_ARG(qweight != nullptr, 1, "null");
    GQ_ARG(K >= 0 && K % 32 == 0, 3, "K must be a multiple of 32");
    GQ_ARG(codes != nullptr, 5, "null");
    GQ_ARG(ldc >= K, 6, "ldc < K");
    if (R == 0 || K == 0) return 0;
    int64_t n = (int64_t)R * (K / 32);
    int grid = cdiv(n, 256);
    cudaStream_t st = as_stream(stream);
    switch (bits) {
        case 2: unpack_kernel<2><<<grid, 256, 0, st>>>(qweight, R, K, codes, ldc); break;
        case 3: unpack_kernel<3><<<grid, 256, 0, st>>>(qweight, R, K, codes, ldc); break;
        case 4: unpack_kernel<4><<<grid, 256, 0, st>>>(qweight, R, K, codes, ldc); break;
        case 8: unpack_kernel<8><<<grid, 256, 0, st>>>(qweight, R, K, codes, ldc); break;
        default: return set_err(-4, "unpack: bits must be 2, 3, 4 or 8 (got %d)", bits);
    }
    GQ_LAUNCHED();
    return 0;
}
