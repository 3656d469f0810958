// (1) Hessian accumulation  H <- beta*H + alpha * X^T X  on the 5th-generation tensor cores.
//
// Reference: GPTQ.add_batch, gptq.py:59-65 -- one fp32 cuBLAS sgemm [K x tokens]x[tokens x K] per
// calibration sample, full square, TF32 off.  Here: fp16/bf16 operands straight from the
// activations, fp32 accumulation in TMEM, lower-triangular tiles only.
//
// Data flow per CTA (persistent, one CTA per SM, 320 threads):
//   warp 0   TMA producer: X is [tokens, K] row-major, so BOTH operands of X^T X are "MN-major"
//            (channel index contiguous).  One stage = 64 tokens: A = 2 panels, B = 4 panels of
//            [64 tokens x 64 channels] (8 KB each, TMA box {64,64}, SWIZZLE_128B).  Tiles on the
//            diagonal (A's channels are a subset of B's) reuse B's panels and skip the A loads.
//   warp 1   MMA issuer: 4 x tcgen05.mma.cta_group::1.kind::f16 (M=128, N=256, K=16) per stage,
//            accumulating into one of two 256-column TMEM buffers; tcgen05.commit frees the stage.
//   warps 2-9 epilogue: tcgen05.ld each finished chunk accumulator (thread = one row x 128 columns)
//            and add it, round-to-nearest, into 128 fp32 registers; at the end of the item
//            H = beta*H + alpha*acc with 128-byte contiguous accesses per thread.
//            Runs while the MMA warp fills the other TMEM buffer.
// Work item = (tile, split): the token range of a split is walked in chunks of `chunk` tokens.
// The tensor core's fp32 accumulator truncates (measured on B200: relative error ~ 6e-8 x chain
// length in MMAs, biased low), so a chunk is kept short (default 512 tokens = 32 MMAs) and the
// chunks are summed in registers with IEEE round-to-nearest adds.  When there are fewer tiles than
// SMs the token range is split across CTAs and a second kernel reduces the partials in a fixed
// order (deterministic; no atomics).
#include "common.cuh"
#include "ptx.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace gq {

constexpr int SY_BM = 128;          // H rows per tile (MMA M)
constexpr int SY_BN = 256;          // H cols per tile (MMA N)
constexpr int SY_BK = 64;           // tokens per pipeline stage
constexpr int SY_STAGES = 4;
constexpr int SY_PANEL = 64 * 128;  // bytes: 64 tokens x 64 channels x 2 B
constexpr int SY_A_BYTES = 2 * SY_PANEL;
constexpr int SY_B_BYTES = 4 * SY_PANEL;
constexpr int SY_STAGE_BYTES = SY_A_BYTES + SY_B_BYTES;
constexpr int SY_THREADS = 320;         // TMA warp, MMA warp, 8 epilogue warps
constexpr int SY_SMEM = SY_STAGES * SY_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

struct SyrkParams {
    int K; int64_t T; int nbm; int n_tiles; int splits;
    int64_t tokens_per_split;   // multiple of SY_BK
    int chunk;                  // multiple of SY_BK
    float* H; int64_t ldh; float alpha; float beta;
    float* partial;             // [splits][n_tiles][128][256] when splits > 1
    uint32_t idesc;
    int sw, sh;                 // tile-order chunk: column blocks x row blocks (see decode_tile)
};

// Tile order.  Valid tiles are (bi, bj) with bi >= 2 bj (128-row x 256-column tiles touching the lower triangle).
// The CTAs of one wave run consecutive items in lockstep over the tokens, so what matters for HBM traffic is how many
// DISTINCT channel panels a wave touches: enumerating column by column makes a wave read ~nbm row panels (all of X,
// over and over: ncu measured 10.8 GB of DRAM reads for the 721 MB slab at K = 11008).  Instead the tiles are walked
// in 2-D chunks of sw column blocks x sh row blocks (default 12 x 12 = 144 tiles, ~one wave: 12 x 128 + 12 x 256 = 4608
// channels; 6 x 24 and 16 x 9 measured 1-2 % slower).

__device__ __forceinline__ void decode_tile(int t, int nbm, int SY_SW, int SY_SH, int& bi, int& bj) {
    const int nbn = (nbm + 1) >> 1;
    for (int sc = 0; sc < nbn; sc += SY_SW) {
        const int w = min(SY_SW, nbn - sc);
        for (int r0 = 2 * sc; r0 < nbm; r0 += SY_SH) {
            const int r1 = min(r0 + SY_SH, nbm);
            for (int j = sc; j < sc + w; ++j) {
                const int lo = max(r0, 2 * j);
                const int cnt = r1 - lo;
                if (cnt <= 0) continue;
                if (t < cnt) { bj = j; bi = lo + t; return; }
                t -= cnt;
            }
        }
    }
    bi = nbm - 1; bj = 0;          // unreachable for t < n_tiles
}

__global__ void __launch_bounds__(SY_THREADS, 1)
syrk_tc_kernel(const __grid_constant__ CUtensorMap tmap, const SyrkParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SY_STAGES * SY_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + SY_STAGES;
    uint64_t* tfull_bar = empty_bar + SY_STAGES;     // [2]
    uint64_t* tempty_bar = tfull_bar + 2;            // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmap);
        for (int s = 0; s < SY_STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { ptx::mbar_init(&tfull_bar[b], 1); ptx::mbar_init(&tempty_bar[b], 8); }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_items = p.n_tiles * p.splits;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int tile = item / p.splits, split = item - tile * p.splits;
                int bi, bj; decode_tile(tile, p.nbm, p.sw, p.sh, bi, bj);
                const bool a_in_b = (bi >> 1) == bj;
                const int64_t t0 = (int64_t)split * p.tokens_per_split;
                const int64_t t1 = min(t0 + p.tokens_per_split, p.T);
                for (int64_t t = t0; t < t1; t += SY_BK) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sA = smem + stage * SY_STAGE_BYTES;
                    uint8_t* sB = sA + SY_A_BYTES;
                    ptx::mbar_arrive_expect_tx(&full_bar[stage], a_in_b ? SY_B_BYTES : SY_STAGE_BYTES);
                    if (!a_in_b) {
                        ptx::tma_load_2d(sA, &tmap, &full_bar[stage], bi * SY_BM, (int)t);
                        ptx::tma_load_2d(sA + SY_PANEL, &tmap, &full_bar[stage], bi * SY_BM + 64, (int)t);
                    }
#pragma unroll
                    for (int pnl = 0; pnl < 4; ++pnl)
                        ptx::tma_load_2d(sB + pnl * SY_PANEL, &tmap, &full_bar[stage], bj * SY_BN + pnl * 64, (int)t);
                    if (++stage == SY_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int buf = 0; uint32_t bphase[2] = {0, 0};
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int tile = item / p.splits, split = item - tile * p.splits;
                int bi, bj; decode_tile(tile, p.nbm, p.sw, p.sh, bi, bj);
                const bool a_in_b = (bi >> 1) == bj;
                const int a_off = a_in_b ? (SY_A_BYTES + (bi & 1) * 2 * SY_PANEL) : 0;
                const int64_t t0 = (int64_t)split * p.tokens_per_split;
                const int64_t t1 = min(t0 + p.tokens_per_split, p.T);
                for (int64_t c0 = t0; c0 < t1; c0 += p.chunk) {
                    const int64_t c1 = min(c0 + (int64_t)p.chunk, t1);
                    ptx::mbar_wait(&tempty_bar[buf], bphase[buf] ^ 1);
                    ptx::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + buf * SY_BN;
                    uint32_t acc = 0;
                    for (int64_t t = c0; t < c1; t += SY_BK) {
                        ptx::mbar_wait(&full_bar[stage], phase);
                        ptx::tc_fence_after();
                        const uint32_t sbase = ptx::smem_u32(smem + stage * SY_STAGE_BYTES);
#pragma unroll
                        for (int ks = 0; ks < SY_BK / 16; ++ks) {
                            // 16 tokens = 2 K-atoms of 8 rows x 128 B
                            const uint64_t ad = ptx::make_smem_desc_sw128(sbase + a_off + ks * 2048, SY_PANEL, 1024);
                            const uint64_t bd = ptx::make_smem_desc_sw128(sbase + SY_A_BYTES + ks * 2048, SY_PANEL, 1024);
                            ptx::mma_f16_ss(d_tmem, ad, bd, p.idesc, acc);
                            acc = 1;
                        }
                        ptx::tc_commit(&empty_bar[stage]);       // stage reusable once these MMAs retire
                        if (++stage == SY_STAGES) { stage = 0; phase ^= 1; }
                    }
                    ptx::tc_commit(&tfull_bar[buf]);             // accumulator complete -> epilogue
                    bphase[buf] ^= 1;
                    buf ^= 1;
                }
            }
        }
    } else {
        // ===================================================== epilogue (8 warps)
        // warp -> (TMEM lane quadrant = warp % 4, column half): thread = one row x 128 columns, kept in
        // 128 fp32 registers across the chunks of an item (second-level accumulator, round-to-nearest).
        const int ew = warp - 2;
        const int quad = warp & 3;
        const int half = ew >> 2;
        int buf = 0; uint32_t bphase[2] = {0, 0};
        const bool hvec = ((p.ldh & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.H) & 15) == 0);
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int tile = item / p.splits, split = item - tile * p.splits;
            int bi, bj; decode_tile(tile, p.nbm, p.sw, p.sh, bi, bj);
            const int64_t t0 = (int64_t)split * p.tokens_per_split;
            const int64_t t1 = min(t0 + p.tokens_per_split, p.T);
            float acc[128];
#pragma unroll
            for (int q = 0; q < 128; ++q) acc[q] = 0.f;
            for (int64_t c0 = t0; c0 < t1; c0 += p.chunk) {
                ptx::mbar_wait(&tfull_bar[buf], bphase[buf]);
                ptx::tc_fence_after();
                const uint32_t taddr = tmem_base + buf * SY_BN + half * 128 + ((uint32_t)(quad * 32) << 16);
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) {
                    float v[32];
                    ptx::tmem_ld_32x32(taddr + cb * 32, v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 32; ++q) acc[cb * 32 + q] = __fadd_rn(acc[cb * 32 + q], v[q]);
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&tempty_bar[buf]);
                bphase[buf] ^= 1;
                buf ^= 1;
            }
            // ---- fold the item into H (or into its split's partial tile)
            const int row = bi * SY_BM + quad * 32 + lane;
            const int col0 = bj * SY_BN + half * 128;
            float* dst; float alpha, beta; bool vec; int rows_ok, cols_ok;
            if (p.splits == 1) {
                dst = p.H + (int64_t)row * p.ldh + col0; alpha = p.alpha; beta = p.beta; vec = hvec;
                rows_ok = row < p.K; cols_ok = min(128, p.K - col0);
            } else {
                dst = p.partial + ((int64_t)split * p.n_tiles + tile) * (SY_BM * SY_BN) + (int64_t)(quad * 32 + lane) * SY_BN + half * 128;
                alpha = 1.f; beta = 0.f; vec = true; rows_ok = 1; cols_ok = 128;
            }
            if (rows_ok) {
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) {
                    float* o = dst + cb * 32;
                    const int nc = cols_ok - cb * 32;
                    if (nc >= 32 && vec) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float4 r;
                            if (beta != 0.f) {
                                const float4 h = *reinterpret_cast<const float4*>(o + q * 4);
                                r.x = fmaf(alpha, acc[cb * 32 + q * 4 + 0], beta * h.x); r.y = fmaf(alpha, acc[cb * 32 + q * 4 + 1], beta * h.y);
                                r.z = fmaf(alpha, acc[cb * 32 + q * 4 + 2], beta * h.z); r.w = fmaf(alpha, acc[cb * 32 + q * 4 + 3], beta * h.w);
                            } else {
                                r.x = alpha * acc[cb * 32 + q * 4 + 0]; r.y = alpha * acc[cb * 32 + q * 4 + 1];
                                r.z = alpha * acc[cb * 32 + q * 4 + 2]; r.w = alpha * acc[cb * 32 + q * 4 + 3];
                            }
                            *reinterpret_cast<float4*>(o + q * 4) = r;
                        }
                    } else if (nc > 0) {
#pragma unroll
                        for (int q = 0; q < 32; ++q)
                            if (q < nc) o[q] = (beta != 0.f) ? fmaf(alpha, acc[cb * 32 + q], beta * o[q]) : alpha * acc[cb * 32 + q];
                    }
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

// H = beta*H + alpha * sum_s partial[s][tile]   (fixed summation order)
__global__ void __launch_bounds__(256)
syrk_reduce_kernel(const float* __restrict__ partial, int n_tiles, int splits, int nbm, int K, float* __restrict__ H,
                   int64_t ldh, float alpha, float beta, int sw, int sh) {
    const int tile = blockIdx.x;
    int bi, bj; decode_tile(tile, nbm, sw, sh, bi, bj);
    for (int e = threadIdx.x; e < SY_BM * SY_BN; e += 256) {
        int r = e >> 8, c = e & 255;
        int gr = bi * SY_BM + r, gc = bj * SY_BN + c;
        if (gr >= K || gc >= K) continue;
        float s = 0.f;
        for (int sp = 0; sp < splits; ++sp) s += partial[((int64_t)sp * n_tiles + tile) * (SY_BM * SY_BN) + e];
        float* o = H + (int64_t)gr * ldh + gc;
        *o = (beta != 0.f) ? fmaf(alpha, s, beta * *o) : alpha * s;
    }
}

// ---------------------------------------------------------------- SIMT cross-check (tests only)
template <typename T>
__global__ void __launch_bounds__(256)
syrk_simt_kernel(const T* __restrict__ X, int64_t tokens, int K, int64_t ldx, float* __restrict__ H, int64_t ldh, float alpha,
                 float beta) {
    __shared__ float xa[32][33], xb[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi) return;                                   // lower triangle only (same contract as the TC kernel)
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;    // 32 x 8
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int64_t t0 = 0; t0 < tokens; t0 += 32) {
        for (int r = ty; r < 32; r += 8) {
            int64_t t = t0 + r;
            int ca = bi * 32 + tx, cb = bj * 32 + tx;
            xa[r][tx] = (t < tokens && ca < K) ? to_f32<T>(X[t * ldx + ca]) : 0.f;
            xb[r][tx] = (t < tokens && cb < K) ? to_f32<T>(X[t * ldx + cb]) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            float b = xb[k][tx];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] = fmaf(xa[k][ty + 8 * q], b, acc[q]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        int gr = bi * 32 + ty + 8 * q, gc = bj * 32 + tx;
        if (gr < K && gc < K) {
            float* o = H + (int64_t)gr * ldh + gc;
            *o = (beta != 0.f) ? fmaf(alpha, acc[q], beta * *o) : alpha * acc[q];
        }
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

struct SyrkPlan { int nbm, nbn, n_tiles, splits; int64_t tokens_per_split; };

static SyrkPlan make_plan(int K, int64_t T) {
    SyrkPlan pl;
    pl.nbm = cdiv(K, SY_BM);
    pl.nbn = cdiv(K, SY_BN);
    pl.n_tiles = 0;
    for (int j = 0; j < pl.nbn; ++j) pl.n_tiles += pl.nbm - 2 * j;
    const int sms = num_sms();
    int64_t kblocks = (T + SY_BK - 1) / SY_BK;
    int splits = 1;
    if (pl.n_tiles < sms) {
        splits = sms / pl.n_tiles;
        int64_t max_splits = (kblocks + 7) / 8;            // at least 8 k-blocks (512 tokens) per split
        if (splits > max_splits) splits = (int)max_splits;
        if (splits < 1) splits = 1;
    }
    int64_t kb_per = (kblocks + splits - 1) / splits;
    pl.tokens_per_split = kb_per * SY_BK;
    pl.splits = (int)((kblocks + kb_per - 1) / kb_per);
    if (pl.splits < 1) pl.splits = 1;
    return pl;
}

// syrk_tc2.cu: the CTA-pair (cta_group::2) kernel
size_t syrk2_workspace_bytes(int K, int64_t tokens);
int syrk2_launch(const CUtensorMap& tmap, int bk, int dtype, int64_t tokens, int K, float* H, int64_t ldh, float alpha, float beta,
                 int chunk, int sw, int sh, void* ws, size_t ws_bytes, cudaStream_t st);

static size_t syrk1_workspace_bytes(int K, int64_t tokens) {
    if (K <= 0 || tokens <= 0) return 0;
    SyrkPlan pl = make_plan(K, tokens);
    if (pl.splits <= 1) return 0;
    return (size_t)pl.splits * pl.n_tiles * SY_BM * SY_BN * sizeof(float);
}

}  // namespace gq

using namespace gq;

extern "C" size_t gptq_b200_syrk_workspace_bytes(int K, int64_t tokens, int variant) {
    variant &= 0xFF;
    return variant == GPTQ_B200_SYRK_CTA1 ? syrk1_workspace_bytes(K, tokens) : syrk2_workspace_bytes(K, tokens);
}

extern "C" int gptq_b200_syrk_accum(const void* X, int dtype, int64_t tokens, int K, int64_t ldx, float* H, int64_t ldh,
                                    float alpha, float beta, int chunk_tokens, int variant, void* ws, size_t ws_bytes,
                                    void* stream) {
    GQ_ARG(X, 1, "null"); GQ_ARG(dtype == GPTQ_B200_F16 || dtype == GPTQ_B200_BF16, 2, "dtype must be f16/bf16");
    GQ_ARG(tokens >= 0 && tokens < (1LL << 31), 3, "tokens out of range"); GQ_ARG(K > 0, 4, "K<=0");
    GQ_ARG(ldx >= K && ldx % 8 == 0, 5, "ldx must be >= K and a multiple of 8 (TMA 16-byte stride)");
    GQ_ARG((reinterpret_cast<uintptr_t>(X) & 15) == 0, 1, "X must be 16-byte aligned");
    GQ_ARG(H, 6, "null"); GQ_ARG(ldh >= K, 7, "ldh<K"); GQ_ARG(chunk_tokens >= 0, 10, "chunk_tokens<0");
    // tuning bits above the kernel selector (perf experiments; 0 = defaults): tile-order chunk sw (bits 8-15) x sh (bits 16-23),
    // bit 24 = TMA L2 promotion 256 B instead of the pair kernel's 128 B (which read 10 % less DRAM at K = 11008)
    const int tune_sw = (variant >> 8) & 0xFF, tune_sh = (variant >> 16) & 0xFF;
    const int tune_p128 = ((variant & 0xFF) == GPTQ_B200_SYRK_CTA1) ? 0 : !((variant >> 24) & 1);
    variant &= 0xFF;
    GQ_ARG(variant >= GPTQ_B200_SYRK_AUTO && variant <= GPTQ_B200_SYRK_CTA2_BK128, 11, "variant must be one of GPTQ_B200_SYRK_*");
    if (tokens == 0) return 0;                 // beta handling for empty input is the caller's business
    cudaStream_t st = as_stream(stream);
    const bool pair = variant != GPTQ_B200_SYRK_CTA1;
    SyrkPlan pl = make_plan(K, tokens);
    const size_t need = pair ? 0 : syrk1_workspace_bytes(K, tokens);      // (the pair kernel checks its own)
    GQ_ARG(need == 0 || (ws && ws_bytes >= need), 13, "workspace too small");

    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return set_err(999, "cuTensorMapEncodeTiled entry point not available");
    CUtensorMap tmap;
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)tokens};
    cuuint64_t gstr[1] = {(cuuint64_t)ldx * 2};
    const int pair_bk = (variant == GPTQ_B200_SYRK_CTA2) ? 64 : 128;
    cuuint32_t box[2] = {64, (cuuint32_t)(pair ? pair_bk : SY_BK)};
    cuuint32_t estr[2] = {1, 1};
    CUresult cr = enc(&tmap, dtype == GPTQ_B200_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                      const_cast<void*>(X), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      tune_p128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return set_err(1000 + (int)cr, "cuTensorMapEncodeTiled failed (%d)", (int)cr);

    int chunk = chunk_tokens > 0 ? chunk_tokens : 512;
    if (pair) return syrk2_launch(tmap, pair_bk, dtype, tokens, K, H, ldh, alpha, beta, chunk, tune_sw, tune_sh, ws, ws_bytes, st);

    GQ_ONCE_PER_DEVICE(GQ_CUDA(cudaFuncSetAttribute(syrk_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SY_SMEM)));
    SyrkParams p;
    p.K = K; p.T = tokens; p.nbm = pl.nbm; p.n_tiles = pl.n_tiles; p.splits = pl.splits;
    p.tokens_per_split = pl.tokens_per_split;
    chunk = ((chunk + SY_BK - 1) / SY_BK) * SY_BK;
    p.chunk = chunk;
    p.H = H; p.ldh = ldh; p.alpha = alpha; p.beta = beta;
    p.partial = static_cast<float*>(ws);
    p.idesc = ptx::make_idesc(dtype == GPTQ_B200_F16 ? 0 : 1, 1, 1, SY_BM, SY_BN);
    p.sw = tune_sw ? tune_sw : 12; p.sh = tune_sh ? tune_sh : 12;          // 144 tiles ~ one wave (6 x 24 and 16 x 9 measured 1-2 % slower)
    const int n_items = pl.n_tiles * pl.splits;
    const int grid = n_items < num_sms() ? n_items : num_sms();
    syrk_tc_kernel<<<grid, SY_THREADS, SY_SMEM, st>>>(tmap, p);
    GQ_LAUNCHED();
    if (pl.splits > 1) {
        syrk_reduce_kernel<<<pl.n_tiles, 256, 0, st>>>(p.partial, pl.n_tiles, pl.splits, pl.nbm, K, H, ldh, alpha, beta, p.sw, p.sh);
        GQ_LAUNCHED();
    }
    return 0;
}

extern "C" int gptq_b200_syrk_accum_simt(const void* X, int dtype, int64_t tokens, int K, int64_t ldx, float* H,
                                         int64_t ldh, float alpha, float beta, void* stream) {
    GQ_ARG(X, 1, "null"); GQ_ARG(dtype >= 0 && dtype <= 2, 2, "dtype must be f16/bf16/f32");
    GQ_ARG(tokens >= 0, 3, "tokens<0"); GQ_ARG(K > 0, 4, "K<=0"); GQ_ARG(ldx >= K, 5, "ldx<K");
    GQ_ARG(H, 6, "null"); GQ_ARG(ldh >= K, 7, "ldh<K");
    int nb = cdiv(K, 32);
    cudaStream_t st = as_stream(stream);
    if (dtype == GPTQ_B200_F32) syrk_simt_kernel<float><<<dim3(nb, nb), 256, 0, st>>>((const float*)X, tokens, K, ldx, H, ldh, alpha, beta);
    else if (dtype == GPTQ_B200_F16) syrk_simt_kernel<__half><<<dim3(nb, nb), 256, 0, st>>>((const __half*)X, tokens, K, ldx, H, ldh, alpha, beta);
    else syrk_simt_kernel<__nv_bfloat16><<<dim3(nb, nb), 256, 0, st>>>((const __nv_bfloat16*)X, tokens, K, ldx, H, ldh, alpha, beta);
    GQ_LAUNCHED();
    return 0;
}
