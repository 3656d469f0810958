// fp32-accurate GEMM on the tensor cores:  C[M,N] <- alpha * op(A) op(B) + beta * C,  fp32 in / out.
//
// Used for the trailing updates of the inverse-Cholesky chain (reference gptq.py:177-179) and for the
// lazy block update W[:, i2:] -= Err1 @ Hinv[i1:i2, i2:] (gptq.py:276).  The reference runs these in
// fp32 with TF32 disabled (gptq.py:18-19); a single TF32 pass is not accurate enough for code parity
// (SURVEY Appendix B), so every operand is split on the fly into hi = tf32(x) and lo = x - hi and the
// product is formed as hi*hi + hi*lo + lo*hi with fp32 accumulation ("3xTF32").
//
// One CTA per 128x256 output tile, 320 threads:
//   warp 0    TMA producer: fp32 operand tiles (BK = 32 = one 128-byte swizzle row) into a 2-stage ring.
//             Either operand may be K-major (k contiguous: box {32 k, 128 rows}) or MN-major
//             (m/n contiguous: boxes {32 mn, 32 k}), so all four transpose combinations run natively.
//   warps 2-9 split the landed stage in place (hi over the raw tile, lo beside it -- same offsets, so the
//             128B swizzle is preserved), fence.proxy.async, signal the MMA warp; the same warps drain
//             the TMEM accumulator every CHUNK k-blocks into 128 fp32 registers per thread with
//             round-to-nearest adds (tcgen05's accumulator truncates: the chain is kept at 12 MMAs),
//             and finally apply alpha/beta and store C.
//   warp 1    MMA issuer: per k-block 4 k-steps x 3 x tcgen05.mma.cta_group::1.kind::tf32 (M128 N256 K8)
//             into one of two 256-column TMEM buffers.
#include "common.cuh"
#include "ptx.cuh"
#include "sgemm.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace gq {

extern int g_gemm_splitk;
extern int g_gemm_splitk_minkb;
extern int g_gemm_trunc;
extern int g_gemm_bk;
extern int g_small_gemm_max;
extern int g_gemm_cprefetch;

constexpr int GT_BM = 128, GT_BN = 256;
constexpr int GT_RING_BYTES = 192 * 1024;             // operand ring: 2 stages of BK = 32 or 4 stages of BK = 16
constexpr int GT_THREADS = 320;
constexpr int GT_SMEM = GT_RING_BYTES + 1024 + 256;

constexpr int GT_TP = GT_BM + 4;                      // pitch of the transposed partial tile in shared memory (split-K):
                                                      // 16-byte aligned columns, conflict-free for float4 reads along the columns

struct GemmTcParams {
    int M, N, Kd; float alpha, beta; float* C; int64_t ldc; int mode; int a_mn, b_mn; uint32_t idesc;
    int trunc;                                        // 1: the MMA reads the raw fp32 tile as `hi` (see the converter)
    int cprefetch;                                    // 1: beta != 0 -> the C tile is pulled into L2 while the main loop runs
    int splits;                                       // cluster size along z: the k-range of a tile is split over the cluster
};

__device__ __forceinline__ float rn_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// GT_BK = k-extent of one pipeline stage.  32: 2 stages of 96 KB (raw + lo of both operands), K-major tiles are one
// 128-byte swizzle row; 16: 4 stages of 48 KB, K-major tiles are 64-byte rows (SWIZZLE_64B) -- same bytes in flight, but
// a TMA round trip overlaps three other stages instead of one.  The accumulation chain is 12 MMAs either way.
// Measured on B200 at 4096^3: BK = 32 -> 160 TFLOP/s, BK = 16 -> 140 TFLOP/s (twice the barrier hand-offs, no gain from the
// deeper ring: the kernel is limited by shared-memory traffic, not by exposed TMA latency), so 32 is the default
// (GPTQ_B200_GEMM_BK=16 selects the other instantiation; both pass the parity tests).
template <int GT_BK>
__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmTcParams p) {
    constexpr int GT_A_BYTES = GT_BM * GT_BK * 4;
    constexpr int GT_B_BYTES = GT_BN * GT_BK * 4;
    constexpr int GT_STAGE_BYTES = 2 * (GT_A_BYTES + GT_B_BYTES);
    constexpr int GT_STAGES = GT_RING_BYTES / GT_STAGE_BYTES;
    constexpr int GT_CHUNK = 32 / GT_BK;              // k-blocks per TMEM accumulation chain: 12 MMAs (24 measured +2 % speed,
                                                      // twice the truncation bias; the drain is not the limiter)
    constexpr int GT_PANEL = 32 * GT_BK * 4;          // one MN-major TMA box: 32 m/n x GT_BK k
    const int m0 = blockIdx.y * GT_BM, n0 = blockIdx.x * GT_BN;
    if ((p.mode & SG_UPPER_TILES) && n0 + GT_BN - 1 < m0) return;      // whole CTA, before any barrier
    int kbeg = 0, kend = p.Kd;
    if (p.mode & SG_A_UPPER) kbeg = (m0 / GT_BK) * GT_BK;
    if (p.mode & SG_B_UPPER) kend = min(p.Kd, n0 + GT_BN);
    int nkb = kend > kbeg ? (kend - kbeg + GT_BK - 1) / GT_BK : 0;
    // split-K: the CTAs of a cluster (same tile, z = rank) take consecutive slices of the k-blocks and reduce their
    // partial tiles through distributed shared memory in rank order (deterministic, no atomics, no workspace)
    const int S = p.splits;
    const int zr = S > 1 ? (int)ptx::cluster_ctarank() : 0;
    if (S > 1) {
        const int per = (nkb + S - 1) / S;
        const int kb0 = zr * per;
        kbeg += kb0 * GT_BK;
        nkb = max(0, min(per, nkb - kb0));
    }

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GT_STAGES * GT_STAGE_BYTES);   // TMA landed
    uint64_t* conv_bar = full_bar + GT_STAGES;                                              // split done
    uint64_t* empty_bar = conv_bar + GT_STAGES;                                             // MMAs retired
    uint64_t* tfull_bar = empty_bar + GT_STAGES;     // [2]
    uint64_t* tempty_bar = tfull_bar + 2;            // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    const int warp = ptx::warp_idx_uniform(), lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmB);
        for (int s = 0; s < GT_STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&conv_bar[s], 8); ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) { ptx::mbar_init(&tfull_bar[b], 1); ptx::mbar_init(&tempty_bar[b], 8); }
        ptx::fence_barrier_init();
    }
    if (warp == 1) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::pdl_launch_dependents();      // the next kernel may start its prologue; it waits for our completion itself
    ptx::pdl_wait();                   // operands (and C when beta != 0) come from the preceding kernels of the stream

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                const int k0 = kbeg + kb * GT_BK;
                ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sA = smem + stage * GT_STAGE_BYTES;
                uint8_t* sB = sA + 2 * GT_A_BYTES;
                ptx::mbar_arrive_expect_tx(&full_bar[stage], GT_A_BYTES + GT_B_BYTES);
                if (p.a_mn) {
#pragma unroll
                    for (int q = 0; q < GT_BM / 32; ++q) ptx::tma_load_2d(sA + q * GT_PANEL, &tmA, &full_bar[stage], m0 + q * 32, k0);
                } else {
                    ptx::tma_load_2d(sA, &tmA, &full_bar[stage], k0, m0);
                }
                if (p.b_mn) {
#pragma unroll
                    for (int q = 0; q < GT_BN / 32; ++q) ptx::tma_load_2d(sB + q * GT_PANEL, &tmB, &full_bar[stage], n0 + q * 32, k0);
                } else {
                    ptx::tma_load_2d(sB, &tmB, &full_bar[stage], k0, n0);
                    ptx::tma_load_2d(sB + GT_B_BYTES / 2, &tmB, &full_bar[stage], k0, n0 + 128);
                }
                if (++stage == GT_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // The whole warp runs the loop and waits on the barriers; one elected lane issues.  Issued from inside
        // `if (lane == 0)` every tcgen05 instruction was wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall (~35
        // dependent instructions per MMA): with 134-cycle tf32 MMAs the single issuing thread could not keep the tensor
        // pipe fed (ncu: 51 % active).  Warp-uniform operands stay in uniform registers.
        {
            int stage = 0; uint32_t phase = 0;
            int buf = 0; uint32_t bphase0 = 0, bphase1 = 0;
            const uint32_t ring = ptx::smem_u32(smem);
            // K-major: SWIZZLE_128B (BK = 32) or SWIZZLE_64B (BK = 16), 8-row atoms (SBO = 8 rows).  MN-major tf32:
            // SWIZZLE_128B_BASE32B, 32-element x 4-k-row atoms: LBO = panel stride (next 32 m/n), SBO = 512.
            constexpr uint32_t KSBO = 8 * GT_BK * 4, KLT = (GT_BK == 32) ? 2 : 4;
            const uint32_t albo = p.a_mn ? GT_PANEL : 16, blbo = p.b_mn ? GT_PANEL : 16;
            const uint32_t asbo = p.a_mn ? 512 : KSBO, bsbo = p.b_mn ? 512 : KSBO;
            const uint32_t alt = p.a_mn ? 1 : KLT, blt = p.b_mn ? 1 : KLT;
            const uint64_t a_hi64 = ptx::make_smem_desc(0, albo, asbo, alt), b_hi64 = ptx::make_smem_desc(0, blbo, bsbo, blt);
            // K-major: +32 B per k-step inside the swizzle row; MN-major: +1024 B (8 k-rows)   (in 16-byte units)
            const uint32_t astep = p.a_mn ? (1024 >> 4) : (32 >> 4), bstep = p.b_mn ? (1024 >> 4) : (32 >> 4);
            for (int kb = 0; kb < nkb; ++kb) {
                const bool chunk_start = (kb % GT_CHUNK) == 0;
                if (chunk_start) {
                    ptx::mbar_wait(&tempty_bar[buf], (buf ? bphase1 : bphase0) ^ 1);
                    ptx::tc_fence_after();
                }
                ptx::mbar_wait(&conv_bar[stage], phase);
                ptx::tc_fence_after();
                // (the descriptor's address field is 14 bits of the CTA-local offset: in a cluster -- split-K -- the shared
                // address of a CTA with rank > 0 carries its rank in the high bits, which must not leak into the LBO field)
                const uint32_t sA = ((ring + stage * GT_STAGE_BYTES) >> 4) & 0x3FFF;
                const uint32_t sB = sA + ((2 * GT_A_BYTES) >> 4);
                const uint32_t d_tmem = tmem_base + buf * GT_BN;
#pragma unroll
                for (int ks = 0; ks < GT_BK / 8; ++ks) {
                    const uint64_t ahi = a_hi64 | (uint64_t)(sA + ks * astep);
                    const uint64_t alo = a_hi64 | (uint64_t)(sA + (GT_A_BYTES >> 4) + ks * astep);
                    const uint64_t bhi = b_hi64 | (uint64_t)(sB + ks * bstep);
                    const uint64_t blo = b_hi64 | (uint64_t)(sB + (GT_B_BYTES >> 4) + ks * bstep);
                    const uint32_t acc0 = (chunk_start && ks == 0) ? 0u : 1u;
                    ptx::mma_tf32_ss_elect(d_tmem, alo, bhi, p.idesc, acc0);        // small terms first
                    ptx::mma_tf32_ss_elect(d_tmem, ahi, blo, p.idesc, 1u);
                    ptx::mma_tf32_ss_elect(d_tmem, ahi, bhi, p.idesc, 1u);
                }
                ptx::tc_commit_elect(&empty_bar[stage]);
                if (++stage == GT_STAGES) { stage = 0; phase ^= 1; }
                if ((kb % GT_CHUNK) == GT_CHUNK - 1 || kb == nkb - 1) {
                    ptx::tc_commit_elect(&tfull_bar[buf]);
                    if (buf) bphase1 ^= 1; else bphase0 ^= 1;
                    buf ^= 1;
                }
            }
        }
    } else {
        // ===================================================== converter + drain + epilogue warps
        const int ew = warp - 2, quad = warp & 3, half = ew >> 2;
        const int ct = threadIdx.x - 64;                       // 0..255
        float acc[128];
#pragma unroll
        for (int q = 0; q < 128; ++q) acc[q] = 0.f;
        int stage = 0; uint32_t phase = 0;
        int dbuf = 0; uint32_t dphase[2] = {0, 0};
        if (p.cprefetch && p.beta != 0.f && S == 1) {
            // the epilogue reads this thread's 512 bytes of C (4 dependent round trips to HBM otherwise): ask for the lines now
            const int prow = m0 + quad * 32 + lane;
            if (prow < p.M) {
                const float* src = p.C + (int64_t)prow * p.ldc + n0 + half * 128;
                const int ok = min(128, p.N - (n0 + half * 128));
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (q * 32 < ok) asm volatile("prefetch.global.L2 [%0];" :: "l"(src + q * 32));
            }
        }
        auto drain = [&]() {
            ptx::mbar_wait(&tfull_bar[dbuf], dphase[dbuf]);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + dbuf * GT_BN + half * 128 + ((uint32_t)(quad * 32) << 16);
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
            if (lane == 0) ptx::mbar_arrive(&tempty_bar[dbuf]);
            dphase[dbuf] ^= 1;
            dbuf ^= 1;
        };
        for (int kb = 0; kb < nkb; ++kb) {
            ptx::mbar_wait(&full_bar[stage], phase);
            uint8_t* base = smem + stage * GT_STAGE_BYTES;
            // A: raw/hi at [0, A), lo at [A, 2A);  B: raw/hi at [2A, 2A+B), lo at [2A+B, 2A+2B)
#pragma unroll
            for (int it = 0; it < (GT_A_BYTES + GT_B_BYTES) / 16 / 256; ++it) {
                const int idx = ct + it * 256;                                  // float4 index over A then B
                const bool isA = idx < GT_A_BYTES / 16;
                uint8_t* hi = isA ? base + idx * 16 : base + 2 * GT_A_BYTES + (idx - GT_A_BYTES / 16) * 16;
                uint8_t* lo = hi + (isA ? GT_A_BYTES : GT_B_BYTES);
                float4 v = *reinterpret_cast<const float4*>(hi);
                float4 h, l;
                if (p.trunc) {
                    // kind::tf32 ignores the 13 low mantissa bits of its 32-bit operands, so the raw tile already IS the
                    // high part (hi = x with those bits cleared) and only lo = rn_tf32(x - hi) has to be written
                    // (x - hi is exact).  One shared-memory store per element instead of two.
                    h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = rn_tf32(v.x - h.x);
                    h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = rn_tf32(v.y - h.y);
                    h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = rn_tf32(v.z - h.z);
                    h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = rn_tf32(v.w - h.w);
                    *reinterpret_cast<float4*>(lo) = l;
                } else {                        // hi = rn_tf32(x); lo = rn_tf32(x - hi)  (x - hi is exact)
                    h.x = rn_tf32(v.x); l.x = rn_tf32(v.x - h.x);
                    h.y = rn_tf32(v.y); l.y = rn_tf32(v.y - h.y);
                    h.z = rn_tf32(v.z); l.z = rn_tf32(v.z - h.z);
                    h.w = rn_tf32(v.w); l.w = rn_tf32(v.w - h.w);
                    *reinterpret_cast<float4*>(hi) = h;
                    *reinterpret_cast<float4*>(lo) = l;
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&conv_bar[stage]);
            if (++stage == GT_STAGES) { stage = 0; phase ^= 1; }
            // drain the previous chunk once the first k-block of this chunk has been handed to the MMA warp
            if (kb >= GT_CHUNK && (kb % GT_CHUNK) == 0) drain();
        }
        if (nkb > 0) drain();                                                   // last chunk

        const int row = m0 + quad * 32 + lane;
        const int col0 = n0 + half * 128;
        if (S > 1) {
            // partial tile -> own shared memory, transposed [col][row] (pitch 129: conflict-free both ways); every MMA
            // that read the stage buffers has retired (final drain) and no TMA is in flight
            float* T = reinterpret_cast<float*>(smem);
            const int r = quad * 32 + lane;
#pragma unroll
            for (int q = 0; q < 128; ++q) T[(half * 128 + q) * GT_TP + r] = acc[q];
        } else if (row < p.M) {
            float* dst = p.C + (int64_t)row * p.ldc + col0;
            const int cols_ok = min(128, p.N - col0);
            const bool vec = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {
                float* o = dst + cb * 32;
                const int nc = cols_ok - cb * 32;
                if (nc >= 32 && vec) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float4 r;
                        if (p.beta != 0.f) {
                            const float4 c = __ldcg(reinterpret_cast<const float4*>(o + q * 4));
                            r.x = fmaf(p.alpha, acc[cb * 32 + q * 4 + 0], p.beta * c.x); r.y = fmaf(p.alpha, acc[cb * 32 + q * 4 + 1], p.beta * c.y);
                            r.z = fmaf(p.alpha, acc[cb * 32 + q * 4 + 2], p.beta * c.z); r.w = fmaf(p.alpha, acc[cb * 32 + q * 4 + 3], p.beta * c.w);
                        } else {
                            r.x = p.alpha * acc[cb * 32 + q * 4 + 0]; r.y = p.alpha * acc[cb * 32 + q * 4 + 1];
                            r.z = p.alpha * acc[cb * 32 + q * 4 + 2]; r.w = p.alpha * acc[cb * 32 + q * 4 + 3];
                        }
                        *reinterpret_cast<float4*>(o + q * 4) = r;
                    }
                } else if (nc > 0) {
#pragma unroll
                    for (int q = 0; q < 32; ++q)
                        if (q < nc) o[q] = (p.beta != 0.f) ? fmaf(p.alpha, acc[cb * 32 + q], p.beta * __ldcg(o + q)) : p.alpha * acc[cb * 32 + q];
                }
            }
        }
    }

    if (S > 1) {
        __syncwarp();
        ptx::cluster_sync();                                   // all partial tiles written
        if (warp >= 2) {
            // rank zr owns columns [zr*W, (zr+1)*W) of the tile: sum the S partials in rank order, apply alpha/beta.
            // item = (32-column group, 4-row group); lane <-> column, so the global accesses are 128-byte rows and the
            // DSMEM reads are one float4 (4 consecutive rows of one column) per rank, all issued before the sums.
            const int W = GT_BN / S;
            const uint32_t tbase = ptx::smem_u32(smem);
            const int ew = warp - 2;
            const int n_items = W;                              // (W/32) column groups x 32 row groups
            for (int it = ew; it < n_items; it += 8) {
                const int cg = it >> 5, rq = it & 31;
                const int col = zr * W + cg * 32 + lane;
                const uint32_t off = (uint32_t)(col * GT_TP + rq * 4) * 4u;
                float4 v[8];
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c < S) v[c] = ptx::ld_dsmem_f32x4(ptx::mapa(tbase, (uint32_t)c) + off);
                float4 s = v[0];
#pragma unroll
                for (int c = 1; c < 8; ++c)
                    if (c < S) { s.x = __fadd_rn(s.x, v[c].x); s.y = __fadd_rn(s.y, v[c].y); s.z = __fadd_rn(s.z, v[c].z); s.w = __fadd_rn(s.w, v[c].w); }
                const float sv[4] = {s.x, s.y, s.z, s.w};
                const int gc = n0 + col;
                if (gc < p.N) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int gr = m0 + rq * 4 + q;
                        if (gr < p.M) {
                            float* o = p.C + (int64_t)gr * p.ldc + gc;
                            *o = (p.beta != 0.f) ? fmaf(p.alpha, sv[q], p.beta * __ldcg(o)) : p.alpha * sv[q];
                        }
                    }
                }
            }
        }
        __syncwarp();
        ptx::cluster_sync();                                   // nobody leaves while a peer still reads its tile
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn2 get_encode_fn2() {
    static EncodeTiledFn2 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn2>(p);
    }
    return fn;
}

// Operand with logical shape [mn_extent x k_extent].  mn_major: stored [k][mn] (mn contiguous), else [mn][k].
static int make_operand_map(CUtensorMap* tm, const float* ptr, int64_t ld, int mn_extent, int k_extent, bool mn_major, int bk) {
    EncodeTiledFn2 enc = get_encode_fn2();
    if (!enc) return set_err(999, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[2], gstr[1];
    cuuint32_t box[2], estr[2] = {1, 1};
    if (mn_major) { gdim[0] = (cuuint64_t)mn_extent; gdim[1] = (cuuint64_t)k_extent; box[0] = 32; box[1] = (cuuint32_t)bk; }
    else          { gdim[0] = (cuuint64_t)k_extent; gdim[1] = (cuuint64_t)mn_extent; box[0] = (cuuint32_t)bk; box[1] = 128; }
    gstr[0] = (cuuint64_t)ld * 4;
    CUresult cr = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE,
                      mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : (bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B),
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return set_err(1000 + (int)cr, "cuTensorMapEncodeTiled(fp32) failed (%d)", (int)cr);
    return 0;
}

bool gemm_tc_supported(const float* A, int64_t lda, const float* B, int64_t ldb, const float* C, int64_t ldc) {
    return (lda % 4 == 0) && (ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
           ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && C != nullptr && ldc > 0;
}

int gemm_tc_launch(int transa, int transb, int M, int N, int Kd, float alpha, const float* A, int64_t lda, const float* B,
                   int64_t ldb, float beta, float* C, int64_t ldc, int mode, cudaStream_t st) {
    if (M <= 0 || N <= 0) return 0;
    CUtensorMap tmA, tmB;
    const int bk = g_gemm_bk == 16 ? 16 : 32;
    int rc = make_operand_map(&tmA, A, lda, M, Kd, transa != 0, bk);
    if (rc) return rc;
    rc = make_operand_map(&tmB, B, ldb, N, Kd, transb == 0, bk);
    if (rc) return rc;
    GQ_ONCE_PER_DEVICE({
        GQ_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM));
        GQ_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM));
    });
    GemmTcParams p;
    p.M = M; p.N = N; p.Kd = Kd; p.alpha = alpha; p.beta = beta; p.C = C; p.ldc = ldc; p.mode = mode;
    p.a_mn = transa ? 1 : 0; p.b_mn = transb ? 0 : 1;
    p.idesc = ptx::make_idesc(2, p.a_mn, p.b_mn, GT_BM, GT_BN);
    const int gx = cdiv(N, GT_BN), gy = cdiv(M, GT_BM);
    // split-K over a thread-block cluster when the tiles alone cannot fill the GPU (the chain's lower levels, the
    // narrow lazy updates): largest power of two with >= 2 k-blocks per CTA and at most one CTA per SM
    int tiles = gx * gy;
    if (mode & SG_UPPER_TILES) {
        tiles = 0;
        for (int y = 0; y < gy; ++y)
            for (int x = 0; x < gx; ++x) tiles += (x * GT_BN + GT_BN - 1 >= y * GT_BM) ? 1 : 0;
    }
    const int nkb = cdiv(Kd, 32);
    int S = 1;
    if (g_gemm_splitk)
        while (S * 2 <= 8 && tiles * S * 2 <= num_sms() && nkb / (S * 2) >= g_gemm_splitk_minkb) S *= 2;
    p.splits = S;
    p.trunc = g_gemm_trunc;
    p.cprefetch = g_gemm_cprefetch;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(gx, gy, S);
    cfg.blockDim = dim3(GT_THREADS, 1, 1);
    cfg.dynamicSmemBytes = GT_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (S > 1) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = 1; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = S;
        ++na;
    }
    if (g_pdl && !(mode & SG_NO_PDL)) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = at; cfg.numAttrs = na;
    if (bk == 16) GQ_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<16>, tmA, tmB, p));
    else GQ_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<32>, tmA, tmB, p));
    GQ_LAUNCHED();
    return 0;
}

static int backend_from_env() {
    const char* e = getenv("GPTQ_B200_GEMM_BACKEND");
    if (e && (e[0] == 's' || e[0] == '0')) return 0;       // "simt"
    return 1;
}
int g_gemm_backend = backend_from_env();
static int flag_from_env(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? (e[0] != '0') : dflt;
}
int g_gemm_splitk = flag_from_env("GPTQ_B200_SPLITK", 1);
static int int_from_env(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
int g_small_gemm_max = int_from_env("GPTQ_B200_SMALL_GEMM_MAX", 512);   // M, N up to which gemm_fp32 uses sgemm_small.cu
int g_gemm_bk = int_from_env("GPTQ_B200_GEMM_BK", 32);                // k-extent of a pipeline stage (32: 2 stages, 16: 4 stages)
int g_gemm_trunc = int_from_env("GPTQ_B200_GEMM_TRUNC", 1);            // see the converter in gemm_tc_kernel
int g_gemm_splitk_minkb = int_from_env("GPTQ_B200_SPLITK_MINKB", 4);   // fewest k-blocks (of 32) a split may get
int g_pdl = flag_from_env("GPTQ_B200_PDL", 1);
int g_gemm_cprefetch = flag_from_env("GPTQ_B200_GEMM_CPREFETCH", 1);

int gemm_fp32(int transa, int transb, int M, int N, int Kd, float alpha, const float* A, int64_t lda, const float* B,
              int64_t ldb, float beta, float* C, int64_t ldc, int mode, cudaStream_t st) {
    // measured on B200 (tools/perf_probe.py hinv): the small-tile kernel wins up to 384x512 outputs, the tensor-core
    // kernel (with split-K) from 512x512 on
    if (M <= g_small_gemm_max && N <= g_small_gemm_max && (int64_t)M * N < (int64_t)g_small_gemm_max * g_small_gemm_max)
        return sgemm_small_launch(transa, transb, M, N, Kd, alpha, A, lda, B, ldb, beta, C, ldc, mode, st);
    // (A CTA-pair variant of the tensor-core kernel -- cta_group::2, 256 x 256 tiles, each SM staging and splitting only its
    // 128 rows of A and 128 columns of B -- was built and measured in round 2: 130 TFLOP/s at 4096^3 against 163 for this
    // kernel, 87 against 100 at the 4096 x 3584 x 512 lazy update; the cross-CTA hand-off of the split operands costs more
    // than the halved B traffic saves, so it was removed again.  profiles/r02_gemm_pair_negative.md keeps the numbers.)
    if (g_gemm_backend == 1 && gemm_tc_supported(A, lda, B, ldb, C, ldc))
        return gemm_tc_launch(transa, transb, M, N, Kd, alpha, A, lda, B, ldb, beta, C, ldc, mode, st);
    return sgemm_launch(transa, transb, M, N, Kd, alpha, A, lda, B, ldb, beta, C, ldc, mode, st);
}

}  // namespace gq

using namespace gq;

extern "C" int gptq_b200_set_gemm_backend(int backend) {
    GQ_ARG(backend == 0 || backend == 1, 1, "0 = CUDA cores, 1 = tensor cores (3xTF32)");
    int old = g_gemm_backend;
    g_gemm_backend = backend;
    return old < 0 ? 0 : 0;
}

extern "C" int gptq_b200_gemm3x(int transa, int transb, int M, int N, int Kd, float alpha, const float* A, int64_t lda,
                                const float* B, int64_t ldb, float beta, float* C, int64_t ldc, void* stream) {
    GQ_ARG(M >= 0, 3, "M<0"); GQ_ARG(N >= 0, 4, "N<0"); GQ_ARG(Kd >= 0, 5, "K<0");
    GQ_ARG(A, 7, "null"); GQ_ARG(B, 9, "null"); GQ_ARG(C, 12, "null");
    GQ_ARG(lda >= (transa ? M : Kd) && lda % 4 == 0, 8, "lda too small or not a multiple of 4");
    GQ_ARG(ldb >= (transb ? Kd : N) && ldb % 4 == 0, 10, "ldb too small or not a multiple of 4");
    GQ_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0, 7, "A/B must be 16-byte aligned");
    GQ_ARG(ldc >= N, 13, "ldc<N");
    return gemm_tc_launch(transa, transb, M, N, Kd, alpha, A, lda, B, ldb, beta, C, ldc, SG_FULL | SG_NO_PDL, as_stream(stream));
}
