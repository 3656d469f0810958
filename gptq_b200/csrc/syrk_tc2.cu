// (1) Hessian accumulation  H <- beta*H + alpha * X^T X  -- CTA-PAIR variant (tcgen05 cta_group::2).
//
// Reference: GPTQ.add_batch, gptq.py:59-65 (one fp32 cuBLAS sgemm per calibration sample, full square).
//
// Why a second kernel: the single-CTA kernel (syrk_tc.cu, M128 x N256 per SM) is bounded by shared-memory bandwidth --
// per MMA every SM reads a 4 KB A slice and an 8 KB B slice from shared memory while TMA writes 48 KB per four MMAs
// (about 180 B/clk against the SM's 128 B/clk: ncu showed the tensor pipe 76 % active).  With cta_group::2 two SMs of
// a TPC execute ONE MMA of M = 256, N = 256: each CTA stages only ITS 128 rows of A and ITS 128 columns of B
// (32 KB per stage instead of 48) and the tensor cores exchange the B halves themselves, so shared-memory traffic per SM
// falls by a third (about 120 B/clk) and L2 -> SM traffic per flop by the same amount.
//
// Work item = one 256 x 256 tile of H touching the lower triangle (bi >= bj), walked over the tokens by a cluster of two
// CTAs (persistent grid, one cluster per TPC):
//   warp 0   TMA producer in BOTH CTAs: per 64-token stage, A half = 2 panels [64 tokens x 64 channels] of channels
//            bi*256 + rank*128 .., B half = 2 panels of channels bj*256 + rank*128 ..  (SWIZZLE_128B, both operands
//            MN-major because X is [tokens, K] row-major).  Completion bytes of both CTAs are credited to the LEADER's
//            full barrier (cp.async.bulk.tensor ... cta_group::2).  On diagonal tiles the two halves are the same
//            channels: they are loaded once and both descriptors point at them.
//   warp 1   MMA issuer, leader CTA only: 4 x tcgen05.mma.cta_group::2.kind::f16 (M256 N256 K16) per stage into one of
//            two 256-column TMEM buffers; tcgen05.commit ... multicast frees the stage in both CTAs and hands the
//            finished chunk to both epilogues.
//   warps 2-9 epilogue in both CTAs: drain this CTA's 128 rows x 256 columns (tcgen05.ld), add round-to-nearest into 128
//            fp32 registers per thread (the tensor core's own accumulator truncates, so a chain is bounded at `chunk`
//            tokens), release the TMEM buffer to the leader (remote mbarrier arrive), fold into H at the end of the item.
// (An L2 prefetch of the producer's future panels, cp.async.bulk.prefetch.tensor 4 / 8 / 16 stages ahead, was measured 8-12 %
// SLOWER at both headline shapes -- profiles/r02_syrk_tc2_ncu_full.md -- and removed.)
// Tile order: 2-D chunks of sw x sh tiles (about one wave of clusters) so that the clusters running at the same time
// share channel panels in L2.  Few tiles (small K): the token range is split across clusters and a second kernel sums
// the partial tiles in a fixed order (deterministic, no atomics).
#include "common.cuh"
#include "ptx.cuh"
#include <cuda.h>

namespace gq {

constexpr int S2_T = 256;                        // tile edge: MMA M across the pair = MMA N
constexpr int S2_THREADS = 320;                  // TMA warp, MMA warp, 8 epilogue warps
constexpr int S2_RING_BYTES = 6 * 32768;         // operand ring per CTA: BK x STAGES = 384 tokens in flight
constexpr int S2_SMEM = S2_RING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
// BK = tokens per pipeline stage (one full/empty barrier round trip and one tcgen05.commit per BK / 16 MMAs):
//   <64, 6>: 32 KB stages;  <128, 3>: 64 KB stages, half the barrier traffic of the single issuing thread.
template <int BK, int STAGES> struct S2Cfg {
    static constexpr int PANEL = BK * 128;              // bytes: BK tokens x 64 channels x 2 B (TMA box {64, BK})
    static constexpr int OP_BYTES = 2 * PANEL;          // one operand half of one CTA: 128 channels x BK tokens
    static constexpr int STAGE_BYTES = 2 * OP_BYTES;    // A half + B half
    static_assert(STAGES * STAGE_BYTES <= S2_RING_BYTES, "ring too large");
};

struct Syrk2Params {
    int K; int64_t T; int nb; int n_tiles; int splits;
    int64_t tokens_per_split;   // multiple of BK
    int chunk;                  // multiple of BK
    float* H; int64_t ldh; float alpha; float beta;
    float* partial;             // [splits][n_tiles][256][256] when splits > 1
    uint32_t idesc;
    int sw, sh;                 // tile-order chunk: column blocks x row blocks
};

// Valid tiles are (bi, bj) with bi >= bj.  Enumerated in chunks of sw column blocks x sh row blocks; inside a chunk column
// by column.  (CPU restatement and footprint model: oracle/chain_algorithms.py, tests/test_oracle_chain_algorithms.py.)
__device__ __forceinline__ void decode_tile2(int t, int nb, int sw, int sh, int& bi, int& bj) {
    for (int sc = 0; sc < nb; sc += sw) {
        const int w = min(sw, nb - sc);
        for (int r0 = sc; r0 < nb; r0 += sh) {
            const int r1 = min(r0 + sh, nb);
            for (int j = sc; j < sc + w; ++j) {
                const int lo = max(r0, j);
                const int cnt = r1 - lo;
                if (cnt <= 0) continue;
                if (t < cnt) { bj = j; bi = lo + t; return; }
                t -= cnt;
            }
        }
    }
    bi = nb - 1; bj = 0;          // unreachable for t < n_tiles
}

// The same enumeration as decode_tile2, walked forward: a role of a cluster visits tiles cid, cid + ncl, ... in increasing order,
// so the position is carried from item to item instead of being searched from the start (decode_tile2 costs O(chunks x sw)
// iterations per call: 3 us at K = 11008, but ~25 us at K = 36864 -- OPT-66B fc2, 10 440 tiles -- on the producer's critical path).
struct TileWalk2 {
    int nb, sw, sh, sc, r0, j, base;
    __device__ __forceinline__ void init(int nb_, int sw_, int sh_) { nb = nb_; sw = sw_; sh = sh_; sc = 0; r0 = 0; j = 0; base = 0; }
    __device__ __forceinline__ void seek(int t, int& bi, int& bj) {          // t must not decrease between calls
        for (;;) {
            const int w = min(sw, nb - sc);
            const int r1 = min(r0 + sh, nb);
            const int lo = max(r0, j);
            const int cnt = r1 - lo;
            if (cnt > 0) {
                if (t < base + cnt) { bj = j; bi = lo + (t - base); return; }
                base += cnt;
            }
            if (++j >= sc + w) {
                r0 += sh; j = sc;
                if (r0 >= nb) { sc += sw; r0 = sc; j = sc; if (sc >= nb) { bi = nb - 1; bj = 0; return; } }
            }
        }
    }
};

template <int S2_BK, int S2_STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(S2_THREADS, 1)
syrk_tc2_kernel(const __grid_constant__ CUtensorMap tmap, const Syrk2Params p) {
    constexpr int S2_PANEL = S2Cfg<S2_BK, S2_STAGES>::PANEL;
    constexpr int S2_OP_BYTES = S2Cfg<S2_BK, S2_STAGES>::OP_BYTES;
    constexpr int S2_STAGE_BYTES = S2Cfg<S2_BK, S2_STAGES>::STAGE_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S2_STAGES * S2_STAGE_BYTES);   // used in the leader only
    uint64_t* empty_bar = full_bar + S2_STAGES;      // one per CTA (multicast commit)
    uint64_t* tfull_bar = empty_bar + S2_STAGES;     // [2] one per CTA (multicast commit)
    uint64_t* tempty_bar = tfull_bar + 2;            // [2] used in the leader only: 8 epilogue warps x 2 CTAs
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = ptx::warp_idx_uniform(), lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmap);
        for (int s = 0; s < S2_STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { ptx::mbar_init(&tfull_bar[b], 1); ptx::mbar_init(&tempty_bar[b], 16); }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc2(tmem_slot, 512);
        ptx::tmem_relinquish2();
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();           // barriers of BOTH CTAs initialised before any remote arrive / multicast commit
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_items = p.n_tiles * p.splits;

    if (warp == 0) {
        // ===================================================== TMA producer (both CTAs)
        if (lane == 0) {
            const uint32_t full0 = ptx::mapa(ptx::smem_u32(&full_bar[0]), 0);      // the leader's full barriers
            int stage = 0; uint32_t phase = 0;
            TileWalk2 walk; walk.init(p.nb, p.sw, p.sh);
            for (int item = cid; item < n_items; item += ncl) {
                const int tile = item / p.splits, split = item - tile * p.splits;
                int bi, bj; walk.seek(tile, bi, bj);
                const bool diag = bi == bj;
                const int ch_a = bi * S2_T + (int)rank * 128, ch_b = bj * S2_T + (int)rank * 128;
                const int64_t t0 = (int64_t)split * p.tokens_per_split;
                const int64_t t1 = min(t0 + p.tokens_per_split, p.T);
                for (int64_t t = t0; t < t1; t += S2_BK) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (rank == 0)      // bytes of BOTH CTAs land on this barrier
                        ptx::mbar_arrive_expect_tx(&full_bar[stage], diag ? 2 * S2_OP_BYTES : 2 * S2_STAGE_BYTES);
                    uint8_t* sA = smem + stage * S2_STAGE_BYTES;
                    uint8_t* sB = sA + S2_OP_BYTES;
                    const uint32_t bar = full0 + stage * 8;
                    if (!diag) {
                        ptx::tma_load_2d_cg2(sA, &tmap, bar, ch_a, (int)t);
                        ptx::tma_load_2d_cg2(sA + S2_PANEL, &tmap, bar, ch_a + 64, (int)t);
                    }
                    ptx::tma_load_2d_cg2(sB, &tmap, bar, ch_b, (int)t);
                    ptx::tma_load_2d_cg2(sB + S2_PANEL, &tmap, bar, ch_b + 64, (int)t);
                    if (++stage == S2_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (leader CTA only)
        // The whole warp runs the loop and waits on the barriers; one elected lane issues (see ptx::mma_f16_ss2_elect).
        if (rank == 0) {
            // descriptor = constant high part | (shared address >> 4): SWIZZLE_128B, LBO = panel, SBO = 1024, version 1
            constexpr uint64_t DESC_HI = ((uint64_t)((S2_PANEL >> 4) & 0x3FFF) << 16) | ((uint64_t)((1024 >> 4) & 0x3FFF) << 32) |
                                         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
            const uint32_t ring_lo = (ptx::smem_u32(smem) >> 4) & 0x3FFF;      // 14-bit CTA-local offset field
            int stage = 0; uint32_t phase = 0;
            int buf = 0; uint32_t bphase0 = 0, bphase1 = 0;
            TileWalk2 walk; walk.init(p.nb, p.sw, p.sh);
            for (int item = cid; item < n_items; item += ncl) {
                const int tile = item / p.splits, split = item - tile * p.splits;
                int bi, bj; walk.seek(tile, bi, bj);
                const uint32_t a_off_lo = (bi == bj) ? (S2_OP_BYTES >> 4) : 0;       // diagonal: A is this CTA's B half
                const int64_t t0 = (int64_t)split * p.tokens_per_split;
                const int64_t t1 = min(t0 + p.tokens_per_split, p.T);
                for (int64_t c0 = t0; c0 < t1; c0 += p.chunk) {
                    const int nst = (int)((min(c0 + (int64_t)p.chunk, t1) - c0 + S2_BK - 1) / S2_BK);
                    ptx::mbar_wait(&tempty_bar[buf], (buf ? bphase1 : bphase0) ^ 1);   // both epilogues have drained this buffer
                    ptx::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + buf * S2_T;
                    uint32_t acc = 0;
                    for (int s_ = 0; s_ < nst; ++s_) {
                        ptx::mbar_wait(&full_bar[stage], phase);
                        ptx::tc_fence_after();
                        const uint32_t st_lo = ring_lo + stage * (S2_STAGE_BYTES >> 4);
#pragma unroll
                        for (int ks = 0; ks < S2_BK / 16; ++ks) {
                            // 16 tokens = 2 K-atoms of 8 rows x 128 B = 2048 B; 128 channels per CTA = 2 MN atoms (LBO = panel)
                            const uint64_t ad = DESC_HI | (uint64_t)(st_lo + a_off_lo + ks * (2048 >> 4));
                            const uint64_t bd = DESC_HI | (uint64_t)(st_lo + (S2_OP_BYTES >> 4) + ks * (2048 >> 4));
                            ptx::mma_f16_ss2_elect(d_tmem, ad, bd, p.idesc, acc);
                            acc = 1;
                        }
                        ptx::tc_commit2_mc_elect(&empty_bar[stage], 3);   // stage reusable in both CTAs once these MMAs retire
                        if (++stage == S2_STAGES) { stage = 0; phase ^= 1; }
                    }
                    ptx::tc_commit2_mc_elect(&tfull_bar[buf], 3);         // chunk complete -> both epilogues
                    if (buf) bphase1 ^= 1; else bphase0 ^= 1;
                    buf ^= 1;
                }
            }
        }
    } else {
        // ===================================================== epilogue (8 warps, both CTAs)
        // warp -> (TMEM lane quadrant = warp % 4, column half): thread = one row x 128 columns of this CTA's 128 x 256 part.
        const int ew = warp - 2;
        const int quad = warp & 3;
        const int half = ew >> 2;
        int buf = 0; uint32_t bphase[2] = {0, 0};
        const bool hvec = ((p.ldh & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.H) & 15) == 0);
        const uint32_t tempty0 = ptx::mapa(ptx::smem_u32(&tempty_bar[0]), 0);
        TileWalk2 walk; walk.init(p.nb, p.sw, p.sh);
        for (int item = cid; item < n_items; item += ncl) {
            const int tile = item / p.splits, split = item - tile * p.splits;
            int bi, bj; walk.seek(tile, bi, bj);
            // the strictly-upper 128 x 128 quarter of a diagonal tile is not needed (the consumers mirror the lower triangle)
            const bool need = !(bi == bj && rank == 0 && half == 1);
            const int64_t t0 = (int64_t)split * p.tokens_per_split;
            const int64_t t1 = min(t0 + p.tokens_per_split, p.T);
            float acc[128];
#pragma unroll
            for (int q = 0; q < 128; ++q) acc[q] = 0.f;
            for (int64_t c0 = t0; c0 < t1; c0 += p.chunk) {
                ptx::mbar_wait(&tfull_bar[buf], bphase[buf]);
                ptx::tc_fence_after();
                if (need) {
                    const uint32_t taddr = tmem_base + buf * S2_T + half * 128 + ((uint32_t)(quad * 32) << 16);
#pragma unroll
                    for (int cb = 0; cb < 4; ++cb) {
                        float v[32];
                        ptx::tmem_ld_32x32(taddr + cb * 32, v);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int q = 0; q < 32; ++q) acc[cb * 32 + q] = __fadd_rn(acc[cb * 32 + q], v[q]);
                    }
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_cluster(tempty0 + buf * 8);
                bphase[buf] ^= 1;
                buf ^= 1;
            }
            if (!need) continue;
            // ---- fold the item into H (or into its split's partial tile)
            const int lrow = (int)rank * 128 + quad * 32 + lane;
            const int row = bi * S2_T + lrow;
            const int col0 = bj * S2_T + half * 128;
            float* dst; float alpha, beta; bool vec; int rows_ok, cols_ok;
            if (p.splits == 1) {
                dst = p.H + (int64_t)row * p.ldh + col0; alpha = p.alpha; beta = p.beta; vec = hvec;
                rows_ok = row < p.K; cols_ok = min(128, p.K - col0);
            } else {
                dst = p.partial + ((int64_t)split * p.n_tiles + tile) * (S2_T * S2_T) + (int64_t)lrow * S2_T + half * 128;
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

    // The leader's MMAs read the peer's shared memory and both CTAs' barriers are signalled remotely: nobody leaves (or
    // frees TMEM) before both CTAs are done.
    ptx::tc_fence_before();
    ptx::cluster_sync();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc2(tmem_base, 512);
    }
}

// H = beta*H + alpha * sum_s partial[s][tile]   (fixed summation order); the strictly-upper quarter of diagonal tiles is skipped
__global__ void __launch_bounds__(256)
syrk2_reduce_kernel(const float* __restrict__ partial, int n_tiles, int splits, int nb, int K, float* __restrict__ H,
                    int64_t ldh, float alpha, float beta, int sw, int sh) {
    const int tile = blockIdx.x;
    int bi, bj; decode_tile2(tile, nb, sw, sh, bi, bj);
    for (int e = threadIdx.x; e < S2_T * S2_T; e += 256) {
        const int r = e >> 8, c = e & 255;
        const int gr = bi * S2_T + r, gc = bj * S2_T + c;
        if (gr >= K || gc >= K) continue;
        if (bi == bj && r < 128 && c >= 128) continue;
        float s = 0.f;
        for (int sp = 0; sp < splits; ++sp) s += partial[((int64_t)sp * n_tiles + tile) * (S2_T * S2_T) + e];
        float* o = H + (int64_t)gr * ldh + gc;
        *o = (beta != 0.f) ? fmaf(alpha, s, beta * *o) : alpha * s;
    }
}

struct Syrk2Plan { int nb, n_tiles, splits, clusters; int64_t tokens_per_split; };

// Clusters of two 197 KB CTAs that can be resident at once (one per TPC: 74 on a B200), cached per device.
constexpr int S2_PLAN_BK = 128;       // splits / chunks are planned in units valid for both stage depths

static int max_clusters() {
    static std::atomic<int> cache[256] = {};
    const int dev = cur_device();
    const int slot = (dev >= 0 && dev < 256) ? dev : 0;
    int v = cache[slot].load(std::memory_order_relaxed);
    if (v == 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(num_sms() & ~1); cfg.blockDim = dim3(S2_THREADS); cfg.dynamicSmemBytes = S2_SMEM;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, syrk_tc2_kernel<64, 6>, &cfg) != cudaSuccess || n <= 0) { (void)cudaGetLastError(); n = num_sms() / 2; }
        if (n > num_sms() / 2) n = num_sms() / 2;
        v = n;
        cache[slot].store(v, std::memory_order_relaxed);
    }
    return v;
}

static Syrk2Plan make_plan2(int K, int64_t T, int clusters) {
    Syrk2Plan pl;
    pl.clusters = clusters;
    pl.nb = cdiv(K, S2_T);
    pl.n_tiles = pl.nb * (pl.nb + 1) / 2;
    const int64_t kblocks = (T + S2_PLAN_BK - 1) / S2_PLAN_BK;
    int splits = 1;
    if (pl.n_tiles < clusters) {
        splits = clusters / pl.n_tiles;
        const int64_t max_splits = (kblocks + 3) / 4;            // at least 4 k-blocks (512 tokens) per split
        if (splits > max_splits) splits = (int)max_splits;
        if (splits < 1) splits = 1;
    }
    const int64_t kb_per = (kblocks + splits - 1) / splits;
    pl.tokens_per_split = kb_per * S2_PLAN_BK;
    pl.splits = (int)((kblocks + kb_per - 1) / kb_per);
    if (pl.splits < 1) pl.splits = 1;
    return pl;
}

size_t syrk2_workspace_bytes(int K, int64_t tokens) {
    if (K <= 0 || tokens <= 0) return 0;
    // the plan depends on the device's cluster capacity; size for the largest split count any capacity <= SMs/2 can give
    const Syrk2Plan pl = make_plan2(K, tokens, num_sms() / 2);
    if (pl.splits <= 1) return 0;
    return (size_t)pl.splits * pl.n_tiles * S2_T * S2_T * sizeof(float);
}

int syrk2_launch(const CUtensorMap& tmap, int bk, int dtype, int64_t tokens, int K, float* H, int64_t ldh, float alpha, float beta,
                 int chunk, int sw, int sh, void* ws, size_t ws_bytes, cudaStream_t st) {
    GQ_ONCE_PER_DEVICE({
        GQ_CUDA(cudaFuncSetAttribute(syrk_tc2_kernel<64, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_SMEM));
        GQ_CUDA(cudaFuncSetAttribute(syrk_tc2_kernel<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_SMEM));
    });
    const Syrk2Plan pl = make_plan2(K, tokens, max_clusters());
    const size_t need = pl.splits > 1 ? (size_t)pl.splits * pl.n_tiles * S2_T * S2_T * sizeof(float) : 0;
    if (need && (!ws || ws_bytes < need)) return set_err(-12, "syrk_accum: workspace too small (%zu < %zu)", ws_bytes, need);
    Syrk2Params p;
    p.K = K; p.T = tokens; p.nb = pl.nb; p.n_tiles = pl.n_tiles; p.splits = pl.splits;
    p.tokens_per_split = pl.tokens_per_split;
    p.chunk = ((chunk + S2_PLAN_BK - 1) / S2_PLAN_BK) * S2_PLAN_BK;
    p.H = H; p.ldh = ldh; p.alpha = alpha; p.beta = beta;
    p.partial = static_cast<float*>(ws);
    p.idesc = ptx::make_idesc(dtype == GPTQ_B200_F16 ? 0 : 1, 1, 1, S2_T, S2_T);
    // 72 tiles ~ one wave of 74 clusters.  ncu at K = 11008 (721 MB slab, profiles/r02_syrk_tile_order_sweep.md): 12 x 6 reads
    // 5.0 GB from DRAM, 8 x 9 5.6 GB, 6 x 12 5.2 GB, 4 x 18 5.9 GB, 43 x 2 8.0 GB; durations within 2 % of each other
    p.sw = sw ? sw : 12; p.sh = sh ? sh : 6;
    const int n_items = pl.n_tiles * pl.splits;
    const int grid = 2 * (n_items < pl.clusters ? n_items : pl.clusters);
    if (bk == 128) syrk_tc2_kernel<128, 3><<<grid, S2_THREADS, S2_SMEM, st>>>(tmap, p);
    else syrk_tc2_kernel<64, 6><<<grid, S2_THREADS, S2_SMEM, st>>>(tmap, p);
    GQ_LAUNCHED();
    if (pl.splits > 1) {
        syrk2_reduce_kernel<<<pl.n_tiles, 256, 0, st>>>(p.partial, pl.n_tiles, pl.splits, pl.nb, K, H, ldh, alpha, beta, p.sw, p.sh);
        GQ_LAUNCHED();
    }
    return 0;
}

}  // namespace gq
