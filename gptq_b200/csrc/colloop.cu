// (3) The fasterquant column loop (reference gptq.py:191-276, standard branch).
//
// Per block of <=128 columns:
//   blockquant_kernel -- rows are independent, so a CTA owns a slab of rows (one CTA per SM; the slab is
//     sized so that ceil(R / slab) <= #SMs where possible).  Four lanes share a row and keep its
//     128-column W1 slice in registers (32 per lane: 8 chunks of 4 consecutive columns); Hinv1 (128x128)
//     and the per-(row,column) scale / zero / 1/scale live in shared memory.  For every column the owner
//     lane broadcasts w with a shuffle, the four lanes quantize it (quant.py:6-10, same operation order
//     and roundings as torch), form err = (w-q)/d and apply the rank-1 update to the columns they own --
//     product and subtraction rounded separately, exactly like
//     W1[:, i:] -= err1.unsqueeze(1).matmul(Hinv1[i, i:].unsqueeze(0)) (gptq.py:270).
//     Finished columns' registers are recycled to hold err, so Q / codes / Err1 leave with 16-byte
//     stores at the end of the block.
//     The two IEEE divisions of the per-column dependency chain (w/scale, (w-q)/d) use reciprocals that
//     are hoisted out of the chain (1/scale per (row,column) at block start, 1/d per column) and the
//     exactly-rounded quotient is recovered with FMAs: q0 = a*y; r = a - q0*b; q1 = q0 + r*y, applied
//     twice (the second pass is the identity on a correctly rounded quotient).  gptq_b200_selftest_div
//     checks bit-equality with div.rn.f32 on the GPU (tests/test_gpu_parity.py).
//   lazy update -- W[:, i2:] -= Err1 @ U[i1:i2, i2:] (gptq.py:276) as an fp32-accurate GEMM, two-level.
// Dynamic groups (gptq.py:253-255) are fitted by find_params on the CURRENT global W when the
// block containing the group start begins, which is what the reference's stale read amounts to.
//
// Compiled with --fmad=false: no a*b+c in this file is contracted; every FMA below is an explicit fmaf.
#include "sgemm.cuh"
#include "ptx.cuh"
#include <stdlib.h>

namespace gq {

constexpr int BQ_B = 128;            // max block width handled by the kernel
constexpr int BQ_SP = BQ_B + 1;      // pitch of the per-(row,column) tiles
constexpr int BQ_SUPER = 4;          // blocks per super-block of the two-level lazy update
constexpr int BQ_MAX_ROWS = 96;      // rows per CTA: 64 KB + 3 * rows * 516 B of shared memory <= 227 KB

struct BQArgs {
    float* W; int R; int K; int64_t ldw; const float* U; int64_t ldu; int i1; int count;
    const int32_t* gidx; const float* scale; const float* zero; int ng; float maxq; int trits;
    float* Q; int64_t ldq; uint8_t* codes; int64_t ldc; float* Err; int64_t lde; float* row_loss; int rows_per_cta;
};

// Exactly rounded a / b given y = RN(1 / b): two residual-correction passes.
__device__ __forceinline__ float div_by_recip(float a, float b, float y) {
    float q = __fmul_rn(a, y);
    float r = fmaf(-q, b, a);
    q = fmaf(r, y, q);
    r = fmaf(-q, b, a);
    q = fmaf(r, y, q);
    return q;
}

// LPR = lanes per row: 4 (32 columns per lane) or 8 (16 columns per lane).  The 8-lane form doubles the warps per
// scheduler, which is what a latency-bound recurrence needs (58 -> 33 us per block at 32 rows per CTA, still 2-3 % ahead
// at 80 rows); the 4-lane form remains for slabs of 81-96 rows (register file: 8 lanes x 96 rows would not fit).
template <int LPR>
__global__ void __launch_bounds__(LPR == 8 ? 640 : BQ_MAX_ROWS * 4, 1)
blockquant_kernel(BQArgs a) {
    constexpr int NW = BQ_B / LPR;           // W1 columns held per lane
    constexpr int NCH = NW / 4;              // chunks of 4 consecutive columns per lane
    constexpr int CW = LPR * 4;              // columns covered by one chunk across the lanes of a row
    extern __shared__ __align__(16) float sm[];
    const int RPC = a.rows_per_cta;
    float* Hs = sm;                          // [128][128]
    float* Ss = Hs + BQ_B * BQ_B;            // [RPC][129]  scale
    float* Zs = Ss + RPC * BQ_SP;            // [RPC][129]  zero
    float* Ys = Zs + RPC * BQ_SP;            // [RPC][129]  RN(1/scale)
    __shared__ float s_rd[BQ_B];             // RN(1 / Hinv1[i][i])

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x;
    const int l4 = lane & (LPR - 1), rl = tid / LPR;
    const int row = blockIdx.x * RPC + rl;
    const bool rvalid = row < a.R;
    const int count = a.count, i1 = a.i1;
    ptx::pdl_launch_dependents();
    // Programmatic dependent launch: this kernel may start while the preceding lazy-update GEMM is still running.  Hinv1,
    // gidx and the quantizer tables do not change during the column loop (mode 0; mode 1 is launched fully serialized),
    // so they are staged BEFORE the wait and overlap the GEMM's tail; only W is read after it (through L2: an
    // early-launched kernel does not start with a clean L1).

    // ---- stage Hinv1 = U[i1:i1+count, i1:i1+count] (zero padded); independent loads, unrolled for memory-level parallelism
    const bool uvec = ((a.ldu & 3) == 0) && ((i1 & 3) == 0) && ((((uintptr_t)a.U) & 15) == 0);
    if (uvec && count == BQ_B) {
#pragma unroll 8
        for (int idx = tid; idx < BQ_B * (BQ_B / 4); idx += nthreads) {
            const int r = idx >> 5, c4 = (idx & 31) * 4;
            *reinterpret_cast<float4*>(&Hs[r * BQ_B + c4]) =
                __ldcg(reinterpret_cast<const float4*>(a.U + (int64_t)(i1 + r) * a.ldu + i1 + c4));
        }
    } else {
        for (int idx = tid; idx < BQ_B * (BQ_B / 4); idx += nthreads) {
            int r = idx >> 5, c4 = (idx & 31) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < count) {
                const float* p = a.U + (int64_t)(i1 + r) * a.ldu + i1 + c4;
                if (c4 + 0 < count) v.x = __ldcg(p + 0);
                if (c4 + 1 < count) v.y = __ldcg(p + 1);
                if (c4 + 2 < count) v.z = __ldcg(p + 2);
                if (c4 + 3 < count) v.w = __ldcg(p + 3);
            }
            *reinterpret_cast<float4*>(&Hs[r * BQ_B + c4]) = v;
        }
    }
    __shared__ int s_g[BQ_B];
    for (int i = tid; i < BQ_B; i += nthreads) {
        float d = (i < count) ? __ldcg(&a.U[(int64_t)(i1 + i) * a.ldu + i1 + i]) : 1.f;
        s_rd[i] = __frcp_rn(d);
        s_g[i] = (a.gidx && i < count) ? a.gidx[i1 + i] : 0;
    }
    __syncthreads();
    // ---- stage the quantizer of every (row, column) of the block (group index from smem: one level of loads).
    // Loads are issued in batches of 16 before anything consumes them (the reciprocal's slow path is a branch the
    // compiler will not hoist loads across: unbatched, every entry paid a full L2 round trip -- 35 % of the kernel).
    // scale / zero / gidx do not change while PDL-launched blocks are in flight (mode 0), so L1-cached loads are safe.
    const int row0 = blockIdx.x * RPC;
    for (int base = 0; base < RPC * BQ_B; base += nthreads * 16) {
        float sv[16], zv[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int idx = base + u * nthreads + tid;
            sv[u] = 1.f; zv[u] = 0.f;
            if (idx < RPC * BQ_B) {
                const int r = idx >> 7, c = idx & 127;
                int gr = row0 + r;
                if (gr >= a.R) gr = a.R - 1;                     // clamp: harmless duplicate of the last row
                const int64_t o = (int64_t)gr * a.ng + s_g[c];
                sv[u] = a.scale[o]; zv[u] = a.zero[o];
            }
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int idx = base + u * nthreads + tid;
            if (idx < RPC * BQ_B) {
                const int r = idx >> 7, c = idx & 127;
                Ss[r * BQ_SP + c] = sv[u];
                Zs[r * BQ_SP + c] = zv[u];
                Ys[r * BQ_SP + c] = __frcp_rn(sv[u]);
            }
        }
    }
    ptx::pdl_wait();
    // ---- W1 slice -> registers: w[kk*4+v] = W[row][i1 + kk*CW + l4*4 + v]
    float w[NW];
    uint32_t cpk[NCH];
    const bool wvec = ((a.ldw & 3) == 0) && ((i1 & 3) == 0) && ((((uintptr_t)a.W) & 15) == 0);
#pragma unroll
    for (int kk = 0; kk < NCH; ++kk) {
        int c = kk * CW + l4 * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rvalid) {
            const float* p = a.W + (int64_t)row * a.ldw + i1 + c;
            if (wvec && c + 3 < count) v = __ldcg(reinterpret_cast<const float4*>(p));   // L2: written by the GEMM just before (PDL)
            else {
                if (c + 0 < count) v.x = __ldcg(p + 0);
                if (c + 1 < count) v.y = __ldcg(p + 1);
                if (c + 2 < count) v.z = __ldcg(p + 2);
                if (c + 3 < count) v.w = __ldcg(p + 3);
            }
        }
        w[kk * 4 + 0] = v.x; w[kk * 4 + 1] = v.y; w[kk * 4 + 2] = v.z; w[kk * 4 + 3] = v.w;
        cpk[kk] = 0u;
    }
    __syncthreads();

    float loss = 0.f;
    const float* srow = Ss + rl * BQ_SP;
    const float* zrow = Zs + rl * BQ_SP;
    const float* yrow = Ys + rl * BQ_SP;
    const int lbase = lane & ~(LPR - 1);

#pragma unroll
    for (int kk = 0; kk < NCH; ++kk) {
#pragma unroll 1
        for (int l = 0; l < LPR; ++l) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const int i = kk * CW + l * 4 + v;
                if (i < count) {                                         // block-uniform
                    const float wi = __shfl_sync(0xffffffffu, w[kk * 4 + v], lbase | l);
                    const float d = Hs[i * BQ_B + i], rd = s_rd[i];
                    const float sc = srow[i], zr = zrow[i], ys = yrow[i];
                    float code, q;
                    if (a.trits) q = quant_trits(wi, sc, zr, &code);
                    else {                                               // quant.py:9-10
                        float t = __fadd_rn(rintf(div_by_recip(wi, sc, ys)), zr);
                        code = fminf(fmaxf(t, 0.f), a.maxq);
                        q = dequant(code, sc, zr);
                    }
                    const float diff = __fsub_rn(wi, q);
                    const float err = div_by_recip(diff, d, rd);           // gptq.py:269
                    loss = fmaf(err, err, loss);                           // gptq.py:267: (w-q)^2/d^2 (reported scalar)
                    // rank-1 update of the columns this lane still owns (gptq.py:270)
#pragma unroll
                    for (int k2 = kk; k2 < NCH; ++k2) {
                        const float4 h = *reinterpret_cast<const float4*>(&Hs[i * BQ_B + k2 * CW + l4 * 4]);
                        const float hv[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
                        for (int v2 = 0; v2 < 4; ++v2) {
                            const bool upd = (k2 > kk) || (l4 > l) || (l4 == l && v2 > v);
                            if (upd) w[k2 * 4 + v2] = __fsub_rn(w[k2 * 4 + v2], __fmul_rn(err, hv[v2]));
                        }
                    }
                    if (l4 == l) {                                       // owner keeps the results
                        w[kk * 4 + v] = err;
                        cpk[kk] |= ((uint32_t)code & 0xFFu) << (8 * v);
                    }
                }
            }
        }
    }

    // ---- outputs: Q (recomputed from the codes with the same two roundings), codes, Err1, loss
    if (rvalid) {
        const bool qvec = ((a.ldq & 3) == 0) && ((i1 & 3) == 0) && ((((uintptr_t)a.Q) & 15) == 0);
        const bool cvec = ((a.ldc & 3) == 0) && ((i1 & 3) == 0) && ((((uintptr_t)a.codes) & 3) == 0);
        const bool evec = ((a.lde & 3) == 0) && ((((uintptr_t)a.Err) & 15) == 0);
#pragma unroll
        for (int kk = 0; kk < NCH; ++kk) {
            const int c = kk * CW + l4 * 4;
            if (c >= count) continue;
            float qv[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const float code = (float)((cpk[kk] >> (8 * v)) & 0xFFu);
                const float sc = srow[c + v], zr = zrow[c + v];
                if (a.trits) qv[v] = __fadd_rn(__fmul_rn(code > 1.5f ? 1.f : 0.f, sc), __fmul_rn(code < 0.5f ? 1.f : 0.f, zr));
                else qv[v] = dequant(code, sc, zr);
            }
            float* qp = a.Q + (int64_t)row * a.ldq + i1 + c;
            uint8_t* cp = a.codes + (int64_t)row * a.ldc + i1 + c;
            float* ep = a.Err + (int64_t)row * a.lde + c;
            if (c + 3 < count) {
                if (qvec) *reinterpret_cast<float4*>(qp) = make_float4(qv[0], qv[1], qv[2], qv[3]);
                else { qp[0] = qv[0]; qp[1] = qv[1]; qp[2] = qv[2]; qp[3] = qv[3]; }
                if (cvec) *reinterpret_cast<uint32_t*>(cp) = cpk[kk];
                else { cp[0] = cpk[kk] & 0xFF; cp[1] = (cpk[kk] >> 8) & 0xFF; cp[2] = (cpk[kk] >> 16) & 0xFF; cp[3] = cpk[kk] >> 24; }
                if (evec) *reinterpret_cast<float4*>(ep) = make_float4(w[kk * 4], w[kk * 4 + 1], w[kk * 4 + 2], w[kk * 4 + 3]);
                else { ep[0] = w[kk * 4]; ep[1] = w[kk * 4 + 1]; ep[2] = w[kk * 4 + 2]; ep[3] = w[kk * 4 + 3]; }
            } else {
#pragma unroll
                for (int v = 0; v < 4; ++v)
                    if (c + v < count) { qp[v] = qv[v]; cp[v] = (cpk[kk] >> (8 * v)) & 0xFF; ep[v] = w[kk * 4 + v]; }
            }
        }
    }
    // Losses1 / 2 (gptq.py:274), accumulated per row over the blocks: one owner lane per row, launches are serialized,
    // so the sum is deterministic (the caller reduces rows, or row ranges, in double).
    if (rvalid && l4 == 0) a.row_loss[row] = __ldcg(&a.row_loss[row]) + 0.5f * loss;
}

static int bq_lpr8_from_env() {
    const char* e = getenv("GPTQ_B200_BQ_LPR8_MAX_ROWS");
    return e ? atoi(e) : 80;
}
static int g_bq_lpr8_max_rows = bq_lpr8_from_env();      // rows per CTA up to which the 8-lanes-per-row kernel is used (<= 80:
                                                          // 640 threads; measured 2-3 % faster than 4 lanes even at 80 rows)

// Counts (a,b) pairs for which div_by_recip(a, b, RN(1/b)) differs from div.rn.f32(a, b).
__global__ void selftest_div_kernel(uint64_t n, uint32_t seed, int mode, unsigned long long* mismatches) {
    uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    for (; idx < n; idx += stride) {
        uint64_t x = idx * 0x9E3779B97F4A7C15ull + seed;
        x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
        uint32_t ma = (uint32_t)x & 0x7FFFFF, mb = (uint32_t)(x >> 23) & 0x7FFFFF;
        int ea = (int)((x >> 46) & 31) - 16, eb = (int)((x >> 51) & 31) - 16;     // exponents in [-16, 15]
        if (mode == 1) mb = 0x7FFFFF - (mb & 0xFF);                                // significand of b near all-ones
        if (mode == 2) mb = mb & 0xFF;                                             // significand of b near 1.0
        float a = __uint_as_float(((uint32_t)(ea + 127) << 23) | ma);
        float b = __uint_as_float(((uint32_t)(eb + 127) << 23) | mb);
        if (x >> 63) a = -a;
        if ((x >> 62) & 1) b = -b;
        float y = __frcp_rn(b);
        if (__float_as_uint(div_by_recip(a, b, y)) != __float_as_uint(__fdiv_rn(a, b))) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

}  // namespace gq

using namespace gq;

extern "C" size_t gptq_b200_colloop_workspace_bytes(int R, int K, int blocksize) {
    (void)K; (void)blocksize;
    return (size_t)(R > 0 ? R : 0) * BQ_SUPER * BQ_B * sizeof(float);
}

extern "C" int gptq_b200_colloop(float* W, int R, int K, int64_t ldw, const float* U, int64_t ldu, int blocksize,
                                 int groupsize, int mode, const int32_t* gidx, float* scale, float* zero, int ngroups,
                                 int maxq, int sym, int mse, float norm, int grid, float maxshrink, float* Q, int64_t ldq,
                                 uint8_t* codes, int64_t ldc, float* row_loss, void* ws, size_t ws_bytes, void* stream) {
    GQ_ARG(W, 1, "null"); GQ_ARG(R >= 0, 2, "R<0"); GQ_ARG(K > 0, 3, "K<=0"); GQ_ARG(ldw >= K, 4, "ldw<K");
    GQ_ARG(U, 5, "null"); GQ_ARG(ldu >= K, 6, "ldu<K");
    GQ_ARG(blocksize > 0 && (blocksize <= BQ_B || blocksize % BQ_B == 0), 7, "blocksize must be in [1,128] or a multiple of 128");
    GQ_ARG(mode == 0 || mode == 1, 9, "mode must be 0 or 1");
    GQ_ARG(mode == 0 || groupsize > 0, 8, "dynamic groups need groupsize > 0");
    GQ_ARG(scale && zero, 11, "null"); GQ_ARG(ngroups >= 1, 13, "ngroups<1");
    GQ_ARG(ngroups == 1 || gidx, 10, "gidx required when ngroups > 1");
    GQ_ARG(maxq != 0, 14, "maxq == 0 (Quantizer not configured)");
    GQ_ARG(Q, 20, "null"); GQ_ARG(ldq >= K, 21, "ldq<K"); GQ_ARG(codes, 22, "null"); GQ_ARG(ldc >= K, 23, "ldc<K");
    GQ_ARG(row_loss, 24, "null"); GQ_ARG(ws, 25, "null");
    GQ_ARG(ws_bytes >= gptq_b200_colloop_workspace_bytes(R, K, blocksize), 26, "workspace too small");
    if (R == 0) return 0;
    cudaStream_t st = as_stream(stream);
    float* Err = static_cast<float*>(ws);
    // one CTA per SM: slab of rows sized so that the grid fits in one wave where possible
    int rpc = ((cdiv(R, num_sms()) + 7) / 8) * 8;
    if (rpc < 8) rpc = 8;
    if (rpc > BQ_MAX_ROWS) rpc = BQ_MAX_ROWS;
    const size_t smem = (size_t)(BQ_B * BQ_B + 3 * rpc * BQ_SP) * sizeof(float);
    GQ_ONCE_PER_DEVICE({
        const size_t smem_max = (size_t)(BQ_B * BQ_B + 3 * BQ_MAX_ROWS * BQ_SP) * sizeof(float);
        GQ_CUDA(cudaFuncSetAttribute(blockquant_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        GQ_CUDA(cudaFuncSetAttribute(blockquant_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    });
    const int lpr = (rpc <= g_bq_lpr8_max_rows) ? 8 : 4;
    GQ_CUDA(cudaMemsetAsync(row_loss, 0, (size_t)R * sizeof(float), st));
    // Two-level lazy update (same arithmetic as gptq.py:276, regrouped): after every block only the rest of the current
    // super-block is brought up to date (narrow GEMM, depth `count`); the columns beyond it receive ONE deep GEMM per
    // super-block.  Dynamic groups are fitted by find_params on the GLOBAL W at the start of the block in which they
    // begin (gptq.py:253-255), and the reference has by then applied every earlier block's update to ALL later columns
    // (gptq.py:276).  Groups start at multiples of `groupsize`, so a group stays inside one super-block -- where the
    // near updates keep W current -- iff the super-block width is a multiple of the group size; otherwise (gs = 96,
    // 100, 1024, blocksize 100 with gs 64, ...) a group would straddle the super-block end and read columns still
    // waiting for the far GEMM, so those cases keep the one-level form.
    // A reference block wider than 128 columns (gptq.py:191 accepts any blocksize) is walked in 128-column kernel blocks:
    // in-block and lazy updates are the same arithmetic, W[:, j] -= err_i * U[i, j], so only the fp32 summation order
    // differs -- except that dynamic groups are fitted on W as of the start of the REFERENCE block (gptq.py:253-255), which
    // is why find_params below is keyed on `blocksize`, not on the kernel block.
    const int kb = blocksize < BQ_B ? blocksize : BQ_B;
    const int64_t sbw = (int64_t)BQ_SUPER * kb;
    const bool two_level = (mode == 0) || (sbw % groupsize == 0 && sbw % blocksize == 0);
    const int sb_blocks = two_level ? BQ_SUPER : 1;
    const int64_t lde = (int64_t)BQ_SUPER * BQ_B;
    int sb_start = 0;
    for (int i1 = 0, bi = 0; i1 < K; i1 += kb, ++bi) {
        const int i2 = (i1 + kb < K) ? i1 + kb : K;
        const int count = i2 - i1;
        if (bi % sb_blocks == 0) sb_start = i1;
        int sb_end = sb_start + sb_blocks * kb; if (sb_end > K) sb_end = K;
        if (mode == 1 && i1 % blocksize == 0) {
            const int r2 = (i1 + blocksize < K) ? i1 + blocksize : K;          // end of the reference block
            const int first = ((i1 + groupsize - 1) / groupsize) * groupsize;
            if (first < r2) {
                const int last = ((r2 - 1) / groupsize) * groupsize;
                int end = last + groupsize; if (end > K) end = K;
                int rc = gptq_b200_find_params(W, R, K, ldw, first, end - first, groupsize, maxq, sym, mse, norm, grid,
                                               maxshrink, scale, zero, ngroups, first / groupsize, stream);
                if (rc) return rc;
            }
        }
        float* Eblk = Err + (i1 - sb_start);
        BQArgs a{W, R, K, ldw, U, ldu, i1, count, gidx, scale, zero, ngroups, (float)maxq, maxq < 0 ? 1 : 0,
                 Q, ldq, codes, ldc, Eblk, lde, row_loss, rpc};
        if (lpr == 8) GQ_CUDA(launch_pdl(/*allow=*/i1 > 0 && mode == 0, blockquant_kernel<8>, dim3(cdiv(R, rpc)), dim3(rpc * 8), smem, st, a));
        else GQ_CUDA(launch_pdl(/*allow=*/i1 > 0 && mode == 0, blockquant_kernel<4>, dim3(cdiv(R, rpc)), dim3(rpc * 4), smem, st, a));
        GQ_LAUNCHED();
        if (i2 < sb_end) {            // near: the rest of this super-block
            int rc = gemm_fp32(0, 0, R, sb_end - i2, count, -1.f, Eblk, lde, U + (int64_t)i1 * ldu + i2, ldu, 1.f,
                               W + i2, ldw, SG_FULL, st);
            if (rc) return rc;
        } else if (i2 < K) {          // far: everything beyond the super-block, one deep GEMM
            int rc = gemm_fp32(0, 0, R, K - sb_end, sb_end - sb_start, -1.f, Err, lde, U + (int64_t)sb_start * ldu + sb_end,
                               ldu, 1.f, W + sb_end, ldw, SG_FULL, st);
            if (rc) return rc;
        }
    }
    return 0;
}

extern "C" int gptq_b200_selftest_div(uint64_t n, uint32_t seed, int mode, unsigned long long* mismatches, void* stream) {
    GQ_ARG(mismatches, 4, "null");
    GQ_ARG(mode >= 0 && mode <= 2, 3, "mode must be 0 (random), 1 (b near 2-ulp) or 2 (b near 1)");
    cudaStream_t st = as_stream(stream);
    GQ_CUDA(cudaMemsetAsync(mismatches, 0, sizeof(unsigned long long), st));
    selftest_div_kernel<<<num_sms() * 8, 256, 0, st>>>(n, seed, mode, mismatches);
    GQ_LAUNCHED();
    return 0;
}
