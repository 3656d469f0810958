// (3) The fasterquant column loop (reference gptq.py:191-276, standard branch).
//
// Per block of <=128 columns:
//   blockquant_kernel -- rows are independent, so a CTA owns 16 rows; 8 lanes share a row and
//     keep its 128-column W1 slice in registers (16 per lane, 4 chunks of 4 consecutive columns),
//     Hinv1 (128x128) and the per-(row,column) scale/zero live in shared memory.  For every column
//     the owner lane broadcasts w with a shuffle, all 8 lanes quantize it (quant.py:6-10, same
//     operation order and IEEE roundings as torch), form err = (w-q)/d and apply the rank-1 update
//     to the columns they own -- product and subtraction rounded separately, exactly like
//     W1[:, i:] -= err1.unsqueeze(1).matmul(Hinv1[i, i:].unsqueeze(0)) (gptq.py:270).
//     Finished columns' registers are recycled to hold err so Q / codes / Err1 leave with
//     16-byte coalesced stores at the end of the block.
//   lazy update -- W[:, i2:] -= Err1 @ U[i1:i2, i2:] (gptq.py:276) as an fp32 GEMM.
// Dynamic groups (gptq.py:253-255) are fitted by find_params on the CURRENT global W when the
// block containing the group start begins, which is what the reference's stale read amounts to.
//
// Compiled with --fmad=false: nothing in this file may be contracted into an FMA.
#include "sgemm.cuh"

namespace gq {

constexpr int BQ_ROWS = 16;
constexpr int BQ_THREADS = 128;
constexpr int BQ_B = 128;            // max block width handled by the kernel
constexpr int BQ_SP = BQ_B + 1;      // pitch of the scale/zero tiles
constexpr int BQ_SUPER = 4;          // blocks per super-block of the two-level lazy update

struct BQArgs {
    float* W; int R; int K; int64_t ldw; const float* U; int64_t ldu; int i1; int count;
    const int32_t* gidx; const float* scale; const float* zero; int ng; float maxq; int trits;
    float* Q; int64_t ldq; uint8_t* codes; int64_t ldc; float* Err; int64_t lde; double* loss;
};

__global__ void __launch_bounds__(BQ_THREADS)
blockquant_kernel(BQArgs a) {
    extern __shared__ __align__(16) float sm[];
    float* Hs = sm;                          // [128][128]
    float* Ss = Hs + BQ_B * BQ_B;            // [16][129]
    float* Zs = Ss + BQ_ROWS * BQ_SP;        // [16][129]
    __shared__ double s_loss[BQ_THREADS / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l8 = lane & 7, rl = warp * 4 + (lane >> 3);
    const int row = blockIdx.x * BQ_ROWS + rl;
    const bool rvalid = row < a.R;
    const int count = a.count, i1 = a.i1;

    // ---- stage Hinv1 = U[i1:i1+count, i1:i1+count] (zero padded)
    const bool uvec = ((a.ldu & 3) == 0) && ((i1 & 3) == 0) && ((((uintptr_t)a.U) & 15) == 0);
    for (int idx = tid; idx < BQ_B * (BQ_B / 4); idx += BQ_THREADS) {
        int r = idx >> 5, c4 = (idx & 31) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < count) {
            const float* p = a.U + (int64_t)(i1 + r) * a.ldu + i1 + c4;
            if (uvec && c4 + 3 < count) v = *reinterpret_cast<const float4*>(p);
            else {
                if (c4 + 0 < count) v.x = p[0];
                if (c4 + 1 < count) v.y = p[1];
                if (c4 + 2 < count) v.z = p[2];
                if (c4 + 3 < count) v.w = p[3];
            }
        }
        *reinterpret_cast<float4*>(&Hs[r * BQ_B + c4]) = v;
    }
    // ---- stage the quantizer of every (row, column) of the block
    for (int idx = tid; idx < BQ_ROWS * BQ_B; idx += BQ_THREADS) {
        int r = idx >> 7, c = idx & 127;
        int gr = blockIdx.x * BQ_ROWS + r;
        float s = 1.f, z = 0.f;
        if (gr < a.R && c < count) {
            int g = a.gidx ? a.gidx[i1 + c] : 0;
            s = a.scale[(int64_t)gr * a.ng + g];
            z = a.zero[(int64_t)gr * a.ng + g];
        }
        Ss[r * BQ_SP + c] = s;
        Zs[r * BQ_SP + c] = z;
    }
    // ---- W1 slice -> registers: w[kk*4+v] = W[row][i1 + kk*32 + l8*4 + v]
    float w[16], qv[16];
    uint32_t cpk[4] = {0u, 0u, 0u, 0u};
    const bool wvec = ((a.ldw & 3) == 0) && ((i1 & 3) == 0) && ((((uintptr_t)a.W) & 15) == 0);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        int c = kk * 32 + l8 * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rvalid) {
            const float* p = a.W + (int64_t)row * a.ldw + i1 + c;
            if (wvec && c + 3 < count) v = *reinterpret_cast<const float4*>(p);
            else {
                if (c + 0 < count) v.x = p[0];
                if (c + 1 < count) v.y = p[1];
                if (c + 2 < count) v.z = p[2];
                if (c + 3 < count) v.w = p[3];
            }
        }
        w[kk * 4 + 0] = v.x; w[kk * 4 + 1] = v.y; w[kk * 4 + 2] = v.z; w[kk * 4 + 3] = v.w;
        qv[kk * 4 + 0] = qv[kk * 4 + 1] = qv[kk * 4 + 2] = qv[kk * 4 + 3] = 0.f;
    }
    __syncthreads();

    double loss = 0.0;
    const float* srow = Ss + rl * BQ_SP;
    const float* zrow = Zs + rl * BQ_SP;
    const int lbase = lane & ~7;

#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll 1
        for (int l = 0; l < 8; ++l) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const int i = kk * 32 + l * 4 + v;
                if (i < count) {                                         // block-uniform
                    const float wi = __shfl_sync(0xffffffffu, w[kk * 4 + v], lbase | l);
                    const float d = Hs[i * BQ_B + i];
                    const float sc = srow[i], zr = zrow[i];
                    float code, q;
                    if (a.trits) q = quant_trits(wi, sc, zr, &code);
                    else { code = quant_code(wi, sc, zr, a.maxq); q = dequant(code, sc, zr); }
                    const float diff = __fsub_rn(wi, q);
                    const float err = __fdiv_rn(diff, d);                  // gptq.py:269
                    loss += (double)__fmul_rn(err, err);      // gptq.py:267: (w-q)^2/d^2 == err^2 up to one rounding (reported scalar only)
                    // rank-1 update of the columns this lane still owns (gptq.py:270)
#pragma unroll
                    for (int k2 = kk; k2 < 4; ++k2) {
                        const float4 h = *reinterpret_cast<const float4*>(&Hs[i * BQ_B + k2 * 32 + l8 * 4]);
                        const float hv[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
                        for (int v2 = 0; v2 < 4; ++v2) {
                            const bool upd = (k2 > kk) || (l8 > l) || (l8 == l && v2 > v);
                            if (upd) w[k2 * 4 + v2] = __fsub_rn(w[k2 * 4 + v2], __fmul_rn(err, hv[v2]));
                        }
                    }
                    if (l8 == l) {                                       // owner keeps the results
                        w[kk * 4 + v] = err;
                        qv[kk * 4 + v] = q;
                        cpk[kk] |= ((uint32_t)code & 0xFFu) << (8 * v);
                    }
                }
            }
        }
    }

    // ---- outputs: Q, codes, Err1 (16-byte stores), loss
    if (rvalid) {
        const bool qvec = ((a.ldq & 3) == 0) && ((i1 & 3) == 0) && ((((uintptr_t)a.Q) & 15) == 0);
        const bool cvec = ((a.ldc & 3) == 0) && ((i1 & 3) == 0) && ((((uintptr_t)a.codes) & 3) == 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int c = kk * 32 + l8 * 4;
            if (c >= count) continue;
            float* qp = a.Q + (int64_t)row * a.ldq + i1 + c;
            uint8_t* cp = a.codes + (int64_t)row * a.ldc + i1 + c;
            float* ep = a.Err + (int64_t)row * a.lde + c;
            if (c + 3 < count) {
                if (qvec) *reinterpret_cast<float4*>(qp) = make_float4(qv[kk * 4], qv[kk * 4 + 1], qv[kk * 4 + 2], qv[kk * 4 + 3]);
                else { qp[0] = qv[kk * 4]; qp[1] = qv[kk * 4 + 1]; qp[2] = qv[kk * 4 + 2]; qp[3] = qv[kk * 4 + 3]; }
                if (cvec) *reinterpret_cast<uint32_t*>(cp) = cpk[kk];
                else { cp[0] = cpk[kk] & 0xFF; cp[1] = (cpk[kk] >> 8) & 0xFF; cp[2] = (cpk[kk] >> 16) & 0xFF; cp[3] = cpk[kk] >> 24; }
                *reinterpret_cast<float4*>(ep) = make_float4(w[kk * 4], w[kk * 4 + 1], w[kk * 4 + 2], w[kk * 4 + 3]);
            } else {
#pragma unroll
                for (int v = 0; v < 4; ++v)
                    if (c + v < count) { qp[v] = qv[kk * 4 + v]; cp[v] = (cpk[kk] >> (8 * v)) & 0xFF; ep[v] = w[kk * 4 + v]; }
            }
        }
    }
    if (!rvalid || l8 != 0) loss = 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    if (lane == 0) s_loss[warp] = loss;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int i = 0; i < BQ_THREADS / 32; ++i) t += s_loss[i];
        atomicAdd(a.loss, t * 0.5);                                      // Losses1 / 2, gptq.py:274
    }
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
                                 uint8_t* codes, int64_t ldc, double* loss, void* ws, size_t ws_bytes, void* stream) {
    GQ_ARG(W, 1, "null"); GQ_ARG(R >= 0, 2, "R<0"); GQ_ARG(K > 0, 3, "K<=0"); GQ_ARG(ldw >= K, 4, "ldw<K");
    GQ_ARG(U, 5, "null"); GQ_ARG(ldu >= K, 6, "ldu<K");
    GQ_ARG(blocksize > 0 && blocksize <= BQ_B, 7, "blocksize must be in [1,128]");
    GQ_ARG(mode == 0 || mode == 1, 9, "mode must be 0 or 1");
    GQ_ARG(mode == 0 || groupsize > 0, 8, "dynamic groups need groupsize > 0");
    GQ_ARG(scale && zero, 11, "null"); GQ_ARG(ngroups >= 1, 13, "ngroups<1");
    GQ_ARG(ngroups == 1 || gidx, 10, "gidx required when ngroups > 1");
    GQ_ARG(maxq != 0, 14, "maxq == 0 (Quantizer not configured)");
    GQ_ARG(Q, 20, "null"); GQ_ARG(ldq >= K, 21, "ldq<K"); GQ_ARG(codes, 22, "null"); GQ_ARG(ldc >= K, 23, "ldc<K");
    GQ_ARG(loss, 24, "null"); GQ_ARG(ws, 25, "null");
    GQ_ARG(ws_bytes >= gptq_b200_colloop_workspace_bytes(R, K, blocksize), 26, "workspace too small");
    if (R == 0) return 0;
    cudaStream_t st = as_stream(stream);
    float* Err = static_cast<float*>(ws);
    const size_t smem = (size_t)(BQ_B * BQ_B + 2 * BQ_ROWS * BQ_SP) * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
        GQ_CUDA(cudaFuncSetAttribute(blockquant_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    GQ_CUDA(cudaMemsetAsync(loss, 0, sizeof(double), st));
    // Two-level lazy update (same arithmetic as gptq.py:276, regrouped): after every block only the rest of the current
    // super-block is brought up to date (narrow GEMM, depth `count`); the columns beyond it receive ONE deep GEMM per
    // super-block.  Dynamic groups wider than a block read W far ahead (gptq.py:255), so they keep the one-level form.
    const int sb_blocks = (mode == 1 && groupsize > blocksize) ? 1 : BQ_SUPER;
    const int64_t lde = (int64_t)BQ_SUPER * BQ_B;
    int sb_start = 0;
    for (int i1 = 0, bi = 0; i1 < K; i1 += blocksize, ++bi) {
        const int i2 = (i1 + blocksize < K) ? i1 + blocksize : K;
        const int count = i2 - i1;
        if (bi % sb_blocks == 0) sb_start = i1;
        int sb_end = sb_start + sb_blocks * blocksize; if (sb_end > K) sb_end = K;
        if (mode == 1) {
            const int first = ((i1 + groupsize - 1) / groupsize) * groupsize;
            if (first < i2) {
                const int last = ((i2 - 1) / groupsize) * groupsize;
                int end = last + groupsize; if (end > K) end = K;
                int rc = gptq_b200_find_params(W, R, K, ldw, first, end - first, groupsize, maxq, sym, mse, norm, grid,
                                               maxshrink, scale, zero, ngroups, first / groupsize, stream);
                if (rc) return rc;
            }
        }
        float* Eblk = Err + (i1 - sb_start);
        BQArgs a{W, R, K, ldw, U, ldu, i1, count, gidx, scale, zero, ngroups, (float)maxq, maxq < 0 ? 1 : 0,
                 Q, ldq, codes, ldc, Eblk, lde, loss};
        blockquant_kernel<<<cdiv(R, BQ_ROWS), BQ_THREADS, smem, st>>>(a);
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
