/*
 * gptq_b200 -- C ABI of the B200-native GPTQ layer-wise quantization hot path.
 *
 * One shared library (libgptq_b200.so), plain pointers and sizes, no torch types.
 * The reference (JerryZhou54/gptq) has NO native boundary on this path: every step below is a
 * chain of eager torch ops inside gptq.py / quant.py.  Each entry point therefore cites the
 * reference Python lines whose arithmetic it replaces; INTEGRATION.md shows the ctypes binding
 * a reference maintainer would add to gptq.py / quant.py.
 *
 * Conventions (all functions):
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - matrices are row-major with an explicit leading dimension in ELEMENTS;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *   - nothing allocates: ops that need scratch take (ws, ws_bytes) sized by *_workspace_bytes();
 *   - nothing synchronises the host; work is enqueued on `stream`;
 *   - return 0 on success, -k if argument k (1-based) is invalid, >0 for a CUDA error code;
 *     gptq_b200_last_error() returns a thread-local message for the last non-zero return.
 *   - no global mutable state besides a per-process cache of kernel attributes.
 */
#ifndef GPTQ_B200_H
#define GPTQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPTQ_B200_ABI_VERSION 3

enum { GPTQ_B200_F16 = 0, GPTQ_B200_BF16 = 1, GPTQ_B200_F32 = 2 };

int         gptq_b200_abi_version(void);
const char* gptq_b200_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py "gpu_launches"). */
int64_t     gptq_b200_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * (1) Hessian accumulation.  Replaces gptq.py:59-65 (H *= n/(n+1); H += X X^T with
 *     X = sqrt(2/n) inp^T): the caller folds the running-mean factors into alpha/beta.
 *
 *     H[0:K,0:K] <- beta * H + alpha * X^T X      (X: [tokens, K], f16 or bf16, ldx % 8 == 0)
 *
 *     tcgen05.mma kind::f16 (fp32 accumulate in TMEM) fed by TMA.  Only the 128x256 tiles that
 *     intersect the LOWER triangle are computed; entries of H above those tiles are left
 *     untouched -- call gptq_b200_symmetrize() before reading H as a full matrix.
 *     `chunk_tokens` (0 = default) bounds how many tokens are summed inside the tensor-core
 *     accumulator before being added to H in fp32 (round-to-nearest) by the epilogue.
 *     `variant`: GPTQ_B200_SYRK_AUTO / _CTA2 = the CTA-pair kernel (tcgen05 cta_group::2, M256 x N256 tiles over two SMs
 *     of a TPC; tiles touching the lower triangle at 256 x 256 granularity), _CTA1 = the single-CTA kernel (M128 x N256).
 *     Both produce the same H up to fp32 summation order inside the tensor core.
 * ------------------------------------------------------------------------------------------ */
enum { GPTQ_B200_SYRK_AUTO = 0 /* = _CTA2_BK128 */, GPTQ_B200_SYRK_CTA1 = 1, GPTQ_B200_SYRK_CTA2 = 2 /* 64-token stages x 6 */,
       GPTQ_B200_SYRK_CTA2_BK128 = 3 /* 128-token stages x 3 */ };
/* (bits 8-24 of `variant` carry tile-order / TMA tuning values for performance experiments; 0 = the shipped defaults) */
size_t gptq_b200_syrk_workspace_bytes(int K, int64_t tokens, int variant);
int gptq_b200_syrk_accum(const void* X, int dtype, int64_t tokens, int K, int64_t ldx,
                         float* H, int64_t ldh, float alpha, float beta, int chunk_tokens, int variant,
                         void* ws, size_t ws_bytes, void* stream);
/* Plain CUDA-core fp32 X^T X with the same contract (test cross-check for the tensor path;
 * never called by the Python package). */
int gptq_b200_syrk_accum_simt(const void* X, int dtype, int64_t tokens, int K, int64_t ldx,
                              float* H, int64_t ldh, float alpha, float beta, void* stream);
/* H[i,j] <- H[j,i] for j > i (mirror the lower triangle). */
int gptq_b200_symmetrize(float* H, int K, int64_t ldh, void* stream);
/* Multi-GPU exchange #1: the all-reduce of the partial accumulators carries the lower triangle only.
 * packed[r (r + 1) / 2 + c] = S[r][c], c <= r  (K (K + 1) / 2 floats); unpack writes the same entries back. */
int gptq_b200_tril_pack(const float* S, int K, int64_t lds, float* packed, void* stream);
int gptq_b200_tril_unpack(const float* packed, int K, float* S, int64_t lds, void* stream);
/* Copy one calibration batch [tokens,K] (any ld) into a staging slab with ld_dst % 8 == 0. */
int gptq_b200_stage_input(const void* src, int dtype, int64_t tokens, int K, int64_t ld_src,
                          void* dst, int64_t ld_dst, void* stream);
/* *flag |= 1 if the 16-bit batches a[tokens,K] and b[tokens,K] differ in any bit (flag is NOT cleared).
 * gptq.py:38-65 recomputes the same H for every linear that sees the same activations (q/k/v,
 * gate/up); the host shares one accumulator between them, after this check proves the inputs equal. */
int gptq_b200_compare_rows(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t tokens, int K,
                           int32_t* flag, void* stream);

/* ------------------------------------------------------------------------------------------
 * (2) fasterquant prologue.  Replaces gptq.py:130-145 (clone().float(), dead columns),
 *     gptq.py:165-169 (act-order argsort + gathers) and gptq.py:174-176 (damp).
 * ------------------------------------------------------------------------------------------ */
/* dead = diag(H)==0: H[dead,dead]=1; diag_out[K] = diag(H); *damp_out = percdamp*mean(diag). */
int gptq_b200_hessian_prepare(float* H, int K, int64_t ldh, float percdamp,
                              float* diag_out, float* damp_out, int32_t* ndead_out, void* stream);
/* W32[r,c] = (float)weight[r,c], 0 where diag[c] was dead (diag_fixed[c]==1 && flagged). */
int gptq_b200_weights_prepare(const void* weight, int dtype, int R, int K, int64_t ldw,
                              const uint8_t* dead /*[K] or NULL*/, float* W32, int64_t ldw32,
                              void* stream);
/* dead[c] = (H[c,c] == 0) BEFORE hessian_prepare fixes it (call first). */
int gptq_b200_dead_columns(const float* H, int K, int64_t ldh, uint8_t* dead, void* stream);
/* perm = argsort(v, descending), ties by ascending index; invperm[perm[i]] = i. */
int gptq_b200_argsort_desc(const float* v, int n, int32_t* perm, int32_t* invperm, void* stream);
/* out[r, c] = W[r, perm[c]] */
int gptq_b200_permute_cols(const float* W, int R, int K, int64_t ldw, const int32_t* perm,
                           float* out, int64_t ldo, void* stream);
/* A[i,j] = H[p(i),p(j)] (+ *damp on the diagonal) for i,j < K; identity on the pad [K,Kp).
 * perm may be NULL (identity).  A is [Kp,Kp], Kp % 128 == 0. */
int gptq_b200_hessian_gather(const float* H, int K, int64_t ldh, const int32_t* perm,
                             const float* damp, float* A, int Kp, int64_t lda, void* stream);

/* ------------------------------------------------------------------------------------------
 *     Inverse-Cholesky chain.  Replaces gptq.py:177-179
 *         H = cholesky(H); H = cholesky_inverse(H); H = cholesky(H, upper=True)
 *     A (damped, [Kp,Kp], Kp % 128 == 0, upper triangle read) is DESTROYED; U ([Kp,Kp], a separate
 *     matrix) receives the upper factor of A^-1 (A^-1 = U^T U) with a zero strict lower triangle.
 *     Computed as U = R^-1 with A = R R^T (R upper): blocked backward Cholesky fused with a
 *     blocked triangular inverse (chol.cu).  From 8 blocks of 128 up to K = 24576 the work is
 *     issued as right-looking panels with look-ahead on three library-internal streams that are
 *     forked from `stream` and joined back into it with events before the call returns: for the
 *     caller the call is asynchronous and ordered on `stream` like every other entry point (A,
 *     U, info and ws must stay valid until work queued on `stream` after the call has run, as
 *     usual); see gptq_b200_set_chain_schedule.  The result does not depend on how the internal
 *     streams interleave (bitwise reproducible).
 *     *info (device int32): 0 ok, i+1 if the pivot at index i was not positive
 *     (torch.linalg.LinAlgError in the reference).
 * ------------------------------------------------------------------------------------------ */
size_t gptq_b200_hinv_upper_workspace_bytes(int Kp);
int gptq_b200_hinv_upper(float* A, int Kp, int64_t lda, float* U, int64_t ldu, int32_t* info,
                         void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * (3) Grid search.  Replaces Quantizer.find_params (quant.py:37-118, weight layout) applied to
 *     W[:, col0 + g*groupsize : col0 + (g+1)*groupsize] for every group g that starts inside
 *     [col0, col0+ncols) (the slice is clipped to K like Python slicing).  groupsize <= 0 means
 *     one group = the whole row range.  Output: scale/zero[r*ldsz + gout0 + g].
 *     maxq < 0 selects trits (quant.py:69-71).
 * ------------------------------------------------------------------------------------------ */
int gptq_b200_find_params(const float* W, int R, int K, int64_t ldw, int col0, int ncols,
                          int groupsize, int maxq, int sym, int mse, float norm, int grid,
                          float maxshrink, float* scale, float* zero, int64_t ldsz, int gout0,
                          void* stream);
/* Dequantised weights from integer codes: out[r][c] = T(scale[r][g(c)] * (code[r][c] - zero[r][g(c)])) (quant.py:10 and the cast
 * of gptq.py:305) -- bit-identical to what the column loop wrote for those codes.  Multi-GPU exchange #2 ships codes (1 byte per
 * weight) and every rank rebuilds the other ranks' rows with this. */
int gptq_b200_dequant_codes(const uint8_t* codes, int64_t ldc, int R, int K, const float* scale,
                            const float* zero, int ngroups, const int32_t* g_idx,
                            void* weight_out, int dtype, int64_t ldwo, void* stream);
/* quant.py:6-10 elementwise: out = scale*(clamp(round(x/scale)+zero,0,maxq)-zero).
 * scale/zero index = r*s_row + c*s_col (so [R,1] -> (1,0); per-tensor -> (0,0); [1,C] -> (0,1)). */
int gptq_b200_quantize(const float* x, int R, int C, int64_t ldx, const float* scale,
                       const float* zero, int64_t s_row, int64_t s_col, int maxq,
                       float* out, int64_t ldo, void* stream);

/* ------------------------------------------------------------------------------------------
 *     Column loop.  Replaces gptq.py:191-276 (standard branch) for all K columns:
 *     per `blocksize` columns, a persistent-tile kernel keeps each row's W1 slice in registers
 *     and Hinv1 in shared memory, quantizes column by column with error feedback
 *     (gptq.py:262-271), then the lazy update W[:, i2:] -= Err1 @ U[i1:i2, i2:] (gptq.py:276)
 *     runs as an fp32-accurate GEMM.
 *
 *     W      [R,K] fp32 working copy (already permuted); destroyed.
 *     U      upper factor from gptq_b200_hinv_upper.
 *     gidx   [K] group index of every (permuted) column into scale/zero [R,ngroups].
 *     mode   0: params given (per-row or static groups) -- scale/zero are inputs;
 *            1: dynamic groups (gptq.py:253-255): params of group g are fitted to the CURRENT
 *               W[:, g*gs:(g+1)*gs] when the block containing column g*gs starts (this
 *               reproduces the reference's stale-weights behaviour for gs < blocksize) and
 *               written to scale/zero.
 *     Q      [R,K] fp32 dequantised weights, permuted order.
 *     codes  [R,K] uint8 integer codes, permuted order (trits: 0/1/2).
 *     row_loss [R] fp32, overwritten: row_loss[r] = sum_k Losses[r,k] (gptq.py:267,274); the scalar printed at
 *            gptq.py:294 is its sum (reduced by the caller, per linear, in double).
 * ------------------------------------------------------------------------------------------ */
size_t gptq_b200_colloop_workspace_bytes(int R, int K, int blocksize);
int gptq_b200_colloop(float* W, int R, int K, int64_t ldw, const float* U, int64_t ldu,
                      int blocksize, int groupsize, int mode, const int32_t* gidx,
                      float* scale, float* zero, int ngroups, int maxq, int sym, int mse,
                      float norm, int grid, float maxshrink,
                      float* Q, int64_t ldq, uint8_t* codes, int64_t ldc, float* row_loss,
                      void* ws, size_t ws_bytes, void* stream);
/* Device self-test of the column loop's division: counts pairs (a,b) out of n pseudo-random ones
 * for which the hoisted-reciprocal quotient differs from div.rn.f32 (must be 0).  *mismatches is a
 * device uint64. */
int gptq_b200_selftest_div(uint64_t n, uint32_t seed, int mode, unsigned long long* mismatches, void* stream);
/* gptq.py:300-305: weight_out[r, c] = (dtype) Q[r, invperm[c]]; codes_out likewise.
 * invperm may be NULL. */
int gptq_b200_finalize(const float* Q, int64_t ldq, const uint8_t* codes, int64_t ldc,
                       int R, int K, const int32_t* invperm, void* weight_out, int dtype,
                       int64_t ldwo, uint8_t* codes_out, int64_t ldco, void* stream);

/* ------------------------------------------------------------------------------------------
 * (4) Packing.  Replaces Quant3Linear.pack's loop (quant.py:161-186; bits=3), the 4-bit layout
 *     of zeroShot/models/quant.py:184-185 (bits=4) and, by analogy, bits=2 and 8.
 *     qweight is [K*bits/32, R] int32 (K % 32 == 0).
 *     pack_codes:   from uint8 codes [R,K].
 *     pack_weights: from dequantised weights, recovering codes as quant.py:153,158 does:
 *                   round((W + zero*scale)/scale), scale/zero = [R,ngroups] indexed by g_idx[c]
 *                   (g_idx NULL => group 0).
 * ------------------------------------------------------------------------------------------ */
int gptq_b200_pack_codes(const uint8_t* codes, int R, int K, int64_t ldc, int bits,
                         int32_t* qweight, void* stream);
int gptq_b200_pack_weights(const void* W, int dtype, int R, int K, int64_t ldw,
                           const float* scale, const float* zero, int ngroups,
                           const int32_t* g_idx, int bits, int32_t* qweight, void* stream);
/* Inverse of pack_codes (for load-time checks and the round-trip tests). */
int gptq_b200_unpack_codes(const int32_t* qweight, int R, int K, int bits,
                           uint8_t* codes, int64_t ldc, void* stream);

/* ------------------------------------------------------------------------------------------
 *     Packed-weight matvec (the consumer of the packed format; SURVEY 8f "next" #1).  Replaces
 *     quant_cuda.vecquant3matmul / vecquant3matmul_faster (quant_cuda.cpp:15-29,
 *     quant_cuda_kernel.cu:88-244) and the unshipped vecquant4matmul:
 *         y[r] += sum_k (scales[r,g(k)] * code[k,r] - zeros_scaled[r,g(k)]) * x[k]
 *     x: [K] f16/bf16/f32 (single token, like the reference); qweight: [K*bits/32, R]; y: [R] fp32,
 *     pre-loaded with the bias by the caller.  g(k) = g_idx[k] if given, else k / groupsize.
 * ------------------------------------------------------------------------------------------ */
int gptq_b200_vecquant_matvec(const void* x, int dtype, const int32_t* qweight, int R, int K, int bits,
                              const float* scales, const float* zeros_scaled, int ngroups,
                              const int32_t* g_idx, int groupsize, float* y, void* stream);

/* ------------------------------------------------------------------------------------------
 *     fp32 GEMM used by the chain and the lazy update (exposed for tests / roofline runs):
 *     C[M,N] <- alpha * op(A) op(B) + beta * C,  op = transpose if the flag is set.
 * ------------------------------------------------------------------------------------------ */
int gptq_b200_sgemm(int transa, int transb, int M, int N, int Kd, float alpha,
                    const float* A, int64_t lda, const float* B, int64_t ldb, float beta,
                    float* C, int64_t ldc, void* stream);
/* Same contract on the tensor cores: tcgen05 kind::tf32, operands split hi/lo on the fly,
 * hi*hi + hi*lo + lo*hi ("3xTF32"), fp32 accumulation drained to registers every 128 k.
 * Requires lda, ldb multiples of 4 and 16-byte aligned A, B. */
int gptq_b200_gemm3x(int transa, int transb, int M, int N, int Kd, float alpha,
                     const float* A, int64_t lda, const float* B, int64_t ldb, float beta,
                     float* C, int64_t ldc, void* stream);
/* Which of the two the chain and the lazy update use: 1 = tensor cores (default), 0 = CUDA cores
 * (also selectable with GPTQ_B200_GEMM_BACKEND=simt).  Process-wide; for tests and A/B timing. */
int gptq_b200_set_gemm_backend(int backend);
/* Schedule of gptq_b200_hinv_upper: lookahead 1 = right-looking panels with look-ahead on three internal streams (default
 * from 8 blocks of 128 on), 0 = the plain block recursion on the caller's stream; panel_blocks = blocks of 128 per panel
 * (0 = chosen by size, the default: 1 up to K = 4608, 2 up to 7168, 8 up to 24576, the plain recursion beyond).  A negative value keeps the current setting.  Also GPTQ_B200_CHAIN_LOOKAHEAD / GPTQ_B200_CHAIN_PANEL.
 * Query gptq_b200_hinv_upper_workspace_bytes AFTER changing it.  Process-wide; for tests and A/B timing. */
int gptq_b200_set_chain_schedule(int lookahead, int panel_blocks);

#ifdef __cplusplus
}
#endif
#endif /* GPTQ_B200_H */
