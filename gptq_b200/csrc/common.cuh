// Shared host/device helpers for libgptq_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/gptq_b200.h"

namespace gq {

extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;

int set_err(int code, const char* fmt, ...);

#define GQ_ARG(cond, k, msg) do { if (!(cond)) return gq::set_err(-(k), "%s: argument %d invalid: %s", __func__, (k), (msg)); } while (0)
#define GQ_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return gq::set_err((int)e__, "%s: %s -> %s", __func__, #expr, cudaGetErrorString(e__)); } while (0)
// Count a launch and surface launch-configuration errors immediately.
#define GQ_LAUNCHED() do { gq::g_launches.fetch_add(1, std::memory_order_relaxed); cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return gq::set_err((int)e__, "%s: launch failed: %s", __func__, cudaGetErrorString(e__)); } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

int num_sms();       // of the CURRENT device (cached per device)
int cur_device();

// Function attributes (cudaFuncSetAttribute) and SM counts are per-DEVICE properties: one process may drive several GPUs
// (a layer on cuda:1 while cuda:0 is current elsewhere), so "done once" is tracked per device ordinal.
struct PerDeviceOnce {
    std::atomic<uint64_t> mask[4] = {};
    bool done(int dev) const { return dev >= 0 && dev < 256 && ((mask[dev >> 6].load(std::memory_order_acquire) >> (dev & 63)) & 1ull); }
    void mark(int dev) { if (dev >= 0 && dev < 256) mask[dev >> 6].fetch_or(1ull << (dev & 63), std::memory_order_release); }
};
#define GQ_ONCE_PER_DEVICE(...) do { static gq::PerDeviceOnce once__; const int dev__ = gq::cur_device(); \
    if (!once__.done(dev__)) { __VA_ARGS__; once__.mark(dev__); } } while (0)

// Programmatic dependent launch (GPTQ_B200_PDL, default on): the kernel may be scheduled while its predecessor in the
// stream is still running; it must execute ptx::pdl_wait() before touching memory the predecessor writes.
// `allow` = false for a kernel that follows a memset / copy in the stream (only kernel -> kernel edges are relaxed).
extern int g_pdl;
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(bool allow, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (g_pdl && allow) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }

// quant.py:6-10 with the reference's operation order and roundings (no FMA contraction).
__device__ __forceinline__ float quant_code(float w, float scale, float zero, float maxq) {
    float t = __fadd_rn(rintf(__fdiv_rn(w, scale)), zero);
    return fminf(fmaxf(t, 0.f), maxq);
}
__device__ __forceinline__ float dequant(float code, float scale, float zero) {
    return __fmul_rn(scale, __fsub_rn(code, zero));
}
// trits grid, quant.py:7-8; returns q and a 0/1/2 code.
__device__ __forceinline__ float quant_trits(float w, float scale, float zero, float* code) {
    float hi = (w > __fmul_rn(scale, 0.5f)) ? 1.f : 0.f;
    float lo = (w < __fmul_rn(zero, 0.5f)) ? 1.f : 0.f;
    *code = 1.f + hi - lo;
    return __fadd_rn(__fmul_rn(hi, scale), __fmul_rn(lo, zero));
}

}  // namespace gq
