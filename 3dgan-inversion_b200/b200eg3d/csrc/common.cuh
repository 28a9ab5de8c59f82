// Shared helpers for the b200eg3d CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#define B200_API extern "C" __attribute__((visibility("default")))

// Last-error string returned by b200_last_error(); thread-local so concurrent host threads do not race.
extern thread_local char g_b200_err[512];

static inline int b200_fail(const char* file, int line, const char* msg) {
    snprintf(g_b200_err, sizeof(g_b200_err), "%s:%d: %s", file, line, msg);
    return 1;
}

#define B200_REQUIRE(cond, msg) do { if (!(cond)) return b200_fail(__FILE__, __LINE__, msg); } while (0)
#define B200_CHECK_LAUNCH() do { cudaError_t e_ = cudaGetLastError(); \
    if (e_ != cudaSuccess) return b200_fail(__FILE__, __LINE__, cudaGetErrorString(e_)); } while (0)
#define B200_CUDA(call) do { cudaError_t e_ = (call); \
    if (e_ != cudaSuccess) return b200_fail(__FILE__, __LINE__, cudaGetErrorString(e_)); } while (0)

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// (a0, a1) += g * (v0, v1) as one sm_100 packed-FP32 instruction (fma.rn.f32x2 -> FFMA2)
__device__ __forceinline__ void fma2(float& a0, float& a1, float g, float v0, float v1) {
    unsigned long long acc, gg, vv;
    asm("mov.b64 %0, {%1, %2};" : "=l"(acc) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(gg) : "f"(g));
    asm("mov.b64 %0, {%1, %2};" : "=l"(vv) : "f"(v0), "f"(v1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(gg), "l"(vv));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(acc));
}

__device__ __forceinline__ float softplus_f(float x) {          // torch softplus, beta=1, threshold=20
    return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }
