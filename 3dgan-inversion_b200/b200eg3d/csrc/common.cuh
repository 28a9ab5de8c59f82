// Shared helpers for the b200eg3d CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <utility>

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

// Per-device one-time state.  cudaFuncSetAttribute and the SM count are properties of a DEVICE, so nothing here may be cached
// in a process-wide flag: a second GPU used from the same process would otherwise launch with the default 48 KB shared-memory
// limit.  Benign race: two threads may both run the set-up for a device, which is idempotent.
constexpr int B200_MAX_DEVICES = 64;
static inline int b200_current_device() { int d = 0; return cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < B200_MAX_DEVICES ? d : 0; }
struct B200PerDeviceFlag {
    volatile unsigned char done[B200_MAX_DEVICES];
    bool test(int dev) const { return done[dev] != 0; }
    void set(int dev) { done[dev] = 1; }
};
// set `attr` of kernel `fn` to `value` once per device
#define B200_FUNC_ATTR_ONCE(fn, attr, value) do { static B200PerDeviceFlag flag_ = {}; const int dev_ = b200_current_device(); \
    if (!flag_.test(dev_)) { B200_CUDA(cudaFuncSetAttribute(fn, attr, value)); flag_.set(dev_); } } while (0)
static inline int b200_sm_count() {
    static int sms[B200_MAX_DEVICES] = {};
    const int dev = b200_current_device();
    if (!sms[dev]) { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148; sms[dev] = v; }
    return sms[dev];
}

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may be
// scheduled while its predecessor in the stream is still draining; it must execute pdl_wait() (griddepcontrol.wait: all
// prerequisite grids complete and their memory visible) before touching global memory.  The step is ~250 dependent launches,
// most of them short: overlapping launch latency / prologue with the previous kernel's tail is worth several percent.
// B200EG3D_PDL=0 disables the attribute (the wait then returns immediately).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Let the next PDL kernel of the stream be scheduled as soon as every CTA of this one has started (it still blocks in its
// own pdl_wait() until this grid has completed): its CTAs then sit ready on the SMs when our last CTA retires.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

extern int g_b200_pdl;            // -1: take B200EG3D_PDL from the environment on first use; 0 / 1: set by b200_set_pdl()
static inline bool b200_pdl_enabled() {
    if (g_b200_pdl < 0) { const char* e = getenv("B200EG3D_PDL"); g_b200_pdl = (e && e[0] == '0') ? 0 : 1; }
    return g_b200_pdl == 1;
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = b200_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

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
