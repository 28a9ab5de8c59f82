// Throughput of the ways to add a 128-byte texel line into global memory, per SM (optimisation aid for the plane-gradient scatter):
//   (a) red.global.add.v4.f32, 8 lanes per line (what the scatter does)      (b) cp.reduce.async.bulk 128 B from shared memory, one thread per line
//   (c) red.global.add.f32 scalar, 32 lanes per line
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__global__ void k_red_v4(float* dst, int lines_mask, int iters) {
    const int lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int i = 0; i < iters; ++i) {
        const uint32_t line = hash32((gw * iters + i) * 4 + (lane >> 3)) & lines_mask;
        float* a = dst + (size_t)line * 32 + (lane & 7) * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(a), "f"(1.f) : "memory");
    }
}
__global__ void k_gather_v4(const float* __restrict__ src, float* out, int lines_mask, int iters) {      // what the sampler's gather does
    const int lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int i = 0; i < iters; ++i) {
        const uint32_t line = hash32((gw * iters + i) * 4 + (lane >> 3)) & lines_mask;
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + (size_t)line * 32 + (lane & 7) * 4));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (acc.x + acc.y + acc.z + acc.w == 1234.5f) out[0] = acc.x;
}
__global__ void k_red_s(float* dst, int lines_mask, int iters) {
    const int lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int i = 0; i < iters; ++i) {
        const uint32_t line = hash32(gw * iters + i) & lines_mask;
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + (size_t)line * 32 + lane), "f"(1.f) : "memory");
    }
}
__global__ void k_bulk(float* dst, int lines_mask, int iters, int bytes) {
    extern __shared__ __align__(128) float sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 1.f;
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t src = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 128;
    for (int i = 0; i < iters; ++i) {
        const uint32_t line = hash32(gt * iters + i) & lines_mask;
        float* a = dst + (size_t)line * 32;
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(a), "r"(src), "r"(bytes) : "memory");
        if ((i & 7) == 7) { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); }
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
int main() {
    const int nlines = 1 << 18;            // 32 MB of 128-byte lines (the plane gradient is 25 MB)
    float* d; cudaMalloc(&d, (size_t)nlines * 128); cudaMemset(d, 0, (size_t)nlines * 128);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int warps = 8; warps <= 32; warps *= 2) {
        const int iters = 2048, mask = (25 << 20) / 128 - 1 >= (1 << 17) ? (1 << 17) - 1 : (1 << 17) - 1;      // 16 MB of lines (power of two below the 25 MB planes)
        k_gather_v4<<<148, warps * 32>>>(d, d, mask, 64);
        cudaEventRecord(e0); k_gather_v4<<<148, warps * 32>>>(d, d, mask, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double lines = 148.0 * warps * iters * 4;
        printf("gather.v4 %2d warps/SM: %8.3f ms  %6.2f clk/line/SM @1.9GHz  %7.1f GB/s (L2-resident 128-byte lines, random)\n", warps, ms,
               ms * 1e-3 * 1.9e9 / (lines / 148), lines * 128 / ms / 1e6);
    }
    for (int warps = 4; warps <= 16; warps *= 2) {
        const int iters = 2048;
        k_red_v4<<<148, warps * 32>>>(d, nlines - 1, 64);
        cudaEventRecord(e0); k_red_v4<<<148, warps * 32>>>(d, nlines - 1, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double lines = 148.0 * warps * iters * 4;
        printf("red.v4   %2d warps/SM: %8.3f ms  %6.2f clk/line/SM @1.9GHz  %7.1f GB/s\n", warps, ms, ms * 1e-3 * 1.9e9 / (lines / 148), lines * 128 / ms / 1e6);
        cudaEventRecord(e0); k_red_s<<<148, warps * 32>>>(d, nlines - 1, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        lines = 148.0 * warps * iters;
        printf("red.f32  %2d warps/SM: %8.3f ms  %6.2f clk/line/SM           %7.1f GB/s\n", warps, ms, ms * 1e-3 * 1.9e9 / (lines / 148), lines * 128 / ms / 1e6);
    }
    for (int threads = 32; threads <= 512; threads *= 4)
        for (int bytes = 128; bytes <= 512; bytes *= 2) {
            const int iters = 1024;
            k_bulk<<<148, threads, 32768>>>(d, nlines - 4, 64, bytes);
            cudaEventRecord(e0); k_bulk<<<148, threads, 32768>>>(d, nlines - 4, iters, bytes); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            double ops = 148.0 * threads * iters;
            printf("bulk %3dB %3d thr/SM:  %8.3f ms  %6.2f clk/op/SM  %7.1f GB/s  (%s)\n", bytes, threads, ms, ms * 1e-3 * 1.9e9 / (ops / 148), ops * bytes / ms / 1e6,
                   cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
