// Weight modulation / demodulation of modulated_conv2d (reference: training/networks_stylegan2.py:58-67)
// and its backward, producing weights directly in the GEMM layout wmod[n][tap][cout][cin].
#include "common.cuh"
#include <cuda_bf16.h>

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float s = 0.f;
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; ++i) s += red[i];
    return s;
}

// grid (cout, n), block 256
__global__ void __launch_bounds__(256) weight_prep_kernel(const float* __restrict__ W, const float* __restrict__ styles,
                                                          float* __restrict__ wmod, __nv_bfloat16* __restrict__ whi,
                                                          __nv_bfloat16* __restrict__ wlo, float* __restrict__ dcoef,
                                                          int cout, int cin, int taps, int demod) {
    __shared__ float red[32];
    const int o = blockIdx.x, n = blockIdx.y;
    const float* Wo = W + (long)o * cin * taps;
    const float* s = styles + (long)n * cin;
    float d = 1.f;
    if (demod) {
        float acc = 0.f;
        for (int idx = threadIdx.x; idx < cin * taps; idx += blockDim.x) {
            const float v = Wo[idx] * s[idx / taps];
            acc = fmaf(v, v, acc);
        }
        acc = block_sum(acc, red);
        d = rsqrtf(acc + 1e-8f);
        if (threadIdx.x == 0 && dcoef) dcoef[(long)n * cout + o] = d;
    }
    const long ob = (long)n * taps * cout * cin;
    for (int idx = threadIdx.x; idx < cin * taps; idx += blockDim.x) {
        const int t = idx / cin, i = idx % cin;
        const float v = Wo[i * taps + t] * s[i] * d;
        const long oi = ob + ((long)t * cout + o) * cin + i;
        if (wmod) wmod[oi] = v;
        if (whi) {
            const __nv_bfloat16 hh = __float2bfloat16_rn(v);
            whi[oi] = hh;
            if (wlo) wlo[oi] = __float2bfloat16_rn(v - __bfloat162float(hh));
        }
    }
}

// grid (cout), block 256.  dW is written (not accumulated); dstyles must be zeroed by the caller (atomics).
__global__ void __launch_bounds__(256) weight_prep_bwd_kernel(const float* __restrict__ W, const float* __restrict__ styles,
                                                              const float* __restrict__ dcoef, const float* __restrict__ dwmod,
                                                              float* __restrict__ dW, float* __restrict__ dstyles,
                                                              int nb, int cout, int cin, int taps, int demod) {
    __shared__ float red[32];
    const int o = blockIdx.x;
    const float* Wo = W + (long)o * cin * taps;
    for (int n = 0; n < nb; ++n) {
        const float* s = styles + (long)n * cin;
        const float* G = dwmod + (long)n * taps * cout * cin;
        float d = 1.f, dot = 0.f;
        if (demod) {
            d = dcoef[(long)n * cout + o];
            float acc = 0.f;
            for (int idx = threadIdx.x; idx < cin * taps; idx += blockDim.x) {
                const int t = idx / cin, i = idx % cin;
                acc = fmaf(G[((long)t * cout + o) * cin + i], Wo[i * taps + t] * s[i], acc);
            }
            dot = block_sum(acc, red);
        }
        const float d3dot = d * d * d * dot;
        for (int i = threadIdx.x; i < cin; i += blockDim.x) {
            const float si = s[i];
            float ds = 0.f;
            for (int t = 0; t < taps; ++t) {
                const float w = Wo[i * taps + t];
                float g = G[((long)t * cout + o) * cin + i] * d;
                if (demod) g -= d3dot * w * si;
                ds = fmaf(w, g, ds);
                float* dst = dW + ((long)o * cin + i) * taps + t;
                if (n == 0) *dst = si * g; else *dst += si * g;
            }
            if (dstyles) atomicAdd(dstyles + (long)n * cin + i, ds);
        }
    }
}

B200_API int b200_modconv_weight_prep(const float* W, const float* styles, float* wmod, void* w_hi, void* w_lo, float* dcoef,
                                      int n, int cout, int cin, int taps, int demod, void* stream) {
    B200_REQUIRE(n > 0 && cout > 0 && cin > 0 && taps > 0, "weight_prep: bad shape");
    B200_REQUIRE(wmod || w_hi, "weight_prep: no output requested");
    weight_prep_kernel<<<dim3(cout, n), 256, 0, (cudaStream_t)stream>>>(W, styles, wmod, (__nv_bfloat16*)w_hi, (__nv_bfloat16*)w_lo,
                                                                        dcoef, cout, cin, taps, demod);
    B200_CHECK_LAUNCH();
    return 0;
}

B200_API int b200_modconv_weight_prep_bwd(const float* W, const float* styles, const float* dcoef, const float* dwmod,
                                          float* dW, float* dstyles, int n, int cout, int cin, int taps, int demod,
                                          void* stream) {
    B200_REQUIRE(n > 0 && cout > 0 && cin > 0 && taps > 0, "weight_prep_bwd: bad shape");
    B200_REQUIRE(!demod || dcoef, "weight_prep_bwd: demodulation needs the saved coefficients");
    if (dstyles) B200_CUDA(cudaMemsetAsync(dstyles, 0, sizeof(float) * (size_t)n * cin, (cudaStream_t)stream));
    weight_prep_bwd_kernel<<<cout, 256, 0, (cudaStream_t)stream>>>(W, styles, dcoef, dwmod, dW, dstyles, n, cout, cin, taps, demod);
    B200_CHECK_LAUNCH();
    return 0;
}
