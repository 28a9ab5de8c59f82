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

// grid (cout, n), block 256, dynamic smem = cin*taps floats (the modulated row W[o]*s, staged once: the global read
// is coalesced in the parameter layout [i][t], the writes are coalesced in the GEMM layout [t][i]).
template <int TAPS>
__global__ void __launch_bounds__(256) weight_prep_kernel(const float* __restrict__ W, const float* __restrict__ styles,
                                                          float* __restrict__ wmod, __nv_bfloat16* __restrict__ whi,
                                                          __nv_bfloat16* __restrict__ wlo, float* __restrict__ dcoef,
                                                          int cout, int cin, int taps_rt, int demod) {
    const int taps = TAPS > 0 ? TAPS : taps_rt;
    extern __shared__ float sW[];
    __shared__ float red[32];
    const int o = blockIdx.x, n = blockIdx.y;
    const float* Wo = W + (long)o * cin * taps;
    const float* s = styles + (long)n * cin;
    float acc = 0.f;
    for (int idx = threadIdx.x; idx < cin * taps; idx += blockDim.x) {
        const float v = Wo[idx] * s[idx / taps];
        sW[idx] = v;
        acc = fmaf(v, v, acc);
    }
    float d = 1.f;
    if (demod) {
        acc = block_sum(acc, red);
        d = rsqrtf(acc + 1e-8f);
        if (threadIdx.x == 0 && dcoef) dcoef[(long)n * cout + o] = d;
    } else {
        __syncthreads();
    }
    const long ob = (long)n * taps * cout * cin;
    if ((cin & 1) == 0) {
        const int half = cin >> 1;
        for (int t = 0; t < taps; ++t)
        for (int ih = threadIdx.x; ih < half; ih += blockDim.x) {
            const int i = ih * 2;
            const float v0 = sW[i * taps + t] * d, v1 = sW[(i + 1) * taps + t] * d;
            const long oi = ob + ((long)t * cout + o) * cin + i;
            if (wmod) *reinterpret_cast<float2*>(wmod + oi) = make_float2(v0, v1);
            if (whi) {
                const __nv_bfloat162 hh = __floats2bfloat162_rn(v0, v1);
                *reinterpret_cast<__nv_bfloat162*>(whi + oi) = hh;
                if (wlo) *reinterpret_cast<__nv_bfloat162*>(wlo + oi) =
                    __floats2bfloat162_rn(v0 - __low2float(hh), v1 - __high2float(hh));
            }
        }
    } else {
        for (int idx = threadIdx.x; idx < cin * taps; idx += blockDim.x) {
            const int t = idx / cin, i = idx % cin;
            const float v = sW[i * taps + t] * d;
            const long oi = ob + ((long)t * cout + o) * cin + i;
            if (wmod) wmod[oi] = v;
            if (whi) {
                const __nv_bfloat16 hh = __float2bfloat16_rn(v);
                whi[oi] = hh;
                if (wlo) wlo[oi] = __float2bfloat16_rn(v - __bfloat162float(hh));
            }
        }
    }
}

// grid (cout), block 256, dynamic smem = 2*cin*taps floats.  dW is written (not accumulated); dstyles must be zeroed by
// the caller (atomics).  The parameter row W[o] and the outgoing dW[o] row are staged in shared memory so that every
// global access is coalesced.
template <int TAPS>
__global__ void __launch_bounds__(256) weight_prep_bwd_kernel(const float* __restrict__ W, const float* __restrict__ styles,
                                                              const float* __restrict__ dcoef, const float* __restrict__ dwmod,
                                                              float* __restrict__ dW, float* __restrict__ dstyles,
                                                              int nb, int cout, int cin, int taps_rt, int demod) {
    const int taps = TAPS > 0 ? TAPS : taps_rt;
    extern __shared__ float sm[];
    float* sW = sm;                       // W[o][i][t]
    float* sD = sm + cin * taps;          // dW[o][i][t]
    __shared__ float red[32];
    const int o = blockIdx.x;
    const float* Wo = W + (long)o * cin * taps;
    for (int idx = threadIdx.x; idx < cin * taps; idx += blockDim.x) { sW[idx] = Wo[idx]; sD[idx] = 0.f; }
    __syncthreads();
    for (int n = 0; n < nb; ++n) {
        const float* s = styles + (long)n * cin;
        const float* G = dwmod + (long)n * taps * cout * cin;
        float d = 1.f, dot = 0.f;
        if (demod) {
            d = dcoef[(long)n * cout + o];
            float acc = 0.f;
            for (int t = 0; t < taps; ++t)
                for (int i = threadIdx.x; i < cin; i += blockDim.x)
                    acc = fmaf(G[((long)t * cout + o) * cin + i], sW[i * taps + t] * s[i], acc);
            dot = block_sum(acc, red);
        }
        const float d3dot = d * d * d * dot;
        for (int i = threadIdx.x; i < cin; i += blockDim.x) {
            const float si = s[i];
            float ds = 0.f;
            for (int t = 0; t < taps; ++t) {
                const float w = sW[i * taps + t];
                float g = G[((long)t * cout + o) * cin + i] * d;
                if (demod) g -= d3dot * w * si;
                ds = fmaf(w, g, ds);
                sD[i * taps + t] += si * g;
            }
            if (dstyles) atomicAdd(dstyles + (long)n * cin + i, ds);
        }
    }
    __syncthreads();
    float* dWo = dW + (long)o * cin * taps;
    for (int idx = threadIdx.x; idx < cin * taps; idx += blockDim.x) dWo[idx] = sD[idx];
}

B200_API int b200_modconv_weight_prep(const float* W, const float* styles, float* wmod, void* w_hi, void* w_lo, float* dcoef,
                                      int n, int cout, int cin, int taps, int demod, void* stream) {
    B200_REQUIRE(n > 0 && cout > 0 && cin > 0 && taps > 0, "weight_prep: bad shape");
    B200_REQUIRE(wmod || w_hi, "weight_prep: no output requested");
    const size_t smem = sizeof(float) * (size_t)cin * taps;
    B200_REQUIRE(smem <= 48 * 1024, "weight_prep: cin*taps too large for the staging tile");
    if (taps == 9)
        weight_prep_kernel<9><<<dim3(cout, n), 256, smem, (cudaStream_t)stream>>>(W, styles, wmod, (__nv_bfloat16*)w_hi,
                                                                                  (__nv_bfloat16*)w_lo, dcoef, cout, cin, taps, demod);
    else if (taps == 1)
        weight_prep_kernel<1><<<dim3(cout, n), 256, smem, (cudaStream_t)stream>>>(W, styles, wmod, (__nv_bfloat16*)w_hi,
                                                                                  (__nv_bfloat16*)w_lo, dcoef, cout, cin, taps, demod);
    else
        weight_prep_kernel<0><<<dim3(cout, n), 256, smem, (cudaStream_t)stream>>>(W, styles, wmod, (__nv_bfloat16*)w_hi,
                                                                                  (__nv_bfloat16*)w_lo, dcoef, cout, cin, taps, demod);
    B200_CHECK_LAUNCH();
    return 0;
}

B200_API int b200_modconv_weight_prep_bwd(const float* W, const float* styles, const float* dcoef, const float* dwmod,
                                          float* dW, float* dstyles, int n, int cout, int cin, int taps, int demod,
                                          void* stream) {
    B200_REQUIRE(n > 0 && cout > 0 && cin > 0 && taps > 0, "weight_prep_bwd: bad shape");
    B200_REQUIRE(!demod || dcoef, "weight_prep_bwd: demodulation needs the saved coefficients");
    if (dstyles) B200_CUDA(cudaMemsetAsync(dstyles, 0, sizeof(float) * (size_t)n * cin, (cudaStream_t)stream));
    const size_t smem = 2 * sizeof(float) * (size_t)cin * taps;
    B200_REQUIRE(smem <= 48 * 1024, "weight_prep_bwd: cin*taps too large for the staging tiles");
    if (taps == 9)
        weight_prep_bwd_kernel<9><<<cout, 256, smem, (cudaStream_t)stream>>>(W, styles, dcoef, dwmod, dW, dstyles, n, cout, cin, taps, demod);
    else if (taps == 1)
        weight_prep_bwd_kernel<1><<<cout, 256, smem, (cudaStream_t)stream>>>(W, styles, dcoef, dwmod, dW, dstyles, n, cout, cin, taps, demod);
    else
        weight_prep_bwd_kernel<0><<<cout, 256, smem, (cudaStream_t)stream>>>(W, styles, dcoef, dwmod, dW, dstyles, n, cout, cin, taps, demod);
    B200_CHECK_LAUNCH();
    return 0;
}
