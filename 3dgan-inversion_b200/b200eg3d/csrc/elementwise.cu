// Elementwise / FIR kernels of the synthesis stack:
//   b200_bias_act        generic bias+activation+gain+clamp, forward and first-order gradient
//                        (semantics of torch_utils/ops/bias_act.cu:28-151, plugin entry bias_act.cpp:36)
//   b200_layer_act_*     the SynthesisLayer epilogue: + noise*strength, + bias, lrelu, gain, clamp
//                        (training/networks_stylegan2.py:318-329) and its backward incl. all reductions
//   b200_upfirdn2d       pad / zero-insert upsample / FIR / decimate on NHWC tensors, optional fused "+ add"
//                        (semantics of torch_utils/ops/upfirdn2d.py:169-213, plugin entry upfirdn2d.cpp:20)
#include "common.cuh"
#include <cuda_bf16.h>

// o[0..3] -> hi = bf16(o), lo = bf16(o - hi), stored as two 8-byte vectors (packed two-at-a-time conversions: F2FP.BF16)
__device__ __forceinline__ void split4(const float* o, uint2* hi, uint2* lo, long i) {
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(o[0], o[1]), h23 = __floats2bfloat162_rn(o[2], o[3]);
    if (hi) hi[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    if (lo) {
        const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
        const __nv_bfloat162 l01 = __floats2bfloat162_rn(o[0] - f01.x, o[1] - f01.y), l23 = __floats2bfloat162_rn(o[2] - f23.x, o[3] - f23.y);
        lo[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
    }
}

// SynthesisLayer activation tail: leaky relu (slope alpha in [0, 1]: max(t, alpha t)), gain, symmetric clamp (clamp < 0: none)
__device__ __forceinline__ float lrelu_gain_clamp(float t, bool lrelu01, int lrelu, float alpha, float gain, float clamp) {
    if (lrelu01) t = fmaxf(t, t * alpha);
    else if (lrelu) t = t > 0.f ? t : t * alpha;
    t *= gain;
    if (clamp >= 0.f) t = fminf(fmaxf(t, -clamp), clamp);
    return t;
}

// ---------------------------------------------------------------------------------------------
// Generic bias_act (mapping network, odd activations; the synthesis layers use the fused epilogues below).
// act codes follow the reference table (bias_act.py:20-30): 1 linear, 2 relu, 3 lrelu, 4 tanh, 5 sigmoid, 6 elu, 7 selu,
// 8 softplus, 9 swish.  Each activation is a small functor:
//     value(u)          the activation of the biased input u
//     slope(u_ref, a)   its derivative, expressed through what the forward SAVED for it -- the un-gained output a = y / gain for
//                       every activation that can be inverted from its output, the biased input u_ref for swish (the only one
//                       that cannot; bias_act.py:20-30 column 'ref')
// and the kernel is instantiated per (activation, direction), so the element loop carries no switch:
//     forward   y  = clamp(value(x + b) * gain)
//     gradient  dx = dy * slope * gain, zeroed where the saved output sits on the clamp
namespace bact {
constexpr float kBig = 80.f;            // exp() saturation guard of the reference kernel (bias_act.cu:92-131): keeps 1/(1+e^-u) etc. finite
struct Linear   { float alpha; __device__ float value(float u) const { return u; }
                  __device__ float slope(float, float) const { return 1.f; } };
struct Relu     { float alpha; __device__ float value(float u) const { return fmaxf(u, 0.f); }
                  __device__ float slope(float, float a) const { return a > 0.f ? 1.f : 0.f; } };
struct LRelu    { float alpha; __device__ float value(float u) const { return u > 0.f ? u : u * alpha; }
                  __device__ float slope(float, float a) const { return a > 0.f ? 1.f : alpha; } };
struct Tanh     { float alpha; __device__ float value(float u) const { return fabsf(u) > kBig ? copysignf(1.f, u) : tanhf(u); }
                  __device__ float slope(float, float a) const { return 1.f - a * a; } };
struct Sigmoid  { float alpha; __device__ float value(float u) const { return u < -kBig ? 0.f : 1.f / (1.f + expf(-u)); }
                  __device__ float slope(float, float a) const { return a * (1.f - a); } };
struct Elu      { float alpha; __device__ float value(float u) const { return u >= 0.f ? u : expm1f(u); }
                  __device__ float slope(float, float a) const { return a >= 0.f ? 1.f : a + 1.f; } };
struct Selu     { float alpha; static constexpr float S = 1.0507009873554804934f, A = 1.6732632423543772848f;
                  __device__ float value(float u) const { return u >= 0.f ? S * u : S * A * expm1f(u); }
                  __device__ float slope(float, float a) const { return a >= 0.f ? S : a + S * A; } };
struct Softplus { float alpha; __device__ float value(float u) const { return u > kBig ? u : log1pf(expf(u)); }
                  __device__ float slope(float, float a) const { return -expm1f(-a); } };           // 1 - e^-softplus = sigmoid(u)
struct Swish    { float alpha; __device__ float value(float u) const { return u < -kBig ? 0.f : u / (1.f + expf(-u)); }
                  __device__ float slope(float u, float) const {                                     // s + u s (1 - s), s = sigmoid(u)
                      if (u > 0.5f * kBig) return 1.f;
                      const float s = u < -kBig ? 0.f : 1.f / (1.f + expf(-u));
                      return s * (1.f + u * (1.f - s)); } };

template <typename Act, bool GRAD>
__global__ void bias_act_generic_kernel(const float* __restrict__ x, const float* __restrict__ b, const float* __restrict__ xref,
                                        const float* __restrict__ yref, const float* __restrict__ dy, float* __restrict__ y, long total,
                                        long stepB, int sizeB, Act act, float gain, float clamp) {
    const float inv_gain = gain != 0.f ? 1.f / gain : 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const float bias = b ? b[(i / stepB) % sizeB] : 0.f;
        const float scale = gain * (dy ? dy[i] : 1.f);
        float out;
        if (!GRAD) {
            out = act.value(x[i] + bias) * scale;
            if (clamp >= 0.f) out = fminf(fmaxf(out, -clamp), clamp);
        } else {
            const float u_ref = (xref ? xref[i] : 0.f) + bias;                       // only swish reads it
            float y_saved = yref ? yref[i] : 0.f;
            if (!yref && xref) y_saved = act.value(u_ref) * gain;                    // swish: the clamp test needs the output, rebuilt from the input
            out = x[i] * act.slope(u_ref, y_saved * inv_gain) * scale;
            if (clamp >= 0.f && !(y_saved > -clamp && y_saved < clamp)) out = 0.f;
        }
        y[i] = out;
    }
}

template <typename Act>
static int launch_bias_act(bool grad, const float* x, const float* b, const float* xref, const float* yref, const float* dy, float* y, long total,
                           long stepB, int sizeB, float alpha, float gain, float clamp, cudaStream_t st) {
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    const Act act{alpha};
    if (grad) bias_act_generic_kernel<Act, true><<<blocks, 256, 0, st>>>(x, b, xref, yref, dy, y, total, stepB, sizeB, act, gain, clamp);
    else bias_act_generic_kernel<Act, false><<<blocks, 256, 0, st>>>(x, b, xref, yref, dy, y, total, stepB, sizeB, act, gain, clamp);
    B200_CHECK_LAUNCH();
    return 0;
}
}  // namespace bact

// Fast path of the two hot uses (ToRGB bias + clamp and its gradient): channel-last bias (stepB == 1), linear / lrelu,
// 4 elements per thread, 32-bit index math.
__global__ void bias_act_nhwc4_kernel(const float4* __restrict__ x, const float* __restrict__ b, const float4* __restrict__ yref,
                                      float4* __restrict__ y, int grad, int n4, int c4, int lrelu, float alpha, float gain, float clamp) {
    pdl_trigger();
    pdl_wait();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        const float4 xv = x[i];
        float v[4] = {xv.x, xv.y, xv.z, xv.w};
        float yr[4] = {0.f, 0.f, 0.f, 0.f};
        if (yref) { const float4 t = yref[i]; yr[0] = t.x; yr[1] = t.y; yr[2] = t.z; yr[3] = t.w; }
        if (grad == 0 && b) {
            const float4 bv = *reinterpret_cast<const float4*>(b + (i % c4) * 4);
            v[0] += bv.x; v[1] += bv.y; v[2] += bv.z; v[3] += bv.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float o = v[j];
            if (lrelu) o = (grad == 0 ? (o > 0.f) : (yr[j] > 0.f)) ? o : o * alpha;      // sign(y/gain) == sign(y) for gain > 0
            o *= gain;
            if (clamp >= 0.f) {
                if (grad == 0) o = (o > -clamp && o < clamp) ? o : (o >= 0.f ? clamp : -clamp);
                else o = (yr[j] > -clamp && yr[j] < clamp) ? o : 0.f;
            }
            v[j] = o;
        }
        y[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

B200_API int b200_bias_act(const float* x, const float* b, const float* xref, const float* yref, const float* dy, float* y,
                           int grad, long sizeX, long stepB, int sizeB, int act, float alpha, float gain, float clamp,
                           void* stream) {
    B200_REQUIRE(grad == 0 || grad == 1, "bias_act: only forward (0) and first-order gradient (1) are implemented");
    B200_REQUIRE(act >= 1 && act <= 9, "bias_act: unknown activation code");
    B200_REQUIRE(!b || (stepB > 0 && sizeB > 0), "bias_act: bad bias indexing");
    if (sizeX <= 0) return 0;
    if ((act == 1 || act == 3) && !xref && !dy && gain > 0.f && sizeX % 4 == 0 && sizeX < (1L << 31) &&
        (!b || (stepB == 1 && sizeB % 4 == 0 && sizeX % sizeB == 0)) && (grad == 0 || act == 1 || yref) &&
        ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && (!yref || (uintptr_t)yref % 16 == 0) && (!b || (uintptr_t)b % 16 == 0)) {
        const int n4 = (int)(sizeX / 4);
        const int blocks4 = (n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16;
        B200_CUDA(launch_pdl(bias_act_nhwc4_kernel, dim3(blocks4), dim3(256), 0, (cudaStream_t)stream, (const float4*)x, grad == 0 ? b : (const float*)nullptr,
                             (const float4*)yref, (float4*)y, grad, n4, (int)(b ? sizeB / 4 : 1), (int)(act == 3), alpha, gain, clamp));
        return 0;
    }
    const long sb = stepB > 0 ? stepB : 1;
    const int nb = sizeB > 0 ? sizeB : 1;
    cudaStream_t st = (cudaStream_t)stream;
    using namespace bact;
    switch (act) {
        case 1: return launch_bias_act<Linear>(grad, x, b, xref, yref, dy, y, sizeX, sb, nb, alpha, gain, clamp, st);
        case 2: return launch_bias_act<Relu>(grad, x, b, xref, yref, dy, y, sizeX, sb, nb, alpha, gain, clamp, st);
        case 3: return launch_bias_act<LRelu>(grad, x, b, xref, yref, dy, y, sizeX, sb, nb, alpha, gain, clamp, st);
        case 4: return launch_bias_act<Tanh>(grad, x, b, xref, yref, dy, y, sizeX, sb, nb, alpha, gain, clamp, st);
        case 5: return launch_bias_act<Sigmoid>(grad, x, b, xref, yref, dy, y, sizeX, sb, nb, alpha, gain, clamp, st);
        case 6: return launch_bias_act<Elu>(grad, x, b, xref, yref, dy, y, sizeX, sb, nb, alpha, gain, clamp, st);
        case 7: return launch_bias_act<Selu>(grad, x, b, xref, yref, dy, y, sizeX, sb, nb, alpha, gain, clamp, st);
        case 8: return launch_bias_act<Softplus>(grad, x, b, xref, yref, dy, y, sizeX, sb, nb, alpha, gain, clamp, st);
        default: return launch_bias_act<Swish>(grad, x, b, xref, yref, dy, y, sizeX, sb, nb, alpha, gain, clamp, st);
    }
}

// ---------------------------------------------------------------------------------------------
// SynthesisLayer epilogue on NHWC [n][hw][c]:  z = clamp(act(y + noise[pix]*strength + bias[c]) * gain, +-clamp)
// noise may be null; noise_bs = per-sample stride of the noise map (0 for the shared 'const' buffer).

template <typename idx_t>
__global__ void layer_act_fwd_kernel(const float4* __restrict__ y, float4* __restrict__ z, uint2* __restrict__ zhi,
                                     uint2* __restrict__ zlo, const float* __restrict__ bias,
                                     const float* __restrict__ noise, const float* __restrict__ strength, long noise_bs,
                                     long total4_, int hw_, int c4_, int lrelu, float alpha, float gain, float clamp) {
    pdl_trigger();
    pdl_wait();
    const float str = (noise && strength) ? *strength : 0.f;
    const idx_t total4 = (idx_t)total4_, c4 = (idx_t)c4_, hw = (idx_t)hw_;
    for (idx_t i = (idx_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (idx_t)gridDim.x * blockDim.x) {
        const int cc = (int)(i % c4);
        const idx_t pix = i / c4;
        float4 v = y[i];
        float add = 0.f;
        if (noise) add = noise[(pix / hw) * noise_bs + pix % hw] * str;
        float4 bv = bias ? *reinterpret_cast<const float4*>(bias + cc * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        float o[4] = {v.x + add + bv.x, v.y + add + bv.y, v.z + add + bv.z, v.w + add + bv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float t = o[j];
            if (lrelu) t = t > 0.f ? t : t * alpha;
            t *= gain;
            if (clamp >= 0.f) t = (t > -clamp && t < clamp) ? t : (t >= 0.f ? clamp : -clamp);
            o[j] = t;
        }
        if (z) z[i] = make_float4(o[0], o[1], o[2], o[3]);
        if (zhi) split4(o, zhi, zlo, i);
    }
}

// scalar fallback for channel counts that are not a multiple of 4
__global__ void layer_act_fwd_kernel_s(const float* __restrict__ y, float* __restrict__ z, const float* __restrict__ bias,
                                       const float* __restrict__ noise, const float* __restrict__ strength, long noise_bs,
                                       long total, int hw, int c, int lrelu, float alpha, float gain, float clamp) {
    const float str = (noise && strength) ? *strength : 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int cc = (int)(i % c);
        const long pix = i / c;
        float t = y[i] + (bias ? bias[cc] : 0.f);
        if (noise) t += noise[(pix / hw) * noise_bs + pix % hw] * str;
        if (lrelu) t = t > 0.f ? t : t * alpha;
        t *= gain;
        if (clamp >= 0.f) t = (t > -clamp && t < clamp) ? t : (t >= 0.f ? clamp : -clamp);
        z[i] = t;
    }
}

B200_API int b200_layer_act_fwd(const float* y, float* z, void* z_hi, void* z_lo, const float* bias, const float* noise,
                                const float* strength, long noise_bs, int n, int hw, int c, int lrelu, float alpha, float gain,
                                float clamp, void* stream) {
    const long total = (long)n * hw * c;
    if (total <= 0) return 0;
    B200_REQUIRE(z || z_hi, "layer_act_fwd: no output requested");
    B200_REQUIRE(!z_hi || c % 4 == 0, "layer_act_fwd: bf16 outputs need a channel count that is a multiple of 4");
    cudaStream_t st = (cudaStream_t)stream;
    if (c % 4 == 0) {
        const long t4 = total / 4;
        const int blocks = (int)((t4 + 255) / 256 < 148 * 16 ? (t4 + 255) / 256 : 148 * 16);
        if (t4 < (1L << 31))
            B200_CUDA(launch_pdl(layer_act_fwd_kernel<unsigned>, dim3(blocks), dim3(256), 0, st, (const float4*)y, (float4*)z, (uint2*)z_hi, (uint2*)z_lo,
                                 bias, noise, strength, noise_bs, t4, hw, c / 4, lrelu, alpha, gain, clamp));
        else
            B200_CUDA(launch_pdl(layer_act_fwd_kernel<long>, dim3(blocks), dim3(256), 0, st, (const float4*)y, (float4*)z, (uint2*)z_hi, (uint2*)z_lo,
                                 bias, noise, strength, noise_bs, t4, hw, c / 4, lrelu, alpha, gain, clamp));
    } else {
        const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
        layer_act_fwd_kernel_s<<<blocks, 256, 0, st>>>(y, z, bias, noise, strength, noise_bs, total, hw, c, lrelu, alpha,
                                                       gain, clamp);
    }
    B200_CHECK_LAUNCH();
    return 0;
}

// Backward: one warp per pixel.  dy = dz * gain * slope(z) * [|z| < clamp]   (uses the saved OUTPUT z, bias_act.cu:73-77,143-145)
// Also reduces dbias[c] += sum_pix dy, dstrength += sum dy*noise, dnoise[pix] += strength * sum_c dy.
template <int R>
__global__ void __launch_bounds__(256) layer_act_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ z,
                                                            const __nv_bfloat16* __restrict__ zhi, const __nv_bfloat16* __restrict__ zlo,
                                                            float* __restrict__ dy, __nv_bfloat16* __restrict__ dyhi,
                                                            __nv_bfloat16* __restrict__ dylo, float* __restrict__ dbias,
                                                            const float* __restrict__ noise, const float* __restrict__ strength,
                                                            long noise_bs, float* __restrict__ dstrength,
                                                            float* __restrict__ dnoise, long npix, int hw, int c, int lrelu,
                                                            float alpha, float gain, float clamp) {
    __shared__ float s_col[R * 32];
    __shared__ float s_str[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < R * 32; i += blockDim.x) s_col[i] = 0.f;
    __syncthreads();
    float col[R];
#pragma unroll
    for (int j = 0; j < R; ++j) col[j] = 0.f;
    float sacc = 0.f;
    const float str = (noise && strength) ? *strength : 0.f;
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    for (long pix = (long)blockIdx.x * (blockDim.x >> 5) + wid; pix < npix; pix += nwarps) {
        const float* dzr = dz + pix * c;
        float row = 0.f;
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int ch = lane + 32 * j;
            if (ch < c) {
                const float zz = z ? z[pix * c + ch] : __bfloat162float(zhi[pix * c + ch]) + (zlo ? __bfloat162float(zlo[pix * c + ch]) : 0.f);
                float g = dzr[ch] * gain;
                if (lrelu && !(zz > 0.f)) g *= alpha;
                if (clamp >= 0.f && !(zz > -clamp && zz < clamp)) g = 0.f;
                if (dy) dy[pix * c + ch] = g;
                if (dyhi) {
                    const __nv_bfloat16 hh = __float2bfloat16_rn(g);
                    dyhi[pix * c + ch] = hh;
                    if (dylo) dylo[pix * c + ch] = __float2bfloat16_rn(g - __bfloat162float(hh));
                }
                col[j] += g;
                row += g;
            }
        }
        if (noise) {
            row = warp_sum(row);
            if (lane == 0) {
                const long nidx = (pix / hw) * noise_bs + pix % hw;
                sacc += row * noise[nidx];
                if (dnoise) atomicAdd(dnoise + nidx, row * str);
            }
        }
    }
    if (dbias) {
#pragma unroll
        for (int j = 0; j < R; ++j) if (lane + 32 * j < c) atomicAdd(&s_col[lane + 32 * j], col[j]);
        __syncthreads();
        for (int i = threadIdx.x; i < c; i += blockDim.x) atomicAdd(dbias + i, s_col[i]);
    }
    if (noise && dstrength) {
        if (lane == 0) s_str[wid] = sacc;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_str[i];
            atomicAdd(dstrength, t);
        }
    }
}

// Vectorised variant for c % 4 == 0 and c <= 512: a group of G = c/4 (<= 128) threads owns one pixel, 4 channels per thread.
// Per-thread column sums are kept in registers across the grid-stride loop (a thread always sees the same 4 channels).
__global__ void __launch_bounds__(256) layer_act_bwd_vec_kernel(const float4* __restrict__ dz, const float4* __restrict__ z,
                                                                const uint2* __restrict__ zhi, const uint2* __restrict__ zlo,
                                                                float4* __restrict__ dy, uint2* __restrict__ dyhi,
                                                                uint2* __restrict__ dylo, float* __restrict__ dbias,
                                                                const float* __restrict__ noise, const float* __restrict__ strength,
                                                                long noise_bs, float* __restrict__ dstrength,
                                                                float* __restrict__ dnoise, long npix, int hw, int c4, int lrelu,
                                                                float alpha, float gain, float clamp) {
    __shared__ float s_col[512];
    __shared__ float s_row[256];
    __shared__ float s_str;
    pdl_trigger();
    pdl_wait();
    const int ppb = blockDim.x / c4;                 // pixels per block iteration
    const int sub = threadIdx.x / c4, cc = threadIdx.x % c4;
    const bool active = sub < ppb;
    const bool shfl_rows = c4 <= 32 && (c4 & (c4 - 1)) == 0;        // a pixel's c4 threads sit in one warp (every thread is active then)
    for (int i = threadIdx.x; i < 4 * c4; i += blockDim.x) s_col[i] = 0.f;
    if (threadIdx.x == 0) s_str = 0.f;
    __syncthreads();
    const float str = (noise && strength) ? *strength : 0.f;
    float col[4] = {0.f, 0.f, 0.f, 0.f};
    float sacc = 0.f;
    for (long p0 = (long)blockIdx.x * ppb; p0 < npix; p0 += (long)gridDim.x * ppb) {
        const long pix = p0 + sub;
        float row = 0.f;
        if (active && pix < npix) {
            const long i = pix * c4 + cc;
            const float4 dd = dz[i];
            float g[4] = {dd.x * gain, dd.y * gain, dd.z * gain, dd.w * gain};
            float zv[4];
            if (z) {
                const float4 zz = z[i];
                zv[0] = zz.x; zv[1] = zz.y; zv[2] = zz.z; zv[3] = zz.w;
            } else {
                // reference output from its split-bf16 copy: hi alone decides the sign (bf16 rounding keeps sign and zero);
                // the clamp test needs hi + lo only where hi itself reaches the clamp (a value just below it rounds up to it)
                const uint2 h2 = zhi[i];
                const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(&h2);
                bool edge = false;
#pragma unroll
                for (int j = 0; j < 4; ++j) { zv[j] = __bfloat162float(hb[j]); edge |= clamp >= 0.f && !(zv[j] > -clamp && zv[j] < clamp); }
                if (edge && zlo) {
                    const uint2 l2 = zlo[i];
                    const __nv_bfloat16* lb = reinterpret_cast<const __nv_bfloat16*>(&l2);
#pragma unroll
                    for (int j = 0; j < 4; ++j) zv[j] += __bfloat162float(lb[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (lrelu && !(zv[j] > 0.f)) g[j] *= alpha;
                if (clamp >= 0.f && !(zv[j] > -clamp && zv[j] < clamp)) g[j] = 0.f;
                col[j] += g[j];
                row += g[j];
            }
            if (dy) dy[i] = make_float4(g[0], g[1], g[2], g[3]);
            if (dyhi) split4(g, dyhi, dylo, i);
        }
        if (noise && shfl_rows) {                     // per-pixel sum over channels -> d noise, d strength: the pixel's threads are lanes of one warp
            float r = row;
            for (int o = c4 >> 1; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
            if (cc == 0 && pix < npix) {
                const long nidx = (pix / hw) * noise_bs + pix % hw;
                sacc += r * noise[nidx];
                if (dnoise) atomicAdd(dnoise + nidx, r * str);
            }
        } else if (noise) {
            s_row[threadIdx.x] = row;
            __syncthreads();
            if (active && cc == 0 && pix < npix) {
                float r = 0.f;
                for (int k = 0; k < c4; ++k) r += s_row[sub * c4 + k];
                const long nidx = (pix / hw) * noise_bs + pix % hw;
                sacc += r * noise[nidx];
                if (dnoise) atomicAdd(dnoise + nidx, r * str);
            }
            __syncthreads();
        }
    }
    if (dbias && active) {
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(&s_col[cc * 4 + j], col[j]);
    }
    if (noise && dstrength && sacc != 0.f) atomicAdd(&s_str, sacc);
    __syncthreads();
    if (dbias) for (int i = threadIdx.x; i < 4 * c4; i += blockDim.x) atomicAdd(dbias + i, s_col[i]);
    if (noise && dstrength && threadIdx.x == 0) atomicAdd(dstrength, s_str);
}

// The same backward, specialised at compile time on where the saved output comes from (fp32 / split pair), where dy goes (fp32 /
// split pair), whether there is a noise input and how many 4-channel vectors a thread owns (VPT: a pixel's c/4/VPT threads are lanes
// of one warp, so its channel sum is a few shuffles): 32-bit indices, no run-time flag tests in the element loop.
// (ncu on the generic kernel: 43 instructions per element, a third of them flag tests and 64-bit index math.)
template <bool ZSPLIT, bool OSPLIT, bool NOISE, int VPT>
__global__ void __launch_bounds__(256) layer_act_bwd_fast_kernel(const float4* __restrict__ dz, const float4* __restrict__ dz2, const float4* __restrict__ z,
                                                                 const uint2* __restrict__ zhi, const uint2* __restrict__ zlo,
                                                                 float4* __restrict__ dy, uint2* __restrict__ dyhi, uint2* __restrict__ dylo,
                                                                 float* __restrict__ dbias, const float* __restrict__ noise,
                                                                 const float* __restrict__ strength, unsigned noise_bs,
                                                                 float* __restrict__ dstrength, float* __restrict__ dnoise, unsigned npix,
                                                                 unsigned hw, int c4, float gain_neg, float gain, float clamp) {
    __shared__ float s_col[512];
    __shared__ float s_str;
    pdl_trigger();
    pdl_wait();
    const unsigned L = (unsigned)c4 / VPT;                 // threads per pixel
    const unsigned ppb = 256u / L, sub = threadIdx.x / L, cc = threadIdx.x % L;
    const bool active = sub < ppb;
    for (int i = threadIdx.x; i < 4 * c4; i += 256) s_col[i] = 0.f;
    if (threadIdx.x == 0) s_str = 0.f;
    __syncthreads();
    const float str = NOISE ? *strength : 0.f;
    const bool clamp_on = clamp >= 0.f;
    float col[VPT][4];
#pragma unroll
    for (int k = 0; k < VPT; ++k) col[k][0] = col[k][1] = col[k][2] = col[k][3] = 0.f;
    float sacc = 0.f;
    for (unsigned p0 = blockIdx.x * ppb; p0 < npix; p0 += gridDim.x * ppb) {
        const unsigned pix = p0 + sub;
        const bool ok = active && pix < npix;
        float r = 0.f;
        if (ok) {
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const unsigned i = pix * (unsigned)c4 + cc + k * L;
                float4 dd = __ldg(dz + i);
                if (dz2) {                            // second consumer's gradient (b200_layer_act_bwd_sum2): summed here, not in a pass of its own
                    const float4 ee = __ldg(dz2 + i);
                    dd.x += ee.x; dd.y += ee.y; dd.z += ee.z; dd.w += ee.w;
                }
                float zv[4];
                if (ZSPLIT) {
                    const uint2 h2 = __ldg(zhi + i);
                    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h2.x)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h2.y));
                    zv[0] = a.x; zv[1] = a.y; zv[2] = b.x; zv[3] = b.y;
                    // hi alone decides the sign; the clamp test needs hi + lo only where hi itself reaches the clamp
                    if (clamp_on && fmaxf(fmaxf(fabsf(zv[0]), fabsf(zv[1])), fmaxf(fabsf(zv[2]), fabsf(zv[3]))) >= clamp) {
                        const uint2 l2 = __ldg(zlo + i);
                        const float2 la = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&l2.x)), lb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&l2.y));
                        zv[0] += la.x; zv[1] += la.y; zv[2] += lb.x; zv[3] += lb.y;
                    }
                } else {
                    const float4 zz = __ldg(z + i);
                    zv[0] = zz.x; zv[1] = zz.y; zv[2] = zz.z; zv[3] = zz.w;
                }
                const float d4[4] = {dd.x, dd.y, dd.z, dd.w};
                float g[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float t = d4[j] * (zv[j] > 0.f ? gain : gain_neg);
                    if (clamp_on) t = fabsf(zv[j]) < clamp ? t : 0.f;
                    g[j] = t;
                    col[k][j] += t;
                }
                if (NOISE) r += (g[0] + g[1]) + (g[2] + g[3]);
                if (OSPLIT) split4(g, dyhi, dylo, i);
                else dy[i] = make_float4(g[0], g[1], g[2], g[3]);
            }
        }
        if (NOISE) {                                  // per-pixel sum over channels -> d noise, d strength (L is a power of two <= 32 here)
            for (unsigned o = L >> 1; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
            if (cc == 0 && ok) {
                const unsigned q = pix / hw, nidx = q * noise_bs + (pix - q * hw);
                sacc = fmaf(r, __ldg(noise + nidx), sacc);
                if (dnoise) atomicAdd(dnoise + nidx, r * str);
            }
        }
    }
    if (dbias && active) {
#pragma unroll
        for (int k = 0; k < VPT; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(&s_col[(cc + k * L) * 4 + j], col[k][j]);
    }
    if (NOISE && dstrength && sacc != 0.f) atomicAdd(&s_str, sacc);
    __syncthreads();
    if (dbias) for (int i = threadIdx.x; i < 4 * c4; i += 256) atomicAdd(dbias + i, s_col[i]);
    if (NOISE && dstrength && threadIdx.x == 0) atomicAdd(dstrength, s_str);
}

// c <= 4 channels (the ToRGB images), no noise input, fp32 in / out: one thread per pixel, the warp reads 32 * c consecutive floats;
// d bias reduced per warp, then per block, one atomic per channel and block.  (The warp-per-pixel kernel above ran 3 of 32 lanes.)
__global__ void __launch_bounds__(256) layer_act_bwd_thin_kernel(const float* __restrict__ dz, const float* __restrict__ z,
                                                                 float* __restrict__ dy, float* __restrict__ dbias, long npix, int c,
                                                                 int lrelu, float alpha, float gain, float clamp) {
    __shared__ float s_col[4];
    if (threadIdx.x < 4) s_col[threadIdx.x] = 0.f;
    __syncthreads();
    float col[4] = {0.f, 0.f, 0.f, 0.f};
    for (long pix = (long)blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += (long)gridDim.x * blockDim.x) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            if (ch < c) {
                const float zz = z[pix * c + ch];
                float g = dz[pix * c + ch] * gain;
                if (lrelu && !(zz > 0.f)) g *= alpha;
                if (clamp >= 0.f && !(zz > -clamp && zz < clamp)) g = 0.f;
                dy[pix * c + ch] = g;
                col[ch] += g;
            }
        }
    }
    if (dbias) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            const float v = warp_sum(col[ch]);
            if ((threadIdx.x & 31) == 0 && ch < c) atomicAdd(&s_col[ch], v);
        }
        __syncthreads();
        if ((int)threadIdx.x < c) atomicAdd(dbias + threadIdx.x, s_col[threadIdx.x]);
    }
}

// Shapes the compile-time specialised kernel takes: channel vectors of a pixel inside one warp (a power-of-two count when the noise
// gradient needs the shuffle sum), 32-bit element indices.
static bool act_bwd_fast_shape(long npix, int c, bool has_noise, long n, long noise_bs) {
    if (c <= 0 || c % 4 != 0 || c > 512) return false;
    const int c4 = c / 4;
    const int vpt = c4 <= 32 ? 1 : (c4 <= 64 ? 2 : 4);
    const int lanes = c4 / vpt;
    const bool pow2 = (lanes & (lanes - 1)) == 0;
    return c4 % vpt == 0 && lanes <= 32 && (pow2 || !has_noise) && npix * c4 < (1L << 31) && n * noise_bs < (1L << 31);
}

static int layer_act_bwd_impl(const float* dz, const float* dz2, const float* z, const void* z_hi, const void* z_lo, float* dy, void* dy_hi, void* dy_lo,
                              float* dbias, const float* noise, const float* strength, long noise_bs, float* dstrength, float* dnoise,
                              int n, int hw, int c, int lrelu, float alpha, float gain, float clamp, void* stream) {
    // dbias / dstrength / dnoise are ACCUMULATED into (callers zero them); any of them may be null.
    // The saved output comes as fp32 z, or (z == NULL) as its split-bf16 copy z_hi (+ z_lo, needed for the clamp test only).
    const long npix = (long)n * hw;
    if (npix <= 0 || c <= 0) return 0;
    B200_REQUIRE(c <= 512, "layer_act_bwd: at most 512 channels");
    B200_REQUIRE(z || z_hi, "layer_act_bwd: the saved output is needed as z or as z_hi / z_lo");
    B200_REQUIRE(z || z_lo || clamp < 0.f, "layer_act_bwd: the clamp test on a split output needs z_lo");
    cudaStream_t st = (cudaStream_t)stream;
    if (c % 4 == 0) {
        const int c4 = c / 4, ppb = 256 / c4;
        const long nb = (npix + ppb - 1) / ppb;
        const int blocks = (int)(nb < 148 * 8 ? nb : 148 * 8);
        const bool osplit = dy_hi && dy_lo && !dy, ofp32 = dy && !dy_hi;
        const bool zsplit = !z;
        // vectors per thread: keep a pixel's threads inside one warp (needed for the shuffle sum when there is a noise input)
        const int vpt = c4 <= 32 ? 1 : (c4 <= 64 ? 2 : 4);
        const int lanes = c4 / vpt;
        if (act_bwd_fast_shape(npix, c, noise != nullptr, n, noise_bs) && (osplit || ofp32) && (!noise || strength) &&
            (zsplit ? (z_lo != nullptr || clamp < 0.f) : true)) {
            const float gneg = lrelu ? gain * alpha : gain;
            const int ppb2 = 256 / lanes;
            const long nb2 = (npix + ppb2 - 1) / ppb2;
            const int blocks2 = (int)(nb2 < 148 * 8 ? nb2 : 148 * 8);
#define LAUNCH_F(ZS, OS, NZ, V) B200_CUDA(launch_pdl(layer_act_bwd_fast_kernel<ZS, OS, NZ, V>, dim3(blocks2), dim3(256), 0, st, (const float4*)dz, \
                (const float4*)dz2, (const float4*)z, (const uint2*)z_hi, (const uint2*)z_lo, (float4*)dy, (uint2*)dy_hi, (uint2*)dy_lo, dbias, noise, strength, \
                (unsigned)noise_bs, dstrength, dnoise, (unsigned)npix, (unsigned)hw, c4, gneg, gain, clamp))
#define LAUNCH_V(ZS, OS, NZ) do { if (vpt == 1) LAUNCH_F(ZS, OS, NZ, 1); else if (vpt == 2) LAUNCH_F(ZS, OS, NZ, 2); else LAUNCH_F(ZS, OS, NZ, 4); } while (0)
            if (zsplit) { if (osplit) { if (noise) LAUNCH_V(true, true, true); else LAUNCH_V(true, true, false); }
                          else { if (noise) LAUNCH_V(true, false, true); else LAUNCH_V(true, false, false); } }
            else { if (osplit) { if (noise) LAUNCH_V(false, true, true); else LAUNCH_V(false, true, false); }
                   else { if (noise) LAUNCH_V(false, false, true); else LAUNCH_V(false, false, false); } }
#undef LAUNCH_V
#undef LAUNCH_F
            return 0;
        }
        B200_REQUIRE(!dz2, "layer_act_bwd_sum2: this shape takes the generic kernels, which read one gradient (b200_layer_act_bwd_sum2_supported)");
        B200_CUDA(launch_pdl(layer_act_bwd_vec_kernel, dim3(blocks), dim3(256), 0, st, (const float4*)dz, (const float4*)z, (const uint2*)z_hi,
                             (const uint2*)z_lo, (float4*)dy, (uint2*)dy_hi,
                             (uint2*)dy_lo, dbias, noise, strength, noise_bs, dstrength, dnoise, npix, hw, c4, lrelu, alpha, gain, clamp));
        return 0;
    }
    B200_REQUIRE(!dy_hi, "layer_act_bwd: bf16 outputs need a channel count that is a multiple of 4");
    B200_REQUIRE(!dz2, "layer_act_bwd_sum2: the channel count must be a multiple of 4 (b200_layer_act_bwd_sum2_supported)");
    if (c <= 4 && z && dy && !noise) {        // the 3-channel ToRGB outputs: a thread per pixel instead of a warp per pixel
        const long nb = (npix + 255) / 256;
        layer_act_bwd_thin_kernel<<<(int)(nb < 148 * 8 ? nb : 148 * 8), 256, 0, st>>>(dz, z, dy, dbias, npix, c, lrelu, alpha, gain, clamp);
        B200_CHECK_LAUNCH();
        return 0;
    }
    const int blocks = (int)((npix + 7) / 8 < 148 * 8 ? (npix + 7) / 8 : 148 * 8);
#define LAUNCH_R(R) layer_act_bwd_kernel<R><<<blocks, 256, 0, st>>>(dz, z, (const __nv_bfloat16*)z_hi, (const __nv_bfloat16*)z_lo, dy, (__nv_bfloat16*)dy_hi, (__nv_bfloat16*)dy_lo, dbias, \
                                                                     noise, strength, noise_bs, dstrength, dnoise, npix, hw, c, lrelu, \
                                                                     alpha, gain, clamp)
    if (c <= 32) LAUNCH_R(1); else if (c <= 64) LAUNCH_R(2); else if (c <= 128) LAUNCH_R(4);
    else if (c <= 256) LAUNCH_R(8); else LAUNCH_R(16);
#undef LAUNCH_R
    B200_CHECK_LAUNCH();
    return 0;
}

B200_API int b200_layer_act_bwd(const float* dz, const float* z, const void* z_hi, const void* z_lo, float* dy, void* dy_hi, void* dy_lo,
                                float* dbias, const float* noise, const float* strength, long noise_bs, float* dstrength, float* dnoise,
                                int n, int hw, int c, int lrelu, float alpha, float gain, float clamp, void* stream) {
    return layer_act_bwd_impl(dz, nullptr, z, z_hi, z_lo, dy, dy_hi, dy_lo, dbias, noise, strength, noise_bs, dstrength, dnoise, n, hw, c, lrelu,
                              alpha, gain, clamp, stream);
}

// The same with the incoming gradient given as two addends (an activation with two consumers -- the next convolution and the block's
// ToRGB layer: autograd would sum the two gradients in a pass of its own, 3 x 4 bytes per element; here the second is one more load).
// dz + dz2 is formed in fp32 exactly as that pass would.  Shapes: b200_layer_act_bwd_sum2_supported.
B200_API int b200_layer_act_bwd_sum2(const float* dz, const float* dz2, const float* z, const void* z_hi, const void* z_lo, float* dy, void* dy_hi,
                                     void* dy_lo, float* dbias, const float* noise, const float* strength, long noise_bs, float* dstrength,
                                     float* dnoise, int n, int hw, int c, int lrelu, float alpha, float gain, float clamp, void* stream) {
    B200_REQUIRE(dz && dz2, "layer_act_bwd_sum2: both gradients are required");
    return layer_act_bwd_impl(dz, dz2, z, z_hi, z_lo, dy, dy_hi, dy_lo, dbias, noise, strength, noise_bs, dstrength, dnoise, n, hw, c, lrelu,
                              alpha, gain, clamp, stream);
}

B200_API int b200_layer_act_bwd_sum2_supported(int n, int hw, int c, int has_noise, long noise_bs) {
    return act_bwd_fast_shape((long)n * hw, c, has_noise != 0, n, noise_bs) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// upfirdn2d on NHWC [n][h][w][c]:  out[y,x] = sum_{a,b} g[a,b] * Upad[y*downy + a, x*downx + b]  (+ add[y,x])
// U = zero-inserted input (U[u*up] = in[u]), Upad shifted by (pady0, padx0) (negative pads crop),
// g = gain * (flip ? f : reversed f)   -- the reference correlates with the reversed filter unless flip_filter.

struct UpfirdnParams {
    const float* x; const float* f; const float* add; float* y; uint2* yhi; uint2* ylo;
    int n, h, w, c, oh, ow, fh, fw, upx, upy, downx, downy, padx0, pady0, flip;
    float gain;
    // optional SynthesisLayer epilogue applied to the filtered value (act != 0)
    int act; const float* bias; const float* noise; const float* strength; long noise_bs; int lrelu; float alpha, act_gain, clamp;
    int separable;          // caller's guarantee: f[a][q] = f[a][0] * f[0][q] / f[0][0] (the [1,3,3,1] x [1,3,3,1] resampling filter)
};

// V = channels per thread (4 or 1).  F > 0: square FxF filter with compile-time UP / DOWN (fully unrolled taps);
// F == 0: generic run-time filter size and factors.
template <int V, int F, int UP, int DOWN>
__global__ void upfirdn2d_kernel(UpfirdnParams p) {
    pdl_trigger();
    pdl_wait();
    const unsigned cv = p.c / V;
    const unsigned total = (unsigned)((long)p.n * p.oh * p.ow * cv);      // launch_upfirdn() guarantees < 2^31 outputs
    const int fh = F > 0 ? F : p.fh, fw = F > 0 ? F : p.fw;
    const int upx = F > 0 ? UP : p.upx, upy = F > 0 ? UP : p.upy, downx = F > 0 ? DOWN : p.downx, downy = F > 0 ? DOWN : p.downy;
    float fr[F > 0 ? F * F : 1];
    if (F > 0) {
#pragma unroll
        for (int i = 0; i < F * F; ++i) fr[i] = p.flip ? p.f[i] : p.f[F * F - 1 - i];
    }
    const float str = (p.act && p.noise && p.strength) ? *p.strength : 0.f;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int cc = (int)(i % cv) * V;
        unsigned r = i / cv;
        const int ox = (int)(r % (unsigned)p.ow); r /= (unsigned)p.ow;
        const int oy = (int)(r % (unsigned)p.oh);
        const int b = (int)(r / (unsigned)p.oh);
        float acc[V];
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = 0.f;
        const float* xb = p.x + (long)b * p.h * p.w * p.c + cc;
        auto tap = [&](int a, int q, float g) {
            const int uy = oy * downy + a - p.pady0, ux = ox * downx + q - p.padx0;
            if (uy < 0 || ux < 0 || uy % upy != 0 || ux % upx != 0) return;
            const int iy = uy / upy, ix = ux / upx;
            if (iy >= p.h || ix >= p.w) return;
            const float* src = xb + ((long)iy * p.w + ix) * p.c;
            if (V == 4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(src));
                acc[0] = fmaf(g, v.x, acc[0]); acc[1 % V] = fmaf(g, v.y, acc[1 % V]);
                acc[2 % V] = fmaf(g, v.z, acc[2 % V]); acc[3 % V] = fmaf(g, v.w, acc[3 % V]);
            } else {
                acc[0] = fmaf(g, __ldg(src), acc[0]);
            }
        };
        if (F == 4 && UP == 2 && DOWN == 1) {
            // zero-insertion upsampling: only the taps whose parity matches the output position land on an input sample (2 x 2 of
            // the 16); pick their rows / columns by parity instead of testing all 16
            const int a0 = (p.pady0 - oy) & 1, q0 = (p.padx0 - ox) & 1;
#pragma unroll
            for (int ia = 0; ia < 2; ++ia)
#pragma unroll
                for (int iq = 0; iq < 2; ++iq) {
                    const float ge = q0 ? fr[(2 * ia) * 4 + 2 * iq + 1] : fr[(2 * ia) * 4 + 2 * iq];
                    const float go = q0 ? fr[(2 * ia + 1) * 4 + 2 * iq + 1] : fr[(2 * ia + 1) * 4 + 2 * iq];
                    tap(a0 + 2 * ia, q0 + 2 * iq, a0 ? go : ge);
                }
        } else if (F > 0) {
#pragma unroll
            for (int a = 0; a < F; ++a)
#pragma unroll
                for (int q = 0; q < F; ++q) tap(a, q, fr[a * (F > 0 ? F : 1) + q]);
        } else {
#pragma unroll 1
            for (int a = 0; a < fh; ++a)
#pragma unroll 1
                for (int q = 0; q < fw; ++q) tap(a, q, p.flip ? p.f[a * fw + q] : p.f[(fh - 1 - a) * fw + (fw - 1 - q)]);
        }
        const long opix = ((long)b * p.oh + oy) * p.ow + ox;
        const long o = opix * p.c + cc;
        float out[V];
#pragma unroll
        for (int j = 0; j < V; ++j) out[j] = acc[j] * p.gain;
        if (p.add) {
            if (V == 4) {
                const float4 ad = __ldg(reinterpret_cast<const float4*>(p.add + o));
                out[0] += ad.x; out[1 % V] += ad.y; out[2 % V] += ad.z; out[3 % V] += ad.w;
            } else {
#pragma unroll
                for (int j = 0; j < V; ++j) out[j] += p.add[o + j];
            }
        }
        if (p.act) {
            float nz = 0.f;
            if (p.noise) nz = p.noise[(long)b * p.noise_bs + (long)oy * p.ow + ox] * str;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                float t = out[j] + nz + (p.bias ? p.bias[cc + j] : 0.f);
                if (p.lrelu) t = t > 0.f ? t : t * p.alpha;
                t *= p.act_gain;
                if (p.clamp >= 0.f) t = (t > -p.clamp && t < p.clamp) ? t : (t >= 0.f ? p.clamp : -p.clamp);
                out[j] = t;
            }
        }
        if (V == 4) {
            if (p.y) *reinterpret_cast<float4*>(p.y + o) = make_float4(out[0], out[1 % V], out[2 % V], out[3 % V]);
            if (p.yhi) split4(out, p.yhi, p.ylo, o / 4);
        } else {
            p.y[o] = out[0];
        }
    }
}

// Separable 4x4 FIR at unit rate as a sliding window down a column strip: one thread owns SW output columns x 4 channels and walks
// SH output rows.  Every input row is loaded once (SW + 3 vector loads), filtered horizontally (4 taps) into a ring of the last four
// h-rows, and each output row is the 4-tap vertical combination of the ring: 8 FMA per output element instead of 16 in the 2 x 4
// patch kernel below (ncu: 67 instructions per element there, most of them index arithmetic and the epilogue's selects).
template <int SH, int SW>
__global__ void __launch_bounds__(128) fir4_col_kernel(UpfirdnParams p) {
    pdl_trigger();
    pdl_wait();
    const unsigned cv = p.c >> 2;
    const unsigned sxn = (p.ow + SW - 1) / SW, syn = (p.oh + SH - 1) / SH;
    const unsigned total = (unsigned)p.n * syn * sxn * cv;
    // effective filter (flip, gain) and its separable factors: fr[a][q] = u[a] * v[q]
    float u[4], v[4];
    {
        const float f00 = (p.flip ? p.f[0] : p.f[15]) * p.gain;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] = (p.flip ? p.f[k] : p.f[15 - k]) * p.gain;
            u[k] = (p.flip ? p.f[4 * k] : p.f[15 - 4 * k]) * p.gain / f00;
        }
    }
    const float str = (p.act && p.noise && p.strength) ? *p.strength : 0.f;
    const bool lrelu01 = p.lrelu && p.alpha >= 0.f && p.alpha <= 1.f;
    const int pitch = p.w * p.c;                              // elements per input row (launch_upfirdn: tensors < 2^31 elements)
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int cc = (int)(i % cv) * 4;
        unsigned r = i / cv;
        const int ox0 = (int)(r % sxn) * SW; r /= sxn;
        const int oy0 = (int)(r % syn) * SH;
        const int b = (int)(r / syn);
        const int ixb = ox0 - p.padx0;
        const float* xb = p.x + (long)b * p.h * pitch + cc + (long)ixb * p.c;       // first column of the window (may lie left of the image)
        const bool xin = ixb >= 0 && ixb + SW + 3 <= p.w;                           // every window column inside: no per-load tests
        float bb[4] = {0.f, 0.f, 0.f, 0.f};
        if (p.act && p.bias) { const float4 bv = *reinterpret_cast<const float4*>(p.bias + cc); bb[0] = bv.x; bb[1] = bv.y; bb[2] = bv.z; bb[3] = bv.w; }

        // input row iy -> registers (zero outside the image: upfirdn2d zero padding); issued one row ahead of its use
        auto load_row = [&](int iy, float4 (&row)[SW + 3]) {
            const bool yin = iy >= 0 && iy < p.h;
            const float* rowp = xb + (yin ? iy : 0) * pitch;
            if (yin && xin) {
#pragma unroll
                for (int rx = 0; rx < SW + 3; ++rx) row[rx] = __ldg(reinterpret_cast<const float4*>(rowp + rx * p.c));
            } else {
#pragma unroll
                for (int rx = 0; rx < SW + 3; ++rx) {
                    const int ix = ixb + rx;
                    row[rx] = (yin && ix >= 0 && ix < p.w) ? __ldg(reinterpret_cast<const float4*>(rowp + rx * p.c)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        };
        // horizontal 4-tap pass: row -> h[0..SW)
        auto filt = [&](const float4 (&row)[SW + 3], float4 (&h)[SW]) {
#pragma unroll
            for (int xx = 0; xx < SW; ++xx) {
                h[xx] = make_float4(v[0] * row[xx].x, v[0] * row[xx].y, v[0] * row[xx].z, v[0] * row[xx].w);
#pragma unroll
                for (int q = 1; q < 4; ++q) {
                    h[xx].x = fmaf(v[q], row[xx + q].x, h[xx].x); h[xx].y = fmaf(v[q], row[xx + q].y, h[xx].y);
                    h[xx].z = fmaf(v[q], row[xx + q].z, h[xx].z); h[xx].w = fmaf(v[q], row[xx + q].w, h[xx].w);
                }
            }
        };
        // output row oy = u0 * ha + u1 * hb + u2 * hc + u3 * hd (h-rows of input rows oy - pady0 .. oy - pady0 + 3), then the consumers
        auto emit = [&](int oy, const float4 (&ha)[SW], const float4 (&hb)[SW], const float4 (&hc)[SW], const float4 (&hd)[SW]) {
            if (oy >= p.oh) return;
            const long orow = (((long)b * p.oh + oy) * p.ow + ox0) * p.c + cc;
            const float* nzp = (p.act && p.noise) ? p.noise + (long)b * p.noise_bs + (long)oy * p.ow + ox0 : nullptr;
#pragma unroll
            for (int xx = 0; xx < SW; ++xx) {
                if (ox0 + xx >= p.ow) continue;
                float out[4];
                out[0] = fmaf(u[0], ha[xx].x, fmaf(u[1], hb[xx].x, fmaf(u[2], hc[xx].x, u[3] * hd[xx].x)));
                out[1] = fmaf(u[0], ha[xx].y, fmaf(u[1], hb[xx].y, fmaf(u[2], hc[xx].y, u[3] * hd[xx].y)));
                out[2] = fmaf(u[0], ha[xx].z, fmaf(u[1], hb[xx].z, fmaf(u[2], hc[xx].z, u[3] * hd[xx].z)));
                out[3] = fmaf(u[0], ha[xx].w, fmaf(u[1], hb[xx].w, fmaf(u[2], hc[xx].w, u[3] * hd[xx].w)));
                const long o = orow + xx * p.c;
                if (p.add) { const float4 ad = __ldg(reinterpret_cast<const float4*>(p.add + o)); out[0] += ad.x; out[1] += ad.y; out[2] += ad.z; out[3] += ad.w; }
                if (p.act) {
                    const float nz = nzp ? __ldg(nzp + xx) * str : 0.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j) out[j] = lrelu_gain_clamp(out[j] + (nz + bb[j]), lrelu01, p.lrelu, p.alpha, p.act_gain, p.clamp);
                }
                if (p.y) *reinterpret_cast<float4*>(p.y + o) = make_float4(out[0], out[1], out[2], out[3]);
                if (p.yhi) split4(out, p.yhi, p.ylo, o >> 2);
            }
        };

        float4 h0[SW], h1[SW], h2[SW], h3[SW], ra[SW + 3], rb[SW + 3];
        const int iy0 = oy0 - p.pady0;
        load_row(iy0, ra); load_row(iy0 + 1, rb);
        filt(ra, h0); load_row(iy0 + 2, ra);
        filt(rb, h1); load_row(iy0 + 3, rb);
        filt(ra, h2); load_row(iy0 + 4, ra);
#pragma unroll 1
        for (int g = 0; g < SH; g += 4) {
            if (oy0 + g >= p.oh) break;
            filt(rb, h3); load_row(iy0 + g + 5, rb); emit(oy0 + g, h0, h1, h2, h3);
            filt(ra, h0); load_row(iy0 + g + 6, ra); emit(oy0 + g + 1, h1, h2, h3, h0);
            filt(rb, h1); load_row(iy0 + g + 7, rb); emit(oy0 + g + 2, h2, h3, h0, h1);
            filt(ra, h2); load_row(iy0 + g + 8, ra); emit(oy0 + g + 3, h3, h0, h1, h2);
        }
    }
}

// Register-blocked 4x4 FIR at unit rate (the filter after every up=2 convolution and its adjoint): one thread produces a
// 2 x 4 patch of outputs for 4 channels from a 5 x 7 input window, i.e. 4.4 vector loads per output instead of 16.
__global__ void __launch_bounds__(256) fir4_strip_kernel(UpfirdnParams p) {
    pdl_trigger();
    pdl_wait();
    const unsigned cv = p.c >> 2;
    const unsigned sxn = (p.ow + 3) >> 2, syn = (p.oh + 1) >> 1;
    const unsigned total = (unsigned)p.n * syn * sxn * cv;
    float fr[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) fr[i] = (p.flip ? p.f[i] : p.f[15 - i]) * p.gain;
    const float str = (p.act && p.noise && p.strength) ? *p.strength : 0.f;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int cc = (int)(i % cv) * 4;
        unsigned r = i / cv;
        const int ox0 = (int)(r % sxn) * 4; r /= sxn;
        const int oy0 = (int)(r % syn) * 2;
        const int b = (int)(r / syn);
        float4 acc[2][4];
#pragma unroll
        for (int yy = 0; yy < 2; ++yy)
#pragma unroll
            for (int xx = 0; xx < 4; ++xx) acc[yy][xx] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* xb = p.x + (long)b * p.h * p.w * p.c + cc;
#pragma unroll
        for (int ry = 0; ry < 5; ++ry) {
            const int iy = oy0 + ry - p.pady0;
            if (iy < 0 || iy >= p.h) continue;
            float4 row[7];
#pragma unroll
            for (int rx = 0; rx < 7; ++rx) {
                const int ix = ox0 + rx - p.padx0;
                row[rx] = (ix >= 0 && ix < p.w) ? __ldg(reinterpret_cast<const float4*>(xb + ((long)iy * p.w + ix) * p.c))
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int yy = 0; yy < 2; ++yy) {
                const int a = ry - yy;                      // filter row applied to this input row
                if (a < 0 || a > 3) continue;
#pragma unroll
                for (int xx = 0; xx < 4; ++xx)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {          // packed FP32 FMA (FFMA2): two channels per instruction
                        const float g = fr[a * 4 + q];
                        const float4 v = row[xx + q];
                        fma2(acc[yy][xx].x, acc[yy][xx].y, g, v.x, v.y);
                        fma2(acc[yy][xx].z, acc[yy][xx].w, g, v.z, v.w);
                    }
            }
        }
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.act && p.bias) bv = *reinterpret_cast<const float4*>(p.bias + cc);
#pragma unroll
        for (int yy = 0; yy < 2; ++yy) {
            const int oy = oy0 + yy;
            if (oy >= p.oh) continue;
#pragma unroll
            for (int xx = 0; xx < 4; ++xx) {
                const int ox = ox0 + xx;
                if (ox >= p.ow) continue;
                const long o = (((long)b * p.oh + oy) * p.ow + ox) * p.c + cc;
                float out[4] = {acc[yy][xx].x, acc[yy][xx].y, acc[yy][xx].z, acc[yy][xx].w};
                if (p.add) { const float4 ad = *reinterpret_cast<const float4*>(p.add + o); out[0] += ad.x; out[1] += ad.y; out[2] += ad.z; out[3] += ad.w; }
                if (p.act) {
                    const float nz = p.noise ? p.noise[(long)b * p.noise_bs + (long)oy * p.ow + ox] * str : 0.f;
                    const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float t = out[j] + nz + bb[j];
                        if (p.lrelu) t = t > 0.f ? t : t * p.alpha;
                        t *= p.act_gain;
                        if (p.clamp >= 0.f) t = (t > -p.clamp && t < p.clamp) ? t : (t >= 0.f ? p.clamp : -p.clamp);
                        out[j] = t;
                    }
                }
                if (p.y) *reinterpret_cast<float4*>(p.y + o) = make_float4(out[0], out[1], out[2], out[3]);
                if (p.yhi) split4(out, p.yhi, p.ylo, o / 4);
            }
        }
    }
}

static int launch_upfirdn(UpfirdnParams& p, int padx1, int pady1, cudaStream_t st) {
    B200_REQUIRE(p.upx >= 1 && p.upy >= 1 && p.downx >= 1 && p.downy >= 1, "upfirdn2d: up/down factors must be >= 1");
    B200_REQUIRE(p.fh >= 1 && p.fw >= 1 && p.fh <= 64 && p.fw <= 64, "upfirdn2d: filter size must be in [1, 64]");
    p.oh = (p.h * p.upy + p.pady0 + pady1 - p.fh + p.downy) / p.downy;
    p.ow = (p.w * p.upx + p.padx0 + padx1 - p.fw + p.downx) / p.downx;
    B200_REQUIRE(p.oh >= 1 && p.ow >= 1, "upfirdn2d: output would be empty");
    const bool v4 = (p.c % 4 == 0);
    B200_REQUIRE(v4 || (!p.yhi && p.y), "upfirdn2d: bf16 outputs need a channel count that is a multiple of 4");
    B200_REQUIRE(p.y || p.yhi, "upfirdn2d: no output requested");
    const long total = (long)p.n * p.oh * p.ow * (v4 ? p.c / 4 : p.c);
    if (total <= 0) return 0;
    B200_REQUIRE(total < (1L << 31) && (long)p.n * p.h * p.w * p.c < (1L << 31), "upfirdn2d: tensor too large (>= 2^31 elements)");
    const int blocks = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    const bool sq4 = p.fh == 4 && p.fw == 4 && p.upx == p.upy && p.downx == p.downy;
    if (v4 && sq4 && p.upx == 1 && p.downx == 1) {
        const long strips = (long)p.n * ((p.oh + 1) / 2) * ((p.ow + 3) / 4) * (p.c / 4);
        // The register-blocked kernel needs enough strips to fill the GPU; below that the one-output-per-thread kernel has
        // 8x the parallelism and finishes sooner (these launches are latency-bound, not bandwidth-bound).
#ifndef B200_FIR_STRIP_MIN
#define B200_FIR_STRIP_MIN (148L * 256)
#endif
#ifndef B200_FIR_SW
#define B200_FIR_SW 2
#endif
#ifndef B200_FIR_T16
#define B200_FIR_T16 4096
#endif
        constexpr int SW = B200_FIR_SW;
        if (p.separable && (long)p.n * ((p.oh + 7) / 8) * ((p.ow + SW - 1) / SW) * (p.c / 4) >= 148L * 256) {
            // column strips: 16 rows per thread when that still leaves >= 8 warps per SM, else 8
            const long t16 = (long)p.n * ((p.oh + 15) / 16) * ((p.ow + SW - 1) / SW) * (p.c / 4);
            if (t16 >= 148L * B200_FIR_T16) {
                const int cb = (int)((t16 + 127) / 128 < 148 * 64 ? (t16 + 127) / 128 : 148 * 64);
                B200_CUDA(launch_pdl(fir4_col_kernel<16, SW>, dim3(cb), dim3(128), 0, st, p));
            } else {
                const long t8 = (long)p.n * ((p.oh + 7) / 8) * ((p.ow + SW - 1) / SW) * (p.c / 4);
                const int cb = (int)((t8 + 127) / 128 < 148 * 64 ? (t8 + 127) / 128 : 148 * 64);
                B200_CUDA(launch_pdl(fir4_col_kernel<8, SW>, dim3(cb), dim3(128), 0, st, p));
            }
        } else if (strips >= B200_FIR_STRIP_MIN) {
            const int sb = (int)((strips + 255) / 256 < 148 * 32 ? (strips + 255) / 256 : 148 * 32);
            B200_CUDA(launch_pdl(fir4_strip_kernel, dim3(sb), dim3(256), 0, st, p));
        } else {
            B200_CUDA(launch_pdl(upfirdn2d_kernel<4, 4, 1, 1>, dim3(blocks), dim3(256), 0, st, p));
        }
    }
    else if (v4 && sq4 && p.upx == 2 && p.downx == 1) B200_CUDA(launch_pdl(upfirdn2d_kernel<4, 4, 2, 1>, dim3(blocks), dim3(256), 0, st, p));
    else if (v4 && sq4 && p.upx == 1 && p.downx == 2) B200_CUDA(launch_pdl(upfirdn2d_kernel<4, 4, 1, 2>, dim3(blocks), dim3(256), 0, st, p));
    else if (v4) B200_CUDA(launch_pdl(upfirdn2d_kernel<4, 0, 1, 1>, dim3(blocks), dim3(256), 0, st, p));
    else if (sq4 && p.upx == 1 && p.downx == 1) B200_CUDA(launch_pdl(upfirdn2d_kernel<1, 4, 1, 1>, dim3(blocks), dim3(256), 0, st, p));
    else if (sq4 && p.upx == 2 && p.downx == 1) B200_CUDA(launch_pdl(upfirdn2d_kernel<1, 4, 2, 1>, dim3(blocks), dim3(256), 0, st, p));
    else if (sq4 && p.upx == 1 && p.downx == 2) B200_CUDA(launch_pdl(upfirdn2d_kernel<1, 4, 1, 2>, dim3(blocks), dim3(256), 0, st, p));
    else B200_CUDA(launch_pdl(upfirdn2d_kernel<1, 0, 1, 1>, dim3(blocks), dim3(256), 0, st, p));
    B200_CHECK_LAUNCH();
    return 0;
}

B200_API int b200_upfirdn2d(const float* x, const float* f, const float* add, float* y, int n, int h, int w, int c, int fh,
                            int fw, int upx, int upy, int downx, int downy, int padx0, int padx1, int pady0, int pady1,
                            int flip, float gain, void* stream) {
    UpfirdnParams p{};
    p.x = x; p.f = f; p.add = add; p.y = y; p.n = n; p.h = h; p.w = w; p.c = c; p.fh = fh; p.fw = fw; p.upx = upx; p.upy = upy;
    p.downx = downx; p.downy = downy; p.padx0 = padx0; p.pady0 = pady0; p.flip = flip; p.gain = gain;
    return launch_upfirdn(p, padx1, pady1, (cudaStream_t)stream);
}

// upfirdn2d with fused consumers: optional SynthesisLayer epilogue (act != 0: + noise*strength + bias, lrelu, gain, clamp --
// conv2d_resample.py:128 followed by networks_stylegan2.py:318-329 in one pass) and optional split-bf16 copies of the result
// (y, y_hi/y_lo: at least one; the bf16 pair feeds the next tensor-core convolution).
B200_API int b200_upfirdn2d_fused(const float* x, const float* f, const float* add, float* y, void* y_hi, void* y_lo, int n, int h,
                                  int w, int c, int fh, int fw, int up, int down, int padx0, int padx1, int pady0, int pady1,
                                  int flip, float gain, int act, const float* bias, const float* noise, const float* strength,
                                  long noise_bs, int lrelu, float alpha, float act_gain, float clamp, int separable, void* stream) {
    UpfirdnParams p{};
    p.separable = separable;
    p.x = x; p.f = f; p.add = add; p.y = y; p.yhi = (uint2*)y_hi; p.ylo = (uint2*)y_lo; p.n = n; p.h = h; p.w = w; p.c = c;
    p.fh = fh; p.fw = fw; p.upx = p.upy = up; p.downx = p.downy = down; p.padx0 = padx0; p.pady0 = pady0; p.flip = flip; p.gain = gain;
    p.act = act; p.bias = bias; p.noise = noise; p.strength = strength; p.noise_bs = noise_bs; p.lrelu = lrelu; p.alpha = alpha;
    p.act_gain = act_gain; p.clamp = clamp;
    return launch_upfirdn(p, padx1, pady1, (cudaStream_t)stream);
}
