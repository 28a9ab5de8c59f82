// Grouped ("bank") style + weight preparation for ALL modulated-conv layers of a synthesis network in four launches.
//
// Styles depend only on the latents ws and the affine parameters (training/networks_stylegan2.py:312 `styles = self.affine(w)`),
// and the modulated / demodulated weights only on the styles and the conv weights (:58-67) -- never on activations.  So the
// 20 (backbone) or 6 (super-resolution) per-layer GEMV + weight-prep launches of a forward pass, and their ~2x as many
// backward launches, are batched here: one grid covers every (layer, output-channel) pair.
//   b200_bank_styles_fwd    styles_l[n][i] = (ws[n][widx_l] . A_l[i] / sqrt(w_dim) + b_l[i]) * post_l
//   b200_bank_weights_fwd   wmod_l = W_l * s_l * d_l  -> fp32 and / or split-bf16, GEMM layout [n][tap][cout][cin]
//   b200_bank_weights_bwd   d wmod_l -> d W_l, d s_l
//   b200_bank_styles_bwd    d s_l -> d A_l, d b_l, d ws
#include "common.cuh"
#include "bank.h"
#include <cuda_bf16.h>

namespace {

constexpr int MAXL = B200_BANK_MAX_LAYERS;

struct BankTable {
    B200BankLayer l[MAXL];
    int start[MAXL + 1];         // prefix sums of per-layer block counts
    int n_layers;
};

__device__ __forceinline__ int find_layer(const BankTable& t, int b) {
    int l = 0;
    while (l + 1 < t.n_layers && b >= t.start[l + 1]) ++l;
    return l;
}

__device__ __forceinline__ float block_sum_b(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    return s;
}

// one warp per style row; block = 8 warps = 8 rows of one layer
__global__ void __launch_bounds__(256) bank_styles_fwd_kernel(const __grid_constant__ BankTable t, const float* __restrict__ ws,
                                                              int n, int num_ws, int w_dim, float wgain) {
    const int l = find_layer(t, blockIdx.x);
    const B200BankLayer& L = t.l[l];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int i = (blockIdx.x - t.start[l]) * 8 + wid;
    if (i >= L.cin) return;
    const float* a = L.affine_w + (long)i * w_dim;
    for (int b = 0; b < n; ++b) {
        const float* w = ws + ((long)b * num_ws + L.widx) * w_dim;
        float acc = 0.f;
        for (int k = lane; k < w_dim; k += 32) acc = fmaf(__ldg(a + k), __ldg(w + k), acc);
        acc = warp_sum(acc);
        if (lane == 0) L.styles[(long)b * L.cin + i] = (acc * wgain + L.affine_b[i]) * L.post_scale;
    }
}

// block per (layer, cout); loops over samples.  Same math as weight_prep_kernel (modconv.cu).
__global__ void __launch_bounds__(256) bank_weights_fwd_kernel(const __grid_constant__ BankTable t, int n) {
    extern __shared__ float sW[];
    __shared__ float red[32];
    const int l = find_layer(t, blockIdx.x);
    const B200BankLayer& L = t.l[l];
    const int o = blockIdx.x - t.start[l];
    const int cin = L.cin, cout = L.cout, taps = L.taps;
    const float* Wo = L.weight + (long)o * cin * taps;
    __nv_bfloat16* whi = (__nv_bfloat16*)L.w_hi;
    __nv_bfloat16* wlo = (__nv_bfloat16*)L.w_lo;
    for (int b = 0; b < n; ++b) {
        const float* s = L.styles + (long)b * cin;
        float acc = 0.f;
        if (taps == 9) {
            for (int idx = threadIdx.x; idx < cin * 9; idx += blockDim.x) { const float v = Wo[idx] * s[idx / 9]; sW[idx] = v; acc = fmaf(v, v, acc); }
        } else {
            for (int idx = threadIdx.x; idx < cin * taps; idx += blockDim.x) { const float v = Wo[idx] * s[idx / taps]; sW[idx] = v; acc = fmaf(v, v, acc); }
        }
        float d = 1.f;
        if (L.demod) {
            acc = block_sum_b(acc, red);
            d = rsqrtf(acc + 1e-8f);
            if (threadIdx.x == 0) L.dcoef[(long)b * cout + o] = d;
        } else {
            __syncthreads();
        }
        const long ob = (long)b * taps * cout * cin;
        if ((cin & 1) == 0) {
            const int half = cin >> 1;
            for (int tp = 0; tp < taps; ++tp)
                for (int ih = threadIdx.x; ih < half; ih += blockDim.x) {
                    const int i = ih * 2;
                    const float v0 = sW[i * taps + tp] * d, v1 = sW[(i + 1) * taps + tp] * d;
                    const long oi = ob + ((long)tp * cout + o) * cin + i;
                    if (L.wmod) *reinterpret_cast<float2*>(L.wmod + oi) = make_float2(v0, v1);
                    if (whi) {
                        const __nv_bfloat162 hh = __floats2bfloat162_rn(v0, v1);
                        *reinterpret_cast<__nv_bfloat162*>(whi + oi) = hh;
                        if (wlo) *reinterpret_cast<__nv_bfloat162*>(wlo + oi) = __floats2bfloat162_rn(v0 - __low2float(hh), v1 - __high2float(hh));
                    }
                }
        } else {
            for (int idx = threadIdx.x; idx < cin * taps; idx += blockDim.x) {
                const int tp = idx / cin, i = idx % cin;
                const float v = sW[i * taps + tp] * d;
                const long oi = ob + ((long)tp * cout + o) * cin + i;
                if (L.wmod) L.wmod[oi] = v;
                if (whi) {
                    const __nv_bfloat16 hh = __float2bfloat16_rn(v);
                    whi[oi] = hh;
                    if (wlo) wlo[oi] = __float2bfloat16_rn(v - __bfloat162float(hh));
                }
            }
        }
        __syncthreads();
    }
}

// block per (layer, cout): d wmod -> d W (written) and d styles (atomics into the zeroed d_styles)
__global__ void __launch_bounds__(256) bank_weights_bwd_kernel(const __grid_constant__ BankTable t, int n) {
    extern __shared__ float sm[];
    __shared__ float red[32];
    const int l = find_layer(t, blockIdx.x);
    const B200BankLayer& L = t.l[l];
    if (!L.dwmod) return;
    const int o = blockIdx.x - t.start[l];
    const int cin = L.cin, cout = L.cout, taps = L.taps;
    float* sW = sm;
    float* sD = sm + cin * taps;
    const float* Wo = L.weight + (long)o * cin * taps;
    for (int idx = threadIdx.x; idx < cin * taps; idx += blockDim.x) { sW[idx] = Wo[idx]; sD[idx] = 0.f; }
    __syncthreads();
    for (int b = 0; b < n; ++b) {
        const float* s = L.styles + (long)b * cin;
        const float* G = L.dwmod + (long)b * taps * cout * cin;
        // styles hold s * post_scale already (the modulation used them as such)
        float d = 1.f, dot = 0.f;
        if (L.demod) {
            d = L.dcoef[(long)b * cout + o];
            float acc = 0.f;
            for (int tp = 0; tp < taps; ++tp)
                for (int i = threadIdx.x; i < cin; i += blockDim.x)
                    acc = fmaf(G[((long)tp * cout + o) * cin + i], sW[i * taps + tp] * s[i], acc);
            dot = block_sum_b(acc, red);
        }
        const float d3dot = d * d * d * dot;
        for (int i = threadIdx.x; i < cin; i += blockDim.x) {
            const float si = s[i];
            float ds = 0.f;
            for (int tp = 0; tp < taps; ++tp) {
                const float w = sW[i * taps + tp];
                float g = G[((long)tp * cout + o) * cin + i] * d;
                if (L.demod) g -= d3dot * w * si;
                ds = fmaf(w, g, ds);
                sD[i * taps + tp] += si * g;
            }
            if (L.d_styles) atomicAdd(L.d_styles + (long)b * cin + i, ds);
        }
    }
    __syncthreads();
    if (L.d_weight) {
        float* dWo = L.d_weight + (long)o * cin * taps;
        for (int idx = threadIdx.x; idx < cin * taps; idx += blockDim.x) dWo[idx] = sD[idx];
    }
}

// one warp per style row: d A_l[i][:], d b_l[i], and (atomically) d ws[n][widx_l][:]
__global__ void __launch_bounds__(256) bank_styles_bwd_kernel(const __grid_constant__ BankTable t, const float* __restrict__ ws,
                                                              float* __restrict__ d_ws, int n, int num_ws, int w_dim, float wgain) {
    const int l = find_layer(t, blockIdx.x);
    const B200BankLayer& L = t.l[l];
    if (!L.dwmod || !L.d_styles) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int i = (blockIdx.x - t.start[l]) * 8 + wid;
    if (i >= L.cin) return;
    const float* a = L.affine_w + (long)i * w_dim;
    float db = 0.f;
    for (int b = 0; b < n; ++b) {
        const float ds = L.d_styles[(long)b * L.cin + i] * L.post_scale;      // d(pre-scale style)
        db += ds;
        if (d_ws) {
            float* dw = d_ws + ((long)b * num_ws + L.widx) * w_dim;
            for (int k = lane; k < w_dim; k += 32) atomicAdd(dw + k, ds * wgain * __ldg(a + k));
        }
    }
    if (L.d_affine_b && lane == 0) L.d_affine_b[i] = db;
    if (L.d_affine_w) {
        float* da = L.d_affine_w + (long)i * w_dim;
        for (int k = lane; k < w_dim; k += 32) {
            float acc = 0.f;
            for (int b = 0; b < n; ++b)
                acc = fmaf(L.d_styles[(long)b * L.cin + i] * L.post_scale * wgain, __ldg(ws + ((long)b * num_ws + L.widx) * w_dim + k), acc);
            da[k] = acc;
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Vectorised weight kernels for the shapes every EG3D layer has (cin a multiple of 8 and <= 2048, 1 or 9 taps): one thread owns
// 8 consecutive input channels x all taps of ONE output channel -- TAPS*8 contiguous floats of W, read with 16-byte loads and
// kept in registers across the demodulation reduction -- and writes, per tap, one 16-byte bf16x8 vector each to w_hi / w_lo
// (cin-contiguous GEMM layout).  A cout's cin/8 threads form one reduction group; a 256-thread block holds 256/(cin/8) couts.
// sum over the `gs` consecutive threads of a group (gs a power of two; groups never straddle a block)
__device__ __forceinline__ float group_sum(float v, int gs, float* red) {
    if (gs <= 32) {
        for (int o = gs >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    v = warp_sum(v);
    const int wid = threadIdx.x >> 5, wpg = gs >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[wid] = v;
    __syncthreads();
    float s = 0.f;
    const int w0 = (wid / wpg) * wpg;
    for (int i = 0; i < wpg; ++i) s += red[w0 + i];
    return s;
}

__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    return make_uint4(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b), *reinterpret_cast<uint32_t*>(&c),
                      *reinterpret_cast<uint32_t*>(&d));
}

__device__ __forceinline__ uint2 pack4_bf16(const float* v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    return make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}

// CPT = channels per thread, chosen per table: 4 when a cout's cin/4 threads fit one block for every layer (cin <= 1024: every
// EG3D network), else 8.  TAPS*4 instead of TAPS*8 live weights per thread halves the register count, so two or three blocks
// share an SM and one block's loads overlap another's reduction barrier and stores (one resident block of 8 warps left the memory
// system idle between its phases).

template <int TAPS, int CPT>
__device__ __forceinline__ void weights_fwd_vec(const B200BankLayer& L, int blk, int n, float* red) {
    const int cin = L.cin, cout = L.cout, gs = cin / CPT, cpb = 256 / gs;     // gs: power of two for every EG3D layer; else fall back
    const int o = blk * cpb + (int)threadIdx.x / gs, i0 = ((int)threadIdx.x % gs) * CPT;
    const bool live = o < cout && (int)threadIdx.x < cpb * gs;
    float w[TAPS * CPT];
    if (live) {
        const float4* src = reinterpret_cast<const float4*>(L.weight + ((long)o * cin + i0) * TAPS);
#pragma unroll
        for (int q = 0; q < TAPS * CPT / 4; ++q) { const float4 v = __ldg(src + q); w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w; }
    }
    __nv_bfloat16* whi = (__nv_bfloat16*)L.w_hi;
    __nv_bfloat16* wlo = (__nv_bfloat16*)L.w_lo;
    for (int b = 0; b < n; ++b) {
        float v[TAPS * CPT];
        float acc = 0.f;
        if (live) {
            float s[CPT];
#pragma unroll
            for (int q = 0; q < CPT / 4; ++q) {
                const float4 sv = *reinterpret_cast<const float4*>(L.styles + (long)b * cin + i0 + 4 * q);
                s[4 * q] = sv.x; s[4 * q + 1] = sv.y; s[4 * q + 2] = sv.z; s[4 * q + 3] = sv.w;
            }
#pragma unroll
            for (int j = 0; j < CPT; ++j)
#pragma unroll
                for (int tp = 0; tp < TAPS; ++tp) { v[j * TAPS + tp] = w[j * TAPS + tp] * s[j]; acc = fmaf(v[j * TAPS + tp], v[j * TAPS + tp], acc); }
        }
        float d = 1.f;
        if (L.demod) {
            acc = group_sum(acc, gs, red);
            d = rsqrtf(acc + 1e-8f);
            if (live && i0 == 0) L.dcoef[(long)b * cout + o] = d;
        }
        if (live) {
            const long ob = (long)b * TAPS * cout * cin;
#pragma unroll
            for (int tp = 0; tp < TAPS; ++tp) {
                float x[CPT], lo[CPT];
#pragma unroll
                for (int j = 0; j < CPT; ++j) x[j] = v[j * TAPS + tp] * d;
                const long oi = ob + ((long)tp * cout + o) * cin + i0;
                if (L.wmod) {
#pragma unroll
                    for (int q = 0; q < CPT / 4; ++q)
                        *reinterpret_cast<float4*>(L.wmod + oi + 4 * q) = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
                }
                if (whi) {
                    __nv_bfloat16 hb[CPT];
                    if (CPT == 8) { const uint4 h = pack8_bf16(x); *reinterpret_cast<uint4*>(whi + oi) = h; *reinterpret_cast<uint4*>(hb) = h; }
                    else { const uint2 h = pack4_bf16(x); *reinterpret_cast<uint2*>(whi + oi) = h; *reinterpret_cast<uint2*>(hb) = h; }
                    if (wlo) {
#pragma unroll
                        for (int j = 0; j < CPT; ++j) lo[j] = x[j] - __bfloat162float(hb[j]);
                        if (CPT == 8) *reinterpret_cast<uint4*>(wlo + oi) = pack8_bf16(lo);
                        else *reinterpret_cast<uint2*>(wlo + oi) = pack4_bf16(lo);
                    }
                }
            }
        }
    }
}

// resident blocks per SM asked of the compiler (CPT = 4): forward 3 (80 registers, 48 B of spill: 48 -> 45 us per step), backward 2
// (128 registers; at 3 it spills 200 B per thread: 71 -> 107 us)
template <int CPT>
__global__ void __launch_bounds__(256, CPT == 4 ? 3 : 1) bank_weights_fwd_vec_kernel(const __grid_constant__ BankTable t, int n) {
    __shared__ float red[8];
    const int l = find_layer(t, blockIdx.x);
    const B200BankLayer& L = t.l[l];
    if (L.taps == 9) weights_fwd_vec<9, CPT>(L, blockIdx.x - t.start[l], n, red);
    else weights_fwd_vec<1, CPT>(L, blockIdx.x - t.start[l], n, red);
}

template <int TAPS, int CPT>
__device__ __forceinline__ void weights_bwd_vec(const B200BankLayer& L, int blk, int n, float* red, float* sds) {
    const int cin = L.cin, cout = L.cout, gs = cin / CPT, cpb = 256 / gs;
    const int o = blk * cpb + (int)threadIdx.x / gs, i0 = ((int)threadIdx.x % gs) * CPT;
    const bool live = o < cout && (int)threadIdx.x < cpb * gs;
    float w[TAPS * CPT], dW[TAPS * CPT];
#pragma unroll
    for (int q = 0; q < TAPS * CPT; ++q) { w[q] = 0.f; dW[q] = 0.f; }
    if (live) {
        const float4* src = reinterpret_cast<const float4*>(L.weight + ((long)o * cin + i0) * TAPS);
#pragma unroll
        for (int q = 0; q < TAPS * CPT / 4; ++q) { const float4 v = __ldg(src + q); w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w; }
    }
    for (int b = 0; b < n; ++b) {
        float s[CPT], g[TAPS * CPT];
        float d = 1.f, acc = 0.f;
        if (live) {
#pragma unroll
            for (int q = 0; q < CPT / 4; ++q) {
                const float4 sv = *reinterpret_cast<const float4*>(L.styles + (long)b * cin + i0 + 4 * q);
                s[4 * q] = sv.x; s[4 * q + 1] = sv.y; s[4 * q + 2] = sv.z; s[4 * q + 3] = sv.w;
            }
            const float* G = L.dwmod + (long)b * TAPS * cout * cin;
#pragma unroll
            for (int tp = 0; tp < TAPS; ++tp) {
                float gg[CPT];
#pragma unroll
                for (int q = 0; q < CPT / 4; ++q) {
                    const float4 gv = *reinterpret_cast<const float4*>(G + ((long)tp * cout + o) * cin + i0 + 4 * q);
                    gg[4 * q] = gv.x; gg[4 * q + 1] = gv.y; gg[4 * q + 2] = gv.z; gg[4 * q + 3] = gv.w;
                }
#pragma unroll
                for (int j = 0; j < CPT; ++j) { g[j * TAPS + tp] = gg[j]; acc = fmaf(gg[j], w[j * TAPS + tp] * s[j], acc); }
            }
        }
        float d3dot = 0.f;
        if (L.demod) {
            const float dot = group_sum(acc, gs, red);
            if (live) d = L.dcoef[(long)b * cout + o];
            d3dot = d * d * d * dot;
        }
        float dsv[CPT];
#pragma unroll
        for (int j = 0; j < CPT; ++j) dsv[j] = 0.f;
        if (live) {
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
#pragma unroll
                for (int tp = 0; tp < TAPS; ++tp) {
                    float gv = g[j * TAPS + tp] * d;
                    if (L.demod) gv -= d3dot * w[j * TAPS + tp] * s[j];
                    dsv[j] = fmaf(w[j * TAPS + tp], gv, dsv[j]);
                    dW[j * TAPS + tp] += s[j] * gv;
                }
            }
        }
        if (L.d_styles) {       // d styles: sum the block's couts in shared memory, then one coalesced atomic per input channel
            __syncthreads();
#pragma unroll
            for (int q = 0; q < CPT / 4; ++q)
                *reinterpret_cast<float4*>(sds + threadIdx.x * CPT + 4 * q) = make_float4(dsv[4 * q], dsv[4 * q + 1], dsv[4 * q + 2], dsv[4 * q + 3]);
            __syncthreads();
            for (int i = threadIdx.x; i < cin; i += 256) {
                float a = 0.f;
                for (int c = 0; c < cpb; ++c) a += sds[c * cin + i];
                atomicAdd(L.d_styles + (long)b * cin + i, a);
            }
        }
    }
    if (live && L.d_weight) {
        float4* dst = reinterpret_cast<float4*>(L.d_weight + ((long)o * cin + i0) * TAPS);
#pragma unroll
        for (int q = 0; q < TAPS * CPT / 4; ++q) dst[q] = make_float4(dW[4 * q], dW[4 * q + 1], dW[4 * q + 2], dW[4 * q + 3]);
    }
}

template <int CPT>
__global__ void __launch_bounds__(256, CPT == 4 ? 2 : 1) bank_weights_bwd_vec_kernel(const __grid_constant__ BankTable t, int n) {
    __shared__ float red[8];
    __shared__ __align__(16) float sds[256 * CPT];
    const int l = find_layer(t, blockIdx.x);
    const B200BankLayer& L = t.l[l];
    if (!L.dwmod) return;
    if (L.taps == 9) weights_bwd_vec<9, CPT>(L, blockIdx.x - t.start[l], n, red, sds);
    else weights_bwd_vec<1, CPT>(L, blockIdx.x - t.start[l], n, red, sds);
}

bool host_vec_ok(const B200BankLayer& L, int cpt) {
    const int gs = L.cin / cpt;
    return (L.taps == 9 || L.taps == 1) && (L.cin & 7) == 0 && L.cin <= 256 * cpt && (gs & (gs - 1)) == 0;
}

// block table of the vectorised kernels: ceil(cout / (256 / (cin/cpt))) blocks per layer; cpt = 4 when every layer allows it
int fill_table_vec(BankTable& t, const B200BankLayer* layers, int n_layers, bool& all_vec, int& cpt) {
    B200_REQUIRE(layers && n_layers > 0 && n_layers <= MAXL, "bank: between 1 and 32 layers");
    t.n_layers = n_layers;
    t.start[0] = 0;
    cpt = 4;
    for (int l = 0; l < n_layers; ++l)
        if (layers[l].cin > 1024) cpt = 8;
    all_vec = true;
    for (int l = 0; l < n_layers; ++l) {
        t.l[l] = layers[l];
        B200_REQUIRE(layers[l].cin > 0 && layers[l].cout > 0 && layers[l].taps > 0, "bank: bad layer shape");
        if (!host_vec_ok(layers[l], cpt)) { all_vec = false; break; }
        const int cpb = 256 / (layers[l].cin / cpt);
        t.start[l + 1] = t.start[l] + (layers[l].cout + cpb - 1) / cpb;
    }
    return 0;
}

int fill_table(BankTable& t, const B200BankLayer* layers, int n_layers, bool per_row8) {
    B200_REQUIRE(layers && n_layers > 0 && n_layers <= MAXL, "bank: between 1 and 32 layers");
    t.n_layers = n_layers;
    t.start[0] = 0;
    for (int l = 0; l < n_layers; ++l) {
        t.l[l] = layers[l];
        B200_REQUIRE(layers[l].cin > 0 && layers[l].cout > 0 && layers[l].taps > 0, "bank: bad layer shape");
        t.start[l + 1] = t.start[l] + (per_row8 ? (layers[l].cin + 7) / 8 : layers[l].cout);
    }
    return 0;
}

size_t max_smem(const B200BankLayer* layers, int n_layers, int mult) {
    size_t m = 0;
    for (int l = 0; l < n_layers; ++l) {
        const size_t s = (size_t)mult * sizeof(float) * layers[l].cin * layers[l].taps;
        if (s > m) m = s;
    }
    return m;
}

}  // namespace

B200_API int b200_bank_styles_fwd(const B200BankLayer* layers, int n_layers, const float* ws, int n, int num_ws, int w_dim,
                                  void* stream) {
    BankTable t;
    if (int e = fill_table(t, layers, n_layers, true)) return e;
    for (int l = 0; l < n_layers; ++l) B200_REQUIRE(layers[l].widx >= 0 && layers[l].widx < num_ws, "bank: latent index out of range");
    bank_styles_fwd_kernel<<<t.start[n_layers], 256, 0, (cudaStream_t)stream>>>(t, ws, n, num_ws, w_dim, 1.f / sqrtf((float)w_dim));
    B200_CHECK_LAUNCH();
    return 0;
}

B200_API int b200_bank_weights_fwd(const B200BankLayer* layers, int n_layers, int n, void* stream) {
    BankTable t;
    bool all_vec = false;
    int cpt = 8;
    if (int e = fill_table_vec(t, layers, n_layers, all_vec, cpt)) return e;
    if (all_vec) {
        if (cpt == 4) bank_weights_fwd_vec_kernel<4><<<t.start[n_layers], 256, 0, (cudaStream_t)stream>>>(t, n);
        else bank_weights_fwd_vec_kernel<8><<<t.start[n_layers], 256, 0, (cudaStream_t)stream>>>(t, n);
        B200_CHECK_LAUNCH();
        return 0;
    }
    if (int e = fill_table(t, layers, n_layers, false)) return e;
    const size_t smem = max_smem(layers, n_layers, 1);
    B200_REQUIRE(smem <= 48 * 1024, "bank: cin*taps too large for the staging tile");
    bank_weights_fwd_kernel<<<t.start[n_layers], 256, smem, (cudaStream_t)stream>>>(t, n);
    B200_CHECK_LAUNCH();
    return 0;
}

B200_API int b200_bank_weights_bwd(const B200BankLayer* layers, int n_layers, int n, void* stream) {
    BankTable t;
    bool all_vec = false;
    int cpt = 8;
    if (int e = fill_table_vec(t, layers, n_layers, all_vec, cpt)) return e;
    if (all_vec) {
        if (cpt == 4) bank_weights_bwd_vec_kernel<4><<<t.start[n_layers], 256, 0, (cudaStream_t)stream>>>(t, n);
        else bank_weights_bwd_vec_kernel<8><<<t.start[n_layers], 256, 0, (cudaStream_t)stream>>>(t, n);
        B200_CHECK_LAUNCH();
        return 0;
    }
    if (int e = fill_table(t, layers, n_layers, false)) return e;
    const size_t smem = max_smem(layers, n_layers, 2);
    B200_REQUIRE(smem <= 48 * 1024, "bank: cin*taps too large for the staging tiles");
    bank_weights_bwd_kernel<<<t.start[n_layers], 256, smem, (cudaStream_t)stream>>>(t, n);
    B200_CHECK_LAUNCH();
    return 0;
}

B200_API int b200_bank_styles_bwd(const B200BankLayer* layers, int n_layers, const float* ws, float* d_ws, int n, int num_ws,
                                  int w_dim, void* stream) {
    BankTable t;
    if (int e = fill_table(t, layers, n_layers, true)) return e;
    bank_styles_bwd_kernel<<<t.start[n_layers], 256, 0, (cudaStream_t)stream>>>(t, ws, d_ws, n, num_ws, w_dim, 1.f / sqrtf((float)w_dim));
    B200_CHECK_LAUNCH();
    return 0;
}
