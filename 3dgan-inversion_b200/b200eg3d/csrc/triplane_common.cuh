// Shared pieces of the fused tri-plane sampler + decoder kernels (triplane.cu: mma.sync generation, triplane_tc.cu: tcgen05
// generation): parameter block, point -> coordinate mapping, bilinear set-up, gather / scatter of texel lines.
#pragma once
#include "common.cuh"
#include <cuda_bf16.h>

namespace tri {
constexpr int C = 32, HID = 64, OUT = 33, PC = 96;
constexpr int SP = 24;      // words per point of the bilinear set-up tile: 12 texel offsets + 12 weights

// MUFU-based activations (ex2 / lg2 approximations): absolute error ~1e-6, far inside the 1e-3 parity budget, and
// ~4x fewer instructions than expf / log1pf (the decoder evaluates 64 softplus + 32 sigmoid per sample point).
__device__ __forceinline__ float softplus_fast(float x) { return x > 20.f ? x : __logf(1.f + __expf(x)); }
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

struct TriplaneParams {
    const float* planes; int n, hp, wp;
    const float* coords;                                   // [n][P][3] or null (ray mode)
    const float* ray_o; const float* ray_d; const float* depths; int S;   // ray mode: point p belongs to ray p / S
    long P;
    float coord_scale;                                     // 2 / box_warp
    const float* W1; const float* b1; const float* W2; const float* b2;
    float w1g, b1g, w2g, b2g;
    float* rgb; float* sigma;
    // backward only
    const float* d_rgb; const float* d_sigma;
    float* d_planes; float* d_coords;
    float* dW1; float* db1; float* dW2; float* db2;
    int fwd_passes;                                         // 3: split operands (default), 1: single pass (non-parity fast mode)
    int ray_w, ray_h;                                       // ray mode: ray image width / height for the column-major work order (0: linear order)
    long M;                                                 // ray mode: rays per sample = P / S
    float* d_ray_o; float* d_ray_d;                         // backward, ray mode: per-ray sums of d point and t * d point (may be null)
    uint8_t* f_save;                                        // forward: [n][ceil(P/128)][16384] image of every tile's layer-1 operand (may be null)
    const uint8_t* f_saved;                                 // backward: the same buffer
};

// Work order of the tcgen05 kernels.  The kernels are bound by L2 bandwidth (12 texel lines gathered per point), so the order
// in which a CTA walks its points decides how much of that traffic its L1 absorbs.  In ray mode with a known image size a tile
// of 128 work points is a PATCH of 8 x 16 neighbouring pixels at ONE depth index (tile T = patch * S + k; the tiles of a CTA are
// consecutive, i.e. it marches one patch front to back): on all three planes the 128 points fall into a compact window of
// texels (neighbouring rays are ~1.3 texels apart), most of which the next depth index touches again.
// ray_w <= 0, or an image that is not a multiple of 8 x 16: linear order.
// Returns the stored point index of work point pp, or -1 past the end.  32-bit arithmetic: the launch code guarantees
// P < 2^31 (64-bit divisions cost ~100 instructions each).
__device__ __forceinline__ int map_point(const TriplaneParams& p, unsigned pp) {
    if (pp >= (unsigned)p.P) return -1;
    if (p.coords || p.ray_w <= 0) return (int)pp;
    const unsigned S = (unsigned)p.S, Rw = (unsigned)p.ray_w;
    const unsigned T = pp >> 7, r = pp & 127u;
    const unsigned patch = T / S, k = T - patch * S;
    const unsigned ppr = Rw >> 3;                               // patches per image row
    const unsigned pr = patch / ppr, pc = patch - pr * ppr;
    const unsigned i = pr * 16u + (r >> 3), j = pc * 8u + (r & 7u);
    return (int)((i * Rw + j) * S + k);
}

// coordinates (grid units) of stored point pi of sample n (pi < 0: zeros); 32-bit index arithmetic
__device__ __forceinline__ void point_coords32(const TriplaneParams& p, int n, int pi, float& cx, float& cy, float& cz) {
    cx = cy = cz = 0.f;
    if (pi < 0) return;
    if (p.coords) {
        const float* c = p.coords + ((long)n * p.P + pi) * 3;
        cx = c[0]; cy = c[1]; cz = c[2];
    } else {
        const unsigned ray = (unsigned)pi / (unsigned)p.S;
        const float t = p.depths[(long)n * p.P + pi];
        const long ro = ((long)n * p.M + ray) * 3;
        const float* o = p.ray_o + ro;
        const float* d = p.ray_d + ro;
        cx = o[0] + t * d[0]; cy = o[1] + t * d[1]; cz = o[2] + t * d[2];
    }
    cx *= p.coord_scale; cy *= p.coord_scale; cz *= p.coord_scale;
}

struct Bilin {
    int x0, y0; float wx0, wx1, wy0, wy1; bool xin0, xin1, yin0, yin1;
};

__device__ __forceinline__ Bilin make_bilin(float gx, float gy, int hp, int wp) {
    // grid_sample unnormalise, align_corners=False: ix = ((x+1)*W - 1)/2 ; zeros padding
    float ix = ((gx + 1.f) * wp - 1.f) * 0.5f, iy = ((gy + 1.f) * hp - 1.f) * 0.5f;
    ix = fminf(fmaxf(ix, -2.f), (float)wp + 1.f);
    iy = fminf(fmaxf(iy, -2.f), (float)hp + 1.f);
    const float fx = floorf(ix), fy = floorf(iy);
    Bilin b;
    b.x0 = (int)fx; b.y0 = (int)fy;
    b.wx1 = ix - fx; b.wx0 = 1.f - b.wx1; b.wy1 = iy - fy; b.wy0 = 1.f - b.wy1;
    b.xin0 = b.x0 >= 0 && b.x0 < wp; b.xin1 = b.x0 + 1 >= 0 && b.x0 + 1 < wp;
    b.yin0 = b.y0 >= 0 && b.y0 < hp; b.yin1 = b.y0 + 1 >= 0 && b.y0 + 1 < hp;
    return b;
}

// per-channel partial derivatives of the bilinear value w.r.t. (ix, iy)
__device__ __forceinline__ void bilin_dcoord(const float* __restrict__ base, const Bilin& b, int wp, float& dix, float& diy) {
    const float* r0 = base + ((long)b.y0 * wp + b.x0) * PC;
    const float* r1 = r0 + (long)wp * PC;
    float t00 = 0.f, t01 = 0.f, t10 = 0.f, t11 = 0.f;
    if (b.yin0 && b.xin0) t00 = __ldg(r0);
    if (b.yin0 && b.xin1) t01 = __ldg(r0 + PC);
    if (b.yin1 && b.xin0) t10 = __ldg(r1);
    if (b.yin1 && b.xin1) t11 = __ldg(r1 + PC);
    dix = b.wy0 * (t01 - t00) + b.wy1 * (t11 - t10);
    diy = b.wx0 * (t10 - t00) + b.wx1 * (t11 - t01);
}

__device__ __forceinline__ void point_coords(const TriplaneParams& p, int n, long pi, float& cx, float& cy, float& cz) {
    cx = cy = cz = 0.f;
    if (pi >= p.P) return;
    if (p.coords) {
        const float* c = p.coords + ((long)n * p.P + pi) * 3;
        cx = c[0]; cy = c[1]; cz = c[2];
    } else {
        const long M = p.P / p.S;
        const long ray = pi / p.S;
        const float t = p.depths[(long)n * p.P + pi];
        const float* o = p.ray_o + ((long)n * M + ray) * 3;
        const float* d = p.ray_d + ((long)n * M + ray) * 3;
        cx = o[0] + t * d[0]; cy = o[1] + t * d[1]; cz = o[2] + t * d[2];
    }
    cx *= p.coord_scale; cy *= p.coord_scale; cz *= p.coord_scale;
}

// Bilinear set-up of one plane for the lane's own point: 4 texel offsets (in floats, plane offset included) and 4 weights
// with the plane mean (1/3) folded in; out-of-range texels get weight 0 and an address clamped onto a valid texel.
__device__ __forceinline__ void plane_setup(float u, float v, int hp, int wp, int plane, float* off, float* wgt) {
    float ix = ((u + 1.f) * wp - 1.f) * 0.5f, iy = ((v + 1.f) * hp - 1.f) * 0.5f;
    ix = fminf(fmaxf(ix, -2.f), (float)wp + 1.f);
    iy = fminf(fmaxf(iy, -2.f), (float)hp + 1.f);
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    const float wx1 = ix - fx, wx0 = 1.f - wx1, wy1 = iy - fy, wy0 = 1.f - wy1;
    const bool xi0 = x0 >= 0 && x0 < wp, xi1 = x0 + 1 >= 0 && x0 + 1 < wp, yi0 = y0 >= 0 && y0 < hp, yi1 = y0 + 1 >= 0 && y0 + 1 < hp;
    const int x0c = min(max(x0, 0), wp - 1), x1c = min(max(x0 + 1, 0), wp - 1);
    const int y0c = min(max(y0, 0), hp - 1), y1c = min(max(y0 + 1, 0), hp - 1);
    const int pc = plane * C;
    const float third = 1.f / 3.f;
    off[0] = __int_as_float((y0c * wp + x0c) * PC + pc); wgt[0] = (yi0 && xi0) ? wy0 * wx0 * third : 0.f;
    off[1] = __int_as_float((y0c * wp + x1c) * PC + pc); wgt[1] = (yi0 && xi1) ? wy0 * wx1 * third : 0.f;
    off[2] = __int_as_float((y1c * wp + x0c) * PC + pc); wgt[2] = (yi1 && xi0) ? wy1 * wx0 * third : 0.f;
    off[3] = __int_as_float((y1c * wp + x1c) * PC + pc); wgt[3] = (yi1 && xi1) ? wy1 * wx1 * third : 0.f;
}

// Extra set-up for the coordinate gradient: fractional weights (wx1, wy1) of one plane and a validity mask of its four texels
// (bit 0: (y0,x0), 1: (y0,x1), 2: (y1,x0), 3: (y1,x1)), same unnormalisation as plane_setup.
__device__ __forceinline__ void plane_setup_frac(float u, float v, int hp, int wp, float& wx1, float& wy1, int& mask) {
    float ix = ((u + 1.f) * wp - 1.f) * 0.5f, iy = ((v + 1.f) * hp - 1.f) * 0.5f;
    ix = fminf(fmaxf(ix, -2.f), (float)wp + 1.f);
    iy = fminf(fmaxf(iy, -2.f), (float)hp + 1.f);
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    wx1 = ix - fx; wy1 = iy - fy;
    const bool xi0 = x0 >= 0 && x0 < wp, xi1 = x0 + 1 >= 0 && x0 + 1 < wp, yi0 = y0 >= 0 && y0 < hp, yi1 = y0 + 1 >= 0 && y0 + 1 < hp;
    mask = (yi0 && xi0 ? 1 : 0) | (yi0 && xi1 ? 2 : 0) | (yi1 && xi0 ? 4 : 0) | (yi1 && xi1 ? 8 : 0);
}

// each lane stages the set-up of its own point: ss[lane*SP + 0..11] = texel offsets, [12..23] = weights
// (plane 0 <- (x,y), plane 1 <- (x,z), plane 2 <- (z,x): renderer.py:23-53)
__device__ __forceinline__ void stage_setup(float* ss, int lane, float cx, float cy, float cz, int hp, int wp) {
    float t[SP];
    plane_setup(cx, cy, hp, wp, 0, t + 0, t + 12);
    plane_setup(cx, cz, hp, wp, 1, t + 4, t + 16);
    plane_setup(cz, cx, hp, wp, 2, t + 8, t + 20);
#pragma unroll
    for (int j = 0; j < SP; j += 4) *reinterpret_cast<float4*>(&ss[lane * SP + j]) = make_float4(t[j], t[j + 1], t[j + 2], t[j + 3]);
}

__device__ __forceinline__ void fma4(float4& a, float w, const float4& v) {          // two packed FP32 FMAs (FFMA2)
    fma2(a.x, a.y, w, v.x, v.y); fma2(a.z, a.w, w, v.z, v.w);
}

// Gather the 32-channel mean feature of the warp's 32 points into sf[point*STRIDE + channel].  Eight lanes share a point
// (4 channels each, one LDG.128 per texel), so one warp instruction fetches the same texel slot of FOUR points = four full
// 128-byte lines; no cross-lane reduction is needed.
template <int STRIDE, int UNROLL = 2>
__device__ __forceinline__ void gather_features(const float* __restrict__ pl, const float* ss, float* sf, int lane) {
    const int pt = lane >> 3, j4 = (lane & 7) * 4;
    const float* pc = pl + j4;
#pragma unroll UNROLL
    for (int q0 = 0; q0 < 32; q0 += 4) {
        const float* s = ss + (q0 + pt) * SP;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 12; k += 4) {
            const int4 o = *reinterpret_cast<const int4*>(s + k);
            const float4 w = *reinterpret_cast<const float4*>(s + 12 + k);
            fma4(acc, w.x, __ldg(reinterpret_cast<const float4*>(pc + o.x)));
            fma4(acc, w.y, __ldg(reinterpret_cast<const float4*>(pc + o.y)));
            fma4(acc, w.z, __ldg(reinterpret_cast<const float4*>(pc + o.z)));
            fma4(acc, w.w, __ldg(reinterpret_cast<const float4*>(pc + o.w)));
        }
        *reinterpret_cast<float4*>(&sf[(q0 + pt) * STRIDE + j4]) = acc;
    }
}

__device__ __forceinline__ void red_add4(float* addr, float w, const float4& g) {
    if (w != 0.f)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(w * g.x), "f"(w * g.y), "f"(w * g.z), "f"(w * g.w)
                     : "memory");
}

// Scatter d_f (sg[point*STRIDE + channel]) into the plane gradient with vector reductions, same lane mapping as the gather.
template <int STRIDE>
__device__ __forceinline__ void scatter_features(float* __restrict__ dpl, const float* ss, const float* sg, int lane, int cnt) {
    const int pt = lane >> 3, j4 = (lane & 7) * 4;
    float* pc = dpl + j4;
    for (int q0 = 0; q0 < 32; q0 += 4) {
        if (q0 + pt >= cnt) continue;
        const float* s = ss + (q0 + pt) * SP;
        const float4 g = *reinterpret_cast<const float4*>(&sg[(q0 + pt) * STRIDE + j4]);
#pragma unroll
        for (int k = 0; k < 12; k += 4) {
            const int4 o = *reinterpret_cast<const int4*>(s + k);
            const float4 w = *reinterpret_cast<const float4*>(s + 12 + k);
            red_add4(pc + o.x, w.x, g); red_add4(pc + o.y, w.y, g); red_add4(pc + o.z, w.z, g); red_add4(pc + o.w, w.w, g);
        }
    }
}

}  // namespace tri
