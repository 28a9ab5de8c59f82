// fp32 SIMT implicit-GEMM convolution kernels over NHWC activations.
//
// These are the exact-fp32 "universal" path of the modulated-conv stack: every conv variant of
// the StyleGAN2 / SR blocks (reference: training/networks_stylegan2.py:34-91,
// torch_utils/ops/conv2d_resample.py:48-143) is expressed as one of two gather-GEMMs
//   pix  : C[pixel, n]     = sum_{tap, c} A[src(pixel, tap), c] * B[tap, c, n]        (fwd, dgrad, 1x1, convT classes)
//   wgrad: C[tap, m, n]    = sum_{pixel}  A[srcA(pixel, tap), m] * B[srcB(pixel, tap), n]
// with a tap table (dy, dx, weight-slice) and integer strides describing the gather.  They handle any
// channel count (guards everywhere) and are used for layers whose shapes do not fit the tcgen05 tiles.
#include "common.cuh"
#include <cuda_bf16.h>
#include "conv_geom.h"

#define BM 128
#define BN 128
#define BK 8

__global__ void __launch_bounds__(256) conv_pix_kernel(ConvPixParams p) {
    __shared__ __align__(16) float As[BK][BM];
    __shared__ __align__(16) float Bs[BK][BN];
    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const float* __restrict__ A = p.A + (long)b * p.a_bs;
    const float* __restrict__ B = p.B + (long)b * p.b_bs;
    float* __restrict__ C = p.C + (long)b * p.c_bs;
    const int M = p.g.Hi * p.g.Wi;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    // A-tile loader coordinates: one pixel row, 4 consecutive channels.
    const int a_row = tid >> 1, a_kq = (tid & 1) * 4;
    const int am = m0 + a_row;
    const bool a_valid = am < M;
    const int a_iy = a_valid ? am / p.g.Wi : 0, a_ix = a_valid ? am % p.g.Wi : 0;
    const bool vecA = (p.g.Ca & 3) == 0;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int ty = tid >> 4, tx = tid & 15;

    for (int t = 0; t < p.g.ntaps; ++t) {
        const int sy = a_iy * p.g.sy + p.g.dy[t], sx = a_ix * p.g.sx + p.g.dx[t];
        const bool inb = a_valid && sy >= 0 && sy < p.g.Hs && sx >= 0 && sx < p.g.Ws;
        const float* arow = A + ((long)sy * p.g.Ws + sx) * p.g.Ca;
        const float* Bt = B + (long)p.g.wt[t] * p.b_ts;
        for (int c0 = 0; c0 < p.g.Ca; c0 += BK) {
            // ---- A tile
            float av[4] = {0.f, 0.f, 0.f, 0.f};
            const int ca = c0 + a_kq;
            if (inb) {
                if (vecA && ca + 3 < p.g.Ca) {
                    float4 v = __ldg(reinterpret_cast<const float4*>(arow + ca));
                    av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (ca + j < p.g.Ca) av[j] = __ldg(arow + ca + j);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) As[a_kq + j][a_row] = av[j];
            // ---- B tile
            if (p.b_mode == 1) {            // K-contiguous: 4 consecutive k for one n
                const int n = tid >> 1, kq = (tid & 1) * 4;
                float bv[4] = {0.f, 0.f, 0.f, 0.f};
                const int c = c0 + kq;
                if (n0 + n < p.N) {
                    const float* bp = Bt + (long)(n0 + n) * p.b_ns + c;
                    if (c + 3 < p.g.Ca) {
                        float4 v = __ldg(reinterpret_cast<const float4*>(bp));
                        bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (c + j < p.g.Ca) bv[j] = __ldg(bp + j);
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) Bs[kq + j][n] = bv[j];
            } else if (p.b_mode == 2) {     // N-contiguous: 4 consecutive n for one k
                const int k = tid >> 5, n4 = (tid & 31) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const int c = c0 + k;
                if (c < p.g.Ca && n0 + n4 + 3 < p.N)
                    v = __ldg(reinterpret_cast<const float4*>(Bt + (long)c * p.b_ks + n0 + n4));
                else if (c < p.g.Ca) {
                    const float* bp = Bt + (long)c * p.b_ks + n0 + n4;
                    if (n0 + n4 + 0 < p.N) v.x = __ldg(bp + 0);
                    if (n0 + n4 + 1 < p.N) v.y = __ldg(bp + 1);
                    if (n0 + n4 + 2 < p.N) v.z = __ldg(bp + 2);
                }
                *reinterpret_cast<float4*>(&Bs[k][n4]) = v;
            } else {                         // generic strides
                const int k = tid >> 5, n4 = (tid & 31) * 4;
                const int c = c0 + k;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v = 0.f;
                    if (c < p.g.Ca && n0 + n4 + j < p.N) v = __ldg(Bt + (long)c * p.b_ks + (long)(n0 + n4 + j) * p.b_ns);
                    Bs[k][n4 + j] = v;
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                float a[8], bb[8];
                *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
                *reinterpret_cast<float4*>(&bb[0]) = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                *reinterpret_cast<float4*>(&bb[4]) = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
    // ---- store
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
        if (m >= M) continue;
        const int iy = m / p.g.Wi, ix = m % p.g.Wi;
        const long opix = (long)(iy * p.g.osy + p.g.ooy) * p.g.Wo + (ix * p.g.osx + p.g.oox);
        float* crow = C + opix * p.ldc;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
            if (n < p.N) {
                float v = acc[i][j];
                if (p.accumulate) v += crow[n];
                crow[n] = v;
            }
        }
    }
}

// C[wt[t]][m][n] (+)= sum_pixels A[srcA(pixel,t)][m] * B[srcB(pixel,t)][n]
__global__ void __launch_bounds__(256) conv_wgrad_kernel(ConvWgradParams p) {
    __shared__ __align__(16) float As[BK][BM];
    __shared__ __align__(16) float Bs[BK][BN];
    const int tid = threadIdx.x;
    int z = blockIdx.z;
    const int ks = z % p.ksplit; z /= p.ksplit;
    const int t = z % p.ntaps;
    const int b = z / p.ntaps;
    const float* __restrict__ A = p.A + (long)b * p.a_bs;
    const float* __restrict__ B = p.B + (long)b * p.b_bs;
    float* __restrict__ C = p.C + (long)b * p.c_bs + (long)p.wt[t] * p.Cm * p.Cn;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int K = p.Hi * p.Wi;
    const int kchunk = ((K + p.ksplit - 1) / p.ksplit + BK - 1) / BK * BK;
    const int kbeg = ks * kchunk, kend = min(K, kbeg + kchunk);

    const int kk = tid >> 5, c4 = (tid & 31) * 4;
    const bool vecA = (p.Cm & 3) == 0, vecB = (p.Cn & 3) == 0;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const int ty = tid >> 4, tx = tid & 15;

    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        const int k = k0 + kk;
        float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
        if (k < kend) {
            const int iy = k / p.Wi, ix = k % p.Wi;
            const int ay = iy * p.sAy + p.dAy[t], ax = ix * p.sAx + p.dAx[t];
            if (ay >= 0 && ay < p.HA && ax >= 0 && ax < p.WA) {
                const float* ap = A + ((long)ay * p.WA + ax) * p.Cm + m0 + c4;
                if (vecA && m0 + c4 + 3 < p.Cm) va = __ldg(reinterpret_cast<const float4*>(ap));
                else {
                    if (m0 + c4 + 0 < p.Cm) va.x = __ldg(ap + 0);
                    if (m0 + c4 + 1 < p.Cm) va.y = __ldg(ap + 1);
                    if (m0 + c4 + 2 < p.Cm) va.z = __ldg(ap + 2);
                    if (m0 + c4 + 3 < p.Cm) va.w = __ldg(ap + 3);
                }
            }
            const int by = iy * p.sBy + p.dBy[t], bx = ix * p.sBx + p.dBx[t];
            if (by >= 0 && by < p.HB && bx >= 0 && bx < p.WB) {
                const float* bp = B + ((long)by * p.WB + bx) * p.Cn + n0 + c4;
                if (vecB && n0 + c4 + 3 < p.Cn) vb = __ldg(reinterpret_cast<const float4*>(bp));
                else {
                    if (n0 + c4 + 0 < p.Cn) vb.x = __ldg(bp + 0);
                    if (n0 + c4 + 1 < p.Cn) vb.y = __ldg(bp + 1);
                    if (n0 + c4 + 2 < p.Cn) vb.z = __ldg(bp + 2);
                    if (n0 + c4 + 3 < p.Cn) vb.w = __ldg(bp + 3);
                }
            }
        }
        *reinterpret_cast<float4*>(&As[kk][c4]) = va;
        *reinterpret_cast<float4*>(&Bs[kk][c4]) = vb;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < BK; ++q) {
            float a[8], bb[8];
            *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[q][ty * 4]);
            *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[q][64 + ty * 4]);
            *reinterpret_cast<float4*>(&bb[0]) = *reinterpret_cast<const float4*>(&Bs[q][tx * 4]);
            *reinterpret_cast<float4*>(&bb[4]) = *reinterpret_cast<const float4*>(&Bs[q][64 + tx * 4]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
        if (m >= p.Cm) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
            if (n < p.Cn) {
                if (p.ksplit > 1) atomicAdd(C + (long)m * p.Cn + n, acc[i][j]);
                else C[(long)m * p.Cn + n] = acc[i][j];
            }
        }
    }
}

// Thin 1x1 weight gradient (Cm <= 4 output channels, e.g. the 3-channel ToRGB of the super-resolution blocks):
// dW[m][n] = sum_pixels dy[pixel][m] * x[pixel][n].  Pure streaming reduction over x (HBM-bound); a 128x128 GEMM tile
// would waste 97% of its rows.
// SPLIT: x comes as its split-bf16 pair (x = hi + lo), the only copy the lean activation path keeps.
template <bool SPLIT>
__global__ void __launch_bounds__(256) conv_wgrad_thin_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ xhi,
                                                              const __nv_bfloat16* __restrict__ xlo, const float* __restrict__ dy,
                                                              float* __restrict__ dW, long npix, int cm, int cn, long x_bs,
                                                              long dy_bs, long c_bs) {
    __shared__ float red[4 * 512];
    const int b = blockIdx.y;
    if (SPLIT) { xhi += (long)b * x_bs; xlo += (long)b * x_bs; } else x += (long)b * x_bs;
    dy += (long)b * dy_bs; dW += (long)b * c_bs;
    const int c4 = cn >> 2, ppb = blockDim.x / c4;
    const int sub = threadIdx.x / c4, cc = threadIdx.x % c4;
    float acc[4][4];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[m][j] = 0.f;
    if (sub < ppb) {
        for (long pix = (long)blockIdx.x * ppb + sub; pix < npix; pix += (long)gridDim.x * ppb) {
            float4 xv;
            if (SPLIT) {
                const uint2 h2 = __ldg(reinterpret_cast<const uint2*>(xhi + pix * cn) + cc), l2 = __ldg(reinterpret_cast<const uint2*>(xlo + pix * cn) + cc);
                const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(&h2);
                const __nv_bfloat16* lb = reinterpret_cast<const __nv_bfloat16*>(&l2);
                xv = make_float4(__bfloat162float(hb[0]) + __bfloat162float(lb[0]), __bfloat162float(hb[1]) + __bfloat162float(lb[1]),
                                 __bfloat162float(hb[2]) + __bfloat162float(lb[2]), __bfloat162float(hb[3]) + __bfloat162float(lb[3]));
            } else {
                xv = __ldg(reinterpret_cast<const float4*>(x + pix * cn) + cc);
            }
            const float* d = dy + pix * cm;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                if (m < cm) {
                    const float dv = __ldg(d + m);
                    acc[m][0] = fmaf(dv, xv.x, acc[m][0]); acc[m][1] = fmaf(dv, xv.y, acc[m][1]);
                    acc[m][2] = fmaf(dv, xv.z, acc[m][2]); acc[m][3] = fmaf(dv, xv.w, acc[m][3]);
                }
            }
        }
    }
    for (int i = threadIdx.x; i < 4 * cn; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    if (sub < ppb) {
#pragma unroll
        for (int m = 0; m < 4; ++m)
            if (m < cm)
#pragma unroll
                for (int j = 0; j < 4; ++j) atomicAdd(&red[m * cn + cc * 4 + j], acc[m][j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cm * cn; i += blockDim.x) atomicAdd(dW + i, red[i]);
}

int launch_conv_wgrad_thin(const float* x, const void* x_hi, const void* x_lo, const float* dy, float* dW, int batch, long npix, int cm, int cn,
                           cudaStream_t st) {
    B200_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)batch * cm * cn, st));
    const int ppb = 256 / (cn / 4);
    const long nb = (npix + ppb - 1) / ppb;
    dim3 grid((unsigned)(nb < 148 * 4 ? nb : 148 * 4), batch);
    if (x)
        conv_wgrad_thin_kernel<false><<<grid, 256, 0, st>>>(x, nullptr, nullptr, dy, dW, npix, cm, cn, npix * cn, npix * cm, (long)cm * cn);
    else
        conv_wgrad_thin_kernel<true><<<grid, 256, 0, st>>>(nullptr, (const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, dy, dW, npix, cm, cn,
                                                          npix * cn, npix * cm, (long)cm * cn);
    B200_CHECK_LAUNCH();
    return 0;
}

// Thin forward of a 1x1 convolution with <= 4 output channels whose input exists as its split-bf16 pair (the ToRGB layers of the
// super-resolution blocks: 64 / 128 input channels at 512^2 / 256^2), with the layer's bias and clamp applied on the way out:
//   y[pix][o] = clamp(sum_c (x_hi + x_lo)[pix][c] * w[o][c] + bias[o]).
// A streaming read of x (HBM-bound); a 128-wide tensor-core tile would compute 125 columns of padding and write the raw sum for a
// second (bias + clamp) pass.  LPP = cin / 8 lanes share a pixel (8 channels each, one 16-byte load per half), their partial sums are
// combined with shuffles; a lane's 8 x cout weights stay in registers.
template <int NCH>
__global__ void __launch_bounds__(256) conv_fwd_thin_kernel(const uint4* __restrict__ xhi, const uint4* __restrict__ xlo,
                                                            const float* __restrict__ w, const float* __restrict__ bias,
                                                            float* __restrict__ y, long npix, int cm, int cn, int lpp, float clamp) {
    const int b = blockIdx.y;
    const long xoff = (long)b * npix * (cn >> 3);
    xhi += xoff; xlo += xoff; y += (long)b * npix * cm; w += (long)b * cm * cn;
    const int lane = threadIdx.x & 31, sub = lane / lpp, cl = lane % lpp;
    const int ppw = 32 / lpp;                                        // pixels per warp and iteration
    float wr[NCH][4][8];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int j = 0; j < 8; ++j) wr[ch][m][j] = m < cm ? w[m * cn + (ch * lpp + cl) * 8 + j] : 0.f;
    float bv[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) bv[m] = (bias && m < cm) ? bias[m] : 0.f;
    const long warp0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long p0 = warp0 * ppw; p0 < npix; p0 += nwarps * ppw) {
        const long pix = p0 + sub;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (pix < npix) {
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const long vi = pix * (cn >> 3) + ch * lpp + cl;
                const uint4 h = __ldg(xhi + vi), l = __ldg(xlo + vi);
                const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {                       // two bf16 per word: low half = even channel
                    const float x0 = __uint_as_float(hw[q] << 16) + __uint_as_float(lw[q] << 16);
                    const float x1 = __uint_as_float(hw[q] & 0xffff0000u) + __uint_as_float(lw[q] & 0xffff0000u);
#pragma unroll
                    for (int m = 0; m < 4; ++m) acc[m] = fmaf(x1, wr[ch][m][2 * q + 1], fmaf(x0, wr[ch][m][2 * q], acc[m]));
                }
            }
        }
        for (int o = lpp >> 1; o > 0; o >>= 1) {
#pragma unroll
            for (int m = 0; m < 4; ++m) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], o);
        }
        if (cl == 0 && pix < npix) {
#pragma unroll
            for (int m = 0; m < 4; ++m)
                if (m < cm) {
                    float t = acc[m] + bv[m];
                    if (clamp >= 0.f) t = fminf(fmaxf(t, -clamp), clamp);
                    y[pix * cm + m] = t;
                }
        }
    }
}

int launch_conv_fwd_thin(const void* x_hi, const void* x_lo, const float* w, const float* bias, float* y, int batch, long npix, int cm,
                         int cn, float clamp, cudaStream_t st) {
    if (npix <= 0 || batch <= 0) return 0;
    int lpp = cn >> 3, nch = 1;
    if (lpp > 32) { nch = lpp / 32; lpp = 32; }
    const long nb = (npix * lpp + 255) / 256;
    dim3 grid((unsigned)(nb < 148 * 8 ? nb : 148 * 8), batch);
    if (nch == 1) conv_fwd_thin_kernel<1><<<grid, 256, 0, st>>>((const uint4*)x_hi, (const uint4*)x_lo, w, bias, y, npix, cm, cn, lpp, clamp);
    else conv_fwd_thin_kernel<2><<<grid, 256, 0, st>>>((const uint4*)x_hi, (const uint4*)x_lo, w, bias, y, npix, cm, cn, lpp, clamp);
    B200_CHECK_LAUNCH();
    return 0;
}

// Thin data gradient of a 1x1 convolution with <= 4 output channels (the ToRGB layers of the super-resolution blocks):
// dx[pix][ci] = sum_o dy[pix][o] * w[o][ci] is a streaming write of the cin-wide activation gradient -- one thread per
// (pixel, 4 input channels), weights in shared memory.
__global__ void __launch_bounds__(256) conv_dgrad_thin_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                              float* __restrict__ dx, long npix, int cm, int cn, long w_bs) {
    __shared__ float sw[4 * 512];
    const int b = blockIdx.y;
    dy += (long)b * npix * cm; dx += (long)b * npix * cn; w += (long)b * w_bs;
    for (int i = threadIdx.x; i < cm * cn; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const int c4 = cn >> 2;
    const long total = npix * c4;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long pix = i / c4;
        const int cc = (int)(i - pix * c4) * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int m = 0; m < 4; ++m)
            if (m < cm) {
                const float d = __ldg(dy + pix * cm + m);
                const float4 wv = *reinterpret_cast<const float4*>(&sw[m * cn + cc]);
                acc.x = fmaf(d, wv.x, acc.x); acc.y = fmaf(d, wv.y, acc.y); acc.z = fmaf(d, wv.z, acc.z); acc.w = fmaf(d, wv.w, acc.w);
            }
        *reinterpret_cast<float4*>(dx + pix * cn + cc) = acc;
    }
}

int launch_conv_dgrad_thin(const float* dy, const float* w, float* dx, int batch, long npix, int cm, int cn, cudaStream_t st) {
    if (npix <= 0 || batch <= 0) return 0;
    const long nb = (npix * (cn / 4) + 255) / 256;
    dim3 grid((unsigned)(nb < 148 * 16 ? nb : 148 * 16), batch);
    conv_dgrad_thin_kernel<<<grid, 256, 0, st>>>(dy, w, dx, npix, cm, cn, (long)cm * cn);
    B200_CHECK_LAUNCH();
    return 0;
}

int launch_conv_pix_simt(const ConvPixParams& p, int batch, cudaStream_t st) {
    const int M = p.g.Hi * p.g.Wi;
    if (M <= 0 || p.N <= 0 || batch <= 0) return 0;
    dim3 grid(cdiv(M, BM), cdiv(p.N, BN), batch);
    conv_pix_kernel<<<grid, 256, 0, st>>>(p);
    B200_CHECK_LAUNCH();
    return 0;
}

int launch_conv_wgrad_simt(ConvWgradParams p, int batch, cudaStream_t st) {
    const int K = p.Hi * p.Wi;
    if (K <= 0 || p.Cm <= 0 || p.Cn <= 0 || batch <= 0) return 0;
    const int tiles = cdiv(p.Cm, BM) * cdiv(p.Cn, BN) * p.ntaps * batch;
    int ksplit = 1;
    if (tiles < 296) ksplit = min(cdiv(296, tiles), max(1, K / 64));
    p.ksplit = ksplit;
    if (ksplit > 1)
        for (int b = 0; b < batch; ++b)
            B200_CUDA(cudaMemsetAsync(p.C + (long)b * p.c_bs, 0, sizeof(float) * (size_t)p.ctaps * p.Cm * p.Cn, st));
    dim3 grid(cdiv(p.Cm, BM), cdiv(p.Cn, BN), p.ntaps * ksplit * batch);
    conv_wgrad_kernel<<<grid, 256, 0, st>>>(p);
    B200_CHECK_LAUNCH();
    return 0;
}
