// Gather-GEMM geometry shared by the SIMT and tcgen05 convolution kernels.
#pragma once
#include <cuda_runtime.h>

struct ConvGeom {
    int Hi, Wi;              // iteration pixel grid (GEMM M = Hi*Wi)
    int Hs, Ws, Ca;          // gathered source tensor [Hs][Ws][Ca] (NHWC, one sample)
    int sy, sx;              // source pixel = (iy*sy + dy[t], ix*sx + dx[t]); out-of-range reads as zero
    int ntaps;
    int dy[9], dx[9], wt[9]; // per tap: source offset and weight-slice index
    int Wo, osy, osx, ooy, oox;  // output pixel = (iy*osy + ooy, ix*osx + oox) in a grid of width Wo
};

struct ConvPixParams {
    ConvGeom g;
    const float* A; long a_bs;                 // activations (+ per-sample stride)
    const float* B; long b_ts, b_ks, b_ns, b_bs;   // weights: B[wt*b_ts + c*b_ks + n*b_ns]
    int b_mode;                                // 1: b_ks == 1 (vector loads along k), 2: b_ns == 1, 0: generic
    int N;                                     // output channels
    float* C; long ldc, c_bs;
    int accumulate;
};

struct ConvWgradParams {
    const float* A; long a_bs; int HA, WA, Cm, sAy, sAx; int dAy[9], dAx[9];
    const float* B; long b_bs; int HB, WB, Cn, sBy, sBx; int dBy[9], dBx[9];
    int Hi, Wi;                                // reduction pixel grid
    int ntaps; int wt[9]; int ctaps;           // taps computed / tap slices present in C
    float* C; long c_bs;                       // C[wt][Cm][Cn]
    int ksplit;
};

int launch_conv_pix_simt(const ConvPixParams& p, int batch, cudaStream_t st);
int launch_conv_wgrad_simt(ConvWgradParams p, int batch, cudaStream_t st);
int launch_conv_wgrad_thin(const float* x, const void* x_hi, const void* x_lo, const float* dy, float* dW, int batch, long npix, int cm, int cn,
                           cudaStream_t st);
int launch_conv_dgrad_thin(const float* dy, const float* w, float* dx, int batch, long npix, int cm, int cn, cudaStream_t st);
int launch_conv_fwd_thin(const void* x_hi, const void* x_lo, const float* w, const float* bias, float* y, int batch, long npix, int cm,
                         int cn, float clamp, cudaStream_t st);
