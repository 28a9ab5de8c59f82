// C-ABI entry points of the convolution family (fp32 universal path).
// Activations are NHWC fp32 per sample; modulated weights are wmod[n][tap][cout][cin] (tap = kh*k + kw).
//
// Replaces, for the synthesis path, the reference's F.conv2d / F.conv_transpose2d calls made through
// torch_utils/ops/conv2d_gradfix.py:37-45 and torch_utils/ops/conv2d_resample.py:113-136.
#include "common.cuh"
#include "conv_geom.h"

thread_local char g_b200_err[512] = "";
int g_b200_pdl = -1;

static void plain_out(ConvGeom& g, int wo) { g.Wo = wo; g.osy = g.osx = 1; g.ooy = g.oox = 0; }

B200_API int b200_conv_fwd(const float* x, const float* wmod, float* y, int n, int h, int w, int cin, int cout,
                           int ksize, int up, void* stream) {
    B200_REQUIRE(ksize == 1 || ksize == 3, "conv_fwd: ksize must be 1 or 3");
    B200_REQUIRE(up == 1 || (up == 2 && ksize == 3), "conv_fwd: up must be 1, or 2 with ksize 3");
    cudaStream_t st = (cudaStream_t)stream;
    const int taps = ksize * ksize;
    ConvPixParams p{};
    p.A = x; p.a_bs = (long)h * w * cin;
    p.B = wmod; p.b_ts = (long)cout * cin; p.b_ks = 1; p.b_ns = cin; p.b_bs = (long)taps * cout * cin; p.b_mode = (cin % 4 == 0) ? 1 : 0;
    p.N = cout; p.C = y; p.ldc = cout; p.accumulate = 0;
    p.g.Hs = h; p.g.Ws = w; p.g.Ca = cin; p.g.sy = p.g.sx = 1;
    if (up == 1) {
        p.c_bs = (long)h * w * cout;
        p.g.Hi = h; p.g.Wi = w; p.g.ntaps = taps; plain_out(p.g, w);
        for (int t = 0; t < taps; ++t) { p.g.dy[t] = t / ksize - ksize / 2; p.g.dx[t] = t % ksize - ksize / 2; p.g.wt[t] = t; }
        return launch_conv_pix_simt(p, n, st);
    }
    // up == 2: stride-2 transposed convolution written straight onto the (2h+1)x(2w+1) grid, one launch per
    // output-parity class (each class has 4, 2, 2 or 1 contributing taps; no zero MACs, no atomics).
    p.c_bs = (long)(2 * h + 1) * (2 * w + 1) * cout;
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            p.g.Hi = h + 1 - py; p.g.Wi = w + 1 - px;
            p.g.Wo = 2 * w + 1; p.g.osy = p.g.osx = 2; p.g.ooy = py; p.g.oox = px;
            int t = 0;
            for (int kh = py; kh < 3; kh += 2)
                for (int kw = px; kw < 3; kw += 2) { p.g.dy[t] = -(kh >> 1); p.g.dx[t] = -(kw >> 1); p.g.wt[t] = kh * 3 + kw; ++t; }
            p.g.ntaps = t;
            if (int e = launch_conv_pix_simt(p, n, st)) return e;
        }
    return 0;
}

B200_API int b200_conv_dgrad(const float* dy, const float* wmod, float* dx, int n, int h, int w, int cin, int cout,
                             int ksize, int up, void* stream) {
    B200_REQUIRE(ksize == 1 || ksize == 3, "conv_dgrad: ksize must be 1 or 3");
    B200_REQUIRE(up == 1 || (up == 2 && ksize == 3), "conv_dgrad: up must be 1, or 2 with ksize 3");
    cudaStream_t st = (cudaStream_t)stream;
    const int taps = ksize * ksize;
    if (ksize == 1 && up == 1 && cout <= 4 && cin % 4 == 0 && cin <= 512)
        return launch_conv_dgrad_thin(dy, wmod, dx, n, (long)h * w, cout, cin, st);
    ConvPixParams p{};
    const int hs = up == 1 ? h : 2 * h + 1, ws = up == 1 ? w : 2 * w + 1;
    p.A = dy; p.a_bs = (long)hs * ws * cout;
    p.B = wmod; p.b_ts = (long)cout * cin; p.b_ks = cin; p.b_ns = 1; p.b_bs = (long)taps * cout * cin;
    p.b_mode = (cin % 4 == 0) ? 2 : 0;
    p.N = cin; p.C = dx; p.ldc = cin; p.c_bs = (long)h * w * cin; p.accumulate = 0;
    p.g.Hi = h; p.g.Wi = w; p.g.Hs = hs; p.g.Ws = ws; p.g.Ca = cout; p.g.sy = p.g.sx = up; p.g.ntaps = taps;
    plain_out(p.g, w);
    for (int t = 0; t < taps; ++t) {
        const int kh = t / ksize, kw = t % ksize;
        p.g.dy[t] = up == 1 ? ksize / 2 - kh : kh;
        p.g.dx[t] = up == 1 ? ksize / 2 - kw : kw;
        p.g.wt[t] = t;
    }
    return launch_conv_pix_simt(p, n, st);
}

B200_API int b200_conv_wgrad(const float* x, const float* dy, float* dwmod, int n, int h, int w, int cin, int cout,
                             int ksize, int up, void* stream) {
    B200_REQUIRE(ksize == 1 || ksize == 3, "conv_wgrad: ksize must be 1 or 3");
    B200_REQUIRE(up == 1 || (up == 2 && ksize == 3), "conv_wgrad: up must be 1, or 2 with ksize 3");
    cudaStream_t st = (cudaStream_t)stream;
    const int taps = ksize * ksize;
    if (ksize == 1 && up == 1 && cout <= 4 && cin % 4 == 0 && cin <= 512)
        return launch_conv_wgrad_thin(x, nullptr, nullptr, dy, dwmod, n, (long)h * w, cout, cin, st);
    ConvWgradParams p{};
    const int hs = up == 1 ? h : 2 * h + 1, ws = up == 1 ? w : 2 * w + 1;
    p.A = dy; p.a_bs = (long)hs * ws * cout; p.HA = hs; p.WA = ws; p.Cm = cout; p.sAy = p.sAx = up;
    p.B = x; p.b_bs = (long)h * w * cin; p.HB = h; p.WB = w; p.Cn = cin; p.sBy = p.sBx = 1;
    p.Hi = h; p.Wi = w; p.ntaps = taps; p.ctaps = taps;
    p.C = dwmod; p.c_bs = (long)taps * cout * cin;
    for (int t = 0; t < taps; ++t) {
        const int kh = t / ksize, kw = t % ksize;
        p.dAy[t] = up == 1 ? 0 : kh; p.dAx[t] = up == 1 ? 0 : kw;
        p.dBy[t] = up == 1 ? kh - ksize / 2 : 0; p.dBx[t] = up == 1 ? kw - ksize / 2 : 0;
        p.wt[t] = t;
    }
    return launch_conv_wgrad_simt(p, n, st);
}

// wgrad of a 1x1 convolution with at most 4 output channels (the ToRGB layers of the super-resolution blocks) whose input exists
// only as a split-bf16 pair (x = x_hi + x_lo): dwmod[n][1][cout][cin] = sum over pixels of dy (x) x.  1 when the shape is handled.
B200_API int b200_conv1x1_thin_supported(int cin, int cout) { return cout >= 1 && cout <= 4 && cin % 4 == 0 && cin >= 4 && cin <= 512; }
B200_API int b200_conv1x1_wgrad_split(const void* x_hi, const void* x_lo, const float* dy, float* dwmod, int n, long npix, int cin, int cout,
                                      void* stream) {
    B200_REQUIRE(b200_conv1x1_thin_supported(cin, cout), "conv1x1_wgrad_split: needs cout <= 4, cin % 4 == 0, cin <= 512");
    B200_REQUIRE(x_hi && x_lo && dy && dwmod, "conv1x1_wgrad_split: null pointer");
    if (n <= 0 || npix <= 0) return 0;
    return launch_conv_wgrad_thin(nullptr, x_hi, x_lo, dy, dwmod, n, npix, cout, cin, (cudaStream_t)stream);
}

// Forward of such a layer from the split pair, bias and clamp included: y [n][npix][cout] = clamp((x_hi + x_lo) . wmod^T + bias).
// wmod [n][cout][cin] fp32; bias [cout] or NULL; clamp < 0: none.  Shapes: b200_conv1x1_fwd_thin_supported.
B200_API int b200_conv1x1_fwd_thin_supported(int cin, int cout) {
    const int lpp = cin >> 3;
    return cout >= 1 && cout <= 4 && cin % 8 == 0 && cin >= 8 && cin <= 512 && (lpp & (lpp - 1)) == 0;
}
B200_API int b200_conv1x1_fwd_thin(const void* x_hi, const void* x_lo, const float* wmod, const float* bias, float* y, int n, long npix,
                                   int cin, int cout, float clamp, void* stream) {
    B200_REQUIRE(b200_conv1x1_fwd_thin_supported(cin, cout), "conv1x1_fwd_thin: needs cout <= 4 and cin a power of two in 8 .. 512");
    B200_REQUIRE(x_hi && x_lo && wmod && y, "conv1x1_fwd_thin: null pointer");
    if (n <= 0 || npix <= 0) return 0;
    return launch_conv_fwd_thin(x_hi, x_lo, wmod, bias, y, n, npix, cout, cin, clamp, (cudaStream_t)stream);
}

B200_API const char* b200_last_error() { return g_b200_err; }
B200_API int b200_version() { return 207; }

// Programmatic dependent launch on (1) / off (0) for subsequent launches; returns the previous setting.  Profiling aid: with PDL a
// traced kernel duration includes the time it waits for its predecessor.
B200_API int b200_set_pdl(int on) {
    const int prev = b200_pdl_enabled() ? 1 : 0;
    g_b200_pdl = on ? 1 : 0;
    return prev;
}
