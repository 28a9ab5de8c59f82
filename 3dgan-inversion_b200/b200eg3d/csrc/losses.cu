// Caller-side reductions of the inversion loops, fused ("next" row f2 of SURVEY.md section 8).
//
// PTI loss without the LPIPS term (training/coaches/base_coach.py:101-126, 294-305):
//     real128 = F.interpolate(real, (R, R), mode='area')
//     loss    = l2_lambda * (mse(image, real) + mse(image_raw, real128)) + tv_lambda * compute_tv_norm(image_depth)
// The reference spends ~25 elementwise / reduction launches on it per step; here the forward is ONE reduction kernel and
// the backward ONE elementwise kernel.  One thread owns one raw-resolution pixel: its f x f block of full-resolution
// pixels (f = H / R), the area-averaged target and the two forward differences of the depth map.
#include "common.cuh"

namespace {

struct LossParams {
    const float* image; long i_sn, i_sc, i_sh, i_sw;       // [n][C][H][W] with element strides (NCHW view of NHWC memory)
    const float* raw; long r_sn, r_sc, r_sh, r_sw;         // [n][C][R][R] with element strides (channel slice of the feature image)
    const float* depth;                                     // [n][R][R] contiguous
    const float* real;                                      // [n][C][H][W] contiguous
    int n, C, H, W, R, f;
    float l2_lambda, tv_lambda;
    float* loss;                                            // [4]: total, mse(image), mse(raw), tv   (accumulated into)
    const float* dloss;                                     // backward: device scalar
    float* d_image; long di_sn, di_sc, di_sh, di_sw;
    float* d_raw; long dr_sn, dr_sc, dr_sh, dr_sw;
    float* d_depth;                                         // [n][R][R]
};

__device__ __forceinline__ float block_sum(float v, float* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;
    if (w == 0) v = warp_sum(v);
    return v;                                               // valid in thread 0
}

// FF / CC: compile-time area factor and channel count (0 = take them from the parameters).  With both known (the inversion's 512 -> 128,
// RGB) the 48 + 48 loads of a thread are unrolled and issued together -- one trip to memory instead of 48 dependent ones: a thread's
// f x f x C loop was a serial chain of load -> accumulate (13 us for 6 MB).  Loads go through the read-only path (__ldg), so the
// backward's stores do not order them.
template <bool BWD, int FF, int CC>
__global__ void pti_loss_kernel(LossParams p) {
    __shared__ float sh[32];
    const int f = FF ? FF : p.f, C = CC ? CC : p.C;
    const long total = (long)p.n * p.R * p.R;
    const float inv_full = 1.f / ((float)p.n * C * p.H * p.W), inv_raw = 1.f / ((float)p.n * C * p.R * p.R);
    const float inv_tv = p.R > 1 ? 1.f / ((float)p.n * (p.R - 1) * (p.R - 1)) : 0.f;
    const float inv_area = 1.f / (float)(f * f);
    float s_full = 0.f, s_raw = 0.f, s_tv = 0.f;
    float g = 0.f;
    if (BWD) g = *p.dloss;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int x = (int)(i % p.R), y = (int)((i / p.R) % p.R), b = (int)(i / ((long)p.R * p.R));
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float area = 0.f;
#pragma unroll
            for (int yy = 0; yy < f; ++yy) {
                const int Y = y * f + yy;
                const float* rrow = p.real + (((long)b * C + c) * p.H + Y) * p.W + x * f;
                float t[4] = {0.f, 0.f, 0.f, 0.f};
                if (FF == 4) {                                      // (W % 4 == 0: the row segment is one aligned 16-byte vector)
                    const float4 tv = __ldg(reinterpret_cast<const float4*>(rrow));
                    t[0] = tv.x; t[1] = tv.y; t[2] = tv.z; t[3] = tv.w;
                }
#pragma unroll
                for (int xx = 0; xx < f; ++xx) {
                    const int X = x * f + xx;
                    const float tt = FF == 4 ? t[xx & 3] : __ldg(rrow + xx);
                    area += tt;
                    if (p.image) {
                        const float d = __ldg(p.image + b * p.i_sn + c * p.i_sc + Y * p.i_sh + X * p.i_sw) - tt;
                        if (BWD) p.d_image[b * p.di_sn + c * p.di_sc + Y * p.di_sh + X * p.di_sw] = g * p.l2_lambda * 2.f * d * inv_full;
                        else s_full += d * d;
                    }
                }
            }
            if (p.raw) {
                const float d = __ldg(p.raw + b * p.r_sn + c * p.r_sc + y * p.r_sh + x * p.r_sw) - area * inv_area;
                if (BWD) p.d_raw[b * p.dr_sn + c * p.dr_sc + y * p.dr_sh + x * p.dr_sw] = g * p.l2_lambda * 2.f * d * inv_raw;
                else s_raw += d * d;
            }
        }
        if (p.depth) {
            const float* dp = p.depth + (long)b * p.R * p.R;
            const float v = dp[y * p.R + x];
            if (!BWD) {
                if (y < p.R - 1 && x < p.R - 1) {
                    const float dx = dp[y * p.R + x + 1] - v, dy = dp[(y + 1) * p.R + x] - v;
                    s_tv += dx * dx + dy * dy;
                }
            } else {
                float a = 0.f;
                if (y < p.R - 1 && x < p.R - 1) a -= 2.f * ((dp[y * p.R + x + 1] - v) + (dp[(y + 1) * p.R + x] - v));
                if (x >= 1 && y < p.R - 1) a += 2.f * (v - dp[y * p.R + x - 1]);
                if (y >= 1 && x < p.R - 1) a += 2.f * (v - dp[(y - 1) * p.R + x]);
                p.d_depth[(long)b * p.R * p.R + y * p.R + x] = g * p.tv_lambda * a * inv_tv;
            }
        }
    }
    if (!BWD) {
        s_full = block_sum(s_full, sh); s_raw = block_sum(s_raw, sh); s_tv = block_sum(s_tv, sh);
        if (threadIdx.x == 0) {
            const float a = s_full * inv_full, r = s_raw * inv_raw, t = s_tv * inv_tv;
            atomicAdd(p.loss + 0, p.l2_lambda * (a + r) + p.tv_lambda * t);
            atomicAdd(p.loss + 1, a); atomicAdd(p.loss + 2, r); atomicAdd(p.loss + 3, t);
        }
    }
}

int fill(LossParams& p, const float* image, const long* istr, const float* raw, const long* rstr, const float* depth, const float* real,
         int n, int C, int H, int W, int R, float l2_lambda, float tv_lambda) {
    B200_REQUIRE(real && n > 0 && C > 0 && H > 0 && W == H && R > 0, "pti_loss: bad shape (square images only)");
    B200_REQUIRE(H % R == 0, "pti_loss: the raw resolution must divide the image resolution (area interpolation with an integer factor)");
    B200_REQUIRE((!image || istr) && (!raw || rstr), "pti_loss: strides missing");
    p.image = image; p.raw = raw; p.depth = depth; p.real = real;
    if (image) { p.i_sn = istr[0]; p.i_sc = istr[1]; p.i_sh = istr[2]; p.i_sw = istr[3]; }
    if (raw) { p.r_sn = rstr[0]; p.r_sc = rstr[1]; p.r_sh = rstr[2]; p.r_sw = rstr[3]; }
    p.n = n; p.C = C; p.H = H; p.W = W; p.R = R; p.f = H / R; p.l2_lambda = l2_lambda; p.tv_lambda = tv_lambda;
    return 0;
}
}  // namespace

// image [n][C][H][W] / raw [n][C][R][R] with element strides {n, c, h, w} (either may be NULL), depth [n][R][R] contiguous
// (may be NULL), real [n][C][H][W] contiguous.  loss[4] = {total, mse(image, real), mse(raw, area(real)), tv(depth)} is
// ACCUMULATED into (zero it first).
B200_API int b200_pti_loss_fwd(const float* image, const long* image_strides, const float* raw, const long* raw_strides,
                               const float* depth, const float* real, int n, int C, int H, int W, int R, float l2_lambda,
                               float tv_lambda, float* loss, void* stream) {
    LossParams p{};
    if (int e = fill(p, image, image_strides, raw, raw_strides, depth, real, n, C, H, W, R, l2_lambda, tv_lambda)) return e;
    B200_REQUIRE(loss, "pti_loss_fwd: null output");
    p.loss = loss;
    const long total = (long)n * R * R;
    const int blocks = (int)((total + 127) / 128 < 148 * 4 ? (total + 127) / 128 : 148 * 4);
    if (p.f == 4 && C == 3 && (W & 3) == 0 && ((uintptr_t)real & 15) == 0) pti_loss_kernel<false, 4, 3><<<blocks, 128, 0, (cudaStream_t)stream>>>(p);
    else pti_loss_kernel<false, 0, 0><<<blocks, 128, 0, (cudaStream_t)stream>>>(p);
    B200_CHECK_LAUNCH();
    return 0;
}

// Gradients of the total loss (scaled by the device scalar *dloss): d_image / d_raw are written with their own element strides,
// d_depth [n][R][R] contiguous; outputs whose input pointer is NULL are skipped.
B200_API int b200_pti_loss_bwd(const float* image, const long* image_strides, const float* raw, const long* raw_strides,
                               const float* depth, const float* real, int n, int C, int H, int W, int R, float l2_lambda,
                               float tv_lambda, const float* dloss, float* d_image, const long* d_image_strides, float* d_raw,
                               const long* d_raw_strides, float* d_depth, void* stream) {
    LossParams p{};
    if (int e = fill(p, image, image_strides, raw, raw_strides, depth, real, n, C, H, W, R, l2_lambda, tv_lambda)) return e;
    B200_REQUIRE(dloss, "pti_loss_bwd: null incoming gradient");
    B200_REQUIRE((!image || (d_image && d_image_strides)) && (!raw || (d_raw && d_raw_strides)) && (!depth || d_depth),
                 "pti_loss_bwd: every present input needs its gradient buffer");
    p.dloss = dloss; p.d_image = d_image; p.d_raw = d_raw; p.d_depth = d_depth;
    if (image) { p.di_sn = d_image_strides[0]; p.di_sc = d_image_strides[1]; p.di_sh = d_image_strides[2]; p.di_sw = d_image_strides[3]; }
    if (raw) { p.dr_sn = d_raw_strides[0]; p.dr_sc = d_raw_strides[1]; p.dr_sh = d_raw_strides[2]; p.dr_sw = d_raw_strides[3]; }
    const long total = (long)n * R * R;
    const int blocks = (int)((total + 127) / 128 < 148 * 4 ? (total + 127) / 128 : 148 * 4);
    if (p.f == 4 && C == 3 && (W & 3) == 0 && ((uintptr_t)real & 15) == 0) pti_loss_kernel<true, 4, 3><<<blocks, 128, 0, (cudaStream_t)stream>>>(p);
    else pti_loss_kernel<true, 0, 0><<<blocks, 128, 0, (cudaStream_t)stream>>>(p);
    B200_CHECK_LAUNCH();
    return 0;
}
