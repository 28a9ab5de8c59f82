// Stage-1 (w-projection) caller-side kernels: SURVEY.md section 8, "next" row f1.
//
//  * warp_uv     training/warping_loss.py:18-43 fused: rays of the predicted camera (ray_sampler.py:24-73) -> surface point
//                o + d * depth (ray_sampler.py:75-93) -> intersection of the line (canonical origin -> point) with the canonical
//                image plane (LinePlaneCollision, warping_loss.py:58-72) -> world-to-canonical-camera -> intrinsics -> uv in
//                [-1, 1].  The reference runs ~40 small launches over [R*R, 3] tensors; here one thread = one pixel, forward and
//                backward (gradients to the predicted extrinsic and to the rendered depth).
//  * noise pyramid   the noise regulariser of w_projector.py:221-237: for every noise buffer and every level of its 2x2
//                average-pool pyramid down to 8x8, (mean(n * roll(n, 1, x)))^2 + (mean(n * roll(n, 1, y)))^2.  One launch per
//                LEVEL for all buffers (<= 7 launches instead of ~900), forward and backward.
//  * noise normalise w_projector.py:262-268: n <- (n - mean) * rsqrt(mean((n - mean)^2)) for all buffers in two launches.
#include "common.cuh"

namespace {

constexpr int MAXB = 32, MAXL = 8;

__device__ __forceinline__ float block_sum_f(float v, float* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;
    if (w == 0) v = warp_sum(v);
    return v;
}

// ------------------------------------------------------------------------------------------------------------ warp_uv
struct WarpParams {
    const float* ext;       // predicted cam2world [4][4]
    const float* init_ext;  // canonical cam2world [4][4]
    const float* w2c;       // inverse of init_ext [4][4]
    const float* K;         // intrinsics [3][3]
    const float* depth;     // [R*R]
    int R;
    float* uv;              // [R*R][2]
    float* min_ndotu;       // device scalar: min |n . v| (the reference raises when it is < 1e-6); encoded as uint bits of a non-negative float
    const float* d_uv;      // backward
    float* d_ext;           // [4][4] accumulated
    float* d_depth;         // [R*R] written
};

template <bool BWD>
__global__ void warp_uv_kernel(WarpParams p) {
    __shared__ float sh[BWD ? 32 : 1];
    __shared__ float c[16 + 16 + 16 + 9];
    if (threadIdx.x < 16) { c[threadIdx.x] = p.ext[threadIdx.x]; c[16 + threadIdx.x] = p.init_ext[threadIdx.x]; c[32 + threadIdx.x] = p.w2c[threadIdx.x]; }
    if (threadIdx.x < 9) c[48 + threadIdx.x] = p.K[threadIdx.x];
    __syncthreads();
    const float* E = c; const float* I = c + 16; const float* W = c + 32; const float* K = c + 48;
    const float fx = K[0], sk = K[1], cx = K[2], fy = K[4], cy = K[5];
    const float oc[3] = {I[3], I[7], I[11]};                                  // canonical camera origin
    const float nrm[3] = {-oc[0], -oc[1], -oc[2]};                            // plane normal
    const float pp[3] = {I[2] + I[3], I[6] + I[7], I[10] + I[11]};            // init_ext @ (0,0,1,1)
    const float wv[3] = {oc[0] - pp[0], oc[1] - pp[1], oc[2] - pp[2]};
    const float nwv = nrm[0] * wv[0] + nrm[1] * wv[1] + nrm[2] * wv[2];
    const int total = p.R * p.R;
    float acc[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) acc[j] = 0.f;
    float mn = 3.0e38f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const float xc = ((float)(i % p.R) + 0.5f) / (float)p.R, yc = ((float)(i / p.R) + 0.5f) / (float)p.R;
        const float xl = (xc - cx + cy * sk / fy - sk * yc / fy) / fx, yl = (yc - cy) / fy;
        float q[3], dir[3], v[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) q[j] = E[4 * j] * xl + E[4 * j + 1] * yl + E[4 * j + 2];
        const float qn = fmaxf(sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]), 1e-12f);
        const float dep = p.depth[i];
#pragma unroll
        for (int j = 0; j < 3; ++j) { dir[j] = q[j] / qn; v[j] = E[4 * j + 3] + dir[j] * dep - oc[j]; }
        const float ndotu = nrm[0] * v[0] + nrm[1] * v[1] + nrm[2] * v[2];
        const float si = -nwv / ndotu;
        float psi[3], u3[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) psi[j] = wv[j] + si * v[j] + pp[j];
#pragma unroll
        for (int j = 0; j < 3; ++j) u3[j] = W[4 * j] * psi[0] + W[4 * j + 1] * psi[1] + W[4 * j + 2] * psi[2] + W[4 * j + 3];
        const float un0 = u3[0] / u3[2], un1 = u3[1] / u3[2];
        if (!BWD) {
            p.uv[2 * i] = (K[0] * un0 + K[1] * un1 + K[2] - 0.5f) * 2.f;
            p.uv[2 * i + 1] = (K[3] * un0 + K[4] * un1 + K[5] - 0.5f) * 2.f;
            mn = fminf(mn, fabsf(ndotu));
        } else {
            const float g0 = 2.f * p.d_uv[2 * i], g1 = 2.f * p.d_uv[2 * i + 1];
            const float dun0 = K[0] * g0 + K[3] * g1, dun1 = K[1] * g0 + K[4] * g1;
            float du3[3] = {dun0 / u3[2], dun1 / u3[2], -(dun0 * u3[0] + dun1 * u3[1]) / (u3[2] * u3[2])};
            float dpsi[3], dv[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) dpsi[j] = W[j] * du3[0] + W[4 + j] * du3[1] + W[8 + j] * du3[2];
            const float dsi = dpsi[0] * v[0] + dpsi[1] * v[1] + dpsi[2] * v[2];
            const float dnd = -dsi * si / ndotu;
            float ddep = 0.f, dd[3], dot = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) { dv[j] = si * dpsi[j] + dnd * nrm[j]; ddep += dv[j] * dir[j]; dd[j] = dv[j] * dep; dot += dir[j] * dd[j]; }
            p.d_depth[i] = ddep;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float dq = (dd[j] - dir[j] * dot) / qn;
                acc[4 * j] += dq * xl; acc[4 * j + 1] += dq * yl; acc[4 * j + 2] += dq; acc[4 * j + 3] += dv[j];
            }
        }
    }
    if (!BWD) {
        for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        if ((threadIdx.x & 31) == 0 && p.min_ndotu) atomicMin(reinterpret_cast<unsigned*>(p.min_ndotu), __float_as_uint(mn));
    } else {
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            const float s = block_sum_f(acc[j], sh);
            if (threadIdx.x == 0) atomicAdd(p.d_ext + j, s);
        }
    }
}

// ------------------------------------------------------------------------------------------------------ noise pyramid
struct PyrParams {
    const float* src[MAXB];     // level input of buffer b (the noise buffer itself at level 0, else the pooled workspace)
    float* dst[MAXB];           // pooled output (null when this is the buffer's last level)
    int size[MAXB];             // side length at this level (0: buffer has no such level)
    float* sums;                // [nbuf][MAXL][2] running sums S_x, S_y
    int level, nbuf;
    // backward
    const float* gsrc[MAXB];    // gradient w.r.t. the pooled (next) level, or null
    float* gdst[MAXB];          // gradient w.r.t. this level (written)
    const float* dreg;
};

__global__ void noise_level_fwd_kernel(PyrParams p) {
    __shared__ float sh[32];
    const int b = blockIdx.y, s = p.size[b];
    if (s == 0) return;
    const float* n = p.src[b];
    float* o = p.dst[b];
    float sx = 0.f, sy = 0.f;
    const int total = s * s;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int y = i / s, x = i - y * s;
        const float v = n[i];
        sx += v * n[y * s + (x == 0 ? s - 1 : x - 1)];
        sy += v * n[(y == 0 ? s - 1 : y - 1) * s + x];
        if (o && !(y & 1) && !(x & 1)) o[(y >> 1) * (s >> 1) + (x >> 1)] = 0.25f * (v + n[i + 1] + n[i + s] + n[i + s + 1]);
    }
    sx = block_sum_f(sx, sh); sy = block_sum_f(sy, sh);
    if (threadIdx.x == 0) { atomicAdd(p.sums + (b * MAXL + p.level) * 2, sx); atomicAdd(p.sums + (b * MAXL + p.level) * 2 + 1, sy); }
}

// reg += sum over buffers and levels of (S_x / numel)^2 + (S_y / numel)^2
__global__ void noise_reg_finalize_kernel(const float* sums, const int* sizes0, int nbuf, float* reg) {
    float acc = 0.f;
    for (int i = threadIdx.x; i < nbuf * MAXL; i += blockDim.x) {
        const int b = i / MAXL, l = i % MAXL;
        int s = sizes0[b];
        bool has = true;
        for (int k = 0; k < l; ++k) { if (s <= 8) has = false; s >>= 1; }
        if (!has || s == 0) continue;
        const float inv = 1.f / (float)(s * s);
        const float mx = sums[i * 2] * inv, my = sums[i * 2 + 1] * inv;
        acc += mx * mx + my * my;
    }
    __shared__ float sh[32];
    acc = block_sum_f(acc, sh);
    if (threadIdx.x == 0) atomicAdd(reg, acc);
}

__global__ void noise_level_bwd_kernel(PyrParams p) {
    const int b = blockIdx.y, s = p.size[b];
    if (s == 0) return;
    const float* n = p.src[b];
    const float* gn = p.gsrc[b];
    float* g = p.gdst[b];
    const float inv = 1.f / (float)(s * s);
    const float d = *p.dreg;
    const float cxm = d * 2.f * p.sums[(b * MAXL + p.level) * 2] * inv * inv, cym = d * 2.f * p.sums[(b * MAXL + p.level) * 2 + 1] * inv * inv;
    const int total = s * s;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int y = i / s, x = i - y * s;
        const float xl = n[y * s + (x == 0 ? s - 1 : x - 1)], xr = n[y * s + (x == s - 1 ? 0 : x + 1)];
        const float yu = n[(y == 0 ? s - 1 : y - 1) * s + x], yd = n[(y == s - 1 ? 0 : y + 1) * s + x];
        float v = cxm * (xl + xr) + cym * (yu + yd);
        if (gn) v += 0.25f * gn[(y >> 1) * (s >> 1) + (x >> 1)];
        g[i] = v;
    }
}

// --------------------------------------------------------------------------------------------------- noise normalise
struct NormParams { float* buf[MAXB]; int size[MAXB]; float* stats; };

__global__ void noise_stats_kernel(NormParams p) {
    __shared__ float sh[32];
    const int b = blockIdx.y, total = p.size[b] * p.size[b];
    float s1 = 0.f, s2 = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) { const float v = p.buf[b][i]; s1 += v; s2 += v * v; }
    s1 = block_sum_f(s1, sh); s2 = block_sum_f(s2, sh);
    if (threadIdx.x == 0) { atomicAdd(p.stats + 2 * b, s1); atomicAdd(p.stats + 2 * b + 1, s2); }
}

__global__ void noise_apply_kernel(NormParams p) {
    const int b = blockIdx.y, total = p.size[b] * p.size[b];
    const float mean = p.stats[2 * b] / (float)total;
    const float var = fmaxf(p.stats[2 * b + 1] / (float)total - mean * mean, 0.f);
    const float r = rsqrtf(var);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) p.buf[b][i] = (p.buf[b][i] - mean) * r;
}

int level_size(int s0, int level) {
    int s = s0;
    for (int k = 0; k < level; ++k) { if (s <= 8) return 0; s >>= 1; }
    return s;
}
}  // namespace

// uv [R*R][2] = canonical-view image coordinates in [-1, 1] of the surface points seen by the predicted camera.
// ext / init_ext / w2c (= inverse of init_ext): device [16]; K: device [9]; depth: device [R*R]; min_ndotu: device scalar
// initialised to a large float (may be NULL) that receives min |n . v| for the caller's "no intersection" check.
B200_API int b200_warp_uv_fwd(const float* ext, const float* init_ext, const float* w2c, const float* K, const float* depth, int R,
                              float* uv, float* min_ndotu, void* stream) {
    B200_REQUIRE(ext && init_ext && w2c && K && depth && uv && R > 0, "warp_uv_fwd: null pointer or bad resolution");
    WarpParams p{};
    p.ext = ext; p.init_ext = init_ext; p.w2c = w2c; p.K = K; p.depth = depth; p.R = R; p.uv = uv; p.min_ndotu = min_ndotu;
    const int blocks = (R * R + 255) / 256 < 148 ? (R * R + 255) / 256 : 148;
    warp_uv_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
    B200_CHECK_LAUNCH();
    return 0;
}

// d_ext [16] is ACCUMULATED into (zero it first; row 3 stays zero), d_depth [R*R] is written.
B200_API int b200_warp_uv_bwd(const float* ext, const float* init_ext, const float* w2c, const float* K, const float* depth, int R,
                              const float* d_uv, float* d_ext, float* d_depth, void* stream) {
    B200_REQUIRE(ext && init_ext && w2c && K && depth && d_uv && d_ext && d_depth && R > 0, "warp_uv_bwd: null pointer or bad resolution");
    WarpParams p{};
    p.ext = ext; p.init_ext = init_ext; p.w2c = w2c; p.K = K; p.depth = depth; p.R = R; p.d_uv = d_uv; p.d_ext = d_ext; p.d_depth = d_depth;
    const int blocks = (R * R + 255) / 256 < 148 ? (R * R + 255) / 256 : 148;
    warp_uv_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
    B200_CHECK_LAUNCH();
    return 0;
}

// Floats of workspace the pyramid needs (pooled levels of all buffers); the gradient workspace has the same size.
B200_API long b200_noise_pyramid_work_floats(int nbuf, const int* sizes) {
    long tot = 0;
    for (int b = 0; b < nbuf; ++b)
        for (int l = 1; l < MAXL; ++l) { const int s = level_size(sizes[b], l); tot += (long)s * s; }
    return tot;
}

// bufs: HOST array of nbuf device pointers (square fp32 buffers of side sizes[b], powers of two); work: device workspace of
// b200_noise_pyramid_work_floats floats; sums: device [nbuf][8][2], ZEROED by the caller (kept for the backward);
// sizes_dev: device copy of sizes; reg: device scalar, ACCUMULATED into.
B200_API int b200_noise_pyramid_fwd(int nbuf, const float* const* bufs, const int* sizes, const int* sizes_dev, float* work, float* sums,
                                    float* reg, void* stream) {
    B200_REQUIRE(nbuf >= 0 && nbuf <= MAXB, "noise_pyramid: at most 32 buffers");
    if (nbuf == 0) return 0;
    B200_REQUIRE(bufs && sizes && sizes_dev && sums && reg, "noise_pyramid_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const float* cur[MAXB];
    long off = 0;
    for (int b = 0; b < nbuf; ++b) {
        B200_REQUIRE(sizes[b] > 0 && (sizes[b] & (sizes[b] - 1)) == 0 && sizes[b] <= 1024, "noise_pyramid: sides must be powers of two <= 1024");
        cur[b] = bufs[b];
    }
    for (int l = 0; l < MAXL; ++l) {
        PyrParams p{};
        p.sums = sums; p.level = l; p.nbuf = nbuf;
        int maxs = 0;
        for (int b = 0; b < nbuf; ++b) {
            const int s = level_size(sizes[b], l), sn = level_size(sizes[b], l + 1);
            p.size[b] = s; p.src[b] = cur[b]; p.dst[b] = nullptr;
            if (s && sn && l + 1 < MAXL) { B200_REQUIRE(work, "noise_pyramid_fwd: workspace missing"); p.dst[b] = work + off; off += (long)sn * sn; cur[b] = p.dst[b]; }
            if (s > maxs) maxs = s;
        }
        if (maxs == 0) break;
        const int bx = (maxs * maxs + 255) / 256 < 64 ? (maxs * maxs + 255) / 256 : 64;
        noise_level_fwd_kernel<<<dim3(bx, nbuf), 256, 0, st>>>(p);
        B200_CHECK_LAUNCH();
    }
    noise_reg_finalize_kernel<<<1, 256, 0, st>>>(sums, sizes_dev, nbuf, reg);
    B200_CHECK_LAUNCH();
    return 0;
}

// dbufs: HOST array of nbuf device pointers receiving d reg / d buffer * (*dreg) (written); gwork: device workspace like `work`.
B200_API int b200_noise_pyramid_bwd(int nbuf, const float* const* bufs, const int* sizes, const float* work, const float* sums,
                                    const float* dreg, float* const* dbufs, float* gwork, void* stream) {
    B200_REQUIRE(nbuf >= 0 && nbuf <= MAXB, "noise_pyramid: at most 32 buffers");
    if (nbuf == 0) return 0;
    B200_REQUIRE(bufs && sizes && sums && dreg && dbufs, "noise_pyramid_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    // level pointers (same walk as the forward)
    const float* lv[MAXB][MAXL]; float* glv[MAXB][MAXL];
    long off = 0;
    for (int b = 0; b < nbuf; ++b) { lv[b][0] = bufs[b]; glv[b][0] = dbufs[b]; }
    for (int l = 0; l + 1 < MAXL; ++l)
        for (int b = 0; b < nbuf; ++b) {
            const int s = level_size(sizes[b], l), sn = level_size(sizes[b], l + 1);
            lv[b][l + 1] = nullptr; glv[b][l + 1] = nullptr;
            if (s && sn) { B200_REQUIRE(work && gwork, "noise_pyramid_bwd: workspace missing"); lv[b][l + 1] = work + off; glv[b][l + 1] = gwork + off; off += (long)sn * sn; }
        }
    for (int l = MAXL - 1; l >= 0; --l) {
        PyrParams p{};
        p.sums = const_cast<float*>(sums); p.level = l; p.nbuf = nbuf; p.dreg = dreg;
        int maxs = 0;
        for (int b = 0; b < nbuf; ++b) {
            const int s = level_size(sizes[b], l);
            p.size[b] = s; p.src[b] = lv[b][l]; p.gdst[b] = glv[b][l];
            p.gsrc[b] = (l + 1 < MAXL && level_size(sizes[b], l + 1)) ? glv[b][l + 1] : nullptr;
            if (s > maxs) maxs = s;
        }
        if (maxs == 0) continue;
        const int bx = (maxs * maxs + 255) / 256 < 64 ? (maxs * maxs + 255) / 256 : 64;
        noise_level_bwd_kernel<<<dim3(bx, nbuf), 256, 0, st>>>(p);
        B200_CHECK_LAUNCH();
    }
    return 0;
}

// In place: buf <- (buf - mean(buf)) * rsqrt(mean((buf - mean)^2)).  stats: device [nbuf][2], ZEROED by the caller.
B200_API int b200_noise_normalize(int nbuf, float* const* bufs, const int* sizes, float* stats, void* stream) {
    B200_REQUIRE(nbuf >= 0 && nbuf <= MAXB, "noise_normalize: at most 32 buffers");
    if (nbuf == 0) return 0;
    B200_REQUIRE(bufs && sizes && stats, "noise_normalize: null pointer");
    NormParams p{};
    int maxs = 0;
    for (int b = 0; b < nbuf; ++b) { p.buf[b] = bufs[b]; p.size[b] = sizes[b]; if (sizes[b] > maxs) maxs = sizes[b]; }
    p.stats = stats;
    const int bx = (maxs * maxs + 255) / 256 < 64 ? (maxs * maxs + 255) / 256 : 64;
    noise_stats_kernel<<<dim3(bx, nbuf), 256, 0, (cudaStream_t)stream>>>(p);
    B200_CHECK_LAUNCH();
    noise_apply_kernel<<<dim3(bx, nbuf), 256, 0, (cudaStream_t)stream>>>(p);
    B200_CHECK_LAUNCH();
    return 0;
}
