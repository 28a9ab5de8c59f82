// Per-ray kernels of the volume renderer: one warp owns one ray.
//   b200_ray_depths_coarse   stratified depths             (renderer.py:224-247)
//   b200_depth_minmax        global min/max of all depths  (ray_marcher.py:50)
//   b200_ray_importance      coarse weights -> smoothed pdf -> inverse-CDF fine depths   (renderer.py:249-308)
//   b200_ray_composite_fwd   merge coarse+fine by depth, mid-point alpha compositing     (renderer.py:212-222, ray_marcher.py:25-57)
//   b200_ray_composite_bwd   its backward (gradients to colours and densities; depths carry no gradient)
#include "common.cuh"
#include <cfloat>

namespace {
constexpr int MAXS = 256;     // max samples per ray after merging (96+96 = 192 for the stress config)

__device__ __forceinline__ unsigned f2ord(float f) {      // order-preserving float -> uint
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
// where a ray's colour rows are prefetched into L2 (measured, 16 384 rays x 48 + 48, forward / backward in us; without prefetch 79 / 132):
//   0  both lists before the depth loads                                58.9 / 105.7   (the depth loads queue behind 96 prefetches)
//   1  coarse rows before the depth loads, fine rows after the ranking  52.6 / 101.1
//   2  both lists right after the depth loads are issued                53.5 / 101.0
//   3  coarse rows after the depth loads, fine rows after the ranking   51.1 /  97.4   <- default
#ifndef RAY_PF_MODE
#define RAY_PF_MODE 3
#endif

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// one ray's colour rows (one 128-byte line per sample) on their way into L2 while its depths are ranked and marched: the
// accumulation loop then waits for L2 hits, not for HBM, six times in a row
__device__ __forceinline__ void prefetch_rows(const float* rows, int count, int lane) {
    if (rows)
        for (int i = lane; i < count; i += 32) prefetch_l2(rows + (long)i * 32);
}

__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// min / max of a block's values -> the two order-preserving words: warp shuffles, then one warp over the per-warp results, ONE
// atomic pair per block (a pair per warp -- ~10 000 same-address atomics -- tripled the time of the kernel that makes the depths)
__device__ __forceinline__ void block_minmax_commit(float lo, float hi, unsigned* __restrict__ mm) {
    __shared__ float s_lo[32], s_hi[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) { s_lo[wid] = lo; s_hi[wid] = hi; }
    __syncthreads();
    if (wid == 0) {
        lo = lane < nw ? s_lo[lane] : FLT_MAX;
        hi = lane < nw ? s_hi[lane] : -FLT_MAX;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0 && lo <= hi) { atomicMin(mm, f2ord(lo)); atomicMax(mm + 1, f2ord(hi)); }
    }
}

__global__ void depths_coarse_kernel(const float* __restrict__ t_base, const float* __restrict__ u, float* __restrict__ t,
                                     long total, int S, float delta, unsigned* __restrict__ mm) {
    float lo = FLT_MAX, hi = -FLT_MAX;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const float v = t_base[i % S] + u[i] * delta;
        t[i] = v;
        lo = fminf(lo, v); hi = fmaxf(hi, v);
    }
    if (mm) block_minmax_commit(lo, hi, mm);         // the depth clamp's global range, gathered where the depths are made
}

__global__ void depth_minmax_kernel(const float* __restrict__ t, long total, unsigned* __restrict__ mm) {
    float lo = FLT_MAX, hi = -FLT_MAX;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const float v = t[i];
        lo = fminf(lo, v); hi = fmaxf(hi, v);
    }
    block_minmax_commit(lo, hi, mm);
}

// alpha_k / weights of the mid-point rule on SORTED samples held in shared memory (lane-strided + lane-0 scan)
__device__ __forceinline__ void march_weights(const float* ts, const float* ss, int S, float* alpha, float* trans, float* w,
                                              int lane) {
    for (int k = lane; k < S - 1; k += 32) {
        const float delta = ts[k + 1] - ts[k];
        const float sp = softplus_f(0.5f * (ss[k] + ss[k + 1]) - 1.f);
        alpha[k] = 1.f - expf(-sp * delta);
    }
    __syncwarp();
    // exclusive running product of (1 - alpha + 1e-10): warp-wide multiplicative scan, 32 intervals per round
    float carry = 1.f;
    for (int k0 = 0; k0 < S - 1; k0 += 32) {
        const int k = k0 + lane;
        const float a = k < S - 1 ? alpha[k] : 0.f;
        float incl = k < S - 1 ? (1.f - a + 1e-10f) : 1.f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl *= up;
        }
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.f;
        const float T = carry * excl;
        if (k < S - 1) { trans[k] = T; w[k] = a * T; }
        carry *= __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(128) ray_importance_kernel(const float* __restrict__ t_c, const float* __restrict__ sigma_c,
                                                             const float* __restrict__ u, float* __restrict__ t_f,
                                                             long n_rays, int S, int S_imp, int SP) {
    extern __shared__ float dyn[];                 // per warp: 5 arrays of SP floats (SP = samples rounded up): more resident warps than a fixed 256
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* ts = dyn + wid * 5 * SP; float* ss = ts + SP; float* alpha = ss + SP; float* trans = alpha + SP; float* w = trans + SP;
    for (long ray = (long)blockIdx.x * 4 + wid; ray < n_rays; ray += (long)gridDim.x * 4) {
        for (int k = lane; k < S; k += 32) { ts[k] = t_c[ray * S + k]; ss[k] = sigma_c[ray * S + k]; }
        // this lane's first two uniform draws: in flight during the march (they are needed last)
        const float u0 = lane < S_imp ? __ldg(u + ray * S_imp + lane) : 0.f;
        const float u1 = lane + 32 < S_imp ? __ldg(u + ray * S_imp + lane + 32) : 0.f;
        __syncwarp();
        march_weights(ts, ss, S, alpha, trans, w, lane);
        const int L = S - 1;                       // number of weights
        // w_hat = avgpool2(maxpool2_pad1(w)) + 0.01        (length L), stored in alpha[]
        for (int i = lane; i < L; i += 32) {
            const float m0 = i == 0 ? w[0] : fmaxf(w[i - 1], w[i]);
            const float m1 = i + 1 <= L - 1 ? fmaxf(w[i], w[i + 1]) : w[L - 1];
            alpha[i] = 0.5f * (m0 + m1) + 0.01f;
        }
        __syncwarp();
        // pdf over w_hat[1 : L-1] + 1e-5 (NB = L-2 entries), cdf with leading zero (NB+1 entries) in trans[]
        const int NB = L - 2;
        {   // (normalise, then cumulative sum -- renderer.py:277-279 -- as a warp reduction and a chunked warp scan)
            float part = 0.f;
            for (int i = lane; i < NB; i += 32) part += alpha[1 + i] + 1e-5f;
            const float tot = warp_sum(part);
            float carry = 0.f;
            if (lane == 0) trans[0] = 0.f;
            for (int i0 = 0; i0 < NB; i0 += 32) {
                const int i = i0 + lane;
                float v = i < NB ? (alpha[1 + i] + 1e-5f) / tot : 0.f;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float up = __shfl_up_sync(0xffffffffu, v, o);
                    if (lane >= o) v += up;
                }
                if (i < NB) trans[1 + i] = carry + v;
                carry += __shfl_sync(0xffffffffu, v, 31);
            }
        }
        __syncwarp();
        for (int j = lane; j < S_imp; j += 32) {
            const float uu = j == lane ? u0 : (j == lane + 32 ? u1 : u[ray * S_imp + j]);
            int lo = 0, hi = NB + 1;                // searchsorted(right=True): first index with cdf > u
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (trans[mid] <= uu) lo = mid + 1; else hi = mid; }
            const int below = max(lo - 1, 0), above = min(lo, NB);
            const float c0 = trans[below], c1 = trans[above];
            const float b0 = 0.5f * (ts[below] + ts[below + 1]), b1 = 0.5f * (ts[above] + ts[above + 1]);
            float den = c1 - c0;
            if (den < 1e-5f) den = 1.f;
            t_f[ray * S_imp + j] = b0 + (uu - c0) / den * (b1 - b0);
        }
        __syncwarp();
    }
}

struct CompositeParams {
    const float* t_c; const float* sigma_c; const float* rgb_c; int S1;
    const float* t_f; const float* sigma_f; const float* rgb_f; int S2;
    const unsigned* minmax; int white_back; long n_rays;
    float* feat; float* depth; float* wsum;
    const float* d_feat; const float* d_depth; const float* d_wsum;
    float* d_rgb_c; float* d_sigma_c; float* d_rgb_f; float* d_sigma_f;
};

// Loads one ray, ranks the merged samples by depth (ties broken by original index) and fills
// ts/ss (sorted depth, density) and rk[i] = sorted position of original sample i.
// Fast path (the renderer's case): the coarse depths are already non-decreasing (stratified sampling), so only the fine
// samples are ranked against each other (S2^2 compares) and the two sorted lists are merged by binary search:
//   rank(coarse i) = i + #{fine < t_i}            rank(fine j) = rank_in_fine(j) + #{coarse <= t_j}
// -- exactly the order a stable sort of [coarse..., fine...] produces (torch.sort in renderer.py:218).  Unsorted coarse
// depths fall back to the all-pairs ranking.  `scratch` holds S2 floats (the sorted fine depths).
__device__ __forceinline__ void load_and_rank(const CompositeParams& p, long ray, float* traw, float* ts, float* ss, int* rk,
                                              float* scratch, int lane) {
    const int S1 = p.S1, S2 = p.S2, S = S1 + S2;
    for (int i = lane; i < S; i += 32) traw[i] = i < S1 ? p.t_c[ray * S1 + i] : p.t_f[ray * S2 + i - S1];
#if RAY_PF_MODE == 2 || RAY_PF_MODE == 3
    prefetch_rows(p.rgb_c + ray * S1 * 32, S1, lane);
#endif
#if RAY_PF_MODE == 2
    prefetch_rows(S2 ? p.rgb_f + ray * S2 * 32 : nullptr, S2, lane);
#endif
    __syncwarp();
    bool sorted = true;
    for (int i = lane; i + 1 < S1; i += 32) sorted = sorted && (traw[i] <= traw[i + 1]);
    sorted = __all_sync(0xffffffffu, sorted);
    if (sorted) {
        const float* tf = traw + S1;
        const bool vec4 = ((S1 | S2) & 3) == 0;               // tf is then 16-byte aligned: four depths per (broadcast) shared load
        // rank among the fine samples = #{t_k < t_j}; distinct depths fill every slot of scratch[] exactly once, so a slot that
        // still holds the NaN it starts with means two equal depths (rare): only then the stable (depth, index) order is counted
        for (int j = lane; j < S2; j += 32) scratch[j] = __int_as_float(0x7fc00000);
        __syncwarp();
        for (int j = lane; j < S2; j += 32) {
            const float tj = tf[j];
            int lt = 0;
            if (vec4) {
                for (int k = 0; k < S2; k += 4) {
                    const float4 t4 = *reinterpret_cast<const float4*>(tf + k);
                    lt += (t4.x < tj) + (t4.y < tj) + (t4.z < tj) + (t4.w < tj);
                }
            } else {
                for (int k = 0; k < S2; ++k) lt += tf[k] < tj;
            }
            rk[S1 + j] = lt;
            scratch[lt] = tj;
        }
        __syncwarp();
        bool hole = false;
        for (int j = lane; j < S2; j += 32) { const float v = scratch[j]; hole = hole || (v != v); }
        if (__any_sync(0xffffffffu, hole)) {
            __syncwarp();
            for (int j = lane; j < S2; j += 32) {
                const float tj = tf[j];
                int r = 0;
                for (int k = 0; k < S2; ++k) { const float tk = tf[k]; r += (tk < tj) || (tk == tj && k < j); }
                rk[S1 + j] = r;
                scratch[r] = tj;
            }
        }
        __syncwarp();
        for (int i = lane; i < S; i += 32) {
            const float ti = traw[i];
            const float sg = i < S1 ? p.sigma_c[ray * S1 + i] : p.sigma_f[ray * S2 + i - S1];   // in flight during the search
            int r;
            if (i < S1) {                                      // lower_bound in the sorted fine depths
                int lo = 0, hi = S2;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (scratch[mid] < ti) lo = mid + 1; else hi = mid; }
                r = i + lo;
            } else {                                           // upper_bound in the coarse depths
                int lo = 0, hi = S1;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (traw[mid] <= ti) lo = mid + 1; else hi = mid; }
                r = rk[i] + lo;
            }
            rk[i] = r;
            ts[r] = ti;
            ss[r] = sg;
        }
    } else {
        for (int i = lane; i < S; i += 32) {
            const float ti = traw[i];
            int r = 0;
            for (int j = 0; j < S; ++j) { const float tj = traw[j]; r += (tj < ti) || (tj == ti && j < i); }
            rk[i] = r;
            ts[r] = ti;
            ss[r] = i < S1 ? p.sigma_c[ray * S1 + i] : p.sigma_f[ray * S2 + i - S1];
        }
    }
    __syncwarp();
}

#ifndef RAY_MINB
#define RAY_MINB 10
#endif
#ifndef RAY_UNROLL
#define RAY_UNROLL 4
#endif
#define RAY_PRAGMA_(x) _Pragma(#x)
#define RAY_PRAGMA(x) RAY_PRAGMA_(x)
__global__ void __launch_bounds__(128, RAY_MINB) ray_composite_fwd_kernel(CompositeParams p, int SP) {
    extern __shared__ __align__(16) float dyn[];                 // per warp: 6 float arrays + the rank array, SP entries each
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* traw = dyn + wid * 7 * SP; float* ts = traw + SP; float* ss = ts + SP; float* alpha = ss + SP;
    float* trans = alpha + SP; float* w = trans + SP;
    int* rk = reinterpret_cast<int*>(w + SP);
    const int S = p.S1 + p.S2;
    const float dmin = ord2f(p.minmax[0]), dmax = ord2f(p.minmax[1]);
    for (long ray = (long)blockIdx.x * 4 + wid; ray < p.n_rays; ray += (long)gridDim.x * 4) {
#if RAY_PF_MODE == 0 || RAY_PF_MODE == 1
        prefetch_rows(p.rgb_c + ray * p.S1 * 32, p.S1, lane);
#endif
#if RAY_PF_MODE == 0
        prefetch_rows(p.S2 ? p.rgb_f + ray * p.S2 * 32 : nullptr, p.S2, lane);
#endif
        load_and_rank(p, ray, traw, ts, ss, rk, alpha, lane);
#if RAY_PF_MODE == 1 || RAY_PF_MODE == 3
        prefetch_rows(p.S2 ? p.rgb_f + ray * p.S2 * 32 : nullptr, p.S2, lane);
#endif
        march_weights(ts, ss, S, alpha, trans, w, lane);
        float ws = 0.f, dn = 0.f;
        for (int k = lane; k < S - 1; k += 32) { ws += w[k]; dn += w[k] * 0.5f * (ts[k] + ts[k + 1]); }
        ws = warp_sum(ws); dn = warp_sum(dn);
        // omega_i = (w[rank(i) - 1] + w[rank(i)]) / 2 per ORIGINAL sample, computed once (lane-parallel) into traw[] (free now)
        for (int i = lane; i < S; i += 32) {
            const int r = rk[i];
            traw[i] = 0.5f * ((r > 0 ? w[r - 1] : 0.f) + (r < S - 1 ? w[r] : 0.f));
        }
        __syncwarp();
        // colour accumulation: eight lanes share one sample row (4 channels each), four rows per warp instruction, coarse and
        // fine rows in separate loops (one base pointer each: no per-row select); the four row groups are summed by two shuffles
        const int sub = lane >> 3, j4 = lane & 7;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        {
            const float4* rc = reinterpret_cast<const float4*>(p.rgb_c + ray * p.S1 * 32) + j4;
RAY_PRAGMA(unroll RAY_UNROLL)
            for (int i = sub; i < p.S1; i += 4) {
                const float om = traw[i];
                const float4 c = __ldg(rc + i * 8);
                acc.x = fmaf(om, c.x, acc.x); acc.y = fmaf(om, c.y, acc.y); acc.z = fmaf(om, c.z, acc.z); acc.w = fmaf(om, c.w, acc.w);
            }
            const float4* rf = reinterpret_cast<const float4*>(p.rgb_f + ray * p.S2 * 32) + j4;
            const float* omf = traw + p.S1;
RAY_PRAGMA(unroll RAY_UNROLL)
            for (int i = sub; i < p.S2; i += 4) {
                const float om = omf[i];
                const float4 c = __ldg(rf + i * 8);
                acc.x = fmaf(om, c.x, acc.x); acc.y = fmaf(om, c.y, acc.y); acc.z = fmaf(om, c.z, acc.z); acc.w = fmaf(om, c.w, acc.w);
            }
        }
#pragma unroll
        for (int o = 8; o < 32; o <<= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
            acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
        }
        if (sub == 0) {
            const float wb = p.white_back ? 1.f - ws : 0.f;
            reinterpret_cast<float4*>(p.feat + ray * 32)[j4] = make_float4((acc.x + wb) * 2.f - 1.f, (acc.y + wb) * 2.f - 1.f,
                                                                           (acc.z + wb) * 2.f - 1.f, (acc.w + wb) * 2.f - 1.f);
        }
        if (lane == 0) {
            float d = dn / ws;
            d = (d != d) ? dmax : fminf(fmaxf(d, dmin), dmax);
            p.depth[ray] = d;
            p.wsum[ray] = ws;
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(128, RAY_MINB) ray_composite_bwd_kernel(CompositeParams p, int SP) {
    extern __shared__ __align__(16) float dyn[];                 // per warp: 8 float arrays + the rank array, SP entries each
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* traw = dyn + wid * 9 * SP; float* ts = traw + SP; float* ss = ts + SP; float* alpha = ss + SP;
    float* trans = alpha + SP; float* w = trans + SP; float* dom = w + SP; float* dsb = dom + SP;
    int* rk = reinterpret_cast<int*>(dsb + SP);
    const int S = p.S1 + p.S2;
    const float dmin = ord2f(p.minmax[0]), dmax = ord2f(p.minmax[1]);
    for (long ray = (long)blockIdx.x * 4 + wid; ray < p.n_rays; ray += (long)gridDim.x * 4) {
#if RAY_PF_MODE == 0 || RAY_PF_MODE == 1
        prefetch_rows(p.rgb_c + ray * p.S1 * 32, p.S1, lane);
#endif
#if RAY_PF_MODE == 0
        prefetch_rows(p.S2 ? p.rgb_f + ray * p.S2 * 32 : nullptr, p.S2, lane);
#endif
        // the ray's incoming gradients: loaded now, used after the march (nothing below depends on them until then)
        const int sub = lane >> 3, j4 = lane & 7;
        float4 g4 = __ldg(reinterpret_cast<const float4*>(p.d_feat + ray * 32) + j4);
        const float gD_in = p.d_depth ? __ldg(p.d_depth + ray) : 0.f;
        const float gW = p.d_wsum ? __ldg(p.d_wsum + ray) : 0.f;
        load_and_rank(p, ray, traw, ts, ss, rk, alpha, lane);
#if RAY_PF_MODE == 1 || RAY_PF_MODE == 3
        prefetch_rows(p.S2 ? p.rgb_f + ray * p.S2 * 32 : nullptr, p.S2, lane);
#endif
        march_weights(ts, ss, S, alpha, trans, w, lane);
        float ws = 0.f, dn = 0.f;
        for (int k = lane; k < S - 1; k += 32) { ws += w[k]; dn += w[k] * 0.5f * (ts[k] + ts[k + 1]); }
        ws = warp_sum(ws); dn = warp_sum(dn);
        const float D = dn / ws;
        const bool d_pass = (D == D) && D >= dmin && D <= dmax;
        const float gD = d_pass ? gD_in : 0.f;
        // through rgb*2-1.  Eight lanes share one sample row (4 channels each): every warp instruction reads / writes four full
        // 128-byte colour rows; d omega_rank(i) = sum_ch g[ch] * c_i[ch] is reduced over the 8 lanes with three shuffles.
        g4.x *= 2.f; g4.y *= 2.f; g4.z *= 2.f; g4.w *= 2.f;
        float gsum = 0.f;                                        // through + 1 - wsum
        if (p.white_back) {
            gsum = g4.x + g4.y + g4.z + g4.w;
            gsum += __shfl_xor_sync(0xffffffffu, gsum, 1); gsum += __shfl_xor_sync(0xffffffffu, gsum, 2); gsum += __shfl_xor_sync(0xffffffffu, gsum, 4);
        }
        // omega_i = (w[rank(i) - 1] + w[rank(i)]) / 2 per ORIGINAL sample, once, lane-parallel (into traw[], free now)
        for (int i = lane; i < S; i += 32) {
            const int r = rk[i];
            traw[i] = 0.5f * ((r > 0 ? w[r - 1] : 0.f) + (r < S - 1 ? w[r] : 0.f));
        }
        __syncwarp();
        // coarse rows, then fine rows (one base pointer per loop, no per-row select); the dot products land in dsb[] by ORIGINAL
        // index and are permuted into rank order afterwards, lane-parallel -- the row loop carries no rank arithmetic
#pragma unroll
        for (int part = 0; part < 2; ++part) {
            const int cnt = part ? p.S2 : p.S1, off = part ? p.S1 : 0;
            const long row0 = part ? ray * p.S2 * 32 : ray * p.S1 * 32;
            const float4* src = reinterpret_cast<const float4*>((part ? p.rgb_f : p.rgb_c) + row0) + j4 + sub * 8;
            float4* dst = reinterpret_cast<float4*>((part ? p.d_rgb_f : p.d_rgb_c) + row0) + j4 + sub * 8;
            const float* omp = traw + off + sub;
            float* tdp = dsb + off + sub;
            const int full = cnt & ~3;
#pragma unroll 4
            for (int i0 = 0; i0 < full; i0 += 4) {
                const float om = omp[i0];
                const float4 c = __ldg(src + i0 * 8);
                float t = g4.x * c.x + g4.y * c.y + g4.z * c.z + g4.w * c.w;
                dst[i0 * 8] = make_float4(om * g4.x, om * g4.y, om * g4.z, om * g4.w);
                t += __shfl_xor_sync(0xffffffffu, t, 1); t += __shfl_xor_sync(0xffffffffu, t, 2); t += __shfl_xor_sync(0xffffffffu, t, 4);
                if (j4 == 0) tdp[i0] = t;
            }
            if (full < cnt) {                                   // sample counts that are not a multiple of four: one partial group
                const bool v = full + sub < cnt;
                float t = 0.f;
                if (v) {
                    const float om = omp[full];
                    const float4 c = __ldg(src + full * 8);
                    t = g4.x * c.x + g4.y * c.y + g4.z * c.z + g4.w * c.w;
                    dst[full * 8] = make_float4(om * g4.x, om * g4.y, om * g4.z, om * g4.w);
                }
                t += __shfl_xor_sync(0xffffffffu, t, 1); t += __shfl_xor_sync(0xffffffffu, t, 2); t += __shfl_xor_sync(0xffffffffu, t, 4);
                if (v && j4 == 0) tdp[full] = t;
            }
        }
        __syncwarp();
        for (int i = lane; i < S; i += 32) dom[rk[i]] = dsb[i];
        __syncwarp();
        // d w_k, stored in dsb[]
        for (int k = lane; k < S - 1; k += 32) {
            float dw = 0.5f * (dom[k] + dom[k + 1]) - gsum + gW;
            if (gD != 0.f) dw += gD * (0.5f * (ts[k] + ts[k + 1]) - D) / ws;
            dsb[k] = dw;
        }
        __syncwarp();
        // reverse scan: d alpha_k (into dom[]).  x_k = dw_k alpha_k + x_{k+1} (1 - alpha_k + 1e-10), x_{S-1} = 0 (the gradient flowing
        // into the transmittance ahead of interval k), d alpha_k = T_k (dw_k - x_{k+1}).  Each interval is the affine map
        // x -> A + B x; a warp-wide suffix scan composes the maps of 32 intervals, chunks run from the far end of the ray
        {
            float carry = 0.f;
            for (int k0 = ((S - 2) >> 5) << 5; k0 >= 0; k0 -= 32) {
                const int k = k0 + lane;
                const bool v = k < S - 1;
                const float dw = v ? dsb[k] : 0.f, al = v ? alpha[k] : 0.f;
                float A = dw * al, B = v ? (1.f - al + 1e-10f) : 1.f;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float A2 = __shfl_down_sync(0xffffffffu, A, o), B2 = __shfl_down_sync(0xffffffffu, B, o);
                    if (lane + o < 32) { A = fmaf(B, A2, A); B *= B2; }
                }
                const float x = fmaf(B, carry, A);                         // x_k
                float xn = __shfl_down_sync(0xffffffffu, x, 1);            // x_{k+1}
                if (lane == 31) xn = carry;
                if (v) dom[k] = trans[k] * (dw - xn);
                carry = __shfl_sync(0xffffffffu, x, 0);
            }
        }
        __syncwarp();
        // d sigma_mid_k (into dsb[])
        for (int k = lane; k < S - 1; k += 32) {
            const float delta = ts[k + 1] - ts[k];
            const float x = 0.5f * (ss[k] + ss[k + 1]) - 1.f;
            const float sp = softplus_f(x);
            const float dsp = dom[k] * delta * expf(-sp * delta);
            dsb[k] = dsp * (x > 20.f ? 1.f : sigmoid_f(x));
        }
        __syncwarp();
        for (int i = lane; i < S; i += 32) {
            const int r = rk[i];
            const float ds = 0.5f * ((r > 0 ? dsb[r - 1] : 0.f) + (r < S - 1 ? dsb[r] : 0.f));
            if (i < p.S1) p.d_sigma_c[ray * p.S1 + i] = ds; else p.d_sigma_f[ray * p.S2 + i - p.S1] = ds;
        }
        __syncwarp();
    }
}
// ---------------------------------------------------------------------------------------------------------------------
// RaySampler.forward (training/volumetric_rendering/ray_sampler.py:24-73): pixel-centre uv, unprojection through the
// intrinsics, rotation into world space, normalisation.  One thread per ray; the backward reduces the 12 cam2world
// gradients per sample (w-projection optimises the pose through them).
template <bool BWD>
__global__ void ray_sampler_kernel(const float* __restrict__ c2w, const float* __restrict__ K, int R, float* __restrict__ ray_o,
                                   float* __restrict__ ray_d, const float* __restrict__ d_o, const float* __restrict__ d_d,
                                   float* __restrict__ d_c2w) {
    __shared__ float red[32];
    const int b = blockIdx.y, M = R * R;
    const float* E = c2w + b * 16;
    const float* Kb = K + b * 9;
    const float fx = Kb[0], sk = Kb[1], cx = Kb[2], fy = Kb[4], cy = Kb[5];
    float acc[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) acc[j] = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        const float xc = (float)(i % R) * (1.f / R) + 0.5f / R, yc = (float)(i / R) * (1.f / R) + 0.5f / R;
        const float xl = (xc - cx + cy * sk / fy - sk * yc / fy) / fx, yl = (yc - cy) / fy;
        float q[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) q[j] = (E[4 * j] * xl + E[4 * j + 1] * yl + E[4 * j + 2] + E[4 * j + 3]) - E[4 * j + 3];
        const float qn = fmaxf(sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]), 1e-12f);
        const long o = ((long)b * M + i) * 3;
        if (!BWD) {
#pragma unroll
            for (int j = 0; j < 3; ++j) { ray_o[o + j] = E[4 * j + 3]; ray_d[o + j] = q[j] / qn; }
        } else {
            float dd[3] = {0.f, 0.f, 0.f}, dot = 0.f, dir[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) { dir[j] = q[j] / qn; if (d_d) dd[j] = d_d[o + j]; dot += dir[j] * dd[j]; }
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float dq = (dd[j] - dir[j] * dot) / qn;
                acc[4 * j] += dq * xl; acc[4 * j + 1] += dq * yl; acc[4 * j + 2] += dq;
                if (d_o) acc[4 * j + 3] += d_o[o + j];
            }
        }
    }
    if (BWD) {
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            float v = warp_sum(acc[j]);
            __syncthreads();
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
            __syncthreads();
            if (threadIdx.x < 32) {
                v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
                v = warp_sum(v);
                if (threadIdx.x == 0) atomicAdd(d_c2w + b * 16 + j, v);
            }
        }
    }
}
}  // namespace

// minmax (may be NULL): the two words of b200_depth_minmax, updated with this call's depths (saves the separate pass over t)
B200_API int b200_ray_depths_coarse(const float* t_base, const float* u, float* t, long n_rays, int S, float delta,
                                    unsigned* minmax, void* stream) {
    const long total = n_rays * S;
    if (total <= 0) return 0;
    const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    depths_coarse_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(t_base, u, t, total, S, delta, minmax);
    B200_CHECK_LAUNCH();
    return 0;
}

// minmax: two uint32 in order-preserving encoding, initialised by the caller to {0xFFFFFFFF, 0}.
B200_API int b200_depth_minmax(const float* t, long total, unsigned* minmax, void* stream) {
    if (total <= 0) return 0;
    const int blocks = (int)((total + 255) / 256 < 148 * 4 ? (total + 255) / 256 : 148 * 4);
    depth_minmax_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(t, total, minmax);
    B200_CHECK_LAUNCH();
    return 0;
}

B200_API int b200_ray_importance(const float* t_c, const float* sigma_c, const float* u, float* t_f, long n_rays, int S,
                                 int S_imp, void* stream) {
    B200_REQUIRE(S >= 4 && S <= MAXS, "ray_importance: need 4 <= depth_resolution <= 256");
    if (n_rays <= 0 || S_imp <= 0) return 0;
    const int SP = (S + 3) & ~3;
    const int blocks = (int)((n_rays + 3) / 4 < (1L << 30) ? (n_rays + 3) / 4 : (1L << 30));   // one ray per warp: the block scheduler balances the SMs
    ray_importance_kernel<<<blocks, 128, 4 * 5 * SP * sizeof(float), (cudaStream_t)stream>>>(t_c, sigma_c, u, t_f, n_rays, S, S_imp, SP);
    B200_CHECK_LAUNCH();
    return 0;
}

B200_API int b200_ray_composite_fwd(const float* t_c, const float* sigma_c, const float* rgb_c, int S1, const float* t_f,
                                    const float* sigma_f, const float* rgb_f, int S2, const unsigned* minmax,
                                    int white_back, long n_rays, float* feat, float* depth, float* wsum, void* stream) {
    B200_REQUIRE(S1 >= 1 && S2 >= 0 && S1 + S2 >= 2 && S1 + S2 <= MAXS, "ray_composite: need 2 <= samples per ray <= 256");
    B200_REQUIRE(((((uintptr_t)rgb_c | (uintptr_t)rgb_f | (uintptr_t)feat) & 15) == 0), "ray_composite: colour rows must be 16-byte aligned");
    if (n_rays <= 0) return 0;
    CompositeParams p{};
    p.t_c = t_c; p.sigma_c = sigma_c; p.rgb_c = rgb_c; p.S1 = S1; p.t_f = t_f; p.sigma_f = sigma_f; p.rgb_f = rgb_f; p.S2 = S2;
    p.minmax = minmax; p.white_back = white_back; p.n_rays = n_rays; p.feat = feat; p.depth = depth; p.wsum = wsum;
    const int SP = (S1 + S2 + 3) & ~3;
    const int blocks = (int)((n_rays + 3) / 4 < (1L << 30) ? (n_rays + 3) / 4 : (1L << 30));   // one ray per warp: the block scheduler balances the SMs
    ray_composite_fwd_kernel<<<blocks, 128, 4 * 7 * SP * sizeof(float), (cudaStream_t)stream>>>(p, SP);
    B200_CHECK_LAUNCH();
    return 0;
}

B200_API int b200_ray_composite_bwd(const float* t_c, const float* sigma_c, const float* rgb_c, int S1, const float* t_f,
                                    const float* sigma_f, const float* rgb_f, int S2, const unsigned* minmax,
                                    int white_back, long n_rays, const float* d_feat, const float* d_depth,
                                    const float* d_wsum, float* d_rgb_c, float* d_sigma_c, float* d_rgb_f, float* d_sigma_f,
                                    void* stream) {
    B200_REQUIRE(S1 >= 1 && S2 >= 0 && S1 + S2 >= 2 && S1 + S2 <= MAXS, "ray_composite: need 2 <= samples per ray <= 256");
    B200_REQUIRE(d_feat, "ray_composite_bwd: d_feat is required");
    B200_REQUIRE(((((uintptr_t)rgb_c | (uintptr_t)rgb_f | (uintptr_t)d_feat | (uintptr_t)d_rgb_c | (uintptr_t)d_rgb_f) & 15) == 0),
                 "ray_composite_bwd: colour rows must be 16-byte aligned");
    if (n_rays <= 0) return 0;
    CompositeParams p{};
    p.t_c = t_c; p.sigma_c = sigma_c; p.rgb_c = rgb_c; p.S1 = S1; p.t_f = t_f; p.sigma_f = sigma_f; p.rgb_f = rgb_f; p.S2 = S2;
    p.minmax = minmax; p.white_back = white_back; p.n_rays = n_rays;
    p.d_feat = d_feat; p.d_depth = d_depth; p.d_wsum = d_wsum;
    p.d_rgb_c = d_rgb_c; p.d_sigma_c = d_sigma_c; p.d_rgb_f = d_rgb_f; p.d_sigma_f = d_sigma_f;
    const int SP = (S1 + S2 + 3) & ~3;
    const int blocks = (int)((n_rays + 3) / 4 < (1L << 30) ? (n_rays + 3) / 4 : (1L << 30));   // one ray per warp: the block scheduler balances the SMs
    ray_composite_bwd_kernel<<<blocks, 128, 4 * 9 * SP * sizeof(float), (cudaStream_t)stream>>>(p, SP);
    B200_CHECK_LAUNCH();
    return 0;
}

// RaySampler.forward: cam2world [n][16], intrinsics [n][9] -> ray_o, ray_d [n][R*R][3] (ray m = i*R + j at pixel centre ((j+.5)/R, (i+.5)/R)).
B200_API int b200_ray_sampler_fwd(const float* cam2world, const float* intrinsics, int n, int R, float* ray_o, float* ray_d, void* stream) {
    B200_REQUIRE(cam2world && intrinsics && ray_o && ray_d && n > 0 && R > 0, "ray_sampler_fwd: null pointer or bad shape");
    const int bx = (R * R + 255) / 256 < 148 ? (R * R + 255) / 256 : 148;
    ray_sampler_kernel<false><<<dim3(bx, n), 256, 0, (cudaStream_t)stream>>>(cam2world, intrinsics, R, ray_o, ray_d, nullptr, nullptr, nullptr);
    B200_CHECK_LAUNCH();
    return 0;
}

// d_cam2world [n][16] is ACCUMULATED into (zero it first; row 3 stays zero); d_ray_o / d_ray_d may each be NULL.
B200_API int b200_ray_sampler_bwd(const float* cam2world, const float* intrinsics, int n, int R, const float* d_ray_o,
                                  const float* d_ray_d, float* d_cam2world, void* stream) {
    B200_REQUIRE(cam2world && intrinsics && d_cam2world && n > 0 && R > 0, "ray_sampler_bwd: null pointer or bad shape");
    const int bx = (R * R + 255) / 256 < 148 ? (R * R + 255) / 256 : 148;
    ray_sampler_kernel<true><<<dim3(bx, n), 256, 0, (cudaStream_t)stream>>>(cam2world, intrinsics, R, nullptr, nullptr, d_ray_o, d_ray_d, d_cam2world);
    B200_CHECK_LAUNCH();
    return 0;
}
