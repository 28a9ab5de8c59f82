// Fused tri-plane sampling + OSG decoder MLP (+ its backward).
//
// Restates, as one kernel per direction, what the reference does with grid_sample -> mean -> FC(32,64)
// -> softplus -> FC(64,33) -> sigmoid (training/volumetric_rendering/renderer.py:39-66,
// training/triplane.py:124-136).  Planes are read in NHWC [n][H][W][96] (plane p = channels 32p..32p+31,
// one 128-byte line per texel and plane), so a warp gathers one texel with one coalesced request.
//
// Work decomposition: a warp owns 32 consecutive sample points.
//   gather / scatter : 8 lanes per point (4 channels each), one warp instruction touches four full 128 B texel lines
//   decoder MLP      : mma.sync m16n8k8 TF32 tiles over the 32 points (see "Tensor-core decoder" below)
// The phases are stitched through small per-warp shared-memory tiles.
#include "common.cuh"
#include "tc_common.cuh"
#include <cstdlib>
#include <cuda_bf16.h>

#include "triplane_common.cuh"
namespace {
using namespace tri;
// ---------------------------------------------------------------------------------------------------------------------
// Tensor-core decoder: the two FC layers of a warp's 32 points as mma.sync.m16n8k8 TF32 tiles with split-float operands
// (x = hi + lo, three MMAs per product: lo*hi + hi*lo + hi*hi, fp32 accumulate -> ~fp32 accuracy).
//   layer 1: H[32x64] = F[32x32] * W1^T   (2 m-tiles x 8 n-tiles x 4 k-steps)
//   layer 2: O[32x40] = H[32x64] * W2^T   (2 m-tiles x 5 n-tiles x 8 k-steps; rows 33..39 of W2 are zero padding)
// Fragment layout (PTX ISA, m16n8k8 .tf32): g = lane/4, t = lane%4
//   A: a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4)     B: b0 (k=t, n=g) b1 (k=t+4, n=g)     C: c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
// Shared-memory strides 36 / 68 (= 4 mod 32) make every fragment load bank-conflict free.
// The hidden layer never touches shared memory: layer 1's C fragment gives lane (g,t) the hidden units {8*nt + 2t, 8*nt + 2t+1}
// of rows g / g+8, and layer 2 consumes exactly these as its A fragment (a0 = c0, a1 = c2, a2 = c1, a3 = c3) when the k-slots
// (t, t+4) of k-step nt are DEFINED to be those two units -- i.e. layer 2's B fragment reads W2[n][8*ks + 2t], W2[n][8*ks + 2t+1].
constexpr int SF = 36, SW2 = 72, OUTP = 40;
constexpr int FW_W = 2 * HID * SF + 2 * OUTP * SW2 + HID + OUTP;          // W1 hi|lo, W2 hi|lo, b1, b2 (floats)
constexpr int FW_WARP = 32 * SF + 32 * SP;
#ifndef B200_FW_WARPS
#define B200_FW_WARPS 16
#endif
constexpr int FW_WARPS = B200_FW_WARPS;
constexpr int FW_SMEM = (FW_W + FW_WARPS * FW_WARP) * 4;

__device__ __forceinline__ uint32_t tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
    hi = tf32_rna(x);
    lo = tf32_rna(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// W1 -> [64][SF] hi / lo (tf32 bit patterns), W2 -> [OUTP][SW2] hi / lo with zero rows 33..39, biases
__device__ __forceinline__ void load_weights_split(const TriplaneParams& p, float* W1h, float* W1l, float* W2h, float* W2l,
                                                   float* b1s, float* b2s) {
    for (int i = threadIdx.x; i < HID * C; i += blockDim.x) {
        uint32_t h, l;
        tf32_split(p.W1[i] * p.w1g, h, l);
        const int o = (i / C) * SF + (i % C);
        W1h[o] = __uint_as_float(h); W1l[o] = __uint_as_float(l);
    }
    for (int i = threadIdx.x; i < OUTP * HID; i += blockDim.x) {
        const int k = i / HID, j = i % HID;
        uint32_t h = 0, l = 0;
        if (k < OUT) tf32_split(p.W2[i] * p.w2g, h, l);
        W2h[k * SW2 + j] = __uint_as_float(h); W2l[k * SW2 + j] = __uint_as_float(l);
    }
    for (int i = threadIdx.x; i < HID; i += blockDim.x) b1s[i] = p.b1[i] * p.b1g;
    for (int i = threadIdx.x; i < OUTP; i += blockDim.x) b2s[i] = i < OUT ? p.b2[i] * p.b2g : 0.f;
}

// SIGMA_ONLY: density queries on voxel grids (single_id_coach.py:118-140) need column 0 of the second layer only.
template <bool SIGMA_ONLY>
__global__ void __launch_bounds__(FW_WARPS * 32, 1) triplane_mlp_fwd_mma_kernel(TriplaneParams p) {
    constexpr int NT2 = SIGMA_ONLY ? 1 : 5;
    extern __shared__ __align__(16) float smem[];
    float* W1h = smem; float* W1l = W1h + HID * SF; float* W2h = W1l + HID * SF; float* W2l = W2h + OUTP * SW2;
    float* b1s = W2l + OUTP * SW2; float* b2s = b1s + HID;
    const int n = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    float* sF = smem + FW_W + wid * FW_WARP;     // [32][SF]  features, later the output staging tile
    float* ss = sF + 32 * SF;                    // [32][SP]  bilinear set-up
    load_weights_split(p, W1h, W1l, W2h, W2l, b1s, b2s);
    __syncthreads();
    const float* pl = p.planes + (long)n * p.hp * p.wp * PC;
    for (long base = ((long)blockIdx.x * FW_WARPS + wid) * 32; base < p.P; base += (long)gridDim.x * (FW_WARPS * 32)) {
        const long pi = base + lane;
        float cx, cy, cz;
        point_coords(p, n, pi, cx, cy, cz);
        stage_setup(ss, lane, cx, cy, cz, p.hp, p.wp);
        __syncwarp();
        gather_features<SF>(pl, ss, sF, lane);
        __syncwarp();
        // ---- layer 1: 16 independent accumulator tiles (2 m-tiles x 8 n-tiles), k-steps outermost
        float c[2][8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float bv0 = b1s[8 * nt + 2 * t], bv1 = b1s[8 * nt + 2 * t + 1];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) { c[mt][nt][0] = c[mt][nt][2] = bv0; c[mt][nt][1] = c[mt][nt][3] = bv1; }
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t ah[2][4], al[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const float* r0 = sF + (g + 16 * mt) * SF + t + 8 * ks;
                tf32_split(r0[0], ah[mt][0], al[mt][0]);
                tf32_split(r0[8 * SF], ah[mt][1], al[mt][1]);
                tf32_split(r0[4], ah[mt][2], al[mt][2]);
                tf32_split(r0[8 * SF + 4], ah[mt][3], al[mt][3]);
            }
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const int bo = (g + 8 * nt) * SF + t + 8 * ks;
                const uint32_t bh0 = __float_as_uint(W1h[bo]), bh1 = __float_as_uint(W1h[bo + 4]);
                const uint32_t bl0 = __float_as_uint(W1l[bo]), bl1 = __float_as_uint(W1l[bo + 4]);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    if (p.fwd_passes == 3) {
                        mma_tf32(c[mt][nt], al[mt], bh0, bh1);
                        mma_tf32(c[mt][nt], ah[mt], bl0, bl1);
                    }
                    mma_tf32(c[mt][nt], ah[mt], bh0, bh1);
                }
            }
        }
        __syncwarp();                            // all lanes are done reading the feature tile
        // ---- layer 2: A fragments come straight from layer 1's accumulators (see the note above)
        float o[2][NT2][4];
#pragma unroll
        for (int nt = 0; nt < NT2; ++nt) {
            const float bv0 = b2s[8 * nt + 2 * t], bv1 = b2s[8 * nt + 2 * t + 1];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) { o[mt][nt][0] = o[mt][nt][2] = bv0; o[mt][nt][1] = o[mt][nt][3] = bv1; }
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            uint32_t hh[2][4], hl[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                tf32_split(softplus_fast(c[mt][ks][0]), hh[mt][0], hl[mt][0]);     // (row g,   unit 8ks+2t)
                tf32_split(softplus_fast(c[mt][ks][2]), hh[mt][1], hl[mt][1]);     // (row g+8, unit 8ks+2t)
                tf32_split(softplus_fast(c[mt][ks][1]), hh[mt][2], hl[mt][2]);     // (row g,   unit 8ks+2t+1)
                tf32_split(softplus_fast(c[mt][ks][3]), hh[mt][3], hl[mt][3]);     // (row g+8, unit 8ks+2t+1)
            }
#pragma unroll
            for (int nt = 0; nt < NT2; ++nt) {
                const int bo = (g + 8 * nt) * SW2 + 8 * ks + 2 * t;
                const float2 bh = *reinterpret_cast<const float2*>(&W2h[bo]);
                const float2 bl = *reinterpret_cast<const float2*>(&W2l[bo]);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    if (p.fwd_passes == 3) {
                        mma_tf32(o[mt][nt], hl[mt], __float_as_uint(bh.x), __float_as_uint(bh.y));
                        mma_tf32(o[mt][nt], hh[mt], __float_as_uint(bl.x), __float_as_uint(bl.y));
                    }
                    mma_tf32(o[mt][nt], hh[mt], __float_as_uint(bh.x), __float_as_uint(bh.y));
                }
            }
        }
        // ---- epilogue: column 0 = sigma (linear), columns 1..32 = rgb = sigmoid(o)*1.002 - 0.001; staged in sF
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT2; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = g + 16 * mt + ((i & 2) ? 8 : 0), col = 8 * nt + 2 * t + (i & 1);
                    if (col < OUT) sF[row * SF + col] = col == 0 ? o[mt][nt][i] : sigmoid_fast(o[mt][nt][i]) * 1.002f - 0.001f;
                }
        __syncwarp();
        if (pi < p.P) p.sigma[(long)n * p.P + pi] = sF[lane * SF];
        if (!SIGMA_ONLY) {
            const int cnt = (int)min((long)32, p.P - base);
            float* out = p.rgb + ((long)n * p.P + base) * C;
            for (int q = 0; q < cnt; ++q) out[q * C + lane] = sF[q * SF + 1 + lane];
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Tensor-core backward of the fused sampler + decoder (mma.sync m16n8k8 TF32, single pass: gradients need ~1e-3 relative
// accuracy, the forward keeps the 3-pass split).  Per warp and 32 points:
//   recompute   C1 = F W1^T (+b1), h = softplus(C1);  O = h W2^T (+b2)             [h feeds layer 2 from registers]
//   d_out       dO = d_rgb * 1.002 * s(1-s) | d_sigma                               [C-fragment layout of O]
//   chain       dh = dO W2  -> d_a = dh * (1 - exp(-h)) -> d_f = d_a W1             [A fragments straight from registers]
//   scatter     d_f (x bilinear weights, 1/3 folded in) -> red.global.add.v4 into the plane gradient
//   parameters  db1 / db2 reduced here.  dW1 = d_a^T F and dW2 = dO^T h are contractions over the POINT axis: every warp
//               stages its 32-point operand tiles as bf16 in shared memory (MN-major, 128-byte swizzle, the layout the
//               tcgen05 weight-gradient kernel of conv_tc.cu gets from TMA) and one lane issues tcgen05.mma into a
//               CTA-wide fp32 accumulator in TMEM -- D[j][0:48] += h^T dO, D[j][48:80] += d_a^T F -- so the reduction over
//               all points of the CTA costs no registers, no atomics and no HBM round trip.  The accumulator is drained
//               once at the end of the kernel (148 x 4160 global atomics).
// Two instantiations: with decoder-parameter gradients (PTI: 12 warps, the operand tiles take 8 KB per warp) and without
// (w-projection: 16 warps, no operand tiles, more shared memory left to the L1).
#ifndef B200_BM_WARPS
#define B200_BM_WARPS 12
#endif
#ifndef B200_BWD_GATHER_UNROLL
#define B200_BWD_GATHER_UNROLL 2      // texel loads in flight per lane = 12 x this
#endif
constexpr int BM_WARPS_WG = B200_BM_WARPS, BM_WARPS_NOWG = 16;
constexpr int BM_W = HID * SF + OUTP * SW2 + HID + OUTP + HID + OUTP;       // W1 [64][36], W2 [40][72], b1, b2, db1, db2
constexpr int BM_WPAD = (BM_W + 3) / 4 * 4;
constexpr int BM_WARP = 32 * SF + 32 * SP;                                  // f (later d_f) | set-up
constexpr int BM_OPER = 8192;                                               // per warp: A tile [32 pts][64] bf16 | B tile [32 pts][<=64] bf16
constexpr int bm_smem(int warps, bool wg) { return (wg ? 1024 + warps * BM_OPER + 8 * warps + 16 : 0) + (BM_WPAD + warps * BM_WARP) * 4; }
constexpr int DW2_COLS = 48, DW1_COL0 = 48, DW_TCOLS = 128;                 // TMEM columns: [0,48) dW2^T, [48,80) dW1

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

// sum over the 8 row-groups (lane bits 2..4) of a C-fragment register; result valid in lanes with g == 0
__device__ __forceinline__ float sum_over_g(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    return v;
}

template <int BM_WARPS, bool WGRAD>
__global__ void __launch_bounds__(BM_WARPS * 32, 1) triplane_mlp_bwd_mma_kernel(TriplaneParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t raw = tc::smem_u32(smem_raw);
    uint8_t* smem_al = smem_raw + (WGRAD ? ((raw + 1023u) & ~1023u) - raw : 0u);   // swizzle-128B operand tiles need 1024-byte alignment
    float* smem = reinterpret_cast<float*>(smem_al + (WGRAD ? BM_WARPS * BM_OPER : 0));
    float* W1s = smem; float* W2s = W1s + HID * SF; float* b1s = W2s + OUTP * SW2; float* b2s = b1s + HID;
    float* ab1 = b2s + OUTP; float* ab2 = ab1 + HID;
    const int n = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    float* sF = smem + BM_WPAD + wid * BM_WARP;                // [32][SF]
    float* ss = sF + 32 * SF;                                  // [32][SP]
    constexpr bool wgrad = WGRAD;
    // decoder weight gradients: per-warp operand tiles, one mbarrier per warp, CTA-wide TMEM accumulator
    uint8_t* opA = smem_al + wid * BM_OPER;                    // [32 pts][64] bf16: h, later d_a          (M side)
    uint8_t* opB = opA + 4096;                                 // [32 pts][64] bf16: dO (48 used), later F (N side)
    const uint32_t opA_u = tc::smem_u32(opA), opB_u = opA_u + 4096;
    uint8_t* bar_base = reinterpret_cast<uint8_t*>(smem + BM_WPAD + BM_WARPS * BM_WARP);
    const uint32_t mybar = tc::smem_u32(bar_base) + 8u * wid;
    const uint32_t tmem_slot = tc::smem_u32(bar_base) + 8u * BM_WARPS;
    uint32_t ncommit = 0;                                      // tcgen05.commit count of this warp (mbarrier phase bookkeeping)
    uint32_t tmem = 0;
    if (wgrad) {
        if (threadIdx.x == 0) {
            for (int w = 0; w < BM_WARPS; ++w) tc::mbar_init(tc::smem_u32(bar_base) + 8u * w, 1);
            tc::fence_barrier_init();
        }
        for (int i = lane; i < 4096 / 16; i += 32) reinterpret_cast<uint4*>(opB)[i] = make_uint4(0u, 0u, 0u, 0u);   // dO columns 40..47 stay zero
        if (wid == 0) tc::tmem_alloc(tmem_slot, DW_TCOLS);
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        tmem = *reinterpret_cast<volatile uint32_t*>(bar_base + 8 * BM_WARPS);
        if (wid < 2) {                                         // zero rows 0..63 (TMEM lanes of warps 0, 1) of the 80 accumulator columns
            for (int c0 = 0; c0 < 80; c0 += 16) tc::tmem_st16_zero(tmem + ((uint32_t)(wid * 32) << 16) + (uint32_t)c0);
            tc::tmem_wait_st();
        }
        tc::tc_fence_before();
    }
    for (int i = threadIdx.x; i < HID * C; i += blockDim.x) W1s[(i / C) * SF + (i % C)] = __uint_as_float(tf32_rna(p.W1[i] * p.w1g));
    for (int i = threadIdx.x; i < OUTP * HID; i += blockDim.x) {
        const int k = i / HID, j = i % HID;
        W2s[k * SW2 + j] = k < OUT ? __uint_as_float(tf32_rna(p.W2[i] * p.w2g)) : 0.f;
    }
    for (int i = threadIdx.x; i < HID; i += blockDim.x) { b1s[i] = p.b1[i] * p.b1g; ab1[i] = 0.f; }
    for (int i = threadIdx.x; i < OUTP; i += blockDim.x) { b2s[i] = i < OUT ? p.b2[i] * p.b2g : 0.f; ab2[i] = 0.f; }
    __syncthreads();
    const float* pl = p.planes + (long)n * p.hp * p.wp * PC;
    float* dpl = p.d_planes ? p.d_planes + (long)n * p.hp * p.wp * PC : nullptr;
    const long row0 = (long)n * p.P;

    for (long base = ((long)blockIdx.x * BM_WARPS + wid) * 32; base < p.P; base += (long)gridDim.x * (BM_WARPS * 32)) {
        const long pi = base + lane;
        const int cnt = (int)min((long)32, p.P - base);
        float cx, cy, cz;
        point_coords(p, n, pi, cx, cy, cz);
        stage_setup(ss, lane, cx, cy, cz, p.hp, p.wp);
        __syncwarp();
        gather_features<SF, B200_BWD_GATHER_UNROLL>(pl, ss, sF, lane);
        __syncwarp();
        if (wgrad && ncommit) tc::mbar_wait(mybar, (ncommit - 1) & 1);      // last iteration's dW1 MMAs have consumed the operand tiles
        // ---- layer 1 (recompute): c1 = b1 + F W1^T, then h = softplus(c1) kept in the accumulator registers
        float c1[2][8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float bv0 = b1s[8 * nt + 2 * t], bv1 = b1s[8 * nt + 2 * t + 1];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) { c1[mt][nt][0] = c1[mt][nt][2] = bv0; c1[mt][nt][1] = c1[mt][nt][3] = bv1; }
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t a[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const float* r0 = sF + (g + 16 * mt) * SF + t + 8 * ks;
                a[mt][0] = tf32_rna(r0[0]); a[mt][1] = tf32_rna(r0[8 * SF]); a[mt][2] = tf32_rna(r0[4]); a[mt][3] = tf32_rna(r0[8 * SF + 4]);
            }
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const int bo = (g + 8 * nt) * SF + t + 8 * ks;
                const uint32_t b0 = __float_as_uint(W1s[bo]), b1 = __float_as_uint(W1s[bo + 4]);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) mma_tf32(c1[mt][nt], a[mt], b0, b1);
            }
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                for (int i = 0; i < 4; ++i) c1[mt][nt][i] = softplus_fast(c1[mt][nt][i]);
                if (wgrad) {       // h as bf16 into the A tile: element (point, unit) at point*128 + ((unit/8 ^ point%8) * 16) + (unit%8)*2
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int row = g + 16 * mt + 8 * hh;
                        *reinterpret_cast<uint32_t*>(opA + row * 128 + ((nt ^ g) << 4) + 4 * t) = pack_bf16(c1[mt][nt][2 * hh], c1[mt][nt][2 * hh + 1]);
                    }
                }
            }
        // ---- layer 2 (recompute) from registers
        float o[2][5][4];
#pragma unroll
        for (int nt = 0; nt < 5; ++nt) {
            const float bv0 = b2s[8 * nt + 2 * t], bv1 = b2s[8 * nt + 2 * t + 1];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) { o[mt][nt][0] = o[mt][nt][2] = bv0; o[mt][nt][1] = o[mt][nt][3] = bv1; }
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            uint32_t a[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                a[mt][0] = tf32_rna(c1[mt][ks][0]); a[mt][1] = tf32_rna(c1[mt][ks][2]);
                a[mt][2] = tf32_rna(c1[mt][ks][1]); a[mt][3] = tf32_rna(c1[mt][ks][3]);
            }
#pragma unroll
            for (int nt = 0; nt < 5; ++nt) {
                const float2 b = *reinterpret_cast<const float2*>(&W2s[(g + 8 * nt) * SW2 + 8 * ks + 2 * t]);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) mma_tf32(o[mt][nt], a[mt], __float_as_uint(b.x), __float_as_uint(b.y));
            }
        }
        // ---- d_out in the C-fragment layout: column 0 = sigma (linear), columns 1..32 = rgb = sigmoid(o)*1.002 - 0.001
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 5; ++nt)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int row = g + 16 * mt + 8 * hh, col = 8 * nt + 2 * t;
                    const bool rv = row < cnt;
                    const long prow = row0 + base + row;
                    float d0 = 0.f, d1 = 0.f;
                    if (rv) {
                        if (col == 0) d0 = __ldg(p.d_sigma + prow);
                        else if (col < OUT) d0 = __ldg(p.d_rgb + prow * C + col - 1);
                        if (col + 1 < OUT) d1 = __ldg(p.d_rgb + prow * C + col);
                    }
                    const float s0 = sigmoid_fast(o[mt][nt][2 * hh]), s1 = sigmoid_fast(o[mt][nt][2 * hh + 1]);
                    if (col != 0) d0 *= 1.002f * s0 * (1.f - s0);
                    d1 *= 1.002f * s1 * (1.f - s1);
                    o[mt][nt][2 * hh] = d0; o[mt][nt][2 * hh + 1] = d1;
                    if (wgrad) *reinterpret_cast<uint32_t*>(opB + row * 128 + ((nt ^ g) << 4) + 4 * t) = pack_bf16(d0, d1);
                }
        if (wgrad) {       // dW2^T[j][k] += sum_p h[p][j] dO[p][k]: two k16 steps over the warp's 32 points
            tc::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                tc::tc_fence_after();
                const uint32_t idesc = tc::instr_desc_bf16(128, DW2_COLS, 1, 1);
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    tc::umma_bf16(tmem, tc::smem_desc(opA_u + k * 2048, 4096, 1024), tc::smem_desc(opB_u + k * 2048, 4096, 1024), idesc, 1);
                tc::umma_commit(mybar);
            }
            ++ncommit;
        }
        if (wgrad) {       // db2[k] += sum over the warp's rows
#pragma unroll
            for (int nt = 0; nt < 5; ++nt) {
                const float s0 = sum_over_g(o[0][nt][0] + o[0][nt][2] + o[1][nt][0] + o[1][nt][2]);
                const float s1 = sum_over_g(o[0][nt][1] + o[0][nt][3] + o[1][nt][1] + o[1][nt][3]);
                if (g == 0) { atomicAdd(&ab2[8 * nt + 2 * t], s0); atomicAdd(&ab2[8 * nt + 2 * t + 1], s1); }
            }
        }
        // ---- dh = d_out W2 (K = 40: the 5 n-tiles of O are the k-steps), two n-tiles at a time; d_a = dh * (1 - exp(-h)) replaces h
        uint32_t ao[2][5][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int ks = 0; ks < 5; ++ks) {
                ao[mt][ks][0] = tf32_rna(o[mt][ks][0]); ao[mt][ks][1] = tf32_rna(o[mt][ks][2]);
                ao[mt][ks][2] = tf32_rna(o[mt][ks][1]); ao[mt][ks][3] = tf32_rna(o[mt][ks][3]);
            }
#pragma unroll
        for (int nt0 = 0; nt0 < 8; nt0 += 2) {
            float dh[2][2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int i = 0; i < 4; ++i) dh[mt][u][i] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 5; ++ks)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const float* r0 = W2s + (8 * ks + 2 * t) * SW2 + g + 8 * (nt0 + u);
                    const uint32_t b0 = __float_as_uint(r0[0]), b1 = __float_as_uint(r0[SW2]);
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) mma_tf32(dh[mt][u], ao[mt][ks], b0, b1);
                }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int i = 0; i < 4; ++i) c1[mt][nt0 + u][i] = dh[mt][u][i] * (1.f - __expf(-c1[mt][nt0 + u][i]));
        }
        if (wgrad) {       // dW1[j][c] += sum_p d_a[p][j] F[p][c]: d_a replaces h in the A tile, F (bf16) replaces dO in the B tile
            tc::mbar_wait(mybar, (ncommit - 1) & 1);                           // the dW2 MMAs have read both tiles
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int row = g + 16 * mt + 8 * hh;
                        *reinterpret_cast<uint32_t*>(opA + row * 128 + ((nt ^ g) << 4) + 4 * t) = pack_bf16(c1[mt][nt][2 * hh], c1[mt][nt][2 * hh + 1]);
                    }
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {                                    // lane = point: 32 channels = four 16-byte chunks
                const float4 v0 = *reinterpret_cast<const float4*>(&sF[lane * SF + 8 * ch]);
                const float4 v1 = *reinterpret_cast<const float4*>(&sF[lane * SF + 8 * ch + 4]);
                *reinterpret_cast<uint4*>(opB + lane * 128 + ((ch ^ (lane & 7)) << 4)) =
                    make_uint4(pack_bf16(v0.x, v0.y), pack_bf16(v0.z, v0.w), pack_bf16(v1.x, v1.y), pack_bf16(v1.z, v1.w));
            }
            tc::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                tc::tc_fence_after();
                const uint32_t idesc = tc::instr_desc_bf16(128, C, 1, 1);
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    tc::umma_bf16(tmem + DW1_COL0, tc::smem_desc(opA_u + k * 2048, 4096, 1024), tc::smem_desc(opB_u + k * 2048, 4096, 1024), idesc, 1);
                tc::umma_commit(mybar);
            }
            ++ncommit;
        }
        if (wgrad) {       // db1[j] += sum over the warp's rows
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float s0 = sum_over_g(c1[0][nt][0] + c1[0][nt][2] + c1[1][nt][0] + c1[1][nt][2]);
                const float s1 = sum_over_g(c1[0][nt][1] + c1[0][nt][3] + c1[1][nt][1] + c1[1][nt][3]);
                if (g == 0) { atomicAdd(&ab1[8 * nt + 2 * t], s0); atomicAdd(&ab1[8 * nt + 2 * t + 1], s1); }
            }
        }
        // ---- d_f = d_a W1 (K = 64: the 8 n-tiles of layer 1 are the k-steps) -> staging tile sF[point][channel]
        {
            float df[2][4][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) df[mt][nt][i] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                uint32_t a[2][4];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    a[mt][0] = tf32_rna(c1[mt][ks][0]); a[mt][1] = tf32_rna(c1[mt][ks][2]);
                    a[mt][2] = tf32_rna(c1[mt][ks][1]); a[mt][3] = tf32_rna(c1[mt][ks][3]);
                }
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const float* r0 = W1s + (8 * ks + 2 * t) * SF + g + 8 * nt;
                    const uint32_t b0 = __float_as_uint(r0[0]), b1 = __float_as_uint(r0[SF]);
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) mma_tf32(df[mt][nt], a[mt], b0, b1);
                }
            }
            __syncwarp();                         // every lane is done with the feature tile (A fragments, bf16 export)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh)
                        *reinterpret_cast<float2*>(&sF[(g + 16 * mt + 8 * hh) * SF + 8 * nt + 2 * t]) =
                            make_float2(df[mt][nt][2 * hh], df[mt][nt][2 * hh + 1]);      // rows >= cnt carry exact zeros (d_out == 0)
        }
        __syncwarp();
        // ---- scatter d_f (the 1/3 of the plane mean is folded into the staged weights)
        if (dpl) scatter_features<SF>(dpl, ss, sF, lane, cnt);
        if (p.d_coords) {
            for (int q = 0; q < cnt; ++q) {
                const float gx = __shfl_sync(0xffffffffu, cx, q), gy = __shfl_sync(0xffffffffu, cy, q), gz = __shfl_sync(0xffffffffu, cz, q);
                const Bilin b0 = make_bilin(gx, gy, p.hp, p.wp), b1 = make_bilin(gx, gz, p.hp, p.wp), b2 = make_bilin(gz, gx, p.hp, p.wp);
                const float gq = sF[q * SF + lane] * (1.f / 3.f);
                float ax, ay, bx, by, ex, ey;
                bilin_dcoord(pl + lane, b0, p.wp, ax, ay);
                bilin_dcoord(pl + C + lane, b1, p.wp, bx, by);
                bilin_dcoord(pl + 2 * C + lane, b2, p.wp, ex, ey);
                const float hx = 0.5f * p.wp, hy = 0.5f * p.hp;
                float dx = gq * (ax * hx + bx * hx + ey * hy);
                float dy = gq * (ay * hy);
                float dz = gq * (by * hy + ex * hx);
                dx = warp_sum(dx); dy = warp_sum(dy); dz = warp_sum(dz);
                if (lane == 0) {
                    float* dcq = p.d_coords + ((long)n * p.P + base + q) * 3;
                    dcq[0] = dx * p.coord_scale; dcq[1] = dy * p.coord_scale; dcq[2] = dz * p.coord_scale;
                }
            }
        }
        __syncwarp();
    }
    if (wgrad) {
        if (ncommit) tc::mbar_wait(mybar, (ncommit - 1) & 1);                  // this warp's MMAs have retired
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        for (int i = threadIdx.x; i < HID; i += blockDim.x) atomicAdd(p.db1 + i, ab1[i] * p.b1g);
        for (int i = threadIdx.x; i < OUT; i += blockDim.x) atomicAdd(p.db2 + i, ab2[i] * p.b2g);
        if (wid < 2) {                                         // drain: TMEM lane j = hidden unit j
            const int j = wid * 32 + lane;
            float v[32];
            tc::tmem_ld32(tmem + ((uint32_t)(wid * 32) << 16), v);             // dW2^T[j][0..31]
#pragma unroll
            for (int k = 0; k < 32; ++k) atomicAdd(p.dW2 + k * HID + j, v[k] * p.w2g);
            tc::tmem_ld32(tmem + ((uint32_t)(wid * 32) << 16) + 32u, v);       // dW2^T[j][32..47] | dW1[j][0..15]
            atomicAdd(p.dW2 + 32 * HID + j, v[0] * p.w2g);
#pragma unroll
            for (int c = 0; c < 16; ++c) atomicAdd(p.dW1 + j * C + c, v[16 + c] * p.w1g);
            tc::tmem_ld32(tmem + ((uint32_t)(wid * 32) << 16) + 64u, v);       // dW1[j][16..31] | unused
#pragma unroll
            for (int c = 0; c < 16; ++c) atomicAdd(p.dW1 + j * C + 16 + c, v[c] * p.w1g);
        }
        tc::tc_fence_before();
        __syncthreads();
        if (wid == 0) tc::tmem_dealloc(tmem, DW_TCOLS);
    }
}

}  // namespace
int triplane_fwd_tc_launch(tri::TriplaneParams& p, bool sigma_only, cudaStream_t st);          // triplane_tc.cu
int triplane_bwd_tc_launch(tri::TriplaneParams& p, cudaStream_t st);
int g_b200_triplane_impl = -1;    // -1: take B200EG3D_TRIPLANE_IMPL from the environment on first use; 0: mma.sync kernels, 1: tcgen05 kernels
static inline int triplane_impl() {
    if (g_b200_triplane_impl < 0) { const char* e = getenv("B200EG3D_TRIPLANE_IMPL"); g_b200_triplane_impl = (e && e[0] == '0') ? 0 : 1; }
    return g_b200_triplane_impl;
}
int g_b200_mlp_passes = 0;        // 0: take B200EG3D_MLP_PASSES from the environment on first use; 1 / 3: set by b200_set_mlp_passes()
namespace {
int fill_common(TriplaneParams& p, const float* planes, int n, int hp, int wp, const float* coords, const float* ray_o,
                const float* ray_d, const float* depths, int S, int ray_w, long P, float box_warp, const float* W1, const float* b1,
                const float* W2, const float* b2, float lr_mul) {
    B200_REQUIRE(planes && W1 && b1 && W2 && b2, "triplane: null plane/weight pointer");
    B200_REQUIRE(coords || (ray_o && ray_d && depths && S > 0 && P % S == 0), "triplane: need coords or (ray_o, ray_d, depths, S)");
    B200_REQUIRE(n > 0 && hp > 0 && wp > 0 && P >= 0 && box_warp != 0.f, "triplane: bad shape");
    p.planes = planes; p.n = n; p.hp = hp; p.wp = wp; p.coords = coords; p.ray_o = ray_o; p.ray_d = ray_d; p.depths = depths;
    B200_REQUIRE(ray_w >= 0 && (ray_w == 0 || coords || (P / S) % ray_w == 0), "triplane: ray_w must divide the ray count");
    p.S = S; p.ray_w = coords ? 0 : ray_w; p.P = P; p.coord_scale = 2.f / box_warp;
    p.M = coords ? 0 : P / S; p.ray_h = p.ray_w > 0 ? (int)(p.M / p.ray_w) : 0;
    if (p.ray_w > 0 && (p.ray_w % 8 != 0 || p.ray_h % 16 != 0)) p.ray_w = p.ray_h = 0;      // patch order needs whole 8 x 16 patches
    p.W1 = W1; p.b1 = b1; p.W2 = W2; p.b2 = b2;
    p.w1g = lr_mul / sqrtf((float)C); p.b1g = lr_mul; p.w2g = lr_mul / sqrtf((float)HID); p.b2g = lr_mul;
    return 0;
}
}  // namespace

// planes: [n][hp][wp][96] fp32.  Either coords [n][P][3], or rays (ray_o/ray_d [n][P/S][3], depths [n][P]).
// W1 [64][32], b1 [64], W2 [33][64], b2 [33] are the raw decoder parameters (gains lr_mul/sqrt(fan_in) applied here,
// training/networks_stylegan2.py:111-112).  Outputs rgb [n][P][32], sigma [n][P].
B200_API int b200_triplane_mlp_fwd(const float* planes, int n, int hp, int wp, const float* coords, const float* ray_o,
                                   const float* ray_d, const float* depths, int S, int ray_w, long P, float box_warp,
                                   const float* W1, const float* b1, const float* W2, const float* b2, float lr_mul,
                                   float* rgb, float* sigma, void* f_save, void* stream) {
    TriplaneParams p{};
    if (int e = fill_common(p, planes, n, hp, wp, coords, ray_o, ray_d, depths, S, ray_w, P, box_warp, W1, b1, W2, b2, lr_mul)) return e;
    B200_REQUIRE(sigma, "triplane_fwd: null output");        // rgb == NULL: density-only query
    if (P == 0) return 0;
    p.rgb = rgb; p.sigma = sigma;
    const long groups = (P + 127) / 128;
    dim3 grid((unsigned)(groups < 148 * 8 ? groups : 148 * 8), n);
    if (g_b200_mlp_passes == 0) {
        const char* e2 = getenv("B200EG3D_MLP_PASSES");
        g_b200_mlp_passes = (e2 && strcmp(e2, "1") == 0) ? 1 : 3;
    }
    B200_FUNC_ATTR_ONCE(triplane_mlp_fwd_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FW_SMEM);
    B200_FUNC_ATTR_ONCE(triplane_mlp_fwd_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FW_SMEM);
    p.fwd_passes = g_b200_mlp_passes;
    p.f_save = triplane_impl() == 1 ? static_cast<uint8_t*>(f_save) : nullptr;      // only the tcgen05 kernels produce / consume it
    if (triplane_impl() == 1 && P < (1L << 31)) return triplane_fwd_tc_launch(p, rgb == nullptr, (cudaStream_t)stream);
    const long g512 = (P + FW_WARPS * 32 - 1) / (FW_WARPS * 32);
    grid.x = (unsigned)(g512 < 148 ? g512 : 148);                      // persistent: one 16-warp CTA per SM
    if (rgb) triplane_mlp_fwd_mma_kernel<false><<<grid, FW_WARPS * 32, FW_SMEM, (cudaStream_t)stream>>>(p);
    else triplane_mlp_fwd_mma_kernel<true><<<grid, FW_WARPS * 32, FW_SMEM, (cudaStream_t)stream>>>(p);
    B200_CHECK_LAUNCH();
    return 0;
}

// Operand passes of the decoder forward: 3 = split operands (fp32-equivalent, the parity default), 1 = single pass (the
// non-parity "fast mode" reported separately by bench.py).  Returns the previous value.
// Implementation of the fused sampler + decoder: 1 = tcgen05 pipeline (default), 0 = the mma.sync kernels of round 1 (kept as
// the on-GPU cross-check; env B200EG3D_TRIPLANE_IMPL=0).  Returns the previous setting.
B200_API int b200_set_triplane_impl(int impl) {
    const int prev = triplane_impl();
    g_b200_triplane_impl = impl ? 1 : 0;
    return prev;
}

B200_API int b200_set_mlp_passes(int passes) {
    const int prev = g_b200_mlp_passes == 0 ? 3 : g_b200_mlp_passes;
    g_b200_mlp_passes = passes == 1 ? 1 : 3;
    return prev;
}

// Size of the optional forward -> backward hand-off buffer: every 128-point tile's mean features as the finished layer-1
// tensor-core operand (split bf16, 128 bytes per point).  With it the backward reloads the features with one bulk copy per tile
// instead of gathering the 12 texel lines per point a second time (the kernels are L2-bandwidth bound on exactly that gather).
B200_API long b200_triplane_fsave_bytes(int n, long P) {
    return (long)n * ((P + 127) / 128) * 16384;
}

// Scratch the backward may need: the per-point coordinate gradients (12 B / point, when per-ray sums are requested without
// d_coords) and the per-point feature gradients handed from the decoder pass to the plane-scatter pass (64 B / point, tcgen05
// implementation).  The caller allocates it once and passes it to b200_triplane_mlp_bwd.
B200_API long b200_triplane_bwd_workspace_bytes(int n, long P) {
    return (long)n * P * (12 + 64) + 1024;
}

namespace {
// d ray_o[m] += sum_k d point[m][k],  d ray_d[m] += sum_k t[m][k] * d point[m][k]      (point = o + t * d, renderer.py:161,178)
__global__ void ray_reduce_dpoints_kernel(const float* __restrict__ d_pts, const float* __restrict__ depths, long rays, int S,
                                          float* __restrict__ d_ro, float* __restrict__ d_rd) {
    const int lane = threadIdx.x & 31;
    const long ray = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ray >= rays) return;
    float so[3] = {0.f, 0.f, 0.f}, sd[3] = {0.f, 0.f, 0.f};
    for (int k = lane; k < S; k += 32) {
        const float t = depths[ray * S + k];
        const float* g = d_pts + (ray * S + k) * 3;
#pragma unroll
        for (int a = 0; a < 3; ++a) { so[a] += g[a]; sd[a] += t * g[a]; }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) { so[a] = warp_sum(so[a]); sd[a] = warp_sum(sd[a]); }
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { d_ro[ray * 3 + a] += so[a]; d_rd[ray * 3 + a] += sd[a]; }
    }
}
}  // namespace

// d_planes [n][hp][wp][96] is ACCUMULATED into (zero it first); d_coords [n][P][3] is written (may be null);
// d_ray_o / d_ray_d [n][P/S][3] (ray mode; both or neither; may be null) are ACCUMULATED into: the per-ray sums of d point and
// t * d point, i.e. the gradients of the ray origins / directions;  dW1/db1/dW2/db2 are ACCUMULATED into (all four null =>
// parameter gradients skipped).  f_saved: the buffer the forward filled through its f_save argument for the SAME points, ray_w and
// implementation setting (may be null: the features are then gathered again).  `workspace`: b200_triplane_bwd_workspace_bytes(n, P) bytes of scratch (may be null when neither
// d_ray_o nor the tcgen05 implementation needs it).
B200_API int b200_triplane_mlp_bwd(const float* planes, int n, int hp, int wp, const float* coords, const float* ray_o,
                                   const float* ray_d, const float* depths, int S, int ray_w, long P, float box_warp,
                                   const float* W1, const float* b1, const float* W2, const float* b2, float lr_mul,
                                   const float* d_rgb, const float* d_sigma, const void* f_saved, float* d_planes, float* d_coords,
                                   float* d_ray_o, float* d_ray_d,
                                   float* dW1, float* db1, float* dW2, float* db2, void* workspace, long workspace_bytes,
                                   void* stream) {
    TriplaneParams p{};
    if (int e = fill_common(p, planes, n, hp, wp, coords, ray_o, ray_d, depths, S, ray_w, P, box_warp, W1, b1, W2, b2, lr_mul)) return e;
    B200_REQUIRE(d_rgb && d_sigma, "triplane_bwd: null incoming gradient");
    B200_REQUIRE((dW1 && db1 && dW2 && db2) || (!dW1 && !db1 && !dW2 && !db2), "triplane_bwd: pass all four parameter gradients or none");
    B200_REQUIRE((d_ray_o != nullptr) == (d_ray_d != nullptr), "triplane_bwd: pass both ray gradients or neither");
    B200_REQUIRE(!d_ray_o || !coords, "triplane_bwd: ray gradients need ray mode");
    if (P == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    p.d_rgb = d_rgb; p.d_sigma = d_sigma; p.d_planes = d_planes; p.d_coords = d_coords;
    p.dW1 = dW1; p.db1 = db1; p.dW2 = dW2; p.db2 = db2;
    p.f_saved = static_cast<const uint8_t*>(f_saved);
    p.d_ray_o = d_ray_o; p.d_ray_d = d_ray_d;
    // tcgen05 pipeline: needs the forward's feature tiles (otherwise the mma.sync kernel gathers the features again)
    if (triplane_impl() == 1 && f_saved && P < (1L << 31)) return triplane_bwd_tc_launch(p, st);
    p.d_ray_o = p.d_ray_d = nullptr;
    if (d_ray_o && !d_coords) {          // per-point coordinate gradients staged in the workspace, reduced per ray below
        B200_REQUIRE(workspace && workspace_bytes >= (long)n * P * 12, "triplane_bwd: workspace too small for the ray gradients");
        p.d_coords = static_cast<float*>(workspace);
    }
    B200_FUNC_ATTR_ONCE((triplane_mlp_bwd_mma_kernel<BM_WARPS_WG, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, bm_smem(BM_WARPS_WG, true));
    B200_FUNC_ATTR_ONCE((triplane_mlp_bwd_mma_kernel<BM_WARPS_NOWG, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, bm_smem(BM_WARPS_NOWG, false));
    const int warps = dW1 ? BM_WARPS_WG : BM_WARPS_NOWG;
    const long gb = (P + warps * 32 - 1) / (warps * 32);
    const int sms = b200_sm_count();
    dim3 grid((unsigned)(gb < sms ? gb : sms), n);                     // persistent: one CTA per SM
    if (dW1) triplane_mlp_bwd_mma_kernel<BM_WARPS_WG, true><<<grid, warps * 32, bm_smem(BM_WARPS_WG, true), st>>>(p);
    else triplane_mlp_bwd_mma_kernel<BM_WARPS_NOWG, false><<<grid, warps * 32, bm_smem(BM_WARPS_NOWG, false), st>>>(p);
    B200_CHECK_LAUNCH();
    if (d_ray_o) {
        const long rays = (long)n * (P / S);
        ray_reduce_dpoints_kernel<<<(unsigned)((rays + 7) / 8), 256, 0, st>>>(p.d_coords, depths, rays, S, d_ray_o, d_ray_d);
        B200_CHECK_LAUNCH();
    }
    return 0;
}
