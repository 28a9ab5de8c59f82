// tcgen05 / TMA implicit-GEMM convolution for sm_100a ("tensor-core path" of the modulated-conv stack).
//
// Operands are bf16 "split-float" pairs: x = hi + lo with hi = bf16(x), lo = bf16(x - hi).  A forward convolution
// issues three MMAs per k-step (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM), which restores ~16 mantissa bits
// -- enough for the 1e-3 max-abs parity bar against the fp32 reference at half the smem bytes of a
// 3xTF32 scheme.  Backward GEMMs may run single-pass (hi*hi).
//
// Pixel-GEMM kernel (forward, dgrad, 1x1, transposed-conv classes): C[pixel, n] = sum_{tap, c} A[src(pixel,tap), c] * B[tap, c, n]
//   A tile  : 128 pixels (th x tw patch) x 64 channels, one TMA box {64, tw, th, 1} of the NHWC tensor at the tap's
//             shifted coordinate (TMA zero-fill == conv zero padding; elementStrides give the stride-2 gather)
//   B tile  : K-major  [BN rows][64 k]  (forward: wmod[tap][cout][cin])        one box {64, BN}
//             MN-major [64 k rows][64 n] x BN/64 (dgrad: wmod[tap][cout][cin] read as [k = cout][n = cin])
//   D       : 128 lanes x BN fp32 columns in TMEM, drained by 4 epilogue warps with tcgen05.ld
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one thread), warps 2-5 = epilogue.
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda_bf16.h>

using namespace tc;

namespace {

constexpr int TILE_M = 128, TILE_K = 64, MAX_BN = 128, STAGES = 3, MAX_ST = 8;
constexpr int A_BYTES = TILE_M * TILE_K * 2;              // 16 KB
constexpr int B_BYTES = MAX_BN * TILE_K * 2;              // 16 KB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;    // hi + lo of both operands
constexpr int RING_BYTES = STAGES * STAGE_BYTES;         // 192 KB operand ring (the pixel-GEMM kernel cuts it into 3..8 stages)
constexpr int SMEM_BYTES = RING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

// One "class" = one iteration grid with its tap list.  Plain convolutions have a single class; the stride-2 transposed
// convolution has four (output parities), merged into ONE launch so that their CTAs fill the GPU together.
struct TcClass {
    int th, tw, tiles_x, ntiles, Hi, Wi, ntaps, ooy, oox;
    int dy[9], dx[9], wt[9];
};

struct TcPixParams {
    CUtensorMap tmA[4][2];       // per class: hi, lo : 4-D {C, W, H, N}  (the box depends on the class's tile shape)
    CUtensorMap tmB[2];          // hi, lo : 2-D {inner, rows}
    TcClass c[4];
    int cta_start[5];            // prefix sums of (tiles * ksplit) per class; CTA pairs (cg == 2): tiles counted in pairs
    int cg;                      // 1: one CTA per 128-pixel tile; 2: a CTA pair issues M = 256 MMAs over two tiles and shares B
    int ncls;
    int kchunks, npass;
    int s;                       // source pixel = iteration pixel * s + (dy, dx)
    int b_rows_per_tap;          // K-major B: total N; MN-major B: total K per tap
    int b_taps;                  // tap slices per sample in B
    int BN, N;
    float* C; long ldc, c_bs; int Wo, osy, osx;
    int ksplit;                  // > 1: the k-blocks of a tile are spread over ksplit CTAs, partial sums reduced with red.global.add
    int ntn, nwork;              // N tiles; work items = cta_start[ncls] * ntn * batch
    int stage_bytes, nstages;    // ring geometry: bytes one k-block really needs, and how many fit in RING_BYTES
    // Optional fused SynthesisLayer epilogue (up == 1 forward, no split-K, N a multiple of 32; networks_stylegan2.py:318-329):
    // z = clamp(lrelu(acc + noise * strength + bias) * act_gain), written as fp32 (C) and as the split-bf16 pair of the next conv.
    int act;
    const float* bias; const float* noise; const float* strength; long noise_bs;
    float alpha, act_gain, clamp;
    __nv_bfloat16* z_hi; __nv_bfloat16* z_lo;
};

__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    return make_uint4(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b), *reinterpret_cast<uint32_t*>(&c),
                      *reinterpret_cast<uint32_t*>(&d));
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Persistent: one CTA per SM walks the work list (pixel tile x k-slice x N tile x sample) with a stride of gridDim.x.  The
// shared-memory ring keeps running across work items (the producer prefetches the next tile while the MMAs of the current
// one drain) and the accumulator is double-buffered in TMEM, so the epilogue of item i overlaps the main loop of item i+1.
//
// CG == 2 (launched as clusters of two CTAs): the pair walks the list together, CTA rank r takes pixel tile 2i + r of pair-tile i
// and loads rows [r * BN/2, (r+1) * BN/2) of the B tile; the leader's single thread issues cta_group::2 MMAs (M = 256) that read
// both CTAs' shared memory and write both CTAs' tensor memory.  Per CTA that is A + B/2 per k-block instead of A + B from L2
// (the L2 -> SM fabric, ~43 B/clk/SM, is what bounds the single-CTA kernel: 64 KB per 768 tensor-clocks), and B is read from
// shared memory once per pair.
template <bool B_MN, int CG>
__global__ void __launch_bounds__(192, 1) conv_tc_pix_kernel(const __grid_constant__ TcPixParams p) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    const uint32_t bars = base + STAGES * STAGE_BYTES;                 // full[MAX_ST], empty[MAX_ST], tfull[2], tempty[2], tmem slot
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (MAX_ST + s); };
    auto tfull = [&](int a) { return bars + 8u * (2 * MAX_ST + a); };
    auto tempty = [&](int a) { return bars + 8u * (2 * MAX_ST + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (2 * MAX_ST + 4);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + STAGES * STAGE_BYTES + 8 * (2 * MAX_ST + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nst = p.nstages;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    const int walker = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, nwalkers = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int bn_cta = p.BN / CG;                                      // B rows (N) this CTA holds
    const uint32_t b_tile = (uint32_t)bn_cta * TILE_K * 2;
    const uint32_t a_lo_off = A_BYTES, b_hi_off = p.npass == 3 ? 2 * A_BYTES : A_BYTES, b_lo_off = b_hi_off + b_tile;
    const uint32_t tcols = 2 * p.BN <= 128 ? 128u : (2 * p.BN <= 256 ? 256u : 512u);

    if (threadIdx.x == 0) {
        for (int s = 0; s < nst; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 4 * CG); }
        fence_barrier_init();
        for (int k = 0; k < p.ncls; ++k) { tma_prefetch_desc(&p.tmA[k][0]); if (p.npass == 3) tma_prefetch_desc(&p.tmA[k][1]); }
        tma_prefetch_desc(&p.tmB[0]);
        if (p.npass == 3) tma_prefetch_desc(&p.tmB[1]);
    }
    if (warp == 1) { if (CG == 2) tmem_alloc_cg2(tmem_slot, tcols); else tmem_alloc(tmem_slot, tcols); }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();       // pair: the peer's barriers must exist before anything signals them
    tc_fence_after();
    const uint32_t tmem = *tmem_slot_ptr;
    pdl_wait();                       // barriers, descriptors and TMEM are set up while the previous kernel drains

    // work item -> (class, tile, k-slice, N tile, sample); every role decodes the same list
    struct Item { int cls, x0, y0, n0, b, it0, nk; };
    const int per_x = p.cta_start[p.ncls];
    auto decode = [&](int w, Item& o) {
        const int lx = w % per_x; int r = w / per_x;
        o.n0 = (r % p.ntn) * p.BN; o.b = r / p.ntn;
        int cls = 0;
        while (cls + 1 < p.ncls && lx >= p.cta_start[cls + 1]) ++cls;
        o.cls = cls;
        const TcClass& kc = p.c[cls];
        const int local = lx - p.cta_start[cls];
        const int tile = (local / p.ksplit) * CG + (int)rank, ksi = local % p.ksplit;
        o.x0 = (tile % kc.tiles_x) * kc.tw; o.y0 = (tile / kc.tiles_x) * kc.th;
        if (CG == 2 && tile >= kc.ntiles) o.y0 = 1 << 20;         // odd tile count: the pair's second CTA multiplies zero-filled rows, stores nothing
        const int nk_all = kc.ntaps * p.kchunks;
        const int per = (nk_all + p.ksplit - 1) / p.ksplit;
        o.it0 = ksi * per;
        o.nk = max(min(nk_all, o.it0 + per) - o.it0, 0);      // 0: nothing to do for this k-slice (ksplit does not divide evenly)
    };

    if (warp == 0) {
        if (lane == 0) {
            const int nh = p.npass == 3 ? 2 : 1;
            const uint32_t bytes = (uint32_t)(A_BYTES + b_tile) * nh;
            int g = 0, s = 0, ph = 0;                              // k-blocks issued; ring position and phase of the next one
            for (int w = walker; w < p.nwork; w += nwalkers) {
                Item o; decode(w, o);
                const TcClass& kc = p.c[o.cls];
                int t = o.it0 / p.kchunks, kc0 = o.it0 % p.kchunks;   // tap and channel chunk of the next k-block
                for (int it = 0; it < o.nk; ++it, ++g) {
                    mbar_wait(empty(s), ph ^ 1);
                    // pair: both CTAs' loads count on the LEADER's full barrier, which the leader arms with the bytes of both
#ifdef B200_DBG_NO_TMA
                    if (rank == 0) mbar_arrive(full(s));
                    if (++s == nst) { s = 0; ph ^= 1; }
                    if (++kc0 == p.kchunks) { kc0 = 0; ++t; }
                    continue;
#endif
                    if (rank == 0) mbar_arrive_expect_tx(full(s), bytes * CG);
                    const uint32_t fbar = CG == 2 ? mapa_u32(full(s), 0) : full(s);
                    const int c0 = kc0 * TILE_K;
                    const uint32_t st = base + s * p.stage_bytes;
                    const int ax = o.x0 * p.s + kc.dx[t], ay = o.y0 * p.s + kc.dy[t];
                    const int nb = o.n0 + (int)rank * bn_cta;          // first B row (N index) of this CTA
                    for (int h = 0; h < nh; ++h) {
                        const uint32_t bdst = st + (h ? b_lo_off : b_hi_off);
                        if (CG == 2) {
                            tma_load_4d_cg2(st + h * a_lo_off, &p.tmA[o.cls][h], fbar, c0, ax, ay, o.b);
                            if (!B_MN) {
                                tma_load_2d_cg2(bdst, &p.tmB[h], fbar, c0, (o.b * p.b_taps + kc.wt[t]) * p.b_rows_per_tap + nb);
                            } else {
                                const int krow = (o.b * p.b_taps + kc.wt[t]) * p.b_rows_per_tap + c0;
                                for (int j = 0; j < bn_cta / 64; ++j) tma_load_2d_cg2(bdst + j * (TILE_K * 128), &p.tmB[h], fbar, nb + j * 64, krow);
                            }
                        } else {
                            tma_load_4d(st + h * a_lo_off, &p.tmA[o.cls][h], fbar, c0, ax, ay, o.b);
                            if (!B_MN) {
                                tma_load_2d(bdst, &p.tmB[h], fbar, c0, (o.b * p.b_taps + kc.wt[t]) * p.b_rows_per_tap + nb);
                            } else {
                                const int krow = (o.b * p.b_taps + kc.wt[t]) * p.b_rows_per_tap + c0;
                                for (int j = 0; j < bn_cta / 64; ++j) tma_load_2d(bdst + j * (TILE_K * 128), &p.tmB[h], fbar, nb + j * 64, krow);
                            }
                        }
                    }
                    if (++s == nst) { s = 0; ph ^= 1; }
                    if (++kc0 == p.kchunks) { kc0 = 0; ++t; }
                }
            }
            if (CG == 2) {
                // pair: the leader's multicast commits still arrive on this CTA's empty barriers after the last load was issued;
                // wait for the release of every stage still in use so that no arrival targets a CTA that has already retired
                for (int j = 0; j < nst && j < g; ++j) {
                    const int gg = g - 1 - j;
                    mbar_wait(empty(gg % nst), (gg / nst) & 1);
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // The whole warp runs this loop converged (descriptor arithmetic stays warp-uniform); lane 0 issues.  Descriptors are
            // built once: per k-block only the 16-byte-unit start addresses move.
            const uint32_t idesc = instr_desc_bf16(TILE_M * CG, p.BN, 0, B_MN ? 1 : 0);
            const uint32_t dhi = smem_desc_hi(1024);
            const uint32_t a_hi0 = smem_desc_lo(base, 0), a_lo0 = smem_desc_lo(base + a_lo_off, 0);
            const uint32_t b_hi0 = smem_desc_lo(base + b_hi_off, B_MN ? TILE_K * 128 : 0), b_lo0 = smem_desc_lo(base + b_lo_off, B_MN ? TILE_K * 128 : 0);
            constexpr uint32_t KA = 32 >> 4, KB = (B_MN ? 2048 : 32) >> 4;          // k16 step of the two operands, in 16-byte units
            const uint32_t stage16 = (uint32_t)p.stage_bytes >> 4;
            const bool three = p.npass == 3;
            int s = 0, ph = 0, li = 0;
            for (int w = walker; w < p.nwork; w += nwalkers) {
                Item o; decode(w, o);
                if (o.nk == 0) continue;
                const int acc = li & 1;
                mbar_wait(tempty(acc), ((li >> 1) & 1) ^ 1);       // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d = tmem + (uint32_t)(acc * p.BN);
                for (int it = 0; it < o.nk; ++it) {
                    mbar_wait(full(s), ph);
                    tc_fence_after();
                    const uint32_t so = (uint32_t)s * stage16;
                    if (elect_one()) {
#ifndef B200_DBG_NO_MMA
#pragma unroll
                        for (int k = 0; k < TILE_K / 16; ++k) {
                            const uint64_t ah = desc_pack(a_hi0 + so + k * KA, dhi), bh = desc_pack(b_hi0 + so + k * KB, dhi);
                            if (CG == 2) umma_bf16_cg2(d, ah, bh, idesc, (it | k) != 0); else umma_bf16(d, ah, bh, idesc, (it | k) != 0);
                            if (three) {
                                const uint64_t al = desc_pack(a_lo0 + so + k * KA, dhi), bl = desc_pack(b_lo0 + so + k * KB, dhi);
                                if (CG == 2) { umma_bf16_cg2(d, ah, bl, idesc, 1); umma_bf16_cg2(d, al, bh, idesc, 1); }
                                else { umma_bf16(d, ah, bl, idesc, 1); umma_bf16(d, al, bh, idesc, 1); }
                            }
                        }
#endif
                        if (CG == 2) umma_commit_cg2(empty(s)); else umma_commit(empty(s));       // pair: frees the stage in both CTAs
                        if (it == o.nk - 1) { if (CG == 2) umma_commit_cg2(tfull(acc)); else umma_commit(tfull(acc)); }
                    }
                    __syncwarp();
                    if (++s == nst) { s = 0; ph ^= 1; }
                }
                ++li;
            }
        }
    } else {
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const bool vec = (p.N & 3) == 0;
        int li = 0;
        for (int w = walker; w < p.nwork; w += nwalkers) {
            Item o; decode(w, o);
            if (o.nk == 0) continue;
            const TcClass& kc = p.c[o.cls];
            const int acc = li & 1;
            mbar_wait(tfull(acc), (li >> 1) & 1);
            tc_fence_after();
            const int m = q * 32 + lane;
            const int iy = o.y0 + m / kc.tw, ix = o.x0 + m % kc.tw;
            const bool valid = iy < kc.Hi && ix < kc.Wi;
            const long opix = (long)(iy * p.osy + kc.ooy) * p.Wo + (ix * p.osx + kc.oox);
            float* crow = p.C + (long)o.b * p.c_bs + opix * p.ldc + o.n0;
            for (int c0 = 0; c0 < p.BN; c0 += 32) {
                float v[32];
#ifdef B200_DBG_NO_EPI
                if (c0 + 32 < p.BN) continue;
#endif
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN + c0), v);
                if (c0 + 32 >= p.BN) {                 // last read of this accumulator: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { if (CG == 2) mbar_arrive_cluster(mapa_u32(tempty(acc), 0)); else mbar_arrive(tempty(acc)); }
                }
                if (valid && p.act) {                  // fused layer epilogue (launch code guarantees ksplit == 1, N % 32 == 0, plain output grid)
                    const float nz = p.noise ? __ldg(p.noise + (long)o.b * p.noise_bs + (long)iy * kc.Wi + ix) * __ldg(p.strength) : 0.f;
                    const float4* bsrc = reinterpret_cast<const float4*>(p.bias + o.n0 + c0);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 bv = __ldg(bsrc + (j >> 2));
                        const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float t = v[j + q] + nz + bb[q];
                            t = t > 0.f ? t : t * p.alpha;
                            t *= p.act_gain;
                            if (p.clamp >= 0.f) t = (t > -p.clamp && t < p.clamp) ? t : (t >= 0.f ? p.clamp : -p.clamp);
                            v[j + q] = t;
                        }
                        if (p.C) *reinterpret_cast<float4*>(crow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                    const long eo = (long)o.b * p.c_bs + opix * p.ldc + o.n0 + c0;
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        const uint4 h = pack8_bf16(v + j);
                        *reinterpret_cast<uint4*>(p.z_hi + eo + j) = h;
                        if (p.z_lo) {
                            const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(&h);
                            float lo[8];
#pragma unroll
                            for (int q = 0; q < 8; ++q) lo[q] = v[j + q] - __bfloat162float(hb[q]);
                            *reinterpret_cast<uint4*>(p.z_lo + eo + j) = pack8_bf16(lo);
                        }
                    }
                } else if (valid) {
#ifdef B200_DBG_NO_STORE
                    if (v[0] != 12345.678f) continue;
#endif
                    if (vec && o.n0 + c0 + 32 <= p.N) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            if (p.ksplit > 1) red_add_v4(crow + c0 + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
                            else *reinterpret_cast<float4*>(crow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (o.n0 + c0 + j < p.N) { if (p.ksplit > 1) atomicAdd(crow + c0 + j, v[j]); else crow[c0 + j] = v[j]; }
                    }
                }
            }
            ++li;
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();       // pair: nobody leaves while the peer may still touch its memory
    if (warp == 1) { if (CG == 2) tmem_dealloc_cg2(tmem, tcols); else tmem_dealloc(tmem, tcols); }
}

// ---------------------------------------------------------------------------------------------------------------------
// Weight-gradient kernel: D[tap][m = cout][n = cin] += sum_{pixels} A[pixA(pixel, tap)][m] * B[pixB(pixel, tap)][n]
// Both operands are MN-major (the reduction index is the pixel, channels are contiguous): a k-block is a th x tw patch
// of 64 pixels, loaded as {64 ch, tw, th} boxes = [64 k rows][64 ch] swizzle-128B atoms.  One CTA owns one tap, one
// 128 x BN output tile and one slice of the pixel range (split-K); partial tiles are reduced with red.global.add.v4.f32.
struct TcWgradParams {
    CUtensorMap tmA[2];          // hi, lo of the M-side tensor (dy / dz): 4-D {C, W, H, N}
    CUtensorMap tmB[2];          // hi, lo of the N-side tensor (x)
    int th, tw, tiles_x, ntiles; // pixel patches (64 pixels each) over the iteration grid
    int sA, sB;                  // pixel strides of the two gathers
    int dAy[9], dAx[9], dBy[9], dBx[9];
    int ntaps, ksplit, npass;
    int BN, Cm, Cn;
    float* C; long c_bs;         // C[b][tap][Cm][Cn]
    int tg;                      // taps per CTA (single-pass only): the operand that does not move with the tap is loaded once
    int shared_a;                // 1: A (dy) is common to the taps of a group, B (x) shifts (up == 1); 0: the reverse (up == 2)
};

__global__ void __launch_bounds__(192, 1) conv_tc_wgrad_kernel(const __grid_constant__ TcWgradParams p) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    const uint32_t bars = base + STAGES * STAGE_BYTES;
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (STAGES + s); };
    const uint32_t accum_bar = bars + 8u * (2 * STAGES);
    const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 1);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 1));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TILE_M, n0 = blockIdx.y * p.BN;
    const int ngroups = p.ntaps / p.tg;
    int z = blockIdx.z;
    const int ks = z % p.ksplit; z /= p.ksplit;
    const int t0 = (z % ngroups) * p.tg;          // first tap of this CTA's group
    const int b = z / ngroups;
    const int per = (p.ntiles + p.ksplit - 1) / p.ksplit;
    const int kb0 = ks * per, kb1 = min(p.ntiles, kb0 + per);
    const int nk = max(kb1 - kb0, 0);
    const uint32_t tcols = p.tg * p.BN <= 128 ? 128u : (p.tg * p.BN <= 256 ? 256u : 512u);
    const uint32_t b_tile = (uint32_t)p.BN * TILE_K * 2;      // bytes of one B tile (N-side), A tile is A_BYTES

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
        tma_prefetch_desc(&p.tmA[0]); tma_prefetch_desc(&p.tmB[0]);
    }
    if (warp == 1) tmem_alloc(tmem_slot, tcols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot_ptr;
    pdl_wait();

    // Stage layout.  tg == 1 (any npass): A_hi | A_lo | B_hi | B_lo as in the pixel-GEMM kernel.
    // tg  > 1 (single pass, hi only): shared tile first, then the tg shifted tiles.
    auto a_addr = [&](uint32_t st, int j, int h) { return p.tg == 1 ? st + h * A_BYTES : (p.shared_a ? st : st + b_tile + j * A_BYTES); };
    auto b_addr = [&](uint32_t st, int j, int h) { return p.tg == 1 ? st + 2 * A_BYTES + h * B_BYTES : (p.shared_a ? st + A_BYTES + j * b_tile : st); };

    if (nk > 0) {
        if (warp == 0) {
            if (lane == 0) {
                const int nh = p.npass == 3 ? 2 : 1;
                const uint32_t bytes = p.tg == 1 ? (A_BYTES + b_tile) * nh
                                                 : (p.shared_a ? A_BYTES + p.tg * b_tile : p.tg * A_BYTES + b_tile);
                for (int it = 0; it < nk; ++it) {
                    const int s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(empty(s), ph ^ 1);
#ifdef B200_DBG_NO_TMA
                    mbar_arrive(full(s));
                    continue;
#endif
                    mbar_arrive_expect_tx(full(s), bytes);
                    const int tile = kb0 + it;
                    const int x0 = (tile % p.tiles_x) * p.tw, y0 = (tile / p.tiles_x) * p.th;
                    const uint32_t st = base + s * STAGE_BYTES;
                    for (int h = 0; h < nh; ++h)
                        for (int j = 0; j < p.tg; ++j) {
                            const int t = t0 + j;
                            if (j == 0 || !p.shared_a)
                                for (int q = 0; q < 2; ++q)
                                    tma_load_4d(a_addr(st, j, h) + q * (TILE_K * 128), &p.tmA[h], full(s), m0 + q * 64,
                                                x0 * p.sA + p.dAx[t], y0 * p.sA + p.dAy[t], b);
                            if (j == 0 || p.shared_a)
                                for (int q = 0; q < p.BN / 64; ++q)
                                    tma_load_4d(b_addr(st, j, h) + q * (TILE_K * 128), &p.tmB[h], full(s), n0 + q * 64,
                                                x0 * p.sB + p.dBx[t], y0 * p.sB + p.dBy[t], b);
                        }
                }
            }
        } else if (warp == 1) {
            // converged warp, one elected lane issues (see the pixel-GEMM kernel): descriptor halves built once per tile
            const uint32_t idesc = instr_desc_bf16(TILE_M, p.BN, 1, 1);
            const uint32_t dhi = smem_desc_hi(1024);
            constexpr uint32_t LBO = TILE_K * 128, KS = 2048 >> 4;             // 64-element MN blocks 8 KB apart; k16 step = 16 rows of 128 B
            const bool three = p.npass == 3;
            int s = 0, ph = 0;
            for (int it = 0; it < nk; ++it) {
                mbar_wait(full(s), ph);
                tc_fence_after();
                const uint32_t st = base + s * STAGE_BYTES;
                if (elect_one()) {
                    for (int j = 0; j < p.tg; ++j) {
                        const uint32_t ah0 = smem_desc_lo(a_addr(st, j, 0), LBO), bh0 = smem_desc_lo(b_addr(st, j, 0), LBO);
                        const uint32_t d = tmem + (uint32_t)(j * p.BN);
#pragma unroll
                        for (int k = 0; k < TILE_K / 16; ++k) {
                            const uint64_t dah = desc_pack(ah0 + k * KS, dhi), dbh = desc_pack(bh0 + k * KS, dhi);
#ifdef B200_DBG_NO_MMA
                            if (it >= 0) continue;
#endif
                            umma_bf16(d, dah, dbh, idesc, (it | k) != 0);
                            if (three) {
                                const uint64_t dal = desc_pack(smem_desc_lo(a_addr(st, j, 1), LBO) + k * KS, dhi);
                                const uint64_t dbl = desc_pack(smem_desc_lo(b_addr(st, j, 1), LBO) + k * KS, dhi);
                                umma_bf16(d, dah, dbl, idesc, 1);
                                umma_bf16(d, dal, dbh, idesc, 1);
                            }
                        }
                    }
                    umma_commit(empty(s));
                    if (it == nk - 1) umma_commit(accum_bar);
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        } else {
            const int q = warp & 3;
            mbar_wait(accum_bar, 0);
            tc_fence_after();
            const int m = m0 + q * 32 + lane;
            const bool vec = (p.Cn & 3) == 0;
            for (int j = 0; j < p.tg; ++j) {
                float* crow = p.C + (long)b * p.c_bs + ((long)(t0 + j) * p.Cm + m) * p.Cn + n0;
                for (int c0 = 0; c0 < p.BN; c0 += 32) {
                    float v[32];
                    tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * p.BN + c0), v);
#ifdef B200_DBG_NO_STORE
                    if (v[0] != 12345.678f) continue;
#endif
                    if (m < p.Cm) {
                        if (vec && n0 + c0 + 32 <= p.Cn) {
#pragma unroll
                            for (int jj = 0; jj < 32; jj += 4) red_add_v4(crow + c0 + jj, v[jj], v[jj + 1], v[jj + 2], v[jj + 3]);
                        } else {
#pragma unroll
                            for (int jj = 0; jj < 32; ++jj)
                                if (n0 + c0 + jj < p.Cn) atomicAdd(crow + c0 + jj, v[jj]);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, tcols);
}

__global__ void split_bf16_kernel(const float4* __restrict__ x, uint2* __restrict__ hi, uint2* __restrict__ lo, long n4) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        const float4 v = x[i];
        const float f[4] = {v.x, v.y, v.z, v.w};
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            h[j] = __float2bfloat16_rn(f[j]);
            l[j] = __float2bfloat16_rn(f[j] - __bfloat162float(h[j]));
        }
        hi[i] = *reinterpret_cast<uint2*>(h);
        if (lo) lo[i] = *reinterpret_cast<uint2*>(l);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)f;
    }
    return fn;
}

// bf16 NHWC tensor {C, W, H, N}; box {64, tw*s, th*s, 1} with traversal stride s on W and H
int make_map_nhwc(CUtensorMap* m, const void* ptr, int n, int h, int w, int c, int tw, int th, int s) {
    EncodeTiledFn f = encode_fn();
    B200_REQUIRE(f, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dim[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t str[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(tw * s), (cuuint32_t)(th * s), 1};
    cuuint32_t es[4] = {1, (cuuint32_t)s, (cuuint32_t)s, 1};
    CUresult r = f(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed for an NHWC activation map");
    return 0;
}

// bf16 row-major matrix {inner, rows}; box {64, box_rows}
int make_map_2d(CUtensorMap* m, const void* ptr, long rows, int inner, int box_rows) {
    EncodeTiledFn f = encode_fn();
    B200_REQUIRE(f, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dim[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
    cuuint64_t str[1] = {(cuuint64_t)inner * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = f(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed for a weight map");
    return 0;
}

void pick_tile(int Hi, int Wi, int& th, int& tw) {
    long best = -1;
    for (int w = 128; w >= 1; w >>= 1) {
        const int h = 128 / w;
        const long tiles = (long)((Wi + w - 1) / w) * ((Hi + h - 1) / h);
        if (best < 0 || tiles < best) { best = tiles; tw = w; th = h; }
    }
}

// Split-K factor for layers with too few output tiles to fill the GPU (the 4x4 .. 32x32 blocks stream 9.4 MB of weights
// through a handful of CTAs otherwise): aim at one CTA per SM (a single wave), keep at least 2 k-blocks per CTA.
int pick_ksplit(int ctas, int nk) {
    if (ctas >= 74 || nk < 4) return 1;
    int ks = 148 / ctas;                 // floor: one wave, never 148 + a few stragglers
    if (ks > nk / 2) ks = nk / 2;
    return ks < 1 ? 1 : ks;
}

int pick_bn(int n) {           // N tile of the K-major-B kernels (a multiple of 32, MMA N)
    if (n <= 32) return 32;
    if (n <= 64) return 64;
    if (n <= 96) return 96;
    return 128;
}

// CTA-pair mode (cta_group::2) switch: -1 = take B200EG3D_CONV_PAIR from the environment on first use (default on)
int g_conv_pair = -1;
bool conv_pair_enabled() {
    if (g_conv_pair < 0) { const char* e = getenv("B200EG3D_CONV_PAIR"); g_conv_pair = (e && e[0] == '0') ? 0 : 1; }
    return g_conv_pair == 1;
}

// how many CTA pairs of this kernel the device runs at once (one CTA per SM: 74 on a full B200); 0 if clusters cannot be placed
template <bool B_MN>
int max_pairs() {
    static int cached[B200_MAX_DEVICES] = {};
    const int dev = b200_current_device();
    if (!cached[dev]) {
        if (cudaFuncSetAttribute(conv_tc_pix_kernel<B_MN, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) {
            cudaGetLastError(); cached[dev] = -1; return 0;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * 148); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = SMEM_BYTES;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, conv_tc_pix_kernel<B_MN, 2>, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
        cached[dev] = n > 0 ? n : -1;
    }
    return cached[dev] > 0 ? cached[dev] : 0;
}

// Pair mode pays when the list still fills the machine with half as many walkers: at least 3/4 of the pairs get an item.
inline bool pair_fills(long pair_items) { return pair_items >= 56; }

template <bool B_MN>
int launch_pix(TcPixParams& p, int batch, cudaStream_t st) {
    p.ntn = (p.N + p.BN - 1) / p.BN;
    p.nwork = p.cta_start[p.ncls] * p.ntn * batch;
    p.stage_bytes = (A_BYTES + (p.BN / p.cg) * TILE_K * 2) * (p.npass == 3 ? 2 : 1);
    p.nstages = RING_BYTES / p.stage_bytes < MAX_ST ? RING_BYTES / p.stage_bytes : MAX_ST;
    if (p.cg == 2) {
        B200_FUNC_ATTR_ONCE((conv_tc_pix_kernel<B_MN, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        const int pairs = max_pairs<B_MN>();
        B200_REQUIRE(pairs > 0, "conv_tc: CTA pairs cannot be scheduled on this device");
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * (p.nwork < pairs ? p.nwork : pairs)); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = st;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = b200_pdl_enabled() ? 2 : 1;
        B200_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_pix_kernel<B_MN, 2>, p));
        B200_CHECK_LAUNCH();
        return 0;
    }
    B200_FUNC_ATTR_ONCE((conv_tc_pix_kernel<B_MN, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    const int sms = b200_sm_count();
    const int grid = p.nwork < sms ? p.nwork : sms;
    B200_CUDA(launch_pdl(conv_tc_pix_kernel<B_MN, 1>, dim3(grid), dim3(192), SMEM_BYTES, st, p));
    B200_CHECK_LAUNCH();
    return 0;
}

}  // namespace

// CTA-pair (cta_group::2) tiles of the forward / dgrad kernels on (1, default; env B200EG3D_CONV_PAIR=0 disables) or off (0):
// the single-CTA kernel stays available as the cross-check.  Returns the previous setting.
B200_API int b200_set_conv_pair(int on) {
    const int prev = conv_pair_enabled() ? 1 : 0;
    g_conv_pair = on ? 1 : 0;
    return prev;
}

// 1 when the tcgen05 path handles this shape; kind: 0 forward, 1 dgrad, 2 wgrad
B200_API int b200_conv_tc_supported(int kind, int h, int w, int cin, int cout, int ksize, int up) {
    if (!(ksize == 1 || ksize == 3) || !(up == 1 || (up == 2 && ksize == 3))) return 0;
    // TMA needs 16-byte global strides: channel counts must be multiples of 8 (bf16).  Ragged K / N edges are
    // handled by TMA zero-fill and by column masking in the epilogue.
    if (kind == 0) return cin % 8 == 0;
    if (kind == 1) return cout % 8 == 0 && cin % 8 == 0;
    if (kind == 2) return cout % 8 == 0 && cin % 8 == 0;
    return 0;
}

B200_API int b200_split_bf16(const float* x, void* hi, void* lo, long count, void* stream) {
    B200_REQUIRE(count % 4 == 0, "split_bf16: element count must be a multiple of 4");
    if (count == 0) return 0;
    const long n4 = count / 4;
    const int blocks = (int)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
    split_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)x, (uint2*)hi, (uint2*)lo, n4);
    B200_CHECK_LAUNCH();
    return 0;
}

// Forward conv on split-bf16 operands.  x_* [n][h][w][cin] bf16, w_* [n][taps][cout][cin] bf16, y fp32 NHWC
// (up == 2: y is the (2h+1)x(2w+1) transposed-conv grid, as in b200_conv_fwd).  npass: 1 (hi*hi) or 3.
struct FwdEpilogue {
    const float* bias; const float* noise; const float* strength; long noise_bs; float alpha, act_gain, clamp; void* z_hi; void* z_lo;
};

static int fwd_ksplit(int n, int h, int w, int cin, int cout, int ksize) {       // split-K factor of an up == 1 forward launch
    int th, tw;
    pick_tile(h, w, th, tw);
    const int tiles = ((w + tw - 1) / tw) * ((h + th - 1) / th), bn = pick_bn(cout);
    return pick_ksplit(tiles * ((cout + bn - 1) / bn) * n, ksize * ksize * ((cin + 63) / 64));
}

static int conv_fwd_tc_impl(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, float* y, int n, int h,
                            int w, int cin, int cout, int ksize, int up, int npass, const FwdEpilogue* ep, int prezeroed, void* stream,
                            int* ksplit_out = nullptr) {
    B200_REQUIRE(b200_conv_tc_supported(0, h, w, cin, cout, ksize, up), "conv_fwd_tc: unsupported shape");
    B200_REQUIRE(npass == 1 || (npass == 3 && x_lo && w_lo), "conv_fwd_tc: npass must be 1, or 3 with lo operands");
    cudaStream_t st = (cudaStream_t)stream;
    const int taps = ksize * ksize;
    TcPixParams p{};
    p.kchunks = (cin + 63) / 64; p.npass = npass; p.s = 1; p.b_rows_per_tap = cout; p.b_taps = taps;
    p.BN = pick_bn(cout); p.N = cout; p.C = y; p.ldc = cout; p.cg = 1;
    const int ntn = (cout + p.BN - 1) / p.BN;
    int tiles[4];
    if (up == 1) {
        p.ncls = 1;
        TcClass& c = p.c[0];
        c.Hi = h; c.Wi = w; pick_tile(h, w, c.th, c.tw); c.tiles_x = (w + c.tw - 1) / c.tw;
        c.ntaps = taps; c.ooy = c.oox = 0;
        for (int t = 0; t < taps; ++t) { c.dy[t] = t / ksize - ksize / 2; c.dx[t] = t % ksize - ksize / 2; c.wt[t] = t; }
        p.c_bs = (long)h * w * cout; p.Wo = w; p.osy = p.osx = 1;
        tiles[0] = c.tiles_x * ((h + c.th - 1) / c.th);
        p.ksplit = pick_ksplit(tiles[0] * ntn * n, taps * p.kchunks);
    } else {
        // stride-2 transposed convolution: four output-parity classes (4, 2, 2, 1 taps), one launch
        p.ncls = 4;
        p.c_bs = (long)(2 * h + 1) * (2 * w + 1) * cout; p.Wo = 2 * w + 1; p.osy = p.osx = 2;
        int total = 0;
        for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
                TcClass& c = p.c[py * 2 + px];
                c.Hi = h + 1 - py; c.Wi = w + 1 - px; pick_tile(c.Hi, c.Wi, c.th, c.tw); c.tiles_x = (c.Wi + c.tw - 1) / c.tw;
                c.ooy = py; c.oox = px;
                int t = 0;
                for (int kh = py; kh < 3; kh += 2)
                    for (int kw = px; kw < 3; kw += 2) { c.dy[t] = -(kh >> 1); c.dx[t] = -(kw >> 1); c.wt[t] = kh * 3 + kw; ++t; }
                c.ntaps = t;
                tiles[py * 2 + px] = c.tiles_x * ((c.Hi + c.th - 1) / c.th);
                total += tiles[py * 2 + px];
            }
        p.ksplit = pick_ksplit(total * ntn * n, p.kchunks);      // the (1,1) class has a single tap: at least kchunks k-blocks
    }
    // CTA pairs for the layers that fill the machine without split-K; 256-wide N tiles when the list stays long enough
    if (ksplit_out) { *ksplit_out = p.ksplit; return 0; }        // planning query (b200_conv_tc_ksplit): no launch
    if (p.ksplit == 1 && conv_pair_enabled() && max_pairs<false>() > 0) {
        long ptiles = 0;
        for (int k = 0; k < p.ncls; ++k) ptiles += (tiles[k] + 1) / 2;
        if (cout >= 256 && pair_fills(ptiles * ((cout + 255) / 256) * n)) { p.cg = 2; p.BN = 256; }
        else if (pair_fills(ptiles * ntn * n)) p.cg = 2;
    }
    for (int i = 0; i < (npass == 3 ? 2 : 1); ++i)
        if (int e = make_map_2d(&p.tmB[i], i ? w_lo : w_hi, (long)n * taps * cout, cin, p.BN / p.cg)) return e;
    p.cta_start[0] = 0;
    for (int k = 0; k < p.ncls; ++k) {
        p.c[k].ntiles = tiles[k];
        p.cta_start[k + 1] = p.cta_start[k] + (p.cg == 2 ? (tiles[k] + 1) / 2 : tiles[k]) * p.ksplit;
        for (int i = 0; i < (npass == 3 ? 2 : 1); ++i)
            if (int e = make_map_nhwc(&p.tmA[k][i], i ? x_lo : x_hi, n, h, w, cin, p.c[k].tw, p.c[k].th, 1)) return e;
    }
    if (ep) {
        B200_REQUIRE(up == 1 && p.ksplit == 1 && cout % 32 == 0 && ep->bias && ep->z_hi && (!ep->noise || ep->strength),
                     "conv_fwd_tc_act: epilogue fusion needs up == 1, no split-K, cout % 32 == 0 (ask b200_conv_tc_act_fusable first)");
        p.act = 1; p.bias = ep->bias; p.noise = ep->noise; p.strength = ep->strength; p.noise_bs = ep->noise_bs;
        p.alpha = ep->alpha; p.act_gain = ep->act_gain; p.clamp = ep->clamp;
        p.z_hi = (__nv_bfloat16*)ep->z_hi; p.z_lo = (__nv_bfloat16*)ep->z_lo;
    }
    if (p.ksplit > 1 && !prezeroed) B200_CUDA(cudaMemsetAsync(y, 0, sizeof(float) * (size_t)n * p.c_bs, st));
    return launch_pix<false>(p, n, st);
}

// prezeroed != 0: y is already zero (one fill for all split-K outputs of a network, made by the caller): a split-K launch then adds its
// partial sums without clearing y first.  Launches that do not split K overwrite y either way.
B200_API int b200_conv_fwd_tc(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, float* y, int n, int h,
                              int w, int cin, int cout, int ksize, int up, int npass, int prezeroed, void* stream) {
    return conv_fwd_tc_impl(x_hi, x_lo, w_hi, w_lo, y, n, h, w, cin, cout, ksize, up, npass, nullptr, prezeroed, stream);
}

// Split-K factor the forward (kind 0) / dgrad (kind 1) launch of this shape uses (1: the output is overwritten, > 1: partial sums are
// added with red.global and the output has to start at zero).  Lets a caller pool the zero-initialised outputs of a whole network.
static int dgrad_ksplit(int n, int h, int w, int cin, int cout, int ksize);
B200_API int b200_conv_tc_ksplit(int kind, int n, int h, int w, int cin, int cout, int ksize, int up) {
    if (!b200_conv_tc_supported(kind, h, w, cin, cout, ksize, up) || n <= 0) return 1;
    if (kind == 1) return dgrad_ksplit(n, h, w, cin, cout, ksize);
    if (kind != 0) return 1;
    int ks = 1;
    if (conv_fwd_tc_impl(nullptr, nullptr, nullptr, nullptr, nullptr, n, h, w, cin, cout, ksize, up, 1, nullptr, 0, nullptr, &ks)) return 1;
    return ks;
}

// 1 when b200_conv_fwd_tc_act can fuse the SynthesisLayer epilogue into the convolution's TMEM drain for this shape.
B200_API int b200_conv_tc_act_fusable(int n, int h, int w, int cin, int cout, int ksize) {
    if (!b200_conv_tc_supported(0, h, w, cin, cout, ksize, 1) || cout % 32 != 0) return 0;
    // The epilogue of tile i runs under the main loop of tile i+1: that hides it only when a tile has enough k-blocks.  Measured:
    // with 9 / 18 k-blocks per tile (64- / 128-channel layers) the fused kernel loses what the separate epilogue launch cost.
    static int min_kb = -1;
    if (min_kb < 0) { const char* e = getenv("B200EG3D_ACT_FUSE_MIN_KBLOCKS"); min_kb = e ? atoi(e) : 36; }
    if (ksize * ksize * ((cin + 63) / 64) < min_kb) return 0;
    return fwd_ksplit(n, h, w, cin, cout, ksize) == 1;
}

// up == 1 forward convolution with the layer epilogue applied while the accumulator leaves tensor memory:
//   z = clamp(lrelu_alpha(conv + noise[pix] * *strength + bias[c]) * act_gain, +-clamp)   (networks_stylegan2.py:318-329)
// written as fp32 z (may be NULL: the split pair is then the only copy) and as split-bf16 z_hi / z_lo (z_lo may be NULL).  noise: [h*w] (noise_bs == 0) or [n][h*w]; may be NULL.
B200_API int b200_conv_fwd_tc_act(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, float* z, void* z_hi,
                                  void* z_lo, const float* bias, const float* noise, const float* strength, long noise_bs, int n,
                                  int h, int w, int cin, int cout, int ksize, int npass, float alpha, float act_gain, float clamp,
                                  void* stream) {
    B200_REQUIRE(z_hi, "conv_fwd_tc_act: the split-bf16 output is required (z may be NULL)");
    FwdEpilogue ep{bias, noise, strength, noise_bs, alpha, act_gain, clamp, z_hi, z_lo};
    return conv_fwd_tc_impl(x_hi, x_lo, w_hi, w_lo, z, n, h, w, cin, cout, ksize, 1, npass, &ep, 0, stream);
}

// dgrad on split-bf16 operands.  dy_* bf16 ([h][w][cout], or the (2h+1)x(2w+1) grid when up == 2), w_* as above, dx fp32.
static int dgrad_ksplit(int n, int h, int w, int cin, int cout, int ksize) {
    int th, tw;
    pick_tile(h, w, th, tw);
    const int tiles0 = ((w + tw - 1) / tw) * ((h + th - 1) / th), bn = cin > 64 ? 128 : 64;
    return pick_ksplit(tiles0 * ((cin + bn - 1) / bn) * n, ksize * ksize * ((cout + 63) / 64));
}

B200_API int b200_conv_dgrad_tc(const void* dy_hi, const void* dy_lo, const void* w_hi, const void* w_lo, float* dx, int n,
                                int h, int w, int cin, int cout, int ksize, int up, int npass, int prezeroed, void* stream) {
    B200_REQUIRE(b200_conv_tc_supported(1, h, w, cin, cout, ksize, up), "conv_dgrad_tc: unsupported shape");
    B200_REQUIRE(npass == 1 || (npass == 3 && dy_lo && w_lo), "conv_dgrad_tc: npass must be 1, or 3 with lo operands");
    cudaStream_t st = (cudaStream_t)stream;
    const int taps = ksize * ksize;
    const int hs = up == 1 ? h : 2 * h + 1, ws = up == 1 ? w : 2 * w + 1;
    TcPixParams p{};
    p.kchunks = (cout + 63) / 64; p.npass = npass; p.s = up; p.b_rows_per_tap = cout; p.b_taps = taps;
    p.BN = cin > 64 ? 128 : 64; p.N = cin; p.C = dx; p.ldc = cin; p.c_bs = (long)h * w * cin; p.cg = 1;
    p.ncls = 1;
    TcClass& c = p.c[0];
    c.Hi = h; c.Wi = w; pick_tile(h, w, c.th, c.tw); c.tiles_x = (w + c.tw - 1) / c.tw;
    c.ntaps = taps; p.Wo = w; p.osy = p.osx = 1; c.ooy = c.oox = 0;
    for (int t = 0; t < taps; ++t) {
        const int kh = t / ksize, kw = t % ksize;
        c.dy[t] = up == 1 ? ksize / 2 - kh : kh; c.dx[t] = up == 1 ? ksize / 2 - kw : kw; c.wt[t] = t;
    }
    for (int i = 0; i < (npass == 3 ? 2 : 1); ++i) {
        if (int e = make_map_nhwc(&p.tmA[0][i], i ? dy_lo : dy_hi, n, hs, ws, cout, c.tw, c.th, up)) return e;
        if (int e = make_map_2d(&p.tmB[i], i ? w_lo : w_hi, (long)n * taps * cout, cin, 64)) return e;
    }
    const int tiles0 = c.tiles_x * ((h + c.th - 1) / c.th);
    c.ntiles = tiles0;
    p.ksplit = pick_ksplit(tiles0 * ((cin + p.BN - 1) / p.BN) * n, taps * p.kchunks);
    // CTA pairs: each CTA holds a 64-multiple of the MN-major B tile's columns, so N tiles of 128 or 256
    if (p.ksplit == 1 && cin >= 128 && conv_pair_enabled() && max_pairs<true>() > 0) {
        const long ptiles = (tiles0 + 1) / 2;
        if (cin >= 256 && pair_fills(ptiles * ((cin + 255) / 256) * n)) { p.cg = 2; p.BN = 256; }
        else if (pair_fills(ptiles * ((cin + 127) / 128) * n)) { p.cg = 2; p.BN = 128; }
    }
    p.cta_start[0] = 0; p.cta_start[1] = (p.cg == 2 ? (tiles0 + 1) / 2 : tiles0) * p.ksplit;
    if (p.ksplit > 1 && !prezeroed) B200_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)n * h * w * cin, st));
    return launch_pix<true>(p, n, st);
}

// wgrad on bf16 operands: dwmod[n][taps][cout][cin] (fp32; overwritten, or added to when accumulate != 0) = sum over pixels of dy (x) x.
// x_* [n][h][w][cin], dy_* [n][h][w][cout] (up == 2: the (2h+1)x(2w+1) transposed-conv grid).
B200_API int b200_conv_wgrad_tc(const void* x_hi, const void* x_lo, const void* dy_hi, const void* dy_lo, float* dwmod, int n,
                                int h, int w, int cin, int cout, int ksize, int up, int npass, int accumulate, void* stream) {
    B200_REQUIRE(b200_conv_tc_supported(2, h, w, cin, cout, ksize, up), "conv_wgrad_tc: unsupported shape");
    B200_REQUIRE(npass == 1 || (npass == 3 && x_lo && dy_lo), "conv_wgrad_tc: npass must be 1, or 3 with lo operands");
    cudaStream_t st = (cudaStream_t)stream;
    const int taps = ksize * ksize;
    const int hs = up == 1 ? h : 2 * h + 1, ws = up == 1 ? w : 2 * w + 1;
    TcWgradParams p{};
    p.npass = npass; p.ntaps = taps; p.Cm = cout; p.Cn = cin; p.BN = cin > 64 ? 128 : 64;
    p.C = dwmod; p.c_bs = (long)taps * cout * cin; p.sA = up; p.sB = 1;
    // 64-pixel patches over the h x w iteration grid
    { long best = -1; for (int tw = 64; tw >= 1; tw >>= 1) { const int th = 64 / tw; const long tl = (long)((w + tw - 1) / tw) * ((h + th - 1) / th);
        if (best < 0 || tl < best) { best = tl; p.tw = tw; p.th = th; } } }
    p.tiles_x = (w + p.tw - 1) / p.tw; p.ntiles = p.tiles_x * ((h + p.th - 1) / p.th);
    for (int t = 0; t < taps; ++t) {
        const int kh = t / ksize, kw = t % ksize;
        p.dAy[t] = up == 1 ? 0 : kh; p.dAx[t] = up == 1 ? 0 : kw;
        p.dBy[t] = up == 1 ? kh - ksize / 2 : 0; p.dBx[t] = up == 1 ? kw - ksize / 2 : 0;
    }
    for (int i = 0; i < (npass == 3 ? 2 : 1); ++i) {
        if (int e = make_map_nhwc(&p.tmA[i], i ? dy_lo : dy_hi, n, hs, ws, cout, p.tw, p.th, up)) return e;
        if (int e = make_map_nhwc(&p.tmB[i], i ? x_lo : x_hi, n, h, w, cin, p.tw, p.th, 1)) return e;
    }
    const int mt = (cout + TILE_M - 1) / TILE_M, nt = (cin + p.BN - 1) / p.BN;
    // single-pass 3x3: one kernel row (3 taps) per CTA -- the operand that does not shift with the tap is fetched once
    p.tg = (npass == 1 && taps == 9) ? 3 : 1;
    p.shared_a = up == 1;
    const int base_ctas = mt * nt * (taps / p.tg) * n;
    const int target = p.tg > 1 ? 148 : 296;      // fewer, heavier CTAs when a CTA owns three taps: less split-K atomic traffic
    int ksplit = p.tg > 1 ? target / base_ctas : (target + base_ctas - 1) / base_ctas;   // tg > 1: stay within one wave
    if (ksplit > p.ntiles) ksplit = p.ntiles;
    if (ksplit < 1) ksplit = 1;
    p.ksplit = ksplit;
    if (!accumulate) B200_CUDA(cudaMemsetAsync(dwmod, 0, sizeof(float) * (size_t)n * taps * cout * cin, st));
    B200_FUNC_ATTR_ONCE(conv_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    dim3 grid(mt, nt, n * (taps / p.tg) * ksplit);
    B200_CUDA(launch_pdl(conv_tc_wgrad_kernel, grid, dim3(192), SMEM_BYTES, st, p));
    B200_CHECK_LAUNCH();
    return 0;
}
