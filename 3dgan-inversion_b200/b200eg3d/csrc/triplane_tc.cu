// Fused tri-plane sampling + OSG decoder, tcgen05 generation (sm_100a).
//
// Same math as triplane.cu (renderer.py:39-66 sample_from_planes, triplane.py:124-136 OSGDecoder), re-organised as a
// warp-specialised pipeline over tiles of 128 sample points so that the decoder runs on the 5th-generation tensor cores with
// its accumulators in tensor memory and the texel gather never waits for it:
//
//   gather warps (8)     point -> 3 plane coordinates -> 12 texel lines (LDG.128, 8 lanes per point) -> mean feature, written
//                        as split bf16 (hi | lo in one 128-byte row) straight into the swizzled K-major A tile of layer 1
//   MMA warp (1 lane)    layer 1: D1[128 x 64] = [F_hi F_lo][W1_hi W1_hi]^T + F_hi W1_lo^T    (6 tcgen05.mma, K = 16 each)
//                        layer 2: D2[128 x 48] = h_hi W2_hi^T + h_hi W2_lo^T + h_lo W2_hi^T     (12 tcgen05.mma)
//   consumer warps (2x4) thread = point = TMEM lane: tcgen05.ld D1 -> +b1 -> softplus -> split bf16 -> A tile of layer 2;
//                        tcgen05.ld D2 -> +b2 -> sigmoid -> staged through shared memory -> coalesced stores
//
// Two tiles are in flight per stage (the two consumer sets alternate), every hand-off is an mbarrier, tensor-core completion is
// signalled with tcgen05.commit.  Split-bf16 operands (x = hi + lo, three products, fp32 accumulate) keep the decoder at
// fp32-equivalent accuracy, like the convolution stack.
//
// Work order: in ray mode the rays are walked column by column and every CTA owns a contiguous range of tiles, see
// tri::map_point (the XZ / ZX texel lines of a pixel column stay in that SM's L1).
#include "triplane_common.cuh"
#include "tc_common.cuh"

namespace {
using namespace tri;
using namespace tc;

constexpr int TILE = 128, OUTP = 48;
constexpr int NCONS = 8, MMA_WARP = 8, GATHER0 = 9, NGRP = 4, NG = 4 * NGRP;   // 4 gather groups of 4 warps = 4 tiles being gathered
constexpr int FW_THREADS = (NCONS + 1 + NG) * 32;              // 672

// shared-memory map (bytes from a 1024-aligned base)
constexpr int OFF_A1 = 0;                                      // NGRP stages x [128 rows][hi 32 ch | lo 32 ch] bf16
constexpr int OFF_A2 = OFF_A1 + NGRP * 16384;                  // 2 sets x ([128][64] hi | [128][64] lo) bf16
constexpr int OFF_W1A = OFF_A2 + 2 * 32768;                    // [64 units][W1_hi | W1_hi]
constexpr int OFF_W1B = OFF_W1A + 8192;                        // [64 units][W1_lo | 0]
constexpr int OFF_W2H = OFF_W1B + 8192;                        // [48 outputs][64 units] hi (rows 33..47 zero)
constexpr int OFF_W2L = OFF_W2H + 6144;
constexpr int OFF_SS = OFF_W2L + 6144;                         // per gather warp: bilinear set-up [32][SP] words
constexpr int OFF_BIAS = OFF_SS + NG * 32 * SP * 4;            // b1[64] | b2[48]
constexpr int OFF_BAR = OFF_BIAS + 512;                        // mbarriers (see bar_*) + TMEM slot
constexpr int FW_SMEM = OFF_BAR + 128 + 1024;                  // + alignment slack
constexpr int TM_COLS = 256;                                   // per set: D1 at +0 (64 columns), D2 at +64 (48 columns)

__device__ __forceinline__ uint32_t sw128(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
// 8 floats -> 8 bf16 (hi) and the bf16 of the remainders (lo)
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = pack2(v[2 * i], v[2 * i + 1]);
        l[i] = pack2(v[2 * i] - bf_lo(h[i]), v[2 * i + 1] - bf_hi(h[i]));
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// Activations in base 2 (MUFU.EX2 / LG2 / RCP are the native special functions); the conversion factors live in the weights:
//   layer 1 is staged as W1 log2(e), b1 log2(e), so the tensor core delivers y = u log2(e) and
//       softplus(u) = ln2 * lg2(1 + 2^y)          -> the kernel keeps h' = lg2(1 + 2^y), ln2 is folded into W2
//   the colour rows of layer 2 are staged as -W2 (= -log2(e) * ln2 * W2), -log2(e) b2, so the tensor core delivers z = -o log2(e) and
//       sigmoid(o) = 1 / (1 + 2^z)
// Equal to torch's softplus (threshold 20) / sigmoid to ~1e-7 absolute.  y is clamped below 2^126 so that huge pre-activations
// stay finite (lg2(1 + 2^y) = y exactly there).
__device__ __forceinline__ float softplus2(float y) { return __log2f(1.f + exp2f(fminf(y, 126.f))); }
__device__ __forceinline__ float sigmoid2(float z) { return __fdividef(1.f, 1.f + exp2f(fminf(z, 126.f))); }
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Decoder weights -> swizzled K-major bf16 B tiles (gains folded in; networks_stylegan2.py:111-112), biases -> fp32.
__device__ void stage_decoder_weights(const TriplaneParams& p, uint8_t* sm) {
    for (int i = threadIdx.x; i < HID * 64; i += blockDim.x) {
        const int j = i >> 6, e = i & 63, c = e & 31;
        const float w = p.W1[j * C + c] * (p.w1g * LOG2E);
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
        const uint32_t o = sw128(j, e >> 3) + (e & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(sm + OFF_W1A + o) = hi;
        *reinterpret_cast<__nv_bfloat16*>(sm + OFF_W1B + o) = e < 32 ? lo : __float2bfloat16_rn(0.f);
    }
    for (int i = threadIdx.x; i < OUTP * 64; i += blockDim.x) {
        const int k = i >> 6, j = i & 63;
        const float w = k < OUT ? p.W2[k * HID + j] * (k == 0 ? p.w2g * LN2 : -p.w2g) : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
        const uint32_t o = sw128(k, j >> 3) + (j & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(sm + OFF_W2H + o) = hi;
        *reinterpret_cast<__nv_bfloat16*>(sm + OFF_W2L + o) = lo;
    }
    float* bias = reinterpret_cast<float*>(sm + OFF_BIAS);
    for (int i = threadIdx.x; i < HID; i += blockDim.x) bias[i] = p.b1[i] * (p.b1g * LOG2E);
    for (int i = threadIdx.x; i < OUTP; i += blockDim.x) bias[HID + i] = i < OUT ? p.b2[i] * (i == 0 ? p.b2g : -p.b2g * LOG2E) : 0.f;
}

// Gather the mean feature of 32 points (rows row0 .. row0+31 of the tile) into the layer-1 A tile as split bf16.
// Eight lanes share a point (4 channels each, one LDG.128 per texel): one warp instruction fetches the same texel slot of four
// points = four full 128-byte lines.  Row layout: chunks 0-3 = hi channels 0..31, chunks 4-7 = lo channels 0..31.
// fsave (may be null): global image of the tile (same swizzled layout) kept for the backward, which then reloads the features
// with one bulk copy per tile instead of gathering the 12 texel lines per point again.
__device__ __forceinline__ void gather_to_tile(const float* __restrict__ pl, const float* ss, uint8_t* a1, int row0, int lane,
                                               uint8_t* __restrict__ fsave = nullptr) {
    const int pt = lane >> 3, l8 = lane & 7;
    const float* pc = pl + l8 * 4;
#pragma unroll 2
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
        const int row = row0 + q0 + pt;
        const uint32_t h01 = pack2(acc.x, acc.y), h23 = pack2(acc.z, acc.w);
        const uint32_t l01 = pack2(acc.x - bf_lo(h01), acc.y - bf_hi(h01)), l23 = pack2(acc.z - bf_lo(h23), acc.w - bf_hi(h23));
        // lane pair (2c, 2c+1) holds channels 8c..8c+7: the even lane collects the 16-byte hi chunk, the odd lane the lo chunk, so
        // every lane issues ONE conflict-free STS.128 (the 8 lanes of a point cover its whole 128-byte row)
        const bool odd = l8 & 1;
        const uint32_t rx = __shfl_xor_sync(0xffffffffu, odd ? h01 : l01, 1), ry = __shfl_xor_sync(0xffffffffu, odd ? h23 : l23, 1);
        const uint4 v = odd ? make_uint4(rx, ry, l01, l23) : make_uint4(h01, h23, rx, ry);
        const uint32_t off = sw128(row, (odd ? 4 : 0) + (l8 >> 1));
        *reinterpret_cast<uint4*>(a1 + off) = v;
        if (fsave) *reinterpret_cast<uint4*>(fsave + off) = v;         // the 8 lanes of a point write one full 128-byte line
    }
}


// mbarrier addresses (8 bytes each) relative to OFF_BAR; plain arithmetic so that runtime stage indices stay in registers
__device__ __forceinline__ uint32_t bar_a1_full(uint32_t b, int s) { return b + 8u * s; }               // NGRP, count 4
__device__ __forceinline__ uint32_t bar_a1_empty(uint32_t b, int s) { return b + 8u * (NGRP + s); }      // NGRP, count 1 (tcgen05.commit)
__device__ __forceinline__ uint32_t bar_d1_full(uint32_t b, int s) { return b + 8u * (2 * NGRP + s); }   // 2
__device__ __forceinline__ uint32_t bar_a2_full(uint32_t b, int s) { return b + 8u * (2 * NGRP + 2 + s); }   // 2, count 4
__device__ __forceinline__ uint32_t bar_d2_full(uint32_t b, int s) { return b + 8u * (2 * NGRP + 4 + s); }   // 2
constexpr int BAR_TMEM_SLOT = 8 * (2 * NGRP + 6);

// Producer / consumer hand-off between two groups of 128 threads on a hardware named barrier: the producers bar.arrive (do not
// block), the consumers bar.sync (block until all `count` threads of both groups have arrived).  Used for the thread-to-thread
// d_f staging hand-off (consumer set <-> its scatter group), where both sides are ordinary shared-memory accesses; the
// hand-offs to and from the tensor core use mbarriers (tcgen05.commit can only signal those).
constexpr int NB_DF_FULL = 1, NB_DF_FREE = 3;       // + set; barrier 0 is __syncthreads
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// Wait that is expected to be long (a producer several pipeline stages away): back off between probes so that the spinning
// warp does not take issue slots from the warps it is waiting for.
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok) __nanosleep(64);
    } while (!ok);
}

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {          // non-blocking probe
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

// Probe by lane 0, result broadcast: the MMA warps below run CONVERGED (descriptor arithmetic stays on the uniform datapath) and
// one elected lane issues -- a `lane == 0` branch around the whole loop makes the compiler rebuild every 64-bit descriptor in
// vector registers and move it across (R2UR), ~100 clocks per MMA, which sat on the tile's phase chain.
__device__ __forceinline__ bool mbar_test_warp(uint32_t bar, uint32_t parity, int lane) {
    uint32_t ok = 0;
    if (lane == 0) ok = mbar_test(bar, parity) ? 1u : 0u;
    return __shfl_sync(0xffffffffu, ok, 0) != 0;
}
// K-major operand, 128-byte swizzle, 8-row groups 1 KB apart: descriptor for start address `a` (high word `hi` = smem_desc_hi(1024))
__device__ __forceinline__ uint64_t kdesc(uint32_t a, uint32_t hi) { return desc_pack(smem_desc_lo(a, 0), hi); }

// SIGMA_ONLY: density queries on voxel grids (single_id_coach.py:118-140) write column 0 of the second layer only.
template <bool SIGMA_ONLY>
__global__ void __launch_bounds__(FW_THREADS, 1) triplane_fwd_tc_kernel(TriplaneParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const uint32_t sm_u = smem_u32(sm);
    const uint32_t B = sm_u + OFF_BAR;
    const int n = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NGRP; ++i) { mbar_init(bar_a1_full(B, i), 4); mbar_init(bar_a1_empty(B, i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_d1_full(B, i), 1); mbar_init(bar_a2_full(B, i), 4); mbar_init(bar_d2_full(B, i), 1); }
        fence_barrier_init();
    }
    if (warp == MMA_WARP) tmem_alloc(B + BAR_TMEM_SLOT, TM_COLS);
    stage_decoder_weights(p, sm);
    fence_proxy_async();                       // the weight tiles were written through the generic proxy; tcgen05.mma reads via the async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + OFF_BAR + BAR_TMEM_SLOT);

    const long ntiles = (p.P + TILE - 1) / TILE;
    const long per = (ntiles + gridDim.x - 1) / gridDim.x;
    const long t0 = (long)blockIdx.x * per;
    const int nloc = (int)max(0L, min(ntiles, t0 + per) - t0);
    const float* pl = p.planes + (long)n * p.hp * p.wp * PC;

    if (warp >= GATHER0) {
        // ------------------------------------------------------------------ gather: tiles lt = group, group + NGRP, ...
        const int g = warp - GATHER0, group = g >> 2, gq = g & 3;
        float* ss = reinterpret_cast<float*>(sm + OFF_SS) + g * 32 * SP;
        uint8_t* a1 = sm + OFF_A1 + group * 16384;
        int it = 0;
        for (int lt = group; lt < nloc; lt += NGRP, ++it) {
            float cx, cy, cz;
            point_coords32(p, n, map_point(p, (unsigned)((t0 + lt) * TILE + gq * 32 + lane)), cx, cy, cz);
            stage_setup(ss, lane, cx, cy, cz, p.hp, p.wp);
            mbar_wait(bar_a1_empty(B, group), (it & 1) ^ 1);           // layer 1 of the tile NGRP back has consumed this stage
            __syncwarp();
            gather_to_tile(pl, ss, a1, gq * 32, lane, p.f_save ? p.f_save + ((long)n * ntiles + t0 + lt) * 16384 : nullptr);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_a1_full(B, group));
        }
    } else if (warp == MMA_WARP) {
        // ------------------------------------------------------------------ tensor-core issue (one lane)
        // Two cursors: n1 = next tile whose layer 1 is to be issued, n2 = next tile whose layer 2 is to be issued.  Whichever
        // operand tile is ready first goes first, so a slow activation stage never holds back the next tile's layer 1.
        // D1[s] / D2[s] (s = tile parity) are free again once layer 2 of the tile two back has been issued: its a2_full arrival
        // follows the consumer's last read of both.
        {
            const uint32_t id1 = instr_desc_bf16(128, HID, 0, 0), id2 = instr_desc_bf16(128, OUTP, 0, 0);
            const uint32_t dhi = smem_desc_hi(1024);
            const bool three = p.fwd_passes == 3;
            int n1 = 0, n2 = 0;
            while (n2 < nloc) {
                bool did = false;
                if (n1 < nloc && n1 - n2 < 2) {
                    const int st = n1 % NGRP, it = n1 / NGRP, s = n1 & 1;
                    if (mbar_test_warp(bar_a1_full(B, st), it & 1, lane)) {
                        tc_fence_after();
                        const uint32_t a = sm_u + OFF_A1 + st * 16384, d1 = tmem + (uint32_t)(s * 128);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)        // [F_hi F_lo] x [W1_hi W1_hi]   (single pass: F_hi x W1_hi only)
                                if (k < 2 || three) umma_bf16(d1, kdesc(a + k * 32, dhi), kdesc(sm_u + OFF_W1A + k * 32, dhi), id1, k != 0);
                            if (three) {
#pragma unroll
                                for (int k = 0; k < 2; ++k)    // F_hi x W1_lo
                                    umma_bf16(d1, kdesc(a + k * 32, dhi), kdesc(sm_u + OFF_W1B + k * 32, dhi), id1, 1);
                            }
                            umma_commit(bar_a1_empty(B, st));
                            umma_commit(bar_d1_full(B, s));
                        }
                        __syncwarp();
                        ++n1;
                        did = true;
                    }
                }
                if (n2 < n1) {
                    const int s = n2 & 1, it = n2 >> 1;
                    if (mbar_test_warp(bar_a2_full(B, s), it & 1, lane)) {
                        tc_fence_after();
                        const uint32_t ah = sm_u + OFF_A2 + s * 32768, al = ah + 16384, d2 = tmem + (uint32_t)(s * 128 + 64);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t dah = kdesc(ah + k * 32, dhi), dbh = kdesc(sm_u + OFF_W2H + k * 32, dhi);
                                umma_bf16(d2, dah, dbh, id2, k != 0);
                                if (three) {
                                    umma_bf16(d2, dah, kdesc(sm_u + OFF_W2L + k * 32, dhi), id2, 1);
                                    umma_bf16(d2, kdesc(al + k * 32, dhi), dbh, id2, 1);
                                }
                            }
                            umma_commit(bar_d2_full(B, s));
                        }
                        __syncwarp();
                        ++n2;
                        did = true;
                    }
                }
                if (!did) __nanosleep(32);
            }
        }
    } else {
        // ------------------------------------------------------------------ consumers: set = tile parity, thread = point = TMEM lane
        const int set = warp >> 2, q = warp & 3, row = q * 32 + lane;
        const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(set * 128);
        uint8_t* a2h = sm + OFF_A2 + set * 32768;
        uint8_t* a2l = a2h + 16384;
        const float* b1s = reinterpret_cast<const float*>(sm + OFF_BIAS);
        const float* b2s = b1s + HID;
        // output staging, 36-float rows (conflict-free for row-per-thread writes and for row-contiguous reads), carved out of the
        // A2 rows this warp itself owns: rows 0..15 in its hi rows, 16..31 in its lo rows
        auto stage_row = [&](int r) -> float* {
            return reinterpret_cast<float*>((r < 16 ? a2h : a2l) + q * 4096 + (r & 15) * 144);
        };
        for (int lt = set; lt < nloc; lt += 2) {
            const int it = lt >> 1;
            const unsigned pp0 = (unsigned)((t0 + lt) * TILE + q * 32);   // first work point of this warp's 32 rows
            mbar_wait(bar_d1_full(B, set), it & 1);
            tc_fence_after();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float v[32];
                tmem_ld32(tq + half * 32, v);
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    float h[8];
#pragma unroll
                    for (int e = 0; e < 8; e += 4) {
                        const float4 bv = *reinterpret_cast<const float4*>(b1s + half * 32 + c8 * 8 + e);
                        h[e] = softplus2(v[c8 * 8 + e] + bv.x); h[e + 1] = softplus2(v[c8 * 8 + e + 1] + bv.y);
                        h[e + 2] = softplus2(v[c8 * 8 + e + 2] + bv.z); h[e + 3] = softplus2(v[c8 * 8 + e + 3] + bv.w);
                    }
                    uint4 hi, lo;
                    split8(h, hi, lo);
                    const uint32_t o = sw128(row, half * 4 + c8);
                    *reinterpret_cast<uint4*>(a2h + o) = hi;
                    *reinterpret_cast<uint4*>(a2l + o) = lo;
                }
            }
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_a2_full(B, set));
            // ---- layer 2 result
            mbar_wait(bar_d2_full(B, set), it & 1);
            tc_fence_after();
            float o[32], o32[8];
            tmem_ld32(tq + 64, o);
            if (!SIGMA_ONLY) tmem_ld8(tq + 96, o32);
            tc_fence_before();
            const int pi = map_point(p, pp0 + lane);
            if (pi >= 0) p.sigma[(long)n * p.P + pi] = o[0] + b2s[0];
            if (!SIGMA_ONLY) {
                // rgb = sigmoid(o[1..32]) * 1.002 - 0.001, staged so that the global stores are full 128-byte rows
                float* mine = stage_row(lane);
#pragma unroll
                for (int c4 = 0; c4 < 32; c4 += 4) {
                    float r[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int col = c4 + e + 1;
                        const float x = (col < 32 ? o[col] : o32[0]) + b2s[col];
                        r[e] = sigmoid2(x) * 1.002f - 0.001f;
                    }
                    *reinterpret_cast<float4*>(mine + c4) = make_float4(r[0], r[1], r[2], r[3]);
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = i * 4 + (lane >> 3);
                    const int pr = __shfl_sync(0xffffffffu, pi, r);
                    const float4 val = *reinterpret_cast<const float4*>(stage_row(r) + (lane & 7) * 4);
                    if (pr >= 0) *reinterpret_cast<float4*>(p.rgb + ((long)n * p.P + pr) * C + (lane & 7) * 4) = val;
                }
                __syncwarp();                       // staging rows are overwritten by the next tile's hidden activations
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem, TM_COLS);
}

// =====================================================================================================================
// Backward.  Per tile of 128 points the tensor core runs four phases, the consumer threads (thread = point) the steps between:
//
//   gather            F (split bf16, as in the forward) -> A1[stage];  bilinear set-up -> SS[stage] (kept for the scatter)
//   P1  D1 = F W1'^T                    S1  h' = lg2(1 + 2^(D1 + b1'))                    -> A2hi | A2lo (split bf16)
//   P2  D2 = h' W2'^T                   S2  s = 1/(1 + 2^(D2 + b2')), dO = d_rgb 1.002 s(1-s) | d_sigma   -> A3 (bf16; = A2lo)
//   P3  D3 = dO W2                      S3  d_a = D3 * 2^y / (1 + 2^y),  y = D1 + b1'  (D1 re-read from TMEM)  -> A2hi (bf16)
//       ACC2 += [h' ; 1]^T dO                                                         (dW2^T rows 0..63, db2 in rows 64..127)
//   P4  D4 = d_a W1                     S4  d_f = D4 -> staging (= A2lo) for the scatter warps
//       ACC1 += [F_hi ; F_lo ; 1]^T d_a                                               (dW1^T rows 0..63, db1 in rows 64..127)
//   scatter           d_f x bilinear weights -> red.global.add.v4.f32 into the plane gradient
//
// The weight / bias gradients are contractions over ALL points: the activation tiles that already sit in shared memory for the
// chain are read a second time as MN-major operands (K = the 128 points of the tile) and accumulated in two CTA-wide TMEM
// accumulators for the whole kernel.  The M = 128 instruction reads a second 64-row block at the descriptor's leading-dimension
// offset; it is pointed at a constant tile of ones, so rows 64..127 of the accumulators are the bias gradients (sum over
// points) at no extra cost.  One drain per CTA at the end (148 x ~5.3 k global atomics).
//
// Gradient operands (dO, d_a) are single bf16 (their rounding errors are independent per point and average out in every
// consumer: plane texels, weight sums); weights are split bf16 everywhere, the forward recompute is the forward's 3-pass.
constexpr int BW_NCONS = 8, BW_MMA_WARP = 8, BW_LOAD_WARP = 9, BW_SCATTER0 = 10, BW_NSCATTER = 8;
constexpr int BW_THREADS = (BW_NCONS + 2 + BW_NSCATTER) * 32;                   // 576
constexpr int BW_ST = 3;                                       // stages of A1 (tile lt uses stage lt % 3); recycled as soon as S1 has copied F out

constexpr int BO_A1 = 0;                                       // BW_ST x 16384
constexpr int BO_FC = BO_A1 + BW_ST * 16384;                   // 2 sets x [128][64] bf16: copy of the tile's F (hi half used) for the dW1 contraction
constexpr int BO_A2 = BO_FC + 2 * 16384;                       // 2 sets x (hi 16384 | lo 16384); lo doubles as A3 (dO) and as the d_rgb / d_f staging
constexpr int BO_W1A = BO_A2 + 2 * 32768;                      // [64][W1'_hi | W1'_hi]   (recompute: activations split, weights hi)
constexpr int BO_W2H = BO_W1A + 8192;                          // [48][64] W2'_hi
constexpr int BO_W2TH = BO_W2H + 6144;                         // [64 units][64 (k < 33 used)] true W2^T, hi / lo
constexpr int BO_W2TL = BO_W2TH + 8192;
constexpr int BO_W1TH = BO_W2TL + 8192;                        // [32 channels][64 units] true W1^T, hi / lo
constexpr int BO_W1TL = BO_W1TH + 4096;
constexpr int BO_ONES = BO_W1TL + 4096;                        // [16][64] bf16 ones: second MN block of every k-step of the weight-gradient A operands
constexpr int BO_SS = BO_ONES + 2048;                          // private set-up tiles of the scatter warps, [32][SP] words each
constexpr int SPX = 8;                                         // extra set-up words per point for the coordinate gradient
constexpr int BO_BIAS = BO_SS + BW_NSCATTER * 32 * (SP + SPX) * 4;   // b1'[64] | b2'[48]
constexpr int BO_BAR = BO_BIAS + 512;
constexpr int BW_SMEM = BO_BAR + 256 + 1024;
constexpr int BW_TM_COLS = 512;                                // set s at 192 s: D1 +0 (64), D2/D4 +64 (48), D3 +128 (64); ACC2 at 384 (48), ACC1 at 432 (64)
constexpr int TM_ACC2 = 384, TM_ACC1 = 432;
static_assert(BW_SMEM <= 227 * 1024, "backward kernel shared memory");

__device__ __forceinline__ uint32_t bb_a1_full(uint32_t b, int s) { return b + 8u * s; }                  // BW_ST, count 4
__device__ __forceinline__ uint32_t bb_a1_free(uint32_t b, int s) { return b + 8u * (BW_ST + s); }         // BW_ST, count 4 (consumer warps after copying F)
__device__ __forceinline__ uint32_t bb_set(uint32_t b, int which, int s) { return b + 8u * (2 * BW_ST + 2 * which + s); }
enum { BB_D1 = 0, BB_A2, BB_D2, BB_A3, BB_D3, BB_A4, BB_D4, BB_DF_FULL, BB_DF_FREE, BB_NSET };
constexpr int BB_ACC_DONE = 8 * (2 * BW_ST + 2 * BB_NSET), BB_TMEM_SLOT = BB_ACC_DONE + 8;

__device__ void stage_decoder_weights_bwd(const TriplaneParams& p, uint8_t* sm) {
    for (int i = threadIdx.x; i < HID * 64; i += blockDim.x) {
        const int j = i >> 6, e = i & 63, c = e & 31;
        const float w = p.W1[j * C + c] * (p.w1g * LOG2E);
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        const uint32_t o = sw128(j, e >> 3) + (e & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(sm + BO_W1A + o) = hi;
        // true W2^T: row j (unit), column k = e (output)
        const float wt = e < OUT ? p.W2[e * HID + j] * p.w2g : 0.f;
        const __nv_bfloat16 th = __float2bfloat16_rn(wt);
        *reinterpret_cast<__nv_bfloat16*>(sm + BO_W2TH + o) = th;
        *reinterpret_cast<__nv_bfloat16*>(sm + BO_W2TL + o) = __float2bfloat16_rn(wt - __bfloat162float(th));
    }
    for (int i = threadIdx.x; i < OUTP * 64; i += blockDim.x) {
        const int k = i >> 6, j = i & 63;
        const float w = k < OUT ? p.W2[k * HID + j] * (k == 0 ? p.w2g * LN2 : -p.w2g) : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        const uint32_t o = sw128(k, j >> 3) + (j & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(sm + BO_W2H + o) = hi;
    }
    for (int i = threadIdx.x; i < C * 64; i += blockDim.x) {            // true W1^T: row c (channel), column j (unit)
        const int c = i >> 6, j = i & 63;
        const float w = p.W1[j * C + c] * p.w1g;
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        const uint32_t o = sw128(c, j >> 3) + (j & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(sm + BO_W1TH + o) = hi;
        *reinterpret_cast<__nv_bfloat16*>(sm + BO_W1TL + o) = __float2bfloat16_rn(w - __bfloat162float(hi));
    }
    for (int i = threadIdx.x; i < 2048 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm + BO_ONES)[i] = 0x3f803f80u;    // bf16 1.0 pairs
    float* bias = reinterpret_cast<float*>(sm + BO_BIAS);
    for (int i = threadIdx.x; i < HID; i += blockDim.x) bias[i] = p.b1[i] * (p.b1g * LOG2E);
    for (int i = threadIdx.x; i < OUTP; i += blockDim.x) bias[HID + i] = i < OUT ? p.b2[i] * (i == 0 ? p.b2g : -p.b2g * LOG2E) : 0.f;
}

__device__ __forceinline__ uint4 pack8(const float* v) {
    return make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
}

// WGRAD: accumulate the decoder weight / bias gradients (PTI); false = frozen decoder (w-projection).
// COORDS: also produce the gradient w.r.t. the sample coordinates (d_coords) and / or its per-ray sums (d_ray_o, d_ray_d): the
// scatter warps then fetch the 12 texel lines of every point once more (the bilinear derivative needs the texel values).
template <bool WGRAD, bool COORDS>
__global__ void __launch_bounds__(BW_THREADS, 1) triplane_bwd_tc_kernel(TriplaneParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const uint32_t sm_u = smem_u32(sm);
    const uint32_t B = sm_u + BO_BAR;
    const int n = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (threadIdx.x == 0) {
        for (int i = 0; i < BW_ST; ++i) { mbar_init(bb_a1_full(B, i), 1); mbar_init(bb_a1_free(B, i), 4); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(bb_set(B, BB_D1, s), 1); mbar_init(bb_set(B, BB_A2, s), 4); mbar_init(bb_set(B, BB_D2, s), 1);
            mbar_init(bb_set(B, BB_A3, s), 4); mbar_init(bb_set(B, BB_D3, s), 1); mbar_init(bb_set(B, BB_A4, s), 4);
            mbar_init(bb_set(B, BB_D4, s), 1);
        }
        mbar_init(B + BB_ACC_DONE, 1);
        fence_barrier_init();
    }
    if (warp == BW_MMA_WARP) tmem_alloc(B + BB_TMEM_SLOT, BW_TM_COLS);
    stage_decoder_weights_bwd(p, sm);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + BO_BAR + BB_TMEM_SLOT);

    const long ntiles = (p.P + TILE - 1) / TILE;
    const long per = (ntiles + gridDim.x - 1) / gridDim.x;
    const long t0 = (long)blockIdx.x * per;
    const int nloc = (int)max(0L, min(ntiles, t0 + per) - t0);
    const float* pl = p.planes + (long)n * p.hp * p.wp * PC;
    float* dpl = p.d_planes ? p.d_planes + (long)n * p.hp * p.wp * PC : nullptr;
    const long row0 = (long)n * p.P;

    if (warp >= BW_SCATTER0) {
        // ------------------------------------------------------------------ scatter: every tile, rows 32 w .. 32 w + 31
        // The bilinear set-up (12 texel offsets + 12 weights per point) is recomputed here from the point's coordinates rather than
        // kept from the gather: that would tie a 12 KB buffer per tile to the whole length of the chain.
        // Eight warps: group sg = (warp - BW_SCATTER0) / 4 serves the tiles of consumer set sg.
        const int sw = warp - BW_SCATTER0, sg = sw >> 2, w = sw & 3, pt = lane >> 3, l8 = lane & 7;
        float* ss = reinterpret_cast<float*>(sm + BO_SS) + sw * 32 * (SP + SPX);
        float* sx = ss + 32 * SP;                                                // [32][SPX]: wx1, wy1 per plane, validity mask
        const uint8_t* stg = sm + BO_A2 + sg * 32768 + 16384;                    // d_f staging = A2lo[set], fp32 rows, 128-byte swizzle
        int it = 0;
        for (int lt = sg; lt < nloc; lt += 2, ++it) {
            float cx, cy, cz;
            const int pi = map_point(p, (unsigned)((t0 + lt) * TILE + w * 32 + lane));
            point_coords32(p, n, pi, cx, cy, cz);
            stage_setup(ss, lane, cx, cy, cz, p.hp, p.wp);
            float tdepth = 0.f;
            if (COORDS) {
                int m0, m1, m2;
                float f[6];
                plane_setup_frac(cx, cy, p.hp, p.wp, f[0], f[1], m0);
                plane_setup_frac(cx, cz, p.hp, p.wp, f[2], f[3], m1);
                plane_setup_frac(cz, cx, p.hp, p.wp, f[4], f[5], m2);
                *reinterpret_cast<float4*>(sx + lane * SPX) = make_float4(f[0], f[1], f[2], f[3]);
                *reinterpret_cast<float4*>(sx + lane * SPX + 4) = make_float4(f[4], f[5], __int_as_float(m0 | (m1 << 4) | (m2 << 8)), 0.f);
                if (pi >= 0 && !p.coords) tdepth = p.depths[row0 + pi];
            }
            named_sync(NB_DF_FULL + sg, 256);                                    // the consumer set has staged this tile's d_f
            float4 g[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] = *reinterpret_cast<const float4*>(stg + sw128(w * 32 + i * 4 + pt, l8));
            named_arrive(NB_DF_FREE + sg, 256);                                  // the staging rows are in registers now
            if (dpl) {
                float* pc = dpl + l8 * 4;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float* s = ss + (i * 4 + pt) * SP;
#pragma unroll
                    for (int k = 0; k < 12; k += 4) {
                        const int4 o = *reinterpret_cast<const int4*>(s + k);
                        const float4 wv = *reinterpret_cast<const float4*>(s + 12 + k);
                        red_add4(pc + o.x, wv.x, g[i]); red_add4(pc + o.y, wv.y, g[i]); red_add4(pc + o.z, wv.z, g[i]); red_add4(pc + o.w, wv.w, g[i]);
                    }
                }
            }
            if (COORDS) {
                // d point = sum over channels of d_f * d(bilinear)/d(coordinate), all three planes (ATen grid_sampler_2d_backward
                // semantics: out-of-range texels count as zeros).  Lane (pt, l8) covers channels 4 l8 .. 4 l8 + 3 of point 4 i + pt.
                const float* pc = pl + l8 * 4;
                const float hx = 0.5f * p.wp * (1.f / 3.f) * p.coord_scale, hy = 0.5f * p.hp * (1.f / 3.f) * p.coord_scale;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int q = i * 4 + pt;
                    const float* s = ss + q * SP;
                    const float4 fa = *reinterpret_cast<const float4*>(sx + q * SPX), fb = *reinterpret_cast<const float4*>(sx + q * SPX + 4);
                    const int mask = __float_as_int(fb.z);
                    const float wx1[3] = {fa.x, fa.z, fb.x}, wy1[3] = {fa.y, fa.w, fb.y};
                    float dix[3], diy[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int4 o = *reinterpret_cast<const int4*>(s + 4 * k);
                        const int mk = mask >> (4 * k);
                        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        const float4 t00 = (mk & 1) ? __ldg(reinterpret_cast<const float4*>(pc + o.x)) : z4;
                        const float4 t01 = (mk & 2) ? __ldg(reinterpret_cast<const float4*>(pc + o.y)) : z4;
                        const float4 t10 = (mk & 4) ? __ldg(reinterpret_cast<const float4*>(pc + o.z)) : z4;
                        const float4 t11 = (mk & 8) ? __ldg(reinterpret_cast<const float4*>(pc + o.w)) : z4;
                        auto dot = [&](const float4& a) { return g[i].x * a.x + g[i].y * a.y + g[i].z * a.z + g[i].w * a.w; };
                        const float s00 = dot(t00), s01 = dot(t01), s10 = dot(t10), s11 = dot(t11);
                        dix[k] = (1.f - wy1[k]) * (s01 - s00) + wy1[k] * (s11 - s10);
                        diy[k] = (1.f - wx1[k]) * (s10 - s00) + wx1[k] * (s11 - s01);
                    }
                    // planes: 0 <- (x, y), 1 <- (x, z), 2 <- (z, x)
                    float dx = (dix[0] + dix[1]) * hx + diy[2] * hy, dy = diy[0] * hy, dz = diy[1] * hy + dix[2] * hx;
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) {
                        dx += __shfl_xor_sync(0xffffffffu, dx, o); dy += __shfl_xor_sync(0xffffffffu, dy, o); dz += __shfl_xor_sync(0xffffffffu, dz, o);
                    }
                    const int pq = __shfl_sync(0xffffffffu, pi, q);
                    const float tq = __shfl_sync(0xffffffffu, tdepth, q);
                    if (l8 == 0 && pq >= 0) {
                        if (p.d_coords) {
                            float* dc = p.d_coords + (row0 + pq) * 3;
                            dc[0] = dx; dc[1] = dy; dc[2] = dz;
                        }
                        if (p.d_ray_o) {
                            const long r3 = ((long)n * p.M + (unsigned)pq / (unsigned)p.S) * 3;
                            atomicAdd(p.d_ray_o + r3, dx); atomicAdd(p.d_ray_o + r3 + 1, dy); atomicAdd(p.d_ray_o + r3 + 2, dz);
                            atomicAdd(p.d_ray_d + r3, tq * dx); atomicAdd(p.d_ray_d + r3 + 1, tq * dy); atomicAdd(p.d_ray_d + r3 + 2, tq * dz);
                        }
                    }
                }
            }
            __syncwarp();
        }
    } else if (warp == BW_LOAD_WARP) {
        // ------------------------------------------------------------------ loader: the forward kept every tile's features as the
        // finished layer-1 operand (split bf16, swizzled): one 16 KB bulk copy per tile, completion counted on the stage's mbarrier
        if (lane == 0) {
            const uint8_t* src = p.f_saved + ((long)n * ntiles + t0) * 16384;
            for (int lt = 0; lt < nloc; ++lt) {
                const int st = lt % BW_ST, use = lt / BW_ST;
                mbar_wait_sleep(bb_a1_free(B, st), (use & 1) ^ 1);    // layer 1 of the tile BW_ST back is done and its F has been copied out
                mbar_arrive_expect_tx(bb_a1_full(B, st), 16384);
                bulk_load(sm_u + BO_A1 + st * 16384, src + (long)lt * 16384, 16384, bb_a1_full(B, st));
            }
        }
    } else if (warp == BW_MMA_WARP) {
        // ------------------------------------------------------------------ tensor-core issue (one lane), four cursors
        {
            const uint32_t id1 = instr_desc_bf16(128, HID, 0, 0), id2 = instr_desc_bf16(128, OUTP, 0, 0);
            const uint32_t id3 = instr_desc_bf16(128, HID, 0, 0), id4 = instr_desc_bf16(128, C, 0, 0);
            const uint32_t idw2 = instr_desc_bf16(128, OUTP, 1, 1), idw1 = instr_desc_bf16(128, HID, 1, 1);
            const uint32_t dhi = smem_desc_hi(1024);
            int n1 = 0, n2 = 0, n3 = 0, n4 = 0;
            while (n4 < nloc) {
                bool did = false;
                if (n4 < n3 && mbar_test_warp(bb_set(B, BB_A4, n4 & 1), (n4 >> 1) & 1, lane)) {          // P4: d_f and dW1
                    tc_fence_after();
                    const int s = n4 & 1;
                    const uint32_t da = sm_u + BO_A2 + s * 32768, d4 = tmem + (uint32_t)(s * 192 + 64);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t a = kdesc(da + k * 32, dhi);
                            umma_bf16(d4, a, kdesc(sm_u + BO_W1TH + k * 32, dhi), id4, k != 0);
                            umma_bf16(d4, a, kdesc(sm_u + BO_W1TL + k * 32, dhi), id4, 1);
                        }
                        if (WGRAD) {
                            const uint32_t f = sm_u + BO_FC + s * 16384;
#pragma unroll
                            for (int k = 0; k < 8; ++k)        // rows 64..127 of the M = 128 operand: the 16 x 64 ones block, for every k-step
                                umma_bf16(tmem + TM_ACC1, desc_pack(smem_desc_lo(f + k * 2048, sm_u + BO_ONES - (f + k * 2048)), dhi),
                                          kdesc(da + k * 2048, dhi), idw1, (n4 | k) != 0);
                        }
                        umma_commit(bb_set(B, BB_D4, s));
                    }
                    __syncwarp();
                    ++n4; did = true;
                }
                if (n3 < n2 && mbar_test_warp(bb_set(B, BB_A3, n3 & 1), (n3 >> 1) & 1, lane)) {          // P3: dh and dW2
                    tc_fence_after();
                    const int s = n3 & 1;
                    const uint32_t hh = sm_u + BO_A2 + s * 32768, dO = hh + 16384, d3 = tmem + (uint32_t)(s * 192 + 128);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const uint64_t a = kdesc(dO + k * 32, dhi);
                            umma_bf16(d3, a, kdesc(sm_u + BO_W2TH + k * 32, dhi), id3, k != 0);
                            umma_bf16(d3, a, kdesc(sm_u + BO_W2TL + k * 32, dhi), id3, 1);
                        }
                        if (WGRAD) {
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                umma_bf16(tmem + TM_ACC2, desc_pack(smem_desc_lo(hh + k * 2048, sm_u + BO_ONES - (hh + k * 2048)), dhi),
                                          kdesc(dO + k * 2048, dhi), idw2, (n3 | k) != 0);
                        }
                        umma_commit(bb_set(B, BB_D3, s));
                    }
                    __syncwarp();
                    ++n3; did = true;
                }
                if (n2 < n1 && mbar_test_warp(bb_set(B, BB_A2, n2 & 1), (n2 >> 1) & 1, lane)) {          // P2: layer 2 (recompute)
                    tc_fence_after();
                    const int s = n2 & 1;
                    const uint32_t ah = sm_u + BO_A2 + s * 32768, al = ah + 16384, d2 = tmem + (uint32_t)(s * 192 + 64);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t dah = kdesc(ah + k * 32, dhi), dbh = kdesc(sm_u + BO_W2H + k * 32, dhi);
                            umma_bf16(d2, dah, dbh, id2, k != 0);
                            umma_bf16(d2, kdesc(al + k * 32, dhi), dbh, id2, 1);
                        }
                        umma_commit(bb_set(B, BB_D2, s));
                    }
                    __syncwarp();
                    ++n2; did = true;
                }
                if (n1 < nloc && n1 - n4 < 2 && mbar_test_warp(bb_a1_full(B, n1 % BW_ST), (n1 / BW_ST) & 1, lane)) {   // P1: layer 1 (recompute); D1[s] is free after S3 of tile n1 - 2
                    tc_fence_after();
                    const int s = n1 & 1;
                    const uint32_t a = sm_u + BO_A1 + (n1 % BW_ST) * 16384, d1 = tmem + (uint32_t)(s * 192);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(d1, kdesc(a + k * 32, dhi), kdesc(sm_u + BO_W1A + k * 32, dhi), id1, k != 0);
                        umma_commit(bb_set(B, BB_D1, s));
                    }
                    __syncwarp();
                    ++n1; did = true;
                }
                if (!did) __nanosleep(32);
            }
            if (elect_one()) umma_commit(B + BB_ACC_DONE);
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ consumers: set = tile parity, thread = point = TMEM lane
        const int set = warp >> 2, q = warp & 3, row = q * 32 + lane;
        const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(set * 192);
        uint8_t* a2h = sm + BO_A2 + set * 32768;
        uint8_t* a2l = a2h + 16384;
        const float* b1s = reinterpret_cast<const float*>(sm + BO_BIAS);
        const float* b2s = b1s + HID;
        int it = 0;
        for (int lt = set; lt < nloc; lt += 2, ++it) {
            const unsigned pp0 = (unsigned)((t0 + lt) * TILE + q * 32);
            const int pi = map_point(p, pp0 + lane);
            // ---- S1: h' = lg2(1 + 2^y)
            mbar_wait(bb_set(B, BB_D1, set), it & 1);
            tc_fence_after();
            if (it > 0) named_sync(NB_DF_FREE + set, 256);                         // the scatter warps have taken the previous d_f out of A2lo
            {   // keep this tile's F (hi half: chunks 0..3 of the row) for the dW1 contraction at the end of the chain; the stage goes back to the gather
                const uint8_t* a1 = sm + BO_A1 + (lt % BW_ST) * 16384;
                uint8_t* fc = sm + BO_FC + set * 16384;
#pragma unroll
                for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(fc + sw128(row, c)) = *reinterpret_cast<const uint4*>(a1 + sw128(row, c));
                __syncwarp();
                if (lane == 0) mbar_arrive(bb_a1_free(B, lt % BW_ST));
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float v[32];
                tmem_ld32(tq + half * 32, v);
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    float h[8];
#pragma unroll
                    for (int e = 0; e < 8; e += 4) {
                        const float4 bv = *reinterpret_cast<const float4*>(b1s + half * 32 + c8 * 8 + e);
                        h[e] = softplus2(v[c8 * 8 + e] + bv.x); h[e + 1] = softplus2(v[c8 * 8 + e + 1] + bv.y);
                        h[e + 2] = softplus2(v[c8 * 8 + e + 2] + bv.z); h[e + 3] = softplus2(v[c8 * 8 + e + 3] + bv.w);
                    }
                    uint4 hi, lo;
                    split8(h, hi, lo);
                    const uint32_t o = sw128(row, half * 4 + c8);
                    *reinterpret_cast<uint4*>(a2h + o) = hi;
                    *reinterpret_cast<uint4*>(a2l + o) = lo;
                }
            }
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bb_set(B, BB_A2, set));
            // ---- S2: dO (A2lo is free once layer 2 has read h' lo).  The incoming gradients of this warp's 32 rows are loaded
            // coalesced (8 lanes per 128-byte row) while layer 2 runs, and exchanged through the staging rows.
            float4 dr[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int pr = __shfl_sync(0xffffffffu, pi, i * 4 + (lane >> 3));
                dr[i] = pr >= 0 ? __ldg(reinterpret_cast<const float4*>(p.d_rgb + (row0 + pr) * C + (lane & 7) * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const float dsig = pi >= 0 ? __ldg(p.d_sigma + row0 + pi) : 0.f;
            mbar_wait(bb_set(B, BB_D2, set), it & 1);
            tc_fence_after();
#pragma unroll
            for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(a2l + sw128(q * 32 + i * 4 + (lane >> 3), lane & 7)) = dr[i];
            __syncwarp();
            {
                float z[32], z32[8], dO[48];
                tmem_ld32(tq + 64, z);
                tmem_ld8(tq + 96, z32);
                dO[0] = dsig;
#pragma unroll
                for (int c4 = 0; c4 < 32; c4 += 4) {
                    const float4 g4 = *reinterpret_cast<const float4*>(a2l + sw128(row, c4 >> 2));      // own row: d_rgb[c4 .. c4+3]
                    const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int col = c4 + e + 1;
                        const float s = sigmoid2((col < 32 ? z[col] : z32[0]) + b2s[col]);
                        dO[col] = gg[e] * 1.002f * s * (1.f - s);
                    }
                }
#pragma unroll
                for (int k = 33; k < 48; ++k) dO[k] = 0.f;
                __syncwarp();                                           // every lane has read its staged d_rgb row
#pragma unroll
                for (int c = 0; c < 6; ++c) *reinterpret_cast<uint4*>(a2l + sw128(row, c)) = pack8(dO + 8 * c);
            }
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bb_set(B, BB_A3, set));
            // ---- S3: d_a = dh * softplus'(u) = dh * 2^y / (1 + 2^y), y re-read from D1
            mbar_wait(bb_set(B, BB_D3, set), it & 1);
            tc_fence_after();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float y[32], dh[32];
                tmem_ld32(tq + half * 32, y);
                tmem_ld32(tq + 128 + half * 32, dh);
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    float da[8];
#pragma unroll
                    for (int e = 0; e < 8; e += 4) {
                        const float4 bv = *reinterpret_cast<const float4*>(b1s + half * 32 + c8 * 8 + e);
                        const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float e2 = exp2f(fminf(y[c8 * 8 + e + u] + bb[u], 126.f));
                            da[e + u] = dh[c8 * 8 + e + u] * __fdividef(e2, 1.f + e2);
                        }
                    }
                    *reinterpret_cast<uint4*>(a2h + sw128(row, half * 4 + c8)) = pack8(da);
                }
            }
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bb_set(B, BB_A4, set));
            // ---- S4: d_f -> staging for the scatter warps
            mbar_wait(bb_set(B, BB_D4, set), it & 1);
            tc_fence_after();
            {
                float df[32];
                tmem_ld32(tq + 64, df);
                tc_fence_before();
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<float4*>(a2l + sw128(row, c)) = make_float4(df[4 * c], df[4 * c + 1], df[4 * c + 2], df[4 * c + 3]);
            }
            named_arrive(NB_DF_FULL + set, 256);
        }
        // ---- drain of the weight / bias gradient accumulators (set 0's four warps cover the 128 TMEM lanes)
        if (WGRAD && set == 0 && nloc > 0) {
            mbar_wait(B + BB_ACC_DONE, 0);
            tc_fence_after();
            const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
            float v[32];
            if (q < 2) {                                    // lanes 0..63: dW2^T[j][k] (h' = h / ln2); lanes 0..31: dW1^T[c][j]
                const int j = q * 32 + lane;
                tmem_ld32(tl + TM_ACC2, v);
#pragma unroll
                for (int k = 0; k < 32; ++k) atomicAdd(p.dW2 + k * HID + j, v[k] * (LN2 * p.w2g));
                tmem_ld32(tl + TM_ACC2 + 32, v);            // columns 32..47 of ACC2, then 16 columns of ACC1
                atomicAdd(p.dW2 + 32 * HID + j, v[0] * (LN2 * p.w2g));
#pragma unroll
                for (int half = 0; half < 2; ++half) {      // ACC1 rows 0..31 = channels (F hi); rows 32..63 multiply the uncopied lo half: ignored
                    tmem_ld32(tl + TM_ACC1 + half * 32, v);
                    if (q == 0) {
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) atomicAdd(p.dW1 + (half * 32 + jj) * C + lane, v[jj] * p.w1g);
                    }
                }
            } else if (q == 2) {                            // lane 64 (first lane of this warp): the ones row = bias gradients
                tmem_ld32(tl + TM_ACC2, v);                 // tcgen05.ld is warp-collective: all lanes load, lane 0 publishes
                if (lane == 0)
                    for (int k = 0; k < 32; ++k) atomicAdd(p.db2 + k, v[k] * p.b2g);
                tmem_ld32(tl + TM_ACC2 + 32, v);
                if (lane == 0) atomicAdd(p.db2 + 32, v[0] * p.b2g);
                for (int half = 0; half < 2; ++half) {
                    tmem_ld32(tl + TM_ACC1 + half * 32, v);
                    if (lane == 0)
                        for (int jj = 0; jj < 32; ++jj) atomicAdd(p.db1 + half * 32 + jj, v[jj] * p.b1g);
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == BW_MMA_WARP) tmem_dealloc(tmem, BW_TM_COLS);
}
}  // namespace

extern int g_b200_mlp_passes;

// Launch helper used by b200_triplane_mlp_fwd (triplane.cu) when the tcgen05 implementation is selected.
int triplane_fwd_tc_launch(tri::TriplaneParams& p, bool sigma_only, cudaStream_t st) {
    B200_FUNC_ATTR_ONCE(triplane_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FW_SMEM);
    B200_FUNC_ATTR_ONCE(triplane_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FW_SMEM);
    const long ntiles = (p.P + TILE - 1) / TILE;
    const int sms = b200_sm_count();
    dim3 grid((unsigned)(ntiles < sms ? ntiles : sms), p.n);              // persistent: one CTA per SM, contiguous tile ranges
    if (sigma_only) triplane_fwd_tc_kernel<true><<<grid, FW_THREADS, FW_SMEM, st>>>(p);
    else triplane_fwd_tc_kernel<false><<<grid, FW_THREADS, FW_SMEM, st>>>(p);
    B200_CHECK_LAUNCH();
    return 0;
}

int triplane_bwd_tc_launch(tri::TriplaneParams& p, cudaStream_t st) {
    B200_FUNC_ATTR_ONCE((triplane_bwd_tc_kernel<true, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM);
    B200_FUNC_ATTR_ONCE((triplane_bwd_tc_kernel<false, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM);
    B200_FUNC_ATTR_ONCE((triplane_bwd_tc_kernel<true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM);
    B200_FUNC_ATTR_ONCE((triplane_bwd_tc_kernel<false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM);
    const long ntiles = (p.P + TILE - 1) / TILE;
    const int sms = b200_sm_count();
    dim3 grid((unsigned)(ntiles < sms ? ntiles : sms), p.n);
    const bool coords = p.d_coords || p.d_ray_o;
    if (p.dW1 && coords) triplane_bwd_tc_kernel<true, true><<<grid, BW_THREADS, BW_SMEM, st>>>(p);
    else if (p.dW1) triplane_bwd_tc_kernel<true, false><<<grid, BW_THREADS, BW_SMEM, st>>>(p);
    else if (coords) triplane_bwd_tc_kernel<false, true><<<grid, BW_THREADS, BW_SMEM, st>>>(p);
    else triplane_bwd_tc_kernel<false, false><<<grid, BW_THREADS, BW_SMEM, st>>>(p);
    B200_CHECK_LAUNCH();
    return 0;
}
