// Adam over a whole parameter list in one launch (the optimiser of the inversion loops: base_coach.py:96-99
// torch.optim.Adam(G.parameters(), lr), w_projector.py:134-140).  Update rule of torch.optim.Adam (amsgrad off, maximize off):
//   g' = g + wd * p;  m = m + (g' - m)(1 - b1);  v = b2 v + (1 - b2) g'^2
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps),   t = *step + 1
// The step counter lives on the device (a captured CUDA graph replays the same launch every step): every block reads it, the last
// block to retire increments it and re-arms the ticket.  HBM-bound: 16 B read + 12 B written per element.
#include "common.cuh"

namespace {

constexpr int ADAM_MAX_TENSORS = 320;        // per launch; 320 * 40 B + 321 * 4 B = 14 KB of kernel parameters (limit 32 764 B)
constexpr int ADAM_CHUNK = 8192;             // elements per block
constexpr int ADAM_THREADS = 256;

struct AdamTensor { float* p; const float* g; float* m; float* v; long n; };

struct AdamBatch {
    AdamTensor t[ADAM_MAX_TENSORS];
    int first_block[ADAM_MAX_TENSORS + 1];   // prefix sums of blocks per tensor
    int count;
    int bump;                                // 1: this launch advances the step counter when its last block retires
    const float* lr_dev; float lr;
    float beta1, beta2, eps, weight_decay;
    float* step; unsigned* ticket;
};

struct AdamCoef { float b1, b2, omb1, omb2, step_size, inv_sqrt_bc2, eps, wd; };

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const AdamCoef& c) {
    g = fmaf(c.wd, p, g);
    m = fmaf(g - m, c.omb1, m);
    v = fmaf(c.omb2 * g, g, c.b2 * v);
    p -= c.step_size * (m / fmaf(sqrtf(v), c.inv_sqrt_bc2, c.eps));
}

__global__ void __launch_bounds__(ADAM_THREADS) adam_multi_kernel(const __grid_constant__ AdamBatch b) {
    // block -> tensor by binary search in the prefix table
    int lo = 0, hi = b.count;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (b.first_block[mid] <= (int)blockIdx.x) lo = mid; else hi = mid; }
    const AdamTensor t = b.t[lo];
    const long e0 = (long)(blockIdx.x - b.first_block[lo]) * ADAM_CHUNK;
    const long e1 = e0 + ADAM_CHUNK < t.n ? e0 + ADAM_CHUNK : t.n;

    const float tstep = *b.step + 1.f;
    const float lr = b.lr_dev ? *b.lr_dev : b.lr;
    AdamCoef c;
    c.b1 = b.beta1; c.b2 = b.beta2; c.omb1 = 1.f - b.beta1; c.omb2 = 1.f - b.beta2; c.eps = b.eps; c.wd = b.weight_decay;
    c.step_size = lr / (1.f - powf(b.beta1, tstep));
    c.inv_sqrt_bc2 = rsqrtf(1.f - powf(b.beta2, tstep));

    const bool vec = ((((uintptr_t)t.p | (uintptr_t)t.g | (uintptr_t)t.m | (uintptr_t)t.v) & 15) == 0);
    if (vec) {
        const long v0 = e0 >> 2, v1 = e1 >> 2;           // e0 is a multiple of ADAM_CHUNK
        float4* p4 = reinterpret_cast<float4*>(t.p); const float4* g4 = reinterpret_cast<const float4*>(t.g);
        float4* m4 = reinterpret_cast<float4*>(t.m); float4* q4 = reinterpret_cast<float4*>(t.v);
        for (long i = v0 + threadIdx.x; i < v1; i += ADAM_THREADS) {
            float4 p = p4[i], m = m4[i], v = q4[i];
            const float4 g = __ldg(g4 + i);
            adam_elem(p.x, g.x, m.x, v.x, c); adam_elem(p.y, g.y, m.y, v.y, c);
            adam_elem(p.z, g.z, m.z, v.z, c); adam_elem(p.w, g.w, m.w, v.w, c);
            p4[i] = p; m4[i] = m; q4[i] = v;
        }
        for (long i = (v1 << 2) + threadIdx.x; i < e1; i += ADAM_THREADS) adam_elem(t.p[i], t.g[i], t.m[i], t.v[i], c);
    } else {
        for (long i = e0 + threadIdx.x; i < e1; i += ADAM_THREADS) adam_elem(t.p[i], t.g[i], t.m[i], t.v[i], c);
    }

    if (b.bump) {                                        // last block out advances the counter (every block has read it by then)
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(b.ticket, 1u) == gridDim.x - 1) { *b.step = tstep; *b.ticket = 0u; __threadfence(); }
        }
    }
}

}  // namespace

// tensors: HOST array of `count` records {float* p; const float* g; float* m; float* v; long n} (device pointers, fp32, n elements
// each; records with n == 0 are skipped).  lr_dev: device scalar learning rate, or NULL to use `lr`.  step: device float, number of
// updates applied so far (0 at the start), advanced by one per call; ticket: device uint32, zero-initialised, owned by the optimiser.
B200_API int b200_adam_step(const void* tensors, int count, const float* lr_dev, float lr, float beta1, float beta2, float eps,
                            float weight_decay, float* step, unsigned* ticket, void* stream) {
    B200_REQUIRE(count >= 0 && (count == 0 || tensors), "adam_step: null tensor table");
    B200_REQUIRE(step && ticket, "adam_step: null step counter / ticket");
    const AdamTensor* src = static_cast<const AdamTensor*>(tensors);
    cudaStream_t st = (cudaStream_t)stream;
    AdamBatch b;
    b.lr_dev = lr_dev; b.lr = lr; b.beta1 = beta1; b.beta2 = beta2; b.eps = eps; b.weight_decay = weight_decay;
    b.step = step; b.ticket = ticket;
    int i = 0;
    // skip trailing empties so that the LAST launch is the one that advances the counter
    while (count > 0 && src[count - 1].n == 0) --count;
    if (count == 0) return 0;                          // nothing to update (torch skips parameters without a gradient as well)
    while (i < count) {
        b.count = 0; b.first_block[0] = 0;
        while (i < count && b.count < ADAM_MAX_TENSORS) {
            if (src[i].n > 0) {
                B200_REQUIRE(src[i].p && src[i].g && src[i].m && src[i].v, "adam_step: null tensor pointer");
                b.t[b.count] = src[i];
                b.first_block[b.count + 1] = b.first_block[b.count] + cdiv(src[i].n, ADAM_CHUNK);
                ++b.count;
            }
            ++i;
        }
        b.bump = i >= count ? 1 : 0;
        adam_multi_kernel<<<b.first_block[b.count], ADAM_THREADS, 0, st>>>(b);
        B200_CHECK_LAUNCH();
    }
    return 0;
}
