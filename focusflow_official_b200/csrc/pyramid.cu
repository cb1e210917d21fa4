// Average-pool correlation pyramid (sm_100a), replaces 3x F.avg_pool2d
// (FF_RAFT_Core/corr.py:24-27).
//
// HBM-bound streaming kernel: level 0 is read ONCE and levels 1..3 are produced
// from shared memory in the same pass.  Algorithmic bytes per query map:
// 4 * (n0 + n1 + n2 + n3), n_i = (h>>i)*(w>>i)  (reference: 3 passes, 4*(n0+2n1+2n2+n3)).
//
// A block owns a group of G consecutive query maps (contiguous in memory, so loads and
// stores are fully coalesced / 16-byte vectorised when the group is 16-byte aligned).
// Window sum order is ATen's: ((a+b)+c)+d then one multiply by 0.25 (== /4 exactly), so
// the result is bit-identical to F.avg_pool2d given the same level 0.
#include <cuda_fp16.h>

#include "common.cuh"

namespace ffcorr {
namespace {

constexpr int kPyrThreads = 256;
constexpr int kFusedLevels = 4;

struct PyrParams {
    const float* l0;
    float* out[kFusedLevels];   // out[1..3] used
    int h[kFusedLevels], w[kFusedLevels];
    int n[kFusedLevels];        // elements per map per level
    int num_levels;             // 2..4
    unsigned magic_w[kFusedLevels];  // ceil(2^32 / w[l])      -> row = umulhi(i, magic)
    unsigned magic_p[kFusedLevels];  // ceil(2^32 / ((w[l]+1)/2))
    int G;                      // maps per block
    int64_t Q;
};

__device__ __forceinline__ float pool4(const float* s, int w) {
    return (((s[0] + s[1]) + s[w]) + s[w + 1]) * 0.25f;
}

// round a float offset up to 16 bytes (level buffers in shared memory start 16-byte aligned)
__host__ __device__ __forceinline__ int align4(int x) { return (x + 3) & ~3; }

// One pooling level for a group of g maps held in shared memory.  Work items are flattened over
// (row, x) so all 256 threads stay busy; the row index comes from a multiply-high by a host-computed
// reciprocal (no integer division).  When the input width is even a thread fetches the two input rows of
// TWO adjacent outputs with two LDS.128, halving the instruction count per output.
__device__ __forceinline__ void pool_level(const PyrParams& p, int l, int g, const float* __restrict__ sin,
                                           float* __restrict__ sout, float* __restrict__ dst, bool keep, int tid) {
    const int wi = p.w[l - 1], ni = p.n[l - 1];
    const int ho = p.h[l], wo = p.w[l], no = p.n[l];
    // 16-byte loads need BOTH input rows 16-byte aligned: wi % 4 == 0 (row 2y+1 starts at (2y+1)*wi)
    const bool vec = ((wi & 3) == 0) && ((((uintptr_t)sin) & 15) == 0) && (g == 1 || (ni & 3) == 0);
    const bool vec2 = ((wi & 1) == 0) && ((((uintptr_t)sin) & 7) == 0) && (g == 1 || (ni & 1) == 0);
    if (vec) {
        const int P = (wo + 1) >> 1;
        const int items = ho * P;
        const bool st2 = ((wo & 1) == 0) && ((no & 1) == 0) && ((((uintptr_t)dst) & 7) == 0) && ((((uintptr_t)sout) & 7) == 0);
        for (int m = 0; m < g; ++m) {
            const float* s0 = sin + m * ni;
            float* so = sout + m * no;
            float* go = dst + (size_t)m * no;
            for (int i = tid; i < items; i += kPyrThreads) {
                const int y = (P == 1) ? i : (int)__umulhi((unsigned)i, p.magic_p[l]);
                const int xp = i - y * P;
                const float4 a = *reinterpret_cast<const float4*>(s0 + (2 * y) * wi + 4 * xp);
                const float4 b = *reinterpret_cast<const float4*>(s0 + (2 * y + 1) * wi + 4 * xp);
                const float v0 = (((a.x + a.y) + b.x) + b.y) * 0.25f;
                const float v1 = (((a.z + a.w) + b.z) + b.w) * 0.25f;
                const int o = y * wo + 2 * xp;
                if (st2) {
                    if (keep) *reinterpret_cast<float2*>(so + o) = make_float2(v0, v1);
                    __stcs(reinterpret_cast<float2*>(go + o), make_float2(v0, v1));
                } else {
                    const bool has1 = 2 * xp + 1 < wo;
                    if (keep) {
                        so[o] = v0;
                        if (has1) so[o + 1] = v1;
                    }
                    __stcs(go + o, v0);
                    if (has1) __stcs(go + o + 1, v1);
                }
            }
        }
    } else {
        const int items = ho * wo;
        for (int m = 0; m < g; ++m) {
            const float* s0 = sin + m * ni;
            float* so = sout + m * no;
            float* go = dst + (size_t)m * no;
            for (int i = tid; i < items; i += kPyrThreads) {
                const int y = (wo == 1) ? i : (int)__umulhi((unsigned)i, p.magic_w[l]);
                const int x = i - y * wo;
                float v;
                if (vec2) {  // even width: each input row of the 2x2 window is one aligned LDS.64
                    const float2 a = *reinterpret_cast<const float2*>(s0 + (2 * y) * wi + 2 * x);
                    const float2 b = *reinterpret_cast<const float2*>(s0 + (2 * y + 1) * wi + 2 * x);
                    v = (((a.x + a.y) + b.x) + b.y) * 0.25f;
                } else {
                    v = pool4(s0 + (2 * y) * wi + 2 * x, wi);
                }
                if (keep) so[i] = v;
                __stcs(go + i, v);
            }
        }
    }
}

template <bool VEC4>
__global__ void __launch_bounds__(kPyrThreads) pyramid_fused_kernel(const PyrParams p) {
    extern __shared__ __align__(128) float sm[];
    const int64_t q0 = (int64_t)blockIdx.x * p.G;
    const int g = (int)min((int64_t)p.G, p.Q - q0);
    const int tid = threadIdx.x;

    // ---- stage level 0 of the group (contiguous) ----
    const int e0 = g * p.n[0];
    const float* __restrict__ src = p.l0 + q0 * p.n[0];
    if (VEC4) {
        // all of a thread's 16-byte loads are issued before the first shared-memory store, so a
        // block has its whole group in flight at once (8 x 16 B x 256 threads = 32 KB per round)
        const float4* __restrict__ s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(sm);
        const int n4 = e0 >> 2;
        constexpr int U = 8;
        for (int base = 0; base < n4; base += U * kPyrThreads) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = base + u * kPyrThreads + tid;
                if (i < n4) v[u] = __ldcs(s4 + i);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = base + u * kPyrThreads + tid;
                if (i < n4) d4[i] = v[u];
            }
        }
        for (int i = (n4 << 2) + tid; i < e0; i += kPyrThreads) sm[i] = __ldcs(src + i);
    } else {
        constexpr int U = 8;
        for (int base = 0; base < e0; base += U * kPyrThreads) {
            float v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = base + u * kPyrThreads + tid;
                if (i < e0) v[u] = __ldcs(src + i);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = base + u * kPyrThreads + tid;
                if (i < e0) sm[i] = v[u];
            }
        }
    }
    __syncthreads();

    // ---- levels 1.. from shared memory ----
    const float* sin = sm;
    int off = align4(p.G * p.n[0]);
#pragma unroll 1
    for (int l = 1; l < p.num_levels; ++l) {
        float* sout = sm + off;
        pool_level(p, l, g, sin, sout, p.out[l] + q0 * p.n[l], l < p.num_levels - 1, tid);
        __syncthreads();
        sin = sout;
        off = align4(off + p.G * p.n[l]);
    }
}

// ------------------------------------------------------------------------------------------
// Persistent, bulk-copy fed variant (group bytes % 16 == 0, 16-byte aligned base): level 0 of each
// group arrives through cp.async.bulk into a 3-stage mbarrier ring, so every block always has
// (stages-1) groups in flight while it pools the current one; no thread spends instructions or
// registers on the 1.7 GB level-0 read.
// ------------------------------------------------------------------------------------------
constexpr int kPyrStages = 3;

__device__ __forceinline__ void py_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void py_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void py_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ void py_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(kPyrThreads) pyramid_bulk_kernel(const PyrParams p, const int64_t num_groups) {
    extern __shared__ __align__(128) float sm[];
    const int tid = threadIdx.x;
    const int stage_floats = p.G * p.n[0];                       // multiple of 4 (host guarantees)
    const uint32_t stage_bytes = (uint32_t)stage_floats * 4u;
    float* scratch = sm + (size_t)kPyrStages * stage_floats;     // levels 1..L-2 of the current group
    const uint32_t sm_u32 = (uint32_t)__cvta_generic_to_shared(sm);
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(scratch + align4(align4(p.G * p.n[1]) + p.G * p.n[2]) + 4) & ~7u;

    if (tid == 0) {
        for (int s = 0; s < kPyrStages; ++s) py_mbar_init(bar0 + 8u * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int64_t grp, int k) {  // thread 0 only; groups are always full here
        const uint32_t bar = bar0 + 8u * (k % kPyrStages);
        py_mbar_expect_tx(bar, stage_bytes);
        py_bulk_load(sm_u32 + (uint32_t)(k % kPyrStages) * stage_bytes, p.l0 + grp * stage_floats, stage_bytes, bar);
    };
    if (tid == 0) {
        int k = 0;
        for (int64_t grp = blockIdx.x; grp < num_groups && k < kPyrStages; grp += gridDim.x, ++k) issue(grp, k);
    }

    int k = 0;
    for (int64_t grp = blockIdx.x; grp < num_groups; grp += gridDim.x, ++k) {
        py_mbar_wait(bar0 + 8u * (k % kPyrStages), (uint32_t)((k / kPyrStages) & 1));
        const int64_t q0 = grp * p.G;
        const float* sin = sm + (size_t)(k % kPyrStages) * stage_floats;
        int off = 0;
#pragma unroll 1
        for (int l = 1; l < p.num_levels; ++l) {
            float* sout = scratch + off;
            pool_level(p, l, p.G, sin, sout, p.out[l] + q0 * p.n[l], l < p.num_levels - 1, tid);
            __syncthreads();
            sin = sout;
            off = align4(off + p.G * p.n[l]);
        }
        // the stage (and the scratch) are free again: refill the stage with the group kPyrStages ahead
        if (tid == 0) {
            const int64_t nxt = grp + (int64_t)kPyrStages * gridDim.x;
            if (nxt < num_groups) issue(nxt, k + kPyrStages);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Tiled ("T4") layout: every level is stored per query map as [th][tw][4][4] floats, th = ceil(h/4),
// tw = ceil(w/4), padding elements are exact zeros.  A 4x4 tile is 64 contiguous bytes -- the granularity
// at which the memory system serves the lookup's gathers -- so a (2r+2)^2 window touches ~10.6 requests
// instead of ~17 with row-major maps.  Pooling is tile-local: one output tile row (4 outputs) reads two
// rows of two adjacent input tiles = four LDS.128, and is stored with one STS.128 + one STG.128.
// ------------------------------------------------------------------------------------------
struct TiledPyrParams {
    const float* l0;
    float* out[kFusedLevels];
    int h[kFusedLevels], w[kFusedLevels];       // true level sizes
    int th[kFusedLevels], tw[kFusedLevels];     // tiles per column / row
    int np[kFusedLevels];                       // floats per map: th*tw*16
    unsigned magic_tw[kFusedLevels];            // ceil(2^32 / tw)
    int num_levels;
};

__device__ __forceinline__ void pool_level_tiled(const TiledPyrParams& p, int l, const float* __restrict__ sin,
                                                 float* __restrict__ sout, float* __restrict__ dst, bool keep, int tid) {
    const int twi = p.tw[l - 1];
    const int tho = p.th[l], two = p.tw[l], ho = p.h[l], wo = p.w[l];
    const int items = tho * two * 4;            // one item = one row of one output tile
    for (int i = tid; i < items; i += kPyrThreads) {
        const int t = i >> 2, iy = i & 3;
        const int ty = (two == 1) ? t : (int)__umulhi((unsigned)t, p.magic_tw[l]);
        const int tx = t - ty * two;
        const int y = 4 * ty + iy;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (y < ho && 4 * tx < wo) {
            // input rows 2y, 2y+1 live in input tile row (2ty + (iy >> 1)), in-tile rows (2*iy)&3 and +1.
            // Bank groups: a 16-byte read of row k of input tile T sits in group (4T + k) mod 8, and an
            // item's four reads cover groups {g, g+1, g+4, g+5}.  The 8 items of a quarter warp are 2 output
            // tiles x 4 rows; odd tiles read the odd row first and, when the input has an even number of tiles
            // per row (tile parity does not alternate with iy), the lower half starts with the right-hand
            // tile -- then every step of the quarter warp hits 8 distinct groups.
            const float* ta = sin + ((2 * ty + (iy >> 1)) * twi + 2 * tx) * 16 + ((2 * iy) & 3) * 4;
            const bool have_b = 4 * tx + 2 < wo;
            const bool swr = (t & 1) != 0;
            const bool swt = have_b && ((twi & 1) == 0) && ((iy >> 1) != 0);
            const float* first = ta + (swt ? 16 : 0);
            const float* second = ta + (swt ? 0 : 16);
            const float4 x0 = *reinterpret_cast<const float4*>(first + (swr ? 4 : 0));
            const float4 x1 = *reinterpret_cast<const float4*>(first + (swr ? 0 : 4));
            // upper / lower input rows of the first-loaded tile
            const float4 u0 = swr ? x1 : x0, l0 = swr ? x0 : x1;
            const float p0 = (((u0.x + u0.y) + l0.x) + l0.y) * 0.25f;
            const float p1 = (((u0.z + u0.w) + l0.z) + l0.w) * 0.25f;
            float q0 = 0.f, q1 = 0.f;
            if (have_b) {
                const float4 x2 = *reinterpret_cast<const float4*>(second + (swr ? 4 : 0));
                const float4 x3 = *reinterpret_cast<const float4*>(second + (swr ? 0 : 4));
                const float4 u1 = swr ? x3 : x2, l1 = swr ? x2 : x3;
                q0 = (((u1.x + u1.y) + l1.x) + l1.y) * 0.25f;
                q1 = (((u1.z + u1.w) + l1.z) + l1.w) * 0.25f;
            }
            // (p0, p1) belong to the left tile unless the tiles were swapped
            o.x = swt ? q0 : p0;
            o.y = swt ? q1 : p1;
            o.z = swt ? p0 : q0;
            o.w = swt ? p1 : q1;
            if (4 * tx + 1 >= wo) o.y = 0.f;
            if (4 * tx + 3 >= wo) o.w = 0.f;
        }
        if (keep) *reinterpret_cast<float4*>(sout + 4 * i) = o;
        __stcs(reinterpret_cast<float4*>(dst + 4 * i), o);
    }
}

__global__ void __launch_bounds__(kPyrThreads) pyramid_tiled_kernel(const TiledPyrParams p, const int64_t Q) {
    extern __shared__ __align__(128) float sm[];
    const int tid = threadIdx.x;
    const int stage_floats = p.np[0];
    const uint32_t stage_bytes = (uint32_t)stage_floats * 4u;
    float* scratch = sm + (size_t)kPyrStages * stage_floats;
    const uint32_t sm_u32 = (uint32_t)__cvta_generic_to_shared(sm);
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(scratch + p.np[1] + p.np[2] + 4) & ~7u;
    if (tid == 0) {
        for (int s = 0; s < kPyrStages; ++s) py_mbar_init(bar0 + 8u * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int64_t q, int k) {  // thread 0 only
        const uint32_t bar = bar0 + 8u * (k % kPyrStages);
        py_mbar_expect_tx(bar, stage_bytes);
        py_bulk_load(sm_u32 + (uint32_t)(k % kPyrStages) * stage_bytes, p.l0 + q * stage_floats, stage_bytes, bar);
    };
    if (tid == 0) {
        int k = 0;
        for (int64_t q = blockIdx.x; q < Q && k < kPyrStages; q += gridDim.x, ++k) issue(q, k);
    }
    int k = 0;
    for (int64_t q = blockIdx.x; q < Q; q += gridDim.x, ++k) {
        py_mbar_wait(bar0 + 8u * (k % kPyrStages), (uint32_t)((k / kPyrStages) & 1));
        const float* sin = sm + (size_t)(k % kPyrStages) * stage_floats;
        int off = 0;
#pragma unroll 1
        for (int l = 1; l < p.num_levels; ++l) {
            float* sout = scratch + off;
            pool_level_tiled(p, l, sin, sout, p.out[l] + q * p.np[l], l < p.num_levels - 1, tid);
            __syncthreads();
            sin = sout;
            off += p.np[l];
        }
        if (tid == 0) {
            const int64_t nxt = q + (int64_t)kPyrStages * gridDim.x;
            if (nxt < Q) issue(nxt, k + kPyrStages);
        }
    }
}

// tiled -> reference row-major [Q, h, w] (CorrBlock.corr_pyramid for inspection / tests)
__global__ void __launch_bounds__(256) untile_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t Q,
                                                     int h, int w, int tw, int np) {
    const int64_t total = Q * h * w;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(idx % w);
        const int64_t t = idx / w;
        const int y = (int)(t % h);
        const int64_t q = t / h;
        dst[idx] = __ldg(src + q * np + ((y >> 2) * tw + (x >> 2)) * 16 + (y & 3) * 4 + (x & 3));
    }
}

// Grouped ("G32") layout <-> reference row-major [B*N, h, w].  Level storage: [B][NG][th*tw][32 queries][4][4] floats,
// NG = ceil(N / 32): the same tile of 32 consecutive queries is contiguous (ffcorr_build_grouped_f32).
__global__ void __launch_bounds__(256) ungroup_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int N,
                                                      int h, int w, int tw, int ntiles) {
    const int64_t total = (int64_t)B * N * h * w;
    const int NG = (N + 31) >> 5;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(idx % w);
        const int64_t t = idx / w;
        const int y = (int)(t % h);
        const int64_t bq = t / h;
        const int b = (int)(bq / N), q = (int)(bq - (int64_t)b * N);
        const int64_t tile = ((int64_t)b * NG + (q >> 5)) * ntiles + (y >> 2) * tw + (x >> 2);
        dst[idx] = __ldg(src + (tile * 32 + (q & 31)) * 16 + (y & 3) * 4 + (x & 3));
    }
}

__global__ void __launch_bounds__(256) group_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int N,
                                                    int h, int w, int tw, int ntiles) {
    const int NG = (N + 31) >> 5;
    const int64_t total = (int64_t)B * NG * ntiles * 512;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(idx & 15), ql = (int)((idx >> 4) & 31);
        const int64_t tile_g = idx >> 9;
        const int tile = (int)(tile_g % ntiles);
        const int64_t bg = tile_g / ntiles;
        const int b = (int)(bg / NG), q = (int)(bg - (int64_t)b * NG) * 32 + ql;
        const int ty = tile / tw, tx = tile - ty * tw;
        const int y = 4 * ty + (e >> 2), x = 4 * tx + (e & 3);
        dst[idx] = (q < N && y < h && x < w) ? __ldg(src + (((int64_t)b * N + q) * h + y) * w + x) : 0.0f;
    }
}

// fp16-stored tiles (ffcorr_build_tiled_f16) -> reference row-major fp32 [Q, h, w]
__global__ void __launch_bounds__(256) untile_half_kernel(const __half* __restrict__ src, float* __restrict__ dst, int64_t Q,
                                                          int h, int w, int tw, int np) {
    const int64_t total = Q * h * w;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(idx % w);
        const int64_t t = idx / w;
        const int y = (int)(t % h);
        const int64_t q = t / h;
        dst[idx] = __half2float(src[q * np + ((y >> 2) * tw + (x >> 2)) * 16 + (y & 3) * 4 + (x & 3)]);
    }
}

// reference row-major fp32 [Q, h, w] -> fp16 tiles (RN, zero padding), used by tests
__global__ void __launch_bounds__(256) tile_half_kernel(const float* __restrict__ src, __half* __restrict__ dst, int64_t Q,
                                                        int h, int w, int tw, int np) {
    const int64_t total = Q * np;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(idx % np);
        const int64_t q = idx / np;
        const int t = e >> 4, iy = (e >> 2) & 3, ix = e & 3;
        const int ty = t / tw, tx = t - ty * tw;
        const int y = 4 * ty + iy, x = 4 * tx + ix;
        dst[idx] = __float2half_rn((y < h && x < w) ? __ldg(src + (q * h + y) * w + x) : 0.0f);
    }
}

// reference row-major [Q, h, w] -> tiled (zero padding), used by tests to feed golden pyramids to the tiled lookup
__global__ void __launch_bounds__(256) tile_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t Q,
                                                   int h, int w, int tw, int np) {
    const int64_t total = Q * np;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(idx % np);
        const int64_t q = idx / np;
        const int t = e >> 4, iy = (e >> 2) & 3, ix = e & 3;
        const int ty = t / tw, tx = t - ty * tw;
        const int y = 4 * ty + iy, x = 4 * tx + ix;
        dst[idx] = (y < h && x < w) ? __ldg(src + (q * h + y) * (int64_t)w + x) : 0.0f;
    }
}

// generic one-level kernel straight from global memory (huge maps / > 4 levels)
__global__ void __launch_bounds__(256) pool_level_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                         int64_t Q, int hi, int wi, int ho, int wo) {
    const int64_t total = Q * ho * wo;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(idx % wo);
        const int64_t t = idx / wo;
        const int y = (int)(t % ho);
        const int64_t q = t / ho;
        const float* s = in + (q * hi + 2 * y) * (int64_t)wi + 2 * x;
        out[idx] = (((__ldg(s) + __ldg(s + 1)) + __ldg(s + wi)) + __ldg(s + wi + 1)) * 0.25f;
    }
}

// adjoint of one pooling level, accumulated in place into the finer level's gradient
__global__ void __launch_bounds__(256) pool_level_bwd_kernel(float* __restrict__ gfine, const float* __restrict__ gcoarse,
                                                             int64_t Q, int hi, int wi, int ho, int wo) {
    const int64_t total = Q * hi * wi;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(idx % wi);
        const int64_t t = idx / wi;
        const int y = (int)(t % hi);
        const int64_t q = t / hi;
        const int yo = y >> 1, xo = x >> 1;
        if (yo < ho && xo < wo) gfine[idx] += __ldg(gcoarse + (q * ho + yo) * (int64_t)wo + xo) * 0.25f;
    }
}

// Same adjoint, one thread per PAIR of coarse elements: two 16-byte read-modify-writes of the fine level (rows 2Y
// and 2Y+1), one 64-bit division per thread instead of three.  VEC needs wi % 4 == 0 and a 16-byte aligned base.
template <bool VEC>
__global__ void __launch_bounds__(256) pool_level_bwd_pair_kernel(float* __restrict__ gfine, const float* __restrict__ gcoarse,
                                                                  int64_t Q, int hi, int wi, int ho, int wo) {
    const int wo2 = (wo + 1) >> 1;
    const int items = ho * wo2;
    const int64_t total = Q * items;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = idx / items;
        const int r = (int)(idx - q * items);
        const int Y = r / wo2;
        const int X = (r - Y * wo2) * 2;
        const float* gc = gcoarse + (q * ho + Y) * (int64_t)wo + X;
        const bool two = X + 1 < wo;
        const float c0 = __ldg(gc) * 0.25f;
        const float c1 = two ? __ldg(gc + 1) * 0.25f : 0.0f;
        float* f0 = gfine + (q * hi + 2 * Y) * (int64_t)wi + 2 * X;
        float* f1 = f0 + wi;
        if (VEC && two) {
            float4 a = *reinterpret_cast<float4*>(f0), b = *reinterpret_cast<float4*>(f1);
            a.x += c0; a.y += c0; a.z += c1; a.w += c1;
            b.x += c0; b.y += c0; b.z += c1; b.w += c1;
            *reinterpret_cast<float4*>(f0) = a;
            *reinterpret_cast<float4*>(f1) = b;
        } else {
            f0[0] += c0; f0[1] += c0; f1[0] += c0; f1[1] += c0;
            if (two) { f0[2] += c1; f0[3] += c1; f1[2] += c1; f1[3] += c1; }
        }
    }
}

int grid_for(int64_t total, int threads) {
    const int64_t want = ceil_div64(total, threads);
    const int64_t cap = (int64_t)sm_count() * 16;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace
}  // namespace ffcorr

using namespace ffcorr;

extern "C" int ffcorr_pyramid_f32(float* const* lvl, int num_levels, int64_t Q, int h, int w, void* stream) {
    FFCORR_REQUIRE(lvl != nullptr, FFCORR_EINVAL, "pyramid: null level table");
    FFCORR_REQUIRE(Q >= 0, FFCORR_EINVAL, "pyramid: Q=%lld", (long long)Q);
    if (int rc = check_levels(num_levels, h, w, "pyramid")) return rc;
    if (Q == 0 || num_levels == 1) return FFCORR_OK;
    for (int i = 0; i < num_levels; ++i)
        FFCORR_REQUIRE(lvl[i] != nullptr, FFCORR_EINVAL, "pyramid: lvl[%d] is null", i);
    cudaStream_t s = (cudaStream_t)stream;

    // fused single-pass path: up to 4 levels, group of maps must fit in shared memory
    const int fused = num_levels < kFusedLevels ? num_levels : kFusedLevels;
    PyrParams p{};
    p.l0 = lvl[0];
    p.num_levels = fused;
    p.Q = Q;
    int64_t per_map = 0;
    for (int i = 0; i < fused; ++i) {
        p.h[i] = h >> i;
        p.w[i] = w >> i;
        p.n[i] = p.h[i] * p.w[i];
        p.out[i] = lvl[i];
        per_map += p.n[i] + 4;   // + alignment slack per level buffer
        p.magic_w[i] = (unsigned)((0x100000000ull + p.w[i] - 1) / (unsigned)p.w[i]);
        const unsigned pr = (unsigned)((p.w[i] + 1) / 2);
        p.magic_p[i] = (unsigned)((0x100000000ull + pr - 1) / pr);
    }
    // maps per block: 16-byte aligned groups, ~8K level-0 elements per block, <= 96 KB smem
    const int n0 = p.n[0];
    int galign = 1;
    while ((int64_t)galign * n0 % 4 != 0) galign *= 2;  // 1, 2 or 4
    int G = (8192 / n0) / galign * galign;
    if (G < galign) G = galign;
    const size_t smem_cap = 96 * 1024;
    while (G > galign && (size_t)G * per_map * sizeof(float) > smem_cap) G -= galign;
    size_t smem = (size_t)G * per_map * sizeof(float);
    bool vec4 = ((uintptr_t)lvl[0] % 16 == 0) && ((int64_t)G * n0 % 4 == 0);
    if (!vec4) {  // alignment not available: any G works with scalar loads
        G = 8192 / n0 > 0 ? 8192 / n0 : 1;
        while (G > 1 && (size_t)G * per_map * sizeof(float) > smem_cap) --G;
        smem = (size_t)G * per_map * sizeof(float);
    }
    int done_levels = 1;
    const size_t bulk_smem = ((size_t)kPyrStages * G * n0 + (size_t)G * (p.n[1] + p.n[2]) + 32) * sizeof(float) + 64;
    const int64_t full_groups = Q / G;
    if (vec4 && fused >= 2 && bulk_smem <= 110 * 1024 && full_groups >= 1 && (uint64_t)G * n0 * 4 < (1u << 20)) {
        // persistent bulk-copy pipeline over the full groups, 2 blocks per SM
        p.G = G;
        const int64_t want = 2ll * sm_count();
        const int grid = (int)(full_groups < want ? full_groups : want);
        FFCORR_CUDA(cudaFuncSetAttribute(pyramid_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bulk_smem));
        pyramid_bulk_kernel<<<grid, kPyrThreads, bulk_smem, s>>>(p, full_groups);
        if (int rc = check_launch("pyramid_bulk_kernel")) return rc;
        const int64_t rest = Q - full_groups * G;
        if (rest > 0) {  // tail maps (fewer than G): the block-per-group kernel, scalar loads
            PyrParams t = p;
            t.l0 = p.l0 + full_groups * G * (int64_t)n0;
            for (int i = 1; i < fused; ++i) t.out[i] = p.out[i] + full_groups * G * (int64_t)p.n[i];
            t.Q = rest;
            t.G = (int)rest;
            const size_t tsm = (size_t)rest * per_map * sizeof(float);
            FFCORR_CUDA(cudaFuncSetAttribute(pyramid_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
            pyramid_fused_kernel<false><<<1, kPyrThreads, tsm, s>>>(t);
            if (int rc = check_launch("pyramid_fused_kernel(tail)")) return rc;
        }
        done_levels = fused;
    } else if (smem <= 200 * 1024) {
        p.G = G;
        const int64_t blocks = ceil_div64(Q, G);
        FFCORR_REQUIRE(blocks < (1ll << 31), FFCORR_EINVAL, "pyramid: grid too large");
        if (vec4) {
            FFCORR_CUDA(cudaFuncSetAttribute(pyramid_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            pyramid_fused_kernel<true><<<(unsigned)blocks, kPyrThreads, smem, s>>>(p);
        } else {
            FFCORR_CUDA(cudaFuncSetAttribute(pyramid_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            pyramid_fused_kernel<false><<<(unsigned)blocks, kPyrThreads, smem, s>>>(p);
        }
        if (int rc = check_launch("pyramid_fused_kernel")) return rc;
        done_levels = fused;
    }
    // remaining levels (or everything, for maps too large for shared memory): one pass per level
    for (int i = done_levels; i < num_levels; ++i) {
        const int hi = h >> (i - 1), wi = w >> (i - 1), ho = h >> i, wo = w >> i;
        pool_level_kernel<<<grid_for(Q * ho * wo, 256), 256, 0, s>>>(lvl[i - 1], lvl[i], Q, hi, wi, ho, wo);
        if (int rc = check_launch("pool_level_kernel")) return rc;
    }
    return FFCORR_OK;
}

extern "C" int ffcorr_pyramid_bwd_f32(float* const* grad_lvl, int num_levels, int64_t Q, int h, int w, void* stream) {
    FFCORR_REQUIRE(grad_lvl != nullptr, FFCORR_EINVAL, "pyramid_bwd: null level table");
    FFCORR_REQUIRE(Q >= 0, FFCORR_EINVAL, "pyramid_bwd: Q=%lld", (long long)Q);
    if (int rc = check_levels(num_levels, h, w, "pyramid_bwd")) return rc;
    if (Q == 0) return FFCORR_OK;
    for (int i = 0; i < num_levels; ++i)
        FFCORR_REQUIRE(grad_lvl[i] != nullptr, FFCORR_EINVAL, "pyramid_bwd: grad_lvl[%d] is null", i);
    cudaStream_t s = (cudaStream_t)stream;
    for (int i = num_levels - 1; i >= 1; --i) {
        const int hi = h >> (i - 1), wi = w >> (i - 1), ho = h >> i, wo = w >> i;
        if ((int64_t)ho * ((wo + 1) / 2) < (1ll << 30)) {
            const int64_t items = Q * ho * ((wo + 1) / 2);
            if (wi % 4 == 0 && (uintptr_t)grad_lvl[i - 1] % 16 == 0)
                pool_level_bwd_pair_kernel<true><<<grid_for(items, 256), 256, 0, s>>>(grad_lvl[i - 1], grad_lvl[i], Q, hi, wi, ho, wo);
            else
                pool_level_bwd_pair_kernel<false><<<grid_for(items, 256), 256, 0, s>>>(grad_lvl[i - 1], grad_lvl[i], Q, hi, wi, ho, wo);
        } else {
            pool_level_bwd_kernel<<<grid_for(Q * hi * wi, 256), 256, 0, s>>>(grad_lvl[i - 1], grad_lvl[i], Q, hi, wi, ho, wo);
        }
        if (int rc = check_launch("pool_level_bwd_kernel")) return rc;
    }
    return FFCORR_OK;
}


// ------------------------------------------------------------------------------------------ tiled layout
extern "C" int64_t ffcorr_tiled_map_elems(int h, int w, int level) {
    if (h < 1 || w < 1 || level < 0 || level >= FFCORR_MAX_LEVELS) return 0;
    const int hl = h >> level, wl = w >> level;
    if (hl < 1 || wl < 1) return 0;
    return (int64_t)tiled_th(hl) * tiled_tw(wl) * 16;
}

static int fill_tiled_params(TiledPyrParams* p, float* const* lvl, int num_levels, int h, int w) {
    for (int i = 0; i < num_levels; ++i) {
        p->h[i] = h >> i;
        p->w[i] = w >> i;
        p->th[i] = tiled_th(p->h[i]);
        p->tw[i] = tiled_tw(p->w[i]);
        p->np[i] = p->th[i] * p->tw[i] * 16;
        p->out[i] = lvl[i];
        p->magic_tw[i] = (unsigned)((0x100000000ull + p->tw[i] - 1) / (unsigned)p->tw[i]);
    }
    p->l0 = lvl[0];
    p->num_levels = num_levels;
    return FFCORR_OK;
}

static size_t tiled_pyramid_smem(const TiledPyrParams& p) {
    return ((size_t)kPyrStages * p.np[0] + (size_t)p.np[1] + p.np[2] + 16) * sizeof(float) + 64;
}

extern "C" int ffcorr_tiled_supported(int num_levels, int h, int w) {
    if (num_levels < 1 || num_levels > kFusedLevels || h < 1 || w < 1 || h > 16384 || w > 16384) return 0;
    if ((h >> (num_levels - 1)) < 1 || (w >> (num_levels - 1)) < 1) return 0;
    // limits of the fused build + tiled lookup: <= 4 levels (pooled in the GEMM epilogue), 32 query maps addressed
    // with 32-bit element offsets, h*w < 2^24 (volume.cu).  The standalone ffcorr_pyramid_tiled_f32 is narrower:
    // it stages whole maps in shared memory (see tiled_pyramid_standalone_ok).
    return (int64_t)h * w < (1ll << 24) && (int64_t)tiled_th(h) * tiled_tw(w) * 16 < (1ll << 25);
}

static bool tiled_pyramid_standalone_ok(const TiledPyrParams& p) {
    return tiled_pyramid_smem(p) <= 200 * 1024 && (uint64_t)p.np[0] * 4 < (1u << 20);
}

extern "C" int ffcorr_pyramid_tiled_f32(float* const* lvl, int num_levels, int64_t Q, int h, int w, void* stream) {
    FFCORR_REQUIRE(lvl != nullptr, FFCORR_EINVAL, "pyramid_tiled: null level table");
    FFCORR_REQUIRE(Q >= 0, FFCORR_EINVAL, "pyramid_tiled: Q=%lld", (long long)Q);
    if (int rc = check_levels(num_levels, h, w, "pyramid_tiled")) return rc;
    FFCORR_REQUIRE(ffcorr_tiled_supported(num_levels, h, w), FFCORR_EINVAL,
                   "pyramid_tiled: %dx%d with %d levels is outside the tiled path (use the row-major entry points)", h, w, num_levels);
    if (Q == 0 || num_levels == 1) return FFCORR_OK;
    for (int i = 0; i < num_levels; ++i) {
        FFCORR_REQUIRE(lvl[i] != nullptr, FFCORR_EINVAL, "pyramid_tiled: lvl[%d] is null", i);
        FFCORR_REQUIRE((uintptr_t)lvl[i] % 16 == 0, FFCORR_EALIGN, "pyramid_tiled: lvl[%d] must be 16-byte aligned", i);
    }
    TiledPyrParams p{};
    fill_tiled_params(&p, lvl, num_levels, h, w);
    FFCORR_REQUIRE(tiled_pyramid_standalone_ok(p), FFCORR_EINVAL,
                   "pyramid_tiled: a %dx%d map does not fit the standalone kernel's shared-memory staging; "
                   "use ffcorr_build_tiled_f32 (pyramid fused into the GEMM)", h, w);
    const size_t smem = tiled_pyramid_smem(p);
    const int per_sm = smem <= 100 * 1024 ? 2 : 1;
    const int64_t want = (int64_t)per_sm * sm_count();
    const int grid = (int)(Q < want ? Q : want);
    FFCORR_CUDA(cudaFuncSetAttribute(pyramid_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pyramid_tiled_kernel<<<grid, kPyrThreads, smem, (cudaStream_t)stream>>>(p, Q);
    return check_launch("pyramid_tiled_kernel");
}

extern "C" int ffcorr_untile_f32(const float* tiled, float* dst, int64_t Q, int h, int w, void* stream) {
    FFCORR_REQUIRE(Q >= 0 && h >= 1 && w >= 1, FFCORR_EINVAL, "untile: bad shape");
    if (Q == 0) return FFCORR_OK;
    FFCORR_REQUIRE(tiled && dst, FFCORR_EINVAL, "untile: null pointer");
    const int tw = tiled_tw(w);
    untile_kernel<<<grid_for(Q * h * w, 256), 256, 0, (cudaStream_t)stream>>>(tiled, dst, Q, h, w, tw, tiled_th(h) * tw * 16);
    return check_launch("untile_kernel");
}

extern "C" int64_t ffcorr_grouped_level_elems(int h, int w, int level, int B, int nq) {
    if (h < 1 || w < 1 || level < 0 || level >= FFCORR_MAX_LEVELS || B < 0 || nq < 0) return 0;
    const int lh = h >> level, lw = w >> level;
    if (lh < 1 || lw < 1) return 0;
    return (int64_t)B * ceil_div(nq, 32) * tiled_th(lh) * tiled_tw(lw) * 512;
}

extern "C" int ffcorr_ungroup_f32(const float* grouped, float* dst, int B, int N, int h, int w, void* stream) {
    FFCORR_REQUIRE(B >= 0 && N >= 1 && h >= 1 && w >= 1, FFCORR_EINVAL, "ungroup: bad shape");
    if (B == 0) return FFCORR_OK;
    FFCORR_REQUIRE(grouped && dst, FFCORR_EINVAL, "ungroup: null pointer");
    const int tw = tiled_tw(w);
    ungroup_kernel<<<grid_for((int64_t)B * N * h * w, 256), 256, 0, (cudaStream_t)stream>>>(grouped, dst, B, N, h, w, tw, tiled_th(h) * tw);
    return check_launch("ungroup_kernel");
}

extern "C" int ffcorr_group_f32(const float* src, float* grouped, int B, int N, int h, int w, void* stream) {
    FFCORR_REQUIRE(B >= 0 && N >= 1 && h >= 1 && w >= 1, FFCORR_EINVAL, "group: bad shape");
    if (B == 0) return FFCORR_OK;
    FFCORR_REQUIRE(grouped && src, FFCORR_EINVAL, "group: null pointer");
    const int tw = tiled_tw(w);
    const int ntiles = tiled_th(h) * tw;
    group_kernel<<<grid_for((int64_t)B * ceil_div(N, 32) * ntiles * 512, 256), 256, 0, (cudaStream_t)stream>>>(src, grouped, B, N, h, w, tw, ntiles);
    return check_launch("group_kernel");
}

extern "C" int ffcorr_untile_f16(const void* tiled, float* dst, int64_t Q, int h, int w, void* stream) {
    FFCORR_REQUIRE(Q >= 0 && h >= 1 && w >= 1, FFCORR_EINVAL, "untile_f16: bad shape");
    if (Q == 0) return FFCORR_OK;
    FFCORR_REQUIRE(tiled && dst, FFCORR_EINVAL, "untile_f16: null pointer");
    const int tw = tiled_tw(w);
    untile_half_kernel<<<grid_for(Q * h * w, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __half*>(tiled), dst, Q, h, w,
                                                                                  tw, tiled_th(h) * tw * 16);
    return check_launch("untile_half_kernel");
}

extern "C" int ffcorr_tile_f16(const float* src, void* tiled, int64_t Q, int h, int w, void* stream) {
    FFCORR_REQUIRE(Q >= 0 && h >= 1 && w >= 1, FFCORR_EINVAL, "tile_f16: bad shape");
    if (Q == 0) return FFCORR_OK;
    FFCORR_REQUIRE(tiled && src, FFCORR_EINVAL, "tile_f16: null pointer");
    const int tw = tiled_tw(w);
    const int np = tiled_th(h) * tw * 16;
    tile_half_kernel<<<grid_for(Q * np, 256), 256, 0, (cudaStream_t)stream>>>(src, reinterpret_cast<__half*>(tiled), Q, h, w, tw, np);
    return check_launch("tile_half_kernel");
}

extern "C" int ffcorr_tile_f32(const float* src, float* tiled, int64_t Q, int h, int w, void* stream) {
    FFCORR_REQUIRE(Q >= 0 && h >= 1 && w >= 1, FFCORR_EINVAL, "tile: bad shape");
    if (Q == 0) return FFCORR_OK;
    FFCORR_REQUIRE(tiled && src, FFCORR_EINVAL, "tile: null pointer");
    const int tw = tiled_tw(w);
    const int np = tiled_th(h) * tw * 16;
    tile_kernel<<<grid_for(Q * np, 256), 256, 0, (cudaStream_t)stream>>>(src, tiled, Q, h, w, tw, np);
    return check_launch("tile_kernel");
}
