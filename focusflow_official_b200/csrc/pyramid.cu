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
    int G;                      // maps per block
    int64_t Q;
};

__device__ __forceinline__ float pool4(const float* s, int w) {
    return (((s[0] + s[1]) + s[w]) + s[w + 1]) * 0.25f;
}

template <bool VEC4>
__global__ void __launch_bounds__(kPyrThreads) pyramid_fused_kernel(const PyrParams p) {
    extern __shared__ __align__(16) float sm[];
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

    // ---- levels 1.. from shared memory: a warp per output row, lanes along x ----
    // (one integer division per ROW, not per element; 8-byte smem reads when the input
    //  width is even so a lane fetches its 2x2 window with two LDS.64)
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int NW = kPyrThreads / 32;
    float* sin = sm;
    float* sout = sm + (size_t)p.G * p.n[0];
#pragma unroll 1
    for (int l = 1; l < p.num_levels; ++l) {
        const int wi = p.w[l - 1], ni = p.n[l - 1];
        const int ho = p.h[l], wo = p.w[l], no = p.n[l];
        float* __restrict__ dst = p.out[l] + q0 * no;
        const int rows = g * ho;
        const bool even = ((wi & 1) == 0) && (((sin - sm) & 1) == 0);
        for (int row = warp; row < rows; row += NW) {
            const int m = row / ho;
            const int y = row - m * ho;
            const float* r0 = sin + m * ni + (2 * y) * wi;
            const int o = m * no + y * wo;
            if (even) {
                const float2* a2 = reinterpret_cast<const float2*>(r0);
                const float2* b2 = reinterpret_cast<const float2*>(r0 + wi);
                for (int x = lane; x < wo; x += 32) {
                    const float2 a = a2[x], bb = b2[x];
                    const float v = (((a.x + a.y) + bb.x) + bb.y) * 0.25f;
                    sout[o + x] = v;
                    dst[o + x] = v;
                }
            } else {
                for (int x = lane; x < wo; x += 32) {
                    const float v = pool4(r0 + 2 * x, wi);
                    sout[o + x] = v;
                    dst[o + x] = v;
                }
            }
        }
        __syncthreads();
        sin = sout;
        sout = sout + (size_t)p.G * no;
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
        per_map += p.n[i];
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
    if (smem <= 200 * 1024) {
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
        pool_level_bwd_kernel<<<grid_for(Q * hi * wi, 256), 256, 0, s>>>(grad_lvl[i - 1], grad_lvl[i], Q, hi, wi, ho, wo);
        if (int rc = check_launch("pool_level_bwd_kernel")) return rc;
    }
    return FFCORR_OK;
}
