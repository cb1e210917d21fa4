// PWC 9x9 local cost volume (max displacement 4, 81 channels) for sm_100a.
//
// Replaces PWCNet_Core/correlation.py:278-328: two padded-NHWC "rearrange" copies
// (3 memsets + 2 full copies) followed by a one-block-per-pixel kernel with 81
// __syncthreads rounds.  Here ONE kernel reads NCHW directly (no padded copy):
//   out[b, (dy+4)*9 + (dx+4), y, x] = (1/C) * sum_c one[b,c,y,x] * two[b,c,y+dy,x+dx]
//
// Decomposition: a block owns an 8 x 32 pixel tile of one batch item and streams the channels
// through a 2-stage cp.async ring, 8 channels per stage (tile of `one`, tile + 4-pixel halo of
// `two`, zero filled outside the image).  Warp w (0..8) owns displacement row dy = w - 4;
// lane -> 8-pixel strip.  Per channel a thread reads 8 + 16 floats (6 x LDS.128) and issues
// 72 FMAs (8 pixels x 9 dx), accumulators stay in registers for all C channels.
// Roofline: 4*(2C+81) bytes and 162*C flop per pixel -> HBM-bound for C = 32, FP32-FMA
// bound for C >= 64.
#include "common.cuh"

namespace ffcorr {
namespace {

constexpr int PT_Y = 8;             // tile rows
constexpr int PT_X = 32;            // tile cols
constexpr int PCC = 8;              // channels per pipeline stage
constexpr int PHALO = 4;
constexpr int PW2 = PT_X + 2 * PHALO;   // 40 columns of `two` per tile row
constexpr int PH2 = PT_Y + 2 * PHALO;   // 16 rows
constexpr int PITCH2 = 44;          // smem row pitches chosen so the 16-byte reads of a quarter warp
constexpr int PITCH1 = 36;          //   (2 rows x 4 strips) fall into 8 distinct bank groups
constexpr int S2 = PH2 * PITCH2;    // floats per channel, `two` tile + halo
constexpr int S1 = PT_Y * PITCH1;   // floats per channel, `one` tile
constexpr int STAGE_FLOATS = PCC * (S2 + S1);
constexpr int PWC_STAGES = 3;
constexpr int PWC_THREADS = 9 * 32;
constexpr int CH2 = PCC * PH2 * (PW2 / 4);   // 16-byte chunks of `two` per stage (1280)
constexpr int CH1 = PCC * PT_Y * (PT_X / 4); // 16-byte chunks of `one` per stage (512)
constexpr int NCHUNK = (CH2 + CH1 + PWC_THREADS - 1) / PWC_THREADS;  // per thread (7)
constexpr size_t PWC_SMEM = (size_t)PWC_STAGES * STAGE_FLOATS * sizeof(float);

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One block = 8 x 32 pixel tile of one batch item; warp w owns displacement row dy = w - 4, a lane owns an
// 8-pixel strip: per channel 2 + 4 LDS.128 feed 72 FMAs (8 pixels x 9 dx), accumulators stay in
// registers across all C channels.  Channels stream through a 2-stage cp.async ring (16-byte copies with
// zero fill outside the image); ALIGNED = false is the scalar-load path for W % 4 != 0 / unaligned bases.
template <bool ALIGNED>
__global__ void __launch_bounds__(PWC_THREADS, 1) pwc81_kernel(const float* __restrict__ one, const float* __restrict__ two,
                                                               float* __restrict__ out, int B, int C, int H, int W,
                                                               int tiles_x, int tiles_y, float leaky_slope) {
    extern __shared__ __align__(16) float psm[];
    const int tid = threadIdx.x;
    const int dyi = tid >> 5;            // 0..8  (warp-uniform)
    const int lane = tid & 31;
    const int row = lane >> 2;           // 0..7
    const int c8 = (lane & 3) * 8;       // strip start column within the tile
    const size_t plane = (size_t)H * W;
    const int nstage = (C + PCC - 1) / PCC;
    const int tiles_per_b = tiles_x * tiles_y;
    const int num_tiles = tiles_per_b * B;
    const uint32_t psm_u32 = (uint32_t)__cvta_generic_to_shared(psm);

    // ---- load cursor: runs PWC_STAGES-1 stages ahead of the compute cursor, across tile boundaries ----
    // plan[i]: bits [0,16) float offset inside a stage, [16,20) channel in stage, 20 inside image,
    //          21 chunk belongs to `two`, 22 chunk exists;  goff[i]: element offset inside the plane.
    int plan[NCHUNK], goff[NCHUNK];
    int l_tile = blockIdx.x, l_stage = 0, l_seq = 0;
    const float* l_one = one;
    const float* l_two = two;

    auto make_plan = [&](int tile) {
        const int b = tile / tiles_per_b;
        const int r2 = tile - b * tiles_per_b;
        const int y0 = (r2 / tiles_x) * PT_Y, x0 = (r2 % tiles_x) * PT_X;
        l_one = one + (size_t)b * C * plane;
        l_two = two + (size_t)b * C * plane;
#pragma unroll
        for (int i = 0; i < NCHUNK; ++i) {
            const int id = tid + i * PWC_THREADS;
            plan[i] = 0;
            goff[i] = 0;
            if (id < CH2) {
                const int c = id / (PH2 * (PW2 / 4));
                const int r = (id / (PW2 / 4)) % PH2;
                const int k = id % (PW2 / 4);
                const int gy = y0 - PHALO + r, gx = x0 - PHALO + 4 * k;
                const bool in = (unsigned)gy < (unsigned)H && gx >= 0 && gx < W;
                plan[i] = (c * S2 + r * PITCH2 + 4 * k) | (c << 16) | ((int)in << 20) | (1 << 21) | (1 << 22);
                goff[i] = gy * W + gx;
            } else if (id < CH2 + CH1) {
                const int id2 = id - CH2;
                const int c = id2 / (PT_Y * (PT_X / 4));
                const int r = (id2 / (PT_X / 4)) % PT_Y;
                const int k = id2 % (PT_X / 4);
                const int gy = y0 + r, gx = x0 + 4 * k;
                const bool in = gy < H && gx < W;
                plan[i] = (PCC * S2 + c * S1 + r * PITCH1 + 4 * k) | (c << 16) | ((int)in << 20) | (1 << 22);
                goff[i] = gy * W + gx;
            }
        }
    };

    auto issue_next = [&]() {
        if (l_tile < num_tiles) {
            if (l_stage == 0) make_plan(l_tile);
            const int cb = l_stage * PCC;
            const int slot = l_seq % PWC_STAGES;
            float* dstf = psm + slot * STAGE_FLOATS;
            const uint32_t dst = psm_u32 + (uint32_t)(slot * STAGE_FLOATS * sizeof(float));
#pragma unroll
            for (int i = 0; i < NCHUNK; ++i) {
                if (!(plan[i] & (1 << 22))) continue;
                const int so = plan[i] & 0xFFFF;
                const int ch = cb + ((plan[i] >> 16) & 15);
                const bool is_two = plan[i] & (1 << 21);
                const float* base = (is_two ? l_two : l_one) + (size_t)ch * plane;
                if (ALIGNED) {
                    const bool ok = (plan[i] & (1 << 20)) && ch < C;
                    cp_async16(dst + (uint32_t)so * 4u, ok ? (const void*)(base + goff[i]) : (const void*)one, ok ? 16u : 0u);
                } else {
                    // scalar path: the 4 columns of a chunk are checked individually (W % 4 != 0)
                    float v[4] = {0.f, 0.f, 0.f, 0.f};
                    if (ch < C) {
                        const int bb = l_tile / tiles_per_b;
                        const int r2 = l_tile - bb * tiles_per_b;
                        const int y0 = (r2 / tiles_x) * PT_Y, x0 = (r2 % tiles_x) * PT_X;
                        const int rem = is_two ? so % S2 : (so - PCC * S2) % S1;
                        const int r = is_two ? rem / PITCH2 : rem / PITCH1;
                        const int k4 = is_two ? rem % PITCH2 : rem % PITCH1;
                        const int gy = is_two ? y0 - PHALO + r : y0 + r;
                        const int gx = is_two ? x0 - PHALO + k4 : x0 + k4;
                        if ((unsigned)gy < (unsigned)H) {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if ((unsigned)(gx + e) < (unsigned)W) v[e] = __ldg(base + (size_t)gy * W + gx + e);
                        }
                    }
                    *reinterpret_cast<float4*>(dstf + so) = make_float4(v[0], v[1], v[2], v[3]);
                }
            }
            if (++l_stage == nstage) {
                l_stage = 0;
                l_tile += gridDim.x;
            }
        }
        ++l_seq;
        cp_async_commit();  // always one group per call (possibly empty) so wait counts stay aligned
    };

#pragma unroll 1
    for (int k = 0; k < PWC_STAGES - 1; ++k) issue_next();

    int c_seq = 0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        float acc[8][9];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 9; ++j) acc[i][j] = 0.0f;

#pragma unroll 1
        for (int st = 0; st < nstage; ++st, ++c_seq) {
            issue_next();                       // refills the slot consumed in the previous iteration
            cp_async_wait<PWC_STAGES - 1>();    // the group of the current compute stage has landed
            __syncthreads();
            const float* s2 = psm + (c_seq % PWC_STAGES) * STAGE_FLOATS;
            const float* s1 = s2 + PCC * S2;
#pragma unroll 2
            for (int c = 0; c < PCC; ++c) {
                const float* ap = s1 + c * S1 + row * PITCH1 + c8;
                const float* tp = s2 + c * S2 + (row + dyi) * PITCH2 + c8;
                const float4 a0 = *reinterpret_cast<const float4*>(ap);
                const float4 a1 = *reinterpret_cast<const float4*>(ap + 4);
                const float4 t0 = *reinterpret_cast<const float4*>(tp);
                const float4 t1 = *reinterpret_cast<const float4*>(tp + 4);
                const float4 t2 = *reinterpret_cast<const float4*>(tp + 8);
                const float4 t3 = *reinterpret_cast<const float4*>(tp + 12);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float t[16] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w,
                                     t2.x, t2.y, t2.z, t2.w, t3.x, t3.y, t3.z, t3.w};
#pragma unroll
                for (int px = 0; px < 8; ++px)
#pragma unroll
                    for (int dx = 0; dx < 9; ++dx) acc[px][dx] = fmaf(a[px], t[px + dx], acc[px][dx]);
            }
            __syncthreads();
        }

        // ---- epilogue: /C (true division like correlation.py:97), optional fused leaky_relu ----
        const int b = tile / tiles_per_b;
        const int r2 = tile - b * tiles_per_b;
        const int gy = (r2 / tiles_x) * PT_Y + row;
        const int gx = (r2 % tiles_x) * PT_X + c8;
        if (gy < H) {
            const float fc = (float)C;
            float* __restrict__ o = out + (((size_t)b * 81 + (size_t)dyi * 9) * H + gy) * W + gx;
            const bool vec = ALIGNED && (gx + 7 < W);
#pragma unroll
            for (int dx = 0; dx < 9; ++dx) {
                float r[8];
#pragma unroll
                for (int px = 0; px < 8; ++px) {
                    float v = __fdiv_rn(acc[px][dx], fc);
                    if (leaky_slope >= 0.0f) v = v > 0.0f ? v : v * leaky_slope;
                    r[px] = v;
                }
                float* od = o + (size_t)dx * plane;
                if (vec) {
                    reinterpret_cast<float4*>(od)[0] = make_float4(r[0], r[1], r[2], r[3]);
                    reinterpret_cast<float4*>(od)[1] = make_float4(r[4], r[5], r[6], r[7]);
                } else {
#pragma unroll
                    for (int px = 0; px < 8; ++px)
                        if (gx + px < W) od[px] = r[px];
                }
            }
        }
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------
// TMA-fed variant (W % 4 == 0): per stage ONE elected thread issues two 4-D tensor copies
// (`two` tile + halo, `one` tile; out-of-image and past-C elements are zero-filled by the TMA unit),
// so the 288 compute threads spend no instructions or registers on staging.  3-stage mbarrier ring.
// ------------------------------------------------------------------------------------------
constexpr int TMA_STAGES = 3;        // stage sizes depend on the tile height: PwcGeom<TY>

__device__ __forceinline__ void pw_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void pw_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pw_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        if (clock64() - t0 > 4000000000ll) __trap();  // protocol bug -> launch error instead of a hang
    }
}
__device__ __forceinline__ void pw_tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// Tile height TY: 8 (9 warps, a warp = one dy, lanes = 8 rows x 4 strips) for levels with enough tiles to fill the SMs;
// 4 (5 warps, a warp = TWO dy, lanes = dy-half x 4 rows x 4 strips; the tenth dy slot idles) for the small levels:
// twice the CTAs at a third less shared memory each, so 28x64 and below -- latency-bound at <= 128 tiles for 148 SMs --
// run two to three CTAs per SM instead of one.
template <int TY> struct PwcGeom {
    static constexpr int PH = TY + 2 * PHALO;                 // rows of `two` per tile
    static constexpr int SZ2 = PH * PITCH2;                   // floats per channel, `two` tile + halo
    static constexpr int SZ1 = TY * PITCH1;                   // floats per channel, `one` tile
    static constexpr int STAGE_BYTES = PCC * (SZ2 + SZ1) * 4;
    static constexpr int THREADS = TY == 8 ? 9 * 32 : 5 * 32;
    static constexpr size_t SMEM = (size_t)TMA_STAGES * STAGE_BYTES + 1024 + 64;
    static_assert((PCC * SZ2 * 4) % 128 == 0 && STAGE_BYTES % 128 == 0, "TMA destinations must stay 128-byte aligned");
    static_assert(72 * THREADS * 4 <= TMA_STAGES * STAGE_BYTES, "the cluster reduce reuses the stage ring");
};

template <bool POW2, int TY>
__global__ void __launch_bounds__(PwcGeom<TY>::THREADS, TY == 8 ? 2 : 3)
pwc81_tma_kernel(const __grid_constant__ CUtensorMap tm_one, const __grid_constant__ CUtensorMap tm_two,
                 float* __restrict__ out, int C, int H, int W, float inv_c, float leaky_slope) {
    using G = PwcGeom<TY>;
    constexpr int S2 = G::SZ2, S1 = G::SZ1, TMA_STAGE_BYTES = G::STAGE_BYTES, PWC_THREADS = G::THREADS, PT_Y = TY;
    // NOTE: index the extern array directly -- rounding the pointer up through uintptr_t makes the
    // compiler lose the shared address space and emit generic LD.E instead of LDS for the hot loop.
    extern __shared__ __align__(1024) uint8_t base[];
    const uint32_t base_u32 = (uint32_t)__cvta_generic_to_shared(base);
    if (base_u32 & 127u) __trap();  // TMA destinations need 128-byte alignment
    const uint32_t bar0 = base_u32 + TMA_STAGES * TMA_STAGE_BYTES;

    // Channel split across a thread-block cluster (small pyramid levels have too few pixel tiles to
    // fill 148 SMs): CTA `crank` of `csize` takes stages crank, crank + csize, ...; partial sums are
    // reduced in rank order through distributed shared memory, so the result is deterministic.
    uint32_t crank, csize;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
    const int b = blockIdx.z;
    const int y0 = blockIdx.y * PT_Y, x0 = (blockIdx.x / csize) * PT_X;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int dyi = TY == 8 ? (tid >> 5) : (tid >> 5) * 2 + (lane >> 4);      // TY == 4: dy slot 9 does not exist
    const bool dy_on = dyi < 9;
    const int row = TY == 8 ? (lane >> 2) : ((lane >> 2) & 3);
    const int c8 = (lane & 3) * 8;
    const int nstage_all = (C + PCC - 1) / PCC;
    const int nstage = (nstage_all - (int)crank + (int)csize - 1) / (int)csize;   // my share

    auto issue = [&](int st_local) {  // thread 0 only
        const int st = (int)crank + st_local * (int)csize;   // global stage -> channel offset
        const int slot = st_local % TMA_STAGES;
        const uint32_t bar = bar0 + 8u * slot;
        const uint32_t dst = base_u32 + (uint32_t)(slot * TMA_STAGE_BYTES);
        pw_mbar_expect_tx(bar, TMA_STAGE_BYTES);
        pw_tma_load_4d(dst, &tm_two, bar, x0 - PHALO, y0 - PHALO, st * PCC, b);
        pw_tma_load_4d(dst + PCC * S2 * 4, &tm_one, bar, x0, y0, st * PCC, b);
    };

    if (tid == 0) {
        for (int s = 0; s < TMA_STAGES; ++s) pw_mbar_init(bar0 + 8u * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const int pre = nstage < TMA_STAGES ? nstage : TMA_STAGES;
        for (int s = 0; s < pre; ++s) issue(s);
    }

    float acc[8][9];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 9; ++j) acc[i][j] = 0.0f;

#pragma unroll 1
    for (int st = 0; st < nstage; ++st) {
        const int slot = st % TMA_STAGES;
        pw_mbar_wait(bar0 + 8u * slot, (uint32_t)((st / TMA_STAGES) & 1));
        const float* s2 = reinterpret_cast<const float*>(base + slot * TMA_STAGE_BYTES);
        const float* s1 = s2 + PCC * S2;
#pragma unroll 1
        for (int c = 0; c < PCC; ++c) {
            const float* ap = s1 + c * S1 + row * PITCH1 + c8;
            const float* tp = s2 + c * S2 + (row + (dy_on ? dyi : 8)) * PITCH2 + c8;   // the idle slot reads a valid row
            const float4 a0 = *reinterpret_cast<const float4*>(ap);
            const float4 a1 = *reinterpret_cast<const float4*>(ap + 4);
            const float4 t0 = *reinterpret_cast<const float4*>(tp);
            const float4 t1 = *reinterpret_cast<const float4*>(tp + 4);
            const float4 t2 = *reinterpret_cast<const float4*>(tp + 8);
            const float4 t3 = *reinterpret_cast<const float4*>(tp + 12);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float t[16] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w,
                                 t2.x, t2.y, t2.z, t2.w, t3.x, t3.y, t3.z, t3.w};
#pragma unroll
            for (int px = 0; px < 8; ++px)
#pragma unroll
                for (int dx = 0; dx < 9; ++dx) acc[px][dx] = fmaf(a[px], t[px + dx], acc[px][dx]);
        }
        __syncthreads();  // every warp is done with this slot
        if (tid == 0 && st + TMA_STAGES < nstage) issue(st + TMA_STAGES);
    }

    if (csize > 1) {
        // Deterministic reduce-scatter through distributed shared memory: every CTA publishes its 72
        // partial sums per thread ([k][tid], conflict-free), then CTA `crank` finishes the slice
        // k in [crank*72/csize, (crank+1)*72/csize): it adds the partials of ranks 0..csize-1 IN ORDER and
        // writes those outputs itself.  (A gather to rank 0 would be DSMEM-bandwidth bound: ~20 B/cycle.)
        float* red = reinterpret_cast<float*>(base);
#pragma unroll
        for (int px = 0; px < 8; ++px)
#pragma unroll
            for (int dx = 0; dx < 9; ++dx) red[(px * 9 + dx) * PWC_THREADS + tid] = acc[px][dx];
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        const int per = 72 / (int)csize;
        const int gy = y0 + row;
        const float fc = (float)C;
        const uint32_t mine = base_u32 + (uint32_t)tid * 4u;
        for (int kk = 0; kk < per; ++kk) {
            const int k = (int)crank * per + kk;
            float sum = 0.0f;
            for (uint32_t r = 0; r < csize; ++r) {
                uint32_t remote;
                float v;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(mine + (uint32_t)(k * PWC_THREADS * 4)), "r"(r));
                asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote));
                sum += v;
            }
            const int px = k / 9, dx = k - px * 9;
            const int gx = x0 + c8 + px;
            if (dy_on && gy < H && gx < W) {
                float v = POW2 ? sum * inv_c : __fdiv_rn(sum, fc);
                if (leaky_slope >= 0.0f) v = v > 0.0f ? v : v * leaky_slope;
                out[(((size_t)b * 81 + (size_t)dyi * 9 + dx) * H + gy) * W + gx] = v;
            }
        }
        // every CTA must stay resident until all peers have read its shared memory
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        return;
    }

    // ---- epilogue: /C (exact reciprocal multiply when C is a power of two, else a true division like
    //      correlation.py:97), optional fused leaky_relu ----
    const int gy = y0 + row;
    if (gy >= H || !dy_on) return;
    const float fc = (float)C;
    const int gx = x0 + c8;
    const size_t plane = (size_t)H * W;
    float* __restrict__ o = out + (((size_t)b * 81 + (size_t)dyi * 9) * H + gy) * W + gx;
    const bool vec = gx + 7 < W;
#pragma unroll
    for (int dx = 0; dx < 9; ++dx) {
        float r[8];
#pragma unroll
        for (int px = 0; px < 8; ++px) {
            float v = POW2 ? acc[px][dx] * inv_c : __fdiv_rn(acc[px][dx], fc);
            if (leaky_slope >= 0.0f) v = v > 0.0f ? v : v * leaky_slope;
            r[px] = v;
        }
        float* od = o + (size_t)dx * plane;
        if (vec) {
            reinterpret_cast<float4*>(od)[0] = make_float4(r[0], r[1], r[2], r[3]);
            reinterpret_cast<float4*>(od)[1] = make_float4(r[4], r[5], r[6], r[7]);
        } else {
#pragma unroll
            for (int px = 0; px < 8; ++px)
                if (gx + px < W) od[px] = r[px];
        }
    }
}

// ------------------------------------------------------------------------------------------
// Gradients (correlation.py:104-232, 331-380) on the forward's machinery.  Both have the form
//   R[c, p] = (1/C) sum_k G_k[p] * T[c, p + s * d_k],      k = (dy, dx), d_k = (dy, dx):
//   grad_one:  T = two, s = +1, G_k[p] = g[k, p]
//   grad_two:  T = one, s = -1, G_k[p] = g[k, p - d_k]      (the gradient plane shifted by its displacement)
// A block owns an 8x32-pixel tile of one batch item and 32 channels: the 32 T planes (+4-pixel halo) are loaded
// once by TMA (4 boxes of 8 channels, zero fill = zero padding), the gradient planes stream through a 2-slot
// ring, 9 planes (one dy) per slot.  For grad_two every plane is fetched with its own row shift; the column shift
// cannot be put into the TMA coordinate (an inner coordinate that is not a multiple of 16 bytes faults), so the
// planes carry a 4-pixel halo and the shift is applied when the registers are filled.  Warp w owns channels 4w..4w+3, lane = 8-pixel strip (as in the forward): per dy a
// thread reads 4 x 16 floats of T and 9 x 8 floats of G (34 LDS.128) for 288 FMAs into 32 accumulators.
// ------------------------------------------------------------------------------------------
constexpr int BW_CH = 32;                         // channels per block
constexpr int BW_CT = 4;                          // channels per thread
constexpr int BW_THREADS = (BW_CH / BW_CT) * 32;  // 256
constexpr int BW_T_BYTES = BW_CH * S2 * 4;        // 90112
constexpr int BW_G_PITCH = PITCH2;                // 44: tile + 4-pixel halo left and right, conflict-free LDS.128
constexpr int BW_G_PLANE = PT_Y * BW_G_PITCH;     // 352 floats
constexpr int BW_G_BYTES = 9 * BW_G_PLANE * 4;    // 12672 per dy
constexpr size_t PWC_BWD_SMEM = (size_t)BW_T_BYTES + 2 * BW_G_BYTES + 64;
static_assert(BW_T_BYTES % 128 == 0 && BW_G_BYTES % 128 == 0 && (BW_G_PLANE * 4) % 128 == 0, "TMA destinations must stay 128-byte aligned");

template <bool NEG, bool POW2>
__global__ void __launch_bounds__(BW_THREADS, 2)
pwc81_bwd_tma_kernel(const __grid_constant__ CUtensorMap tm_t, const __grid_constant__ CUtensorMap tm_g,
                     float* __restrict__ out, int C, int H, int W, int cstages, float inv_c) {
    extern __shared__ __align__(1024) uint8_t base[];
    const uint32_t base_u32 = (uint32_t)__cvta_generic_to_shared(base);
    if (base_u32 & 127u) __trap();
    const uint32_t g_u32 = base_u32 + BW_T_BYTES;
    const uint32_t bar_t = g_u32 + 2 * BW_G_BYTES;
    const uint32_t bar_g = bar_t + 8;             // two barriers
    const int b = blockIdx.z;
    const int y0 = blockIdx.y * PT_Y;
    const int x0 = (blockIdx.x / cstages) * PT_X;
    const int cbase = (blockIdx.x % cstages) * BW_CH;
    const int tid = threadIdx.x;
    const int cg = tid >> 5;                      // warp = channel group
    const int lane = tid & 31;
    const int row = lane >> 2;
    const int c8 = (lane & 3) * 8;

    auto issue_g = [&](int dyi) {                 // thread 0 only: the 9 gradient planes of displacement row dyi
        const uint32_t bar = bar_g + 8u * (dyi & 1);
        const uint32_t dst = g_u32 + (uint32_t)((dyi & 1) * BW_G_BYTES);
        pw_mbar_expect_tx(bar, BW_G_BYTES);
#pragma unroll 1
        for (int dxi = 0; dxi < 9; ++dxi) {
            const int gy = NEG ? y0 - (dyi - 4) : y0;
            pw_tma_load_4d(dst + (uint32_t)(dxi * BW_G_PLANE * 4), &tm_g, bar, x0 - PHALO, gy, dyi * 9 + dxi, b);
        }
    };

    if (tid == 0) {
        pw_mbar_init(bar_t, 1);
        pw_mbar_init(bar_g, 1);
        pw_mbar_init(bar_g + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        pw_mbar_expect_tx(bar_t, BW_T_BYTES);
#pragma unroll 1
        for (int i = 0; i < BW_CH / PCC; ++i)
            pw_tma_load_4d(base_u32 + (uint32_t)(i * PCC * S2 * 4), &tm_t, bar_t, x0 - PHALO, y0 - PHALO, cbase + i * PCC, b);
        issue_g(0);
        issue_g(1);
    }

    float acc[BW_CT][8];
#pragma unroll
    for (int j = 0; j < BW_CT; ++j)
#pragma unroll
        for (int px = 0; px < 8; ++px) acc[j][px] = 0.0f;

    pw_mbar_wait(bar_t, 0);
    const float* Ts = reinterpret_cast<const float*>(base) + (size_t)(cg * BW_CT) * S2 + c8;
#pragma unroll 1
    for (int dyi = 0; dyi < 9; ++dyi) {
        pw_mbar_wait(bar_g + 8u * (dyi & 1), (uint32_t)((dyi >> 1) & 1));
        const float* Gs = reinterpret_cast<const float*>(base + BW_T_BYTES + (dyi & 1) * BW_G_BYTES) + row * BW_G_PITCH + c8;
        const int trow = row + (NEG ? 8 - dyi : dyi);
        float t[BW_CT][16];
#pragma unroll
        for (int j = 0; j < BW_CT; ++j) {
            const float* tp = Ts + j * S2 + trow * PITCH2;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 v = *reinterpret_cast<const float4*>(tp + 4 * q);
                t[j][4 * q] = v.x; t[j][4 * q + 1] = v.y; t[j][4 * q + 2] = v.z; t[j][4 * q + 3] = v.w;
            }
        }
#pragma unroll
        for (int dxi = 0; dxi < 9; ++dxi) {
            // plane column of pixel px: c8 + px + 4 (halo) - dx for grad_two, c8 + px + 4 for grad_one
            const int gs = NEG ? 8 - dxi : PHALO;                 // compile-time after unrolling
            const float* gp = Gs + dxi * BW_G_PLANE + (gs & ~3);
            const float4 w0 = *reinterpret_cast<const float4*>(gp);
            const float4 w1 = *reinterpret_cast<const float4*>(gp + 4);
            float4 w2 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gs & 3) w2 = *reinterpret_cast<const float4*>(gp + 8);
            const float wv[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
            float g8[8];
#pragma unroll
            for (int px = 0; px < 8; ++px) g8[px] = wv[(gs & 3) + px];
            const int sh = NEG ? 8 - dxi : dxi;
#pragma unroll
            for (int j = 0; j < BW_CT; ++j)
#pragma unroll
                for (int px = 0; px < 8; ++px) acc[j][px] = fmaf(g8[px], t[j][px + sh], acc[j][px]);
        }
        __syncthreads();                           // everyone is done with this gradient slot
        if (tid == 0 && dyi + 2 < 9) issue_g(dyi + 2);
    }

    const int gy = y0 + row, gx = x0 + c8;
    if (gy >= H || gx >= W) return;
    const float fc = (float)C;
#pragma unroll
    for (int j = 0; j < BW_CT; ++j) {
        const int c = cbase + cg * BW_CT + j;
        if (c >= C) break;
        float r[8];
#pragma unroll
        for (int px = 0; px < 8; ++px) r[px] = POW2 ? acc[j][px] * inv_c : __fdiv_rn(acc[j][px], fc);
        float* od = out + (((size_t)b * C + c) * H + gy) * W + gx;
        if (gx + 7 < W) {
            reinterpret_cast<float4*>(od)[0] = make_float4(r[0], r[1], r[2], r[3]);
            reinterpret_cast<float4*>(od)[1] = make_float4(r[4], r[5], r[6], r[7]);
        } else {
#pragma unroll
            for (int px = 0; px < 8; ++px)
                if (gx + px < W) od[px] = r[px];
        }
    }
}

// Fallback gradients (W % 4 != 0 or unaligned bases).  One thread per input element, 81 taps each.
//   gone[b,c,y,x] = (1/C) sum_{p,o} g[b,(p,o),y,x]       * two[b,c,y+p,x+o]
//   gtwo[b,c,y,x] = (1/C) sum_{p,o} g[b,(p,o),y-p,x-o]   * one[b,c,y-p,x-o]
__global__ void __launch_bounds__(256) pwc81_bwd_kernel(const float* __restrict__ one, const float* __restrict__ two,
                                                        const float* __restrict__ g, float* __restrict__ gone,
                                                        float* __restrict__ gtwo, int B, int C, int H, int W) {
    const int64_t total = (int64_t)B * C * H * W;
    const size_t plane = (size_t)H * W;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(idx % W);
        int64_t t = idx / W;
        const int y = (int)(t % H);
        t /= H;
        const int c = (int)(t % C);
        const int b = (int)(t / C);
        const float* gb = g + (size_t)b * 81 * plane;
        const float* one_c = one + ((size_t)b * C + c) * plane;
        const float* two_c = two + ((size_t)b * C + c) * plane;
        float s1 = 0.0f, s2 = 0.0f;
        for (int p = -4; p <= 4; ++p) {
            for (int o = -4; o <= 4; ++o) {
                const int op = (p + 4) * 9 + (o + 4);
                if (gone) {
                    const int yy = y + p, xx = x + o;
                    if ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W)
                        s1 = fmaf(__ldg(gb + op * plane + (size_t)y * W + x), __ldg(two_c + (size_t)yy * W + xx), s1);
                }
                if (gtwo) {
                    const int yy = y - p, xx = x - o;
                    if ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W)
                        s2 = fmaf(__ldg(gb + op * plane + (size_t)yy * W + xx), __ldg(one_c + (size_t)yy * W + xx), s2);
                }
            }
        }
        if (gone) gone[idx] = __fdiv_rn(s1, (float)C);
        if (gtwo) gtwo[idx] = __fdiv_rn(s2, (float)C);
    }
}

}  // namespace
}  // namespace ffcorr

using namespace ffcorr;

extern "C" int ffcorr_pwc81_f32(const float* one, const float* two, float* out, int B, int C, int H, int W,
                                float leaky_slope, void* stream) {
    FFCORR_REQUIRE(B >= 0 && C >= 1 && H >= 1 && W >= 1, FFCORR_EINVAL, "pwc81: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
    if (B == 0) return FFCORR_OK;
    FFCORR_REQUIRE(one && two && out, FFCORR_EINVAL, "pwc81: null pointer");
    FFCORR_REQUIRE(leaky_slope < 1.0f, FFCORR_EINVAL, "pwc81: leaky_slope=%f must be < 1 (negative = off)", leaky_slope);
    if (B == 0) return FFCORR_OK;
    const int tiles_x = ceil_div(W, PT_X), tiles_y = ceil_div(H, PT_Y);
    const int64_t num_tiles = (int64_t)tiles_x * tiles_y * B;
    FFCORR_REQUIRE(num_tiles < (1ll << 31), FFCORR_EINVAL, "pwc81: too many tiles");
    FFCORR_REQUIRE((int64_t)H * W < (1ll << 31), FFCORR_EINVAL, "pwc81: H*W too large");
    const int grid = (int)(num_tiles < sm_count() ? num_tiles : sm_count());  // persistent: one block per SM
    const bool aligned = (W % 4 == 0) && ((uintptr_t)one % 16 == 0) && ((uintptr_t)two % 16 == 0) && ((uintptr_t)out % 16 == 0);
    if (aligned && tiles_y < 65536 && B < 65536) {
        CUtensorMap tm_one, tm_two;
        const uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)C, (uint64_t)B};
        const uint64_t strides[3] = {(uint64_t)W * 4, (uint64_t)H * W * 4, (uint64_t)C * H * W * 4};
        // small levels: half-height tiles (see PwcGeom)
        const bool half_tiles = num_tiles <= (int64_t)sm_count() * 7 / 2 && H > 4;
        const int ty = half_tiles ? 4 : 8;
        const int tiles_y_k = ceil_div(H, ty);
        const int64_t tiles_k = (int64_t)tiles_x * tiles_y_k * B;
        const uint32_t box_two[4] = {PITCH2, (uint32_t)(ty + 2 * PHALO), PCC, 1};
        const uint32_t box_one[4] = {PITCH1, (uint32_t)ty, PCC, 1};
        if (int rc = encode_tensor_map(&tm_two, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, two, dims, strides, box_two,
                                       CU_TENSOR_MAP_SWIZZLE_NONE, "pwc two")) return rc;
        if (int rc = encode_tensor_map(&tm_one, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, one, dims, strides, box_one,
                                       CU_TENSOR_MAP_SWIZZLE_NONE, "pwc one")) return rc;
        // channel split across a cluster only while the grid stays within ONE CTA per SM (two for the half-height
        // tiles, which fit three): the DSMEM reduce-scatter and its two cluster barriers cost more than a few extra
        // channel stages (28x64, C=96: split 4 = 47 us, no split = 128 CTAs x 12 stages), so splitting pays only for
        // the levels that would leave most SMs idle.
        const int nstage_all = ceil_div(C, PCC);
        const int64_t cta_cap = (int64_t)sm_count() * (half_tiles ? 2 : 1);
        int split = 1;
        while (split < 8 && tiles_k * split * 2 <= cta_cap && split * 2 <= nstage_all) split *= 2;
        const bool pow2 = (C & (C - 1)) == 0;
        const float inv_c = 1.0f / (float)C;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(tiles_x * split, tiles_y_k, B);
        cfg.blockDim = dim3(half_tiles ? PwcGeom<4>::THREADS : PwcGeom<8>::THREADS);
        cfg.dynamicSmemBytes = half_tiles ? PwcGeom<4>::SMEM : PwcGeom<8>::SMEM;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = split;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
#define FF_PWC_LAUNCH(P2, TYV)                                                                                                  \
    do {                                                                                                                        \
        FFCORR_CUDA(cudaFuncSetAttribute(pwc81_tma_kernel<P2, TYV>, cudaFuncAttributeMaxDynamicSharedMemorySize,                \
                                         (int)PwcGeom<TYV>::SMEM));                                                             \
        FFCORR_CUDA(cudaLaunchKernelEx(&cfg, pwc81_tma_kernel<P2, TYV>, tm_one, tm_two, out, C, H, W, inv_c, leaky_slope));     \
    } while (0)
        if (half_tiles) { if (pow2) FF_PWC_LAUNCH(true, 4); else FF_PWC_LAUNCH(false, 4); }
        else            { if (pow2) FF_PWC_LAUNCH(true, 8); else FF_PWC_LAUNCH(false, 8); }
#undef FF_PWC_LAUNCH
        return check_launch("pwc81_tma_kernel");
    }
    if (aligned) {
        FFCORR_CUDA(cudaFuncSetAttribute(pwc81_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PWC_SMEM));
        pwc81_kernel<true><<<grid, PWC_THREADS, PWC_SMEM, (cudaStream_t)stream>>>(one, two, out, B, C, H, W, tiles_x, tiles_y, leaky_slope);
    } else {
        FFCORR_CUDA(cudaFuncSetAttribute(pwc81_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PWC_SMEM));
        pwc81_kernel<false><<<grid, PWC_THREADS, PWC_SMEM, (cudaStream_t)stream>>>(one, two, out, B, C, H, W, tiles_x, tiles_y, leaky_slope);
    }
    return check_launch("pwc81_kernel");
}

extern "C" int ffcorr_pwc81_bwd_f32(const float* one, const float* two, const float* grad_out, float* grad_one,
                                    float* grad_two, int B, int C, int H, int W, void* stream) {
    FFCORR_REQUIRE(B >= 0 && C >= 1 && H >= 1 && W >= 1, FFCORR_EINVAL, "pwc81_bwd: bad shape");
    if (B == 0) return FFCORR_OK;
    FFCORR_REQUIRE(one && two && grad_out, FFCORR_EINVAL, "pwc81_bwd: null pointer");
    if (B == 0 || (!grad_one && !grad_two)) return FFCORR_OK;
    const int tiles_x = ceil_div(W, PT_X), tiles_y = ceil_div(H, PT_Y);
    const int cstages = ceil_div(C, BW_CH);
    const bool aligned = (W % 4 == 0) && ((uintptr_t)one % 16 == 0) && ((uintptr_t)two % 16 == 0) &&
                         ((uintptr_t)grad_out % 16 == 0) && (!grad_one || (uintptr_t)grad_one % 16 == 0) &&
                         (!grad_two || (uintptr_t)grad_two % 16 == 0);
    if (aligned && tiles_y < 65536 && B < 65536 && (int64_t)tiles_x * cstages < (1ll << 31)) {
        CUtensorMap tm_one, tm_two, tm_g;
        const uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)C, (uint64_t)B};
        const uint64_t strides[3] = {(uint64_t)W * 4, (uint64_t)H * W * 4, (uint64_t)C * H * W * 4};
        const uint64_t gdims[4] = {(uint64_t)W, (uint64_t)H, 81, (uint64_t)B};
        const uint64_t gstrides[3] = {(uint64_t)W * 4, (uint64_t)H * W * 4, (uint64_t)81 * H * W * 4};
        const uint32_t box_t[4] = {PITCH2, PH2, PCC, 1};
        const uint32_t box_g[4] = {BW_G_PITCH, PT_Y, 1, 1};
        if (int rc = encode_tensor_map(&tm_g, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, grad_out, gdims, gstrides, box_g,
                                       CU_TENSOR_MAP_SWIZZLE_NONE, "pwc grad_out")) return rc;
        const bool pow2 = (C & (C - 1)) == 0;
        const float inv_c = 1.0f / (float)C;
        const dim3 grid((unsigned)(tiles_x * cstages), (unsigned)tiles_y, (unsigned)B);
        cudaStream_t s = (cudaStream_t)stream;
#define FF_PWC_BWD(NEG, P2, TM, OUT)                                                                                       \
    do {                                                                                                                   \
        FFCORR_CUDA(cudaFuncSetAttribute(pwc81_bwd_tma_kernel<NEG, P2>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                         (int)PWC_BWD_SMEM));                                                              \
        pwc81_bwd_tma_kernel<NEG, P2><<<grid, BW_THREADS, PWC_BWD_SMEM, s>>>(TM, tm_g, OUT, C, H, W, cstages, inv_c);      \
        if (int rc = check_launch("pwc81_bwd_tma_kernel")) return rc;                                                      \
    } while (0)
        if (grad_one) {
            if (int rc = encode_tensor_map(&tm_two, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, two, dims, strides, box_t,
                                           CU_TENSOR_MAP_SWIZZLE_NONE, "pwc two")) return rc;
            if (pow2) FF_PWC_BWD(false, true, tm_two, grad_one); else FF_PWC_BWD(false, false, tm_two, grad_one);
        }
        if (grad_two) {
            if (int rc = encode_tensor_map(&tm_one, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, one, dims, strides, box_t,
                                           CU_TENSOR_MAP_SWIZZLE_NONE, "pwc one")) return rc;
            if (pow2) FF_PWC_BWD(true, true, tm_one, grad_two); else FF_PWC_BWD(true, false, tm_one, grad_two);
        }
#undef FF_PWC_BWD
        return FFCORR_OK;
    }
    const int64_t total = (int64_t)B * C * H * W;
    const int64_t want = ceil_div64(total, 256);
    const int64_t cap = (int64_t)sm_count() * 16;
    const int grid = (int)(want < cap ? want : cap);
    pwc81_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(one, two, grad_out, grad_one, grad_two, B, C, H, W);
    return check_launch("pwc81_bwd_kernel");
}
