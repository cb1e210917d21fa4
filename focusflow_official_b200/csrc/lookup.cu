// Fused multi-level bilinear window lookup of the correlation pyramid (sm_100a).
//
// Replaces, per refinement iteration, the reference's 4x F.grid_sample + ~60 glue
// kernels (FF_RAFT_Core/corr.py:29-50, utils/utils.py:57-71) with ONE launch.
//
// HBM-bound gather: per query and level a (2r+2)^2 window of that query's private
// [h_i, w_i] correlation map is read once, (2r+1)^2 bilinear samples are written.
// Algorithmic bytes per query (r = 4, L = 4): 4*100*4 read + 324*4 write + 8 = 2904 B.
//
// Kernels: lookup_kernel (row-major pyramid, described below); on the tiled pyramid lookup_tiled_stream_kernel (NCHW
// output), lookup_tiled_nhwc_kernel (channels-last output, the host model's default), lookup_tiled_nhwc_h_kernel (fp16
// storage) and lookup_convc1_kernel (the lookup fused with its consumer convc1 + ReLU on the tensor cores) -- each
// described at its definition; lookup_bwd_kernel (adjoint w.r.t. the pyramid).
//
// Work decomposition
//   unit   = (level, tile of 32 consecutive queries of one batch item), one warp each;
//   block  = 4 warps = 4 consecutive tiles; blocks are ordered level-major so the
//            expensive level-0 units are scheduled first and the cheap ones fill the tail.
//   phase A (lane = query)  : window origin / extent from the first and last tap
//   phase B (lanes = window): the 32 windows are gathered cooperatively, 32 consecutive
//            window elements per warp load (3 rows of one private map -> 3-6 sectors),
//            zero-filled outside the map (grid_sample padding_mode='zeros'), into smem
//   phase C (lane = query)  : every lane evaluates its own 81 samples from smem and the
//            warp stores 32 consecutive queries of one output channel = one 128 B line.
//
// Numerics: the reference normalises pixel coordinates to [-1, 1] (utils.py:61-62) and
// ATen un-normalises them again (GridSampler.cuh:26).  That fp32 round trip perturbs
// every tap differently (up to ~2e-5 px), which is visible at the 1e-5 parity bar, so
// it is reproduced per tap with explicitly rounded intrinsics (no FMA contraction).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace ffcorr {

namespace {

constexpr int kWarpsPerBlock = 2;
constexpr int kTile = 32;  // queries per warp

struct LookupParams {
    const float* lvl[FFCORR_MAX_LEVELS];
    int lh[FFCORR_MAX_LEVELS];
    int lw[FFCORR_MAX_LEVELS];
    const float* coords;  // [B, 2, N]
    float* out;           // [B, L*K*K, N]
    int B, N, num_levels;
    int tiles_per_batch;
    int blocks_per_batch;
    int64_t cstride;      // floats between consecutive output channels of one query: N (NCHW) or 1 (NHWC)
    int64_t qstride;      // floats between consecutive queries of one channel:       1 (NCHW) or L*K*K (NHWC)
};

// utils.py:61-62 then GridSampler.cuh:26, every op rounded to fp32 like the reference.  The one place where the
// reference's CPU and GPU runs differ is the division by (size - 1) in utils.py:61-62: ATen's CPU kernel divides,
// its CUDA kernel multiplies by the fp32 reciprocal of the scalar (<= 1 ulp of the normalised coordinate, ~1e-5 px at
// w = 156).  CUDA_SEM selects the latter (the `sampler` argument of every lookup entry point).
template <bool CUDA_SEM>
__device__ __forceinline__ float source_index(float x, float size_m1, float rcp_size_m1) {
    const float t = __fmul_rn(2.0f, x);
    const float g = __fsub_rn(CUDA_SEM ? __fmul_rn(t, rcp_size_m1) : __fdiv_rn(t, size_m1), 1.0f);
    return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), size_m1);
}

// |index| beyond this is outside every supported map (h, w <= 16384): all taps are zero.
constexpr float kWildLimit = 3.0e4f;

template <int R, int QU, bool CUDA_SEM>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) lookup_kernel(const LookupParams p) {
    constexpr int K = 2 * R + 1;
    constexpr int W2 = K + 2;        // window extent incl. the +-1 floor deviation of the round trip
    constexpr int WIN = W2 * W2;     // odd -> lane-per-query smem reads are conflict-free
    constexpr int NLOAD = (WIN + 31) / 32;
    static_assert(WIN % 2 == 1, "window stride must be odd");
    static_assert(W2 <= 15, "row/column masks are 16 bits");

    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    float* swin = smem + warp * (kTile * WIN);

    // block -> (level, batch, tiles); level-major so level 0 goes first
    int bid = blockIdx.x;
    const int per_level = p.B * p.blocks_per_batch;
    const int level = bid / per_level;
    bid -= level * per_level;
    const int b = bid / p.blocks_per_batch;
    const int tile = (bid - b * p.blocks_per_batch) * kWarpsPerBlock + warp;
    if (tile >= p.tiles_per_batch) return;  // warp-uniform; no block-level sync below

    const int N = p.N;
    const int n0 = tile * kTile;
    const int n = n0 + lane;
    const bool valid = n < N;
    const int lh = p.lh[level], lw = p.lw[level];
    const float* __restrict__ lvl = p.lvl[level];
    const float inv_scale = __int_as_float((127 - level) << 23);  // 2^-level, exact (corr.py:40)

    // ---------------- phase A: per-query window origin and in-bounds masks ----------------
    float cx = 0.f, cy = 0.f;
    if (valid) {
        const float* c = p.coords + (size_t)b * 2 * N + n;
        cx = __ldg(c) * inv_scale;
        cy = __ldg(c + N) * inv_scale;
    }
    const float sx = (float)(lw - 1), sy = (float)(lh - 1);
    const float rsx = __frcp_rn(sx), rsy = __frcp_rn(sy);     // ATen: 1.0 / scalar, in fp32
    const float ixf = source_index<CUDA_SEM>(__fadd_rn(cx, (float)(-R)), sx, rsx);
    const float ixl = source_index<CUDA_SEM>(__fadd_rn(cx, (float)(R)), sx, rsx);
    const float iyf = source_index<CUDA_SEM>(__fadd_rn(cy, (float)(-R)), sy, rsy);
    const float iyl = source_index<CUDA_SEM>(__fadd_rn(cy, (float)(R)), sy, rsy);
    const bool wild = !(fabsf(ixf) < kWildLimit) || !(fabsf(ixl) < kWildLimit) ||
                      !(fabsf(iyf) < kWildLimit) || !(fabsf(iyl) < kWildLimit);
    const int map_elems = lh * lw;                // < 2^24 (checked on the host)
    int x_lo = 0, y_lo = 0;
    int my_mask = 0;                              // bits [0,16): valid window rows, [16,32): valid columns
    int my_qoff = 0;                              // element offset of the window origin from the tile's first map
    if (valid && !wild) {
        x_lo = (int)floorf(ixf);
        y_lo = (int)floorf(iyf);
        const int ncols = min((int)floorf(ixl) + 2 - x_lo, W2);
        const int nrows = min((int)floorf(iyl) + 2 - y_lo, W2);
        const int rlo = max(0, -y_lo), rhi = min(nrows, lh - y_lo);
        const int clo = max(0, -x_lo), chi = min(ncols, lw - x_lo);
        const int rm = (rhi > rlo) ? (((1 << rhi) - 1) & ~((1 << rlo) - 1)) : 0;
        const int cm = (chi > clo) ? (((1 << chi) - 1) & ~((1 << clo) - 1)) : 0;
        my_mask = (rm && cm) ? (rm | (cm << 16)) : 0;
        my_qoff = lane * map_elems + y_lo * lw + x_lo;   // |.| < 2^31: 31*2^24 + 3e4*2^14
    }

    const float* __restrict__ tile_base = lvl + ((int64_t)b * N + n0) * (int64_t)map_elems;

    // ---------------- phase B: cooperative gather into smem ----------------
    // element e = lane + 32 j of a window -> (row, col) = (e / W2, e % W2); its in-bounds test is
    // one AND against the query's mask, its address one IMAD.WIDE from a per-lane base pointer.
    const float* pj[NLOAD];
    int bits[NLOAD];
#pragma unroll
    for (int j = 0; j < NLOAD; ++j) {
        const int e = lane + 32 * j;
        const int er = e / W2, ec = e - er * W2;
        pj[j] = tile_base + (er * lw + ec);
        bits[j] = (e < WIN) ? ((1 << er) | (1 << (16 + ec))) : 0x80008000;  // never matches
    }
    float* sdst = swin + lane;
#pragma unroll 1
    for (int q0 = 0; q0 < kTile; q0 += QU) {
        float v[QU][NLOAD];
#pragma unroll
        for (int u = 0; u < QU; ++u) {
            const int qoff = __shfl_sync(0xffffffffu, my_qoff, q0 + u);
            const int msk = __shfl_sync(0xffffffffu, my_mask, q0 + u);
#pragma unroll
            for (int j = 0; j < NLOAD; ++j) {
                const bool ok = (msk & bits[j]) == bits[j];
                v[u][j] = ok ? __ldg(pj[j] + qoff) : 0.0f;
            }
        }
#pragma unroll
        for (int u = 0; u < QU; ++u) {
#pragma unroll
            for (int j = 0; j < NLOAD; ++j) {
                if (32 * j + 31 < WIN || lane + 32 * j < WIN) sdst[u * WIN + 32 * j] = v[u][j];
            }
        }
        sdst += QU * WIN;
    }
    __syncwarp();

    // ---------------- phase C: lane-per-query evaluation ----------------
    int rx[K], ry[K];
    float wx0[K], wx1[K], wy0[K], wy1[K];
    bool deviated = false;
#pragma unroll
    for (int a = 0; a < K; ++a) {
        const float ix = source_index<CUDA_SEM>(__fadd_rn(cx, (float)(a - R)), sx, rsx);
        const float fx = floorf(ix);
        const float iy = source_index<CUDA_SEM>(__fadd_rn(cy, (float)(a - R)), sy, rsy);
        const float fy = floorf(iy);
        // corner distances of the ATen CUDA kernel: (ix_se - ix), (ix - ix_nw) with ix_se = ix_nw + 1
        wx1[a] = wild ? 0.f : __fsub_rn(ix, fx);
        wx0[a] = wild ? 0.f : __fsub_rn(__fadd_rn(fx, 1.0f), ix);
        wy1[a] = wild ? 0.f : __fsub_rn(iy, fy);
        wy0[a] = wild ? 0.f : __fsub_rn(__fadd_rn(fy, 1.0f), iy);
        rx[a] = wild ? a : min(max((int)fx - x_lo, 0), W2 - 2);
        ry[a] = wild ? a : min(max((int)fy - y_lo, 0), W2 - 2);
        deviated |= (rx[a] != a) | (ry[a] != a);
    }
    if (!valid) deviated = false;  // padding lanes of the last tile must not force the slow path

    const float* sq = swin + lane * WIN;
    const int CT = p.num_levels * K * K;
    // NCHW: channel c of query n at ((b*CT + c)*N + n); NHWC: at ((b*N + n)*CT + c)
    float* __restrict__ op = p.out + (int64_t)b * CT * N + (int64_t)level * K * K * p.cstride + (int64_t)n * p.qstride;
    const int64_t cs = p.cstride;

    if (!__any_sync(0xffffffffu, deviated)) {
        // Fast path (no tap of any query in the warp changed its floor through the round trip):
        // tap (a, b) sits at window (b, a), so every smem offset is a compile-time constant and the
        // bilinear form factors into a horizontal pass shared by the two outputs that use a row.
        // out = wy0*(wx0*v00 + wx1*v01) + wy1*(wx0*v10 + wx1*v11): same value as the reference's
        // nw/ne/sw/se form up to fp32 rounding (~1e-7 relative).
        float tprev[K];
#pragma unroll
        for (int r = 0; r <= K; ++r) {
            float vrow[K + 1];
#pragma unroll
            for (int c = 0; c <= K; ++c) vrow[c] = sq[r * W2 + c];
            float tcur[K];
#pragma unroll
            for (int a = 0; a < K; ++a) tcur[a] = __fmaf_rn(wx1[a], vrow[a + 1], __fmul_rn(wx0[a], vrow[a]));
            if (r > 0) {
                const int bb = r - 1;
#pragma unroll
                for (int a = 0; a < K; ++a) {
                    const float o = __fmaf_rn(wy1[bb], tcur[a], __fmul_rn(wy0[bb], tprev[a]));
                    if (valid) op[(int64_t)(a * K + bb) * cs] = o;
                }
            }
#pragma unroll
            for (int a = 0; a < K; ++a) tprev[a] = tcur[a];
        }
    } else {
        // Exact general path (integer / near-integer coordinates, e.g. the first iteration): per-tap
        // window indices, corners weighted and accumulated in ATen's order nw, ne, sw, se.
#pragma unroll
        for (int a = 0; a < K; ++a) {
#pragma unroll
            for (int bb = 0; bb < K; ++bb) {
                const float* s = sq + ry[bb] * W2 + rx[a];
                const float v00 = s[0], v01 = s[1], v10 = s[W2], v11 = s[W2 + 1];
                const float nw = __fmul_rn(wx0[a], wy0[bb]);
                const float ne = __fmul_rn(wx1[a], wy0[bb]);
                const float sw = __fmul_rn(wx0[a], wy1[bb]);
                const float se = __fmul_rn(wx1[a], wy1[bb]);
                float o = __fmul_rn(v00, nw);
                o = __fmaf_rn(v01, ne, o);
                o = __fmaf_rn(v10, sw, o);
                o = __fmaf_rn(v11, se, o);
                if (valid) op[(int64_t)(a * K + bb) * cs] = o;
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// Tiled ("T4") pyramid variant.  Levels are stored per query map as 4x4-pixel tiles of 64 contiguous
// bytes with exact-zero padding (pyramid.cu), so zero padding of the sampler needs no per-element bounds
// test -- only "does this tile exist" -- and a window costs ~10.6 64-byte requests instead of ~17.
// (A first version staged whole 11x11 windows like lookup_kernel: 15.5 KB of shared memory per warp, 0.081 ms at
// config 2.  The row-streaming kernel below replaced it.)
// ---------------------------------------------------------------------------------
struct LookupTiledParams {
    const float* lvl[FFCORR_MAX_LEVELS];
    int lh[FFCORR_MAX_LEVELS], lw[FFCORR_MAX_LEVELS];     // true level sizes (for the coordinate round trip)
    int th[FFCORR_MAX_LEVELS], tw[FFCORR_MAX_LEVELS];     // tiles per column / row
    const float* coords;
    float* out;
    int B, N, num_levels;
    int tiles_per_batch;
    int blocks_per_batch;
    int64_t coords_stride;   // floats between the x and y planes of coords (N, or the full map size for a query chunk)
    int64_t out_stride;      // floats between output channels (N, or the full map size for a query chunk)
    // Storage order of the tiles.  grouped == 0 ("T4"): every query owns a contiguous map [th][tw][16].  grouped == 1
    // ("G32", ffcorr_build_grouped_f32): 32 consecutive queries share one tile grid, [group][th*tw][32 queries][16] -- the
    // same tile of neighbouring queries (whose windows overlap under any smooth flow) is contiguous in memory.
    int grouped;
    int NG;                  // groups of 32 queries per batch item (grouped only)
};

// ---------------------------------------------------------------------------------
// Row-streaming tiled lookup: instead of staging whole 11x11 windows (15.5 KB of shared
// memory per warp -> 14 warps per SM), the warp streams the 32 windows ROW BY ROW through a small ring of
// row buffers filled by cp.async (LDGSTS: global -> shared without a register round trip, zero-fill for
// tiles outside the map).  Gather role: lane = (query, tile column), one 16-byte piece of a tile row per
// slot.  Evaluate role: lane = query; it folds window row r into its 9 running horizontal interpolants and
// emits the 9 outputs whose lower row it completes.  kStreamStages - 1 rows are in flight per warp while
// one is evaluated, so the memory system sees ~2 KB x (stages-1) x resident warps of outstanding requests.
// ---------------------------------------------------------------------------------
constexpr int kStreamWarps = 4;
#ifndef FFCORR_STREAM_STAGES
#define FFCORR_STREAM_STAGES 3
#endif
constexpr int kStreamStages = FFCORR_STREAM_STAGES;
constexpr int kRowPitch = 20;     // floats per query row: the 16 columns of the 4x4 tile block + 4 (16-byte aligned)

// FFCORR_CP_L2 (development switch): L2 prefetch size hint of the gather's cp.async (0 = none, 64, 128, 256)
#ifndef FFCORR_CP_L2
#define FFCORR_CP_L2 0
#endif
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
#if FFCORR_CP_L2 == 64
    asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
#elif FFCORR_CP_L2 == 128
    asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
#elif FFCORR_CP_L2 == 256
    asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
}

template <int R, bool CUDA_SEM>
__global__ void __launch_bounds__(kStreamWarps * 32) lookup_tiled_stream_kernel(const LookupTiledParams p) {
    constexpr int K = 2 * R + 1;
    constexpr int W2 = K + 2;
    constexpr int S = kStreamStages;
    static_assert(S >= 2, "need at least a double buffer");
    static_assert(W2 + 3 <= 16, "window + sub-tile shift must fit in the 16 columns of the tile block");
    constexpr int ROWBUF = kTile * kRowPitch;

    __shared__ __align__(16) float sring[kStreamWarps][S][ROWBUF];
    __shared__ float s_iy[kStreamWarps][K][kTile];      // y source index of tap bb, per query
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    float* ring = &sring[warp][0][0];
    float* siy = &s_iy[warp][0][0];

    int bid = blockIdx.x;
    const int per_level = p.B * p.blocks_per_batch;
    const int level = bid / per_level;
    bid -= level * per_level;
    const int b = bid / p.blocks_per_batch;
    const int tile = (bid - b * p.blocks_per_batch) * kStreamWarps + warp;
    if (tile >= p.tiles_per_batch) return;

    const int N = p.N;
    const int n0 = tile * kTile;
    // Lanes past the end of the last tile recompute the LAST query (identical values, identical address):
    // benign duplicate stores instead of a per-store predicate in the hot loop.
    const int n = min(n0 + lane, N - 1);
    const int lh = p.lh[level], lw = p.lw[level];
    const int th = p.th[level], tw = p.tw[level];
    const int twr = (lw + 3) >> 2;      // tile columns that hold pixels; the column that evens the pitch is never read
    const int map_elems = th * tw * 16;
    const int TS = p.grouped ? 16 * kTile : 16;                  // floats between horizontally adjacent tiles
    const int QS = p.grouped ? 16 : map_elems;                   // floats between the same tile of consecutive queries
    const float* __restrict__ lvl = p.lvl[level];
    const float inv_scale = __int_as_float((127 - level) << 23);

    // ---------------- phase A (lane = query) ----------------
    const float* cptr = p.coords + (size_t)b * 2 * p.coords_stride + n;
    const float cx = __ldg(cptr) * inv_scale;
    const float cy = __ldg(cptr + p.coords_stride) * inv_scale;
    const float sx = (float)(lw - 1), sy = (float)(lh - 1);
    const float rsx = __frcp_rn(sx), rsy = __frcp_rn(sy);     // ATen: 1.0 / scalar, in fp32
    const float ixf = source_index<CUDA_SEM>(__fadd_rn(cx, (float)(-R)), sx, rsx);
    const float ixl = source_index<CUDA_SEM>(__fadd_rn(cx, (float)(R)), sx, rsx);
    const float iyf = source_index<CUDA_SEM>(__fadd_rn(cy, (float)(-R)), sy, rsy);
    const float iyl = source_index<CUDA_SEM>(__fadd_rn(cy, (float)(R)), sy, rsy);
    const bool wild = !(fabsf(ixf) < kWildLimit) || !(fabsf(ixl) < kWildLimit) ||
                      !(fabsf(iyf) < kWildLimit) || !(fabsf(iyl) < kWildLimit);
    int x_lo = 0, y_lo = 0;
    int my_base = 0;    // float offset of the window's first tile inside the query map: (ty0*tw + tx0)*16
    int my_pack = 0;    // bits [0,2) x_lo&3, [2,4) y_lo&3, bit 4+k: tile k = tyi*4+txi of the 4x4 block exists
    if (!wild) {
        x_lo = (int)floorf(ixf);
        y_lo = (int)floorf(iyf);
        const int tx0 = x_lo >> 2, ty0 = y_lo >> 2;
        my_base = (ty0 * tw + tx0) * TS;
        int tmask = 0;
#pragma unroll
        for (int tyi = 0; tyi < 4; ++tyi)
#pragma unroll
            for (int txi = 0; txi < 4; ++txi)
                if ((unsigned)(ty0 + tyi) < (unsigned)th && (unsigned)(tx0 + txi) < (unsigned)twr) tmask |= 1 << (tyi * 4 + txi);
        my_pack = (x_lo & 3) | ((y_lo & 3) << 2) | (tmask << 4);
    }

    // x-tap weights stay in registers; the y source indices go to shared memory and the y-tap weights are
    // rebuilt from them per row step.  Floor deviations of the round trip (-1/0/+1 per tap) are packed
    // 2 bits per tap: px for columns, py for rows.
    float wx0[K], wx1[K];
    unsigned px = 0, py = 0;
    bool deviated = false;
#pragma unroll
    for (int a = 0; a < K; ++a) {
        const float ix = source_index<CUDA_SEM>(__fadd_rn(cx, (float)(a - R)), sx, rsx);
        const float fx = floorf(ix);
        const float iy = source_index<CUDA_SEM>(__fadd_rn(cy, (float)(a - R)), sy, rsy);
        const float fy = floorf(iy);
        siy[a * kTile + lane] = iy;
        wx1[a] = wild ? 0.f : __fsub_rn(ix, fx);
        wx0[a] = wild ? 0.f : __fsub_rn(__fadd_rn(fx, 1.0f), ix);
        const int dxa = wild ? 0 : min(max((int)fx - x_lo, 0), W2 - 2) - a;   // in {-1, 0, 1}
        const int dya = wild ? 0 : min(max((int)fy - y_lo, 0), W2 - 2) - a;
        px |= (unsigned)((dxa + 1) & 3) << (2 * a);
        py |= (unsigned)((dya + 1) & 3) << (2 * a);
        deviated |= (dxa != 0) | (dya != 0);
    }
    const bool slow = __any_sync(0xffffffffu, deviated);
    auto ytap = [&](int bb, float& w0, float& w1) {      // only the lane's own entries: no sync needed
        const float iy = siy[bb * kTile + lane];
        const float fy = floorf(iy);
        w1 = wild ? 0.f : __fsub_rn(iy, fy);
        w0 = wild ? 0.f : __fsub_rn(__fadd_rn(fy, 1.0f), iy);
    };

    // Row buffer layout: query q owns kRowPitch floats at slot(q) * kRowPitch.  The tile block is stored
    // unshifted (16-byte cp.async destinations); the evaluate role reads its 10 columns starting at x_lo & 3,
    // so two queries collide on a bank when their slots are equal mod 8 AND their x_lo & 3 agree.  For a
    // locally smooth flow x_lo advances by one every 2^level queries: slot = m*8 + g with m = (q >> level) & 3
    // and g = the other three bits of q puts exactly the four queries with distinct x_lo & 3 on one phase.
    const int lsh = min(level, 3);
    auto slot_of = [lsh](int q) {
        const int m = (q >> lsh) & 3;
        const int g = (q & ((1 << lsh) - 1)) | ((q >> (lsh + 2)) << lsh);
        return (m << 3) | g;
    };

    // ---------------- gather slots (lane = (query qj, tile column txi), j = 0..3) ----------------
    const int txi = lane & 3;
    // n0 is a multiple of 32: the warp's queries are exactly one group of the grouped layout
    const float* __restrict__ tile_base = p.grouped ? lvl + ((int64_t)b * p.NG + (n0 >> 5)) * (int64_t)map_elems * kTile
                                                    : lvl + ((int64_t)b * N + n0) * (int64_t)map_elems;
    int goff[4];        // float offset from tile_base of the slot's 16-byte piece in the current window row
    unsigned gmask[4];  // bit r: the tile under window row r exists; bit 16+r: row r is the last row of its tile
    unsigned gdst[4];   // byte offset of the slot inside a row buffer
    const int last_q = N - 1 - n0;              // last real query of this tile (>= 0)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int qj = (lane >> 2) + 8 * j;
        const int base = __shfl_sync(0xffffffffu, my_base, qj);
        const int pk = __shfl_sync(0xffffffffu, my_pack, qj);
        const int gy = (pk >> 2) & 3;
        goff[j] = min(qj, last_q) * QS + base + txi * TS + gy * 4;    // < 2^31 (host check: 32 maps)
        const unsigned e = (unsigned)pk >> (4 + txi);
        unsigned okT = ((e & 1u) * 0xFu) | (((e >> 4) & 1u) * 0xF0u) | (((e >> 8) & 1u) * 0xF00u) |
                       (((e >> 12) & 1u) * 0xF000u);
#ifndef FFCORR_NO_COLUMN_MASK
        // only the tile columns the window's columns [x_lo & 3, (x_lo & 3) + ncols) reach: the fourth one is needed by a
        // quarter of the windows only (fetching all four cost 23 % more DRAM sectors)
        if (txi * 4 >= (pk & 3) + (slow ? W2 : K + 1)) okT = 0;
#endif
        gmask[j] = ((okT >> gy) & 0xFFFFu) | ((0x8888u >> gy) << 16);
        gdst[j] = (unsigned)(slot_of(qj) * kRowPitch + 4 * txi) * 4u;
    }
    const int next_tile_row = tw * TS - 12;     // from in-tile row 3 to row 0 of the tile below
    const uint32_t ring_saddr = (uint32_t)__cvta_generic_to_shared(ring);
    int rows_issued = 0;
    uint32_t issue_saddr = ring_saddr;          // stage the next issued row lands in
    // issues window row `rows_issued` (rows go out in order: the offsets advance); always commits a group
    auto issue_row = [&](bool active) {
        if (active) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned m = gmask[j] >> rows_issued;
                const bool ok = m & 1u;
                const float* src = tile_base + (ok ? goff[j] : 0);
                cp_async16_zfill(issue_saddr + gdst[j], src, ok ? 16u : 0u);
                goff[j] += (m & 0x10000u) ? next_tile_row : 4;
            }
            ++rows_issued;
            issue_saddr += ROWBUF * 4;
            if (issue_saddr == ring_saddr + S * ROWBUF * 4) issue_saddr = ring_saddr;
        }
        cp_async_commit();
    };

    const int CT = p.num_levels * K * K;
    const int64_t ostride = p.out_stride;
    float* __restrict__ op = p.out + ((int64_t)b * CT + (int64_t)level * K * K) * ostride + n;
    const int64_t stride_a = (int64_t)K * ostride;
    const int my_row = slot_of(lane) * kRowPitch + (x_lo & 3);

#pragma unroll
    for (int r = 0; r < S - 1; ++r) issue_row(true);       // W2 > S - 1 for every radius

    if (!slow) {
        // fast path: tap (a, b) reads window (b, a); row r feeds outputs bb = r - 1 (as its lower row)
        float tprev[K];
#pragma unroll
        for (int a = 0; a < K; ++a) tprev[a] = 0.f;
        float* o_row = op;
        const float* sq = ring + my_row;
#pragma unroll 1
        for (int r = 0; r <= K; ++r) {
            cp_async_wait<S - 2>();                         // row r has landed (this lane's pieces) ...
            __syncwarp();                                   // ... and everyone's; row r-1 is fully consumed
            issue_row(r + S - 1 <= K);                      // refill the stage row r-1 occupied
            float vrow[K + 1];
#pragma unroll
            for (int c = 0; c <= K; ++c) vrow[c] = sq[c];
            float tcur[K];
#pragma unroll
            for (int a = 0; a < K; ++a) tcur[a] = __fmaf_rn(wx1[a], vrow[a + 1], __fmul_rn(wx0[a], vrow[a]));
            if (r > 0) {
                float w0, w1;
                ytap(r - 1, w0, w1);
                float* o_ptr = o_row;                       // channel a*K + (r-1); next a is K*N floats further
#pragma unroll
                for (int a = 0; a < K; ++a) {
                    *o_ptr = __fmaf_rn(w1, tcur[a], __fmul_rn(w0, tprev[a]));
                    o_ptr += stride_a;
                }
                o_row += ostride;
            }
#pragma unroll
            for (int a = 0; a < K; ++a) tprev[a] = tcur[a];
            sq += ROWBUF;
            if (sq == ring + my_row + S * ROWBUF) sq = ring + my_row;
        }
    } else {
        // exact general path (integer / near-integer coordinates): row r completes the outputs whose lower
        // corner row is r, i.e. ry[bb] + 1 == r; rows r-1 and r are both in the ring, so the refill of row
        // r-1's stage waits until row r has been evaluated.  ATen's nw/ne/sw/se order.
        const float* cur = ring + my_row;
        const float* prev = cur;                            // unused at r = 0
#pragma unroll 1
        for (int r = 0; r < W2; ++r) {
            cp_async_wait<S - 2>();
            __syncwarp();
#pragma unroll 1
            for (int bb = 0; bb < K; ++bb) {
                const int ryb = bb + (int)((py >> (2 * bb)) & 3u) - 1;
                if (ryb + 1 == r) {
                    float wy0b, wy1b;
                    ytap(bb, wy0b, wy1b);
#pragma unroll
                    for (int a = 0; a < K; ++a) {
                        const int rxa = a + (int)((px >> (2 * a)) & 3u) - 1;
                        const float v00 = prev[rxa], v01 = prev[rxa + 1], v10 = cur[rxa], v11 = cur[rxa + 1];
                        const float nw = __fmul_rn(wx0[a], wy0b);
                        const float ne = __fmul_rn(wx1[a], wy0b);
                        const float sw = __fmul_rn(wx0[a], wy1b);
                        const float se = __fmul_rn(wx1[a], wy1b);
                        float o = __fmul_rn(v00, nw);
                        o = __fmaf_rn(v01, ne, o);
                        o = __fmaf_rn(v10, sw, o);
                        o = __fmaf_rn(v11, se, o);
                        op[(int64_t)(a * K + bb) * ostride] = o;
                    }
                }
            }
            __syncwarp();                                   // row r-1 no longer needed by anyone
            issue_row(r + S - 1 < W2);
            prev = cur;
            cur += ROWBUF;
            if (cur == ring + my_row + S * ROWBUF) cur = ring + my_row;
        }
    }
}


// ---------------------------------------------------------------------------------
// NHWC ("channels last") row-streaming tiled lookup: out[b, n, lvl*K*K + a*K + bb] -- the layout the consumer of
// the lookup, the 1x1 convolution convc1 (update.py:82-83,90), runs in on tensor cores; the NCHW kernel above forces
// a transposing copy of the whole result (76 MB at config 2) per refinement iteration on the host side.
//
// Unit of work = 8 consecutive queries x ALL (<= 4) levels per warp: window wi = lane = level * 8 + query.  The
// warp's 8 x L*K*K results are one CONTIGUOUS piece of the output (10 368 bytes for r = 4, L = 4): every lane drops
// its samples into a shared-memory staging block (conflict-free: (4 q + 17 lvl) mod 32 are 32 distinct banks) and one
// elected lane hands the block to the TMA engine (cp.async.bulk shared -> global), so the SM spends no issue slots on
// the 76 MB of stores.  Mixing the levels in a warp also makes every warp cost the same (the level-major kernel has
// cheap level-3 blocks in its tail).  Gather role: slot j of lane (query = lane >> 2, tile column = lane & 3) serves
// level j, so all level constants of a slot are warp-uniform.  Evaluate role: the row buffer is read with four
// 16-byte loads per lane (8 consecutive lanes hit 8 distinct 16-byte bank groups with the 80-byte pitch, whatever
// the flow field) and shifted by x_lo & 3 in registers.
// ---------------------------------------------------------------------------------
#ifndef FFCORR_NHWC_WARPS
#define FFCORR_NHWC_WARPS 2
#endif
constexpr int kNhwcWarps = FFCORR_NHWC_WARPS;
constexpr int kNhwcQ = 8;          // queries per warp

// Storage type T of the pyramid: float here; the fp16-stored pyramid (ffcorr_build_tiled_f16) has its own row-pair
// kernel below.  Row-buffer pitch per window: 80 bytes (20 floats), which puts 8 consecutive lanes on 8 distinct 16-byte
// bank groups.
template <typename T> struct RowGeom;
template <> struct RowGeom<float> { static constexpr int PITCH = 20; };

template <typename T>
__host__ __device__ constexpr int nhwc_warp_bytes(int K, int CT) {
    return kStreamStages * kTile * RowGeom<T>::PITCH * (int)sizeof(T) + (K * kTile + kNhwcQ * CT) * 4;
}

// Samples of one window, as the unit function below produces them: row() hands over the K samples (a = 0 .. K-1) of one
// window row bb (fast path), tap() a single sample (exact per-tap path).  Output channel of (a, bb) is a*K + bb (corr.py:37-43).
struct StageSink {                 // the fp32 [8 queries][CT] staging block of lookup_tiled_nhwc_kernel
    float* sout;                   // this window's K*K samples
    template <int K>
    __device__ __forceinline__ void row(int bb, const float (&v)[K]) {
#pragma unroll
        for (int a = 0; a < K; ++a) sout[a * K + bb] = v[a];
    }
    template <int K>
    __device__ __forceinline__ void tap(int a, int bb, float v) { sout[a * K + bb] = v; }
};
struct StageSinks {
    float* stage;
    int CT, KK;
    __device__ __forceinline__ StageSink make(int q, int level) const { return StageSink{stage + q * CT + level * KK}; }
};

// One unit of work of the channels-last lookups: 8 consecutive queries [n0, n0 + 8) of batch item b x all levels, by one
// warp (lane = level * 8 + query), streamed row by row through the warp's cp.async ring.  `ring` = S row buffers
// of 32 windows x PITCH floats, `siy` = K x 32 floats of scratch.  All cp.async groups are drained on return.
template <int R, bool CUDA_SEM, int S = kStreamStages, typename Sinks>
__device__ __forceinline__ void nhwc_lookup_unit(const LookupTiledParams& p, const int b, const int n0, float* ring, float* siy,
                                                 const int lane, const Sinks& sinks) {
    using T = float;
    constexpr bool HALF = false;
    constexpr int PITCH = RowGeom<T>::PITCH;             // elements per window row buffer
    constexpr int K = 2 * R + 1;
    constexpr int W2 = K + 2;
    constexpr int ROWBUF = kTile * PITCH;                // elements per ring stage
    static_assert(W2 + 3 <= 16, "window + sub-tile shift must fit in the 16 columns of the tile block");

    const int N = p.N;
    const int lvl_of_lane = lane >> 3;
    const bool lvl_on = lvl_of_lane < p.num_levels;
    const int level = lvl_on ? lvl_of_lane : 0;
    const int q = lane & 7;
    const int n = min(n0 + q, N - 1);                    // tail lanes recompute the last query (never stored)
    const int lh = p.lh[level], lw = p.lw[level];
    const int th = p.th[level], tw = p.tw[level];
    const int twr = (lw + 3) >> 2;
    const float inv_scale = __int_as_float((127 - level) << 23);

    // ---------------- phase A (lane = window) ----------------
    const float* cptr = p.coords + (size_t)b * 2 * p.coords_stride + n;
    const float cx = __ldg(cptr) * inv_scale;
    const float cy = __ldg(cptr + p.coords_stride) * inv_scale;
    const float sx = (float)(lw - 1), sy = (float)(lh - 1);
    const float rsx = __frcp_rn(sx), rsy = __frcp_rn(sy);
    const float ixf = source_index<CUDA_SEM>(__fadd_rn(cx, (float)(-R)), sx, rsx);
    const float ixl = source_index<CUDA_SEM>(__fadd_rn(cx, (float)(R)), sx, rsx);
    const float iyf = source_index<CUDA_SEM>(__fadd_rn(cy, (float)(-R)), sy, rsy);
    const float iyl = source_index<CUDA_SEM>(__fadd_rn(cy, (float)(R)), sy, rsy);
    const bool wild = !lvl_on || !(fabsf(ixf) < kWildLimit) || !(fabsf(ixl) < kWildLimit) ||
                      !(fabsf(iyf) < kWildLimit) || !(fabsf(iyl) < kWildLimit);
    int x_lo = 0, y_lo = 0;
    int my_base = 0, my_pack = 0;
    if (!wild) {
        x_lo = (int)floorf(ixf);
        y_lo = (int)floorf(iyf);
        const int tx0 = x_lo >> 2, ty0 = y_lo >> 2;
        my_base = (ty0 * tw + tx0) * (p.grouped ? 16 * kTile : 16);
        int tmask = 0;
#pragma unroll
        for (int tyi = 0; tyi < 4; ++tyi)
#pragma unroll
            for (int txi = 0; txi < 4; ++txi)
                if ((unsigned)(ty0 + tyi) < (unsigned)th && (unsigned)(tx0 + txi) < (unsigned)twr) tmask |= 1 << (tyi * 4 + txi);
        my_pack = (x_lo & 3) | ((y_lo & 3) << 2) | (tmask << 4);
    }

    float wx0[K], wx1[K];
    unsigned px = 0, py = 0;
    bool deviated = false;
#pragma unroll
    for (int a = 0; a < K; ++a) {
        const float ix = source_index<CUDA_SEM>(__fadd_rn(cx, (float)(a - R)), sx, rsx);
        const float fx = floorf(ix);
        const float iy = source_index<CUDA_SEM>(__fadd_rn(cy, (float)(a - R)), sy, rsy);
        const float fy = floorf(iy);
        siy[a * kTile + lane] = iy;
        wx1[a] = wild ? 0.f : __fsub_rn(ix, fx);
        wx0[a] = wild ? 0.f : __fsub_rn(__fadd_rn(fx, 1.0f), ix);
        const int dxa = wild ? 0 : min(max((int)fx - x_lo, 0), W2 - 2) - a;
        const int dya = wild ? 0 : min(max((int)fy - y_lo, 0), W2 - 2) - a;
        px |= (unsigned)((dxa + 1) & 3) << (2 * a);
        py |= (unsigned)((dya + 1) & 3) << (2 * a);
        deviated |= (dxa != 0) | (dya != 0);
    }
    const bool slow = __any_sync(0xffffffffu, deviated);
    auto ytap = [&](int bb, float& w0, float& w1) {
        const float iy = siy[bb * kTile + lane];
        const float fy = floorf(iy);
        w1 = wild ? 0.f : __fsub_rn(iy, fy);
        w0 = wild ? 0.f : __fsub_rn(__fadd_rn(fy, 1.0f), iy);
    };

    // ---------------- gather slots: slot j = level j, lane = (query lane >> 2, tile column lane & 3) ----------------
    const int txi = lane & 3;
    const int qj = lane >> 2;
    const int last_q = N - 1 - n0;
    const T* tbase[4];
    int goff[4], gnext[4];
    unsigned gmask[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int lj = j < p.num_levels ? j : 0;
        const int map_elems = p.th[lj] * p.tw[lj] * 16;
        const int TS = p.grouped ? 16 * kTile : 16, QS = p.grouped ? 16 : map_elems;
        // grouped: the warp's 8 queries sit inside group n0 >> 5 at slots (n0 & 31) ..
        tbase[j] = p.grouped ? reinterpret_cast<const T*>(p.lvl[lj]) + ((int64_t)b * p.NG + (n0 >> 5)) * (int64_t)map_elems * kTile + (n0 & 31) * 16
                             : reinterpret_cast<const T*>(p.lvl[lj]) + ((int64_t)b * N + n0) * (int64_t)map_elems;
        const int base = __shfl_sync(0xffffffffu, my_base, j * 8 + qj);
        const int pk = __shfl_sync(0xffffffffu, my_pack, j * 8 + qj);       // 0 for a wild / absent window
        const int gy = (pk >> 2) & 3;
        goff[j] = min(qj, last_q) * QS + base + txi * TS + gy * 4;
        const unsigned e = (unsigned)pk >> (4 + txi);
        unsigned okT = ((e & 1u) * 0xFu) | (((e >> 4) & 1u) * 0xF0u) | (((e >> 8) & 1u) * 0xF00u) |
                       (((e >> 12) & 1u) * 0xF000u);
#ifndef FFCORR_NO_COLUMN_MASK
        if (txi * 4 >= (pk & 3) + (slow ? W2 : K + 1)) okT = 0;       // only the tile columns the window reaches
#endif
        gmask[j] = ((okT >> gy) & 0xFFFFu) | ((0x8888u >> gy) << 16);
        gnext[j] = p.tw[lj] * TS - 12;
    }
    const unsigned gdst = (unsigned)(qj * PITCH + 4 * txi) * (unsigned)sizeof(T);     // + j * 8 * PITCH * sizeof(T) per slot
    const uint32_t ring_saddr = (uint32_t)__cvta_generic_to_shared(ring);
    int rows_issued = 0;
    uint32_t issue_saddr = ring_saddr;
    auto issue_row = [&](bool active) {
        if (active) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned m = gmask[j] >> rows_issued;
                const bool ok = m & 1u;
                const T* src = tbase[j] + (ok ? goff[j] : 0);
                cp_async16_zfill(issue_saddr + gdst + (unsigned)(j * 8 * PITCH * sizeof(T)), src, ok ? 16u : 0u);
                goff[j] += (m & 0x10000u) ? gnext[j] : 4;
            }
            ++rows_issued;
            issue_saddr += ROWBUF * sizeof(T);
            if (issue_saddr == ring_saddr + S * ROWBUF * sizeof(T)) issue_saddr = ring_saddr;
        }
        cp_async_commit();
    };

    auto sink = sinks.make(q, level);                     // where this window's K*K samples go
    const int shift = x_lo & 3;
    const T* my_row = ring + lane * PITCH;

#pragma unroll
    for (int r = 0; r < S - 1; ++r) issue_row(true);

    if (!slow) {
        float tprev[K];
#pragma unroll
        for (int a = 0; a < K; ++a) tprev[a] = 0.f;
        const T* sq = my_row;
#pragma unroll 1
        for (int r = 0; r <= K; ++r) {
            cp_async_wait<S - 2>();
            __syncwarp();
            issue_row(r + S - 1 <= K);
            float f[16];
            if constexpr (HALF) {
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    const uint4 t4 = reinterpret_cast<const uint4*>(sq)[v];
                    const uint32_t wds[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 c2 = __half22float2(*reinterpret_cast<const __half2*>(&wds[e]));
                        f[8 * v + 2 * e] = c2.x;
                        f[8 * v + 2 * e + 1] = c2.y;
                    }
                }
            } else {
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const float4 t4 = reinterpret_cast<const float4*>(sq)[v];
                    f[4 * v] = t4.x; f[4 * v + 1] = t4.y; f[4 * v + 2] = t4.z; f[4 * v + 3] = t4.w;
                }
            }
            // barrel shift by x_lo & 3 so that f[c] is window column c
#pragma unroll
            for (int c = 0; c < 15; ++c) f[c] = (shift & 1) ? f[c + 1] : f[c];
#pragma unroll
            for (int c = 0; c < 13; ++c) f[c] = (shift & 2) ? f[c + 2] : f[c];
            float tcur[K];
#pragma unroll
            for (int a = 0; a < K; ++a) tcur[a] = __fmaf_rn(wx1[a], f[a + 1], __fmul_rn(wx0[a], f[a]));
            if (r > 0) {
                float w0, w1;
                ytap(r - 1, w0, w1);
                if (lvl_on) {
                    float o[K];
#pragma unroll
                    for (int a = 0; a < K; ++a) o[a] = __fmaf_rn(w1, tcur[a], __fmul_rn(w0, tprev[a]));
                    sink.template row<K>(r - 1, o);
                }
            }
#pragma unroll
            for (int a = 0; a < K; ++a) tprev[a] = tcur[a];
            sq += ROWBUF;
            if (sq == my_row + S * ROWBUF) sq = my_row;
        }
    } else {
        const T* cur = my_row + shift;
        const T* prev = cur;
#pragma unroll 1
        for (int r = 0; r < W2; ++r) {
            cp_async_wait<S - 2>();
            __syncwarp();
#pragma unroll 1
            for (int bb = 0; bb < K; ++bb) {
                const int ryb = bb + (int)((py >> (2 * bb)) & 3u) - 1;
                if (ryb + 1 == r) {
                    float wy0b, wy1b;
                    ytap(bb, wy0b, wy1b);
#pragma unroll
                    for (int a = 0; a < K; ++a) {
                        const int rxa = a + (int)((px >> (2 * a)) & 3u) - 1;
                        const float v00 = (float)prev[rxa], v01 = (float)prev[rxa + 1], v10 = (float)cur[rxa], v11 = (float)cur[rxa + 1];
                        const float nw = __fmul_rn(wx0[a], wy0b);
                        const float ne = __fmul_rn(wx1[a], wy0b);
                        const float sw = __fmul_rn(wx0[a], wy1b);
                        const float se = __fmul_rn(wx1[a], wy1b);
                        float o = __fmul_rn(v00, nw);
                        o = __fmaf_rn(v01, ne, o);
                        o = __fmaf_rn(v10, sw, o);
                        o = __fmaf_rn(v11, se, o);
                        if (lvl_on) sink.template tap<K>(a, bb, o);
                    }
                }
            }
            __syncwarp();
            issue_row(r + S - 1 < W2);
            prev = cur;
            cur += ROWBUF;
            if (cur == my_row + shift + S * ROWBUF) cur = my_row + shift;
        }
    }
    cp_async_wait<0>();
}

template <int R, bool CUDA_SEM, typename T>
__global__ void __launch_bounds__(kNhwcWarps * 32) lookup_tiled_nhwc_kernel(const LookupTiledParams p) {
    static_assert(sizeof(T) == 4, "fp32 tiles; see lookup_tiled_nhwc_h_kernel for fp16 storage");
    constexpr int PITCH = RowGeom<T>::PITCH;
    constexpr int K = 2 * R + 1;
    constexpr int KK = K * K;
    constexpr int S = kStreamStages;
    constexpr int ROWBUF = kTile * PITCH;

    extern __shared__ __align__(16) unsigned char smem_nhwc[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int CT = p.num_levels * KK;
    T* ring = reinterpret_cast<T*>(smem_nhwc + (size_t)warp * nhwc_warp_bytes<T>(K, CT));
    float* siy = reinterpret_cast<float*>(ring + S * ROWBUF);
    float* stage = siy + K * kTile;                      // [8 queries][CT]

    const int b = blockIdx.x / p.blocks_per_batch;
    const int unit = (blockIdx.x - b * p.blocks_per_batch) * kNhwcWarps + warp;
    if (unit >= p.tiles_per_batch) return;               // warp-uniform; no block-level sync below
    const int N = p.N;
    const int n0 = unit * kNhwcQ;

    nhwc_lookup_unit<R, CUDA_SEM>(p, b, n0, ring, siy, lane, StageSinks{stage, CT, KK});

    // ---------------- flush: the warp's 8 x CT block is contiguous in the NHWC output ----------------
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the TMA engine
    __syncwarp();
    if (lane == 0) {
        const int nq = min(kNhwcQ, N - n0);
        float* gdst_ptr = p.out + ((int64_t)b * p.out_stride + n0) * CT;
        const uint32_t bytes = (uint32_t)(nq * CT) * 4u;               // CT*4 is a multiple of 16 only for even CT/4 ...
        if ((bytes & 15u) == 0 && (((uintptr_t)gdst_ptr) & 15u) == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst_ptr),
                         "r"((uint32_t)__cvta_generic_to_shared(stage)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory must outlive the read
        } else {
            for (int i = 0; i < nq * CT; ++i) gdst_ptr[i] = stage[i];        // odd tails only (never at r = 4, L = 4)
        }
    }
}

// ---------------------------------------------------------------------------------
// Channels-last lookup of the fp16-STORED pyramid (ffcorr_build_tiled_f16): row-PAIR streaming.
// A 4x4 tile of halfs is 32 bytes = one DRAM sector; its rows (r, r+1) for even r are 16 contiguous bytes, so the
// gather moves TWO window rows per cp.async.cg 16 (the 8-byte copy a single row needs exists only as cp.async.ca, which
// allocates in L1 and ran at 84 us; this kernel: half the copy instructions of the fp32 one, L2-only).  Ring stage per
// window = 4 tile columns x 16 bytes (+16 pad -> 80-byte pitch, conflict-free 16-byte reads).  A window's first row may
// be the odd row of its pair (y_lo & 1), so a lane evaluates window row 2s - (y_lo & 1) + rho at pair-step s, rho = 0, 1.
// Work decomposition, staging of the 8 x L*K*K results and the TMA bulk store are those of lookup_tiled_nhwc_kernel.
// ---------------------------------------------------------------------------------
__host__ __device__ constexpr int nhwc_h_warp_bytes(int K, int CT) { return kStreamStages * kTile * 80 + (K * kTile + kNhwcQ * CT) * 4; }

template <int R, bool CUDA_SEM>
__global__ void __launch_bounds__(kNhwcWarps * 32) lookup_tiled_nhwc_h_kernel(const LookupTiledParams p) {
    constexpr int K = 2 * R + 1;
    constexpr int W2 = K + 2;
    constexpr int KK = K * K;
    constexpr int S = kStreamStages;
    constexpr int PITCHB = 80;                            // bytes per window per stage
    constexpr int STAGEB = kTile * PITCHB;
    constexpr int NPAIR_FAST = (K + 3) / 2;               // pair-steps covering (y_lo & 1) + K + 1 rows
    constexpr int NPAIR_SLOW = (W2 + 2) / 2;              //                     (y_lo & 1) + W2 rows
    static_assert(W2 + 3 <= 16 && NPAIR_SLOW <= 8, "window must fit the 4x4 tile block");

    extern __shared__ __align__(16) unsigned char smem_nhwc_h[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int CT = p.num_levels * KK;
    unsigned char* ring = smem_nhwc_h + (size_t)warp * nhwc_h_warp_bytes(K, CT);
    float* siy = reinterpret_cast<float*>(ring + S * STAGEB);
    float* stage = siy + K * kTile;

    const int b = blockIdx.x / p.blocks_per_batch;
    const int unit = (blockIdx.x - b * p.blocks_per_batch) * kNhwcWarps + warp;
    if (unit >= p.tiles_per_batch) return;

    const int N = p.N;
    const int n0 = unit * kNhwcQ;
    const int lvl_of_lane = lane >> 3;
    const bool lvl_on = lvl_of_lane < p.num_levels;
    const int level = lvl_on ? lvl_of_lane : 0;
    const int q = lane & 7;
    const int n = min(n0 + q, N - 1);
    const int lh = p.lh[level], lw = p.lw[level];
    const int th = p.th[level], tw = p.tw[level];
    const int twr = (lw + 3) >> 2;
    const float inv_scale = __int_as_float((127 - level) << 23);

    // ---------------- phase A (lane = window): identical to the fp32 kernel ----------------
    const float* cptr = p.coords + (size_t)b * 2 * p.coords_stride + n;
    const float cx = __ldg(cptr) * inv_scale;
    const float cy = __ldg(cptr + p.coords_stride) * inv_scale;
    const float sx = (float)(lw - 1), sy = (float)(lh - 1);
    const float rsx = __frcp_rn(sx), rsy = __frcp_rn(sy);
    const float ixf = source_index<CUDA_SEM>(__fadd_rn(cx, (float)(-R)), sx, rsx);
    const float ixl = source_index<CUDA_SEM>(__fadd_rn(cx, (float)(R)), sx, rsx);
    const float iyf = source_index<CUDA_SEM>(__fadd_rn(cy, (float)(-R)), sy, rsy);
    const float iyl = source_index<CUDA_SEM>(__fadd_rn(cy, (float)(R)), sy, rsy);
    const bool wild = !lvl_on || !(fabsf(ixf) < kWildLimit) || !(fabsf(ixl) < kWildLimit) ||
                      !(fabsf(iyf) < kWildLimit) || !(fabsf(iyl) < kWildLimit);
    int x_lo = 0, y_lo = 0;
    int my_base = 0, my_pack = 0;
    if (!wild) {
        x_lo = (int)floorf(ixf);
        y_lo = (int)floorf(iyf);
        const int tx0 = x_lo >> 2, ty0 = y_lo >> 2;
        my_base = (ty0 * tw + tx0) * 16;
        int tmask = 0;
#pragma unroll
        for (int tyi = 0; tyi < 4; ++tyi)
#pragma unroll
            for (int txi = 0; txi < 4; ++txi)
                if ((unsigned)(ty0 + tyi) < (unsigned)th && (unsigned)(tx0 + txi) < (unsigned)twr) tmask |= 1 << (tyi * 4 + txi);
        my_pack = (x_lo & 3) | ((y_lo & 3) << 2) | (tmask << 4);
    }
    float wx0[K], wx1[K];
    unsigned px = 0, py = 0;
    bool deviated = false;
#pragma unroll
    for (int a = 0; a < K; ++a) {
        const float ix = source_index<CUDA_SEM>(__fadd_rn(cx, (float)(a - R)), sx, rsx);
        const float fx = floorf(ix);
        const float iy = source_index<CUDA_SEM>(__fadd_rn(cy, (float)(a - R)), sy, rsy);
        const float fy = floorf(iy);
        siy[a * kTile + lane] = iy;
        wx1[a] = wild ? 0.f : __fsub_rn(ix, fx);
        wx0[a] = wild ? 0.f : __fsub_rn(__fadd_rn(fx, 1.0f), ix);
        const int dxa = wild ? 0 : min(max((int)fx - x_lo, 0), W2 - 2) - a;
        const int dya = wild ? 0 : min(max((int)fy - y_lo, 0), W2 - 2) - a;
        px |= (unsigned)((dxa + 1) & 3) << (2 * a);
        py |= (unsigned)((dya + 1) & 3) << (2 * a);
        deviated |= (dxa != 0) | (dya != 0);
    }
    const bool slow = __any_sync(0xffffffffu, deviated);
    auto ytap = [&](int bb, float& w0, float& w1) {
        const float iy = siy[bb * kTile + lane];
        const float fy = floorf(iy);
        w1 = wild ? 0.f : __fsub_rn(iy, fy);
        w0 = wild ? 0.f : __fsub_rn(__fadd_rn(fy, 1.0f), iy);
    };

    // ---------------- gather slots: slot j = level j, lane = (query lane >> 2, tile column lane & 3) ----------------
    const int txi = lane & 3;
    const int qj = lane >> 2;
    const int last_q = N - 1 - n0;
    const __half* tbase[4];
    int goff[4], gnext[4], ppos[4], nstep[4];
    unsigned texist[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int lj = j < p.num_levels ? j : 0;
        const int map_elems = p.th[lj] * p.tw[lj] * 16;
        tbase[j] = reinterpret_cast<const __half*>(p.lvl[lj]) + ((int64_t)b * N + n0) * (int64_t)map_elems;
        const int base = __shfl_sync(0xffffffffu, my_base, j * 8 + qj);
        const int pk = __shfl_sync(0xffffffffu, my_pack, j * 8 + qj);
        const int gy = (pk >> 2) & 3;
        ppos[j] = gy >> 1;                                   // first pair: rows (0,1) or (2,3) of the first tile row
        goff[j] = min(qj, last_q) * map_elems + base + txi * 16 + ppos[j] * 8;
        const unsigned e = (unsigned)pk >> (4 + txi);
        unsigned ex = (e & 1u) | (((e >> 4) & 1u) << 1) | (((e >> 8) & 1u) << 2) | (((e >> 12) & 1u) << 3);
        if (txi * 4 >= (pk & 3) + (slow ? W2 : K + 1)) ex = 0;                  // only the tile columns the window reaches
        texist[j] = ex;
        nstep[j] = ((gy & 1) + (slow ? W2 : K + 1) + 1) >> 1;                   // pairs this window needs
        gnext[j] = p.tw[lj] * 16 - 8;
    }
    const unsigned gdst = (unsigned)(qj * PITCHB + txi * 16);
    const uint32_t ring_saddr = (uint32_t)__cvta_generic_to_shared(ring);
    int pairs_issued = 0;
    uint32_t issue_saddr = ring_saddr;
    auto issue_pair = [&](bool active) {
        if (active) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int P = ppos[j];
                const bool ok = ((texist[j] >> (P >> 1)) & 1u) && pairs_issued < nstep[j];
                const __half* src = tbase[j] + (ok ? goff[j] : 0);
                cp_async16_zfill(issue_saddr + gdst + (unsigned)(j * 8 * PITCHB), src, ok ? 16u : 0u);
                goff[j] += (P & 1) ? gnext[j] : 8;
                ppos[j] = P + 1;
            }
            ++pairs_issued;
            issue_saddr += STAGEB;
            if (issue_saddr == ring_saddr + S * STAGEB) issue_saddr = ring_saddr;
        }
        cp_async_commit();
    };

    float* sout = stage + q * CT + level * KK;
    const int shift = x_lo & 3;
    const int par = y_lo & 1;                              // the window's first row is the odd row of its pair
    const unsigned char* my_row = ring + lane * PITCHB;
    const int npair = slow ? NPAIR_SLOW : NPAIR_FAST;

#pragma unroll
    for (int r = 0; r < S - 1; ++r) issue_pair(true);      // NPAIR_* > S - 1 for every radius

    if (!slow) {
        float tprev[K];
#pragma unroll
        for (int a = 0; a < K; ++a) tprev[a] = 0.f;
        const unsigned char* sq = my_row;
#pragma unroll 1
        for (int s = 0; s < NPAIR_FAST; ++s) {
            cp_async_wait<S - 2>();
            __syncwarp();
            issue_pair(s + S - 1 < NPAIR_FAST);
            uint4 pc[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) pc[v] = reinterpret_cast<const uint4*>(sq)[v];
#pragma unroll
            for (int rho = 0; rho < 2; ++rho) {
                const int r = 2 * s - par + rho;            // window row of this half of the pair (per lane)
                float f[16];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const uint32_t w0 = rho ? pc[v].z : pc[v].x, w1 = rho ? pc[v].w : pc[v].y;
                    const float2 a2 = __half22float2(*reinterpret_cast<const __half2*>(&w0));
                    const float2 b2 = __half22float2(*reinterpret_cast<const __half2*>(&w1));
                    f[4 * v] = a2.x; f[4 * v + 1] = a2.y; f[4 * v + 2] = b2.x; f[4 * v + 3] = b2.y;
                }
#pragma unroll
                for (int c = 0; c < 15; ++c) f[c] = (shift & 1) ? f[c + 1] : f[c];
#pragma unroll
                for (int c = 0; c < 13; ++c) f[c] = (shift & 2) ? f[c + 2] : f[c];
                float tcur[K];
#pragma unroll
                for (int a = 0; a < K; ++a) tcur[a] = __fmaf_rn(wx1[a], f[a + 1], __fmul_rn(wx0[a], f[a]));
                if (r >= 1 && r <= K && lvl_on) {
                    float w0, w1;
                    ytap(r - 1, w0, w1);
#pragma unroll
                    for (int a = 0; a < K; ++a) sout[a * K + (r - 1)] = __fmaf_rn(w1, tcur[a], __fmul_rn(w0, tprev[a]));
                }
                if (r >= 0) {
#pragma unroll
                    for (int a = 0; a < K; ++a) tprev[a] = tcur[a];
                }
            }
            sq += STAGEB;
            if (sq == my_row + S * STAGEB) sq = my_row;
        }
    } else {
        // exact per-tap path: window row r sits in pair-step s = (r + par) >> 1, half rho = (r + par) & 1
        auto elem = [&](int step, int rho, int c) -> float {
            const __half* rowp = reinterpret_cast<const __half*>(my_row + (step % S) * STAGEB);
            const int cc = shift + c;
            return __half2float(rowp[(cc >> 2) * 8 + rho * 4 + (cc & 3)]);
        };
#pragma unroll 1
        for (int s = 0; s < NPAIR_SLOW; ++s) {
            cp_async_wait<S - 2>();
            __syncwarp();
#pragma unroll 1
            for (int rho = 0; rho < 2; ++rho) {
                const int r = 2 * s - par + rho;
                if (r < 1 || r > W2 - 1) continue;
                const int ps = (r - 1 + par) >> 1, prho = (r - 1 + par) & 1;       // where row r - 1 lives
#pragma unroll 1
                for (int bb = 0; bb < K; ++bb) {
                    const int ryb = bb + (int)((py >> (2 * bb)) & 3u) - 1;
                    if (ryb + 1 != r) continue;
                    float wy0b, wy1b;
                    ytap(bb, wy0b, wy1b);
#pragma unroll
                    for (int a = 0; a < K; ++a) {
                        const int rxa = a + (int)((px >> (2 * a)) & 3u) - 1;
                        const float v00 = elem(ps, prho, rxa), v01 = elem(ps, prho, rxa + 1);
                        const float v10 = elem(s, rho, rxa), v11 = elem(s, rho, rxa + 1);
                        const float nw = __fmul_rn(wx0[a], wy0b);
                        const float ne = __fmul_rn(wx1[a], wy0b);
                        const float sw = __fmul_rn(wx0[a], wy1b);
                        const float se = __fmul_rn(wx1[a], wy1b);
                        float o = __fmul_rn(v00, nw);
                        o = __fmaf_rn(v01, ne, o);
                        o = __fmaf_rn(v10, sw, o);
                        o = __fmaf_rn(v11, se, o);
                        if (lvl_on) sout[a * K + bb] = o;
                    }
                }
            }
            __syncwarp();                                   // the previous pair is no longer needed by anyone
            issue_pair(s + S - 1 < NPAIR_SLOW);
        }
    }
    (void)npair;
    cp_async_wait<0>();

    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        const int nq = min(kNhwcQ, N - n0);
        float* gdst_ptr = p.out + ((int64_t)b * p.out_stride + n0) * CT;
        const uint32_t bytes = (uint32_t)(nq * CT) * 4u;
        if ((bytes & 15u) == 0 && (((uintptr_t)gdst_ptr) & 15u) == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst_ptr),
                         "r"((uint32_t)__cvta_generic_to_shared(stage)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        } else {
            for (int i = 0; i < nq * CT; ++i) gdst_ptr[i] = stage[i];
        }
    }
}


// ---------------------------------------------------------------------------------
// Lookup fused with its consumer: cor = relu(convc1(lookup(coords)))  -- BasicMotionEncoder,
// update.py:82-83,90: a 1x1 convolution 324 -> 256 + ReLU over the lookup result, i.e. per query a [256 x 324] x [324]
// product.  The 324 samples of a query never leave the SM: the gather warps write them as fp16 into a UMMA operand
// tile in shared memory, the 5th-gen tensor cores multiply by the weights, and only the 256 outputs per query are written
// (NHWC, what the next convolution of the update block reads).  Saves the 76 MB lookup result round trip and the
// convolution launch per refinement iteration.
//
//   D^T[256 ch x 64 queries] = W'[256 x K'] * V^T[K' x 64 queries],  K' = 384 (360 used)
//   * W' (A operand) lives in TENSOR MEMORY for the whole kernel: 2 halves of 128 channels x 192 columns (two fp16 per
//     32-bit column), loaded once per CTA with tcgen05.st; the accumulators take the other 128 columns (512 in total, so
//     one persistent CTA per SM).  No weight traffic per tile.
//   * V (B operand): [64 queries][K'] fp16, K-major, 128B swizzle (6 atoms of 8 KB), double buffered.  K' orders the
//     samples [level][window row bb][a] with 10 slots per (level, bb) (9 samples + a zero), so the row-streaming lookup
//     emits one window row as five aligned half2 stores; the weights are permuted the same way on the host side
//     (ffcorr_pack_convc1_weight).  Precision: fp16 operands (11 significant bits, like the TF32 convolution the
//     reference runs under ALLOW_TF32), fp32 accumulate; samples saturate at +-65504.
//   * warps 0-7: gather (one unit of 8 queries each per tile, nhwc_lookup_unit); warps 8-11: epilogue (TMEM -> +bias -> ReLU
//     -> global; lane = channel, so a warp writes 128 contiguous bytes per query); warp 12: MMA issue (one thread).
// ---------------------------------------------------------------------------------
constexpr int kMoQ = 64;                       // queries per tile == UMMA N
constexpr int kMoGatherWarps = kMoQ / kNhwcQ;  // 8
constexpr int kMoEpiWarps = 4;
constexpr int kMoThreads = (kMoGatherWarps + kMoEpiWarps + 1) * 32;   // 416
constexpr int kMoSlots = 10;                   // K' slots per (level, window row): 9 samples + 1 zero
constexpr int kMoKLevel = 9 * kMoSlots;        // 90
constexpr int kMoK = 384;                      // padded K'
constexpr int kMoKSteps = 23;                  // k-steps of 16 that hold data (K' < 368)
constexpr int kMoCout = 256;
constexpr int kMoAtomBytes = kMoQ * 128;       // one 64-wide K atom of the sample tile
constexpr int kMoVBytes = (kMoK / 64) * kMoAtomBytes;                 // 48 KB
#ifndef FFCORR_MO_STAGES
#define FFCORR_MO_STAGES 4
#endif
constexpr int kMoStages = FFCORR_MO_STAGES;    // 8 gather warps per SM (the plain kernel has 10); 4, 5 and 6 stages measure the same
constexpr int kMoRingBytes = kMoStages * kTile * RowGeom<float>::PITCH * 4 + 9 * kTile * 4;   // ring + siy per gather warp
constexpr int kMoOffRing = 2 * kMoVBytes;
constexpr int kMoOffBar = kMoOffRing + kMoGatherWarps * kMoRingBytes;
constexpr int kMoSmem = kMoOffBar + 128;
constexpr int kMoWCols = kMoK / 2;             // TMEM columns per channel half of the weights (192)
constexpr int kMoAccCol = 2 * kMoWCols;        // first accumulator column (384)
static_assert(kMoAccCol + 2 * kMoQ == 512, "weights + accumulators fill the tensor memory exactly");
static_assert(kMoSmem <= 227 * 1024, "shared memory budget");
static_assert(kMoOffRing % 1024 == 0, "sample tiles must stay 1024-byte aligned (128B swizzle)");

__device__ __forceinline__ uint32_t pack_h2_sat(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));   // low half = a
    return r;
}

struct SampleTileSink {            // one query row of the fp16 operand tile
    uint32_t row_addr;             // shared address of the row inside atom 0
    uint32_t row7;                 // row & 7: the swizzle key
    int kbase;                     // level * 90
    __device__ __forceinline__ uint32_t addr(int k) const {
        return row_addr + (uint32_t)(k >> 6) * kMoAtomBytes + ((((uint32_t)(k & 63) >> 3) ^ row7) << 4) + ((uint32_t)(k & 7) << 1);
    }
    template <int K>
    __device__ __forceinline__ void row(int bb, const float (&v)[K]) {
        static_assert(K == 9, "the fused kernel is written for radius 4");
        const int k0 = kbase + bb * kMoSlots;          // even: a half2 never straddles a 16-byte chunk
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const uint32_t h2 = pack_h2_sat(v[2 * i], i < 4 ? v[i < 4 ? 2 * i + 1 : 0] : 0.0f);   // slot 9 of the row: zero
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr(k0 + 2 * i)), "r"(h2) : "memory");
        }
    }
    template <int K>
    __device__ __forceinline__ void tap(int a, int bb, float v) {
        const uint32_t h2 = pack_h2_sat(v, 0.0f);
        asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr(kbase + bb * kMoSlots + a)), "h"((unsigned short)(h2 & 0xffffu)) : "memory");
    }
};
struct SampleTileSinks {
    uint32_t tile_addr;            // shared address of the operand tile
    int row0;                      // first row of this unit
    __device__ __forceinline__ SampleTileSink make(int q, int level) const {
        const uint32_t r = (uint32_t)(row0 + q);
        return SampleTileSink{tile_addr + (r >> 3) * 1024u + (r & 7u) * 128u, r & 7u, level * kMoKLevel};
    }
};

__device__ __forceinline__ void umma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct ConvC1Params {
    const uint32_t* wpacked;       // [256][192]: fp16 pairs of W' (ffcorr_pack_convc1_weight)
    const float* bias;             // [256]
    float* out;                    // [B, N, 256] (NHWC)
    int num_tiles;                 // B * tiles_per_batch, tiles of 64 queries
    int* tile_counter;             // zero at launch: the tile scheduler's atomic
    uint32_t idesc;
};

#ifdef FFCORR_MO_TRACE      // development build: per-role clock64() stamps of every CTA (ffcorr_debug_mo_trace reads them)
__device__ long long g_mo_trace[148 * 128];
#define MO_TRACE(slot) do { if (lane == 0 && blockIdx.x < 148 && (slot) < 128) g_mo_trace[blockIdx.x * 128 + (slot)] = clock64(); } while (0)
#else
#define MO_TRACE(slot) do { } while (0)
#endif

template <bool CUDA_SEM>
__global__ void __launch_bounds__(kMoThreads, 1) lookup_convc1_kernel(const LookupTiledParams p, const ConvC1Params cp) {
    constexpr int R = 4;
    constexpr int S = kMoStages;
    constexpr int ROWBUF = kTile * RowGeom<float>::PITCH;
    extern __shared__ __align__(1024) uint8_t smem_mo[];
    const uint32_t smem_base = smem_u32(smem_mo);
    if (smem_base & 1023u) __trap();
    const uint32_t bar_base = smem_base + kMoOffBar;
    auto vfull = [&](int i) { return bar_base + 8u * i; };            // sample tile i written (8 gather warps)
    auto vempty = [&](int i) { return bar_base + 8u * (2 + i); };     // the MMAs that read sample tile i retired
    auto tready = [&](int i) { return bar_base + 8u * (4 + i); };     // tile_q[i] names the next tile for sample tile i
    const uint32_t accfull = bar_base + 48u, accempty = bar_base + 56u, wready = bar_base + 64u;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_mo + kMoOffBar + 72);
    volatile int* tile_q = reinterpret_cast<volatile int*>(smem_mo + kMoOffBar + 80);      // [2]
    volatile int* acc_tile = reinterpret_cast<volatile int*>(smem_mo + kMoOffBar + 88);    // [2]: tile of accumulator use k (k & 1)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp == 0) MO_TRACE(120);

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(vfull(i), kMoGatherWarps);
            mbar_init(vempty(i), 1);
            mbar_init(tready(i), 1);
        }
        mbar_init(accfull, 1);
        mbar_init(accempty, kMoEpiWarps);
        mbar_init(wready, kMoEpiWarps);
        fence_barrier_init();
    }
    if (warp == kMoGatherWarps + kMoEpiWarps) {
        tmem_alloc(smem_u32(tmem_slot), 512);
        MO_TRACE(121);
    }
    // both sample tiles start as zeros: the pad slots (a = 9 of every window row, K' >= 360) are never written with
    // anything else, and a 0 x NaN from uninitialised shared memory would poison the accumulators
    for (int i = threadIdx.x; i < 2 * kMoVBytes / 16; i += kMoThreads)
        reinterpret_cast<uint4*>(smem_mo)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == 0) MO_TRACE(122);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 0) MO_TRACE(0);

    // Tiles of 64 queries are handed out DYNAMICALLY (one atomicAdd on *cp.tile_counter per tile, by the MMA issuer, two
    // tiles ahead): a unit that takes the exact per-tap path costs 3x a normal one, and with a static round-robin the CTA
    // that met it finished 12-17 us after the rest (measured with the trace build).
    const int tiles_per_batch = p.tiles_per_batch;

    if (warp < kMoGatherWarps) {
        // ===================== gather warps: one unit of 8 queries per tile =====================
        float* ring = reinterpret_cast<float*>(smem_mo + kMoOffRing + warp * kMoRingBytes);
        float* siy = ring + S * ROWBUF;
        for (int it = 0;; ++it) {
            const int buf = it & 1;
            const uint32_t ph = ((uint32_t)it >> 1) & 1u;
            mbar_wait(tready(buf), ph);
            const int t = tile_q[buf];
            if (t >= cp.num_tiles) break;
            mbar_wait(vempty(buf), ph ^ 1u);                                // the MMAs that read this sample tile two rounds ago retired
            if (warp == 0) MO_TRACE(50 + it * 2);
            if (warp == 7) MO_TRACE(80 + it * 2);
            const int b = t / tiles_per_batch;
            const int n0 = (t - b * tiles_per_batch) * kMoQ + warp * kNhwcQ;
            if (n0 < p.N)
                nhwc_lookup_unit<R, CUDA_SEM, S>(p, b, n0, ring, siy, lane, SampleTileSinks{smem_base + (uint32_t)(buf * kMoVBytes), warp * kNhwcQ});
            if (warp == 0) MO_TRACE(51 + it * 2);
            if (warp == 7) MO_TRACE(81 + it * 2);
            fence_proxy_async_smem();                                       // generic writes -> visible to the tensor core's reads
            __syncwarp();
            if (lane == 0) mbar_arrive(vfull(buf));
        }
    } else if (warp < kMoGatherWarps + kMoEpiWarps) {
        // ===================== epilogue warps (first: the weights -> tensor memory) =====================
        const int e = warp - kMoGatherWarps;                     // == warp % 4: the TMEM lane quarter this warp may touch
        const uint32_t t_quarter = tmem_base + ((uint32_t)(e * 32) << 16);
        constexpr int PIECES = 2 * (kMoWCols / 32);              // 12 pieces of 32 columns; every CTA starts at another one
#pragma unroll 1                                                 // so that 148 CTAs do not read the same lines at the same time
        for (int i0 = 0; i0 < PIECES; ++i0) {
            const int c = (i0 + (int)blockIdx.x) % PIECES;
            const int half = c / (kMoWCols / 32), j = c - half * (kMoWCols / 32);
            const uint32_t* wrow = cp.wpacked + (size_t)(half * 128 + e * 32 + lane) * kMoWCols + j * 32;
            uint32_t rr[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(wrow) + i);
                rr[4 * i] = v.x; rr[4 * i + 1] = v.y; rr[4 * i + 2] = v.z; rr[4 * i + 3] = v.w;
            }
            tmem_st_32x32(t_quarter + (uint32_t)(half * kMoWCols + j * 32), rr);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(wready);          // only the MMA issuer waits for the weights; the gather starts at once
        if (warp == kMoGatherWarps) MO_TRACE(1);

        const uint32_t t_lane = t_quarter + (uint32_t)kMoAccCol;
        const float bias0 = __ldg(cp.bias + e * 32 + lane), bias1 = __ldg(cp.bias + 128 + e * 32 + lane);
        for (uint32_t k = 0;; ++k) {
            mbar_wait_relaxed(accfull, k & 1u, 256);
            const int t = acc_tile[k & 1u];
            if (t < 0) break;
            if (e == 0) MO_TRACE(100 + (int)k * 2);
            tc_fence_after();
            const int b = t / tiles_per_batch;
            const int n_tile = (t - b * tiles_per_batch) * kMoQ;
            float* orow = cp.out + ((size_t)b * p.N + n_tile) * kMoCout + e * 32 + lane;
            const int nq = min(kMoQ, p.N - n_tile);
            uint32_t v[64];
            tmem_ld_32x64(t_lane, v);                  // channels e*32+lane, queries 0..63
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < kMoQ; ++j)
                if (j < nq) orow[(size_t)j * kMoCout] = fmaxf(__uint_as_float(v[j]) + bias0, 0.0f);
            tmem_ld_32x64(t_lane + kMoQ, v);           // channels 128+e*32+lane
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(accempty);      // the accumulators may be overwritten
#pragma unroll
            for (int j = 0; j < kMoQ; ++j)
                if (j < nq) orow[(size_t)j * kMoCout + 128] = fmaxf(__uint_as_float(v[j]) + bias1, 0.0f);
            if (e == 0) MO_TRACE(101 + (int)k * 2);
        }
    } else if (lane == 0) {
        // ===================== tile scheduler + MMA issuer =====================
        int cur[2];
        auto hand_out = [&](int buf) {               // next tile for sample tile `buf` (>= num_tiles: none left)
            cur[buf] = atomicAdd(cp.tile_counter, 1);
            tile_q[buf] = cur[buf];
            mbar_arrive(tready(buf));                // release: the gather warps read tile_q after their acquire
        };
        hand_out(0);
        hand_out(1);
        mbar_wait(wready, 0);
        MO_TRACE(2);
        uint32_t k = 0;
        for (int it = 0;; ++it) {
            const int buf = it & 1;
            const int t = cur[buf];
            if (t >= cp.num_tiles) break;            // tiles come in increasing order: nothing later is valid either
            mbar_wait_relaxed(vfull(buf), ((uint32_t)it >> 1) & 1u, 128);
            MO_TRACE(10 + it * 4);
            hand_out(buf);                           // all gather warps are done reading tile_q[buf]
            mbar_wait(accempty, (k & 1u) ^ 1u);
            MO_TRACE(11 + it * 4);
            acc_tile[k & 1u] = t;
            __threadfence_block();
            ++k;
            tc_fence_after();
            // B: atom ks >> 2 of the sample tile, +32 bytes per k-step inside the atom; both in units of 16 bytes of the
            // descriptor's address field (shared addresses stay below 2^18, so the field never overflows)
            const uint64_t b0 = make_smem_desc(smem_base + (uint32_t)(buf * kMoVBytes));
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const uint32_t d_tmem = tmem_base + (uint32_t)(kMoAccCol + half * kMoQ);
                const uint32_t a_tmem = tmem_base + (uint32_t)(half * kMoWCols);
#pragma unroll
                for (int ks = 0; ks < kMoKSteps; ++ks)
                    umma_ts_f16(d_tmem, a_tmem + (uint32_t)(ks * 8), b0 + (uint64_t)((ks >> 2) * (kMoAtomBytes >> 4) + (ks & 3) * 2),
                                cp.idesc, (uint32_t)(ks != 0));
            }
            umma_commit(vempty(buf));
            umma_commit(accfull);
            MO_TRACE(12 + it * 4);
        }
        // tell the epilogue there is nothing more
        mbar_wait(accempty, (k & 1u) ^ 1u);
        acc_tile[k & 1u] = -1;
        __threadfence_block();
        mbar_arrive(accfull);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) MO_TRACE(127);
    if (warp == kMoGatherWarps + kMoEpiWarps) {
        tmem_dealloc(tmem_base, 512);
        MO_TRACE(123);
    }
}

// W [256][L*81] (convc1.weight, input channel = level*81 + a*9 + bb, corr.py:37-43) -> W' [256][384] fp16 in the K' order of
// the sample tile (level*90 + bb*10 + a; zero elsewhere), packed two per 32-bit word.
__global__ void pack_convc1_weight_kernel(const float* __restrict__ w, uint32_t* __restrict__ packed, int cin) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= kMoCout * kMoWCols) return;
    const int row = idx / kMoWCols, j = idx - row * kMoWCols;
    float v[2];
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
        const int k = 2 * j + hlf;
        const int level = k / kMoKLevel, rem = k - level * kMoKLevel;
        const int bb = rem / kMoSlots, a = rem - bb * kMoSlots;
        const int ci = level * 81 + a * 9 + bb;
        v[hlf] = (a < 9 && ci < cin && level * 81 < cin) ? w[(size_t)row * cin + ci] : 0.0f;
    }
    packed[idx] = pack_h2(v[0], v[1]);
}

template <int R>
int launch_lookup_tiled_nhwc_h(const LookupTiledParams& p0, int sampler, cudaStream_t stream) {
    constexpr int K = 2 * R + 1;
    LookupTiledParams p = p0;
    FFCORR_REQUIRE(p.num_levels <= 4, FFCORR_EINVAL, "lookup_tiled_f16: at most 4 levels, got %d", p.num_levels);
    p.tiles_per_batch = ceil_div(p.N, kNhwcQ);
    p.blocks_per_batch = ceil_div(p.tiles_per_batch, kNhwcWarps);
    const int64_t blocks = (int64_t)p.B * p.blocks_per_batch;
    FFCORR_REQUIRE(blocks < (1ll << 31), FFCORR_EINVAL, "lookup_tiled_f16: grid too large");
    const size_t smem = (size_t)kNhwcWarps * nhwc_h_warp_bytes(K, p.num_levels * K * K);
    FFCORR_CUDA(cudaFuncSetAttribute(lookup_tiled_nhwc_h_kernel<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FFCORR_CUDA(cudaFuncSetAttribute(lookup_tiled_nhwc_h_kernel<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (sampler == FFCORR_SAMPLER_ATEN_CUDA)
        lookup_tiled_nhwc_h_kernel<R, true><<<(unsigned)blocks, kNhwcWarps * 32, smem, stream>>>(p);
    else
        lookup_tiled_nhwc_h_kernel<R, false><<<(unsigned)blocks, kNhwcWarps * 32, smem, stream>>>(p);
    return check_launch("lookup_tiled_nhwc_h_kernel");
}

template <int R, typename T>
int launch_lookup_tiled_nhwc(const LookupTiledParams& p0, int sampler, cudaStream_t stream) {
    constexpr int K = 2 * R + 1;
    LookupTiledParams p = p0;
    FFCORR_REQUIRE(p.num_levels <= 4, FFCORR_EINVAL, "lookup_tiled (channels last): at most 4 levels, got %d", p.num_levels);
    p.tiles_per_batch = ceil_div(p.N, kNhwcQ);
    p.blocks_per_batch = ceil_div(p.tiles_per_batch, kNhwcWarps);
    const int64_t blocks = (int64_t)p.B * p.blocks_per_batch;
    FFCORR_REQUIRE(blocks < (1ll << 31), FFCORR_EINVAL, "lookup_tiled: grid too large");
    const size_t smem = (size_t)kNhwcWarps * nhwc_warp_bytes<T>(K, p.num_levels * K * K);
    FFCORR_CUDA(cudaFuncSetAttribute(lookup_tiled_nhwc_kernel<R, false, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FFCORR_CUDA(cudaFuncSetAttribute(lookup_tiled_nhwc_kernel<R, true, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (sampler == FFCORR_SAMPLER_ATEN_CUDA)
        lookup_tiled_nhwc_kernel<R, true, T><<<(unsigned)blocks, kNhwcWarps * 32, smem, stream>>>(p);
    else
        lookup_tiled_nhwc_kernel<R, false, T><<<(unsigned)blocks, kNhwcWarps * 32, smem, stream>>>(p);
    return check_launch("lookup_tiled_nhwc_kernel");
}

template <int R>
int launch_lookup_tiled_stream(const LookupTiledParams& p0, int sampler, cudaStream_t stream) {
    LookupTiledParams p = p0;
    p.blocks_per_batch = ceil_div(p.tiles_per_batch, kStreamWarps);
    const int64_t blocks = (int64_t)p.num_levels * p.B * p.blocks_per_batch;
    FFCORR_REQUIRE(blocks < (1ll << 31), FFCORR_EINVAL, "lookup_tiled: grid too large");
    if (sampler == FFCORR_SAMPLER_ATEN_CUDA)
        lookup_tiled_stream_kernel<R, true><<<(unsigned)blocks, kStreamWarps * 32, 0, stream>>>(p);
    else
        lookup_tiled_stream_kernel<R, false><<<(unsigned)blocks, kStreamWarps * 32, 0, stream>>>(p);
    return check_launch("lookup_tiled_stream_kernel");
}

// ---------------------------------------------------------------------------------
// adjoint w.r.t. the pyramid: one thread per output-gradient element, 4 atomics each.
// ---------------------------------------------------------------------------------
struct LookupBwdParams {
    float* glvl[FFCORR_MAX_LEVELS];
    int lh[FFCORR_MAX_LEVELS];
    int lw[FFCORR_MAX_LEVELS];
    const float* coords;
    const float* gout;
    int B, N, num_levels, radius;
};

// Mirror image of lookup_kernel.  Phase A (lane = query) as in the forward; then every lane builds the gradient
// of ITS window in shared memory -- no atomics, the window is private -- using the same separable form: the 81
// output gradients are folded along x into 9 horizontal rows, and window row r = wy0[r] * hrow[r] +
// wy1[r-1] * hrow[r-1]; finally the warp flushes the 32 windows cooperatively, 32 consecutive window elements
// per instruction, with red.global.add (the pyramid gradient accumulates over the lookups of all iterations).
// ~100 reductions per (query, level) in row-contiguous runs instead of 324 scattered ones.
template <int R, bool CUDA_SEM>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) lookup_bwd_kernel(const LookupBwdParams p, const int tiles_per_batch,
                                                                         const int blocks_per_batch) {
    constexpr int K = 2 * R + 1;
    constexpr int W2 = K + 2;
    constexpr int WIN = W2 * W2;
    constexpr int NLOAD = (WIN + 31) / 32;
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    float* swin = smem + warp * (kTile * WIN);

    int bid = blockIdx.x;
    const int per_level = p.B * blocks_per_batch;
    const int level = bid / per_level;
    bid -= level * per_level;
    const int b = bid / blocks_per_batch;
    const int tile = (bid - b * blocks_per_batch) * kWarpsPerBlock + warp;
    if (tile >= tiles_per_batch) return;

    const int N = p.N;
    const int n0 = tile * kTile;
    const int n = n0 + lane;
    const bool valid = n < N;
    const int lh = p.lh[level], lw = p.lw[level];
    const float inv_scale = __int_as_float((127 - level) << 23);

    // ---------------- phase A: window origin and in-bounds masks (as lookup_kernel) ----------------
    float cx = 0.f, cy = 0.f;
    if (valid) {
        const float* c = p.coords + (size_t)b * 2 * N + n;
        cx = __ldg(c) * inv_scale;
        cy = __ldg(c + N) * inv_scale;
    }
    const float sx = (float)(lw - 1), sy = (float)(lh - 1);
    const float rsx = __frcp_rn(sx), rsy = __frcp_rn(sy);     // ATen: 1.0 / scalar, in fp32
    const float ixf = source_index<CUDA_SEM>(__fadd_rn(cx, (float)(-R)), sx, rsx);
    const float ixl = source_index<CUDA_SEM>(__fadd_rn(cx, (float)(R)), sx, rsx);
    const float iyf = source_index<CUDA_SEM>(__fadd_rn(cy, (float)(-R)), sy, rsy);
    const float iyl = source_index<CUDA_SEM>(__fadd_rn(cy, (float)(R)), sy, rsy);
    const bool wild = !(fabsf(ixf) < kWildLimit) || !(fabsf(ixl) < kWildLimit) ||
                      !(fabsf(iyf) < kWildLimit) || !(fabsf(iyl) < kWildLimit);
    const int map_elems = lh * lw;
    int x_lo = 0, y_lo = 0;
    int my_mask = 0, my_qoff = 0;
    if (valid && !wild) {
        x_lo = (int)floorf(ixf);
        y_lo = (int)floorf(iyf);
        const int ncols = min((int)floorf(ixl) + 2 - x_lo, W2);
        const int nrows = min((int)floorf(iyl) + 2 - y_lo, W2);
        const int rlo = max(0, -y_lo), rhi = min(nrows, lh - y_lo);
        const int clo = max(0, -x_lo), chi = min(ncols, lw - x_lo);
        const int rm = (rhi > rlo) ? (((1 << rhi) - 1) & ~((1 << rlo) - 1)) : 0;
        const int cm = (chi > clo) ? (((1 << chi) - 1) & ~((1 << clo) - 1)) : 0;
        my_mask = (rm && cm) ? (rm | (cm << 16)) : 0;
        my_qoff = lane * map_elems + y_lo * lw + x_lo;
    }

    int rx[K], ry[K];
    float wx0[K], wx1[K], wy0[K], wy1[K];
    bool deviated = false;
#pragma unroll
    for (int a = 0; a < K; ++a) {
        const float ix = source_index<CUDA_SEM>(__fadd_rn(cx, (float)(a - R)), sx, rsx);
        const float fx = floorf(ix);
        const float iy = source_index<CUDA_SEM>(__fadd_rn(cy, (float)(a - R)), sy, rsy);
        const float fy = floorf(iy);
        const bool dead = wild || !valid;
        wx1[a] = dead ? 0.f : __fsub_rn(ix, fx);
        wx0[a] = dead ? 0.f : __fsub_rn(__fadd_rn(fx, 1.0f), ix);
        wy1[a] = dead ? 0.f : __fsub_rn(iy, fy);
        wy0[a] = dead ? 0.f : __fsub_rn(__fadd_rn(fy, 1.0f), iy);
        rx[a] = dead ? a : min(max((int)fx - x_lo, 0), W2 - 2);
        ry[a] = dead ? a : min(max((int)fy - y_lo, 0), W2 - 2);
        deviated |= (rx[a] != a) | (ry[a] != a);
    }

    // ---------------- per-lane window gradient in shared memory ----------------
    float* sq = swin + lane * WIN;
    const int CT = p.num_levels * K * K;
    const float* __restrict__ gp = p.gout + ((int64_t)b * CT + (int64_t)level * K * K) * N + min(n, N - 1);
    if (!__any_sync(0xffffffffu, deviated)) {
        float hprev[K + 1];
#pragma unroll
        for (int c = 0; c <= K; ++c) hprev[c] = 0.f;
#pragma unroll
        for (int r = 0; r <= K; ++r) {
            float hcur[K + 1];
#pragma unroll
            for (int c = 0; c <= K; ++c) hcur[c] = 0.f;
            if (r < K) {
#pragma unroll
                for (int a = 0; a < K; ++a) {
                    const float g = __ldg(gp + (int64_t)(a * K + r) * N);
                    hcur[a] = fmaf(g, wx0[a], hcur[a]);
                    hcur[a + 1] = fmaf(g, wx1[a], hcur[a + 1]);
                }
            }
#pragma unroll
            for (int c = 0; c <= K; ++c) {
                float v = (r > 0) ? wy1[r > 0 ? r - 1 : 0] * hprev[c] : 0.f;
                if (r < K) v = fmaf(wy0[r < K ? r : 0], hcur[c], v);
                sq[r * W2 + c] = v;
            }
            sq[r * W2 + K + 1] = 0.f;
#pragma unroll
            for (int c = 0; c <= K; ++c) hprev[c] = hcur[c];
        }
#pragma unroll
        for (int c = 0; c < W2; ++c) sq[(K + 1) * W2 + c] = 0.f;
    } else {
        for (int e = 0; e < WIN; ++e) sq[e] = 0.f;
#pragma unroll
        for (int a = 0; a < K; ++a) {
#pragma unroll
            for (int bb = 0; bb < K; ++bb) {
                const float g = __ldg(gp + (int64_t)(a * K + bb) * N);
                float* s = sq + ry[bb] * W2 + rx[a];
                s[0] = fmaf(g, wx0[a] * wy0[bb], s[0]);
                s[1] = fmaf(g, wx1[a] * wy0[bb], s[1]);
                s[W2] = fmaf(g, wx0[a] * wy1[bb], s[W2]);
                s[W2 + 1] = fmaf(g, wx1[a] * wy1[bb], s[W2 + 1]);
            }
        }
    }
    __syncwarp();

    // ---------------- cooperative flush: red.global.add of the in-bounds window elements ----------------
    float* __restrict__ tile_base = p.glvl[level] + ((int64_t)b * N + n0) * (int64_t)map_elems;
    int poff[NLOAD], bits[NLOAD];
#pragma unroll
    for (int j = 0; j < NLOAD; ++j) {
        const int e = lane + 32 * j;
        const int er = e / W2, ec = e - er * W2;
        poff[j] = er * lw + ec;
        bits[j] = (e < WIN) ? ((1 << er) | (1 << (16 + ec))) : 0x80008000;  // never matches
    }
#pragma unroll 4
    for (int q = 0; q < kTile; ++q) {
        const int qoff = __shfl_sync(0xffffffffu, my_qoff, q);
        const int msk = __shfl_sync(0xffffffffu, my_mask, q);
#pragma unroll
        for (int j = 0; j < NLOAD; ++j) {
            if ((msk & bits[j]) == bits[j]) {
                const float v = swin[q * WIN + lane + 32 * j];
                if (v != 0.f) atomicAdd(tile_base + (poff[j] + qoff), v);
            }
        }
    }
}

template <int R>
int launch_lookup_bwd(const LookupBwdParams& p, int sampler, cudaStream_t stream) {
    constexpr int K = 2 * R + 1;
    constexpr int WIN = (K + 2) * (K + 2);
    const size_t smem = (size_t)kWarpsPerBlock * kTile * WIN * sizeof(float);
    FFCORR_CUDA(cudaFuncSetAttribute(lookup_bwd_kernel<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FFCORR_CUDA(cudaFuncSetAttribute(lookup_bwd_kernel<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles_per_batch = ceil_div(p.N, kTile);
    const int blocks_per_batch = ceil_div(tiles_per_batch, kWarpsPerBlock);
    const int64_t blocks = (int64_t)p.num_levels * p.B * blocks_per_batch;
    FFCORR_REQUIRE(blocks < (1ll << 31), FFCORR_EINVAL, "lookup_bwd: grid too large");
    if (sampler == FFCORR_SAMPLER_ATEN_CUDA)
        lookup_bwd_kernel<R, true><<<(unsigned)blocks, kWarpsPerBlock * 32, smem, stream>>>(p, tiles_per_batch, blocks_per_batch);
    else
        lookup_bwd_kernel<R, false><<<(unsigned)blocks, kWarpsPerBlock * 32, smem, stream>>>(p, tiles_per_batch, blocks_per_batch);
    return check_launch("lookup_bwd_kernel");
}

template <int R, int QU>
int launch_lookup(const LookupParams& p, int sampler, cudaStream_t stream) {
    constexpr int K = 2 * R + 1;
    constexpr int WIN = (K + 2) * (K + 2);
    const size_t smem = (size_t)kWarpsPerBlock * kTile * WIN * sizeof(float);
    static thread_local int configured_dev = -1;
    int dev = 0;
    FFCORR_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        FFCORR_CUDA(cudaFuncSetAttribute(lookup_kernel<R, QU, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FFCORR_CUDA(cudaFuncSetAttribute(lookup_kernel<R, QU, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured_dev = dev;
    }
    const int64_t blocks = (int64_t)p.num_levels * p.B * p.blocks_per_batch;
    FFCORR_REQUIRE(blocks < (1ll << 31), FFCORR_EINVAL, "lookup: grid too large (%lld blocks)", (long long)blocks);
    if (sampler == FFCORR_SAMPLER_ATEN_CUDA)
        lookup_kernel<R, QU, true><<<(unsigned)blocks, kWarpsPerBlock * 32, smem, stream>>>(p);
    else
        lookup_kernel<R, QU, false><<<(unsigned)blocks, kWarpsPerBlock * 32, smem, stream>>>(p);
    return check_launch("lookup_kernel");
}

}  // namespace
}  // namespace ffcorr

using namespace ffcorr;

static int check_sampler(int sampler, const char* who) {
    FFCORR_REQUIRE(sampler == FFCORR_SAMPLER_ATEN_CPU || sampler == FFCORR_SAMPLER_ATEN_CUDA, FFCORR_EINVAL,
                   "%s: sampler=%d is neither FFCORR_SAMPLER_ATEN_CPU nor FFCORR_SAMPLER_ATEN_CUDA", who, sampler);
    return FFCORR_OK;
}

extern "C" int ffcorr_lookup_f32(const float* const* lvl, int num_levels, const float* coords, float* out,
                                 int B, int h, int w, int radius, int sampler, int out_channels_last, void* stream) {
    FFCORR_REQUIRE(B >= 0, FFCORR_EINVAL, "lookup: B=%d", B);
    if (B == 0) return FFCORR_OK;  // empty batch: pointers may legitimately be null
    FFCORR_REQUIRE(lvl && coords && out, FFCORR_EINVAL, "lookup: null pointer");
    FFCORR_REQUIRE(radius >= 1 && radius <= 4, FFCORR_EINVAL, "lookup: radius=%d outside [1,4]", radius);
    if (int rc = check_sampler(sampler, "lookup")) return rc;
    if (int rc = check_levels(num_levels, h, w, "lookup")) return rc;
    // the kernel keeps a 32-query tile's element offsets in 32 bits: 31 * h*w + window offset < 2^31
    FFCORR_REQUIRE((int64_t)h * w < (1ll << 24), FFCORR_EINVAL, "lookup: h*w = %lld must be below 2^24", (long long)h * w);
    LookupParams p{};
    for (int i = 0; i < num_levels; ++i) {
        FFCORR_REQUIRE(lvl[i] != nullptr, FFCORR_EINVAL, "lookup: lvl[%d] is null", i);
        p.lvl[i] = lvl[i];
        p.lh[i] = h >> i;
        p.lw[i] = w >> i;
    }
    p.coords = coords;
    p.out = out;
    p.B = B;
    p.N = h * w;
    p.num_levels = num_levels;
    p.tiles_per_batch = ceil_div(p.N, kTile);
    p.blocks_per_batch = ceil_div(p.tiles_per_batch, kWarpsPerBlock);
    const int CT = num_levels * (2 * radius + 1) * (2 * radius + 1);
    p.cstride = out_channels_last ? 1 : p.N;
    p.qstride = out_channels_last ? CT : 1;
    cudaStream_t s = (cudaStream_t)stream;
    switch (radius) {
        case 1: return launch_lookup<1, 4>(p, sampler, s);
        case 2: return launch_lookup<2, 4>(p, sampler, s);
        case 3: return launch_lookup<3, 4>(p, sampler, s);
        default: return launch_lookup<4, 8>(p, sampler, s);
    }
}

extern "C" int ffcorr_lookup_bwd_f32(float* const* grad_lvl, int num_levels, const float* coords,
                                     const float* grad_out, int B, int h, int w, int radius, int sampler, void* stream) {
    FFCORR_REQUIRE(B >= 0, FFCORR_EINVAL, "lookup_bwd: B=%d", B);
    if (B == 0) return FFCORR_OK;
    FFCORR_REQUIRE(grad_lvl && coords && grad_out, FFCORR_EINVAL, "lookup_bwd: null pointer");
    FFCORR_REQUIRE(radius >= 1 && radius <= 4, FFCORR_EINVAL, "lookup_bwd: radius=%d outside [1,4]", radius);
    if (int rc = check_sampler(sampler, "lookup_bwd")) return rc;
    if (int rc = check_levels(num_levels, h, w, "lookup_bwd")) return rc;
    LookupBwdParams p{};
    for (int i = 0; i < num_levels; ++i) {
        FFCORR_REQUIRE(grad_lvl[i] != nullptr, FFCORR_EINVAL, "lookup_bwd: grad_lvl[%d] is null", i);
        p.glvl[i] = grad_lvl[i];
        p.lh[i] = h >> i;
        p.lw[i] = w >> i;
    }
    p.coords = coords;
    p.gout = grad_out;
    p.B = B;
    p.N = h * w;
    p.num_levels = num_levels;
    p.radius = radius;
    FFCORR_REQUIRE((int64_t)(h) * w < (1ll << 24), FFCORR_EINVAL, "lookup_bwd: h*w too large");
    cudaStream_t s = (cudaStream_t)stream;
    switch (radius) {
        case 1: return launch_lookup_bwd<1>(p, sampler, s);
        case 2: return launch_lookup_bwd<2>(p, sampler, s);
        case 3: return launch_lookup_bwd<3>(p, sampler, s);
        default: return launch_lookup_bwd<4>(p, sampler, s);
    }
}


static int lookup_tiled_impl(const float* const* lvl, int num_levels, const float* coords, float* out, int B, int h, int w,
                             int nq, int64_t coords_stride, int64_t out_stride, int radius, int sampler, int out_channels_last,
                             void* stream, const char* who, bool half_storage = false, bool grouped = false) {
    FFCORR_REQUIRE(B >= 0, FFCORR_EINVAL, "%s: B=%d", who, B);
    if (B == 0) return FFCORR_OK;
    FFCORR_REQUIRE(lvl && coords && out, FFCORR_EINVAL, "%s: null pointer", who);
    FFCORR_REQUIRE(radius >= 1 && radius <= 4, FFCORR_EINVAL, "%s: radius=%d outside [1,4]", who, radius);
    if (int rc = check_sampler(sampler, who)) return rc;
    if (int rc = check_levels(num_levels, h, w, who)) return rc;
    FFCORR_REQUIRE(nq >= 1 && nq <= h * w, FFCORR_EINVAL, "%s: %d queries for a %dx%d map", who, nq, h, w);
    FFCORR_REQUIRE((int64_t)h * w < (1ll << 24), FFCORR_EINVAL, "%s: h*w = %lld must be below 2^24", who, (long long)h * w);
    LookupTiledParams p{};
    for (int i = 0; i < num_levels; ++i) {
        FFCORR_REQUIRE(lvl[i] != nullptr, FFCORR_EINVAL, "%s: lvl[%d] is null", who, i);
        FFCORR_REQUIRE((uintptr_t)lvl[i] % 16 == 0, FFCORR_EALIGN, "%s: lvl[%d] must be 16-byte aligned", who, i);
        p.lvl[i] = lvl[i];
        p.lh[i] = h >> i;
        p.lw[i] = w >> i;
        p.th[i] = tiled_th(p.lh[i]);
        p.tw[i] = tiled_tw(p.lw[i]);
    }
    p.coords = coords;
    p.out = out;
    p.B = B;
    p.N = nq;
    p.coords_stride = coords_stride;
    p.out_stride = out_stride;
    p.num_levels = num_levels;
    p.grouped = grouped ? 1 : 0;
    p.NG = ceil_div(nq, kTile);
    FFCORR_REQUIRE(!(grouped && half_storage), FFCORR_EINVAL, "%s: the grouped layout stores fp32", who);
    p.tiles_per_batch = ceil_div(p.N, kTile);
    p.blocks_per_batch = ceil_div(p.tiles_per_batch, kWarpsPerBlock);
    cudaStream_t s = (cudaStream_t)stream;
    if (half_storage) {
        FFCORR_REQUIRE(out_channels_last, FFCORR_EINVAL, "%s: the fp16-stored pyramid is read by the channels-last kernel only", who);
        switch (radius) {
            case 1: return launch_lookup_tiled_nhwc_h<1>(p, sampler, s);
            case 2: return launch_lookup_tiled_nhwc_h<2>(p, sampler, s);
            case 3: return launch_lookup_tiled_nhwc_h<3>(p, sampler, s);
            default: return launch_lookup_tiled_nhwc_h<4>(p, sampler, s);
        }
    }
    if (out_channels_last) {
        switch (radius) {
            case 1: return launch_lookup_tiled_nhwc<1, float>(p, sampler, s);
            case 2: return launch_lookup_tiled_nhwc<2, float>(p, sampler, s);
            case 3: return launch_lookup_tiled_nhwc<3, float>(p, sampler, s);
            default: return launch_lookup_tiled_nhwc<4, float>(p, sampler, s);
        }
    }
    switch (radius) {
        case 1: return launch_lookup_tiled_stream<1>(p, sampler, s);
        case 2: return launch_lookup_tiled_stream<2>(p, sampler, s);
        case 3: return launch_lookup_tiled_stream<3>(p, sampler, s);
        default: return launch_lookup_tiled_stream<4>(p, sampler, s);
    }
}

extern "C" int ffcorr_lookup_tiled_f32(const float* const* lvl, int num_levels, const float* coords, float* out,
                                       int B, int h, int w, int radius, int sampler, int out_channels_last, void* stream) {
    return lookup_tiled_impl(lvl, num_levels, coords, out, B, h, w, h * w, (int64_t)h * w, (int64_t)h * w, radius, sampler,
                             out_channels_last, stream, "lookup_tiled");
}

#ifdef FFCORR_MO_TRACE
extern "C" int ffcorr_debug_mo_trace(void* dst) {
    FFCORR_CUDA(cudaMemcpyFromSymbol(dst, g_mo_trace, sizeof(g_mo_trace)));
    return FFCORR_OK;
}
#endif

extern "C" size_t ffcorr_convc1_packed_bytes(void) { return (size_t)kMoCout * kMoWCols * 4; }

extern "C" int ffcorr_pack_convc1_weight(const float* weight, int cout, int cin, void* packed, void* stream) {
    FFCORR_REQUIRE(weight && packed, FFCORR_EINVAL, "pack_convc1_weight: null pointer");
    FFCORR_REQUIRE(cout == kMoCout && cin == 4 * 81, FFCORR_EINVAL,
                   "pack_convc1_weight: the fused kernel is built for convc1 = Conv2d(324, 256, 1) (update.py:82), got %d -> %d", cin, cout);
    const int total = kMoCout * kMoWCols;
    pack_convc1_weight_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(weight, reinterpret_cast<uint32_t*>(packed), cin);
    return check_launch("pack_convc1_weight_kernel");
}

extern "C" int ffcorr_lookup_convc1_tiled_f32(const float* const* lvl, int num_levels, const float* coords, const void* packed_weight,
                                              const float* bias, float* out, int* tile_counter, int B, int h, int w, int radius,
                                              int sampler, void* stream) {
    const char* who = "lookup_convc1_tiled";
    FFCORR_REQUIRE(B >= 0, FFCORR_EINVAL, "%s: B=%d", who, B);
    if (B == 0) return FFCORR_OK;
    FFCORR_REQUIRE(lvl && coords && packed_weight && bias && out && tile_counter, FFCORR_EINVAL, "%s: null pointer", who);
    FFCORR_REQUIRE(num_levels == 4 && radius == 4, FFCORR_EINVAL,
                   "%s: built for the reference's 4 levels x radius 4 (324 planes, update.py:82), got %d levels, radius %d", who,
                   num_levels, radius);
    if (int rc = check_sampler(sampler, who)) return rc;
    if (int rc = check_levels(num_levels, h, w, who)) return rc;
    FFCORR_REQUIRE((int64_t)h * w < (1ll << 24), FFCORR_EINVAL, "%s: h*w = %lld must be below 2^24", who, (long long)h * w);
    FFCORR_REQUIRE((uintptr_t)packed_weight % 16 == 0 && (uintptr_t)out % 16 == 0, FFCORR_EALIGN, "%s: packed weight / out must be 16-byte aligned", who);
    LookupTiledParams p{};
    for (int i = 0; i < num_levels; ++i) {
        FFCORR_REQUIRE(lvl[i] != nullptr, FFCORR_EINVAL, "%s: lvl[%d] is null", who, i);
        FFCORR_REQUIRE((uintptr_t)lvl[i] % 16 == 0, FFCORR_EALIGN, "%s: lvl[%d] must be 16-byte aligned", who, i);
        p.lvl[i] = lvl[i];
        p.lh[i] = h >> i;
        p.lw[i] = w >> i;
        p.th[i] = tiled_th(p.lh[i]);
        p.tw[i] = tiled_tw(p.lw[i]);
    }
    p.coords = coords;
    p.out = nullptr;
    p.B = B;
    p.N = h * w;
    p.coords_stride = (int64_t)h * w;
    p.out_stride = (int64_t)h * w;
    p.num_levels = num_levels;
    p.tiles_per_batch = ceil_div(p.N, kMoQ);
    p.blocks_per_batch = p.tiles_per_batch;
    const int64_t num_tiles = (int64_t)B * p.tiles_per_batch;
    FFCORR_REQUIRE(num_tiles < (1ll << 31), FFCORR_EINVAL, "%s: too many tiles", who);
    ConvC1Params cp{};
    cp.wpacked = reinterpret_cast<const uint32_t*>(packed_weight);
    cp.bias = bias;
    cp.out = out;
    cp.num_tiles = (int)num_tiles;
    cp.tile_counter = tile_counter;
    // instruction descriptor: D = f32, A = B = f16, both K-major, N >> 3, M >> 4 (M = 128 channels, N = 64 queries)
    cp.idesc = (1u << 4) | ((uint32_t)(kMoQ >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const int grid = (int)(num_tiles < sm_count() ? num_tiles : sm_count());
    cudaStream_t s = (cudaStream_t)stream;
    FFCORR_CUDA(cudaFuncSetAttribute(lookup_convc1_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMoSmem));
    FFCORR_CUDA(cudaFuncSetAttribute(lookup_convc1_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMoSmem));
    if (sampler == FFCORR_SAMPLER_ATEN_CUDA)
        lookup_convc1_kernel<true><<<grid, kMoThreads, kMoSmem, s>>>(p, cp);
    else
        lookup_convc1_kernel<false><<<grid, kMoThreads, kMoSmem, s>>>(p, cp);
    return check_launch("lookup_convc1_kernel");
}

extern "C" int ffcorr_lookup_tiled_f16(const void* const* lvl, int num_levels, const float* coords, float* out,
                                       int B, int h, int w, int radius, int sampler, int out_channels_last, void* stream) {
    return lookup_tiled_impl(reinterpret_cast<const float* const*>(lvl), num_levels, coords, out, B, h, w, h * w, (int64_t)h * w,
                             (int64_t)h * w, radius, sampler, out_channels_last, stream, "lookup_tiled_f16", true);
}

extern "C" int ffcorr_lookup_grouped_f32(const float* const* lvl, int num_levels, const float* coords, float* out,
                                         int B, int h, int w, int radius, int sampler, int out_channels_last, void* stream) {
    return lookup_tiled_impl(lvl, num_levels, coords, out, B, h, w, h * w, (int64_t)h * w, (int64_t)h * w, radius, sampler,
                             out_channels_last, stream, "lookup_grouped", false, true);
}

extern "C" int ffcorr_lookup_tiled_chunk_f32(const float* const* lvl, int num_levels, const float* coords, float* out,
                                             int B, int h, int w, int q0, int nq, int radius, int sampler, int out_channels_last,
                                             void* stream) {
    FFCORR_REQUIRE(q0 >= 0 && nq >= 1 && (int64_t)q0 + nq <= (int64_t)h * w, FFCORR_EINVAL,
                   "lookup_tiled_chunk: query range [%d, %d) outside the %dx%d map", q0, q0 + nq, h, w);
    if (B == 0) return FFCORR_OK;
    FFCORR_REQUIRE(coords && out, FFCORR_EINVAL, "lookup_tiled_chunk: null pointer");
    FFCORR_REQUIRE(num_levels >= 1 && radius >= 1 && radius <= 4, FFCORR_EINVAL, "lookup_tiled_chunk: bad levels / radius");
    // coords / out are the FULL [B, 2, h*w] / [B, L*K*K, h*w] (or [B, h*w, L*K*K]) tensors; the chunk reads and writes
    // its slice in place
    const int64_t CT = (int64_t)num_levels * (2 * radius + 1) * (2 * radius + 1);
    float* out_chunk = out_channels_last ? out + (int64_t)q0 * CT : out + q0;
    return lookup_tiled_impl(lvl, num_levels, coords + q0, out_chunk, B, h, w, nq, (int64_t)h * w, (int64_t)h * w, radius,
                             sampler, out_channels_last, stream, "lookup_tiled_chunk");
}
