// All-pairs correlation volume (level 0 of the pyramid) for sm_100a, optionally with the whole avg-pool
// pyramid fused into the GEMM epilogue, and the two backward GEMMs.
//
// Replaces torch.matmul(fmap1^T, fmap2) followed by a separate "/ sqrt(D)" pass
// (FF_RAFT_Core/corr.py:52-60) -- and, in the fused build, the three avg_pool2d passes of corr.py:24-27.
// C[b,i,j] = sum_d f1[b,d,i] f2[b,d,j] / sqrt(D).
//
// Tensor-core path (default):
//   1. operand pre-pass: [B, D, N] fp32 (MN-major, as the encoder hands it over) ->
//      [B, N, Dp] K-major fp16 / bf16-split / fp32(tf32), zero padded along K; the B operand's rows are
//      permuted into the storage order of the output (row-major pixels, 4x4 tiles, or 16x16 super-groups).
//      Traffic: 2*B*N*D*(4 + e) bytes, ~2% of the volume write.
//   2. persistent warp-specialised GEMM, one CTA per SM, tile 128 x 256:
//        warp 0    TMA producer   (cp.async.bulk.tensor, 128B swizzle, mbarrier ring)
//        warp 1    MMA issuer     (tcgen05.mma cta_group::1, M=128 N=256, fp32 accumulators in
//                                  TMEM, 2 x 256 columns so tile t+1 overlaps the epilogue of t)
//        warps 2.. epilogue       (tcgen05.ld -> scale by 1/sqrt(D) [-> pool levels 1..3] -> swizzled smem ->
//                                  TMA store; each warp owns 32 TMEM lanes and a private store ring)
//   The kernel is HBM-WRITE bound, not tensor bound: per tile 128 KB of fp32 leave the SM for
//   2*128*256*D flop (AI ~ 120 flop/B at D=256 < B200 ridge ~215), so the design spends shared
//   memory on store buffering rather than on a deep operand pipeline.  See Cfg<> and DESIGN.md 3.1-3.2.
// Exact path: FFCORR_PREC_FP32, a CUDA-core SGEMM (also the exact backward).
// Backward: gf1 = g * f2^T, gf2 = g^T * f1 on the same tcgen05 kernel with kind::tf32 reading the fp32 arrays
// in place (both operands are K-major as they lie; g is transposed in place for the second product).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"

namespace ffcorr {
namespace {

// ------------------------------------------------------------------------------------------
// GEMM configuration
// ------------------------------------------------------------------------------------------
constexpr int BM = 128;              // tile rows (queries i)     == UMMA M == TMEM lanes
constexpr int BN = 256;              // tile cols (targets j)     == UMMA N == TMEM columns per accumulator
constexpr int BK_BYTES = 128;        // one 128B swizzle atom along K per stage
constexpr int UMMA_K_BYTES = 32;     // K extent of one tcgen05.mma: 16 x 16-bit or 8 x tf32
constexpr int ACC_STAGES = 2;
constexpr int STORE_COLS = 32;       // fp32 columns per store box (128 B rows)
constexpr int SMEM_A_STAGE = BM * BK_BYTES;   // 16 KB
constexpr int SMEM_B_STAGE = BN * BK_BYTES;   // 32 KB
constexpr int SMEM_STORE_BUF = 32 * STORE_COLS * 4;  // 4 KB: 32 rows x 128 B

// Per-variant resources.
//   plain volume : 4 epilogue warps (one per TMEM lane quarter, 4 chunks each), 2 store pairs per warp, 3 operand stages.
//   fused build  : 8 epilogue warps -- two per lane quarter, chunks {0,1} and {2,3} -- each with one store pair + one
//                  level-1 box, 2 operand stages.  The epilogue does ~2.3x the instructions per chunk (pooling of
//                  three levels), so it gets two warps per scheduler; measured, this is no faster than the 4-warp
//                  version (0.486 ms either way at config 2): the kernel runs at the speed of its WRITE PATTERN --
//                  tools/mb/mb_scatter_write.cu reproduces 0.48 ms with plain stores and no GEMM at all (level 0
//                  alone 0.31 ms; the 32-byte level-2 and 8-byte level-3 pieces cost 0.08 ms for 150 MB).
template <bool FUSED>
struct Cfg {
    static constexpr int STAGES = FUSED ? 2 : 3;
    static constexpr int EPI_WARPS = FUSED ? 8 : 4;
    static constexpr int THREADS = (2 + EPI_WARPS) * 32;
    static constexpr int STORE_BUFS = FUSED ? 2 : 4;            // level-0 boxes per epilogue warp
    static constexpr int STORE_PAIRS = STORE_BUFS / 2;
    static constexpr int WARP_BUFS = STORE_BUFS + (FUSED ? 1 : 0);   // + the level-1 box
    static constexpr int SMEM_STORE = EPI_WARPS * WARP_BUFS * SMEM_STORE_BUF;
    static constexpr int OFF_A = SMEM_STORE;
    static constexpr int OFF_B = OFF_A + STAGES * SMEM_A_STAGE;
    static constexpr int OFF_BAR = OFF_B + STAGES * SMEM_B_STAGE;
    static constexpr int SMEM_TOTAL = OFF_BAR + 128;            // + barriers and the TMEM slot
    static_assert(SMEM_TOTAL <= 227 * 1024, "shared memory budget");
    static_assert((OFF_A % 1024) == 0, "operand stages must stay 1024-byte aligned (128B swizzle)");
};
static_assert(BN == 256 && STORE_COLS == 32, "the epilogue is written for 4 x 64-column chunks");

struct GemmParams {
    int N;          // rows of the volume per batch item (queries)
    int Ncols;      // columns (targets): N, or the padded tile-major count for the tiled layout
    int B;
    int num_kb;     // K blocks of 128 bytes
    int tiles_m, tiles_n;
    float scale;    // 1/sqrt(D) (exact when D is a power of 4) or 1 when divide != 0
    float divisor;  // sqrt(D)
    int use_div;    // 1: fp32 divide like the reference; 0: multiply by the exact reciprocal
    float* out;     // only used by the non-TMA epilogue (N % 4 != 0, backward GEMMs)
    int row_offset;       // first A row (query) of this launch: the chunked build computes rows [row_offset, row_offset + N)
    int out_transposed;   // non-TMA epilogue: write C^T, i.e. out[b][col][row] with row pitch ldc (backward GEMMs)
    int64_t ldc;
    const uint32_t* amax_bits;   // fp16 operands: bit patterns of max|fmap1[b]| at [b], max|fmap2[b]| at [B + b]; else null
};

// Power-of-two block scaling of the fp16 operands (one exponent per batch item and operand).  The pre-pass multiplies a
// feature map by 2^-k with k chosen so that its largest magnitude lands in [2^13, 2^14) -- inside fp16's range whatever
// the activations' scale (|x| up to ~1e33 or down to ~1e-25) -- and the GEMM epilogue multiplies the accumulator by
// 2^(k1 + k2).  Both are exact, so for data that was in range anyway the result is unchanged except that values that
// used to fall into fp16's subnormals now keep their 10 bits.
__host__ __device__ __forceinline__ int scale_exponent(uint32_t amax_bits) {
    const int e = (int)((amax_bits >> 23) & 0xffu);
    if (e == 0 || e == 255) return 0;                 // zero / subnormal / inf / nan: leave the data alone
    const int k = (e - 127) - 13;
    return k < -96 ? -96 : (k > 96 ? 96 : k);
}
__host__ __device__ __forceinline__ float exp2i(int k) {   // 2^k for k in [-126, 127]
    k = k < -126 ? -126 : (k > 127 ? 127 : k);
#ifdef __CUDA_ARCH__
    return __int_as_float((127 + k) << 23);
#else
    union { uint32_t u; float f; } c;
    c.u = (uint32_t)(127 + k) << 23;
    return c.f;
#endif
}

// Fused pyramid build (ffcorr_build_tiled_f32): the GEMM's N order is made of 16x16-pixel SUPER-GROUPS
// (4x4 tiles of 4x4 pixels; one 256-column UMMA tile each), chunk c of the epilogue = tile row c of the
// super-group, so a thread (= one query) holds everything the 2x2 poolings of levels 1..3 need in registers:
//   level 1: every level-0 tile pools to a 2x2 patch; two chunks complete two level-1 tiles (128 B)
//   level 2: one value per level-0 tile; the four chunks complete one level-2 tile (64 B)
//   level 3: one 2x2 patch per super-group (a quarter of a level-3 tile)
// The memory layout of every level stays [tile row][tile column][4][4]; only the GEMM column order changes.
struct FusedParams {
    int levels;                 // levels written (2..4)
    int sgw, sgh;               // super-groups per row / column of the level-0 map
    int th0, tw0, th1, tw1;     // tiles of levels 0 and 1 (TMA clips, these only skip empty boxes)
    int lh1, lw1, lh2, lw2, lh3, lw3;   // true level sizes (floor semantics of avg_pool2d)
    int th2, tw2, th3, tw3;
    int map2, map3;             // floats per query map, levels 2 and 3
    float* l2;
    float* l3;
    int ng;                     // GROUPED layout only: groups of 32 queries per batch item
};

// HALF (fused build only): the pyramid is STORED as fp16 (ffcorr_build_tiled_f16) -- the same 4x4-pixel tiles, 32 bytes
// each; accumulation and the poolings stay fp32, every level is rounded once (RN) when it is written.  Level 0 of a
// chunk is then 128 bytes per query (ONE store box instead of two), which halves the bytes of the kernel's bound.
// GROUPED (fused build only, fp32): the levels are stored [group of 32 queries][tile][query][4][4] (ffcorr_build_grouped_f32).
// An epilogue warp owns 32 consecutive accumulator rows = exactly one group, so what it writes per chunk is CONTIGUOUS:
// 8 KB of level 0 (4 tiles x 32 queries x 64 B), 4 KB of level 1, 1 KB / 256 B regions of levels 2 / 3 -- instead of 32
// pieces 30 KB apart.  Staged [tile][query][16 floats] in 64-byte rows (SWIZZLE_64B) and stored through 2-D tensor maps
// over the level viewed as rows of one tile-of-one-query each.
template <bool TF32, bool TMA_STORE, bool DIV, bool FUSED, bool HALF = false, bool GROUPED = false>
__global__ void __launch_bounds__(Cfg<FUSED>::THREADS, 1)
volume_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_l1,
                   const GemmParams p, const FusedParams fp, const uint32_t idesc) {
    static_assert(!FUSED || TMA_STORE, "the fused build stores through TMA");
    static_assert(!HALF || FUSED, "fp16 storage exists for the fused build only");
    static_assert(!GROUPED || (FUSED && !HALF), "the grouped layout is written by the fp32 fused build");
    using C = Cfg<FUSED>;
    constexpr int STAGES = C::STAGES;
    constexpr int STORE_PAIRS = C::STORE_PAIRS;
    // Index the extern array directly: rounding the pointer up through uintptr_t loses the shared
    // address space (generic ST.E instead of STS in the epilogue).  128B-swizzle needs 1024-byte alignment.
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t smem_base = smem_u32(smem);
    if (smem_base & 1023u) __trap();
    const uint32_t bar_base = smem_base + C::OFF_BAR;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + ACC_STAGES + s); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::OFF_BAR + 8 * (2 * STAGES + 2 * ACC_STAGES));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        if (TMA_STORE) tma_prefetch_desc(&tmap_c);
        if (FUSED) tma_prefetch_desc(&tmap_l1);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < ACC_STAGES; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), C::EPI_WARPS);
        }
        fence_barrier_init();
    } else if (warp == 1) {
        tmem_alloc(smem_u32(tmem_slot), ACC_STAGES * BN);  // 512 columns
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Tile t goes to CTA t % gridDim.x with n fastest, so at any moment the CTAs are writing ~5 adjacent
    // 128-row blocks of the volume.  Measured alternatives that keep an operand resident in shared memory
    // (B resident + walking down a 256-column strip: 0.63 ms; A resident + a contiguous run of tiles per
    // CTA: 0.52 ms, vs 0.44 ms for this order at config 2) cut the operand traffic by 33-40 % but lose more
    // through the output write pattern: the kernel is bound by how well the 1.7 GB of stores combine in
    // L2/DRAM, not by operand feed (skipping the stores entirely brings it to 0.16 ms).
    const int tiles_per_batch = p.tiles_m * p.tiles_n;
    const int num_tiles = tiles_per_batch * p.B;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int b = t / tiles_per_batch;
                const int r = t - b * tiles_per_batch;
                const int m0 = (r / p.tiles_n) * BM;
                const int n0 = (r % p.tiles_n) * BN;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    mbar_arrive_expect_tx(full_bar(stage), SMEM_A_STAGE + SMEM_B_STAGE);
                    const int k0 = kb * (TF32 ? BK_BYTES / 4 : BK_BYTES / 2);
                    tma_load_3d(smem_base + C::OFF_A + stage * SMEM_A_STAGE, &tmap_a, full_bar(stage), k0, m0 + p.row_offset, b);
                    tma_load_3d(smem_base + C::OFF_B + stage * SMEM_B_STAGE, &tmap_b, full_bar(stage), k0, n0, b);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint64_t adesc = make_smem_desc(smem_base + C::OFF_A + stage * SMEM_A_STAGE);
                    const uint64_t bdesc = make_smem_desc(smem_base + C::OFF_B + stage * SMEM_B_STAGE);
#pragma unroll
                    for (int k = 0; k < BK_BYTES / UMMA_K_BYTES; ++k) {
                        // advance K inside the swizzle atom: +32 B == +2 in the (addr >> 4) field
                        umma<TF32>(d_tmem, adesc + (uint64_t)(k * (UMMA_K_BYTES >> 4)),
                                   bdesc + (uint64_t)(k * (UMMA_K_BYTES >> 4)), idesc, (uint32_t)((kb | k) != 0));
                    }
                    umma_commit(empty_bar(stage));               // frees the smem stage when the MMAs retire
                    if (kb == p.num_kb - 1) umma_commit(tfull_bar(acc));  // accumulator ready
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===================== epilogue warps =====================
        const int quarter = warp & 3;                 // TMEM lanes [32*quarter, 32*quarter + 32)
        const int ew = warp - 2;                      // private store ring
        uint8_t* my_bufs = smem + (size_t)ew * C::WARP_BUFS * SMEM_STORE_BUF;
        const uint32_t my_bufs_u32 = smem_base + (uint32_t)(ew * C::WARP_BUFS * SMEM_STORE_BUF);
        int pair = 0;   // store ring position
        int it = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
            const int b = t / tiles_per_batch;
            const int r = t - b * tiles_per_batch;
            const int m0 = (r / p.tiles_n) * BM;
            const int nt = r % p.tiles_n;
            const int n0 = nt * BN;
            const int acc = it & 1;
            const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
            const int row0 = m0 + quarter * 32;
            // undo the operands' block scaling (fp16 mode): an exact power of two per batch item
            float unscale = 1.0f;
            if (p.amax_bits != nullptr)
                unscale = exp2i(scale_exponent(__ldg(p.amax_bits + b)) + scale_exponent(__ldg(p.amax_bits + p.B + b)));
            const float mul = DIV ? unscale : p.scale * unscale;
            // raw accumulators -> volume values, in place (fp32 bit patterns stay in the same registers)
            auto scale_chunk = [&](uint32_t (&v)[64]) {
#pragma unroll
                for (int i = 0; i < 64; ++i)
                    v[i] = __float_as_uint(DIV ? __fdiv_rn(__uint_as_float(v[i]) * mul, p.divisor) : __uint_as_float(v[i]) * mul);
            };
            auto scaled = [](uint32_t bits) { return __uint_as_float(bits); };
            // stages one 64-column chunk as two 128B-swizzled 32x32 boxes in store pair `pair`
            auto stage_pair = [&](const uint32_t (&v)[64]) {
                uint8_t* dst = my_bufs + (size_t)(pair * 2) * SMEM_STORE_BUF + lane * 128;
                if constexpr (GROUPED) {
                    // tile t of the chunk -> box t >> 1, row (t & 1) * 32 + lane of 64 bytes; 16-byte piece j = tile row j
                    uint8_t* gb = my_bufs + (size_t)(pair * 2) * SMEM_STORE_BUF;
#pragma unroll
                    for (int tt = 0; tt < 4; ++tt) {
                        const int r = (tt & 1) * 32 + lane;
                        uint8_t* rowp = gb + (tt >> 1) * SMEM_STORE_BUF + r * 64;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            *reinterpret_cast<float4*>(rowp + ((j ^ ((r >> 1) & 3)) << 4)) =
                                make_float4(scaled(v[tt * 16 + 4 * j]), scaled(v[tt * 16 + 4 * j + 1]),
                                            scaled(v[tt * 16 + 4 * j + 2]), scaled(v[tt * 16 + 4 * j + 3]));
                    }
                } else if constexpr (HALF) {
                    // 64 halfs = one 128-byte row of ONE 128B-swizzled box
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<uint4*>(dst + ((j ^ (lane & 7)) << 4)) =
                            make_uint4(pack_h2(scaled(v[8 * j]), scaled(v[8 * j + 1])), pack_h2(scaled(v[8 * j + 2]), scaled(v[8 * j + 3])),
                                       pack_h2(scaled(v[8 * j + 4]), scaled(v[8 * j + 5])), pack_h2(scaled(v[8 * j + 6]), scaled(v[8 * j + 7])));
                } else {
#pragma unroll
                    for (int hlf = 0; hlf < 2; ++hlf)
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int i0 = hlf * 32 + 4 * j;
                            *reinterpret_cast<float4*>(dst + hlf * SMEM_STORE_BUF + ((j ^ (lane & 7)) << 4)) =
                                make_float4(scaled(v[i0]), scaled(v[i0 + 1]), scaled(v[i0 + 2]), scaled(v[i0 + 3]));
                        }
                }
            };

            if constexpr (FUSED) {
                // ---- fused pyramid build: this warp owns chunks c0 (even) and c0 + 1 of its 32 queries ----
                const int c0 = (ew >> 2) * 2;
                const int sgy = nt / fp.sgw, sgx = nt - sgy * fp.sgw;
                const bool rows_live = row0 < p.N;               // warp-uniform
                const int row = row0 + lane;
                float pe[16];                                    // level-1 values of the even chunk
                float q2[8];                                     // level-2 values: [chunk parity][tile]
                uint32_t v[64];
#pragma unroll
                for (int ci = 0; ci < 2; ++ci) {
                    const int c = c0 + ci;
                    const int ty = sgy * 4 + c;                   // level-0 tile row of this chunk
                    tmem_ld_32x64(t_row + (uint32_t)(c * 64), v);
                    tmem_ld_wait();
                    if (ci == 1) {
                        // both chunks are in registers: hand the TMEM stage back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty_bar(acc));
                    }
                    if (!rows_live) continue;
                    scale_chunk(v);
                    // ---- level 1: 2x2 means inside every tile, ATen's (((a+b)+c)+d)/4 order, zero beyond the floor size ----
                    float p1[16];   // [tile t][py][px]
#pragma unroll
                    for (int tt = 0; tt < 4; ++tt)
#pragma unroll
                        for (int py = 0; py < 2; ++py)
#pragma unroll
                            for (int px = 0; px < 2; ++px) {
                                const int i0 = tt * 16 + py * 8 + px * 2;
                                const float sum = __fadd_rn(__fadd_rn(__fadd_rn(scaled(v[i0]), scaled(v[i0 + 1])), scaled(v[i0 + 4])),
                                                            scaled(v[i0 + 5]));
                                const bool ok = (ty * 2 + py < fp.lh1) && ((sgx * 4 + tt) * 2 + px < fp.lw1);
                                p1[tt * 4 + py * 2 + px] = ok ? __fmul_rn(sum, 0.25f) : 0.0f;
                            }
                    // ---- level 2: one value per level-0 tile ----
#pragma unroll
                    for (int tt = 0; tt < 4; ++tt) {
                        const float sum = __fadd_rn(__fadd_rn(__fadd_rn(p1[tt * 4], p1[tt * 4 + 1]), p1[tt * 4 + 2]), p1[tt * 4 + 3]);
                        const bool ok = (ty < fp.lh2) && (sgx * 4 + tt < fp.lw2);
                        q2[ci * 4 + tt] = ok ? __fmul_rn(sum, 0.25f) : 0.0f;
                    }
                    if (lane == 0) tma_store_wait_read<STORE_PAIRS - 1>();   // this warp's previous group has left smem
                    __syncwarp();
                    // ---- level 0: tiles (ty, 4 sgx .. 4 sgx + 3) = 256 contiguous bytes per query ----
                    const bool l0_any = ty < fp.th0;
                    if (l0_any) stage_pair(v);
                    bool l1_any = false;
                    if (ci == 1) {
                        // level-1 tiles (2 sgy + c/2, 2 sgx .. + 1): rows 0-1 from the even chunk, rows 2-3 from this one
                        const int ty1 = sgy * 2 + (c >> 1);
                        l1_any = (ty1 < fp.th1) && (sgx * 2 < fp.tw1);
                        if (l1_any) {
                            uint8_t* dst = my_bufs + (size_t)C::STORE_BUFS * SMEM_STORE_BUF + lane * (HALF ? 64 : 128);
#pragma unroll
                            for (int tt = 0; tt < 2; ++tt)
#pragma unroll
                                for (int rr = 0; rr < 4; ++rr) {
                                    const float* src = (rr < 2) ? pe : p1;
                                    const int py = rr & 1;
                                    const float4 q = make_float4(src[(2 * tt) * 4 + py * 2], src[(2 * tt) * 4 + py * 2 + 1],
                                                                 src[(2 * tt + 1) * 4 + py * 2], src[(2 * tt + 1) * 4 + py * 2 + 1]);
                                    const int j = tt * 4 + rr;
                                    if constexpr (GROUPED) {   // level-1 tile tt of this query = row tt * 32 + lane, piece rr
                                        const int r = tt * 32 + lane;
                                        *reinterpret_cast<float4*>(my_bufs + (size_t)C::STORE_BUFS * SMEM_STORE_BUF + r * 64 +
                                                                   ((rr ^ ((r >> 1) & 3)) << 4)) = q;
                                    } else if constexpr (HALF)   // two tiles x 4 rows x 8 bytes = 64 bytes per query, unswizzled box
                                        *reinterpret_cast<uint2*>(dst + j * 8) = make_uint2(pack_h2(q.x, q.y), pack_h2(q.z, q.w));
                                    else
                                        *reinterpret_cast<float4*>(dst + ((j ^ (lane & 7)) << 4)) = q;
                                }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) pe[i] = p1[i];
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (GROUPED) {
                            const int grp = b * fp.ng + (row0 >> 5);
                            if (l0_any) {
                                const uint32_t src = my_bufs_u32 + (uint32_t)(pair * 2 * SMEM_STORE_BUF);
                                const int r0 = (grp * (fp.th0 * fp.tw0) + ty * fp.tw0 + 4 * sgx) * 32;     // < 2^31 (host check)
                                tma_store_2d(&tmap_c, src, 0, r0);
                                if (sgx * 4 + 2 < fp.tw0) tma_store_2d(&tmap_c, src + SMEM_STORE_BUF, 0, r0 + 64);
                            }
                            if (l1_any)
                                tma_store_2d(&tmap_l1, my_bufs_u32 + (uint32_t)(C::STORE_BUFS * SMEM_STORE_BUF), 0,
                                             (grp * (fp.th1 * fp.tw1) + (sgy * 2 + (c >> 1)) * fp.tw1 + 2 * sgx) * 32);
                        } else {
                            if (l0_any) {
                                const uint32_t src = my_bufs_u32 + (uint32_t)(pair * 2 * SMEM_STORE_BUF);
                                tma_store_4d(&tmap_c, src, sgx * 64, ty, row0, b);
                                if (!HALF && sgx * 4 + 2 < fp.tw0) tma_store_4d(&tmap_c, src + SMEM_STORE_BUF, sgx * 64 + 32, ty, row0, b);
                            }
                            if (l1_any)
                                tma_store_4d(&tmap_l1, my_bufs_u32 + (uint32_t)(C::STORE_BUFS * SMEM_STORE_BUF), sgx * 32,
                                             sgy * 2 + (c >> 1), row0, b);
                        }
                        tma_store_commit();
                    }
                    pair = (pair + 1) % STORE_PAIRS;
                    if (ci == 1 && row < p.N) {
                        // ---- level 2: rows c0, c0 + 1 of this super-group's tile, 32 contiguous bytes per query ----
                        if (fp.levels >= 3 && sgy < fp.th2 && sgx < fp.tw2) {
                            const int64_t e2 = GROUPED
                                ? ((((int64_t)b * fp.ng + (row >> 5)) * (fp.th2 * fp.tw2) + sgy * fp.tw2 + sgx) * 32 + (row & 31)) * 16 + c0 * 4
                                : ((int64_t)b * p.N + row) * fp.map2 + (sgy * fp.tw2 + sgx) * 16 + c0 * 4;
                            if constexpr (HALF) {
                                *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(fp.l2) + e2) =
                                    make_uint4(pack_h2(q2[0], q2[1]), pack_h2(q2[2], q2[3]), pack_h2(q2[4], q2[5]), pack_h2(q2[6], q2[7]));
                            } else {
                                float4* d2 = reinterpret_cast<float4*>(fp.l2 + e2);
                                d2[0] = make_float4(q2[0], q2[1], q2[2], q2[3]);
                                d2[1] = make_float4(q2[4], q2[5], q2[6], q2[7]);
                            }
                        }
                        // ---- level 3: one row of the super-group's 2x2 patch, 8 bytes per query ----
                        const int ty3 = sgy >> 1, tx3 = sgx >> 1;
                        if (fp.levels >= 4 && ty3 < fp.th3 && tx3 < fp.tw3) {
                            float2 o;
                            o.x = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(q2[0], q2[1]), q2[4]), q2[5]), 0.25f);
                            o.y = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(q2[2], q2[3]), q2[6]), q2[7]), 0.25f);
                            const int y3 = sgy * 2 + (c >> 1);
                            if (!(y3 < fp.lh3 && sgx * 2 < fp.lw3)) o.x = 0.0f;
                            if (!(y3 < fp.lh3 && sgx * 2 + 1 < fp.lw3)) o.y = 0.0f;
                            const int64_t e3 = (GROUPED
                                ? ((((int64_t)b * fp.ng + (row >> 5)) * (fp.th3 * fp.tw3) + ty3 * fp.tw3 + tx3) * 32 + (row & 31)) * 16
                                : ((int64_t)b * p.N + row) * fp.map3 + (ty3 * fp.tw3 + tx3) * 16) +
                                               ((sgy & 1) * 2 + (c >> 1)) * 4 + (sgx & 1) * 2;
                            // quarters of this level-3 tile whose super-group does not exist stay exact zeros
                            const bool ghost_x = (sgx == fp.sgw - 1) && !(sgx & 1);
                            const bool ghost_y = (sgy == fp.sgh - 1) && !(sgy & 1);
                            if constexpr (HALF) {
                                uint32_t* d3 = reinterpret_cast<uint32_t*>(reinterpret_cast<__half*>(fp.l3) + e3);   // 2 halfs
                                d3[0] = pack_h2(o.x, o.y);
                                if (ghost_x) d3[1] = 0u;
                                if (ghost_y) d3[4] = 0u;
                                if (ghost_x && ghost_y) d3[5] = 0u;
                            } else {
                                float* d3 = fp.l3 + e3;
                                *reinterpret_cast<float2*>(d3) = o;
                                const float2 z = make_float2(0.0f, 0.0f);
                                if (ghost_x) *reinterpret_cast<float2*>(d3 + 2) = z;
                                if (ghost_y) *reinterpret_cast<float2*>(d3 + 8) = z;
                                if (ghost_x && ghost_y) *reinterpret_cast<float2*>(d3 + 10) = z;
                            }
                        }
                    }
                }
            } else {
                // ---- plain volume: 4 chunks of 64 columns, software pipelined: the TMEM load of chunk c+1 is in
                // flight while chunk c is scaled, staged (two boxes, ONE proxy fence) and handed to TMA ----
                auto process = [&](uint32_t (&v)[64], int c) {
                    const int col0 = n0 + c * 64;
                    if (TMA_STORE) {
                        if (row0 >= p.N || col0 >= p.Ncols) return;   // warp-uniform: both boxes outside
                        if (lane == 0) tma_store_wait_read<STORE_PAIRS - 1>();   // the pair used STORE_PAIRS groups ago has been read
                        __syncwarp();
                        scale_chunk(v);
                        stage_pair(v);
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            const uint32_t src = my_bufs_u32 + (uint32_t)(pair * 2 * SMEM_STORE_BUF);
                            tma_store_3d(&tmap_c, src, col0, row0, b);
                            if (col0 + 32 < p.Ncols) tma_store_3d(&tmap_c, src + SMEM_STORE_BUF, col0 + 32, row0, b);
                            tma_store_commit();
                        }
                        pair = (pair + 1) % STORE_PAIRS;
                    } else {
                        const int row = row0 + lane;
                        if (row < p.N) {
                            scale_chunk(v);
                            if (p.out_transposed) {
                                // 32 lanes = 32 consecutive rows -> every store instruction writes one 128-byte line
                                float* o = p.out + ((int64_t)b * p.Ncols + col0) * p.ldc + row;
#pragma unroll
                                for (int i = 0; i < 64; ++i)
                                    if (col0 + i < p.Ncols) o[(int64_t)i * p.ldc] = scaled(v[i]);
                            } else {
                                float* o = p.out + ((int64_t)b * p.N + row) * (int64_t)p.Ncols + col0;
#pragma unroll
                                for (int i = 0; i < 64; ++i)
                                    if (col0 + i < p.Ncols) o[i] = scaled(v[i]);
                            }
                        }
                    }
                };
                uint32_t va[64], vb[64];
                tmem_ld_32x64(t_row, va);
                tmem_ld_wait();
                tmem_ld_32x64(t_row + 64u, vb);
                process(va, 0);
                tmem_ld_wait();
                tmem_ld_32x64(t_row + 128u, va);
                process(vb, 1);
                tmem_ld_wait();
                tmem_ld_32x64(t_row + 192u, vb);
                process(va, 2);
                tmem_ld_wait();
                // accumulator fully drained into registers: hand the TMEM stage back to the MMA warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(acc));
                process(vb, 3);
            }
        }
        if (TMA_STORE && lane == 0) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, ACC_STAGES * BN);
    }
}

// ------------------------------------------------------------------------------------------
// operand pre-pass: [B, D, N] fp32 -> [B, N, Dp] K-major in the MMA operand type
// ------------------------------------------------------------------------------------------
enum : int { CVT_F16 = 0, CVT_BF16X3_A = 1, CVT_BF16X3_B = 2, CVT_F32 = 3 };

constexpr int PP_CH = 16;        // channels per thread: one 32-byte K-major piece per pixel in the 16-bit modes
constexpr int PP_THREADS = 128;

__device__ __forceinline__ uint32_t pack_bf2(__nv_bfloat16 a, __nv_bfloat16 b) {
    __nv_bfloat162 h;
    h.x = a;
    h.y = b;
    return *reinterpret_cast<uint32_t*>(&h);
}

// Transposing convert, register-only.  A thread owns 4 consecutive output rows (pixels) x 16 channels: it issues 16
// independent 16-byte loads along the pixel axis (a warp reads 512 contiguous bytes of one channel plane), transposes
// by register naming, and writes one 32-byte K-major piece per pixel.  (The first version staged 32 x 128 tiles in
// shared memory and was bound by its scalar shared-memory column reads: 48 us / 19.8 M instructions at config 2.)
// Tiled ("T4") target order for the B operand: row n' of the staged operand is pixel
// (y, x) = (4*ty + iy, 4*tx + ix) with n' = (ty*TW + tx)*16 + iy*4 + ix, zero rows for the padding, so the GEMM
// writes every query's map directly as 4x4-pixel tiles of 64 contiguous bytes.
// Super-group order (enabled == 2, fused pyramid build): n' = sg*256 + c*64 + t*16 + iy*4 + ix is pixel
// (y, x) = (16*sgy + 4*c + iy, 16*sgx + 4*t + ix), sg = sgy*SGW + sgx; padding up to multiples of 16.
struct TiledB {
    int enabled;   // 0: row-major pixels, 1: 4x4 tiles in row-major tile order, 2: 16x16 super-groups of 4x4 tiles
    int h, w;      // feature-map size
    int tw;        // enabled == 1: tiles per row (tiled_tw); enabled == 2: super-groups per row, ceil(w / 16)
    int np;        // padded pixel count = rows of the staged operand
};

// max |x| of every feature map (batch item x operand) as a float bit pattern (monotonic for non-negative floats):
// amax_bits[which * B + b].  The maps were just written by the encoder, so most of this read hits L2.
__global__ void __launch_bounds__(256) operand_amax_kernel(const float* __restrict__ f1, const float* __restrict__ f2,
                                                           uint32_t* __restrict__ amax_bits, int B, int64_t per_item) {
    const int which = blockIdx.y / B, b = blockIdx.y - which * B;
    const float* __restrict__ in = (which == 0 ? f1 : f2) + (int64_t)b * per_item;
    uint32_t m = 0;
    const bool vec = (per_item & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
    if (vec) {
        const float4* in4 = reinterpret_cast<const float4*>(in);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_item / 4; i += (int64_t)gridDim.x * blockDim.x) {
            const float4 v = __ldg(in4 + i);
            m = max(max(m, __float_as_uint(fabsf(v.x))), max(__float_as_uint(fabsf(v.y)), max(__float_as_uint(fabsf(v.z)), __float_as_uint(fabsf(v.w)))));
        }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_item; i += (int64_t)gridDim.x * blockDim.x)
            m = max(m, __float_as_uint(fabsf(__ldg(in + i))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m != 0) atomicMax(amax_bits + blockIdx.y, m);
}

template <int MODE>
__global__ void __launch_bounds__(PP_THREADS) operand_prepass_kernel(const float* __restrict__ f1, const float* __restrict__ f2,
                                                                     void* __restrict__ o1, void* __restrict__ o2,
                                                                     int B, int D, int N, int Dp /* padded K per segment */,
                                                                     const TiledB tb, const uint32_t* __restrict__ amax_bits = nullptr) {
    const int which = blockIdx.z / B;          // 0: fmap1 (A operand), 1: fmap2 (B operand)
    const int b = blockIdx.z - which * B;
    const float* __restrict__ in = (which == 0 ? f1 : f2) + (size_t)b * D * N;
    void* __restrict__ outp = which == 0 ? o1 : o2;
    const bool tiled = tb.enabled && which == 1;
    const int Nout = tiled ? tb.np : N;        // rows of the staged operand
    const int n = (blockIdx.x * PP_THREADS + threadIdx.x) * 4;   // first of this thread's 4 output rows
    if (n >= Nout) return;
    const int d0 = blockIdx.y * PP_CH;

    int src_n = n, lim = N - n;                // source pixel of output row n, and how many of the 4 exist
    if (tiled) {                               // n is a multiple of 4: one row of one 4x4 tile
        int y, x;
        if (tb.enabled == 2) {
            const int sg = n >> 8, c = (n >> 6) & 3, t = (n >> 4) & 3, iy = (n >> 2) & 3;
            const int sgy = sg / tb.tw, sgx = sg - sgy * tb.tw;
            y = 16 * sgy + 4 * c + iy;
            x = 16 * sgx + 4 * t;
        } else {
            const int t = n >> 4, iy = (n >> 2) & 3;
            const int ty = t / tb.tw, tx = t - ty * tb.tw;
            y = 4 * ty + iy;
            x = 4 * tx;
        }
        src_n = y * tb.w + x;
        lim = (y < tb.h) ? tb.w - x : 0;
    }
    const bool vec_ok = ((N & 3) == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0) && (!tiled || (tb.w & 3) == 0);
    // fp16 mode: exact power-of-two block scaling into fp16's range (see scale_exponent)
    const float pre = (MODE == CVT_F16 && amax_bits != nullptr) ? exp2i(-scale_exponent(__ldg(amax_bits + which * B + b))) : 1.0f;

    float4 v[PP_CH];                           // v[k] = channel d0 + k at the 4 pixels
#pragma unroll
    for (int k = 0; k < PP_CH; ++k) {
        v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int d = d0 + k;
        if (d < D && lim > 0) {
            const float* src = in + (size_t)d * N + src_n;
            if (vec_ok && lim >= 4) {
                v[k] = __ldg(reinterpret_cast<const float4*>(src));
            } else {
                v[k].x = __ldg(src + 0);
                if (lim > 1) v[k].y = __ldg(src + 1);
                if (lim > 2) v[k].z = __ldg(src + 2);
                if (lim > 3) v[k].w = __ldg(src + 3);
            }
        }
    }
    const int rows = min(4, Nout - n);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j >= rows) break;
        float x[PP_CH];
#pragma unroll
        for (int k = 0; k < PP_CH; ++k) {
            x[k] = j == 0 ? v[k].x : (j == 1 ? v[k].y : (j == 2 ? v[k].z : v[k].w));
            if (MODE == CVT_F16) x[k] *= pre;
        }
        const size_t row = (size_t)b * Nout + n + j;
        if (MODE == CVT_F16) {
            __half* o = reinterpret_cast<__half*>(outp) + row * Dp + d0;
            reinterpret_cast<uint4*>(o)[0] = make_uint4(pack_h2(x[0], x[1]), pack_h2(x[2], x[3]), pack_h2(x[4], x[5]), pack_h2(x[6], x[7]));
            reinterpret_cast<uint4*>(o)[1] = make_uint4(pack_h2(x[8], x[9]), pack_h2(x[10], x[11]), pack_h2(x[12], x[13]), pack_h2(x[14], x[15]));
        } else if (MODE == CVT_F32) {
            float* o = reinterpret_cast<float*>(outp) + row * Dp + d0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                reinterpret_cast<float4*>(o)[k] = make_float4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
        } else {
            // bf16 hi/lo split; A rows hold [hi | hi | lo], B rows [hi | lo | hi] so that one
            // K = 3*Dp GEMM computes hi*hi + hi*lo + lo*hi with fp32 accumulation.
            __nv_bfloat16 hi[PP_CH], lo[PP_CH];
#pragma unroll
            for (int k = 0; k < PP_CH; ++k) {
                hi[k] = __float2bfloat16_rn(x[k]);
                lo[k] = __float2bfloat16_rn(x[k] - __bfloat162float(hi[k]));
            }
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(outp) + row * (3 * (size_t)Dp) + d0;
#pragma unroll
            for (int seg = 0; seg < 3; ++seg) {
                const bool use_lo = (which == 0) ? (seg == 2) : (seg == 1);
                const __nv_bfloat16* q = use_lo ? lo : hi;
                reinterpret_cast<uint4*>(o + (size_t)seg * Dp)[0] =
                    make_uint4(pack_bf2(q[0], q[1]), pack_bf2(q[2], q[3]), pack_bf2(q[4], q[5]), pack_bf2(q[6], q[7]));
                reinterpret_cast<uint4*>(o + (size_t)seg * Dp)[1] =
                    make_uint4(pack_bf2(q[8], q[9]), pack_bf2(q[10], q[11]), pack_bf2(q[12], q[13]), pack_bf2(q[14], q[15]));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// CUDA-core SGEMM with arbitrary operand strides (exact fp32 path + backward GEMMs)
//   C[m, n] = (sum_k A[m,k] * B[k,n]) / divisor,   batched over blockIdx.z
// ------------------------------------------------------------------------------------------
constexpr int SG_T = 64, SG_K = 16;

template <bool A_MCONTIG, bool B_NCONTIG>
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ C,
                                                    int M, int Nn, int K, int64_t a_sm, int64_t a_sk, int64_t b_sk, int64_t b_sn,
                                                    int64_t c_ld, int64_t a_batch, int64_t b_batch, int64_t c_batch, float divisor) {
    __shared__ float As[SG_K][SG_T + 4];
    __shared__ float Bs[SG_K][SG_T + 4];
    A += (int64_t)blockIdx.z * a_batch;
    Bm += (int64_t)blockIdx.z * b_batch;
    C += (int64_t)blockIdx.z * c_batch;
    const int m0 = blockIdx.y * SG_T, n0 = blockIdx.x * SG_T;
    const int t = threadIdx.x;
    const int tx = t & 15, ty = t >> 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += SG_K) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int m, k;
            if (A_MCONTIG) { m = t & 63; k = (t >> 6) + 4 * i; } else { k = t & 15; m = (t >> 4) + 16 * i; }
            const int gm = m0 + m, gk = k0 + k;
            As[k][m] = (gm < M && gk < K) ? __ldg(A + gm * a_sm + gk * a_sk) : 0.0f;
            int n, kk;
            if (B_NCONTIG) { n = t & 63; kk = (t >> 6) + 4 * i; } else { kk = t & 15; n = (t >> 4) + 16 * i; }
            const int gn = n0 + n, gkb = k0 + kk;
            Bs[kk][n] = (gn < Nn && gkb < K) ? __ldg(Bm + gkb * b_sk + gn * b_sn) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SG_K; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn < Nn) C[gm * c_ld + gn] = __fdiv_rn(acc[i][j], divisor);
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
int encode_3d(CUtensorMap* m, CUtensorMapDataType dt, int elem_bytes, void* base, uint64_t d0, uint64_t d1, uint64_t d2,
              uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, const char* what) {
    (void)elem_bytes;
    const uint64_t dims[3] = {d0, d1, d2};
    const uint64_t strides[2] = {stride1_bytes, stride2_bytes};
    const uint32_t box[3] = {b0, b1, 1};
    return encode_tensor_map(m, dt, 3, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, what);
}

struct PrecInfo {
    int elem_bytes;   // operand element size
    int k_mult;       // K multiplier (3 for the bf16 split)
    int k_align;      // per-segment K padding in elements (one 128 B swizzle atom)
};

bool prec_info(int precision, PrecInfo* pi) {
    switch (precision) {
        case FFCORR_PREC_FP16:   *pi = {2, 1, 64}; return true;
        case FFCORR_PREC_BF16X3: *pi = {2, 3, 64}; return true;
        case FFCORR_PREC_TF32:   *pi = {4, 1, 32}; return true;
        default: return false;
    }
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <typename K>
int set_smem(K kernel, int bytes) {
    FFCORR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return FFCORR_OK;
}

// In-place transpose of B square [n x n] fp32 matrices (the level-0 gradient before the second backward GEMM):
// block (ti, tj), ti <= tj, swaps tiles (ti, tj) and (tj, ti) through shared memory.
__global__ void __launch_bounds__(256) transpose_inplace_kernel(float* __restrict__ g, int n, int tiles) {
    __shared__ float ta[32][33], tb[32][33];
    // linear block index -> (ti <= tj) of the upper triangle
    int k = blockIdx.x, ti = 0;
    while (k >= tiles - ti) { k -= tiles - ti; ++ti; }
    const int tj = ti + k;
    float* m = g + (size_t)blockIdx.y * n * n;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int ia = ti * 32 + r, ja = tj * 32 + tx;          // element of tile (ti, tj)
        const int ib = tj * 32 + r, jb = ti * 32 + tx;          // element of tile (tj, ti)
        ta[r][tx] = (ia < n && ja < n) ? m[(size_t)ia * n + ja] : 0.f;
        tb[r][tx] = (ib < n && jb < n) ? m[(size_t)ib * n + jb] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int ia = ti * 32 + r, ja = tj * 32 + tx;
        const int ib = tj * 32 + r, jb = ti * 32 + tx;
        if (ia < n && ja < n) m[(size_t)ia * n + ja] = tb[tx][r];
        if (ti != tj && ib < n && jb < n) m[(size_t)ib * n + jb] = ta[tx][r];
    }
}

// C^T[b][n][m] = (sum_k A[b][m][k] * Bm[b][n][k]) / divisor on the tensor cores (kind::tf32 straight from the fp32
// arrays: both operands are K-major as they lie in memory, so there is no pre-pass).  K*4 must be a multiple of 16.
int launch_gemm_nt_tf32(const float* A, const float* Bm, float* Ct, int M, int Nn, int K, int batches, float divisor,
                        cudaStream_t s);

int launch_sgemm(bool a_mcontig, bool b_ncontig, const float* A, const float* Bm, float* C, int M, int Nn, int K,
                 int64_t a_sm, int64_t a_sk, int64_t b_sk, int64_t b_sn, int64_t c_ld, int64_t a_batch, int64_t b_batch,
                 int64_t c_batch, int batches, float divisor, cudaStream_t s) {
    dim3 grid(ceil_div(Nn, SG_T), ceil_div(M, SG_T), batches);
    FFCORR_REQUIRE(grid.y < 65536 && grid.z < 65536, FFCORR_EINVAL, "sgemm: grid too large");
#define FF_SGEMM(AM, BN_)                                                                                           \
    sgemm_kernel<AM, BN_><<<grid, 256, 0, s>>>(A, Bm, C, M, Nn, K, a_sm, a_sk, b_sk, b_sn, c_ld, a_batch, b_batch, \
                                               c_batch, divisor)
    if (a_mcontig && b_ncontig) FF_SGEMM(true, true);
    else if (a_mcontig) FF_SGEMM(true, false);
    else if (b_ncontig) FF_SGEMM(false, true);
    else FF_SGEMM(false, false);
#undef FF_SGEMM
    return check_launch("sgemm_kernel");
}

}  // namespace
}  // namespace ffcorr

using namespace ffcorr;

extern "C" size_t ffcorr_volume_workspace_bytes(int B, int D, int h, int w, int precision) {
    PrecInfo pi;
    if (!prec_info(precision, &pi) || B <= 0 || D <= 0 || h <= 0 || w <= 0) return 0;
    const size_t Np = align_up((size_t)h, 16) * align_up((size_t)w, 16); // >= h*w; covers the tiled and super-group orders too
    const size_t Dp = align_up((size_t)D, pi.k_align);
    const size_t one = align_up((size_t)B * Np * Dp * pi.k_mult * pi.elem_bytes, 256);
    // fp16 operands: + the block-scaling exponents' source, max|fmap| per batch item and operand (2*B words)
    return 2 * one + (precision == FFCORR_PREC_FP16 ? align_up((size_t)2 * B * sizeof(uint32_t), 256) : 0);
}

enum : int { OUT_ROWMAJOR = 0, OUT_TILED = 1, OUT_FUSED_PYRAMID = 2 };
enum : int { PHASE_ALL = 0, PHASE_STAGE = 1, PHASE_GEMM = 2 };

static int volume_impl(const float* fmap1, const float* fmap2, float* lvl0, int B, int D, int h, int w, int precision,
                       void* workspace, size_t workspace_bytes, void* stream, int out_mode, float* const* lvl = nullptr,
                       int num_levels = 1, int q0 = 0, int nq = -1, int phase = PHASE_ALL, float divisor = 0.0f,
                       bool out_half = false, bool grouped = false) {
    // q0 / nq: only the queries [q0, q0 + nq) are computed (chunked build); phase: stage the operands, run the
    // GEMM on already staged operands, or both.
    const bool tiled = out_mode != OUT_ROWMAJOR;
    const bool fused = out_mode == OUT_FUSED_PYRAMID;
    FFCORR_REQUIRE(B >= 0 && D >= 1 && h >= 1 && w >= 1, FFCORR_EINVAL, "volume: bad shape B=%d D=%d h=%d w=%d", B, D, h, w);
    FFCORR_REQUIRE((int64_t)h * w < (1ll << 24), FFCORR_EINVAL, "volume: h*w=%lld too large", (long long)h * w);
    if (B == 0) return FFCORR_OK;  // empty batch: pointers may legitimately be null
    FFCORR_REQUIRE((phase == PHASE_GEMM || (fmap1 && fmap2)) && (phase == PHASE_STAGE || lvl0), FFCORR_EINVAL, "volume: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const int N = h * w;
    const int Nq = nq < 0 ? N : nq;               // queries computed by this call
    FFCORR_REQUIRE(q0 >= 0 && Nq >= 1 && q0 + Nq <= N, FFCORR_EINVAL, "volume: query range [%d, %d) outside [0, %d)", q0, q0 + Nq, N);
    const float sqrt_d = divisor > 0.0f ? divisor : sqrtf((float)D);   // the value every dot product is divided by

    FFCORR_REQUIRE(!(tiled && precision == FFCORR_PREC_FP32), FFCORR_EINVAL,
                   "volume: the tiled layout is produced by the tensor-core paths only");
    if (precision == FFCORR_PREC_FP32) {
        // C[i,j] = sum_d f1[d,i] f2[d,j]: A[m=i,k=d] (m contiguous), B[k=d,n=j] (n contiguous)
        return launch_sgemm(true, true, fmap1, fmap2, lvl0, N, N, D, 1, N, N, 1, N, (int64_t)D * N, (int64_t)D * N,
                            (int64_t)N * N, B, sqrt_d, s);
    }

    PrecInfo pi;
    FFCORR_REQUIRE(prec_info(precision, &pi), FFCORR_EINVAL, "volume: unknown precision %d", precision);
    const size_t need = ffcorr_volume_workspace_bytes(B, D, h, w, precision);
    FFCORR_REQUIRE(workspace != nullptr && workspace_bytes >= need, FFCORR_EWORKSPACE,
                   "volume: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
    FFCORR_REQUIRE((uintptr_t)workspace % 256 == 0, FFCORR_EALIGN, "volume: workspace must be 256-byte aligned");
    FFCORR_REQUIRE(phase == PHASE_STAGE || (uintptr_t)lvl0 % 16 == 0, FFCORR_EALIGN, "volume: lvl0 must be 16-byte aligned");

    const int Dp = (int)align_up((size_t)D, pi.k_align);
    const int Kt = Dp * pi.k_mult;  // total K in elements
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    const size_t amax_bytes = precision == FFCORR_PREC_FP16 ? align_up((size_t)2 * B * sizeof(uint32_t), 256) : 0;
    const size_t one_op = (need - amax_bytes) / 2;
    void* opA = ws;
    void* opB = ws + one_op;
    uint32_t* amax_bits = amax_bytes ? reinterpret_cast<uint32_t*>(ws + 2 * one_op) : nullptr;

    // ---- 1. operand pre-pass ----
    TiledB tlb{};
    tlb.enabled = fused ? 2 : (tiled ? 1 : 0);
    tlb.h = h;
    tlb.w = w;
    tlb.tw = fused ? ceil_div(w, 16) : tiled_tw(w);
    tlb.np = fused ? ceil_div(h, 16) * tlb.tw * 256 : tiled_th(h) * tlb.tw * 16;
    const int Ncols = tiled ? tlb.np : N;       // columns of the volume == rows of the staged B operand
    if (phase != PHASE_GEMM) {
        // rows of the larger operand in groups of 4 per thread; row-major A (N rows) and B (Ncols >= N rows) share the grid
        dim3 grid(ceil_div(ceil_div(Ncols > N ? Ncols : N, 4), PP_THREADS), Dp / PP_CH, 2 * B);
        FFCORR_REQUIRE(grid.y < 65536 && grid.z < 65536, FFCORR_EINVAL, "volume: pre-pass grid too large");
        if (precision == FFCORR_PREC_FP16) {
            FFCORR_CUDA(cudaMemsetAsync(amax_bits, 0, (size_t)2 * B * sizeof(uint32_t), s));
            const int64_t per_item = (int64_t)D * N;
            const int bx = (int)std::min<int64_t>(ceil_div64(per_item, 256 * 4 * 4), 64);
            operand_amax_kernel<<<dim3((unsigned)bx, (unsigned)(2 * B)), 256, 0, s>>>(fmap1, fmap2, amax_bits, B, per_item);
            if (int rc = check_launch("operand_amax_kernel")) return rc;
            operand_prepass_kernel<CVT_F16><<<grid, PP_THREADS, 0, s>>>(fmap1, fmap2, opA, opB, B, D, N, Dp, tlb, amax_bits);
        }
        else if (precision == FFCORR_PREC_TF32)
            operand_prepass_kernel<CVT_F32><<<grid, PP_THREADS, 0, s>>>(fmap1, fmap2, opA, opB, B, D, N, Dp, tlb);
        else
            operand_prepass_kernel<CVT_BF16X3_A><<<grid, PP_THREADS, 0, s>>>(fmap1, fmap2, opA, opB, B, D, N, Dp, tlb);
        if (int rc = check_launch("operand_prepass_kernel")) return rc;
    }
    if (phase == PHASE_STAGE) return FFCORR_OK;

    // ---- 2. tensor maps ----
    CUtensorMap ta, tb, tc, tl1;
    const bool tf32 = precision == FFCORR_PREC_TF32;
    const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                        : (precision == FFCORR_PREC_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                                                         : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
    const uint32_t box_k = (uint32_t)(BK_BYTES / pi.elem_bytes);
    const uint64_t row_bytes = (uint64_t)Kt * pi.elem_bytes;
    if (int rc = encode_3d(&ta, dt, pi.elem_bytes, opA, Kt, N, B, row_bytes, row_bytes * N, box_k, BM, "A")) return rc;
    if (int rc = encode_3d(&tb, dt, pi.elem_bytes, opB, Kt, Ncols, B, row_bytes, row_bytes * Ncols, box_k, BN, "B")) return rc;
    const bool tma_store = (Ncols % 4 == 0);
    FusedParams fp{};
    if (fused) {
        // 4-D views [B][N][tile row][tile-row floats] of levels 0 and 1; boxes of 32 queries x 128 bytes
        const int th0 = tiled_th(h), tw0 = tiled_tw(w);
        const int lh1 = h >> 1, lw1 = w >> 1;
        const int th1 = tiled_th(lh1), tw1 = tiled_tw(lw1);
        // fp32 storage: boxes of 32 queries x 32 floats (128 B); fp16 storage: level 0 in boxes of 32 queries x 64 halfs
        // (128 B, one per chunk), level 1 in unswizzled boxes of 32 queries x 32 halfs (64 B)
        const uint64_t eb = out_half ? 2 : 4;
        const CUtensorMapDataType odt = out_half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
        if (grouped) {
            // levels 0 and 1 as 2-D arrays of 64-byte rows (one tile of one query each): [B*NG*th*tw*32][16 floats]
            FFCORR_REQUIRE(!out_half, FFCORR_EINVAL, "volume: the grouped layout stores fp32");
            const int64_t ng = ceil_div(Nq, 32);
            const int64_t rows0 = (int64_t)B * ng * th0 * tw0 * 32, rows1 = (int64_t)B * ng * th1 * tw1 * 32;
            FFCORR_REQUIRE(rows0 < (1ll << 31), FFCORR_EINVAL, "volume: grouped level 0 has too many tile rows (%lld)", (long long)rows0);
            const uint32_t box[2] = {16, 64};
            const uint64_t strides[1] = {64};
            const uint64_t d0[2] = {16, (uint64_t)rows0}, d1[2] = {16, (uint64_t)rows1};
            if (int rc = encode_tensor_map(&tc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, lvl[0], d0, strides, box, CU_TENSOR_MAP_SWIZZLE_64B, "G0"))
                return rc;
            if (int rc = encode_tensor_map(&tl1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, lvl[1], d1, strides, box, CU_TENSOR_MAP_SWIZZLE_64B, "G1"))
                return rc;
            fp.ng = (int)ng;
        } else {
        {
            const uint32_t box[4] = {out_half ? 64u : 32u, 1, 32, 1};
            const uint64_t dims[4] = {(uint64_t)tw0 * 16, (uint64_t)th0, (uint64_t)Nq, (uint64_t)B};
            const uint64_t map_b = (uint64_t)th0 * tw0 * 16 * eb;
            const uint64_t strides[3] = {(uint64_t)tw0 * 16 * eb, map_b, map_b * Nq};
            if (int rc = encode_tensor_map(&tc, odt, 4, lvl[0], dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, "L0"))
                return rc;
        }
        {
            const uint32_t box[4] = {32, 1, 32, 1};
            const uint64_t dims[4] = {(uint64_t)tw1 * 16, (uint64_t)th1, (uint64_t)Nq, (uint64_t)B};
            const uint64_t map_b = (uint64_t)th1 * tw1 * 16 * eb;
            const uint64_t strides[3] = {(uint64_t)tw1 * 16 * eb, map_b, map_b * Nq};
            if (int rc = encode_tensor_map(&tl1, odt, 4, lvl[1], dims, strides, box,
                                           out_half ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B, "L1"))
                return rc;
        }
        }
        fp.levels = num_levels;
        fp.sgw = tlb.tw;
        fp.sgh = ceil_div(h, 16);
        fp.th0 = th0; fp.tw0 = tw0; fp.th1 = th1; fp.tw1 = tw1;
        fp.lh1 = lh1; fp.lw1 = lw1;
        fp.lh2 = h >> 2; fp.lw2 = w >> 2;
        fp.lh3 = h >> 3; fp.lw3 = w >> 3;
        fp.th2 = tiled_th(fp.lh2); fp.tw2 = tiled_tw(fp.lw2);
        fp.th3 = tiled_th(fp.lh3); fp.tw3 = tiled_tw(fp.lw3);
        fp.map2 = fp.th2 * fp.tw2 * 16;
        fp.map3 = fp.th3 * fp.tw3 * 16;
        fp.l2 = num_levels >= 3 ? lvl[2] : nullptr;
        fp.l3 = num_levels >= 4 ? lvl[3] : nullptr;
    } else if (tma_store) {
        if (int rc = encode_3d(&tc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, lvl0, Ncols, Nq, B, (uint64_t)Ncols * 4,
                               (uint64_t)Nq * Ncols * 4, STORE_COLS, 32, "C"))
            return rc;
    } else {
        tc = ta;  // unused by the kernel
    }

    // ---- 3. GEMM ----
    GemmParams p{};
    p.N = Nq;
    p.row_offset = q0;
    p.Ncols = Ncols;
    p.B = B;
    p.num_kb = Kt * pi.elem_bytes / BK_BYTES;
    p.tiles_m = ceil_div(Nq, BM);
    p.tiles_n = ceil_div(Ncols, BN);
    p.divisor = sqrt_d;
    int e = 0;
    const float mant = frexpf(sqrt_d, &e);
    p.use_div = (mant == 0.5f) ? 0 : 1;  // sqrt(D) a power of two -> exact reciprocal multiply
    p.scale = 1.0f / sqrt_d;
    p.out = lvl0;
    p.amax_bits = amax_bits;
    // instruction descriptor: D=f32, A/B format, K-major both, N>>3, M>>4
    const uint32_t fmt = tf32 ? 2u : (precision == FFCORR_PREC_FP16 ? 0u : 1u);
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const int64_t num_tiles = (int64_t)p.tiles_m * p.tiles_n * B;
    FFCORR_REQUIRE(num_tiles < (1ll << 31), FFCORR_EINVAL, "volume: too many tiles");
    const int grid = (int)(num_tiles < sm_count() ? num_tiles : sm_count());
#define FF_GEMM_F(TF, TS, DV, FU)                                                                                     \
    do {                                                                                                              \
        if (int rc = set_smem(volume_gemm_kernel<TF, TS, DV, FU>, Cfg<FU>::SMEM_TOTAL)) return rc;                    \
        volume_gemm_kernel<TF, TS, DV, FU><<<grid, Cfg<FU>::THREADS, Cfg<FU>::SMEM_TOTAL, s>>>(ta, tb, tc, tl1, p, fp, \
                                                                                             idesc);                  \
    } while (0)
#define FF_GEMM(TF, TS, DV) FF_GEMM_F(TF, TS, DV, false)
#define FF_GEMM_DV(TF, TS)            \
    do {                              \
        if (p.use_div) FF_GEMM(TF, TS, true); \
        else FF_GEMM(TF, TS, false);  \
    } while (0)
#define FF_GEMM_H(TF, DV)                                                                                             \
    do {                                                                                                              \
        if (int rc = set_smem(volume_gemm_kernel<TF, true, DV, true, true>, Cfg<true>::SMEM_TOTAL)) return rc;        \
        volume_gemm_kernel<TF, true, DV, true, true><<<grid, Cfg<true>::THREADS, Cfg<true>::SMEM_TOTAL, s>>>(ta, tb, tc, tl1, p, fp, \
                                                                                                            idesc);   \
    } while (0)
    FFCORR_REQUIRE(!out_half || fused, FFCORR_EINVAL, "volume: fp16 storage is produced by the fused build (2-4 levels) only");
    FFCORR_REQUIRE(!grouped || fused, FFCORR_EINVAL, "volume: the grouped layout is produced by the fused build (2-4 levels) only");
#define FF_GEMM_G(TF, DV)                                                                                             \
    do {                                                                                                              \
        if (int rc = set_smem(volume_gemm_kernel<TF, true, DV, true, false, true>, Cfg<true>::SMEM_TOTAL)) return rc; \
        volume_gemm_kernel<TF, true, DV, true, false, true><<<grid, Cfg<true>::THREADS, Cfg<true>::SMEM_TOTAL, s>>>(ta, tb, tc, tl1, p, fp, \
                                                                                                                   idesc);   \
    } while (0)
    if (fused && grouped) {
        if (tf32) { if (p.use_div) FF_GEMM_G(true, true); else FF_GEMM_G(true, false); }
        else      { if (p.use_div) FF_GEMM_G(false, true); else FF_GEMM_G(false, false); }
    } else if (fused && out_half) {
        if (tf32) { if (p.use_div) FF_GEMM_H(true, true); else FF_GEMM_H(true, false); }
        else      { if (p.use_div) FF_GEMM_H(false, true); else FF_GEMM_H(false, false); }
    } else if (fused) {
        if (tf32) { if (p.use_div) FF_GEMM_F(true, true, true, true); else FF_GEMM_F(true, true, false, true); }
        else      { if (p.use_div) FF_GEMM_F(false, true, true, true); else FF_GEMM_F(false, true, false, true); }
    }
    else if (tf32 && tma_store) FF_GEMM_DV(true, true);
    else if (tf32) FF_GEMM_DV(true, false);
    else if (tma_store) FF_GEMM_DV(false, true);
    else FF_GEMM_DV(false, false);
#undef FF_GEMM_DV
#undef FF_GEMM
#undef FF_GEMM_F
#undef FF_GEMM_H
#undef FF_GEMM_G
    return check_launch("volume_gemm_kernel");
}

namespace ffcorr {
namespace {
int launch_gemm_nt_tf32(const float* A, const float* Bm, float* Ct, int M, int Nn, int K, int batches, float divisor,
                        cudaStream_t s) {
    CUtensorMap ta, tb;
    const uint64_t row_bytes = (uint64_t)K * 4;
    if (int rc = encode_3d(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(A), K, M, batches, row_bytes,
                           row_bytes * M, BK_BYTES / 4, BM, "bwd A")) return rc;
    if (int rc = encode_3d(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(Bm), K, Nn, batches, row_bytes,
                           row_bytes * Nn, BK_BYTES / 4, BN, "bwd B")) return rc;
    GemmParams p{};
    p.N = M;
    p.Ncols = Nn;
    p.B = batches;
    p.num_kb = (int)((row_bytes + BK_BYTES - 1) / BK_BYTES);   // the K tail is zero-filled by TMA
    p.tiles_m = ceil_div(M, BM);
    p.tiles_n = ceil_div(Nn, BN);
    p.divisor = divisor;
    int e = 0;
    const float mant = frexpf(divisor, &e);
    p.use_div = (mant == 0.5f) ? 0 : 1;
    p.scale = 1.0f / divisor;
    p.out = Ct;
    p.out_transposed = 1;
    p.ldc = M;
    FusedParams fp{};
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const int64_t num_tiles = (int64_t)p.tiles_m * p.tiles_n * batches;
    FFCORR_REQUIRE(num_tiles < (1ll << 31), FFCORR_EINVAL, "volume_bwd: too many tiles");
    const int grid = (int)(num_tiles < sm_count() ? num_tiles : sm_count());
    if (p.use_div) {
        if (int rc = set_smem(volume_gemm_kernel<true, false, true, false>, Cfg<false>::SMEM_TOTAL)) return rc;
        volume_gemm_kernel<true, false, true, false><<<grid, Cfg<false>::THREADS, Cfg<false>::SMEM_TOTAL, s>>>(ta, tb, ta, ta, p, fp, idesc);
    } else {
        if (int rc = set_smem(volume_gemm_kernel<true, false, false, false>, Cfg<false>::SMEM_TOTAL)) return rc;
        volume_gemm_kernel<true, false, false, false><<<grid, Cfg<false>::THREADS, Cfg<false>::SMEM_TOTAL, s>>>(ta, tb, ta, ta, p, fp, idesc);
    }
    return check_launch("volume_gemm_kernel (backward)");
}
}  // namespace
}  // namespace ffcorr

extern "C" int ffcorr_volume_f32(const float* fmap1, const float* fmap2, float* lvl0, int B, int D, int h, int w,
                                 int precision, void* workspace, size_t workspace_bytes, void* stream) {
    return volume_impl(fmap1, fmap2, lvl0, B, D, h, w, precision, workspace, workspace_bytes, stream, OUT_ROWMAJOR);
}

extern "C" int ffcorr_volume_scaled_f32(const float* fmap1, const float* fmap2, float* lvl0, int B, int D, int h, int w,
                                        int precision, float divisor, void* workspace, size_t workspace_bytes, void* stream) {
    FFCORR_REQUIRE(divisor > 0.0f && divisor < 3.0e38f, FFCORR_EINVAL, "volume_scaled: divisor=%g must be positive and finite", (double)divisor);
    return volume_impl(fmap1, fmap2, lvl0, B, D, h, w, precision, workspace, workspace_bytes, stream, OUT_ROWMAJOR, nullptr, 1,
                       0, -1, PHASE_ALL, divisor);
}

extern "C" int ffcorr_volume_tiled_f32(const float* fmap1, const float* fmap2, float* lvl0_tiled, int B, int D, int h, int w,
                                       int precision, void* workspace, size_t workspace_bytes, void* stream) {
    return volume_impl(fmap1, fmap2, lvl0_tiled, B, D, h, w, precision, workspace, workspace_bytes, stream, OUT_TILED);
}

extern "C" int ffcorr_build_tiled_f32(const float* fmap1, const float* fmap2, float* const* lvl, int num_levels, int B, int D,
                                      int h, int w, int precision, void* workspace, size_t workspace_bytes, void* stream) {
    FFCORR_REQUIRE(lvl != nullptr, FFCORR_EINVAL, "build_tiled: null level table");
    if (int rc = check_levels(num_levels, h, w, "build_tiled")) return rc;
    FFCORR_REQUIRE(ffcorr_tiled_supported(num_levels, h, w), FFCORR_EINVAL,
                   "build_tiled: %dx%d with %d levels is outside the tiled path (use the row-major entry points)", h, w, num_levels);
    FFCORR_REQUIRE(B >= 0, FFCORR_EINVAL, "build_tiled: B=%d", B);
    if (B == 0) return FFCORR_OK;
    for (int i = 0; i < num_levels; ++i) {
        FFCORR_REQUIRE(lvl[i] != nullptr, FFCORR_EINVAL, "build_tiled: lvl[%d] is null", i);
        FFCORR_REQUIRE((uintptr_t)lvl[i] % 16 == 0, FFCORR_EALIGN, "build_tiled: lvl[%d] must be 16-byte aligned", i);
    }
    if (num_levels == 1)
        return volume_impl(fmap1, fmap2, lvl[0], B, D, h, w, precision, workspace, workspace_bytes, stream, OUT_TILED);
    return volume_impl(fmap1, fmap2, lvl[0], B, D, h, w, precision, workspace, workspace_bytes, stream, OUT_FUSED_PYRAMID,
                       lvl, num_levels);
}

extern "C" int ffcorr_build_tiled_f16(const float* fmap1, const float* fmap2, void* const* lvl, int num_levels, int B, int D,
                                      int h, int w, int precision, void* workspace, size_t workspace_bytes, void* stream) {
    FFCORR_REQUIRE(lvl != nullptr, FFCORR_EINVAL, "build_tiled_f16: null level table");
    if (int rc = check_levels(num_levels, h, w, "build_tiled_f16")) return rc;
    FFCORR_REQUIRE(num_levels >= 2 && ffcorr_tiled_supported(num_levels, h, w), FFCORR_EINVAL,
                   "build_tiled_f16: %dx%d with %d levels is outside the fused build (2-4 levels)", h, w, num_levels);
    FFCORR_REQUIRE(precision != FFCORR_PREC_FP32, FFCORR_EINVAL, "build_tiled_f16: needs a tensor-core operand precision");
    FFCORR_REQUIRE(B >= 0, FFCORR_EINVAL, "build_tiled_f16: B=%d", B);
    if (B == 0) return FFCORR_OK;
    for (int i = 0; i < num_levels; ++i) {
        FFCORR_REQUIRE(lvl[i] != nullptr, FFCORR_EINVAL, "build_tiled_f16: lvl[%d] is null", i);
        FFCORR_REQUIRE((uintptr_t)lvl[i] % 16 == 0, FFCORR_EALIGN, "build_tiled_f16: lvl[%d] must be 16-byte aligned", i);
    }
    // the level pointers travel as float* through the shared implementation; the HALF kernel reinterprets them
    return volume_impl(fmap1, fmap2, reinterpret_cast<float*>(lvl[0]), B, D, h, w, precision, workspace, workspace_bytes, stream,
                       OUT_FUSED_PYRAMID, reinterpret_cast<float* const*>(lvl), num_levels, 0, -1, PHASE_ALL, 0.0f, true);
}

extern "C" int ffcorr_build_grouped_f32(const float* fmap1, const float* fmap2, float* const* lvl, int num_levels, int B, int D,
                                        int h, int w, int precision, void* workspace, size_t workspace_bytes, void* stream) {
    FFCORR_REQUIRE(lvl != nullptr, FFCORR_EINVAL, "build_grouped: null level table");
    if (int rc = check_levels(num_levels, h, w, "build_grouped")) return rc;
    FFCORR_REQUIRE(num_levels >= 2 && ffcorr_tiled_supported(num_levels, h, w), FFCORR_EINVAL,
                   "build_grouped: %dx%d with %d levels is outside the fused build (2-4 levels)", h, w, num_levels);
    FFCORR_REQUIRE(precision != FFCORR_PREC_FP32, FFCORR_EINVAL, "build_grouped: needs a tensor-core operand precision");
    FFCORR_REQUIRE(B >= 0, FFCORR_EINVAL, "build_grouped: B=%d", B);
    if (B == 0) return FFCORR_OK;
    for (int i = 0; i < num_levels; ++i) {
        FFCORR_REQUIRE(lvl[i] != nullptr, FFCORR_EINVAL, "build_grouped: lvl[%d] is null", i);
        FFCORR_REQUIRE((uintptr_t)lvl[i] % 16 == 0, FFCORR_EALIGN, "build_grouped: lvl[%d] must be 16-byte aligned", i);
    }
    return volume_impl(fmap1, fmap2, lvl[0], B, D, h, w, precision, workspace, workspace_bytes, stream, OUT_FUSED_PYRAMID, lvl,
                       num_levels, 0, -1, PHASE_ALL, 0.0f, false, true);
}

extern "C" int ffcorr_stage_operands_f32(const float* fmap1, const float* fmap2, int num_levels, int B, int D, int h, int w,
                                         int precision, void* workspace, size_t workspace_bytes, void* stream) {
    if (int rc = check_levels(num_levels, h, w, "stage_operands")) return rc;
    FFCORR_REQUIRE(ffcorr_tiled_supported(num_levels, h, w), FFCORR_EINVAL, "stage_operands: shape outside the tiled path");
    float dummy = 0.f;   // PHASE_STAGE never touches the output
    return volume_impl(fmap1, fmap2, &dummy, B, D, h, w, precision, workspace, workspace_bytes, stream,
                       num_levels == 1 ? OUT_TILED : OUT_FUSED_PYRAMID, nullptr, num_levels, 0, -1, PHASE_STAGE);
}

extern "C" int ffcorr_build_tiled_chunk_f32(float* const* lvl, int num_levels, int B, int D, int h, int w, int q0, int nq,
                                            int precision, void* workspace, size_t workspace_bytes, void* stream) {
    FFCORR_REQUIRE(lvl != nullptr, FFCORR_EINVAL, "build_tiled_chunk: null level table");
    if (int rc = check_levels(num_levels, h, w, "build_tiled_chunk")) return rc;
    FFCORR_REQUIRE(ffcorr_tiled_supported(num_levels, h, w), FFCORR_EINVAL, "build_tiled_chunk: shape outside the tiled path");
    FFCORR_REQUIRE(B >= 0, FFCORR_EINVAL, "build_tiled_chunk: B=%d", B);
    if (B == 0) return FFCORR_OK;
    for (int i = 0; i < num_levels; ++i) {
        FFCORR_REQUIRE(lvl[i] != nullptr, FFCORR_EINVAL, "build_tiled_chunk: lvl[%d] is null", i);
        FFCORR_REQUIRE((uintptr_t)lvl[i] % 16 == 0, FFCORR_EALIGN, "build_tiled_chunk: lvl[%d] must be 16-byte aligned", i);
    }
    return volume_impl(nullptr, nullptr, lvl[0], B, D, h, w, precision, workspace, workspace_bytes, stream,
                       num_levels == 1 ? OUT_TILED : OUT_FUSED_PYRAMID, lvl, num_levels, q0, nq, PHASE_GEMM);
}

extern "C" int ffcorr_volume_bwd_f32(float* grad_lvl0, const float* fmap1, const float* fmap2, float* gfmap1,
                                     float* gfmap2, int B, int D, int h, int w, int precision, void* stream) {
    FFCORR_REQUIRE(B >= 0 && D >= 1 && h >= 1 && w >= 1, FFCORR_EINVAL, "volume_bwd: bad shape");
    if (B == 0) return FFCORR_OK;
    FFCORR_REQUIRE(grad_lvl0 && fmap1 && fmap2, FFCORR_EINVAL, "volume_bwd: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const int N = h * w;
    const float sqrt_d = sqrtf((float)D);
    const int64_t fb = (int64_t)D * N, gb = (int64_t)N * N;
    const bool aligned = ((uintptr_t)grad_lvl0 % 16 == 0) && ((uintptr_t)fmap1 % 16 == 0) && ((uintptr_t)fmap2 % 16 == 0);
    if (precision != FFCORR_PREC_FP32 && N % 4 == 0 && aligned && N >= 32) {
        // tensor-core path (tf32 operands, fp32 accumulate -- what the reference's cuBLAS backward does under
        // ALLOW_TF32).  gf1^T[i,d] = sum_j g[i,j] f2[d,j]: both operands K-major in place.  gf2^T[j,d] = sum_i
        // g[i,j] f1[d,i] needs g^T: the gradient is dead after this call, so it is transposed in place.
        if (gfmap1)
            if (int rc = launch_gemm_nt_tf32(grad_lvl0, fmap2, gfmap1, N, D, N, B, sqrt_d, s)) return rc;
        if (gfmap2) {
            const int tiles = ceil_div(N, 32);
            FFCORR_REQUIRE(B < 65536, FFCORR_EINVAL, "volume_bwd: B=%d too large", B);
            transpose_inplace_kernel<<<dim3((unsigned)((int64_t)tiles * (tiles + 1) / 2), B), 256, 0, s>>>(grad_lvl0, N, tiles);
            if (int rc = check_launch("transpose_inplace_kernel")) return rc;
            if (int rc = launch_gemm_nt_tf32(grad_lvl0, fmap1, gfmap2, N, D, N, B, sqrt_d, s)) return rc;
        }
        return FFCORR_OK;
    }
    if (gfmap1) {
        // gf1[d,i] = sum_j f2[d,j] g[i,j]: A[m=d,k=j] = f2 (k contiguous), B[k=j,n=i] = g[i*N+j] (k contiguous)
        if (int rc = launch_sgemm(false, false, fmap2, grad_lvl0, gfmap1, D, N, N, N, 1, 1, N, N, fb, gb, fb, B, sqrt_d, s))
            return rc;
    }
    if (gfmap2) {
        // gf2[d,j] = sum_i f1[d,i] g[i,j]: A[m=d,k=i] = f1 (k contiguous), B[k=i,n=j] = g (n contiguous)
        if (int rc = launch_sgemm(false, true, fmap1, grad_lvl0, gfmap2, D, N, N, N, 1, N, 1, N, fb, gb, fb, B, sqrt_d, s))
            return rc;
    }
    return FFCORR_OK;
}
