/*
 * ffcorr.h -- C ABI of libffcorr.so: the B200 (sm_100a) correlation hot path of
 * FocusFlow (FocusRAFT CorrBlock + FF-PWC 9x9 local cost volume).
 *
 * Drop-in boundary.  The reference has no FFI layer: its hot path is two Python
 * call surfaces over torch / CuPy.  Each entry point below replaces the device
 * work behind one of them (paths relative to the reference root):
 *
 *   ffcorr_volume_f32      core/models/ff-raft/FF_RAFT_Core/corr.py:52-60   CorrBlock.corr
 *                          (torch.matmul + separate "/ sqrt(D)" kernel)
 *   ffcorr_pyramid_f32     corr.py:24-27   3x F.avg_pool2d(corr, 2, stride=2)
 *   ffcorr_build_tiled_f32 corr.py:13-27   CorrBlock.__init__ = volume + pyramid, fused (tiled layout)
 *   ffcorr_*_chunk_f32     corr.py:63-91   AlternateCorrBlock: memory-bounded, recomputed per lookup
 *   ffcorr_lookup_f32      corr.py:29-50   CorrBlock.__call__  +  utils/utils.py:57-71
 *                          bilinear_sampler (4x F.grid_sample + ~60 glue kernels)
 *   ffcorr_lookup_bwd_f32  autograd of the above w.r.t. the pyramid (coords are detached
 *                          by the caller, raft.py:216-217)
 *   ffcorr_pyramid_bwd_f32 autograd of corr.py:24-27
 *   ffcorr_volume_bwd_f32  autograd of corr.py:52-60
 *   ffcorr_pwc81_f32       core/models/ff-pwcnet/PWCNet_Core/correlation.py:278-328
 *                          _FunctionCorrelation.forward (rearrange x2 + updateOutput)
 *   ffcorr_pwc81_bwd_f32   correlation.py:331-380 _FunctionCorrelation.backward
 *   ffcorr_backwarp_f32    core/models/ff-pwcnet/PWCNet_Core/ff_pwcnet.py:27-46 backwarp (caller side of the PWC op)
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch's caching
 *     allocator in the shipped host code); the library allocates no device memory;
 *   - all tensors are dense, contiguous, fp32, in the reference's layouts;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *     synchronises the host;
 *   - return value: 0 on success, a negative FFCORR_E* code for argument errors,
 *     a positive cudaError_t when the CUDA runtime reports one.  Nothing throws.
 *     ffcorr_last_error() returns a thread-local message for the last failure.
 *   - re-entrant: no global mutable state besides cached function attributes (every behavioural switch --
 *     operand precision, sampler semantics, output layout -- is an argument of the call it affects).
 */
#ifndef FFCORR_H_
#define FFCORR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFCORR_VERSION 201

#define FFCORR_OK            0
#define FFCORR_EINVAL      (-1)   /* bad shape / null pointer / unsupported argument   */
#define FFCORR_EWORKSPACE  (-2)   /* workspace missing or too small                     */
#define FFCORR_EDEVICE     (-3)   /* not an sm_100 device / driver entry point missing  */
#define FFCORR_EALIGN      (-4)   /* pointer alignment the kernel needs is not met      */

#define FFCORR_MAX_LEVELS    8

/* operand precision of the all-pairs contraction (accumulation is always fp32) */
#define FFCORR_PREC_FP16     0    /* tcgen05 kind::f16, operands rounded to fp16 (RN): default          */
#define FFCORR_PREC_FP32     1    /* CUDA-core fp32 FMA: exact path for training / debugging            */
#define FFCORR_PREC_BF16X3   2    /* tcgen05, bf16 hi/lo split, 3 MMAs per k-step: ~fp32 accuracy       */
#define FFCORR_PREC_TF32     3    /* tcgen05 kind::tf32 on fp32 operands                                */

int         ffcorr_version(void);
const char* ffcorr_last_error(void);

/* number of SMs / compute capability of the current device (for tests and the bench) */
int ffcorr_device_info(int* sm_count, int* cc_major, int* cc_minor);

/*
 * Device-wide hint for the DRAM->L2 fetch size (cudaLimitMaxL2FetchGranularity: 32, 64 or 128).
 * Diagnostic only: on B200 the measured DRAM traffic of the lookup did not change with it
 * (the fetch unit stays 64 bytes), so the host code leaves the driver default alone.
 */
int ffcorr_set_l2_fetch_granularity(int bytes);
int ffcorr_get_l2_fetch_granularity(int* bytes);

/*
 * All-pairs correlation volume, level 0 of the pyramid.
 *   fmap1, fmap2 : [B, D, h, w]          (raft.py:191-193 hands them over as fp32 NCHW)
 *   lvl0         : [B*h*w, h, w]         == reference [B*N, 1, h, w], N = h*w
 *   lvl0[b*N + i, j] = sum_d fmap1[b,d,i] * fmap2[b,d,j] / sqrt(D)
 * workspace: ffcorr_volume_workspace_bytes() bytes, 256-byte aligned (operand staging
 * for the tensor-core paths; may be NULL/0 for FFCORR_PREC_FP32).
 */
size_t ffcorr_volume_workspace_bytes(int B, int D, int h, int w, int precision);
int ffcorr_volume_f32(const float* fmap1, const float* fmap2, float* lvl0,
                      int B, int D, int h, int w, int precision,
                      void* workspace, size_t workspace_bytes, void* stream);

/*
 * Same contraction with an explicit divisor instead of sqrt(D): FlowFormer's cost volume
 * (FF_FlowFormer_Core/FlowFormer/LatentCostFormer/encoder.py:335-347, one head) is the unscaled one, divisor = 1.
 */
int ffcorr_volume_scaled_f32(const float* fmap1, const float* fmap2, float* lvl0,
                             int B, int D, int h, int w, int precision, float divisor,
                             void* workspace, size_t workspace_bytes, void* stream);

/*
 * Average-pool pyramid.  lvl[0] is the input [Q, h, w]; lvl[i] (i >= 1) is written,
 * [Q, h>>i, w>>i] with floor semantics and the ((a+b)+c+d)/4 summation order of
 * ATen avg_pool2d, so the result is bit-identical to the reference given the same lvl[0].
 * `lvl` is a HOST array of num_levels device pointers.
 */
int ffcorr_pyramid_f32(float* const* lvl, int num_levels, int64_t Q, int h, int w, void* stream);

/*
 * `sampler`: which of the reference's two runs the lookups reproduce bit-closely.  bilinear_sampler (utils/utils.py:61-62)
 * normalises with x / (W-1): ATen's CPU kernel divides, its CUDA kernel multiplies by the fp32 reciprocal of the
 * scalar -- a <= 1-ulp difference of the normalised coordinate (~1e-5 px at w = 156, ~3e-5 of max|value| in the
 * output).  FFCORR_SAMPLER_ATEN_CUDA is what a GPU user of the reference gets (and the Python default);
 * FFCORR_SAMPLER_ATEN_CPU is the form the golden vectors (reference run on CPU) were made with.  Per call.
 */
#define FFCORR_SAMPLER_ATEN_CPU   0
#define FFCORR_SAMPLER_ATEN_CUDA  1

/*
 * Fused multi-level bilinear window lookup (one launch per refinement iteration).
 *   lvl     : HOST array of num_levels device pointers, lvl[i] = [B*h*w, h>>i, w>>i]
 *   coords  : [B, 2, h, w]  channel 0 = x, channel 1 = y   (utils.py:74-77)
 *   out     : [B, num_levels*(2r+1)^2, h, w]; channel = lvl*(2r+1)^2 + a*(2r+1) + b samples
 *             (x/2^lvl + a - r, y/2^lvl + b - r)            (corr.py:37-43)
 *             out_channels_last != 0: the same values stored [B, h, w, num_levels*(2r+1)^2] (NHWC) -- the layout the
 *             consumer, the 1x1 convolution `convc1` (update.py:82-83,90), wants on tensor cores; every query's
 *             channels are then one contiguous run.
 * Zero padding, align_corners=True and the normalise/un-normalise fp32 round trip of
 * bilinear_sampler + grid_sample are reproduced tap by tap.  radius in [1, 4]; h*w < 2^24.
 */
int ffcorr_lookup_f32(const float* const* lvl, int num_levels, const float* coords, float* out,
                      int B, int h, int w, int radius, int sampler, int out_channels_last, void* stream);

/*
 * Tiled ("T4") pyramid layout -- the inference fast path.  Every level is stored per query map as
 * [ceil(h_i/4)][tw_i][4][4] floats with tw_i = ceil(w_i/4) rounded up to an even number, so that every tile
 * row starts on a 128-byte line (ffcorr_tiled_map_elems() per map).  Pixels beyond the level size inside
 * tiles that hold pixels are exact zeros; the extra tile column that evens the pitch is never read.
 * A 4x4 tile is 64 contiguous bytes, the granularity at which the memory system serves the lookup's
 * gathers.  The three entry points below are drop-ins for ffcorr_volume_f32 / ffcorr_pyramid_f32 /
 * ffcorr_lookup_f32 on that layout (same arguments, same results); ffcorr_untile_f32 / ffcorr_tile_f32
 * convert a level to / from the reference's row-major [Q, h_i, w_i] (CorrBlock.corr_pyramid, tests).
 * ffcorr_tiled_supported() tells whether a shape fits ffcorr_build_tiled_f32 + ffcorr_lookup_tiled_f32 (<= 4 levels);
 * the standalone ffcorr_pyramid_tiled_f32 additionally needs the level-0 map to fit its shared-memory staging (~48 KB).
 */
int64_t ffcorr_tiled_map_elems(int h, int w, int level);
int ffcorr_tiled_supported(int num_levels, int h, int w);
int ffcorr_volume_tiled_f32(const float* fmap1, const float* fmap2, float* lvl0_tiled,
                            int B, int D, int h, int w, int precision,
                            void* workspace, size_t workspace_bytes, void* stream);
int ffcorr_pyramid_tiled_f32(float* const* lvl, int num_levels, int64_t Q, int h, int w, void* stream);
/*
 * CorrBlock.__init__ (corr.py:13-27) in ONE GEMM launch: level 0 and the pooled levels 1..num_levels-1 are
 * all written by the GEMM epilogue, so level 0 is never read back (the standalone pyramid re-reads it).
 * `lvl` is a HOST array of num_levels device pointers to tiled levels; results are bit-identical to
 * ffcorr_volume_tiled_f32 followed by ffcorr_pyramid_tiled_f32.  Same workspace as ffcorr_volume_f32.
 */
int ffcorr_build_tiled_f32(const float* fmap1, const float* fmap2, float* const* lvl, int num_levels,
                           int B, int D, int h, int w, int precision,
                           void* workspace, size_t workspace_bytes, void* stream);
int ffcorr_lookup_tiled_f32(const float* const* lvl, int num_levels, const float* coords, float* out,
                            int B, int h, int w, int radius, int sampler, int out_channels_last, void* stream);

/*
 * The lookup fused with its consumer: out = relu(convc1(lookup(coords))), the first layer of
 * BasicMotionEncoder (update.py:82-83,90: Conv2d(324, 256, 1) + ReLU on the lookup result).  The 324 samples of a
 * query are written as fp16 into a tensor-core operand tile in shared memory and never reach global memory; the
 * weights stay in tensor memory for the whole launch.  out is [B, h, w, 256] fp32 (NHWC: what the next convolution
 * of the update block reads under channels_last).  Same sample values as ffcorr_lookup_tiled_f32 before the
 * rounding to fp16 (which has the 11 significant bits of the TF32 convolution the reference runs, common.py:25-27);
 * fp32 accumulation; samples saturate at +-65504.  Built for the reference's 4 levels x radius 4 only.
 *   ffcorr_pack_convc1_weight: convc1.weight [256, 324] fp32 (device) -> the kernel's operand order, fp16,
 *   ffcorr_convc1_packed_bytes() bytes; do it once per set of weights.
 *   tile_counter: one device int32 that is ZERO when the launch starts (stream-ordered) and is not shared with another
 *   launch in flight; the persistent kernel hands its tiles of 64 queries out through it and leaves it non-zero.
 */
size_t ffcorr_convc1_packed_bytes(void);
int ffcorr_pack_convc1_weight(const float* weight, int cout, int cin, void* packed, void* stream);
int ffcorr_lookup_convc1_tiled_f32(const float* const* lvl, int num_levels, const float* coords, const void* packed_weight,
                                   const float* bias, float* out, int* tile_counter, int B, int h, int w, int radius,
                                   int sampler, void* stream);

/*
 * Opt-in HALF-PRECISION STORAGE of the pyramid (the Python side: CorrBlock(..., storage="fp16")).  Same geometry as
 * the tiled fp32 layout -- level i per query map = [ceil(h_i/4)][tw_i][4][4] elements, ffcorr_tiled_map_elems() of
 * them -- but the elements are IEEE fp16 (a 4x4 tile is one 32-byte DRAM sector).  Accumulation (fp32, TMEM), the
 * 1/sqrt(D) scale and the 2x2 poolings are unchanged; every level is rounded to fp16 (RN) once, when it is written,
 * and the lookup interpolates in fp32.  Halves the bytes of both HBM-bound kernels: the build's 2.3 GB write and the
 * lookup's window gathers.  Values: volume within 1e-3 of the reference (2e-4 storage rounding on top of the operand
 * rounding), lookups exact to 1e-5 w.r.t. the stored pyramid; |corr| must stay below 65504.
 * ffcorr_build_tiled_f16 needs 2 <= num_levels <= 4 and a tensor-core operand precision; ffcorr_lookup_tiled_f16 writes
 * channels-last output only (out_channels_last must be non-zero); ffcorr_untile_f16 / ffcorr_tile_f16 convert a level
 * to / from the reference's row-major fp32 [Q, h_i, w_i].
 */
int ffcorr_build_tiled_f16(const float* fmap1, const float* fmap2, void* const* lvl, int num_levels,
                           int B, int D, int h, int w, int precision,
                           void* workspace, size_t workspace_bytes, void* stream);
int ffcorr_lookup_tiled_f16(const void* const* lvl, int num_levels, const float* coords, float* out,
                            int B, int h, int w, int radius, int sampler, int out_channels_last, void* stream);
int ffcorr_untile_f16(const void* tiled, float* dst, int64_t Q, int h_level, int w_level, void* stream);
int ffcorr_tile_f16(const float* src, void* tiled, int64_t Q, int h_level, int w_level, void* stream);

/*
 * Grouped ("G32") tile order (the Python side: CorrBlock(..., layout="grouped")).  Same 4x4-pixel tiles, same values,
 * but 32 consecutive queries share one tile grid: level i = [B][NG = ceil(N/32)][th_i * tw_i][32 queries][4][4] floats,
 * ffcorr_grouped_level_elems() of them.  Neighbouring queries' windows overlap under any smooth flow, so the tiles a
 * lookup warp gathers -- and the pieces an epilogue warp of the build writes (8 KB of level 0, 4 KB of level 1 per chunk)
 * -- are contiguous in memory instead of one map (~30 KB) apart.  ffcorr_build_grouped_f32 / ffcorr_lookup_grouped_f32
 * take the arguments of ffcorr_build_tiled_f32 / ffcorr_lookup_tiled_f32 and return the same values (2 <= num_levels
 * <= 4, tensor-core precisions); ffcorr_ungroup_f32 / ffcorr_group_f32 convert a level to / from the reference's
 * row-major [B*N, h_i, w_i].
 */
int64_t ffcorr_grouped_level_elems(int h, int w, int level, int B, int num_queries);
int ffcorr_build_grouped_f32(const float* fmap1, const float* fmap2, float* const* lvl, int num_levels,
                             int B, int D, int h, int w, int precision,
                             void* workspace, size_t workspace_bytes, void* stream);
int ffcorr_lookup_grouped_f32(const float* const* lvl, int num_levels, const float* coords, float* out,
                              int B, int h, int w, int radius, int sampler, int out_channels_last, void* stream);
int ffcorr_ungroup_f32(const float* grouped, float* dst, int B, int N, int h_level, int w_level, void* stream);
int ffcorr_group_f32(const float* src, float* grouped, int B, int N, int h_level, int w_level, void* stream);

/*
 * Query-chunked build + lookup: AlternateCorrBlock semantics (corr.py:63-91) -- the same lookup values with
 * O(nq * h*w) instead of O((h*w)^2) pyramid memory, recomputed per lookup.  Stage the GEMM operands once per
 * image pair (ffcorr_stage_operands_f32, same workspace as ffcorr_volume_f32), then per lookup and per chunk of
 * queries [q0, q0+nq): ffcorr_build_tiled_chunk_f32 writes the tiled pyramid of those queries only
 * (lvl[i] = [B*nq, ffcorr_tiled_map_elems(h, w, i)]), ffcorr_lookup_tiled_chunk_f32 reads the chunk's slice of the
 * FULL coords [B, 2, h, w] and writes its slice of the FULL out [B, L*(2r+1)^2, h, w] in place.  Values are
 * bit-identical to ffcorr_build_tiled_f32 + ffcorr_lookup_tiled_f32.  num_levels must agree between the calls.
 */
int ffcorr_stage_operands_f32(const float* fmap1, const float* fmap2, int num_levels, int B, int D, int h, int w,
                              int precision, void* workspace, size_t workspace_bytes, void* stream);
int ffcorr_build_tiled_chunk_f32(float* const* lvl, int num_levels, int B, int D, int h, int w, int q0, int nq,
                                 int precision, void* workspace, size_t workspace_bytes, void* stream);
int ffcorr_lookup_tiled_chunk_f32(const float* const* lvl, int num_levels, const float* coords, float* out,
                                  int B, int h, int w, int q0, int nq, int radius, int sampler, int out_channels_last,
                                  void* stream);
int ffcorr_untile_f32(const float* tiled, float* dst, int64_t Q, int h_level, int w_level, void* stream);
int ffcorr_tile_f32(const float* src, float* tiled, int64_t Q, int h_level, int w_level, void* stream);

/*
 * Adjoint of the lookup w.r.t. the pyramid.  grad_lvl[i] must be ZEROED by the caller;
 * contributions are accumulated with atomics (red.global.add.f32).
 */
int ffcorr_lookup_bwd_f32(float* const* grad_lvl, int num_levels, const float* coords,
                          const float* grad_out, int B, int h, int w, int radius, int sampler, void* stream);

/*
 * Adjoint of the pyramid: grad_lvl[i-1] += upsample(grad_lvl[i]) / 4 for i = L-1 .. 1
 * (in place, so after the call grad_lvl[0] holds d loss / d lvl0).
 */
int ffcorr_pyramid_bwd_f32(float* const* grad_lvl, int num_levels, int64_t Q, int h, int w, void* stream);

/*
 * Adjoint of the volume: gfmap1[b,d,i] = sum_j g[b,i,j] fmap2[b,d,j] / sqrt(D),
 *                         gfmap2[b,d,j] = sum_i g[b,i,j] fmap1[b,d,i] / sqrt(D).
 * precision FFCORR_PREC_FP32: exact CUDA-core path.  Any other precision: tcgen05 kind::tf32 straight from the
 * fp32 arrays (what the reference's cuBLAS backward does under ALLOW_TF32; needs h*w % 4 == 0, else the
 * CUDA-core path runs).  grad_lvl0 is SCRATCH: the tensor-core path transposes it in place for the second
 * product, so its content is unspecified afterwards.  Either gradient pointer may be NULL.
 */
int ffcorr_volume_bwd_f32(float* grad_lvl0, const float* fmap1, const float* fmap2,
                          float* gfmap1, float* gfmap2, int B, int D, int h, int w, int precision, void* stream);

/*
 * PWC local cost volume, max displacement 4 (81 channels).
 *   one, two : [B, C, H, W];  out : [B, 81, H, W]
 *   out[b, (dy+4)*9 + (dx+4), y, x] = (1/C) sum_c one[b,c,y,x] * two[b,c,y+dy,x+dx]   (0 outside)
 * leaky_slope < 0 writes the raw volume (reference semantics); 0 <= leaky_slope < 1 fuses
 * the leaky_relu the reference applies right after the call (ff_pwcnet.py:317,325).
 */
int ffcorr_pwc81_f32(const float* one, const float* two, float* out,
                     int B, int C, int H, int W, float leaky_slope, void* stream);

/* gradients of the raw cost volume; either output may be NULL. */
int ffcorr_pwc81_bwd_f32(const float* one, const float* two, const float* grad_out,
                         float* grad_one, float* grad_two, int B, int C, int H, int W, void* stream);

/*
 * backwarp of FF-PWC (core/models/ff-pwcnet/PWCNet_Core/ff_pwcnet.py:27-46): the second feature map warped towards
 * the first by the up-sampled flow, before the correlation of decoder levels 5..2 (ff_pwcnet.py:322-325).
 *   input : [B, C, H, W];  flow : [B, 2, H, W] (channel 0 = x);  out : [B, C, H, W]
 *   grid_x[W], grid_y[H] : the pixel-centre grids linspace(-1 + 1/size, 1 - 1/size, size) the reference caches
 *                          (ff_pwcnet.py:29-31), passed in so that they are bit-identical to torch.linspace
 *   out = grid_sample(input, grid + flow * flow_scale / ((size-1)/2), bilinear, zeros, align_corners=False)
 *         * (sum of in-bounds tap weights > 0.999)
 * One kernel instead of the reference's ~9 (mul, div x2, cat x2, add, permute, grid_sample over C+1 planes, mask, mul).
 */
int ffcorr_backwarp_f32(const float* input, const float* flow, const float* grid_x, const float* grid_y, float* out,
                        int B, int C, int H, int W, float flow_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FFCORR_H_ */
