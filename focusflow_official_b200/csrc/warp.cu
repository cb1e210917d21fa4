// backwarp of FF-PWC (PWCNet_Core/ff_pwcnet.py:27-46) as ONE kernel (sm_100a).
//
// The reference warps the second feature map towards the first with the up-sampled flow before every
// correlation of decoder levels 5..2 (ff_pwcnet.py:322-325):
//     grid  = linspace(-1 + 1/W, 1 - 1/W, W)  (x)  /  linspace(-1 + 1/H, 1 - 1/H, H)  (y)      pixel centres
//     g     = grid + flow * fltBackwarp / ((size - 1) / 2)                                     (sic: size - 1)
//     out   = grid_sample(cat(input, ones), g, bilinear, zeros, align_corners=False)
//     mask  = out[:, -1] > 0.999 ? 1 : 0 ;   result = out[:, :-1] * mask
// i.e. a chain of ~9 PyTorch kernels (mul, 2x div, 2x cat, add, permute, grid_sample over C+1 planes, 2 masked
// assignments, mul) that materialises [B, C+1, H, W] twice.  Here: a thread owns one pixel, derives the four taps,
// their weights and the validity mask once, and streams the C channels (coalesced along x).
// Arithmetic follows ATen's CUDA kernels operation by operation (GridSampler.cuh: grid_sampler_unnormalize,
// grid_sampler_2d_kernel; scalar division = multiplication by the fp32 reciprocal), so the result agrees with the
// reference formula run through torch on the same GPU to rounding.
// Algorithmic bytes: 4 * B*H*W * (2C + 2): read `input` once (taps overlap between neighbours), write the result.
#include <algorithm>

#include "common.cuh"

namespace ffcorr {
namespace {

// Block = TX x (256 / TX) pixels (TX = 64, 32 or 16: the small decoder levels are 32 and 16 pixels wide) x one CHUNK of
// channels: blockIdx.z = b * chunks + chunk.  Splitting the channels over the grid is what keeps the small levels from
// being latency-bound (level 5 is 14 x 32 pixels x 128 channels: one thread per pixel looping over all channels was 64
// blocks of 128 dependent iterations).
template <int TX>
__global__ void __launch_bounds__(256) backwarp_kernel(const float* __restrict__ in, const float* __restrict__ flow,
                                                       const float* __restrict__ gx, const float* __restrict__ gy,
                                                       float* __restrict__ out, int C, int H, int W, float flow_scale,
                                                       float rcp_half_wm1, float rcp_half_hm1, int chunks, int chunk_c) {
    const int x = blockIdx.x * TX + (threadIdx.x % TX);
    const int y = blockIdx.y * (256 / TX) + (threadIdx.x / TX);
    const int b = blockIdx.z / chunks;
    const int c0 = (blockIdx.z - b * chunks) * chunk_c;
    const int c1 = min(C, c0 + chunk_c);
    if (x >= W || y >= H) return;
    const size_t plane = (size_t)H * W;
    const size_t pix = (size_t)y * W + x;
    const float* fl = flow + (size_t)b * 2 * plane + pix;
    // tenFlow * fltBackwarp, then / ((size - 1) / 2): ATen multiplies by the fp32 reciprocal of the scalar
    const float fx = __fmul_rn(__fmul_rn(__ldg(fl), flow_scale), rcp_half_wm1);
    const float fy = __fmul_rn(__fmul_rn(__ldg(fl + plane), flow_scale), rcp_half_hm1);
    const float cxn = __fadd_rn(__ldg(gx + x), fx);
    const float cyn = __fadd_rn(__ldg(gy + y), fy);
    // grid_sampler_unnormalize(align_corners = false): ((coord + 1) * size - 1) / 2, contracted like nvcc does
    const float ix = __fmul_rn(__fmaf_rn(__fadd_rn(cxn, 1.f), (float)W, -1.f), 0.5f);
    const float iy = __fmul_rn(__fmaf_rn(__fadd_rn(cyn, 1.f), (float)H, -1.f), 0.5f);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    // corner weights exactly as GridSampler.cu: nw = (ix_se - ix)(iy_se - iy), ne = (ix - ix_sw)(iy_sw - iy), ...
    const float x1 = __fadd_rn(fx0, 1.f), y1 = __fadd_rn(fy0, 1.f);
    const float wnw = __fmul_rn(__fsub_rn(x1, ix), __fsub_rn(y1, iy));
    const float wne = __fmul_rn(__fsub_rn(ix, fx0), __fsub_rn(y1, iy));
    const float wsw = __fmul_rn(__fsub_rn(x1, ix), __fsub_rn(iy, fy0));
    const float wse = __fmul_rn(__fsub_rn(ix, fx0), __fsub_rn(iy, fy0));
    // non-finite or absurd coordinates: every tap is out of bounds (ATen's int conversion is UB there)
    const bool sane = fabsf(ix) < 1.0e8f && fabsf(iy) < 1.0e8f;
    const int xi = sane ? (int)fx0 : -10, yi = sane ? (int)fy0 : -10;
    const bool okx0 = (unsigned)xi < (unsigned)W, okx1 = (unsigned)(xi + 1) < (unsigned)W;
    const bool oky0 = (unsigned)yi < (unsigned)H, oky1 = (unsigned)(yi + 1) < (unsigned)H;
    const bool onw = okx0 && oky0, one_ = okx1 && oky0, osw = okx0 && oky1, ose = okx1 && oky1;
    // the ones channel: the sum of the in-bounds weights, thresholded at 0.999 (ff_pwcnet.py:44)
    float m = 0.f;
    if (onw) m = __fadd_rn(m, wnw);
    if (one_) m = __fadd_rn(m, wne);
    if (osw) m = __fadd_rn(m, wsw);
    if (ose) m = __fadd_rn(m, wse);
    const float mask = (m > 0.999f) ? 1.f : 0.f;
    const int o00 = yi * W + xi;
    const float* ip = in + (size_t)b * C * plane;
    float* op = out + (size_t)b * C * plane + pix;
    if (mask == 0.f) {
        for (int c = c0; c < c1; ++c) op[(size_t)c * plane] = 0.f;
        return;
    }
#pragma unroll 4
    for (int c = c0; c < c1; ++c) {
        const float* p = ip + (size_t)c * plane + o00;
        float acc = 0.f;
        if (onw) acc = __fmaf_rn(__ldg(p), wnw, acc);
        if (one_) acc = __fmaf_rn(__ldg(p + 1), wne, acc);
        if (osw) acc = __fmaf_rn(__ldg(p + W), wsw, acc);
        if (ose) acc = __fmaf_rn(__ldg(p + W + 1), wse, acc);
        op[(size_t)c * plane] = __fmul_rn(acc, mask);
    }
}

}  // namespace
}  // namespace ffcorr

using namespace ffcorr;

extern "C" int ffcorr_backwarp_f32(const float* input, const float* flow, const float* grid_x, const float* grid_y,
                                   float* out, int B, int C, int H, int W, float flow_scale, void* stream) {
    FFCORR_REQUIRE(B >= 0 && C >= 1 && H >= 2 && W >= 2 && H <= 65535 * 4, FFCORR_EINVAL, "backwarp: B=%d C=%d H=%d W=%d", B, C, H, W);
    if (B == 0) return FFCORR_OK;
    FFCORR_REQUIRE(input && flow && grid_x && grid_y && out, FFCORR_EINVAL, "backwarp: null pointer");
    FFCORR_REQUIRE(B <= 65535, FFCORR_EINVAL, "backwarp: B=%d exceeds the grid z extent", B);
    const int tx = W > 32 ? 64 : (W > 16 ? 32 : 16);
    // channel chunks: enough blocks for ~8 per SM, at least 8 channels per block (the per-pixel set-up is repeated per chunk)
    const int64_t pix_blocks = (int64_t)ceil_div(W, tx) * ceil_div(H, 256 / tx) * B;
    int chunks = (int)std::min<int64_t>(ceil_div(C, 8), std::max<int64_t>(1, ceil_div64((int64_t)sm_count() * 8, pix_blocks)));
    const int chunk_c = ceil_div(C, chunks);
    chunks = ceil_div(C, chunk_c);
    FFCORR_REQUIRE((int64_t)B * chunks <= 65535, FFCORR_EINVAL, "backwarp: B * channel chunks = %lld exceeds the grid z extent", (long long)B * chunks);
    const dim3 grid((unsigned)ceil_div(W, tx), (unsigned)ceil_div(H, 256 / tx), (unsigned)(B * chunks));
    // (size - 1.0) / 2.0 is a Python double in the reference; ATen turns "tensor / scalar" into a multiplication by
    // the fp32 reciprocal of the scalar cast to float
    const float rw = 1.0f / (float)((W - 1.0) / 2.0);
    const float rh = 1.0f / (float)((H - 1.0) / 2.0);
    if (tx == 64)
        backwarp_kernel<64><<<grid, 256, 0, (cudaStream_t)stream>>>(input, flow, grid_x, grid_y, out, C, H, W, flow_scale, rw, rh, chunks, chunk_c);
    else if (tx == 32)
        backwarp_kernel<32><<<grid, 256, 0, (cudaStream_t)stream>>>(input, flow, grid_x, grid_y, out, C, H, W, flow_scale, rw, rh, chunks, chunk_c);
    else
        backwarp_kernel<16><<<grid, 256, 0, (cudaStream_t)stream>>>(input, flow, grid_x, grid_y, out, C, H, W, flow_scale, rw, rh, chunks, chunk_c);
    return check_launch("backwarp_kernel");
}
