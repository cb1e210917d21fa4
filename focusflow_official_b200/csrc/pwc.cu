// PWC 9x9 local cost volume (max displacement 4, 81 channels) for sm_100a.
//
// Replaces PWCNet_Core/correlation.py:278-328: two padded-NHWC "rearrange" copies
// (3 memsets + 2 full copies) followed by a one-block-per-pixel kernel with 81
// __syncthreads rounds.  Here ONE kernel reads NCHW directly (no padded copy):
//   out[b, (dy+4)*9 + (dx+4), y, x] = (1/C) * sum_c one[b,c,y,x] * two[b,c,y+dy,x+dx]
//
// Decomposition: a block owns a 4 x 32 pixel tile of one batch item and loops over the
// channels in chunks of 8 staged in shared memory (tile of `one`, tile + 4-pixel halo of
// `two`, zero filled outside the image).  Warp w (0..8) owns displacement row dy = w - 4;
// lane -> 4-pixel strip.  Per channel a thread reads 4 + 12 floats (4 x LDS.128) and issues
// 36 FMAs (4 pixels x 9 dx), accumulators stay in registers for all C channels.
// Roofline: 4*(2C+81) bytes and 162*C flop per pixel -> HBM-bound for C = 32, FP32-FMA
// bound for C >= 64.
#include "common.cuh"

namespace ffcorr {
namespace {

constexpr int PT_Y = 4;            // tile rows
constexpr int PT_X = 32;           // tile cols
constexpr int PCC = 8;             // channels per smem chunk
constexpr int PHALO = 4;
constexpr int PW2 = PT_X + 2 * PHALO;   // 40
constexpr int PH2 = PT_Y + 2 * PHALO;   // 12
constexpr int PWC_THREADS = 9 * 32;

__global__ void __launch_bounds__(PWC_THREADS) pwc81_kernel(const float* __restrict__ one, const float* __restrict__ two,
                                                            float* __restrict__ out, int C, int H, int W,
                                                            float leaky_slope) {
    __shared__ __align__(16) float s_one[PCC][PT_Y][PT_X];
    __shared__ __align__(16) float s_two[PCC][PH2][PW2];

    const int b = blockIdx.z;
    const int y0 = blockIdx.y * PT_Y, x0 = blockIdx.x * PT_X;
    const int tid = threadIdx.x;
    const int dyi = tid >> 5;            // 0..8  (warp-uniform)
    const int lane = tid & 31;
    const int row = lane >> 3;           // 0..3
    const int c4 = (lane & 7) * 4;       // strip start column within the tile

    const size_t plane = (size_t)H * W;
    const float* __restrict__ one_b = one + (size_t)b * C * plane;
    const float* __restrict__ two_b = two + (size_t)b * C * plane;

    float acc[4][9];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 9; ++j) acc[i][j] = 0.0f;

    for (int cb = 0; cb < C; cb += PCC) {
        // ---- stage `one` tile: PCC*4*32 = 1024 elements ----
        for (int i = tid; i < PCC * PT_Y * PT_X; i += PWC_THREADS) {
            const int x = i & (PT_X - 1);
            const int y = (i >> 5) & (PT_Y - 1);
            const int c = i >> 7;
            const int gy = y0 + y, gx = x0 + x, gc = cb + c;
            float v = 0.0f;
            if (gc < C && gy < H && gx < W) v = __ldg(one_b + (size_t)gc * plane + (size_t)gy * W + gx);
            (&s_one[0][0][0])[i] = v;
        }
        // ---- stage `two` tile + halo: PCC*12*40 = 3840 elements ----
        for (int i = tid; i < PCC * PH2 * PW2; i += PWC_THREADS) {
            const int c = i / (PH2 * PW2);
            const int r = i - c * (PH2 * PW2);
            const int y = r / PW2;
            const int x = r - y * PW2;
            const int gy = y0 + y - PHALO, gx = x0 + x - PHALO, gc = cb + c;
            float v = 0.0f;
            if (gc < C && (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W)
                v = __ldg(two_b + (size_t)gc * plane + (size_t)gy * W + gx);
            (&s_two[0][0][0])[i] = v;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < PCC; ++c) {
            const float4 a4 = *reinterpret_cast<const float4*>(&s_one[c][row][c4]);
            const float* tr = &s_two[c][row + dyi][c4];
            const float4 t0 = *reinterpret_cast<const float4*>(tr);
            const float4 t1 = *reinterpret_cast<const float4*>(tr + 4);
            const float4 t2 = *reinterpret_cast<const float4*>(tr + 8);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float t[12] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w, t2.x, t2.y, t2.z, t2.w};
#pragma unroll
            for (int px = 0; px < 4; ++px)
#pragma unroll
                for (int dx = 0; dx < 9; ++dx) acc[px][dx] = fmaf(a[px], t[px + dx], acc[px][dx]);
        }
        __syncthreads();
    }

    // ---- epilogue: /C (true division like correlation.py:97), optional fused leaky_relu ----
    const int gy = y0 + row;
    if (gy >= H) return;
    const float fc = (float)C;
    const int gx = x0 + c4;
    float* __restrict__ o = out + (((size_t)b * 81 + (size_t)dyi * 9) * H + gy) * W + gx;
    const bool vec = ((W & 3) == 0) && (gx + 3 < W) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
#pragma unroll
    for (int dx = 0; dx < 9; ++dx) {
        float r[4];
#pragma unroll
        for (int px = 0; px < 4; ++px) {
            float v = __fdiv_rn(acc[px][dx], fc);
            if (leaky_slope >= 0.0f) v = v > 0.0f ? v : v * leaky_slope;
            r[px] = v;
        }
        float* od = o + (size_t)dx * plane;
        if (vec) {
            *reinterpret_cast<float4*>(od) = make_float4(r[0], r[1], r[2], r[3]);
        } else {
#pragma unroll
            for (int px = 0; px < 4; ++px)
                if (gx + px < W) od[px] = r[px];
        }
    }
}

// Gradients (correlation.py:104-232).  One thread per input element, 81 taps each.
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
    dim3 grid(ceil_div(W, PT_X), ceil_div(H, PT_Y), B);
    FFCORR_REQUIRE(grid.y < 65536 && grid.z < 65536, FFCORR_EINVAL, "pwc81: grid too large");
    pwc81_kernel<<<grid, PWC_THREADS, 0, (cudaStream_t)stream>>>(one, two, out, C, H, W, leaky_slope);
    return check_launch("pwc81_kernel");
}

extern "C" int ffcorr_pwc81_bwd_f32(const float* one, const float* two, const float* grad_out, float* grad_one,
                                    float* grad_two, int B, int C, int H, int W, void* stream) {
    FFCORR_REQUIRE(B >= 0 && C >= 1 && H >= 1 && W >= 1, FFCORR_EINVAL, "pwc81_bwd: bad shape");
    if (B == 0) return FFCORR_OK;
    FFCORR_REQUIRE(one && two && grad_out, FFCORR_EINVAL, "pwc81_bwd: null pointer");
    if (B == 0 || (!grad_one && !grad_two)) return FFCORR_OK;
    const int64_t total = (int64_t)B * C * H * W;
    const int64_t want = ceil_div64(total, 256);
    const int64_t cap = (int64_t)sm_count() * 16;
    const int grid = (int)(want < cap ? want : cap);
    pwc81_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(one, two, grad_out, grad_one, grad_two, B, C, H, W);
    return check_launch("pwc81_bwd_kernel");
}
