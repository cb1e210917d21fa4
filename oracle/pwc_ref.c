/*
 * CPU restatement of the PWC 9x9 local cost volume of FocusFlow's FF-PWC variant.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by tests/ as the
 * checker and by bench.py as the timed CPU baseline.  Never linked into the
 * product library.
 *
 * Reference: /root/reference/core/models/ff-pwcnet/PWCNet_Core/correlation.py
 *   :7-32    kernel_Correlation_rearrange    NCHW -> zero-padded (pad 4) NHWC
 *   :34-102  kernel_Correlation_updateOutput 81 displacement channels, /C
 *   :104-166 kernel_Correlation_updateGradOne
 *   :168-232 kernel_Correlation_updateGradTwo
 * The reference cannot run without CuPy + a GPU (:320-321 raises on CPU), so
 * this file restates the kernels' index arithmetic AND their fp32 summation
 * order: 32 lane-strided partial sums over channels, then lane 0 adds the 32
 * partials serially and divides by C (:84-98).
 *
 * Build: gcc -O2 -shared -fPIC -ffp-contract=off (see oracle/Makefile).  The image's gcc has no
 * libgomp, so the omp pragmas are inert here; oracle/pwc_oracle.py threads over the batch instead.
 */
#include <stdlib.h>
#include <string.h>

#define PAD 4
#define NDISP 9
#define LANES 32

/* correlation.py:19-30: out[b, y+4, x+4, c] = in[b, c, y, x], border stays 0. */
static void pad_to_nhwc(const float *in, float *out, int B, int C, int H, int W)
{
    const int PH = H + 2 * PAD, PW = W + 2 * PAD;
    memset(out, 0, sizeof(float) * (size_t)B * PH * PW * C);
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                float *dst = out + (((size_t)b * PH + (y + PAD)) * PW + (x + PAD)) * C;
                for (int c = 0; c < C; ++c)
                    dst[c] = in[(((size_t)b * C + c) * H + y) * W + x];
            }
}

/*
 * correlation.py:46-98.  top[b, ch, y, x] with ch = (dy+4)*9 + (dx+4):
 * s2o = ch % 9 - 4 shifts x, s2p = ch / 9 - 4 shifts y (:71-72).
 */
int pwc_ref_forward(const float *one, const float *two, float *top,
                    int B, int C, int H, int W)
{
    const int PH = H + 2 * PAD, PW = W + 2 * PAD;
    const size_t padded = (size_t)B * PH * PW * C;
    float *r0 = (float *)malloc(sizeof(float) * padded);
    float *r1 = (float *)malloc(sizeof(float) * padded);
    if (!r0 || !r1) { free(r0); free(r1); return -1; }
    pad_to_nhwc(one, r0, B, C, H, W);
    pad_to_nhwc(two, r1, B, C, H, W);

#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const float *p0 = r0 + (((size_t)b * PH + (y + PAD)) * PW + (x + PAD)) * C;
                for (int ch = 0; ch < NDISP * NDISP; ++ch) {
                    const int dx = ch % NDISP - PAD, dy = ch / NDISP - PAD;
                    const float *p1 = r1 + (((size_t)b * PH + (y + PAD + dy)) * PW + (x + PAD + dx)) * C;
                    float part[LANES];
                    for (int l = 0; l < LANES; ++l) {
                        float s = 0.0f;
                        for (int c = l; c < C; c += LANES)
                            s += p0[c] * p1[c];
                        part[l] = s;
                    }
                    float tot = 0.0f;
                    for (int l = 0; l < LANES; ++l)
                        tot += part[l];
                    top[(((size_t)b * (NDISP * NDISP) + ch) * H + y) * W + x] = tot / (float)C;
                }
            }
    free(r0);
    free(r1);
    return 0;
}

/*
 * correlation.py:115-165 (stride 1 => the x/y ranges collapse to one position):
 * gradOne[b,c,y,x] = (1/C) sum_{p,o} gradOut[b,(p+4)*9+(o+4),y,x] * two_pad[b,y+p,x+o,c]
 * summed in (p, o) order in fp32.
 */
int pwc_ref_backward_one(const float *two, const float *gout, float *gone,
                         int B, int C, int H, int W)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    float s = 0.0f;
                    for (int p = -PAD; p <= PAD; ++p)
                        for (int o = -PAD; o <= PAD; ++o) {
                            const int yy = y + p, xx = x + o;
                            const float t = (yy >= 0 && yy < H && xx >= 0 && xx < W)
                                ? two[(((size_t)b * C + c) * H + yy) * W + xx] : 0.0f;
                            const int op = (p + PAD) * NDISP + (o + PAD);
                            s += gout[(((size_t)b * 81 + op) * H + y) * W + x] * t;
                        }
                    gone[(((size_t)b * C + c) * H + y) * W + x] = s / (float)C;
                }
    return 0;
}

/*
 * correlation.py:179-231:
 * gradTwo[b,c,y,x] = (1/C) sum_{p,o : (y-p, x-o) inside} gradOut[b,op,y-p,x-o] * one[b,c,y-p,x-o]
 */
int pwc_ref_backward_two(const float *one, const float *gout, float *gtwo,
                         int B, int C, int H, int W)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    float s = 0.0f;
                    for (int p = -PAD; p <= PAD; ++p)
                        for (int o = -PAD; o <= PAD; ++o) {
                            const int yy = y - p, xx = x - o;
                            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                            const int op = (p + PAD) * NDISP + (o + PAD);
                            s += gout[(((size_t)b * 81 + op) * H + yy) * W + xx]
                               * one[(((size_t)b * C + c) * H + yy) * W + xx];
                        }
                    gtwo[(((size_t)b * C + c) * H + y) * W + x] = s / (float)C;
                }
    return 0;
}
