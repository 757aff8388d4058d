/*
 * bilateral_oracle.c -- CPU restatement of cv2.bilateralFilter(src, d, sigmaColor, sigmaSpace) for 8-bit single-channel
 * images, as the reference's loader applies it to every frame (reference src/loader/loader.py:16-20,86: d = 5,
 * sigmaColor = sigmaSpace = 1.5).  TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load it.
 *
 * The arithmetic lives in OpenCV (imgproc/src/bilateral_filter.dispatch.cpp, bilateral_filter.simd.hpp; the reference
 * pins opencv=4.4.0, setup/conda_env.yml).  Restated here:
 *   B.1  radius = d / 2 (d > 0), at least 1; border BORDER_REFLECT_101 by `radius`;
 *   B.2  color_weight[i] = (float)exp(i * i * (-0.5 / sigmaColor^2)), i = 0..255 (double exp, then float);
 *   B.3  neighbours (i, j) in [-radius, radius]^2, rows outer, with sqrt(i^2 + j^2) <= radius, in that order:
 *        space_weight[k] = (float)exp(r^2 * (-0.5 / sigmaSpace^2));
 *   B.4  per pixel, float32: for k in order:  w = space_weight[k] * color_weight[|v_k - v_0|];  wsum += w;
 *        sum = fma(v_k, w, sum)   (OpenCV's universal intrinsics: v_muladd is a fused multiply-add in the AVX2 build);
 *   B.5  columns x >= W - W % 8 (scalar tail of the 8-lane loop): neighbours are taken four at a time, k = 4g .. 4g+3:
 *        wsum += (w0 + w2) + (w1 + w3);  sum += (p0 + p2) + (p1 + p3)  with p = v * w rounded (v_reduce_sum of a
 *        4-lane vector; no fused multiply-add), then the remaining neighbours one by one as  wsum += w; sum += v * w
 *        (fused or not: `tail_fma`, pinned against the wheel in tests/test_bilateral_cpu.py);
 *   B.6  dst = cvRound(sum / wsum) (float division, round half to even); no saturation needed.
 * Pinned bit-for-bit against live cv2.bilateralFilter with cv2.ipp.setUseIPP(False) -- OpenCV's own code path, which is
 * what an OpenCV build without IPP (e.g. the reference's conda package) runs; the IPP-enabled wheel of this image routes
 * the call to a closed-source primitive that differs by +-1 on about half of the pixels (tolerance test, same file).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = (p < 0) ? -p : 2 * len - 2 - p;
    return p;
}

/* space weights / offsets of B.3; returns the number of neighbours (<= d*d) */
int klt_oracle_bilateral_kernel(int d, double sigma_space, float* space_weight, int* di, int* dj)
{
    if (sigma_space <= 0) sigma_space = 1;
    const double gauss_space_coeff = -0.5 / (sigma_space * sigma_space);
    int radius = (d <= 0) ? (int)lrint(sigma_space * 1.5) : d / 2;
    if (radius < 1) radius = 1;
    int maxk = 0;
    for (int i = -radius; i <= radius; ++i)
        for (int j = -radius; j <= radius; ++j) {
            const double r = sqrt((double)i * i + (double)j * j);
            if (r > radius) continue;
            space_weight[maxk] = (float)exp(r * r * gauss_space_coeff);
            di[maxk] = i; dj[maxk] = j;
            ++maxk;
        }
    return maxk;
}

void klt_oracle_bilateral_color_weights(double sigma_color, float* color_weight /*[256]*/)
{
    if (sigma_color <= 0) sigma_color = 1;
    const double gauss_color_coeff = -0.5 / (sigma_color * sigma_color);
    for (int i = 0; i < 256; ++i) color_weight[i] = (float)exp(i * i * gauss_color_coeff);
}

static uint8_t round_div(float sum, float wsum) { return (uint8_t)lrintf(sum / wsum); }

/* simd_lanes: width of the vector loop whose scalar tail takes the B.5 path (8 for the AVX2 wheel; 0: no tail path) */
int klt_oracle_bilateral_filter(const uint8_t* src, int w, int h, int64_t pitch, int d, double sigma_color, double sigma_space,
                                int simd_lanes, int tail_fma, uint8_t* dst, int64_t dst_pitch)
{
    if (!src || !dst || w <= 0 || h <= 0 || pitch < w || dst_pitch < w) return -1;
    int radius = (d <= 0) ? (int)lrint((sigma_space <= 0 ? 1 : sigma_space) * 1.5) : d / 2;
    if (radius < 1) radius = 1;
    const int dd = 2 * radius + 1;
    float* sw = (float*)malloc(sizeof(float) * dd * dd);
    int* di = (int*)malloc(sizeof(int) * dd * dd);
    int* dj = (int*)malloc(sizeof(int) * dd * dd);
    float cw[256];
    if (!sw || !di || !dj) { free(sw); free(di); free(dj); return -4; }
    const int maxk = klt_oracle_bilateral_kernel(d, sigma_space, sw, di, dj);
    klt_oracle_bilateral_color_weights(sigma_color, cw);
    const int x_tail = (simd_lanes > 0) ? w - w % simd_lanes : w;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const int v0 = src[(int64_t)y * pitch + x];
            float sum = 0.f, wsum = 0.f;
            if (x < x_tail) {
                for (int k = 0; k < maxk; ++k) {
                    const int v = src[(int64_t)reflect101(y + di[k], h) * pitch + reflect101(x + dj[k], w)];
                    const float wt = sw[k] * cw[abs(v - v0)];
                    wsum += wt;
                    sum = fmaf((float)v, wt, sum);
                }
            } else {
                int k = 0;
                for (; k + 4 <= maxk; k += 4) {
                    float wt[4], p[4];
                    for (int q = 0; q < 4; ++q) {
                        const int v = src[(int64_t)reflect101(y + di[k + q], h) * pitch + reflect101(x + dj[k + q], w)];
                        wt[q] = sw[k + q] * cw[abs(v - v0)];
                        p[q] = (float)v * wt[q];
                    }
                    wsum += (wt[0] + wt[2]) + (wt[1] + wt[3]);
                    sum += (p[0] + p[2]) + (p[1] + p[3]);
                }
                for (; k < maxk; ++k) {
                    const int v = src[(int64_t)reflect101(y + di[k], h) * pitch + reflect101(x + dj[k], w)];
                    const float wt = sw[k] * cw[abs(v - v0)];
                    wsum += wt;
                    if (tail_fma) sum = fmaf((float)v, wt, sum);
                    else sum += (float)v * wt;
                }
            }
            dst[(int64_t)y * dst_pitch + x] = round_div(sum, wsum);
        }
    free(sw); free(di); free(dj);
    return 0;
}
