/*
 * klt_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded CPU restatement of the arithmetic that the reference
 * (JonasFrey96/Visual-Odom-Pipeline) delegates to OpenCV at
 *     src/extractor/extractor.py:44,45,65,66   (cv2.calcOpticalFlowPyrLK call sites)
 *     src/extractor/extractor.py:16-19         (winSize / maxLevel / criteria)
 *     notebooks/tracking.py:39-40              (same call pattern, prototype)
 * The arithmetic itself lives in a third-party dependency that is NOT vendored under
 * /root/reference: OpenCV (pinned `opencv=4.4.0`, setup/conda_env.yml:57,78,87), functions
 * calcOpticalFlowPyrLK / buildOpticalFlowPyramid (modules/video/src/lkpyramid.cpp),
 * pyrDown (modules/imgproc/src/pyramids.cpp) and the Scharr derivative.  This file restates the
 * published algorithm as specified in SURVEY.md Appendix A (A.1 - A.7).
 *
 * Parity pinning: the reference has no tests / golden vectors for this path.
 * The oracle is therefore pinned against outputs of the reference's own implementation run
 * here -- the `cv2` module that the reference imports (4.13.0 in this image) -- in
 * tests/test_oracle.py (live) and against the committed .npz vectors in tests/golden
 * (made by tests/golden/make_golden.py).  Target: bit-exact nextPts / status / err(status==1).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 * The product path (visual-odom-pipeline_b200/) never links, imports or calls it.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: every float op rounds once, no FMA,
 * as in OpenCV's x86-64 SSE build).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KLT_TERM_COUNT 1
#define KLT_TERM_EPS 2
#define KLT_USE_INITIAL_FLOW 4
#define KLT_GET_MIN_EIGENVALS 8

/* BORDER_REFLECT_101 index map, any distance (-1->1, -2->2, n->n-2, ...) */
static int reflect101(int p, int len)
{
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p;
        else p = 2 * len - 2 - p;
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

/* ---- A.2 pyrDown: 5x5 separable [1 4 6 4 1], REFLECT_101, (sum+128)>>8 ------------------ */
int klt_oracle_pyr_down(const uint8_t* src, int w, int h, int64_t src_pitch,
                        uint8_t* dst, int64_t dst_pitch)
{
    static const int k[5] = {1, 4, 6, 4, 1};
    int dw = (w + 1) / 2, dh = (h + 1) / 2;
    if (w <= 0 || h <= 0) return -1;
    int* row = (int*)malloc(sizeof(int) * (size_t)w);
    if (!row) return -2;
    for (int y = 0; y < dh; ++y) {
        /* vertical pass at full horizontal resolution */
        for (int x = 0; x < w; ++x) {
            int s = 0;
            for (int j = -2; j <= 2; ++j)
                s += k[j + 2] * src[(int64_t)reflect101(2 * y + j, h) * src_pitch + x];
            row[x] = s;
        }
        for (int x = 0; x < dw; ++x) {
            int s = 0;
            for (int i = -2; i <= 2; ++i) s += k[i + 2] * row[reflect101(2 * x + i, w)];
            dst[(int64_t)y * dst_pitch + x] = (uint8_t)((s + 128) >> 8);
        }
    }
    free(row);
    return 0;
}

/* A.2: number of the last pyramid level actually built (buildOpticalFlowPyramid's return) */
int klt_oracle_pyr_max_level(int w, int h, int win_w, int win_h, int max_level)
{
    int level = 0;
    while (level < max_level) {
        int nw = (w + 1) / 2, nh = (h + 1) / 2;
        if (nw <= win_w || nh <= win_h) break;
        w = nw; h = nh; ++level;
    }
    return level;
}

/* ---- A.3 Scharr derivative, int16 interleaved (Ix,Iy), REFLECT_101 inside the image ---- */
int klt_oracle_scharr(const uint8_t* img, int w, int h, int64_t pitch, int16_t* deriv /* h*w*2 */)
{
    if (w <= 0 || h <= 0) return -1;
    int* t0 = (int*)malloc(sizeof(int) * (size_t)(w + 2));
    int* t1 = (int*)malloc(sizeof(int) * (size_t)(w + 2));
    if (!t0 || !t1) { free(t0); free(t1); return -2; }
    for (int y = 0; y < h; ++y) {
        const uint8_t* r0 = img + (int64_t)reflect101(y - 1, h) * pitch;
        const uint8_t* r1 = img + (int64_t)y * pitch;
        const uint8_t* r2 = img + (int64_t)reflect101(y + 1, h) * pitch;
        for (int x = -1; x <= w; ++x) {
            int xs = reflect101(x, w);
            t0[x + 1] = 3 * (r0[xs] + r2[xs]) + 10 * r1[xs]; /* vertical smooth */
            t1[x + 1] = r2[xs] - r0[xs];                     /* vertical difference */
        }
        for (int x = 0; x < w; ++x) {
            deriv[((int64_t)y * w + x) * 2 + 0] = (int16_t)(t0[x + 2] - t0[x]);
            deriv[((int64_t)y * w + x) * 2 + 1] = (int16_t)(3 * (t1[x] + t1[x + 2]) + 10 * t1[x + 1]);
        }
    }
    free(t0); free(t1);
    return 0;
}

/* one pyramid level with its win-sized borders: REFLECT_101 image, zero-padded derivative */
typedef struct {
    int w, h;          /* level size without border */
    int pw, ph;        /* border sizes (= win_w, win_h) */
    int64_t ipitch;    /* padded image pitch, bytes  */
    int64_t dpitch;    /* padded deriv pitch, int16 elements */
    uint8_t* ibuf;     /* padded image storage */
    int16_t* dbuf;     /* padded deriv storage (NULL for the `next` pyramid) */
} level_t;

static const uint8_t* lvl_img(const level_t* L) { return L->ibuf + (int64_t)L->ph * L->ipitch + L->pw; }
static const int16_t* lvl_der(const level_t* L) { return L->dbuf + (int64_t)L->ph * L->dpitch + 2 * (int64_t)L->pw; }

static int level_alloc(level_t* L, int w, int h, int pw, int ph, int with_deriv)
{
    L->w = w; L->h = h; L->pw = pw; L->ph = ph;
    L->ipitch = w + 2 * pw;
    L->dpitch = 2 * (int64_t)(w + 2 * pw);
    L->ibuf = (uint8_t*)malloc((size_t)L->ipitch * (size_t)(h + 2 * ph));
    L->dbuf = with_deriv ? (int16_t*)calloc((size_t)L->dpitch * (size_t)(h + 2 * ph), sizeof(int16_t)) : NULL;
    return (L->ibuf && (!with_deriv || L->dbuf)) ? 0 : -2;
}

static void level_free(level_t* L) { free(L->ibuf); free(L->dbuf); L->ibuf = NULL; L->dbuf = NULL; }

/* fill the borders of the padded image by REFLECT_101 of the level itself (A.2, last sentence) */
static void level_make_border(level_t* L)
{
    uint8_t* c = L->ibuf + (int64_t)L->ph * L->ipitch + L->pw;
    for (int y = -L->ph; y < L->h + L->ph; ++y) {
        int ys = reflect101(y, L->h);
        for (int x = -L->pw; x < L->w + L->pw; ++x) {
            if (y >= 0 && y < L->h && x >= 0 && x < L->w) continue;
            c[(int64_t)y * L->ipitch + x] = c[(int64_t)ys * L->ipitch + reflect101(x, L->w)];
        }
    }
}

static int build_pyramid(const uint8_t* img, int w, int h, int64_t pitch, int win_w, int win_h,
                         int max_level, int with_deriv, level_t* levels /* max_level+1 */)
{
    int rc = level_alloc(&levels[0], w, h, win_w, win_h, with_deriv);
    if (rc) return rc;
    uint8_t* c = (uint8_t*)lvl_img(&levels[0]);
    for (int y = 0; y < h; ++y) memcpy(c + (int64_t)y * levels[0].ipitch, img + (int64_t)y * pitch, (size_t)w);
    for (int l = 0;; ++l) {
        level_t* L = &levels[l];
        level_make_border(L);
        if (with_deriv) {
            int16_t* tmp = (int16_t*)malloc(sizeof(int16_t) * 2 * (size_t)L->w * (size_t)L->h);
            if (!tmp) return -2;
            klt_oracle_scharr(lvl_img(L), L->w, L->h, L->ipitch, tmp);
            int16_t* d = (int16_t*)lvl_der(L);
            for (int y = 0; y < L->h; ++y)
                memcpy(d + (int64_t)y * L->dpitch, tmp + (int64_t)y * L->w * 2, sizeof(int16_t) * 2 * (size_t)L->w);
            free(tmp);
        }
        if (l == max_level) break;
        int nw = (L->w + 1) / 2, nh = (L->h + 1) / 2;
        rc = level_alloc(&levels[l + 1], nw, nh, win_w, win_h, with_deriv);
        if (rc) return rc;
        klt_oracle_pyr_down(lvl_img(L), L->w, L->h, L->ipitch, (uint8_t*)lvl_img(&levels[l + 1]), levels[l + 1].ipitch);
    }
    return 0;
}

/* A.5 accumulator: four float lanes + a scalar tail, combined as t + ((q0+q2)+(q1+q3)) */
typedef struct { float q[4]; float t; } acc_t;
static void acc_zero(acc_t* a) { a->q[0] = a->q[1] = a->q[2] = a->q[3] = 0.f; a->t = 0.f; }
static float acc_total(const acc_t* a)
{
    float s02 = a->q[0] + a->q[2];
    float s13 = a->q[1] + a->q[3];
    float s = s02 + s13;
    return a->t + s;
}

/* A.4 step 3: Q14 bilinear weights */
static void q14_weights(float a, float b, int* w00, int* w01, int* w10, int* w11)
{
    float oa = 1.f - a, ob = 1.f - b;
    float p00 = oa * ob, p01 = a * ob, p10 = oa * b;
    *w00 = (int)lrintf(p00 * 16384.f);
    *w01 = (int)lrintf(p01 * 16384.f);
    *w10 = (int)lrintf(p10 * 16384.f);
    *w11 = 16384 - *w00 - *w01 - *w10;
}

static int out_of_range(int ix, int iy, int win_w, int win_h, int w, int h)
{
    return ix < -win_w || ix >= w || iy < -win_h || iy >= h;
}

/*
 * A.1 - A.6: the whole cv2.calcOpticalFlowPyrLK call on 1-channel u8 images.
 * next_pts is in/out (read only with KLT_USE_INITIAL_FLOW).  err of points whose err cv2 leaves
 * uninitialised (A.6) is written as 0.  iters_out (optional, n entries) receives the total number
 * of LK iterations executed per point over all levels (SURVEY.md s8d work model).
 * Returns the max level used (>= 0) or a negative error code.
 */
int klt_oracle_calc_optical_flow_pyr_lk(
    const uint8_t* prev_img, int64_t prev_pitch, const uint8_t* next_img, int64_t next_pitch,
    int w, int h, const float* prev_pts, float* next_pts, uint8_t* status, float* err, int n,
    int win_w, int win_h, int max_level, int crit_type, int crit_max_count, double crit_eps,
    int flags, double min_eig_threshold, int32_t* iters_out)
{
    if (max_level < 0 || win_w <= 2 || win_h <= 2 || w <= 0 || h <= 0 || n < 0) return -1;
    if (n == 0) return 0;

    /* A.2 criteria normalisation */
    int max_count; double eps;
    if ((crit_type & KLT_TERM_COUNT) == 0) max_count = 30;
    else max_count = crit_max_count < 0 ? 0 : (crit_max_count > 100 ? 100 : crit_max_count);
    if ((crit_type & KLT_TERM_EPS) == 0) eps = 0.01;
    else eps = crit_eps < 0. ? 0. : (crit_eps > 10. ? 10. : crit_eps);
    const double eps2 = eps * eps;
    const float min_eig_thr = (float)min_eig_threshold;

    max_level = klt_oracle_pyr_max_level(w, h, win_w, win_h, max_level);
    level_t* P = (level_t*)calloc((size_t)max_level + 1, sizeof(level_t));
    level_t* N = (level_t*)calloc((size_t)max_level + 1, sizeof(level_t));
    int16_t* Ibuf = (int16_t*)malloc(sizeof(int16_t) * (size_t)win_w * (size_t)win_h);
    int16_t* Dbuf = (int16_t*)malloc(sizeof(int16_t) * 2 * (size_t)win_w * (size_t)win_h);
    int rc = (P && N && Ibuf && Dbuf) ? 0 : -2;
    if (!rc) rc = build_pyramid(prev_img, w, h, prev_pitch, win_w, win_h, max_level, 1, P);
    if (!rc) rc = build_pyramid(next_img, w, h, next_pitch, win_w, win_h, max_level, 0, N);
    if (rc) goto done;

    for (int i = 0; i < n; ++i) { status[i] = 1; err[i] = 0.f; if (iters_out) iters_out[i] = 0; }

    const float hwx = (float)(win_w - 1) * 0.5f, hwy = (float)(win_h - 1) * 0.5f;
    const int nv = 8 * (win_w / 8); /* A.5: width of the vectorised part of each window row */

    for (int level = max_level; level >= 0; --level) {
        const level_t* LI = &P[level];
        const level_t* LJ = &N[level];
        const uint8_t* I = lvl_img(LI);
        const int16_t* dI = lvl_der(LI);
        const uint8_t* J = lvl_img(LJ);
        const int64_t stepI = LI->ipitch, stepJ = LJ->ipitch, dstep = LI->dpitch;
        const int lw = LI->w, lh = LI->h;
        const float scale = (float)(1. / (double)(1 << level));

        for (int p = 0; p < n; ++p) {
            /* A.4 step 1 */
            float px = prev_pts[2 * p] * scale, py = prev_pts[2 * p + 1] * scale;
            float nx, ny;
            if (level == max_level) {
                if (flags & KLT_USE_INITIAL_FLOW) { nx = next_pts[2 * p] * scale; ny = next_pts[2 * p + 1] * scale; }
                else { nx = px; ny = py; }
            } else { nx = next_pts[2 * p] * 2.f; ny = next_pts[2 * p + 1] * 2.f; }
            next_pts[2 * p] = nx; next_pts[2 * p + 1] = ny;

            /* step 2 */
            px -= hwx; py -= hwy;
            int ipx = (int)floorf(px), ipy = (int)floorf(py);
            if (!(px == px) || !(py == py) || out_of_range(ipx, ipy, win_w, win_h, lw, lh)) {
                if (level == 0) { status[p] = 0; err[p] = 0.f; }
                continue;
            }
            /* step 3 */
            int w00, w01, w10, w11;
            q14_weights(px - (float)ipx, py - (float)ipy, &w00, &w01, &w10, &w11);

            /* steps 4+5: patch of I (Q5) and of the derivative, covariance matrix G */
            acc_t a11, a12, a22; acc_zero(&a11); acc_zero(&a12); acc_zero(&a22);
            for (int y = 0; y < win_h; ++y) {
                const uint8_t* src = I + (int64_t)(y + ipy) * stepI + ipx;
                const int16_t* dsrc = dI + (int64_t)(y + ipy) * dstep + 2 * (int64_t)ipx;
                for (int x = 0; x < win_w; ++x) {
                    int iv = (src[x] * w00 + src[x + 1] * w01 + src[x + stepI] * w10 + src[x + stepI + 1] * w11 + (1 << 8)) >> 9;
                    int gx = (dsrc[2 * x] * w00 + dsrc[2 * x + 2] * w01 + dsrc[2 * x + dstep] * w10 + dsrc[2 * x + dstep + 2] * w11 + (1 << 13)) >> 14;
                    int gy = (dsrc[2 * x + 1] * w00 + dsrc[2 * x + 3] * w01 + dsrc[2 * x + dstep + 1] * w10 + dsrc[2 * x + dstep + 3] * w11 + (1 << 13)) >> 14;
                    Ibuf[y * win_w + x] = (int16_t)iv;
                    Dbuf[2 * (y * win_w + x)] = (int16_t)gx;
                    Dbuf[2 * (y * win_w + x) + 1] = (int16_t)gy;
                    if (x < nv) {
                        a11.q[x & 3] += (float)(gx * gx);
                        a12.q[x & 3] += (float)(gx * gy);
                        a22.q[x & 3] += (float)(gy * gy);
                    } else {
                        a11.t += (float)(gx * gx);
                        a12.t += (float)(gx * gy);
                        a22.t += (float)(gy * gy);
                    }
                }
            }
            const float FLT_SCALE = 1.f / (float)(1 << 20);
            float A11 = acc_total(&a11) * FLT_SCALE;
            float A12 = acc_total(&a12) * FLT_SCALE;
            float A22 = acc_total(&a22) * FLT_SCALE;
            float m1 = A11 * A22, m2 = A12 * A12;
            float D = m1 - m2;
            float dA = A11 - A22;
            float sq = dA * dA, fa = 4.f * A12, fb = fa * A12;
            float rad = sqrtf(sq + fb);
            float sum = A22 + A11;
            float minEig = (sum - rad) / (float)(2 * win_w * win_h);
            if (flags & KLT_GET_MIN_EIGENVALS) err[p] = minEig;
            if (minEig < min_eig_thr || D < 1.1920929e-7f) {
                if (level == 0) status[p] = 0;
                continue;
            }
            D = 1.f / D;

            /* step 6: iterations */
            nx -= hwx; ny -= hwy;
            float pdx = 0.f, pdy = 0.f;
            for (int j = 0; j < max_count; ++j) {
                int inx = (int)floorf(nx), iny = (int)floorf(ny);
                if (!(nx == nx) || !(ny == ny) || out_of_range(inx, iny, win_w, win_h, lw, lh)) {
                    if (level == 0) status[p] = 0;
                    break;
                }
                if (iters_out) iters_out[p]++;
                q14_weights(nx - (float)inx, ny - (float)iny, &w00, &w01, &w10, &w11);
                acc_t b1a, b2a; acc_zero(&b1a); acc_zero(&b2a);
                for (int y = 0; y < win_h; ++y) {
                    const uint8_t* Jp = J + (int64_t)(y + iny) * stepJ + inx;
                    const int16_t* Ip = Ibuf + y * win_w;
                    const int16_t* dp = Dbuf + 2 * y * win_w;
                    int x = 0;
#define DIFF(x_) (((Jp[(x_)] * w00 + Jp[(x_) + 1] * w01 + Jp[(x_) + stepJ] * w10 + Jp[(x_) + stepJ + 1] * w11 + (1 << 8)) >> 9) - Ip[(x_)])
                    for (; x < nv; x += 8) {
                        for (int l = 0; l < 4; ++l) {
                            int d0 = DIFF(x + l), d1 = DIFF(x + l + 4);
                            b1a.q[l] += (float)(d0 * dp[2 * (x + l)] + d1 * dp[2 * (x + l + 4)]);
                            b2a.q[l] += (float)(d0 * dp[2 * (x + l) + 1] + d1 * dp[2 * (x + l + 4) + 1]);
                        }
                    }
                    for (; x < win_w; ++x) {
                        int d = DIFF(x);
                        b1a.t += (float)(d * dp[2 * x]);
                        b2a.t += (float)(d * dp[2 * x + 1]);
                    }
                }
                float b1 = acc_total(&b1a) * FLT_SCALE;
                float b2 = acc_total(&b2a) * FLT_SCALE;
                float t1 = A12 * b2, t2 = A22 * b1, t3 = A12 * b1, t4 = A11 * b2;
                float dx = (t1 - t2) * D;
                float dy = (t3 - t4) * D;
                nx += dx; ny += dy;
                next_pts[2 * p] = nx + hwx; next_pts[2 * p + 1] = ny + hwy;
                if ((double)dx * (double)dx + (double)dy * (double)dy <= eps2) break;
                if (j > 0 && fabs((double)(dx + pdx)) < 0.01 && fabs((double)(dy + pdy)) < 0.01) {
                    next_pts[2 * p] -= dx * 0.5f; next_pts[2 * p + 1] -= dy * 0.5f;
                    break;
                }
                pdx = dx; pdy = dy;
            }

            /* step 7: err at level 0 */
            if (status[p] && level == 0 && (flags & KLT_GET_MIN_EIGENVALS) == 0) {
                float qx = next_pts[2 * p] - hwx, qy = next_pts[2 * p + 1] - hwy;
                int iqx = (int)floorf(qx), iqy = (int)floorf(qy);
                if (!(qx == qx) || !(qy == qy) || out_of_range(iqx, iqy, win_w, win_h, lw, lh)) {
                    status[p] = 0;
                    continue;
                }
                q14_weights(qx - (float)iqx, qy - (float)iqy, &w00, &w01, &w10, &w11);
                float e = 0.f;
                for (int y = 0; y < win_h; ++y) {
                    const uint8_t* Jp = J + (int64_t)(y + iqy) * stepJ + iqx;
                    const int16_t* Ip = Ibuf + y * win_w;
                    for (int x = 0; x < win_w; ++x) {
                        int d = DIFF(x);
                        e += fabsf((float)d);
                    }
                }
#undef DIFF
                err[p] = e * 1.f / (float)(32 * win_w * win_h);
            }
        }
    }
    rc = max_level;
done:
    if (P) for (int l = 0; l <= max_level; ++l) level_free(&P[l]);
    if (N) for (int l = 0; l <= max_level; ++l) level_free(&N[l]);
    free(P); free(N); free(Ibuf); free(Dbuf);
    return rc;
}
