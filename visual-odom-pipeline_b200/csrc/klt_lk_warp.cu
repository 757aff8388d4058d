// K4 (bulk shape): fused Scharr + pyramidal LK, ONE WARP PER KEYPOINT, patch in shared memory.
//
// Same arithmetic as klt_lk_fast.cu / klt_lk.cu (bit-exact with cv2.calcOpticalFlowPyrLK as called
// at reference src/extractor/extractor.py:44,45,65,66).  What differs is the shape, chosen from the ncu profile of the
// team kernel (profiles/r01: 41.3 M warp-instructions per 2000-point launch, 9.3 thread-instructions per algorithmic
// MAC, 27 % warps active): there the per-iteration scalar work (floor / range test, Q14 weights, 2x2 solve, termination
// tests) is replicated in every warp of a point's team and the patch lives in registers, which caps occupancy.  Here
//  * one warp owns a point, so the scalar work is issued once per iteration;
//  * the Q5 intensity / Q14 derivative patch of a level lives in shared memory (40 bytes per 4-pixel unit), not in
//    registers: ~64 registers per thread, every point of a 2000-point frame pair is resident at once (no waves);
//  * the intensity is stored as the constant 2^8 - 2^9 * I, which is fed to the first dp2a of the bilinear tap as its
//    accumulator: ((J_bilinear + 2^8) >> 9) - I  ==  (J_bilinear + 2^8 - 2^9 I) >> 9  (arithmetic shift), so the
//    subtraction and one unpack per pixel disappear;
//  * the window is cut into 4-pixel units enumerated SIMD part first, scalar tail (x >= 8 * floor(w / 8), A.5) last, so
//    that a lane's unit slot is compile-time "SIMD lane classes" or "tail" (one mixed slot) and the five per-class
//    integer sums of b1, b2 and of the exactness bound come out of the pixel loop directly: there is one test per
//    iteration (every accumulation class of OpenCV's float32 sums stays below 2^24 => float sum == integer sum)
//    instead of the team kernel's whole-window test followed by the per-class test;
//  * a point that spends `budget` iterations on one level is handed to the long-point kernel (klt_lk_fast.cu); the
//    rare iteration that fails the class test replays OpenCV's accumulation order serially inside the warp.
#include "klt_common.cuh"

namespace klt {

namespace {

constexpr int kWarpsPerCta = 4;
constexpr int kM = 3;  // margin of the staged next-image region
constexpr unsigned kFull = 0xffffffffu;
constexpr int kExact = 1 << 24;

__host__ __device__ constexpr int r4(int v) { return (v + 3) / 4 * 4; }
__host__ __device__ constexpr int r16(int v) { return (v + 15) / 16 * 16; }
__host__ __device__ constexpr int cmax(int a, int b) { return a > b ? a : b; }

template <int WW, int WH>
struct WC {
    static constexpr int NV = 8 * (WW / 8);           // width of OpenCV's SIMD part (A.5)
    static constexpr int TL = WW - NV;                // scalar tail
    static constexpr int NS = NV / 8;                 // 8-pixel SIMD steps per row
    static constexpr int SPR = NV / 4;                // SIMD units per window row
    static constexpr int TPR = (TL + 3) / 4;          // tail units per window row
    static constexpr int NSU = WH * SPR, NTU = WH * TPR, NU = NSU + NTU;
    static constexpr int UPT = (NU + 31) / 32;        // unit slots per lane
    static constexpr int NUP = 32 * UPT;
    static constexpr int SI = r4(WW + 6), IR = WH + 3;              // prev-image region: row stride (bytes), rows
    static constexpr int SD = r4(WW + 2), DR = WH + 1;              // Scharr region (border windows): words per row, rows
    static constexpr int JW = WW + 1 + 2 * kM, JR = WH + 1 + 2 * kM, SJ = r4(JW + 3);   // next-image region
    static constexpr int RPR = (WW + 1 + 3) / 4, NRUN = DR * RPR;   // Scharr: 4-position runs
    // shared-memory slice of one warp (bytes)
    static constexpr int OFF_J = 0;
    static constexpr int OFF_PC = OFF_J + r16(SJ * JR);             // int  [NUP][4]: 2^8 - 2^9 * I (Q5)
    static constexpr int OFF_PG = OFF_PC + NUP * 16;                // u32  [NUP][4]: gx | gy << 16 (Q14-weighted Scharr)
    static constexpr int OFF_PM = OFF_PG + NUP * 16;                // u16  [NUP][4]: max(|gx|, |gy|)
    static constexpr int OFF_I = OFF_PM + NUP * 8;                  // prev-image region, then the Scharr region
    static constexpr int OFF_D = OFF_I + r16(SI * IR);
    static constexpr int WARP_BYTES = (OFF_D + r16(4 * SD * DR) + 32 + 127) / 128 * 128;   // + 32: padded pixels of the last unit read past the last row
    static_assert(WW * WH <= 2056, "err pass assumes an exact float32 sum");
    static_assert(UPT <= 8, "per-lane bound accumulators would overflow");
    static_assert(TPR >= 1 && TL >= 1, "windows whose width is a multiple of 8 are served by the generic kernel");
};

__device__ __forceinline__ int dp2a_lo(uint32_t w, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi(uint32_t w, uint32_t b, int c)
{
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int lo16(uint32_t w) { return (int)(short)(w & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t w) { return ((int)w) >> 16; }

__device__ __forceinline__ void q14_weights(float a, float b, int& w00, int& w01, int& w10, int& w11)
{
    const float oa = __fsub_rn(1.f, a), ob = __fsub_rn(1.f, b);
    w00 = __float2int_rn(__fmul_rn(__fmul_rn(oa, ob), 16384.f));
    w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, ob), 16384.f));
    w10 = __float2int_rn(__fmul_rn(__fmul_rn(oa, b), 16384.f));
    w11 = 16384 - w00 - w01 - w10;
}

// cvFloor + the window range test of A.4 step 2, in the float domain: floor(x) is an integer-valued float, so
// "-win <= floor(x) < len" can be decided on it directly; NaN / inf fail every comparison (out of range, like cv2's
// INT_MIN conversion).  fx / fy: the floors (as floats: nx - fx is the exact fractional part OpenCV computes).
__device__ __forceinline__ bool floor_in_range(float x, float y, int win_w, int win_h, float lwm1, float lhm1,
                                               float& fx, float& fy)
{
    fx = floorf(x); fy = floorf(y);
    return (fx >= (float)(-win_w)) && (fx <= lwm1) && (fy >= (float)(-win_h)) && (fy <= lhm1);
}

__device__ __forceinline__ float combine5(float q0, float q1, float q2, float q3, float t)
{
    const float s = __fadd_rn(__fadd_rn(q0, q2), __fadd_rn(q1, q3));
    return __fmul_rn(__fadd_rn(t, s), 9.5367431640625e-07f);
}

// Stage ROWS x STRIDE bytes whose top-left image coordinate is (ax, y0) (ax % 4 == 0) into smem with the 32 lanes of a
// warp.  Columns [c0, c0 + need) of every row are the ones later read.  Fast path (the staged columns lie inside the
// image and rows are 4-byte aligned): a lane keeps one word column and walks down the rows, RPT rows per trip, with all
// loads in flight before the first store; REFLECT_101 is needed for the rows only (warp-uniform branch).
template <int ROWS, int STRIDE>
__device__ __forceinline__ void stage(uint8_t* __restrict__ dst, const LevelView& lv, const uint8_t* __restrict__ img,
                                      int ax, int y0, int c0, int need, int lane)
{
    constexpr int NWR = STRIDE / 4;          // words per row
    constexpr int RPT = 32 / NWR;            // rows per trip
    constexpr int TRIPS = (ROWS + RPT - 1) / RPT;
    static_assert(RPT >= 1, "row wider than a warp");
    const bool fast = lv.aligned4 && ax >= 0 && (ax + STRIDE <= lv.w);
    if (fast) {  // warp-uniform
        const int rl = lane / NWR, cl = lane - rl * NWR;
        const bool lane_on = rl < RPT;
        uint32_t v[TRIPS];
        if ((y0 >= 0) && (y0 + ROWS <= lv.h)) {
            const uint8_t* __restrict__ p = img + (long long)(y0 + rl) * lv.pitch + ax + 4 * cl;
            const long long step = (long long)RPT * lv.pitch;
#pragma unroll
            for (int t = 0; t < TRIPS; ++t) {
                v[t] = (lane_on && (rl + t * RPT < ROWS)) ? __ldg(reinterpret_cast<const uint32_t*>(p)) : 0u;
                p += step;
            }
        } else {
#pragma unroll
            for (int t = 0; t < TRIPS; ++t) {
                const int yy = reflect101(y0 + rl + t * RPT, lv.h);
                v[t] = (lane_on && (rl + t * RPT < ROWS)) ? __ldg(reinterpret_cast<const uint32_t*>(img + (long long)yy * lv.pitch + ax) + cl) : 0u;
            }
        }
        uint32_t* __restrict__ d = reinterpret_cast<uint32_t*>(dst) + rl * NWR + cl;
#pragma unroll
        for (int t = 0; t < TRIPS; ++t)
            if (lane_on && (rl + t * RPT < ROWS)) d[t * RPT * NWR] = v[t];
    } else {
        // the columns cross the image border (or rows are unaligned): a lane keeps one (reflected) column, bytes
        constexpr int CSETS = (STRIDE + 31) / 32;
#pragma unroll
        for (int cs = 0; cs < CSETS; ++cs) {
            const int c = c0 + lane + 32 * cs;
            if (c < c0 + need) {
                const uint8_t* __restrict__ col = img + reflect101(ax + c, lv.w);
                uint8_t* __restrict__ d = dst + c;
#pragma unroll 4
                for (int r = 0; r < ROWS; ++r) d[r * STRIDE] = __ldg(col + (long long)reflect101(y0 + r, lv.h) * lv.pitch);
            }
        }
    }
}

// geometry of unit slot k of this lane: window row, first column, number of window pixels (0: slot unused)
template <int WW, int WH>
__device__ __forceinline__ void unit_geom(int k, int lane, int& y, int& x0, int& nvalid)
{
    using C = WC<WW, WH>;
    const int u = lane + 32 * k;
    if (u < C::NSU) {
        y = u / C::SPR; x0 = 4 * (u - y * C::SPR); nvalid = 4;
    } else if (u < C::NU) {
        const int t = u - C::NSU;
        y = t / C::TPR; x0 = C::NV + 4 * (t - y * C::TPR); nvalid = min(4, WW - x0);
    } else {
        y = 0; x0 = 0; nvalid = 0;
    }
}

// Adds the four per-pixel values of unit slot k into the class accumulators: SIMD slots feed lane class (x & 3),
// tail slots the tail class; the slot that straddles the boundary selects per lane.
template <int WW, int WH, typename T>
__device__ __forceinline__ void class_add(int k, int lane, T (&q)[4], T& t, const T (&v)[4])
{
    using C = WC<WW, WH>;
    if (32 * k + 32 <= C::NSU) {
#pragma unroll
        for (int j = 0; j < 4; ++j) q[j] += v[j];
    } else if (32 * k >= C::NSU) {
        t += v[0] + v[1] + v[2] + v[3];
    } else {
        const bool tail = lane + 32 * k >= C::NSU;
#pragma unroll
        for (int j = 0; j < 4; ++j) q[j] += tail ? (T)0 : v[j];
        t += tail ? (v[0] + v[1] + v[2] + v[3]) : (T)0;
    }
}

// word index of window pixel (y, x) in the patch arrays
template <int WW, int WH>
__device__ __forceinline__ int patch_index(int y, int x)
{
    using C = WC<WW, WH>;
    return x < C::NV ? (y * C::SPR + (x >> 2)) * 4 + (x & 3) : (C::NSU + y * C::TPR + ((x - C::NV) >> 2)) * 4 + ((x - C::NV) & 3);
}

template <int WW, int WH>
__device__ __forceinline__ void warp_point(const LKLaunch& L, const long long gid, uint8_t* ws, const int lane)
{
    using C = WC<WW, WH>;
    const int bidx = (int)(gid / L.n_per_pair);
    uint8_t* jreg = ws + C::OFF_J;
    int* pc = reinterpret_cast<int*>(ws + C::OFF_PC);
    uint32_t* pg = reinterpret_cast<uint32_t*>(ws + C::OFF_PG);
    uint16_t* pm = reinterpret_cast<uint16_t*>(ws + C::OFF_PM);
    uint8_t* ireg = ws + C::OFF_I;
    uint32_t* dreg = reinterpret_cast<uint32_t*>(ws + C::OFF_D);

    // word offset of each unit slot inside the staged next-image region
    int jw[C::UPT];
#pragma unroll
    for (int k = 0; k < C::UPT; ++k) {
        int y, x0, nvl;
        unit_geom<WW, WH>(k, lane, y, x0, nvl);
        jw[k] = (y * C::SJ + x0) >> 2;
    }

    const long long t_start = clock64();
    int n_t2 = 0, n_g = 0;
    const float2 p0 = reinterpret_cast<const float2*>(L.prev_pts)[gid];
    float2 outp = make_float2(0.f, 0.f);
    if (L.flags & KLT_OPTFLOW_USE_INITIAL_FLOW) outp = reinterpret_cast<const float2*>(L.next_pts)[gid];
    int status = 1;
    float err = 0.f;
    int iters = 0;
    const float hwx = (float)(WW - 1) * 0.5f, hwy = (float)(WH - 1) * 0.5f;
    const int top = L.prev.top;
    const int budget = (L.wl != nullptr) ? L.budget : 0;   // iterations per level before the hand-off (0: never)

    for (int level = top; level >= 0; --level) {
        const LevelView lvI = L.prev.lv[level];
        const LevelView lvJ = L.next.lv[level];
        const uint8_t* __restrict__ imgI = lvI.data + (long long)bidx * lvI.batch_stride;
        const uint8_t* __restrict__ imgJ = lvJ.data + (long long)bidx * lvJ.batch_stride;
        const int lw = lvI.w, lh = lvI.h;
        const float lwm1 = (float)(lw - 1), lhm1 = (float)(lh - 1);
        const float scale = __int_as_float((127 - level) << 23);

        float px = __fmul_rn(p0.x, scale), py = __fmul_rn(p0.y, scale);
        float nx, ny;
        if (level == top) {
            if (L.flags & KLT_OPTFLOW_USE_INITIAL_FLOW) { nx = __fmul_rn(outp.x, scale); ny = __fmul_rn(outp.y, scale); }
            else { nx = px; ny = py; }
        } else {
            nx = __fmul_rn(outp.x, 2.f); ny = __fmul_rn(outp.y, 2.f);
        }
        outp = make_float2(nx, ny);

        px = __fsub_rn(px, hwx); py = __fsub_rn(py, hwy);
        float fpx, fpy;
        if (!floor_in_range(px, py, WW, WH, lwm1, lhm1, fpx, fpy)) {
            if (level == 0) { status = 0; err = 0.f; }
            continue;
        }
        const int ipx = (int)fpx, ipy = (int)fpy;
        int w00, w01, w10, w11;
        q14_weights(__fsub_rn(px, fpx), __fsub_rn(py, fpy), w00, w01, w10, w11);

        nx = __fsub_rn(nx, hwx); ny = __fsub_rn(ny, hwy);
        // ---- stage both neighbourhoods (the previous level's readers are this warp) ---------------------------------
        __syncwarp();
        int jax = 0, jy0 = INT_MIN / 2, jx0 = INT_MIN / 2;  // region origin: smem col 0 <-> image x = jax; window columns start at jx0
        {
            float fnx, fny;
            if (floor_in_range(nx, ny, WW, WH, lwm1, lhm1, fnx, fny)) {
                jx0 = (int)fnx - kM; jy0 = (int)fny - kM; jax = jx0 & ~3;
                stage<C::JR, C::SJ>(jreg, lvJ, imgJ, jax, jy0, jx0 - jax, C::JW, lane);
            }
        }
        const int iax = (ipx - 1) & ~3;
        const int oi = (ipx - 1) - iax;
        stage<C::IR, C::SI>(ireg, lvI, imgI, iax, ipy - 1, oi, WW + 3, lane);
        __syncwarp();

        // ---- patch pass: Q5 intensity + Q14 derivative patch into shared memory, integer class sums of G -------------
        const uint32_t W0 = (uint32_t)(w00 & 0xffff) | ((uint32_t)w01 << 16);
        const uint32_t W1 = (uint32_t)(w10 & 0xffff) | ((uint32_t)w11 << 16);
        unsigned q11[4] = {0, 0, 0, 0}, q22[4] = {0, 0, 0, 0}, t11 = 0, t22 = 0;
        int q12[4] = {0, 0, 0, 0}, t12 = 0;
        // all (WW+1) x (WH+1) derivative positions inside the image <=> no zero-masking of the derivative
        const bool interior = (ipx >= 0) && (ipy >= 0) && (ipx + WW < lw) && (ipy + WH < lh);
        if (!interior) {
            // border window: Scharr at the (WW+1) x (WH+1) integer positions, zero outside the image (A.3)
            for (int u = lane; u < C::NRUN; u += 32) {
                const int dy = u / C::RPR;
                const int dx0 = 4 * (u - dy * C::RPR);
                const uint8_t* r0 = ireg + dy * C::SI + oi + dx0;
                const uint8_t* r1 = r0 + C::SI;
                const uint8_t* r2 = r1 + C::SI;
                int t0[6], t1[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const int a = r0[k], b = r1[k], cc = r2[k];
                    t0[k] = 3 * (a + cc) + 10 * b;
                    t1[k] = cc - a;
                }
                const bool yin = (unsigned)(ipy + dy) < (unsigned)lh;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int dx = dx0 + k;
                    const int gx = t0[k + 2] - t0[k];
                    const int gy = 3 * (t1[k] + t1[k + 2]) + 10 * t1[k + 1];
                    const bool in = yin && ((unsigned)(ipx + dx) < (unsigned)lw);
                    if (dx < C::SD) dreg[dy * C::SD + dx] = in ? (((uint32_t)gx & 0xffffu) | ((uint32_t)gy << 16)) : 0u;
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int k = 0; k < C::UPT; ++k) {
            int y, x0, nvl;
            unit_geom<WW, WH>(k, lane, y, x0, nvl);
            int iv[4], gx[4], gy[4];
            if (interior) {
                // Scharr is linear and so is the Q14 bilinear tap, so  sum_c w_c * Scharr(I)(p + c)  ==  Scharr(T)(p)  with
                // T(q) = sum_c w_c * I(q + c) the UNROUNDED bilinear sum (<= 255 * 2^14); exact in int32 (|.| < 2^27).
                const int sh = (oi & 3) * 8;
                uint32_t pa[4], pb[4], pcw[4], pd[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const uint32_t* wp = reinterpret_cast<const uint32_t*>(ireg + (y + r) * C::SI) + ((oi + x0) >> 2);
                    const uint32_t v0 = wp[0], v1 = wp[1], v2 = wp[2];
                    pa[r] = __funnelshift_r(v0, v1, sh);        // bytes c0 .. c0+3   (c0 = column of window x0-1)
                    pb[r] = __funnelshift_rc(v0, v1, sh + 8);   // bytes c0+1 .. c0+4
                    pcw[r] = __funnelshift_r(v1, v2, sh);       // bytes c0+4 .. c0+7
                    pd[r] = __funnelshift_rc(v1, v2, sh + 8);   // bytes c0+5 .. c0+8
                }
                int T[3][6];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    T[r][0] = dp2a_lo(W1, pa[r + 1], dp2a_lo(W0, pa[r], 0));
                    T[r][1] = dp2a_lo(W1, pb[r + 1], dp2a_lo(W0, pb[r], 0));
                    T[r][2] = dp2a_hi(W1, pa[r + 1], dp2a_hi(W0, pa[r], 0));
                    T[r][3] = dp2a_hi(W1, pb[r + 1], dp2a_hi(W0, pb[r], 0));
                    T[r][4] = dp2a_lo(W1, pcw[r + 1], dp2a_lo(W0, pcw[r], 0));
                    T[r][5] = dp2a_lo(W1, pd[r + 1], dp2a_lo(W0, pd[r], 0));
                }
                int t0[6], t1[6];
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    t0[c] = 3 * (T[0][c] + T[2][c]) + 10 * T[1][c];
                    t1[c] = T[2][c] - T[0][c];
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    iv[j] = (T[1][j + 1] + (1 << 8)) >> 9;
                    gx[j] = (t0[j + 2] - t0[j] + (1 << 13)) >> 14;
                    gy[j] = (3 * (t1[j] + t1[j + 2]) + 10 * t1[j + 1] + (1 << 13)) >> 14;
                }
            } else {
                const int o = oi + 1 + x0;
                const uint32_t* wa = reinterpret_cast<const uint32_t*>(ireg + (y + 1) * C::SI) + (o >> 2);
                const uint32_t* wb = reinterpret_cast<const uint32_t*>(ireg + (y + 2) * C::SI) + (o >> 2);
                const int s = (o & 3) * 8;
                const uint32_t a0 = __funnelshift_r(wa[0], wa[1], s), b0 = __funnelshift_rc(wa[0], wa[1], s + 8);
                const uint32_t a1 = __funnelshift_r(wb[0], wb[1], s), b1 = __funnelshift_rc(wb[0], wb[1], s + 8);
                iv[0] = dp2a_lo(W1, a1, dp2a_lo(W0, a0, 256)) >> 9;
                iv[1] = dp2a_lo(W1, b1, dp2a_lo(W0, b0, 256)) >> 9;
                iv[2] = dp2a_hi(W1, a1, dp2a_hi(W0, a0, 256)) >> 9;
                iv[3] = dp2a_hi(W1, b1, dp2a_hi(W0, b0, 256)) >> 9;
                const uint32_t* d0 = dreg + y * C::SD + x0;
                const uint32_t* d1 = d0 + C::SD;
                const uint4 e0 = *reinterpret_cast<const uint4*>(d0);
                const uint4 e1 = *reinterpret_cast<const uint4*>(d1);
                const uint32_t r0w[5] = {e0.x, e0.y, e0.z, e0.w, d0[4]};
                const uint32_t r1w[5] = {e1.x, e1.y, e1.z, e1.w, d1[4]};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    gx[j] = (lo16(r0w[j]) * w00 + lo16(r0w[j + 1]) * w01 + lo16(r1w[j]) * w10 + lo16(r1w[j + 1]) * w11 + (1 << 13)) >> 14;
                    gy[j] = (hi16(r0w[j]) * w00 + hi16(r0w[j + 1]) * w01 + hi16(r1w[j]) * w10 + hi16(r1w[j + 1]) * w11 + (1 << 13)) >> 14;
                }
            }
            unsigned u11[4], u22[4];
            int u12[4];
            int cst[4];
            uint32_t gw[4];
            uint32_t gmax[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool valid = j < nvl;
                const int vx = valid ? gx[j] : 0, vy = valid ? gy[j] : 0;
                cst[j] = 256 - 512 * iv[j];
                gw[j] = ((uint32_t)vx & 0xffffu) | ((uint32_t)vy << 16);
                gmax[j] = (uint32_t)max(abs(vx), abs(vy));
                u11[j] = (unsigned)(vx * vx); u12[j] = vx * vy; u22[j] = (unsigned)(vy * vy);
            }
            const int u = lane + 32 * k;
            *reinterpret_cast<int4*>(pc + 4 * u) = make_int4(cst[0], cst[1], cst[2], cst[3]);
            *reinterpret_cast<uint4*>(pg + 4 * u) = make_uint4(gw[0], gw[1], gw[2], gw[3]);
            *reinterpret_cast<uint2*>(pm + 4 * u) = make_uint2(gmax[0] | (gmax[1] << 16), gmax[2] | (gmax[3] << 16));
            class_add<WW, WH>(k, lane, q11, t11, u11);
            class_add<WW, WH>(k, lane, q12, t12, u12);
            class_add<WW, WH>(k, lane, q22, t22, u22);
        }
        __syncwarp();

        float A11, A12, A22;
        {
            const unsigned cap = 1u << 25;  // keeps the warp totals below 2^31; a capped lane fails the test below anyway
            unsigned g11[5], g22[5];
            int g12[5];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                g11[j] = __reduce_add_sync(kFull, min(q11[j], cap));
                g22[j] = __reduce_add_sync(kFull, min(q22[j], cap));
                g12[j] = __reduce_add_sync(kFull, max(min(q12[j], (int)cap), -(int)cap));
            }
            g11[4] = __reduce_add_sync(kFull, min(t11, cap));
            g22[4] = __reduce_add_sync(kFull, min(t22, cap));
            g12[4] = __reduce_add_sync(kFull, max(min(t12, (int)cap), -(int)cap));
            // A11 / A22: non-negative terms, exact iff every class total <= 2^24; A12: |gx gy| <= (gx^2 + gy^2) / 2
            bool ok = true;
#pragma unroll
            for (int j = 0; j < 5; ++j) ok = ok && g11[j] <= (unsigned)kExact && g22[j] <= (unsigned)kExact;
            if (ok) {
                A11 = combine5((float)g11[0], (float)g11[1], (float)g11[2], (float)g11[3], (float)g11[4]);
                A12 = combine5((float)g12[0], (float)g12[1], (float)g12[2], (float)g12[3], (float)g12[4]);
                A22 = combine5((float)g22[0], (float)g22[1], (float)g22[2], (float)g22[3], (float)g22[4]);
            } else {
                // serial replay in OpenCV's order (A.5), products formed on the fly from the patch: lane = 5 * sum + class
                // (classes 0..3: SIMD lanes x & 3 == class over x < NV; class 4: scalar tail), rows in order
                ++n_g;
                const int sm = lane / 5, cl = lane - 5 * sm;
                const bool live = lane < 15;
                float acc = 0.f;
                for (int y = 0; y < WH; ++y) {
#pragma unroll 1
                    for (int s = 0; s < cmax(C::SPR, C::TL); ++s) {
                        const int x = (cl < 4) ? (4 * s + cl) : (C::NV + s);
                        const bool on = live && ((cl < 4) ? (s < C::SPR) : (s < C::TL));
                        const uint32_t g = pg[patch_index<WW, WH>(y, on ? x : 0)];
                        const int vx = lo16(g), vy = hi16(g);
                        const int prod = (sm == 0 ? vx : vy) * (sm == 2 ? vy : vx);   // gx*gx, gx*gy (sm 1: vy*vx), gy*gy
                        if (on) acc = __fadd_rn(acc, (float)prod);
                    }
                }
                A11 = combine5(__shfl_sync(kFull, acc, 0), __shfl_sync(kFull, acc, 1), __shfl_sync(kFull, acc, 2), __shfl_sync(kFull, acc, 3),
                               __shfl_sync(kFull, acc, 4));
                A12 = combine5(__shfl_sync(kFull, acc, 5), __shfl_sync(kFull, acc, 6), __shfl_sync(kFull, acc, 7), __shfl_sync(kFull, acc, 8),
                               __shfl_sync(kFull, acc, 9));
                A22 = combine5(__shfl_sync(kFull, acc, 10), __shfl_sync(kFull, acc, 11), __shfl_sync(kFull, acc, 12), __shfl_sync(kFull, acc, 13),
                               __shfl_sync(kFull, acc, 14));
            }
        }
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dA = __fsub_rn(A11, A22);
        const float rad = __fsqrt_rn(__fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12)));
        const float min_eig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), rad), (float)(2 * WW * WH));
        if (L.flags & KLT_OPTFLOW_LK_GET_MIN_EIGENVALS) err = min_eig;
        if (min_eig < L.min_eig_thr || D < 1.1920929e-7f) {
            if (level == 0) status = 0;
            continue;
        }
        D = __fdiv_rn(1.f, D);

        // ---- iterations ------------------------------------------------------------------------------------------
        // make sure the staged next-image region covers the window at (inx, iny)
        auto ensure_j = [&](int inx, int iny) {
            if ((unsigned)(inx - jx0) > 2u * kM || (unsigned)(iny - jy0) > 2u * kM) {
                __syncwarp();
                jx0 = inx - kM; jy0 = iny - kM; jax = jx0 & ~3;
                stage<C::JR, C::SJ>(jreg, lvJ, imgJ, jax, jy0, jx0 - jax, C::JW, lane);
                __syncwarp();
            }
        };
        const int4* __restrict__ pcl = reinterpret_cast<const int4*>(pc) + lane;
        const uint4* __restrict__ pgl = reinterpret_cast<const uint4*>(pg) + lane;
        const uint2* __restrict__ pml = reinterpret_cast<const uint2*>(pm) + lane;

        float pdx = 0.f, pdy = 0.f;
        for (int j = 0; j < L.max_count; ++j) {
            if (j == budget && budget > 0) {
                // Long point (99.4 % of the (point, level) pairs of a KITTI frame converge within 6 iterations): it would
                // bound the launch, so a team of the long-point kernel continues it from here.
                if (lane == 0) {
                    const int slot = atomicAdd(L.wl_ctrl, 1);
                    push_entry(L, slot, gid, level, j, nx, ny, pdx, pdy, iters);
                }
                return;
            }
            float fnx, fny;
            if (!floor_in_range(nx, ny, WW, WH, lwm1, lhm1, fnx, fny)) {
                if (level == 0) status = 0;
                break;
            }
            ++iters;
            const int inx = (int)fnx, iny = (int)fny;
            ensure_j(inx, iny);
            int v00, v01, v10, v11;
            q14_weights(__fsub_rn(nx, fnx), __fsub_rn(ny, fny), v00, v01, v10, v11);
            const uint32_t V0 = (uint32_t)(v00 & 0xffff) | ((uint32_t)v01 << 16);
            const uint32_t V1 = (uint32_t)(v10 & 0xffff) | ((uint32_t)v11 << 16);
            const int cb = (iny - jy0) * C::SJ + (inx - jax);
            const uint32_t* __restrict__ jbase = reinterpret_cast<const uint32_t*>(jreg) + (cb >> 2);
            const int sh = (cb & 3) * 8;
            // class sums: [0..3] SIMD lanes x & 3, t = scalar tail; bound in units of 16 (rounded up per pixel)
            int s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0}, bd[4] = {0, 0, 0, 0}, s1t = 0, s2t = 0, bdt = 0;
#pragma unroll
            for (int k = 0; k < C::UPT; ++k) {
                const uint32_t* __restrict__ r0 = jbase + jw[k];
                const uint32_t* __restrict__ r1 = r0 + C::SJ / 4;
                const uint32_t p0w = r0[0], p1w = r0[1], q0w = r1[0], q1w = r1[1];
                const int4 cc = pcl[32 * k];
                const uint4 gg = pgl[32 * k];
                const uint2 mm = pml[32 * k];
                const uint32_t a0 = __funnelshift_r(p0w, p1w, sh), b0 = __funnelshift_rc(p0w, p1w, sh + 8);
                const uint32_t a1 = __funnelshift_r(q0w, q1w, sh), b1 = __funnelshift_rc(q0w, q1w, sh + 8);
                int d[4];
                d[0] = dp2a_lo(V1, a1, dp2a_lo(V0, a0, cc.x)) >> 9;
                d[1] = dp2a_lo(V1, b1, dp2a_lo(V0, b0, cc.y)) >> 9;
                d[2] = dp2a_hi(V1, a1, dp2a_hi(V0, a0, cc.z)) >> 9;
                d[3] = dp2a_hi(V1, b1, dp2a_hi(V0, b0, cc.w)) >> 9;
                const uint32_t gwv[4] = {gg.x, gg.y, gg.z, gg.w};
                const int gmv[4] = {(int)(mm.x & 0xffffu), (int)(mm.x >> 16), (int)(mm.y & 0xffffu), (int)(mm.y >> 16)};
                int u1[4], u2[4], ub[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    u1[jj] = d[jj] * lo16(gwv[jj]);
                    u2[jj] = d[jj] * hi16(gwv[jj]);
                    ub[jj] = (abs(d[jj]) * gmv[jj] + 15) >> 4;
                }
                class_add<WW, WH>(k, lane, s1, s1t, u1);
                class_add<WW, WH>(k, lane, s2, s2t, u2);
                class_add<WW, WH>(k, lane, bd, bdt, ub);
            }
            // per-lane |sum| <= UPT * 8160 * 4080 < 2^31; bounds: <= UPT * 4 * 2^21, clamped so the warp total cannot wrap
            int c1[5], c2[5], cbd[5];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                c1[q] = __reduce_add_sync(kFull, s1[q]);
                c2[q] = __reduce_add_sync(kFull, s2[q]);
                cbd[q] = __reduce_add_sync(kFull, min(bd[q], 1 << 22));
            }
            c1[4] = __reduce_add_sync(kFull, s1t);
            c2[4] = __reduce_add_sync(kFull, s2t);
            cbd[4] = __reduce_add_sync(kFull, min(bdt, 1 << 22));
            float b1, b2;
            if (max(max(max(cbd[0], cbd[1]), max(cbd[2], cbd[3])), cbd[4]) <= (kExact >> 4)) {
                // every float32 partial sum OpenCV forms in these classes is an exact integer
                b1 = combine5((float)c1[0], (float)c1[1], (float)c1[2], (float)c1[3], (float)c1[4]);
                b2 = combine5((float)c2[0], (float)c2[1], (float)c2[2], (float)c2[3], (float)c2[4]);
            } else if (L.wl != nullptr) {
                // the float32 sums really round: the long-point kernel has the fast serial replay; it redoes this iteration
                if (lane == 0) {
                    const int slot = atomicAdd(L.wl_ctrl, 1);
                    push_entry(L, slot, gid, level, j, nx, ny, pdx, pdy, iters - 1);
                }
                return;
            } else {
                // Serial replay in OpenCV's order (A.5), mismatch recomputed on the fly.  Lanes 0..7: SIMD chain
                // (sum = lane / 4, class = lane & 3): per 8-pixel step the pair (x, x + 4) is summed in int32, converted,
                // added; lanes 8, 9: the scalar-tail chains of b1, b2.
                ++n_t2;
                const bool simd = lane < 8, live = lane < 10;
                const int sm = simd ? (lane >> 2) : (lane - 8), cl = lane & 3;
                float acc = 0.f;
                const int cbx = inx - jax, cby = iny - jy0;
                for (int y = 0; y < WH; ++y) {
                    const uint8_t* jr0 = jreg + (cby + y) * C::SJ + cbx;
                    int pair = 0;
#pragma unroll 1
                    for (int s = 0; s < cmax(2 * C::NS, C::TL); ++s) {
                        const bool on = live && (simd ? (s < 2 * C::NS) : (s < C::TL));
                        const int x = on ? (simd ? (8 * (s >> 1) + cl + 4 * (s & 1)) : (C::NV + s)) : 0;
                        const int pi = patch_index<WW, WH>(y, x);
                        const uint8_t* jp = jr0 + x;
                        const int dv = (jp[0] * v00 + jp[1] * v01 + jp[C::SJ] * v10 + jp[C::SJ + 1] * v11 + pc[pi]) >> 9;
                        const uint32_t g = pg[pi];
                        const int prod = dv * (sm == 0 ? lo16(g) : hi16(g));
                        if (on) {
                            if (simd) {
                                if (s & 1) acc = __fadd_rn(acc, (float)(pair + prod));
                                else pair = prod;
                            } else {
                                acc = __fadd_rn(acc, (float)prod);
                            }
                        }
                    }
                }
                b1 = combine5(__shfl_sync(kFull, acc, 0), __shfl_sync(kFull, acc, 1), __shfl_sync(kFull, acc, 2), __shfl_sync(kFull, acc, 3),
                              __shfl_sync(kFull, acc, 8));
                b2 = combine5(__shfl_sync(kFull, acc, 4), __shfl_sync(kFull, acc, 5), __shfl_sync(kFull, acc, 6), __shfl_sync(kFull, acc, 7),
                              __shfl_sync(kFull, acc, 9));
            }
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
            outp = make_float2(__fadd_rn(nx, hwx), __fadd_rn(ny, hwy));
            {   // termination tests of A.4 6f / 6g without double-precision instructions on the common path
                const float s2f = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                bool small = s2f <= L.eps2_lo;
                if (!small && !(s2f >= L.eps2_hi))
                    small = __dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= L.eps2;
                if (small) break;
            }
            // (double)f < 0.01  <=>  f <= 0.01f: the float nearest to 0.01 lies below it, the next float above it
            if (j > 0 && fabsf(__fadd_rn(dx, pdx)) <= 0.01f && fabsf(__fadd_rn(dy, pdy)) <= 0.01f) {
                outp.x = __fsub_rn(outp.x, __fmul_rn(dx, 0.5f));
                outp.y = __fsub_rn(outp.y, __fmul_rn(dy, 0.5f));
                break;
            }
            pdx = dx; pdy = dy;
        }

        // ---- err at level 0 ------------------------------------------------------------------------------------------
        if (status && level == 0 && (L.flags & KLT_OPTFLOW_LK_GET_MIN_EIGENVALS) == 0) {
            const float qx = __fsub_rn(outp.x, hwx), qy = __fsub_rn(outp.y, hwy);
            float fqx, fqy;
            if (!floor_in_range(qx, qy, WW, WH, lwm1, lhm1, fqx, fqy)) {
                status = 0;
                continue;
            }
            const int iqx = (int)fqx, iqy = (int)fqy;
            ensure_j(iqx, iqy);
            int v00, v01, v10, v11;
            q14_weights(__fsub_rn(qx, fqx), __fsub_rn(qy, fqy), v00, v01, v10, v11);
            const uint32_t V0 = (uint32_t)(v00 & 0xffff) | ((uint32_t)v01 << 16);
            const uint32_t V1 = (uint32_t)(v10 & 0xffff) | ((uint32_t)v11 << 16);
            const int cb = (iqy - jy0) * C::SJ + (iqx - jax);
            const uint32_t* __restrict__ jbase = reinterpret_cast<const uint32_t*>(jreg) + (cb >> 2);
            const int sh = (cb & 3) * 8;
            int e = 0;
#pragma unroll
            for (int k = 0; k < C::UPT; ++k) {
                int y, x0, nvl;
                unit_geom<WW, WH>(k, lane, y, x0, nvl);
                const uint32_t* __restrict__ r0 = jbase + jw[k];
                const uint32_t* __restrict__ r1 = r0 + C::SJ / 4;
                const uint32_t p0w = r0[0], p1w = r0[1], q0w = r1[0], q1w = r1[1];
                const int4 cc = pcl[32 * k];
                const uint32_t a0 = __funnelshift_r(p0w, p1w, sh), b0 = __funnelshift_rc(p0w, p1w, sh + 8);
                const uint32_t a1 = __funnelshift_r(q0w, q1w, sh), b1 = __funnelshift_rc(q0w, q1w, sh + 8);
                int d[4];
                d[0] = dp2a_lo(V1, a1, dp2a_lo(V0, a0, cc.x)) >> 9;
                d[1] = dp2a_lo(V1, b1, dp2a_lo(V0, b0, cc.y)) >> 9;
                d[2] = dp2a_hi(V1, a1, dp2a_hi(V0, a0, cc.z)) >> 9;
                d[3] = dp2a_hi(V1, b1, dp2a_hi(V0, b0, cc.w)) >> 9;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) e += (jj < nvl) ? abs(d[jj]) : 0;
            }
            e = __reduce_add_sync(kFull, e);
            // |d| <= 8160 and WW*WH <= 2056 for the instantiated windows: e <= 2^24, so OpenCV's float32 running sum is exact
            err = __fdiv_rn(__fmul_rn((float)e, 1.f), (float)(32 * WW * WH));
        }
    }

    if (lane == 0) {
        reinterpret_cast<float2*>(L.next_pts)[gid] = outp;
        L.status[gid] = (uint8_t)status;
        L.err[gid] = err;
        if (L.iters) {
            // debug flag 0x100: cycles / 64 in the low 20 bits, replay count above (profiling aid)
            L.iters[gid] = (L.flags & 0x100) ? (int)(((clock64() - t_start) >> 6) & 0xfffff) | (min(n_g, 63) << 20) | (min(n_t2, 63) << 26) : iters;
        }
    }
}

template <int WW, int WH>
__global__ void __launch_bounds__(kWarpsPerCta * 32, (WW * WH <= 21 * 21) ? 6 : 3)
lk_warp_kernel(const __grid_constant__ LKLaunch L)
{
    using C = WC<WW, WH>;
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int wic = threadIdx.x >> 5;
    const long long gid = (long long)blockIdx.x * kWarpsPerCta + wic;
    if (gid >= (long long)L.n_per_pair * L.batch) return;  // whole warp
    warp_point<WW, WH>(L, gid, smem + wic * C::WARP_BYTES, lane);
    if (L.wl != nullptr && lane == 0) {
        __threadfence();            // the work-list entry (if any) is visible before the sign-off
        atomicAdd(L.wl_ctrl + kCtrlFinished, 1);
    }
}

template <int WW, int WH>
klt_status launch_warp(const LKLaunch& L, cudaStream_t stream)
{
    using C = WC<WW, WH>;
    static PerDeviceOnce configured;
    const size_t smem = (size_t)C::WARP_BYTES * kWarpsPerCta;
    if (configured.needed()) {
        cudaError_t e = cudaFuncSetAttribute(lk_warp_kernel<WW, WH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (klt_status)e;
    }
    if (L.n_per_pair < 0) return KLT_OK;   // configure only
    const long long total = (long long)L.n_per_pair * L.batch;
    const long long blocks = (total + kWarpsPerCta - 1) / kWarpsPerCta;
    if (blocks > 0x7fffffffLL) return KLT_ERR_UNSUPPORTED;
    lk_warp_kernel<WW, WH><<<(unsigned)blocks, kWarpsPerCta * 32, smem, stream>>>(L);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? KLT_OK : (klt_status)e;
}

}  // namespace

// Bulk launch of the warp-per-point shape; KLT_ERR_UNSUPPORTED when no specialisation exists.
klt_status lk_launch_warp(const LKLaunch& L, cudaStream_t stream)
{
    if (L.win_w == 21 && L.win_h == 21) return launch_warp<21, 21>(L, stream);
    if (L.win_w == 31 && L.win_h == 31) return launch_warp<31, 31>(L, stream);
    return KLT_ERR_UNSUPPORTED;
}

}  // namespace klt
