// K4 (specialised): fused Scharr + pyramidal LK for compile-time window sizes.
//
// Same arithmetic as klt_lk.cu (bit-exact with cv2.calcOpticalFlowPyrLK as called
// at reference src/extractor/extractor.py:44,45,65,66); this variant is what the BASELINE configs
// (winSize 21 and 31) run.  Structure:
//  * a CTA of 4 warps tracks S in {1,2,4} keypoints ("slots") through all pyramid levels in one launch;
//  * LEADERS: everything that is scalar per point (window position, range tests, Q14 weights, the 2x2 solve, the
//    termination tests, the exactness tests of the float32 sums) is executed by ONE warp: warp s is the leader of
//    slot s.  A leader is a small state machine that consumes the partial sums of its last COMMAND and posts the next
//    one (level setup, iteration, class sums, serial replay, err pass) in an 8-word mailbox in shared memory;
//  * per-pixel work is done by all 128 threads for one slot after the other (4-pixel units, one or two per thread
//    and slot, patch in registers), so the CTA alternates between a scalar phase (S leaders in parallel, each on its
//    own point) and a per-pixel phase (all warps, slot by slot), two block barriers per round.
//    Round 1 ran one point per team of warps and every warp of the team executed the scalar part redundantly: 9.3
//    thread-instructions per algorithmic MAC, 75 % of them outside the per-pixel loops (profiles/r01);
//  * bilinear taps use dp2a (two 14-bit weights x two u8 pixels per instruction, exact);
//  * neighbourhoods are staged with 32-bit loads (all loads in flight before the first store);
//  * the mismatch sums are reduced in three tiers: (0) if sum|d*gx| and sum|d*gy| over the WHOLE window are
//    <= 2^24 every float32 partial sum OpenCV forms is an exact integer, so b = float(sum) and only 4 values cross
//    the warps; (1) otherwise per-accumulation-class sums (4 SIMD lanes + tail) with a bound per class; (2) otherwise
//    a serial float32 replay in OpenCV's order.
#include "klt_common.cuh"

#include <cstdlib>

namespace klt {

namespace {

constexpr int kThreads = 128;
constexpr int kWarps = 4;
constexpr int kM = 3;  // margin of the staged next-image region
constexpr unsigned kFull = 0xffffffffu;
constexpr int kExact = 1 << 24;

__host__ __device__ constexpr int r4(int v) { return (v + 3) / 4 * 4; }
__host__ __device__ constexpr int r16(int v) { return (v + 15) / 16 * 16; }

// ---- serial float32 replay in OpenCV's accumulation order -------------------------------
// Scratch layout: one zero-padded row per accumulation chain (4 SIMD lanes with x % 4 == l, x < NV, then the scalar
// tail x >= NV), elements in OpenCV's visiting order.  Padding with zeros makes every chain the same length, so the
// replay loop has no bounds checks: adding +0.0f is an exact no-op (the accumulator is never -0).
template <int WW, int WH>
struct Chains {
    static constexpr int NV = 8 * (WW / 8), TL = WW - NV;
    // G scratch: 12 SIMD chains (3 sums x 4 lanes, one float product per pixel) of GQ4 floats, then 3 tail chains of
    // GT4 floats (zero padded to multiples of 4), then 16 result words
    static constexpr int GQ = WH * (NV / 4), GT = WH * TL;
    static constexpr int GQ4 = (GQ + 3) / 4 * 4, GT4 = (GT + 3) / 4 * 4;
    static constexpr int G_RES = 12 * GQ4 + 3 * GT4;
    static constexpr int G_WORDS = G_RES + 16;
    // b scratch: 8 SIMD chains (2 sums x 4 lanes) of int pairs (x, x+4) in visiting order, then 2 tail chains of floats
    static constexpr int NS = NV / 8;                                // 8-pixel SIMD steps per window row
    static constexpr int SUSED = 2 * WH * NS;                        // ints per SIMD chain
    static constexpr int SLEN = (SUSED + 3) / 4 * 4;                 // chain stride (zero padded: 0 + 0 -> +0.0f is a no-op)
    static constexpr int TLEN = (WH * TL + 3) / 4 * 4;               // floats per tail chain (zero padded)
    static constexpr int B_RES = 8 * SLEN + 2 * TLEN;                // 12 result words: q[0..3] of b1, of b2, t of b1, of b2
    static constexpr int B_WORDS = (B_RES + 12 + 3) / 4 * 4;
};

template <int WW, int WH, int WPP>
struct Cfg {
    static constexpr int UPR = (WW + 3) / 4;          // 4-pixel units per window row
    static constexpr int NV = 8 * (WW / 8);           // width of OpenCV's SIMD part (A.5)
    static constexpr int TL = WW - NV;                // scalar tail
    static constexpr int NU = WH * UPR;
    static constexpr int NT = 32 * WPP;               // threads per point
    static constexpr int UPT = (NU + NT - 1) / NT;    // units per thread
    static constexpr int PPC = kThreads / NT;         // points per CTA
    static constexpr bool PACK = UPT > 2;             // register budget: pack the per-pixel state
    static constexpr int SI = r4(WW + 6);             // prev-image region: row stride (bytes)
    static constexpr int IR = WH + 3;                 //                    rows
    static constexpr int SD = r4(WW + 2);             // derivative region: row stride (words)
    static constexpr int DR = WH + 1;
    static constexpr int JW = WW + 1 + 2 * kM;        // next-image region: logical size
    static constexpr int JR = WH + 1 + 2 * kM;
    static constexpr int SJ = r4(JW + 3);             //                    row stride (bytes)
    static constexpr int RPR = (WW + 1 + 3) / 4;      // Scharr: 4-position runs per row
    static constexpr int NRUN = DR * RPR;
    // shared-memory slice of one point (bytes)
    static constexpr int OFF_J = 0;
    static constexpr int OFF_D = OFF_J + r16(SJ * JR);        // dreg; fallback: packed derivative patch
    static constexpr int CHAIN_WORDS = r4(Chains<WW, WH>::B_WORDS > Chains<WW, WH>::G_WORDS ? Chains<WW, WH>::B_WORDS : Chains<WW, WH>::G_WORDS);   // replay scratch
    static constexpr int D_BYTES = r16(4 * (SD * DR > CHAIN_WORDS ? SD * DR : CHAIN_WORDS));  // dreg, or the replay chains
    static constexpr int OFF_I = OFF_D + D_BYTES;             // ireg
    static constexpr int I_BYTES = r16(SI * IR);
    static constexpr int NS = NV / 8;                         // 8-pixel SIMD steps per window row
    static constexpr int OFF_R3 = OFF_I + I_BYTES;            // [4 warps] int4: partial sums of a command
    static constexpr int OFF_R16 = OFF_R3 + 4 * 16;           // [16 values][4 warps] int: class sums of a command
    static constexpr int OFF_CMD = OFF_R16 + 256;             // mailbox: 2 x int4
    static constexpr int POINT_BYTES = (OFF_CMD + 32 + 127) / 128 * 128;
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
__device__ __forceinline__ unsigned long long gtimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ int lo16(uint32_t w) { return (int)(short)(w & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t w) { return ((int)w) >> 16; }

// Per-pixel patch state of one window pixel: Q5 intensity, Q14 derivative (gx, gy).
// Plain registers when a thread owns few pixels, two registers per pixel otherwise.
template <bool PACK> struct PxStore;
template <> struct PxStore<false> {
    int iv_, gx_, gy_;
    __device__ __forceinline__ void set(int iv, int gx, int gy) { iv_ = iv; gx_ = gx; gy_ = gy; }
    __device__ __forceinline__ int iv() const { return iv_; }
    __device__ __forceinline__ int gx() const { return gx_; }
    __device__ __forceinline__ int gy() const { return gy_; }
};
template <> struct PxStore<true> {
    int a_; uint32_t g_;
    __device__ __forceinline__ void set(int iv, int gx, int gy) { a_ = iv; g_ = ((uint32_t)gx & 0xffffu) | ((uint32_t)gy << 16); }
    __device__ __forceinline__ int iv() const { return a_; }
    __device__ __forceinline__ int gx() const { return (int)(short)(g_ & 0xffffu); }
    __device__ __forceinline__ int gy() const { return ((int)g_) >> 16; }
};
// the 4 mismatch values of one unit
template <bool PACK> struct DiffStore;
template <> struct DiffStore<false> {
    int d_[4];
    __device__ __forceinline__ void set(int j, int d) { d_[j] = d; }
    __device__ __forceinline__ int get(int j) const { return d_[j]; }
};
template <> struct DiffStore<true> {
    uint32_t p_[2];
    __device__ __forceinline__ void set(int j, int d) { if (j & 1) p_[j >> 1] |= (uint32_t)d << 16; else p_[j >> 1] = (uint32_t)d & 0xffffu; }
    __device__ __forceinline__ int get(int j) const { return (j & 1) ? (((int)p_[j >> 1]) >> 16) : (int)(short)(p_[j >> 1] & 0xffffu); }
};

__device__ __forceinline__ void q14_weights(float a, float b, int& w00, int& w01, int& w10, int& w11)
{
    const float oa = __fsub_rn(1.f, a), ob = __fsub_rn(1.f, b);
    w00 = __float2int_rn(__fmul_rn(__fmul_rn(oa, ob), 16384.f));
    w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, ob), 16384.f));
    w10 = __float2int_rn(__fmul_rn(__fmul_rn(oa, b), 16384.f));
    w11 = 16384 - w00 - w01 - w10;
}

__device__ __forceinline__ bool floor_in_range(float x, float y, int win_w, int win_h, int lw, int lh, int& ix, int& iy)
{
    const bool finite = (fabsf(x) < 1.0e9f) && (fabsf(y) < 1.0e9f);
    ix = finite ? __float2int_rd(x) : INT_MIN;
    iy = finite ? __float2int_rd(y) : INT_MIN;
    return finite && !(ix < -win_w || ix >= lw || iy < -win_h || iy >= lh);
}

__device__ __forceinline__ float combine5(float q0, float q1, float q2, float q3, float t)
{
    const float s = __fadd_rn(__fadd_rn(q0, q2), __fadd_rn(q1, q3));
    return __fmul_rn(__fadd_rn(t, s), 9.5367431640625e-07f);
}

// Stage ROWS x STRIDE bytes whose top-left image coordinate is (ax, y0) (ax % 4 == 0) into smem.  Columns
// [c0, c0 + need) of every row are the ones later read; the rest may hold anything.
template <int ROWS, int STRIDE, int NT>
__device__ __forceinline__ void stage(uint8_t* __restrict__ dst, const LevelView& lv, const uint8_t* __restrict__ img,
                                      int ax, int y0, int c0, int need, int tid)
{
    constexpr int NWR = STRIDE / 4;
    constexpr int NWORDS = ROWS * NWR;
    constexpr int PER = (NWORDS + NT - 1) / NT;
    // 32-bit loads whenever the staged columns lie inside the image; rows are reflected per row (REFLECT_101), which
    // is the common border case on the small coarse levels.  Only horizontal crossings take the byte path.
    const bool fast = lv.aligned4 && ax >= 0 && (ax + STRIDE <= lv.w);
    if (fast) {  // uniform over the point's threads
        const uint8_t* __restrict__ base = img + ax;
        const bool rows_inside = (y0 >= 0) && (y0 + ROWS <= lv.h);
        uint32_t v[PER];
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int i = tid + k * NT;
            const int r = i / NWR, c = i - r * NWR;
            const int yy = rows_inside ? (y0 + r) : reflect101(y0 + r, lv.h);
            v[k] = (i < NWORDS) ? __ldg(reinterpret_cast<const uint32_t*>(base + (long long)yy * lv.pitch) + c) : 0u;
        }
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int i = tid + k * NT;
            if (i < NWORDS) reinterpret_cast<uint32_t*>(dst)[i] = v[k];
        }
    } else {
        for (int i = tid; i < ROWS * need; i += NT) {
            const int r = i / need, c = c0 + (i - r * need);
            dst[r * STRIDE + c] = __ldg(img + (long long)reflect101(y0 + r, lv.h) * lv.pitch + reflect101(ax + c, lv.w));
        }
    }
}

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// The same staging without waiting for the data: 32-bit cp.async copies (LDGSTS: no registers, nobody waits); the
// consumer runs cp_async_wait_all() + a block barrier one round later.  Windows that cross the left / right image
// border (or unaligned images) take the synchronous byte path of stage().
template <int ROWS, int STRIDE, int NT>
__device__ __forceinline__ void stage_async(uint8_t* __restrict__ dst, const LevelView& lv, const uint8_t* __restrict__ img,
                                            int ax, int y0, int c0, int need, int tid)
{
    constexpr int NWR = STRIDE / 4;
    constexpr int NWORDS = ROWS * NWR;
    constexpr int PER = (NWORDS + NT - 1) / NT;
    const bool fast = lv.aligned4 && ax >= 0 && (ax + STRIDE <= lv.w);
    if (fast) {  // uniform over the CTA
        const uint8_t* __restrict__ base = img + ax;
        const bool rows_inside = (y0 >= 0) && (y0 + ROWS <= lv.h);
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int i = tid + k * NT;
            const int r = i / NWR, c = i - r * NWR;
            const int yy = rows_inside ? (y0 + r) : reflect101(y0 + r, lv.h);
            if (i < NWORDS) cp_async4(reinterpret_cast<uint32_t*>(dst) + i, reinterpret_cast<const uint32_t*>(base + (long long)yy * lv.pitch) + c);
        }
    } else {
        for (int i = tid; i < ROWS * need; i += NT) {
            const int r = i / need, c = c0 + (i - r * need);
            dst[r * STRIDE + c] = __ldg(img + (long long)reflect101(y0 + r, lv.h) * lv.pitch + reflect101(ax + c, lv.w));
        }
    }
}

// two words holding bytes [o, o+4) and [o+1, o+5) of an smem row (o = arbitrary byte offset)
__device__ __forceinline__ void load5(const uint8_t* row, int o, uint32_t& a, uint32_t& b)
{
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(row) + (o >> 2);
    const uint32_t w0 = wp[0], w1 = wp[1];
    const int s = (o & 3) * 8;
    a = __funnelshift_r(w0, w1, s);
    b = __funnelshift_rc(w0, w1, s + 8);
}

// serial float32 sum of one zero-padded chain of n4 floats (n4 % 4 == 0), in order
__device__ __forceinline__ float chain_sum(const float* __restrict__ p, int n4)
{
    const float4* __restrict__ src = reinterpret_cast<const float4*>(p);
    float acc = 0.f;
#pragma unroll 4
    for (int e = 0; e < n4 / 4; ++e) {
        const float4 v = src[e];
        acc = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc, v.x), v.y), v.z), v.w);
    }
    return acc;
}

// G replay, split over the point's warps: warp 0 runs the 12 SIMD chains (lane = chain), warp TW the 3 tail chains.
// The products were converted to float by the threads that own the pixels (|gx*gy| < 2^24: exact).
template <int WW, int WH, int WPP>
__device__ __forceinline__ void replay_g(float* __restrict__ gf, int wip, int lane)
{
    using CH = Chains<WW, WH>;
    constexpr int TW = WPP > 1 ? 1 : 0;
    if (wip == 0) {
        const float acc = chain_sum(gf + (lane < 12 ? lane : 0) * CH::GQ4, CH::GQ4);
        if (lane < 12) gf[CH::G_RES + (lane >> 2) * 5 + (lane & 3)] = acc;   // result layout: [sum][q0 q1 q2 q3 t]
    }
    if (wip == TW) {
        const float acc = chain_sum(gf + 12 * CH::GQ4 + (lane < 3 ? lane : 0) * CH::GT4, CH::GT4);
        if (lane < 3) gf[CH::G_RES + lane * 5 + 4] = acc;
    }
}

// zero the padding of the two tail chains of one b scratch buffer
template <int WW, int WH, int NT>
__device__ __forceinline__ void zero_pad_b(int* buf, int tid)
{
    using CH = Chains<WW, WH>;
    constexpr int PAD = CH::TLEN - WH * CH::TL;
    if (tid < 2 * PAD) buf[8 * CH::SLEN + (tid / (PAD > 0 ? PAD : 1)) * CH::TLEN + WH * CH::TL + tid % (PAD > 0 ? PAD : 1)] = 0;
    constexpr int SPAD = CH::SLEN - CH::SUSED;
    static_assert(8 * SPAD <= 32 && 2 * PAD <= 32, "one thread per padding word");
    if (tid < 8 * SPAD) buf[(tid / (SPAD > 0 ? SPAD : 1)) * CH::SLEN + CH::SUSED + tid % (SPAD > 0 ? SPAD : 1)] = 0;
}

// b replay, split over the point's warps: warp 0 runs the 8 SIMD chains (lane = chain; int pair -> float -> add), warp
// TW the 2 tail chains (floats, converted by the threads that own the pixels).  Results go to buf[B_RES ..].
template <int WW, int WH, int WPP>
__device__ __forceinline__ void replay_b(int* __restrict__ buf, int wip, int lane)
{
    using CH = Chains<WW, WH>;
    constexpr int TW = WPP > 1 ? 1 : 0;
    if (wip == 0) {
        const int2* __restrict__ src = reinterpret_cast<const int2*>(buf + (lane & 7) * CH::SLEN);
        float acc = 0.f;
#pragma unroll 6
        for (int e = 0; e < CH::SLEN / 2; ++e) {
            const int2 v = src[e];
            acc = __fadd_rn(acc, (float)(v.x + v.y));
        }
        if (lane < 8) reinterpret_cast<float*>(buf)[CH::B_RES + lane] = acc;
    }
    if (wip == TW) {
        const float4* __restrict__ src = reinterpret_cast<const float4*>(buf + 8 * CH::SLEN + (lane & 1) * CH::TLEN);
        float acc = 0.f;
#pragma unroll 4
        for (int e = 0; e < CH::TLEN / 4; ++e) {
            const float4 v = src[e];
            acc = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc, v.x), v.y), v.z), v.w);
        }
        if (lane < 2) reinterpret_cast<float*>(buf)[CH::B_RES + 8 + lane] = acc;
    }
}


// ---- commands (leader -> CTA) ---------------------------------------------------------------------------------------
// word 0 of the mailbox: op in bits 0..7, flags in bits 8..15, pyramid level in bits 16..23
enum : int { OP_NONE = 0, OP_STAGE = 1, OP_LEVEL = 2, OP_GREPLAY = 3, OP_ITER = 4, OP_TIER1 = 5, OP_REPLAY = 6, OP_ERR = 7 };
enum : int { F_RESTAGE = 0x100, F_CLASSES = 0x200, F_JVALID = 0x400 };
// leader state machine: what the leader does next (it consumes the partial sums of the command it issued last)
enum : int { PH_LEVEL_START, PH_STAGED, PH_AFTER_LEVEL, PH_AFTER_GREPLAY, PH_HAVE_G, PH_ITER_NEXT, PH_AFTER_SUM3, PH_AFTER_TIER1,
             PH_AFTER_REPLAY, PH_SOLVE, PH_LEVEL_END, PH_AFTER_ERR };

// OP_STAGE: both neighbourhoods of the slot's new level, asynchronously (consumed by OP_LEVEL in the next round)
template <int WW, int WH>
__device__ __forceinline__ void exec_stage(const LKLaunch& L, const int c_op, const int c_a, const int c_b, const int c_jx0, const int c_jy0,
                                           const int bidx, uint8_t* ws, const int tid)
{
    using C = Cfg<WW, WH, kWarps>;
    uint8_t* jreg = ws + C::OFF_J;
    uint8_t* ireg = ws + C::OFF_I;
    const int cur_level = (c_op >> 16) & 0xff;
    const LevelView lvI = L.prev.lv[cur_level];
    const LevelView lvJ = L.next.lv[cur_level];
    const uint8_t* __restrict__ imgI = lvI.data + (long long)bidx * lvI.batch_stride;
    const uint8_t* __restrict__ imgJ = lvJ.data + (long long)bidx * lvJ.batch_stride;
    const int ipx = c_a, ipy = c_b;
    if (c_op & F_JVALID) {
        const int sax = c_jx0 & ~3;
        stage_async<C::JR, C::SJ, C::NT>(jreg, lvJ, imgJ, sax, c_jy0, c_jx0 - sax, C::JW, tid);
    }
    const int iax = (ipx - 1) & ~3;
    stage_async<C::IR, C::SI, C::NT>(ireg, lvI, imgI, iax, ipy - 1, (ipx - 1) - iax, WW + 3, tid);
}

// ---- the per-pixel part of one command, executed by all threads of the CTA for the slot whose state is passed in ------
template <int WW, int WH>
__device__ __forceinline__ void exec_cmd(const LKLaunch& L, const int c_op, const int c_a, const int c_b, const int c_jx0, const int c_jy0,
                                         const uint32_t c_W0, const uint32_t c_W1, const int bidx, uint8_t* ws, const int tid,
                                         PxStore<Cfg<WW, WH, kWarps>::PACK> (&pxs)[Cfg<WW, WH, kWarps>::UPT][4],
                                         DiffStore<Cfg<WW, WH, kWarps>::PACK> (&dd)[Cfg<WW, WH, kWarps>::UPT], bool& pads_zeroed)
{
    using C = Cfg<WW, WH, kWarps>;
    using CH = Chains<WW, WH>;
    const int lane = tid & 31;
    const int wip = tid >> 5;
    uint8_t* jreg = ws + C::OFF_J;
    uint32_t* dreg = reinterpret_cast<uint32_t*>(ws + C::OFF_D);
    uint8_t* ireg = ws + C::OFF_I;
    int4* red3 = reinterpret_cast<int4*>(ws + C::OFF_R3);
    int* red16 = reinterpret_cast<int*>(ws + C::OFF_R16);
    // units of this thread: unit u = tid + k*NT covers window pixels (y, x0..x0+3)
    auto unit_y = [&](int k) { const int u = tid + k * C::NT; return (u < C::NU ? u : 0) / C::UPR; };
    auto unit_x0 = [&](int k) { const int u = tid + k * C::NT; const int uu = (u < C::NU ? u : 0); return 4 * (uu - (uu / C::UPR) * C::UPR); };
    auto unit_ok = [&](int k) { return tid + k * C::NT < C::NU; };
    int jw[C::UPT];     // word offset of the unit inside the staged next-image region
#pragma unroll
    for (int k = 0; k < C::UPT; ++k) jw[k] = (unit_y(k) * C::SJ + unit_x0(k)) >> 2;
    const int op = c_op & 0xff;
    if (op == OP_LEVEL) {
        const int cur_level = (c_op >> 16) & 0xff;
        pads_zeroed = false;
        const int tlw = L.prev.lv[cur_level].w, tlh = L.prev.lv[cur_level].h;
        const int ipx = c_a, ipy = c_b;
        const uint32_t W0 = c_W0, W1 = c_W1;
        const int oi = (ipx - 1) - ((ipx - 1) & ~3);
        cp_async_wait_all();
        __syncthreads();

        // ---- patch pass: Q5 intensity + Q14 derivative patch into registers, integer class sums of G ------------
        int vals[16];
        unsigned q11[4] = {0, 0, 0, 0}, q22[4] = {0, 0, 0, 0}, t11 = 0, t22 = 0;
        int q12[4] = {0, 0, 0, 0}, t12 = 0;
        // all (WW+1) x (WH+1) derivative positions inside the image <=> no zero-masking of the derivative
        const bool interior = (ipx >= 0) && (ipy >= 0) && (ipx + WW < tlw) && (ipy + WH < tlh);
        if (interior) {
            // Scharr is linear and so is the Q14 bilinear tap, so  sum_c w_c * Scharr(I)(p + c)  ==  Scharr(T)(p)  with
            // T(q) = sum_c w_c * I(q + c) the UNROUNDED bilinear sum (<= 255 * 2^14); exact in int32 (|.| < 2^27).
            const int sh = (oi & 3) * 8;
#pragma unroll
            for (int k = 0; k < C::UPT; ++k) {
                const int y = unit_y(k), x0 = unit_x0(k);
                const bool ok = unit_ok(k);
                uint32_t pa[4], pb[4], pc[4], pd[4];   // byte pairs (c,c+1) of 4 region rows
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const uint32_t* wp = reinterpret_cast<const uint32_t*>(ireg + (y + r) * C::SI) + ((oi + x0) >> 2);
                    const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
                    pa[r] = __funnelshift_r(w0, w1, sh);        // bytes c0 .. c0+3   (c0 = column of window x0-1)
                    pb[r] = __funnelshift_rc(w0, w1, sh + 8);   // bytes c0+1 .. c0+4
                    pc[r] = __funnelshift_r(w1, w2, sh);        // bytes c0+4 .. c0+7
                    pd[r] = __funnelshift_rc(w1, w2, sh + 8);   // bytes c0+5 .. c0+8
                }
                int T[3][6];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    T[r][0] = dp2a_lo(W1, pa[r + 1], dp2a_lo(W0, pa[r], 0));
                    T[r][1] = dp2a_lo(W1, pb[r + 1], dp2a_lo(W0, pb[r], 0));
                    T[r][2] = dp2a_hi(W1, pa[r + 1], dp2a_hi(W0, pa[r], 0));
                    T[r][3] = dp2a_hi(W1, pb[r + 1], dp2a_hi(W0, pb[r], 0));
                    T[r][4] = dp2a_lo(W1, pc[r + 1], dp2a_lo(W0, pc[r], 0));
                    T[r][5] = dp2a_lo(W1, pd[r + 1], dp2a_lo(W0, pd[r], 0));
                }
                int t0[6], t1[6];
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    t0[c] = 3 * (T[0][c] + T[2][c]) + 10 * T[1][c];
                    t1[c] = T[2][c] - T[0][c];
                }
                unsigned u11[4], u22[4];
                int u12[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const bool valid = ok && (x0 + jj) < WW;
                    const int iv = (T[1][jj + 1] + (1 << 8)) >> 9;
                    int gx = (t0[jj + 2] - t0[jj] + (1 << 13)) >> 14;
                    int gy = (3 * (t1[jj] + t1[jj + 2]) + 10 * t1[jj + 1] + (1 << 13)) >> 14;
                    gx = valid ? gx : 0; gy = valid ? gy : 0;
                    pxs[k][jj].set(valid ? iv : 0, gx, gy);
                    u11[jj] = (unsigned)(gx * gx); u12[jj] = gx * gy; u22[jj] = (unsigned)(gy * gy);
                }
                const bool tail = x0 >= C::NV;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    q11[jj] += tail ? 0u : u11[jj];
                    q12[jj] += tail ? 0 : u12[jj];
                    q22[jj] += tail ? 0u : u22[jj];
                }
                t11 += tail ? (u11[0] + u11[1] + u11[2] + u11[3]) : 0u;
                t12 += tail ? (u12[0] + u12[1] + u12[2] + u12[3]) : 0;
                t22 += tail ? (u22[0] + u22[1] + u22[2] + u22[3]) : 0u;
            }
        } else {
            // border window: Scharr at the (WW+1) x (WH+1) integer positions, zero outside the image, then bilinear
            const int w00 = lo16(W0), w01 = hi16(W0), w10 = lo16(W1), w11 = hi16(W1);
            for (int u = tid; u < C::NRUN; u += C::NT) {
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
                const bool yin = (unsigned)(ipy + dy) < (unsigned)tlh;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int dx = dx0 + k;
                    const int gx = t0[k + 2] - t0[k];
                    const int gy = 3 * (t1[k] + t1[k + 2]) + 10 * t1[k + 1];
                    const bool in = yin && ((unsigned)(ipx + dx) < (unsigned)tlw);
                    if (dx < C::SD) dreg[dy * C::SD + dx] = in ? (((uint32_t)gx & 0xffffu) | ((uint32_t)gy << 16)) : 0u;
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < C::UPT; ++k) {
                const int y = unit_y(k), x0 = unit_x0(k);
                const bool ok = unit_ok(k);
                uint32_t a0, b0, a1, b1_;
                load5(ireg + (y + 1) * C::SI, oi + 1 + x0, a0, b0);
                load5(ireg + (y + 2) * C::SI, oi + 1 + x0, a1, b1_);
                int iv[4];
                iv[0] = dp2a_lo(W1, a1, dp2a_lo(W0, a0, 256)) >> 9;
                iv[1] = dp2a_lo(W1, b1_, dp2a_lo(W0, b0, 256)) >> 9;
                iv[2] = dp2a_hi(W1, a1, dp2a_hi(W0, a0, 256)) >> 9;
                iv[3] = dp2a_hi(W1, b1_, dp2a_hi(W0, b0, 256)) >> 9;
                const uint32_t* d0 = dreg + y * C::SD + x0;
                const uint32_t* d1 = d0 + C::SD;
                const uint4 e0 = *reinterpret_cast<const uint4*>(d0);
                const uint4 e1 = *reinterpret_cast<const uint4*>(d1);
                const uint32_t r0w[5] = {e0.x, e0.y, e0.z, e0.w, d0[4]};
                const uint32_t r1w[5] = {e1.x, e1.y, e1.z, e1.w, d1[4]};
                unsigned u11[4], u22[4];
                int u12[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    int gx = (lo16(r0w[jj]) * w00 + lo16(r0w[jj + 1]) * w01 + lo16(r1w[jj]) * w10 + lo16(r1w[jj + 1]) * w11 + (1 << 13)) >> 14;
                    int gy = (hi16(r0w[jj]) * w00 + hi16(r0w[jj + 1]) * w01 + hi16(r1w[jj]) * w10 + hi16(r1w[jj + 1]) * w11 + (1 << 13)) >> 14;
                    const bool valid = ok && (x0 + jj) < WW;
                    gx = valid ? gx : 0; gy = valid ? gy : 0;
                    pxs[k][jj].set(valid ? iv[jj] : 0, gx, gy);
                    u11[jj] = (unsigned)(gx * gx); u12[jj] = gx * gy; u22[jj] = (unsigned)(gy * gy);
                }
                const bool tail = x0 >= C::NV;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    q11[jj] += tail ? 0u : u11[jj];
                    q12[jj] += tail ? 0 : u12[jj];
                    q22[jj] += tail ? 0u : u22[jj];
                }
                t11 += tail ? (u11[0] + u11[1] + u11[2] + u11[3]) : 0u;
                t12 += tail ? (u12[0] + u12[1] + u12[2] + u12[3]) : 0;
                t22 += tail ? (u22[0] + u22[1] + u22[2] + u22[3]) : 0u;
            }
            // the replay scratch aliases dreg: the Scharr words are dead once every thread has passed the exchange below
        }
        {
            const unsigned cap = (1u << 25) / kWarps;  // keeps the point-wide totals below 2^31
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                vals[jj] = (int)min(q11[jj], cap);
                vals[5 + jj] = max(min(q12[jj], (int)cap), -(int)cap);
                vals[10 + jj] = (int)min(q22[jj], cap);
            }
            vals[4] = (int)min(t11, cap);
            vals[9] = max(min(t12, (int)cap), -(int)cap);
            vals[14] = (int)min(t22, cap);
            vals[15] = 0;
        }
        int mine = 0;
#pragma unroll
        for (int i = 0; i < 15; ++i) {
            const int t = __reduce_add_sync(kFull, vals[i]);
            mine = (lane == i) ? t : mine;
        }
        if (lane < 15) red16[lane * 4 + wip] = mine;
    } else if (op == OP_ITER || op == OP_ERR) {
        if (c_op & F_RESTAGE) {
            // (everybody has finished reading the old region: the previous command is complete)
            const LevelView lvJ = L.next.lv[(c_op >> 16) & 0xff];
            const uint8_t* __restrict__ imgJ = lvJ.data + (long long)bidx * lvJ.batch_stride;
            const int sax = c_jx0 & ~3;
            stage<C::JR, C::SJ, C::NT>(jreg, lvJ, imgJ, sax, c_jy0, c_jx0 - sax, C::JW, tid);
            __syncthreads();
        }
        const uint32_t W0 = c_W0, W1 = c_W1;
        const int cb = c_a;
        const uint32_t* __restrict__ jbase = reinterpret_cast<const uint32_t*>(jreg) + (cb >> 2);
        const int sh = (cb & 3) * 8;
        if (op == OP_ERR) {
            int e = 0;
#pragma unroll
            for (int k = 0; k < C::UPT; ++k) {
                const uint32_t* __restrict__ r0 = jbase + jw[k];
                const uint32_t* __restrict__ r1 = r0 + C::SJ / 4;
                const uint32_t p0_ = r0[0], p1_ = r0[1], q0 = r1[0], q1 = r1[1];
                const uint32_t a0 = __funnelshift_r(p0_, p1_, sh), b0 = __funnelshift_rc(p0_, p1_, sh + 8);
                const uint32_t a1 = __funnelshift_r(q0, q1, sh), b1_ = __funnelshift_rc(q0, q1, sh + 8);
                int jv[4];
                jv[0] = dp2a_lo(W1, a1, dp2a_lo(W0, a0, 256)) >> 9;
                jv[1] = dp2a_lo(W1, b1_, dp2a_lo(W0, b0, 256)) >> 9;
                jv[2] = dp2a_hi(W1, a1, dp2a_hi(W0, a0, 256)) >> 9;
                jv[3] = dp2a_hi(W1, b1_, dp2a_hi(W0, b0, 256)) >> 9;
                const bool ok = unit_ok(k);
                const int x0 = unit_x0(k);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) e += (ok && (x0 + jj) < WW) ? abs(jv[jj] - pxs[k][jj].iv()) : 0;
            }
            e = __reduce_add_sync(kFull, e);
            if (lane == 0) red3[wip] = make_int4(e, 0, 0, 0);
        } else {
            // invalid pixels carry gx = gy = 0, so they drop out of all sums without a select
            int s1 = 0, s2 = 0;
            unsigned bx = 0, by = 0;
#pragma unroll
            for (int k = 0; k < C::UPT; ++k) {
                const uint32_t* __restrict__ r0 = jbase + jw[k];
                const uint32_t* __restrict__ r1 = r0 + C::SJ / 4;
                const uint32_t p0_ = r0[0], p1_ = r0[1], q0 = r1[0], q1 = r1[1];
                const uint32_t a0 = __funnelshift_r(p0_, p1_, sh), b0 = __funnelshift_rc(p0_, p1_, sh + 8);
                const uint32_t a1 = __funnelshift_r(q0, q1, sh), b1_ = __funnelshift_rc(q0, q1, sh + 8);
                int jv[4];
                jv[0] = dp2a_lo(W1, a1, dp2a_lo(W0, a0, 256)) >> 9;
                jv[1] = dp2a_lo(W1, b1_, dp2a_lo(W0, b0, 256)) >> 9;
                jv[2] = dp2a_hi(W1, a1, dp2a_hi(W0, a0, 256)) >> 9;
                jv[3] = dp2a_hi(W1, b1_, dp2a_hi(W0, b0, 256)) >> 9;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int d = jv[jj] - pxs[k][jj].iv();
                    dd[k].set(jj, d);
                    const int u1 = d * pxs[k][jj].gx(), u2 = d * pxs[k][jj].gy();
                    s1 += u1; s2 += u2;
                    bx += (unsigned)abs(u1); by += (unsigned)abs(u2);
                }
            }
            if (!(c_op & F_CLASSES)) {
                // per-thread bounds <= UPT*4*8160*4080 < 2^32 for UPT <= 8; clamp so the point totals cannot wrap
                bx = min(bx, (1u << 25) / kWarps); by = min(by, (1u << 25) / kWarps);
                s1 = __reduce_add_sync(kFull, s1);
                s2 = __reduce_add_sync(kFull, s2);
                bx = __reduce_add_sync(kFull, bx);
                by = __reduce_add_sync(kFull, by);
                if (lane == 0) red3[wip] = make_int4(s1, s2, (int)bx, (int)by);
            }
        }
    }
    if (op == OP_TIER1 || (op == OP_ITER && (c_op & F_CLASSES))) {
        // class sums (4 SIMD lanes + tail) of d*gx, d*gy and of the bound |d| * max(|gx|,|gy|) in units of 16
        int cv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) cv[i] = 0;
#pragma unroll
        for (int k = 0; k < C::UPT; ++k) {
            int u1[4], u2[4], ub[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int d = dd[k].get(jj);
                const int gx = pxs[k][jj].gx(), gy = pxs[k][jj].gy();
                u1[jj] = d * gx;
                u2[jj] = d * gy;
                ub[jj] = (abs(d) * max(abs(gx), abs(gy)) + 15) >> 4;
            }
            const bool tail = unit_x0(k) >= C::NV;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                cv[jj] += tail ? 0 : u1[jj];
                cv[5 + jj] += tail ? 0 : u2[jj];
                cv[10 + jj] += tail ? 0 : ub[jj];
            }
            cv[4] += tail ? (u1[0] + u1[1] + u1[2] + u1[3]) : 0;
            cv[9] += tail ? (u2[0] + u2[1] + u2[2] + u2[3]) : 0;
            cv[14] += tail ? (ub[0] + ub[1] + ub[2] + ub[3]) : 0;
        }
#pragma unroll
        for (int i = 10; i < 15; ++i) cv[i] = min(cv[i], (1 << 22) / kWarps);
        int mine = 0;
#pragma unroll
        for (int i = 0; i < 15; ++i) {
            const int t = __reduce_add_sync(kFull, cv[i]);
            mine = (lane == i) ? t : mine;
        }
        if (lane < 15) red16[lane * 4 + wip] = mine;
    } else if (op == OP_REPLAY) {
        // tier 2: serial replay in OpenCV's order (pairs (l, l+4) summed in int32 first; A.5)
        int* buf = reinterpret_cast<int*>(dreg);
        if (!pads_zeroed) { zero_pad_b<WW, WH, C::NT>(buf, tid); pads_zeroed = true; }  // pads survive until the next level
#pragma unroll
        for (int k = 0; k < C::UPT; ++k)
            if (unit_ok(k)) {
                const int y = unit_y(k), x0 = unit_x0(k);
                if (x0 >= C::NV) {
                    float* tf = reinterpret_cast<float*>(buf) + 8 * CH::SLEN + y * CH::TL + (x0 - C::NV);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
                        if (x0 + jj < WW) {
                            const int d = dd[k].get(jj);
                            tf[jj] = (float)(d * pxs[k][jj].gx());
                            tf[CH::TLEN + jj] = (float)(d * pxs[k][jj].gy());
                        }
                } else {
                    int* si = buf + (y * CH::NS + (x0 >> 3)) * 2 + ((x0 >> 2) & 1);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int d = dd[k].get(jj);
                        si[jj * CH::SLEN] = d * pxs[k][jj].gx();
                        si[(4 + jj) * CH::SLEN] = d * pxs[k][jj].gy();
                    }
                }
            }
        __syncthreads();
        replay_b<WW, WH, kWarps>(buf, wip, lane);
    } else if (op == OP_GREPLAY) {
        // serial replay of the G sums in OpenCV's order (A.5): every thread stores the float products of its pixels in
        // chain order; the scratch (dreg) is dead here
        float* gf = reinterpret_cast<float*>(dreg);
#pragma unroll
        for (int k = 0; k < C::UPT; ++k)
            if (unit_ok(k)) {
                const int y = unit_y(k), x0 = unit_x0(k);
                const bool tail = x0 >= C::NV;
                float* g0 = tail ? gf + 12 * CH::GQ4 + y * CH::TL + (x0 - C::NV) : gf + y * (C::NV / 4) + (x0 >> 2);
                const int sj = tail ? 1 : CH::GQ4;            // next pixel: next element of the tail / next lane chain
                const int ss = tail ? CH::GT4 : 4 * CH::GQ4;  // next sum
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                    if (x0 + jj < WW) {
                        const int gx = pxs[k][jj].gx(), gy = pxs[k][jj].gy();
                        g0[jj * sj] = (float)(gx * gx);
                        g0[jj * sj + ss] = (float)(gx * gy);
                        g0[jj * sj + 2 * ss] = (float)(gy * gy);
                    }
            }
        {   // zero pads
            constexpr int PQ = CH::GQ4 - CH::GQ, PT = CH::GT4 - CH::GT;
            if (tid < 12 * PQ) gf[(tid / (PQ > 0 ? PQ : 1)) * CH::GQ4 + CH::GQ + tid % (PQ > 0 ? PQ : 1)] = 0.f;
            if (tid < 3 * PT) gf[12 * CH::GQ4 + (tid / (PT > 0 ? PT : 1)) * CH::GT4 + CH::GT + tid % (PT > 0 ? PT : 1)] = 0.f;
        }
        pads_zeroed = false;   // the G chains overlap the pads of the b chains
        __syncthreads();
        replay_g<WW, WH, kWarps>(gf, wip, lane);
    }
}

template <int S> struct MinBlocks { static constexpr int v = (S == 4 ? 4 : 5); };

// One CTA = 4 warps = S keypoints (slots): point blockIdx.x * S + s is led by warp s.
template <int WW, int WH, int S>
__global__ void __launch_bounds__(kThreads, MinBlocks<S>::v)
lk_fast_kernel(const __grid_constant__ LKLaunch L)
{
    using C = Cfg<WW, WH, kWarps>;
    using CH = Chains<WW, WH>;
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int wip = tid >> 5;
    // (a vote makes the predicate warp-uniform for the compiler: no convergence barriers around the leader's collectives)
    const bool is_leader = (S == kWarps) ? true : __all_sync(kFull, wip < S);
    const long long total = (long long)L.n_per_pair * L.batch;
    const long long gid = (long long)blockIdx.x * S + wip;        // the point this warp leads (if it leads one)

    // ---- per-pixel state of the CTA: one patch per slot (registers; valid for the slot's current level) ----------------
    PxStore<C::PACK> pxs[S][C::UPT][4];
    DiffStore<C::PACK> dd[S][C::UPT];
    bool pads_zeroed[S];
#pragma unroll
    for (int s = 0; s < S; ++s) pads_zeroed[s] = false;

    // ---- state of this warp as the leader of slot `wip` ------------------------------------------------------------------
    uint8_t* ws = smem + (is_leader ? wip : 0) * C::POINT_BYTES;
    uint32_t* dreg = reinterpret_cast<uint32_t*>(ws + C::OFF_D);
    const int4* red3 = reinterpret_cast<const int4*>(ws + C::OFF_R3);
    const int* red16 = reinterpret_cast<const int*>(ws + C::OFF_R16);
    int4* mbox = reinterpret_cast<int4*>(ws + C::OFF_CMD);
    const long long t_start = clock64();
#ifdef KLT_LK_TIMELINE
    const unsigned long long t_g0 = (L.flags & 0x400) ? gtimer() : 0ull;   // debug flag 0x400: start / end stamps (128 ns units)
#endif
    int n_t1 = 0, n_t2 = 0;
    bool finished = !is_leader || gid >= total;
    float2 p0 = make_float2(0.f, 0.f), outp = make_float2(0.f, 0.f);
    int bidx = 0;
    if (!finished) {
        p0 = reinterpret_cast<const float2*>(L.prev_pts)[gid];
        if (L.flags & KLT_OPTFLOW_USE_INITIAL_FLOW) outp = reinterpret_cast<const float2*>(L.next_pts)[gid];
        bidx = (int)(gid / L.n_per_pair);
    } else if (is_leader && lane == 0) {
        mbox[0] = make_int4(OP_NONE, 0, 0, 0);
    }
    int status = 1;
    float err = 0.f;
    int iters = 0;
    const float hwx = (float)(WW - 1) * 0.5f, hwy = (float)(WH - 1) * 0.5f;
    int level = L.prev.top;
    int phase = PH_LEVEL_START;
    int lw = 0, lh = 0;
    float nx = 0.f, ny = 0.f, pdx = 0.f, pdy = 0.f;
    float A11 = 0.f, A12 = 0.f, A22 = 0.f, D = 0.f, b1 = 0.f, b2 = 0.f;
    int jx0 = 0, jy0 = 0, jax = 0, j = 0;   // staged next-image region: smem col 0 <-> image x = jax; window columns start at jx0
    bool jvalid = false, sticky = false;
    int lv_ipx = 0, lv_ipy = 0;            // parameters of the level being staged
    uint32_t lv_W0 = 0, lv_W1 = 0;

    for (;;) {
        if (!finished) {
            // ================= leader step: consume the partial sums of the last command, advance to the next command ===
            int c_op = 0, c_a = 0, c_b = 0, c_jx0 = 0, c_jy0 = 0;
            uint32_t c_W0 = 0, c_W1 = 0;
            bool produced = false;
#pragma unroll 1
            while (!produced) {
                switch (phase) {
                case PH_LEVEL_START: {
                    if (level < 0) { c_op = OP_NONE; finished = true; produced = true; break; }
                    lw = L.prev.lv[level].w; lh = L.prev.lv[level].h;
                    const float scale = __int_as_float((127 - level) << 23);
                    float px = __fmul_rn(p0.x, scale), py = __fmul_rn(p0.y, scale);
                    if (level == L.prev.top) {
                        if (L.flags & KLT_OPTFLOW_USE_INITIAL_FLOW) { nx = __fmul_rn(outp.x, scale); ny = __fmul_rn(outp.y, scale); }
                        else { nx = px; ny = py; }
                    } else {
                        nx = __fmul_rn(outp.x, 2.f); ny = __fmul_rn(outp.y, 2.f);
                    }
                    outp = make_float2(nx, ny);
                    px = __fsub_rn(px, hwx); py = __fsub_rn(py, hwy);
                    int ipx, ipy;
                    if (!floor_in_range(px, py, WW, WH, lw, lh, ipx, ipy)) {
                        if (level == 0) { status = 0; err = 0.f; }
                        --level;
                        break;
                    }
                    int w00, w01, w10, w11;
                    q14_weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy), w00, w01, w10, w11);
                    nx = __fsub_rn(nx, hwx); ny = __fsub_rn(ny, hwy);
                    int inx, iny;
                    jvalid = floor_in_range(nx, ny, WW, WH, lw, lh, inx, iny);
                    if (jvalid) { jx0 = inx - kM; jy0 = iny - kM; jax = jx0 & ~3; }
                    c_op = OP_STAGE | (jvalid ? F_JVALID : 0);
                    c_a = ipx; c_b = ipy;
                    lv_W0 = (uint32_t)(w00 & 0xffff) | ((uint32_t)w01 << 16);
                    lv_W1 = (uint32_t)(w10 & 0xffff) | ((uint32_t)w11 << 16);
                    lv_ipx = ipx; lv_ipy = ipy;
                    c_jx0 = jx0; c_jy0 = jy0;
                    phase = PH_STAGED; produced = true;
                    break;
                }
                case PH_STAGED: {
                    c_op = OP_LEVEL;
                    c_a = lv_ipx; c_b = lv_ipy; c_W0 = lv_W0; c_W1 = lv_W1;
                    phase = PH_AFTER_LEVEL; produced = true;
                    break;
                }
                case PH_AFTER_LEVEL: {
                    // lane L < 15 holds class total L: [0..4] = gx*gx (4 SIMD lanes, tail), [5..9] = gx*gy, [10..14] = gy*gy
                    const int4 t = reinterpret_cast<const int4*>(red16)[lane & 15];
                    const int gtot = t.x + t.y + t.z + t.w;
                    // A11 / A22: non-negative terms, exact iff every class total <= 2^24; A12: |gx gy| <= (gx^2 + gy^2) / 2
                    const unsigned partner = (unsigned)__shfl_sync(kFull, gtot, (lane + 10) & 31);
                    const bool ok = (lane >= 5) || ((unsigned)gtot <= (unsigned)kExact && partner <= (unsigned)kExact &&
                                                    (unsigned)gtot + partner <= 2u * (unsigned)kExact);
                    if (__all_sync(kFull, ok)) {
                        const float f = (float)gtot;
                        A11 = combine5(__shfl_sync(kFull, f, 0), __shfl_sync(kFull, f, 1), __shfl_sync(kFull, f, 2), __shfl_sync(kFull, f, 3),
                                       __shfl_sync(kFull, f, 4));
                        A12 = combine5(__shfl_sync(kFull, f, 5), __shfl_sync(kFull, f, 6), __shfl_sync(kFull, f, 7), __shfl_sync(kFull, f, 8),
                                       __shfl_sync(kFull, f, 9));
                        A22 = combine5(__shfl_sync(kFull, f, 10), __shfl_sync(kFull, f, 11), __shfl_sync(kFull, f, 12), __shfl_sync(kFull, f, 13),
                                       __shfl_sync(kFull, f, 14));
                        phase = PH_HAVE_G;
                    } else {
                        c_op = OP_GREPLAY; phase = PH_AFTER_GREPLAY; produced = true;
                    }
                    break;
                }
                case PH_AFTER_GREPLAY: {
                    const float* r = reinterpret_cast<const float*>(dreg) + CH::G_RES;
                    A11 = combine5(r[0], r[1], r[2], r[3], r[4]);
                    A12 = combine5(r[5], r[6], r[7], r[8], r[9]);
                    A22 = combine5(r[10], r[11], r[12], r[13], r[14]);
                    phase = PH_HAVE_G;
                    break;
                }
                case PH_HAVE_G: {
                    D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
                    const float dA = __fsub_rn(A11, A22);
                    const float rad = __fsqrt_rn(__fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12)));
                    const float min_eig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), rad), (float)(2 * WW * WH));
                    if (L.flags & KLT_OPTFLOW_LK_GET_MIN_EIGENVALS) err = min_eig;
                    if (min_eig < L.min_eig_thr || D < 1.1920929e-7f) {
                        if (level == 0) status = 0;
                        --level; phase = PH_LEVEL_START;
                        break;
                    }
                    D = __fdiv_rn(1.f, D);
                    j = 0; pdx = 0.f; pdy = 0.f; sticky = false;
                    phase = PH_ITER_NEXT;
                    break;
                }
                case PH_ITER_NEXT: {
                    if (j >= L.max_count) { phase = PH_LEVEL_END; break; }
                    int inx, iny;
                    if (!floor_in_range(nx, ny, WW, WH, lw, lh, inx, iny)) {
                        if (level == 0) status = 0;
                        phase = PH_LEVEL_END;
                        break;
                    }
                    ++iters;
                    // make sure the staged next-image region covers the window at (inx, iny)
                    const bool restage = !jvalid || (unsigned)(inx - jx0) > (unsigned)(2 * kM) || (unsigned)(iny - jy0) > (unsigned)(2 * kM);
                    if (restage) { jx0 = inx - kM; jy0 = iny - kM; jax = jx0 & ~3; jvalid = true; }
                    int v00, v01, v10, v11;
                    q14_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), v00, v01, v10, v11);
                    c_W0 = (uint32_t)(v00 & 0xffff) | ((uint32_t)v01 << 16);
                    c_W1 = (uint32_t)(v10 & 0xffff) | ((uint32_t)v11 << 16);
                    c_a = (iny - jy0) * C::SJ + (inx - jax);
                    c_jx0 = jx0; c_jy0 = jy0;
                    // Tier 0 (whole-window bounds) unless the previous iteration of this point already failed it ("sticky"):
                    // diverging points fail it every time, and they are the ones that bound the launch latency.
                    c_op = OP_ITER | (restage ? F_RESTAGE : 0) | (sticky ? F_CLASSES : 0);
                    if (sticky) ++n_t1;
                    phase = sticky ? PH_AFTER_TIER1 : PH_AFTER_SUM3;
                    produced = true;
                    break;
                }
                case PH_AFTER_SUM3: {
                    int s1 = 0, s2 = 0;
                    unsigned bx = 0, by = 0;
#pragma unroll
                    for (int w = 0; w < kWarps; ++w) {
                        const int4 v = red3[w];
                        s1 += v.x; s2 += v.y; bx += (unsigned)v.z; by += (unsigned)v.w;
                    }
                    if (bx <= (unsigned)kExact && by <= (unsigned)kExact) {
                        // every float32 partial sum OpenCV forms (lanes, tail, final combine) is an exact integer
                        b1 = __fmul_rn((float)s1, 9.5367431640625e-07f);
                        b2 = __fmul_rn((float)s2, 9.5367431640625e-07f);
                        phase = PH_SOLVE;
                    } else {
                        sticky = true; ++n_t1;
                        c_op = OP_TIER1; phase = PH_AFTER_TIER1; produced = true;
                    }
                    break;
                }
                case PH_AFTER_TIER1: {
                    // tier 1: per accumulation class (4 SIMD lanes + tail), bound in units of 16 (rounded up per pixel)
                    const int4 t = reinterpret_cast<const int4*>(red16)[lane & 15];
                    const int ctot = t.x + t.y + t.z + t.w;
                    const bool is_bound = (lane >= 10) && (lane < 15);
                    const bool exact = __all_sync(kFull, !is_bound || ctot <= (kExact >> 4));
                    // leave sticky mode once the whole-window bound would pass again (converging point)
                    sticky = __reduce_add_sync(kFull, is_bound ? ctot : 0) > (kExact >> 4);
                    if (exact) {
                        const float f = (float)ctot;
                        b1 = combine5(__shfl_sync(kFull, f, 0), __shfl_sync(kFull, f, 1), __shfl_sync(kFull, f, 2), __shfl_sync(kFull, f, 3),
                                      __shfl_sync(kFull, f, 4));
                        b2 = combine5(__shfl_sync(kFull, f, 5), __shfl_sync(kFull, f, 6), __shfl_sync(kFull, f, 7), __shfl_sync(kFull, f, 8),
                                      __shfl_sync(kFull, f, 9));
                        phase = PH_SOLVE;
                    } else {
                        ++n_t2;
                        c_op = OP_REPLAY; phase = PH_AFTER_REPLAY; produced = true;
                    }
                    break;
                }
                case PH_AFTER_REPLAY: {
                    const int* buf = reinterpret_cast<const int*>(dreg);
                    const float4 r1 = *reinterpret_cast<const float4*>(buf + CH::B_RES);
                    const float4 r2 = *reinterpret_cast<const float4*>(buf + CH::B_RES + 4);
                    const float2 rt = *reinterpret_cast<const float2*>(buf + CH::B_RES + 8);
                    b1 = combine5(r1.x, r1.y, r1.z, r1.w, rt.x);
                    b2 = combine5(r2.x, r2.y, r2.z, r2.w, rt.y);
                    phase = PH_SOLVE;
                    break;
                }
                case PH_SOLVE: {
                    const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
                    const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
                    nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
                    outp = make_float2(__fadd_rn(nx, hwx), __fadd_rn(ny, hwy));
                    phase = PH_ITER_NEXT;
                    {   // termination tests of A.4 6f / 6g without double-precision instructions on the common path
                        const float s2f = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                        bool small = s2f <= L.eps2_lo;
                        if (!small && !(s2f >= L.eps2_hi))
                            small = __dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= L.eps2;
                        if (small) { phase = PH_LEVEL_END; break; }
                    }
                    // (double)f < 0.01  <=>  f <= 0.01f: the float nearest to 0.01 lies below it, the next float above it
                    if (j > 0 && fabsf(__fadd_rn(dx, pdx)) <= 0.01f && fabsf(__fadd_rn(dy, pdy)) <= 0.01f) {
                        outp.x = __fsub_rn(outp.x, __fmul_rn(dx, 0.5f));
                        outp.y = __fsub_rn(outp.y, __fmul_rn(dy, 0.5f));
                        phase = PH_LEVEL_END;
                        break;
                    }
                    pdx = dx; pdy = dy; ++j;
                    break;
                }
                case PH_LEVEL_END: {
                    // ---- err at level 0 -------------------------------------------------------------------------------
                    if (status && level == 0 && (L.flags & KLT_OPTFLOW_LK_GET_MIN_EIGENVALS) == 0) {
                        const float qx = __fsub_rn(outp.x, hwx), qy = __fsub_rn(outp.y, hwy);
                        int iqx, iqy;
                        if (!floor_in_range(qx, qy, WW, WH, lw, lh, iqx, iqy)) {
                            status = 0;
                            --level; phase = PH_LEVEL_START;
                            break;
                        }
                        const bool restage = !jvalid || (unsigned)(iqx - jx0) > (unsigned)(2 * kM) || (unsigned)(iqy - jy0) > (unsigned)(2 * kM);
                        if (restage) { jx0 = iqx - kM; jy0 = iqy - kM; jax = jx0 & ~3; jvalid = true; }
                        int v00, v01, v10, v11;
                        q14_weights(__fsub_rn(qx, (float)iqx), __fsub_rn(qy, (float)iqy), v00, v01, v10, v11);
                        c_W0 = (uint32_t)(v00 & 0xffff) | ((uint32_t)v01 << 16);
                        c_W1 = (uint32_t)(v10 & 0xffff) | ((uint32_t)v11 << 16);
                        c_a = (iqy - jy0) * C::SJ + (iqx - jax);
                        c_jx0 = jx0; c_jy0 = jy0;
                        c_op = OP_ERR | (restage ? F_RESTAGE : 0);
                        phase = PH_AFTER_ERR; produced = true;
                    } else {
                        --level; phase = PH_LEVEL_START;
                    }
                    break;
                }
                default: {   // PH_AFTER_ERR
                    int e = 0;
#pragma unroll
                    for (int w = 0; w < kWarps; ++w) e += red3[w].x;
                    // |d| <= 8160 and WW*WH <= 2056 for the instantiated windows: e <= 2^24, so OpenCV's float32 running sum is exact
                    err = __fdiv_rn(__fmul_rn((float)e, 1.f), (float)(32 * WW * WH));
                    --level; phase = PH_LEVEL_START;
                    break;
                }
                }
            }
            if (finished) {
                if (lane == 0) {
                    reinterpret_cast<float2*>(L.next_pts)[gid] = outp;
                    L.status[gid] = (uint8_t)status;
                    L.err[gid] = err;
                    if (L.iters) {
                        // debug flag 0x100: cycles / 64 in the low 20 bits, tier-1 and tier-2 counts above (profiling aid)
#ifdef KLT_LK_TIMELINE
                        if (L.flags & 0x400) L.iters[gid] = (int)((t_g0 >> 7) & 0x7fff) | (int)(((gtimer() >> 7) & 0x7fff) << 15);
                        else
#endif
                        L.iters[gid] = (L.flags & 0x100) ? (int)(((clock64() - t_start) >> 6) & 0xfffff) | (min(n_t1, 63) << 20) | (min(n_t2, 63) << 26) : iters;
                    }
                }
            } else {
                c_op |= level << 16;
            }
            if (lane == 0) {
                mbox[0] = make_int4(c_op, c_a, (int)c_W0, (int)c_W1);
                mbox[1] = make_int4(c_jx0, c_jy0, c_b, bidx);
            }
        }
        __syncthreads();     // the commands of this round are posted

        // ================= all threads execute the commands, slot by slot ==================================================
        bool any = false;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            uint8_t* wss = smem + s * C::POINT_BYTES;
            const int4 m0 = reinterpret_cast<const int4*>(wss + C::OFF_CMD)[0];
            const int op = m0.x & 0xff;
            if (op == OP_NONE) continue;     // uniform over the CTA
            any = true;
            if (op == OP_STAGE) continue;    // below: the copies are issued after the round's other work
            const int4 m1 = reinterpret_cast<const int4*>(wss + C::OFF_CMD)[1];
            exec_cmd<WW, WH>(L, m0.x, m0.y, m1.z, m1.x, m1.y, (uint32_t)m0.z, (uint32_t)m0.w, m1.w, wss, tid, pxs[s], dd[s], pads_zeroed[s]);
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            uint8_t* wss = smem + s * C::POINT_BYTES;
            const int4 m0 = reinterpret_cast<const int4*>(wss + C::OFF_CMD)[0];
            if ((m0.x & 0xff) != OP_STAGE) continue;
            const int4 m1 = reinterpret_cast<const int4*>(wss + C::OFF_CMD)[1];
            exec_stage<WW, WH>(L, m0.x, m0.y, m1.z, m1.x, m1.y, m1.w, wss, tid);
        }
        if (!any) break;
        __syncthreads();     // partial sums / replay results are in shared memory
    }
}

template <int WW, int WH, int S>
klt_status launch_fast(const LKLaunch& L, cudaStream_t stream)
{
    using C = Cfg<WW, WH, kWarps>;
    static_assert(WW * WH <= 2056, "err pass assumes an exact float32 sum");
    static_assert(C::UPT <= 8, "per-thread bound accumulators would overflow");
    static PerDeviceOnce configured;
    const size_t smem = (size_t)C::POINT_BYTES * S;
    if (configured.needed()) {
        const cudaError_t e = cudaFuncSetAttribute(lk_fast_kernel<WW, WH, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (klt_status)e;
    }
    const long long total = (long long)L.n_per_pair * L.batch;
    const long long blocks = (total + S - 1) / S;
    if (blocks > 0x7fffffffLL) return KLT_ERR_UNSUPPORTED;
    lk_fast_kernel<WW, WH, S><<<(unsigned)blocks, kThreads, smem, stream>>>(L);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? KLT_OK : (klt_status)e;
}

template <int WW, int WH>
klt_status launch_window(const LKLaunch& L, int slots, cudaStream_t stream)
{
    if (slots == 1) return launch_fast<WW, WH, 1>(L, stream);
    if (slots == 2) return launch_fast<WW, WH, 2>(L, stream);
    return launch_fast<WW, WH, 4>(L, stream);
}

}  // namespace

// Returns KLT_ERR_UNSUPPORTED when no specialisation exists (the caller then uses the generic kernel).
klt_status lk_launch_fast(const LKLaunch& L, int sm_count, int forced_slots, cudaStream_t stream)
{
    // keypoints per CTA (KLT_LK_SLOTS forces it): 21x21 -> 4; 31x31 (two units per thread and slot) -> 2
    int slots = (L.win_w * L.win_h <= 21 * 21) ? 4 : 2;
    if (forced_slots == 1 || forced_slots == 2 || forced_slots == 4) slots = forced_slots;
    (void)sm_count;
    if (L.win_w == 21 && L.win_h == 21) return launch_window<21, 21>(L, slots, stream);
    if (L.win_w == 31 && L.win_h == 31) return launch_window<31, 31>(L, slots, stream);
    return KLT_ERR_UNSUPPORTED;
}

}  // namespace klt
