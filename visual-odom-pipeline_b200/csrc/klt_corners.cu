// K5: Shi-Tomasi corner detection -- what the reference delegates to
// cv2.goodFeaturesToTrack(img, mask=mask, maxCorners=1000, qualityLevel=0.03, minDistance, blockSize=31) at
// reference src/extractor/extractor.py:21-24,110-111 (called once per frame from src/pipeline/pipeline.py:159-163).
//
// Arithmetic: oracle/gftt_oracle.c G.1-G.7 (float32 Sobel with OpenCV's AVX2 fused-multiply-add pattern, float32
// products, DOUBLE running box sums in OpenCV's order, non-fused eigenvalue formula) -- bit-exact with the cv2 wheel.
// Four kernels per image batch:
//   cov_kernel        32x32 pixel tiles: u8 tile + halo in shared memory -> Dx, Dy -> (Dx^2, DxDy, Dy^2), written
//                     TRANSPOSED ([channel][x][y]) so that the row scan reads it coalesced;
//   row_scan_kernel   one lane per image row (a warp = 32 rows of one channel): the serial double-precision running
//                     sum along x (the order is part of the result: the sums are not exact), loads prefetched a chunk
//                     ahead, results transposed through shared memory and written row-major;
//   col_scan_kernel   one thread per image column, the three channels as three independent chains: running sum along
//                     y, float32 conversion, eigenvalue, masked maximum (REDUX + one atomicMax per warp);
//   candidates_kernel threshold (maxVal * qualityLevel), 3x3 dilation and local-maximum test fused; survivors are
//                     appended as 64-bit keys (ordered float << 32 | y*W+x) with one atomicAdd per warp.
// The sort of the (few thousand) keys and the greedy minimum-distance selection (G.8) are inherently sequential
// and run on the host (klt_capi.cu), like the tail of cv::cuda::GoodFeaturesToTrackDetector.
#include "klt_common.cuh"

namespace klt {

namespace {

constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ unsigned ordered_from_float(float f)
{
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_from_ordered(unsigned u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// ---- cov_kernel -----------------------------------------------------------------------------------------------
constexpr int kTile = 32;

__global__ void __launch_bounds__(256)
cov_kernel(const uint8_t* __restrict__ img, long long pitch, long long batch_stride, int w, int h,
           float* __restrict__ covT, int hp, float k1, float k0)
{
    __shared__ uint8_t sI[kTile + 2][kTile + 4];
    __shared__ float sT[kTile + 2][kTile + 1];   // smoothed rows (for Dy)
    __shared__ int sR[kTile + 2][kTile + 1];     // horizontal differences (for Dx)
    __shared__ float sC[3][kTile][kTile + 1];    // [channel][x][y]
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
    const uint8_t* __restrict__ src = img + (long long)blockIdx.z * batch_stride;
    for (int i = ty * 32 + tx; i < (kTile + 2) * (kTile + 2); i += 256) {
        const int r = i / (kTile + 2), c = i - r * (kTile + 2);
        const int yy = reflect101(min(y0 + r - 1, h), h), xx = reflect101(min(x0 + c - 1, w), w);
        sI[r][c] = __ldg(src + (long long)yy * pitch + xx);
    }
    __syncthreads();
    const int nv = 32 * (w / 32);
    for (int r = ty; r < kTile + 2; r += 8) {
        const float L = (float)sI[r][tx], C = (float)sI[r][tx + 1], R = (float)sI[r][tx + 2];
        sR[r][tx] = (int)sI[r][tx + 2] - (int)sI[r][tx];
        float t;
        if (x0 + tx < nv) t = __fmaf_rn(k1, R, __fmaf_rn(k0, C, __fmul_rn(k1, L)));                       // G.3, SIMD part
        else t = __fadd_rn(__fadd_rn(__fmul_rn(k1, L), __fmul_rn(k0, C)), __fmul_rn(k1, R));             // scalar tail
        sT[r][tx] = t;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int y = ty + 8 * k;
        const float dx = __fmaf_rn(k1, (float)(sR[y][tx] + sR[y + 2][tx]), __fmul_rn(k0, (float)sR[y + 1][tx]));   // G.2
        const float dy = __fsub_rn(sT[y + 2][tx], sT[y][tx]);
        sC[0][tx][y] = __fmul_rn(dx, dx);
        sC[1][tx][y] = __fmul_rn(dx, dy);
        sC[2][tx][y] = __fmul_rn(dy, dy);
    }
    __syncthreads();
    float* __restrict__ dst = covT + (long long)blockIdx.z * 3 * w * hp;
    if (y0 + tx < h) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int x = ty + 8 * k;
            if (x0 + x < w) {
#pragma unroll
                for (int c = 0; c < 3; ++c) dst[((long long)c * w + (x0 + x)) * hp + (y0 + tx)] = sC[c][x][tx];
            }
        }
    }
}

// ---- row_scan_kernel ------------------------------------------------------------------------------------------
constexpr int kChunk = 16;

// padded column i (0 <= i < w + block - 1) -> source column (reflect-101 of i - anchor)
__device__ __forceinline__ int src_col(int i, int an, int w) { return reflect101(i - an, w); }

__global__ void __launch_bounds__(32)
row_scan_kernel(const float* __restrict__ covT, int w, int h, int hp, int block, double* __restrict__ rows, int wd)
{
    __shared__ double sD[32][kChunk + 1];
    const int lane = threadIdx.x;
    const int y0 = blockIdx.x * 32, c = blockIdx.y;
    const int y = min(y0 + lane, h - 1);
    const int an = block / 2;
    const float* __restrict__ S = covT + ((long long)blockIdx.z * 3 + c) * w * hp + y;   // element x at S[x * hp]
    double* __restrict__ D = rows + (((long long)blockIdx.z * 3 + c) * h + y0) * wd;      // row r at D + r * wd
    const int nrows = min(32, h - y0);

    auto flush = [&](int xbase, int n) {   // write columns [xbase, xbase + n) of the 32 rows, coalesced
        __syncwarp();
        for (int r = 0; r < nrows; ++r)
            if (lane < n) D[(long long)r * wd + xbase + lane] = sD[r][lane];
        __syncwarp();
    };

    if (block == 3 || block == 5) {
        // OpenCV's RowSum forms fresh left-to-right sums for these two kernel sizes
        for (int xb = 0; xb < w; xb += kChunk) {
            const int n = min(kChunk, w - xb);
            for (int j = 0; j < n; ++j) {
                double s = (double)S[(long long)src_col(xb + j, an, w) * hp];
                for (int k = 1; k < block; ++k) s = __dadd_rn(s, (double)S[(long long)src_col(xb + j + k, an, w) * hp]);
                sD[lane][j] = s;
            }
            flush(xb, n);
        }
        return;
    }
    double s = 0.0;
    for (int i = 0; i < block; ++i) s = __dadd_rn(s, (double)S[(long long)src_col(i, an, w) * hp]);
    // D[0] = s, then D[x + 1] = (s += new - old) for x = 0 .. w - 2: chunk k covers outputs [k*kChunk, (k+1)*kChunk)
    float lead[kChunk], trail[kChunk];
    auto fetch = [&](int xb) {   // operands of outputs xb + 1 + j, j = 0 .. kChunk-1 (steps x = xb + j)
#pragma unroll
        for (int j = 0; j < kChunk; ++j) {
            const int x = min(xb + j, w - 2 >= 0 ? w - 2 : 0);
            lead[j] = S[(long long)src_col(x + block, an, w) * hp];
            trail[j] = S[(long long)src_col(x, an, w) * hp];
        }
    };
    // output 0
    sD[lane][0] = s;
    int fill = 1, xbase = 0;   // sD holds outputs [xbase, xbase + fill)
    for (int xb = 0; xb < w - 1; xb += kChunk) {
        fetch(xb);
        const int n = min(kChunk, w - 1 - xb);
#pragma unroll
        for (int j = 0; j < kChunk; ++j) {
            if (j < n) {
                s = __dadd_rn(s, __dsub_rn((double)lead[j], (double)trail[j]));
                sD[lane][fill] = s;
                if (++fill == kChunk) { flush(xbase, kChunk); xbase += kChunk; fill = 0; }
            }
        }
    }
    if (fill) flush(xbase, fill);
}

// ---- col_scan_kernel ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
col_scan_kernel(const double* __restrict__ rows, int w, int h, int wd, int block, float* __restrict__ eig, long long eig_pitch,
                long long eig_batch_stride, const uint8_t* __restrict__ mask, long long mask_pitch, long long mask_batch_stride,
                unsigned* __restrict__ max_out)
{
    const int x = blockIdx.x * 64 + threadIdx.x;
    const int xc = min(x, w - 1);
    const int an = block / 2;
    const long long plane = (long long)h * wd;
    const double* __restrict__ R0 = rows + (long long)blockIdx.z * 3 * plane + xc;
    const double* __restrict__ R1 = R0 + plane;
    const double* __restrict__ R2 = R1 + plane;
    float* __restrict__ E = eig + (long long)blockIdx.z * eig_batch_stride;
    const uint8_t* __restrict__ M = mask ? mask + (long long)blockIdx.z * mask_batch_stride : nullptr;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int i = 0; i < block - 1; ++i) {
        const long long o = (long long)reflect101(i - an, h) * wd;
        s0 = __dadd_rn(s0, R0[o]); s1 = __dadd_rn(s1, R1[o]); s2 = __dadd_rn(s2, R2[o]);
    }
    unsigned best = 0;   // 0 = nothing seen (every real float maps above it)
#pragma unroll 4
    for (int y = 0; y < h; ++y) {
        const long long op = (long long)reflect101(y + block - 1 - an, h) * wd;
        const long long om = (long long)reflect101(y - an, h) * wd;
        const double p0 = R0[op], p1 = R1[op], p2 = R2[op];
        const double m0 = R0[om], m1 = R1[om], m2 = R2[om];
        const double t0 = __dadd_rn(s0, p0), t1 = __dadd_rn(s1, p1), t2 = __dadd_rn(s2, p2);
        s0 = __dsub_rn(t0, m0); s1 = __dsub_rn(t1, m1); s2 = __dsub_rn(t2, m2);
        const float a = __fmul_rn(__double2float_rn(t0), 0.5f), b = __double2float_rn(t1), c = __fmul_rn(__double2float_rn(t2), 0.5f);
        const float d = __fsub_rn(a, c);
        const float e = __fsub_rn(__fadd_rn(a, c), __fsqrt_rn(__fadd_rn(__fmul_rn(d, d), __fmul_rn(b, b))));   // G.6
        if (x < w) {
            E[(long long)y * eig_pitch + x] = e;
            if (max_out && (!M || M[(long long)y * mask_pitch + x])) best = max(best, ordered_from_float(e));
        }
    }
    if (max_out) {
        best = __reduce_max_sync(kFullMask, best);
        if ((threadIdx.x & 31) == 0 && best) atomicMax(max_out + blockIdx.z, best);
    }
}

// ---- candidates_kernel ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
candidates_kernel(const float* __restrict__ eig, long long eig_pitch, long long eig_batch_stride, int w, int h,
                  const uint8_t* __restrict__ mask, long long mask_pitch, long long mask_batch_stride,
                  const unsigned* __restrict__ max_in, double quality, unsigned long long* __restrict__ keys,
                  long long keys_batch_stride, int capacity, unsigned* __restrict__ count)
{
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int b = blockIdx.z;
    const unsigned mo = max_in[b];
    const float max_val = mo ? float_from_ordered(mo) : 0.f;                 // minMaxLoc over an empty mask gives 0
    const float thr = __double2float_rn(__dmul_rn((double)max_val, quality));   // G.7
    bool is_cand = false;
    float v = 0.f;
    if (x >= 1 && x < w - 1 && y >= 1 && y < h - 1) {
        const float* __restrict__ E = eig + (long long)b * eig_batch_stride + (long long)y * eig_pitch + x;
        v = E[0];
        v = v > thr ? v : 0.f;
        if (v != 0.f && (!mask || mask[(long long)b * mask_batch_stride + (long long)y * mask_pitch + x])) {
            float m = v;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    float u = E[(long long)dy * eig_pitch + dx];
                    u = u > thr ? u : 0.f;
                    m = fmaxf(m, u);
                }
            is_cand = (v == m);
        }
    }
    const unsigned ballot = __ballot_sync(kFullMask, is_cand);
    if (ballot) {
        const int lane = threadIdx.x;
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(count + b, (unsigned)__popc(ballot));
        base = __shfl_sync(kFullMask, base, 0);
        if (is_cand) {
            const unsigned slot = base + __popc(ballot & ((1u << lane) - 1u));
            if (slot < (unsigned)capacity)
                keys[(long long)b * keys_batch_stride + slot] = ((unsigned long long)ordered_from_float(v) << 32) | (unsigned)(y * w + x);
        }
    }
}

}  // namespace

long long corners_ws_bytes(int w, int h, int batch)
{
    const long long hp = (h + 31) / 32 * 32, wd = (w + 3) / 4 * 4;
    const long long cov = 3LL * w * hp * 4, rows = 3LL * h * wd * 8;
    return ((cov + 255) / 256 * 256 + (rows + 255) / 256 * 256) * batch;
}

klt_status corner_min_eig_launch(const uint8_t* img, long long pitch, long long batch_stride, int w, int h, int batch,
                                 int block, float* eig, long long eig_pitch, long long eig_batch_stride,
                                 const uint8_t* mask, long long mask_pitch, long long mask_batch_stride,
                                 unsigned* max_out, void* ws, cudaStream_t stream)
{
    if (w < 1 || h < 1 || batch < 1 || block < 1) return KLT_ERR_INVALID_ARG;
    if (block / 2 >= w || block / 2 >= h || batch > 65535) return KLT_ERR_UNSUPPORTED;
    const int hp = (h + 31) / 32 * 32, wd = (w + 3) / 4 * 4;
    const long long cov_bytes = (3LL * w * hp * 4 * batch + 255) / 256 * 256;
    float* covT = static_cast<float*>(ws);
    double* rows = reinterpret_cast<double*>(static_cast<uint8_t*>(ws) + cov_bytes);
    const float k1 = (float)(1.0 / (4.0 * (double)block * 255.0));   // G.1
    const float k0 = (float)(2.0 / (4.0 * (double)block * 255.0));
    cov_kernel<<<dim3((w + kTile - 1) / kTile, (h + kTile - 1) / kTile, batch), dim3(32, 8), 0, stream>>>(
        img, pitch, batch_stride, w, h, covT, hp, k1, k0);
    row_scan_kernel<<<dim3((h + 31) / 32, 3, batch), 32, 0, stream>>>(covT, w, h, hp, block, rows, wd);
    col_scan_kernel<<<dim3((w + 63) / 64, 1, batch), 64, 0, stream>>>(rows, w, h, wd, block, eig, eig_pitch, eig_batch_stride,
                                                                       mask, mask_pitch, mask_batch_stride, max_out);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? KLT_OK : (klt_status)e;
}

klt_status corner_candidates_launch(const float* eig, long long eig_pitch, long long eig_batch_stride, int w, int h, int batch,
                                    const uint8_t* mask, long long mask_pitch, long long mask_batch_stride,
                                    const unsigned* max_in, double quality, unsigned long long* keys,
                                    long long keys_batch_stride, int capacity, unsigned* count, cudaStream_t stream)
{
    if (w < 1 || h < 1 || batch < 1 || capacity < 0 || !(quality > 0)) return KLT_ERR_INVALID_ARG;
    if (batch > 65535 || (long long)w * h > 0xffffffffLL) return KLT_ERR_UNSUPPORTED;
    candidates_kernel<<<dim3((w + 31) / 32, (h + 7) / 8, batch), dim3(32, 8), 0, stream>>>(
        eig, eig_pitch, eig_batch_stride, w, h, mask, mask_pitch, mask_batch_stride, max_in, quality, keys, keys_batch_stride,
        capacity, count);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? KLT_OK : (klt_status)e;
}

}  // namespace klt
