// Row f3 of the scope table: the loader's pre-filter  cv2.bilateralFilter(img, d = 5, sigmaColor = 1.5, sigmaSpace = 1.5)
// (reference src/loader/loader.py:16-20,86) on 8-bit single-channel frames, batched, on the device.
//
// Arithmetic = OpenCV's own code path (oracle/bilateral_oracle.c B.1-B.6): neighbours inside the circle of radius d/2 in
// row-major order, float32  w = space[k] * color[|v_k - v_0|],  wsum += w,  sum = fma(v_k, w, sum),  dst = round-half-even
// (sum / wsum); the last W % 8 columns take the 4-at-a-time order of OpenCV's scalar tail.  The two weight tables are
// computed on the host in double precision exactly as OpenCV does and passed in.
//
// One thread = 4 horizontally adjacent pixels (one 32-bit store); a block of 32 x 8 threads stages its (8 + 2r) x (128 + 2r)
// neighbourhood in shared memory with BORDER_REFLECT_101 resolved at load time.  HBM-bound by construction (1 byte in,
// 1 byte out per pixel); at the sizes of the path (one 0.47 MB KITTI frame) the launch is latency-bound.
#include "klt_common.cuh"

#include <cmath>
#include <cstring>

namespace klt {

namespace {

constexpr int kBX = 32, kBY = 8;          // threads
constexpr int kTW = 4 * kBX, kTH = kBY;   // output tile
constexpr int kMaxRadius = 7;
constexpr int kMaxTaps = (2 * kMaxRadius + 1) * (2 * kMaxRadius + 1);

__global__ void __launch_bounds__(kBX * kBY)
bilateral_kernel(const uint8_t* __restrict__ src, int w, int h, long long spitch, long long sbatch, uint8_t* __restrict__ dst,
                 long long dpitch, long long dbatch, int radius, int n_taps, int x_tail, const float* __restrict__ tab)
{
    // tab: [256] color weights, [n_taps] space weights, [n_taps] packed offsets (dy << 16 | (dx & 0xffff)) as int bits
    extern __shared__ __align__(16) uint8_t smem_raw[];
    float* s_color = reinterpret_cast<float*>(smem_raw);
    float* s_space = s_color + 256;
    int* s_ofs = reinterpret_cast<int*>(s_space + kMaxTaps);
    uint8_t* tile = reinterpret_cast<uint8_t*>(s_ofs + kMaxTaps);
    const int tw = kTW + 2 * radius;               // tile width in bytes (pitch rounded up to 4)
    const int tp = (tw + 3) & ~3;
    const int th = kTH + 2 * radius;
    const int tid = threadIdx.y * kBX + threadIdx.x;
    for (int i = tid; i < 256 + 2 * n_taps; i += kBX * kBY) {
        if (i < 256) s_color[i] = tab[i];
        else if (i < 256 + n_taps) s_space[i - 256] = tab[i];
        else {
            const int o = __float_as_int(tab[i]);                       // dy << 16 | (dx & 0xffff)
            s_ofs[i - 256 - n_taps] = (o >> 16) * tp + (int)(short)(o & 0xffff);   // byte offset inside the tile
        }
    }
    const uint8_t* __restrict__ img = src + (long long)blockIdx.z * sbatch;
    const int X0 = blockIdx.x * kTW, Y0 = blockIdx.y * kTH;
    for (int i = tid; i < th * tw; i += kBX * kBY) {
        const int r = i / tw, c = i - r * tw;
        tile[r * tp + c] = __ldg(img + (long long)reflect101(Y0 - radius + r, h) * spitch + reflect101(X0 - radius + c, w));
    }
    __syncthreads();
    const int y = Y0 + threadIdx.y;
    const int xb = X0 + 4 * threadIdx.x;
    if (y >= h || xb >= w) return;
    const uint8_t* centre = tile + (threadIdx.y + radius) * tp + 4 * threadIdx.x + radius;
    uint32_t out = 0;
#pragma unroll
    for (int px = 0; px < 4; ++px) {
        const int x = xb + px;
        const uint8_t* c0 = centre + px;
        const int v0 = *c0;
        float sum = 0.f, wsum = 0.f;
        if (x < x_tail) {
            for (int k = 0; k < n_taps; ++k) {
                const int v = c0[s_ofs[k]];
                const float wt = __fmul_rn(s_space[k], s_color[abs(v - v0)]);
                wsum = __fadd_rn(wsum, wt);
                sum = __fmaf_rn((float)v, wt, sum);
            }
        } else {
            // OpenCV's scalar tail (columns past the last full 8-lane vector): neighbours four at a time, the sums of the
            // four weights / products formed as (0 + 2) + (1 + 3), products rounded (no fused multiply-add)
            int k = 0;
            for (; k + 4 <= n_taps; k += 4) {
                float wt[4], p[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int v = c0[s_ofs[k + q]];
                    wt[q] = __fmul_rn(s_space[k + q], s_color[abs(v - v0)]);
                    p[q] = __fmul_rn((float)v, wt[q]);
                }
                wsum = __fadd_rn(wsum, __fadd_rn(__fadd_rn(wt[0], wt[2]), __fadd_rn(wt[1], wt[3])));
                sum = __fadd_rn(sum, __fadd_rn(__fadd_rn(p[0], p[2]), __fadd_rn(p[1], p[3])));
            }
            for (; k < n_taps; ++k) {
                const int v = c0[s_ofs[k]];
                const float wt = __fmul_rn(s_space[k], s_color[abs(v - v0)]);
                wsum = __fadd_rn(wsum, wt);
                sum = __fadd_rn(sum, __fmul_rn((float)v, wt));
            }
        }
        const int r = __float2int_rn(__fdiv_rn(sum, wsum));
        out |= (uint32_t)(r & 0xff) << (8 * px);
    }
    uint8_t* drow = dst + (long long)blockIdx.z * dbatch + (long long)y * dpitch + xb;
    if (xb + 4 <= w && ((reinterpret_cast<uintptr_t>(drow) & 3) == 0)) {
        *reinterpret_cast<uint32_t*>(drow) = out;
    } else {
        for (int px = 0; px < 4 && xb + px < w; ++px) drow[px] = (uint8_t)(out >> (8 * px));
    }
}

}  // namespace

int bilateral_radius(int d, double sigma_space)
{
    if (sigma_space <= 0) sigma_space = 1;
    int radius = (d <= 0) ? (int)lrint(sigma_space * 1.5) : d / 2;
    return radius < 1 ? 1 : radius;
}

// Host side of B.2 / B.3: fills tab (256 + 2 * taps floats); returns the number of taps or -1 when radius > kMaxRadius.
int bilateral_tables(int d, double sigma_color, double sigma_space, float* tab, int capacity)
{
    if (sigma_color <= 0) sigma_color = 1;
    if (sigma_space <= 0) sigma_space = 1;
    const int radius = bilateral_radius(d, sigma_space);
    if (radius > kMaxRadius) return -1;
    const double gauss_color_coeff = -0.5 / (sigma_color * sigma_color);
    const double gauss_space_coeff = -0.5 / (sigma_space * sigma_space);
    float space[kMaxTaps];
    int ofs[kMaxTaps];
    int n = 0;
    for (int i = -radius; i <= radius; ++i)
        for (int j = -radius; j <= radius; ++j) {
            const double r = std::sqrt((double)i * i + (double)j * j);
            if (r > radius) continue;
            space[n] = (float)std::exp(r * r * gauss_space_coeff);
            ofs[n] = (int)(((unsigned)i << 16) | ((unsigned)j & 0xffffu));
            ++n;
        }
    if (capacity < 256 + 2 * n) return -1;
    for (int i = 0; i < 256; ++i) tab[i] = (float)std::exp(i * i * gauss_color_coeff);
    for (int k = 0; k < n; ++k) {
        tab[256 + k] = space[k];
        std::memcpy(&tab[256 + n + k], &ofs[k], 4);
    }
    return n;
}

int bilateral_table_capacity() { return 256 + 2 * kMaxTaps; }

klt_status bilateral_launch(const uint8_t* src, int w, int h, long long spitch, long long sbatch, uint8_t* dst, long long dpitch,
                            long long dbatch, int batch, int radius, int n_taps, const float* d_tab, cudaStream_t stream)
{
    if (!src || !dst || !d_tab || w <= 0 || h <= 0 || batch <= 0 || spitch < w || dpitch < w) return KLT_ERR_INVALID_ARG;
    if (radius < 1 || radius > kMaxRadius || n_taps < 1 || n_taps > kMaxTaps || batch > 65535) return KLT_ERR_UNSUPPORTED;
    const int tp = (kTW + 2 * radius + 3) & ~3;
    const size_t smem = (size_t)(256 + 2 * kMaxTaps) * 4 + (size_t)tp * (kTH + 2 * radius);
    static PerDeviceOnce configured;
    if (configured.needed()) {
        const cudaError_t e = cudaFuncSetAttribute(bilateral_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return (klt_status)e;
    }
    const dim3 grid((w + kTW - 1) / kTW, (h + kTH - 1) / kTH, batch);
    if (grid.y > 65535) return KLT_ERR_UNSUPPORTED;
    bilateral_kernel<<<grid, dim3(kBX, kBY), smem, stream>>>(src, w, h, spitch, sbatch, dst, dpitch, dbatch, radius, n_taps,
                                                              w - w % 8, d_tab);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? KLT_OK : (klt_status)e;
}

}  // namespace klt
