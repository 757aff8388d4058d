// Bidirectional-error / bounds filter of the reference's tracking step on the device.
//
// Replaces, per point, reference src/extractor/extractor.py:46-47,53 (extend_tracks) and :67-68,75 (extend_landmarks):
//     d = abs(p0 - p0r).reshape(-1, 2).max(-1);  good = d < max_bidir_error
//     keep = good and 0 <= x <= W and 0 <= y <= H          (x, y) = p1, inclusive bounds
// where p1 = LK(im0, im1, p0) and p0r = LK(im0, im1, p1) -- the reference's second call runs in the SAME direction,
// started from the forward result; it is reproduced as is.  float32 arithmetic like numpy's: |a - b| rounds once,
// max() and the comparisons are exact, NaN fails every test.
#include "klt_common.cuh"

namespace klt {

namespace {

__global__ void __launch_bounds__(256)
track_filter_kernel(const float2* __restrict__ p0, const float2* __restrict__ p1, const float2* __restrict__ p0r,
                    long long n, float max_bidir_error, float w, float h, uint8_t* __restrict__ keep,
                    float* __restrict__ bidir)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 a = p0[i], b = p0r[i], q = p1[i];
    const float dx = fabsf(__fsub_rn(a.x, b.x)), dy = fabsf(__fsub_rn(a.y, b.y));
    // numpy's max propagates NaN; fmaxf would drop it
    const float d = (dx != dx || dy != dy) ? __int_as_float(0x7fc00000) : fmaxf(dx, dy);
    const bool good = d < max_bidir_error;
    const bool inside = (0.f <= q.x) && (q.x <= w) && (0.f <= q.y) && (q.y <= h);
    keep[i] = (good && inside) ? 1 : 0;
    if (bidir) bidir[i] = d;
}

}  // namespace

klt_status track_filter_launch(const float* p0, const float* p1, const float* p0r, long long n, float max_bidir_error,
                               int w, int h, uint8_t* keep, float* bidir, cudaStream_t stream)
{
    if (n < 0 || w <= 0 || h <= 0) return KLT_ERR_INVALID_ARG;
    if (n == 0) return KLT_OK;
    if (!p0 || !p1 || !p0r || !keep) return KLT_ERR_INVALID_ARG;
    const long long blocks = (n + 255) / 256;
    if (blocks > 0x7fffffffLL) return KLT_ERR_UNSUPPORTED;
    track_filter_kernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const float2*>(p0), reinterpret_cast<const float2*>(p1),
                                                               reinterpret_cast<const float2*>(p0r), n, max_bidir_error, (float)w, (float)h,
                                                               keep, bidir);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? KLT_OK : (klt_status)e;
}

}  // namespace klt
