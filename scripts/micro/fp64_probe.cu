// Micro-benchmark: cost of FP64-pipe instructions on this GPU (cycles per warp-instruction) as a function of active
// lanes, dependence and warps per SM.  Build: nvcc -arch=sm_100a -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dep_dadd(double* out, int lanes, int iters, long long* cyc)
{
    double s = threadIdx.x * 1e-3, d = 1.000001;
    long long t0 = 0, t1 = 0;
    if ((threadIdx.x & 31) < lanes) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 16; ++j) s = __dadd_rn(s, d);
        }
        t1 = clock64();
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void indep_dadd(double* out, int lanes, int iters, long long* cyc)
{
    double s[8];
    for (int j = 0; j < 8; ++j) s[j] = threadIdx.x * 1e-3 + j;
    const double d = 1.000001;
    long long t0 = 0, t1 = 0;
    if ((threadIdx.x & 31) < lanes) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int j = 0; j < 8; ++j) s[j] = __dadd_rn(s[j], d);
        }
        t1 = clock64();
    }
    double a = 0; for (int j = 0; j < 8; ++j) a += s[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void cvt_f2d(double* out, const float* in, int iters, long long* cyc)
{
    float f[8];
    for (int j = 0; j < 8; ++j) f[j] = in[threadIdx.x + j];
    unsigned long long acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) { double d = (double)f[j]; acc ^= (unsigned long long)__double_as_longlong(d); f[j] = __uint_as_float(__float_as_uint(f[j]) + 1u); }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = __longlong_as_double((long long)acc);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void cvt_d2f(float* out, const double* in, int iters, long long* cyc)
{
    double f[8];
    for (int j = 0; j < 8; ++j) f[j] = in[threadIdx.x + j];
    unsigned acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) { float d = __double2float_rn(f[j]); acc ^= __float_as_uint(d); f[j] = __longlong_as_double(__double_as_longlong(f[j]) + 12345); }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double* out; long long* cyc; float* fin; double* din;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8); cudaMalloc(&fin, 1 << 16); cudaMalloc(&din, 1 << 16);
    cudaMemset(fin, 0x3f, 1 << 16); cudaMemset(din, 0x3f, 1 << 16);
    const int iters = 2000;
    long long h;
    for (int threads : {32, 64, 128, 256}) {
        for (int lanes : {32, 16, 8, 1}) {
            dep_dadd<<<1, threads>>>(out, lanes, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            const double dep = (double)h / (iters * 16.0);
            indep_dadd<<<1, threads>>>(out, lanes, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("threads/SM %3d lanes %2d: dependent DADD %.1f cyc/instr, independent DADD %.1f cyc/instr (per warp)\n", threads, lanes, dep,
                   (double)h / (iters * 16.0));
        }
    }
    for (int threads : {32, 128}) {
        cvt_f2d<<<1, threads>>>(out, fin, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("threads %3d: F2F.F64.F32 (+int ops) %.1f cyc/instr\n", threads, (double)h / (iters * 16.0));
        cvt_d2f<<<1, threads>>>((float*)out, din, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("threads %3d: F2F.F32.F64 (+int ops) %.1f cyc/instr\n", threads, (double)h / (iters * 16.0));
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
