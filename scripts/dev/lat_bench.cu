// Dev microbenchmark: dependent-chain latencies (cycles) of fp64 operations on one thread.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int OP> __global__ void k(double a, double b, double* out, long long* cyc)
{
    __shared__ double sm[64];
    double x = a, y = b;
    sm[threadIdx.x] = a;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = __fma_rn(x, y, y);
        if (OP == 1) x = __dmul_rn(x, y);
        if (OP == 2) x = __dadd_rn(x, y);
        if (OP == 3) x = __ddiv_rn(y, x) + 1.0;
        if (OP == 4) x = __dsqrt_rn(x) + 1.0;
        if (OP == 5) x = __drcp_rn(x) + 1.0;
        if (OP == 6) { sm[(i & 31)] = x; __syncwarp(); x = sm[(i & 31)] + 0.0; }   // smem store->load
        if (OP == 7) { x = __dmul_rn(x, y); x = __dsub_rn(x, b); }
        if (OP == 8) { __syncwarp(); x = __dadd_rn(x, y); }
        if (OP == 9) { float f = (float)x; f = sqrtf(f); x = (double)f + 1.0; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double* out; long long* cyc; cudaMalloc(&out, 512); cudaMallocManaged(&cyc, 8);
    const char* names[] = {"dfma", "dmul", "dadd", "ddiv(+dadd)", "dsqrt(+dadd)", "drcp(+dadd)", "smem st->ld(+dadd)", "dmul+dsub", "syncwarp+dadd", "f32 cvt+sqrtf+cvt+dadd"};
#define RUN(OP) k<OP><<<1, 32>>>(1.000001, 0.999999, out, cyc); cudaDeviceSynchronize(); k<OP><<<1, 32>>>(1.000001, 0.999999, out, cyc); cudaDeviceSynchronize(); printf("%-24s %.1f cycles/iter\n", names[OP], (double)*cyc / N);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
    return 0;
}
