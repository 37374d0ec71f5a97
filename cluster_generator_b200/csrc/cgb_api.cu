// Error plumbing, device queries and the pipe-peak microbenchmarks behind the C ABI.
#include <string.h>
#include "cgb_common.cuh"

static thread_local char g_err[512] = "";

void cgb_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int cgb_version(void) { return CGB_VERSION; }
extern "C" const char* cgb_last_error(void) { return g_err; }

// Bounds-checked build (make checked, -DCGB_BOUNDS_CHECK): every table / tile-buffer / output
// index of the kernels is asserted on the device.  cgb_bounds_check_enabled says which build
// this is; cgb_test_bounds_check launches one thread that checks index i against size n and
// reports whether the device trapped -- the control that proves the asserts are live (a trap
// leaves the CUDA context unusable, so run it in a process of its own).
extern "C" int cgb_bounds_check_enabled(void)
{
#ifdef CGB_BOUNDS_CHECK
    return 1;
#else
    return 0;
#endif
}

__global__ void cgb_bounds_probe_kernel(long long i, long long n, int* out)
{
    CGB_CHECK_INDEX(i, n);
    *out = 1;
}

extern "C" int cgb_test_bounds_check(long long i, long long n)
{
    int* d = nullptr;
    CGB_CUDA_CHECK(cudaMalloc(&d, sizeof(int)));
    cgb_bounds_probe_kernel<<<1, 1>>>(i, n, d);
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cgb_set_error("cgb_test_bounds_check: index %lld of %lld trapped: %s", i, n, cudaGetErrorString(e));
        return CGB_E_CUDA;
    }
    CGB_CUDA_CHECK(cudaFree(d));
    return CGB_OK;
}

extern "C" int cgb_device_sm_count(int* sm_count)
{
    CGB_REQUIRE(sm_count != nullptr, "cgb_device_sm_count: null pointer");
    int dev = 0;
    CGB_CUDA_CHECK(cudaGetDevice(&dev));
    CGB_CUDA_CHECK(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    return CGB_OK;
}

// Register-resident FMA chains: 8 independent accumulators per thread, 1024 threads per
// SM-resident CTA pair, enough ILP to saturate the pipe.  Roofline denominator only.
template <typename T>
__global__ void __launch_bounds__(512) fma_peak_kernel(int iters, T* out)
{
    T a0 = (T)threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
      a6 = a0 + 6, a7 = a0 + 7;
    const T m = (T)0.999, c = (T)0.001;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
            a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
        }
    }
    const T s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == (T)-1) out[0] = s;
}

template <typename T>
static int measure_fma(int iters, double* tflops, void* stream)
{
    CGB_REQUIRE(tflops != nullptr && iters > 0, "measure_fma: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    int sms = 0;
    int rc = cgb_device_sm_count(&sms);
    if (rc) return rc;
    T* d_out = nullptr;
    CGB_CUDA_CHECK(cudaMalloc(&d_out, sizeof(T)));
    cudaEvent_t e0, e1;
    CGB_CUDA_CHECK(cudaEventCreate(&e0));
    CGB_CUDA_CHECK(cudaEventCreate(&e1));
    const int blocks = sms * 4;
    fma_peak_kernel<T><<<blocks, 512, 0, s>>>(iters / 4 + 1, d_out);  // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CGB_CUDA_CHECK(cudaEventRecord(e0, s));
        fma_peak_kernel<T><<<blocks, 512, 0, s>>>(iters, d_out);
        CGB_CUDA_CHECK(cudaEventRecord(e1, s));
        CGB_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CGB_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CGB_CUDA_CHECK(cudaGetLastError());
    const double flops = 2.0 * 64.0 * (double)iters * 512.0 * (double)blocks;
    *tflops = flops / ((double)best * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return CGB_OK;
}

extern "C" int cgb_measure_dfma_tflops(int iters, double* tflops, void* stream)
{
    return measure_fma<double>(iters, tflops, stream);
}

extern "C" int cgb_measure_ffma_tflops(int iters, double* tflops, void* stream)
{
    return measure_fma<float>(iters, tflops, stream);
}

// ---------------------------------------------------------------------------
// peer-visible buffers (CUDA IPC) for the fused moment all-reduce
// ---------------------------------------------------------------------------
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");

extern "C" int cgb_peer_alloc(size_t bytes, void** d_ptr, unsigned char handle_out[64])
{
    CGB_REQUIRE(bytes > 0 && d_ptr && handle_out, "cgb_peer_alloc: bad argument");
    void* p = nullptr;
    CGB_CUDA_CHECK(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        cgb_set_error("cgb_peer_alloc: %s", cudaGetErrorString(e));
        return CGB_E_CUDA;
    }
    memcpy(handle_out, &h, 64);
    *d_ptr = p;
    return CGB_OK;
}

extern "C" int cgb_peer_open(const unsigned char handle[64], void** d_ptr)
{
    CGB_REQUIRE(handle && d_ptr, "cgb_peer_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CGB_CUDA_CHECK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return CGB_OK;
}

extern "C" int cgb_peer_close(void* d_ptr)
{
    if (d_ptr) CGB_CUDA_CHECK(cudaIpcCloseMemHandle(d_ptr));
    return CGB_OK;
}

extern "C" int cgb_peer_free(void* d_ptr)
{
    if (d_ptr) CGB_CUDA_CHECK(cudaFree(d_ptr));
    return CGB_OK;
}
