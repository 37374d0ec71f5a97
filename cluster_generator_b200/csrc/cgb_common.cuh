// Shared device/host helpers for libcgb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/cgb200.h"

// ---------------------------------------------------------------------------
// error plumbing: nothing throws across the C ABI
// ---------------------------------------------------------------------------
void cgb_set_error(const char* fmt, ...);

#define CGB_CUDA_CHECK(expr)                                                          \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            cgb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                          __FILE__, __LINE__);                                        \
            return CGB_E_CUDA;                                                        \
        }                                                                             \
    } while (0)

#define CGB_REQUIRE(cond, ...)                                                        \
    do {                                                                              \
        if (!(cond)) {                                                                \
            cgb_set_error(__VA_ARGS__);                                               \
            return CGB_E_INVALID;                                                     \
        }                                                                             \
    } while (0)

// ---------------------------------------------------------------------------
// `make checked` (-DCGB_BOUNDS_CHECK) turns every table / tile-buffer / output index of the
// kernels into a device-side assert: the build the bounds test runs, since compute-sanitizer
// is not available on the GPU pool.  Compiles to nothing in the shipped library.
// ---------------------------------------------------------------------------
#ifdef CGB_BOUNDS_CHECK
#include <assert.h>
#define CGB_CHECK_INDEX(i, n) assert((long long)(i) >= 0 && (long long)(i) < (long long)(n))
#define CGB_CHECK(cond) assert(cond)
#else
#define CGB_CHECK_INDEX(i, n) ((void)0)
#define CGB_CHECK(cond) ((void)0)
#endif

static inline size_t cgb_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. SC'11).  Contract in include/cgb200.h.
// ---------------------------------------------------------------------------
#define CGB_PHILOX_M0 0xD2511F53u
#define CGB_PHILOX_M1 0xCD9E8D57u
#define CGB_PHILOX_W0 0x9E3779B9u
#define CGB_PHILOX_W1 0xBB67AE85u

struct Philox4 {
    uint32_t x, y, z, w;
};

// 52-bit uniform in [0,1): [1,2) - 1 with mantissa (hi:32 | lo>>12)
__device__ __forceinline__ double u52(uint32_t hi, uint32_t lo)
{
    const uint32_t top = 0x3FF00000u | (hi >> 12);
    const uint32_t bot = (hi << 20) | (lo >> 12);
    return __hiloint2double((int)top, (int)bot) - 1.0;
}

// exact (double) value of a u32 without the I2F pipe
__device__ __forceinline__ double u32_to_double(uint32_t w)
{
    return __hiloint2double(0x43300000, (int)w) - 4503599627370496.0;
}

// ---------------------------------------------------------------------------
// FITPACK-style cubic B-spline helpers (fpbspl / splev / splder arithmetic)
// ---------------------------------------------------------------------------
// largest l in [3, n-5] with t[l] <= x (else 3); n = number of knots
__device__ __forceinline__ int cgb_find_interval(const double* __restrict__ t, int n, double x)
{
    int lo = 3, hi = n - 4;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (t[mid] <= x) lo = mid; else hi = mid;
    }
    CGB_CHECK(n < 8 || (lo >= 3 && lo <= n - 5));
    return lo;
}

template <int K>
__device__ __forceinline__ void cgb_fpbspl(const double* __restrict__ t, double x, int l, double* h)
{
    double hh[4];
    h[0] = 1.0;
#pragma unroll
    for (int j = 1; j <= K; ++j) {
#pragma unroll
        for (int i = 0; i < j; ++i) hh[i] = h[i];
        h[0] = 0.0;
#pragma unroll
        for (int i = 1; i <= j; ++i) {
            const double tli = t[l + i], tlj = t[l + i - j];
            if (tli == tlj) { h[i] = 0.0; continue; }
            const double f = __ddiv_rn(hh[i - 1], __dsub_rn(tli, tlj));
            h[i - 1] = __dadd_rn(h[i - 1], __dmul_rn(f, __dsub_rn(tli, x)));
            h[i] = __dmul_rn(f, __dsub_rn(x, tlj));
        }
    }
}

// s(x), FITPACK operation order, no FMA contraction (bit-faithful to splev)
__device__ __forceinline__ double cgb_splev(const double* __restrict__ t, int n,
                                            const double* __restrict__ c, double x)
{
    const int l = cgb_find_interval(t, n, x);
    double h[4];
    cgb_fpbspl<3>(t, x, l, h);
    double sp = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) sp = __dadd_rn(sp, __dmul_rn(c[l - 3 + j], h[j]));
    return sp;
}

// s'(x) from precomputed splder coefficients wrk[i] = 3 (c[i+1]-c[i]) / (t[i+4]-t[i+1])
__device__ __forceinline__ double cgb_splder1(const double* __restrict__ t, int n,
                                              const double* __restrict__ wrk, double x)
{
    const int l = cgb_find_interval(t, n, x);
    double h[4];
    cgb_fpbspl<2>(t, x, l, h);
    double sp = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) sp = __dadd_rn(sp, __dmul_rn(wrk[l - 3 + j], h[j]));
    return sp;
}
