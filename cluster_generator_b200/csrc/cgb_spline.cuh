// internal launchers shared between translation units
#pragma once
#include <cuda_runtime.h>

// d_work: n*18 doubles.  reuse_factor = 1: d_work still holds the Givens factorisation of a
// previous CGB_FIT_FITPACK fit over the SAME x (only the right-hand side is processed).
// d_wrk (nullable, n-1 doubles): also emit the splder coefficients of the fitted spline.
int cgb_launch_spline_fit(const double* d_x, const double* d_y, int n, double* d_t, double* d_c,
                          void* d_work, cudaStream_t s, int method, int reuse_factor, double* d_wrk);
// d_wrk: n-1 doubles
int cgb_launch_splder_coeffs(const double* d_t, const double* d_c, int n, double* d_wrk, cudaStream_t s);
// d_brk: n-2 doubles, d_pp: (n-3)*4 doubles
int cgb_launch_pp_convert(const double* d_t, const double* d_c, int n, double* d_brk, double* d_pp,
                          cudaStream_t s);
