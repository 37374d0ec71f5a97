// D: Eddington f(E) build   -- replaces VirialEquilibrium._generate_df (virial.py:43-68)
// C: virial self-check      -- replaces VirialEquilibrium.check_virial (virial.py:78-106)
//
// One warp per energy bin.  Each warp runs a decision-for-decision restatement of
// QUADPACK dqagse (what scipy.integrate.quad calls at virial.py:55-62,99): 21-point
// Gauss-Kronrod on lanes 0..20, bisection of the worst sub-interval, Wynn epsilon
// extrapolation (dqelg), limit = 50.  The 21 integrand values are computed in
// parallel; every sum is then formed by all lanes in QUADPACK's sequential order, so
// the arithmetic equals the scalar algorithm's bit for bit.  The spline tables of the
// integrand are staged in shared memory once per CTA.
//
// Compiled with -fmad=false (plain IEEE double, no contraction), like the oracle.
#include <math.h>
#include <float.h>
#include "cgb_common.cuh"
#include "cgb_spline.cuh"

#define QAGS_LIMIT 50
#define QAGS_NEED_VALUES (-99)  // replay ran out of cached integrand values (CachedIntegrand)
#define EPMACH DBL_EPSILON
#define UFLOW DBL_MIN
#define OFLOW DBL_MAX

__constant__ double c_xgk[11] = {
    0.995657163025808080735527280689003, 0.973906528517171720077964012084452,
    0.930157491355708226001207180059508, 0.865063366688984510732096688423493,
    0.780817726586416897063717578345042, 0.679409568299024406234327365114874,
    0.562757134668604683339000099272694, 0.433395394129247190799265943165784,
    0.294392862701460198131126603103866, 0.148874338981631210884826001129720, 0.0};
__constant__ double c_wgk[11] = {
    0.011694638867371874278064396062192, 0.032558162307964727478818972459390,
    0.054755896574351996031381300244580, 0.075039674810919952767043140916190,
    0.093125454583697605535065465083366, 0.109387158802297641899210590325805,
    0.123491976262065851077958109585166, 0.134709217311473325928054001771707,
    0.142775938577060080797094273138717, 0.147739104901338491374841515972068,
    0.149445554002916905664936468389821};
__constant__ double c_wg[5] = {
    0.066671344308688137593568809893332, 0.149451349150580593145776339657697,
    0.219086362515982043995534934228163, 0.269266719309996355091226921569469,
    0.295524224714752870173815619188769};

// per-warp QAGS scratch in shared memory
struct QagsScratch {
    double alist[QAGS_LIMIT], blist[QAGS_LIMIT], rlist[QAGS_LIMIT], elist[QAGS_LIMIT];
    double rlist2[52];
    double res3la[3];
    int iord[QAGS_LIMIT + 2];
};

// virial.py:47  dgdp = lambda t, e: 2 * density_spline(e - t*t, 1)
struct DgdpIntegrand {
    __device__ __forceinline__ void advance() const {}
    __device__ __forceinline__ bool exhausted() const { return false; }
    const double* t;   // knots of the density spline (shared memory)
    const double* wrk; // splder coefficients (shared memory)
    int nk;
    double e;
    __device__ __forceinline__ double operator()(double tt) const
    {
        return 2.0 * cgb_splder1(t, nk, wrk, e - tt * tt);
    }
};

// virial.py:96  rho_int = lambda e, psi: self.f(e) * np.sqrt(2 * (psi - e))
struct RhoIntegrand {
    __device__ __forceinline__ void advance() const {}
    __device__ __forceinline__ bool exhausted() const { return false; }
    const double* t; // knots of f (shared memory)
    const double* c; // coefficients of f (shared memory)
    int nk;
    double psi;
    __device__ __forceinline__ double operator()(double e) const
    {
        return cgb_splev(t, nk, c, e) * sqrt(2.0 * (psi - e));
    }
};

// x^1.5 rounded to nearest (error < 0.5 ulp + 2^-100): sqrt residual and product in
// double-double with explicit FMAs.  QUADPACK computes this with libm pow (glibc: < 0.52
// ulp), so the two agree except in rare near-tie cases; CUDA's pow (2 ulp) would not.
__device__ __forceinline__ double pow15(double x)
{
    const double s = sqrt(x);
    const double r = fma(-s, s, x);      // x - s*s, exact
    const double c = r / (2.0 * s);      // sqrt(x) = s + c
    const double hi = x * s;
    double lo = fma(x, s, -hi);          // x*s = hi + lo, exact
    lo = fma(x, c, lo);
    return hi + lo;
}

// model.py:219-224 / utils.py:137-142:  profile(r) = spline(r) * r**power, the integrand of
// the potential table (power 1) and of spline-tabulated mass integrals (power 2)
struct PowSplineIntegrand {
    __device__ __forceinline__ void advance() const {}
    __device__ __forceinline__ bool exhausted() const { return false; }
    const double* t;
    const double* c;
    int nk;
    int power;
    __device__ __forceinline__ double operator()(double x) const
    {
        double v = cgb_splev(t, nk, c, x);
        for (int k = 0; k < power; ++k) v = v * x;
        return v;
    }
};

// Warp-cooperative dqk21.  All lanes return identical values.
template <class F>
__device__ __forceinline__ void dqk21_warp(const F& f, double a, double b, double& result,
                                           double& abserr, double& resabs, double& resasc)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const double centr = 0.5 * (a + b), hlgth = 0.5 * (b - a), dhlgth = fabs(hlgth);
    // lane 0: centre; lane 1+2m (+1): node m of the evaluation order, minus (plus) side.
    // evaluation order m = 0..4: Gauss nodes xgk[2m+1]; m = 5..9: Kronrod nodes xgk[2(m-5)]
    double fval = 0.0;
    if (lane < 21) {
        double x = centr;
        if (lane > 0) {
            const int m = (lane - 1) >> 1;
            const int idx = (m < 5) ? (2 * m + 1) : (2 * (m - 5));
            const double absc = hlgth * c_xgk[idx];
            x = ((lane - 1) & 1) ? (centr + absc) : (centr - absc);
        }
        fval = f(x);
    }
    f.advance();  // one Gauss-Kronrod call consumed (no-op for analytic integrands)
    double fv1[10], fv2[10];
    const double fc = __shfl_sync(FULL, fval, 0);
#pragma unroll
    for (int m = 0; m < 10; ++m) {
        const int idx = (m < 5) ? (2 * m + 1) : (2 * (m - 5));
        fv1[idx] = __shfl_sync(FULL, fval, 1 + 2 * m);
        fv2[idx] = __shfl_sync(FULL, fval, 2 + 2 * m);
    }
    double resg = 0.0, resk = c_wgk[10] * fc;
    resabs = fabs(resk);
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int jtw = 2 * j + 1;
        const double fsum = fv1[jtw] + fv2[jtw];
        resg += c_wg[j] * fsum;
        resk += c_wgk[jtw] * fsum;
        resabs += c_wgk[jtw] * (fabs(fv1[jtw]) + fabs(fv2[jtw]));
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int jtwm1 = 2 * j;
        const double fsum = fv1[jtwm1] + fv2[jtwm1];
        resk += c_wgk[jtwm1] * fsum;
        resabs += c_wgk[jtwm1] * (fabs(fv1[jtwm1]) + fabs(fv2[jtwm1]));
    }
    const double reskh = resk * 0.5;
    resasc = c_wgk[10] * fabs(fc - reskh);
#pragma unroll
    for (int j = 0; j < 10; ++j)
        resasc += c_wgk[j] * (fabs(fv1[j] - reskh) + fabs(fv2[j] - reskh));
    result = resk * hlgth;
    resabs *= dhlgth;
    resasc *= dhlgth;
    abserr = fabs((resk - resg) * hlgth);
    if (resasc != 0.0 && abserr != 0.0) {
        const double r = pow15(200.0 * abserr / resasc);
        abserr = resasc * (r < 1.0 ? r : 1.0);
    }
    if (resabs > UFLOW / (50.0 * EPMACH)) {
        const double e = (EPMACH * 50.0) * resabs;
        if (e > abserr) abserr = e;
    }
}

// dqelg: epsilon algorithm.  epstab (52) and res3la (3) live in shared memory; every
// lane executes the same scalar code and writes the same values.
__device__ void dqelg_dev(int& n, double* epstab_, double& result, double& abserr, double* res3la_,
                          int& nres)
{
    double* epstab = epstab_ - 1;
    double* res3la = res3la_ - 1;
    ++nres;
    abserr = OFLOW;
    result = epstab[n];
    if (n >= 3) {
        const int limexp = 50;
        __syncwarp();
        epstab[n + 2] = epstab[n];
        const int newelm = (n - 1) / 2;
        epstab[n] = OFLOW;
        __syncwarp();
        const int num = n;
        int k1 = n;
        int i;
        for (i = 1; i <= newelm; ++i) {
            const int k2 = k1 - 1, k3 = k1 - 2;
            double res = epstab[k1 + 2];
            const double e0 = epstab[k3], e1 = epstab[k2], e2 = res;
            const double e1abs = fabs(e1);
            const double delta2 = e2 - e1, err2 = fabs(delta2);
            const double tol2 = fmax(fabs(e2), e1abs) * EPMACH;
            const double delta3 = e1 - e0, err3 = fabs(delta3);
            const double tol3 = fmax(e1abs, fabs(e0)) * EPMACH;
            if (!(err2 > tol2 || err3 > tol3)) {
                result = res;
                abserr = err2 + err3;
                abserr = fmax(abserr, 5.0 * EPMACH * fabs(result));
                return;
            }
            const double e3 = epstab[k1];
            __syncwarp();
            epstab[k1] = e1;
            __syncwarp();
            const double delta1 = e1 - e3, err1 = fabs(delta1);
            const double tol1 = fmax(e1abs, fabs(e3)) * EPMACH;
            if (err1 <= tol1 || err2 <= tol2 || err3 <= tol3) { n = i + i - 1; break; }
            const double ss = 1.0 / delta1 + 1.0 / delta2 - 1.0 / delta3;
            const double epsinf = fabs(ss * e1);
            if (!(epsinf > 1.0e-4)) { n = i + i - 1; break; }
            res = e1 + 1.0 / ss;
            epstab[k1] = res;
            __syncwarp();
            k1 -= 2;
            const double error = err2 + fabs(res - e2) + err3;
            if (error > abserr) continue;
            abserr = error;
            result = res;
        }
        if (n == limexp) n = 2 * (limexp / 2) - 1;
        int ib = ((num / 2) * 2 == num) ? 2 : 1;
        const int ie = newelm + 1;
        for (i = 1; i <= ie; ++i) {
            const int ib2 = ib + 2;
            const double v = epstab[ib2];
            __syncwarp();
            epstab[ib] = v;
            __syncwarp();
            ib = ib2;
        }
        if (num != n) {
            int indx = num - n + 1;
            for (i = 1; i <= n; ++i) {
                const double v = epstab[indx];
                __syncwarp();
                epstab[i] = v;
                __syncwarp();
                ++indx;
            }
        }
        if (nres < 4) {
            res3la[nres] = result;
            __syncwarp();
            abserr = OFLOW;
        } else {
            abserr = fabs(result - res3la[3]) + fabs(result - res3la[2]) + fabs(result - res3la[1]);
            const double r2 = res3la[2], r3 = res3la[3];
            __syncwarp();
            res3la[1] = r2;
            res3la[2] = r3;
            res3la[3] = result;
            __syncwarp();
        }
    }
    abserr = fmax(abserr, 5.0 * EPMACH * fabs(result));
}

// dqpsrt: maintain the descending ordering of error estimates.  1-based lists.
__device__ void dqpsrt_dev(int limit, int last, int& maxerr, double& ermax, const double* elist_,
                           int* iord_, int& nrmax)
{
    const double* elist = elist_ - 1;
    int* iord = iord_ - 1;
    __syncwarp();
    if (last <= 2) {
        iord[1] = 1;
        iord[2] = 2;
    } else {
        const double errmax = elist[maxerr];
        if (nrmax != 1) {
            const int ido = nrmax - 1;
            for (int i = 1; i <= ido; ++i) {
                const int isucc = iord[nrmax - 1];
                if (errmax <= elist[isucc]) break;
                __syncwarp();
                iord[nrmax] = isucc;
                __syncwarp();
                --nrmax;
            }
        }
        int jupbn = last;
        if (last > (limit / 2 + 2)) jupbn = limit + 3 - last;
        const double errmin = elist[last];
        const int jbnd = jupbn - 1;
        const int ibeg = nrmax + 1;
        int i = ibeg;
        bool found = false;
        if (ibeg <= jbnd) {
            for (i = ibeg; i <= jbnd; ++i) {
                const int isucc = iord[i];
                if (errmax >= elist[isucc]) { found = true; break; }
                __syncwarp();
                iord[i - 1] = isucc;
                __syncwarp();
            }
        }
        if (!found) {
            __syncwarp();
            iord[jbnd] = maxerr;
            iord[jupbn] = last;
        } else {
            __syncwarp();
            iord[i - 1] = maxerr;
            __syncwarp();
            int k = jbnd;
            bool placed = false;
            for (int j = i; j <= jbnd; ++j) {
                const int isucc = iord[k];
                if (errmin < elist[isucc]) {
                    __syncwarp();
                    iord[k + 1] = last;
                    placed = true;
                    break;
                }
                __syncwarp();
                iord[k + 1] = isucc;
                __syncwarp();
                --k;
            }
            if (!placed) {
                __syncwarp();
                iord[i] = last;
            }
        }
    }
    __syncwarp();
    CGB_CHECK(nrmax >= 1 && nrmax <= last);
    maxerr = iord[nrmax];
    CGB_CHECK(maxerr >= 1 && maxerr <= last);
    ermax = elist[maxerr];
}

// Warp-cooperative dqagse.  Every lane executes the control flow; all lanes hold
// identical scalars.  Returns ier; result/abserr/last by reference.
template <class F>
__device__ int dqagse_warp(const F& f, double a, double b, double epsabs, double epsrel, int limit,
                           QagsScratch& S, double& result, double& abserr, int& last_out)
{
    double* alist = S.alist - 1;
    double* blist = S.blist - 1;
    double* rlist = S.rlist - 1;
    double* elist = S.elist - 1;
    int* iord = S.iord - 1;
    double* rlist2 = S.rlist2;
    double area, area1, area12, area2, a1, a2, b1, b2, correc = 0.0, defabs, defab1, defab2, dres,
           erlarg = 0.0, erlast, errbnd, errmax, error1, error2, erro12, errsum, ertest = 0.0,
           resabs, reseps, small = 0.0, abseps;
    int id, ier, ierro, iroff1, iroff2, iroff3, jupbnd, k, ksgn, ktmin, last, maxerr, nres, nrmax,
        numrl2;
    bool extrap, noext;

    ier = 0; last = 0; result = 0.0; abserr = 0.0;
    __syncwarp();
    alist[1] = a; blist[1] = b; rlist[1] = 0.0; elist[1] = 0.0;
    if (epsabs <= 0.0 && epsrel < fmax(50.0 * EPMACH, 0.5e-28)) { last_out = 0; return 6; }

    ierro = 0;
    dqk21_warp(f, a, b, result, abserr, defabs, resabs);
    if (f.exhausted()) { last_out = 0; return QAGS_NEED_VALUES; }
    dres = fabs(result);
    errbnd = fmax(epsabs, epsrel * dres);
    last = 1;
    rlist[1] = result; elist[1] = abserr; iord[1] = 1;
    __syncwarp();
    if (abserr <= 100.0 * EPMACH * defabs && abserr > errbnd) ier = 2;
    if (limit == 1) ier = 1;
    if (ier != 0 || (abserr <= errbnd && abserr != resabs) || abserr == 0.0) {
        last_out = last;
        return ier;
    }

    rlist2[0] = result;
    __syncwarp();
    errmax = abserr; maxerr = 1; area = result; errsum = abserr; abserr = OFLOW;
    nrmax = 1; nres = 0; numrl2 = 2; ktmin = 0; extrap = false; noext = false;
    iroff1 = 0; iroff2 = 0; iroff3 = 0;
    ksgn = -1;
    if (dres >= (1.0 - 50.0 * EPMACH) * defabs) ksgn = 1;

    bool goto115 = false;
    for (last = 2; last <= limit; ++last) {
        CGB_CHECK(limit <= QAGS_LIMIT && maxerr >= 1 && maxerr < last);
        a1 = alist[maxerr]; b1 = 0.5 * (alist[maxerr] + blist[maxerr]); a2 = b1; b2 = blist[maxerr];
        erlast = errmax;
        dqk21_warp(f, a1, b1, area1, error1, resabs, defab1);
        dqk21_warp(f, a2, b2, area2, error2, resabs, defab2);
        if (f.exhausted()) { last_out = last; return QAGS_NEED_VALUES; }
        area12 = area1 + area2; erro12 = error1 + error2;
        errsum = errsum + erro12 - errmax;
        area = area + area12 - rlist[maxerr];
        if (!(defab1 == error1 || defab2 == error2)) {
            if (!(fabs(rlist[maxerr] - area12) > 1.0e-5 * fabs(area12) || erro12 < 0.99 * errmax)) {
                if (extrap) ++iroff2; else ++iroff1;
            }
            if (last > 10 && erro12 > errmax) ++iroff3;
        }
        __syncwarp();
        rlist[maxerr] = area1; rlist[last] = area2;
        errbnd = fmax(epsabs, epsrel * fabs(area));
        if (iroff1 + iroff2 >= 10 || iroff3 >= 20) ier = 2;
        if (iroff2 >= 5) ierro = 3;
        if (last == limit) ier = 1;
        if (fmax(fabs(a1), fabs(b2)) <= (1.0 + 100.0 * EPMACH) * (fabs(a2) + 1000.0 * UFLOW)) ier = 4;
        if (error2 > error1) {
            alist[maxerr] = a2; alist[last] = a1; blist[last] = b1;
            rlist[maxerr] = area2; rlist[last] = area1;
            elist[maxerr] = error2; elist[last] = error1;
        } else {
            alist[last] = a2; blist[maxerr] = b1; blist[last] = b2;
            elist[maxerr] = error1; elist[last] = error2;
        }
        __syncwarp();
        dqpsrt_dev(limit, last, maxerr, errmax, S.elist, S.iord, nrmax);
        if (errsum <= errbnd) { goto115 = true; break; }
        if (ier != 0) break;
        if (last == 2) {
            small = fabs(b - a) * 0.375;
            erlarg = errsum; ertest = errbnd;
            rlist2[1] = area;
            __syncwarp();
            continue;
        }
        if (noext) continue;
        erlarg -= erlast;
        if (fabs(b1 - a1) > small) erlarg += erro12;
        if (!extrap) {
            if (fabs(blist[maxerr] - alist[maxerr]) > small) continue;
            extrap = true;
            nrmax = 2;
        }
        if (ierro != 3 && erlarg > ertest) {
            id = nrmax;
            jupbnd = last;
            if (last > (2 + limit / 2)) jupbnd = limit + 3 - last;
            bool jumped = false;
            for (k = id; k <= jupbnd; ++k) {
                maxerr = iord[nrmax];
                errmax = elist[maxerr];
                if (fabs(blist[maxerr] - alist[maxerr]) > small) { jumped = true; break; }
                ++nrmax;
            }
            if (jumped) continue;
        }
        ++numrl2;
        __syncwarp();
        CGB_CHECK_INDEX(numrl2 - 1, 52);
        rlist2[numrl2 - 1] = area;
        __syncwarp();
        dqelg_dev(numrl2, rlist2, reseps, abseps, S.res3la, nres);
        ++ktmin;
        if (ktmin > 5 && abserr < 1.0e-3 * errsum) ier = 5;
        if (abseps < abserr) {
            ktmin = 0;
            abserr = abseps;
            result = reseps;
            correc = erlarg;
            ertest = fmax(epsabs, epsrel * fabs(reseps));
            if (abserr <= ertest) break;
        }
        if (numrl2 == 1) noext = true;
        if (ier == 5) break;
        maxerr = iord[1];
        errmax = elist[maxerr];
        nrmax = 1;
        extrap = false;
        small *= 0.5;
        erlarg = errsum;
    }
    if (last > limit) last = limit;

    bool goto130 = false;
    if (!goto115) {
        if (abserr == OFLOW) {
            goto115 = true;
        } else {
            if (ier + ierro != 0) {
                if (ierro == 3) abserr += correc;
                if (ier == 0) ier = 3;
                if (result != 0.0 && area != 0.0) {
                    if (abserr / fabs(result) > errsum / fabs(area)) goto115 = true;
                } else {
                    if (abserr > errsum) goto115 = true;
                    else if (area == 0.0) goto130 = true;
                }
            }
            if (!goto115 && !goto130) {
                if (!(ksgn == -1 && fmax(fabs(result), fabs(area)) <= defabs * 0.01)) {
                    if (0.01 > (result / area) || (result / area) > 100.0 || errsum > fabs(area)) ier = 6;
                }
            }
        }
    }
    if (goto115) {
        result = 0.0;
        for (k = 1; k <= last; ++k) result += rlist[k];
        abserr = errsum;
    }
    if (ier > 2) --ier;
    last_out = last;
    return ier;
}

#define EDD_WARPS 8

// g[i] = quad(dgdp, 0, sqrt(ee[i]), epsabs, epsrel, args=ee[i])[0]     (virial.py:54-63)
__global__ void __launch_bounds__(EDD_WARPS * 32)
cgb_eddington_g_kernel(const double* __restrict__ ee, int n, const double* __restrict__ t,
                       const double* __restrict__ wrk, double epsabs, double epsrel,
                       double* __restrict__ g, int* __restrict__ last_out, int* __restrict__ ier_out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int nk = n + 4;
    double* s_t = (double*)smem;
    double* s_w = s_t + nk;
    QagsScratch* scr = (QagsScratch*)(s_w + n);
    for (int i = threadIdx.x; i < nk; i += blockDim.x) s_t[i] = t[i];
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_w[i] = (i < n - 1) ? wrk[i] : 0.0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bin = blockIdx.x * EDD_WARPS + warp;
    if (bin >= n) return;
    DgdpIntegrand f;
    f.t = s_t; f.wrk = s_w; f.nk = nk; f.e = ee[bin];
    double result, abserr;
    int last;
    const int ier = dqagse_warp(f, 0.0, sqrt(ee[bin]), epsabs, epsrel, QAGS_LIMIT, scr[warp], result,
                                abserr, last);
    if (lane == 0) {
        g[bin] = result;
        if (last_out) last_out[bin] = last;
        if (ier_out) ier_out[bin] = ier;
    }
}

// rho[i] = 4 pi quad(rho_int, 0, ee[i], args=(ee[i],))[0]; outputs reversed to radius
// order, chk = (rho - pden) / pden                                   (virial.py:94-100)
__global__ void __launch_bounds__(EDD_WARPS * 32)
cgb_check_virial_kernel(const double* __restrict__ ee, const double* __restrict__ pden, int n,
                        const double* __restrict__ t, const double* __restrict__ c,
                        double* __restrict__ rho_out, double* __restrict__ chk_out,
                        int* __restrict__ last_out, int* __restrict__ ier_out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int nk = n + 4;
    double* s_t = (double*)smem;
    double* s_c = s_t + nk;
    QagsScratch* scr = (QagsScratch*)(s_c + n);
    for (int i = threadIdx.x; i < nk; i += blockDim.x) s_t[i] = t[i];
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_c[i] = c[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bin = blockIdx.x * EDD_WARPS + warp;
    if (bin >= n) return;
    RhoIntegrand f;
    f.t = s_t; f.c = s_c; f.nk = nk; f.psi = ee[bin];
    double result, abserr;
    int last;
    const int ier = dqagse_warp(f, 0.0, ee[bin], 1.49e-8, 1.49e-8, QAGS_LIMIT, scr[warp], result,
                                abserr, last);
    if (lane == 0) {
        const double rho = 4.0 * M_PI * result;
        const double pd = pden[bin];
        rho_out[n - 1 - bin] = rho;
        chk_out[n - 1 - bin] = (rho - pd) / pd;
        if (last_out) last_out[bin] = last;
        if (ier_out) ier_out[bin] = ier;
    }
}

// out[i] = quad(lambda x: s(x) * x**power, a[i], b[i], epsabs, epsrel)[0]
__global__ void __launch_bounds__(EDD_WARPS * 32)
cgb_integrate_spline_kernel(const double* __restrict__ t, const double* __restrict__ c, int n,
                            const double* __restrict__ a, const double* __restrict__ b, int m, int power,
                            double epsabs, double epsrel, double* __restrict__ out,
                            int* __restrict__ last_out, int* __restrict__ ier_out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int nk = n + 4;
    double* s_t = (double*)smem;
    double* s_c = s_t + nk;
    QagsScratch* scr = (QagsScratch*)(s_c + n);
    for (int i = threadIdx.x; i < nk; i += blockDim.x) s_t[i] = t[i];
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_c[i] = c[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bin = blockIdx.x * EDD_WARPS + warp;
    if (bin >= m) return;
    PowSplineIntegrand f;
    f.t = s_t; f.c = s_c; f.nk = nk; f.power = power;
    double result, abserr;
    int last;
    const int ier = dqagse_warp(f, a[bin], b[bin], epsabs, epsrel, QAGS_LIMIT, scr[warp], result, abserr, last);
    if (lane == 0) {
        out[bin] = result;
        if (last_out) last_out[bin] = last;
        if (ier_out) ier_out[bin] = ier;
    }
}

// Integrals of HOST callables (utils.py:129-142: integrate_mass / integrate over a Python
// profile).  The integrand cannot run on the device, the adaptive decisions can: dqagse is
// REPLAYED against a cache of integrand values.  Gauss-Kronrod call number k of an integral
// takes its 21 values from cache[k]; the first call(s) beyond the cache -- the whole interval,
// later the two halves of the bisected interval -- record their abscissae instead and the
// routine stops (QAGS_NEED_VALUES).  The host evaluates the profile on the requested abscissae
// of ALL integrals in one vectorised call, appends the values, and launches the next replay:
// one host call per bisection level (<= ~10) instead of one Python callback per abscissa
// (1000 x ~86).  dqagse is deterministic, so every replay retraces the same decisions.
struct CachedIntegrand {
    const double* cache;  // [QAGS_MAX_CALLS][21] values, evaluation order of dqk21_warp's lanes
    double* xreq;         // [2][21] abscissae of the next call(s)
    int ncached;
    mutable int k;
    __device__ __forceinline__ double operator()(double x) const
    {
        const int lane = threadIdx.x & 31;
        if (k < ncached) return cache[(size_t)k * 21 + lane];
        if (k - ncached < 2) xreq[(k - ncached) * 21 + lane] = x;
        return 0.0;
    }
    __device__ __forceinline__ void advance() const { ++k; }
    __device__ __forceinline__ bool exhausted() const { return k > ncached; }
};
#define QAGS_MAX_CALLS (2 * QAGS_LIMIT - 1)

// status[i]: 0 = running, 1 = done (out / last / ier written).  nreq[i]: abscissa sets requested.
__global__ void __launch_bounds__(EDD_WARPS * 32)
cgb_qags_replay_kernel(const double* __restrict__ a, const double* __restrict__ b, int m, double epsabs,
                       double epsrel, const double* __restrict__ cache, const int* __restrict__ ncached,
                       double* __restrict__ xreq, int* __restrict__ nreq, int* __restrict__ status,
                       double* __restrict__ out, int* __restrict__ last_out, int* __restrict__ ier_out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    QagsScratch* scr = (QagsScratch*)smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bin = blockIdx.x * EDD_WARPS + warp;
    if (bin >= m || status[bin]) return;
    CachedIntegrand f;
    f.cache = cache + (size_t)bin * QAGS_MAX_CALLS * 21;
    f.xreq = xreq + (size_t)bin * 42;
    f.ncached = ncached[bin];
    f.k = 0;
    double result, abserr;
    int last;
    const int ier = dqagse_warp(f, a[bin], b[bin], epsabs, epsrel, QAGS_LIMIT, scr[warp], result, abserr, last);
    __syncwarp();
    if (lane == 0) {
        if (ier == QAGS_NEED_VALUES) {
            nreq[bin] = f.k - f.ncached;
        } else {
            status[bin] = 1;
            nreq[bin] = 0;
            out[bin] = result;
            if (last_out) last_out[bin] = last;
            if (ier_out) ier_out[bin] = ier;
        }
    }
}

// ff = g'(ee) / (sqrt(8) pi^2)                                       (virial.py:66)
__global__ void cgb_ff_kernel(const double* __restrict__ t, const double* __restrict__ wrk, int n,
                              const double* __restrict__ ee, double inv_norm_den,
                              double* __restrict__ ff)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ff[i] = cgb_splder1(t, n + 4, wrk, ee[i]) / inv_norm_den;
}

// One integral, with the full QAGS state dumped: out = [result, abserr, last, ier,
// alist[50], blist[50], rlist[50], elist[50]].  kind 0: dgdp integrand of the density
// spline (t, c) with e = par; kind 1: rho_int integrand of the f spline with psi = par.
__global__ void __launch_bounds__(32)
cgb_qags_trace_kernel(const double* __restrict__ t, const double* __restrict__ c, int n, double par,
                      double a, double b, double epsabs, double epsrel, int kind,
                      double* __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int nk = n + 4;
    double* s_t = (double*)smem;
    double* s_c = s_t + nk;
    QagsScratch* scr = (QagsScratch*)(s_c + n);
    for (int i = threadIdx.x; i < nk; i += 32) s_t[i] = t[i];
    for (int i = threadIdx.x; i < n; i += 32) {
        if (kind == 0) {
            double w = 0.0;
            if (i < n - 1) {
                const double fac = t[i + 4] - t[i + 1];
                w = (fac > 0.0) ? 3.0 * (c[i + 1] - c[i]) / fac : 0.0;
            }
            s_c[i] = w;
        } else {
            s_c[i] = c[i];
        }
    }
    __syncthreads();
    double result, abserr;
    int last, ier;
    if (kind == 0) {
        DgdpIntegrand f;
        f.t = s_t; f.wrk = s_c; f.nk = nk; f.e = par;
        ier = dqagse_warp(f, a, b, epsabs, epsrel, QAGS_LIMIT, scr[0], result, abserr, last);
    } else {
        RhoIntegrand f;
        f.t = s_t; f.c = s_c; f.nk = nk; f.psi = par;
        ier = dqagse_warp(f, a, b, epsabs, epsrel, QAGS_LIMIT, scr[0], result, abserr, last);
    }
    __syncwarp();
    if (threadIdx.x == 0) {
        out[0] = result; out[1] = abserr; out[2] = (double)last; out[3] = (double)ier;
        for (int i = 0; i < QAGS_LIMIT; ++i) {
            out[4 + i] = scr[0].alist[i];
            out[4 + QAGS_LIMIT + i] = scr[0].blist[i];
            out[4 + 2 * QAGS_LIMIT + i] = scr[0].rlist[i];
            out[4 + 3 * QAGS_LIMIT + i] = scr[0].elist[i];
        }
    }
}

extern "C" int cgb_qags_trace(const double* d_t, const double* d_c, int n, double par, double a,
                              double b, double epsabs, double epsrel, int kind, double* d_out,
                              void* stream)
{
    CGB_REQUIRE(n >= 4 && d_t && d_c && d_out, "cgb_qags_trace: bad arguments");
    CGB_REQUIRE(kind == 0 || kind == 1, "cgb_qags_trace: kind must be 0 or 1");
    const size_t smem = (size_t)(2 * n + 4) * sizeof(double) + sizeof(QagsScratch);
    CGB_REQUIRE(smem <= 227 * 1024, "cgb_qags_trace: n = %d needs %zu B of shared memory", n, smem);
    CGB_CUDA_CHECK(cudaFuncSetAttribute(cgb_qags_trace_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cgb_qags_trace_kernel<<<1, 32, smem, (cudaStream_t)stream>>>(d_t, d_c, n, par, a, b, epsabs, epsrel,
                                                                kind, d_out);
    CGB_CUDA_CHECK(cudaGetLastError());
    return CGB_OK;
}

static size_t edd_smem_bytes(int n)
{
    return (size_t)(2 * n + 4) * sizeof(double) + EDD_WARPS * sizeof(QagsScratch);
}

extern "C" size_t cgb_eddington_workspace_bytes(int n)
{
    // t_rho (n+4), c_rho (n), wrk (n), t_g (n+4), c_g (n), wrk_g (n), fit scratch 18n
    // (the three fits share one Givens factorisation of the energy grid)
    return n > 0 ? (size_t)(24 * (size_t)n + 8) * sizeof(double) : 0;
}

extern "C" int cgb_eddington_df(const double* d_ee, const double* d_pden, int n, double epsabs,
                                double epsrel, double* d_g, double* d_ff, double* d_f_t,
                                double* d_f_c, int* d_last, int* d_ier, void* d_work,
                                size_t work_bytes, int fit_method, void* stream)
{
    CGB_REQUIRE(n >= 8, "cgb_eddington_df: need at least 8 radial points, got %d", n);
    // CGB_FIT_REUSE_FACTOR: d_work still holds the Givens factorisation of this d_ee from an
    // earlier CGB_FIT_FITPACK call (the other species of the same model share the energy grid)
    const int reuse_first = (fit_method & CGB_FIT_REUSE_FACTOR) ? 1 : 0;
    fit_method &= ~CGB_FIT_REUSE_FACTOR;
    CGB_REQUIRE(fit_method == CGB_FIT_FITPACK || fit_method == CGB_FIT_PCR,
                "cgb_eddington_df: unknown fit method %d", fit_method);
    CGB_REQUIRE(!reuse_first || fit_method == CGB_FIT_FITPACK,
                "cgb_eddington_df: CGB_FIT_REUSE_FACTOR needs CGB_FIT_FITPACK");
    CGB_REQUIRE(d_ee && d_pden && d_g && d_ff && d_f_t && d_f_c && d_work,
                "cgb_eddington_df: null pointer");
    if (work_bytes < cgb_eddington_workspace_bytes(n)) {
        cgb_set_error("cgb_eddington_df: workspace %zu < %zu bytes", work_bytes,
                      cgb_eddington_workspace_bytes(n));
        return CGB_E_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    // the fit scratch (Givens factorisation, re-usable) comes first, then the per-call arrays
    double* w = (double*)d_work;
    void* fit_scratch = w;          w += 18 * (size_t)n;
    double* t_rho = w;              w += n + 4;
    double* c_rho = w;              w += n;
    double* wrk = w;                w += n;
    double* t_g = w;                w += n + 4;
    double* c_g = w;                w += n;
    double* wrk_g = w;              w += n;
    int rc;
    // density_spline = IUS(ee, pden)                                  (virial.py:45)
    if ((rc = cgb_launch_spline_fit(d_ee, d_pden, n, t_rho, c_rho, fit_scratch, s, fit_method, reuse_first, wrk))) return rc;
    const size_t smem = edd_smem_bytes(n);
    CGB_REQUIRE(smem <= 227 * 1024, "cgb_eddington_df: n = %d needs %zu B of shared memory", n, smem);
    CGB_CUDA_CHECK(cudaFuncSetAttribute(cgb_eddington_g_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cgb_eddington_g_kernel<<<(n + EDD_WARPS - 1) / EDD_WARPS, EDD_WARPS * 32, smem, s>>>(
        d_ee, n, t_rho, wrk, epsabs, epsrel, d_g, d_last, d_ier);
    CGB_CUDA_CHECK(cudaGetLastError());
    // g_spline = IUS(ee, g); ff = g_spline(ee, 1) / (sqrt(8) pi^2)    (virial.py:65-66)
    if ((rc = cgb_launch_spline_fit(d_ee, d_g, n, t_g, c_g, fit_scratch, s, fit_method, fit_method == CGB_FIT_FITPACK, wrk_g))) return rc;
    const double den = sqrt(8.0) * (M_PI * M_PI);
    cgb_ff_kernel<<<(n + 255) / 256, 256, 0, s>>>(t_g, wrk_g, n, d_ee, den, d_ff);
    CGB_CUDA_CHECK(cudaGetLastError());
    // self.f = IUS(ee, ff)                                            (virial.py:67)
    if ((rc = cgb_launch_spline_fit(d_ee, d_ff, n, d_f_t, d_f_c, fit_scratch, s, fit_method, fit_method == CGB_FIT_FITPACK, nullptr))) return rc;
    return CGB_OK;
}

extern "C" int cgb_integrate_spline(const double* d_t, const double* d_c, int n, const double* d_a,
                                    const double* d_b, int m, int power, double epsabs, double epsrel,
                                    double* d_out, int* d_last, int* d_ier, void* stream)
{
    CGB_REQUIRE(n >= 4 && d_t && d_c, "cgb_integrate_spline: bad spline");
    CGB_REQUIRE(power >= 0 && power <= 4, "cgb_integrate_spline: power must be 0..4, got %d", power);
    CGB_REQUIRE(m >= 0, "cgb_integrate_spline: negative count");
    if (m == 0) return CGB_OK;
    CGB_REQUIRE(d_a && d_b && d_out, "cgb_integrate_spline: null pointer");
    const size_t smem = edd_smem_bytes(n);
    CGB_REQUIRE(smem <= 227 * 1024, "cgb_integrate_spline: n = %d needs %zu B of shared memory", n, smem);
    CGB_CUDA_CHECK(cudaFuncSetAttribute(cgb_integrate_spline_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cgb_integrate_spline_kernel<<<(m + EDD_WARPS - 1) / EDD_WARPS, EDD_WARPS * 32, smem, (cudaStream_t)stream>>>(
        d_t, d_c, n, d_a, d_b, m, power, epsabs, epsrel, d_out, d_last, d_ier);
    CGB_CUDA_CHECK(cudaGetLastError());
    return CGB_OK;
}

extern "C" size_t cgb_qags_replay_cache_doubles(int m)
{
    return m > 0 ? (size_t)m * QAGS_MAX_CALLS * 21 : 0;
}

extern "C" int cgb_qags_replay(const double* d_a, const double* d_b, int m, double epsabs, double epsrel,
                               const double* d_cache, const int* d_ncached, double* d_xreq, int* d_nreq,
                               int* d_status, double* d_out, int* d_last, int* d_ier, void* stream)
{
    CGB_REQUIRE(m >= 0, "cgb_qags_replay: negative count");
    if (m == 0) return CGB_OK;
    CGB_REQUIRE(d_a && d_b && d_cache && d_ncached && d_xreq && d_nreq && d_status && d_out,
                "cgb_qags_replay: null pointer");
    const size_t smem = EDD_WARPS * sizeof(QagsScratch);
    CGB_CUDA_CHECK(cudaFuncSetAttribute(cgb_qags_replay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cgb_qags_replay_kernel<<<(m + EDD_WARPS - 1) / EDD_WARPS, EDD_WARPS * 32, smem, (cudaStream_t)stream>>>(
        d_a, d_b, m, epsabs, epsrel, d_cache, d_ncached, d_xreq, d_nreq, d_status, d_out, d_last, d_ier);
    CGB_CUDA_CHECK(cudaGetLastError());
    return CGB_OK;
}

extern "C" int cgb_check_virial(const double* d_ee, const double* d_pden, int n, const double* d_f_t,
                                const double* d_f_c, double* d_rho, double* d_chk, int* d_last,
                                int* d_ier, void* stream)
{
    CGB_REQUIRE(n >= 8, "cgb_check_virial: need at least 8 radial points, got %d", n);
    CGB_REQUIRE(d_ee && d_pden && d_f_t && d_f_c && d_rho && d_chk, "cgb_check_virial: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = edd_smem_bytes(n);
    CGB_REQUIRE(smem <= 227 * 1024, "cgb_check_virial: n = %d needs %zu B of shared memory", n, smem);
    CGB_CUDA_CHECK(cudaFuncSetAttribute(cgb_check_virial_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cgb_check_virial_kernel<<<(n + EDD_WARPS - 1) / EDD_WARPS, EDD_WARPS * 32, smem, s>>>(
        d_ee, d_pden, n, d_f_t, d_f_c, d_rho, d_chk, d_last, d_ier);
    CGB_CUDA_CHECK(cudaGetLastError());
    return CGB_OK;
}
