/*
 * TEST INFRASTRUCTURE ONLY -- plain-C CPU oracle for the cluster_generator
 * particle-sampling hot path.  Never linked into or called by the product.
 *
 * What is restated here, and from where (file:line into the reference tree
 * /root/reference/cluster_generator unless stated otherwise):
 *
 *   - generate_velocities (row V): the Cython source opt/cython_utils.pyx is
 *     ABSENT from the snapshot; restated from its call-site contract
 *     virial.py:191-206 and docs/source/particles.rst:102-105.  PARITY UNPINNED.
 *   - FITPACK splev/splder/fpbspl (third-party, scipy >= 1.11.4 per setup.py:35):
 *     published algorithm (P. Dierckx, "Curve and Surface Fitting with Splines");
 *     call sites virial.py:45,47,65-67,152,191-193.  Pinned against scipy in
 *     tests/test_oracle_cpu.py.
 *   - QUADPACK dqagse/dqk21/dqelg/dqpsrt (third-party via scipy.integrate.quad):
 *     published algorithm (Piessens et al., "QUADPACK", 1983); call sites
 *     virial.py:55-62 (epsabs=epsrel=1.49e-5) and virial.py:99 (defaults 1.49e-8).
 *     Pinned decision-for-decision against scipy's full_output in
 *     tests/test_oracle_cpu.py.
 *   - Philox4x32-10 (Salmon et al., SC'11): the counter-based generator both the
 *     oracle and the CUDA path use for "mode B"; pinned by the Random123 KATs.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (no FMA contraction, so the
 * arithmetic is plain IEEE double in the written order).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

/* ------------------------------------------------------------------------- */
/* Philox4x32-10                                                             */
/* ------------------------------------------------------------------------- */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void cgo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* 52-bit uniform in [0,1): mantissa = (hi:32 | lo>>12), value = [1,2) - 1.
 * Contract shared with the CUDA path (include/cgb200.h). */
static inline double u01_from_words(uint32_t hi, uint32_t lo)
{
    uint64_t bits = 0x3FF0000000000000ull | ((uint64_t)hi << 20) | (uint64_t)(lo >> 12);
    double d;
    memcpy(&d, &bits, 8);
    return d - 1.0;
}

static inline void philox_pair(uint64_t seed, uint64_t id, uint32_t trial, uint32_t stream,
                               double *ua, double *ub)
{
    uint32_t ctr[4] = {(uint32_t)id, (uint32_t)(id >> 32), trial, stream};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t o[4];
    cgo_philox4x32_10(ctr, key, o);
    *ua = u01_from_words(o[0], o[1]);
    *ub = u01_from_words(o[2], o[3]);
}

void cgo_philox_uniforms(uint64_t seed, uint32_t stream, int64_t first_id, int64_t n, double *out)
{
    for (int64_t i = 0; i < n; ++i)
        philox_pair(seed, (uint64_t)(first_id + i), 0u, stream, &out[2 * i], &out[2 * i + 1]);
}

/* ------------------------------------------------------------------------- */
/* FITPACK cubic B-spline evaluation (splev / splder, ext = 0: extrapolate)  */
/* ------------------------------------------------------------------------- */
/* interval search as splev: t[l] <= x < t[l+1], l clamped to [3, n-5] (0-based) */
static inline int64_t find_interval(const double *t, int64_t n, double x)
{
    int64_t lo = 3, hi = n - 4; /* answer l in [lo, hi-1] */
    /* largest l in [3, n-5] with t[l] <= x, else 3 */
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (t[mid] <= x) lo = mid; else hi = mid;
    }
    return lo;
}

/* fpbspl: the k+1 non-zero B-splines of degree k at x, t[l] <= x < t[l+1] */
static inline void fpbspl(const double *t, int k, double x, int64_t l, double *h)
{
    double hh[5];
    h[0] = 1.0;
    for (int j = 1; j <= k; ++j) {
        for (int i = 0; i < j; ++i) hh[i] = h[i];
        h[0] = 0.0;
        for (int i = 1; i <= j; ++i) {
            int64_t li = l + i, lj = li - j;
            if (t[li] == t[lj]) { h[i] = 0.0; continue; }
            double f = hh[i - 1] / (t[li] - t[lj]);
            h[i - 1] = h[i - 1] + f * (t[li] - x);
            h[i] = f * (x - t[lj]);
        }
    }
}

static inline double splev1(const double *t, int64_t n, const double *c, double x)
{
    int64_t l = find_interval(t, n, x);
    double h[5];
    fpbspl(t, 3, x, l, h);
    double sp = 0.0;
    const double *cc = c + (l - 3);
    for (int j = 0; j < 4; ++j) sp += cc[j] * h[j];
    return sp;
}

void cgo_splev(const double *t, int64_t n, const double *c, const double *x, double *y, int64_t m)
{
    for (int64_t i = 0; i < m; ++i) y[i] = splev1(t, n, c, x[i]);
}

/* splder nu=1 for k=3: derivative coefficients d_i = 3 (c[i+1]-c[i])/(t[i+4]-t[i+1]),
 * evaluated as a degree-2 spline on the same knots. wrk has n-5 entries. */
static void splder1_coeffs(const double *t, int64_t n, const double *c, double *wrk)
{
    int64_t nk1 = n - 4;
    for (int64_t i = 0; i < nk1 - 1; ++i) {
        double fac = t[i + 4] - t[i + 1];
        wrk[i] = (fac > 0.0) ? 3.0 * (c[i + 1] - c[i]) / fac : 0.0;
    }
}

static inline double splder1_eval(const double *t, int64_t n, const double *wrk, double x)
{
    int64_t l = find_interval(t, n, x);
    double h[5];
    fpbspl(t, 2, x, l, h);
    double sp = 0.0;
    const double *w = wrk + (l - 3);
    for (int j = 0; j < 3; ++j) sp += w[j] * h[j];
    return sp;
}

void cgo_splder1(const double *t, int64_t n, const double *c, const double *x, double *y, int64_t m)
{
    double *wrk = (double *)malloc(sizeof(double) * (size_t)n);
    splder1_coeffs(t, n, c, wrk);
    for (int64_t i = 0; i < m; ++i) y[i] = splder1_eval(t, n, wrk, x[i]);
    free(wrk);
}

/* Interpolating cubic spline with FITPACK's not-a-knot knot vector
 * t = [x0]*4 + x[2:-2] + [x_{m-1}]*4 (what fpcurf0 produces for s=0, k=3;
 * virial.py:45).  Solves the collocation system (bandwidth kl=ku=2) by
 * banded LU without pivoting (the matrix is totally positive). */
int cgo_spline_fit(const double *x, const double *y, int64_t m, double *t, double *c)
{
    if (m < 4) return -1;
    int64_t n = m + 4;
    for (int i = 0; i < 4; ++i) { t[i] = x[0]; t[n - 1 - i] = x[m - 1]; }
    for (int64_t i = 2; i <= m - 3; ++i) t[i + 2] = x[i];
    /* band storage: a[i][0..4] = columns i-2 .. i+2 */
    double *a = (double *)calloc((size_t)m * 5, sizeof(double));
    double *b = (double *)malloc(sizeof(double) * (size_t)m);
    if (!a || !b) return -2;
    for (int64_t i = 0; i < m; ++i) {
        int64_t l = find_interval(t, n, x[i]);
        double h[5];
        fpbspl(t, 3, x[i], l, h);
        if (i == m - 1) { h[0] = h[1] = h[2] = 0.0; h[3] = 1.0; } /* right end: B_{m-1}=1 */
        for (int j = 0; j < 4; ++j) {
            int64_t col = l - 3 + j;
            int64_t off = col - i + 2;
            if (off < 0 || off > 4) {
                if (h[j] != 0.0) { free(a); free(b); return -3; }
                continue;
            }
            a[i * 5 + off] = h[j];
        }
        b[i] = y[i];
    }
    for (int64_t p = 0; p < m; ++p) {
        double piv = a[p * 5 + 2];
        for (int64_t r = p + 1; r <= p + 2 && r < m; ++r) {
            int64_t off = p - r + 2; /* column p in row r */
            double f = a[r * 5 + off] / piv;
            if (f == 0.0) continue;
            for (int64_t cidx = p; cidx <= p + 2 && cidx < m; ++cidx) {
                int64_t o_r = cidx - r + 2, o_p = cidx - p + 2;
                a[r * 5 + o_r] -= f * a[p * 5 + o_p];
            }
            b[r] -= f * b[p];
        }
    }
    for (int64_t p = m - 1; p >= 0; --p) {
        double s = b[p];
        for (int64_t cidx = p + 1; cidx <= p + 2 && cidx < m; ++cidx)
            s -= a[p * 5 + (cidx - p + 2)] * c[cidx];
        c[p] = s / a[p * 5 + 2];
    }
    free(a); free(b);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* QUADPACK: dqk21, dqelg, dqpsrt, dqagse                                    */
/* ------------------------------------------------------------------------- */
typedef double (*cgo_integrand)(double, void *);

static const double XGK[11] = {
    0.995657163025808080735527280689003, 0.973906528517171720077964012084452,
    0.930157491355708226001207180059508, 0.865063366688984510732096688423493,
    0.780817726586416897063717578345042, 0.679409568299024406234327365114874,
    0.562757134668604683339000099272694, 0.433395394129247190799265943165784,
    0.294392862701460198131126603103866, 0.148874338981631210884826001129720, 0.0};
static const double WGK[11] = {
    0.011694638867371874278064396062192, 0.032558162307964727478818972459390,
    0.054755896574351996031381300244580, 0.075039674810919952767043140916190,
    0.093125454583697605535065465083366, 0.109387158802297641899210590325805,
    0.123491976262065851077958109585166, 0.134709217311473325928054001771707,
    0.142775938577060080797094273138717, 0.147739104901338491374841515972068,
    0.149445554002916905664936468389821};
static const double WG[5] = {
    0.066671344308688137593568809893332, 0.149451349150580593145776339657697,
    0.219086362515982043995534934228163, 0.269266719309996355091226921569469,
    0.295524224714752870173815619188769};

#define EPMACH DBL_EPSILON
#define UFLOW DBL_MIN
#define OFLOW DBL_MAX

static void dqk21(cgo_integrand f, void *ctx, double a, double b, double *result, double *abserr,
                  double *resabs, double *resasc)
{
    double fv1[10], fv2[10];
    double centr = 0.5 * (a + b), hlgth = 0.5 * (b - a), dhlgth = fabs(hlgth);
    double resg = 0.0, fc = f(centr, ctx), resk = WGK[10] * fc;
    *resabs = fabs(resk);
    for (int j = 0; j < 5; ++j) {
        int jtw = 2 * j + 1;
        double absc = hlgth * XGK[jtw];
        double f1 = f(centr - absc, ctx), f2 = f(centr + absc, ctx);
        fv1[jtw] = f1; fv2[jtw] = f2;
        double fsum = f1 + f2;
        resg += WG[j] * fsum;
        resk += WGK[jtw] * fsum;
        *resabs += WGK[jtw] * (fabs(f1) + fabs(f2));
    }
    for (int j = 0; j < 5; ++j) {
        int jtwm1 = 2 * j;
        double absc = hlgth * XGK[jtwm1];
        double f1 = f(centr - absc, ctx), f2 = f(centr + absc, ctx);
        fv1[jtwm1] = f1; fv2[jtwm1] = f2;
        double fsum = f1 + f2;
        resk += WGK[jtwm1] * fsum;
        *resabs += WGK[jtwm1] * (fabs(f1) + fabs(f2));
    }
    double reskh = resk * 0.5;
    *resasc = WGK[10] * fabs(fc - reskh);
    for (int j = 0; j < 10; ++j)
        *resasc += WGK[j] * (fabs(fv1[j] - reskh) + fabs(fv2[j] - reskh));
    *result = resk * hlgth;
    *resabs *= dhlgth;
    *resasc *= dhlgth;
    *abserr = fabs((resk - resg) * hlgth);
    if (*resasc != 0.0 && *abserr != 0.0) {
        double r = pow(200.0 * *abserr / *resasc, 1.5);
        *abserr = *resasc * (r < 1.0 ? r : 1.0);
    }
    if (*resabs > UFLOW / (50.0 * EPMACH)) {
        double e = (EPMACH * 50.0) * *resabs;
        if (e > *abserr) *abserr = e;
    }
}

/* epstab has 52 entries, res3la 3; n, nres in/out. 1-based inside. */
static void dqelg(int *n_io, double *epstab_, double *result, double *abserr, double *res3la_,
                  int *nres)
{
    double *epstab = epstab_ - 1, *res3la = res3la_ - 1;
    int n = *n_io;
    ++*nres;
    *abserr = OFLOW;
    *result = epstab[n];
    if (n < 3) goto L100;
    {
        const int limexp = 50;
        epstab[n + 2] = epstab[n];
        int newelm = (n - 1) / 2;
        epstab[n] = OFLOW;
        int num = n, k1 = n;
        int i;
        for (i = 1; i <= newelm; ++i) {
            int k2 = k1 - 1, k3 = k1 - 2;
            double res = epstab[k1 + 2];
            double e0 = epstab[k3], e1 = epstab[k2], e2 = res;
            double e1abs = fabs(e1);
            double delta2 = e2 - e1, err2 = fabs(delta2);
            double tol2 = fmax(fabs(e2), e1abs) * EPMACH;
            double delta3 = e1 - e0, err3 = fabs(delta3);
            double tol3 = fmax(e1abs, fabs(e0)) * EPMACH;
            if (!(err2 > tol2 || err3 > tol3)) {
                *result = res;
                *abserr = err2 + err3;
                *abserr = fmax(*abserr, 5.0 * EPMACH * fabs(*result));
                *n_io = n;
                return; /* goto 100 with the max already applied (idempotent) */
            }
            double e3 = epstab[k1];
            epstab[k1] = e1;
            double delta1 = e1 - e3, err1 = fabs(delta1);
            double tol1 = fmax(e1abs, fabs(e3)) * EPMACH;
            if (err1 <= tol1 || err2 <= tol2 || err3 <= tol3) { n = i + i - 1; break; }
            double ss = 1.0 / delta1 + 1.0 / delta2 - 1.0 / delta3;
            double epsinf = fabs(ss * e1);
            if (!(epsinf > 1.0e-4)) { n = i + i - 1; break; }
            res = e1 + 1.0 / ss;
            epstab[k1] = res;
            k1 -= 2;
            double error = err2 + fabs(res - e2) + err3;
            if (error > *abserr) continue;
            *abserr = error;
            *result = res;
        }
        /* shift the table */
        if (n == limexp) n = 2 * (limexp / 2) - 1;
        int ib = ((num / 2) * 2 == num) ? 2 : 1;
        int ie = newelm + 1;
        for (i = 1; i <= ie; ++i) {
            int ib2 = ib + 2;
            epstab[ib] = epstab[ib2];
            ib = ib2;
        }
        if (num != n) {
            int indx = num - n + 1;
            for (i = 1; i <= n; ++i) { epstab[i] = epstab[indx]; ++indx; }
        }
        if (*nres < 4) {
            res3la[*nres] = *result;
            *abserr = OFLOW;
        } else {
            *abserr = fabs(*result - res3la[3]) + fabs(*result - res3la[2]) + fabs(*result - res3la[1]);
            res3la[1] = res3la[2];
            res3la[2] = res3la[3];
            res3la[3] = *result;
        }
    }
L100:
    *abserr = fmax(*abserr, 5.0 * EPMACH * fabs(*result));
    *n_io = n;
}

/* all index arrays 1-based inside */
static void dqpsrt(int limit, int last, int *maxerr, double *ermax, const double *elist_, int *iord_,
                   int *nrmax)
{
    const double *elist = elist_ - 1;
    int *iord = iord_ - 1;
    int i, j, k, isucc, jbnd, jupbn, ibeg, ido;
    double errmax, errmin;
    if (last <= 2) { iord[1] = 1; iord[2] = 2; goto L90; }
    errmax = elist[*maxerr];
    if (*nrmax != 1) {
        ido = *nrmax - 1;
        for (i = 1; i <= ido; ++i) {
            isucc = iord[*nrmax - 1];
            if (errmax <= elist[isucc]) break;
            iord[*nrmax] = isucc;
            --*nrmax;
        }
    }
    jupbn = last;
    if (last > (limit / 2 + 2)) jupbn = limit + 3 - last;
    errmin = elist[last];
    jbnd = jupbn - 1;
    ibeg = *nrmax + 1;
    if (ibeg <= jbnd) {
        for (i = ibeg; i <= jbnd; ++i) {
            isucc = iord[i];
            if (errmax >= elist[isucc]) goto L60;
            iord[i - 1] = isucc;
        }
    }
    iord[jbnd] = *maxerr;
    iord[jupbn] = last;
    goto L90;
L60:
    iord[i - 1] = *maxerr;
    k = jbnd;
    for (j = i; j <= jbnd; ++j) {
        isucc = iord[k];
        if (errmin < elist[isucc]) { iord[k + 1] = last; goto L90; }
        iord[k + 1] = isucc;
        --k;
    }
    iord[i] = last;
L90:
    *maxerr = iord[*nrmax];
    *ermax = elist[*maxerr];
}

#define QAGS_LIMIT_MAX 64

static int dqagse(cgo_integrand f, void *ctx, double a, double b, double epsabs, double epsrel,
                  int limit, double *result, double *abserr, int *neval, int *ier_out,
                  double *alist_, double *blist_, double *rlist_, double *elist_, int *iord_,
                  int *last_out)
{
    double *alist = alist_ - 1, *blist = blist_ - 1, *rlist = rlist_ - 1, *elist = elist_ - 1;
    int *iord = iord_ - 1;
    double rlist2[52], res3la[3];
    double area, area1, area12, area2, a1, a2, b1, b2, correc = 0.0, defabs, defab1, defab2, dres,
           erlarg = 0.0, erlast, errbnd, errmax, error1, error2, erro12, errsum, ertest = 0.0,
           resabs, reseps, small = 0.0, abseps;
    int id, ier, ierro, iroff1, iroff2, iroff3, jupbnd, k, ksgn, ktmin, last, maxerr, nres, nrmax,
        numrl2, extrap, noext;

    ier = 0; *neval = 0; last = 0; *result = 0.0; *abserr = 0.0;
    alist[1] = a; blist[1] = b; rlist[1] = 0.0; elist[1] = 0.0;
    if (epsabs <= 0.0 && epsrel < fmax(50.0 * EPMACH, 0.5e-28)) { *ier_out = 6; *last_out = 0; return 6; }

    ierro = 0;
    dqk21(f, ctx, a, b, result, abserr, &defabs, &resabs);
    dres = fabs(*result);
    errbnd = fmax(epsabs, epsrel * dres);
    last = 1; rlist[1] = *result; elist[1] = *abserr; iord[1] = 1;
    if (*abserr <= 100.0 * EPMACH * defabs && *abserr > errbnd) ier = 2;
    if (limit == 1) ier = 1;
    if (ier != 0 || (*abserr <= errbnd && *abserr != resabs) || *abserr == 0.0) goto L140;

    rlist2[0] = *result;
    errmax = *abserr; maxerr = 1; area = *result; errsum = *abserr; *abserr = OFLOW;
    nrmax = 1; nres = 0; numrl2 = 2; ktmin = 0; extrap = 0; noext = 0;
    iroff1 = 0; iroff2 = 0; iroff3 = 0;
    ksgn = -1;
    if (dres >= (1.0 - 50.0 * EPMACH) * defabs) ksgn = 1;

    for (last = 2; last <= limit; ++last) {
        a1 = alist[maxerr]; b1 = 0.5 * (alist[maxerr] + blist[maxerr]); a2 = b1; b2 = blist[maxerr];
        erlast = errmax;
        dqk21(f, ctx, a1, b1, &area1, &error1, &resabs, &defab1);
        dqk21(f, ctx, a2, b2, &area2, &error2, &resabs, &defab2);
        area12 = area1 + area2; erro12 = error1 + error2;
        errsum = errsum + erro12 - errmax;
        area = area + area12 - rlist[maxerr];
        if (!(defab1 == error1 || defab2 == error2)) {
            if (!(fabs(rlist[maxerr] - area12) > 1.0e-5 * fabs(area12) || erro12 < 0.99 * errmax)) {
                if (extrap) ++iroff2; else ++iroff1;
            }
            if (last > 10 && erro12 > errmax) ++iroff3;
        }
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
        dqpsrt(limit, last, &maxerr, &errmax, elist_, iord_, &nrmax);
        if (errsum <= errbnd) goto L115;
        if (ier != 0) break;
        if (last == 2) {
            small = fabs(b - a) * 0.375;
            erlarg = errsum; ertest = errbnd; rlist2[1] = area;
            continue;
        }
        if (noext) continue;
        erlarg -= erlast;
        if (fabs(b1 - a1) > small) erlarg += erro12;
        if (!extrap) {
            if (fabs(blist[maxerr] - alist[maxerr]) > small) continue;
            extrap = 1;
            nrmax = 2;
        }
        if (ierro != 3 && erlarg > ertest) {
            id = nrmax;
            jupbnd = last;
            if (last > (2 + limit / 2)) jupbnd = limit + 3 - last;
            int jumped = 0;
            for (k = id; k <= jupbnd; ++k) {
                maxerr = iord[nrmax];
                errmax = elist[maxerr];
                if (fabs(blist[maxerr] - alist[maxerr]) > small) { jumped = 1; break; }
                ++nrmax;
            }
            if (jumped) continue;
        }
        ++numrl2;
        rlist2[numrl2 - 1] = area;
        dqelg(&numrl2, rlist2, &reseps, &abseps, res3la, &nres);
        ++ktmin;
        if (ktmin > 5 && *abserr < 1.0e-3 * errsum) ier = 5;
        if (abseps < *abserr) {
            ktmin = 0;
            *abserr = abseps;
            *result = reseps;
            correc = erlarg;
            ertest = fmax(epsabs, epsrel * fabs(reseps));
            if (*abserr <= ertest) break;
        }
        if (numrl2 == 1) noext = 1;
        if (ier == 5) break;
        maxerr = iord[1];
        errmax = elist[maxerr];
        nrmax = 1;
        extrap = 0;
        small *= 0.5;
        erlarg = errsum;
    }
    if (last > limit) last = limit; /* loop ran to completion */

    /* L100: set final result and error estimate */
    if (*abserr == OFLOW) goto L115;
    if (ier + ierro != 0) {
        if (ierro == 3) *abserr += correc;
        if (ier == 0) ier = 3;
        if (*result != 0.0 && area != 0.0) {
            if (*abserr / fabs(*result) > errsum / fabs(area)) goto L115;
        } else {
            if (*abserr > errsum) goto L115;
            if (area == 0.0) goto L130;
        }
    }
    /* L110: test on divergence */
    if (ksgn == -1 && fmax(fabs(*result), fabs(area)) <= defabs * 0.01) goto L130;
    if (0.01 > (*result / area) || (*result / area) > 100.0 || errsum > fabs(area)) ier = 6;
    goto L130;
L115:
    *result = 0.0;
    for (k = 1; k <= last; ++k) *result += rlist[k];
    *abserr = errsum;
L130:
    if (ier > 2) --ier;
L140:
    *neval = 42 * last - 21;
    *ier_out = ier;
    *last_out = last;
    return ier;
}

/* ---- integrands ---- */
typedef struct { const double *t; int64_t n; const double *wrk; double e; } dgdp_ctx;

/* virial.py:47  dgdp = lambda t, e: 2 * density_spline(e - t*t, 1) */
static double dgdp_fn(double tt, void *p)
{
    dgdp_ctx *c = (dgdp_ctx *)p;
    return 2.0 * splder1_eval(c->t, c->n, c->wrk, c->e - tt * tt);
}

typedef struct { const double *t; int64_t n; const double *c; double psi; } rho_ctx;

/* virial.py:96  rho_int = lambda e, psi: self.f(e) * np.sqrt(2 * (psi - e)) */
static double rho_int_fn(double e, void *p)
{
    rho_ctx *c = (rho_ctx *)p;
    return splev1(c->t, c->n, c->c, e) * sqrt(2.0 * (c->psi - e));
}

/* generic test hook: integrand kind 0 = dgdp with (t,c) the density spline and par = e;
 * kind 1 = rho_int with (t,c) the f spline and par = psi. Lists must hold `limit` entries. */
int cgo_qags(const double *t, int64_t n, const double *c, double par, double a, double b,
             double eps, int kind, double *result, double *abserr, int *last, int *ier,
             double *alist, double *blist, double *rlist, double *elist)
{
    int iord[QAGS_LIMIT_MAX], neval;
    if (kind == 0) {
        double *wrk = (double *)malloc(sizeof(double) * (size_t)n);
        splder1_coeffs(t, n, c, wrk);
        dgdp_ctx ctx = {t, n, wrk, par};
        dqagse(dgdp_fn, &ctx, a, b, eps, eps, 50, result, abserr, &neval, ier, alist, blist, rlist,
               elist, iord, last);
        free(wrk);
    } else {
        rho_ctx ctx = {t, n, c, par};
        dqagse(rho_int_fn, &ctx, a, b, eps, eps, 50, result, abserr, &neval, ier, alist, blist,
               rlist, elist, iord, last);
    }
    return 0;
}

/* generic-callback version for validating the QAGS restatement on arbitrary integrands */
int cgo_qags_cb(cgo_integrand f, double a, double b, double epsabs, double epsrel, int limit,
                double *result, double *abserr, int *last, int *ier, double *alist, double *blist,
                double *rlist, double *elist, int *iord)
{
    int neval;
    if (limit > QAGS_LIMIT_MAX) return -1;
    dqagse(f, NULL, a, b, epsabs, epsrel, limit, result, abserr, &neval, ier, alist, blist, rlist,
           elist, iord, last);
    return 0;
}

/* virial.py:54-63: g[i] = quad(dgdp, 0, sqrt(ee[i]), epsabs=epsrel=1.49e-5, args=ee[i])[0] */
int cgo_eddington_g(const double *ee, int64_t m, const double *t, const double *c, double epsabs,
                    double epsrel, double *g, int *last_out, int *ier_out)
{
    int64_t n = m + 4;
    double *wrk = (double *)malloc(sizeof(double) * (size_t)n);
    splder1_coeffs(t, n, c, wrk);
    for (int64_t i = 0; i < m; ++i) {
        double al[QAGS_LIMIT_MAX], bl[QAGS_LIMIT_MAX], rl[QAGS_LIMIT_MAX], el[QAGS_LIMIT_MAX];
        int iord[QAGS_LIMIT_MAX], neval, ier, last;
        double res, err;
        dgdp_ctx ctx = {t, n, wrk, ee[i]};
        dqagse(dgdp_fn, &ctx, 0.0, sqrt(ee[i]), epsabs, epsrel, 50, &res, &err, &neval, &ier, al, bl,
               rl, el, iord, &last);
        g[i] = res;
        if (last_out) last_out[i] = last;
        if (ier_out) ier_out[i] = ier;
    }
    free(wrk);
    return 0;
}

/* virial.py:94-99: rho[i] = 4 pi quad(rho_int, 0, ee[i], args=(ee[i],))[0] */
int cgo_check_virial_rho(const double *ee, int64_t m, const double *t, const double *c, double *rho,
                         int *last_out, int *ier_out)
{
    int64_t n = m + 4;
    for (int64_t i = 0; i < m; ++i) {
        double al[QAGS_LIMIT_MAX], bl[QAGS_LIMIT_MAX], rl[QAGS_LIMIT_MAX], el[QAGS_LIMIT_MAX];
        int iord[QAGS_LIMIT_MAX], neval, ier, last;
        double res, err;
        rho_ctx ctx = {t, n, c, ee[i]};
        dqagse(rho_int_fn, &ctx, 0.0, ee[i], 1.49e-8, 1.49e-8, 50, &res, &err, &neval, &ier, al, bl,
               rl, el, iord, &last);
        rho[i] = 4.0 * M_PI * res;
        if (last_out) last_out[i] = last;
        if (ier_out) ier_out[i] = ier;
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Row V: acceptance-rejection speed draw                                    */
/* ------------------------------------------------------------------------- */
static double *full_knots(const double *knots, int64_t nknots, int k, int64_t *n_out)
{
    /* get_knots() == t[k:-k]; rebuild t by padding k copies at each end */
    int64_t n = nknots + 2 * k;
    double *t = (double *)malloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < k; ++i) { t[i] = knots[0]; t[n - 1 - i] = knots[nknots - 1]; }
    memcpy(t + k, knots, sizeof(double) * (size_t)nknots);
    *n_out = n;
    return t;
}

/* one trial; returns 1 on accept and stores v^2 */
static inline int trial(const double *t, int64_t n, const double *c, double psi, double vesc,
                        double fv2esc, double u1, double u2, double *v2_out)
{
    double v = u1 * vesc;
    double v2 = v * v;
    double e = psi - 0.5 * v2;
    double F = splev1(t, n, c, e);
    *v2_out = v2;
    return !(F * v2 < u2 * fv2esc);
}

#define STREAM_VEL_TRIAL 2u
#define STREAM_SEQ 15u

/* The trial stream spends one Philox4x32-10 call on TWO trials: trial 2p uses
 * words (w0,w1), trial 2p+1 words (w2,w3); u = w * 2^-32 (exact in double).
 * Contract shared with the CUDA path (include/cgb200.h). */
static inline void trial_uniforms(uint64_t seed, uint64_t id, uint64_t trial_idx, uint32_t stream,
                                  double *u1, double *u2)
{
    uint32_t ctr[4] = {(uint32_t)id, (uint32_t)(id >> 32), (uint32_t)(trial_idx >> 1), stream};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t o[4];
    cgo_philox4x32_10(ctr, key, o);
    int odd = (int)(trial_idx & 1);
    *u1 = (double)o[2 * odd] * 0x1.0p-32;
    *u2 = (double)o[2 * odd + 1] * 0x1.0p-32;
}

/* Mode A: one private sequential stream for all particles. */
int cgo_generate_velocities_seq(const double *psi, const double *vesc, const double *fv2esc,
                                const double *knots, int64_t nknots, const double *coeffs,
                                int64_t ncoeffs, int k, uint64_t seed, double *vel, int64_t *trials,
                                int64_t npart)
{
    if (k != 3 || ncoeffs != nknots + 2) return -1;
    int64_t n;
    double *t = full_knots(knots, nknots, k, &n);
    uint64_t draw = 0;
    for (int64_t i = 0; i < npart; ++i) {
        int64_t nt = 0;
        double v2;
        for (;;) {
            double u1, u2;
            /* the stream position is the "id": (draw >> 33 .. ) never overflows in practice */
            trial_uniforms(seed, draw >> 33, draw & 0x1FFFFFFFFull, STREAM_SEQ, &u1, &u2);
            ++draw; ++nt;
            if (trial(t, n, coeffs, psi[i], vesc[i], fv2esc[i], u1, u2, &v2)) break;
        }
        vel[i] = sqrt(v2);
        if (trials) trials[i] = nt;
    }
    free(t);
    return 0;
}

/* Mode B: uniforms are a pure function of (particle id, trial). */
int cgo_generate_velocities_counter(const double *psi, const double *vesc, const double *fv2esc,
                                    const double *knots, int64_t nknots, const double *coeffs,
                                    int64_t ncoeffs, int k, uint64_t seed, int64_t first_id,
                                    double *vel, int64_t *trials, int64_t npart)
{
    if (k != 3 || ncoeffs != nknots + 2) return -1;
    int64_t n;
    double *t = full_knots(knots, nknots, k, &n);
    for (int64_t i = 0; i < npart; ++i) {
        uint64_t tr = 0;
        double v2 = 0.0;
        int ok = 0;
        for (;;) {
            double u1, u2;
            trial_uniforms(seed, (uint64_t)(first_id + i), tr, STREAM_VEL_TRIAL, &u1, &u2);
            ok = trial(t, n, coeffs, psi[i], vesc[i], fv2esc[i], u1, u2, &v2);
            if (ok || tr == 0x1FFFFFFFFull) break;
            ++tr;
        }
        vel[i] = ok ? sqrt(v2) : NAN;
        if (trials) trials[i] = ok ? (int64_t)tr + 1 : -1;
    }
    free(t);
    return 0;
}

/* Mode B, literal host-supplied uniforms u[npart][tcap][2]. */
int cgo_generate_velocities_array(const double *psi, const double *vesc, const double *fv2esc,
                                  const double *knots, int64_t nknots, const double *coeffs,
                                  int64_t ncoeffs, int k, const double *u, int64_t tcap, double *vel,
                                  int64_t *trials, int64_t npart)
{
    if (k != 3 || ncoeffs != nknots + 2) return -1;
    int64_t n;
    double *t = full_knots(knots, nknots, k, &n);
    for (int64_t i = 0; i < npart; ++i) {
        double v2 = 0.0;
        int64_t acc = -1;
        for (int64_t tr = 0; tr < tcap; ++tr) {
            const double *uu = u + ((size_t)i * (size_t)tcap + (size_t)tr) * 2;
            if (trial(t, n, coeffs, psi[i], vesc[i], fv2esc[i], uu[0], uu[1], &v2)) { acc = tr + 1; break; }
        }
        vel[i] = acc > 0 ? sqrt(v2) : NAN;
        if (trials) trials[i] = acc;
    }
    free(t);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* FITPACK fpcurf for the interpolating case (s = 0, k = 3, unit weights):   */
/* observation rows rotated into an upper-triangular band matrix by Givens   */
/* transformations (fpgivs / fprota), then back substitution (fpback).       */
/* Restated from P. Dierckx's published algorithm; pinned bit-for-bit        */
/* against scipy's InterpolatedUnivariateSpline in tests/test_oracle_cpu.py. */
/* ------------------------------------------------------------------------- */
int cgo_spline_fit_fitpack(const double *x, const double *y, int64_t m, double *t, double *c)
{
    if (m < 4) return -1;
    const int k = 3, k1 = 4;
    int64_t n = m + 4, nk1 = n - k1;
    for (int i = 0; i < 4; ++i) { t[i] = x[0]; t[n - 1 - i] = x[m - 1]; }
    for (int64_t i = 2; i <= m - 3; ++i) t[i + 2] = x[i];
    double *a = (double *)calloc((size_t)nk1 * k1, sizeof(double)); /* a[j][0..3] */
    double *z = (double *)calloc((size_t)nk1, sizeof(double));
    if (!a || !z) return -2;
    int64_t l = k1; /* 1-based knot interval index as in fpcurf */
    for (int64_t it = 1; it <= m; ++it) {
        double xi = x[it - 1], yi = y[it - 1], h[5];
        /* search for knot interval t(l) <= xi < t(l+1) */
        while (!(xi < t[l] || l == nk1)) ++l; /* t(l+1) in 1-based == t[l] 0-based */
        fpbspl(t, k, xi, l - 1, h);          /* 0-based interval index l-1 */
        int64_t j = l - k1;                  /* 1-based row j+1 .. */
        for (int i = 1; i <= k1; ++i) {
            ++j;
            double piv = h[i - 1];
            if (piv == 0.0) continue;
            double *aj = a + (size_t)(j - 1) * k1;
            /* fpgivs */
            double ww = aj[0], store = fabs(piv), dd, co, si;
            if (store >= ww) dd = store * sqrt(1.0 + (ww / piv) * (ww / piv));
            else dd = ww * sqrt(1.0 + (piv / ww) * (piv / ww));
            co = ww / dd; si = piv / dd; aj[0] = dd;
            /* fprota on (yi, z(j)) */
            { double s1 = yi, s2 = z[j - 1]; z[j - 1] = co * s2 + si * s1; yi = co * s1 - si * s2; }
            if (i == k1) break;
            int i2 = 0;
            for (int i1 = i + 1; i1 <= k1; ++i1) {
                ++i2;
                double s1 = h[i1 - 1], s2 = aj[i2];
                aj[i2] = co * s2 + si * s1;
                h[i1 - 1] = co * s1 - si * s2;
            }
        }
    }
    /* fpback */
    c[nk1 - 1] = z[nk1 - 1] / a[(size_t)(nk1 - 1) * k1];
    for (int64_t i = nk1 - 1; i >= 1; --i) {
        double store = z[i - 1];
        int64_t i1 = (nk1 - i < k1 - 1) ? nk1 - i : k1 - 1;
        int64_t mm = i;
        for (int64_t ll = 1; ll <= i1; ++ll) { ++mm; store -= c[mm - 1] * a[(size_t)(i - 1) * k1 + ll]; }
        c[i - 1] = store / a[(size_t)(i - 1) * k1];
    }
    free(a); free(z);
    return 0;
}
