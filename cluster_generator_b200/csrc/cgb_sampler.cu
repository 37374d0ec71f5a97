// R+E+V+P fused particle sampler and its table preparation.
//
//   replaces VirialEquilibrium.generate_particles (virial.py:108-236):
//     R  inverse-CDF radius draw           generate_particle_radii (utils.py:154-166)
//     E  psi(r), envelope                  virial.py:152-154,191-194
//     V  acceptance-rejection speed draw   opt/cython_utils.pyx::generate_velocities
//                                          (source absent; call site virial.py:198-206)
//     P  isotropic directions, SoA stores  virial.py:175-187,213-227,229-234
//   and ClusterModel.generate_gas_particles / generate_tracer_particles
//   (model.py:838-947, 766-836) for the position-only species.
//
// Design (B200, 148 SMs, one persistent 512-thread CTA per SM):
//   * every quantity of particle `id` is a pure function of (tables, seed, id): Philox
//     counter = (id, trial pair), so results do not depend on scheduling or on how
//     the index range is sharded over GPUs.
//   * each WARP owns two rolling buffers of TILE consecutive ids, claimed from a
//     global atomic counter.  Phase 0 (uniform): radius, psi, envelope for a whole
//     buffer into shared memory, two particles per lane at a time so that their table walks
//     (dependent L1/L2 loads) overlap.  Phase 1 (divergent): one particle per lane, lanes
//     refill from the buffers as they accept (ballot + popc), two trials per Philox
//     call; the next buffer is filled while stragglers of the previous one finish.
//     A particle that exceeds LANE_CAP_PAIRS trial pairs, or is left over when the
//     index range runs dry, is finished by the whole warp cooperatively (64 trials
//     per step, first accepted trial wins); one that exceeds HARD_PAIRS pairs is
//     deferred to cgb_hard_kernel, which spreads its trials over the whole grid.
//     Phase 2 (uniform): directions, spherical->Cartesian, coalesced plane stores.
//   * the accept test f(psi - v^2/2) v^2 >= u2 * fv2esc is first tried on a
//     conservative fp32 interval screen: a 4096-bucket table of rigorous lower/upper
//     bounds of f over each energy bucket, staged into shared memory by one TMA bulk
//     copy per CTA (one 8-byte load per trial).  Only when the screen cannot decide
//     (~4e-4 of trials) is the exact fp64 piecewise-polynomial spline evaluated, so
//     accept/reject decisions are identical to the exact evaluation by construction.
#include <math.h>
#include <float.h>
#include "cgb_common.cuh"
#include "cgb_spline.cuh"

#define NB 4096            // screen buckets
// Phase-1 shape, measured on B200 (1e7 Vikhlinin / Hernquist / NFW DM, 2e6 BCG stars, ms):
//   1 pair/iteration 0.923 1.696 1.513 2.705 | 2: 0.882 1.551 1.373 2.640 | 3: 0.884 1.527 1.332 2.526
//   3 + all screens up front 0.868 1.543 1.310 2.471 | 4: 0.906 1.620 1.344 2.508 | 4 + all 0.938 1.718 1.401 2.510
#ifndef LEAN_PAIRS
#define LEAN_PAIRS 3       // trial pairs (independent Philox chains) per lane per phase-1 iteration
#endif
#ifndef LEAN_SCREEN_ALL
#define LEAN_SCREEN_ALL 1  // screen all 2*LEAN_PAIRS trials before looking at any result
#endif
#ifndef PH0_Q
#define PH0_Q 2            // particles per lane whose phase-0 table walks are interleaved
#endif
#ifndef WARPS
#define WARPS 16           // warps per CTA
#endif
#ifndef TILE
#define TILE 128           // particles per warp tile buffer
#endif
#ifndef NBUF
#define NBUF 2             // tile buffers per warp (ring): a straggler may keep its lane for
#endif                     // NBUF-1 tile lifetimes before the warp has to finish it.  Measured
                           // (1e7 Vikhlinin DM): 2x128 1.06 ms, 3x96 1.13, 4x64 1.15, 4x96 1.19
// Shape sweep on the round-2 code (1e7 Vikhlinin / Hernquist, ms; shipped 16 warps x 2x128: 0.871 / 1.456):
//   2x96 1.011 / 1.623 | 2x160 1.021 / 1.652 | 14 warps 2x192 1.217 / 1.929 | 12 warps 2x224 1.079 / 1.694
//   20 warps at 96 registers: 2x96 1.045 / 1.591, 2x128 1.056 / 1.657 | 24 warps at 80 registers, 2x64: 1.048 / 1.735
// (the shared-memory carve-out is quantised at 132/164/196/228 KB: every other shape either
// shrinks L1 under the 116 KB of phase-0 tables or the tile under the stragglers)
#define NBC 1024           // buckets of the CDF start-index table
#define NBG 2048           // finer start-index table of the position-only (gas) fast path
#define GAS_A_STRIDE 6      // doubles per record A: 48 B = 3 x 16 B, so consecutive records rotate through
                           // all eight 16-byte bank groups (a 32-byte stride uses only four of them)
#ifndef LANE_CAP_PAIRS
#define LANE_CAP_PAIRS 384 // trial pairs a particle may spend in its lane before the warp takes over (48/96/192/384: Hernquist 1.47/1.56/1.45/1.45 ms, stars 2.57/2.55/2.43/2.41)
#endif
#define HARD_PAIRS 32768   // pair index beyond which a particle is deferred to the grid-wide pass
#define HARD_CAP 4096      // capacity of the deferred-particle list
#define TRI_DEFERRED (-2147483647)
#define SCREEN_BYTES ((NB + 2) * 8)  // screen table + sentinel, 16-byte multiple
#define SCREEN_REL 4.0e-6  // relative widening of the screen bounds (covers fp32 rounding)
#define SCREEN_EDGE 0.02   // bucket range widening, in bucket widths

// ---------------------------------------------------------------------------
// workspace layout: header + tables.  Offsets are in doubles from the base.
// ---------------------------------------------------------------------------
struct SamplerHeader {
    int n_cdf, n_psi, n_f, n_aux, n_aux_coef, pad0;
    long long off_cdf_P, off_cdf_r, off_cdf_s;
    long long off_psi_brk, off_psi_pp;
    long long off_f_brk, off_f_pp;
    long long off_screen;
    long long off_cdf_lut;  // unsigned[NBC]: lo | hi << 16 search range per u bucket
    long long off_cdf_iv;   // int2[n_cdf]: starting psi / f piece for CDF interval j
    long long off_aux_brk[2], off_aux_pp[2];
    // position-only species (gas / tracer): packed tables of the fast path
    // (one contiguous block, staged into shared memory by ONE bulk copy)
    long long off_gas_lut;  // unsigned[NBG]: lo | hi << 16 search range per (finer) u bucket
    long long off_gas_P;    // double[n_cdf]: copy of the CDF for the search
    long long off_gas_A;    // double[6][n_cdf]: P_j, r_j, slope_j, r_{j+1} (NaN: irregular interval), 2 pad
    long long off_gas_B;    // double[10][n_psi - 3]: brk_lo, brk_hi, aux0[4], aux1[4]
    long long gas_block_doubles;  // length of that block
    int gas_fast_ok;        // set to 0 by the prepare kernels when the packing does not apply
    int pad2;
    long long total;
    // scalars computed on the device by the prepare kernels
    double e_hi;      // upper end of the screen grid
    double inv_w;     // NB / e_hi
    double f_scale;   // power of two applied to f in the screen table
    double pad1;
};

static SamplerHeader make_layout(const cgb_sampler_desc* d)
{
    SamplerHeader h;
    memset(&h, 0, sizeof(h));
    h.n_cdf = d->n_cdf; h.n_psi = d->n_psi; h.n_f = d->d_f_t ? d->n_f : 0;
    h.n_aux = d->n_aux; h.n_aux_coef = d->n_aux_coef;
    long long o = (long long)cgb_align_up(sizeof(SamplerHeader), 256) / 8;
    auto take = [&](long long n) { long long r = o; o += (n + 31) / 32 * 32; return r; };
    h.off_cdf_P = take(h.n_cdf); h.off_cdf_r = take(h.n_cdf); h.off_cdf_s = take(h.n_cdf);
    h.off_psi_brk = take(h.n_psi - 2); h.off_psi_pp = take(4LL * (h.n_psi - 3));
    h.off_cdf_lut = take(NBC / 2); h.off_cdf_iv = take(h.n_cdf);
    if (h.n_f) {
        h.off_f_brk = take(h.n_f - 2); h.off_f_pp = take(4LL * (h.n_f - 3));
        h.off_screen = take(NB);  // float2 per bucket = one double slot
    }
    for (int k = 0; k < h.n_aux; ++k) {
        h.off_aux_brk[k] = take(h.n_aux_coef - 2);
        h.off_aux_pp[k] = take(4LL * (h.n_aux_coef - 3));
    }
    if (!h.n_f) {
        h.off_gas_lut = take(NBG / 2);
        h.off_gas_P = take(h.n_cdf);
        h.off_gas_A = take((long long)GAS_A_STRIDE * h.n_cdf);
        h.off_gas_B = take(10LL * (h.n_psi - 3));
        h.gas_block_doubles = o - h.off_gas_lut;
        h.gas_fast_ok = 1;
    }
    h.total = o;
    return h;
}

// per-call scratch (caller-provided, cgb_sample_scratch_bytes()): dynamic tile counter
// and the list of particles whose trial count is so large that the whole GRID finishes
// them in a second pass (the trial-count distribution has a ~1e-7 tail of >1e6 trials)
struct HardEntry {
    unsigned long long id;         // global particle id (Philox counter)
    long long loc;                 // local output index
    double psi, vesc, env;
    unsigned long long next_pair;  // first pair index not yet examined
    unsigned long long best;       // smallest accepted trial index found so far (~0: none)
    unsigned long long pad;
};
struct ScratchHeader {
    unsigned long long tile_counter;
    unsigned int hard_count;
    unsigned int done_ctas;
    unsigned long long pad[6];
};
static_assert(sizeof(HardEntry) == 64 && sizeof(ScratchHeader) == 64, "scratch layout");

// ---------------------------------------------------------------------------
// piecewise-polynomial evaluation (exact path)
// ---------------------------------------------------------------------------
// largest j in [0, nint-1] with brk[j] <= x (else 0); brk has nint+1 entries
__device__ __forceinline__ int pp_find(const double* __restrict__ brk, int nint, double x)
{
    int lo = 0, hi = nint;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(brk + mid) <= x) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ double pp_eval_at(const double* __restrict__ brk,
                                             const double* __restrict__ pp, int j, double x)
{
    CGB_CHECK(j >= 0);
    const double dx = x - __ldg(brk + j);
    const double2 a01 = __ldg((const double2*)(pp + 4 * (size_t)j));
    const double2 a23 = __ldg((const double2*)(pp + 4 * (size_t)j) + 1);
    return fma(fma(fma(a23.y, dx, a23.x), dx, a01.y), dx, a01.x);
}

__device__ __forceinline__ double pp_eval(const double* __restrict__ brk,
                                          const double* __restrict__ pp, int nint, double x)
{
    return pp_eval_at(brk, pp, pp_find(brk, nint, x), x);
}

// start at piece j0 and walk to the piece that holds x (same rule as pp_find)
__device__ __forceinline__ int pp_walk(const double* __restrict__ brk, int nint, int j0, double x)
{
    int j = min(max(j0, 0), nint - 1);
    while (j > 0 && x < __ldg(brk + j)) --j;
    while (j < nint - 1 && x >= __ldg(brk + j + 1)) ++j;
    return j;
}

// pp_walk for Q independent arguments at once: the breakpoint loads of the Q walks are issued
// together instead of one dependent chain after the other (phase 0 of the sampler is bound by
// exactly this load latency).  Same fixed point as pp_walk: breakpoints are strictly
// increasing, so the piece with brk[j] <= x < brk[j+1] (clamped at the ends) is unique.
template <int Q>
__device__ __forceinline__ void pp_walk_q(const double* __restrict__ brk, int nint, int (&j)[Q],
                                          const double (&x)[Q])
{
#pragma unroll
    for (int q = 0; q < Q; ++q) j[q] = min(max(j[q], 0), nint - 1);
    bool moved;
    do {
        moved = false;
        double b0[Q], b1[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) { b0[q] = __ldg(brk + j[q]); b1[q] = __ldg(brk + j[q] + 1); }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            if (j[q] > 0 && x[q] < b0[q]) { --j[q]; moved = true; }
            else if (j[q] < nint - 1 && x[q] >= b1[q]) { ++j[q]; moved = true; }
        }
    } while (moved);
}

// monotone bucket of a uniform in [0,1); the SAME function builds the table, so the
// search range it yields always contains the answer
template <int NBK = NBC>
__device__ __forceinline__ int cdf_bucket(double u)
{
    return min(max(__double2int_rd(__dmul_rn(u, (double)NBK)), 0), NBK - 1);
}

// ---------------------------------------------------------------------------
// prepare kernels
// ---------------------------------------------------------------------------
// cdf_lut[b] = lo | hi << 16 with lo = #{j >= 1 : bucket(P[j]) < b}, hi = #{j >= 1 :
// bucket(P[j]) <= b}: the largest j with P[j] <= u lies in [lo, hi] for bucket(u) = b.
template <int NBK>
__global__ void prep_cdf_lut_kernel(const double* __restrict__ P, int n, unsigned* __restrict__ lut)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= NBK) return;
    // P is non-decreasing: count by binary search on the bucket value
    auto count_lt = [&](int bb) {  // #{j in [1, n-1] : bucket(P[j]) < bb}
        int lo = 1, hi = n;        // first j in [1,n) with bucket(P[j]) >= bb
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (cdf_bucket<NBK>(P[mid]) < bb) lo = mid + 1; else hi = mid;
        }
        return lo - 1;
    };
    const int lo = count_lt(b), hi = min(count_lt(b + 1), n - 2);
    lut[b] = (unsigned)min(lo, n - 2) | ((unsigned)hi << 16);
}

// cdf_iv[j] = (psi piece, f piece) that hold the middle of CDF interval j
__global__ void prep_cdf_iv_kernel(const SamplerHeader* __restrict__ hdr, double* __restrict__ base)
{
    const SamplerHeader h = *hdr;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= h.n_cdf) return;
    const double* rc = base + h.off_cdf_r;
    const double rm = (j + 1 < h.n_cdf) ? 0.5 * (rc[j] + rc[j + 1]) : rc[j];
    const int ip = pp_find(base + h.off_psi_brk, h.n_psi - 3, rm);
    int jf = 0;
    if (h.n_f) {
        const double psi = pp_eval_at(base + h.off_psi_brk, base + h.off_psi_pp, ip, rm);
        jf = pp_find(base + h.off_f_brk, h.n_f - 3, psi);
    }
    ((int2*)(base + h.off_cdf_iv))[j] = make_int2(ip, jf);
}

// Packed tables of the position-only fast path.  All radial splines of a gas / tracer sampler
// are fitted over the model's radius grid and the CDF's r-grid is that grid (plus r = 0), so a
// CDF interval lies inside ONE spline piece and the aux / psi splines share their breakpoints:
// record A[j] carries np.interp's (P_j, r_j, slope_j) and r_{j+1}, record B[piece] the
// breakpoints and the coefficients of every spline -- two table reads per particle instead of
// a dozen.  For a REGULAR interval -- its spline piece is j - 2 with breakpoints exactly
// (r_j, r_{j+1}), true of all but the first three and the last two -- the kernel needs
// nothing but record A to place r in its piece (r < r_{j+1}: piece j - 2, dx = r - r_j);
// the others carry NaN there and take the breakpoint walk.  Where a caller's tables do not
// have that structure at all gas_fast_ok is cleared and the kernel takes the general path.
__global__ void prep_gas_pack_kernel(SamplerHeader* __restrict__ hdr, double* __restrict__ base)
{
    const SamplerHeader h = *hdr;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int nint = h.n_psi - 3;
    const double* pb = base + h.off_psi_brk;
    bool ok = true;
    if (h.n_aux > 0 && h.n_aux_coef != h.n_psi) ok = false;
    if (t < h.n_cdf) {
        const double* rc = base + h.off_cdf_r;
        const double r0 = rc[t], r1 = (t + 1 < h.n_cdf) ? rc[t + 1] : rc[t];
        const int pj = pp_find(pb, nint, r0);
        if (!(pb[pj] <= r0 || pj == 0) || !(r1 <= pb[pj + 1] || pj == nint - 1)) ok = false;
        const bool regular = t + 1 < h.n_cdf && pj == t - 2 && pb[pj] == r0 && pb[pj + 1] == r1 && r1 > r0;
        double* A = base + h.off_gas_A + (long long)GAS_A_STRIDE * t;
        A[0] = base[h.off_cdf_P + t]; A[1] = r0; A[2] = base[h.off_cdf_s + t];
        A[3] = regular ? r1 : __longlong_as_double(0x7ff8000000000000LL);
        A[4] = 0.0; A[5] = 0.0;
        base[h.off_gas_P + t] = A[0];
    }
    if (t < nint) {
        double* B = base + h.off_gas_B + 10LL * t;
        B[0] = pb[t]; B[1] = pb[t + 1];
        for (int k = 0; k < 2; ++k) {
            for (int c = 0; c < 4; ++c)
                B[2 + 4 * k + c] = (k < h.n_aux && ok) ? base[h.off_aux_pp[k] + 4LL * t + c] : 0.0;
            if (k < h.n_aux && ok && (base[h.off_aux_brk[k] + t] != pb[t] || base[h.off_aux_brk[k] + t + 1] != pb[t + 1]))
                ok = false;
        }
    }
    if (!ok) atomicAnd(&hdr->gas_fast_ok, 0);
}

__global__ void prep_cdf_kernel(const double* __restrict__ P, const double* __restrict__ r, int n,
                                double* __restrict__ oP, double* __restrict__ oR,
                                double* __restrict__ oS)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    oP[j] = P[j];
    oR[j] = r[j];
    // np.interp: slope = (fp[j+1]-fp[j]) / (xp[j+1]-xp[j])                (utils.py:165)
    oS[j] = (j + 1 < n) ? __ddiv_rn(__dsub_rn(r[j + 1], r[j]), __dsub_rn(P[j + 1], P[j])) : 0.0;
}

// e_hi = the largest energy any trial can see: max(f knots' upper end, psi(r) over
// the reachable radii incl. the extrapolated core r in [0, r_0)), padded.
__global__ void __launch_bounds__(256) prep_scalars_kernel(SamplerHeader* __restrict__ hdr,
                                                           double* __restrict__ base)
{
    __shared__ double s_e[256], s_f[256];
    const SamplerHeader h = *hdr;
    const double* fb = base + h.off_f_brk;
    const double* fp = base + h.off_f_pp;
    const double* pb = base + h.off_psi_brk;
    const double* pp = base + h.off_psi_pp;
    const int nfi = h.n_f - 3, npi = h.n_psi - 3;
    double e_hi = fb[nfi], fmaxabs = 0.0;
    // psi over [0, r_max]: sample knots and the extrapolated core
    for (int j = threadIdx.x; j <= npi; j += blockDim.x) e_hi = fmax(e_hi, pp_eval(pb, pp, npi, pb[j]));
    for (int k = threadIdx.x; k <= 64; k += blockDim.x)
        e_hi = fmax(e_hi, pp_eval(pb, pp, npi, pb[0] * (double)k / 64.0));
    for (int j = threadIdx.x; j < nfi; j += blockDim.x) fmaxabs = fmax(fmaxabs, fabs(fp[4 * j]));
    if (threadIdx.x == 0) fmaxabs = fmax(fmaxabs, fabs(pp_eval(fb, fp, nfi, fb[nfi])));
    s_e[threadIdx.x] = e_hi;
    s_f[threadIdx.x] = fmaxabs;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) {
            s_e[threadIdx.x] = fmax(s_e[threadIdx.x], s_e[threadIdx.x + o]);
            s_f[threadIdx.x] = fmax(s_f[threadIdx.x], s_f[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    e_hi = s_e[0] * (1.0 + 1.0e-6);
    fmaxabs = s_f[0];
    // f scale: bring max |f| near 2^40 so fp32 products stay far from denormals
    int ex = 0;
    if (fmaxabs > 0.0 && isfinite(fmaxabs)) frexp(fmaxabs, &ex);
    hdr->e_hi = e_hi;
    hdr->inv_w = (double)NB / e_hi;
    hdr->f_scale = ldexp(1.0, 40 - ex);
}

// rigorous bounds of the cubic piece j over [a, b]
__device__ void piece_minmax(const double* __restrict__ brk, const double* __restrict__ pp, int j,
                             double a, double b, double& lo, double& hi)
{
    const double x0 = brk[j];
    const double a0 = pp[4 * j], a1 = pp[4 * j + 1], a2 = pp[4 * j + 2], a3 = pp[4 * j + 3];
    auto ev = [&](double x) {
        const double dx = x - x0;
        return fma(fma(fma(a3, dx, a2), dx, a1), dx, a0);
    };
    auto upd = [&](double v) { lo = fmin(lo, v); hi = fmax(hi, v); };
    upd(ev(a));
    upd(ev(b));
    // critical points: 3 a3 dx^2 + 2 a2 dx + a1 = 0
    const double A = 3.0 * a3, B = 2.0 * a2, C = a1;
    if (A != 0.0) {
        const double disc = B * B - 4.0 * A * C;
        if (disc >= 0.0) {
            const double sq = sqrt(disc);
            const double q = -0.5 * (B + copysign(sq, B));
            const double r1 = q / A;
            const double r2 = (q != 0.0) ? C / q : r1;
            upd(ev(fmin(fmax(x0 + r1, a), b)));
            upd(ev(fmin(fmax(x0 + r2, a), b)));
        }
    } else if (B != 0.0) {
        upd(ev(fmin(fmax(x0 - C / B, a), b)));
    }
}

// screen[b] = (lo, hi) fp32 bounds of f_scale * f(e) over the energies whose fp32
// bucket index can be b; (-inf, +inf) disables the screen for that bucket.
__global__ void prep_screen_kernel(const SamplerHeader* __restrict__ hdr,
                                   double* __restrict__ base)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= NB) return;
    const SamplerHeader h = *hdr;
    const double* fb = base + h.off_f_brk;
    const double* fp = base + h.off_f_pp;
    const int nint = h.n_f - 3;
    const double w = h.e_hi / (double)NB;
    const double edge = fmax(SCREEN_EDGE * w, 2.0e-6 * h.e_hi);
    const double a = (double)b * w - edge, c = (double)(b + 1) * w + edge;
    double lo = INFINITY, hi = -INFINITY;
    int j0 = pp_find(fb, nint, a), j1 = pp_find(fb, nint, c);
    for (int j = j0; j <= j1; ++j) {
        // piece j covers [fb[j], fb[j+1]); the end pieces extrapolate
        const double pa = (j == 0) ? a : fmax(a, fb[j]);
        const double pc = (j == nint - 1) ? c : fmin(c, fb[j + 1]);
        if (pa <= pc) piece_minmax(fb, fp, j, pa, pc, lo, hi);
    }
    lo *= h.f_scale;
    hi *= h.f_scale;
    lo -= fabs(lo) * SCREEN_REL;
    hi += fabs(hi) * SCREEN_REL;
    float flo = __double2float_rd(lo), fhi = __double2float_ru(hi);
    // tiny or non-finite bounds: leave the decision to the exact path
    const bool bad = !(isfinite(flo) && isfinite(fhi)) ||
                     (fhi != 0.0f && fabsf(fhi) < 1.0e-20f) || (flo != 0.0f && fabsf(flo) < 1.0e-20f);
    if (bad) { flo = -INFINITY; fhi = INFINITY; }
    float2* screen = (float2*)(base + h.off_screen);
    screen[b] = make_float2(flo, fhi);
}

// ---------------------------------------------------------------------------
// sampling kernel
// ---------------------------------------------------------------------------
struct SampleParams {
    const double* base;  // workspace (header at base)
    long long id_begin, id_end;
    double particle_mass;
    unsigned rk[20];     // Philox round keys (k0 + r W0, k1 + r W1), r = 0..9: constant-bank operands
    void* pos; void* vel; long long ld;
    void* mass; void* pot;
    double* radius; double* speed; int* trials; double* moments;
    cgb_host_uniforms hu;
    ScratchHeader* scratch;
    // velocity-only mode (cgb_generate_velocities): per-particle inputs
    const double* in_psi; const double* in_vesc; const double* in_env;
    long long first_id;
    long long radius_period;  // gas sub_sample: radius of id taken from id % radius_period
    // n_peers > 0: the moments go to every rank's buffer (peer-mapped over NVLink) instead of
    // `moments` alone -- the all-reduce of the path, fused into the kernels' final flush
    int n_peers;
    double* peer_moments[CGB_MAX_PEERS];
};

__device__ __forceinline__ void mom_add(const SampleParams& p, int k, double v)
{
    if (p.n_peers > 0) {
        for (int r = 0; r < p.n_peers; ++r) atomicAdd_system(p.peer_moments[r] + k, v);
    } else {
        atomicAdd(p.moments + k, v);
    }
}

// max via integer atomicMax on the bits (non-negative doubles order like integers)
__device__ __forceinline__ void mom_max(const SampleParams& p, int k, double v)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    if (p.n_peers > 0) {
        for (int r = 0; r < p.n_peers; ++r) atomicMax_system((unsigned long long*)(p.peer_moments[r] + k), b);
    } else {
        atomicMax((unsigned long long*)(p.moments + k), b);
    }
}

// Philox4x32-10 with the key schedule taken from the kernel parameters
__device__ __forceinline__ Philox4 philox_rk(const unsigned (&rk)[20], unsigned c0, unsigned c1,
                                             unsigned c2, unsigned c3)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(CGB_PHILOX_M0, c0), lo0 = CGB_PHILOX_M0 * c0;
        const unsigned hi1 = __umulhi(CGB_PHILOX_M1, c2), lo1 = CGB_PHILOX_M1 * c2;
        c0 = hi1 ^ c1 ^ rk[2 * r];
        c1 = lo1;
        c2 = hi0 ^ c3 ^ rk[2 * r + 1];
        c3 = lo0;
    }
    Philox4 o;
    o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

struct Tables {
    const double *cdf_P, *cdf_r, *cdf_s;
    const double *psi_brk, *psi_pp;
    const double *f_brk, *f_pp;
    const unsigned* cdf_lut;
    const int2* cdf_iv;
    int n_cdf, n_psi_int, n_f_int;
    float inv_w_f;
    double f_scale;
};

__device__ __forceinline__ void load_tables(const double* base, Tables& T)
{
    const SamplerHeader* hdr = (const SamplerHeader*)base;
    T.cdf_P = base + hdr->off_cdf_P; T.cdf_r = base + hdr->off_cdf_r; T.cdf_s = base + hdr->off_cdf_s;
    T.psi_brk = base + hdr->off_psi_brk; T.psi_pp = base + hdr->off_psi_pp;
    T.f_brk = base + hdr->off_f_brk; T.f_pp = base + hdr->off_f_pp;
    T.cdf_lut = (const unsigned*)(base + hdr->off_cdf_lut);
    T.cdf_iv = (const int2*)(base + hdr->off_cdf_iv);
    T.n_cdf = hdr->n_cdf; T.n_psi_int = hdr->n_psi - 3; T.n_f_int = hdr->n_f - 3;
    T.inv_w_f = (float)hdr->inv_w;
    T.f_scale = hdr->f_scale;
}

// The screen table in shared memory: NB buckets + one sentinel (-inf, +inf) that every
// out-of-range bucket index is clamped to, so such trials come out "undecided".
// The 32 KB table is staged with one TMA bulk copy (cp.async.bulk, completion on an
// mbarrier) issued by thread 0 while the other threads set up their state.
__device__ __forceinline__ void load_screen(const double* base, float2* s_screen, unsigned long long* mbar)
{
    const SamplerHeader* hdr = (const SamplerHeader*)base;
    const float2* g = (const float2*)(base + hdr->off_screen);
    const unsigned mbar_a = (unsigned)__cvta_generic_to_shared(mbar);
    const unsigned dst_a = (unsigned)__cvta_generic_to_shared(s_screen);
    constexpr unsigned BYTES = NB * sizeof(float2);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_a) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"(BYTES) : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_a),
            "l"(g), "r"(BYTES), "r"(mbar_a)
            : "memory");
        s_screen[NB] = make_float2(-INFINITY, INFINITY);
    }
    __syncthreads();  // the barrier is initialised and armed for everyone
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
            : "=r"(done)
            : "r"(mbar_a)
            : "memory");
    }
}

// largest j in [0, n_cdf-2] with P[j] <= u       (np.interp's interval, utils.py:165)
__device__ __forceinline__ int cdf_find(const Tables& T, double u)
{
    const unsigned lh = __ldg(T.cdf_lut + cdf_bucket<NBC>(u));
    int lo = (int)(lh & 0xffffu), hi = (int)(lh >> 16);
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        CGB_CHECK_INDEX(mid, T.n_cdf);
        if (__ldg(T.cdf_P + mid) <= u) lo = mid; else hi = mid - 1;
    }
    CGB_CHECK_INDEX(lo, T.n_cdf);
    return lo;
}

// envelope fv2esc = (2 psi) * f(psi)                                   (virial.py:192-193)
__device__ __forceinline__ double envelope_of(const Tables& T, double psi)
{
    return __dmul_rn(__dmul_rn(2.0, psi), pp_eval(T.f_brk, T.f_pp, T.n_f_int, psi));
}

// exact accept test, the oracle's operation order (oracle/cg_oracle_c.c: trial()):
//   v = u1*vesc; v2 = v*v; e = psi - 0.5*v2; accept iff !(f(e)*v2 < u2*fv2esc)
__device__ __forceinline__ bool exact_accept(const Tables& T, double u1, double u2, double psi,
                                             double vesc, double env, double& v2_out)
{
    const double v = __dmul_rn(u1, vesc);
    const double v2 = __dmul_rn(v, v);
    const double e = __dsub_rn(psi, __dmul_rn(0.5, v2));
    const double F = pp_eval(T.f_brk, T.f_pp, T.n_f_int, e);
    v2_out = v2;
    return !(__dmul_rn(F, v2) < __dmul_rn(u2, env));
}

// the three fp32 numbers the screen needs for one particle:
//   ps = float(psi), vs = float(vesc * 2^-32), es = float(env * f_scale * 2^-32);
// es = NaN (every comparison false => always "undecided") for particles the fp32
// screen cannot represent safely
__device__ __forceinline__ float screen_es(double psi, double vesc, double env, double f_scale)
{
    float es = (float)(env * f_scale * 0x1.0p-32);
    const float v2max = (float)(vesc * vesc);
    if (!(v2max > 1.0e-12f && v2max < 1.0e12f) || !(fabsf(es) > 1.0e-30f) || !(psi > 0.0))
        es = __int_as_float(0x7fc00000);
    return es;
}

// (-1)^k (pi/2)^(2k+1) / (2k+1)!  and  (-1)^k (pi/2)^(2k) / (2k)!, correctly rounded
__constant__ double SC2PI_SIN[9] = {0x1.921fb54442d18p+0,  -0x1.4abbce625be53p-1, 0x1.466bc6775aae2p-4,
                                    -0x1.32d2cce62bd86p-8, 0x1.50783487ee782p-13, -0x1.e3074fde8871fp-19,
                                    0x1.e8f434d018d63p-25, -0x1.6fadb9f155744p-31, 0x1.aaec32af93359p-38};
__constant__ double SC2PI_COS[9] = {1.0,                   -0x1.3bd3cc9be45dep+0, 0x1.03c1f081b5ac4p-2,
                                    -0x1.55d3c7e3cbffap-6, 0x1.e1f506891babbp-11, -0x1.a6d1f2a204a8cp-16,
                                    0x1.f9d38a3763cc3p-22, -0x1.b6e24f44b128fp-28, 0x1.20c62c2f2d7f5p-34};

// sin(2 pi u), cos(2 pi u) for u in [0, 1) (virial.py:176,180-182: phi = 2 pi u).  The quadrant
// reduction r = 4u - rint(4u) is exact; sin/cos(pi/2 r) on |r| <= 1/2 are Taylor polynomials
// whose truncation error is < 3e-18 relative, so the result is within ~1 ulp like sincospi(2u),
// at under half its instruction count (no general-argument, huge-argument or NaN paths).
__device__ __forceinline__ void sincos2pi(double u, double* sn, double* cs)
{
    const double t = __dmul_rn(u, 4.0);
    const double big = __dadd_rn(t, 6755399441055744.0);  // 1.5 * 2^52: low word holds rint(t)
    const int q = __double2loint(big);
    const double r = __dsub_rn(t, __dsub_rn(big, 6755399441055744.0));
    const double x = __dmul_rn(r, r);
    // coefficients come from the constant bank: DFMA takes c[][] operands directly, while
    // 64-bit immediates would each cost two uniform moves
    double ps = SC2PI_SIN[8];
#pragma unroll
    for (int k = 7; k >= 0; --k) ps = fma(ps, x, SC2PI_SIN[k]);
    ps = __dmul_rn(ps, r);
    double pc = SC2PI_COS[8];
#pragma unroll
    for (int k = 7; k >= 0; --k) pc = fma(pc, x, SC2PI_COS[k]);
    const bool odd = q & 1;  // quadrants 1, 3: sin <- cos, cos <- sin
    const double s = odd ? pc : ps, c = odd ? ps : pc;
    // sin < 0 in quadrants 2, 3; cos < 0 in quadrants 1, 2 (sign bit of the high word)
    *sn = __hiloint2double(__double2hiint(s) ^ ((q & 2) << 30), __double2loint(s));
    *cs = __hiloint2double(__double2hiint(c) ^ (((q + 1) & 2) << 30), __double2loint(c));
}

// Shared-memory accesses of the trial loop by 32-bit shared-window address held in a register
// the compiler cannot see through: given a pointer it re-derives the window base
// (S2UR SR_CgaCtaId + ULEA) at every use -- three times per loop iteration, with the S2UR
// latency exposed each time.
struct ScreenAddr { unsigned a; };
__device__ __forceinline__ float2 screen_load(const float2* __restrict__ s, unsigned idx) { return s[idx]; }
__device__ __forceinline__ float2 screen_load(ScreenAddr s, unsigned idx)
{
    float2 v;
    asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(s.a + idx * 8u));
    return v;
}
__device__ __forceinline__ unsigned ldsc_u32(unsigned a)
{
    unsigned v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double ldsc_f64(unsigned a)
{
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double2 ldsc_v2f64(unsigned a)
{
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds_f64(unsigned a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds_f32(unsigned a)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u32(unsigned a, unsigned v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// fp32 screen: 0 = reject, 1 = accept, 2 = undecided.
template <typename ScreenRef>
__device__ __forceinline__ int screen_trial(ScreenRef s_screen, float inv_w_f,
                                            unsigned wa, unsigned wb, float ps, float vs, float es)
{
    const float v = __uint2float_rn(wa) * vs;
    const float v2 = v * v;
    const float e = fmaf(-0.5f, v2, ps);
    const unsigned b = (unsigned)__float2int_rd(e * inv_w_f);  // negative -> huge -> sentinel
    const float B = __uint2float_rn(wb) * es;
    CGB_CHECK_INDEX(min(b, (unsigned)NB), NB + 1);
    const float2 fb = screen_load(s_screen, min(b, (unsigned)NB));
    if (fb.y * v2 < B) return 0;
    if (fb.x * v2 >= B) return 1;
    return 2;
}

struct PairResult {
    int acc;      // 0 none, 1 first trial of the pair, 2 second
    int n_exact;  // exact evaluations spent
};

// Resolve the undecided trials of a pair on the exact fp64 path.  sa/sb are the screen
// results; penv == nullptr: the envelope is recomputed from psi (fused mode).
__device__ __noinline__ PairResult resolve_pair(const Tables& T, Philox4 w, int sa, int sb,
                                                double psi, double vesc, const double* penv)
{
    PairResult r;
    r.acc = 0; r.n_exact = 0;
    const double env = penv ? *penv : envelope_of(T, psi);
    double v2;
    if (sa == 2) {
        sa = exact_accept(T, u32_to_double(w.x) * 0x1.0p-32, u32_to_double(w.y) * 0x1.0p-32, psi, vesc,
                          env, v2) ? 1 : 0;
        ++r.n_exact;
    }
    if (sa == 1) { r.acc = 1; return r; }
    if (sb == 2) {
        sb = exact_accept(T, u32_to_double(w.z) * 0x1.0p-32, u32_to_double(w.w) * 0x1.0p-32, psi, vesc,
                          env, v2) ? 1 : 0;
        ++r.n_exact;
    }
    if (sb == 1) r.acc = 2;
    return r;
}

// where the exact path finds a particle's fp64 psi and vesc: by pointer, or by shared-window
// address (tile slot), read only when a trial is undecided
struct PsiVescPtr {
    const double *ppsi, *pvesc;
    __device__ __forceinline__ double psi() const { return *ppsi; }
    __device__ __forceinline__ double vesc() const { return *pvesc; }
};
struct PsiVescSlot {
    unsigned a_psi, a_ves;
    __device__ __forceinline__ double psi() const { return lds_f64(a_psi); }
    __device__ __forceinline__ double vesc() const { return lds_f64(a_ves); }
};

// outcome of a pair from its two screen results; the exact path only where it matters
template <typename PV>
__device__ __forceinline__ PairResult decide_pair(const Tables& T, const Philox4& w, int sa, int sb,
                                                  const PV& pv, const double* penv)
{
    if (sa == 2 || (sa == 0 && sb == 2)) return resolve_pair(T, w, sa, sb, pv.psi(), pv.vesc(), penv);
    PairResult r;
    r.n_exact = 0;
    r.acc = (sa == 1) ? 1 : ((sb == 1) ? 2 : 0);
    return r;
}

// evaluate the two trials of one Philox call for one particle
template <typename ScreenRef, typename PV>
__device__ __forceinline__ PairResult eval_pair(const Tables& T, ScreenRef s_screen,
                                                const Philox4& w, float ps, float vs, float es,
                                                const PV& pv, const double* penv)
{
    const int sa = screen_trial(s_screen, T.inv_w_f, w.x, w.y, ps, vs, es);
    const int sb = screen_trial(s_screen, T.inv_w_f, w.z, w.w, ps, vs, es);
    return decide_pair(T, w, sa, sb, pv, penv);
}

template <typename OutT>
__device__ __forceinline__ void store_out(void* p, long long i, double v)
{
    ((OutT*)p)[i] = (OutT)v;
}

// speed from the accepted first uniform: v = u1 * vesc.  The oracle returns sqrt(v*v);
// in IEEE binary64 round-to-nearest sqrt(fl(v*v)) == |v| exactly whenever v*v neither
// overflows nor underflows (v is a speed in kpc/Myr, O(1)), so the sqrt is only kept for
// the out-of-range case.
__device__ __forceinline__ double speed_of(unsigned wa, double vesc)
{
    const double v = __dmul_rn(u32_to_double(wa) * 0x1.0p-32, vesc);
    const double v2 = __dmul_rn(v, v);
    return (v2 > 1.0e-280 && v2 < 1.0e280) ? fabs(v) : sqrt(v2);
}

// MODE 0: fused DM/star sampler.  MODE 1: velocity-only (Cython seam mirror).
//
// Each warp keeps TWO tile buffers in shared memory (rolling window).  Lanes claim
// particles from the current buffer; when it runs dry the warp sets up the next tile
// in the other buffer and lanes keep refilling from there while stragglers of the
// previous tile finish in their lanes.  A tile is written out (phase 2) as soon as
// all its particles are done.  When a buffer must be recycled while it still holds
// stragglers (the long tail of the trial-count distribution), when a particle has
// spent LANE_CAP_PAIRS in its lane, and at the very end of the kernel, the whole warp
// finishes a particle cooperatively (128 trials per step; the first accepted trial in
// sequence order wins, so the result is independent of scheduling).
template <typename OutT, int MODE, bool HOST_U>
__global__ void __launch_bounds__(WARPS * 32, 1) cgb_sample_kernel(const __grid_constant__ SampleParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    float2* s_screen = (float2*)smem;
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // opaque to the compiler, otherwise it re-reads SR_TID.X (a slow S2R) all over the loop
    asm volatile("" : "+r"(warp), "+r"(lane));
    // per-warp slot arrays, 2 buffers x TILE slots, 32 B per slot
    unsigned char* wbase = smem + SCREEN_BYTES + (size_t)warp * (NBUF * TILE) * 32;
    double* s_psi = (double*)wbase;            // relative potential
    double* s_ves = s_psi + NBUF * TILE;          // escape speed
    double* s_rad = s_ves + NBUF * TILE;          // radius (MODE 0)
    float* s_es = (float*)(s_rad + NBUF * TILE);  // fp32 envelope for the screen; trial count once done
    unsigned* s_wa = (unsigned*)(s_es + NBUF * TILE);  // accepted first uniform (32 bits)
    int* s_tri = (int*)s_es;
    double* s_mom = (double*)(smem + SCREEN_BYTES + (size_t)WARPS * (NBUF * TILE) * 32) + (size_t)warp * CGB_NMOMENTS;
    const unsigned FULL = 0xffffffffu;
    const unsigned lt_mask = (1u << lane) - 1u;
    // the same arrays by shared-window address, opaque to the compiler (see ScreenAddr)
    ScreenAddr scr;
    scr.a = (unsigned)__cvta_generic_to_shared(s_screen);
    unsigned a_psi = (unsigned)__cvta_generic_to_shared(s_psi), a_ves = (unsigned)__cvta_generic_to_shared(s_ves),
             a_es = (unsigned)__cvta_generic_to_shared(s_es), a_wa = (unsigned)__cvta_generic_to_shared(s_wa);
    asm volatile("" : "+r"(scr.a), "+r"(a_psi), "+r"(a_ves), "+r"(a_es), "+r"(a_wa));

    __shared__ unsigned long long s_mbar;
    Tables T;
    load_tables(p.base, T);
    if (lane < CGB_NMOMENTS) s_mom[lane] = 0.0;
    load_screen(p.base, s_screen, &s_mbar);
    __syncthreads();

    unsigned long long m_exact = 0;  // per-lane exact-evaluation counter
    unsigned m_coop = 0;             // warp-uniform
    bool no_defer = false;           // warp-uniform: the deferred list overflowed

    const long long n_local = p.id_end - p.id_begin;
    const long long id_base = (MODE == 0 ? p.id_begin : p.first_id);
    // warp-uniform buffer state
    // (scalars selected by ternaries: dynamically indexed arrays would go to local memory)
    // (read / written only through fully unrolled compare-selects so they stay in registers)
    long long b_loc0[NBUF];
    int b_cnt[NBUF];
    bool b_live[NBUF];
#pragma unroll
    for (int k = 0; k < NBUF; ++k) { b_loc0[k] = 0; b_cnt[k] = 0; b_live[k] = false; }
#define BGET(name, b) ([&]() { auto _v = name[0]; _Pragma("unroll") for (int _k = 1; _k < NBUF; ++_k) if ((b) == _k) _v = name[_k]; return _v; }())
#define BSET(name, b, v) do { _Pragma("unroll") for (int _k = 0; _k < NBUF; ++_k) if ((b) == _k) name[_k] = (v); } while (0)
    bool more_tiles = true;
    int cur = 0;

    // ---- phase 0: radius, psi, envelope for a freshly claimed tile into buffer b ----
    auto acquire = [&](int b) -> bool {
        long long tile = 0;
        if (lane == 0) tile = atomicAdd(&p.scratch->tile_counter, 1ULL);
        tile = __shfl_sync(FULL, tile, 0);
        const long long loc0 = tile * TILE;
        if (loc0 >= n_local) { more_tiles = false; return false; }
        const int cnt = (int)min((long long)TILE, n_local - loc0);
        const long long id0 = id_base + loc0;
        int nonpos = 0;
        if (MODE == 0) {
            // PH0_Q particles per lane at a time: the table walks below are chains of dependent
            // L1/L2 loads (ncu: 31 % of the kernel's warp samples sat here at 13 % of its
            // instructions, 56 % of them long-scoreboard), so independent chains are interleaved
            for (int i0 = lane; i0 < cnt; i0 += 32 * PH0_Q) {
                double u[PH0_Q], r[PH0_Q], psi[PH0_Q], env[PH0_Q];
                int lo[PH0_Q], hi[PH0_Q], jp[PH0_Q], jf[PH0_Q];
#pragma unroll
                for (int q = 0; q < PH0_Q; ++q) {
                    const int ic = min(i0 + 32 * q, cnt - 1);  // lanes past the end redo the last particle
                    const unsigned long long id = (unsigned long long)(id0 + ic);
                    if (HOST_U && p.hu.d_u_r) u[q] = p.hu.d_u_r[loc0 + ic];
                    else {
                        const Philox4 w = philox_rk(p.rk, (unsigned)id, (unsigned)(id >> 32), 0u, CGB_STREAM_RADIUS);
                        u[q] = u52(w.x, w.y);
                    }
                }
                // np.interp(u, P_r, r): j = searchsorted(P_r, u, 'right') - 1        (utils.py:165)
#pragma unroll
                for (int q = 0; q < PH0_Q; ++q) {
                    const unsigned lh = __ldg(T.cdf_lut + cdf_bucket<NBC>(u[q]));
                    lo[q] = (int)(lh & 0xffffu); hi[q] = (int)(lh >> 16);
                }
                bool open;
                do {
                    open = false;
#pragma unroll
                    for (int q = 0; q < PH0_Q; ++q) {
                        if (lo[q] < hi[q]) {
                            const int mid = (lo[q] + hi[q] + 1) >> 1;
                            CGB_CHECK_INDEX(mid, T.n_cdf);
                            if (__ldg(T.cdf_P + mid) <= u[q]) lo[q] = mid; else hi[q] = mid - 1;
                            open = open || lo[q] < hi[q];
                        }
                    }
                } while (open);
#pragma unroll
                for (int q = 0; q < PH0_Q; ++q) {
                    const int j = lo[q];
                    CGB_CHECK_INDEX(j, T.n_cdf);
                    const int2 iv = __ldg(T.cdf_iv + j);
                    r[q] = __dadd_rn(__dmul_rn(__ldg(T.cdf_s + j), __dsub_rn(u[q], __ldg(T.cdf_P + j))),
                                     __ldg(T.cdf_r + j));
                    jp[q] = iv.x; jf[q] = iv.y;
                }
                // psi = energy_spline(r); vesc = 2 psi; fv2esc = vesc * f(psi); vesc = sqrt(vesc)
                //                                                           (virial.py:191-194)
                pp_walk_q<PH0_Q>(T.psi_brk, T.n_psi_int, jp, r);
#pragma unroll
                for (int q = 0; q < PH0_Q; ++q) {
                    CGB_CHECK_INDEX(jp[q], T.n_psi_int);
                    psi[q] = pp_eval_at(T.psi_brk, T.psi_pp, jp[q], r[q]);
                }
                pp_walk_q<PH0_Q>(T.f_brk, T.n_f_int, jf, psi);
#pragma unroll
                for (int q = 0; q < PH0_Q; ++q) {
                    CGB_CHECK_INDEX(jf[q], T.n_f_int);
                    env[q] = __dmul_rn(__dmul_rn(2.0, psi[q]), pp_eval_at(T.f_brk, T.f_pp, jf[q], psi[q]));
                }
#pragma unroll
                for (int q = 0; q < PH0_Q; ++q) {
                    const int i = i0 + 32 * q;
                    if (i < cnt) {
                        const int s = b * TILE + i;
                        CGB_CHECK_INDEX(s, NBUF * TILE);
                        const double vesc = sqrt(__dmul_rn(2.0, psi[q]));
                        s_rad[s] = r[q]; s_psi[s] = psi[q]; s_ves[s] = vesc;
                        s_es[s] = screen_es(psi[q], vesc, env[q], T.f_scale);
                        if (!(env[q] > 0.0)) ++nonpos;
                    }
                }
            }
        } else {
            for (int i = lane; i < cnt; i += 32) {
                const int s = b * TILE + i;
                CGB_CHECK_INDEX(s, NBUF * TILE);
                CGB_CHECK_INDEX(loc0 + i, n_local);
                const double psi = p.in_psi[loc0 + i], vesc = p.in_vesc[loc0 + i], env = p.in_env[loc0 + i];
                s_psi[s] = psi; s_ves[s] = vesc;
                s_es[s] = screen_es(psi, vesc, env, T.f_scale);
                if (!(env > 0.0)) ++nonpos;
            }
        }
        nonpos = __reduce_add_sync(FULL, nonpos);
        if (lane == 0 && nonpos) s_mom[11] += (double)nonpos;
        __syncwarp();
        BSET(b_loc0, b, loc0); BSET(b_cnt, b, cnt); BSET(b_live, b, true);
        return true;
    };

    // ---- phase 2: directions and stores for the finished tile in buffer b ----
    auto epilogue = [&](int b) {
        const long long loc0 = BGET(b_loc0, b);
        const int cnt = BGET(b_cnt, b);
        const long long id0 = id_base + loc0;
        double a_v2 = 0.0, a_psi = 0.0, a_r = 0.0, a_px = 0.0, a_py = 0.0, a_pz = 0.0, a_vx = 0.0,
               a_vy = 0.0, a_vz = 0.0;
        unsigned long long a_tr = 0;  // 64-bit: a single long-tail particle can need > 2^31 trials
        unsigned a_tmax = 0, a_fail = 0;
        __syncwarp();
        for (int i = lane; i < cnt; i += 32) {
            const int s = b * TILE + i;
            const long long loc = loc0 + i;
            CGB_CHECK_INDEX(s, NBUF * TILE);
            CGB_CHECK_INDEX(loc, n_local);
            CGB_CHECK(MODE == 1 || p.ld >= n_local);
            const double psi = s_psi[s];
            const int nt = s_tri[s];
            const bool deferred = nt == TRI_DEFERRED;  // speed comes from cgb_hard_kernel
            // literal host-uniform mode stashed the speed itself in the vesc slot
            const double v = (nt > 0) ? ((HOST_U && p.hu.d_u_trial) ? s_ves[s] : speed_of(s_wa[s], s_ves[s]))
                                      : __longlong_as_double(0x7ff8000000000000LL);
            if (nt > 0) { a_tr += (unsigned)nt; a_tmax = max(a_tmax, (unsigned)nt); }
            else if (!deferred) ++a_fail;
            if (MODE == 1) {
                if (!deferred) {
                    p.speed[loc] = v;
                    if (p.trials) p.trials[loc] = nt;
                }
                continue;
            }
            const unsigned long long id = (unsigned long long)(id0 + i);
            const double r = s_rad[s];
            double uc, up, ucv, upv;
            if (HOST_U && p.hu.d_u_ct) { uc = p.hu.d_u_ct[loc]; up = p.hu.d_u_ph[loc]; }
            else {
                const Philox4 w = philox_rk(p.rk, (unsigned)id, (unsigned)(id >> 32), 0u, CGB_STREAM_POS_DIR);
                uc = u52(w.x, w.y); up = u52(w.z, w.w);
            }
            if (HOST_U && p.hu.d_u_ctv) { ucv = p.hu.d_u_ctv[loc]; upv = p.hu.d_u_phv[loc]; }
            else {
                const Philox4 w = philox_rk(p.rk, (unsigned)id, (unsigned)(id >> 32), 0u, CGB_STREAM_VEL_DIR);
                ucv = u52(w.x, w.y); upv = u52(w.z, w.w);
            }
            // theta = arccos(-1 + 2u), phi = 2 pi u                     (virial.py:175-176)
            const double ct = __dadd_rn(-1.0, __dmul_rn(2.0, uc));
            const double st = sqrt(__dmul_rn(__dsub_rn(1.0, ct), __dadd_rn(1.0, ct)));
            double sp, cp;
            sincos2pi(up, &sp, &cp);
            const double rs = __dmul_rn(r, st);
            const double px = __dmul_rn(rs, cp), py = __dmul_rn(rs, sp), pz = __dmul_rn(r, ct);
            const double ctv = __dadd_rn(-1.0, __dmul_rn(2.0, ucv));
            const double stv = sqrt(__dmul_rn(__dsub_rn(1.0, ctv), __dadd_rn(1.0, ctv)));
            double spv, cpv;
            sincos2pi(upv, &spv, &cpv);
            const double vsn = __dmul_rn(v, stv);
            const double vx = __dmul_rn(vsn, cpv), vy = __dmul_rn(vsn, spv), vz = __dmul_rn(v, ctv);
            store_out<OutT>(p.pos, loc, px);
            store_out<OutT>(p.pos, p.ld + loc, py);
            store_out<OutT>(p.pos, 2 * p.ld + loc, pz);
            if (!deferred) {
                store_out<OutT>(p.vel, loc, vx);
                store_out<OutT>(p.vel, p.ld + loc, vy);
                store_out<OutT>(p.vel, 2 * p.ld + loc, vz);
                if (p.speed) p.speed[loc] = v;
                if (p.trials) p.trials[loc] = nt;
            }
            if (p.mass) store_out<OutT>(p.mass, loc, p.particle_mass);
            if (p.pot) store_out<OutT>(p.pot, loc, -psi);
            if (p.radius) p.radius[loc] = r;
            if (nt > 0 || deferred) {
                a_psi += psi; a_r += r; a_px += px; a_py += py; a_pz += pz;
            }
            if (nt > 0) { a_v2 += v * v; a_vx += vx; a_vy += vy; a_vz += vz; }
        }
        if (p.moments) {
            auto wsum = [&](double x) {
                for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
                return x;
            };
            unsigned long long tr = a_tr;
            for (int o = 16; o; o >>= 1) tr += __shfl_xor_sync(FULL, tr, o);
            const unsigned tm = __reduce_max_sync(FULL, a_tmax);
            const unsigned fl = __reduce_add_sync(FULL, a_fail);
            a_v2 = wsum(a_v2); a_psi = wsum(a_psi); a_r = wsum(a_r);
            a_px = wsum(a_px); a_py = wsum(a_py); a_pz = wsum(a_pz);
            a_vx = wsum(a_vx); a_vy = wsum(a_vy); a_vz = wsum(a_vz);
            if (lane == 0) {
                s_mom[0] += (double)cnt; s_mom[1] += (double)tr; s_mom[2] = fmax(s_mom[2], (double)tm);
                s_mom[3] += a_v2; s_mom[4] += a_psi; s_mom[5] += a_px; s_mom[6] += a_py; s_mom[7] += a_pz;
                s_mom[8] += a_vx; s_mom[9] += a_vy; s_mom[10] += a_vz; s_mom[13] += (double)fl;
                s_mom[14] += a_r;
            }
        }
        __syncwarp();
        BSET(b_live, b, false);
    };

    // ---- per-lane trial state ----
    int slot = -1;
    unsigned pair = 0, id_lo = 0, id_hi = 0;
    float ps = 0.f, vs = 0.f, es = 0.f;

    // the exact path's envelope: an input array in MODE 1, recomputed from psi in MODE 0
    auto env_ptr = [&](int s) -> const double* {
        if (MODE == 0) return nullptr;
        const int b = s / TILE;
        return p.in_env + (BGET(b_loc0, b) + (s - b * TILE));
    };

    // Finish the particle held by lane `src` with the whole warp: 64 consecutive trial pairs
    // (128 trials) per step in two independent Philox chains per lane; the first accepted
    // trial in sequence order wins -- the answer of the serial loop.
    auto coop_finish = [&](int src) {
        const int s = __shfl_sync(FULL, slot, src);
        const unsigned p0 = __shfl_sync(FULL, pair, src);
        const float cps = __shfl_sync(FULL, ps, src), cvs = __shfl_sync(FULL, vs, src),
                    ces = __shfl_sync(FULL, es, src);
        const unsigned cid_lo = __shfl_sync(FULL, id_lo, src), cid_hi = __shfl_sync(FULL, id_hi, src);
        CGB_CHECK_INDEX(s, NBUF * TILE);
        const double* penv = env_ptr(s);
        const PsiVescSlot cpv = {a_psi + 8u * (unsigned)s, a_ves + 8u * (unsigned)s};
        unsigned long long pr = p0;
        for (;;) {
            const unsigned long long mine0 = pr + lane, mine1 = pr + 32 + lane;
            PairResult r0, r1;
            r0.acc = r1.acc = 0; r0.n_exact = r1.n_exact = 0;
            Philox4 w0, w1;
            w0.x = w0.y = w0.z = w0.w = 0;
            w1 = w0;
            // the first step looks at 32 pairs only: most stragglers need just a few more
            const bool two = pr != p0;
            if (two && mine1 <= 0xFFFFFFFFull) {
                w0 = philox_rk(p.rk, cid_lo, cid_hi, (unsigned)mine0, CGB_STREAM_VEL_TRIAL);
                w1 = philox_rk(p.rk, cid_lo, cid_hi, (unsigned)mine1, CGB_STREAM_VEL_TRIAL);
                r0 = eval_pair(T, scr, w0, cps, cvs, ces, cpv, penv);
                r1 = eval_pair(T, scr, w1, cps, cvs, ces, cpv, penv);
            } else if (mine0 <= 0xFFFFFFFFull) {
                w0 = philox_rk(p.rk, cid_lo, cid_hi, (unsigned)mine0, CGB_STREAM_VEL_TRIAL);
                r0 = eval_pair(T, scr, w0, cps, cvs, ces, cpv, penv);
            }
            m_exact += r0.n_exact + r1.n_exact;
            const unsigned hit0 = __ballot_sync(FULL, r0.acc != 0);
            const unsigned hit1 = __ballot_sync(FULL, r1.acc != 0);
            if (hit0 | hit1) {
                const bool first = hit0 != 0;
                if (lane == __ffs(first ? hit0 : hit1) - 1) {
                    const int acc = first ? r0.acc : r1.acc;
                    const unsigned long long mine = first ? mine0 : mine1;
                    s_wa[s] = first ? (acc == 1 ? w0.x : w0.z) : (acc == 1 ? w1.x : w1.z);
                    s_tri[s] = (int)min(2ULL * mine + (unsigned)acc, 0x7fffffffULL);
                }
                break;
            }
            pr += two ? 64 : 32;
            if (pr >= HARD_PAIRS && !no_defer) {
                // long tail: hand the particle to the grid-wide pass (cgb_hard_kernel)
                unsigned h = 0;
                if (lane == 0) h = atomicAdd(&p.scratch->hard_count, 1u);
                h = __shfl_sync(FULL, h, 0);
                if (h < HARD_CAP) {
                    if (lane == 0) {
                        const int bb = s / TILE;
                        HardEntry* e = (HardEntry*)(p.scratch + 1) + h;
                        e->id = ((unsigned long long)cid_hi << 32) | cid_lo;
                        e->loc = BGET(b_loc0, bb) + (s - bb * TILE);
                        CGB_CHECK_INDEX(e->loc, n_local);
                        e->psi = s_psi[s]; e->vesc = s_ves[s];
                        e->env = penv ? *penv : envelope_of(T, s_psi[s]);
                        e->next_pair = pr; e->best = ~0ULL;
                        s_tri[s] = TRI_DEFERRED;
                    }
                    break;
                }
                no_defer = true;  // list full: keep going in the warp
            }
            if (pr > 0xFFFFFFFFull) {  // 2^33 trials exhausted: report failure
                if (lane == 0) s_tri[s] = -1;
                break;
            }
        }
        ++m_coop;
        if (lane == src) slot = -1;
        __syncwarp();
    };

    // cooperatively finish every lane in `mask`
    auto finish_lanes = [&](unsigned mask) {
        while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            coop_finish(src);
        }
    };

    if (HOST_U && p.hu.d_u_trial) {
        // literal host-supplied (u1,u2)[t_cap] per particle: exact path only, tile by tile
        while (acquire(0)) {
            for (int i = lane; i < b_cnt[0]; i += 32) {
                const double psi = s_psi[i], vesc = s_ves[i];
                const double env = (MODE == 0) ? envelope_of(T, psi) : p.in_env[b_loc0[0] + i];
                const double* uu = p.hu.d_u_trial + (size_t)(b_loc0[0] + i) * (size_t)p.hu.t_cap * 2;
                int acc = -1;
                double v2 = 0.0;
                for (long long t = 0; t < p.hu.t_cap; ++t) {
                    if (exact_accept(T, uu[2 * t], uu[2 * t + 1], psi, vesc, env, v2)) { acc = (int)(t + 1); break; }
                }
                // keep the literal double uniform: stash the speed itself in the vesc slot
                s_ves[i] = (acc > 0) ? sqrt(v2) : __longlong_as_double(0x7ff8000000000000LL);
                s_wa[i] = 0u;
                s_tri[i] = acc;
            }
            epilogue(0);
        }
    } else {
        // One call site each for acquire() and epilogue() (they are large and inlined): the
        // loop is a small state machine.  `pending` = buffers whose tile is complete and must
        // be written out.  The hot path keeps the CURRENT buffer's claim state in plain
        // registers (cur_cnt / cur_next / cur_id0) and only tracks how many particles of the
        // OTHER (older) buffer are still unfinished (oth_left), so an iteration costs two
        // ballots beyond the trial itself.
        static_assert(NBUF == 2, "the lean phase-1 loop assumes two tile buffers per warp");
        unsigned pending = 0;
        bool exiting = false;
        cur = 1;
        int cur_cnt = 0, cur_next = 0, oth_left = 0;
        bool oth_live = false;
        unsigned cur_id0_lo = 0, cur_id0_hi = 0;
        // cooperative finishes are requested by setting fin_mask / fin_kind and served at ONE
        // place (coop_finish is large; a single inlined copy keeps the kernel's code footprint down)
        unsigned fin_mask = 0;
        int fin_kind = 0, fin_aux = 0;  // 1: final drain, 2: recycle buffer fin_aux, 3: long runners (fin_aux of them in the older tile)
        for (;;) {
            if (fin_kind) {
                finish_lanes(fin_mask);
                if (fin_kind == 1) {
                    if (b_live[0]) pending |= 1u;
                    if (b_live[1]) pending |= 2u;
                    exiting = true;
                } else if (fin_kind == 2) {
                    oth_live = false;
                    pending |= 1u << fin_aux;
                } else {
                    oth_left -= fin_aux;
                    if (oth_live && oth_left == 0) { oth_live = false; pending |= 1u << (cur ^ 1); }
                }
                fin_kind = 0;
            }
            while (pending) {
                const int k = __ffs(pending) - 1;
                pending &= pending - 1;
                epilogue(k);
            }
            if (exiting) break;
            // ---- refill idle lanes from the current buffer ----
            const unsigned need = __ballot_sync(FULL, slot < 0);
            if (need) {
                if (cur_next < cur_cnt) {
                    const int i = cur_next + __popc(need & lt_mask);
                    cur_next += __popc(need);
                    if (slot < 0 && i < cur_cnt) {
                        slot = cur * TILE + i;
                        CGB_CHECK_INDEX(slot, NBUF * TILE);
                        ps = (float)lds_f64(a_psi + 8u * (unsigned)slot);
                        vs = (float)(lds_f64(a_ves + 8u * (unsigned)slot) * 0x1.0p-32);
                        es = lds_f32(a_es + 4u * (unsigned)slot);
                        pair = 0;
                        id_lo = cur_id0_lo + (unsigned)i;
                        id_hi = cur_id0_hi + (id_lo < cur_id0_lo ? 1u : 0u);
                    }
                } else {
                    // current buffer exhausted (rare path): roll to the next tile or drain
                    const unsigned act = __ballot_sync(FULL, slot >= 0);
                    if (!more_tiles) {
                        fin_mask = act; fin_kind = 1;  // nothing left to claim anywhere: final drain
                        continue;
                    }
                    const int nxt = cur ^ 1;
                    if (oth_live) {
                        // recycle: the older tile's stragglers are finished by the whole warp,
                        // then it is written out at the top of the loop
                        fin_mask = __ballot_sync(FULL, slot >= 0 && (slot >= TILE) == (nxt == 1));
                        fin_kind = 2; fin_aux = nxt;
                        continue;
                    }
                    if (!acquire(nxt)) continue;  // more_tiles is now false: drains next round
                    // the buffer we were claiming from becomes the older one
                    oth_left = __popc(__ballot_sync(FULL, slot >= 0));
                    oth_live = cur_cnt > 0;
                    if (oth_live && oth_left == 0) { oth_live = false; pending |= 1u << cur; }
                    cur = nxt;
                    cur_cnt = BGET(b_cnt, cur);
                    cur_next = 0;
                    const unsigned long long id0 = (unsigned long long)(id_base + BGET(b_loc0, cur));
                    cur_id0_lo = (unsigned)id0; cur_id0_hi = (unsigned)(id0 >> 32);
                    continue;
                }
            }
            // ---- LEAN_PAIRS Philox calls = 2 * LEAN_PAIRS trials for every active lane ----
            bool done_oth = false;
            if (slot >= 0) {
                static_assert(LEAN_PAIRS >= 2 && LEAN_PAIRS <= 4, "2..4 trial pairs per iteration");
                // LEAN_PAIRS independent Philox chains per iteration (trial pairs pair ..
                // pair + LEAN_PAIRS - 1): more ILP and less loop overhead per trial; the later
                // pairs' work is discarded when an earlier one accepts
                const Philox4 w0 = philox_rk(p.rk, id_lo, id_hi, pair, CGB_STREAM_VEL_TRIAL);
                const Philox4 w1 = philox_rk(p.rk, id_lo, id_hi, pair + 1u, CGB_STREAM_VEL_TRIAL);
#if LEAN_PAIRS >= 3
                const Philox4 w2 = philox_rk(p.rk, id_lo, id_hi, pair + 2u, CGB_STREAM_VEL_TRIAL);
#endif
#if LEAN_PAIRS >= 4
                const Philox4 w3 = philox_rk(p.rk, id_lo, id_hi, pair + 3u, CGB_STREAM_VEL_TRIAL);
#endif
                int acc = 0;
                unsigned wacc = 0, ntr = 0;
#if LEAN_SCREEN_ALL
                // screen every trial of the iteration up front (independent fp32 chains), then
                // take the first accepted one; any undecided trial sends the lane to the
                // sequential path below
                const int t0 = screen_trial(scr, T.inv_w_f, w0.x, w0.y, ps, vs, es);
                const int t1 = screen_trial(scr, T.inv_w_f, w0.z, w0.w, ps, vs, es);
                const int t2 = screen_trial(scr, T.inv_w_f, w1.x, w1.y, ps, vs, es);
                const int t3 = screen_trial(scr, T.inv_w_f, w1.z, w1.w, ps, vs, es);
                int any = t0 | t1 | t2 | t3;
#if LEAN_PAIRS >= 3
                const int t4 = screen_trial(scr, T.inv_w_f, w2.x, w2.y, ps, vs, es);
                const int t5 = screen_trial(scr, T.inv_w_f, w2.z, w2.w, ps, vs, es);
                any |= t4 | t5;
#endif
#if LEAN_PAIRS >= 4
                const int t6 = screen_trial(scr, T.inv_w_f, w3.x, w3.y, ps, vs, es);
                const int t7 = screen_trial(scr, T.inv_w_f, w3.z, w3.w, ps, vs, es);
                any |= t6 | t7;
#endif
                if (!(any & 2)) {
                    if (any) {
                        acc = 1;
#if LEAN_PAIRS >= 4
                        if (t7) { wacc = w3.z; ntr = 8u; }
                        if (t6) { wacc = w3.x; ntr = 7u; }
#endif
#if LEAN_PAIRS >= 3
                        if (t5) { wacc = w2.z; ntr = 6u; }
                        if (t4) { wacc = w2.x; ntr = 5u; }
#endif
                        if (t3) { wacc = w1.z; ntr = 4u; }
                        if (t2) { wacc = w1.x; ntr = 3u; }
                        if (t1) { wacc = w0.z; ntr = 2u; }
                        if (t0) { wacc = w0.x; ntr = 1u; }
                        ntr += 2u * pair;
                    }
                } else
#endif
                {
#if LEAN_SCREEN_ALL
#define CGB_PAIR(k, w, ta, tb) decide_pair(T, w, ta, tb, PsiVescSlot{a_psi + 8u * (unsigned)slot, a_ves + 8u * (unsigned)slot}, env_ptr(slot))
#else
#define CGB_PAIR(k, w, ta, tb) eval_pair(T, scr, w, ps, vs, es, PsiVescSlot{a_psi + 8u * (unsigned)slot, a_ves + 8u * (unsigned)slot}, env_ptr(slot))
                    const int t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0, t6 = 0, t7 = 0;
                    (void)t0; (void)t1; (void)t2; (void)t3; (void)t4; (void)t5; (void)t6; (void)t7;
#endif
                    PairResult r = CGB_PAIR(0, w0, t0, t1);
                    m_exact += r.n_exact;
                    wacc = (r.acc == 1) ? w0.x : w0.z;
                    ntr = 2u * pair + (unsigned)r.acc;
                    if (!r.acc) {
                        r = CGB_PAIR(1, w1, t2, t3);
                        m_exact += r.n_exact;
                        wacc = (r.acc == 1) ? w1.x : w1.z;
                        ntr = 2u * pair + 2u + (unsigned)r.acc;
#if LEAN_PAIRS >= 3
                        if (!r.acc) {
                            r = CGB_PAIR(2, w2, t4, t5);
                            m_exact += r.n_exact;
                            wacc = (r.acc == 1) ? w2.x : w2.z;
                            ntr = 2u * pair + 4u + (unsigned)r.acc;
#if LEAN_PAIRS >= 4
                            if (!r.acc) {
                                r = CGB_PAIR(3, w3, t6, t7);
                                m_exact += r.n_exact;
                                wacc = (r.acc == 1) ? w3.x : w3.z;
                                ntr = 2u * pair + 6u + (unsigned)r.acc;
                            }
#endif
                        }
#endif
                    }
#undef CGB_PAIR
                    acc = r.acc;
                }
                if (acc) {
                    CGB_CHECK_INDEX(slot, NBUF * TILE);
                    CGB_CHECK(ntr >= 1u);
                    sts_u32(a_wa + 4u * (unsigned)slot, wacc);
                    sts_u32(a_es + 4u * (unsigned)slot, ntr);  // s_tri aliases s_es
                    done_oth = (slot >= TILE) != (cur == 1);
                    slot = -1;
                } else {
                    pair += (unsigned)LEAN_PAIRS;
                }
            }
            // particles of the older tile that just finished; the tile goes out when all did
            const unsigned dn = __ballot_sync(FULL, done_oth);
            if (dn) {
                oth_left -= __popc(dn);
                if (oth_live && oth_left == 0) { oth_live = false; pending |= 1u << (cur ^ 1); }
            }
            // long runners leave their lane: the warp finishes them 128 trials at a time
            const unsigned lng = __ballot_sync(FULL, slot >= 0 && pair >= LANE_CAP_PAIRS);
            if (lng) {
                const unsigned in_oth = __ballot_sync(FULL, slot >= 0 && (slot >= TILE) != (cur == 1));
                fin_mask = lng; fin_kind = 3; fin_aux = __popc(lng & in_oth);
            }
        }
    }

    // ---------------- moments ----------------
    // per-warp partial sums (shared memory) -> one CTA total -> one flush per CTA; with peers
    // the flush goes to every rank's buffer over NVLink, so it is kept to 16 words per CTA
    if (p.moments) {
        unsigned long long ex = m_exact;
        for (int o = 16; o; o >>= 1) ex += __shfl_xor_sync(FULL, ex, o);
        __syncwarp();
        if (lane == 0) {
            s_mom[12] = (double)m_coop;
            s_mom[15] = (double)ex;
        }
    }
    __syncthreads();
    if (p.moments && warp == 0 && lane < CGB_NMOMENTS) {
        const double* all = (const double*)(smem + SCREEN_BYTES + (size_t)WARPS * (NBUF * TILE) * 32);
        double v = 0.0;
        if (lane == 2) {
            for (int w = 0; w < WARPS; ++w) v = fmax(v, all[w * CGB_NMOMENTS + 2]);
            mom_max(p, 2, v);
        } else {
            for (int w = 0; w < WARPS; ++w) v += all[w * CGB_NMOMENTS + lane];
            if (v != 0.0) mom_add(p, lane, v);
        }
    }
}

// Grid-wide pass over the deferred particles: CTA c examines trial chunks c, c+G, ...
// (512 pairs each) until a chunk starts beyond the smallest accepted trial found so far
// (atomicMin).  The last CTA to finish writes the speeds / velocity vectors.  The
// result is the first accepted trial in sequence order -- identical to the serial loop.
template <typename OutT, int MODE, bool HOST_U>
__global__ void __launch_bounds__(512, 1) cgb_hard_kernel(const __grid_constant__ SampleParams p)
{
    ScratchHeader* sh = p.scratch;
    const unsigned n_hard = min(sh->hard_count, (unsigned)HARD_CAP);
    if (n_hard == 0) return;
    extern __shared__ __align__(16) unsigned char smem[];
    float2* s_screen = (float2*)smem;
    __shared__ unsigned long long s_best[16];
    __shared__ unsigned long long s_bcast;
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    __shared__ unsigned long long s_mbar;
    Tables T;
    load_tables(p.base, T);
    load_screen(p.base, s_screen, &s_mbar);
    __syncthreads();
    HardEntry* ent = (HardEntry*)(sh + 1);
    for (unsigned h = 0; h < n_hard; ++h) {
        const unsigned long long e_id = ent[h].id, e_next = ent[h].next_pair;
        const double e_psi = ent[h].psi, e_vesc = ent[h].vesc, e_env = ent[h].env;
        const float cps = (float)e_psi, cvs = (float)(e_vesc * 0x1.0p-32);
        const float ces = screen_es(e_psi, e_vesc, e_env, T.f_scale);
        volatile unsigned long long* bestp = &ent[h].best;
        for (unsigned long long chunk = blockIdx.x;; chunk += gridDim.x) {
            const unsigned long long start = e_next + chunk * 512ULL;
            if (threadIdx.x == 0) s_bcast = *bestp;
            __syncthreads();
            const unsigned long long best = s_bcast;
            if (start > 0xFFFFFFFFull || 2ULL * start >= best) break;
            const unsigned long long mine = start + threadIdx.x;
            unsigned long long t = ~0ULL;
            if (mine <= 0xFFFFFFFFull) {
                const Philox4 w = philox_rk(p.rk, (unsigned)e_id, (unsigned)(e_id >> 32), (unsigned)mine,
                                            CGB_STREAM_VEL_TRIAL);
                const PairResult r = eval_pair(T, s_screen, w, cps, cvs, ces, PsiVescPtr{&e_psi, &e_vesc}, &e_env);
                if (r.acc) t = 2ULL * mine + (unsigned)(r.acc - 1);
            }
            for (int o = 16; o; o >>= 1) {
                const unsigned long long u = __shfl_xor_sync(FULL, t, o);
                t = u < t ? u : t;
            }
            if (lane == 0) s_best[warp] = t;
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned long long m = ~0ULL;
                for (int k = 0; k < 16; ++k) m = s_best[k] < m ? s_best[k] : m;
                if (m != ~0ULL) atomicMin((unsigned long long*)&ent[h].best, m);
                s_bcast = m;
            }
            __syncthreads();
            const bool found = s_bcast != ~0ULL;
            __syncthreads();
            if (found) break;
        }
        __syncthreads();
    }
    // last CTA finalises
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&sh->done_ctas, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (unsigned h = threadIdx.x; h < n_hard; h += blockDim.x) {
        const unsigned long long e_id = ent[h].id;
        const double e_vesc = ent[h].vesc;
        const long long loc = ent[h].loc;
        CGB_CHECK_INDEX(loc, p.id_end - p.id_begin);
        const unsigned long long best = *(volatile unsigned long long*)&ent[h].best;
        double v;
        int nt;
        if (best == ~0ULL) {
            v = __longlong_as_double(0x7ff8000000000000LL);
            nt = -1;
        } else {
            const Philox4 w = philox_rk(p.rk, (unsigned)e_id, (unsigned)(e_id >> 32), (unsigned)(best >> 1),
                                        CGB_STREAM_VEL_TRIAL);
            v = speed_of((best & 1ULL) ? w.z : w.x, e_vesc);
            nt = (int)min(best + 1ULL, 0x7fffffffULL);
        }
        if (MODE == 1) {
            p.speed[loc] = v;
            if (p.trials) p.trials[loc] = nt;
        } else {
            double ucv, upv;
            if (HOST_U && p.hu.d_u_ctv) { ucv = p.hu.d_u_ctv[loc]; upv = p.hu.d_u_phv[loc]; }
            else {
                const Philox4 w = philox_rk(p.rk, (unsigned)e_id, (unsigned)(e_id >> 32), 0u, CGB_STREAM_VEL_DIR);
                ucv = u52(w.x, w.y); upv = u52(w.z, w.w);
            }
            const double ctv = __dadd_rn(-1.0, __dmul_rn(2.0, ucv));
            const double stv = sqrt(__dmul_rn(__dsub_rn(1.0, ctv), __dadd_rn(1.0, ctv)));
            double spv, cpv;
            sincos2pi(upv, &spv, &cpv);
            const double vsn = __dmul_rn(v, stv);
            const double vx = __dmul_rn(vsn, cpv), vy = __dmul_rn(vsn, spv), vz = __dmul_rn(v, ctv);
            store_out<OutT>(p.vel, loc, vx);
            store_out<OutT>(p.vel, p.ld + loc, vy);
            store_out<OutT>(p.vel, 2 * p.ld + loc, vz);
            if (p.speed) p.speed[loc] = v;
            if (p.trials) p.trials[loc] = nt;
            if (p.moments && nt > 0) {
                mom_add(p, 3, v * v);
                mom_add(p, 8, vx); mom_add(p, 9, vy); mom_add(p, 10, vz);
            }
        }
        if (p.moments) {
            if (nt > 0) {
                mom_add(p, 1, (double)nt);
                mom_max(p, 2, (double)nt);
            } else {
                mom_add(p, 13, 1.0);
            }
        }
    }
}

static size_t sample_smem_bytes()
{
    return SCREEN_BYTES + (size_t)WARPS * (NBUF * TILE) * 32 + (size_t)WARPS * CGB_NMOMENTS * sizeof(double);
}

template <typename OutT, int MODE, bool HOST_U>
static int launch_sample(const SampleParams& sp, cudaStream_t s)
{
    int dev = 0, sms = 0;
    CGB_CUDA_CHECK(cudaGetDevice(&dev));
    CGB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t smem = sample_smem_bytes();
    CGB_CUDA_CHECK(cudaFuncSetAttribute(cgb_sample_kernel<OutT, MODE, HOST_U>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long n_local = sp.id_end - sp.id_begin;
    const long long tiles = (n_local + TILE - 1) / TILE;
    long long ctas = (tiles + WARPS - 1) / WARPS;
    if (ctas > sms) ctas = sms;
    if (ctas < 1) ctas = 1;
    CGB_CUDA_CHECK(cudaMemsetAsync(sp.scratch, 0, sizeof(ScratchHeader), s));
    cgb_sample_kernel<OutT, MODE, HOST_U><<<(unsigned)ctas, WARPS * 32, smem, s>>>(sp);
    // grid-wide pass over the (rare) deferred particles; exits at once when there are none
    cgb_hard_kernel<OutT, MODE, HOST_U><<<sms, 512, SCREEN_BYTES, s>>>(sp);
    CGB_CUDA_CHECK(cudaGetLastError());
    return CGB_OK;
}

// ---------------------------------------------------------------------------
// gas / tracer kernel and the sub_sample assembly kernel (elementwise, HBM-bound)
// ---------------------------------------------------------------------------
// Two consecutive particles per thread: their independent chains interleave (the table
// walks are latency-bound) and every plane takes ONE 16-byte (fp64) / 8-byte (fp32) store
// per pair instead of two scalar ones.  The velocities of gas and tracer particles are
// identically zero (model.py:829-831, 933-935): they are written as vector stores of
// zeros, or not at all when the caller passes d_vel == NULL for storage it zeroed itself.
template <typename OutT> struct Pair2;
template <> struct Pair2<double> { typedef double2 type; };
template <> struct Pair2<float> { typedef float2 type; };

template <typename OutT>
__device__ __forceinline__ void store_pair(void* p, long long i, double a, double b, bool both, bool vec)
{
    typedef typename Pair2<OutT>::type V;
    if (vec && both) {
        V v;
        v.x = (OutT)a; v.y = (OutT)b;
        *(V*)((OutT*)p + i) = v;
    } else {
        ((OutT*)p)[i] = (OutT)a;
        if (both) ((OutT*)p)[i + 1] = (OutT)b;
    }
}

template <typename OutT, bool HOST_U>
__global__ void __launch_bounds__(256) cgb_gas_kernel(const __grid_constant__ SampleParams p, void* aux0, void* aux1)
{
    const SamplerHeader* hdr = (const SamplerHeader*)p.base;
    const double* cdf_P = p.base + hdr->off_cdf_P;
    const double* cdf_r = p.base + hdr->off_cdf_r;
    const double* cdf_s = p.base + hdr->off_cdf_s;
    const unsigned* cdf_lut = (const unsigned*)(p.base + hdr->off_cdf_lut);
    const int2* cdf_iv = (const int2*)(p.base + hdr->off_cdf_iv);
    const int npi = hdr->n_psi - 3, nai = hdr->n_aux_coef - 3;
    // table pointers hoisted out of the particle loop (the header lives in global memory)
    const double* const a0_brk = p.base + hdr->off_aux_brk[0];
    const double* const a0_pp = p.base + hdr->off_aux_pp[0];
    const double* const a1_brk = p.base + hdr->off_aux_brk[1];
    const double* const a1_pp = p.base + hdr->off_aux_pp[1];
    const double* const ps_brk = p.base + hdr->off_psi_brk;
    const double* const ps_pp = p.base + hdr->off_psi_pp;
    const long long n_local = p.id_end - p.id_begin;
    const long long n_pairs = (n_local + 1) >> 1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    // vector stores need even plane strides and 2-element aligned bases
    const size_t amask = 2 * sizeof(OutT) - 1;
    const bool vec = ((p.ld & 1) == 0) && (((size_t)p.pos | (size_t)p.vel | (size_t)p.mass | (size_t)p.pot |
                                            (size_t)aux0 | (size_t)aux1) & amask) == 0;
    for (long long pr = (long long)blockIdx.x * blockDim.x + threadIdx.x; pr < n_pairs; pr += stride) {
        const long long loc = 2 * pr;
        const bool both = loc + 1 < n_local;
        double u[2], uc[2], up[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const long long lq = both ? loc + q : loc;  // the odd tail computes particle `loc` twice
            const unsigned long long id = (unsigned long long)(p.id_begin + lq);
            // sub_sample: np.tile(radius_sub, s)[:n]                   (model.py:887-890)
            const unsigned long long rid = p.radius_period > 0 ? id % (unsigned long long)p.radius_period : id;
            if (HOST_U && p.hu.d_u_r) u[q] = p.hu.d_u_r[p.radius_period > 0 ? (long long)rid : lq];
            else {
                const Philox4 w = philox_rk(p.rk, (unsigned)rid, (unsigned)(rid >> 32), 0u, CGB_STREAM_RADIUS);
                u[q] = u52(w.x, w.y);
            }
            if (HOST_U && p.hu.d_u_ct) { uc[q] = p.hu.d_u_ct[lq]; up[q] = p.hu.d_u_ph[lq]; }
            else {
                const Philox4 w = philox_rk(p.rk, (unsigned)id, (unsigned)(id >> 32), 0u, CGB_STREAM_POS_DIR);
                uc[q] = u52(w.x, w.y); up[q] = u52(w.z, w.w);
            }
        }
        // bucket table + short search for np.interp's interval (same rule as the DM sampler);
        // the two searches advance together
        int lo[2], hi[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const unsigned lh = __ldg(cdf_lut + cdf_bucket(u[q]));
            lo[q] = (int)(lh & 0xffffu); hi[q] = (int)(lh >> 16);
        }
        while (lo[0] < hi[0] || lo[1] < hi[1]) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (lo[q] < hi[q]) {
                    const int mid = (lo[q] + hi[q] + 1) >> 1;
                    if (__ldg(cdf_P + mid) <= u[q]) lo[q] = mid; else hi[q] = mid - 1;
                }
            }
        }
        double r[2], px[2], py[2], pz[2];
        int jw[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            CGB_CHECK_INDEX(lo[q], hdr->n_cdf);
            r[q] = __dadd_rn(__dmul_rn(__ldg(cdf_s + lo[q]), __dsub_rn(u[q], __ldg(cdf_P + lo[q]))),
                             __ldg(cdf_r + lo[q]));
            jw[q] = __ldg(cdf_iv + lo[q]).x;  // starting piece for every spline over `radius`
            const double ct = __dadd_rn(-1.0, __dmul_rn(2.0, uc[q]));
            const double st = sqrt(__dmul_rn(__dsub_rn(1.0, ct), __dadd_rn(1.0, ct)));
            double sp, cp;
            sincos2pi(up[q], &sp, &cp);
            const double rs = __dmul_rn(r[q], st);
            px[q] = __dmul_rn(rs, cp); py[q] = __dmul_rn(rs, sp); pz[q] = __dmul_rn(r[q], ct);
        }
        CGB_CHECK_INDEX(loc + (both ? 1 : 0), n_local);
        CGB_CHECK(p.ld >= n_local);
        store_pair<OutT>(p.pos, loc, px[0], px[1], both, vec);
        store_pair<OutT>(p.pos, p.ld + loc, py[0], py[1], both, vec);
        store_pair<OutT>(p.pos, 2 * p.ld + loc, pz[0], pz[1], both, vec);
        if (p.vel) {
            store_pair<OutT>(p.vel, loc, 0.0, 0.0, both, vec);
            store_pair<OutT>(p.vel, p.ld + loc, 0.0, 0.0, both, vec);
            store_pair<OutT>(p.vel, 2 * p.ld + loc, 0.0, 0.0, both, vec);
        }
        if (p.mass) store_pair<OutT>(p.mass, loc, p.particle_mass, p.particle_mass, both, vec);
        // the splines normally share the radius knots: each walk starts at the previous
        // answer and only verifies it (pp_walk corrects it if the knots differ)
        if (aux0) {
            double v[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                jw[q] = pp_walk(a0_brk, nai, jw[q], r[q]);
                v[q] = pp_eval_at(a0_brk, a0_pp, jw[q], r[q]);
            }
            store_pair<OutT>(aux0, loc, v[0], v[1], both, vec);
        }
        if (aux1) {
            double v[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                jw[q] = pp_walk(a1_brk, nai, jw[q], r[q]);
                v[q] = pp_eval_at(a1_brk, a1_pp, jw[q], r[q]);
            }
            store_pair<OutT>(aux1, loc, v[0], v[1], both, vec);
        }
        if (p.pot) {
            double v[2];
#pragma unroll
            for (int q = 0; q < 2; ++q)
                v[q] = -pp_eval_at(ps_brk, ps_pp, pp_walk(ps_brk, npi, jw[q], r[q]), r[q]);
            store_pair<OutT>(p.pot, loc, v[0], v[1], both, vec);
        }
        if (p.radius) {
            p.radius[loc] = r[0];
            if (both) p.radius[loc + 1] = r[1];
        }
    }
}

// Fast path of the position-only species: one persistent 1024-thread CTA per SM with the packed
// tables (prep_gas_pack_kernel: start-index table, CDF, record A per CDF interval, record B per
// spline piece; 128 KB for the default 1000-point model) staged into shared memory by ONE TMA
// bulk copy.  The kernel is a stream of random table gathers: from L1 every gather costs one
// wavefront per distinct cache line (ncu: l1tex at 89 % of peak, the limiter of the general
// kernel), from shared memory a handful of bank passes.  Same arithmetic as the general path.
#define GAS_THREADS 1024
template <typename OutT, bool HOST_U>
__global__ void __launch_bounds__(GAS_THREADS, 1)
cgb_gas_smem_kernel(const __grid_constant__ SampleParams p, void* aux0, void* aux1)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ unsigned long long s_mbar;
    const SamplerHeader* hdr = (const SamplerHeader*)p.base;
    const int npi = hdr->n_psi - 3;
    const long long blk = hdr->gas_block_doubles;
    {
        const unsigned mbar_a = (unsigned)__cvta_generic_to_shared(&s_mbar);
        const unsigned dst_a = (unsigned)__cvta_generic_to_shared(smem);
        const unsigned bytes = (unsigned)(blk * 8);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_a) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"(bytes) : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_a),
                "l"(p.base + hdr->off_gas_lut), "r"(bytes), "r"(mbar_a)
                : "memory");
        }
        __syncthreads();
        unsigned done = 0;
        while (!done) {
            asm volatile(
                "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                : "=r"(done)
                : "r"(mbar_a)
                : "memory");
        }
    }
    // the tables by 32-bit shared-window address, opaque to the compiler: through pointers it
    // re-derived the window base (S2UR SR_CgaCtaId + ULEA) before each of the seven gathers
    unsigned lut = (unsigned)__cvta_generic_to_shared(smem);
    unsigned cdf_P = lut + 8u * (unsigned)(hdr->off_gas_P - hdr->off_gas_lut);
    unsigned tabA = lut + 8u * (unsigned)(hdr->off_gas_A - hdr->off_gas_lut);
    unsigned tabB = lut + 8u * (unsigned)(hdr->off_gas_B - hdr->off_gas_lut);
    asm volatile("" : "+r"(lut), "+r"(cdf_P), "+r"(tabA), "+r"(tabB));
    const double* const ps_brk = p.base + hdr->off_psi_brk;
    const double* const ps_pp = p.base + hdr->off_psi_pp;
    const long long n_local = p.id_end - p.id_begin;
    const long long n_pairs = (n_local + 1) >> 1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const size_t amask = 2 * sizeof(OutT) - 1;
    const bool vec = ((p.ld & 1) == 0) && (((size_t)p.pos | (size_t)p.vel | (size_t)p.mass | (size_t)p.pot |
                                            (size_t)aux0 | (size_t)aux1) & amask) == 0;
    for (long long pr = (long long)blockIdx.x * blockDim.x + threadIdx.x; pr < n_pairs; pr += stride) {
        const long long loc = 2 * pr;
        const bool both = loc + 1 < n_local;
        double u[2], uc[2], up[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const long long lq = both ? loc + q : loc;  // the odd tail computes particle `loc` twice
            const unsigned long long id = (unsigned long long)(p.id_begin + lq);
            // sub_sample: np.tile(radius_sub, s)[:n]                   (model.py:887-890)
            const unsigned long long rid = p.radius_period > 0 ? id % (unsigned long long)p.radius_period : id;
            if (HOST_U && p.hu.d_u_r) u[q] = p.hu.d_u_r[p.radius_period > 0 ? (long long)rid : lq];
            else {
                const Philox4 w = philox_rk(p.rk, (unsigned)rid, (unsigned)(rid >> 32), 0u, CGB_STREAM_RADIUS);
                u[q] = u52(w.x, w.y);
            }
            if (HOST_U && p.hu.d_u_ct) { uc[q] = p.hu.d_u_ct[lq]; up[q] = p.hu.d_u_ph[lq]; }
            else {
                const Philox4 w = philox_rk(p.rk, (unsigned)id, (unsigned)(id >> 32), 0u, CGB_STREAM_POS_DIR);
                uc[q] = u52(w.x, w.y); up[q] = u52(w.z, w.w);
            }
        }
        // np.interp's interval: largest j with P[j] <= u                  (utils.py:165)
        int lo[2], hi[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const unsigned lh = ldsc_u32(lut + 4u * (unsigned)cdf_bucket<NBG>(u[q]));
            lo[q] = (int)(lh & 0xffffu); hi[q] = (int)(lh >> 16);
        }
        while (lo[0] < hi[0] || lo[1] < hi[1]) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (lo[q] < hi[q]) {
                    const int mid = (lo[q] + hi[q] + 1) >> 1;
                    CGB_CHECK_INDEX(mid, hdr->n_cdf);
                    if (ldsc_f64(cdf_P + 8u * (unsigned)mid) <= u[q]) lo[q] = mid; else hi[q] = mid - 1;
                }
            }
        }
        double r[2], px[2], py[2], pz[2], v0[2], v1[2], vp[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            CGB_CHECK_INDEX(lo[q], hdr->n_cdf);
            const double2 a01 = ldsc_v2f64(tabA + (8u * GAS_A_STRIDE) * (unsigned)lo[q]);        // P_j, r_j
            const double2 a23 = ldsc_v2f64(tabA + (8u * GAS_A_STRIDE) * (unsigned)lo[q] + 16u);  // slope_j, r_{j+1} | NaN
            r[q] = __dadd_rn(__dmul_rn(a23.x, __dsub_rn(u[q], a01.x)), a01.y);
            int pj = lo[q] - 2;
            double brk_lo = a01.y;
            if (!(r[q] < a23.y)) {
                // irregular interval, or r rounded onto the next breakpoint: the piece FITPACK's
                // interval search picks
                pj = pp_walk(ps_brk, npi, pj, r[q]);
                brk_lo = ldsc_f64(tabB + 80u * (unsigned)pj);
            }
            CGB_CHECK_INDEX(pj, npi);
            const unsigned B = tabB + 80u * (unsigned)pj;
            const double dx = r[q] - brk_lo;
            if (aux0) {
                const double2 c01 = ldsc_v2f64(B + 16u), c23 = ldsc_v2f64(B + 32u);
                v0[q] = fma(fma(fma(c23.y, dx, c23.x), dx, c01.y), dx, c01.x);
            }
            if (aux1) {
                const double2 c01 = ldsc_v2f64(B + 48u), c23 = ldsc_v2f64(B + 64u);
                v1[q] = fma(fma(fma(c23.y, dx, c23.x), dx, c01.y), dx, c01.x);
            }
            if (p.pot) vp[q] = -pp_eval_at(ps_brk, ps_pp, pj, r[q]);  // same breakpoints (checked by the prepare kernel)
            const double ct = __dadd_rn(-1.0, __dmul_rn(2.0, uc[q]));
            const double st = sqrt(__dmul_rn(__dsub_rn(1.0, ct), __dadd_rn(1.0, ct)));
            double sp, cp;
            sincos2pi(up[q], &sp, &cp);
            const double rs = __dmul_rn(r[q], st);
            px[q] = __dmul_rn(rs, cp); py[q] = __dmul_rn(rs, sp); pz[q] = __dmul_rn(r[q], ct);
        }
        CGB_CHECK_INDEX(loc + (both ? 1 : 0), n_local);
        CGB_CHECK(p.ld >= n_local);
        store_pair<OutT>(p.pos, loc, px[0], px[1], both, vec);
        store_pair<OutT>(p.pos, p.ld + loc, py[0], py[1], both, vec);
        store_pair<OutT>(p.pos, 2 * p.ld + loc, pz[0], pz[1], both, vec);
        if (p.vel) {
            store_pair<OutT>(p.vel, loc, 0.0, 0.0, both, vec);
            store_pair<OutT>(p.vel, p.ld + loc, 0.0, 0.0, both, vec);
            store_pair<OutT>(p.vel, 2 * p.ld + loc, 0.0, 0.0, both, vec);
        }
        if (p.mass) store_pair<OutT>(p.mass, loc, p.particle_mass, p.particle_mass, both, vec);
        if (aux0) store_pair<OutT>(aux0, loc, v0[0], v0[1], both, vec);
        if (aux1) store_pair<OutT>(aux1, loc, v1[0], v1[1], both, vec);
        if (p.pot) store_pair<OutT>(p.pot, loc, vp[0], vp[1], both, vec);
        if (p.radius) {
            p.radius[loc] = r[0];
            if (both) p.radius[loc + 1] = r[1];
        }
    }
}

template <typename OutT, bool HOST_U>
__global__ void __launch_bounds__(256)
cgb_assemble_kernel(SampleParams p, const double* __restrict__ r_sub, const double* __restrict__ v_sub,
                    const double* __restrict__ psi_sub, long long n_sub)
{
    const long long n_local = p.id_end - p.id_begin;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long loc = (long long)blockIdx.x * blockDim.x + threadIdx.x; loc < n_local; loc += stride) {
        const unsigned long long id = (unsigned long long)(p.id_begin + loc);
        const long long js = (long long)(id % (unsigned long long)n_sub);  // np.tile(x_sub, s)[:n]
        const double r = r_sub[js], v = v_sub[js];
        double uc, up, ucv, upv;
        if (HOST_U && p.hu.d_u_ct) { uc = p.hu.d_u_ct[loc]; up = p.hu.d_u_ph[loc]; }
        else {
            const Philox4 w = philox_rk(p.rk, (unsigned)id, (unsigned)(id >> 32), 0u, CGB_STREAM_POS_DIR);
            uc = u52(w.x, w.y); up = u52(w.z, w.w);
        }
        if (HOST_U && p.hu.d_u_ctv) { ucv = p.hu.d_u_ctv[loc]; upv = p.hu.d_u_phv[loc]; }
        else {
            const Philox4 w = philox_rk(p.rk, (unsigned)id, (unsigned)(id >> 32), 0u, CGB_STREAM_VEL_DIR);
            ucv = u52(w.x, w.y); upv = u52(w.z, w.w);
        }
        const double ct = __dadd_rn(-1.0, __dmul_rn(2.0, uc));
        const double st = sqrt(__dmul_rn(__dsub_rn(1.0, ct), __dadd_rn(1.0, ct)));
        double sp, cp;
        sincos2pi(up, &sp, &cp);
        const double rs = __dmul_rn(r, st);
        store_out<OutT>(p.pos, loc, __dmul_rn(rs, cp));
        store_out<OutT>(p.pos, p.ld + loc, __dmul_rn(rs, sp));
        store_out<OutT>(p.pos, 2 * p.ld + loc, __dmul_rn(r, ct));
        const double ctv = __dadd_rn(-1.0, __dmul_rn(2.0, ucv));
        const double stv = sqrt(__dmul_rn(__dsub_rn(1.0, ctv), __dadd_rn(1.0, ctv)));
        double spv, cpv;
        sincos2pi(upv, &spv, &cpv);
        const double vsn = __dmul_rn(v, stv);
        store_out<OutT>(p.vel, loc, __dmul_rn(vsn, cpv));
        store_out<OutT>(p.vel, p.ld + loc, __dmul_rn(vsn, spv));
        store_out<OutT>(p.vel, 2 * p.ld + loc, __dmul_rn(v, ctv));
        if (p.mass) store_out<OutT>(p.mass, loc, p.particle_mass);
        if (p.pot && psi_sub) store_out<OutT>(p.pot, loc, -psi_sub[js]);
    }
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
static int check_desc(const cgb_sampler_desc* d, const char* who)
{
    CGB_REQUIRE(d != nullptr, "%s: null descriptor", who);
    CGB_REQUIRE(d->n_cdf >= 2 && d->d_cdf_P && d->d_cdf_r, "%s: CDF table needs >= 2 entries", who);
    // the start-index tables pack two 16-bit interval numbers per bucket
    CGB_REQUIRE(d->n_cdf < 65535, "%s: n_cdf must be < 65535 (got %d)", who, d->n_cdf);
    CGB_REQUIRE(d->n_psi >= 4 && d->d_psi_t && d->d_psi_c, "%s: psi spline needs >= 4 coefficients", who);
    CGB_REQUIRE((d->d_f_t == nullptr) == (d->d_f_c == nullptr), "%s: f knots/coeffs must both be given", who);
    if (d->d_f_t) CGB_REQUIRE(d->n_f >= 4, "%s: f spline needs >= 4 coefficients", who);
    CGB_REQUIRE(d->n_buckets == 0 || d->n_buckets == NB, "%s: n_buckets must be 0 or %d", who, NB);
    CGB_REQUIRE(d->n_aux >= 0 && d->n_aux <= 2, "%s: n_aux must be 0..2", who);
    for (int k = 0; k < d->n_aux; ++k)
        CGB_REQUIRE(d->d_aux_t[k] && d->d_aux_c[k] && d->n_aux_coef >= 4, "%s: aux spline %d incomplete", who, k);
    return CGB_OK;
}

extern "C" size_t cgb_sampler_workspace_bytes(const cgb_sampler_desc* desc)
{
    if (check_desc(desc, "cgb_sampler_workspace_bytes") != CGB_OK) return 0;
    return (size_t)make_layout(desc).total * sizeof(double);
}

extern "C" int cgb_sampler_prepare(const cgb_sampler_desc* desc, void* d_workspace,
                                   size_t workspace_bytes, void* stream)
{
    int rc = check_desc(desc, "cgb_sampler_prepare");
    if (rc) return rc;
    CGB_REQUIRE(d_workspace != nullptr, "cgb_sampler_prepare: null workspace");
    const SamplerHeader h = make_layout(desc);
    if (workspace_bytes < (size_t)h.total * sizeof(double) || ((uintptr_t)d_workspace & 255)) {
        cgb_set_error("cgb_sampler_prepare: workspace %zu B (need %zu, 256-byte aligned)",
                      workspace_bytes, (size_t)h.total * sizeof(double));
        return CGB_E_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    double* base = (double*)d_workspace;
    CGB_CUDA_CHECK(cudaMemcpyAsync(base, &h, sizeof(h), cudaMemcpyHostToDevice, s));
    prep_cdf_kernel<<<(h.n_cdf + 255) / 256, 256, 0, s>>>(desc->d_cdf_P, desc->d_cdf_r, h.n_cdf,
                                                         base + h.off_cdf_P, base + h.off_cdf_r,
                                                         base + h.off_cdf_s);
    prep_cdf_lut_kernel<NBC><<<NBC / 256, 256, 0, s>>>(desc->d_cdf_P, h.n_cdf, (unsigned*)(base + h.off_cdf_lut));
    if ((rc = cgb_launch_pp_convert(desc->d_psi_t, desc->d_psi_c, h.n_psi, base + h.off_psi_brk,
                                    base + h.off_psi_pp, s))) return rc;
    for (int k = 0; k < h.n_aux; ++k)
        if ((rc = cgb_launch_pp_convert(desc->d_aux_t[k], desc->d_aux_c[k], h.n_aux_coef,
                                        base + h.off_aux_brk[k], base + h.off_aux_pp[k], s))) return rc;
    if (h.n_f) {
        if ((rc = cgb_launch_pp_convert(desc->d_f_t, desc->d_f_c, h.n_f, base + h.off_f_brk,
                                        base + h.off_f_pp, s))) return rc;
        prep_scalars_kernel<<<1, 256, 0, s>>>((SamplerHeader*)base, base);
        prep_screen_kernel<<<NB / 256, 256, 0, s>>>((const SamplerHeader*)base, base);
    }
    prep_cdf_iv_kernel<<<(h.n_cdf + 255) / 256, 256, 0, s>>>((const SamplerHeader*)base, base);
    if (!h.n_f) {
        prep_cdf_lut_kernel<NBG><<<NBG / 256, 256, 0, s>>>(desc->d_cdf_P, h.n_cdf, (unsigned*)(base + h.off_gas_lut));
        const int nt = max(h.n_cdf, h.n_psi - 3);
        prep_gas_pack_kernel<<<(nt + 255) / 256, 256, 0, s>>>((SamplerHeader*)base, base);
    }
    CGB_CUDA_CHECK(cudaGetLastError());
    return CGB_OK;
}

static void set_round_keys(SampleParams& sp, uint64_t seed)
{
    unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        sp.rk[2 * r] = k0; sp.rk[2 * r + 1] = k1;
        k0 += CGB_PHILOX_W0; k1 += CGB_PHILOX_W1;
    }
}

static int fill_params(const cgb_sample_args* a, SampleParams& sp, const char* who, bool need_vel)
{
    CGB_REQUIRE(a != nullptr, "%s: null args", who);
    CGB_REQUIRE(a->d_workspace != nullptr, "%s: null workspace", who);
    CGB_REQUIRE(a->id_end >= a->id_begin && a->id_begin >= 0, "%s: bad id range [%lld, %lld)", who,
                (long long)a->id_begin, (long long)a->id_end);
    CGB_REQUIRE(a->out_dtype == CGB_DTYPE_F64 || a->out_dtype == CGB_DTYPE_F32, "%s: bad out_dtype %d",
                who, a->out_dtype);
    CGB_REQUIRE(a->d_pos != nullptr, "%s: null position buffer", who);
    if (need_vel) CGB_REQUIRE(a->d_vel != nullptr, "%s: null velocity buffer", who);
    CGB_REQUIRE(a->ld >= a->id_end - a->id_begin, "%s: ld %lld < particle count", who, (long long)a->ld);
    CGB_REQUIRE((a->hu.d_u_ct == nullptr) == (a->hu.d_u_ph == nullptr) &&
                (a->hu.d_u_ctv == nullptr) == (a->hu.d_u_phv == nullptr),
                "%s: direction uniforms must come in (cos, phi) pairs", who);
    CGB_REQUIRE(a->hu.d_u_trial == nullptr || a->hu.t_cap > 0, "%s: t_cap must be positive", who);
    memset(&sp, 0, sizeof(sp));
    sp.base = (const double*)a->d_workspace;
    sp.id_begin = a->id_begin; sp.id_end = a->id_end;
    sp.particle_mass = a->particle_mass;
    set_round_keys(sp, a->seed);
    sp.pos = a->d_pos; sp.vel = a->d_vel; sp.ld = a->ld; sp.mass = a->d_mass; sp.pot = a->d_pot;
    sp.radius = a->d_radius; sp.speed = a->d_speed; sp.trials = a->d_trials; sp.moments = a->d_moments;
    sp.hu = a->hu;
    sp.radius_period = a->radius_period;
    CGB_REQUIRE(a->radius_period >= 0, "%s: negative radius_period", who);
    CGB_REQUIRE(a->n_peers >= 0 && a->n_peers <= CGB_MAX_PEERS, "%s: n_peers must be 0..%d", who, CGB_MAX_PEERS);
    sp.n_peers = a->n_peers;
    for (int r = 0; r < a->n_peers; ++r) {
        CGB_REQUIRE(a->d_peer_moments[r] != nullptr, "%s: null peer moment buffer %d", who, r);
        sp.peer_moments[r] = a->d_peer_moments[r];
    }
    if (a->n_peers > 0) sp.moments = a->d_peer_moments[0];  // "moments requested" for the kernels
    return CGB_OK;
}

static bool any_host_u(const cgb_host_uniforms& hu)
{
    return hu.d_u_r || hu.d_u_ct || hu.d_u_ctv || hu.d_u_trial;
}

// flags word for cgb_sample_gas: CGB_SAMPLE_GAS_PACKED | (packed-block bytes / 256) << 8 when the
// prepared workspace carries valid packed tables, else 0.  Synchronises `stream` (one 4-byte and
// one 8-byte read-back; call it once per prepared workspace).
extern "C" int cgb_sampler_gas_flags(const void* d_workspace, int32_t* flags_out, void* stream)
{
    CGB_REQUIRE(d_workspace && flags_out, "cgb_sampler_gas_flags: null pointer");
    SamplerHeader h;
    CGB_CUDA_CHECK(cudaMemcpyAsync(&h, d_workspace, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CGB_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    *flags_out = 0;
    if (!h.n_f && h.gas_fast_ok && h.gas_block_doubles > 0) {
        const long long units = (h.gas_block_doubles * 8 + 255) / 256;
        if (units < (1 << 22)) *flags_out = CGB_SAMPLE_GAS_PACKED | (int32_t)(units << 8);
    }
    return CGB_OK;
}

extern "C" size_t cgb_sample_scratch_bytes(void)
{
    return sizeof(ScratchHeader) + (size_t)HARD_CAP * sizeof(HardEntry);
}

extern "C" int cgb_sample_particles(const cgb_sample_args* args, void* d_scratch, size_t scratch_bytes,
                                    void* stream)
{
    SampleParams sp;
    int rc = fill_params(args, sp, "cgb_sample_particles", true);
    if (rc) return rc;
    if (d_scratch == nullptr || scratch_bytes < cgb_sample_scratch_bytes() || ((uintptr_t)d_scratch & 63)) {
        cgb_set_error("cgb_sample_particles: scratch %zu B (need %zu, 64-byte aligned)", scratch_bytes,
                      cgb_sample_scratch_bytes());
        return CGB_E_WORKSPACE;
    }
    if (sp.id_end == sp.id_begin) return CGB_OK;
    sp.scratch = (ScratchHeader*)d_scratch;
    cudaStream_t s = (cudaStream_t)stream;
    const bool f32 = args->out_dtype == CGB_DTYPE_F32;
    if (any_host_u(args->hu))
        return f32 ? launch_sample<float, 0, true>(sp, s) : launch_sample<double, 0, true>(sp, s);
    return f32 ? launch_sample<float, 0, false>(sp, s) : launch_sample<double, 0, false>(sp, s);
}

extern "C" int cgb_assemble_particles(const cgb_sample_args* args, const double* d_radius_sub,
                                      const double* d_speed_sub, const double* d_psi_sub,
                                      int64_t n_sub, void* stream)
{
    SampleParams sp;
    int rc = fill_params(args, sp, "cgb_assemble_particles", true);
    if (rc) return rc;
    CGB_REQUIRE(d_radius_sub && d_speed_sub && n_sub > 0, "cgb_assemble_particles: bad sub-sample arrays");
    const long long n_local = sp.id_end - sp.id_begin;
    if (n_local == 0) return CGB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)min((n_local + 255) / 256, 148LL * 16);
    const bool f32 = args->out_dtype == CGB_DTYPE_F32, hu = any_host_u(args->hu);
    if (f32) {
        if (hu) cgb_assemble_kernel<float, true><<<blocks, 256, 0, s>>>(sp, d_radius_sub, d_speed_sub, d_psi_sub, n_sub);
        else cgb_assemble_kernel<float, false><<<blocks, 256, 0, s>>>(sp, d_radius_sub, d_speed_sub, d_psi_sub, n_sub);
    } else {
        if (hu) cgb_assemble_kernel<double, true><<<blocks, 256, 0, s>>>(sp, d_radius_sub, d_speed_sub, d_psi_sub, n_sub);
        else cgb_assemble_kernel<double, false><<<blocks, 256, 0, s>>>(sp, d_radius_sub, d_speed_sub, d_psi_sub, n_sub);
    }
    CGB_CUDA_CHECK(cudaGetLastError());
    return CGB_OK;
}

extern "C" int cgb_sample_gas(const cgb_sample_args* args, void* d_aux_out0, void* d_aux_out1, void* stream)
{
    SampleParams sp;
    int rc = fill_params(args, sp, "cgb_sample_gas", false);
    if (rc) return rc;
    const long long n_local = sp.id_end - sp.id_begin;
    if (n_local == 0) return CGB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const bool f32 = args->out_dtype == CGB_DTYPE_F32, hu = any_host_u(args->hu);
    // fast path: packed tables in shared memory.  The table size is a function of the sizes in
    // the header, which the caller's descriptor fixed; whether the packing applies is known on
    // the device only (gas_fast_ok), so that flag travels with args->flags once the caller has
    // read it back -- the wrappers do so when the tables are prepared.
    if ((args->flags & CGB_SAMPLE_GAS_PACKED) && n_local >= 4096) {
        int dev = 0, sms = 0;
        CGB_CUDA_CHECK(cudaGetDevice(&dev));
        CGB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const size_t smem = (size_t)(args->flags >> 8) * 256;  // packed-block size in 256-byte units
        if (smem > 0 && smem <= 220 * 1024) {
#define CGB_LAUNCH_GAS_SMEM(T, H) do { \
            CGB_CUDA_CHECK(cudaFuncSetAttribute(cgb_gas_smem_kernel<T, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            cgb_gas_smem_kernel<T, H><<<sms, GAS_THREADS, smem, s>>>(sp, d_aux_out0, d_aux_out1); } while (0)
            if (f32) { if (hu) CGB_LAUNCH_GAS_SMEM(float, true); else CGB_LAUNCH_GAS_SMEM(float, false); }
            else { if (hu) CGB_LAUNCH_GAS_SMEM(double, true); else CGB_LAUNCH_GAS_SMEM(double, false); }
#undef CGB_LAUNCH_GAS_SMEM
            CGB_CUDA_CHECK(cudaGetLastError());
            return CGB_OK;
        }
    }
    const unsigned blocks = (unsigned)min(((n_local + 1) / 2 + 255) / 256, 148LL * 16);
    if (f32) {
        if (hu) cgb_gas_kernel<float, true><<<blocks, 256, 0, s>>>(sp, d_aux_out0, d_aux_out1);
        else cgb_gas_kernel<float, false><<<blocks, 256, 0, s>>>(sp, d_aux_out0, d_aux_out1);
    } else {
        if (hu) cgb_gas_kernel<double, true><<<blocks, 256, 0, s>>>(sp, d_aux_out0, d_aux_out1);
        else cgb_gas_kernel<double, false><<<blocks, 256, 0, s>>>(sp, d_aux_out0, d_aux_out1);
    }
    CGB_CUDA_CHECK(cudaGetLastError());
    return CGB_OK;
}

// ---- screen audit (diagnostics / tests) --------------------------------------
// Every trial of every particle in [id_begin, id_end) is evaluated BOTH by the fp32 interval
// screen (screen_trial, the function and table the production kernel uses, with the same
// per-particle fp32 inputs) and by the exact fp64 test, in the production kernel's trial
// order (Philox counter (id, pair), two trials per call), until the exact test accepts or
// max_trials is reached.  A contradiction -- the screen rejects what the exact test accepts
// or accepts what it rejects -- would mean the screen changed a decision; the claim
// "identical to the exact evaluation" holds iff counts[2] stays 0.
//   counts[0] trials audited   [1] undecided by the screen   [2] CONTRADICTIONS
//   [3] particles stopped at max_trials   [4] screen accepts   [5] screen rejects
//   [6] particles whose screen is disabled (es = NaN)   [7] sum of accepted trial numbers
__global__ void __launch_bounds__(256) cgb_screen_audit_kernel(const __grid_constant__ SampleParams p,
                                                               unsigned long long* __restrict__ counts,
                                                               long long max_trials)
{
    extern __shared__ __align__(16) unsigned char smem[];
    float2* s_screen = (float2*)smem;
    __shared__ unsigned long long s_mbar;
    __shared__ unsigned long long s_cnt[8];
    Tables T;
    load_tables(p.base, T);
    if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
    load_screen(p.base, s_screen, &s_mbar);
    __syncthreads();
    unsigned long long c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long n_local = p.id_end - p.id_begin;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long loc = (long long)blockIdx.x * blockDim.x + threadIdx.x; loc < n_local; loc += stride) {
        const unsigned long long id = (unsigned long long)(p.id_begin + loc);
        // phase 0 of cgb_sample_kernel
        const Philox4 wr = philox_rk(p.rk, (unsigned)id, (unsigned)(id >> 32), 0u, CGB_STREAM_RADIUS);
        const double u = u52(wr.x, wr.y);
        const int j = cdf_find(T, u);
        const double r = __dadd_rn(__dmul_rn(__ldg(T.cdf_s + j), __dsub_rn(u, __ldg(T.cdf_P + j))), __ldg(T.cdf_r + j));
        const int2 iv = __ldg(T.cdf_iv + j);
        const double psi = pp_eval_at(T.psi_brk, T.psi_pp, pp_walk(T.psi_brk, T.n_psi_int, iv.x, r), r);
        const double ve2 = __dmul_rn(2.0, psi);
        const double env = __dmul_rn(ve2, pp_eval_at(T.f_brk, T.f_pp, pp_walk(T.f_brk, T.n_f_int, iv.y, psi), psi));
        const double vesc = sqrt(ve2);
        const float ps = (float)psi, vs = (float)(vesc * 0x1.0p-32);
        const float es = screen_es(psi, vesc, env, T.f_scale);
        if (es != es) ++c[6];
        long long t = 0;
        bool done = false;
        for (unsigned pair = 0; !done && t < max_trials; ++pair) {
            const Philox4 w = philox_rk(p.rk, (unsigned)id, (unsigned)(id >> 32), pair, CGB_STREAM_VEL_TRIAL);
#pragma unroll
            for (int h = 0; h < 2 && !done; ++h) {
                const unsigned wa = h ? w.z : w.x, wb = h ? w.w : w.y;
                const int sc = screen_trial(s_screen, T.inv_w_f, wa, wb, ps, vs, es);
                double v2;
                const bool ex = exact_accept(T, u32_to_double(wa) * 0x1.0p-32, u32_to_double(wb) * 0x1.0p-32,
                                             psi, vesc, env, v2);
                ++t;
                ++c[0];
                if (sc == 2) ++c[1];
                else if ((sc == 1) != ex) ++c[2];
                if (sc == 1) ++c[4];
                if (sc == 0) ++c[5];
                if (ex) { done = true; c[7] += (unsigned long long)t; }
            }
        }
        if (!done) ++c[3];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        unsigned long long v = c[k];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_cnt[k], v);
    }
    __syncthreads();
    if (threadIdx.x < 8 && s_cnt[threadIdx.x]) atomicAdd(counts + threadIdx.x, s_cnt[threadIdx.x]);
}

extern "C" int cgb_screen_audit(const void* d_workspace, int64_t id_begin, int64_t id_end, uint64_t seed,
                                int64_t max_trials, unsigned long long* d_counts, void* stream)
{
    CGB_REQUIRE(d_workspace && d_counts, "cgb_screen_audit: null pointer");
    CGB_REQUIRE(id_end >= id_begin && id_begin >= 0 && max_trials > 0, "cgb_screen_audit: bad range or trial cap");
    if (id_end == id_begin) return CGB_OK;
    SampleParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.base = (const double*)d_workspace;
    sp.id_begin = id_begin; sp.id_end = id_end;
    set_round_keys(sp, seed);
    int dev = 0, sms = 0;
    CGB_CUDA_CHECK(cudaGetDevice(&dev));
    CGB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CGB_CUDA_CHECK(cudaFuncSetAttribute(cgb_screen_audit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)SCREEN_BYTES));
    cgb_screen_audit_kernel<<<sms * 6, 256, SCREEN_BYTES, (cudaStream_t)stream>>>(sp, d_counts, (long long)max_trials);
    CGB_CUDA_CHECK(cudaGetLastError());
    return CGB_OK;
}

// ---- V: Cython-seam mirror ------------------------------------------------
// knots = t[k:-k] -> rebuild the padded knot vector on the device
__global__ void pad_knots_kernel(const double* __restrict__ knots, int n_knots, double* __restrict__ t)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = n_knots + 6;
    if (i >= n) return;
    t[i] = (i < 3) ? knots[0] : (i >= n - 3) ? knots[n_knots - 1] : knots[i - 3];
}

// a workspace with only the f tables: the psi "spline" slot is a dummy so e_hi
// falls back to the f knots' upper end; velocities-only particles with psi beyond
// it are routed to the exact path by the bucket range check.
extern "C" size_t cgb_generate_velocities_workspace_bytes(int n_coeffs)
{
    if (n_coeffs < 4) return 0;
    cgb_sampler_desc d;
    memset(&d, 0, sizeof(d));
    d.n_cdf = 2; d.n_psi = 4; d.n_f = n_coeffs;
    d.d_cdf_P = d.d_cdf_r = d.d_psi_t = d.d_psi_c = d.d_f_t = d.d_f_c = (const double*)16;
    // header+tables, then padded knots (n+4), 64 dummy doubles, tile counter
    return (size_t)make_layout(&d).total * sizeof(double) +
           cgb_align_up((size_t)(n_coeffs + 4) * 8, 256) + 512 + cgb_sample_scratch_bytes();
}

__global__ void vel_dummy_kernel(double* __restrict__ dummy)
{
    // cdf P = r = {0, 1}; psi knots = {0,0,0,0,1,1,1,1}; psi coeffs = 0
    const int i = threadIdx.x;
    if (i < 2) { dummy[i] = (double)i; }
    if (i < 8) dummy[8 + i] = (i < 4) ? 0.0 : 1.0;
    if (i < 4) dummy[16 + i] = 0.0;
}

extern "C" int cgb_generate_velocities(const double* d_psi, const double* d_vesc,
                                       const double* d_fv2esc, const double* d_knots, int n_knots,
                                       const double* d_coeffs, int n_coeffs, int k,
                                       const double* d_u_trial, int64_t t_cap, uint64_t seed,
                                       int64_t first_id, double* d_vel, int32_t* d_trials, int64_t n,
                                       void* d_work, size_t work_bytes, void* stream)
{
    CGB_REQUIRE(k == 3, "cgb_generate_velocities: only cubic splines (k=3) are supported, got k=%d", k);
    CGB_REQUIRE(n_coeffs >= 4 && n_knots == n_coeffs - 2,
                "cgb_generate_velocities: expected n_knots == n_coeffs - 2 (got %d, %d)", n_knots, n_coeffs);
    CGB_REQUIRE(n >= 0 && first_id >= 0, "cgb_generate_velocities: negative count or id");
    if (n == 0) return CGB_OK;
    CGB_REQUIRE(d_psi && d_vesc && d_fv2esc && d_knots && d_coeffs && d_vel && d_work,
                "cgb_generate_velocities: null pointer");
    CGB_REQUIRE(d_u_trial == nullptr || t_cap > 0, "cgb_generate_velocities: t_cap must be positive");
    if (work_bytes < cgb_generate_velocities_workspace_bytes(n_coeffs) || ((uintptr_t)d_work & 255)) {
        cgb_set_error("cgb_generate_velocities: workspace %zu B (need %zu, 256-byte aligned)", work_bytes,
                      cgb_generate_velocities_workspace_bytes(n_coeffs));
        return CGB_E_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    cgb_sampler_desc d;
    memset(&d, 0, sizeof(d));
    d.n_cdf = 2; d.n_psi = 4; d.n_f = n_coeffs;
    d.d_cdf_P = d.d_cdf_r = d.d_psi_t = d.d_psi_c = d.d_f_t = d.d_f_c = (const double*)16;
    const size_t tab_bytes = (size_t)make_layout(&d).total * sizeof(double);
    unsigned char* w = (unsigned char*)d_work;
    double* t_full = (double*)(w + tab_bytes);
    double* dummy = (double*)(w + tab_bytes + cgb_align_up((size_t)(n_coeffs + 4) * 8, 256));
    ScratchHeader* scratch = (ScratchHeader*)((unsigned char*)dummy + 512);
    pad_knots_kernel<<<(n_knots + 6 + 255) / 256, 256, 0, s>>>(d_knots, n_knots, t_full);
    vel_dummy_kernel<<<1, 32, 0, s>>>(dummy);
    d.d_cdf_P = dummy; d.d_cdf_r = dummy; d.d_psi_t = dummy + 8; d.d_psi_c = dummy + 16;
    d.d_f_t = t_full; d.d_f_c = d_coeffs;
    int rc = cgb_sampler_prepare(&d, d_work, tab_bytes, stream);
    if (rc) return rc;
    SampleParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.base = (const double*)d_work;
    sp.id_begin = 0; sp.id_end = n; sp.first_id = first_id;
    set_round_keys(sp, seed);
    sp.speed = d_vel; sp.trials = d_trials;
    sp.in_psi = d_psi; sp.in_vesc = d_vesc; sp.in_env = d_fv2esc;
    sp.hu.d_u_trial = d_u_trial; sp.hu.t_cap = t_cap;
    sp.scratch = scratch;
    if (d_u_trial) return launch_sample<double, 1, true>(sp, s);
    return launch_sample<double, 1, false>(sp, s);
}
