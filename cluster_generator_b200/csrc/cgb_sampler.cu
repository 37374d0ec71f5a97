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
//   * each WARP owns a tile of TILE consecutive ids, claimed from a global atomic
//     counter.  Phase 0 (uniform): radius, psi, envelope for the whole tile into
//     shared memory.  Phase 1 (divergent): one particle per lane, lanes refill from
//     the tile as they accept (ballot + popc), two trials per Philox call.  When the
//     tile runs dry, or a particle exceeds LANE_CAP trials, the warp finishes that
//     particle cooperatively (64 trials per step, first accepted trial wins).
//     Phase 2 (uniform): directions, spherical->Cartesian, coalesced plane stores.
//   * the accept test f(psi - v^2/2) v^2 >= u2 * fv2esc is first tried on a
//     conservative fp32 interval screen: a 4096-bucket table of rigorous lower/upper
//     bounds of f over each energy bucket, resident in shared memory (one 8-byte
//     load per trial).  Only when the screen cannot decide (~0.3 % of trials) is the
//     exact fp64 piecewise-polynomial spline evaluated, so accept/reject decisions
//     are identical to the exact evaluation by construction.
#include <math.h>
#include <float.h>
#include "cgb_common.cuh"
#include "cgb_spline.cuh"

#define NB 4096            // screen buckets
#define WARPS 16           // warps per CTA
#define TILE 256           // particles per warp tile
#define LANE_CAP_PAIRS 96  // trial pairs in lane mode before the warp takes over
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
    long long off_aux_brk[2], off_aux_pp[2];
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
    if (h.n_f) {
        h.off_f_brk = take(h.n_f - 2); h.off_f_pp = take(4LL * (h.n_f - 3));
        h.off_screen = take(NB);  // float2 per bucket = one double slot
    }
    for (int k = 0; k < h.n_aux; ++k) {
        h.off_aux_brk[k] = take(h.n_aux_coef - 2);
        h.off_aux_pp[k] = take(4LL * (h.n_aux_coef - 3));
    }
    h.total = o;
    return h;
}

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

// ---------------------------------------------------------------------------
// prepare kernels
// ---------------------------------------------------------------------------
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
__global__ void prep_scalars_kernel(SamplerHeader* __restrict__ hdr, double* __restrict__ base)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const SamplerHeader h = *hdr;
    const double* fb = base + h.off_f_brk;
    const double* fp = base + h.off_f_pp;
    const double* pb = base + h.off_psi_brk;
    const double* pp = base + h.off_psi_pp;
    const int nfi = h.n_f - 3, npi = h.n_psi - 3;
    double e_hi = fb[nfi];
    // psi over [0, r_max]: sample knots and the extrapolated core
    for (int j = 0; j <= npi; ++j) e_hi = fmax(e_hi, pp_eval(pb, pp, npi, pb[j]));
    for (int k = 0; k <= 64; ++k) e_hi = fmax(e_hi, pp_eval(pb, pp, npi, pb[0] * (double)k / 64.0));
    e_hi *= 1.0 + 1.0e-6;
    // f scale: bring max |f| near 2^40 so fp32 products stay far from denormals
    double fmaxabs = 0.0;
    for (int j = 0; j < nfi; ++j) fmaxabs = fmax(fmaxabs, fabs(fp[4 * j]));
    fmaxabs = fmax(fmaxabs, fabs(pp_eval(fb, fp, nfi, fb[nfi])));
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
    unsigned seed_lo, seed_hi;
    void* pos; void* vel; long long ld;
    void* mass; void* pot;
    double* radius; double* speed; int* trials; double* moments;
    cgb_host_uniforms hu;
    long long* tile_counter;
    // velocity-only mode (cgb_generate_velocities): per-particle inputs
    const double* in_psi; const double* in_vesc; const double* in_env;
    long long first_id;
};

struct Tables {
    const double *cdf_P, *cdf_r, *cdf_s;
    const double *psi_brk, *psi_pp;
    const double *f_brk, *f_pp;
    int n_cdf, n_psi_int, n_f_int;
    float inv_w_f;
    double f_scale;
};

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

// fp32 screen: 0 = reject, 1 = accept, 2 = undecided.
//   vs = float(vesc * 2^-32), es = float(env * f_scale * 2^-32), ps = float(psi)
__device__ __forceinline__ int screen_trial(const float2* __restrict__ s_screen, float inv_w_f,
                                            unsigned wa, unsigned wb, float ps, float vs, float es)
{
    const float v = __uint2float_rn(wa) * vs;
    const float v2 = v * v;
    const float e = fmaf(-0.5f, v2, ps);
    const int b = __float2int_rd(e * inv_w_f);
    const float B = __uint2float_rn(wb) * es;
    const float2 fb = s_screen[min(max(b, 0), NB - 1)];
    const bool inr = (unsigned)b < (unsigned)NB;
    if (inr && fb.y * v2 < B) return 0;
    if (inr && fb.x * v2 >= B) return 1;
    return 2;
}

struct PairResult {
    int acc;      // 0 none, 1 first trial of the pair, 2 second
    int n_exact;  // exact evaluations spent
};

// evaluate the two trials of one Philox call for one particle
__device__ __forceinline__ PairResult eval_pair(const Tables& T, const float2* s_screen,
                                                const Philox4& w, float ps, float vs, float es,
                                                double psi, double vesc, double env)
{
    PairResult r;
    r.acc = 0; r.n_exact = 0;
    int sa = screen_trial(s_screen, T.inv_w_f, w.x, w.y, ps, vs, es);
    int sb = screen_trial(s_screen, T.inv_w_f, w.z, w.w, ps, vs, es);
    if (sa == 2) {
        double v2;
        sa = exact_accept(T, u32_to_double(w.x) * 0x1.0p-32, u32_to_double(w.y) * 0x1.0p-32, psi,
                          vesc, env, v2) ? 1 : 0;
        ++r.n_exact;
    }
    if (sa == 1) { r.acc = 1; return r; }
    if (sb == 2) {
        double v2;
        sb = exact_accept(T, u32_to_double(w.z) * 0x1.0p-32, u32_to_double(w.w) * 0x1.0p-32, psi,
                          vesc, env, v2) ? 1 : 0;
        ++r.n_exact;
    }
    if (sb == 1) r.acc = 2;
    return r;
}

template <typename OutT>
__device__ __forceinline__ void store_out(void* p, long long i, double v)
{
    ((OutT*)p)[i] = (OutT)v;
}

// MODE 0: fused DM/star sampler.  MODE 1: velocity-only (Cython seam mirror).
template <typename OutT, int MODE, bool HOST_U>
__global__ void __launch_bounds__(WARPS * 32, 1) cgb_sample_kernel(SampleParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    float2* s_screen = (float2*)smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* s_psi = (double*)(smem + NB * sizeof(float2)) + (size_t)warp * 4 * TILE;
    double* s_ves = s_psi + TILE;  // vesc, overwritten by the accepted speed
    double* s_env = s_ves + TILE;
    double* s_rad = s_env + TILE;  // radius (MODE 0); trial count as double bits (both modes)
    const unsigned FULL = 0xffffffffu;

    const SamplerHeader* hdr = (const SamplerHeader*)p.base;
    Tables T;
    T.cdf_P = p.base + hdr->off_cdf_P; T.cdf_r = p.base + hdr->off_cdf_r; T.cdf_s = p.base + hdr->off_cdf_s;
    T.psi_brk = p.base + hdr->off_psi_brk; T.psi_pp = p.base + hdr->off_psi_pp;
    T.f_brk = p.base + hdr->off_f_brk; T.f_pp = p.base + hdr->off_f_pp;
    T.n_cdf = hdr->n_cdf; T.n_psi_int = hdr->n_psi - 3; T.n_f_int = hdr->n_f - 3;
    T.inv_w_f = (float)hdr->inv_w;
    T.f_scale = hdr->f_scale;
    {
        const float2* g = (const float2*)(p.base + hdr->off_screen);
        for (int i = threadIdx.x; i < NB; i += blockDim.x) s_screen[i] = g[i];
    }
    __syncthreads();

    // per-lane moment accumulators (reduced once at the end)
    double m_v2 = 0.0, m_psi = 0.0, m_r = 0.0, m_px = 0.0, m_py = 0.0, m_pz = 0.0, m_vx = 0.0,
           m_vy = 0.0, m_vz = 0.0;
    unsigned long long m_trials = 0, m_n = 0, m_nonpos = 0, m_coop = 0, m_fail = 0, m_exact = 0;
    unsigned m_tmax = 0;

    const long long n_local = p.id_end - p.id_begin;
    for (;;) {
        long long tile = 0;
        if (lane == 0) tile = atomicAdd((unsigned long long*)p.tile_counter, 1ULL);
        tile = __shfl_sync(FULL, tile, 0);
        const long long loc0 = tile * TILE;  // local index of the tile's first particle
        if (loc0 >= n_local) break;
        const int cnt = (int)min((long long)TILE, n_local - loc0);
        const long long id0 = (MODE == 0 ? p.id_begin : p.first_id) + loc0;

        // ---------------- phase 0: radius, psi, envelope ----------------
        for (int i = lane; i < cnt; i += 32) {
            double psi, vesc, env;
            if (MODE == 0) {
                const unsigned long long id = (unsigned long long)(id0 + i);
                double u;
                if (HOST_U && p.hu.d_u_r) u = p.hu.d_u_r[loc0 + i];
                else {
                    const Philox4 w = philox4x32_10((unsigned)id, (unsigned)(id >> 32), 0u,
                                                    CGB_STREAM_RADIUS, p.seed_lo, p.seed_hi);
                    u = u52(w.x, w.y);
                }
                // np.interp(u, P_r, r): j = searchsorted(P_r, u, 'right') - 1        (utils.py:165)
                int lo = 0, hi = T.n_cdf - 1;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (__ldg(T.cdf_P + mid) <= u) lo = mid; else hi = mid;
                }
                const double r = __dadd_rn(
                    __dmul_rn(__ldg(T.cdf_s + lo), __dsub_rn(u, __ldg(T.cdf_P + lo))),
                    __ldg(T.cdf_r + lo));
                // psi = energy_spline(r); vesc = 2 psi; fv2esc = vesc * f(psi); vesc = sqrt(vesc)
                //                                                           (virial.py:191-194)
                psi = pp_eval(T.psi_brk, T.psi_pp, T.n_psi_int, r);
                const double ve2 = __dmul_rn(2.0, psi);
                env = __dmul_rn(ve2, pp_eval(T.f_brk, T.f_pp, T.n_f_int, psi));
                vesc = sqrt(ve2);
                s_rad[i] = r;
            } else {
                psi = p.in_psi[loc0 + i];
                vesc = p.in_vesc[loc0 + i];
                env = p.in_env[loc0 + i];
            }
            s_psi[i] = psi; s_ves[i] = vesc; s_env[i] = env;
            if (!(env > 0.0)) ++m_nonpos;
        }
        __syncwarp();

        // ---------------- phase 1: rejection loop ----------------
        int* s_tri = (int*)(s_env);  // trial counts reuse the env slot once a particle is done
        int idx = (lane < cnt) ? lane : -1;
        int next = 32;
        unsigned pair = 0;
        float ps = 0.f, vs = 0.f, es = 0.f;
        auto load_lane = [&](int j) {
            const double psi = s_psi[j], vesc = s_ves[j], env = s_env[j];
            ps = (float)psi;
            vs = (float)(vesc * 0x1.0p-32);
            es = (float)(env * T.f_scale * 0x1.0p-32);
            // particles the fp32 screen cannot represent safely -> NaN disables it
            const float v2max = (float)(vesc * vesc);
            if (!(v2max > 1.0e-12f && v2max < 1.0e12f) || !(fabsf(es) > 1.0e-30f) || !(psi > 0.0))
                es = __int_as_float(0x7fc00000);
            pair = 0;
        };
        if (idx >= 0) load_lane(idx);

        // finish particle j cooperatively starting at pair index pair0 (warp-uniform args)
        auto coop_finish = [&](int j, unsigned pair0) {
            const double psi = s_psi[j], vesc = s_ves[j], env = s_env[j];
            float cps = (float)psi, cvs = (float)(vesc * 0x1.0p-32),
                  ces = (float)(env * T.f_scale * 0x1.0p-32);
            const float v2max = (float)(vesc * vesc);
            if (!(v2max > 1.0e-12f && v2max < 1.0e12f) || !(fabsf(ces) > 1.0e-30f) || !(psi > 0.0))
                ces = __int_as_float(0x7fc00000);
            const unsigned long long id = (unsigned long long)(id0 + j);
            unsigned long long pr = pair0;
            for (;;) {
                const unsigned long long mine = pr + lane;
                PairResult r;
                r.acc = 0; r.n_exact = 0;
                Philox4 w;
                w.x = w.y = w.z = w.w = 0;
                if (mine <= 0xFFFFFFFFull) {
                    w = philox4x32_10((unsigned)id, (unsigned)(id >> 32), (unsigned)mine,
                                      CGB_STREAM_VEL_TRIAL, p.seed_lo, p.seed_hi);
                    r = eval_pair(T, s_screen, w, cps, cvs, ces, psi, vesc, env);
                }
                m_exact += r.n_exact;
                const unsigned hit = __ballot_sync(FULL, r.acc != 0);
                if (hit) {
                    const int src = __ffs(hit) - 1;
                    if (lane == src) {
                        const unsigned wa = (r.acc == 1) ? w.x : w.z;
                        const double v = __dmul_rn(u32_to_double(wa) * 0x1.0p-32, vesc);
                        s_ves[j] = sqrt(__dmul_rn(v, v));
                        const unsigned long long nt = 2ULL * mine + (unsigned)r.acc;
                        s_tri[2 * j] = (int)min(nt, 0x7fffffffULL);
                    }
                    break;
                }
                pr += 32;
                if (pr > 0xFFFFFFFFull) {  // trial cap reached: report failure
                    if (lane == 0) { s_ves[j] = __longlong_as_double(0x7ff8000000000000LL); s_tri[2 * j] = -1; }
                    break;
                }
            }
            __syncwarp();
        };

        if (HOST_U && p.hu.d_u_trial) {
            // literal host-supplied (u1,u2)[t_cap] per particle: exact path only
            for (int i = lane; i < cnt; i += 32) {
                const double psi = s_psi[i], vesc = s_ves[i], env = s_env[i];
                const double* uu = p.hu.d_u_trial + (size_t)(loc0 + i) * (size_t)p.hu.t_cap * 2;
                int acc = -1;
                double v2 = 0.0;
                for (long long t = 0; t < p.hu.t_cap; ++t) {
                    if (exact_accept(T, uu[2 * t], uu[2 * t + 1], psi, vesc, env, v2)) { acc = (int)(t + 1); break; }
                }
                s_ves[i] = (acc > 0) ? sqrt(v2) : __longlong_as_double(0x7ff8000000000000LL);
                s_tri[2 * i] = acc;
            }
            __syncwarp();
        } else {
            for (;;) {
                const unsigned act = __ballot_sync(FULL, idx >= 0);
                if (act == 0) break;
                if (next >= cnt) {
                    // tile exhausted: drain the remaining particles cooperatively
                    unsigned rem = act;
                    while (rem) {
                        const int src = __ffs(rem) - 1;
                        rem &= rem - 1;
                        const int j = __shfl_sync(FULL, idx, src);
                        const unsigned p0 = __shfl_sync(FULL, pair, src);
                        coop_finish(j, p0);
                        ++m_coop;
                    }
                    break;
                }
                bool done = false, hard = false;
                if (idx >= 0) {
                    const unsigned long long id = (unsigned long long)(id0 + idx);
                    const Philox4 w = philox4x32_10((unsigned)id, (unsigned)(id >> 32), pair,
                                                    CGB_STREAM_VEL_TRIAL, p.seed_lo, p.seed_hi);
                    int sa = screen_trial(s_screen, T.inv_w_f, w.x, w.y, ps, vs, es);
                    int sb = screen_trial(s_screen, T.inv_w_f, w.z, w.w, ps, vs, es);
                    if (sa == 2 || (sa == 0 && sb == 2)) {
                        const PairResult r = eval_pair(T, s_screen, w, ps, vs, es, s_psi[idx],
                                                       s_ves[idx], s_env[idx]);
                        m_exact += r.n_exact;
                        sa = (r.acc == 1); sb = (r.acc == 2);
                    }
                    if (sa == 1 || sb == 1) {
                        const unsigned wa = (sa == 1) ? w.x : w.z;
                        const double vesc = s_ves[idx];
                        const double v = __dmul_rn(u32_to_double(wa) * 0x1.0p-32, vesc);
                        s_ves[idx] = sqrt(__dmul_rn(v, v));
                        s_tri[2 * idx] = (int)(2u * pair + (sa == 1 ? 1u : 2u));
                        done = true;
                    } else {
                        ++pair;
                        hard = pair >= LANE_CAP_PAIRS;
                    }
                }
                // long-running particles: the whole warp takes over
                unsigned hm = __ballot_sync(FULL, hard);
                while (hm) {
                    const int src = __ffs(hm) - 1;
                    hm &= hm - 1;
                    const int j = __shfl_sync(FULL, idx, src);
                    const unsigned p0 = __shfl_sync(FULL, pair, src);
                    coop_finish(j, p0);
                    ++m_coop;
                    if (lane == src) done = true;
                }
                // refill finished lanes from the tile
                const unsigned need = __ballot_sync(FULL, done);
                if (need) {
                    if (done) {
                        const int j = next + __popc(need & ((1u << lane) - 1u));
                        idx = (j < cnt) ? j : -1;
                        if (idx >= 0) load_lane(idx);
                    }
                    next += __popc(need);
                }
            }
            __syncwarp();
        }

        // ---------------- phase 2: directions and stores ----------------
        for (int i = lane; i < cnt; i += 32) {
            const long long loc = loc0 + i;
            const double psi = s_psi[i];
            const double v = s_ves[i];
            const int nt = s_tri[2 * i];
            ++m_n;
            if (nt > 0) { m_trials += (unsigned)nt; m_tmax = max(m_tmax, (unsigned)nt); }
            else ++m_fail;
            if (MODE == 1) {
                p.speed[loc] = v;
                if (p.trials) p.trials[loc] = nt;
                continue;
            }
            const unsigned long long id = (unsigned long long)(id0 + i);
            const double r = s_rad[i];
            double uc, up, ucv, upv;
            if (HOST_U && p.hu.d_u_ct) { uc = p.hu.d_u_ct[loc]; up = p.hu.d_u_ph[loc]; }
            else {
                const Philox4 w = philox4x32_10((unsigned)id, (unsigned)(id >> 32), 0u,
                                                CGB_STREAM_POS_DIR, p.seed_lo, p.seed_hi);
                uc = u52(w.x, w.y); up = u52(w.z, w.w);
            }
            if (HOST_U && p.hu.d_u_ctv) { ucv = p.hu.d_u_ctv[loc]; upv = p.hu.d_u_phv[loc]; }
            else {
                const Philox4 w = philox4x32_10((unsigned)id, (unsigned)(id >> 32), 0u,
                                                CGB_STREAM_VEL_DIR, p.seed_lo, p.seed_hi);
                ucv = u52(w.x, w.y); upv = u52(w.z, w.w);
            }
            // theta = arccos(-1 + 2u), phi = 2 pi u                     (virial.py:175-176)
            const double ct = __dadd_rn(-1.0, __dmul_rn(2.0, uc));
            const double st = sqrt(__dmul_rn(__dsub_rn(1.0, ct), __dadd_rn(1.0, ct)));
            double sp, cp;
            sincospi(__dmul_rn(2.0, up), &sp, &cp);
            const double rs = __dmul_rn(r, st);
            const double px = __dmul_rn(rs, cp), py = __dmul_rn(rs, sp), pz = __dmul_rn(r, ct);
            const double ctv = __dadd_rn(-1.0, __dmul_rn(2.0, ucv));
            const double stv = sqrt(__dmul_rn(__dsub_rn(1.0, ctv), __dadd_rn(1.0, ctv)));
            double spv, cpv;
            sincospi(__dmul_rn(2.0, upv), &spv, &cpv);
            const double vsn = __dmul_rn(v, stv);
            const double vx = __dmul_rn(vsn, cpv), vy = __dmul_rn(vsn, spv), vz = __dmul_rn(v, ctv);
            store_out<OutT>(p.pos, loc, px);
            store_out<OutT>(p.pos, p.ld + loc, py);
            store_out<OutT>(p.pos, 2 * p.ld + loc, pz);
            store_out<OutT>(p.vel, loc, vx);
            store_out<OutT>(p.vel, p.ld + loc, vy);
            store_out<OutT>(p.vel, 2 * p.ld + loc, vz);
            if (p.mass) store_out<OutT>(p.mass, loc, p.particle_mass);
            if (p.pot) store_out<OutT>(p.pot, loc, -psi);
            if (p.radius) p.radius[loc] = r;
            if (p.speed) p.speed[loc] = v;
            if (p.trials) p.trials[loc] = nt;
            if (nt > 0) {
                m_v2 += v * v; m_psi += psi; m_r += r;
                m_px += px; m_py += py; m_pz += pz; m_vx += vx; m_vy += vy; m_vz += vz;
            }
        }
        __syncwarp();
    }

    // ---------------- moments ----------------
    if (p.moments) {
        auto wsum = [&](double x) {
            for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
            return x;
        };
        double vals[CGB_NMOMENTS];
        vals[0] = (double)m_n; vals[1] = (double)m_trials; vals[2] = 0.0;
        vals[3] = m_v2; vals[4] = m_psi; vals[5] = m_px; vals[6] = m_py; vals[7] = m_pz;
        vals[8] = m_vx; vals[9] = m_vy; vals[10] = m_vz; vals[11] = (double)m_nonpos;
        vals[12] = (double)m_coop / 32.0;  // every lane counted each cooperative finish
        vals[13] = (double)m_fail; vals[14] = m_r; vals[15] = (double)m_exact;
        unsigned tmax = m_tmax;
        for (int o = 16; o; o >>= 1) tmax = max(tmax, __shfl_xor_sync(FULL, tmax, o));
#pragma unroll
        for (int k = 0; k < CGB_NMOMENTS; ++k) {
            if (k == 2) continue;
            const double s = wsum(vals[k]);
            if (lane == 0 && s != 0.0) atomicAdd(p.moments + k, s);
        }
        if (lane == 0) {
            // max via CAS on the double's bits (non-negative doubles order like integers)
            unsigned long long* a = (unsigned long long*)(p.moments + 2);
            atomicMax(a, (unsigned long long)__double_as_longlong((double)tmax));
        }
    }
}

static size_t sample_smem_bytes()
{
    return NB * sizeof(float2) + (size_t)WARPS * 4 * TILE * sizeof(double);
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
    cgb_sample_kernel<OutT, MODE, HOST_U><<<(unsigned)ctas, WARPS * 32, smem, s>>>(sp);
    CGB_CUDA_CHECK(cudaGetLastError());
    return CGB_OK;
}

// ---------------------------------------------------------------------------
// gas / tracer kernel and the sub_sample assembly kernel (elementwise, HBM-bound)
// ---------------------------------------------------------------------------
template <typename OutT, bool HOST_U>
__global__ void __launch_bounds__(256) cgb_gas_kernel(SampleParams p, void* aux0, void* aux1)
{
    const SamplerHeader* hdr = (const SamplerHeader*)p.base;
    const double* cdf_P = p.base + hdr->off_cdf_P;
    const double* cdf_r = p.base + hdr->off_cdf_r;
    const double* cdf_s = p.base + hdr->off_cdf_s;
    const int n_cdf = hdr->n_cdf, npi = hdr->n_psi - 3, nai = hdr->n_aux_coef - 3;
    const long long n_local = p.id_end - p.id_begin;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long loc = (long long)blockIdx.x * blockDim.x + threadIdx.x; loc < n_local; loc += stride) {
        const unsigned long long id = (unsigned long long)(p.id_begin + loc);
        double u, uc, up;
        if (HOST_U && p.hu.d_u_r) u = p.hu.d_u_r[loc];
        else {
            const Philox4 w = philox4x32_10((unsigned)id, (unsigned)(id >> 32), 0u, CGB_STREAM_RADIUS,
                                            p.seed_lo, p.seed_hi);
            u = u52(w.x, w.y);
        }
        if (HOST_U && p.hu.d_u_ct) { uc = p.hu.d_u_ct[loc]; up = p.hu.d_u_ph[loc]; }
        else {
            const Philox4 w = philox4x32_10((unsigned)id, (unsigned)(id >> 32), 0u, CGB_STREAM_POS_DIR,
                                            p.seed_lo, p.seed_hi);
            uc = u52(w.x, w.y); up = u52(w.z, w.w);
        }
        int lo = 0, hi = n_cdf - 1;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(cdf_P + mid) <= u) lo = mid; else hi = mid;
        }
        const double r = __dadd_rn(__dmul_rn(__ldg(cdf_s + lo), __dsub_rn(u, __ldg(cdf_P + lo))),
                                   __ldg(cdf_r + lo));
        const double ct = __dadd_rn(-1.0, __dmul_rn(2.0, uc));
        const double st = sqrt(__dmul_rn(__dsub_rn(1.0, ct), __dadd_rn(1.0, ct)));
        double sp, cp;
        sincospi(__dmul_rn(2.0, up), &sp, &cp);
        const double rs = __dmul_rn(r, st);
        store_out<OutT>(p.pos, loc, __dmul_rn(rs, cp));
        store_out<OutT>(p.pos, p.ld + loc, __dmul_rn(rs, sp));
        store_out<OutT>(p.pos, 2 * p.ld + loc, __dmul_rn(r, ct));
        if (p.vel) {
            store_out<OutT>(p.vel, loc, 0.0);
            store_out<OutT>(p.vel, p.ld + loc, 0.0);
            store_out<OutT>(p.vel, 2 * p.ld + loc, 0.0);
        }
        if (p.mass) store_out<OutT>(p.mass, loc, p.particle_mass);
        if (aux0) store_out<OutT>(aux0, loc, pp_eval(p.base + hdr->off_aux_brk[0], p.base + hdr->off_aux_pp[0], nai, r));
        if (aux1) store_out<OutT>(aux1, loc, pp_eval(p.base + hdr->off_aux_brk[1], p.base + hdr->off_aux_pp[1], nai, r));
        if (p.pot) store_out<OutT>(p.pot, loc, -pp_eval(p.base + hdr->off_psi_brk, p.base + hdr->off_psi_pp, npi, r));
        if (p.radius) p.radius[loc] = r;
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
            const Philox4 w = philox4x32_10((unsigned)id, (unsigned)(id >> 32), 0u, CGB_STREAM_POS_DIR,
                                            p.seed_lo, p.seed_hi);
            uc = u52(w.x, w.y); up = u52(w.z, w.w);
        }
        if (HOST_U && p.hu.d_u_ctv) { ucv = p.hu.d_u_ctv[loc]; upv = p.hu.d_u_phv[loc]; }
        else {
            const Philox4 w = philox4x32_10((unsigned)id, (unsigned)(id >> 32), 0u, CGB_STREAM_VEL_DIR,
                                            p.seed_lo, p.seed_hi);
            ucv = u52(w.x, w.y); upv = u52(w.z, w.w);
        }
        const double ct = __dadd_rn(-1.0, __dmul_rn(2.0, uc));
        const double st = sqrt(__dmul_rn(__dsub_rn(1.0, ct), __dadd_rn(1.0, ct)));
        double sp, cp;
        sincospi(__dmul_rn(2.0, up), &sp, &cp);
        const double rs = __dmul_rn(r, st);
        store_out<OutT>(p.pos, loc, __dmul_rn(rs, cp));
        store_out<OutT>(p.pos, p.ld + loc, __dmul_rn(rs, sp));
        store_out<OutT>(p.pos, 2 * p.ld + loc, __dmul_rn(r, ct));
        const double ctv = __dadd_rn(-1.0, __dmul_rn(2.0, ucv));
        const double stv = sqrt(__dmul_rn(__dsub_rn(1.0, ctv), __dadd_rn(1.0, ctv)));
        double spv, cpv;
        sincospi(__dmul_rn(2.0, upv), &spv, &cpv);
        const double vsn = __dmul_rn(v, stv);
        store_out<OutT>(p.vel, loc, __dmul_rn(vsn, cpv));
        store_out<OutT>(p.vel, p.ld + loc, __dmul_rn(vsn, spv));
        store_out<OutT>(p.vel, 2 * p.ld + loc, __dmul_rn(v, ctv));
        if (p.mass) store_out<OutT>(p.mass, loc, p.particle_mass);
        if (p.pot && psi_sub) store_out<OutT>(p.pot, loc, -psi_sub[js]);
    }
}

// psi at given radii (for the sub_sample potential): psi = energy_spline(radius_sub)
__global__ void cgb_psi_at_kernel(const double* __restrict__ base, const double* __restrict__ r,
                                  double* __restrict__ psi, long long n)
{
    const SamplerHeader* hdr = (const SamplerHeader*)base;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) psi[i] = pp_eval(base + hdr->off_psi_brk, base + hdr->off_psi_pp, hdr->n_psi - 3, r[i]);
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
static int check_desc(const cgb_sampler_desc* d, const char* who)
{
    CGB_REQUIRE(d != nullptr, "%s: null descriptor", who);
    CGB_REQUIRE(d->n_cdf >= 2 && d->d_cdf_P && d->d_cdf_r, "%s: CDF table needs >= 2 entries", who);
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
    if ((rc = cgb_launch_pp_convert(desc->d_psi_t, desc->d_psi_c, h.n_psi, base + h.off_psi_brk,
                                    base + h.off_psi_pp, s))) return rc;
    for (int k = 0; k < h.n_aux; ++k)
        if ((rc = cgb_launch_pp_convert(desc->d_aux_t[k], desc->d_aux_c[k], h.n_aux_coef,
                                        base + h.off_aux_brk[k], base + h.off_aux_pp[k], s))) return rc;
    if (h.n_f) {
        if ((rc = cgb_launch_pp_convert(desc->d_f_t, desc->d_f_c, h.n_f, base + h.off_f_brk,
                                        base + h.off_f_pp, s))) return rc;
        prep_scalars_kernel<<<1, 32, 0, s>>>((SamplerHeader*)base, base);
        prep_screen_kernel<<<NB / 256, 256, 0, s>>>((const SamplerHeader*)base, base);
    }
    CGB_CUDA_CHECK(cudaGetLastError());
    return CGB_OK;
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
    sp.seed_lo = (unsigned)a->seed; sp.seed_hi = (unsigned)(a->seed >> 32);
    sp.pos = a->d_pos; sp.vel = a->d_vel; sp.ld = a->ld; sp.mass = a->d_mass; sp.pot = a->d_pot;
    sp.radius = a->d_radius; sp.speed = a->d_speed; sp.trials = a->d_trials; sp.moments = a->d_moments;
    sp.hu = a->hu;
    return CGB_OK;
}

static bool any_host_u(const cgb_host_uniforms& hu)
{
    return hu.d_u_r || hu.d_u_ct || hu.d_u_ctv || hu.d_u_trial;
}

extern "C" int cgb_sample_particles(const cgb_sample_args* args, int64_t* d_tile_counter, void* stream)
{
    SampleParams sp;
    int rc = fill_params(args, sp, "cgb_sample_particles", true);
    if (rc) return rc;
    CGB_REQUIRE(d_tile_counter != nullptr, "cgb_sample_particles: null tile counter");
    if (sp.id_end == sp.id_begin) return CGB_OK;
    sp.tile_counter = (long long*)d_tile_counter;
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
    const unsigned blocks = (unsigned)min((n_local + 255) / 256, 148LL * 16);
    const bool f32 = args->out_dtype == CGB_DTYPE_F32, hu = any_host_u(args->hu);
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
           cgb_align_up((size_t)(n_coeffs + 4) * 8, 256) + 512 + 256;
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
    long long* counter = (long long*)((unsigned char*)dummy + 512);
    pad_knots_kernel<<<(n_knots + 6 + 255) / 256, 256, 0, s>>>(d_knots, n_knots, t_full);
    vel_dummy_kernel<<<1, 32, 0, s>>>(dummy);
    CGB_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(long long), s));
    d.d_cdf_P = dummy; d.d_cdf_r = dummy; d.d_psi_t = dummy + 8; d.d_psi_c = dummy + 16;
    d.d_f_t = t_full; d.d_f_c = d_coeffs;
    int rc = cgb_sampler_prepare(&d, d_work, tab_bytes, stream);
    if (rc) return rc;
    SampleParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.base = (const double*)d_work;
    sp.id_begin = 0; sp.id_end = n; sp.first_id = first_id;
    sp.seed_lo = (unsigned)seed; sp.seed_hi = (unsigned)(seed >> 32);
    sp.speed = d_vel; sp.trials = d_trials;
    sp.in_psi = d_psi; sp.in_vesc = d_vesc; sp.in_env = d_fv2esc;
    sp.hu.d_u_trial = d_u_trial; sp.hu.t_cap = t_cap;
    sp.tile_counter = counter;
    if (d_u_trial) return launch_sample<double, 1, true>(sp, s);
    return launch_sample<double, 1, false>(sp, s);
}
