/*
 * cgb200.h -- C ABI of libcgb200.so: the B200 (sm_100a) particle-sampling engine
 * that stands in for the native/numeric part of cluster_generator's particle path.
 *
 * Every entry point cites the reference interface it replaces (file:line into
 * /root/reference/cluster_generator).  Conventions:
 *
 *   - all functions return int: 0 = ok, <0 = CGB_E_*; the message for the last
 *     failure on the calling thread is cgb_last_error().  Nothing throws or aborts
 *     across this boundary.
 *   - every pointer named d_* (and every pointer inside the structs) is a DEVICE
 *     pointer owned by the caller (PyTorch: tensor.data_ptr()); the library never
 *     allocates, frees or retains caller memory.  Scratch is caller-provided
 *     ("workspace"), sized by the *_workspace_bytes functions.
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *     calls are asynchronous on that stream unless stated otherwise.
 *   - one host thread per GPU; the library holds no global mutable state except
 *     the thread-local error string.
 *
 * Random-number contract (shared with oracle/cg_oracle_c.c, "mode B"):
 *   Philox4x32-10, key = (seed_lo, seed_hi), counter = (id_lo, id_hi, c2, stream)
 *   with id the GLOBAL particle index, so output is a pure function of
 *   (tables, seed, id) and identical for any sharding over GPUs.
 *     stream CGB_STREAM_RADIUS   c2=0: u_r   = u52(w0,w1)
 *     stream CGB_STREAM_POS_DIR  c2=0: u_cos = u52(w0,w1), u_phi = u52(w2,w3)
 *     stream CGB_STREAM_VEL_TRIAL c2=p: trial 2p   uses u1 = w0*2^-32, u2 = w1*2^-32
 *                                       trial 2p+1 uses u1 = w2*2^-32, u2 = w3*2^-32
 *     stream CGB_STREAM_VEL_DIR  c2=0: u_cos = u52(w0,w1), u_phi = u52(w2,w3)
 *   u52(hi,lo) = bits_to_double(0x3FF0000000000000 | hi<<20 | lo>>12) - 1.0
 */
#ifndef CGB200_H
#define CGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGB_VERSION 100

#define CGB_OK 0
#define CGB_E_INVALID (-1)   /* bad argument (ValueError on the Python side)          */
#define CGB_E_CUDA (-2)      /* CUDA runtime error (RuntimeError)                      */
#define CGB_E_WORKSPACE (-3) /* workspace too small / misaligned                       */
#define CGB_E_NOTMONO (-4)   /* abscissa not strictly increasing ("x must be increasing",
                                what InterpolatedUnivariateSpline raises at virial.py:45) */
#define CGB_E_IO (-5)        /* file could not be created / read / parsed (IOError)      */

#define CGB_MAX_PEERS 8 /* ranks of one NVSwitch domain whose moment buffers a kernel can update */

#define CGB_STREAM_RADIUS 0u
#define CGB_STREAM_POS_DIR 1u
#define CGB_STREAM_VEL_TRIAL 2u
#define CGB_STREAM_VEL_DIR 3u

#define CGB_DTYPE_F64 0
#define CGB_DTYPE_F32 1

int cgb_version(void);
const char *cgb_last_error(void);

/* Number of SMs / FP64 FMA microbenchmark used for the roofline denominator
 * (bench.py).  cgb_measure_dfma_tflops launches a register-resident DFMA loop and
 * returns achieved TFLOP/s (2 flop per FMA) through *tflops (host pointer; syncs). */
int cgb_device_sm_count(int *sm_count);
int cgb_measure_dfma_tflops(int iters, double *tflops, void *stream);
int cgb_measure_ffma_tflops(int iters, double *tflops, void *stream);

/* ------------------------------------------------------------------------ */
/* Cubic interpolating splines, FITPACK-compatible                           */
/*   replaces scipy.interpolate.InterpolatedUnivariateSpline(x, y) (k=3,     */
/*   s=0, ext=0) as used at virial.py:41,45,65,67,152 and model.py:909,921,939 */
/* ------------------------------------------------------------------------ */
/* Fit: knots t = [x0]*4 + x[2:-2] + [x_{n-1}]*4 (n+4 values), n coefficients.
 * d_work: >= cgb_spline_fit_workspace_bytes(n) bytes.  The call is asynchronous and does
 * not inspect the data: x must be strictly increasing (scipy raises ValueError for that at
 * virial.py:45); the Python wrapper checks it on the host, where the tables come from.
 * method: CGB_FIT_FITPACK = fpcurf's own Givens-rotation solve, coefficients bit-equal to
 * scipy's (a serial dependency chain run by one thread from registers, ~0.26 ms for n = 1000,
 * ~0.08 ms when the factorisation is re-used; used by the f(E) build, whose quadrature
 * decisions and spline derivative amplify coefficient noise); CGB_FIT_PCR = parallel cyclic
 * reduction of the same collocation system, ~20 us, coefficients within 1e-13 of FITPACK's. */
#define CGB_FIT_FITPACK 0
#define CGB_FIT_PCR 1
/* cgb_eddington_df only, OR-ed onto CGB_FIT_FITPACK: d_work was used by an earlier
 * CGB_FIT_FITPACK call with the SAME d_ee values and has not been written since -- its Givens
 * factorisation of the energy grid is re-used (the dark-matter and stellar f(E) of one model
 * share the grid, model.py:111-127), all three fits only rotate their right-hand side. */
#define CGB_FIT_REUSE_FACTOR 0x100
size_t cgb_spline_fit_workspace_bytes(int n);
int cgb_spline_fit(const double *d_x, const double *d_y, int n, double *d_t, double *d_c,
                   void *d_work, size_t work_bytes, int method, void *stream);
/* Test hook: d_q[i] = d_x[i] / d_b[i] computed the way the back-substitution divides (correctly
 * rounded reciprocal + two FMAs); tests assert it equals the IEEE quotient bit for bit. */
int cgb_test_div_by_rcp(const double *d_x, const double *d_b, double *d_q, int64_t n, void *stream);
/* Profiling hook: clock64 ticks of the four phases (serial prologue, register-window run, serial
 * epilogue, back-substitution) of the last FITPACK-exact fit: [0..3] factorising, [4..7] re-using. */
int cgb_test_givens_cycles(long long *host8);
/* Test hooks of the factorising fit's rotations.  cgb_test_fpgivs: FITPACK's fpgivs on n operand
 * pairs (piv != 0, ww > 0) in its branch-free form and with the plain operators; d_fast / d_plain
 * receive (dd, cos, sin) per pair, d_flag the range test that sends the fit back to the plain
 * operators -- tests assert the two agree bit for bit wherever the flag is 0.
 * cgb_test_fpgivs_redo: *was_redone = whether a factorising fit since the last call repeated its
 * triangularisation with the plain operators; `force` != 0 makes the following fits do so. */
int cgb_test_fpgivs(const double *d_piv, const double *d_ww, double *d_fast, double *d_plain, int *d_flag,
                    int64_t n, void *stream);
int cgb_test_fpgivs_redo(int force, int *was_redone);

/* Bounds-checked build (`make -C cluster_generator_b200/csrc checked` -> libcgb200_checked.so,
 * -DCGB_BOUNDS_CHECK): every table, tile-buffer and output index of the kernels is a device-side
 * assert.  It stands in for compute-sanitizer, which the GPU pool does not offer; the reference
 * has no counterpart (its loops are numpy / Cython with Python's own bounds checks).
 * cgb_bounds_check_enabled: 1 in that build, 0 in the shipped library.
 * cgb_test_bounds_check: one device thread checks index i against size n; CGB_OK when the
 * device did not trap (always, in the shipped library), CGB_E_CUDA when it did -- after which
 * the process's CUDA context is unusable, so call it from a process of its own. */
int cgb_bounds_check_enabled(void);
int cgb_test_bounds_check(long long i, long long n);
/* Evaluate the spline (nu = 0) or its first derivative (nu = 1) at m points
 * (FITPACK splev/splder arithmetic, extrapolating outside the knots). */
int cgb_spline_eval(const double *d_t, const double *d_c, int n, const double *d_x, double *d_y,
                    int64_t m, int nu, void *stream);

/* ------------------------------------------------------------------------ */
/* D: Eddington distribution function                                        */
/*   replaces VirialEquilibrium._generate_df  (virial.py:43-68)              */
/* ------------------------------------------------------------------------ */
/* in : d_ee[n] ascending relative potential, d_pden[n] density in the same order.
 * out: d_g[n] (the QAGS integrals, virial.py:54-63), d_ff[n] = g'(E)/(sqrt(8) pi^2)
 *      (virial.py:65-66), and the spline of ff: d_f_t[n+4], d_f_c[n] (virial.py:67).
 *      d_last[n], d_ier[n] (nullable): QAGS sub-interval count / error flag per bin.
 * The quadrature is a decision-for-decision restatement of QUADPACK dqagse
 * (21-point Gauss-Kronrod, bisection, epsilon extrapolation, limit = 50).
 * fit_method: CGB_FIT_FITPACK makes g, ff and the f spline bit-equal to scipy's (the three
 * spline fits share one factorisation by fpcurf's own solve, ~0.55 ms of kernels in all); CGB_FIT_PCR is ~2.2x faster overall and
 * within 1e-11 of it. */
size_t cgb_eddington_workspace_bytes(int n);
int cgb_eddington_df(const double *d_ee, const double *d_pden, int n, double epsabs, double epsrel,
                     double *d_g, double *d_ff, double *d_f_t, double *d_f_c, int *d_last,
                     int *d_ier, void *d_work, size_t work_bytes, int fit_method, void *stream);

/* ------------------------------------------------------------------------ */
/* C: virial self-check                                                      */
/*   replaces VirialEquilibrium.check_virial  (virial.py:78-106)             */
/* ------------------------------------------------------------------------ */
/* rho_chk(psi_i) = 4 pi int_0^{psi_i} f(E) sqrt(2 (psi_i - E)) dE, QAGS with the
 * scipy defaults epsabs = epsrel = 1.49e-8, limit = 50.
 * d_rho[n], d_chk[n] come back in RADIUS order like the reference (rho[::-1]);
 * d_pden[n] is in ascending-psi order (the caller's reversed density). */
int cgb_check_virial(const double *d_ee, const double *d_pden, int n, const double *d_f_t,
                     const double *d_f_c, double *d_rho, double *d_chk, int *d_last, int *d_ier,
                     void *stream);

/* T (next row 3): m QAGS integrals of spline(x) * x**power over [d_a[i], d_b[i]] with the
 * spline given by FITPACK knots d_t[n+4] / coefficients d_c[n].
 *   replaces utils.integrate (utils.py:137-142) for the potential table, whose integrand
 *   is IUS(r, total_density)(r) * r  (model.py:219-224, power = 1), one quad per radius;
 *   same decision-for-decision dqagse restatement as cgb_eddington_df. */
int cgb_integrate_spline(const double *d_t, const double *d_c, int n, const double *d_a,
                         const double *d_b, int m, int power, double epsabs, double epsrel,
                         double *d_out, int *d_last, int *d_ier, void *stream);

/* T (tables): quad(profile, a_i, b_i) for a HOST callable -- integrate_mass / integrate
 * (utils.py:129-142) without one Python callback per abscissa.  QUADPACK's dqagse is replayed on
 * the device against a cache of integrand values: replay number r consumes d_cache[i][k][21]
 * for the d_ncached[i] Gauss-Kronrod calls evaluated so far and, for every integral that needs
 * more, writes the 21 abscissae of its next one or two calls to d_xreq[i][2][21] and their
 * number to d_nreq[i]; the caller evaluates the profile on all of them at once (vectorised),
 * appends the values to the cache in the same order, bumps d_ncached and calls again until
 * every d_status[i] is 1 (d_out / d_last / d_ier written then).  Caller zeroes d_status and
 * d_ncached before the first call; d_cache holds cgb_qags_replay_cache_doubles(m) doubles. */
size_t cgb_qags_replay_cache_doubles(int m);
int cgb_qags_replay(const double *d_a, const double *d_b, int m, double epsabs, double epsrel,
                    const double *d_cache, const int *d_ncached, double *d_xreq, int *d_nreq,
                    int *d_status, double *d_out, int *d_last, int *d_ier, void *stream);

/* Diagnostic: ONE QAGS integral with its full adaptive state, so tests can compare
 * the device quadrature decision for decision with QUADPACK (scipy's
 * quad(..., full_output=1): last, alist, blist, rlist, elist).
 * kind 0: integrand 2*rho'(par - t^2) of the density spline (d_t, d_c)  (virial.py:47)
 * kind 1: integrand f(e) sqrt(2 (par - e)) of the f spline (d_t, d_c)   (virial.py:96)
 * d_out[4 + 4*50] = result, abserr, last, ier, alist[50], blist[50], rlist[50], elist[50]. */
int cgb_qags_trace(const double *d_t, const double *d_c, int n, double par, double a, double b,
                   double epsabs, double epsrel, int kind, double *d_out, void *stream);

/* ------------------------------------------------------------------------ */
/* Sampler tables                                                            */
/* ------------------------------------------------------------------------ */
typedef struct cgb_sampler_desc {
    /* inverse-CDF table of generate_particle_radii (utils.py:159-164):
       P_r = [0, m[:ridx]] / m[ridx-1],  r = [0, r[:ridx]] */
    const double *d_cdf_P;
    const double *d_cdf_r;
    int32_t n_cdf;
    /* energy_spline = IUS(radius, psi) (virial.py:152): knots [n_psi+4], coeffs [n_psi] */
    int32_t n_psi;
    const double *d_psi_t;
    const double *d_psi_c;
    /* f(E) spline self.f (virial.py:67): knots [n_f+4], coeffs [n_f]; may be NULL for
       a position-only sampler (gas / tracer) */
    const double *d_f_t;
    const double *d_f_c;
    int32_t n_f;
    /* number of buckets of the fp32 screening table (power of two, 0 = default) */
    int32_t n_buckets;
    /* optional extra radial splines evaluated at the drawn radius (gas:
       thermal energy model.py:908-916, density model.py:921-930); same knot layout */
    int32_t n_aux;      /* 0, 1 or 2 */
    int32_t n_aux_coef; /* coefficients per aux spline */
    const double *d_aux_t[2];
    const double *d_aux_c[2];
} cgb_sampler_desc;

size_t cgb_sampler_workspace_bytes(const cgb_sampler_desc *desc);
/* Builds the derived device tables (CDF slopes, piecewise-polynomial forms, fp32
 * screening bounds) into the caller's workspace.  Asynchronous. */
int cgb_sampler_prepare(const cgb_sampler_desc *desc, void *d_workspace, size_t workspace_bytes,
                        void *stream);

/* Host-supplied uniforms (parity mode); all nullable, indexed by LOCAL particle
 * index (id - id_begin).  When a pointer is NULL the Philox stream is used. */
typedef struct cgb_host_uniforms {
    const double *d_u_r;    /* draw #1 utils.py:161            */
    const double *d_u_ct;   /* draw #2 virial.py:175 (u of uniform(-1,1) = -1+2u) */
    const double *d_u_ph;   /* draw #3 virial.py:176           */
    const double *d_u_ctv;  /* draw #4 virial.py:213           */
    const double *d_u_phv;  /* draw #5 virial.py:214           */
    const double *d_u_trial; /* [n_local][t_cap][2] (u1,u2) per trial, mode B literal */
    int64_t t_cap;
} cgb_host_uniforms;

#define CGB_NMOMENTS 16
/* moments[ 0] particles           [ 1] sum trials        [ 2] max trials
 *        [ 3] sum v^2             [ 4] sum psi           [ 5..7] sum x,y,z
 *        [ 8..10] sum vx,vy,vz    [11] non-positive envelopes
 *        [12] particles finished cooperatively (> lane cap)  [13] failures (trial cap)
 *        [14] sum r               [15] exact (fp64) spline evaluations in the trial loop */

typedef struct cgb_sample_args {
    const void *d_workspace; /* prepared by cgb_sampler_prepare */
    int64_t id_begin;        /* global index range [id_begin, id_end) generated by this call */
    int64_t id_end;
    double particle_mass;    /* mtot / num_particles (virial.py:225-227) */
    uint64_t seed;
    int32_t out_dtype;       /* CGB_DTYPE_F64 / CGB_DTYPE_F32 for pos/vel/mass/pot */
    int32_t flags;           /* 0, or the value cgb_sampler_gas_flags returned (cgb_sample_gas only) */
    /* outputs, local index = id - id_begin; pos/vel are (3, ld) plane-major ("SoA":
       the memory layout of unyt_array([x,y,z]).T, virial.py:180-187,216-223) */
    void *d_pos;
    void *d_vel;
    int64_t ld;              /* plane stride in elements (>= id_end - id_begin) */
    void *d_mass;            /* nullable */
    void *d_pot;             /* nullable: particle_potential = -psi (virial.py:229-234) */
    double *d_radius;        /* nullable, always f64 */
    double *d_speed;         /* nullable, always f64 */
    int32_t *d_trials;       /* nullable */
    double *d_moments;       /* nullable, CGB_NMOMENTS doubles, ACCUMULATED (caller zeroes) */
    cgb_host_uniforms hu;
    /* cgb_sample_gas only: sub_sample support (model.py:887-890, 911-930).  > 0: the radius
       (and everything derived from it) of particle id is that of id % radius_period, i.e.
       np.tile(radius_sub, sub_sample); directions stay per particle.  d_u_r is then the
       FULL radius_period-long array, indexed by id % radius_period. */
    int64_t radius_period;
    /* n_peers > 0 (cgb_sample_particles / cgb_generate_velocities path): the moment all-reduce of
       the path fused into the kernel -- instead of d_moments, the
       kernel's final flush adds its partial sums to d_peer_moments[0..n_peers), one buffer per
       rank (this rank's own and the peers', mapped with cgb_peer_open), with system-scope
       atomics over NVLink.  After every rank's kernel has completed, each buffer holds the
       global moments; no collective launch follows the kernel. */
    int32_t n_peers;
    int32_t pad_peers;
    double *d_peer_moments[CGB_MAX_PEERS];
} cgb_sample_args;

/* P+R+E+V fused: replaces VirialEquilibrium.generate_particles (virial.py:108-236)
 * for sub_sample == 1: radius draw (utils.py:154-166), psi / envelope
 * (virial.py:191-194), rejection-sampled speed (the absent opt/cython_utils.pyx
 * generate_velocities, call site virial.py:198-206), isotropic directions and the
 * SoA field stores.  d_scratch: >= cgb_sample_scratch_bytes() bytes of device memory,
 * 64-byte aligned, private to this call while it runs (dynamic tile counter and the
 * list of long-tail particles finished by a grid-wide second pass); the library
 * initialises it. */
size_t cgb_sample_scratch_bytes(void);
int cgb_sample_particles(const cgb_sample_args *args, void *d_scratch, size_t scratch_bytes,
                         void *stream);

/* Screen audit (diagnostics): runs the fp32 interval screen AND the exact fp64 accept test on
 * every trial of particles [id_begin, id_end) of a prepared DM/star sampler, in the production
 * kernel's trial order, up to max_trials per particle.  d_counts: uint64[8], ACCUMULATED
 * (caller zeroes): [0] trials  [1] undecided by the screen  [2] contradictions (must stay 0:
 * the screen never changes a decision of virial.py:198-206's accept test)  [3] particles that
 * hit max_trials  [4] screen accepts  [5] screen rejects  [6] particles with the screen
 * disabled  [7] sum of accepted trial numbers (== moments[1] when [3] == 0). */
int cgb_screen_audit(const void *d_workspace, int64_t id_begin, int64_t id_end, uint64_t seed,
                     int64_t max_trials, unsigned long long *d_counts, void *stream);

/* Directions-only pass used for sub_sample > 1 (virial.py:170-176,208-223): particle
 * i takes radius/speed d_radius_sub[i % n_sub], d_speed_sub[i % n_sub] and fresh
 * directions; d_psi_sub (nullable) feeds the potential. */
int cgb_assemble_particles(const cgb_sample_args *args, const double *d_radius_sub,
                           const double *d_speed_sub, const double *d_psi_sub, int64_t n_sub,
                           void *stream);

/* V: mirror of the Cython seam generate_velocities(psi, vesc, fv2esc, knots, coeffs,
 * k, quiet) -> float64[n]  (call site virial.py:198-206).  knots = t[k:-k]
 * (get_knots()), coeffs = c[:n_coeffs] (get_coeffs()), k must be 3.
 * d_u_trial (nullable) = [n][t_cap][2] host-supplied uniforms, else Philox stream
 * CGB_STREAM_VEL_TRIAL with ids first_id + i.  d_trials (nullable): accepted trial
 * count, -1 where no trial within t_cap accepted (speed = NaN there).
 * d_work: >= cgb_generate_velocities_workspace_bytes(n_coeffs). */
size_t cgb_generate_velocities_workspace_bytes(int n_coeffs);
int cgb_generate_velocities(const double *d_psi, const double *d_vesc, const double *d_fv2esc,
                            const double *d_knots, int n_knots, const double *d_coeffs,
                            int n_coeffs, int k, const double *d_u_trial, int64_t t_cap,
                            uint64_t seed, int64_t first_id, double *d_vel, int32_t *d_trials,
                            int64_t n, void *d_work, size_t work_bytes, void *stream);

/* Gs: replaces the numeric part of ClusterModel.generate_gas_particles
 * (model.py:838-947) and generate_tracer_particles (model.py:766-836): radius from
 * the gas-mass CDF, isotropic position, aux splines (thermal energy, density) and
 * optional potential at the radius; velocity planes are zero-filled when d_vel != NULL
 * (d_vel == NULL: the caller owns zeroed storage).  Two particles per thread, 16-byte stores.
 * d_aux_out[k]: nullable f64/f32 arrays (out_dtype) for aux spline k. */
int cgb_sample_gas(const cgb_sample_args *args, void *d_aux_out0, void *d_aux_out1, void *stream);
/* Fast path of cgb_sample_gas.  cgb_sampler_prepare packs the tables of a position-only sampler
 * (d_f_t == NULL) for shared-memory residence when every radial spline shares the radius grid of
 * the CDF, as ClusterModel's tables do (model.py:879-941).  Whether that applied is known on the
 * device; cgb_sampler_gas_flags reads it back ONCE per prepared workspace (it synchronises the
 * stream) and returns the value to put into cgb_sample_args.flags for cgb_sample_gas: with it the
 * call launches one persistent CTA per SM with the tables staged by a TMA bulk copy, without it
 * (flags = 0) the general kernel.  Results are identical either way. */
#define CGB_SAMPLE_GAS_PACKED 1
int cgb_sampler_gas_flags(const void *d_workspace, int32_t *flags_out, void *stream);

/* ------------------------------------------------------------------------ */
/* Multi-halo gas re-sampling ("next" row 1 of SURVEY.md section 8(f))        */
/*   replaces _sample_clusters (particles.py:587-639), the numeric part of    */
/*   combine_two/three_clusters and resample_*_clusters (particles.py:642-795) */
/* ------------------------------------------------------------------------ */
/* Piecewise-polynomial form of a FITPACK cubic spline (knots d_t[n+4], coefficients
 * d_c[n]) in caller memory; what cgb_resample_gas evaluates. */
size_t cgb_ppform_workspace_bytes(int n);
int cgb_ppform_prepare(const double *d_t, const double *d_c, int n, void *d_workspace,
                       size_t workspace_bytes, void *stream);

typedef struct cgb_halo_tables {
    const double *d_dens_pp; /* ppform of IUS(radius, density)            (particles.py:615) */
    const double *d_ener_pp; /* ppform of IUS(radius, 1.5 pressure/density) (particles.py:617-618) */
    int32_t n;               /* radial points of both splines; NULL tables = halo without gas */
    int32_t pad;
    double center[3];        /* kpc     */
    double velocity[3];      /* kpc/Myr */
    double radius_cut;       /* radii[i] of the resample mask (particles.py:597-601) */
} cgb_halo_tables;

/* halos: HOST array of num_halos (1..3) entries.  d_pos / d_vel: (3, ld) planes; d_density,
 * d_thermal: (n,) in/out; d_mass: NULL, or the particle masses to rescale (resample=True);
 * use_radii != 0: only particles within radius_cut of at least one centre are updated. */
int cgb_resample_gas(const cgb_halo_tables *halos, int num_halos, const void *d_pos, int64_t ld,
                     int64_t n, int dtype, void *d_density, void *d_thermal, void *d_vel,
                     void *d_mass, int use_radii, void *stream);

/* Passive scalars of _sample_clusters (particles.py:606-637, 745-795): out_s = sum_k S_ks(r_k)
 * rho_k(r_k) / sum_k rho_k(r_k).  halos as for cgb_resample_gas (only d_dens_pp, n, center and
 * radius_cut are read); d_scalar_pp: HOST array [num_halos * num_scalars] of ppform tables
 * (cgb_ppform_prepare) of IUS(radius, scalar), halo-major, NULL for halos without gas;
 * d_out: HOST array of num_scalars (1..8) device arrays (n,), updated where the mask allows. */
int cgb_resample_scalars(const cgb_halo_tables *halos, int num_halos, const void *const *d_scalar_pp,
                         int num_scalars, const void *d_pos, int64_t ld, int64_t n, int dtype,
                         void *const *d_out, int use_radii, void *stream);

/* Virial-check histograms (north_star item 4): counts of r (log-spaced
 * bins on [r_lo, r_hi)), |v| (linear on [0, v_max)) and E = psi(r) - v^2/2 (linear on
 * [0, e_max)) of a particle set in halo-centric coordinates; out-of-range values fall into the
 * first / last bin.  d_psi_pp: ppform (cgb_ppform_prepare) of the relative-potential spline
 * psi(r) = IUS(radius, -Phi) (virial.py:152) with n_psi points.  d_pos / d_vel: (3, ld) planes.
 * d_hist: uint64[3][nbins] (r, |v|, E), ACCUMULATED (caller zeroes).  Integer counts: the sum
 * over index shards equals the unsharded histogram exactly, so ranks all-reduce them (SUM). */
int cgb_particle_histograms(const double *d_psi_pp, int n_psi, const void *d_pos, const void *d_vel,
                            int64_t ld, int64_t n, int dtype, double r_lo, double r_hi, double v_max,
                            double e_max, int nbins, unsigned long long *d_hist, void *stream);

/* 3-D field -> particles ("next" row 4): tri-linear interpolation of up to three gridded
 * components at the particle positions, zero outside the grid.
 *   replaces ClusterField.map_field_to_particles (fields.py:308-338), i.e.
 *   RegularGridInterpolator((x, y, z), F, bounds_error=False, fill_value=0.0)(positions).
 * d_gx/gy/gz: grid coordinates [nx], [ny], [nz] (ascending); d_fx/fy/fz: C-ordered
 * (nx, ny, nz) components (d_fy, d_fz nullable); d_pos, d_out: (3, ld) / (ncomp, ld) planes. */
int cgb_map_field(const double *d_gx, const double *d_gy, const double *d_gz, int nx, int ny, int nz,
                  const double *d_fx, const double *d_fy, const double *d_fz, const void *d_pos,
                  int64_t ld, int64_t n, int dtype, void *d_out, void *stream);

/* ------------------------------------------------------------------------ */
/* Particle files ("next" row 2): a native writer/reader for the HDF5 subset the reference's
 * particle files use -- h5py/libhdf5 do the same job there:
 *   ClusterParticles.write_particles / from_file            (particles.py:239-274, 337-363):
 *       group per particle type, dataset per field, "units" string attribute (unyt write_hdf5)
 *   ClusterParticles.write_to_gadget_file / from_gadget_file (particles.py:280-335, 453-552):
 *       PartType{0,1,2,4,5} groups, Header group attributes
 * Layout written: superblock v0, symbol-table groups, v1 object headers, contiguous
 * little-endian datasets, variable-length UTF-8 string attributes -- what h5py writes with its
 * default libver, so either side reads the other's files.  Host-only except
 * cgb_h5_write_dataset_device.  Object paths are '/'-separated; parents must exist.  Handles
 * are not thread-safe.  Errors: CGB_E_IO (file), CGB_E_INVALID (argument). */
#define CGB_H5_F32 1
#define CGB_H5_F64 2
#define CGB_H5_I32 3
#define CGB_H5_U32 4
#define CGB_H5_I64 5
#define CGB_H5_U64 6
#define CGB_H5_STR 7 /* attributes only: scalar string (NUL-terminated char* in, bytes out) */

typedef struct cgb_h5_file cgb_h5_file;

int cgb_h5_create(const char *path, int overwrite, cgb_h5_file **out);
int cgb_h5_create_group(cgb_h5_file *f, const char *path);
/* host_data: C-ordered dims[0..rank) elements of dtype; rank 0..4 (0 = scalar dataset) */
int cgb_h5_write_dataset(cgb_h5_file *f, const char *path, int dtype, int rank, const int64_t *dims,
                         const void *host_data);
/* d_src: (ncomp, ld) device planes of src_dtype (f32/f64).  Written as an (n, ncomp) dataset
 * ((n,) if ncomp == 1) of out_dtype with value (out_dtype)((double)x * scale) -- the unit
 * conversion and cast of particles.py:470-472.  Pack kernel + D2H of chunk k+1 overlap the file
 * write of chunk k (two pinned 32 MiB slots owned by the handle).  Returns after the data is
 * in the file. */
int cgb_h5_write_dataset_device(cgb_h5_file *f, const char *path, int out_dtype, const void *d_src,
                                int src_dtype, int64_t n, int ncomp, int64_t ld, double scale,
                                void *stream);
/* count < 0: scalar; count >= 0: 1-D array of count elements.  CGB_H5_STR: data = char*. */
int cgb_h5_set_attr(cgb_h5_file *f, const char *obj_path, const char *name, int dtype, int64_t count,
                    const void *data);
/* writer: lays out the metadata, writes the superblock, frees the handle (also on error) */
int cgb_h5_close(cgb_h5_file *f);

int cgb_h5_open(const char *path, cgb_h5_file **out);
int cgb_h5_num_children(cgb_h5_file *f, const char *group_path, int *n);
int cgb_h5_child_name(cgb_h5_file *f, const char *group_path, int i, char *buf, int buflen,
                      int *is_group);
/* file_offset: absolute byte offset of the contiguous data (-1 for compact / empty datasets) */
int cgb_h5_dataset_info(cgb_h5_file *f, const char *path, int *dtype, int *rank, int64_t *dims4,
                        int64_t *file_offset);
int cgb_h5_read_dataset(cgb_h5_file *f, const char *path, void *host_out, size_t bytes);
/* (n,) or (n, k) f32/f64 dataset -> (k, ld) device planes of dst_dtype, value
 * (dst_dtype)((double)x * scale): the read direction of cgb_h5_write_dataset_device (host copy out
 * of the page cache overlaps upload + unpack of the previous chunk).  Synchronises `stream`. */
int cgb_h5_read_dataset_device(cgb_h5_file *f, const char *path, void *d_dst, int dst_dtype, int64_t ld,
                               double scale, void *stream);
int cgb_h5_num_attrs(cgb_h5_file *f, const char *obj_path, int *n);
int cgb_h5_attr_name(cgb_h5_file *f, const char *obj_path, int i, char *buf, int buflen);
/* dtype 0 = attribute exists but its type is outside the subset; bytes = payload size */
int cgb_h5_attr_info(cgb_h5_file *f, const char *obj_path, const char *name, int *dtype,
                     int64_t *count, size_t *bytes);
int cgb_h5_read_attr(cgb_h5_file *f, const char *obj_path, const char *name, void *out, size_t bytes);

/* ------------------------------------------------------------------------ */
/* Peer-visible device buffers for the fused moment all-reduce (one process per GPU):
 * cgb_peer_alloc: zero-filled device allocation + its 64-byte CUDA IPC handle, which the caller
 * sends to the other ranks (any transport; torch.distributed in the host mirror);
 * cgb_peer_open maps another rank's allocation into this process (peer access over NVLink is
 * enabled as needed); cgb_peer_close / cgb_peer_free undo them. */
int cgb_peer_alloc(size_t bytes, void **d_ptr, unsigned char handle_out[64]);
int cgb_peer_open(const unsigned char handle[64], void **d_ptr);
int cgb_peer_close(void *d_ptr);
int cgb_peer_free(void *d_ptr);

#ifdef __cplusplus
}
#endif
#endif /* CGB200_H */
