// Multi-halo gas re-sampling  -- replaces cluster_generator.particles._sample_clusters
// (particles.py:587-639), the numeric part of combine_two/three_clusters and
// resample_one/two/three_clusters (particles.py:642-795).
//
// For every gas particle: distance to each halo centre, density and thermal energy of
// each halo's radial profile at that distance (cubic interpolating splines of
// (radius, density) and (radius, 1.5 P / density), extrapolating like FITPACK ext=0),
// then total density, mass-weighted specific energy and mass-weighted bulk velocity.
// Elementwise, one thread per particle: 24 B read + 40 B written per particle.
#include <math.h>
#include "cgb_common.cuh"
#include "cgb_spline.cuh"

#define MAX_HALOS 3

struct HaloDev {
    const double* dens;  // pp workspace: brk[n-2 padded to 32] then pp[4 (n-3)]
    const double* ener;
    int n;
    double cx, cy, cz, vx, vy, vz, rcut;
};

struct ResampleParams {
    HaloDev h[MAX_HALOS];
    int num_halos, use_radii, resample_mass;
    const void* pos; long long ld; long long n;
    void* density; void* thermal; void* vel; void* mass;
};

__device__ __forceinline__ int rs_find(const double* __restrict__ brk, int nint, double x)
{
    int lo = 0, hi = nint;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(brk + mid) <= x) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ double rs_eval(const double* __restrict__ ws, int n, int j, double x)
{
    const long long off = ((long long)(n - 2) + 31) / 32 * 32;
    CGB_CHECK_INDEX(j, n - 3);
    const double dx = x - __ldg(ws + j);
    const double* pp = ws + off + 4 * (long long)j;
    const double2 a01 = __ldg((const double2*)pp), a23 = __ldg((const double2*)pp + 1);
    return fma(fma(fma(a23.y, dx, a23.x), dx, a01.y), dx, a01.x);
}

template <typename T>
__global__ void __launch_bounds__(256) cgb_resample_kernel(const __grid_constant__ ResampleParams p)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const T* pos = (const T*)p.pos;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
        const double x = (double)pos[i], y = (double)pos[p.ld + i], z = (double)pos[2 * p.ld + i];
        double dsum = 0.0, esum = 0.0, mx = 0.0, my = 0.0, mz = 0.0;
        bool inside = (p.use_radii == 0);
#pragma unroll
        for (int k = 0; k < MAX_HALOS; ++k) {
            if (k >= p.num_halos) break;
            const HaloDev& h = p.h[k];
            // r = sqrt(((pos - c)**2).sum(axis=1))                       (particles.py:594-596)
            const double dx = __dsub_rn(x, h.cx), dy = __dsub_rn(y, h.cy), dz = __dsub_rn(z, h.cz);
            const double r = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
            if (p.use_radii && r <= h.rcut) inside = true;
            if (h.dens == nullptr) continue;          // halo without gas (particles.py:611-613)
            const int j = rs_find(h.dens, h.n - 3, r);  // both splines share the radius knots
            const double d = rs_eval(h.dens, h.n, j, r);
            const double e = __dmul_rn(rs_eval(h.ener, h.n, j, r), d);
            dsum = __dadd_rn(dsum, d);
            esum = __dadd_rn(esum, e);
            mx = __dadd_rn(mx, __dmul_rn(h.vx, d));
            my = __dadd_rn(my, __dmul_rn(h.vy, d));
            mz = __dadd_rn(mz, __dmul_rn(h.vz, d));
        }
        if (!inside) continue;
        if (p.resample_mass) {
            // vol = mass / density(old); mass = dens * vol                (particles.py:626-628)
            T* m = (T*)p.mass;
            const double vol = (double)m[i] / (double)((const T*)p.density)[i];
            m[i] = (T)__dmul_rn(dsum, vol);
        }
        ((T*)p.density)[i] = (T)dsum;
        ((T*)p.thermal)[i] = (T)(esum / dsum);
        T* v = (T*)p.vel;
        v[i] = (T)(mx / dsum);
        v[p.ld + i] = (T)(my / dsum);
        v[2 * p.ld + i] = (T)(mz / dsum);
    }
}

extern "C" size_t cgb_ppform_workspace_bytes(int n)
{
    if (n < 4) return 0;
    return (size_t)(((long long)(n - 2) + 31) / 32 * 32 + 4LL * (n - 3)) * sizeof(double);
}

extern "C" int cgb_ppform_prepare(const double* d_t, const double* d_c, int n, void* d_workspace,
                                  size_t workspace_bytes, void* stream)
{
    CGB_REQUIRE(n >= 4 && d_t && d_c && d_workspace, "cgb_ppform_prepare: bad arguments");
    if (workspace_bytes < cgb_ppform_workspace_bytes(n) || ((uintptr_t)d_workspace & 15)) {
        cgb_set_error("cgb_ppform_prepare: workspace %zu B (need %zu, 16-byte aligned)", workspace_bytes,
                      cgb_ppform_workspace_bytes(n));
        return CGB_E_WORKSPACE;
    }
    double* ws = (double*)d_workspace;
    return cgb_launch_pp_convert(d_t, d_c, n, ws, ws + ((long long)(n - 2) + 31) / 32 * 32, (cudaStream_t)stream);
}

extern "C" int cgb_resample_gas(const cgb_halo_tables* halos, int num_halos, const void* d_pos, int64_t ld,
                                int64_t n, int dtype, void* d_density, void* d_thermal, void* d_vel,
                                void* d_mass, int use_radii, void* stream)
{
    CGB_REQUIRE(halos != nullptr && num_halos >= 1 && num_halos <= MAX_HALOS,
                "cgb_resample_gas: num_halos must be 1..3, got %d", num_halos);
    CGB_REQUIRE(dtype == CGB_DTYPE_F64 || dtype == CGB_DTYPE_F32, "cgb_resample_gas: bad dtype %d", dtype);
    CGB_REQUIRE(n >= 0 && ld >= n, "cgb_resample_gas: bad sizes");
    if (n == 0) return CGB_OK;
    CGB_REQUIRE(d_pos && d_density && d_thermal && d_vel, "cgb_resample_gas: null pointer");
    ResampleParams p;
    memset(&p, 0, sizeof(p));
    bool any = false;
    for (int k = 0; k < num_halos; ++k) {
        const cgb_halo_tables& h = halos[k];
        CGB_REQUIRE((h.d_dens_pp == nullptr) == (h.d_ener_pp == nullptr),
                    "cgb_resample_gas: halo %d needs both density and energy tables", k);
        if (h.d_dens_pp) CGB_REQUIRE(h.n >= 4, "cgb_resample_gas: halo %d has %d radial points", k, h.n);
        any = any || h.d_dens_pp;
        p.h[k].dens = h.d_dens_pp; p.h[k].ener = h.d_ener_pp; p.h[k].n = h.n;
        p.h[k].cx = h.center[0]; p.h[k].cy = h.center[1]; p.h[k].cz = h.center[2];
        p.h[k].vx = h.velocity[0]; p.h[k].vy = h.velocity[1]; p.h[k].vz = h.velocity[2];
        p.h[k].rcut = h.radius_cut;
    }
    CGB_REQUIRE(any, "cgb_resample_gas: no halo carries a gas profile");
    p.num_halos = num_halos; p.use_radii = use_radii; p.resample_mass = d_mass != nullptr;
    p.pos = d_pos; p.ld = ld; p.n = n;
    p.density = d_density; p.thermal = d_thermal; p.vel = d_vel; p.mass = d_mass;
    const unsigned blocks = (unsigned)min((long long)((n + 255) / 256), 148LL * 16);
    if (dtype == CGB_DTYPE_F64) cgb_resample_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
    else cgb_resample_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
    CGB_CUDA_CHECK(cudaGetLastError());
    return CGB_OK;
}

// Passive scalars of _sample_clusters (particles.py:606-637): for every scalar s,
//   out_s = sum_k S_ks(r_k) rho_k(r_k) / sum_k rho_k(r_k)
// with S_ks the cubic spline of halo k's radial profile of that scalar.  Same per-halo radii,
// densities and mask as cgb_resample_gas; the particle's own fields are not read, so the call
// may run before or after it.
#define MAX_SCALARS 8
struct ScalarParams {
    HaloDev h[MAX_HALOS];
    const double* sc[MAX_HALOS][MAX_SCALARS];
    void* out[MAX_SCALARS];
    int num_halos, num_scalars, use_radii;
    const void* pos; long long ld; long long n;
};

template <typename T>
__global__ void __launch_bounds__(256) cgb_resample_scalars_kernel(const __grid_constant__ ScalarParams p)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const T* pos = (const T*)p.pos;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
        const double x = (double)pos[i], y = (double)pos[p.ld + i], z = (double)pos[2 * p.ld + i];
        double dsum = 0.0;
        double acc[MAX_SCALARS];
#pragma unroll
        for (int s = 0; s < MAX_SCALARS; ++s) acc[s] = 0.0;
        bool inside = (p.use_radii == 0);
#pragma unroll
        for (int k = 0; k < MAX_HALOS; ++k) {
            if (k >= p.num_halos) break;
            const HaloDev& h = p.h[k];
            const double dx = __dsub_rn(x, h.cx), dy = __dsub_rn(y, h.cy), dz = __dsub_rn(z, h.cz);
            const double r = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
            if (p.use_radii && r <= h.rcut) inside = true;
            if (h.dens == nullptr) continue;
            const int j = rs_find(h.dens, h.n - 3, r);
            const double d = rs_eval(h.dens, h.n, j, r);
            dsum = __dadd_rn(dsum, d);
#pragma unroll
            for (int s = 0; s < MAX_SCALARS; ++s)
                if (s < p.num_scalars)  // s[i, j, :] = get_scalar(r[i, :]) * d[i, :]   (particles.py:621-623)
                    acc[s] = __dadd_rn(acc[s], __dmul_rn(rs_eval(p.sc[k][s], h.n, j, r), d));
        }
        if (!inside) continue;
#pragma unroll
        for (int s = 0; s < MAX_SCALARS; ++s)
            if (s < p.num_scalars) ((T*)p.out[s])[i] = (T)(acc[s] / dsum);
    }
}

extern "C" int cgb_resample_scalars(const cgb_halo_tables* halos, int num_halos,
                                    const void* const* d_scalar_pp, int num_scalars, const void* d_pos,
                                    int64_t ld, int64_t n, int dtype, void* const* d_out, int use_radii,
                                    void* stream)
{
    CGB_REQUIRE(halos != nullptr && num_halos >= 1 && num_halos <= MAX_HALOS,
                "cgb_resample_scalars: num_halos must be 1..3, got %d", num_halos);
    CGB_REQUIRE(num_scalars >= 1 && num_scalars <= MAX_SCALARS, "cgb_resample_scalars: 1..%d scalars, got %d",
                MAX_SCALARS, num_scalars);
    CGB_REQUIRE(dtype == CGB_DTYPE_F64 || dtype == CGB_DTYPE_F32, "cgb_resample_scalars: bad dtype %d", dtype);
    CGB_REQUIRE(n >= 0 && ld >= n, "cgb_resample_scalars: bad sizes");
    if (n == 0) return CGB_OK;
    CGB_REQUIRE(d_pos && d_scalar_pp && d_out, "cgb_resample_scalars: null pointer");
    ScalarParams p;
    memset(&p, 0, sizeof(p));
    bool any = false;
    for (int k = 0; k < num_halos; ++k) {
        const cgb_halo_tables& h = halos[k];
        if (h.d_dens_pp) CGB_REQUIRE(h.n >= 4, "cgb_resample_scalars: halo %d has %d radial points", k, h.n);
        any = any || h.d_dens_pp;
        p.h[k].dens = h.d_dens_pp; p.h[k].n = h.n;
        p.h[k].cx = h.center[0]; p.h[k].cy = h.center[1]; p.h[k].cz = h.center[2];
        p.h[k].rcut = h.radius_cut;
        for (int s = 0; s < num_scalars; ++s) {
            p.sc[k][s] = (const double*)d_scalar_pp[k * num_scalars + s];
            CGB_REQUIRE(!h.d_dens_pp || p.sc[k][s], "cgb_resample_scalars: halo %d lacks the table of scalar %d", k, s);
        }
    }
    CGB_REQUIRE(any, "cgb_resample_scalars: no halo carries a gas profile");
    for (int s = 0; s < num_scalars; ++s) {
        CGB_REQUIRE(d_out[s] != nullptr, "cgb_resample_scalars: null output %d", s);
        p.out[s] = d_out[s];
    }
    p.num_halos = num_halos; p.num_scalars = num_scalars; p.use_radii = use_radii;
    p.pos = d_pos; p.ld = ld; p.n = n;
    const unsigned blocks = (unsigned)min((long long)((n + 255) / 256), 148LL * 16);
    if (dtype == CGB_DTYPE_F64) cgb_resample_scalars_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
    else cgb_resample_scalars_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
    CGB_CUDA_CHECK(cudaGetLastError());
    return CGB_OK;
}

// ---------------------------------------------------------------------------
// 3-D field -> particles, tri-linear  -- replaces ClusterField.map_field_to_particles
// (fields.py:308-338), i.e. scipy.interpolate.RegularGridInterpolator(method="linear",
// bounds_error=False, fill_value=0.0) evaluated at the particle positions, for the three
// vector components at once.  Gather-bound (8 corner loads per component).
// ---------------------------------------------------------------------------
struct MapFieldParams {
    const double *gx, *gy, *gz;
    int nx, ny, nz;
    const double* comp[3];
    int ncomp;
    const void* pos; long long ld; long long n;
    void* out;  // (ncomp, ld) planes
};

// RegularGridInterpolator._find_indices: i = searchsorted(grid, x) - 1 clipped to [0, n-2],
// norm distance (x - grid[i]) / (grid[i+1] - grid[i]); out of bounds if x outside [g0, g_last]
__device__ __forceinline__ bool grid_locate(const double* __restrict__ g, int n, double x, int& i, double& t)
{
    int lo = 0, hi = n;  // first index with g[idx] >= x  (searchsorted side='left')
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(g + mid) < x) lo = mid + 1; else hi = mid;
    }
    i = min(max(lo - 1, 0), n - 2);
    const double g0 = __ldg(g + i), g1 = __ldg(g + i + 1);
    t = __ddiv_rn(__dsub_rn(x, g0), __dsub_rn(g1, g0));
    return !(x < __ldg(g) || x > __ldg(g + n - 1));
}

template <typename T>
__global__ void __launch_bounds__(256) cgb_map_field_kernel(const __grid_constant__ MapFieldParams p)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const T* pos = (const T*)p.pos;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
        int ix, iy, iz;
        double tx, ty, tz;
        bool in = grid_locate(p.gx, p.nx, (double)pos[i], ix, tx);
        in = grid_locate(p.gy, p.ny, (double)pos[p.ld + i], iy, ty) && in;
        in = grid_locate(p.gz, p.nz, (double)pos[2 * p.ld + i], iz, tz) && in;
        for (int c = 0; c < p.ncomp; ++c) {
            double v = 0.0;
            if (in) {
                const double* f = p.comp[c];
                // corners in itertools.product order, weight = ((1*wx)*wy)*wz, sequential sum
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int b = 0; b < 2; ++b)
#pragma unroll
                        for (int d = 0; d < 2; ++d) {
                            const double wx = a ? tx : __dsub_rn(1.0, tx);
                            const double wy = b ? ty : __dsub_rn(1.0, ty);
                            const double wz = d ? tz : __dsub_rn(1.0, tz);
                            const double w = __dmul_rn(__dmul_rn(wx, wy), wz);
                            const size_t idx = ((size_t)(ix + a) * p.ny + (size_t)(iy + b)) * p.nz + (size_t)(iz + d);
                            CGB_CHECK(ix + a < p.nx && iy + b < p.ny && iz + d < p.nz && ix >= 0 && iy >= 0 && iz >= 0);
                            v = __dadd_rn(v, __dmul_rn(w, __ldg(f + idx)));
                        }
            }
            ((T*)p.out)[(long long)c * p.ld + i] = (T)v;
        }
    }
}

extern "C" int cgb_map_field(const double* d_gx, const double* d_gy, const double* d_gz, int nx, int ny,
                             int nz, const double* d_fx, const double* d_fy, const double* d_fz,
                             const void* d_pos, int64_t ld, int64_t n, int dtype, void* d_out, void* stream)
{
    CGB_REQUIRE(nx >= 2 && ny >= 2 && nz >= 2 && d_gx && d_gy && d_gz, "cgb_map_field: grid needs >= 2 points per axis");
    CGB_REQUIRE(d_fx != nullptr, "cgb_map_field: null field");
    CGB_REQUIRE(dtype == CGB_DTYPE_F64 || dtype == CGB_DTYPE_F32, "cgb_map_field: bad dtype %d", dtype);
    CGB_REQUIRE(n >= 0 && ld >= n, "cgb_map_field: bad sizes");
    if (n == 0) return CGB_OK;
    CGB_REQUIRE(d_pos && d_out, "cgb_map_field: null pointer");
    MapFieldParams p;
    p.gx = d_gx; p.gy = d_gy; p.gz = d_gz; p.nx = nx; p.ny = ny; p.nz = nz;
    p.comp[0] = d_fx; p.comp[1] = d_fy; p.comp[2] = d_fz;
    p.ncomp = d_fz ? 3 : (d_fy ? 2 : 1);
    p.pos = d_pos; p.ld = ld; p.n = n; p.out = d_out;
    const unsigned blocks = (unsigned)min((long long)((n + 255) / 256), 148LL * 16);
    if (dtype == CGB_DTYPE_F64) cgb_map_field_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
    else cgb_map_field_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
    CGB_CUDA_CHECK(cudaGetLastError());
    return CGB_OK;
}

// ---------------------------------------------------------------------------
// virial-check histograms: r, |v| and E = psi(r) - v^2/2 of a particle set
// ---------------------------------------------------------------------------
struct HistParams {
    const double* psi;  // ppform workspace of psi(r)
    int n_psi, nbins;
    const void* pos; const void* vel; long long ld, n;
    double r_lo, inv_dlog, inv_dv, inv_de;
    unsigned long long* hist;
};

// HBM-bound (48 B read per fp64 particle); per-CTA shared-memory counters, one flush per CTA
template <typename T>
__global__ void __launch_bounds__(256) cgb_hist_kernel(const __grid_constant__ HistParams p)
{
    extern __shared__ unsigned int s_h[];  // [3][nbins]
    const int nb = p.nbins;
    for (int i = threadIdx.x; i < 3 * nb; i += blockDim.x) s_h[i] = 0u;
    __syncthreads();
    const T* pos = (const T*)p.pos;
    const T* vel = (const T*)p.vel;
    const int nint = p.n_psi - 3;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
        const double x = (double)pos[i], y = (double)pos[p.ld + i], z = (double)pos[2 * p.ld + i];
        const double vx = (double)vel[i], vy = (double)vel[p.ld + i], vz = (double)vel[2 * p.ld + i];
        const double r = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
        const double v2 = __dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz));
        const double psi = rs_eval(p.psi, p.n_psi, rs_find(p.psi, nint, r), r);
        const double e = __dsub_rn(psi, __dmul_rn(0.5, v2));
        // out-of-range (and NaN) values land in the end bins
        const int br = min(max(__double2int_rd(__dmul_rn(log(r / p.r_lo), p.inv_dlog)), 0), nb - 1);
        const int bv = min(max(__double2int_rd(__dmul_rn(sqrt(v2), p.inv_dv)), 0), nb - 1);
        const int be = min(max(__double2int_rd(__dmul_rn(e, p.inv_de)), 0), nb - 1);
        atomicAdd(s_h + br, 1u);
        atomicAdd(s_h + nb + bv, 1u);
        atomicAdd(s_h + 2 * nb + be, 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * nb; i += blockDim.x)
        if (s_h[i]) atomicAdd(p.hist + i, (unsigned long long)s_h[i]);
}

extern "C" int cgb_particle_histograms(const double* d_psi_pp, int n_psi, const void* d_pos, const void* d_vel,
                                       int64_t ld, int64_t n, int dtype, double r_lo, double r_hi, double v_max,
                                       double e_max, int nbins, unsigned long long* d_hist, void* stream)
{
    CGB_REQUIRE(d_psi_pp && n_psi >= 4, "cgb_particle_histograms: psi spline needs >= 4 points");
    CGB_REQUIRE(nbins >= 1 && nbins <= 4096, "cgb_particle_histograms: nbins must be 1..4096");
    CGB_REQUIRE(r_lo > 0.0 && r_hi > r_lo && v_max > 0.0 && e_max > 0.0, "cgb_particle_histograms: bad ranges");
    CGB_REQUIRE(n >= 0 && ld >= n && d_hist, "cgb_particle_histograms: bad shape");
    CGB_REQUIRE(dtype == CGB_DTYPE_F64 || dtype == CGB_DTYPE_F32, "cgb_particle_histograms: dtype %d", dtype);
    if (n == 0) return CGB_OK;
    CGB_REQUIRE(d_pos && d_vel, "cgb_particle_histograms: null particle buffers");
    HistParams p;
    p.psi = d_psi_pp; p.n_psi = n_psi; p.nbins = nbins;
    p.pos = d_pos; p.vel = d_vel; p.ld = ld; p.n = n;
    p.r_lo = r_lo; p.inv_dlog = (double)nbins / log(r_hi / r_lo);
    p.inv_dv = (double)nbins / v_max; p.inv_de = (double)nbins / e_max;
    p.hist = d_hist;
    const unsigned blocks = (unsigned)min((long long)((n + 255) / 256), 148LL * 8);
    const size_t smem = (size_t)3 * nbins * sizeof(unsigned int);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == CGB_DTYPE_F32) cgb_hist_kernel<float><<<blocks, 256, smem, s>>>(p);
    else cgb_hist_kernel<double><<<blocks, 256, smem, s>>>(p);
    CGB_CUDA_CHECK(cudaGetLastError());
    return CGB_OK;
}
