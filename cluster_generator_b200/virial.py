"""
Equilibrium sampler: Eddington f(E), virial check, DM/star particle generation.

Drop-in mirror of ``cluster_generator.virial.VirialEquilibrium`` (virial.py:16-236):
same constructor, ``f / df / ee / ff``, ``check_virial()`` and
``generate_particles()`` with the reference's argument meaning and error behaviour.
All numerics run in hand-written sm_100a CUDA behind the C ABI of
``include/cgb200.h``; there is no CPU fallback.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from . import _native as N
from .particles import ClusterParticles
from .utils import mylog, parse_prng, seed_from_prng, unit_array

#: Philox key used for the speed trials when the caller asks for numpy-ordered
#: uniforms (rng="numpy"): like the reference's Cython loop, the trial stream is then
#: private and does not consume the caller's ``prng``.
PRIVATE_TRIAL_SEED = 0xC6A4A7935BD1E995

_SPECIES = {"dark_matter": 1, "stellar": 2, "gas": 3, "tracer": 4}

#: Spline solver of the f(E) build: "fitpack" (default) reproduces scipy's numbers to the
#: bit (fpcurf's Givens solve, ~0.85 ms per f(E)); "pcr" is ~2.6x faster and within 1e-11.
FE_SPLINE_FIT = "fitpack"


class DeviceSpline:
    """Cubic interpolating spline living on the GPU.

    Stands in for ``scipy.interpolate.InterpolatedUnivariateSpline`` where the
    reference keeps one as ``VirialEquilibrium.f`` (virial.py:41,67): callable with
    ``(x, nu=0|1)``, ``get_knots()``, ``get_coeffs()`` and ``_eval_args``.
    """

    def __init__(self, t: torch.Tensor, c: torch.Tensor):
        self.t, self.c = t, c
        self.n = c.numel()

    @classmethod
    def fit(cls, x, y, method="pcr") -> "DeviceSpline":
        """IUS(x, y): raises ValueError unless x is strictly increasing (as scipy does).
        ``method="fitpack"``: FITPACK's own Givens solve, coefficients bit-equal to scipy's;
        ``"pcr"``: parallel cyclic reduction, ~15x faster, within 1e-13."""
        if method not in ("fitpack", "pcr"):
            raise ValueError("method must be 'fitpack' or 'pcr'")
        xh = np.asarray(x, dtype=np.float64)
        yh = np.asarray(y, dtype=np.float64)
        if xh.ndim != 1 or xh.shape != yh.shape:
            raise ValueError("x and y should be 1-d arrays of the same length")
        if xh.size < 4:
            raise ValueError("need at least 4 points for a cubic spline")
        if not np.all(np.diff(xh) > 0.0):
            raise ValueError("x must be strictly increasing")
        if not (np.all(np.isfinite(xh)) and np.all(np.isfinite(yh))):
            raise ValueError("x and y must be finite")
        dev = N.require_cuda()
        n = xh.size
        xd, yd = N.dev_f64(xh, dev), N.dev_f64(yh, dev)
        t = torch.empty(n + 4, dtype=torch.float64, device=dev)
        c = torch.empty(n, dtype=torch.float64, device=dev)
        nb = N.lib().cgb_spline_fit_workspace_bytes(n)
        work = N.workspace(nb, dev)
        N.check(N.lib().cgb_spline_fit(N.ptr(xd), N.ptr(yd), n, N.ptr(t), N.ptr(c), N.ptr(work), nb,
                                       N.CGB_FIT_FITPACK if method == "fitpack" else N.CGB_FIT_PCR,
                                       N.stream_ptr()))
        return cls(t, c)

    def __call__(self, x, nu=0):
        dev = self.t.device
        xh = np.atleast_1d(np.asarray(x, dtype=np.float64))
        xd = N.dev_f64(xh.ravel(), dev)
        yd = torch.empty_like(xd)
        N.check(N.lib().cgb_spline_eval(N.ptr(self.t), N.ptr(self.c), self.n, N.ptr(xd), N.ptr(yd),
                                        xd.numel(), int(nu), N.stream_ptr()))
        y = yd.cpu().numpy().reshape(xh.shape)
        return y if np.ndim(x) else float(y[0])

    def get_knots(self):
        return self.t[3:-3].cpu().numpy()

    def get_coeffs(self):
        return self.c.cpu().numpy()

    @property
    def _eval_args(self):
        return self.t.cpu().numpy(), self.c.cpu().numpy(), 3


class SamplerTables:
    """Device tables of one (species, r_max): CDF, psi(r), f(E), screen; see
    ``cgb_sampler_prepare``."""

    def __init__(self, P_r, r_cdf, mtot, psi_spline: DeviceSpline, f_spline: DeviceSpline | None,
                 aux: list[DeviceSpline] | None = None):
        dev = N.require_cuda()
        self.mtot = float(mtot)
        self.P_r, self.r_cdf = N.dev_f64(P_r, dev), N.dev_f64(r_cdf, dev)
        self.psi_spline, self.f_spline, self.aux = psi_spline, f_spline, aux or []
        d = N.SamplerDesc()
        d.d_cdf_P, d.d_cdf_r, d.n_cdf = self.P_r.data_ptr(), self.r_cdf.data_ptr(), self.P_r.numel()
        d.n_psi, d.d_psi_t, d.d_psi_c = psi_spline.n, psi_spline.t.data_ptr(), psi_spline.c.data_ptr()
        if f_spline is not None:
            d.n_f, d.d_f_t, d.d_f_c = f_spline.n, f_spline.t.data_ptr(), f_spline.c.data_ptr()
        d.n_aux = len(self.aux)
        for k, s in enumerate(self.aux):
            d.n_aux_coef = s.n
            d.d_aux_t[k], d.d_aux_c[k] = s.t.data_ptr(), s.c.data_ptr()
        nb = N.lib().cgb_sampler_workspace_bytes(N.ctypes.byref(d))
        if nb == 0:
            N.check(N.CGB_E_INVALID)
        self.workspace = N.workspace(nb, dev)
        N.check(N.lib().cgb_sampler_prepare(N.ctypes.byref(d), N.ptr(self.workspace), nb, N.stream_ptr()))
        self.scratch = N.workspace(N.lib().cgb_sample_scratch_bytes(), dev)
        #: cgb_sample_args.flags for cgb_sample_gas (packed shared-memory tables), read back once
        self.gas_flags = 0
        if f_spline is None:
            fl = N.ctypes.c_int32(0)
            N.check(N.lib().cgb_sampler_gas_flags(N.ptr(self.workspace), N.ctypes.byref(fl), N.stream_ptr()))
            self.gas_flags = int(fl.value)


def radius_cdf_table(r, m, r_max=None):
    """The (P_r, r) table of generate_particle_radii (utils.py:155-164), built on the
    host from the N-point model arrays (the per-particle draw happens on the GPU)."""
    ridx = r.size if r_max is None else int(np.searchsorted(r, r_max))
    if ridx < 1:
        raise ValueError(f"r_max = {r_max} lies inside the innermost radius {r[0]}")
    mtot = m[ridx - 1]
    P_r = np.insert(m[:ridx], 0, 0.0)
    P_r /= P_r[-1]
    rr = np.insert(r[:ridx], 0, 0.0)
    return P_r, rr, float(mtot)


def shard_bounds(n: int, shard):
    """Contiguous index range of rank k of G: [k n / G, (k+1) n / G)."""
    if shard is None:
        return 0, n
    rank, world = shard
    if not (0 <= rank < world):
        raise ValueError(f"bad shard {shard}")
    return (rank * n) // world, ((rank + 1) * n) // world


def _out_code(dtype):
    if dtype == torch.float64:
        return N.CGB_DTYPE_F64
    if dtype == torch.float32:
        return N.CGB_DTYPE_F32
    raise ValueError(f"dtype must be torch.float64 or torch.float32, got {dtype}")


class VirialEquilibrium:
    def __init__(self, model, ptype="dark_matter", df=None):
        r"""
        Generate a virial equilibrium model from a profile (virial.py:17-41).

        Parameters
        ----------
        model : ClusterModel
        ptype : "dark_matter" or "stellar"
        df : array, optional
            The particle distribution function in radius order. If not supplied, it
            is generated on the GPU.
        """
        if ptype not in ("dark_matter", "stellar"):
            raise KeyError(ptype)
        self.num_elements = model.num_elements
        self.ptype = ptype
        self.model = model
        self._tables = {}
        self._psi_pp = None
        self._g_dev = self._last_dev = self._ier_dev = None
        if df is None:
            self._generate_df()
        else:
            self.df = unit_array(np.asarray(df, dtype=np.float64), "Msun*Myr**3/kpc**6")
            self.f = DeviceSpline.fit(self.ee, self.df.d[::-1], method=FE_SPLINE_FIT)

    def _generate_df(self):
        """Eddington inversion on the device (virial.py:43-68)."""
        n = self.num_elements
        ee = np.ascontiguousarray(self.ee)
        pden = np.ascontiguousarray(self.model[f"{self.ptype}_density"].d[::-1])
        if not np.all(np.diff(ee) > 0.0):
            raise ValueError("x must be strictly increasing")
        dev = N.require_cuda()
        L = N.lib()
        both = N.dev_f64(np.stack([ee, pden]), dev)       # one host -> device copy
        ee_d, pd_d = both[0], both[1]
        out = torch.empty(4 * n + 4, dtype=torch.float64, device=dev)
        g, ff, fc, ft = out[:n], out[n:2 * n], out[2 * n:3 * n], out[3 * n:]
        flags = torch.empty(2 * n, dtype=torch.int32, device=dev)
        last, ier = flags[:n], flags[n:]
        nb = L.cgb_eddington_workspace_bytes(n)
        method = N.CGB_FIT_FITPACK if FE_SPLINE_FIT == "fitpack" else N.CGB_FIT_PCR
        # the species of one model share the energy grid: the workspace -- with the Givens
        # factorisation of that grid inside -- is kept on the model and re-used
        cache = getattr(self.model, "_fe_fit_cache", None)
        key = (n, ee.tobytes(), str(dev))
        if method == N.CGB_FIT_FITPACK and cache is not None and cache[0] == key:
            work = cache[1]
            method |= N.CGB_FIT_REUSE_FACTOR
        else:
            work = N.workspace(nb, dev)
        N.check(L.cgb_eddington_df(N.ptr(ee_d), N.ptr(pd_d), n, 1.49e-05, 1.49e-05, N.ptr(g), N.ptr(ff),
                                   N.ptr(ft), N.ptr(fc), N.ptr(last), N.ptr(ier), N.ptr(work), nb,
                                   method, N.stream_ptr()))
        if (method & ~N.CGB_FIT_REUSE_FACTOR) == N.CGB_FIT_FITPACK:
            try:
                self.model._fe_fit_cache = (key, work)
            except AttributeError:
                pass
        self.f = DeviceSpline(ft, fc)
        # g and QUADPACK's per-bin diagnostics stay on the device until somebody asks for them
        self._g_dev, self._last_dev, self._ier_dev = g, last, ier
        # ... and so does f(E) itself: the sampler reads the device spline; `df` / `ff` copy on demand
        self._df, self._ff_dev = None, ff

    @property
    def df(self):
        """f(E) in radius order (virial.py:68), a host array like the reference's; copied from the
        device the first time it is asked for."""
        if self._df is None and getattr(self, "_ff_dev", None) is not None:
            self._df = unit_array(self._ff_dev.cpu().numpy()[::-1].copy(), "Msun*Myr**3/kpc**6")
        return self._df

    @df.setter
    def df(self, value):
        self._df, self._ff_dev = value, None

    @property
    def g(self):
        """g(E_i) = int_0^sqrt(E_i) 2 rho'(E_i - t^2) dt, the N quadratures of virial.py:54-63."""
        return None if self._g_dev is None else self._g_dev.cpu().numpy()

    @property
    def qags_last(self):
        """QUADPACK's number of sub-intervals per energy bin."""
        return None if self._last_dev is None else self._last_dev.cpu().numpy()

    @property
    def qags_ier(self):
        """QUADPACK's error flag per energy bin."""
        return None if self._ier_dev is None else self._ier_dev.cpu().numpy()

    @property
    def ee(self):
        """Relative potential in ascending order (virial.py:70-72)."""
        return -self.model["gravitational_potential"].d[::-1]

    @property
    def ff(self):
        """f(E) in ascending-E order (virial.py:74-76)."""
        return self.df.d[::-1]

    def check_virial(self):
        r"""
        Density profile implied by f(E) and its relative deviation from the input
        profile (virial.py:78-106).  Returns ``(rho, chk)`` in radius order.
        """
        n = self.num_elements
        dev = N.require_cuda()
        ee_d = N.dev_f64(np.ascontiguousarray(self.ee), dev)
        pd_d = N.dev_f64(np.ascontiguousarray(self.model[f"{self.ptype}_density"].d[::-1]), dev)
        rho = torch.empty(n, dtype=torch.float64, device=dev)
        chk = torch.empty(n, dtype=torch.float64, device=dev)
        N.check(N.lib().cgb_check_virial(N.ptr(ee_d), N.ptr(pd_d), n, N.ptr(self.f.t), N.ptr(self.f.c),
                                         N.ptr(rho), N.ptr(chk), None, None, N.stream_ptr()))
        rho, chk = rho.cpu().numpy(), chk.cpu().numpy()
        mylog.info(
            "The maximum relative deviation of this profile from virial equilibrium is %g",
            np.abs(chk).max(),
        )
        return rho, chk

    # ------------------------------------------------------------------ #
    def particle_histograms(self, particles, nbins=256, all_reduce=False):
        """
        Virial-check histograms of this species' particles (north_star item 4): counts of r
        (``nbins`` log-spaced bins over the model's radial range), |v| (linear up to the largest
        escape speed) and E = psi(r) - v^2/2 (linear up to psi_max), computed on the device
        (``cgb_particle_histograms``).  The particles must be in halo-centric coordinates.
        Counts are integers, so summing the histograms of index shards reproduces the unsharded
        histogram exactly; ``all_reduce=True`` sums them over the default process group (NCCL).

        Returns a dict: ``r``, ``v``, ``E`` (int64 CUDA tensors), ``r_edges``, ``v_edges``,
        ``E_edges`` (NumPy), ``n``.
        """
        key = {"dark_matter": "dm", "stellar": "star"}[self.ptype]
        pos, vel = particles[key, "particle_position"], particles[key, "particle_velocity"]
        ppos, pvel = pos.T, vel.T
        if not (ppos.is_contiguous() and pvel.is_contiguous()):
            ppos, pvel = ppos.contiguous(), pvel.contiguous()
        n = pos.shape[0]
        if getattr(self, "_psi_pp", None) is None:
            sp = DeviceSpline.fit(self.model["radius"].d, self.ee[::-1])
            nb = N.lib().cgb_ppform_workspace_bytes(sp.n)
            ws = N.workspace(nb)
            N.check(N.lib().cgb_ppform_prepare(N.ptr(sp.t), N.ptr(sp.c), sp.n, N.ptr(ws), nb, N.stream_ptr()))
            self._psi_pp = (ws, sp.n)
        ws, n_psi = self._psi_pp
        rr = self.model["radius"].d
        r_lo, r_hi = float(rr[0]), float(rr[-1])
        e_max = float(self.ee.max())
        v_max = float(np.sqrt(2.0 * e_max))
        hist = torch.zeros((3, nbins), dtype=torch.int64, device=pos.device)
        N.check(N.lib().cgb_particle_histograms(
            N.ptr(ws), n_psi, N.ptr(ppos), N.ptr(pvel), ppos.stride(0) if n else 0, n, _out_code(pos.dtype),
            r_lo, r_hi, v_max, e_max, nbins, N.ptr(hist), N.stream_ptr()))
        if all_reduce:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                dist.all_reduce(hist, op=dist.ReduceOp.SUM)
                nn = torch.tensor([n], dtype=torch.int64, device=pos.device)
                dist.all_reduce(nn, op=dist.ReduceOp.SUM)
                n = int(nn.item())
        return {
            "r": hist[0], "v": hist[1], "E": hist[2], "n": n,
            "r_edges": r_lo * np.exp(np.arange(nbins + 1) * (np.log(r_hi / r_lo) / nbins)),
            "v_edges": np.arange(nbins + 1) * (v_max / nbins),
            "E_edges": np.arange(nbins + 1) * (e_max / nbins),
        }

    def radial_histogram_chi2(self, hist, r_max=None):
        """Chi-square of the radial histogram against the distribution the radius draw samples
        exactly: P(r) is piecewise linear between the enclosed-mass table's nodes
        (utils.py:154-166).  Returns ``(chi2, dof)`` over the bins with >= 5 expected counts."""
        density, mass = f"{self.ptype}_density", f"{self.ptype}_mass"
        nonzero = self.model[density].d > 0.0
        P_r, r_cdf, _ = radius_cdf_table(self.model["radius"].d[nonzero], self.model[mass].d[nonzero], r_max)
        cdf = np.interp(hist["r_edges"], r_cdf, P_r, left=0.0, right=1.0)
        cdf[0], cdf[-1] = 0.0, 1.0                       # the end bins collect everything outside
        expected = hist["n"] * np.diff(cdf)
        obs = hist["r"].cpu().numpy().astype(np.float64)
        ok = expected >= 5.0
        chi2 = float((((obs - expected) ** 2)[ok] / expected[ok]).sum())
        return chi2, int(ok.sum()) - 1

    # ------------------------------------------------------------------ #
    def sampler_tables(self, r_max=None) -> SamplerTables:
        """Device tables for this species and r_max (cached)."""
        key = None if r_max is None else float(r_max)
        if key not in self._tables:
            density = f"{self.ptype}_density"
            mass = f"{self.ptype}_mass"
            nonzero = self.model[density].d > 0.0  # virial.py:161
            P_r, r_cdf, mtot = radius_cdf_table(
                self.model["radius"].d[nonzero], self.model[mass].d[nonzero], r_max
            )
            # energy_spline = IUS(radius, psi)                        (virial.py:152-154)
            psi_spline = DeviceSpline.fit(self.model["radius"].d, self.ee[::-1])
            self._tables[key] = SamplerTables(P_r, r_cdf, mtot, psi_spline, self.f)
        return self._tables[key]

    def screen_audit(self, num_particles, r_max=None, seed=0, max_trials=1 << 20, first_id=0):
        """Diagnostics: run the fp32 interval screen and the exact fp64 accept test side by
        side on EVERY trial of ``num_particles`` particles (``cgb_screen_audit``) and return
        the counters as a dict; ``contradictions`` must be 0."""
        dev = N.require_cuda()
        tabs = self.sampler_tables(r_max)
        counts = torch.zeros(8, dtype=torch.int64, device=dev)
        N.check(N.lib().cgb_screen_audit(N.ptr(tabs.workspace), int(first_id), int(first_id) + int(num_particles),
                                         int(seed) & 0xFFFFFFFFFFFFFFFF, int(max_trials), N.ptr(counts),
                                         N.stream_ptr()))
        names = ("trials", "undecided", "contradictions", "capped_particles", "screen_accepts",
                 "screen_rejects", "screen_disabled_particles", "sum_accepted_trials")
        return dict(zip(names, (int(v) for v in counts.cpu())))

    def generate_particles(
        self,
        num_particles,
        r_max=None,
        sub_sample=1,
        compute_potential=False,
        prng=None,
        *,
        dtype=torch.float64,
        rng="device",
        shard=None,
        diagnostics=True,
        peer_moments=None,
        host=False,
        chunk=None,
        wait=True,
    ):
        """
        Generate a set of dark matter or star particles in virial equilibrium
        (virial.py:108-236).

        Parameters (reference signature)
        --------------------------------
        num_particles, r_max, sub_sample, compute_potential, prng : as the reference.

        Extensions (keyword-only)
        -------------------------
        dtype : torch.float64 (default) or torch.float32 output buffers.
        rng : "device" (default) -- every uniform comes from Philox4x32-10 keyed by a
            seed drawn from ``prng``; "numpy" -- radii and directions use uniforms
            drawn from ``prng`` in the reference's order (virial.py:162,175-176,
            213-214), so those fields equal the reference's for the same ``prng``.
        shard : (rank, world) -- generate only this rank's contiguous index range;
            output is bit-identical to the corresponding slice of a single-GPU run.
        peer_moments : :class:`~cluster_generator_b200.peer.PeerMoments` -- fuse the moment
            all-reduce into the kernel (sub_sample == 1): the diagnostics hold the global
            sums after ``peer_moments.fence()``, no NCCL call follows the kernel.
        host : True, or a host-resident ClusterParticles returned by an earlier ``host=True``
            call (its pinned buffers are re-used) -- return the fields as pinned host tensors
            like the reference's host arrays.  The index range is generated in ``chunk``-sized
            pieces whose device->host copies overlap the next piece's kernels
            (:mod:`cluster_generator_b200.hoststream`); values are bit-identical to the
            device path.
        wait : with ``host``: False returns while the last copies are still in flight, so the
            next species' table preparation and first chunks overlap them; the arrays are
            complete after ``particles.wait()`` (``particles[key]`` calls it).
        """
        num_particles = int(num_particles)
        sub_sample = int(sub_sample)
        if num_particles < 0:
            raise ValueError("num_particles must be non-negative")
        if sub_sample < 1:
            raise ValueError("sub_sample must be >= 1")
        if num_particles % sub_sample != 0:
            # the reference fails later with a broadcasting ValueError (virial.py:182)
            raise ValueError("num_particles must be a multiple of sub_sample")
        if rng not in ("device", "numpy"):
            raise ValueError("rng must be 'device' or 'numpy'")
        num_particles_sub = num_particles // sub_sample
        key = {"dark_matter": "dm", "stellar": "star"}[self.ptype]
        prng = parse_prng(prng)
        mylog.info("We will be assigning %s %s particles.", num_particles, self.ptype)

        dev = N.require_cuda()
        L = N.lib()
        tabs = self.sampler_tables(r_max)
        code = _out_code(dtype)
        lo, hi = shard_bounds(num_particles, shard)
        n_loc = hi - lo

        hu = {}
        if rng == "numpy":
            # reference draw order: u_r[n_sub], cos(theta)[n], phi[n], (speeds), cos, phi
            hu["u_r"] = prng.uniform(size=num_particles_sub)
            hu["u_ct"] = prng.uniform(size=num_particles)
            hu["u_ph"] = prng.uniform(size=num_particles)
            hu["u_ctv"] = prng.uniform(size=num_particles)
            hu["u_phv"] = prng.uniform(size=num_particles)
            seed = PRIVATE_TRIAL_SEED
        else:
            seed = seed_from_prng(prng, _SPECIES[self.ptype])

        streamed = bool(host) and sub_sample == 1 and peer_moments is None
        n_dev = 0 if streamed else n_loc   # host=True stages through two chunk-sized slots instead
        pos = torch.empty((3, n_dev), dtype=dtype, device=dev)
        vel = torch.empty((3, n_dev), dtype=dtype, device=dev)
        mass = torch.empty(n_dev, dtype=dtype, device=dev)
        pot = torch.empty(n_dev, dtype=dtype, device=dev) if compute_potential else None
        moments = torch.zeros(N.CGB_NMOMENTS, dtype=torch.float64, device=dev) if diagnostics else None
        peer_slot = None
        if peer_moments is not None:
            if sub_sample != 1:
                raise ValueError("peer_moments needs sub_sample == 1")
            peer_slot = peer_moments.next_slot()  # taken on every rank, also by ranks with no particles
            moments = peer_moments.slot(peer_slot)
        pmass = tabs.mtot / num_particles if num_particles > 0 else 0.0  # virial.py:225-227

        def base_args(begin, end, ld):
            a = N.SampleArgs()
            a.d_workspace = tabs.workspace.data_ptr()
            a.id_begin, a.id_end = begin, end
            a.particle_mass = pmass
            a.seed = seed & 0xFFFFFFFFFFFFFFFF
            a.out_dtype = code
            a.ld = ld
            return a

        keep = []  # keep device uniforms alive until the kernels are enqueued

        def host_u(a, begin, end, sub):
            if not hu:
                return
            def put(name, b, e):
                t = N.dev_f64(hu[name][b:e], dev)
                keep.append(t)
                return t.data_ptr()
            if sub:
                a.hu.d_u_r = put("u_r", begin, end)
            a.hu.d_u_ct, a.hu.d_u_ph = put("u_ct", begin, end), put("u_ph", begin, end)
            a.hu.d_u_ctv, a.hu.d_u_phv = put("u_ctv", begin, end), put("u_phv", begin, end)

        if host and sub_sample == 1 and peer_moments is None:
            from .hoststream import HostStreamer

            layout = OrderedDict([("particle_position", 3), ("particle_velocity", 3), ("particle_mass", 1)])
            if compute_potential:
                layout["particle_potential"] = 1
            prev = getattr(host, "_host_buffers", {}).get(key) if isinstance(host, ClusterParticles) else None
            hs = HostStreamer(layout, n_loc, dtype, dev, chunk, out=prev)

            def launch(stage, b0, b1):
                a = base_args(lo + b0, lo + b1, hs.chunk)
                a.d_pos, a.d_vel = stage["particle_position"].data_ptr(), stage["particle_velocity"].data_ptr()
                a.d_mass = stage["particle_mass"].data_ptr()
                a.d_pot = stage["particle_potential"].data_ptr() if compute_potential else None
                a.d_moments = moments.data_ptr() if moments is not None else None
                host_u(a, lo + b0, lo + b1, True)
                N.check(L.cgb_sample_particles(N.ctypes.byref(a), N.ptr(tabs.scratch), tabs.scratch.numel(),
                                               N.stream_ptr()))

            hb = hs.run(launch, wait=wait) if n_loc > 0 else hs.host
            if wait:
                hs.copy_stream.synchronize()
            fields = OrderedDict()
            fields[key, "particle_position"] = hb.arrays["particle_position"].T
            fields[key, "particle_velocity"] = hb.arrays["particle_velocity"].T
            fields[key, "particle_mass"] = hb.arrays["particle_mass"]
            if compute_potential:
                fields[key, "particle_potential"] = hb.arrays["particle_potential"]
            diag = {key: SamplerDiagnostics(moments, shard)} if moments is not None else {}
            del keep
            out = ClusterParticles(key, fields, diag)
            out._host_buffers = {key: hb}
            out.d2h_bytes = hs.bytes
            if not wait:
                out._pending.append(hs)  # keeps the staging slots alive until the copies have left them
            return out

        if n_loc > 0 and sub_sample == 1:
            a = base_args(lo, hi, n_loc)
            a.d_pos, a.d_vel, a.d_mass = pos.data_ptr(), vel.data_ptr(), mass.data_ptr()
            a.d_pot = pot.data_ptr() if pot is not None else None
            a.d_moments = moments.data_ptr() if moments is not None else None
            if peer_slot is not None:
                peer_moments.attach(a, peer_slot)
            host_u(a, lo, hi, True)
            N.check(L.cgb_sample_particles(N.ctypes.byref(a), N.ptr(tabs.scratch), tabs.scratch.numel(),
                                           N.stream_ptr()))
        elif n_loc > 0:
            # unique (r, |v|) for n_sub particles, tiled over all n (virial.py:170-173,208-211)
            ns = num_particles_sub
            r_sub = torch.empty(ns, dtype=torch.float64, device=dev)
            v_sub = torch.empty(ns, dtype=torch.float64, device=dev)
            psi_sub = torch.empty(ns, dtype=torch.float64, device=dev)
            scratch = torch.empty((2, 3, ns), dtype=dtype, device=dev)
            a = base_args(0, ns, ns)
            a.d_pos, a.d_vel = scratch[0].data_ptr(), scratch[1].data_ptr()
            a.d_radius, a.d_speed = r_sub.data_ptr(), v_sub.data_ptr()
            a.d_pot = None
            a.d_moments = moments.data_ptr() if moments is not None else None
            if hu:
                t = N.dev_f64(hu["u_r"], dev)
                keep.append(t)
                a.hu.d_u_r = t.data_ptr()
            N.check(L.cgb_sample_particles(N.ctypes.byref(a), N.ptr(tabs.scratch), tabs.scratch.numel(),
                                           N.stream_ptr()))
            if compute_potential:
                N.check(L.cgb_spline_eval(N.ptr(tabs.psi_spline.t), N.ptr(tabs.psi_spline.c),
                                          tabs.psi_spline.n, N.ptr(r_sub), N.ptr(psi_sub), ns, 0,
                                          N.stream_ptr()))
            b = base_args(lo, hi, n_loc)
            b.d_pos, b.d_vel, b.d_mass = pos.data_ptr(), vel.data_ptr(), mass.data_ptr()
            b.d_pot = pot.data_ptr() if pot is not None else None
            host_u(b, lo, hi, False)
            N.check(L.cgb_assemble_particles(N.ctypes.byref(b), N.ptr(r_sub), N.ptr(v_sub),
                                             N.ptr(psi_sub) if compute_potential else None, ns,
                                             N.stream_ptr()))

        fields = OrderedDict()
        fields[key, "particle_position"] = pos.T
        fields[key, "particle_velocity"] = vel.T
        fields[key, "particle_mass"] = mass
        if compute_potential:
            fields[key, "particle_potential"] = pot
        diag = {}
        if host:
            fields = OrderedDict((k, _pinned_copy(v)) for k, v in fields.items())
            torch.cuda.current_stream().synchronize()
        if moments is not None:
            # sub_sample > 1: every rank draws the whole sub-sample, so its moments already are
            # the global ones (summing them over ranks would count the sub-sample world times)
            diag[key] = SamplerDiagnostics(moments, shard, reduced=peer_slot is not None or sub_sample > 1)
            if peer_slot is not None:
                peer_moments.track(diag[key])
        del keep
        return ClusterParticles(key, fields, diag)


def _pinned_copy(t: torch.Tensor) -> torch.Tensor:
    """Device field -> pinned host tensor with the same logical shape and plane-major layout."""
    if t.dim() == 2:
        h = torch.empty((t.shape[1], t.shape[0]), dtype=t.dtype, pin_memory=True)
        h.copy_(t.T, non_blocking=True)
        return h.T
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    return h


class SamplerDiagnostics:
    """Kernel-side moments of one generation call (lazy D2H; optional all-reduce)."""

    def __init__(self, moments: torch.Tensor, shard=None, reduced=False):
        self.moments = moments
        self.shard = shard
        #: True when the kernels already summed over ranks (PeerMoments); all_reduce is then a no-op
        self.reduced = reduced

    def all_reduce(self):
        """Sum the moments over the ranks of the default process group (NCCL on GPUs);
        the only collective of the path."""
        import torch.distributed as dist

        if self.reduced:
            return self
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            mx = self.moments[2:3].clone()
            dist.all_reduce(self.moments, op=dist.ReduceOp.SUM)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            self.moments[2:3] = mx
        return self

    def as_dict(self):
        m = self.moments.cpu().numpy()
        d = dict(zip(N.MOMENT_NAMES, (float(x) for x in m)))
        n = max(d["particles"] - d["failures"], 1.0)
        d["mean_trials"] = d["sum_trials"] / n
        d["mean_v2"] = d["sum_v2"] / n
        d["mean_psi"] = d["sum_psi"] / n
        return d
