"""
Host-side utilities of the particle path: constants, the ``quad``-based table
integrators, prng parsing and the tiny unit-tagged ndarray used for model fields.

Mirrors ``cluster_generator/utils.py`` (reference) for the names the hot path uses:
``G, mp, mu, mue`` (utils.py:116-123), ``integrate_mass`` (:129-134), ``integrate``
(:137-142), ``parse_prng`` (:184-188), ``mylog`` (:66-78).
"""
from __future__ import annotations

import logging

import numpy as np
from numpy.random import RandomState
from scipy.integrate import quad

mylog = logging.getLogger("cluster_generator_b200")
if not mylog.handlers:
    _h = logging.StreamHandler()
    _h.setFormatter(logging.Formatter("%(name)-3s : [%(levelname)-9s] %(asctime)s %(message)s"))
    mylog.addHandler(_h)
    mylog.setLevel(logging.WARNING)
    mylog.propagate = False

# unyt is not a dependency here; these are the values unyt's CODATA-2014 tables give
# for the conversions at utils.py:116-119 (recorded in DESIGN.md).
_G_MKS = 6.67408e-11
_MSUN_KG = 1.98841586e30
_SEC_PER_MYR = 31557600.0 * 1.0e6
_M_PER_KPC = 3.0856775809623245e19
_MP_KG = 1.672621898e-27
_J_PER_KEV = 1.6021766208e-16

# unyt's conversion rule: value * old.base_value / new.base_value (tests/_refshim pins it)
G = _G_MKS * (1.0 / (_MSUN_KG**-1.0 * _SEC_PER_MYR**-2.0 * _M_PER_KPC**3.0))  #: kpc**3/Msun/Myr**2
mp = _MP_KG / _MSUN_KG  #: Msun
keV = _J_PER_KEV / _MSUN_KG * _SEC_PER_MYR**2 / _M_PER_KPC**2  #: Msun*kpc**2/Myr**2
kpc_to_cm = _M_PER_KPC * 100.0
_MSUN_G = _MSUN_KG * 1.0e3
_MH_G = 1.007947 * 1.660539040e-27 * 1.0e3  # unyt's mass_hydrogen

X_H = 0.76  # bin/config.yaml: physics.hydrogen_abundance
mu = 1.0 / (2.0 * X_H + 0.75 * (1.0 - X_H))
mue = 1.0 / (X_H + 0.5 * (1.0 - X_H))


class unit_array(np.ndarray):
    """ndarray with a ``units`` string and unyt's ``.d`` / ``.v`` accessors."""

    def __new__(cls, value, units=""):
        obj = np.asarray(value, dtype=np.float64).view(cls)
        obj.units = units
        return obj

    def __array_finalize__(self, obj):
        self.units = getattr(obj, "units", "")

    @property
    def d(self):
        return np.asarray(self)

    v = d
    value = d


def integrate_mass(profile, rr):
    """M(<r) = 4 pi int_0^r rho r^2 dr, one ``quad`` per radius (utils.py:129-134)."""
    mass_int = lambda r: profile(r) * r * r
    mass = np.zeros(rr.shape)
    for i, r in enumerate(rr):
        mass[i] = 4.0 * np.pi * quad(mass_int, 0, r)[0]
    return mass


def integrate(profile, rr):
    """int_r^rmax profile dr, one ``quad`` per radius (utils.py:137-142)."""
    ret = np.zeros(rr.shape)
    rmax = rr[-1]
    for i, r in enumerate(rr):
        ret[i] = quad(profile, r, rmax)[0]
    return ret


def integrate_toinf(profile, rr):
    """utils.py:145-151"""
    ret = np.zeros(rr.shape)
    rmax = rr[-1]
    for i, r in enumerate(rr):
        ret[i] = quad(profile, r, rmax)[0]
    ret[:] += quad(profile, rmax, np.inf, limit=100)[0]
    return ret


def parse_prng(prng):
    """utils.py:184-188"""
    if isinstance(prng, RandomState):
        return prng
    return RandomState(prng)


def seed_from_prng(prng: RandomState, species: int) -> int:
    """64-bit Philox key for one generation call.

    Drawn from the caller's MT19937 stream, so a fixed ``prng`` reproduces the whole
    sequence of calls (like the reference's shared ``RandomState``) while successive
    calls and different species get independent Philox streams.
    """
    lo, hi = (int(x) for x in prng.randint(0, 2**32, size=2, dtype=np.uint64))
    return (lo | (hi << 32)) ^ ((species * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)


def ensure_list(x):
    if isinstance(x, (list, tuple, set)):
        return list(x)
    return [x]


def integrate_spline_device(spline, a, b, power=1, epsabs=1.49e-8, epsrel=1.49e-8):
    """``[quad(lambda x: spline(x) * x**power, a_i, b_i)[0] for i]`` on the GPU.

    ``spline`` is a scipy ``InterpolatedUnivariateSpline`` (its FITPACK knots and
    coefficients are uploaded as they are, so the integrand is the reference's to the
    bit); the quadrature is the same QUADPACK ``dqagse`` restatement that builds f(E).
    Used for the potential table (model.py:219-224) when ``integrator="device"``.
    """
    import torch

    from . import _native as N

    t, c, k = spline._eval_args
    if k != 3:
        raise ValueError("cubic splines only")
    a = np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), np.shape(b) or np.shape(a)))
    b = np.ascontiguousarray(np.broadcast_to(np.asarray(b, dtype=np.float64), a.shape))
    n = len(t) - 4
    td, cd = N.dev_f64(t), N.dev_f64(np.asarray(c)[:n])
    ad, bd = N.dev_f64(a), N.dev_f64(b)
    out = torch.empty(a.size, dtype=torch.float64, device=td.device)
    N.check(N.lib().cgb_integrate_spline(N.ptr(td), N.ptr(cd), n, N.ptr(ad), N.ptr(bd), a.size, int(power),
                                         epsabs, epsrel, N.ptr(out), None, None, N.stream_ptr()))
    return out.cpu().numpy().reshape(a.shape)


def quad_callable_device(fn, a, b, epsabs=1.49e-8, epsrel=1.49e-8, return_info=False):
    """``[quad(fn, a_i, b_i)[0] for i]`` for a vectorised HOST callable ``fn`` with QUADPACK's
    adaptive decisions on the GPU (``cgb_qags_replay``).

    ``scipy.integrate.quad`` calls ``fn`` once per abscissa -- ~86 000 Python callbacks for the
    1000 integrals of ``integrate_mass``.  Here ``dqagse`` is replayed on the device against a
    cache of integrand values; after every replay the abscissae all unfinished integrals need
    next (the two halves of their worst interval) come back in ONE array, ``fn`` is evaluated on
    it in one vectorised call and the values are appended to the cache: one host call per
    bisection level.  The decisions are QUADPACK's (the same restatement that builds f(E)); the
    integrand values are ``fn``'s own, evaluated on arrays instead of Python scalars -- identical
    wherever NumPy's array arithmetic equals its scalar arithmetic (everything but ``**`` with
    non-integer exponents on some CPUs), so tables agree with ``quad`` to the last bit or two.
    """
    import torch

    from . import _native as N

    a = np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), np.shape(b) or np.shape(a)))
    b = np.ascontiguousarray(np.broadcast_to(np.asarray(b, dtype=np.float64), a.shape))
    shape, m = a.shape, a.size
    if m == 0:
        return np.zeros(shape)
    L = N.lib()
    dev = N.require_cuda()
    ad, bd = N.dev_f64(a.ravel(), dev), N.dev_f64(b.ravel(), dev)
    ncalls = int(L.cgb_qags_replay_cache_doubles(m)) // (21 * m)
    cache = torch.empty((m, ncalls, 21), dtype=torch.float64, device=dev)
    ncached = torch.zeros(m, dtype=torch.int32, device=dev)
    status = torch.zeros(m, dtype=torch.int32, device=dev)
    nreq = torch.zeros(m, dtype=torch.int32, device=dev)
    xreq = torch.empty((m, 2, 21), dtype=torch.float64, device=dev)
    out = torch.empty(m, dtype=torch.float64, device=dev)
    last = torch.zeros(m, dtype=torch.int32, device=dev)
    ier = torch.zeros(m, dtype=torch.int32, device=dev)
    have = np.zeros(m, dtype=np.int64)
    levels = 0
    for levels in range(1, 2 * 50 + 2):
        N.check(L.cgb_qags_replay(N.ptr(ad), N.ptr(bd), m, epsabs, epsrel, N.ptr(cache), N.ptr(ncached),
                                  N.ptr(xreq), N.ptr(nreq), N.ptr(status), N.ptr(out), N.ptr(last), N.ptr(ier),
                                  N.stream_ptr()))
        nr = nreq.cpu().numpy().astype(np.int64)
        idx = np.nonzero(nr)[0]
        if idx.size == 0:
            break
        x = xreq.cpu().numpy()
        # requested call sets, integral-major: rows (i, j) for j < nr[i]
        rows = np.repeat(idx, nr[idx])
        cols = np.concatenate([np.arange(k) for k in nr[idx]])
        vals = np.asarray(fn(x[rows, cols].ravel()), dtype=np.float64).reshape(-1, 21)
        rows_d = torch.as_tensor(rows, device=dev)
        kpos_d = torch.as_tensor(have[rows] + cols, device=dev)
        cache[rows_d, kpos_d] = torch.as_tensor(vals, device=dev)
        have[idx] += nr[idx]
        ncached.copy_(torch.as_tensor(have, dtype=torch.int32))
    else:
        raise RuntimeError("cgb_qags_replay did not converge")
    res = out.cpu().numpy().reshape(shape)
    if return_info:
        return res, dict(last=last.cpu().numpy().reshape(shape), ier=ier.cpu().numpy().reshape(shape), levels=levels)
    return res


def integrate_mass_device(profile, rr):
    """``integrate_mass`` (utils.py:129-134) through :func:`quad_callable_device`."""
    rr = np.asarray(rr, dtype=np.float64)
    return 4.0 * np.pi * quad_callable_device(lambda r: profile(r) * r * r, np.zeros_like(rr), rr)


def integrate_device(profile, rr):
    """``integrate`` (utils.py:137-142) through :func:`quad_callable_device`."""
    rr = np.asarray(rr, dtype=np.float64)
    return quad_callable_device(profile, rr, np.full_like(rr, rr[-1]))
