"""
TESTS ONLY -- drives the reference's own code (loaded unmodified by ``tests/_refshim``)
for the BASELINE configs, and records the intermediate quantities the parity tests
and ``scripts/make_golden.py`` need.

Nothing of the reference is edited or copied: intermediate values are observed by
wrapping the *scipy / utils callables the reference's modules look up by name*
(``virial.InterpolatedUnivariateSpline``, ``virial.generate_particle_radii``) with
recording pass-throughs, and by the injected ``generate_velocities`` seam.
"""
from __future__ import annotations

import contextlib
import warnings

import numpy as np
from numpy.random import RandomState

from tests._refshim import available, load_reference  # noqa: F401

VIKH_DENS = dict(rho_0=4.2e5, r_c=100.0, r_s=1000.0, alpha=1.0, beta=0.67, epsilon=3.0)
VIKH_TEMP = dict(T_0=10.0, a=0.0, b=2.0, c=1.0, r_t=1500.0, T_min=6.0, r_cool=80.0, a_cool=1.0)


class ReplayState(RandomState):
    """A ``RandomState`` whose ``uniform`` hands out caller-supplied [0,1) uniforms in
    call order, transformed as numpy does (``low + (high - low) * u``).  It passes
    ``utils.parse_prng`` untouched (utils.py:184-188), which is how device-RNG
    uniforms reach the reference's unmodified ``generate_particles``."""

    def __new__(cls, *a, **k):
        return super().__new__(cls)

    def __init__(self, draws):
        super().__init__(0)
        self._draws = [np.asarray(d, dtype=np.float64) for d in draws]

    def uniform(self, low=0.0, high=1.0, size=None):
        u = self._draws.pop(0)
        n = int(np.prod(size)) if size is not None else 1
        assert u.size >= n, f"replay stream too short: {u.size} < {n}"
        u = u[:n]
        if low == 0.0 and high == 1.0:
            return u.copy()
        return low + (high - low) * u


def replay_prng(uniforms, n_sub=None, gas=False):
    """Draw order of virial.py:161-214 (u_r, u_ct, u_ph, [velocities], u_ctv, u_phv)
    or model.py:879-890 for gas / tracers (u_r, u_ct, u_ph)."""
    ur = np.asarray(uniforms["u_r"])
    if n_sub is not None:
        ur = ur[:n_sub]
    draws = [ur, uniforms["u_ct"], uniforms["u_ph"]]
    if not gas:
        draws += [uniforms["u_ctv"], uniforms["u_phv"]]
    return ReplayState(draws)


def reference_model(name, num_points=1000):
    """The BASELINE configs built by the reference's ClusterModel."""
    R = load_reference()
    rp, CM = R.radial_profiles, R.model.ClusterModel
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if name == "hernquist":
            return CM.no_gas(0.1, 1.0e4, rp.hernquist_density_profile(1.0e15, 500.0), num_points=num_points)
        if name == "nfw_bcg":
            star = rp.hernquist_density_profile(1.0e12, 20.0)
            total = rp.nfw_density_profile(1.0e7, 100.0) + star
            return CM.no_gas(0.1, 1.0e4, total, stellar_density=star, num_points=num_points)
        if name == "vikhlinin":
            return CM.from_dens_and_temp(0.1, 1.0e4, rp.vikhlinin_density_profile(**VIKH_DENS),
                                         rp.vikhlinin_temperature_profile(**VIKH_TEMP), num_points=num_points)
    raise KeyError(name)


def oracle_model(name, num_points=1000):
    from oracle import cg_oracle as O

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if name == "hernquist":
            return O.no_gas(0.1, 1.0e4, O.hernquist_density_profile(1.0e15, 500.0), num_points=num_points)
        if name == "nfw_bcg":
            star = O.hernquist_density_profile(1.0e12, 20.0)
            total = O.add_profiles(O.nfw_density_profile(1.0e7, 100.0), star)
            return O.no_gas(0.1, 1.0e4, total, stellar_density=star, num_points=num_points)
        if name == "vikhlinin":
            return O.from_dens_and_temp(0.1, 1.0e4, O.vikhlinin_density_profile(**{**VIKH_DENS}),
                                        O.vikhlinin_temperature_profile(**VIKH_TEMP), num_points=num_points)
    raise KeyError(name)


@contextlib.contextmanager
def recording(R):
    """Record, inside the reference's ``virial`` module, every spline it builds
    ((x, y) of each ``InterpolatedUnivariateSpline``) and every radius draw."""
    rec = {"splines": [], "radii": []}
    IUS = R.virial.InterpolatedUnivariateSpline
    gpr = R.virial.generate_particle_radii

    class RecordingIUS(IUS):
        def __init__(self, x, y, *a, **k):
            rec["splines"].append((np.array(x, dtype=np.float64), np.array(y, dtype=np.float64)))
            super().__init__(x, y, *a, **k)

    def rec_radii(*a, **k):
        out = gpr(*a, **k)
        rec["radii"].append(out)
        return out

    R.virial.InterpolatedUnivariateSpline = RecordingIUS
    R.virial.generate_particle_radii = rec_radii
    try:
        yield rec
    finally:
        R.virial.InterpolatedUnivariateSpline = IUS
        R.virial.generate_particle_radii = gpr


class _CaptureVelocities:
    """Wraps the injected seam for one call and keeps what went in and came out."""

    def __init__(self, R, mode, seed, first_id=0):
        self.R, self.mode, self.seed, self.first_id = R, mode, seed, first_id

    def __enter__(self):
        v = self.R.velocities
        self._saved = (v.mode, v.seed, v.first_id)
        v.mode, v.seed, v.first_id = self.mode, self.seed, self.first_id
        cy = self.R.virial
        self._orig = cy.generate_velocities
        cap = self

        def wrapped(psi, vesc, fv2esc, knots, coeffs, k, quiet):
            cap.psi, cap.vesc, cap.fv2esc = np.array(psi), np.array(vesc), np.array(fv2esc)
            cap.knots, cap.coeffs, cap.k, cap.quiet = np.array(knots), np.array(coeffs), k, quiet
            cap.speed = self._orig(psi, vesc, fv2esc, knots, coeffs, k, quiet)
            cap.trials = v.last_trials
            return cap.speed

        cy.generate_velocities = wrapped
        return self

    def __exit__(self, *a):
        v = self.R.velocities
        v.mode, v.seed, v.first_id = self._saved
        self.R.virial.generate_velocities = self._orig
        return False


def velocities(R, mode="sequential", seed=None, first_id=0):
    return _CaptureVelocities(R, mode, seed, first_id)


def reference_virial_record(model, ptype):
    """Run the reference's ``VirialEquilibrium(model, ptype)`` and return it together
    with the ``g`` array of ``_generate_df`` (the y of the second spline it builds,
    virial.py:64) -- the only intermediate the reference does not keep."""
    R = load_reference()
    with recording(R) as rec, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        v = R.virial.VirialEquilibrium(model, ptype)
    assert len(rec["splines"]) == 3, "virial.py:45,64,66 build exactly three splines"
    return v, rec["splines"][1][1]


def check_virial(v):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return v.check_virial()


def plain(x):
    return np.array(getattr(x, "d", x))
