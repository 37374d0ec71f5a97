"""
TEST INFRASTRUCTURE ONLY -- CPU restatement of cluster_generator's particle path.

Bare-ndarray restatement of the reference's Eddington f(E) build, virial check,
inverse-CDF radius draw and particle assembly, calling the *same* scipy / numpy
entry points in the *same* order as the reference so that the numbers are the
reference's numbers by construction -- and, since round 2, by test: the reference's
own modules run unmodified behind ``tests/_refshim`` and ``tests/test_reference_cpu.py``
asserts bitwise equality with this file (see ``oracle/__init__.py``).

Citations are ``file:line`` into the reference tree (``cluster_generator/``).

Parity status: rows D/C/R/E/T/P/Gs pinned bitwise to the reference's executed code;
row V (the absent Cython rejection loop) **parity unpinned**, restated from its
call-site contract -- see ``oracle/__init__.py``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from collections import OrderedDict

import numpy as np
from numpy.random import RandomState
from scipy.integrate import cumulative_trapezoid as cumtrapz
from scipy.integrate import quad
from scipy.interpolate import InterpolatedUnivariateSpline

# --------------------------------------------------------------------------- #
# Constants (utils.py:116-123).  unyt is not installed, so the values unyt would
# produce are hard-coded from its CODATA-2014 tables; both the oracle and the
# product share them, so kernel-vs-oracle parity does not depend on the vintage.
# --------------------------------------------------------------------------- #
_G_MKS = 6.67408e-11
_MSUN_KG = 1.98841586e30
_SEC_PER_MYR = 31557600.0 * 1.0e6
_M_PER_KPC = 3.0856775809623245e19
_MP_KG = 1.672621898e-27
_J_PER_KEV = 1.6021766208e-16

# unyt's conversion rule: value * old.base_value / new.base_value (tests/_refshim pins it)
G = _G_MKS * (1.0 / (_MSUN_KG**-1.0 * _SEC_PER_MYR**-2.0 * _M_PER_KPC**3.0))  # kpc**3/Msun/Myr**2
mp = _MP_KG / _MSUN_KG  # Msun
KEV = _J_PER_KEV / _MSUN_KG * _SEC_PER_MYR**2 / _M_PER_KPC**2  # Msun*kpc**2/Myr**2
X_H = 0.76  # bin/config.yaml physics.hydrogen_abundance
mu = 1.0 / (2.0 * X_H + 0.75 * (1.0 - X_H))  # utils.py:122
mue = 1.0 / (X_H + 0.5 * (1.0 - X_H))  # utils.py:123


# --------------------------------------------------------------------------- #
# Radial profiles (radial_profiles.py)
# --------------------------------------------------------------------------- #
def hernquist_density_profile(M_0, a):
    """radial_profiles.py:173-186"""
    return lambda r: M_0 / (2.0 * np.pi * a**3) / ((r / a) * (1.0 + r / a) ** 3)


def nfw_density_profile(rho_s, r_s):
    """radial_profiles.py:249-262"""
    return lambda r: rho_s / ((r / r_s) * (1.0 + r / r_s) ** 2)


def vikhlinin_density_profile(rho_0, r_c, r_s, alpha, beta, epsilon, gamma=None):
    """radial_profiles.py:605-637"""
    if gamma is None:
        gamma = 3.0
    return (
        lambda r: rho_0
        * (r / r_c) ** (-0.5 * alpha)
        * (1.0 + (r / r_c) ** 2) ** (-1.5 * beta + 0.25 * alpha)
        * (1.0 + (r / r_s) ** gamma) ** (-0.5 * epsilon / gamma)
    )


def vikhlinin_temperature_profile(T_0, a, b, c, r_t, T_min, r_cool, a_cool):
    """radial_profiles.py:640-671"""

    def _temp(r):
        x = (r / r_cool) ** a_cool
        t = (r / r_t) ** (-a) / ((1.0 + (r / r_t) ** b) ** (c / b))
        return T_0 * t * (x + T_min / T_0) / (x + 1)

    return _temp


def add_profiles(p, q):
    """radial_profiles.py:16-25 (RadialProfile.__add__)"""
    return lambda r: np.add(p(r), q(r))


# --------------------------------------------------------------------------- #
# Integrators and the radius draw (utils.py)
# --------------------------------------------------------------------------- #
def integrate_mass(profile, rr):
    """utils.py:129-134"""
    mass_int = lambda r: profile(r) * r * r
    mass = np.zeros(rr.shape)
    for i, r in enumerate(rr):
        mass[i] = 4.0 * np.pi * quad(mass_int, 0, r)[0]
    return mass


def integrate(profile, rr):
    """utils.py:137-142"""
    ret = np.zeros(rr.shape)
    rmax = rr[-1]
    for i, r in enumerate(rr):
        ret[i] = quad(profile, r, rmax)[0]
    return ret


def parse_prng(prng):
    """utils.py:184-188"""
    if isinstance(prng, RandomState):
        return prng
    return RandomState(prng)


def generate_particle_radii(r, m, num_particles, r_max=None, prng=None, u=None):
    """utils.py:154-166.  ``u`` (optional) substitutes the prng draw so a caller
    can feed the same uniforms to the CUDA path."""
    if r_max is None:
        ridx = r.size
    else:
        ridx = np.searchsorted(r, r_max)
    mtot = m[ridx - 1]
    if u is None:
        u = parse_prng(prng).uniform(size=num_particles)
    P_r = np.insert(m[:ridx], 0, 0.0)
    P_r /= P_r[-1]
    r = np.insert(r[:ridx], 0, 0.0)
    radius = np.interp(u, P_r, r, left=0.0, right=1.0)
    return radius, mtot


def radius_cdf_table(r, m, r_max=None):
    """The (P_r, r) table generate_particle_radii interpolates in (utils.py:155-164)."""
    ridx = r.size if r_max is None else int(np.searchsorted(r, r_max))
    mtot = m[ridx - 1]
    P_r = np.insert(m[:ridx], 0, 0.0)
    P_r /= P_r[-1]
    rr = np.insert(r[:ridx], 0, 0.0)
    return P_r, rr, mtot


# --------------------------------------------------------------------------- #
# Model tables (model.py)
# --------------------------------------------------------------------------- #
class Model:
    """Field dictionary on bare ndarrays (model.py:51-110 without unyt)."""

    def __init__(self, num_elements, fields):
        self.num_elements = num_elements
        self.fields = fields
        self._dm_virial = None
        self._star_virial = None

    def __getitem__(self, key):
        return self.fields[key]

    def __contains__(self, key):
        return key in self.fields

    @property
    def dm_virial(self):
        """model.py:111-118"""
        if self._dm_virial is None:
            self._dm_virial = VirialEquilibrium(self, "dark_matter")
        return self._dm_virial

    @property
    def star_virial(self):
        """model.py:120-127"""
        if self._star_virial is None and "stellar_density" in self:
            self._star_virial = VirialEquilibrium(self, "stellar")
        return self._star_virial


def _from_scratch(fields, stellar_density=None):
    """model.py:213-269"""
    rr = fields["radius"]
    tdens_func = InterpolatedUnivariateSpline(rr, fields["total_density"])
    gpot_profile = lambda r: tdens_func(r) * r
    gpot1 = fields["total_mass"] / fields["radius"]
    gpot2 = 4.0 * np.pi * integrate(gpot_profile, rr)
    fields["gravitational_potential"] = -G * (gpot1 + gpot2)

    if "density" in fields and "gas_mass" not in fields:
        m0 = fields["density"][0] * rr[0] ** 3 / 3.0
        fields["gas_mass"] = (
            4.0 * np.pi * cumtrapz(fields["density"] * rr * rr, x=rr, initial=0.0) + m0
        )

    if stellar_density is not None:
        fields["stellar_density"] = stellar_density(rr)
        fields["stellar_mass"] = integrate_mass(stellar_density, rr)

    mdm = fields["total_mass"].copy()
    ddm = fields["total_density"].copy()
    if "density" in fields:
        mdm -= fields["gas_mass"]
        ddm -= fields["density"]
    if "stellar_mass" in fields:
        mdm -= fields["stellar_mass"]
        ddm -= fields["stellar_density"]
    mdm[ddm < 0.0] = mdm.max()
    ddm[ddm < 0.0] = 0.0
    fields["dark_matter_density"] = ddm
    fields["dark_matter_mass"] = mdm

    if "density" in fields:
        fields["gas_fraction"] = fields["gas_mass"] / fields["total_mass"]
    return Model(rr.size, fields)


def no_gas(rmin, rmax, total_density, stellar_density=None, num_points=1000):
    """model.py:620-656"""
    rr = np.logspace(np.log10(rmin), np.log10(rmax), num_points, endpoint=True)
    fields = OrderedDict()
    fields["radius"] = rr
    fields["total_density"] = total_density(rr)
    fields["total_mass"] = integrate_mass(total_density, rr)
    fields["gravitational_field"] = -G * fields["total_mass"] / (fields["radius"] ** 2)
    return _from_scratch(fields, stellar_density=stellar_density)


def from_dens_and_temp(rmin, rmax, density, temperature, stellar_density=None, num_points=1000):
    """model.py:474-523"""
    rr = np.logspace(np.log10(rmin), np.log10(rmax), num_points, endpoint=True)
    fields = OrderedDict()
    fields["radius"] = rr
    fields["density"] = density(rr)
    fields["temperature"] = temperature(rr)  # keV
    fields["pressure"] = fields["density"] * fields["temperature"]
    fields["pressure"] = fields["pressure"] / (mu * mp)
    fields["pressure"] = fields["pressure"] * KEV  # -> Msun/(Myr**2*kpc)
    pressure_spline = InterpolatedUnivariateSpline(rr, fields["pressure"])
    dPdr = pressure_spline(rr, 1)
    fields["gravitational_field"] = dPdr / fields["density"]
    fields["gas_mass"] = integrate_mass(density, rr)
    fields["total_mass"] = -fields["radius"] ** 2 * fields["gravitational_field"] / G
    total_mass_spline = InterpolatedUnivariateSpline(rr, fields["total_mass"])
    dMdr = total_mass_spline(rr, nu=1)
    fields["total_density"] = dMdr / (4.0 * np.pi * fields["radius"] ** 2)
    return _from_scratch(fields, stellar_density=stellar_density)


# --------------------------------------------------------------------------- #
# Virial equilibrium (virial.py)
# --------------------------------------------------------------------------- #
class VirialEquilibrium:
    """virial.py:16-236"""

    def __init__(self, model, ptype="dark_matter", df=None):
        self.num_elements = model.num_elements
        self.ptype = ptype
        self.model = model
        self.qags_info = None
        if df is None:
            self._generate_df()
        else:
            self.df = df
            self.f = InterpolatedUnivariateSpline(self.ee, df[::-1])

    def _generate_df(self, record=False):
        """virial.py:43-68"""
        pden = self.model[f"{self.ptype}_density"][::-1]
        density_spline = InterpolatedUnivariateSpline(self.ee, pden)
        g = np.zeros(self.num_elements)
        dgdp = lambda t, e: 2 * density_spline(e - t * t, 1)
        info = []
        ee = self.ee
        for i in range(self.num_elements):
            if record:
                res = quad(
                    dgdp, 0.0, np.sqrt(ee[i]), epsabs=1.49e-05, epsrel=1.49e-05,
                    args=(ee[i]), full_output=1,
                )
                g[i] = res[0]
                info.append((res[1], res[2]))
            else:
                g[i] = quad(
                    dgdp, 0.0, np.sqrt(ee[i]), epsabs=1.49e-05, epsrel=1.49e-05, args=(ee[i])
                )[0]
        g_spline = InterpolatedUnivariateSpline(ee, g)
        ff = g_spline(ee, 1) / (np.sqrt(8.0) * np.pi**2)
        self.g = g
        self.f = InterpolatedUnivariateSpline(ee, ff)
        self.df = ff[::-1]
        if record:
            self.qags_info = info

    @property
    def ee(self):
        """virial.py:70-72"""
        return -self.model["gravitational_potential"][::-1]

    @property
    def ff(self):
        """virial.py:74-76"""
        return self.df[::-1]

    def check_virial(self):
        """virial.py:78-106"""
        import warnings

        n = self.num_elements
        rho = np.zeros(n)
        pden = self.model[f"{self.ptype}_density"]
        rho_int = lambda e, psi: self.f(e) * np.sqrt(2 * (psi - e))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for i, e in enumerate(self.ee):
                rho[i] = 4.0 * np.pi * quad(rho_int, 0.0, e, args=(e,))[0]
        chk = (rho[::-1] - pden) / pden
        return rho[::-1], chk

    def sampler_tables(self, r_max=None):
        """Everything generate_particles derives before drawing (virial.py:148-166)."""
        density = f"{self.ptype}_density"
        mass = f"{self.ptype}_mass"
        nonzero = self.model[density] > 0.0
        P_r, rr, mtot = radius_cdf_table(
            self.model["radius"][nonzero], self.model[mass][nonzero], r_max
        )
        return dict(
            P_r=P_r, r_cdf=rr, mtot=float(mtot),
            radius=self.model["radius"], psi_r=self.ee[::-1],
            f_t=np.asarray(self.f._eval_args[0]), f_c=np.asarray(self.f._eval_args[1]),
        )

    def generate_particles(
        self, num_particles, r_max=None, sub_sample=1, compute_potential=False,
        prng=None, velocity_fn=None, uniforms=None,
    ):
        """virial.py:108-236.

        ``velocity_fn(psi, vesc, fv2esc, knots, coeffs, k)`` stands in for the
        absent ``generate_velocities``; default = mode-A sequential C loop.
        ``uniforms`` (dict with u_r, u_ct, u_ph, u_ctv, u_phv in [0,1)) replaces
        the five prng draws, preserving their order and meaning: cos(theta) =
        -1 + 2*u (what ``uniform(low=-1, high=1)`` computes) and phi = 2*pi*u.
        """
        num_particles_sub = num_particles // sub_sample
        key = {"dark_matter": "dm", "stellar": "star"}[self.ptype]
        density = f"{self.ptype}_density"
        mass = f"{self.ptype}_mass"
        energy_spline = InterpolatedUnivariateSpline(self.model["radius"], self.ee[::-1])

        prng = parse_prng(prng)

        def draw(name, lo_hi=None):
            if uniforms is not None:
                u = np.asarray(uniforms[name], dtype=np.float64)
                return u if lo_hi is None else lo_hi[0] + (lo_hi[1] - lo_hi[0]) * u
            if lo_hi is None:
                return prng.uniform(size=num_particles)
            return prng.uniform(low=lo_hi[0], high=lo_hi[1], size=num_particles)

        nonzero = self.model[density] > 0.0
        radius_sub, mtot = generate_particle_radii(
            self.model["radius"][nonzero], self.model[mass][nonzero], num_particles_sub,
            r_max=r_max, prng=prng,
            u=None if uniforms is None else np.asarray(uniforms["u_r"])[:num_particles_sub],
        )
        radius = np.tile(radius_sub, sub_sample)[:num_particles] if sub_sample > 1 else radius_sub

        theta = np.arccos(draw("u_ct", (-1.0, 1.0)))
        phi = 2.0 * np.pi * draw("u_ph")

        fields = OrderedDict()
        fields[key, "particle_position"] = np.array(
            [
                radius * np.sin(theta) * np.cos(phi),
                radius * np.sin(theta) * np.sin(phi),
                radius * np.cos(theta),
            ]
        ).T

        psi = energy_spline(radius_sub)
        vesc = 2.0 * psi
        fv2esc = vesc * self.f(psi)
        vesc = np.sqrt(vesc)

        if velocity_fn is None:
            velocity_fn = generate_velocities_sequential
        out = velocity_fn(psi, vesc, fv2esc, self.f.get_knots(), self.f.get_coeffs(),
                          self.f._eval_args[2])
        velocity_sub, self.last_trials = out if isinstance(out, tuple) else (out, None)
        velocity = np.tile(velocity_sub, sub_sample)[:num_particles] if sub_sample > 1 else velocity_sub

        theta = np.arccos(draw("u_ctv", (-1.0, 1.0)))
        phi = 2.0 * np.pi * draw("u_phv")

        fields[key, "particle_velocity"] = np.array(
            [
                velocity * np.sin(theta) * np.cos(phi),
                velocity * np.sin(theta) * np.sin(phi),
                velocity * np.cos(theta),
            ]
        ).T
        fields[key, "particle_mass"] = np.array([mtot / num_particles] * num_particles)
        if compute_potential:
            fields[key, "particle_potential"] = -np.tile(psi, sub_sample) if sub_sample > 1 else -psi
        self.last = dict(radius=radius, psi=psi, vesc=vesc, fv2esc=fv2esc, speed=velocity)
        return fields


def generate_gas_particles(model, num_particles, r_max=None, sub_sample=1,
                           compute_potential=False, prng=None, uniforms=None):
    """model.py:838-947"""
    prng = parse_prng(prng)
    num_particles_sub = num_particles // sub_sample
    radius_sub, mtot = generate_particle_radii(
        model["radius"], model["gas_mass"], num_particles_sub, r_max=r_max, prng=prng,
        u=None if uniforms is None else np.asarray(uniforms["u_r"])[:num_particles_sub],
    )
    radius = np.tile(radius_sub, sub_sample)[:num_particles] if sub_sample > 1 else radius_sub
    if uniforms is None:
        theta = np.arccos(prng.uniform(low=-1.0, high=1.0, size=num_particles))
        phi = 2.0 * np.pi * prng.uniform(size=num_particles)
    else:
        theta = np.arccos(-1.0 + 2.0 * np.asarray(uniforms["u_ct"]))
        phi = 2.0 * np.pi * np.asarray(uniforms["u_ph"])
    fields = OrderedDict()
    fields["gas", "particle_position"] = np.array(
        [
            radius * np.sin(theta) * np.cos(phi),
            radius * np.sin(theta) * np.sin(phi),
            radius * np.cos(theta),
        ]
    ).T
    e_arr = 1.5 * model["pressure"] / model["density"]
    get_energy = InterpolatedUnivariateSpline(model["radius"], e_arr)
    if sub_sample > 1:
        energy = np.tile(get_energy(radius_sub), sub_sample)[:num_particles]
    else:
        energy = get_energy(radius)
    fields["gas", "thermal_energy"] = energy
    fields["gas", "particle_mass"] = np.array([mtot / num_particles] * num_particles)
    get_density = InterpolatedUnivariateSpline(model["radius"], model["density"])
    if sub_sample > 1:
        density = np.tile(get_density(radius_sub), sub_sample)[:num_particles]
    else:
        density = get_density(radius)
    fields["gas", "density"] = density
    fields["gas", "particle_velocity"] = np.zeros((num_particles, 3))
    if compute_potential:
        energy_spline = InterpolatedUnivariateSpline(model["radius"], -model["gravitational_potential"])
        phi = -energy_spline(radius_sub)
        if sub_sample > 1:
            phi = np.tile(phi, sub_sample)
        fields["gas", "particle_potential"] = phi
    return fields


# --------------------------------------------------------------------------- #
# Multi-halo combination (particles.py:587-715, ics.py:22-57) on dict-of-ndarray particles
# --------------------------------------------------------------------------- #
def compute_centers_for_binary(center, d, b, a=0.0):
    """ics.py:22-57"""
    d = np.sqrt(d * d - b * b - a * a)
    diff = np.array([d, b, a])
    return center - 0.5 * diff, center + 0.5 * diff


def sample_clusters(fields, hses, center, velocity, radii=None, resample=False, passive_scalars=None):
    """particles.py:587-639; ``fields`` is updated in place."""
    num_halos = len(hses)
    pos = fields["gas", "particle_position"]
    n = pos.shape[0]
    r = np.zeros((num_halos, n))
    for i, c in enumerate(center):
        r[i, :] = ((pos - np.asarray(c)) ** 2).sum(axis=1)
    np.sqrt(r, r)
    if radii is None:
        idxs = slice(None, None, None)
    else:
        radii = np.array(radii)
        idxs = np.any(r <= radii[:, np.newaxis], axis=0)
    d = np.zeros((num_halos, n))
    e = np.zeros((num_halos, n))
    m = np.zeros((num_halos, 3, n))
    num_scalars = 0
    if passive_scalars is not None:
        num_scalars = len(passive_scalars)
        s = np.zeros((num_halos, num_scalars, n))
    for i, hse in enumerate(hses):
        if "density" not in hse:
            continue
        get_density = InterpolatedUnivariateSpline(hse["radius"], hse["density"])
        d[i, :] = get_density(r[i, :])
        e_arr = 1.5 * hse["pressure"] / hse["density"]
        get_energy = InterpolatedUnivariateSpline(hse["radius"], e_arr)
        e[i, :] = get_energy(r[i, :]) * d[i, :]
        m[i, :, :] = np.asarray(velocity[i])[:, np.newaxis] * d[i, :]
        for j in range(num_scalars):
            get_scalar = InterpolatedUnivariateSpline(hse["radius"], hse[passive_scalars[j]])
            s[i, j, :] = get_scalar(r[i, :]) * d[i, :]
    dens = d.sum(axis=0)
    eint = e.sum(axis=0) / dens
    mom = m.sum(axis=0) / dens
    if num_scalars > 0:
        ps = s.sum(axis=0) / dens
    if resample:
        vol = fields["gas", "particle_mass"] / fields["gas", "density"]
        fields["gas", "particle_mass"][idxs] = dens[idxs] * vol[idxs]
    fields["gas", "density"][idxs] = dens[idxs]
    fields["gas", "thermal_energy"][idxs] = eint[idxs]
    fields["gas", "particle_velocity"][idxs] = mom.T[idxs]
    for j in range(num_scalars):
        fields["gas", passive_scalars[j]][idxs] = ps[j, idxs]
    return fields


def combine_clusters(parts, hses, centers, velocities):
    """particles.py:642-715: ``parts`` = list of field dicts, one per halo."""
    out = OrderedDict()
    for f, c, v in zip(parts, centers, velocities):
        for (pt, name), arr in f.items():
            arr = np.array(arr, dtype=np.float64)
            if name == "particle_position":
                arr = arr + np.asarray(c)
            elif name == "particle_velocity" and pt != "gas":
                arr = arr + np.asarray(v)
            out[pt, name] = arr if (pt, name) not in out else np.concatenate([out[pt, name], arr])
    if any(k[0] == "gas" for k in out):
        sample_clusters(out, hses, centers, velocities)
    return out


# --------------------------------------------------------------------------- #
# Row V: the rejection loop (C restatement in oracle/cg_oracle_c.c)
# --------------------------------------------------------------------------- #
_HERE = os.path.dirname(os.path.abspath(__file__))
_CLIB = None


def build_c_oracle(force=False):
    """gcc-compile oracle/cg_oracle_c.c into oracle/_build/libcgoracle.so."""
    src = os.path.join(_HERE, "cg_oracle_c.c")
    out_dir = os.path.join(_HERE, "_build")
    out = os.path.join(out_dir, "libcgoracle.so")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", out, src, "-lm"]
        )
    return out


def clib():
    global _CLIB
    if _CLIB is None:
        lib = ctypes.CDLL(build_c_oracle())
        dp = ctypes.POINTER(ctypes.c_double)
        lp = ctypes.POINTER(ctypes.c_int64)
        lib.cgo_generate_velocities_seq.argtypes = [
            dp, dp, dp, dp, ctypes.c_int64, dp, ctypes.c_int64, ctypes.c_int,
            ctypes.c_uint64, dp, lp, ctypes.c_int64]
        lib.cgo_generate_velocities_seq.restype = ctypes.c_int
        lib.cgo_generate_velocities_counter.argtypes = [
            dp, dp, dp, dp, ctypes.c_int64, dp, ctypes.c_int64, ctypes.c_int,
            ctypes.c_uint64, ctypes.c_int64, dp, lp, ctypes.c_int64]
        lib.cgo_generate_velocities_counter.restype = ctypes.c_int
        lib.cgo_generate_velocities_array.argtypes = [
            dp, dp, dp, dp, ctypes.c_int64, dp, ctypes.c_int64, ctypes.c_int,
            dp, ctypes.c_int64, dp, lp, ctypes.c_int64]
        lib.cgo_generate_velocities_array.restype = ctypes.c_int
        lib.cgo_philox4x32_10.argtypes = [ctypes.POINTER(ctypes.c_uint32),
                                          ctypes.POINTER(ctypes.c_uint32),
                                          ctypes.POINTER(ctypes.c_uint32)]
        lib.cgo_philox_uniforms.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int64,
                                            ctypes.c_int64, dp]
        lib.cgo_splev.argtypes = [dp, ctypes.c_int64, dp, dp, dp, ctypes.c_int64]
        lib.cgo_qags.argtypes = [dp, ctypes.c_int64, dp, ctypes.c_double, ctypes.c_double,
                                 ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                 dp, dp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                 dp, dp, dp, dp]
        lib.cgo_qags.restype = ctypes.c_int
        lib.cgo_eddington_g.argtypes = [dp, ctypes.c_int64, dp, dp, ctypes.c_double,
                                        ctypes.c_double, dp, ctypes.POINTER(ctypes.c_int),
                                        ctypes.POINTER(ctypes.c_int)]
        lib.cgo_eddington_g.restype = ctypes.c_int
        lib.cgo_check_virial_rho.argtypes = [dp, ctypes.c_int64, dp, dp, dp,
                                             ctypes.POINTER(ctypes.c_int),
                                             ctypes.POINTER(ctypes.c_int)]
        lib.cgo_check_virial_rho.restype = ctypes.c_int
        lib.cgo_spline_fit.argtypes = [dp, dp, ctypes.c_int64, dp, dp]
        lib.cgo_spline_fit.restype = ctypes.c_int
        lib.cgo_spline_fit_fitpack.argtypes = [dp, dp, ctypes.c_int64, dp, dp]
        lib.cgo_spline_fit_fitpack.restype = ctypes.c_int
        _CLIB = lib
    return _CLIB


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _lptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


def _vel_args(psi, vesc, fv2esc, knots, coeffs):
    psi = np.ascontiguousarray(psi, dtype=np.float64)
    vesc = np.ascontiguousarray(vesc, dtype=np.float64)
    fv2esc = np.ascontiguousarray(fv2esc, dtype=np.float64)
    knots = np.ascontiguousarray(knots, dtype=np.float64)
    coeffs = np.ascontiguousarray(coeffs, dtype=np.float64)
    return psi, vesc, fv2esc, knots, coeffs


SEQ_SEED = 0x9E3779B97F4A7C15  # fixed: the Cython RNG is private and deterministic


def generate_velocities_sequential(psi, vesc, fv2esc, knots, coeffs, k, seed=SEQ_SEED):
    """Mode A: one private sequential stream shared by all particles
    (call contract virial.py:198-206; algorithm docs/source/particles.rst:102-105)."""
    psi, vesc, fv2esc, knots, coeffs = _vel_args(psi, vesc, fv2esc, knots, coeffs)
    n = psi.size
    vel = np.empty(n)
    trials = np.empty(n, dtype=np.int64)
    rc = clib().cgo_generate_velocities_seq(
        _dptr(psi), _dptr(vesc), _dptr(fv2esc), _dptr(knots), knots.size, _dptr(coeffs),
        coeffs.size, int(k), ctypes.c_uint64(seed), _dptr(vel), _lptr(trials), n)
    if rc != 0:
        raise RuntimeError(f"cgo_generate_velocities_seq failed rc={rc}")
    return vel, trials


def generate_velocities_counter(psi, vesc, fv2esc, knots, coeffs, k, seed=0, first_id=0):
    """Mode B: (u1,u2) = Philox4x32-10(counter=(particle id, trial), key=(seed, VEL stream))."""
    psi, vesc, fv2esc, knots, coeffs = _vel_args(psi, vesc, fv2esc, knots, coeffs)
    n = psi.size
    vel = np.empty(n)
    trials = np.empty(n, dtype=np.int64)
    rc = clib().cgo_generate_velocities_counter(
        _dptr(psi), _dptr(vesc), _dptr(fv2esc), _dptr(knots), knots.size, _dptr(coeffs),
        coeffs.size, int(k), ctypes.c_uint64(seed), int(first_id), _dptr(vel), _lptr(trials), n)
    if rc != 0:
        raise RuntimeError(f"cgo_generate_velocities_counter failed rc={rc}")
    return vel, trials


def generate_velocities_array(psi, vesc, fv2esc, knots, coeffs, k, uniforms):
    """Mode B with literal host-supplied uniforms ``uniforms[n, T_cap, 2]``.
    trials[i] == -1 (and vel NaN) where no trial within T_cap accepted."""
    psi, vesc, fv2esc, knots, coeffs = _vel_args(psi, vesc, fv2esc, knots, coeffs)
    uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
    n = psi.size
    assert uniforms.shape[0] == n and uniforms.shape[2] == 2
    vel = np.empty(n)
    trials = np.empty(n, dtype=np.int64)
    rc = clib().cgo_generate_velocities_array(
        _dptr(psi), _dptr(vesc), _dptr(fv2esc), _dptr(knots), knots.size, _dptr(coeffs),
        coeffs.size, int(k), _dptr(uniforms), uniforms.shape[1], _dptr(vel), _lptr(trials), n)
    if rc != 0:
        raise RuntimeError(f"cgo_generate_velocities_array failed rc={rc}")
    return vel, trials


def philox4x32_10(ctr, key):
    c = (ctypes.c_uint32 * 4)(*ctr)
    k = (ctypes.c_uint32 * 2)(*key)
    o = (ctypes.c_uint32 * 4)()
    clib().cgo_philox4x32_10(c, k, o)
    return tuple(int(x) for x in o)


# Philox stream ids (key word 1 low bits); shared with the CUDA path by contract,
# see include/cgb200.h.
STREAM_RADIUS, STREAM_POS_DIR, STREAM_VEL_TRIAL, STREAM_VEL_DIR = 0, 1, 2, 3


def philox_uniforms(seed, stream, first_id, n):
    """The two fp64 uniforms per particle id the device draws on a non-trial stream
    (counter = (id_lo, id_hi, 0, 0)); returns array [n, 2]."""
    out = np.empty((n, 2))
    clib().cgo_philox_uniforms(ctypes.c_uint64(seed), int(stream), int(first_id), n, _dptr(out))
    return out


def device_rng_uniforms(seed, first_id, n):
    """The u_r/u_ct/u_ph/u_ctv/u_phv set the device RNG uses for ids first_id..+n."""
    ur = philox_uniforms(seed, STREAM_RADIUS, first_id, n)
    up = philox_uniforms(seed, STREAM_POS_DIR, first_id, n)
    uv = philox_uniforms(seed, STREAM_VEL_DIR, first_id, n)
    return dict(u_r=ur[:, 0].copy(), u_ct=up[:, 0].copy(), u_ph=up[:, 1].copy(),
                u_ctv=uv[:, 0].copy(), u_phv=uv[:, 1].copy())


def splev_c(t, c, x):
    """FITPACK-style cubic B-spline evaluation (ext=0) by the C restatement."""
    t = np.ascontiguousarray(t, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    clib().cgo_splev(_dptr(t), t.size, _dptr(c), _dptr(x), _dptr(y), x.size)
    return y
