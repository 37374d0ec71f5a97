"""
Analytic radial profiles consumed by the particle path.

Mirrors ``cluster_generator/radial_profiles.py``: the callable ``RadialProfile``
algebra (radial_profiles.py:6-118) and the factories the BASELINE configs use
(Hernquist :173-186, NFW :249-262, Vikhlinin density :605-637 and temperature
:640-671), plus a handful of closed forms that cost nothing.  Profiles are host
closures evaluated once on ~1000 radii; they are inputs to the path, not part of it.
"""
import numpy as np


class RadialProfile:
    def __init__(self, profile):
        self.profile = profile.profile if isinstance(profile, RadialProfile) else profile

    def __call__(self, r):
        return self.profile(r)

    def _do_op(self, other, op):
        if hasattr(other, "profile"):
            return lambda r: op(self.profile(r), other.profile(r))
        return lambda r: op(self.profile(r), other)

    def __add__(self, other):
        return RadialProfile(self._do_op(other, np.add))

    def __mul__(self, other):
        return RadialProfile(self._do_op(other, np.multiply))

    __radd__ = __add__
    __rmul__ = __mul__

    def __pow__(self, power):
        return RadialProfile(lambda r: self.profile(r) ** power)

    def add_core(self, r_core, alpha):
        def _core(r):
            x = r / r_core
            return self.profile(r) * (1.0 - np.exp(-(x**alpha)))

        return RadialProfile(_core)

    def cutoff(self, r_cut, k=5):
        def _cutoff(r):
            x = r / r_cut
            step = 1.0 / (1.0 + np.exp(-2 * k * (x - 1)))
            return self.profile(r) * (1.0 - step)

        return RadialProfile(_cutoff)

    @classmethod
    def from_array(cls, r, f_r):
        from scipy.interpolate import UnivariateSpline

        return cls(UnivariateSpline(r, f_r))


def constant_profile(const):
    return RadialProfile(lambda r: const * np.ones_like(r))


def power_law_profile(A, r_s, alpha):
    return RadialProfile(lambda r: A * (r / r_s) ** alpha)


def beta_model_profile(rho_c, r_c, beta):
    return RadialProfile(lambda r: rho_c * ((1 + (r / r_c) ** 2) ** (-1.5 * beta)))


def hernquist_density_profile(M_0, a):
    """Hernquist (1990) density; M_0 in Msun, a in kpc (radial_profiles.py:173-186)."""
    return RadialProfile(lambda r: M_0 / (2.0 * np.pi * a**3) / ((r / a) * (1.0 + r / a) ** 3))


def hernquist_mass_profile(M_0, a):
    return RadialProfile(lambda r: M_0 * r**2 / (r + a) ** 2)


def nfw_density_profile(rho_s, r_s):
    """NFW density; rho_s in Msun/kpc**3, r_s in kpc (radial_profiles.py:249-262)."""
    return RadialProfile(lambda r: rho_s / ((r / r_s) * (1.0 + r / r_s) ** 2))


def nfw_mass_profile(rho_s, r_s):
    def _nfw(r):
        x = r / r_s
        return 4 * np.pi * rho_s * r_s**3 * (np.log(1 + x) - x / (1 + x))

    return RadialProfile(_nfw)


def vikhlinin_density_profile(rho_0, r_c, r_s, alpha, beta, epsilon, gamma=None):
    """Vikhlinin et al. (2006) modified beta model (radial_profiles.py:605-637)."""
    if gamma is None:
        gamma = 3.0
    return RadialProfile(
        lambda r: rho_0
        * (r / r_c) ** (-0.5 * alpha)
        * (1.0 + (r / r_c) ** 2) ** (-1.5 * beta + 0.25 * alpha)
        * (1.0 + (r / r_s) ** gamma) ** (-0.5 * epsilon / gamma)
    )


def vikhlinin_temperature_profile(T_0, a, b, c, r_t, T_min, r_cool, a_cool):
    """Vikhlinin et al. (2006) temperature profile in keV (radial_profiles.py:640-671)."""

    def _temp(r):
        x = (r / r_cool) ** a_cool
        t = (r / r_t) ** (-a) / ((1.0 + (r / r_t) ** b) ** (c / b))
        return T_0 * t * (x + T_min / T_0) / (x + 1)

    return RadialProfile(_temp)
