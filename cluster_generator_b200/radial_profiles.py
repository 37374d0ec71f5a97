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


# --------------------------------------------------------------------------- #
# the remaining closed-form families of radial_profiles.py (inputs to model construction)
# --------------------------------------------------------------------------- #
def _nfw_mu(c):
    """ln(1 + c) - c / (1 + c): the NFW mass shape function."""
    return np.log1p(c) - c / (1.0 + c)


def cored_hernquist_density_profile(M_0, a, b):
    """Hernquist density with the central cusp softened on the scale a / b (radial_profiles.py:189-209)."""
    norm = M_0 * b / (2.0 * np.pi * a**3)
    return RadialProfile(lambda r: norm / ((1.0 + b * (r / a)) * (1.0 + r / a) ** 3))


def convert_nfw_to_hernquist(M_200, r_200, conc):
    """Hernquist (M_0, a) with the same mass inside r_200 and the same inner density as an NFW
    halo of concentration ``conc`` (radial_profiles.py:228-246)."""
    a = r_200 / (np.sqrt(0.5 * conc**2 / _nfw_mu(conc)) - 1.0)
    return M_200 * ((r_200 + a) / r_200) ** 2, a


def tnfw_density_profile(rho_s, r_s, r_t):
    """NFW truncated by 1 / (1 + (r / r_t)^2) (radial_profiles.py:316-336)."""
    return RadialProfile(lambda r: rho_s / ((r / r_s) * (1.0 + r / r_s) ** 2 * (1.0 + (r / r_t) ** 2)))


def tnfw_mass_profile(rho_s, r_s, r_t):
    """Enclosed mass of the truncated NFW profile (radial_profiles.py:339-366 integrates the same
    integrand symbolically; here by adaptive quadrature per radius)."""
    from scipy.integrate import quad

    t = r_t / r_s

    def _shape(x):
        return quad(lambda y: y / ((1.0 + y) ** 2 * (1.0 + (y / t) ** 2)), 0.0, x, epsabs=0.0, epsrel=1e-12)[0]

    return RadialProfile(lambda r: 4.0 * np.pi * rho_s * r_s**3 * np.vectorize(_shape)(np.asarray(r, dtype=np.float64) / r_s))


def snfw_density_profile(M, a):
    """"Super-NFW" density, total mass M, scale a (radial_profiles.py:369-386)."""
    norm = 3.0 * M / (16.0 * np.pi * a**3)
    return RadialProfile(lambda r: norm / ((r / a) * (1.0 + r / a) ** 2.5))


def snfw_mass_profile(M, a):
    """Enclosed mass of the sNFW profile (radial_profiles.py:389-406)."""
    def _mass(r):
        x = r / a
        return M * (1.0 - (1.0 + 1.5 * x) / (1.0 + x) ** 1.5)

    return RadialProfile(_mass)


def snfw_total_mass(mass, radius, a):
    """Total sNFW mass for which ``mass`` lies inside ``radius`` (radial_profiles.py:409-425)."""
    return mass / snfw_mass_profile(1.0, a)(radius)


def snfw_conc(conc_nfw):
    """sNFW concentration matching an NFW concentration (radial_profiles.py:482-494)."""
    return 0.76 * conc_nfw + 1.36


def cored_snfw_density_profile(M, a, r_c):
    """sNFW with a core of radius r_c (radial_profiles.py:428-450)."""
    b = a / r_c
    norm = 3.0 * M * b / (16.0 * np.pi * a**3)
    return RadialProfile(lambda r: norm / ((1.0 + b * (r / a)) * (1.0 + r / a) ** 2.5))


def cored_snfw_mass_profile(M, a, r_c):
    """Enclosed mass of the cored sNFW profile (radial_profiles.py:453-479); complex arithmetic
    covers both b = a / r_c < 1 and b > 1 with one expression."""
    b = a / r_c

    def _mass(r):
        y = np.sqrt(np.asarray(r, dtype=np.float64) / a + 1.0).astype(np.complex128)
        d = np.sqrt(np.complex128(b / (1.0 - b)))
        t = (1.0 - 1.0 / y) * (b - 2.0) / (b - 1.0) ** 2 + (y**-3 - 1.0) / (3.0 * (b - 1.0))
        t = t + d * (np.arctan(y * d) - np.arctan(d)) / (b * (b - 1.0) ** 2)
        return 1.5 * M * b * np.real(t)

    return RadialProfile(_mass)


def cored_snfw_total_mass(mass, radius, a, r_c):
    """radial_profiles.py:497-515"""
    return mass / cored_snfw_mass_profile(1.0, a, r_c)(radius)


def _einasto_dn(n):
    return 3.0 * n - 1.0 / 3.0 + 8.0 / (1215.0 * n) + 184.0 / (229635.0 * n * n)


def einasto_density_profile(M, r_s, n):
    """Einasto density of total mass M, half-mass-like scale r_s, index n (radial_profiles.py:521-546)."""
    from scipy.special import gamma

    h = r_s / _einasto_dn(n) ** n
    rho_0 = M / (4.0 * np.pi * h**3 * n * gamma(3.0 * n))
    return RadialProfile(lambda r: rho_0 * np.exp(-((r / h) ** (1.0 / n))))


def einasto_mass_profile(M, r_s, n):
    """radial_profiles.py:549-573"""
    from scipy.special import gammainc

    h = r_s / _einasto_dn(n) ** n
    return RadialProfile(lambda r: M * gammainc(3.0 * n, (r / h) ** (1.0 / n)))


def am06_density_profile(rho_0, a, a_c, c, n):
    """Ascasibar & Markevitch (2006) gas density (radial_profiles.py:576-602)."""
    alpha = -1.0 - n * (c - 1.0) / (c - a / a_c)
    beta = 1.0 - n * (1.0 - a / a_c) / (c - a / a_c)
    return RadialProfile(
        lambda r: rho_0 * (1.0 + r / a_c) * (1.0 + r / (a_c * c)) ** alpha * (1.0 + r / a) ** beta)


def am06_temperature_profile(T_0, a, a_c, c):
    """Ascasibar & Markevitch (2006) temperature (radial_profiles.py:674-692)."""
    return RadialProfile(lambda r: T_0 * (c + r / a_c) / ((1.0 + r / a) * (1.0 + r / a_c)))


def baseline_entropy_profile(K_0, K_200, r_200, alpha):
    """K_0 + K_200 (r / r_200)^alpha (radial_profiles.py:695-712)."""
    return RadialProfile(lambda r: K_0 + K_200 * (r / r_200) ** alpha)


def broken_entropy_profile(r_s, K_scale, alpha, K_0=0.0):
    """radial_profiles.py:715-721: slope alpha inside r_s, 1.1 outside."""
    def _entr(r):
        x = r / r_s
        return K_scale * (K_0 + x**alpha * (1.0 + x**5) ** (0.2 * (1.1 - alpha)))

    return RadialProfile(_entr)


def walker_entropy_profile(r_200, A, B, K_scale, alpha=1.1):
    """Walker et al. (2012) entropy with the outer flattening (radial_profiles.py:724-729)."""
    def _entr(r):
        x = r / r_200
        return K_scale * A * x**alpha * np.exp(-((x / B) ** 2))

    return RadialProfile(_entr)


def rescale_profile_by_mass(profile, mass, radius):
    """Scale a density profile so that it holds ``mass`` inside ``radius`` (radial_profiles.py:732-762)."""
    from scipy.integrate import quad

    inside = 4.0 * np.pi * quad(lambda r: profile(r) * r * r, 0.0, radius)[0]
    return (mass / inside) * profile
