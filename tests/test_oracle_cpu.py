"""CPU suite, part 1: the oracle pinned against everything the reference (and its
third-party arithmetic) lets us pin here -- see SURVEY.md section 8(c)."""
import ctypes
import os
import warnings

import numpy as np
import pytest
from numpy.random import RandomState
from scipy.integrate import quad
from scipy.interpolate import InterpolatedUnivariateSpline as IUS

from tests.configs import oracle_model

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DP = ctypes.POINTER(ctypes.c_double)
IP = ctypes.POINTER(ctypes.c_int)


# ---- the reference's own known answers ------------------------------------------------ #
def test_integrate_mass_hernquist_known_answer(oracle):
    """cluster_generator/tests/test_radial_profiles.py:177-194"""
    rr = np.geomspace(0.1, 1.0e4, 1000)
    m = oracle.integrate_mass(oracle.hernquist_density_profile(1.0, 500.0), rr)
    np.testing.assert_allclose(m, (rr / (500.0 + rr)) ** 2, rtol=1e-5)


def test_mt19937_stream_known_answer():
    """RandomState(25) as seeded at tests/test_particles.py:27"""
    np.testing.assert_allclose(RandomState(25).uniform(size=5),
                               [0.87012414, 0.58227693, 0.27883894, 0.18591123, 0.41110013], atol=5e-9)


def test_philox4x32_10_random123_known_answers(oracle):
    kat = [((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
           ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
           ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
            (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))]
    for ctr, key, out in kat:
        assert oracle.philox4x32_10(ctr, key) == out


def test_philox_uniforms_are_in_unit_interval_and_52_bit(oracle):
    u = oracle.philox_uniforms(7, oracle.STREAM_POS_DIR, 0, 100_000)
    assert u.min() >= 0.0 and u.max() < 1.0
    assert np.all(u * 2.0**52 == np.floor(u * 2.0**52))
    assert abs(u.mean() - 0.5) < 0.005 and abs(np.corrcoef(u[:, 0], u[:, 1])[0, 1]) < 0.01


# ---- third-party arithmetic: FITPACK --------------------------------------------------- #
def test_c_splev_and_splder_are_bit_equal_to_fitpack(oracle):
    rng = RandomState(1)
    x = np.sort(rng.uniform(0.4, 9.0, 1000))
    y = np.exp(rng.normal(size=1000) * 0.1) * x**3
    s = IUS(x, y)
    t, c, _ = s._eval_args
    t, c = np.ascontiguousarray(t), np.ascontiguousarray(c[:1000])
    xx = rng.uniform(-0.5, 10.0, 50_000)  # includes extrapolation on both sides
    np.testing.assert_array_equal(oracle.splev_c(t, c, xx), s(xx))
    lib = oracle.clib()
    lib.cgo_splder1.argtypes = [DP, ctypes.c_int64, DP, DP, DP, ctypes.c_int64]
    yd = np.empty_like(xx)
    lib.cgo_splder1(oracle._dptr(t), t.size, oracle._dptr(c), oracle._dptr(xx), oracle._dptr(yd), xx.size)
    np.testing.assert_array_equal(yd, s(xx, 1))


@pytest.mark.parametrize("name", ["hernquist", "nfw_bcg"])
def test_c_spline_fit_reproduces_fitpack(oracle, name):
    om = oracle_model(name)
    x = np.ascontiguousarray(-om["gravitational_potential"][::-1])
    y = np.ascontiguousarray(om["dark_matter_density"][::-1])
    t_ref, c_ref, _ = IUS(x, y)._eval_args
    t, c = np.empty(x.size + 4), np.empty(x.size)
    assert oracle.clib().cgo_spline_fit(oracle._dptr(x), oracle._dptr(y), x.size, oracle._dptr(t),
                                        oracle._dptr(c)) == 0
    np.testing.assert_array_equal(t, t_ref)
    np.testing.assert_allclose(c, c_ref[: x.size], rtol=1e-13)


@pytest.mark.parametrize("name", ["hernquist", "nfw_bcg", "vikhlinin"])
def test_c_fpcurf_givens_fit_is_bit_equal_to_fitpack(oracle, name):
    """fpcurf's own solve (Givens rotations + fpback) restated: coefficients equal scipy's bits."""
    om = oracle_model(name)
    for x, y in ((-om["gravitational_potential"][::-1], om["dark_matter_density"][::-1]),
                 (om["radius"], -om["gravitational_potential"])):
        x, y = np.ascontiguousarray(x), np.ascontiguousarray(y)
        t_ref, c_ref, _ = IUS(x, y)._eval_args
        t, c = np.empty(x.size + 4), np.empty(x.size)
        assert oracle.clib().cgo_spline_fit_fitpack(oracle._dptr(x), oracle._dptr(y), x.size,
                                                    oracle._dptr(t), oracle._dptr(c)) == 0
        np.testing.assert_array_equal(t, t_ref)
        np.testing.assert_array_equal(c, c_ref[: x.size])


# ---- third-party arithmetic: QUADPACK -------------------------------------------------- #
def _c_quad(oracle, f, a, b, eps, limit=50):
    CB = ctypes.CFUNCTYPE(ctypes.c_double, ctypes.c_double, ctypes.c_void_p)
    lib = oracle.clib()
    lib.cgo_qags_cb.argtypes = [CB, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                ctypes.c_int, DP, DP, IP, IP, DP, DP, DP, DP, IP]
    cb = CB(lambda x, _: f(x))
    res, err, last, ier = ctypes.c_double(), ctypes.c_double(), ctypes.c_int(), ctypes.c_int()
    lists = [np.zeros(64) for _ in range(4)]
    iord = np.zeros(64, dtype=np.int32)
    lib.cgo_qags_cb(cb, a, b, eps, eps, limit, ctypes.byref(res), ctypes.byref(err), ctypes.byref(last),
                    ctypes.byref(ier), *[oracle._dptr(x) for x in lists], iord.ctypes.data_as(IP))
    return res.value, err.value, last.value, ier.value, lists


INTEGRANDS = [
    (lambda x: np.sqrt(abs(x - 0.3)), 0, 1), (lambda x: np.exp(-x * x), 0, 3),
    (lambda x: 1 / np.sqrt(x) if x > 0 else 0.0, 0, 1), (lambda x: np.log(x) if x > 0 else 0.0, 0, 1),
    (lambda x: np.sin(50 * x) * x, 0, 2), (lambda x: abs(x - 0.123) ** 0.1, 0, 1),
    (lambda x: 1 / (1e-4 + (x - 0.5) ** 2), 0, 1), (lambda x: x**-0.9 if x > 0 else 0.0, 0, 1),
    (lambda x: np.cos(200 * x), 0, 1), (lambda x: 1.0 if x < 0.3333 else 0.0, 0, 1),
    (lambda x: np.sqrt(x) * np.log(x) if x > 0 else 0.0, 0, 1), (lambda x: 1 / (x - 0.5) if x != 0.5 else 0.0, 0, 1.1),
]


@pytest.mark.parametrize("eps", [1.49e-8, 1.49e-5, 1e-12])
def test_qags_restatement_matches_quadpack_decision_for_decision(oracle, eps):
    """Bisection order, epsilon extrapolation, roundoff flags: every list scipy exposes
    through full_output must be reproduced bit for bit."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for f, a, b in INTEGRANDS:
            r = quad(f, a, b, full_output=1, epsabs=eps, epsrel=eps)
            info = r[2]
            res, err, last, ier, (al, bl, rl, el) = _c_quad(oracle, f, a, b, eps)
            L = info["last"]
            assert last == L
            assert res == r[0] and err == r[1]
            for mine, ref in ((al, "alist"), (bl, "blist"), (rl, "rlist"), (el, "elist")):
                np.testing.assert_array_equal(mine[:L], info[ref][:L])
            assert (ier != 0) == (len(r) > 3)


# ---- D / C against scipy on the real tables -------------------------------------------- #
def test_c_eddington_integrals_are_bit_equal_to_scipy(oracle):
    om = oracle_model("hernquist")
    ov = oracle.VirialEquilibrium(om, "dark_matter")   # scipy quad loop, virial.py:54-63
    ee = np.ascontiguousarray(ov.ee)
    t, c, _ = IUS(ee, om["dark_matter_density"][::-1])._eval_args
    t, c = np.ascontiguousarray(t), np.ascontiguousarray(c[: ee.size])
    g = np.empty(ee.size)
    last = np.zeros(ee.size, dtype=np.int32)
    ier = np.zeros(ee.size, dtype=np.int32)
    oracle.clib().cgo_eddington_g(oracle._dptr(ee), ee.size, oracle._dptr(t), oracle._dptr(c), 1.49e-5,
                                  1.49e-5, oracle._dptr(g), last.ctypes.data_as(IP), ier.ctypes.data_as(IP))
    np.testing.assert_array_equal(g, ov.g)
    assert ier.max() == 0 and 1 <= last.min() and last.max() <= 10
    assert np.all(ov.ff > 0)


def test_c_check_virial_is_bit_equal_to_scipy_on_sampled_bins(oracle):
    om = oracle_model("hernquist")
    ov = om.dm_virial
    ee = np.ascontiguousarray(ov.ee)
    t, c, _ = ov.f._eval_args
    t, c = np.ascontiguousarray(t), np.ascontiguousarray(c[: ee.size])
    rho = np.empty(ee.size)
    oracle.clib().cgo_check_virial_rho(oracle._dptr(ee), ee.size, oracle._dptr(t), oracle._dptr(c),
                                       oracle._dptr(rho), None, None)
    rho_int = lambda e, psi: ov.f(e) * np.sqrt(2 * (psi - e))   # virial.py:96
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(3, ee.size, 97):
            assert rho[i] == 4.0 * np.pi * quad(rho_int, 0.0, ee[i], args=(ee[i],))[0]
    chk = (rho[::-1] - om["dark_matter_density"]) / om["dark_matter_density"]
    assert np.all(chk < 1e-4)      # tests/test_models.py:116-130 (signed comparison)


# ---- golden fixtures (scripts/make_golden.py) ------------------------------------------ #
def _golden_model(oracle, name):
    npts = 300
    if name == "hernquist":
        return oracle.no_gas(0.1, 1.0e4, oracle.hernquist_density_profile(1.0e15, 500.0), num_points=npts)
    if name == "nfw_bcg":
        star = oracle.hernquist_density_profile(1.0e12, 20.0)
        tot = oracle.add_profiles(oracle.nfw_density_profile(1.0e7, 100.0), star)
        return oracle.no_gas(0.1, 1.0e4, tot, stellar_density=star, num_points=npts)
    raise KeyError(name)


@pytest.mark.parametrize("name,ptype,key", [("hernquist", "dark_matter", "dm"), ("nfw_bcg", "stellar", "star")])
def test_oracle_reproduces_golden_vectors(oracle, name, ptype, key):
    z = np.load(os.path.join(GOLDEN, f"{name}_{key}.npz"))
    om = _golden_model(oracle, name)
    np.testing.assert_allclose(om["gravitational_potential"], z["potential"], rtol=1e-13)
    v = oracle.VirialEquilibrium(om, ptype)
    np.testing.assert_allclose(v.g, z["g"], rtol=1e-9)
    np.testing.assert_allclose(v.ff, z["ff"], rtol=1e-8)
    seed, n = int(z["seed"]), z["p_radius"].size
    f = v.generate_particles(n, uniforms=oracle.device_rng_uniforms(seed, 0, n), compute_potential=True,
                             velocity_fn=lambda *a: oracle.generate_velocities_counter(*a, seed=seed))
    np.testing.assert_array_equal(v.last_trials, z["p_trials"])
    np.testing.assert_allclose(f[key, "particle_position"], z["p_pos"], rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(f[key, "particle_velocity"], z["p_vel"], rtol=1e-10, atol=1e-12)


def test_golden_qags_traces_match_c_restatement(oracle):
    """scipy's adaptive state frozen at fixture time vs the C restatement today."""
    z = np.load(os.path.join(GOLDEN, "qags_traces.npz"))
    om = _golden_model(oracle, "hernquist")
    v = oracle.VirialEquilibrium(om, "dark_matter")
    ee = np.ascontiguousarray(v.ee)
    t, c, _ = IUS(ee, om["dark_matter_density"][::-1])._eval_args
    tf, cf, _ = v.f._eval_args
    lib = oracle.clib()
    for key in z.files:
        kind, b = (0 if key[0] == "d" else 1), int(key[1:])
        tt, cc = (t, c) if kind == 0 else (tf, cf)
        tt, cc = np.ascontiguousarray(tt), np.ascontiguousarray(cc[: ee.size])
        res, err, last, ier = ctypes.c_double(), ctypes.c_double(), ctypes.c_int(), ctypes.c_int()
        lists = [np.zeros(64) for _ in range(4)]
        upper = np.sqrt(ee[b]) if kind == 0 else ee[b]
        lib.cgo_qags(oracle._dptr(tt), tt.size, oracle._dptr(cc), ee[b], 0.0, upper,
                     1.49e-5 if kind == 0 else 1.49e-8, kind, ctypes.byref(res), ctypes.byref(err),
                     ctypes.byref(last), ctypes.byref(ier), *[oracle._dptr(x) for x in lists])
        ref = z[key]
        L = int(ref[2])
        assert last.value == L, key
        np.testing.assert_allclose(res.value, ref[0], rtol=1e-12, err_msg=key)
        np.testing.assert_allclose(lists[0][:L], ref[3:3 + L], rtol=1e-15, err_msg=key)      # alist
        np.testing.assert_allclose(lists[2][:L], ref[3 + 2 * L:3 + 3 * L], rtol=1e-9, atol=1e-300, err_msg=key)


# ---- V: the restated rejection loop ----------------------------------------------------- #
def test_rejection_loop_invariants_and_mode_agreement(oracle):
    """tests/test_particles.py:95-152 invariants + mode A (sequential stream) and
    mode B (counter) sample the same distribution."""
    from scipy.stats import ks_2samp

    om = oracle_model("hernquist")
    ov = om.dm_virial
    n = 100_000
    fa = ov.generate_particles(n, prng=25)
    va, ta = ov.last["speed"].copy(), ov.last_trials.copy()
    fb = ov.generate_particles(n, prng=26, velocity_fn=lambda *a: oracle.generate_velocities_counter(*a, seed=3))
    vb, psi_b = ov.last["speed"], ov.last["psi"]
    for f in (fa, fb):
        assert f["dm", "particle_position"].shape == (n, 3)
        assert all(np.isfinite(f["dm", k]).all() for k in ("particle_position", "particle_velocity", "particle_mass"))
        assert (f["dm", "particle_mass"] > 0).all()
        assert np.linalg.norm(f["dm", "particle_position"], axis=1).max() < 50_000
    assert ks_2samp(va, vb).pvalue > 0.001
    assert np.all(vb**2 < 2.0 * psi_b)              # bound: v < v_esc
    assert 20 < ta.mean() < 60                      # Hernquist T-bar 31-51
    # literal host-uniform mode agrees with the definition on a tiny case
    u = RandomState(0).uniform(size=(50, 4000, 2))
    vel, tr = oracle.generate_velocities_array(ov.last["psi"][:50], ov.last["vesc"][:50], ov.last["fv2esc"][:50],
                                               ov.f.get_knots(), ov.f.get_coeffs(), 3, u)
    for i in range(50):
        k = tr[i] - 1
        v = u[i, k, 0] * ov.last["vesc"][i]
        assert ov.f(ov.last["psi"][i] - 0.5 * v * v) * v * v >= u[i, k, 1] * ov.last["fv2esc"][i]
        assert vel[i] == np.sqrt(v * v)


def test_sub_sample_and_gas_oracle_shapes(oracle):
    om = oracle_model("vikhlinin")
    f = om.dm_virial.generate_particles(4000, sub_sample=4, prng=1, compute_potential=True)
    r = np.linalg.norm(f["dm", "particle_position"], axis=1)
    np.testing.assert_allclose(r[:1000], r[1000:2000], rtol=1e-12)   # np.tile(radius_sub, 4)
    g = oracle.generate_gas_particles(om, 3000, prng=2, compute_potential=True)
    assert g["gas", "particle_velocity"].shape == (3000, 3) and not g["gas", "particle_velocity"].any()
    assert (g["gas", "density"] > 0).all() and (g["gas", "thermal_energy"] > 0).all()
