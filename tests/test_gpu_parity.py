"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Tolerances are BASELINE.json's: f(E) rtol 1e-8; coordinates on identical uniforms
fp64 rtol 1e-10 (+ atol 1e-12 * |r| resp. |v| for components that nearly vanish,
SURVEY.md section 7 hard part 5), fp32 rtol 1e-5; accept/reject decisions exact.
"""
import numpy as np
import pytest
import torch

from tests.configs import assert_tables_equal, oracle_model, product_model

pytestmark = pytest.mark.gpu

RTOL64 = 1e-10
RTOL32 = 1e-5


def _np(t):
    return t.detach().cpu().numpy()


def _vec_close(a, b, rtol, scale):
    """|a-b| <= rtol*|b| + 1e-12*scale (scale = |r| or |v| per particle)."""
    atol = (1e-12 if rtol < 1e-6 else 1e-7) * scale[:, None]
    bad = np.abs(a - b) > rtol * np.abs(b) + atol
    assert not bad.any(), f"{bad.sum()} components differ, worst {np.abs(a - b)[bad].max()}"


# --------------------------------------------------------------------------- #
# splines
# --------------------------------------------------------------------------- #
def test_spline_fit_and_eval_match_fitpack():
    from scipy.interpolate import InterpolatedUnivariateSpline as IUS

    from cluster_generator_b200.virial import DeviceSpline

    om = oracle_model("hernquist")
    x = -om["gravitational_potential"][::-1]
    y = om["dark_matter_density"][::-1]
    s_ref = IUS(x, y)
    s = DeviceSpline.fit(x, y, method="fitpack")
    t, c, k = s._eval_args
    np.testing.assert_array_equal(t, s_ref._eval_args[0])
    np.testing.assert_array_equal(c, s_ref.get_coeffs())          # fpcurf's Givens solve, bit for bit
    c_pcr = DeviceSpline.fit(x, y, method="pcr").get_coeffs()     # fast solver: same system
    np.testing.assert_allclose(c_pcr, s_ref.get_coeffs(), rtol=1e-12)
    np.testing.assert_array_equal(s.get_knots(), s_ref.get_knots())
    xx = np.random.RandomState(0).uniform(x[0], x[-1], 5000)
    np.testing.assert_allclose(s(xx), s_ref(xx), rtol=1e-10)
    np.testing.assert_allclose(s(xx, 1), s_ref(xx, 1), rtol=1e-9)
    # extrapolation (ext=0) below and above the knots amplifies the ~1e-15 coefficient noise
    xe = np.concatenate([np.linspace(0.0, x[0], 50), np.linspace(x[-1], 1.02 * x[-1], 20)])
    np.testing.assert_allclose(s(xe), s_ref(xe), rtol=1e-8)
    np.testing.assert_allclose(s(xe, 1), s_ref(xe, 1), rtol=1e-8)


def test_spline_rejects_non_increasing_x():
    from cluster_generator_b200.virial import DeviceSpline

    with pytest.raises(ValueError):
        DeviceSpline.fit(np.array([0.0, 1.0, 1.0, 2.0, 3.0]), np.zeros(5))


# --------------------------------------------------------------------------- #
# D: Eddington f(E)
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("name,ptype", [("hernquist", "dark_matter"), ("nfw_bcg", "dark_matter"),
                                        ("nfw_bcg", "stellar"), ("vikhlinin", "dark_matter")])
def test_eddington_df_matches_reference_arithmetic(oracle, name, ptype):
    om, pm = oracle_model(name), product_model(name)
    assert_tables_equal(om, pm)
    ov = oracle.VirialEquilibrium(om, ptype)      # scipy quad + FITPACK, as virial.py:43-68
    pv = pm.dm_virial if ptype == "dark_matter" else pm.star_virial
    np.testing.assert_allclose(pv.g, ov.g, rtol=1e-9)
    np.testing.assert_allclose(pv.df.d, ov.df, rtol=1e-8)          # north_star: f(E) rtol 1e-8
    np.testing.assert_allclose(pv.ff, ov.ff, rtol=1e-8)
    # with FITPACK's own solve and QUADPACK's own decisions the numbers are scipy's to the bit
    print(name, ptype, "bit-equal g", (pv.g == ov.g).mean(), "ff", (pv.ff == ov.ff).mean(),
          "f coeffs", (pv.f.get_coeffs() == ov.f.get_coeffs()).mean())
    np.testing.assert_array_equal(pv.g, ov.g)
    np.testing.assert_array_equal(pv.ff, ov.ff)
    np.testing.assert_array_equal(pv.f.get_coeffs(), ov.f.get_coeffs())
    # decision-for-decision: sub-interval counts equal the C QAGS restatement's
    import ctypes
    from scipy.interpolate import InterpolatedUnivariateSpline as IUS
    ee = np.ascontiguousarray(ov.ee)
    t, c, _ = IUS(ee, om[f"{ptype}_density"][::-1])._eval_args
    t, c = np.ascontiguousarray(t), np.ascontiguousarray(c[: ee.size])
    g = np.empty(ee.size)
    last = np.zeros(ee.size, dtype=np.int32)
    ier = np.zeros(ee.size, dtype=np.int32)
    ip = ctypes.POINTER(ctypes.c_int)
    oracle.clib().cgo_eddington_g(oracle._dptr(ee), ee.size, oracle._dptr(t), oracle._dptr(c), 1.49e-5,
                                  1.49e-5, oracle._dptr(g), last.ctypes.data_as(ip), ier.ctypes.data_as(ip))
    np.testing.assert_array_equal(g, ov.g)   # the C restatement is bit-equal to scipy
    np.testing.assert_array_equal(pv.qags_last, last)  # same spline, same decisions
    np.testing.assert_array_equal(pv.qags_ier, ier)


def test_eddington_df_supplied(oracle):
    om, pm = oracle_model("hernquist"), product_model("hernquist")
    from cluster_generator_b200 import VirialEquilibrium

    ov = oracle.VirialEquilibrium(om, "dark_matter")
    pv = VirialEquilibrium(pm, "dark_matter", df=ov.df)
    x = np.linspace(ov.ee[0], ov.ee[-1], 1000)
    np.testing.assert_allclose(pv.f(x), ov.f(x), rtol=1e-10)
    np.testing.assert_array_equal(pv.df.d, ov.df)


# --------------------------------------------------------------------------- #
# C: check_virial
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("name", ["hernquist", "nfw_bcg"])
def test_check_virial_matches_reference_arithmetic(oracle, name):
    import ctypes

    om, pm = oracle_model(name), product_model(name)
    ov = oracle.VirialEquilibrium(om, "dark_matter")
    pv = pm.dm_virial
    rho, chk = pv.check_virial()
    # oracle: C restatement (bit-equal to scipy's quad, see tests/test_oracle_cpu.py) on
    # the oracle's own f spline
    ee = np.ascontiguousarray(ov.ee)
    t, c, _ = ov.f._eval_args
    t, c = np.ascontiguousarray(t), np.ascontiguousarray(c[: ee.size])
    rho_o = np.empty(ee.size)
    oracle.clib().cgo_check_virial_rho(oracle._dptr(ee), ee.size, oracle._dptr(t), oracle._dptr(c),
                                       oracle._dptr(rho_o), None, None)
    rho_o = rho_o[::-1]
    # (a) same f spline in, same numbers out: the device QAGS against the C restatement
    from cluster_generator_b200 import _native as N
    n = ee.size
    dev_rho = torch.empty(n, dtype=torch.float64, device="cuda")
    dev_chk = torch.empty(n, dtype=torch.float64, device="cuda")
    last = torch.empty(n, dtype=torch.int32, device="cuda")
    ier = torch.empty(n, dtype=torch.int32, device="cuda")
    last_o = np.zeros(n, dtype=np.int32)
    ier_o = np.zeros(n, dtype=np.int32)
    ip = ctypes.POINTER(ctypes.c_int)
    oracle.clib().cgo_check_virial_rho(oracle._dptr(ee), n, oracle._dptr(t), oracle._dptr(c),
                                       oracle._dptr(np.empty(n)), last_o.ctypes.data_as(ip),
                                       ier_o.ctypes.data_as(ip))
    pd = np.ascontiguousarray(om["dark_matter_density"][::-1])
    ee_d, pd_d, t_d, c_d = N.dev_f64(ee), N.dev_f64(pd), N.dev_f64(t), N.dev_f64(c)  # keep alive
    N.check(N.lib().cgb_check_virial(N.ptr(ee_d), N.ptr(pd_d), n, N.ptr(t_d), N.ptr(c_d),
                                     N.ptr(dev_rho), N.ptr(dev_chk), N.ptr(last), N.ptr(ier),
                                     N.stream_ptr()))
    print("same-spline: last mismatches", (_np(last) != last_o).sum(), "ier mismatches",
          (_np(ier) != ier_o).sum(), "rho bit-equal", (_np(dev_rho) == rho_o).sum(), "of", n,
          "max rel", np.abs(_np(dev_rho) / rho_o - 1).max())
    np.testing.assert_array_equal(_np(last), last_o)
    np.testing.assert_array_equal(_np(ier), ier_o)
    np.testing.assert_allclose(_np(dev_rho), rho_o, rtol=1e-13)
    chk_o = (rho_o - om["dark_matter_density"]) / om["dark_matter_density"]
    # (b) end to end: the device f spline is FITPACK's to the bit, so the check is too
    np.testing.assert_array_equal(rho, rho_o)
    assert np.all(chk < 1e-4)                               # test_models.py:127-130 (signed)
    assert np.all(chk_o < 1e-4)


# --------------------------------------------------------------------------- #
# V: the Cython seam mirror
# --------------------------------------------------------------------------- #
def _seam_inputs(oracle, name, ptype, n, seed=11):
    om = oracle_model(name)
    ov = om.dm_virial if ptype == "dark_matter" else om.star_virial
    u = oracle.device_rng_uniforms(seed, 0, n)
    ov.generate_particles(n, uniforms=u,
                          velocity_fn=lambda *a: oracle.generate_velocities_counter(*a, seed=seed))
    return ov, ov.last, ov.last_trials


def _call_seam(ov, last, n, seed, u_trial=None, first_id=0):
    from cluster_generator_b200 import _native as N

    dev = N.require_cuda()
    knots, coeffs = ov.f.get_knots(), ov.f.get_coeffs()
    psi, vesc, env = (N.dev_f64(last[k]) for k in ("psi", "vesc", "fv2esc"))
    kd, cd = N.dev_f64(knots), N.dev_f64(coeffs)
    vel = torch.empty(n, dtype=torch.float64, device=dev)
    trials = torch.empty(n, dtype=torch.int32, device=dev)
    nb = N.lib().cgb_generate_velocities_workspace_bytes(coeffs.size)
    work = N.workspace(nb)
    ut = None if u_trial is None else N.dev_f64(u_trial)
    tcap = 0 if u_trial is None else u_trial.shape[1]
    N.check(N.lib().cgb_generate_velocities(N.ptr(psi), N.ptr(vesc), N.ptr(env), N.ptr(kd), knots.size,
                                            N.ptr(cd), coeffs.size, 3, N.ptr(ut), tcap, seed, first_id,
                                            N.ptr(vel), N.ptr(trials), n, N.ptr(work), nb, N.stream_ptr()))
    return _np(vel), _np(trials)


@pytest.mark.parametrize("name,ptype,n", [("hernquist", "dark_matter", 100_000),
                                          ("nfw_bcg", "stellar", 20_000)])
def test_generate_velocities_counter_mode_exact(oracle, name, ptype, n):
    ov, last, trials_o = _seam_inputs(oracle, name, ptype, n)
    vel, trials = _call_seam(ov, last, n, 11)
    np.testing.assert_array_equal(trials, trials_o)          # accept/reject decisions identical
    np.testing.assert_allclose(vel, last["speed"], rtol=RTOL64)


def test_generate_velocities_long_tail_goes_through_the_grid_wide_pass(oracle):
    """Envelopes inflated by 1e4..3e4 => 1e5..1e6 trials for a few particles: these leave
    the warp loop and are finished by the grid-wide second pass; the accepted trial must
    still be the first accepted one in sequence order."""
    n = 3000
    ov, last, _ = _seam_inputs(oracle, "hernquist", "dark_matter", n)
    last = {k: v.copy() for k, v in last.items()}
    idx = np.arange(7, n, 150)
    last["fv2esc"][idx] *= np.linspace(1.0e3, 6.0e4, idx.size)
    vel_o, tr_o = oracle.generate_velocities_counter(last["psi"], last["vesc"], last["fv2esc"],
                                                     ov.f.get_knots(), ov.f.get_coeffs(), 3, seed=99)
    assert tr_o[idx].min() > 1000 and (tr_o > 70_000).sum() >= 5   # both sides of the deferral threshold
    vel, tr = _call_seam(ov, last, n, 99)
    np.testing.assert_array_equal(tr, tr_o)
    np.testing.assert_allclose(vel, vel_o, rtol=RTOL64)


def test_generate_velocities_host_uniform_array(oracle):
    n, tcap = 4000, 600
    ov, last, _ = _seam_inputs(oracle, "hernquist", "dark_matter", n)
    u = np.random.RandomState(3).uniform(size=(n, tcap, 2))
    vel_o, tr_o = oracle.generate_velocities_array(last["psi"], last["vesc"], last["fv2esc"],
                                                   ov.f.get_knots(), ov.f.get_coeffs(), 3, u)
    vel, tr = _call_seam(ov, last, n, 0, u_trial=u)
    np.testing.assert_array_equal(tr, tr_o)
    ok = tr_o > 0
    assert ok.sum() > 0.9 * n and (~ok).sum() > 0             # both accepted and capped cases seen
    np.testing.assert_allclose(vel[ok], vel_o[ok], rtol=RTOL64)
    assert np.isnan(vel[~ok]).all()


def test_generate_velocities_rejects_bad_arguments():
    from cluster_generator_b200 import _native as N

    w = N.workspace(1 << 20)
    d = torch.zeros(1000, dtype=torch.float64, device="cuda")
    rc = N.lib().cgb_generate_velocities(N.ptr(d), N.ptr(d), N.ptr(d), N.ptr(d), 998, N.ptr(d), 1000, 2,
                                         None, 0, 0, 0, N.ptr(d), None, 10, N.ptr(w), 1 << 20, None)
    with pytest.raises(ValueError, match="k=3"):
        N.check(rc)


# --------------------------------------------------------------------------- #
# P: fused sampler
# --------------------------------------------------------------------------- #
def _oracle_particles(oracle, name, ptype, n, seed, **kw):
    om = oracle_model(name)
    ov = om.dm_virial if ptype == "dark_matter" else om.star_virial
    u = oracle.device_rng_uniforms(seed, 0, n)
    f = ov.generate_particles(n, uniforms=u,
                              velocity_fn=lambda *a: oracle.generate_velocities_counter(*a, seed=seed), **kw)
    return ov, f


def _device_particles(name, ptype, n, seed, **kw):
    """Device-RNG run with an explicit Philox seed (bypassing the prng->seed draw)."""
    import cluster_generator_b200.virial as V

    pm = product_model(name)
    pv = pm.dm_virial if ptype == "dark_matter" else pm.star_virial
    orig = V.seed_from_prng
    V.seed_from_prng = lambda prng, species: seed
    try:
        return pv, pv.generate_particles(n, prng=0, **kw)
    finally:
        V.seed_from_prng = orig


@pytest.mark.parametrize("name,ptype,n", [("hernquist", "dark_matter", 100_000),
                                          ("nfw_bcg", "dark_matter", 50_000),
                                          ("nfw_bcg", "stellar", 20_000),
                                          ("vikhlinin", "dark_matter", 50_000)])
def test_fused_sampler_matches_oracle_on_identical_uniforms(oracle, name, ptype, n):
    key = {"dark_matter": "dm", "stellar": "star"}[ptype]
    ov, fo = _oracle_particles(oracle, name, ptype, n, 77, compute_potential=True)
    pv, p = _device_particles(name, ptype, n, 77, compute_potential=True)
    pos, vel = _np(p[key, "particle_position"]), _np(p[key, "particle_velocity"])
    assert pos.shape == (n, 3) and pos.strides == (8, 8 * n)      # SoA planes behind (n,3)
    r_o = ov.last["radius"]
    _vec_close(pos, fo[key, "particle_position"], RTOL64, r_o)
    _vec_close(vel, fo[key, "particle_velocity"], RTOL64, ov.last["speed"])
    np.testing.assert_allclose(np.linalg.norm(vel, axis=1), ov.last["speed"], rtol=1e-9)
    np.testing.assert_allclose(_np(p[key, "particle_mass"]), fo[key, "particle_mass"], rtol=1e-15)
    np.testing.assert_allclose(_np(p[key, "particle_potential"]), fo[key, "particle_potential"], rtol=RTOL64)
    d = p.diagnostics[key].as_dict()
    assert d["particles"] == n and d["failures"] == 0
    assert d["sum_trials"] == ov.last_trials.sum()                # every accept/reject decision
    assert d["max_trials"] == ov.last_trials.max()
    assert d["exact_evaluations"] < 0.05 * d["sum_trials"]        # the fp32 screen does the work


def test_fused_sampler_numpy_rng_reproduces_reference_draw_order(oracle):
    """rng='numpy': radii, positions and velocity directions come from the caller's
    RandomState in the reference's draw order (virial.py:162,175-176,213-214)."""
    from cluster_generator_b200.virial import PRIVATE_TRIAL_SEED

    n = 30_000
    om, pm = oracle_model("hernquist"), product_model("hernquist")
    ov = om.dm_virial
    fo = ov.generate_particles(
        n, prng=25,
        velocity_fn=lambda *a: oracle.generate_velocities_counter(*a, seed=PRIVATE_TRIAL_SEED))
    p = pm.generate_dm_particles(n, prng=25, rng="numpy")
    _vec_close(_np(p["dm", "particle_position"]), fo["dm", "particle_position"], RTOL64, ov.last["radius"])
    _vec_close(_np(p["dm", "particle_velocity"]), fo["dm", "particle_velocity"], RTOL64, ov.last["speed"])


def test_fused_sampler_fp32_outputs(oracle):
    n = 50_000
    ov, fo = _oracle_particles(oracle, "hernquist", "dark_matter", n, 5)
    pv, p = _device_particles("hernquist", "dark_matter", n, 5, dtype=torch.float32)
    pos = _np(p["dm", "particle_position"])
    assert pos.dtype == np.float32
    _vec_close(pos.astype(np.float64), fo["dm", "particle_position"], RTOL32, ov.last["radius"])
    _vec_close(_np(p["dm", "particle_velocity"]).astype(np.float64), fo["dm", "particle_velocity"], RTOL32,
               ov.last["speed"])
    assert p.diagnostics["dm"].as_dict()["sum_trials"] == ov.last_trials.sum()   # decisions stay fp64-exact


def test_sharded_generation_is_bit_identical():
    n = 100_003
    _, full = _device_particles("hernquist", "dark_matter", n, 9)
    parts = [_device_particles("hernquist", "dark_matter", n, 9, shard=(k, 3))[1] for k in range(3)]
    for f in ("particle_position", "particle_velocity", "particle_mass"):
        cat = torch.cat([q["dm", f] for q in parts])
        assert torch.equal(cat, full["dm", f]), f
    tot = sum(q.diagnostics["dm"].as_dict()["sum_trials"] for q in parts)
    assert tot == full.diagnostics["dm"].as_dict()["sum_trials"]


def test_sub_sample_tiles_radius_and_speed(oracle):
    n, s = 40_000, 4
    ov, fo = _oracle_particles(oracle, "hernquist", "dark_matter", n, 21, sub_sample=s, compute_potential=True)
    pv, p = _device_particles("hernquist", "dark_matter", n, 21, sub_sample=s, compute_potential=True)
    r_o = np.tile(ov.last["radius"][: n // s], s)
    _vec_close(_np(p["dm", "particle_position"]), fo["dm", "particle_position"], RTOL64, r_o)
    _vec_close(_np(p["dm", "particle_velocity"]), fo["dm", "particle_velocity"], RTOL64, ov.last["speed"])
    np.testing.assert_allclose(_np(p["dm", "particle_potential"]), fo["dm", "particle_potential"], rtol=RTOL64)
    with pytest.raises(ValueError):
        pv.generate_particles(10, sub_sample=3)


def test_r_max_and_edge_counts(oracle):
    om, pm = oracle_model("hernquist"), product_model("hernquist")
    for n in (0, 1, 31, 257):
        p = pm.generate_dm_particles(n, prng=1)
        assert p.num_particles["dm"] == n and p["dm", "particle_position"].shape == (n, 3)
    ov, fo = _oracle_particles(oracle, "hernquist", "dark_matter", 20_000, 4, r_max=2000.0)
    pv, p = _device_particles("hernquist", "dark_matter", 20_000, 4, r_max=2000.0)
    pos = _np(p["dm", "particle_position"])
    _vec_close(pos, fo["dm", "particle_position"], RTOL64, ov.last["radius"])
    assert np.linalg.norm(pos, axis=1).max() < 2000.0
    np.testing.assert_allclose(_np(p["dm", "particle_mass"]), fo["dm", "particle_mass"], rtol=1e-15)


def test_non_monotone_potential_raises_like_the_reference():
    import cluster_generator_b200 as cg

    pm = product_model("hernquist")
    fields = {k: np.array(v) for k, v in pm.fields.items()}
    fields["gravitational_potential"][500] = fields["gravitational_potential"][499]
    bad = cg.ClusterModel.from_arrays(fields)
    with pytest.raises(ValueError, match="increasing"):
        bad.dm_virial


# --------------------------------------------------------------------------- #
# Gs: gas particles
# --------------------------------------------------------------------------- #
def test_gas_particles_match_oracle(oracle):
    n = 100_000
    om, pm = oracle_model("vikhlinin"), product_model("vikhlinin")
    fo = oracle.generate_gas_particles(om, n, prng=25, compute_potential=True)
    p = pm.generate_gas_particles(n, prng=25, rng="numpy", compute_potential=True)
    pos = _np(p["gas", "particle_position"])
    _vec_close(pos, fo["gas", "particle_position"], RTOL64, np.linalg.norm(fo["gas", "particle_position"], axis=1))
    for f in ("thermal_energy", "density", "particle_mass", "particle_potential"):
        np.testing.assert_allclose(_np(p["gas", f]), fo["gas", f], rtol=1e-9, err_msg=f)
    assert torch.count_nonzero(p["gas", "particle_velocity"]) == 0
    assert p.num_particles["gas"] == n


def test_gas_and_tracer_sub_sample_and_tracers(oracle):
    n, s = 40_000, 4
    om, pm = oracle_model("vikhlinin"), product_model("vikhlinin")
    fo = oracle.generate_gas_particles(om, n, prng=8, sub_sample=s, compute_potential=True)
    p = pm.generate_gas_particles(n, prng=8, rng="numpy", sub_sample=s, compute_potential=True)
    pos = _np(p["gas", "particle_position"])
    _vec_close(pos, fo["gas", "particle_position"], RTOL64, np.linalg.norm(fo["gas", "particle_position"], axis=1))
    for f in ("thermal_energy", "density", "particle_mass", "particle_potential"):
        np.testing.assert_allclose(_np(p["gas", f]), fo["gas", f], rtol=1e-9, err_msg=f)
    r = np.linalg.norm(pos, axis=1)
    np.testing.assert_allclose(r[: n // s], r[n // s: 2 * n // s], rtol=1e-12)     # tiled radii
    with pytest.raises(ValueError):
        pm.generate_gas_particles(10, sub_sample=3)
    # tracers: gas-mass CDF, zero mass and velocity (model.py:766-836)
    t = pm.generate_tracer_particles(5000, prng=3)
    assert t.num_particles["tracer"] == 5000 and not t["tracer", "particle_mass"].any()
    assert not t["tracer", "particle_velocity"].any() and torch.isfinite(t["tracer", "particle_position"]).all()
    # sharded sub-sampled gas equals the single-call result
    full = pm.generate_gas_particles(n, prng=9, sub_sample=s)
    parts = [pm.generate_gas_particles(n, prng=9, sub_sample=s, shard=(k, 2)) for k in range(2)]
    for f in ("particle_position", "density"):
        assert torch.equal(torch.cat([q["gas", f] for q in parts]), full["gas", f])


# --------------------------------------------------------------------------- #
# reference test-suite invariants (tests/test_particles.py:95-186)
# --------------------------------------------------------------------------- #
def test_reference_particle_invariants():
    pm = product_model("nfw_bcg")
    prng = np.random.RandomState(25)
    dm = pm.generate_dm_particles(1000, prng=prng)
    st = pm.generate_star_particles(1000, prng=prng)
    assert dm.num_particles["dm"] == 1000 and st.num_particles["star"] == 1000
    allp = dm + st
    assert allp.num_particles == {"dm": 1000, "star": 1000}
    for pt in ("dm", "star"):
        for f in ("particle_position", "particle_velocity", "particle_mass"):
            assert torch.isfinite(allp[pt, f]).all()
        assert (allp[pt, "particle_mass"] > 0).all()
        assert torch.linalg.norm(allp[pt, "particle_position"], dim=1).max() < 50_000
    both = dm + dm
    assert both.num_particles["dm"] == 2000 and both["dm", "particle_position"].shape == (2000, 3)
    dm.add_offsets([100.0, 0.0, 0.0], [0.0, 1.0, 0.0])


# --------------------------------------------------------------------------- #
# statistics with the device RNG (KS on r, |v|, E)
# --------------------------------------------------------------------------- #
def test_device_rng_is_statistically_indistinguishable_from_oracle(oracle):
    from scipy.stats import ks_2samp

    n = 1_000_000
    om, pm = oracle_model("hernquist"), product_model("hernquist")
    ov = om.dm_virial
    fo = ov.generate_particles(n, prng=123)          # mode A: independent sequential stream
    p = pm.generate_dm_particles(n, prng=456, compute_potential=True)
    pos, vel = _np(p["dm", "particle_position"]), _np(p["dm", "particle_velocity"])
    r, v = np.linalg.norm(pos, axis=1), np.linalg.norm(vel, axis=1)
    E = -_np(p["dm", "particle_potential"]) - 0.5 * v * v
    E_o = ov.last["psi"] - 0.5 * ov.last["speed"] ** 2
    for a, b, what in ((r, ov.last["radius"], "r"), (v, ov.last["speed"], "|v|"), (E, E_o, "E")):
        assert ks_2samp(a, b).pvalue > 0.01, what
    # isotropy
    assert abs((pos[:, 2] / r).mean()) < 4 / np.sqrt(3 * n)
    assert abs(np.arctan2(pos[:, 1], pos[:, 0]).mean()) < 4 * np.pi / np.sqrt(3 * n)
    assert (E > 0).all()                              # bound particles only


def test_full_size_properties_config2():
    """BASELINE config 2 at full size (1e7 DM): size-independent properties."""
    n = 10_000_000
    pm = product_model("vikhlinin")
    p = pm.generate_dm_particles(n, prng=7, compute_potential=True)
    d = p.diagnostics["dm"].as_dict()
    assert d["particles"] == n and d["failures"] == 0
    assert 5.0 < d["mean_trials"] < 30.0              # SURVEY: T-bar ~ 11.7
    pos, vel = p["dm", "particle_position"], p["dm", "particle_velocity"]
    assert torch.isfinite(pos).all() and torch.isfinite(vel).all()
    r = torch.linalg.norm(pos, dim=1)
    assert r.max() < 1.0e4 * (1 + 1e-12)
    E = -p["dm", "particle_potential"] - 0.5 * (vel * vel).sum(dim=1)
    assert (E > 0).all()
    m = p["dm", "particle_mass"]
    np.testing.assert_allclose(float(m.sum()), pm.dm_virial.sampler_tables().mtot, rtol=1e-9)
    # idempotence: same prng -> same particles
    q = pm.generate_dm_particles(n, prng=7)
    assert torch.equal(q["dm", "particle_position"], pos)


# --------------------------------------------------------------------------- #
# committed golden vectors (tests/golden, scripts/make_golden.py): no oracle at run time
# --------------------------------------------------------------------------- #
import os

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _golden_product_model(name):
    import cluster_generator_b200 as cg

    if name == "hernquist":
        return cg.ClusterModel.no_gas(0.1, 1.0e4, cg.hernquist_density_profile(1.0e15, 500.0), num_points=300)
    if name == "nfw_bcg":
        star = cg.hernquist_density_profile(1.0e12, 20.0)
        return cg.ClusterModel.no_gas(0.1, 1.0e4, cg.nfw_density_profile(1.0e7, 100.0) + star,
                                      stellar_density=star, num_points=300)
    dens = cg.vikhlinin_density_profile(4.2e5, 100.0, 1000.0, 1.0, 0.67, 3.0)
    temp = cg.vikhlinin_temperature_profile(10.0, 0.0, 2.0, 1.0, 1500.0, 6.0, 80.0, 1.0)
    return cg.ClusterModel.from_dens_and_temp(0.1, 1.0e4, dens, temp, num_points=300)


@pytest.mark.parametrize("name,ptype,key", [("hernquist", "dark_matter", "dm"), ("nfw_bcg", "dark_matter", "dm"),
                                            ("nfw_bcg", "stellar", "star"), ("vikhlinin", "dark_matter", "dm")])
def test_cuda_path_reproduces_golden_vectors(name, ptype, key):
    import cluster_generator_b200.virial as V

    z = np.load(os.path.join(GOLDEN, f"{name}_{key}.npz"))
    pm = _golden_product_model(name)
    np.testing.assert_allclose(np.asarray(pm["gravitational_potential"]), z["potential"], rtol=1e-13)
    pv = pm.dm_virial if ptype == "dark_matter" else pm.star_virial
    np.testing.assert_allclose(pv.g, z["g"], rtol=1e-9)
    np.testing.assert_allclose(pv.ff, z["ff"], rtol=1e-8)
    rho, chk = pv.check_virial()
    np.testing.assert_allclose(rho, z["check_rho"], rtol=2e-4)
    seed, n = int(z["seed"]), z["p_radius"].size
    orig = V.seed_from_prng
    V.seed_from_prng = lambda prng, species: seed
    try:
        p = pv.generate_particles(n, prng=0, compute_potential=True)
    finally:
        V.seed_from_prng = orig
    d = p.diagnostics[key].as_dict()
    assert d["sum_trials"] == z["p_trials"].sum() and d["max_trials"] == z["p_trials"].max()
    _vec_close(_np(p[key, "particle_position"]), z["p_pos"], RTOL64, z["p_radius"])
    _vec_close(_np(p[key, "particle_velocity"]), z["p_vel"], RTOL64, z["p_speed"])
    np.testing.assert_allclose(_np(p[key, "particle_potential"]), z["p_pot"], rtol=RTOL64)
    np.testing.assert_allclose(_np(p[key, "particle_mass"]), z["p_mass"], rtol=1e-14)


def test_cuda_gas_path_reproduces_golden_vectors():
    import cluster_generator_b200.model as M

    z = np.load(os.path.join(GOLDEN, "vikhlinin_gas.npz"))
    pm = _golden_product_model("vikhlinin")
    seed, n = int(z["seed"]), z["p_mass"].size
    orig = M.seed_from_prng
    M.seed_from_prng = lambda prng, species: seed
    try:
        p = pm.generate_gas_particles(n, prng=0, compute_potential=True)
    finally:
        M.seed_from_prng = orig
    pos = z["p_pos"]
    _vec_close(_np(p["gas", "particle_position"]), pos, RTOL64, np.linalg.norm(pos, axis=1))
    np.testing.assert_allclose(_np(p["gas", "thermal_energy"]), z["p_therm"], rtol=1e-9)
    np.testing.assert_allclose(_np(p["gas", "density"]), z["p_dens"], rtol=1e-9)
    np.testing.assert_allclose(_np(p["gas", "particle_potential"]), z["p_pot"], rtol=1e-9)
    np.testing.assert_allclose(_np(p["gas", "particle_mass"]), z["p_mass"], rtol=1e-14)


def test_device_qags_matches_scipy_traces_decision_for_decision():
    """tests/golden/qags_traces.npz holds scipy's quad(..., full_output=1) state (last,
    alist, blist, rlist, elist) for a few energy bins of both integrands."""
    from scipy.interpolate import InterpolatedUnivariateSpline as IUS

    from cluster_generator_b200 import _native as N

    z = np.load(os.path.join(GOLDEN, "qags_traces.npz"))
    g = np.load(os.path.join(GOLDEN, "hernquist_dm.npz"))
    ee = g["ee"]
    t, c, _ = IUS(ee, g["density"][::-1])._eval_args
    tabs = {0: (N.dev_f64(t), N.dev_f64(c[: ee.size])), 1: (N.dev_f64(g["f_t"]), N.dev_f64(g["f_c"]))}
    out = torch.zeros(204, dtype=torch.float64, device="cuda")
    for key in z.files:
        kind, b = (0 if key[0] == "d" else 1), int(key[1:])
        td, cd = tabs[kind]
        eps = 1.49e-5 if kind == 0 else 1.49e-8
        upper = float(np.sqrt(ee[b])) if kind == 0 else float(ee[b])
        N.check(N.lib().cgb_qags_trace(N.ptr(td), N.ptr(cd), ee.size, float(ee[b]), 0.0, upper, eps, eps,
                                       kind, N.ptr(out), N.stream_ptr()))
        o, ref = out.cpu().numpy(), z[key]
        L = int(ref[2])
        assert int(o[2]) == L, key
        np.testing.assert_array_equal(o[4:4 + L], ref[3:3 + L], err_msg=key)                 # alist
        np.testing.assert_array_equal(o[54:54 + L], ref[3 + L:3 + 2 * L], err_msg=key)       # blist
        np.testing.assert_allclose(o[104:104 + L], ref[3 + 2 * L:3 + 3 * L], rtol=1e-13, atol=1e-300, err_msg=key)
        np.testing.assert_allclose(o[0], ref[0], rtol=1e-13, err_msg=key)


# --------------------------------------------------------------------------- #
# "next" row 1: multi-halo ICs (ics.py, particles.py:587-715)
# --------------------------------------------------------------------------- #
def _oracle_halo(oracle, om, n_dm, n_gas, seed, r_max):
    u = oracle.device_rng_uniforms(seed, 0, n_dm)
    dm = om.dm_virial.generate_particles(
        n_dm, r_max=r_max, uniforms=u,
        velocity_fn=lambda *a: oracle.generate_velocities_counter(*a, seed=seed))
    gas = oracle.generate_gas_particles(om, n_gas, r_max=r_max, uniforms=oracle.device_rng_uniforms(seed + 1, 0, n_gas))
    f = dict(dm)
    f.update(gas)
    return f


def test_two_cluster_ics_match_oracle(oracle):
    """BASELINE config 4 shape at test size: binary merger, d = 3000, b = 500, v = +-500
    (tests/test_ics.py:94-96), gas re-sampled from both profiles."""
    import cluster_generator_b200 as cg
    import cluster_generator_b200.model as M
    import cluster_generator_b200.virial as V

    om, pm = oracle_model("vikhlinin"), product_model("vikhlinin")
    c1, c2 = cg.compute_centers_for_binary(np.array([7000.0, 7000.0, 7000.0]), 3000.0, 500.0)
    v1, v2 = np.array([500.0, 0.0, 0.0]), np.array([-500.0, 0.0, 0.0])
    ics = cg.ClusterICs("t", 2, [pm, pm], [c1, c2], [v1, v2],
                        num_particles={"dm": 60_000, "gas": 60_000}, r_max=5000.0)
    assert ics.num_particles["dm"] == [30_000, 30_000] and ics.num_particles["gas"] == [30_000, 30_000]
    seeds = iter([101, 102, 201, 202])          # dm0, gas0, dm1, gas1 in the driver's call order
    orig_v, orig_m = V.seed_from_prng, M.seed_from_prng
    V.seed_from_prng = M.seed_from_prng = lambda prng, species: next(seeds)
    try:
        allp = ics.setup_particle_ics(prng=0)
    finally:
        V.seed_from_prng, M.seed_from_prng = orig_v, orig_m
    h0 = _oracle_halo(oracle, om, 30_000, 30_000, 101, 5000.0)
    h1 = _oracle_halo(oracle, om, 30_000, 30_000, 201, 5000.0)
    ref = oracle.combine_clusters([h0, h1], [om, om], [c1, c2], [v1, v2])
    assert allp.num_particles == {"dm": 60_000, "gas": 60_000}
    for key in (("dm", "particle_position"), ("gas", "particle_position")):
        np.testing.assert_allclose(_np(allp[key]), ref[key], rtol=1e-10, atol=1e-8)
    np.testing.assert_allclose(_np(allp["dm", "particle_velocity"]), ref["dm", "particle_velocity"], rtol=1e-10, atol=1e-10)
    for f in ("density", "thermal_energy"):
        np.testing.assert_allclose(_np(allp["gas", f]), ref["gas", f], rtol=1e-8, err_msg=f)
    np.testing.assert_allclose(_np(allp["gas", "particle_velocity"]), ref["gas", "particle_velocity"], rtol=1e-8, atol=1e-9)
    # the overlap really changes things: densities exceed the single-halo profile somewhere
    assert (ref["gas", "density"] > 0).all()
    vx = _np(allp["gas", "particle_velocity"])[:, 0]
    assert vx.min() < -400 and vx.max() > 400 and np.abs(vx).min() < 50


def test_resample_with_radii_matches_oracle(oracle):
    import cluster_generator_b200 as cg

    om, pm = oracle_model("vikhlinin"), product_model("vikhlinin")
    n = 40_000
    p = pm.generate_gas_particles(n, prng=5)
    fields = {k: _np(v).copy() for k, v in p.fields.items()}
    c1, c2 = np.zeros(3), np.array([1500.0, 200.0, 0.0])
    v1, v2 = np.array([1.0, 0.0, 0.0]), np.array([-2.0, 0.5, 0.0])
    radii = [800.0, 600.0]
    ref = oracle.sample_clusters(fields, [om, om], [c1, c2], [v1, v2], radii=radii, resample=True)
    out = cg.resample_two_clusters(p, pm, pm, c1, c2, v1, v2, radii)
    for f in ("density", "thermal_energy", "particle_mass"):
        np.testing.assert_allclose(_np(out["gas", f]), ref["gas", f], rtol=1e-8, err_msg=f)
    np.testing.assert_allclose(_np(out["gas", "particle_velocity"]), ref["gas", "particle_velocity"], rtol=1e-8, atol=1e-12)
    untouched = np.linalg.norm(fields["gas", "particle_position"] - c1, axis=1) > 800.0
    untouched &= np.linalg.norm(fields["gas", "particle_position"] - c2, axis=1) > 600.0
    assert untouched.any() and not _np(out["gas", "particle_velocity"])[untouched].any()


# --------------------------------------------------------------------------- #
# "next" row 3: the potential table on the device (model.py:219-224, utils.py:137-142)
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("name", ["hernquist", "vikhlinin"])
def test_device_potential_integral_matches_scipy(name):
    import time

    import cluster_generator_b200 as cg

    om = oracle_model(name)
    t0 = time.perf_counter()
    if name == "hernquist":
        pm = cg.ClusterModel.no_gas(0.1, 1.0e4, cg.hernquist_density_profile(1.0e15, 500.0), integrator="device")
    else:
        dens = cg.vikhlinin_density_profile(4.2e5, 100.0, 1000.0, 1.0, 0.67, 3.0)
        temp = cg.vikhlinin_temperature_profile(10.0, 0.0, 2.0, 1.0, 1500.0, 6.0, 80.0, 1.0)
        pm = cg.ClusterModel.from_dens_and_temp(0.1, 1.0e4, dens, temp, integrator="device")
    print(name, "model build with device integrator:", time.perf_counter() - t0, "s")
    phi, phi_o = np.asarray(pm["gravitational_potential"]), om["gravitational_potential"]
    np.testing.assert_allclose(phi, phi_o, rtol=1e-12)
    assert (phi == phi_o).mean() > 0.9            # same QUADPACK decisions: mostly bit-equal
    with pytest.raises(ValueError):
        cg.ClusterModel.no_gas(0.1, 1.0e4, cg.hernquist_density_profile(1.0e15, 500.0), integrator="gpu")


# --------------------------------------------------------------------------- #
# SURVEY.md section 8(d) config-2 hazard (i): strong cool core => f(E) < 0 in a band =>
# non-positive envelopes whose accept test degenerates.  Reference behaviour, reproduced.
# --------------------------------------------------------------------------- #
def test_negative_distribution_function_band_is_reproduced(oracle):
    import cluster_generator_b200 as cg
    import cluster_generator_b200.virial as V

    dens_o = oracle.vikhlinin_density_profile(4.2e5, 100.0, 1000.0, 1.0, 0.67, 3.0)
    temp_o = oracle.vikhlinin_temperature_profile(10.0, 0.0, 2.0, 1.0, 1500.0, 4.0, 80.0, 1.5)
    om = oracle.from_dens_and_temp(0.1, 1.0e4, dens_o, temp_o)
    dens = cg.vikhlinin_density_profile(4.2e5, 100.0, 1000.0, 1.0, 0.67, 3.0)
    temp = cg.vikhlinin_temperature_profile(10.0, 0.0, 2.0, 1.0, 1500.0, 4.0, 80.0, 1.5)
    pm = cg.ClusterModel.from_dens_and_temp(0.1, 1.0e4, dens, temp)
    ov, pv = om.dm_virial, pm.dm_virial
    assert (ov.ff < 0).sum() > 10                                  # the band exists
    np.testing.assert_allclose(pv.ff, ov.ff, rtol=1e-8, atol=1e-8 * np.abs(ov.ff).max() * 1e-6)
    n, seed = 200_000, 31
    f = ov.generate_particles(n, uniforms=oracle.device_rng_uniforms(seed, 0, n),
                              velocity_fn=lambda *a: oracle.generate_velocities_counter(*a, seed=seed))
    nonpos = int((ov.last["fv2esc"] <= 0).sum())
    assert nonpos > 0
    orig = V.seed_from_prng
    V.seed_from_prng = lambda prng, species: seed
    try:
        p = pv.generate_particles(n, prng=0)
    finally:
        V.seed_from_prng = orig
    d = p.diagnostics["dm"].as_dict()
    assert d["nonpositive_envelopes"] == nonpos                     # counted and reported
    assert d["sum_trials"] == ov.last_trials.sum() and d["max_trials"] == ov.last_trials.max()
    _vec_close(_np(p["dm", "particle_velocity"]), f["dm", "particle_velocity"], RTOL64, ov.last["speed"])


def test_full_size_properties_config3_headline():
    """BASELINE config 3 (headline): NFW DM + Hernquist BCG stars, 1e8 particles split by
    enclosed mass like ClusterICs (ics.py:132-153), here on one GPU.  Size-independent
    properties only (the oracle would need ~10 CPU-minutes for this)."""
    import cluster_generator_b200 as cg

    pm = product_model("nfw_bcg")
    ics = cg.ClusterICs("c3", 1, [pm], [0.0, 0.0, 0.0], [0.0, 0.0, 0.0],
                        num_particles={"dm": 99_000_000, "star": 1_000_000}, r_max=1.0e4)
    assert ics.num_particles["dm"] == [99_000_000] and ics.num_particles["star"] == [1_000_000]
    p = ics.setup_particle_ics(prng=11, compute_potential=False)
    assert p.num_particles == {"dm": 99_000_000, "star": 1_000_000}
    for pt, lo, hi in (("dm", 15.0, 40.0), ("star", 150.0, 600.0)):      # SURVEY section 6 T-bar: 23 / 293
        d = p.diagnostics[pt].as_dict()
        assert d["failures"] == 0 and lo < d["mean_trials"] < hi, (pt, d["mean_trials"])
        pos, vel = p[pt, "particle_position"], p[pt, "particle_velocity"]
        assert torch.isfinite(pos).all() and torch.isfinite(vel).all()
        r = torch.linalg.norm(pos, dim=1)
        assert float(r.max()) < 1.0e4 * (1 + 1e-12)
        # isotropy of both vectors: |<cos theta>| within 5 sigma
        n = pos.shape[0]
        assert abs(float((pos[:, 2] / r).mean())) < 5.0 / (3.0 * n) ** 0.5
        v = torch.linalg.norm(vel, dim=1)
        assert abs(float((vel[:, 2] / v).mean())) < 5.0 / (3.0 * n) ** 0.5
        # moments reported by the kernel agree with the data it wrote
        np.testing.assert_allclose(d["sum_v2"], float((v.double() ** 2).sum()), rtol=1e-9)
        np.testing.assert_allclose(d["sum_r"], float(r.double().sum()), rtol=1e-9)
    # stars are more concentrated than the dark matter
    assert float(torch.linalg.norm(p["star", "particle_position"], dim=1).median()) < \
        0.2 * float(torch.linalg.norm(p["dm", "particle_position"], dim=1).median())


# --------------------------------------------------------------------------- #
# "next" row 4: 3-D field -> particles (fields.py:308-338)
# --------------------------------------------------------------------------- #
def test_map_field_to_particles_matches_regular_grid_interpolator():
    from scipy.interpolate import RegularGridInterpolator

    import cluster_generator_b200 as cg

    rng = np.random.RandomState(4)
    x = np.linspace(-3000.0, 3000.0, 33)
    y = np.linspace(-2500.0, 2500.0, 29)
    z = np.sort(rng.uniform(-2000.0, 2000.0, 31))          # non-uniform axis too
    comps = [rng.normal(size=(33, 29, 31)) for _ in range(3)]
    pm = product_model("vikhlinin")
    p = pm.generate_gas_particles(50_000, prng=2)
    fld = cg.GridField("magnetic_field", x, y, z, comps, units="gauss")
    fld.map_field_to_particles(p, ptype="gas")
    pos = _np(p["gas", "particle_position"])
    got = _np(p["gas", "magnetic_field"])
    assert got.shape == (50_000, 3) and p.field_units["gas", "magnetic_field"] == "gauss"
    for i in range(3):
        ref = RegularGridInterpolator((x, y, z), comps[i], bounds_error=False, fill_value=0.0)(pos)
        np.testing.assert_allclose(got[:, i], ref, rtol=1e-12, atol=1e-14)
    outside = (np.abs(pos[:, 0]) > 3000.0) | (np.abs(pos[:, 1]) > 2500.0) | (pos[:, 2] < z[0]) | (pos[:, 2] > z[-1])
    assert outside.any() and not got[outside].any() and np.abs(got[~outside]).max() > 0
    with pytest.raises(ValueError):
        cg.GridField("b", x[::-1], y, z, comps)


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 127, 128, 129, 255, 257, 4095, 4097, 65_537])
def test_tile_boundaries_parity_sweep(oracle, n):
    """Particle counts around the warp / tile / CTA boundaries of the sampler (32 lanes,
    128-particle tiles, 16 warps): every particle must still equal the oracle's."""
    seed = 1000 + n
    ov, fo = _oracle_particles(oracle, "nfw_bcg", "dark_matter", n, seed)
    pv, p = _device_particles("nfw_bcg", "dark_matter", n, seed)
    _vec_close(_np(p["dm", "particle_position"]), fo["dm", "particle_position"], RTOL64, ov.last["radius"])
    _vec_close(_np(p["dm", "particle_velocity"]), fo["dm", "particle_velocity"], RTOL64, ov.last["speed"])
    d = p.diagnostics["dm"].as_dict()
    assert d["particles"] == n and d["sum_trials"] == ov.last_trials.sum() and d["failures"] == 0
    # a shard that starts in the middle of a tile reproduces its slice
    if n > 40:
        _, q = _device_particles("nfw_bcg", "dark_matter", n, seed, shard=(1, 3))
        lo, hi = n // 3, (2 * n) // 3
        assert torch.equal(q["dm", "particle_velocity"], p["dm", "particle_velocity"][lo:hi])


def test_integration_md_cython_seam_stub_runs_as_documented(oracle):
    """The ctypes stub INTEGRATION.md tells a maintainer to add is executed verbatim."""
    import re

    from cluster_generator_b200 import _native as N

    md = open(os.path.join(os.path.dirname(__file__), "..", "INTEGRATION.md")).read()
    code = re.search(r"```python\n(import ctypes, numpy as np, torch.*?)```", md, re.S).group(1)
    ns = {}
    exec(code.replace('"libcgb200.so"', repr(N.LIB_PATH)), ns)
    n = 20_000
    ov, last, trials_o = _seam_inputs(oracle, "hernquist", "dark_matter", n, seed=0)
    vel = ns["generate_velocities"](last["psi"], last["vesc"], last["fv2esc"], ov.f.get_knots(),
                                    ov.f.get_coeffs(), 3, -2)
    np.testing.assert_allclose(vel, last["speed"], rtol=RTOL64)
