"""Round-2 GPU tests: screen audit, full-size parity against the C oracle, statistical
acceptance beyond Hernquist, the re-worked gas kernel and spline solve."""
import os

import numpy as np
import pytest
import torch

from tests.configs import oracle_model, product_model

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


def _seeded(seed):
    """Context manager: make the product's device RNG use exactly ``seed``."""
    import contextlib

    import cluster_generator_b200.model as M
    import cluster_generator_b200.virial as V

    @contextlib.contextmanager
    def cm():
        ov, om = V.seed_from_prng, M.seed_from_prng
        V.seed_from_prng = M.seed_from_prng = lambda prng, species: seed
        try:
            yield
        finally:
            V.seed_from_prng, M.seed_from_prng = ov, om
    return cm()


def _negband_model():
    import cluster_generator_b200 as cg

    dens = cg.vikhlinin_density_profile(4.2e5, 100.0, 1000.0, 1.0, 0.67, 3.0)
    temp = cg.vikhlinin_temperature_profile(10.0, 0.0, 2.0, 1.0, 1500.0, 4.0, 80.0, 1.5)
    return cg.ClusterModel.from_dens_and_temp(0.1, 1.0e4, dens, temp)


AUDIT = [("hernquist", "dm", 3_000_000), ("nfw_bcg", "dm", 3_000_000), ("nfw_bcg", "star", 300_000),
         ("vikhlinin", "dm", 4_000_000), ("negband", "dm", 2_000_000)]


@pytest.mark.parametrize("name,key,n", AUDIT)
def test_screen_never_contradicts_the_exact_accept_test(name, key, n):
    """the fp32 interval screen and the exact fp64 test, side by side on
    EVERY trial (cgb_screen_audit): zero contradictions, and the audited trial sequence is
    the production kernel's (same sum of accepted trial numbers)."""
    pm = _negband_model() if name == "negband" else product_model(name)
    v = pm.dm_virial if key == "dm" else pm.star_virial
    seed = 4242
    a = v.screen_audit(n, seed=seed, max_trials=1 << 22)
    assert a["contradictions"] == 0, a
    assert a["capped_particles"] == 0, a
    assert a["trials"] > 10 * n and a["trials"] == a["undecided"] + a["screen_accepts"] + a["screen_rejects"]
    assert a["undecided"] < 2e-3 * a["trials"], a              # the screen decides > 99.8 % of trials
    # (negative envelopes keep the screen: the fp32 comparison is sign-correct for them too)
    with _seeded(seed):
        p = v.generate_particles(n, prng=0)
    d = p.diagnostics[key].as_dict()
    assert d["sum_trials"] == a["sum_accepted_trials"]
    print(name, key, a)


def test_gas_kernel_odd_counts_unaligned_shards_and_null_velocity(oracle):
    """The pair-per-thread gas kernel: odd particle counts (scalar tail, odd plane stride),
    shards starting at odd ids, fp32, and d_vel == NULL."""
    from cluster_generator_b200 import _native as N

    om, pm = oracle_model("vikhlinin"), product_model("vikhlinin")
    n, seed = 100_001, 77
    og = oracle.generate_gas_particles(om, n, compute_potential=True,
                                       uniforms=oracle.device_rng_uniforms(seed, 0, n))
    with _seeded(seed):
        p = pm.generate_gas_particles(n, prng=0, compute_potential=True)
        shards = [pm.generate_gas_particles(n, prng=0, compute_potential=True, shard=(r, 3)) for r in range(3)]
        p32 = pm.generate_gas_particles(n, prng=0, dtype=torch.float32)
    scale = np.linalg.norm(og["gas", "particle_position"], axis=1)[:, None]
    assert np.all(np.abs(_np(p["gas", "particle_position"]) - og["gas", "particle_position"])
                  <= 1e-10 * np.abs(og["gas", "particle_position"]) + 1e-12 * scale)
    for f in ("thermal_energy", "density", "particle_mass", "particle_potential"):
        np.testing.assert_allclose(_np(p["gas", f]), og["gas", f], rtol=1e-10, err_msg=f)
    assert not p["gas", "particle_velocity"].any()
    for f in p.fields:
        cat = torch.cat([s[f] for s in shards])
        assert torch.equal(cat, p[f]), f                       # odd shard starts: bit-identical
    np.testing.assert_allclose(_np(p32["gas", "particle_position"]), og["gas", "particle_position"],
                               rtol=1e-5, atol=1e-7 * scale.max())
    # C ABI: d_vel == NULL leaves the caller's (zeroed) storage alone
    tabs = pm._gas_sampler_tables(None, False)
    dev = p["gas", "particle_mass"].device
    pos = torch.empty((3, 4096), dtype=torch.float64, device=dev)
    a = N.SampleArgs()
    a.d_workspace, a.id_begin, a.id_end, a.ld = tabs.workspace.data_ptr(), 0, 4096, 4096
    a.particle_mass, a.seed, a.out_dtype = 1.0, seed, N.CGB_DTYPE_F64
    a.d_pos = pos.data_ptr()
    N.check(N.lib().cgb_sample_gas(N.ctypes.byref(a), None, None, N.stream_ptr()))
    assert torch.equal(pos.T, p["gas", "particle_position"][:4096])
    # the packed shared-memory path and the general kernel give identical bits
    assert tabs.gas_flags & 1
    pos2 = torch.empty_like(pos)
    for flags, out in ((tabs.gas_flags, pos), (0, pos2)):
        a.flags, a.d_pos = flags, out.data_ptr()
        N.check(N.lib().cgb_sample_gas(N.ctypes.byref(a), None, None, N.stream_ptr()))
    assert torch.equal(pos, pos2)


def test_set_field_invalidates_cached_device_tables():
    """generate_gas_particles after set_field must sample the new field."""
    import cluster_generator_b200 as cg

    dens = cg.vikhlinin_density_profile(4.2e5, 100.0, 1000.0, 1.0, 0.67, 3.0)
    temp = cg.vikhlinin_temperature_profile(10.0, 0.0, 2.0, 1.0, 1500.0, 6.0, 80.0, 1.0)
    pm = cg.ClusterModel.from_dens_and_temp(0.1, 1.0e4, dens, temp, num_points=200)
    a = pm.generate_gas_particles(1000, prng=1)
    pm.set_field("density", 2.0 * pm["density"].d)
    b = pm.generate_gas_particles(1000, prng=1)
    np.testing.assert_allclose(_np(b["gas", "density"]), 2.0 * _np(a["gas", "density"]), rtol=1e-12)


def test_back_substitution_quotient_is_the_ieee_quotient():
    """fpback divides by a stored correctly-rounded reciprocal + two FMAs (Markstein's final
    step) instead of a full division; it must be the IEEE quotient bit for bit."""
    from cluster_generator_b200 import _native as N

    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(5)
    n = 1 << 24
    for spread in (1.0, 40.0, 250.0):
        x = (torch.rand(n, generator=g, device=dev, dtype=torch.float64) + 0.5) * torch.exp2(
            torch.randint(-int(spread), int(spread) + 1, (n,), generator=g, device=dev).double())
        x = x * (torch.randint(0, 2, (n,), generator=g, device=dev).double() * 2 - 1)
        b = (torch.rand(n, generator=g, device=dev, dtype=torch.float64) + 0.5) * torch.exp2(
            torch.randint(-int(spread), int(spread) + 1, (n,), generator=g, device=dev).double())
        q = torch.empty_like(x)
        N.check(N.lib().cgb_test_div_by_rcp(N.ptr(x), N.ptr(b), N.ptr(q), n, N.stream_ptr()))
        assert torch.equal(q, x / b), spread
    # special operands
    x = torch.tensor([0.0, -0.0, 1e-310, 1e300, 3.0, float("inf")], dtype=torch.float64, device=dev)
    b = torch.tensor([3.0, 3.0, 7.0, 1e-10, 3.0, 2.0], dtype=torch.float64, device=dev)
    q = torch.empty_like(x)
    N.check(N.lib().cgb_test_div_by_rcp(N.ptr(x), N.ptr(b), N.ptr(q), x.numel(), N.stream_ptr()))
    assert torch.equal(q, x / b)


def test_branch_free_fpgivs_equals_the_plain_operators_and_the_fit_survives_a_redo():
    """The factorising fit computes FITPACK's rotations with the compiler's own division / sqrt /
    reciprocal sequences minus their special-operand branches.  (dd, cos, sin) must equal the
    plain-operator results bit for bit on every operand pair its range test lets through, the
    test must flag what it cannot vouch for, and a fit forced back to the plain operators must
    give the same coefficients."""
    import ctypes

    from scipy.interpolate import InterpolatedUnivariateSpline as IUS

    from cluster_generator_b200 import _native as N
    from cluster_generator_b200.virial import DeviceSpline

    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(9)
    n = 1 << 22

    def operands(lo, hi):
        def one():
            m = torch.rand(n, generator=g, device=dev, dtype=torch.float64) + 1.0
            return m * torch.exp2(torch.randint(lo, hi + 1, (n,), generator=g, device=dev).double())
        sign = torch.randint(0, 2, (n,), generator=g, device=dev).double() * 2 - 1
        return one() * sign, one()

    def run(piv, ww):
        m = piv.numel()
        fast = torch.empty(3 * m, dtype=torch.float64, device=dev)
        plain = torch.empty_like(fast)
        flag = torch.empty(m, dtype=torch.int32, device=dev)
        N.check(N.lib().cgb_test_fpgivs(N.ptr(piv), N.ptr(ww), N.ptr(fast), N.ptr(plain), N.ptr(flag), m,
                                        N.stream_ptr()))
        return fast.view(torch.int64), plain.view(torch.int64), flag

    # the whole admitted range, near-equal magnitudes, and the extreme ratios inside it
    for lo, hi in ((-299, 298), (-3, 3), (0, 0)):
        fast, plain, flag = run(*operands(lo, hi))
        assert int(flag.sum()) == 0
        assert torch.equal(fast, plain), (lo, hi)
    piv, ww = operands(-299, -290)
    fast, plain, flag = run(piv, ww * 2.0 ** 588)
    assert int(flag.sum()) == 0 and torch.equal(fast, plain)
    fast, plain, flag = run(piv * 2.0 ** 588, ww)
    assert int(flag.sum()) == 0 and torch.equal(fast, plain)
    # outside: flagged (whatever the fast form returns there is discarded by the fit)
    bad_p = torch.tensor([1e-300, 1.0, 1e300, float("inf"), float("nan"), 0.0, 1e-310, 1.0], dtype=torch.float64, device=dev)
    bad_w = torch.tensor([1.0, 1e-300, 1.0, 1.0, 1.0, 1.0, 1.0, 1e95], dtype=torch.float64, device=dev)
    _, _, flag = run(bad_p, bad_w)
    assert flag.tolist() == [1] * 8

    # a fit sent back to the plain operators
    om = oracle_model("vikhlinin")
    x, y = -om["gravitational_potential"][::-1], om["dark_matter_density"][::-1]
    was = ctypes.c_int(0)
    N.check(N.lib().cgb_test_fpgivs_redo(0, ctypes.byref(was)))
    c0 = DeviceSpline.fit(x, y, method="fitpack").get_coeffs()
    N.check(N.lib().cgb_test_fpgivs_redo(1, ctypes.byref(was)))
    assert was.value == 0                       # the normal fit never needed it
    try:
        c1 = DeviceSpline.fit(x, y, method="fitpack").get_coeffs()
    finally:
        N.check(N.lib().cgb_test_fpgivs_redo(0, ctypes.byref(was)))
    assert was.value == 1                       # the forced one did
    np.testing.assert_array_equal(c0, c1)
    np.testing.assert_array_equal(c0, IUS(x, y).get_coeffs())


def test_fitpack_exact_fit_small_and_odd_sizes_bit_equal():
    """The register-window factorisation at full, odd and minimal sizes equals scipy's
    FITPACK coefficients bit for bit (the factor re-use path is covered by the f(E) tests)."""
    from scipy.interpolate import InterpolatedUnivariateSpline as IUS

    from cluster_generator_b200.virial import DeviceSpline

    for name in ("hernquist", "vikhlinin"):
        om = oracle_model(name)
        x = -om["gravitational_potential"][::-1]
        for n in (x.size, 37, 8):
            xs = x[:n]
            for y in (om["dark_matter_density"][::-1][:n], np.sin(np.arange(n)) * 1e-7, np.ones(n)):
                c = DeviceSpline.fit(xs, y, method="fitpack").get_coeffs()
                np.testing.assert_array_equal(c, IUS(xs, y).get_coeffs())


def test_host_streaming_is_bit_identical_to_the_device_path():
    """generate_*_particles(host=True): chunked generation with overlapped device->host copies
    returns pinned host tensors equal to the one-shot device result bit for bit, for chunk
    sizes that do and do not divide n, odd shards, fp32, and with buffer re-use."""
    pm = product_model("vikhlinin")
    n = 300_001
    for kw in (dict(), dict(compute_potential=True), dict(shard=(1, 3)), dict(dtype=torch.float32)):
        with _seeded(11):
            d = pm.generate_dm_particles(n, prng=0, **kw)
            h = pm.generate_dm_particles(n, prng=0, host=True, chunk=65_536, **kw)
            g = pm.generate_gas_particles(n, prng=0, **kw)
            gh = pm.generate_gas_particles(n, prng=0, host=True, chunk=100_000, **kw)
        for dev_p, host_p in ((d, h), (g, gh)):
            assert list(dev_p.fields) == list(host_p.fields)
            for k in dev_p.fields:
                t = host_p[k]
                assert not t.is_cuda and t.is_pinned() and t.shape == dev_p[k].shape
                assert torch.equal(t, dev_p[k].cpu()), k
        assert h["dm", "particle_position"].stride() == d["dm", "particle_position"].stride()
        assert h.diagnostics["dm"].as_dict()["sum_trials"] == d.diagnostics["dm"].as_dict()["sum_trials"]
    # buffer re-use: the same pinned storage is overwritten
    with _seeded(12):
        h2 = pm.generate_dm_particles(n, prng=0, host=h, chunk=65_536, dtype=torch.float32)
        d2 = pm.generate_dm_particles(n, prng=0, dtype=torch.float32)
    assert h2["dm", "particle_mass"].data_ptr() == h["dm", "particle_mass"].data_ptr()
    assert torch.equal(h2["dm", "particle_velocity"], d2["dm", "particle_velocity"].cpu())
    # wait=False: the call returns with copies in flight, a second species is generated meanwhile,
    # and the first access to the fields completes them
    with _seeded(14):
        ha = pm.generate_dm_particles(n, prng=0, host=True, chunk=65_536, wait=False)
        ga = pm.generate_gas_particles(n, prng=0, host=True, chunk=65_536, wait=False)
        dd = pm.generate_dm_particles(n, prng=0)
        gd = pm.generate_gas_particles(n, prng=0)
    assert len(ha._pending) == 1 and len(ga._pending) == 1
    assert ha.wait() is ha and not ha._pending
    for k in dd.fields:
        assert torch.equal(ha[k], dd[k].cpu()), k
    for k in gd.fields:   # no explicit wait: the field access does it
        assert torch.equal(ga[k], gd[k].cpu()), k
    assert not ga._pending
    # sub_sample > 1 and rng="numpy" take the copy-after path
    with _seeded(13):
        a = pm.generate_dm_particles(40_000, prng=3, sub_sample=4, host=True)
        b = pm.generate_dm_particles(40_000, prng=3, sub_sample=4)
    assert torch.equal(a["dm", "particle_velocity"], b["dm", "particle_velocity"].cpu())


# --------------------------------------------------------------------------- #
# parity at BASELINE sizes: the C oracle forked over the host cores
# --------------------------------------------------------------------------- #
_SHARD_CTX = {}


def _oracle_shard(args):
    """Particles [a, b) of an n-particle run, by the oracle in counter mode; returns the trial
    statistics of the shard and the coordinates of every `step`-th particle."""
    a, b, step = args
    from oracle import cg_oracle as O

    ov, seed, n = _SHARD_CTX["virial"], _SHARD_CTX["seed"], _SHARD_CTX["n"]
    u = O.device_rng_uniforms(seed, a, b - a)
    f = ov.generate_particles(b - a, uniforms=u,
                              velocity_fn=lambda *x: O.generate_velocities_counter(*x, seed=seed, first_id=a))
    key = "dm" if ov.ptype == "dark_matter" else "star"
    tr = ov.last_trials
    pick = np.arange((-a) % step, b - a, step)
    return (int(tr.sum()), int(tr.max()), int((tr <= 0).sum()), a + pick,
            f[key, "particle_position"][pick], f[key, "particle_velocity"][pick],
            ov.last["radius"][pick], ov.last["speed"][pick])


def _oracle_full_size(ov, seed, n, step):
    import multiprocessing as mp

    _SHARD_CTX.update(virial=ov, seed=seed, n=n)
    cores = min(os.cpu_count() or 1, 32)
    chunk = max(50_000, -(-n // (4 * cores)))
    jobs = [(a, min(a + chunk, n), step) for a in range(0, n, chunk)]
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_oracle_shard, jobs)
    return (sum(r[0] for r in res), max(r[1] for r in res), sum(r[2] for r in res),
            np.concatenate([r[3] for r in res]), np.concatenate([r[4] for r in res]),
            np.concatenate([r[5] for r in res]), np.concatenate([r[6] for r in res]),
            np.concatenate([r[7] for r in res]))


@pytest.mark.parametrize("name,key,n", [("vikhlinin", "dm", 10_000_000), ("nfw_bcg", "dm", 10_000_000),
                                        ("nfw_bcg", "star", 100_000), ("hernquist", "dm", 4_000_000)])
def test_parity_at_baseline_sizes_against_the_c_oracle(oracle, name, key, n):
    """configs[1] at its stated 1e7 DM particles, configs[2] at 1e7 DM + 1e5 stars, configs[0]'s
    model at 4e6: the sum and the maximum of the accepted trial numbers equal the oracle's
    EXACTLY (every accept / reject decision, incl. the long-tail paths -- lane cap, cooperative
    finish, deferred grid-wide pass -- that small runs never reach), and every 997th particle's
    coordinates agree to rtol 1e-10."""
    om, pm = oracle_model(name), product_model(name)
    ov = om.dm_virial if key == "dm" else om.star_virial
    pv = pm.dm_virial if key == "dm" else pm.star_virial
    seed = 20261018
    tsum, tmax, nfail, ids, opos, ovel, orad, ospd = _oracle_full_size(ov, seed, n, 997)
    with _seeded(seed):
        p = pv.generate_particles(n, prng=0)
    d = p.diagnostics[key].as_dict()
    assert d["particles"] == n and d["failures"] == 0 and nfail == 0
    assert d["sum_trials"] == tsum, (d["sum_trials"], tsum)
    assert d["max_trials"] == tmax, (d["max_trials"], tmax)
    idx = torch.as_tensor(ids, device=p[key, "particle_mass"].device)
    pos = _np(p[key, "particle_position"][idx])
    vel = _np(p[key, "particle_velocity"][idx])
    # the oracle scales mass by its own (shard) particle count: positions and velocities only
    assert np.all(np.abs(pos - opos) <= 1e-10 * np.abs(opos) + 1e-12 * orad[:, None])
    assert np.all(np.abs(vel - ovel) <= 1e-10 * np.abs(ovel) + 1e-12 * ospd[:, None])
    print(name, key, n, "sum trials", tsum, "max", tmax, "deferred/coop", d.get("cooperative_finishes"))


# --------------------------------------------------------------------------- #
# statistical acceptance with the device RNG beyond Hernquist (north_star: KS on r, |v|, E)
# --------------------------------------------------------------------------- #
def _ks_2samp_stat(a, b):
    from scipy.stats import ks_2samp

    return ks_2samp(a, b)


@pytest.mark.parametrize("name,key", [("nfw_bcg", "dm"), ("nfw_bcg", "star"), ("vikhlinin", "dm"), ("hernquist", "dm")])
def test_device_rng_is_statistically_indistinguishable_from_the_oracle(oracle, name, key):
    """Two-sample KS (alpha = 0.01) of r, |v| and E = psi(r) - v^2/2 between the device-RNG
    particles and the oracle run on an independent numpy stream with its private sequential
    speed RNG (the reference's own mode), n = 1e6 each; isotropy of both direction sets."""
    from scipy.interpolate import InterpolatedUnivariateSpline as IUS

    om, pm = oracle_model(name), product_model(name)
    ov = om.dm_virial if key == "dm" else om.star_virial
    pv = pm.dm_virial if key == "dm" else pm.star_virial
    n = 1_000_000
    fo = ov.generate_particles(n, prng=123)
    p = pv.generate_particles(n, prng=456)
    psi_of = IUS(om["radius"], -om["gravitational_potential"])

    def rve(pos, vel):
        r = np.linalg.norm(pos, axis=1)
        v = np.linalg.norm(vel, axis=1)
        return r, v, psi_of(r) - 0.5 * v * v

    a = rve(_np(p[key, "particle_position"]), _np(p[key, "particle_velocity"]))
    b = rve(fo[key, "particle_position"], fo[key, "particle_velocity"])
    for what, x, y in zip(("r", "|v|", "E"), a, b):
        ks = _ks_2samp_stat(x, y)
        assert ks.pvalue > 0.01, (name, key, what, ks)
        print(name, key, what, "KS D = %.2e p = %.3f" % (ks.statistic, ks.pvalue))
    # bound orbits only, like the reference sampler: E > 0
    assert (a[2] > 0).all()
    from scipy.stats import kstest

    for fld in ("particle_position", "particle_velocity"):
        x = _np(p[key, fld])
        nrm = np.linalg.norm(x, axis=1)
        cos_t = x[:, 2] / nrm
        phi = np.mod(np.arctan2(x[:, 1], x[:, 0]), 2 * np.pi)
        assert kstest(cos_t, "uniform", args=(-1, 2)).pvalue > 0.01
        assert kstest(phi, "uniform", args=(0, 2 * np.pi)).pvalue > 0.01


@pytest.mark.parametrize("name,ptype", [("nfw_bcg", "stellar"), ("vikhlinin", "dark_matter")])
def test_check_virial_matches_reference_arithmetic_other_species(oracle, name, ptype):
    """check_star_virial (model.py:705-716, tests/test_models.py:132-148) and the Vikhlinin
    check_dm_virial: densities bit-equal to the scipy-calling oracle, and the reference's signed
    acceptance criterion."""
    om, pm = oracle_model(name), product_model(name)
    ov = om.star_virial if ptype == "stellar" else om.dm_virial
    rho_o, chk_o = ov.check_virial()
    rho, chk = pm.check_star_virial() if ptype == "stellar" else pm.check_dm_virial()
    np.testing.assert_array_equal(np.asarray(rho), rho_o)
    np.testing.assert_array_equal(np.asarray(chk), chk_o)
    if ptype == "stellar":
        assert np.all(np.asarray(chk) < 1.0e-4)              # tests/test_models.py:144-147 (signed)
    print(name, ptype, "max |chk|", np.abs(chk_o).max(), "max chk", chk_o.max())


# --------------------------------------------------------------------------- #
# 8(f)1 leftovers: three-halo combination, passive scalars
# --------------------------------------------------------------------------- #
def _oracle_halo(oracle, om, nd, ns, ng, seed, r_max):
    """One halo's particles with the device RNG streams of the product (seed, seed + 1, seed + 2
    for dm / star / gas in the driver's call order)."""
    out = {}
    if nd:
        out.update(om.dm_virial.generate_particles(
            nd, r_max=r_max, uniforms=oracle.device_rng_uniforms(seed, 0, nd),
            velocity_fn=lambda *a: oracle.generate_velocities_counter(*a, seed=seed)))
    if ns:
        out.update(om.star_virial.generate_particles(
            ns, r_max=r_max, uniforms=oracle.device_rng_uniforms(seed + 1, 0, ns),
            velocity_fn=lambda *a: oracle.generate_velocities_counter(*a, seed=seed + 1)))
    if ng:
        out.update(oracle.generate_gas_particles(om, ng, r_max=r_max,
                                                 uniforms=oracle.device_rng_uniforms(seed + 2, 0, ng)))
    return out


def test_three_cluster_ics_match_oracle(oracle):
    """combine_three_clusters (particles.py:669-715, ics.py:349-374): two gas halos and a
    gas-free NFW + BCG halo (the "no density field" skip), against the oracle's restatement --
    which tests/test_reference_cpu.py pins to the reference's own code for this geometry."""
    import cluster_generator_b200 as cg
    import cluster_generator_b200.model as M
    import cluster_generator_b200.virial as V

    om_a, pm_a = oracle_model("vikhlinin"), product_model("vikhlinin")
    om_b, pm_b = oracle_model("nfw_bcg"), product_model("nfw_bcg")
    c1, c2 = cg.compute_centers_for_binary(np.array([7000.0] * 3), 3000.0, 500.0)
    cen = [c1, c2, np.array([9000.0, 4000.0, 7000.0])]
    vel = [np.array([500.0, 0.0, 0.0]), np.array([-500.0, 0.0, 0.0]), np.array([0.0, 300.0, -100.0])]
    ics = cg.ClusterICs("t3", 3, [pm_a, pm_b, pm_a], cen, vel,
                        num_particles={"dm": 90_000, "gas": 40_000, "star": 5_000}, r_max=5000.0)
    nd, ng, ns = ics.num_particles["dm"], ics.num_particles["gas"], ics.num_particles["star"]
    assert ng[1] == 0 and ns[0] == ns[2] == 0 and ns[1] == 5_000 and abs(sum(nd) - 90_000) <= 1
    seeds = {0: 100, 1: 200, 2: 300}
    calls = []
    for i in range(3):  # the driver's call order per halo: dm, star, gas (ics.py:173-192)
        calls.append(seeds[i])
        if ns[i]:
            calls.append(seeds[i] + 1)
        if ng[i]:
            calls.append(seeds[i] + 2)
    it = iter(calls)
    orig_v, orig_m = V.seed_from_prng, M.seed_from_prng
    V.seed_from_prng = M.seed_from_prng = lambda prng, species: next(it)
    try:
        allp = ics.setup_particle_ics(prng=0)
    finally:
        V.seed_from_prng, M.seed_from_prng = orig_v, orig_m
    oms = [om_a, om_b, om_a]
    parts = [_oracle_halo(oracle, oms[i], nd[i], ns[i], ng[i], seeds[i], 5000.0) for i in range(3)]
    ref = oracle.combine_clusters(parts, oms, cen, vel)
    assert {k: int(v) for k, v in allp.num_particles.items()} == {"dm": sum(nd), "star": 5_000, "gas": sum(ng)}
    for key in ref:
        a, b = _np(allp[key]), ref[key]
        if key[1] in ("particle_position", "particle_velocity"):
            np.testing.assert_allclose(a, b, rtol=1e-8, atol=1e-8, err_msg=str(key))
        else:
            np.testing.assert_allclose(a, b, rtol=1e-8, err_msg=str(key))
    vx = _np(allp["gas", "particle_velocity"])
    assert vx[:, 0].max() > 400 and vx[:, 1].max() > 200        # both gas halos (0 and 2) contribute


def test_resample_with_passive_scalars_matches_oracle(oracle):
    """Passive scalars in _sample_clusters (particles.py:606-637): mass-weighted scalar profiles
    of two and three halos inside the resample radii, untouched outside."""
    import cluster_generator_b200 as cg

    om, pm = oracle_model("vikhlinin"), product_model("vikhlinin")
    z = 0.3 * (1.0 + om["radius"] / 500.0) ** -0.5
    y = 1.0e-3 * np.exp(-om["radius"] / 2000.0)
    om.fields["metallicity"], om.fields["dye"] = z.copy(), y.copy()
    pm.set_field("metallicity", z.copy())
    pm.set_field("dye", y.copy())
    try:
        n = 40_000
        c = [np.zeros(3), np.array([1500.0, 200.0, 0.0]), np.array([-900.0, 0.0, 400.0])]
        v = [np.array([1.0, 0.0, 0.0]), np.array([-2.0, 0.5, 0.0]), np.array([0.0, 0.0, 3.0])]
        radii = [800.0, 600.0, 500.0]
        for nh in (2, 3):
            p = pm.generate_gas_particles(n, prng=5)
            fields = {k: _np(t).copy() for k, t in p.fields.items()}
            p.set_field("gas", "metallicity", np.full(n, 0.1), passive_scalar=True)
            p.set_field("gas", "dye", np.zeros(n), passive_scalar=True)
            fields["gas", "metallicity"], fields["gas", "dye"] = np.full(n, 0.1), np.zeros(n)
            assert p.passive_scalars == ["metallicity", "dye"]
            ref = oracle.sample_clusters(fields, [om] * nh, c[:nh], v[:nh], radii=radii[:nh], resample=True,
                                         passive_scalars=["metallicity", "dye"])
            if nh == 2:
                out = cg.resample_two_clusters(p, pm, pm, c[0], c[1], v[0], v[1], radii[:2],
                                               passive_scalars=p.passive_scalars)
            else:
                out = cg.resample_three_clusters(p, pm, pm, pm, *c, *v, radii, passive_scalars=p.passive_scalars)
            for f in ("density", "thermal_energy", "particle_mass", "metallicity", "dye"):
                np.testing.assert_allclose(_np(out["gas", f]), ref["gas", f], rtol=1e-8, err_msg=f)
            zz = _np(out["gas", "metallicity"])
            assert (zz == 0.1).any() and (zz != 0.1).any()
    finally:
        om.fields.pop("metallicity"), om.fields.pop("dye")
        pm.fields.pop("metallicity"), pm.fields.pop("dye")


# --------------------------------------------------------------------------- #
# 8(f)3: table construction with the device integrator (utils.py:129-142, model.py:213-269)
# --------------------------------------------------------------------------- #
def test_quad_of_host_callables_replays_quadpack_on_the_device():
    """cgb_qags_replay against scipy.integrate.quad(full_output=1): same number of
    sub-intervals and error flag for every integral, values equal to the last bits."""
    import warnings

    from scipy.integrate import quad

    import cluster_generator_b200 as cg
    from cluster_generator_b200.utils import quad_callable_device

    rr = np.logspace(-1, 4, 400)
    cases = [(cg.hernquist_density_profile(1.0e15, 500.0), "hernquist"),
             (cg.vikhlinin_density_profile(4.2e5, 100.0, 1000.0, 1.0, 0.67, 3.0), "vikhlinin"),
             (cg.nfw_density_profile(1.0e7, 100.0) + cg.hernquist_density_profile(1.0e12, 20.0), "nfw+bcg")]
    for prof, name in cases:
        fn = lambda r: prof(r) * r * r                       # integrate_mass's integrand (utils.py:130)
        got, info = quad_callable_device(fn, np.zeros_like(rr), rr, return_info=True)
        want, last = np.empty_like(rr), np.empty(rr.size, int)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for i, r in enumerate(rr):
                res = quad(fn, 0, r, full_output=1)
                want[i], last[i] = res[0], res[2]["last"]
        assert (info["last"] == last).mean() > 0.995, (name, (info["last"] != last).sum())
        np.testing.assert_allclose(got, want, rtol=1e-13, err_msg=name)
        print(name, "levels", info["levels"], "bit-equal", (got == want).mean(), "max last", last.max())
    # integrate(): a spline-times-r integrand from r_i to r_max, incl. the empty last interval
    from scipy.interpolate import InterpolatedUnivariateSpline as IUS

    spl = IUS(rr, cases[0][0](rr))
    got = quad_callable_device(lambda r: spl(r) * r, rr, np.full_like(rr, rr[-1]))
    want = np.array([quad(lambda r: spl(r) * r, r, rr[-1])[0] for r in rr])
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-300)
    assert got[-1] == 0.0


@pytest.mark.parametrize("name", ["hernquist", "nfw_bcg", "vikhlinin"])
def test_cold_model_build_on_the_device_matches_the_host_tables(name):
    """ClusterModel.*(integrator="device"): every quad loop of the constructors on the device;
    the tables agree with the host (reference-pinned) construction to 1e-12 and the f(E) built
    from them to 1e-8; the cold build drops from seconds to tens of milliseconds."""
    import time

    import cluster_generator_b200 as cg

    pm = product_model(name)
    t0 = time.perf_counter()
    if name == "hernquist":
        dm = cg.ClusterModel.no_gas(0.1, 1.0e4, cg.hernquist_density_profile(1.0e15, 500.0), integrator="device")
    elif name == "nfw_bcg":
        star = cg.hernquist_density_profile(1.0e12, 20.0)
        dm = cg.ClusterModel.no_gas(0.1, 1.0e4, cg.nfw_density_profile(1.0e7, 100.0) + star, stellar_density=star,
                                    integrator="device")
    else:
        dm = cg.ClusterModel.from_dens_and_temp(
            0.1, 1.0e4, cg.vikhlinin_density_profile(4.2e5, 100.0, 1000.0, 1.0, 0.67, 3.0),
            cg.vikhlinin_temperature_profile(10.0, 0.0, 2.0, 1.0, 1500.0, 6.0, 80.0, 1.0), integrator="device")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(name, "cold model build with the device integrator: %.1f ms" % (1e3 * dt))
    assert set(dm.fields) == set(pm.fields)
    for k in pm.fields:
        np.testing.assert_allclose(np.asarray(dm[k]), np.asarray(pm[k]), rtol=1e-11, atol=1e-300, err_msg=k)
    np.testing.assert_allclose(dm.dm_virial.df.d, pm.dm_virial.df.d, rtol=1e-8)
    assert dt < 1.0


# --------------------------------------------------------------------------- #
# edge cases of the round-2 paths: empty and tiny inputs, tables too large for shared memory
# --------------------------------------------------------------------------- #
def test_round2_paths_edge_cases(oracle):
    import cluster_generator_b200 as cg
    from cluster_generator_b200 import _native as N
    from cluster_generator_b200.utils import quad_callable_device

    pm = product_model("vikhlinin")
    # host streaming with nothing, one particle, fewer particles than a chunk
    for n in (0, 1, 3, 4097):
        with _seeded(5):
            h = pm.generate_gas_particles(n, prng=0, host=True)
            d = pm.generate_gas_particles(n, prng=0)
            hd = pm.generate_dm_particles(n, prng=0, host=True, compute_potential=True)
            dd = pm.generate_dm_particles(n, prng=0, compute_potential=True)
        for a, b in ((h, d), (hd, dd)):
            assert a.num_particles == b.num_particles
            for k in b.fields:
                assert a[k].shape == b[k].shape and torch.equal(a[k], b[k].cpu()), (n, k)
    # the screen audit of nothing
    assert pm.dm_virial.screen_audit(0)["trials"] == 0
    # quad of host callables: no integrals, one integral, a zero-width interval
    assert quad_callable_device(np.sin, np.zeros(0), np.zeros(0)).shape == (0,)
    one = quad_callable_device(np.sin, np.array([0.0]), np.array([np.pi]))
    np.testing.assert_allclose(one, [2.0], rtol=1e-14)
    assert quad_callable_device(np.sin, np.array([1.5]), np.array([1.5]))[0] == 0.0
    # a radial grid whose packed gas tables exceed the shared-memory budget: the general kernel
    # runs (flags without the size fit), same numbers as a direct evaluation of the splines
    dens = cg.vikhlinin_density_profile(4.2e5, 100.0, 1000.0, 1.0, 0.67, 3.0)
    temp = cg.vikhlinin_temperature_profile(10.0, 0.0, 2.0, 1.0, 1500.0, 6.0, 80.0, 1.0)
    big = cg.ClusterModel.from_dens_and_temp(0.1, 1.0e4, dens, temp, num_points=2600, integrator="device")
    tabs = big._gas_sampler_tables(None, False)
    assert tabs.gas_flags & 1 and (tabs.gas_flags >> 8) * 256 > 220 * 1024
    with _seeded(9):
        g = big.generate_gas_particles(50_001, prng=0)
    r = torch.linalg.norm(g["gas", "particle_position"], dim=1).cpu().numpy()
    from scipy.interpolate import InterpolatedUnivariateSpline as IUS
    want = IUS(big["radius"].d, big["density"].d)(r)
    np.testing.assert_allclose(_np(g["gas", "density"]), want, rtol=1e-9)
    assert r.max() <= 1.0e4 * (1 + 1e-12) and np.isfinite(_np(g["gas", "thermal_energy"])).all()
    # the start-index tables pack 16-bit interval numbers: clear message beyond that
    d = N.SamplerDesc()
    d.n_cdf, d.n_psi = 70_000, 8
    dummy = torch.zeros(16, dtype=torch.float64, device="cuda")
    d.d_cdf_P = d.d_cdf_r = d.d_psi_t = d.d_psi_c = dummy.data_ptr()
    assert N.lib().cgb_sampler_workspace_bytes(N.ctypes.byref(d)) == 0
    assert b"n_cdf must be < 65535" in N.lib().cgb_last_error()
