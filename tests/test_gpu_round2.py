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


def test_fitpack_exact_fit_small_and_odd_sizes_bit_equal():
    """The two-lane wavefront factorisation at full, odd and minimal sizes equals scipy's
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
    # sub_sample > 1 and rng="numpy" take the copy-after path
    with _seeded(13):
        a = pm.generate_dm_particles(40_000, prng=3, sub_sample=4, host=True)
        b = pm.generate_dm_particles(40_000, prng=3, sub_sample=4)
    assert torch.equal(a["dm", "particle_velocity"], b["dm", "particle_velocity"].cpu())
