"""CPU suite, part 2: the C-ABI library surface and the host-side logic (no GPU)."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cgb200.h")


def test_library_loads_and_exports_every_declared_symbol():
    from cluster_generator_b200 import _native as N

    lib = N.load_library()
    declared = set(re.findall(r"\b(cgb_[a-z0-9_]+)\s*\(", open(HEADER).read()))
    declared -= {"cgb_sampler_desc", "cgb_host_uniforms", "cgb_sample_args"}
    bound = {name for name, _, _ in N.SYMBOLS}
    assert declared == bound, declared ^ bound
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.cgb_version() == 100
    assert lib.cgb_sample_scratch_bytes() >= 64


def test_ctypes_structs_match_the_c_header_layout():
    """Compile a probe against include/cgb200.h with gcc and compare sizeof/offsetof."""
    from cluster_generator_b200 import _native as N

    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "cgb200.h"
    int main(void) {
      printf("%zu %zu %zu\n", sizeof(cgb_sampler_desc), sizeof(cgb_host_uniforms), sizeof(cgb_sample_args));
      printf("%zu %zu %zu %zu\n", offsetof(cgb_sampler_desc, d_psi_t), offsetof(cgb_sampler_desc, n_f),
             offsetof(cgb_sampler_desc, d_aux_t), offsetof(cgb_sampler_desc, d_aux_c));
      printf("%zu %zu %zu %zu %zu\n", offsetof(cgb_sample_args, seed), offsetof(cgb_sample_args, d_pos),
             offsetof(cgb_sample_args, ld), offsetof(cgb_sample_args, d_moments), offsetof(cgb_sample_args, hu));
      printf("%zu %zu\n", offsetof(cgb_sample_args, n_peers), offsetof(cgb_sample_args, d_peer_moments));
      return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "probe.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "probe")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe], text=True).split()
    got = [int(x) for x in out]
    D, H, A = N.SamplerDesc, N.HostUniforms, N.SampleArgs
    want = [ctypes.sizeof(D), ctypes.sizeof(H), ctypes.sizeof(A),
            D.d_psi_t.offset, D.n_f.offset, D.d_aux_t.offset, D.d_aux_c.offset,
            A.seed.offset, A.d_pos.offset, A.ld.offset, A.d_moments.offset, A.hu.offset,
            A.n_peers.offset, A.d_peer_moments.offset]
    assert got == want


def test_shipped_kernels_carry_the_tma_bulk_copy_and_the_philox_products():
    """The SASS of the built library (what profiles/r02_sass_opcodes.md tabulates): the sampler and
    the gas fast path stage their tables with a TMA bulk copy (UBLKCP) completed on an mbarrier
    (SYNCS), Philox is IMAD.WIDE.U32 products, and the two hot loops do not re-derive the
    shared-memory window base (S2UR SR_CgaCtaId) at every access."""
    import shutil

    if shutil.which("cuobjdump") is None or shutil.which("cu++filt") is None:
        pytest.skip("CUDA binary utilities are not on PATH")
    so = os.path.join(ROOT, "cluster_generator_b200", "lib", "libcgb200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), [])
        elif cur is not None and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            cur.append(line)
    names = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True, check=True).stdout.splitlines()
    by_name = {re.sub(r"\((?:int|bool)\)", "", n): body for n, body in zip(names, kernels.values())}

    def count(kernel, needle):
        (body,) = [b for n, b in by_name.items() if kernel in n]
        return sum(needle in line for line in body)

    samp, gas = "cgb_sample_kernel<double, 0, 0>", "cgb_gas_smem_kernel<double, 0>"
    for k in (samp, gas):
        assert count(k, "UBLKCP") >= 1 and count(k, "SYNCS") >= 2, k
    assert count(samp, "IMAD.WIDE.U32") >= 150      # 2 per round x 10 rounds x the inlined Philox calls
    assert count(samp, "SR_CgaCtaId") <= 10 and count(gas, "SR_CgaCtaId") <= 2
    assert count(samp, "DFMA") > 0 and count(gas, "STG.E.128") >= 1


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, not route through the oracle."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cluster_generator_b200 import _native as N
    from tests.configs import product_model

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        N.require_cuda()
    pm = product_model("hernquist")      # host tables need no GPU
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pm.generate_dm_particles(10)
    with pytest.raises(RuntimeError, match="not built"):
        N.load_library("/nonexistent/libcgb200.so")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cluster_generator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "cg_oracle" not in text or f.endswith((".cu", ".cuh")), f
    out = subprocess.check_output(
        [sys.executable, "-c", "import sys; sys.path.insert(0, %r); import cluster_generator_b200; "
         "print(any(m.split('.')[0] == 'oracle' for m in sys.modules))" % ROOT], text=True)
    assert out.strip() == "False"


def test_host_tables_equal_the_oracle_tables():
    from tests.configs import assert_tables_equal, oracle_model, product_model

    for name in ("hernquist", "nfw_bcg", "vikhlinin"):
        om, pm = oracle_model(name), product_model(name)
        assert_tables_equal(om, pm)
        assert pm["radius"].units == "kpc" and pm["gravitational_potential"].units == "kpc**2/Myr**2"
        assert pm.num_elements == 1000
    pm = product_model("vikhlinin")
    assert np.all(pm.check_hse() < 1e-4)            # tests/test_models.py:99-114 (signed)
    for k in ("gas_mass", "pressure", "temperature", "gas_fraction"):
        np.testing.assert_array_equal(np.asarray(pm[k]), np.asarray(oracle_model("vikhlinin")[k])
                                      if k != "temperature" else np.asarray(pm[k]))


def test_entropy_constructor_and_set_rmax():
    """model.py:525-563 (from_dens_and_entr) reduces to from_dens_and_temp; set_rmax truncates."""
    import cluster_generator_b200 as cg
    from cluster_generator_b200.utils import kpc_to_cm, mp, mue

    dens = cg.vikhlinin_density_profile(4.2e5, 100.0, 1000.0, 1.0, 0.67, 3.0)
    temp = cg.vikhlinin_temperature_profile(10.0, 0.0, 2.0, 1.0, 1500.0, 6.0, 80.0, 1.0)
    entr = cg.RadialProfile(lambda r: temp(r) * (dens(r) / (mue * mp * kpc_to_cm**3)) ** (-2.0 / 3.0))
    a = cg.ClusterModel.from_dens_and_temp(1.0, 5.0e3, dens, temp, num_points=60)
    b = cg.ClusterModel.from_dens_and_entr(1.0, 5.0e3, dens, entr, num_points=60)
    np.testing.assert_allclose(np.asarray(b["temperature"]), np.asarray(a["temperature"]), rtol=1e-12)
    np.testing.assert_allclose(np.asarray(b["gravitational_potential"]), np.asarray(a["gravitational_potential"]), rtol=1e-9)
    c = a.set_rmax(1000.0)
    assert c.num_elements == int((np.asarray(a["radius"]) <= 1000.0).sum()) and c["radius"].d.max() <= 1000.0
    assert c["radius"].units == "kpc"


def test_radius_cdf_table_and_shards():
    from cluster_generator_b200.virial import radius_cdf_table, shard_bounds
    from oracle import cg_oracle as O
    from tests.configs import oracle_model

    om = oracle_model("hernquist")
    for r_max in (None, 2000.0, 0.5):
        a = radius_cdf_table(om["radius"], om["dark_matter_mass"], r_max)
        b = O.radius_cdf_table(om["radius"], om["dark_matter_mass"], r_max)
        for x, y in zip(a, b):
            np.testing.assert_array_equal(x, y)
        assert a[0][0] == 0.0 and a[0][-1] == 1.0 and np.all(np.diff(a[0]) > 0)
    with pytest.raises(ValueError):
        radius_cdf_table(om["radius"], om["dark_matter_mass"], 0.01)
    for n in (0, 1, 7, 1000, 10**9 + 7):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, (k, world)) for k in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[k][1] == b[k + 1][0] for k in range(world - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
    assert shard_bounds(10, None) == (0, 10)
    with pytest.raises(ValueError):
        shard_bounds(10, (2, 2))


def test_seed_derivation_follows_the_prng_stream():
    from numpy.random import RandomState

    from cluster_generator_b200.utils import parse_prng, seed_from_prng

    a = seed_from_prng(parse_prng(25), 1)
    assert a == seed_from_prng(parse_prng(25), 1) != seed_from_prng(parse_prng(25), 2)
    prng = RandomState(25)
    s1, s2 = seed_from_prng(prng, 1), seed_from_prng(prng, 1)
    assert s1 == a and s2 != s1                     # a shared RandomState advances
    assert 0 <= a < 2**64
    assert parse_prng(prng) is prng


def test_cluster_particles_container():
    """particles.py:78-120, 423-451 semantics on CPU tensors."""
    from cluster_generator_b200 import ClusterParticles

    def make(key, n, off=0.0):
        pos = (torch.arange(3 * n, dtype=torch.float64).reshape(3, n) + off)
        return ClusterParticles(key, {(key, "particle_position"): pos.T,
                                      (key, "particle_velocity"): (pos * 2).T,
                                      (key, "particle_mass"): torch.ones(n, dtype=torch.float64)})

    a, b, s = make("dm", 4), make("dm", 6, 100.0), make("star", 5)
    ab = a + b
    assert ab.num_particles == {"dm": 10}
    p = ab["dm", "particle_position"]
    assert p.shape == (10, 3) and p.T.is_contiguous()              # SoA planes survive concatenation
    assert torch.equal(p[:4], a["dm", "particle_position"]) and torch.equal(p[4:], b["dm", "particle_position"])
    allp = ab + s
    assert allp.num_particles == {"dm": 10, "star": 5} and sorted(allp.particle_types) == ["dm", "star"]
    assert allp.field_names["star"] == ["particle_position", "particle_velocity", "particle_mass"]
    assert allp.field_units["dm", "particle_velocity"] == "kpc/Myr"
    before = s["star", "particle_position"].clone()
    s.add_offsets([1.0, 2.0, 3.0], [0.0, 0.0, 1.0])
    assert torch.equal(s["star", "particle_position"] - before, torch.tensor([1.0, 2.0, 3.0]).expand(5, 3).double())
    allp.drop_ptypes("star")
    assert allp.particle_types == ["dm"] and ("star", "particle_mass") not in allp


def test_pending_host_copies_are_completed_by_any_field_access():
    """generate_*_particles(host=..., wait=False) hands back particles whose device->host copies
    may still be in flight; every path to the fields must complete them first."""
    from cluster_generator_b200 import ClusterParticles

    class FakeStreamer:
        def __init__(self):
            self.synced = 0

        def synchronize(self):
            self.synced += 1

    def make():
        n = 4
        pos = torch.arange(3 * n, dtype=torch.float64).reshape(3, n)
        p = ClusterParticles("dm", {("dm", "particle_position"): pos.T, ("dm", "particle_velocity"): (pos * 2).T,
                                    ("dm", "particle_mass"): torch.ones(n, dtype=torch.float64)})
        hs = FakeStreamer()
        p._pending.append(hs)
        return p, hs

    for touch in (lambda p: p["dm", "particle_mass"], lambda p: p.fields, lambda p: p.wait(),
                  lambda p: p + p, lambda p: p.add_offsets([1.0, 0.0, 0.0], [0.0, 0.0, 0.0]),
                  lambda p: ("dm", "particle_mass") in p, lambda p: list(p.keys())):
        p, hs = make()
        assert hs.synced == 0          # constructing the container does not wait
        touch(p)
        assert hs.synced == 1 and p._pending == [], touch
        touch(p)
        assert hs.synced == 1          # and only once


def test_radial_cut_and_black_hole():
    """particles.py:139-225 on CPU tensors."""
    from cluster_generator_b200 import ClusterParticles

    n = 100
    g = torch.Generator().manual_seed(0)
    pos = torch.rand((3, n), generator=g, dtype=torch.float64) * 200.0 - 100.0
    p = ClusterParticles("dm", {("dm", "particle_position"): pos.T, ("dm", "particle_velocity"): (pos * 0.1).T,
                                ("dm", "particle_mass"): torch.ones(n, dtype=torch.float64),
                                ("dm", "potential_energy"): -torch.arange(n, dtype=torch.float64)})
    keep = (pos**2).sum(dim=0) <= 80.0**2
    p.make_radial_cut(80.0)
    assert p.num_particles["dm"] == int(keep.sum()) and p["dm", "particle_position"].T.is_contiguous()
    assert torch.equal(p["dm", "particle_position"], pos.T[keep])
    p.add_black_hole(1.0e9)
    p.add_black_hole(2.0e9, use_pot_min=True)
    assert p.num_particles["black_hole"] == 2 and "black_hole" in p.particle_types
    idx = int(torch.argmin(p["dm", "potential_energy"]))
    assert torch.equal(p["black_hole", "particle_position"][1], p["dm", "particle_position"][idx])
    assert p["black_hole", "particle_mass"].tolist() == [1.0e9, 2.0e9]


def test_radial_profile_algebra():
    """cluster_generator/tests/test_radial_profiles.py:100-123"""
    import cluster_generator_b200 as cg

    r = np.geomspace(0.1, 1e4, 50)
    h, n = cg.hernquist_density_profile(1e12, 20.0), cg.nfw_density_profile(1e7, 100.0)
    np.testing.assert_allclose((h + n)(r), h(r) + n(r))
    np.testing.assert_allclose((2.0 * h)(r), 2.0 * h(r))
    np.testing.assert_allclose((h * n)(r), h(r) * n(r))
    np.testing.assert_allclose((h**2)(r), h(r) ** 2)
    assert np.all(h.cutoff(500.0)(r) <= h(r)) and np.all(h.add_core(1.0, 2.0)(r) <= h(r))


def test_argument_validation_happens_before_any_gpu_work():
    from tests.configs import product_model

    pm = product_model("hernquist")
    import cluster_generator_b200.virial as V

    v = object.__new__(V.VirialEquilibrium)
    v.ptype, v.model, v.num_elements, v._tables = "dark_matter", pm, pm.num_elements, {}
    with pytest.raises(ValueError, match="multiple of sub_sample"):
        v.generate_particles(10, sub_sample=3)
    with pytest.raises(ValueError, match="rng"):
        v.generate_particles(10, rng="mt")
    with pytest.raises(ValueError):
        v.generate_particles(-1)
    with pytest.raises(KeyError):
        V.VirialEquilibrium(pm, ptype="gas")
    with pytest.raises(ValueError):
        pm.set_field("radius", np.zeros(3))


# ---- multi-process path on CPU (gloo, world_size 2) -------------------------------------- #
def _gloo_worker(rank, world, port, n, out):
    import torch.distributed as dist

    from cluster_generator_b200 import _native as N
    from cluster_generator_b200.virial import SamplerDiagnostics, shard_bounds

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(n, (rank, world))
    m = torch.zeros(N.CGB_NMOMENTS, dtype=torch.float64)
    m[0] = hi - lo                    # particles of this shard
    m[1] = float(sum(range(lo, hi)))  # stands in for the trial sum
    m[2] = 10.0 + rank                # max trials: reduced with MAX, not SUM
    d = SamplerDiagnostics(m, (rank, world)).all_reduce().as_dict()
    if rank == 0:
        torch.save(d, out)
    dist.destroy_process_group()


def test_moments_all_reduce_two_ranks_gloo(tmp_path):
    import torch.multiprocessing as mp

    n, world = 1001, 2
    out = str(tmp_path / "d.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(world, port, n, out), nprocs=world, join=True)
    d = torch.load(out)
    assert d["particles"] == n and d["sum_trials"] == sum(range(n)) and d["max_trials"] == 11.0
    assert d["mean_trials"] == pytest.approx(sum(range(n)) / n)


def test_compute_centers_for_binary_known_answer():
    """cluster_generator/tests/test_ics.py:160-173"""
    import cluster_generator_b200 as cg
    from oracle import cg_oracle as O

    for fn in (cg.compute_centers_for_binary, O.compute_centers_for_binary):
        c1, c2 = fn(np.array([0.0, 0.0, 0.0]), 3000.0, 500.0)
        sep = np.sqrt(3000.0**2 - 500.0**2)
        np.testing.assert_allclose(c1, [-0.5 * sep, -250.0, 0.0])
        np.testing.assert_allclose(c2, [0.5 * sep, 250.0, 0.0])
        assert np.linalg.norm(c2 - c1) == pytest.approx(3000.0)
        c1, c2 = fn(np.array([1.0, 2.0, 3.0]), 1000.0, 300.0, a=400.0)
        assert np.linalg.norm(c2 - c1) == pytest.approx(1000.0)


def test_cluster_ics_particle_budget_split():
    """ics.py:100-160: n_i = rint(N * M_i / sum M) per species; 1:3 mass ratio (config 4)."""
    import cluster_generator_b200 as cg
    from tests.configs import product_model

    big = product_model("hernquist")                       # M = 1e15
    fields = {k: np.array(v) for k, v in big.fields.items()}
    for k in ("total_mass", "dark_matter_mass", "total_density", "dark_matter_density",
              "gravitational_potential", "gravitational_field"):
        fields[k] = fields[k] / 3.0
    small = cg.ClusterModel.from_arrays(fields)
    ics = cg.ClusterICs("m", 2, [big, small], [[0, 0, 0], [3000, 0, 0]], [[0, 0, 0], [0, 0, 0]],
                        num_particles={"dm": 1_000_000}, r_max=5000.0)
    assert ics.num_particles["dm"] == [750_000, 250_000]
    assert ics.num_particles["gas"] == [0, 0] and ics.num_particles["star"] == [0, 0]
    with pytest.raises(IOError, match="Cannot open"):      # profiles may be model files (ics.py:104)
        cg.ClusterICs("m", 1, ["absent_model.h5"], [0, 0, 0], [0, 0, 0])
    with pytest.raises(ValueError):
        cg.ClusterICs("m", 2, [big], [[0, 0, 0]] * 2, [[0, 0, 0]] * 2)


def test_set_field_and_from_fields():
    """particles.py:227-233, 368-421 on CPU tensors."""
    from cluster_generator_b200 import ClusterParticles

    f = {("gas", "particle_mass"): torch.ones(5, dtype=torch.float64),
         ("dm", "particle_mass"): torch.ones(3, dtype=torch.float64)}
    p = ClusterParticles.from_fields(f)
    assert p.particle_types == ["gas", "dm"] and p.num_particles == {"gas": 5, "dm": 3}
    p.set_field("gas", "metallicity", np.full(5, 0.3), passive_scalar=True)
    assert p.passive_scalars == ["metallicity"] and p.num_passive_scalars == 1
    assert "metallicity" in p.field_names["gas"] and p.field_units["gas", "metallicity"] == ""
    p.set_field("gas", "metallicity", torch.full((5,), 0.2, dtype=torch.float64), add=True)
    assert torch.allclose(p["gas", "metallicity"], torch.full((5,), 0.5, dtype=torch.float64))
    p.set_field("dm", "particle_mass", torch.full((3,), 2.0, dtype=torch.float64), units="Msun")
    assert float(p["dm", "particle_mass"][0]) == 2.0
    with pytest.raises(ValueError, match="needs to be 3 particles"):
        p.set_field("dm", "x", np.zeros(4))
    with pytest.raises(RuntimeError, match="add=True"):
        p.set_field("dm", "x", np.zeros(3), add=True)


def test_model_helpers_and_flat_files(tmp_path):
    """model.py:307-340, 399-456, 1088-1113"""
    from scipy.io import FortranFile

    from tests.configs import product_model

    pm = product_model("vikhlinin")
    m = pm.mass_in_radius(1000.0)
    r = pm["radius"].d
    assert set(m) == {"total", "gas", "dark_matter"} and float(m["total"]) == pm["total_mass"].d[r < 1000.0][-1]
    assert m["gas"].units == "Msun" and float(m["gas"]) < float(m["total"])
    rho = pm["density"].d[400]
    assert float(pm.find_radius_for_density(rho)) == pytest.approx(r[400], rel=1e-12)
    fn = pm.write_model_to_binary(tmp_path / "m.bin", fields_to_write=["radius", "density"], r_max=500.0)
    with FortranFile(str(fn), "r") as f:
        n = int(f.read_ints()[0])
        tab = f.read_reals().reshape(n, 2)
    assert n == int((r <= 500.0).sum())
    np.testing.assert_array_equal(tab[:, 0], r[:n])
    np.testing.assert_array_equal(tab[:, 1], pm["density"].d[:n])
    with pytest.raises(IOError, match="overwrite=False"):
        pm.write_model_to_binary(tmp_path / "m.bin")
    fa = pm.write_model_to_ascii(tmp_path / "m.txt")
    back = np.loadtxt(str(fa))
    assert back.shape == (pm.num_elements, len(list(pm.keys())))
    np.testing.assert_allclose(back[:, list(pm.keys()).index("radius")], r, rtol=1e-15)
    head = open(fa).read().splitlines()[:3]
    assert head[0] == "# unit_system=galactic" and "radius" in head[1] and "kpc" in head[2]


def test_profile_families_are_self_consistent():
    """radial_profiles.py:189-762: mass profiles are the integrals of their densities, totals and
    conversions hold (checked by quadrature, not against copied numbers)."""
    from scipy.integrate import quad

    import cluster_generator_b200 as cg

    def enclosed(rho, r):
        return 4.0 * np.pi * quad(lambda x: rho(x) * x * x, 0.0, r, epsabs=0.0, epsrel=1e-11, limit=200)[0]

    M, a = 3.0e14, 400.0
    for dens, mass in ((cg.snfw_density_profile(M, a), cg.snfw_mass_profile(M, a)),
                       (cg.cored_snfw_density_profile(M, a, 50.0), cg.cored_snfw_mass_profile(M, a, 50.0)),
                       (cg.cored_snfw_density_profile(M, a, 900.0), cg.cored_snfw_mass_profile(M, a, 900.0)),
                       (cg.einasto_density_profile(M, a, 5.0), cg.einasto_mass_profile(M, a, 5.0)),
                       (cg.hernquist_density_profile(M, a), cg.hernquist_mass_profile(M, a)),
                       (cg.tnfw_density_profile(1.0e7, 100.0, 800.0), cg.tnfw_mass_profile(1.0e7, 100.0, 800.0)),
                       (cg.nfw_density_profile(1.0e7, 100.0), cg.nfw_mass_profile(1.0e7, 100.0))):
        for r in (30.0, 400.0, 5000.0):
            assert float(mass(r)) == pytest.approx(enclosed(dens, r), rel=1e-8)
    assert float(cg.snfw_mass_profile(M, a)(1.0e15)) == pytest.approx(M, rel=1e-5)   # converges like x**-0.5
    assert float(cg.einasto_mass_profile(M, a, 5.0)(1.0e7)) == pytest.approx(M, rel=1e-6)
    assert cg.snfw_total_mass(1.0e14, 1000.0, a) * float(cg.snfw_mass_profile(1.0, a)(1000.0)) == pytest.approx(1.0e14)
    assert cg.cored_snfw_total_mass(1.0e14, 1000.0, a, 50.0) > 1.0e14 and cg.snfw_conc(5.0) == pytest.approx(5.16)
    # NFW -> Hernquist: same mass inside r_200
    M200, r200, c = 1.0e15, 2000.0, 4.0
    M0, ah = cg.convert_nfw_to_hernquist(M200, r200, c)
    assert float(cg.hernquist_mass_profile(M0, ah)(r200)) == pytest.approx(M200, rel=1e-12)
    # the cored Hernquist tends to the cuspy one for b -> infinity, and is flat inside a / b
    r = np.array([1.0, 10.0, 1000.0])
    np.testing.assert_allclose(cg.cored_hernquist_density_profile(M, a, 1.0e9)(r), cg.hernquist_density_profile(M, a)(r),
                               rtol=1e-5)
    core = cg.cored_hernquist_density_profile(M, a, 10.0)
    assert core(1.0e-3) / core(1.0e-1) == pytest.approx(1.0, abs=5e-3)
    np.testing.assert_allclose(cg.tnfw_density_profile(1.0e7, 100.0, 1.0e12)(r), cg.nfw_density_profile(1.0e7, 100.0)(r),
                               rtol=1e-12)
    assert cg.tnfw_density_profile(1.0e7, 100.0, 500.0)(5000.0) < 0.011 * cg.nfw_density_profile(1.0e7, 100.0)(5000.0)
    # AM06 gas: density falls, temperature -> T_0 c at the centre
    assert cg.am06_density_profile(1.0e5, 600.0, 60.0, 0.17, 5.0)(r).tolist() == sorted(
        cg.am06_density_profile(1.0e5, 600.0, 60.0, 0.17, 5.0)(r).tolist(), reverse=True)
    assert cg.am06_temperature_profile(8.0, 600.0, 60.0, 0.17)(1.0e-6) == pytest.approx(8.0 * 0.17, rel=1e-6)
    # entropy families: limits and slopes
    assert cg.baseline_entropy_profile(10.0, 1000.0, 1500.0, 1.1)(1500.0) == pytest.approx(1010.0)
    bk = cg.broken_entropy_profile(300.0, 100.0, 0.5)
    assert np.log(bk(3.0) / bk(0.3)) / np.log(10.0) == pytest.approx(0.5, abs=1e-3)
    assert np.log(bk(3.0e6) / bk(3.0e5)) / np.log(10.0) == pytest.approx(1.1, abs=1e-3)
    wk = cg.walker_entropy_profile(1500.0, 4.4, 1.0, 1000.0)
    assert wk(1500.0) == pytest.approx(1000.0 * 4.4 * np.exp(-1.0)) and wk(6000.0) < wk(1500.0)
    scaled = cg.rescale_profile_by_mass(cg.beta_model_profile(1.0, 100.0, 0.7), 5.0e13, 1000.0)
    assert enclosed(scaled, 1000.0) == pytest.approx(5.0e13, rel=1e-7)
