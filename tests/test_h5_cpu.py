"""
CPU tests of the native HDF5 subset (cgb_h5_* in include/cgb200.h; host-only entry points,
no GPU needed) and of the particle / model file layouts built on it.

Pinning: h5py/libhdf5 are not installed, so the format oracle (oracle/h5_oracle.py, an
independent pure-Python restatement of the HDF5 format specification that also enforces
libhdf5's open-time invariants) and the native reader are both checked against a file
written by the real libhdf5 that ships with scipy's test data; the native writer is then
checked against the oracle.
"""
import os
import subprocess
import sys
from collections import OrderedDict

import numpy as np
import pytest
import torch

import cluster_generator_b200 as cg
from cluster_generator_b200.h5io import H5Reader, H5Writer
from cluster_generator_b200.particles import KMS_PER_KPCMYR, ClusterParticles
from oracle import h5_oracle as H


def _libhdf5_file():
    import scipy.io

    p = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(p):
        pytest.skip("scipy test data (a libhdf5-written file) not installed")
    return p


def test_real_libhdf5_file_oracle_and_native_reader():
    """KAT: MATLAB v7.3 file = HDF5 (libhdf5 1.6, 512-byte user block); scipy's own test
    expects testdouble = 0, pi/4, ..., 2 pi."""
    p = _libhdf5_file()
    want = np.arange(0, 2 * np.pi + 1e-9, np.pi / 4).reshape(9, 1)
    root = H.load(p)
    assert list(root["children"]) == ["testdouble"]
    ds = root["children"]["testdouble"]
    np.testing.assert_allclose(ds["data"], want, rtol=0, atol=1e-15)
    assert ds["attrs"] == {"MATLAB_class": "double"}
    with H5Reader(p) as r:
        assert r.keys("/") == ["testdouble"]
        dtype, shape, off = r.info("/testdouble")
        assert dtype == np.float64 and shape == (9, 1) and off == ds["offset"]
        np.testing.assert_array_equal(r.read("/testdouble"), ds["data"])
        assert r.attrs("/testdouble") == {"MATLAB_class": "double"}
    raw = np.fromfile(p, dtype="<f8", count=9, offset=off)
    np.testing.assert_array_equal(raw, ds["data"].ravel())


def _write_sample(path, nfields=40):
    rng = np.random.default_rng(7)
    data, attrs = {}, {}
    dts = [np.float32, np.float64, np.int32, np.uint32, np.int64, np.uint64]
    with H5Writer(path) as w:
        for g in ("dm", "gas", "Header", "empty"):
            w.create_group("/" + g)
        w.create_group("/dm/sub")
        w.create_group("/dm/sub/subsub")
        for i in range(nfields):  # 40 links -> 5 symbol-table nodes under one B-tree node
            dt = dts[i % 6]
            shape = [(11,), (17, 3), (2, 3, 4), (5, 1, 2, 3)][i % 4]
            a = (rng.normal(size=shape) * 100).astype(dt)
            name = f"/gas/field_{i:02d}" if i % 5 else f"/gas/Z{i}"
            w.write(name, a)
            data[name] = a
            u = ["kpc/Myr", "", "Msun/kpc**3", "größe µ"][i % 4]
            w.set_attr(name, "units", u)
            attrs[name] = {"units": u}
        w.write("/dm/none", np.zeros((0, 3)))
        data["/dm/none"] = np.zeros((0, 3))
        w.write("/dm/sub/subsub/x", np.arange(5, dtype=np.float32))
        data["/dm/sub/subsub/x"] = np.arange(5, dtype=np.float32)
        w.write("/scalar", np.asarray(1000, dtype=np.int64))
        data["/scalar"] = np.asarray(1000, dtype=np.int64)
        hdr = OrderedDict([("Time", 0.0), ("NumFilesPerSnapshot", 1),
                           ("NumPart_ThisFile", np.array([1, 2, 3, 0, 4, 5], dtype=np.uint32)),
                           ("MassTable", np.arange(6.0)), ("f32", np.float32(1.5)),
                           ("long", "x" * 5000), ("none", np.zeros(0))])
        for k, v in hdr.items():
            w.set_attr("/Header", k, v)
        w.set_attr("/Header", "Time", 2.5)  # overwrite keeps one copy
        hdr["Time"] = 2.5
        attrs["/Header"] = hdr
        w.set_attr("/", "unit_system", "galactic")
        attrs["/"] = {"unit_system": "galactic"}
    return data, attrs


def _same_attrs(got, want):
    assert list(got) == list(want)
    for k, v in want.items():
        if isinstance(v, str):
            assert got[k] == v
        else:
            assert np.asarray(got[k]).dtype == (np.asarray(v).dtype if not isinstance(v, (int, float)) else
                                                np.dtype("int64" if isinstance(v, int) else "float64"))
            np.testing.assert_array_equal(np.asarray(got[k]), np.asarray(v))


def test_writer_output_satisfies_format_oracle(tmp_path):
    p = str(tmp_path / "sample.h5")
    data, attrs = _write_sample(p)
    root = H.load(p)  # raises on any structural violation
    assert root["eof"] == os.path.getsize(p)

    def get(path):
        n = root
        for part in [s for s in path.split("/") if s]:
            n = n["children"][part]
        return n

    for path, a in data.items():
        d = get(path)["data"]
        assert d.dtype == a.dtype and d.shape == a.shape
        np.testing.assert_array_equal(d, a)
    for path, a in attrs.items():
        _same_attrs(get(path)["attrs"], a)
    assert get("/empty") == {"attrs": {}, "kind": "group", "children": {}}
    assert sorted(n.encode() for n in get("/gas")["children"]) == [n.encode() for n in get("/gas")["children"]]
    # big datasets are page-aligned, and every dataset sits where the header says
    for path, a in data.items():
        if a.size:
            off = get(path)["offset"]
            np.testing.assert_array_equal(np.fromfile(p, dtype=a.dtype, count=a.size, offset=off), a.ravel())


def test_native_reader_round_trip(tmp_path):
    p = str(tmp_path / "sample.h5")
    data, attrs = _write_sample(p)
    with H5Reader(p) as r:
        assert r.keys("/") == ["Header", "dm", "empty", "gas", "scalar"]
        assert r.is_group("/", "dm") and not r.is_group("/", "scalar")
        assert r.keys("/empty") == [] and len(r.keys("/gas")) == 40
        for path, a in data.items():
            b = r.read(path)
            assert b.dtype == a.dtype and b.shape == a.shape
            np.testing.assert_array_equal(b, a)
        for path, a in attrs.items():
            _same_attrs(r.attrs(path), a)
        with pytest.raises(ValueError, match="no dataset"):
            r.read("/gas/missing")
        with pytest.raises(ValueError, match="no group"):
            r.keys("/nope")


def test_large_dataset_alignment(tmp_path):
    p = str(tmp_path / "big.h5")
    a = np.arange(300_000, dtype=np.float64)
    with H5Writer(p) as w:
        w.write("/a", a)
        w.write("/b", a[:10])
        w.write("/c", a)
    with H5Reader(p) as r:
        assert r.info("/a")[2] % 4096 == 0 and r.info("/c")[2] % 4096 == 0
        np.testing.assert_array_equal(r.read("/c"), a)
    H.load(p)


def test_writer_errors(tmp_path):
    p = str(tmp_path / "e.h5")
    with H5Writer(p) as w:
        w.create_group("/g")
        with pytest.raises(ValueError, match="already exists"):
            w.create_group("/g")
        with pytest.raises(ValueError, match="does not exist"):
            w.write("/missing/x", np.zeros(3))
        with pytest.raises(ValueError, match="does not exist"):
            w.set_attr("/nope", "a", 1)
        with pytest.raises(ValueError, match="outside the HDF5 subset"):
            w.write("/g/c", np.zeros(3, dtype=np.complex128))
    with pytest.raises(IOError, match="Cannot create"):
        H5Writer(p)
    H5Writer(p, overwrite=True).close()
    assert H.load(p)["children"] == {}
    p2 = str(tmp_path / "many.h5")
    w = H5Writer(p2)
    for i in range(257):
        w.write(f"/d{i}", np.zeros(1))
    with pytest.raises(ValueError, match="more than 256"):
        w.close()


def test_reader_rejects_garbage(tmp_path):
    p = tmp_path / "junk.h5"
    p.write_bytes(b"not an hdf5 file" * 100)
    with pytest.raises(IOError, match="not an HDF5 file"):
        H5Reader(str(p))
    with pytest.raises(IOError, match="Cannot open"):
        H5Reader(str(tmp_path / "absent.h5"))


def test_reader_survives_corruption(tmp_path):
    """Truncations and byte flips in the metadata must give IOError/ValueError, never a crash:
    run in a child process so a segfault would be seen as a non-zero exit."""
    p = str(tmp_path / "sample.h5")
    _write_sample(p, nfields=12)
    code = r"""
import sys, numpy as np
from cluster_generator_b200.h5io import H5Reader
raw = bytearray(open(sys.argv[1], 'rb').read())
rng = np.random.default_rng(0)
meta0 = len(raw) - 12000
bad = sys.argv[1] + '.bad'
n_err = 0
for trial in range(300):
    b = bytearray(raw)
    if trial % 3 == 0:
        b = b[:int(rng.integers(100, len(raw)))]
    else:
        for _ in range(int(rng.integers(1, 6))):
            pos = int(rng.integers(0, 96)) if rng.random() < 0.2 else int(rng.integers(max(meta0, 96), len(b)))
            b[pos] = int(rng.integers(0, 256))
    open(bad, 'wb').write(b)
    try:
        r = H5Reader(bad)
        for k in r.keys('/'):
            if r.is_group('/', k):
                for j in r.keys('/' + k):
                    try:
                        if not r.is_group('/' + k, j):
                            r.read('/' + k + '/' + j); r.attrs('/' + k + '/' + j)
                    except (IOError, ValueError, KeyError, UnicodeDecodeError):
                        n_err += 1
            r.attrs('/' + k)
        r.close()
    except (IOError, ValueError, KeyError, UnicodeDecodeError):
        n_err += 1
print('errors', n_err)
"""
    out = subprocess.run([sys.executable, "-c", code, p], capture_output=True, text=True, timeout=300,
                         cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert out.returncode == 0, out.stderr[-2000:]
    assert int(out.stdout.split()[-1]) > 50


# --------------------------------------------------------------------------- #
# particle / model files on host tensors (device tensors: tests/test_gpu_h5.py)
# --------------------------------------------------------------------------- #
def _planes(a):
    return torch.from_numpy(np.ascontiguousarray(a.T)).T


def _host_particles(n=1000, seed=1):
    rng = np.random.default_rng(seed)
    F = OrderedDict()
    F["dm", "particle_position"] = _planes(rng.uniform(-100, 1100, (n, 3)))
    F["dm", "particle_velocity"] = _planes(rng.normal(size=(n, 3)))
    F["dm", "particle_mass"] = torch.full((n,), 3.5e9, dtype=torch.float64)
    F["gas", "particle_position"] = _planes(rng.uniform(0, 1000, (n, 3)))
    F["gas", "particle_velocity"] = _planes(np.zeros((n, 3)))
    F["gas", "particle_mass"] = torch.from_numpy(rng.uniform(1e8, 2e8, n))
    F["gas", "thermal_energy"] = torch.from_numpy(rng.uniform(1, 2, n))
    F["gas", "density"] = torch.from_numpy(rng.uniform(1e3, 2e3, n))
    return ClusterParticles(["dm", "gas"], F)


def test_write_particles_layout_and_round_trip(tmp_path):
    """particles.py:337-363: group per ptype, dataset per field, units attribute."""
    P = _host_particles()
    p = P.write_particles(tmp_path / "p.h5")
    root = H.load(p)
    assert list(root["children"]) == ["dm", "gas"]
    assert sorted(root["children"]["gas"]["children"]) == sorted(P.field_names["gas"])
    for (ptype, name), t in P.fields.items():
        ds = root["children"][ptype]["children"][name]
        np.testing.assert_array_equal(ds["data"], t.numpy())
        assert ds["attrs"] == {"units": P.field_units[ptype, name]}
    with pytest.raises(IOError, match="overwrite=False"):
        P.write_particles(tmp_path / "p.h5")
    P.write_particles(tmp_path / "p.h5", overwrite=True)
    Q = ClusterParticles.from_file(p, device="cpu")
    assert Q.particle_types == ["dm", "gas"] and Q.num_particles == P.num_particles
    for k in P.fields:
        assert torch.equal(P[k], Q[k]) and Q.field_units[k] == P.field_units[k]
        if Q[k].dim() == 2:
            assert Q[k].T.is_contiguous()  # plane-major storage, as the samplers produce
    only = ClusterParticles.from_h5_file(p, ptypes="gas", device="cpu")
    assert only.particle_types == ["gas"] and ("dm", "particle_mass") not in only
    # the NumPy archive variant still works
    z = P.write_particles(tmp_path / "p.npz")
    Z = ClusterParticles.from_file(z, device="cpu")
    for k in P.fields:
        assert torch.equal(P[k], Z[k])


def test_gadget_file_layout(tmp_path):
    """particles.py:453-552: unit conversions, casts, clipping, ids, header."""
    P = _host_particles()
    box = 1000.0
    g = P.write_to_gadget_file(tmp_path / "ic.h5", box)
    root = H.load(g)
    assert list(root["children"]) == ["Header", "PartType0", "PartType1"]
    keep = P._clip_to_box("dm", box).numpy()
    assert 0 < keep.sum() < keep.size
    dm = root["children"]["PartType1"]["children"]
    assert sorted(dm) == ["Coordinates", "Masses", "ParticleIDs", "Velocities"]
    f32 = lambda a: a.astype(np.float32)
    np.testing.assert_array_equal(dm["Coordinates"]["data"], f32(P["dm", "particle_position"].numpy()[keep]))
    np.testing.assert_array_equal(dm["Velocities"]["data"],
                                  f32(P["dm", "particle_velocity"].numpy()[keep] * KMS_PER_KPCMYR))
    np.testing.assert_array_equal(dm["Masses"]["data"], f32(P["dm", "particle_mass"].numpy()[keep] * 1e-10))
    gas = root["children"]["PartType0"]["children"]
    assert sorted(gas) == ["Coordinates", "Density", "InternalEnergy", "Masses", "ParticleIDs", "Velocities"]
    np.testing.assert_array_equal(gas["InternalEnergy"]["data"],
                                  f32(P["gas", "thermal_energy"].numpy() * KMS_PER_KPCMYR**2))
    np.testing.assert_array_equal(gas["Density"]["data"], f32(P["gas", "density"].numpy() * 1e-10))
    # ids: 1-based, sequential across types in particle_types order (dm first here)
    n_dm = int(keep.sum())
    np.testing.assert_array_equal(dm["ParticleIDs"]["data"], np.arange(1, n_dm + 1, dtype=np.uint32))
    np.testing.assert_array_equal(gas["ParticleIDs"]["data"], np.arange(n_dm + 1, n_dm + 1001, dtype=np.uint32))
    h = root["children"]["Header"]["attrs"]
    assert list(h) == ["Time", "Redshift", "BoxSize", "Omega0", "OmegaLambda", "HubbleParam", "NumPart_ThisFile",
                       "NumPart_Total", "NumPart_Total_HighWord", "NumFilesPerSnapshot", "MassTable", "Flag_Sfr",
                       "Flag_Cooling", "Flag_StellarAge", "Flag_Metals", "Flag_Feedback", "Flag_DoublePrecision",
                       "Flag_IC_Info"]
    np.testing.assert_array_equal(h["NumPart_ThisFile"], np.array([1000, n_dm, 0, 0, 0, 0], dtype=np.uint32))
    assert h["NumPart_ThisFile"].dtype == np.uint32 and h["NumFilesPerSnapshot"].dtype == np.int64
    assert h["BoxSize"] == box and h["HubbleParam"] == 1.0
    want_mt = np.zeros(6)
    want_mt[1] = np.float32(3.5e9 * 1e-10)
    np.testing.assert_array_equal(h["MassTable"], want_mt)
    # float64 + arepo flavour
    g64 = P.write_to_gadget_file(tmp_path / "ic64.h5", box, dtype="float64", code="arepo")
    r64 = H.load(g64)
    assert r64["children"]["Config"]["attrs"] == {"VORONOI": 1}
    np.testing.assert_array_equal(r64["children"]["PartType1"]["children"]["Velocities"]["data"],
                                  P["dm", "particle_velocity"].numpy()[keep] * KMS_PER_KPCMYR)
    # read back into galactic units
    G = ClusterParticles.from_gadget_file(g64, device="cpu")
    assert G.particle_types == ["gas", "dm"] and G.num_particles == {"gas": 1000, "dm": n_dm}
    np.testing.assert_allclose(G["dm", "particle_velocity"].numpy(), P["dm", "particle_velocity"].numpy()[keep],
                               rtol=1e-15)
    np.testing.assert_allclose(G["gas", "thermal_energy"].numpy(), P["gas", "thermal_energy"].numpy(), rtol=1e-15)
    np.testing.assert_allclose(G["gas", "particle_mass"].numpy(), P["gas", "particle_mass"].numpy(), rtol=1e-15)
    assert G["dm", "particle_index"].dtype == torch.int64
    only = ClusterParticles.from_gadget_file(g64, ptypes=["dm"], device="cpu")
    assert only.particle_types == ["dm"]


def test_gadget_mass_table_fallback(tmp_path):
    """particles.py:326-332: no Masses dataset -> Header.MassTable."""
    p = str(tmp_path / "snap.h5")
    with H5Writer(p) as w:
        w.create_group("/PartType1")
        w.write("/PartType1/Coordinates", np.zeros((4, 3), dtype=np.float32))
        w.write("/PartType1/ParticleIDs", np.arange(1, 5, dtype=np.uint32))
        w.create_group("/Header")
        w.set_attr("/Header", "MassTable", np.array([0, 0.25, 0, 0, 0, 0.0]))
    G = ClusterParticles.from_gadget_file(p, device="cpu")
    np.testing.assert_array_equal(G["dm", "particle_mass"].numpy(), np.full(4, 0.25e10))


def test_model_file_round_trip(tmp_path):
    """model.py:342-397 / 149-211 layout: num_elements, unit_system, fields/<name> + units."""
    m = cg.ClusterModel.from_arrays(OrderedDict(
        radius=np.logspace(-1, 4, 50), density=np.linspace(2, 1, 50), total_mass=np.logspace(8, 15, 50),
        gas_fraction=np.linspace(0.1, 0.2, 50)))
    p = m.write_model_to_h5(tmp_path / "m.h5")
    root = H.load(p)
    assert root["attrs"] == {"unit_system": "galactic"}
    assert root["children"]["num_elements"]["data"].shape == () and root["children"]["num_elements"]["data"] == 50
    assert root["children"]["fields"]["children"]["density"]["attrs"] == {"units": "Msun/kpc**3"}
    assert root["children"]["fields"]["children"]["gas_fraction"]["attrs"] == {"units": "dimensionless"}
    m2 = cg.ClusterModel.from_h5_file(p)
    assert m2.num_elements == 50 and list(m2.keys()) == sorted(m.keys())
    for k in m.keys():
        np.testing.assert_array_equal(m[k].d, m2[k].d)
        assert m[k].units == m2[k].units
    m3 = cg.ClusterModel.from_h5_file(p, r_min=1.0, r_max=100.0)
    assert m3.num_elements == int(((m["radius"].d >= 1.0) & (m["radius"].d <= 100.0)).sum())
    with pytest.raises(IOError, match="overwrite=False"):
        m.write_model_to_h5(tmp_path / "m.h5")


def test_ics_budget_split_from_model_files(tmp_path):
    """ics.py:100-160 opens each profile file; the split equals the one from the models."""
    from tests.configs import product_model

    pm = product_model("vikhlinin")
    mf = pm.write_model_to_h5(tmp_path / "halo.h5")
    num = {"dm": 100_000, "gas": 50_000}
    a = cg.ClusterICs("a", 2, [mf, mf], [[0, 0, 0], [1, 1, 1]], [[0, 0, 0], [0, 0, 0]], num_particles=num,
                      r_max=[5000.0, 2000.0])
    b = cg.ClusterICs("b", 2, [pm, pm], [[0, 0, 0], [1, 1, 1]], [[0, 0, 0], [0, 0, 0]], num_particles=num,
                      r_max=[5000.0, 2000.0])
    assert dict(a.num_particles) == dict(b.num_particles)
    assert sum(a.num_particles["dm"]) == 100_000 and a.num_particles["dm"][0] > a.num_particles["dm"][1]
    assert a.particle_files == [None, None]


def _move_attrs_to_continuation(src, dst, obj_path):
    """Rewrite ``obj_path``'s v1 object header the way libhdf5 lays out a header that outgrew its
    first chunk: the attribute messages move to a continuation block at the end of the file, the
    first chunk keeps the other messages, a continuation message (0x0010) and NIL padding."""
    import struct

    raw = bytearray(open(src, "rb").read())
    root_hdr = struct.unpack_from("<Q", raw, 24 + 32 + 8)[0]

    def header_of(path):
        addr = root_hdr
        for part in [p for p in path.split("/") if p]:
            nm, _, hs = struct.unpack_from("<HII", raw, addr + 2)
            o, stab = addr + 16, None
            while o < addr + 16 + hs:
                t, sz = struct.unpack_from("<HH", raw, o)
                if t == 0x11:
                    stab = struct.unpack_from("<QQ", raw, o + 8)
                o += 8 + sz
            bt, hp = stab
            seg = struct.unpack_from("<Q", raw, hp + 24)[0]
            used = struct.unpack_from("<H", raw, bt + 6)[0]
            found = None
            for k in range(used):
                snod = struct.unpack_from("<Q", raw, bt + 24 + 16 * k + 8)[0]
                for e in range(struct.unpack_from("<H", raw, snod + 6)[0]):
                    noff, oaddr = struct.unpack_from("<QQ", raw, snod + 8 + 40 * e)
                    name = bytes(raw[seg + noff:raw.index(b"\0", seg + noff)]).decode()
                    if name == part:
                        found = oaddr
            addr = found
        return addr

    addr = header_of(obj_path)
    nm, _, hs = struct.unpack_from("<HII", raw, addr + 2)
    keep, moved, o = [], [], addr + 16
    while o < addr + 16 + hs:
        t, sz = struct.unpack_from("<HH", raw, o)
        (moved if t == 0x0C else keep).append(bytes(raw[o:o + 8 + sz]))
        o += 8 + sz
    assert moved, "object has no attributes"
    cont = b"".join(moved)
    cont_addr = (len(raw) + 7) // 8 * 8
    first = b"".join(keep) + struct.pack("<HHBxxxQQ", 0x10, 16, 0, cont_addr, len(cont))
    room = hs - len(first)
    assert room >= 0 and room % 8 == 0
    n_nil = 0
    while room > 0:                                   # NIL messages fill what is left
        body = min(room - 8, 65528)
        first += struct.pack("<HHBxxx", 0, body, 0) + b"\0" * body
        room -= 8 + body
        n_nil += 1
    raw[addr + 16:addr + 16 + hs] = first
    struct.pack_into("<H", raw, addr + 2, len(keep) + 1 + n_nil + len(moved))
    raw += b"\0" * (cont_addr - len(raw)) + cont
    struct.pack_into("<Q", raw, 24 + 16, len(raw))    # end-of-file address
    open(dst, "wb").write(raw)


def test_reader_follows_object_header_continuation_blocks(tmp_path):
    """Headers that h5py grows after creation (e.g. a Gadget ``Header`` group whose attributes are
    added one by one) spill into continuation blocks; both readers must follow them."""
    p = str(tmp_path / "ic.h5")
    P = _host_particles(50)
    P.write_to_gadget_file(p, 2000.0)
    q = str(tmp_path / "ic_cont.h5")
    _move_attrs_to_continuation(p, q, "/Header")
    a, b = H.load(p), H.load(q)
    assert list(b["children"]["Header"]["attrs"]) == list(a["children"]["Header"]["attrs"])
    with H5Reader(p) as r0, H5Reader(q) as r1:
        h0, h1 = r0.attrs("/Header"), r1.attrs("/Header")
        assert list(h0) == list(h1) and len(h1) == 18
        for k in h0:
            np.testing.assert_array_equal(h0[k], h1[k])
        np.testing.assert_array_equal(r0.read("/PartType1/Masses"), r1.read("/PartType1/Masses"))
    G = ClusterParticles.from_gadget_file(q, device="cpu")
    assert G.num_particles["gas"] == 50


def test_unit_tables():
    from cluster_generator_b200.particles import _to_galactic, gadget_unit_scale

    assert gadget_unit_scale("Velocities", "kpc/Myr") == KMS_PER_KPCMYR
    assert gadget_unit_scale("Masses", "Msun") == 1e-10 and gadget_unit_scale("Coordinates", "kpc") == 1.0
    assert gadget_unit_scale("InternalEnergy", "kpc**2/Myr**2") == KMS_PER_KPCMYR**2
    assert _to_galactic("km/s") * KMS_PER_KPCMYR == pytest.approx(1.0, rel=1e-15)
    assert KMS_PER_KPCMYR == pytest.approx(977.79222, rel=1e-7)
    with pytest.raises(ValueError, match="not one this package can convert"):
        _to_galactic("furlong/fortnight")


def test_ics_parameter_file_round_trip(tmp_path):
    """ics.py:202-320: YAML keys basename / num_halos / profileN / centerN / velocityN / num_*_particles / r_max."""
    import yaml

    from tests.configs import product_model

    pm = product_model("vikhlinin")
    mf = pm.write_model_to_h5(tmp_path / "halo.h5")
    a = cg.ClusterICs("bin", 2, [mf, mf], [[0.0, 1.0, 2.0], [3.0, 4.0, 5.0]], [[0.5, 0, 0], [-0.5, 0, 0]],
                      num_particles={"dm": 1000, "gas": 500}, r_max=[5000.0, 2000.0],
                      particle_files=[None, "halo2_particles.h5"])
    fn = a.to_file(tmp_path / "ics.yml")
    raw = yaml.safe_load(open(fn))
    assert list(raw)[:5] == ["basename", "num_halos", "profile1", "center1", "velocity1"]
    assert raw["particle_file2"] == "halo2_particles.h5" and "particle_file1" not in raw
    assert raw["num_dm_particles"] == 1000 and raw["r_max"] == [5000.0, 2000.0]
    b = cg.ClusterICs.from_file(fn)
    assert b.basename == "bin" and b.num_halos == 2 and b.profiles == [str(mf), str(mf)]
    np.testing.assert_array_equal(b.center, a.center)
    np.testing.assert_array_equal(b.velocity, a.velocity)
    assert dict(b.num_particles) == dict(a.num_particles) and b.particle_files == a.particle_files
    with pytest.raises(RuntimeError, match="overwrite=False"):
        a.to_file(fn)
    with pytest.raises(ValueError, match="model file names"):
        cg.ClusterICs("x", 1, [pm], [0, 0, 0], [0, 0, 0]).to_file(tmp_path / "x.yml")


def test_gadget_multi_file_layout(tmp_path):
    """One file per rank of an index-sharded run: same particles, ids and totals as the single file."""
    P = _host_particles(600, seed=4)
    box = 1000.0
    single = H.load(P.write_to_gadget_file(tmp_path / "one.h5", box))
    cut = {"dm": 250, "gas": 100}

    def shard(lo_hi):
        f = OrderedDict()
        for (ptype, name), v in P.fields.items():
            lo, hi = lo_hi[ptype]
            f[ptype, name] = _planes(v[lo:hi].numpy()) if v.dim() == 2 else v[lo:hi].clone()
        return ClusterParticles(["dm", "gas"], f)

    parts = [shard({k: (0, c) for k, c in cut.items()}), shard({k: (c, 600) for k, c in cut.items()})]
    counts = [{t: int(p._clip_to_box(t, box).sum()) for t in ("dm", "gas")} for p in parts]
    files = [p.write_to_gadget_file(tmp_path / "multi.h5", box, shard=(r, 2), shard_counts=counts)
             for r, p in enumerate(parts)]
    assert [os.path.basename(f) for f in files] == ["multi.0.h5", "multi.1.h5"]
    roots = [H.load(f) for f in files]
    hs, h0, h1 = (r["children"]["Header"]["attrs"] for r in [single] + roots)
    assert h0["NumFilesPerSnapshot"] == 2 and hs["NumFilesPerSnapshot"] == 1
    np.testing.assert_array_equal(h0["NumPart_Total"], hs["NumPart_Total"])
    np.testing.assert_array_equal(h0["NumPart_ThisFile"].astype(np.int64) + h1["NumPart_ThisFile"], hs["NumPart_Total"])
    np.testing.assert_array_equal(h1["NumPart_Total_HighWord"], np.zeros(6, dtype=np.uint32))
    np.testing.assert_array_equal(h0["MassTable"], hs["MassTable"])
    for g in ("PartType0", "PartType1"):
        for name, ds in single["children"][g]["children"].items():
            both = np.concatenate([r["children"][g]["children"][name]["data"] for r in roots])
            np.testing.assert_array_equal(both, ds["data"])       # ids included: they continue across files
    with pytest.raises(ValueError, match="shard_counts"):
        parts[0].write_to_gadget_file(tmp_path / "bad.h5", box, shard=(0, 2), shard_counts=counts[:1])
    G = ClusterParticles.from_gadget_file(files[1], device="cpu")
    assert G.num_particles == {"gas": counts[1]["gas"], "dm": counts[1]["dm"]}
    A = ClusterParticles.from_gadget_file(files[0], device="cpu", all_files=True)
    S = ClusterParticles.from_gadget_file(tmp_path / "one.h5", device="cpu", all_files=True)
    assert A.num_particles == S.num_particles
    for k in S.fields:
        assert torch.equal(A[k], S[k]), k


def test_h5py_opens_the_written_files(tmp_path):
    """Interchange with the real libhdf5: runs wherever h5py is installed -- it is not
    in the build image, where this test skips and ``oracle/h5_oracle.py`` (pinned by a file libhdf5
    wrote) stands in.  Datasets, shapes, dtypes, string and array attributes must come back."""
    h5py = pytest.importorskip("h5py")
    if getattr(h5py, "__cgb_test_stub__", False):
        pytest.skip("only the in-memory stand-in of tests/_refshim is importable")
    p = str(tmp_path / "x.h5")
    a = np.arange(12, dtype=np.float64).reshape(4, 3)
    b = np.arange(5, dtype=np.float32)
    with H5Writer(p) as w:
        w.create_group("/dm")
        w.write("/dm/particle_position", a)
        w.write("/dm/particle_mass", b)
        w.set_attr("/dm/particle_position", "units", "kpc")
        w.create_group("/Header")
        w.set_attr("/Header", "NumPart_ThisFile", np.array([0, 4, 0, 0, 0, 0], dtype=np.uint32))
        w.set_attr("/Header", "Time", 0.0)
    with h5py.File(p, "r") as f:
        assert sorted(f.keys()) == ["Header", "dm"]
        np.testing.assert_array_equal(f["dm/particle_position"][:], a)
        np.testing.assert_array_equal(f["dm/particle_mass"][:], b)
        u = f["dm/particle_position"].attrs["units"]
        assert (u.decode() if isinstance(u, bytes) else u) == "kpc"
        np.testing.assert_array_equal(f["Header"].attrs["NumPart_ThisFile"], [0, 4, 0, 0, 0, 0])
        assert f["Header"].attrs["Time"] == 0.0
    # and the other direction: a file h5py wrote, read by the native reader
    q = str(tmp_path / "y.h5")
    with h5py.File(q, "w") as f:
        g = f.create_group("gas")
        d = g.create_dataset("density", data=b.astype(np.float64))
        d.attrs["units"] = "Msun/kpc**3"
    with H5Reader(q) as r:
        np.testing.assert_array_equal(r.read("/gas/density"), b.astype(np.float64))
        assert r.attrs("/gas/density")["units"] == "Msun/kpc**3"
