"""GPU tests of the particle / model files written straight from device buffers
(cgb_h5_write_dataset_device: pack kernel + pinned staging + file write), checked against
the independent format oracle (oracle/h5_oracle.py) and against NumPy arithmetic."""
import os

import numpy as np
import pytest
import torch

import cluster_generator_b200 as cg
from cluster_generator_b200.h5io import H5Reader, H5Writer
from cluster_generator_b200.particles import KMS_PER_KPCMYR, ClusterParticles
from oracle import h5_oracle as H
from tests.configs import product_model

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


def _particles(n=200_000, dtype=torch.float64):
    pm = product_model("vikhlinin")
    p = pm.generate_dm_particles(n, prng=25, compute_potential=True, dtype=dtype)
    return p + pm.generate_gas_particles(n, prng=25, dtype=dtype), pm


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_write_particles_from_device(tmp_path, dtype):
    P, _ = _particles(dtype=dtype)
    p = P.write_particles(tmp_path / "p.h5")
    root = H.load(p)
    assert list(root["children"]) == ["dm", "gas"]
    for (ptype, name), t in P.fields.items():
        ds = root["children"][ptype]["children"][name]
        assert ds["data"].dtype == _np(t).dtype and ds["data"].shape == tuple(t.shape)
        np.testing.assert_array_equal(ds["data"], _np(t))
        assert ds["attrs"] == {"units": P.field_units[ptype, name]}
    Q = ClusterParticles.from_file(p)
    for k in P.fields:
        assert Q[k].is_cuda and Q[k].dtype == torch.float64
        assert torch.equal(Q[k], P[k].double())


def test_device_write_spans_many_staging_chunks(tmp_path):
    """(n, 3) fp64 rows of 3e6 particles = 72 MB = three 32 MiB staging chunks, last one partial."""
    n = 3_000_001
    g = torch.Generator(device="cuda").manual_seed(3)
    planes = torch.rand(3, n, generator=g, device="cuda", dtype=torch.float64)
    one = torch.rand(n, generator=g, device="cuda", dtype=torch.float32)
    rows = torch.rand(n // 7, 3, generator=g, device="cuda", dtype=torch.float64)  # row-major input
    p = str(tmp_path / "big.h5")
    with H5Writer(p) as w:
        w.write_device("/pos", planes.T)
        w.write_device("/pos32s", planes.T, out_dtype="float32", scale=KMS_PER_KPCMYR)
        w.write_device("/one", one, out_dtype="float32")
        w.write_device("/one64", one, out_dtype="float64", scale=0.5)
        w.write_device("/rows", rows)
        w.write_device("/empty", planes.T[:0])
    with H5Reader(p) as r:
        np.testing.assert_array_equal(r.read("/pos"), _np(planes.T))
        np.testing.assert_array_equal(r.read("/pos32s"), (_np(planes.T) * KMS_PER_KPCMYR).astype(np.float32))
        np.testing.assert_array_equal(r.read("/one"), _np(one))
        np.testing.assert_array_equal(r.read("/one64"), _np(one).astype(np.float64) * 0.5)
        np.testing.assert_array_equal(r.read("/rows"), _np(rows))
        assert r.info("/empty")[:2] == (np.float64, (0, 3))
        # the read direction: chunked upload + unpack into plane-major storage
        back = r.read_device("/pos", "cuda")
        assert back.shape == (n, 3) and back.T.is_contiguous() and torch.equal(back, planes.T)
        assert torch.equal(r.read_device("/pos32s", "cuda", torch.float32), (planes.T * KMS_PER_KPCMYR).float())
        assert torch.equal(r.read_device("/pos32s", "cuda", torch.float64, 2.0),
                           (planes.T * KMS_PER_KPCMYR).float().double() * 2.0)
        assert torch.equal(r.read_device("/one", "cuda", torch.float32), one)
        assert torch.equal(r.read_device("/one64", "cuda"), one.double() * 0.5)
        assert torch.equal(r.read_device("/rows", "cuda"), rows)
        assert r.read_device("/empty", "cuda").shape == (0, 3)
    H.load(p)


def test_gadget_file_from_device(tmp_path):
    P, _ = _particles()
    box = 2000.0
    P.add_offsets([1000.0, 1000.0, 1000.0], [0.1, 0.0, 0.0])
    g = P.write_to_gadget_file(tmp_path / "ic.h5", box)
    root = H.load(g)
    for ptype, gp in (("dm", "PartType1"), ("gas", "PartType0")):
        keep = _np(P._clip_to_box(ptype, box))
        assert 0 < keep.sum() < keep.size
        grp = root["children"][gp]["children"]
        np.testing.assert_array_equal(grp["Coordinates"]["data"],
                                      _np(P[ptype, "particle_position"])[keep].astype(np.float32))
        np.testing.assert_array_equal(grp["Velocities"]["data"],
                                      (_np(P[ptype, "particle_velocity"])[keep] * KMS_PER_KPCMYR).astype(np.float32))
        np.testing.assert_array_equal(grp["Masses"]["data"],
                                      (_np(P[ptype, "particle_mass"])[keep] * 1e-10).astype(np.float32))
    gas = root["children"]["PartType0"]["children"]
    keep = _np(P._clip_to_box("gas", box))
    np.testing.assert_array_equal(gas["InternalEnergy"]["data"],
                                  (_np(P["gas", "thermal_energy"])[keep] * KMS_PER_KPCMYR**2).astype(np.float32))
    h = root["children"]["Header"]["attrs"]
    n_dm, n_gas = int(_np(P._clip_to_box("dm", box)).sum()), int(keep.sum())
    np.testing.assert_array_equal(h["NumPart_ThisFile"], np.array([n_gas, n_dm, 0, 0, 0, 0], dtype=np.uint32))
    assert h["MassTable"][1] == np.float32(_np(P["dm", "particle_mass"])[0] * 1e-10)
    G = ClusterParticles.from_gadget_file(g)
    assert G.num_particles == {"gas": n_gas, "dm": n_dm}
    np.testing.assert_allclose(_np(G["dm", "particle_velocity"]),
                               _np(P["dm", "particle_velocity"])[_np(P._clip_to_box("dm", box))], rtol=1e-6, atol=1e-9)


def test_model_file_with_df_reproduces_particles(tmp_path):
    """model.py:392-397 / 197-209: a stored dm_df is reused, so the particles of the re-read
    model are bit-identical to the original's."""
    pm = product_model("hernquist")
    a = pm.generate_dm_particles(50_000, prng=11)
    p = pm.write_model_to_h5(tmp_path / "m.h5")
    root = H.load(p)
    np.testing.assert_array_equal(root["children"]["dm_df"]["data"], pm.dm_virial.df.d)
    m2 = cg.ClusterModel.from_h5_file(p)
    assert m2._dm_virial is not None
    b = m2.generate_dm_particles(50_000, prng=11)
    for k in a.fields:
        assert torch.equal(a[k], b[k])


def test_ics_particle_files(tmp_path):
    """ics.py:173-199: profiles given as model files; particles written per halo and reused."""
    pm = product_model("vikhlinin")
    mf = pm.write_model_to_h5(tmp_path / "halo.h5")
    c1, c2 = cg.compute_centers_for_binary(np.array([5000.0, 5000.0, 5000.0]), 3000.0, 500.0)
    num = {"dm": 40_000, "gas": 40_000}
    ics = cg.ClusterICs("bin", 2, [mf, mf], [c1, c2], [[0.5, 0, 0], [-0.5, 0, 0]], num_particles=num, r_max=5000.0)
    first = ics.setup_particle_ics(output_directory=str(tmp_path), prng=3)
    files = [str(tmp_path / f"bin_{i}_particles.h5") for i in range(2)]
    assert ics.particle_files == files and all(os.path.exists(f) for f in files)
    H.load(files[0])
    again = ics.setup_particle_ics(output_directory=str(tmp_path), prng=999)  # read back, prng unused
    for k in first.fields:
        assert torch.equal(first[k], again[k]), k
    ics2 = cg.ClusterICs("bin", 2, [pm, pm], [c1, c2], [[0.5, 0, 0], [-0.5, 0, 0]], num_particles=num,
                         r_max=5000.0, particle_files=files)
    third = ics2.setup_particle_ics()
    for k in first.fields:
        assert torch.equal(first[k], third[k]), k
    # ics.py:377-422: the relaxed-particle path re-samples the same fields from the same profiles
    dens0 = first["gas", "density"].clone()
    again2 = ics.resample_particle_ics(first)
    assert torch.isfinite(again2["gas", "thermal_energy"]).all()
    inside = ((first["gas", "particle_position"] - torch.as_tensor(c1, device="cuda")).norm(dim=1) < 5000.0)
    torch.testing.assert_close(again2["gas", "density"][inside], dens0[inside], rtol=1e-12, atol=0.0)


def test_write_amr_particles(tmp_path):
    """codes.py:16-73: the flat GAMER / FLASH / RAMSES particle file, hdf5 and fortran formats."""
    from scipy.io import FortranFile

    P, _ = _particles(n=50_000)
    ptype_num = {"dm": 2, "gas": 1}
    fn = cg.write_amr_particles(P, tmp_path / "amr.h5", ["dm", "gas"], ptype_num)
    root = H.load(fn)
    assert sorted(root["children"]) == ["particle_mass", "particle_position", "particle_type", "particle_velocity"]
    for f in ("particle_position", "particle_velocity", "particle_mass"):
        want = np.concatenate([_np(P["dm", f]), _np(P["gas", f])], axis=0)
        np.testing.assert_array_equal(root["children"][f]["data"], want)
    np.testing.assert_array_equal(root["children"]["particle_type"]["data"],
                                  np.concatenate([2.0 * np.ones(50_000), np.ones(50_000)]))
    with pytest.raises(IOError):
        cg.write_amr_particles(P, fn, ["dm"], ptype_num, overwrite=False)
    # cgs output scales by the kpc / Msun / Myr definitions of utils.py
    fn2 = cg.write_amr_particles(P, tmp_path / "amr_cgs.h5", ["dm"], ptype_num, in_cgs=True)
    pos = H.load(fn2)["children"]["particle_position"]["data"]
    np.testing.assert_allclose(pos, _np(P["dm", "particle_position"]) * 3.0856775809623245e21, rtol=1e-15)
    fn3 = cg.write_amr_particles(P, tmp_path / "amr.dat", ["dm", "gas"], ptype_num, format="fortran")
    with FortranFile(fn3, "r") as f:
        assert int(f.read_ints(np.int32)[0]) == 100_000
        rec = f.read_reals(np.float64).reshape(100_000, 7)
    np.testing.assert_array_equal(rec[:50_000, :3], _np(P["dm", "particle_position"]))
    np.testing.assert_array_equal(rec[50_000:, 6], _np(P["gas", "particle_mass"]))
