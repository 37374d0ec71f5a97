"""Virial-check histograms (cgb_particle_histograms; north_star item 4)."""
import numpy as np
import pytest
import torch
from scipy.interpolate import InterpolatedUnivariateSpline as IUS

from tests.configs import product_model

pytestmark = pytest.mark.gpu


def _host_hist(pv, p, key, nbins):
    """The kernel's binning rule in NumPy."""
    pos, vel = p[key, "particle_position"].double().cpu().numpy(), p[key, "particle_velocity"].double().cpu().numpy()
    r = np.sqrt(pos[:, 0] * pos[:, 0] + pos[:, 1] * pos[:, 1] + pos[:, 2] * pos[:, 2])
    v2 = vel[:, 0] * vel[:, 0] + vel[:, 1] * vel[:, 1] + vel[:, 2] * vel[:, 2]
    rr = pv.model["radius"].d
    psi = IUS(rr, pv.ee[::-1])(r)
    e_max = float(pv.ee.max())
    v_max = float(np.sqrt(2.0 * e_max))
    clip = lambda b: np.clip(b, 0, nbins - 1).astype(np.int64)
    br = clip(np.floor(np.log(r / rr[0]) * (nbins / np.log(rr[-1] / rr[0]))))
    bv = clip(np.floor(np.sqrt(v2) * (nbins / v_max)))
    be = clip(np.floor((psi - 0.5 * v2) * (nbins / e_max)))
    return [np.bincount(b, minlength=nbins) for b in (br, bv, be)]


@pytest.mark.parametrize("name,ptype,dtype", [("hernquist", "dark_matter", torch.float64),
                                              ("nfw_bcg", "stellar", torch.float64),
                                              ("vikhlinin", "dark_matter", torch.float32)])
def test_histograms_match_numpy_and_the_sampled_distribution(name, ptype, dtype):
    pm = product_model(name)
    pv = pm.dm_virial if ptype == "dark_matter" else pm.star_virial
    key = "dm" if ptype == "dark_matter" else "star"
    n, nbins = 1_000_000, 128
    p = pv.generate_particles(n, prng=5, dtype=dtype)
    h = pv.particle_histograms(p, nbins=nbins)
    assert h["n"] == n and all(int(h[k].sum()) == n for k in ("r", "v", "E"))
    want = _host_hist(pv, p, key, nbins)
    for k, w in zip(("r", "v", "E"), want):
        diff = np.abs(h[k].cpu().numpy() - w).sum()     # bin-edge rounding (log / spline ulps) only
        assert diff <= (4 if dtype == torch.float64 else 2000), (k, diff)
    # bound particles only: nothing beyond the escape speed / above psi_max
    assert int(h["E"][0]) < n and h["v_edges"][-1] == pytest.approx(np.sqrt(2 * pv.ee.max()))
    # the radial counts follow the piecewise-linear CDF the radius draw inverts
    chi2, dof = pv.radial_histogram_chi2(h)
    assert dof > 20 and chi2 < dof + 5.0 * np.sqrt(2.0 * dof), (chi2, dof)
    assert h["r_edges"].shape == (nbins + 1,) and h["r_edges"][0] == pytest.approx(pm["radius"].d[0])


def test_shard_histograms_sum_to_the_unsharded_ones():
    pm = product_model("hernquist")
    pv = pm.dm_virial
    n = 300_001
    full = pv.particle_histograms(pv.generate_particles(n, prng=9), nbins=64)
    acc = {k: torch.zeros_like(full[k]) for k in ("r", "v", "E")}
    for rank in range(3):
        part = pv.particle_histograms(pv.generate_particles(n, prng=9, shard=(rank, 3)), nbins=64)
        for k in acc:
            acc[k] += part[k]
    for k in acc:
        assert torch.equal(acc[k], full[k])
    empty = pv.particle_histograms(pv.generate_particles(0, prng=9), nbins=64)
    assert empty["n"] == 0 and int(empty["r"].sum()) == 0
    with pytest.raises(ValueError, match="nbins"):
        pv.particle_histograms(pv.generate_particles(10, prng=9), nbins=0)
