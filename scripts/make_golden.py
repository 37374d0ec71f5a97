#!/usr/bin/env python
"""Generate tests/golden/*.npz from the CPU oracle (run in the build container).

The reference package itself cannot be imported here (unyt/h5py/yt absent and its
Cython module is missing from the snapshot), so the golden
vectors are outputs of oracle/cg_oracle.py -- which calls the reference's own
scipy/numpy entry points in the reference's order -- frozen so that (a) drift of the
oracle or of scipy/numpy shows up in the CPU suite and (b) the GPU tests can compare
the CUDA path with committed numbers without running the oracle.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cg_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SEED = 20261018
N_PART = 512


def dump(name, model, ptype, key):
    v = O.VirialEquilibrium(model, ptype)
    u = O.device_rng_uniforms(SEED, 0, N_PART)
    f = v.generate_particles(N_PART, uniforms=u, compute_potential=True,
                             velocity_fn=lambda *a: O.generate_velocities_counter(*a, seed=SEED))
    rho, chk = v.check_virial()
    t, c, k = v.f._eval_args
    np.savez_compressed(
        os.path.join(OUT, f"{name}_{key}.npz"),
        radius=model["radius"], potential=model["gravitational_potential"],
        density=model[f"{ptype}_density"], mass=model[f"{ptype}_mass"],
        ee=v.ee, g=v.g, ff=v.ff, f_t=t, f_c=c[: v.num_elements],
        check_rho=rho, check_chk=chk, seed=np.uint64(SEED),
        p_radius=v.last["radius"], p_psi=v.last["psi"], p_speed=v.last["speed"], p_trials=v.last_trials,
        p_pos=f[key, "particle_position"], p_vel=f[key, "particle_velocity"],
        p_mass=f[key, "particle_mass"], p_pot=f[key, "particle_potential"],
    )
    print(name, key, "Tbar", v.last_trials.mean(), "max", v.last_trials.max())


def main():
    os.makedirs(OUT, exist_ok=True)
    npts = 300  # small tables keep the fixtures ~100 KB; the algorithms are size-independent
    m1 = O.no_gas(0.1, 1.0e4, O.hernquist_density_profile(1.0e15, 500.0), num_points=npts)
    dump("hernquist", m1, "dark_matter", "dm")
    star = O.hernquist_density_profile(1.0e12, 20.0)
    m3 = O.no_gas(0.1, 1.0e4, O.add_profiles(O.nfw_density_profile(1.0e7, 100.0), star),
                  stellar_density=star, num_points=npts)
    dump("nfw_bcg", m3, "dark_matter", "dm")
    dump("nfw_bcg", m3, "stellar", "star")
    dens = O.vikhlinin_density_profile(4.2e5, 100.0, 1000.0, 1.0, 0.67, 3.0)
    temp = O.vikhlinin_temperature_profile(10.0, 0.0, 2.0, 1.0, 1500.0, 6.0, 80.0, 1.0)
    m2 = O.from_dens_and_temp(0.1, 1.0e4, dens, temp, num_points=npts)
    dump("vikhlinin", m2, "dark_matter", "dm")
    g = O.generate_gas_particles(m2, N_PART, compute_potential=True,
                                 uniforms=O.device_rng_uniforms(SEED, 0, N_PART))
    np.savez_compressed(
        os.path.join(OUT, "vikhlinin_gas.npz"), radius=m2["radius"], gas_mass=m2["gas_mass"],
        pressure=m2["pressure"], density=m2["density"], potential=m2["gravitational_potential"],
        seed=np.uint64(SEED), p_pos=g["gas", "particle_position"], p_mass=g["gas", "particle_mass"],
        p_therm=g["gas", "thermal_energy"], p_dens=g["gas", "density"], p_pot=g["gas", "particle_potential"])
    # QAGS traces: scipy's own adaptive state for a few bins (decision-for-decision pin)
    from scipy.integrate import quad
    from scipy.interpolate import InterpolatedUnivariateSpline as IUS
    v = O.VirialEquilibrium(m1, "dark_matter")
    ee = v.ee
    spl = IUS(ee, m1["dark_matter_density"][::-1])
    rec = {}
    for b in (0, 17, 150, 299):
        r = quad(lambda t, e: 2 * spl(e - t * t, 1), 0.0, np.sqrt(ee[b]), epsabs=1.49e-5, epsrel=1.49e-5,
                 args=(ee[b],), full_output=1)
        L = r[2]["last"]
        rec[f"d{b}"] = np.concatenate([[r[0], r[1], L], r[2]["alist"][:L], r[2]["blist"][:L],
                                       r[2]["rlist"][:L], r[2]["elist"][:L]])
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r = quad(lambda e, psi: v.f(e) * np.sqrt(2 * (psi - e)), 0.0, ee[b], args=(ee[b],), full_output=1)
        L = r[2]["last"]
        rec[f"c{b}"] = np.concatenate([[r[0], r[1], L], r[2]["alist"][:L], r[2]["blist"][:L],
                                       r[2]["rlist"][:L], r[2]["elist"][:L]])
    np.savez_compressed(os.path.join(OUT, "qags_traces.npz"), **rec)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
