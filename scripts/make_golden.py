#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN CODE in the build container.

``tests/_refshim`` imports ``/root/reference/cluster_generator/{utils,virial,model,
particles,radial_profiles}.py`` unmodified (stand-ins for the absent unyt / h5py and
an injected ``generate_velocities``, the one function whose source is missing from
the snapshot).  Every array written here comes out of the reference's
``ClusterModel.no_gas / from_dens_and_temp``, ``VirialEquilibrium._generate_df``,
``check_virial``, ``generate_particles`` and ``generate_gas_particles``; the device-RNG
uniforms reach the reference through a replaying ``RandomState`` (tests/refrun.py).
Only the speed magnitudes pass through the injected seam (counter-mode rejection
loop of oracle/cg_oracle_c.c) -- the one part that stays "parity unpinned".

The script refuses to write if the oracle disagrees with the reference anywhere
(bitwise), so the fixtures pin the oracle instead of freezing it.  The GPU box has
no /root/reference: the ``-m gpu`` tests read these files only.
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cg_oracle as O  # noqa: E402
from tests import refrun  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SEED = 20261018
N_PART = 2048
NPTS = 300  # small tables keep the fixtures ~100 KB; the algorithms are size-independent


def same(what, ref, orc):
    a, b = refrun.plain(ref), np.asarray(orc)
    if not np.array_equal(a, b):
        raise SystemExit(f"oracle != reference for {what}: max rel {np.nanmax(np.abs(a - b) / np.abs(b)):.3g}")


def dump(R, name, rmodel, omodel, ptype, key):
    rv, g = refrun.reference_virial_record(rmodel, ptype)
    ov = O.VirialEquilibrium(omodel, ptype)
    same("g", g, ov.g)
    same("df", rv.df, ov.df)
    u = O.device_rng_uniforms(SEED, 0, N_PART)
    with refrun.recording(R) as rec, refrun.velocities(R, "counter", SEED) as cap:
        parts = rv.generate_particles(N_PART, prng=refrun.replay_prng(u), compute_potential=True)
    radius = rec["radii"][0][0]
    of = ov.generate_particles(N_PART, uniforms=u, compute_potential=True,
                               velocity_fn=lambda *a: O.generate_velocities_counter(*a, seed=SEED))
    for fld in of:
        same(fld, parts[fld], of[fld])
    rho, chk = refrun.check_virial(rv)
    orho, ochk = ov.check_virial()
    same("check_virial rho", rho, orho)
    same("check_virial chk", chk, ochk)
    t, c, k = rv.f._eval_args
    np.savez_compressed(
        os.path.join(OUT, f"{name}_{key}.npz"),
        radius=rmodel["radius"].d, potential=rmodel["gravitational_potential"].d,
        density=rmodel[f"{ptype}_density"].d, mass=rmodel[f"{ptype}_mass"].d,
        ee=rv.ee, g=g, ff=rv.ff, f_t=t, f_c=c[: rv.num_elements],
        check_rho=rho, check_chk=refrun.plain(chk), seed=np.uint64(SEED),
        p_radius=radius, p_psi=cap.psi, p_speed=cap.speed, p_trials=cap.trials,
        p_pos=parts[key, "particle_position"].d, p_vel=parts[key, "particle_velocity"].d,
        p_mass=parts[key, "particle_mass"].d, p_pot=parts[key, "particle_potential"].d,
    )
    print(name, key, "Tbar", cap.trials.mean(), "max", cap.trials.max())


def main():
    if not refrun.available():
        raise SystemExit("needs the reference tree (/root/reference); run in the build container")
    os.makedirs(OUT, exist_ok=True)
    R = refrun.load_reference()
    models = {n: (refrun.reference_model(n, NPTS), refrun.oracle_model(n, NPTS))
              for n in ("hernquist", "nfw_bcg", "vikhlinin")}
    for n, (rm, om) in models.items():
        for k in om.fields:
            same(f"{n} table {k}", rm[k], om[k])
    dump(R, "hernquist", *models["hernquist"], "dark_matter", "dm")
    dump(R, "nfw_bcg", *models["nfw_bcg"], "dark_matter", "dm")
    dump(R, "nfw_bcg", *models["nfw_bcg"], "stellar", "star")
    dump(R, "vikhlinin", *models["vikhlinin"], "dark_matter", "dm")
    m2, o2 = models["vikhlinin"]
    u = O.device_rng_uniforms(SEED, 0, N_PART)
    g = m2.generate_gas_particles(N_PART, compute_potential=True, prng=refrun.replay_prng(u, gas=True))
    og = O.generate_gas_particles(o2, N_PART, compute_potential=True, uniforms=u)
    for fld in og:
        same(fld, g[fld], og[fld])
    np.savez_compressed(
        os.path.join(OUT, "vikhlinin_gas.npz"), radius=m2["radius"].d, gas_mass=m2["gas_mass"].d,
        pressure=m2["pressure"].d, density=m2["density"].d, potential=m2["gravitational_potential"].d,
        seed=np.uint64(SEED), p_pos=g["gas", "particle_position"].d, p_mass=g["gas", "particle_mass"].d,
        p_therm=g["gas", "thermal_energy"].d, p_dens=g["gas", "density"].d,
        p_pot=g["gas", "particle_potential"].d)
    # QAGS traces: scipy's own adaptive state for a few bins of the reference's integrands
    # (virial.py:47,55-62 and :98-99), on the reference-built tables
    from scipy.integrate import quad
    from scipy.interpolate import InterpolatedUnivariateSpline as IUS
    m1 = models["hernquist"][0]
    v, _ = refrun.reference_virial_record(m1, "dark_matter")
    ee = v.ee
    spl = IUS(ee, m1["dark_matter_density"].d[::-1])
    rec = {}
    for b in (0, 17, 150, 299):
        r = quad(lambda t, e: 2 * spl(e - t * t, 1), 0.0, np.sqrt(ee[b]), epsabs=1.49e-5, epsrel=1.49e-5,
                 args=(ee[b],), full_output=1)
        L = r[2]["last"]
        rec[f"d{b}"] = np.concatenate([[r[0], r[1], L], r[2]["alist"][:L], r[2]["blist"][:L],
                                       r[2]["rlist"][:L], r[2]["elist"][:L]])
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
