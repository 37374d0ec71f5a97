#!/usr/bin/env python
"""BASELINE.json configs 3 and 4 end to end through the public API on ONE GPU (timed with the
wall clock around a device synchronise; table builds excluded, reported separately).

config 3: NFW halo + Hernquist BCG, 1e8 particles split DM / stars in proportion to mass.
config 4: binary merger of two Vikhlinin-model clusters (mass ratio 3:1 by density scaling),
          n DM + n gas per halo, combined with gas re-sampling, written to HDF5 and to a Gadget
          IC file, read back.
usage: configs_demo.py [n_total_config3] [n_per_species_per_halo_config4] [out_dir]"""
import os, sys, tempfile, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cluster_generator_b200 as cg
from tests.configs import product_model

n3 = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
n4 = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000_000
out = sys.argv[3] if len(sys.argv) > 3 else tempfile.mkdtemp()


def timed(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"  {label}: {dt*1e3:9.1f} ms", flush=True)
    return r, dt


print(f"== config 3: NFW + BCG, {n3:.3g} particles on one GPU")
t0 = time.perf_counter(); pm = product_model("nfw_bcg"); print(f"  host tables: {time.perf_counter()-t0:.2f} s")
_, t = timed("f(E) builds (DM + stars, bit-equal to scipy)", lambda: (pm.dm_virial, pm.star_virial))
idx = pm["radius"].d < 5000.0
m_dm, m_st = pm["dark_matter_mass"].d[idx][-1], pm["stellar_mass"].d[idx][-1]
n_st = int(round(n3 * m_st / (m_dm + m_st)))                                    # budget in proportion to mass
ics = cg.ClusterICs("c3", 1, [pm], [0.0, 0.0, 0.0], [0.0, 0.0, 0.0],
                    num_particles={"dm": n3 - n_st, "star": n_st}, r_max=5000.0)
print("  split:", dict(ics.num_particles))
pm.generate_dm_particles(1000, prng=1)                                          # warm up the sampler tables
parts, t = timed("setup_particle_ics (first call: device allocations)", lambda: ics.setup_particle_ics(prng=25))
del parts
parts, t = timed("setup_particle_ics (second call)", lambda: ics.setup_particle_ics(prng=26))
n = sum(parts.num_particles.values())
print(f"  -> {n/t:.3g} particles/s; mean trials DM {parts.diagnostics['dm'].as_dict()['mean_trials']:.1f}, "
      f"stars {parts.diagnostics['star'].as_dict()['mean_trials']:.1f}")
del parts; torch.cuda.empty_cache()

print(f"== config 4: binary merger, {n4:.3g} DM + {n4:.3g} gas per halo, one GPU, files in {out}")
pm1 = product_model("vikhlinin")
# hydrostatic mass scales with the temperature normalisation: T / 3 (and rho / 3) -> mass ratio 3:1
dens = cg.vikhlinin_density_profile(4.2e5 / 3.0, 100.0, 1000.0, 1.0, 0.67, 3.0)
temp = cg.vikhlinin_temperature_profile(10.0 / 3.0, 0.0, 2.0, 1.0, 1500.0, 6.0 / 3.0, 80.0, 1.0)
t0 = time.perf_counter(); pm2 = cg.ClusterModel.from_dens_and_temp(0.1, 1.0e4, dens, temp, integrator="device")
print(f"  second model (device potential integrator): {time.perf_counter()-t0:.2f} s, "
      f"mass ratio {pm1['total_mass'].d[-1] / pm2['total_mass'].d[-1]:.2f}")
c1, c2 = cg.compute_centers_for_binary(np.array([7000.0, 7000.0, 7000.0]), 3000.0, 500.0)
ics = cg.ClusterICs("c4", 2, [pm1, pm2], [c1, c2], [[0.5, 0, 0], [-0.5, 0, 0]],
                    num_particles={"dm": 2 * n4, "gas": 2 * n4}, r_max=5000.0)
print("  split:", dict(ics.num_particles))
timed("f(E) builds", lambda: (pm1.dm_virial, pm2.dm_virial))
parts, t = timed("setup_particle_ics (generate 2 halos + combine + gas re-sampling; first call)", lambda: ics.setup_particle_ics(prng=7))
del parts
parts, t = timed("setup_particle_ics (second call)", lambda: ics.setup_particle_ics(prng=8))
n = sum(parts.num_particles.values())
print(f"  -> {n/t:.3g} particles/s")
nbytes = sum(v.numel() * v.element_size() for v in parts.fields.values())
f, t = timed(f"write_particles ({nbytes/1e9:.2f} GB HDF5)", lambda: parts.write_particles(os.path.join(out, "c4_particles.h5"), overwrite=True))
print(f"  -> {nbytes/1e9/t:.2f} GB/s")
g, t = timed("write_to_gadget_file (fp32, box 14 Mpc)", lambda: parts.write_to_gadget_file(os.path.join(out, "c4_ic.h5"), 14000.0, overwrite=True))
print(f"  -> {os.path.getsize(g)/1e9:.2f} GB, {os.path.getsize(g)/1e9/t:.2f} GB/s")
back, t = timed("from_file", lambda: cg.ClusterParticles.from_file(f))
assert all(torch.equal(parts[k], back[k]) for k in parts.fields)
print("  round trip bit-equal")
for x in ("c4_particles.h5", "c4_ic.h5"):
    os.remove(os.path.join(out, x))
