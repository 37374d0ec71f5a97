"""Small end-to-end run of every kernel on the path, for the bounds-checked library
(`make -C cluster_generator_b200/csrc checked`, selected with CGB200_LIB) or compute-sanitizer.
Odd sizes, tile-boundary sizes, fp32, sub_sample, host streaming, the velocity seam, the
deferred long-tail pass, gas / tracers, ICs + resampling with passive scalars, field mapping,
histograms.  Prints "ok ..." when everything ran and the device raised nothing."""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
import cluster_generator_b200 as cg
from cluster_generator_b200 import _native as N

print("bounds checks enabled:", N.lib().cgb_bounds_check_enabled())
star = cg.hernquist_density_profile(1.0e12, 20.0)
pm = cg.ClusterModel.no_gas(0.1, 1.0e4, cg.nfw_density_profile(1.0e7, 100.0) + star, stellar_density=star,
                            num_points=200)
dm = pm.generate_dm_particles(20_001, prng=1, compute_potential=True)
for n in (1, 31, 127, 128, 129, 4097):
    pm.generate_dm_particles(n, prng=n)
st = pm.generate_star_particles(3_000, prng=2)          # mean trials ~300: cooperative finishes
sub = pm.generate_dm_particles(8_000, prng=3, sub_sample=4, dtype=torch.float32)
hst = pm.generate_dm_particles(10_001, prng=5, host=True, chunk=2048)
rho, chk = pm.check_dm_virial()
rho_s, chk_s = pm.check_star_virial()
# the Cython seam mirror (velocity-only mode of the sampler) on this model's f(E)
v = pm.dm_virial
n_seam = 5_003
r = np.geomspace(1.0, 5.0e3, n_seam)
psi = -np.interp(r, pm["radius"].d, pm["gravitational_potential"].d)
knots, coeffs = v.f.get_knots(), v.f.get_coeffs()
args = [N.dev_f64(a) for a in (psi, np.sqrt(2.0 * psi), 2.0 * psi * np.interp(psi, v.ee, v.ff), knots, coeffs)]
speed = torch.empty(n_seam, dtype=torch.float64, device="cuda")
trials = torch.empty(n_seam, dtype=torch.int32, device="cuda")
nb = N.lib().cgb_generate_velocities_workspace_bytes(coeffs.size)
work = N.workspace(nb)
N.check(N.lib().cgb_generate_velocities(N.ptr(args[0]), N.ptr(args[1]), N.ptr(args[2]), N.ptr(args[3]), knots.size,
                                        N.ptr(args[4]), coeffs.size, 3, N.ptr(None), 0, 17, 0,
                                        N.ptr(speed), N.ptr(trials), n_seam, N.ptr(work), nb, N.stream_ptr()))
# heavy tail: Hernquist sends a few particles per 1e6 to the deferred grid-wide pass
hq = cg.ClusterModel.no_gas(0.1, 1.0e4, cg.hernquist_density_profile(1.0e15, 500.0), num_points=300)
tail = hq.generate_dm_particles(1_000_003, prng=6)
dens = cg.vikhlinin_density_profile(4.2e5, 100.0, 1000.0, 1.0, 0.67, 3.0)
temp = cg.vikhlinin_temperature_profile(10.0, 0.0, 2.0, 1.0, 1500.0, 6.0, 80.0, 1.0)
gm = cg.ClusterModel.from_dens_and_temp(0.1, 1.0e4, dens, temp, num_points=200, integrator="device")
for n in (1, 2, 4095, 4096, 4097, 20_001):                # general kernel below 4096, packed above
    gm.generate_gas_particles(n, prng=n, compute_potential=True)
gm.generate_gas_particles(9_999, prng=7, sub_sample=3)
gm.generate_gas_particles(9_999, prng=8, host=True, chunk=1000)
gm.generate_tracer_particles(5_001, prng=9)
ics = cg.ClusterICs("s", 2, [gm, gm], [[0, 0, 0], [2000, 300, 0]], [[1, 0, 0], [-1, 0, 0]],
                    num_particles={"dm": 6000, "gas": 6000}, r_max=4000.0)
allp = ics.setup_particle_ics(prng=4)
ics.resample_particle_ics(allp)
torch.cuda.synchronize()
print("ok", dm.diagnostics["dm"].as_dict()["mean_trials"], st.diagnostics["star"].as_dict()["mean_trials"],
      tail.diagnostics["dm"].as_dict()["max_trials"], float(chk.max()), float(chk_s.max()), allp.num_particles)
