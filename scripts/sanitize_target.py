"""Small end-to-end run for compute-sanitizer (memcheck / racecheck)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
import cluster_generator_b200 as cg

star = cg.hernquist_density_profile(1.0e12, 20.0)
pm = cg.ClusterModel.no_gas(0.1, 1.0e4, cg.nfw_density_profile(1.0e7, 100.0) + star, stellar_density=star,
                            num_points=200)
dm = pm.generate_dm_particles(20_001, prng=1, compute_potential=True)
st = pm.generate_star_particles(3_000, prng=2)
sub = pm.generate_dm_particles(8_000, prng=3, sub_sample=4, dtype=torch.float32)
rho, chk = pm.check_dm_virial()
dens = cg.vikhlinin_density_profile(4.2e5, 100.0, 1000.0, 1.0, 0.67, 3.0)
temp = cg.vikhlinin_temperature_profile(10.0, 0.0, 2.0, 1.0, 1500.0, 6.0, 80.0, 1.0)
gm = cg.ClusterModel.from_dens_and_temp(0.1, 1.0e4, dens, temp, num_points=200, integrator="device")
ics = cg.ClusterICs("s", 2, [gm, gm], [[0, 0, 0], [2000, 300, 0]], [[1, 0, 0], [-1, 0, 0]],
                    num_particles={"dm": 6000, "gas": 6000}, r_max=4000.0)
allp = ics.setup_particle_ics(prng=4)
torch.cuda.synchronize()
print("ok", dm.diagnostics["dm"].as_dict()["mean_trials"], st.diagnostics["star"].as_dict()["mean_trials"],
      float(chk.max()), allp.num_particles)
