#!/usr/bin/env python
"""The fp32 interval screen against the exact fp64 accept test on >= 1e9 trials per model
(cgb_screen_audit).  Prints one line per model; `contradictions` must be 0.
usage: python scripts/screen_audit.py [trials_per_model=1e9]"""
import sys
import time

sys.path.insert(0, ".")
import torch  # noqa: E402

import cluster_generator_b200 as cg  # noqa: E402
from tests.configs import product_model  # noqa: E402

target = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0e9
dens = cg.vikhlinin_density_profile(4.2e5, 100.0, 1000.0, 1.0, 0.67, 3.0)
temp = cg.vikhlinin_temperature_profile(10.0, 0.0, 2.0, 1.0, 1500.0, 4.0, 80.0, 1.5)
negband = cg.ClusterModel.from_dens_and_temp(0.1, 1.0e4, dens, temp)
cases = [("hernquist DM", product_model("hernquist").dm_virial, 28.5), ("nfw DM", product_model("nfw_bcg").dm_virial, 23.3),
         ("bcg stars", product_model("nfw_bcg").star_virial, 323.0), ("vikhlinin DM", product_model("vikhlinin").dm_virial, 11.4),
         ("negative-f band DM", negband.dm_virial, 20.0)]
tot = {"trials": 0, "contradictions": 0, "undecided": 0}
for name, v, tbar in cases:
    n = int(1.1 * target / tbar)
    t0 = time.perf_counter()
    a = v.screen_audit(n, seed=987654321, max_trials=1 << 26)
    torch.cuda.synchronize()
    for k in tot:
        tot[k] += a[k]
    print(f"{name:20s} particles {n:10d} trials {a['trials']:13d} undecided {a['undecided']:9d} "
          f"({a['undecided'] / a['trials']:.2e}) contradictions {a['contradictions']} capped {a['capped_particles']} "
          f"screen-disabled particles {a['screen_disabled_particles']}  [{time.perf_counter() - t0:.1f} s]")
print(f"TOTAL trials {tot['trials']} undecided {tot['undecided']} contradictions {tot['contradictions']}")
assert tot["contradictions"] == 0
