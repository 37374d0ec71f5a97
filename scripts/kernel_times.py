"""Dev tool: one pass over the path's kernels for an ncu launch list
(ncu --metrics gpu__time_duration.sum --clock-control none --csv python scripts/kernel_times.py)
and CUDA-event timings of the pieces when run plainly."""
import sys
sys.path.insert(0, ".")
import torch
from tests.configs import product_model

name = sys.argv[1] if len(sys.argv) > 1 else "vikhlinin"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000_000
pm = product_model(name)


def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


v = pm.dm_virial
def fresh():
    pm._fe_fit_cache = None
    v._generate_df()
print("f(E) build exact, fresh ms", timed(fresh))
print("f(E) build exact, shared grid ms", timed(lambda: v._generate_df()))
print("check_virial ms", timed(lambda: v.check_virial(), 2))
if "gas_mass" in pm.fields:
    pm.generate_gas_particles(n, prng=1)
    print("gas ms (incl. allocations)", timed(lambda: pm.generate_gas_particles(n, prng=1)))
pm.generate_dm_particles(n, prng=1)
print("dm ms (incl. allocations)", timed(lambda: pm.generate_dm_particles(n, prng=1)))
