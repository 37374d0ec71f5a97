"""Kernel-only timing of the fused sampler for one model (dev tool).
usage: quick_bench.py [model] [n] ; env CGB200_LIB selects the library variant."""
import sys
sys.path.insert(0, ".")
import torch
from tests.configs import product_model
from cluster_generator_b200 import _native as N

name = sys.argv[1] if len(sys.argv) > 1 else "vikhlinin"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000_000
ptype = sys.argv[3] if len(sys.argv) > 3 else "dark_matter"
pm = product_model(name)
pv = pm.dm_virial if ptype == "dark_matter" else pm.star_virial
tabs = pv.sampler_tables()
dev = torch.device("cuda")
pos = torch.empty((3, n), dtype=torch.float64, device=dev); vel = torch.empty_like(pos)
mass = torch.empty(n, dtype=torch.float64, device=dev)
mom = torch.zeros(16, dtype=torch.float64, device=dev)
L = N.lib()
def run(seed):
    a = N.SampleArgs(); a.d_workspace = tabs.workspace.data_ptr(); a.id_begin, a.id_end, a.ld = 0, n, n
    a.particle_mass = 1.0; a.seed = seed; a.out_dtype = 0
    a.d_pos, a.d_vel, a.d_mass, a.d_moments = pos.data_ptr(), vel.data_ptr(), mass.data_ptr(), mom.data_ptr()
    N.check(L.cgb_sample_particles(N.ctypes.byref(a), N.ptr(tabs.scratch), tabs.scratch.numel(), N.stream_ptr()))
for i in range(3): run(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for i in range(10):
    mom.zero_(); e0.record(); run(10 + i); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
m = mom.cpu().numpy()
ts.sort()
hc = tabs.scratch[:16].cpu().numpy().view("uint32")
print(f"hard_count {hc[2]} tiles {hc[0]}", end="  ")
print(f"{name}/{ptype} n={n:.0e} median {ts[len(ts)//2]:.3f} ms  min {ts[0]:.3f} ms  {n/ts[len(ts)//2]/1e6:.2f} G particles/s  Tbar {m[1]/m[0]:.2f} coop {m[12]:.0f} exact/trial {m[15]/m[1]:.2e}")
