"""Dev tool: host-streaming throughput vs chunk size (one GPU)."""
import sys, time
sys.path.insert(0, ".")
import torch
from tests.configs import product_model

pm = product_model("vikhlinin")
n = 10_000_000
pm.dm_virial
src = torch.empty(130_000_000, dtype=torch.float64, device="cuda")
dst = torch.empty(130_000_000, dtype=torch.float64, pin_memory=True)
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
    print("ceiling GB/s", 1.04 / (time.perf_counter() - t0))
del src, dst
for chunk in (1 << 19, 1 << 20, 1 << 21, 1 << 22, 1 << 23):
    hd = hg = True
    ts = []
    for i in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        hd = pm.generate_dm_particles(n, prng=i, host=hd, chunk=chunk)
        t1 = time.perf_counter()
        hg = pm.generate_gas_particles(n, prng=i + 1, host=hg, chunk=chunk)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        ts.append((t2 - t0, t1 - t0, t2 - t1))
    b = hd.d2h_bytes + hg.d2h_bytes
    t = min(ts)
    print(f"chunk {chunk:8d}: {t[0]*1e3:6.2f} ms (dm {t[1]*1e3:5.2f} gas {t[2]*1e3:5.2f})  {b/t[0]/1e9:5.1f} GB/s  dm {hd.d2h_bytes/t[1]/1e9:5.1f} gas {hg.d2h_bytes/t[2]/1e9:5.1f}")
