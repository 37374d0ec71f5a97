"""Dev tool: where the host-streaming path loses against the one-big-copy D2H ceiling."""
import sys, time
sys.path.insert(0, ".")
import torch
from tests.configs import product_model

dev = torch.device("cuda")
N8 = 130_000_000
src = torch.empty(N8, dtype=torch.float64, device=dev)
dst = torch.empty(N8, dtype=torch.float64, pin_memory=True)
def timed(fn, reps=4):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best
t = timed(lambda: dst.copy_(src, non_blocking=True))
print(f"one copy of 1.04 GB: {t*1e3:.2f} ms  {1.04/t:.1f} GB/s")
for mib in (2, 4, 8, 16, 32, 64):
    w = mib * (1 << 20) // 8
    k = N8 // w
    def many():
        for i in range(k):
            dst[i * w:(i + 1) * w].copy_(src[i * w:(i + 1) * w], non_blocking=True)
    t = timed(many)
    print(f"{k:4d} copies of {mib:3d} MiB: {t*1e3:.2f} ms  {k*w*8/1e9/t:.1f} GB/s")
pm = product_model("vikhlinin")
n = 10_000_000
pm.dm_virial
for chunk in (1 << 19, 1 << 20, 1 << 21, 1 << 22):
    hd = hg = True
    ts = []
    for i in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        hd = pm.generate_dm_particles(n, prng=i, host=hd, chunk=chunk)
        hg = pm.generate_gas_particles(n, prng=i + 1, host=hg, chunk=chunk)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    b = hd.d2h_bytes + hg.d2h_bytes
    print(f"pipeline chunk {chunk:8d}: {min(ts)*1e3:6.2f} ms  {b/min(ts)/1e9:5.1f} GB/s")
