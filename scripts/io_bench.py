#!/usr/bin/env python
"""Time ClusterParticles.write_particles / write_to_gadget_file / from_file on device-resident
particles (bench.py's workload: 1e7 Vikhlinin DM + 1e7 gas, 1.28 GB of fields)."""
import os, sys, tempfile, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cluster_generator_b200 as cg
from tests.configs import product_model

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
pm = product_model("vikhlinin")
P = pm.generate_dm_particles(n, prng=1) + pm.generate_gas_particles(n, prng=1)
nbytes = sum(t.numel() * t.element_size() for t in P.fields.values())
for root in ("/dev/shm", tempfile.gettempdir()):
    d = tempfile.mkdtemp(dir=root)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        f = P.write_particles(os.path.join(d, "p.h5"), overwrite=True)
        t1 = time.perf_counter()
        print(f"{root}: write_particles {nbytes/1e9:.2f} GB in {t1-t0:.3f} s = {nbytes/1e9/(t1-t0):.2f} GB/s "
              f"({2*n/(t1-t0):.3g} particles/s)", flush=True)
    t0 = time.perf_counter()
    g = P.write_to_gadget_file(os.path.join(d, "g.h5"), 1e9, overwrite=True)
    t1 = time.perf_counter()
    print(f"{root}: write_to_gadget_file (fp32, {os.path.getsize(g)/1e9:.2f} GB) in {t1-t0:.3f} s", flush=True)
    t0 = time.perf_counter()
    Q = cg.ClusterParticles.from_file(f)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"{root}: from_file in {t1-t0:.3f} s = {nbytes/1e9/(t1-t0):.2f} GB/s", flush=True)
    assert all(torch.equal(P[k], Q[k]) for k in P.fields)
    del Q
    for x in os.listdir(d):
        os.remove(os.path.join(d, x))
    os.rmdir(d)
