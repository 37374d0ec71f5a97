#!/usr/bin/env python
"""
bench.py -- particles/sec of the particle-sampling hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # CUDA arm (this repo)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port of the
                                                           # reference's own CPU path

Workload (BASELINE.json configs[1]): Vikhlinin
density+temperature hydrostatic model, 1e7 DM particles (positions + velocities + mass,
Eddington f(E) + rejection sampling) plus 1e7 gas particles (positions, zero velocity,
mass, thermal energy, density) per GPU.  One "step" = one pass of both generators with
the tables resident in HBM.  Weak scaling: rank k of N generates the k-th contiguous
index shard of N * 1e7 particles per species; no data-path collective.  The only
collective, the sum of the 16 moment words over ranks, is inside the timed region: fused into
the sampler kernel over peer-mapped buffers (default), or an NCCL all_reduce per step with
CGB_BENCH_MOMENTS=nccl.

`e2e` is the same metric through the public API with host tables in and pinned host arrays
out (generate_*_particles(host=True): chunked generation, D2H copies overlapped with the
kernels), graded against a concurrent pinned-D2H ceiling measured in the same run
(`e2e.roofline`).  `configs` carries the other BASELINE configs at their stated sizes
(configs[0] 1e5 Hernquist + CPU oracle, configs[2] NFW+BCG at 1e8 and 1e9 through ClusterICs
over the N ranks, configs[3] binary merger + HDF5 write at N = 1, configs[4] sweep).

Prints ONE JSON line (rank 0).  Timing: CUDA events on the launching stream, barrier +
synchronize on both sides, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_PER_GPU = 10_000_000
WORKLOAD = "vikhlinin_dens_temp_hse: 1e7 DM (pos+vel+mass) + 1e7 gas (pos+vel0+mass+e+rho) per GPU, fp64"
METRIC = "particles/sec (pos+vel, device-timed)"
UNIT = "particles/s"
F_FIXED, F_TRIAL = 200.0, 52.0   # algorithmic flop per DM particle = 200 + 52 * T
B_DM, B_GAS = 56.0, 72.0         # algorithmic bytes written per particle (fp64)
NCU_PROFILE = os.path.join(ROOT, "profiles", "r02_ncu.json")  # written by scripts/make_profiles.py
SAMPLER_SRC = os.path.join(ROOT, "cluster_generator_b200", "csrc", "cgb_sampler.cu")


def code_sha256(path):
    """sha256 of a CUDA source with comments and white space removed: the hash follows the
    code the compiler sees, not its documentation."""
    import hashlib
    import re

    text = open(path, encoding="utf-8").read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)   # block comments
    text = re.sub(r"//[^\n]*", "", text)                 # line comments (no string in the sources holds "//")
    return hashlib.sha256("".join(text.split()).encode()).hexdigest()


def load_ncu_profile():
    """Per-launch ncu counters of the dominant kernel on THIS workload (smsp__inst_executed,
    dram bytes), captured by scripts/make_profiles.py together with the sha256 of the kernel
    source (code only, see code_sha256) they were measured on; `stale` when the code has
    changed since."""
    try:
        prof = json.load(open(NCU_PROFILE))
    except Exception:
        return None
    prof["stale"] = prof.get("sampler_sha256") != code_sha256(SAMPLER_SRC)
    return prof


def build_product_model():
    import cluster_generator_b200 as cg

    dens = cg.vikhlinin_density_profile(4.2e5, 100.0, 1000.0, 1.0, 0.67, 3.0)
    temp = cg.vikhlinin_temperature_profile(10.0, 0.0, 2.0, 1.0, 1500.0, 6.0, 80.0, 1.0)
    return cg.ClusterModel.from_dens_and_temp(0.1, 1.0e4, dens, temp)


def build_oracle_model():
    from oracle import cg_oracle as O

    dens = O.vikhlinin_density_profile(4.2e5, 100.0, 1000.0, 1.0, 0.67, 3.0)
    temp = O.vikhlinin_temperature_profile(10.0, 0.0, 2.0, 1.0, 1500.0, 6.0, 80.0, 1.0)
    return O.from_dens_and_temp(0.1, 1.0e4, dens, temp)


# --------------------------------------------------------------------------- #
# clocks during the timed region (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------- #
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread
    (5 ms cadence, so even a 40 ms region gets samples); nvidia-smi is the fallback."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.f = gpu_index, None, None
        self.thread, self.stop_flag, self.samples, self.reasons, self.max_mhz = None, False, [], set(), None

    def _poll(self):
        import pynvml as nv

        h = self.handle
        bits = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import threading

            import pynvml as nv

            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            self.handle = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            sm = sorted(float(x) for x in self.samples)
            if sm:
                out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=self.max_mhz, samples=len(sm), source="nvml 5 ms poll")
            out["reasons"] = sorted(self.reasons)
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.flush()
        self.f.seek(0)
        sm, reasons = [], set()
        for line in self.f:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                out["sm_max_mhz"] = float(p[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            sm.sort()
            top = sm[len(sm) // 2:]
            out["sm_mhz"] = top[len(top) // 2]
            out["samples"] = len(sm)
            out["source"] = "nvidia-smi 100 ms"
        out["reasons"] = sorted(reasons)
        return out


# --------------------------------------------------------------------------- #
# CPU baseline: the oracle port timed on the host
# --------------------------------------------------------------------------- #
def _cpu_worker(args):
    n, seed = args
    from oracle import cg_oracle as O

    om = _cpu_worker.model
    t0 = time.perf_counter()
    om.dm_virial.generate_particles(n, prng=seed)          # mode A sequential C loop + numpy
    O.generate_gas_particles(om, n, prng=seed + 1)
    return time.perf_counter() - t0


_ORACLE_MODEL = None


def cpu_baseline(n_sample, cores):
    """Oracle port of the reference CPU path: `cores` independent processes, each
    generating n_sample DM + n_sample gas particles.  Returns (particles/s, seconds)."""
    global _ORACLE_MODEL
    if _ORACLE_MODEL is None:
        _ORACLE_MODEL = build_oracle_model()
        _ORACLE_MODEL.dm_virial  # tables and f(E) outside the timed region, like the GPU arm
    _cpu_worker.model = _ORACLE_MODEL
    if cores == 1:
        t0 = time.perf_counter()
        _cpu_worker((n_sample, 25))
        dt = time.perf_counter() - t0
    else:
        import multiprocessing as mp

        ctx = mp.get_context("fork")
        with ctx.Pool(cores) as pool:
            pool.map(_cpu_worker, [(1000, 1)] * cores)  # warm the workers
            t0 = time.perf_counter()
            pool.map(_cpu_worker, [(n_sample, 25 + 2 * i) for i in range(cores)])
            dt = time.perf_counter() - t0
    return 2.0 * n_sample * cores / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    n_sample = 100_000
    vals = []
    if args.warmup > 0:
        cpu_baseline(20_000, cores)
    t_all = 0.0
    for _ in range(args.steps):
        v, dt = cpu_baseline(n_sample, cores)
        vals.append(v)
        t_all += dt
    value = sum(vals) / len(vals)
    sample = f"{cores} processes x ({n_sample} DM + {n_sample} gas) particles per step, oracle mode A"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------- #
# the other BASELINE configs, measured at their stated sizes
# --------------------------------------------------------------------------- #
def run_configs(args, rank, world, dev, barrier):
    """configs[0] (1e5 Hernquist DM, with the CPU oracle at the same size), configs[2] (NFW DM +
    BCG stars through ClusterICs at 1e8 and at the north-star size 1e9, sharded over the N
    ranks), configs[3] (binary merger incl. the HDF5 write; N = 1 only) and the configs[4]
    sweep (1e6 ... 1e9 Hernquist DM particles).  Device-timed (CUDA events, max over ranks,
    best of 3 after one warm-up call) plus the wall clock of the same call."""
    import shutil

    import numpy as np
    import torch
    import torch.distributed as dist

    import cluster_generator_b200 as cg
    from cluster_generator_b200 import _native as N

    f64 = torch.float64
    dfma = N.ctypes.c_double()
    N.check(N.lib().cgb_measure_dfma_tflops(4096, N.ctypes.byref(dfma), N.stream_ptr()))

    def timed(fn, reps=3):
        res = fn()
        best = (float("inf"), float("inf"))
        for _ in range(reps):
            del res
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            w0 = time.perf_counter()
            e0.record()
            res = fn()
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1), 1e3 * (time.perf_counter() - w0)], dtype=f64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, (float(t[0]), float(t[1])))
        return res, best[0], best[1]

    def species(p, key, n, ms):
        d = p.diagnostics[key].all_reduce().as_dict()
        flops = F_FIXED + F_TRIAL * d["mean_trials"]
        return {"particles": n, "device_ms": ms, "particles_per_s": n / (ms * 1e-3), "mean_trials": d["mean_trials"],
                "max_trials": d["max_trials"], "failures": d["failures"],
                "fp64_roofline_frac": n * flops / (ms * 1e-3) / 1e12 / (dfma.value * world)}

    out = {"note": "device_ms: CUDA events around the public-API call, max over ranks, best of 3; "
                   "wall_ms: barrier-to-barrier host clock of the same call",
           "n_gpus": world, "dfma_peak_tflops": dfma.value}
    shard = (rank, world)

    # ---- configs[0]: Hernquist DM-only, 1e5 DM particles, and the reference CPU path at that size
    m1 = cg.ClusterModel.no_gas(0.1, 1.0e4, cg.hernquist_density_profile(1.0e15, 500.0))
    m1.dm_virial
    n1 = 100_000
    p, ms, wall = timed(lambda: m1.generate_dm_particles(n1, prng=25, shard=shard))
    c1 = species(p, "dm", n1, ms)
    c1["wall_ms"] = wall
    if rank == 0:
        from oracle import cg_oracle as O

        om = O.no_gas(0.1, 1.0e4, O.hernquist_density_profile(1.0e15, 500.0))
        om.dm_virial
        t0 = time.perf_counter()
        om.dm_virial.generate_particles(n1, prng=25)
        c1["cpu_oracle_s"] = time.perf_counter() - t0
        c1["cpu_oracle_particles_per_s"] = n1 / c1["cpu_oracle_s"]
        c1["cpu_oracle_what"] = "oracle mode A (sequential C loop + numpy), 1 core, same size"
    out["config1_hernquist_1e5"] = c1
    del p

    # ---- configs[4]: sweep 1e6 ... 1e9 Hernquist DM particles over the N ranks (strong scaling in n_gpus)
    sweep = []
    for n in (1_000_000, 10_000_000, 100_000_000, 1_000_000_000):
        p, ms, wall = timed(lambda: m1.generate_dm_particles(n, prng=7, shard=shard))
        r = species(p, "dm", n, ms)
        r["wall_ms"] = wall
        sweep.append(r)
        del p
        torch.cuda.empty_cache()
    out["config5_sweep_hernquist_dm"] = sweep

    # ---- configs[2] (headline): NFW DM + Hernquist BCG stars through ClusterICs, split like ics.py:132-153
    star = cg.hernquist_density_profile(1.0e12, 20.0)
    m3 = cg.ClusterModel.no_gas(0.1, 1.0e4, cg.nfw_density_profile(1.0e7, 100.0) + star, stellar_density=star)
    m3.dm_virial, m3.star_virial
    mdm, mst = float(m3["dark_matter_mass"].d[-1]), float(m3["stellar_mass"].d[-1])
    c3 = []
    for ntot in (100_000_000, 1_000_000_000):
        nst = int(np.rint(ntot * mst / (mdm + mst)))   # equal-mass particles
        ndm = ntot - nst
        ics = cg.ClusterICs("c3", 1, [m3], [0.0, 0.0, 0.0], [0.0, 0.0, 0.0],
                            num_particles={"dm": ndm, "star": nst}, r_max=1.0e4)
        p, ms, wall = timed(lambda: ics.setup_particle_ics(prng=11, shard=shard))
        r = {"total_particles": ntot, "device_ms": ms, "wall_ms": wall, "particles_per_s": ntot / (ms * 1e-3),
             "what": "ClusterICs.setup_particle_ics(shard=(rank, N)): DM + star particles, positions + velocities + masses"}
        del p
        torch.cuda.empty_cache()
        pd, msd, _ = timed(lambda: m3.generate_dm_particles(ndm, r_max=1.0e4, prng=11, shard=shard))
        r["dm"] = species(pd, "dm", ndm, msd)
        del pd
        torch.cuda.empty_cache()
        ps, mss, _ = timed(lambda: m3.generate_star_particles(nst, r_max=1.0e4, prng=12, shard=shard))
        r["star"] = species(ps, "star", nst, mss)
        del ps
        if ntot == 1_000_000_000:
            r["north_star_target"] = {"target": ">= 1e9 DM+star particles in < 10 s on 8 x B200",
                                      "wall_s": wall * 1e-3, "n_gpus": world, "met": wall * 1e-3 < 10.0}
        c3.append(r)
        torch.cuda.empty_cache()
    out["config3_nfw_bcg"] = c3

    # ---- configs[3]: 1:3 binary merger, 1e8 DM + 1e8 gas split by mass, written to HDF5 (N = 1 only)
    if world == 1 and os.environ.get("CGB_BENCH_CONFIG4", "1") == "1":
        vd = dict(r_c=100.0, r_s=1000.0, alpha=1.0, beta=0.67, epsilon=3.0)
        vt = dict(a=0.0, b=2.0, c=1.0, r_t=1500.0, r_cool=80.0, a_cool=1.0)
        ma = cg.ClusterModel.from_dens_and_temp(0.1, 1.0e4, cg.vikhlinin_density_profile(rho_0=4.2e5, **vd),
                                                cg.vikhlinin_temperature_profile(T_0=10.0, T_min=6.0, **vt),
                                                integrator="device")
        mb = cg.ClusterModel.from_dens_and_temp(0.1, 1.0e4, cg.vikhlinin_density_profile(rho_0=1.4e5, **vd),
                                                cg.vikhlinin_temperature_profile(T_0=10.0 / 3.0, T_min=2.0, **vt),
                                                integrator="device")
        ma.dm_virial, mb.dm_virial
        ca, cb = cg.compute_centers_for_binary(np.array([7000.0] * 3), 3000.0, 500.0)
        ics = cg.ClusterICs("c4", 2, [ma, mb], [ca, cb], [[500.0, 0.0, 0.0], [-500.0, 0.0, 0.0]],
                            num_particles={"dm": 100_000_000, "gas": 100_000_000}, r_max=5000.0)
        p, ms, wall = timed(lambda: ics.setup_particle_ics(prng=3), reps=2)
        c4 = {"num_particles": {k: [int(x) for x in v] for k, v in ics.num_particles.items() if sum(v)},
              "mass_ratio": float(ma["total_mass"].d[-1] / mb["total_mass"].d[-1]),
              "device_ms": ms, "wall_ms": wall, "particles_per_s": 2.0e8 / (ms * 1e-3),
              "what": "ClusterICs.setup_particle_ics, 2 halos: generate, offset, concatenate, re-sample the gas of both halos"}
        nbytes = sum(t.numel() * t.element_size() for t in p.fields.values())
        target = None
        for d in ("/dev/shm", "/tmp"):
            try:
                if shutil.disk_usage(d).free > 2.5 * nbytes:
                    target = d
                    break
            except OSError:
                pass
        if target is not None:
            fn = os.path.join(target, f"cgb_bench_c4_{os.getpid()}.h5")
            try:
                t0 = time.perf_counter()
                p.write_particles(fn, overwrite=True)
                c4["hdf5_write_s"] = time.perf_counter() - t0
                c4["hdf5_bytes"] = os.path.getsize(fn)
                c4["hdf5_write_gbs"] = c4["hdf5_bytes"] / c4["hdf5_write_s"] / 1e9
                c4["hdf5_where"] = target
            finally:
                if os.path.exists(fn):
                    os.unlink(fn)
        else:
            c4["hdf5_write_s"] = None
            c4["hdf5_note"] = "no scratch file system with room for the file"
        out["config4_binary_merger"] = c4
        del p
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------- #
# CUDA arm
# --------------------------------------------------------------------------- #
def run_cuda(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import cluster_generator_b200 as cg
    from cluster_generator_b200 import _native as N

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # pinned host memory must come from the socket this GPU hangs off: bind before allocating
    from cluster_generator_b200.hoststream import bind_host_to_gpu

    cpus = bind_host_to_gpu(local) if os.environ.get("CGB_BENCH_BIND", "1") == "1" else []
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = N.lib()
    stream = torch.cuda.current_stream()

    # ---- tables (outside the timed region): host quad loops + device f(E) build ----
    pm = build_product_model()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pm.dm_virial  # first build (includes lazy module/kernel loading)
    import cluster_generator_b200.virial as V

    fe_ms, fe_fast_ms, fe_shared_ms = [], [], []
    for method, acc in (("pcr", fe_fast_ms), ("fitpack", fe_ms)):   # ends on the default (exact) one
        V.FE_SPLINE_FIT = method
        for _ in range(3):
            pm._fe_fit_cache = None   # a fresh build: the Givens factorisation of the energy grid included
            ev0.record()
            pm.dm_virial._generate_df()
            ev1.record()
            torch.cuda.synchronize()
            acc.append(ev0.elapsed_time(ev1))
    for _ in range(3):                # a further species of the same model re-uses the factorisation
        ev0.record()
        pm.dm_virial._generate_df()
        ev1.record()
        torch.cuda.synchronize()
        fe_shared_ms.append(ev0.elapsed_time(ev1))
    cv_ms = []
    for _ in range(3):
        ev0.record()
        pm.dm_virial.check_virial()
        ev1.record()
        torch.cuda.synchronize()
        cv_ms.append(ev0.elapsed_time(ev1))
    dm_tabs = pm.dm_virial.sampler_tables()
    gas_tabs = pm._gas_sampler_tables(None, False)

    n_glob = N_PER_GPU * world
    lo, hi = rank * N_PER_GPU, (rank + 1) * N_PER_GPU
    n = hi - lo
    f64 = torch.float64
    dm_pos, dm_vel = torch.empty((3, n), dtype=f64, device=dev), torch.empty((3, n), dtype=f64, device=dev)
    dm_mass = torch.empty(n, dtype=f64, device=dev)
    g_pos, g_vel = torch.empty((3, n), dtype=f64, device=dev), torch.empty((3, n), dtype=f64, device=dev)
    g_mass, g_e, g_rho = (torch.empty(n, dtype=f64, device=dev) for _ in range(3))
    moments = torch.zeros(N.CGB_NMOMENTS, dtype=f64, device=dev)

    def dm_args(seed):
        a = N.SampleArgs()
        a.d_workspace = dm_tabs.workspace.data_ptr()
        a.id_begin, a.id_end, a.ld = lo, hi, n
        a.particle_mass = dm_tabs.mtot / n_glob
        a.seed, a.out_dtype = seed, N.CGB_DTYPE_F64
        a.d_pos, a.d_vel, a.d_mass = dm_pos.data_ptr(), dm_vel.data_ptr(), dm_mass.data_ptr()
        a.d_moments = moments.data_ptr()
        return a

    def gas_args(seed):
        a = N.SampleArgs()
        a.d_workspace = gas_tabs.workspace.data_ptr()
        a.id_begin, a.id_end, a.ld = lo, hi, n
        a.particle_mass = gas_tabs.mtot / n_glob
        a.seed, a.out_dtype = seed, N.CGB_DTYPE_F64
        a.d_pos, a.d_vel, a.d_mass = g_pos.data_ptr(), g_vel.data_ptr(), g_mass.data_ptr()
        a.flags = gas_tabs.gas_flags
        return a

    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    g_ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]

    # The 16-word moment all-reduce is the path's only collective.  Default ("peer"): fused into
    # the sampler kernel -- every rank's final flush adds its partial sums to all ranks' buffers
    # over NVLink (cluster_generator_b200/peer.py), a fresh arena slot per step, nothing to wait
    # for but the kernels themselves.  "nccl": dist.all_reduce per step, asynchronously into one
    # of two buffers and drained before the closing event (the baseline this replaces).
    mom_mode = os.environ.get("CGB_BENCH_MOMENTS", "peer") if world > 1 else "local"
    if mom_mode not in ("peer", "nccl", "local"):
        raise SystemExit("CGB_BENCH_MOMENTS must be peer or nccl")
    peer = cg.PeerMoments(slots=args.steps + args.warmup + 8) if mom_mode == "peer" else None
    mom_ring = [moments, torch.zeros_like(moments)]
    pending = [None, None]
    last = {"mom": moments}

    def step(i, timed_idx=None):
        a, g = dm_args(1000 + i), gas_args(5000 + i)
        if peer is not None:
            slot = peer.next_slot()
            peer.attach(a, slot)
            last["mom"] = peer.slot(slot)
        else:
            moments = mom_ring[i & 1]
            if pending[i & 1] is not None:
                pending[i & 1].wait()
                pending[i & 1] = None
            moments.zero_()
            a.d_moments = moments.data_ptr()
            last["mom"] = moments
        sp = N.stream_ptr()
        if timed_idx is not None:
            k_ev[timed_idx][0].record()
        N.check(L.cgb_sample_particles(N.ctypes.byref(a), N.ptr(dm_tabs.scratch), dm_tabs.scratch.numel(), sp))
        if timed_idx is not None:
            k_ev[timed_idx][1].record()
        N.check(L.cgb_sample_gas(N.ctypes.byref(g), N.ptr(g_e), N.ptr(g_rho), sp))
        if timed_idx is not None:
            g_ev[timed_idx].record()
        if mom_mode == "nccl":
            pending[i & 1] = dist.all_reduce(last["mom"], async_op=True)

    def drain():
        for k in (0, 1):
            if pending[k] is not None:
                pending[k].wait()
                pending[k] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    drain()
    barrier()
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()  # started right before the timed region: every sample is under load
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        step(args.warmup + i, i)
    drain()
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1)
    clk = clocks.stop() if rank == 0 else None
    k_ms = sum(a.elapsed_time(b) for a, b in k_ev) / args.steps
    g_ms = sum(k[1].elapsed_time(g) for k, g in zip(k_ev, g_ev)) / args.steps
    mom = last["mom"].cpu().numpy()  # last step, summed over ranks when world > 1
    tt = torch.tensor([ms, k_ms, g_ms], dtype=f64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms, k_ms, g_ms = float(tt[0]), float(tt[1]), float(tt[2])
    ms_per_step = ms / args.steps
    value = 2.0 * n_glob / (ms_per_step * 1e-3)
    tbar = mom[1] / max(mom[0] - mom[13], 1.0)

    # ---- e2e: the public API with host tables in and host particles out (rank-local shard) ----
    # generate_*_particles(host=...) generates in id-range chunks and overlaps every chunk's
    # device->host copy with the next chunk's kernels (cluster_generator_b200/hoststream.py).
    host_dm = host_gas = True

    def e2e_step(i):
        nonlocal host_dm, host_gas
        # host -> device: the radial tables (rebuilt from the host arrays every step)
        pm.dm_virial._tables.clear()
        pm._gas_tables.clear()
        h2d = 8 * (12 * pm.num_elements + 4)  # CDF + spline-fit inputs of both species
        # wait=False: the gas species' table preparation and first chunks overlap the tail of the
        # dark-matter copies; both species are complete after the two wait() calls below
        host_dm = pm.generate_dm_particles(n_glob, prng=i, shard=(rank, world), host=host_dm, wait=False)
        host_gas = pm.generate_gas_particles(n_glob, prng=i + 1, shard=(rank, world), host=host_gas, wait=False)
        # device -> host read of the step's result: the moments (mean trials) of this step
        _ = host_dm.diagnostics["dm"].as_dict()["mean_trials"]
        host_dm.wait()
        host_gas.wait()
        return h2d, host_dm.d2h_bytes + host_gas.d2h_bytes

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step(0)   # allocates the pinned buffers, re-used by the timed steps
    barrier()
    w0 = time.perf_counter()
    for i in range(e2e_steps):
        h2d, d2h = e2e_step(1 + i)
    barrier()
    e2e_s = (time.perf_counter() - w0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=f64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te[0])
    e2e_value = 2.0 * n_glob / e2e_s

    # the ceiling the e2e number is graded against: every rank copies the same d2h bytes from a
    # device buffer into its pinned host buffer with plain cudaMemcpyAsync, same barrier
    # discipline, best of 3 -- what the host can ingest from all N GPUs at once
    src = torch.empty(d2h // 8, dtype=f64, device=dev)
    dst = torch.empty(d2h // 8, dtype=f64, pin_memory=True)
    ceil_s = []
    for _ in range(4):
        barrier()
        w0 = time.perf_counter()
        dst.copy_(src, non_blocking=True)
        barrier()
        tc = torch.tensor([time.perf_counter() - w0], dtype=f64, device=dev)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        ceil_s.append(float(tc[0]))
    ceil_s = min(ceil_s[1:])
    del src, dst
    e2e_roofline = {"bound": "pcie_d2h", "achieved": world * d2h / e2e_s / 1e9, "peak": world * d2h / ceil_s / 1e9,
                    "unit": "GB/s", "frac": ceil_s / e2e_s,
                    "peak_source": f"measured live: {world} ranks x {d2h / 1e9:.2f} GB pinned cudaMemcpyAsync D2H at once",
                    "host_cpus_bound": len(cpus)}
    del host_dm, host_gas

    # ---- the other BASELINE configs at their stated sizes, device-timed ----
    configs = run_configs(args, rank, world, dev, barrier) if os.environ.get("CGB_BENCH_CONFIGS", "1") == "1" else None

    if peer is not None:
        mom = mom.copy()
        peer.close()  # collective: unmap the peers' buffers, then free ours
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (cgb_sample_kernel) ----
    dfma = N.ctypes.c_double()
    ffma = N.ctypes.c_double()
    sms = N.ctypes.c_int()
    N.check(L.cgb_device_sm_count(N.ctypes.byref(sms)))
    N.check(L.cgb_measure_dfma_tflops(4096, N.ctypes.byref(dfma), N.stream_ptr()))
    N.check(L.cgb_measure_ffma_tflops(4096, N.ctypes.byref(ffma), N.stream_ptr()))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    prof = load_ncu_profile()
    winst = prof["warp_instructions_per_launch"] * n / prof["n_particles"] if prof else None
    flops_per_particle = F_FIXED + F_TRIAL * tbar
    k_s = k_ms * 1e-3
    ach_tflops = n * flops_per_particle / k_s / 1e12
    ach_gbs = n * B_DM / k_s / 1e9
    t_pipe = n * flops_per_particle / (dfma.value * 1e12)
    t_hbm = n * B_DM / (hbm_peak * 1e9)
    roofline = {
        "kernel": "cgb_sample_kernel<double,0,false>",
        "bound": "fp64_pipe" if t_pipe >= t_hbm else "hbm",
        "achieved": ach_tflops, "peak": dfma.value, "unit": "TFLOP/s", "frac": ach_tflops / dfma.value,
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel on this
        # workload, from the ncu --set full capture summarised in profiles/r01_ncu_full_summary.md
        "traffic": prof["dram_bytes_per_launch"] * n / prof["n_particles"] if prof else None,
        "ncu_profile": {"file": os.path.relpath(NCU_PROFILE, ROOT), "stale": prof["stale"]} if prof else None,
        "peak_source": "DFMA microbenchmark measured live (cgb_measure_dfma_tflops); MEASURED_PEAKS.json has no FP64 figure",
        "algorithmic_flops_per_particle": flops_per_particle, "mean_trials": tbar,
        "kernel_ms": k_ms, "dm_particles_per_s_kernel": n / k_s,
        "ffma_peak_tflops": ffma.value, "frac_of_fp32_pipe": ach_tflops / ffma.value,
        "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
        "exact_fp64_evaluations_per_trial": mom[15] / max(mom[1], 1.0),
        # what actually bounds the kernel: warp-instruction issue slots (4 schedulers per SM, one
        # warp-instruction per clock each); instruction count from the committed ncu capture,
        # duration and clock measured live
        "issue": {"warp_instructions": winst, "achieved": winst / k_s,
                  "peak": sms.value * 4 * (clk["sm_mhz"] or 1965.0) * 1.0e6,
                  "unit": "warp-instructions/s",
                  "frac": (winst / k_s) / (sms.value * 4 * (clk["sm_mhz"] or 1965.0) * 1.0e6)} if (prof and clk) else None,
    }
    # the gas kernel (HBM-write bound: 72 B per particle incl. the zero velocities)
    gas_s = g_ms * 1e-3
    roofline["gas_kernel"] = {"kernel": "cgb_gas_smem_kernel<double,false>", "bound": "hbm", "kernel_ms": g_ms,
                              "achieved": n * B_GAS / gas_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                              "frac": n * B_GAS / gas_s / 1e9 / hbm_peak, "particles_per_s": n / gas_s}

    # ---- CPU baseline (bounded sample, 1 core: the reference is single-threaded) ----
    n_cpu = 2_000_000
    cpu_v, cpu_dt = cpu_baseline(n_cpu, 1)
    t0c = time.perf_counter()
    from oracle import cg_oracle as O
    O.VirialEquilibrium(_ORACLE_MODEL, "dark_matter")
    fe_cpu_s = time.perf_counter() - t0c

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "dm_per_gpu": n, "gas_per_gpu": n, "num_points": pm.num_elements,
                   "rng": "Philox4x32-10 device", "l2": "outputs 1.28 GB/step/GPU exceed the 126 MB L2",
                   "parallelism": f"index-sharded x{world}",
                   "moments_reduce": {"peer": "fused into the sampler kernel over peer-mapped buffers (NVLink)",
                                      "nccl": "NCCL all_reduce per step", "local": "single GPU"}[mom_mode]},
        "fE_build_ms": min(fe_ms), "fE_build_note": "bit-equal to scipy (FITPACK-exact spline fits); host tables in, f(E) spline resident on the device (the host copy of df is made on demand)",
        "fE_build_shared_grid_ms": min(fe_shared_ms),
        "fE_build_shared_grid_note": "a further species of the same model (stars after dark matter): the Givens factorisation of the shared energy grid is re-used, bit-equal too",
        "fE_build_fast_ms": min(fe_fast_ms), "fE_build_fast_note": "PCR spline fits, f within 1e-11 of scipy",
        "fE_build_cpu_ms": 1e3 * fe_cpu_s, "check_virial_ms": min(cv_ms),
        "dm_particles_per_s": n_glob / (k_ms * 1e-3), "gas_particles_per_s": n_glob / (g_ms * 1e-3),
        "value_note": "value counts DM (pos+vel) and gas (pos, zero vel) particles of one step; the DM-only rate is dm_particles_per_s",
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "roofline": e2e_roofline,
                "what": "ClusterModel.generate_dm_particles + generate_gas_particles(host=..., wait=False) + wait() from host tables: chunked generation, every chunk's D2H copy to pinned host memory overlapped with the next chunk's kernels and with the next species' table preparation; gas velocities (identically zero) are host-side zeros"},
        "configs": configs,
        # per step: cgb_sample_kernel + cgb_hard_kernel (deferred long-tail pass) + cgb_gas_smem_kernel
        "gpu_launches": 3 * args.steps,
        "roofline": roofline,
        "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"{n_cpu} DM + {n_cpu} gas particles, oracle mode A C loop + numpy, {cpu_dt:.1f} s",
                         "host_cores_available": os.cpu_count()},
        "clocks": clk,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
