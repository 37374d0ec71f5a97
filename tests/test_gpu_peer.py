"""The moment all-reduce fused into the sampler (cgb_sample_args.d_peer_moments, peer.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from cluster_generator_b200 import _native as N
from tests.configs import product_model

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(tabs, lo, hi, n, seed, moments=None, peers=None):
    dev = torch.device("cuda")
    pos = torch.empty((3, hi - lo), dtype=torch.float64, device=dev)
    vel = torch.empty_like(pos)
    a = N.SampleArgs()
    a.d_workspace = tabs.workspace.data_ptr()
    a.id_begin, a.id_end, a.ld = lo, hi, hi - lo
    a.particle_mass, a.seed, a.out_dtype = tabs.mtot / n, seed, N.CGB_DTYPE_F64
    a.d_pos, a.d_vel = pos.data_ptr(), vel.data_ptr()
    if peers is not None:
        a.n_peers = len(peers)
        for r, t in enumerate(peers):
            a.d_peer_moments[r] = t.data_ptr()
    elif moments is not None:
        a.d_moments = moments.data_ptr()
    N.check(N.lib().cgb_sample_particles(N.ctypes.byref(a), N.ptr(tabs.scratch), tabs.scratch.numel(), N.stream_ptr()))
    return pos, vel


def test_kernel_adds_its_moments_to_every_peer_buffer():
    """Two shards on one GPU, each flushing into both 'rank' buffers: both end up with the
    moments of the unsharded run (counts and maxima exactly, fp sums to rounding)."""
    pm = product_model("hernquist")          # heavy tail: also exercises the deferred grid pass
    tabs = pm.dm_virial.sampler_tables()
    n, seed = 300_001, 77
    dev = torch.device("cuda")
    ref = torch.zeros(N.CGB_NMOMENTS, dtype=torch.float64, device=dev)
    p_ref, v_ref = _run(tabs, 0, n, n, seed, moments=ref)
    bufs = [torch.zeros(N.CGB_NMOMENTS, dtype=torch.float64, device=dev) for _ in range(2)]
    cut = 123_457
    p0, v0 = _run(tabs, 0, cut, n, seed, peers=bufs)
    p1, v1 = _run(tabs, cut, n, n, seed, peers=bufs)
    torch.cuda.synchronize()
    assert torch.equal(torch.cat([p0, p1], 1), p_ref) and torch.equal(torch.cat([v0, v1], 1), v_ref)
    r, b0, b1 = ref.cpu().numpy(), bufs[0].cpu().numpy(), bufs[1].cpu().numpy()
    np.testing.assert_allclose(b0, b1, rtol=1e-12, atol=1e-9 * n)   # same addends, different atomic order
    for k in (0, 1, 2, 11, 13):              # particles, trials, max trials, envelopes, failures
        assert b0[k] == b1[k] == r[k], (k, b0[k], b1[k], r[k])
    np.testing.assert_allclose(b0[[3, 4, 14]], r[[3, 4, 14]], rtol=1e-12)
    np.testing.assert_allclose(b0[5:11], r[5:11], rtol=1e-9, atol=1e-6 * n)
    a = N.SampleArgs()
    a.d_workspace = tabs.workspace.data_ptr()
    a.id_begin, a.id_end, a.ld = 0, 10, 10
    a.d_pos, a.d_vel = p_ref.data_ptr(), v_ref.data_ptr()
    a.n_peers = 9
    with pytest.raises(ValueError, match="n_peers"):
        N.check(N.lib().cgb_sample_particles(N.ctypes.byref(a), N.ptr(tabs.scratch), tabs.scratch.numel(),
                                             N.stream_ptr()))


_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)) % torch.cuda.device_count())
dist.init_process_group("gloo" if torch.cuda.device_count() < world else "nccl")
import cluster_generator_b200 as cg
from tests.configs import product_model
pm = product_model("vikhlinin")
peer = cg.PeerMoments(slots=4)
n = 200_001
out = []
for call in range(4):
    p = pm.generate_dm_particles(n, prng=call, shard=(rank, world), peer_moments=peer)
    out.append(p.diagnostics["dm"])
try:                                         # a used slot is never handed out again
    pm.generate_dm_particles(n, prng=4, shard=(rank, world), peer_moments=peer)
    raise SystemExit("expected the exhausted arena to raise")
except RuntimeError as e:
    assert "exhausted" in str(e)
peer.fence()
assert torch.equal(out[1].moments, peer.slot(1))
peer.reset()                                 # collective: earlier diagnostics keep their values
for call in range(4, 6):
    p = pm.generate_dm_particles(n, prng=call % 4, shard=(rank, world), peer_moments=peer)
    out.append(p.diagnostics["dm"])
peer.fence()
single = pm.generate_dm_particles(n, prng=1)
want = single.diagnostics["dm"].as_dict()
for k in (1, 5):                             # calls 1 and 5 drew the same particles: per-call sums
    got = out[k].all_reduce().as_dict()      # all_reduce is a no-op: already global
    assert got["particles"] == n and got["sum_trials"] == want["sum_trials"], (k, got, want)
    assert got["max_trials"] == want["max_trials"]
    assert abs(got["sum_v2"] - want["sum_v2"]) <= 1e-11 * got["sum_v2"]
assert torch.equal(out[5].moments, peer.slot(1))
# virial-check histograms: integer counts, summed over ranks with the process group's all_reduce
pv = pm.dm_virial
mine = pm.generate_dm_particles(n, prng=1, shard=(rank, world))
h = pv.particle_histograms(mine, nbins=32, all_reduce=True)
hs = pv.particle_histograms(single, nbins=32)
assert h["n"] == n and all(torch.equal(h[k], hs[k]) for k in ("r", "v", "E")), (h, hs)
# one Gadget file per rank (counts exchanged with all_gather_object)
from cluster_generator_b200.h5io import H5Reader
mine.add_offsets([1.0e4] * 3, [0.0] * 3)
mine.write_to_gadget_file(os.path.join(sys.argv[2], "ic.h5"), 2.0e4, overwrite=True, shard=(rank, world))
dist.barrier()
seen, ids = 0, []
for r in range(world):
    with H5Reader(os.path.join(sys.argv[2], f"ic.{r}.h5")) as rd:
        hd = rd.attrs("/Header")
        assert hd["NumFilesPerSnapshot"] == world
        seen += int(hd["NumPart_ThisFile"][1])
        ids.append(rd.read("/PartType1/ParticleIDs"))
assert seen == int(hd["NumPart_Total"][1]) <= n
assert np.array_equal(np.concatenate(ids), np.arange(1, seen + 1, dtype=np.uint32))
peer.close()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_peer_moments_across_two_processes(tmp_path):
    """Two ranks (two GPUs when present, else both on cuda:0 with gloo for the handle exchange):
    IPC-mapped buffers, fused reduction, arena exhaustion + reset, close."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(tmp_path)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, cwd=ROOT) for r in range(2)]
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=600)[0])
        except subprocess.TimeoutExpired:
            p.kill()
            outs.append(p.communicate()[0])
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"rank {r} ok" in o, o[-3000:]
