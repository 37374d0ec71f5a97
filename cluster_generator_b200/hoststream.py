"""
Host-resident output of the particle generators: ``generate_*_particles(..., host=True)``.

The reference returns host arrays.  Here the kernels write device buffers; when the caller
wants the particles on the host (to hand them to a CPU code, or because n exceeds HBM), the
index range is generated in id-range chunks into two device staging slots and every chunk
is copied to pinned host memory on a second stream, so chunk i's device->host copy overlaps
chunk i+1's kernels.  One ``cudaMemcpyAsync`` per plane and chunk (both sides contiguous),
nothing else on the copy stream.  Every particle is a pure function of (tables, seed, id),
so the result is bit-identical to the one-shot device path.

``generate_*_particles(..., host=..., wait=False)`` returns while the last copies are still in
flight: the caller's next calls (another species' table preparation and first chunks) overlap
them.  The returned ``ClusterParticles`` keeps the streamer -- and with it the staging slots the
copies read from -- until ``wait()``, which any access to its fields calls.

``bind_host_to_gpu()`` pins the calling process to the CPU cores NVML reports as local to
its GPU before any pinned memory is allocated: pinned pages are then taken from that
socket's DRAM (first touch) and the copies of the 8 ranks of a box do not funnel through
the inter-socket link.
"""
from __future__ import annotations

import os
from collections import OrderedDict

import torch

#: particles per chunk: 8 MiB per fp64 plane -- long enough for PCIe to reach its plateau,
#: short enough that the staging slots (2 x 9 planes) stay small
DEFAULT_CHUNK = 1 << 20


def bind_host_to_gpu(device_index=None) -> list[int]:
    """Restrict this process to the CPUs close to its GPU (NVML's ideal CPU affinity).
    Returns the CPU list, or [] when NVML cannot tell (nothing is changed then)."""
    try:
        import pynvml as nv

        nv.nvmlInit()
        idx = torch.cuda.current_device() if device_index is None else int(device_index)
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            parts = vis.split(",")
            if idx < len(parts) and parts[idx].strip().isdigit():
                idx = int(parts[idx])
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = nv.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return []


class HostBuffers:
    """Pinned host arrays of one species: name -> (rows, n) or (n,) tensor."""

    def __init__(self, layout, n, dtype):
        self.n, self.dtype = int(n), dtype
        self.arrays = OrderedDict()
        for name, rows in layout.items():
            shape = (rows, self.n) if rows > 1 else (self.n,)
            self.arrays[name] = torch.empty(shape, dtype=dtype, pin_memory=True)

    def matches(self, layout, n, dtype):
        return (self.n == int(n) and self.dtype == dtype and list(self.arrays) == list(layout)
                and all((a.shape[0] if a.dim() == 2 else 1) == r for a, r in zip(self.arrays.values(), layout.values())))


class HostStreamer:
    """Double-buffered staging + copy stream.

    ``launch(stage, a, b)`` must enqueue, on the current stream, the kernels that fill the
    device tensors ``stage[name][..., :b-a]`` with local particles [a, b)."""

    def __init__(self, layout, n, dtype, device, chunk=None, out: HostBuffers | None = None):
        self.layout, self.n, self.dtype, self.dev = OrderedDict(layout), int(n), dtype, device
        self.chunk = int(min(max(int(chunk or DEFAULT_CHUNK), 2), max(self.n, 2)))
        self.chunk += self.chunk & 1  # even plane strides keep the kernels' 16-byte stores
        self.host = out if (out is not None and out.matches(self.layout, n, dtype)) else HostBuffers(self.layout, n, dtype)
        self.stage = [OrderedDict((name, torch.empty((rows, self.chunk) if rows > 1 else (self.chunk,),
                                                     dtype=dtype, device=device))
                                  for name, rows in self.layout.items()) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=device)
        self.filled = [torch.cuda.Event(), torch.cuda.Event()]
        self.drained = [torch.cuda.Event(), torch.cuda.Event()]
        self.bytes = 0

    def run(self, launch, wait=True):
        """``wait=False``: do not make the compute stream wait for the copies -- the caller keeps
        this streamer alive and calls :meth:`synchronize` (``ClusterParticles.wait``) before the
        host arrays are read; kernels enqueued meanwhile (the next species' tables and chunks)
        overlap the tail of these copies."""
        compute = torch.cuda.current_stream(self.dev)
        for c, a in enumerate(range(0, self.n, self.chunk)):
            b = min(a + self.chunk, self.n)
            s = c & 1
            if c >= 2:
                compute.wait_event(self.drained[s])   # the slot's previous copy has left the device
            launch(self.stage[s], a, b)
            self.filled[s].record(compute)
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self.filled[s])
                for name, rows in self.layout.items():
                    src, dst = self.stage[s][name], self.host.arrays[name]
                    if rows > 1:
                        for k in range(rows):
                            dst[k, a:b].copy_(src[k, : b - a], non_blocking=True)
                    else:
                        dst[a:b].copy_(src[: b - a], non_blocking=True)
                    self.bytes += rows * (b - a) * src.element_size()
                self.drained[s].record(self.copy_stream)
        if wait:
            # the caller's stream sees the host buffers complete
            compute.wait_stream(self.copy_stream)
        return self.host

    def synchronize(self):
        """Block the host until every copy issued by :meth:`run` has landed."""
        self.copy_stream.synchronize()
