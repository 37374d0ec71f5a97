"""
Moment all-reduce fused into the sampler kernel (one process per GPU, NVLink / NVSwitch).

The path's only collective is the sum of 16 moment words over ranks.
Instead of launching an NCCL all-reduce after the kernel, every rank maps the other ranks'
moment buffers into its address space (CUDA IPC handles exchanged once through
``torch.distributed``) and the sampler's final flush adds its partial sums to ALL ranks'
buffers with system-scope atomics over NVLink (``cgb_sample_args.d_peer_moments``).  When every
rank's kernel has completed, each rank's buffer holds the global moments.

A :class:`PeerMoments` owns an arena of ``slots`` 16-word buffers.  Calls made in the same order
on every rank (collective semantics, like NCCL) take consecutive slots; nothing is zeroed while
the arena is in use, so ranks may drift apart by up to ``slots`` calls without a race.  Reading a
slot requires :meth:`fence` (local stream sync + barrier) first.  A slot is never handed out
twice: when the arena is used up :meth:`next_slot` raises, and :meth:`reset` (collective:
fence, snapshot the diagnostics already handed out, zero the own arena, barrier) makes all
slots available again.  Counts in a slot are exact; the floating-point sums are added with
system-scope atomics in arrival order, so their last bits may differ from run to run.
"""
from __future__ import annotations

import ctypes

import torch

from . import _native as N

_WORDS = N.CGB_NMOMENTS
_MAX_PEERS = 8


class _DevView:
    """Expose a raw device pointer to torch through the CUDA array interface (no copy)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


class PeerMoments:
    def __init__(self, group=None, slots: int = 1024):
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerMoments needs an initialised torch.distributed process group")
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _MAX_PEERS:
            raise ValueError(f"at most {_MAX_PEERS} ranks (one NVSwitch domain) can share moment buffers")
        self.slots = int(slots)
        self.device = N.require_cuda()
        nbytes = self.slots * _WORDS * 8
        own = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        N.check(N.lib().cgb_peer_alloc(nbytes, ctypes.byref(own), handle))
        self._own = own.value
        # exchange the 64-byte handles (device tensors for NCCL, host tensors otherwise)
        on_gpu = dist.get_backend(group) == "nccl"
        mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).clone()
        mine = mine.to(self.device) if on_gpu else mine
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine, group=group)
        self._ptrs, self._opened = [], []
        for r, h in enumerate(gathered):
            if r == self.rank:
                self._ptrs.append(self._own)
                continue
            p = ctypes.c_void_p()
            N.check(N.lib().cgb_peer_open(bytes(h.cpu().numpy().tobytes()), ctypes.byref(p)))
            self._ptrs.append(p.value)
            self._opened.append(p.value)
        self._next = 0
        self._handed_out = []
        self._view = torch.as_tensor(_DevView(self._own, self.slots * _WORDS), device=self.device)
        dist.barrier(group=group)  # every rank has mapped every buffer before anyone writes

    def next_slot(self) -> int:
        if self._next >= self.slots:
            raise RuntimeError(
                f"PeerMoments arena exhausted after {self.slots} calls: call reset() on every rank "
                f"(collective) or construct it with more slots")
        s = self._next
        self._next += 1
        return s

    def track(self, diagnostics) -> None:
        """Remember a SamplerDiagnostics that views one of our slots, so that reset() can
        give it a private copy before the slot is zeroed."""
        self._handed_out.append(diagnostics)

    def reset(self) -> None:
        """Collective.  All slots become available again; diagnostics handed out so far keep
        their values (they are detached from the arena first)."""
        import torch.distributed as dist

        self.fence()
        for d in self._handed_out:
            d.moments = d.moments.clone()
        self._handed_out = []
        self._view.zero_()
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)  # nobody adds into a slot before every arena is clean
        self._next = 0

    def attach(self, args: "N.SampleArgs", slot: int) -> None:
        """Point a sampler call at slot ``slot`` of every rank's arena."""
        args.d_moments = None
        args.n_peers = self.world
        for r, base in enumerate(self._ptrs):
            args.d_peer_moments[r] = base + slot * _WORDS * 8

    def fence(self) -> None:
        """All ranks' kernels enqueued so far have completed and their sums are visible."""
        import torch.distributed as dist

        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)

    def slot(self, slot: int) -> torch.Tensor:
        """This rank's copy of a slot (global sums once :meth:`fence` has returned)."""
        return self._view[slot * _WORDS:(slot + 1) * _WORDS]

    def close(self) -> None:
        import torch.distributed as dist

        if self._own is None:
            return
        self.fence()
        for p in self._opened:
            N.check(N.lib().cgb_peer_close(ctypes.c_void_p(p)))
        dist.barrier(group=self.group)  # nobody frees while a peer still has the buffer mapped
        self._view = None
        N.check(N.lib().cgb_peer_free(ctypes.c_void_p(self._own)))
        self._own, self._opened, self._ptrs = None, [], []
