"""
3-D gridded vector field -> particle fields ("next" row 4 of SURVEY.md section 8(f)).

Only the particle-facing part of ``cluster_generator.fields.ClusterField`` is here:
``map_field_to_particles`` (fields.py:308-338).  Generating the fields themselves (FFT
Gaussian random fields, divergence cleaning) is a different algorithm family and out of
scope; any gridded field can be wrapped in :class:`GridField`.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as N
from .particles import _planes


class GridField:
    def __init__(self, name, x, y, z, components, units=""):
        """``components``: three (nx, ny, nz) arrays (x, y, z components) on the grid
        ``x`` x ``y`` x ``z`` (1-D ascending coordinates, kpc)."""
        self._name, self.units = name, units
        self.x, self.y, self.z = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z))
        comps = [np.ascontiguousarray(c, dtype=np.float64) for c in components]
        shape = (self.x.size, self.y.size, self.z.size)
        if len(comps) != 3 or any(c.shape != shape for c in comps):
            raise ValueError(f"expected three components of shape {shape}")
        if min(shape) < 2:
            raise ValueError("the grid needs at least two points per axis")
        for g in (self.x, self.y, self.z):
            if not np.all(np.diff(g) > 0.0):
                raise ValueError("The points in dimension must be strictly ascending")
        self._dev = None
        self._host = comps

    def _device(self):
        if self._dev is None:
            self._dev = ([N.dev_f64(g) for g in (self.x, self.y, self.z)], [N.dev_f64(c) for c in self._host])
        return self._dev

    def map_field_to_particles(self, cluster_particles, ptype="gas", units=None):
        """Tri-linear interpolation of the field at the particle positions; adds the
        ``(ptype, name)`` particle field of shape (n, 3) (fields.py:308-338)."""
        (gx, gy, gz), (fx, fy, fz) = self._device()
        pos = _planes(cluster_particles[ptype, "particle_position"])
        n = pos.shape[1]
        code = N.CGB_DTYPE_F64 if pos.dtype == torch.float64 else N.CGB_DTYPE_F32
        out = torch.empty((3, n), dtype=pos.dtype, device=pos.device)
        N.check(N.lib().cgb_map_field(N.ptr(gx), N.ptr(gy), N.ptr(gz), self.x.size, self.y.size, self.z.size,
                                      N.ptr(fx), N.ptr(fy), N.ptr(fz), N.ptr(pos), n, n, code, N.ptr(out),
                                      N.stream_ptr()))
        cluster_particles[ptype, self._name] = out.T
        cluster_particles.field_units[ptype, self._name] = units or self.units
        cluster_particles._update_field_names()
        return cluster_particles
