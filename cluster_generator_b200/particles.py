"""
Particle container on device buffers.

Mirrors ``cluster_generator.particles.ClusterParticles`` (particles.py:78-120,
423-451): a dict ``(ptype, field) -> array``, ``num_particles``, ``field_names``,
``+`` concatenation and ``add_offsets``.  Arrays are ``torch`` CUDA tensors; the
``(n, 3)`` position / velocity fields are transposed views of contiguous ``(3, n)``
plane buffers -- the memory layout of ``unyt_array([x, y, z]).T`` in the reference
(virial.py:180-187).  Units live in ``field_units`` (the reference attaches them
through unyt).
"""
from __future__ import annotations

from collections import OrderedDict, defaultdict

import torch

from .utils import ensure_list

FIELD_UNITS = {
    "particle_position": "kpc",
    "particle_velocity": "kpc/Myr",
    "particle_mass": "Msun",
    "particle_potential": "kpc**2/Myr**2",
    "thermal_energy": "kpc**2/Myr**2",
    "density": "Msun/kpc**3",
}


class ClusterParticles:
    def __init__(self, particle_types, fields, diagnostics=None):
        self.particle_types = ensure_list(particle_types)
        self.fields = fields
        self.field_units = {k: FIELD_UNITS.get(k[1], "") for k in fields}
        self._update_num_particles()
        self._update_field_names()
        self.passive_scalars = []
        #: per-ptype sampler diagnostics (moments from the kernel: trials/particle ...)
        self.diagnostics = diagnostics or {}

    def __getitem__(self, key):
        return self.fields[key]

    def __setitem__(self, key, value):
        self.fields[key] = value
        self.field_units.setdefault(key, FIELD_UNITS.get(key[1], ""))

    def __contains__(self, key):
        return key in self.fields

    def keys(self):
        return self.fields.keys()

    def _update_num_particles(self):
        self.num_particles = {}
        for ptype in self.particle_types:
            self.num_particles[ptype] = self.fields[ptype, "particle_mass"].numel()

    def _update_field_names(self):
        self.field_names = defaultdict(list)
        for field in self.fields:
            self.field_names[field[0]].append(field[1])

    def __add__(self, other):
        """particles.py:109-119 (same-key fields concatenate, others union)."""
        fields = OrderedDict(self.fields)
        for field in other.fields:
            if field in fields:
                a, b = self[field], other[field]
                if a.dim() == 2:  # keep the plane-major (3, n) storage behind the (n, 3) view
                    fields[field] = torch.cat([a.T, b.T], dim=1).contiguous().T
                else:
                    fields[field] = torch.cat([a, b])
            else:
                fields[field] = other[field]
        particle_types = list(dict.fromkeys(self.particle_types + other.particle_types))
        diag = dict(self.diagnostics)
        diag.update(other.diagnostics)
        return ClusterParticles(particle_types, fields, diag)

    def add_offsets(self, r_ctr, v_ctr, ptypes=None):
        """particles.py:423-451"""
        if ptypes is None:
            ptypes = self.particle_types
        for ptype in ensure_list(ptypes):
            pos = self.fields[ptype, "particle_position"]
            vel = self.fields[ptype, "particle_velocity"]
            pos += torch.as_tensor(r_ctr, dtype=pos.dtype, device=pos.device)
            vel += torch.as_tensor(v_ctr, dtype=vel.dtype, device=vel.device)

    def drop_ptypes(self, ptypes):
        """particles.py:125-137"""
        ptypes = ensure_list(ptypes)
        for ptype in ptypes:
            self.particle_types.remove(ptype)
            for name in list(self.fields.keys()):
                if name[0] == ptype:
                    self.fields.pop(name)
        self._update_num_particles()
        self._update_field_names()

    def to_numpy(self):
        """Host copy: dict (ptype, field) -> ndarray."""
        return {k: v.detach().cpu().numpy() for k, v in self.fields.items()}
