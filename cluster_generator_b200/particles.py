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

    def make_radial_cut(self, r_max, center=None, ptypes=None):
        """Remove all particles outside ``r_max`` of ``center`` (particles.py:139-167)."""
        rm2 = r_max * r_max
        if ptypes is None:
            ptypes = self.particle_types
        for part in ensure_list(ptypes):
            pos = self.fields[part, "particle_position"]
            c = torch.as_tensor([0.0] * 3 if center is None else center, dtype=pos.dtype, device=pos.device)
            cidx = ((pos - c) ** 2).sum(dim=1) <= rm2
            for field in self.field_names[part]:
                f = self.fields[part, field]
                # keep the plane-major storage of the (n, 3) fields
                self.fields[part, field] = f.T[:, cidx].contiguous().T if f.dim() == 2 else f[cidx]
        self._update_num_particles()

    def add_black_hole(self, bh_mass, pos=None, vel=None, use_pot_min=False):
        """Add a black hole particle (particles.py:169-225)."""
        ref = self.fields[self.particle_types[0], "particle_mass"]
        kw = dict(dtype=ref.dtype, device=ref.device)
        mass = torch.tensor([bh_mass], **kw)
        if use_pot_min:
            if ("dm", "potential_energy") not in self.fields:
                raise KeyError("('dm', 'potential_energy') is not available!")
            idx = int(torch.argmin(self.fields["dm", "potential_energy"]))
            pos = self.fields["dm", "particle_position"][idx].reshape(1, 3).clone()
            vel = self.fields["dm", "particle_velocity"][idx].reshape(1, 3).clone()
        else:
            pos = torch.zeros((1, 3), **kw) if pos is None else torch.as_tensor(pos, **kw).reshape(1, 3)
            vel = torch.zeros((1, 3), **kw) if vel is None else torch.as_tensor(vel, **kw).reshape(1, 3)
        if "black_hole" not in self.particle_types:
            self.particle_types.append("black_hole")
            self.fields["black_hole", "particle_position"] = pos
            self.fields["black_hole", "particle_velocity"] = vel
            self.fields["black_hole", "particle_mass"] = mass
        else:
            for name, new in (("particle_position", pos), ("particle_velocity", vel), ("particle_mass", mass)):
                self.fields["black_hole", name] = torch.cat([self.fields["black_hole", name], new], dim=0)
        for k in (("black_hole", "particle_position"), ("black_hole", "particle_velocity"),
                  ("black_hole", "particle_mass")):
            self.field_units.setdefault(k, FIELD_UNITS.get(k[1], ""))
        self._update_num_particles()
        self._update_field_names()

    # -- persistence ------------------------------------------------------------
    # The reference writes HDF5 (group per ptype, dataset per field, "units" attribute,
    # particles.py:337-363).  There is no HDF5 library in this environment, so the same
    # layout is stored in a NumPy .npz archive: key "<ptype>/<field>", plus "<ptype>/<field>.units".
    def write_particles(self, output_filename, overwrite=False):
        """Write the particles to ``output_filename`` (.npz; see the note above)."""
        import os

        import numpy as np

        output_filename = str(output_filename)
        if output_filename.endswith((".h5", ".hdf5")):
            raise NotImplementedError("HDF5 output needs an HDF5 library, which is not available here; "
                                      "use a .npz file name")
        if not output_filename.endswith(".npz"):
            output_filename += ".npz"
        if os.path.exists(output_filename) and not overwrite:
            raise IOError(f"Cannot create {output_filename}. It exists and overwrite=False.")
        data = {}
        for (ptype, name), t in self.fields.items():
            data[f"{ptype}/{name}"] = t.detach().cpu().numpy()
            data[f"{ptype}/{name}.units"] = np.array(self.field_units.get((ptype, name), ""))
        np.savez(output_filename, **data)
        return output_filename

    @classmethod
    def from_file(cls, filename, ptypes=None, device=None):
        """Read particles written by :meth:`write_particles` (particles.py:235-274)."""
        import numpy as np

        device = device or ("cuda" if torch.cuda.is_available() else "cpu")
        z = np.load(str(filename))
        keys = [k for k in z.files if not k.endswith(".units")]
        all_ptypes = list(dict.fromkeys(k.split("/")[0] for k in keys))
        ptypes = all_ptypes if ptypes is None else ensure_list(ptypes)
        fields = OrderedDict()
        units = {}
        for k in keys:
            ptype, name = k.split("/", 1)
            if ptype not in ptypes:
                continue
            a = torch.from_numpy(z[k]).to(device)
            fields[ptype, name] = a.T.contiguous().T if a.dim() == 2 else a   # plane-major storage
            units[ptype, name] = str(z[k + ".units"]) if k + ".units" in z.files else ""
        out = cls(ptypes, fields)
        out.field_units.update(units)
        return out

    def to_numpy(self):
        """Host copy: dict (ptype, field) -> ndarray."""
        return {k: v.detach().cpu().numpy() for k, v in self.fields.items()}


# --------------------------------------------------------------------------- #
# multi-halo combination / gas re-sampling (particles.py:587-795)
# --------------------------------------------------------------------------- #
def _planes(t: torch.Tensor) -> torch.Tensor:
    """(n, 3) field -> its contiguous (3, n) plane storage (copying only if it has none)."""
    p = t.T
    return p if p.is_contiguous() else p.contiguous()


def _sample_clusters(particles, hses, center, velocity, radii=None, resample=False, passive_scalars=None):
    """Re-sample the gas fields from the overlap of up to three radial profiles
    (particles.py:587-639): total density, mass-weighted specific thermal energy and
    mass-weighted bulk velocity at every gas particle.  Runs in ``cgb_resample_gas``."""
    import numpy as np

    from . import _native as N
    from .virial import DeviceSpline

    if passive_scalars:
        raise NotImplementedError("passive scalars are outside the particle-sampling path")
    num_halos = len(hses)
    if not 1 <= num_halos <= 3:
        raise ValueError("between one and three halos are supported")
    n = particles.num_particles["gas"]
    pos = _planes(particles["gas", "particle_position"])
    vel = _planes(particles["gas", "particle_velocity"])
    dens, therm = particles["gas", "density"], particles["gas", "thermal_energy"]
    mass = particles["gas", "particle_mass"]
    dtype = pos.dtype
    code = N.CGB_DTYPE_F64 if dtype == torch.float64 else N.CGB_DTYPE_F32
    for t in (vel, dens, therm, mass):
        if t.dtype != dtype or not t.is_contiguous():
            raise ValueError("gas fields must share one dtype and be contiguous")
    L = N.lib()
    keep = []
    halos = (N.HaloTables * num_halos)()
    for i, hse in enumerate(hses):
        c = np.asarray(center[i], dtype=np.float64).reshape(3)
        v = np.asarray(velocity[i], dtype=np.float64).reshape(3)
        halos[i].center[:] = c.tolist()
        halos[i].velocity[:] = v.tolist()
        halos[i].radius_cut = float(radii[i]) if radii is not None else -1.0
        if "density" not in hse:
            continue  # halo without gas (particles.py:611-613)
        rr = np.asarray(hse["radius"], dtype=np.float64)
        rho = np.asarray(hse["density"], dtype=np.float64)
        e_arr = 1.5 * np.asarray(hse["pressure"], dtype=np.float64) / rho
        halos[i].n = rr.size
        for name, y in (("d_dens_pp", rho), ("d_ener_pp", e_arr)):
            s = DeviceSpline.fit(rr, y)
            nb = L.cgb_ppform_workspace_bytes(s.n)
            ws = N.workspace(nb)
            N.check(L.cgb_ppform_prepare(N.ptr(s.t), N.ptr(s.c), s.n, N.ptr(ws), nb, N.stream_ptr()))
            keep.append((s, ws))
            setattr(halos[i], name, ws.data_ptr())
    N.check(L.cgb_resample_gas(halos, num_halos, N.ptr(pos), n, n, code, N.ptr(dens), N.ptr(therm),
                               N.ptr(vel), N.ptr(mass) if resample else None,
                               1 if radii is not None else 0, N.stream_ptr()))
    particles.fields["gas", "particle_velocity"] = vel.T
    del keep
    return particles


def _combine(parts, hses, centers, velocities):
    """combine_two_clusters / combine_three_clusters (particles.py:642-715): gas is offset in
    position only, the collisionless species in position and velocity; the gas fields are
    then re-sampled from the overlapping profiles."""
    for p, c, v in zip(parts, centers, velocities):
        if "gas" in p.particle_types:
            p.add_offsets(c, [0.0] * 3, ptypes=["gas"])
        others = [t for t in p.particle_types if t != "gas"]
        p.add_offsets(c, v, ptypes=others)
    particles = parts[0]
    for p in parts[1:]:
        particles = particles + p
    if "gas" in particles.particle_types:
        particles = _sample_clusters(particles, hses, centers, velocities)
    return particles


def combine_two_clusters(particles1, particles2, hse1, hse2, center1, center2, velocity1, velocity2):
    """particles.py:642-666"""
    return _combine([particles1, particles2], [hse1, hse2], [center1, center2], [velocity1, velocity2])


def combine_three_clusters(particles1, particles2, particles3, hse1, hse2, hse3, center1, center2,
                           center3, velocity1, velocity2, velocity3):
    """particles.py:669-715"""
    return _combine([particles1, particles2, particles3], [hse1, hse2, hse3],
                    [center1, center2, center3], [velocity1, velocity2, velocity3])


def resample_one_cluster(particles, hse, center, velocity):
    """particles.py:718-744"""
    if "gas" not in particles.particle_types:
        return particles
    return _sample_clusters(particles, [hse], [center], [velocity], resample=True)


def resample_two_clusters(particles, hse1, hse2, center1, center2, velocity1, velocity2, radii):
    """particles.py:747-768"""
    return _sample_clusters(particles, [hse1, hse2], [center1, center2], [velocity1, velocity2],
                            radii=radii, resample=True)


def resample_three_clusters(particles, hse1, hse2, hse3, center1, center2, center3, velocity1,
                            velocity2, velocity3, radii):
    """particles.py:771-795"""
    return _sample_clusters(particles, [hse1, hse2, hse3], [center1, center2, center3],
                            [velocity1, velocity2, velocity3], radii=radii, resample=True)
