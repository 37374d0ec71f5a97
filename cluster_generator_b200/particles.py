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

from .utils import _M_PER_KPC, _SEC_PER_MYR, ensure_list

FIELD_UNITS = {
    "particle_position": "kpc",
    "particle_velocity": "kpc/Myr",
    "particle_mass": "Msun",
    "particle_potential": "kpc**2/Myr**2",
    "thermal_energy": "kpc**2/Myr**2",
    "density": "Msun/kpc**3",
}

# Gadget-HDF5 naming (particles.py:13-75)
gadget_fields = {
    "dm": ["Coordinates", "Velocities", "Masses", "ParticleIDs", "Potential"],
    "gas": ["Coordinates", "Velocities", "Masses", "ParticleIDs", "InternalEnergy", "MagneticField",
            "Density", "Potential"],
    "star": ["Coordinates", "Velocities", "Masses", "ParticleIDs", "Potential"],
    "black_hole": ["Coordinates", "Velocities", "Masses", "ParticleIDs"],
    "tracer": ["Coordinates"],
}
gadget_field_map = {
    "Coordinates": "particle_position", "Velocities": "particle_velocity", "Masses": "particle_mass",
    "Density": "density", "Potential": "potential_energy", "InternalEnergy": "thermal_energy",
    "MagneticField": "magnetic_field", "Metallicity": "metallicity", "GFM_Metallicity": "metallicity",
}
gadget_field_units = {
    "Coordinates": "kpc", "Velocities": "km/s", "Masses": "1e10*Msun", "Density": "1e10*Msun/kpc**3",
    "InternalEnergy": "km**2/s**2", "Potential": "km**2/s**2", "PassiveScalars": "",
    "MagneticField": "1e5*sqrt(Msun)*km/s/(kpc**1.5)", "Metallicity": "", "GFM_Metallicity": "",
}
code_fields = {"arepo": {"gas": ["PassiveScalars", "GFM_Metallicity"]}}
ptype_map = OrderedDict([("PartType0", "gas"), ("PartType1", "dm"), ("PartType2", "tracer"),
                         ("PartType4", "star"), ("PartType5", "black_hole")])
rptype_map = OrderedDict([(v, k) for k, v in ptype_map.items()])

#: km/s per kpc/Myr, from the same kpc and Julian-year definitions as utils.G
KMS_PER_KPCMYR = _M_PER_KPC / 1.0e3 / _SEC_PER_MYR

# unit string -> factor to the galactic base (kpc, Msun, Myr); the strings this package and the
# Gadget layout use (the reference resolves arbitrary strings through unyt).  sqrt(Msun)*kpc**-0.5/Myr
# is the galactic base of the magnetic-field unit.
_GALACTIC = {
    "": 1.0, "dimensionless": 1.0, "kpc": 1.0, "kpc/Myr": 1.0, "Msun": 1.0, "kpc**2/Myr**2": 1.0,
    "Msun/kpc**3": 1.0, "Myr": 1.0, "sqrt(Msun)/(sqrt(kpc)*Myr)": 1.0,
    "km/s": 1.0 / KMS_PER_KPCMYR, "1e10*Msun": 1.0e10, "1e10*Msun/kpc**3": 1.0e10,
    "km**2/s**2": 1.0 / KMS_PER_KPCMYR**2,
    "1e5*sqrt(Msun)*km/s/(kpc**1.5)": 1.0e5 / KMS_PER_KPCMYR,
}
_GALACTIC_NAME = {
    "km/s": "kpc/Myr", "1e10*Msun": "Msun", "1e10*Msun/kpc**3": "Msun/kpc**3", "km**2/s**2": "kpc**2/Myr**2",
    "1e5*sqrt(Msun)*km/s/(kpc**1.5)": "sqrt(Msun)/(sqrt(kpc)*Myr)", "dimensionless": "",
}


def _to_galactic(unit: str) -> float:
    """Multiplier taking values in ``unit`` to kpc/Msun/Myr (``in_base("galactic")``)."""
    try:
        return _GALACTIC[unit.strip()]
    except KeyError:
        raise ValueError(f"unit {unit!r} is not one this package can convert "
                         f"(known: {sorted(_GALACTIC)})") from None


def _galactic_units(unit: str) -> str:
    return _GALACTIC_NAME.get(unit.strip(), unit.strip())


# galactic base -> Gadget unit, as direct factors (not reciprocals of _GALACTIC, so that a field
# already in galactic units is scaled by exactly KMS_PER_KPCMYR, 1e-10, ...)
_FROM_GALACTIC = {
    "": 1.0, "kpc": 1.0, "km/s": KMS_PER_KPCMYR, "1e10*Msun": 1.0e-10, "1e10*Msun/kpc**3": 1.0e-10,
    "km**2/s**2": KMS_PER_KPCMYR**2, "1e5*sqrt(Msun)*km/s/(kpc**1.5)": KMS_PER_KPCMYR * 1.0e-5,
}


def gadget_unit_scale(gadget_field: str, units: str) -> float:
    """Multiplier from a field held in ``units`` to the Gadget unit of ``gadget_field``
    (``fd.to(units)`` at particles.py:470-471)."""
    return _to_galactic(units) * _FROM_GALACTIC[gadget_field_units[gadget_field]]


def _to_device_planes(a, device) -> torch.Tensor:
    """Host array -> device tensor; (n, k) arrays get plane-major (k, n) storage."""
    import numpy as np

    t = torch.from_numpy(np.ascontiguousarray(a)).to(device)
    return t.T.contiguous().T if t.dim() == 2 else t



class ClusterParticles:
    def __init__(self, particle_types, fields, diagnostics=None):
        #: host streamers whose device->host copies may still be in flight (``host=..., wait=False``)
        self._pending = []
        self.particle_types = ensure_list(particle_types)
        self.fields = fields
        self.field_units = {k: FIELD_UNITS.get(k[1], "") for k in fields}
        self._update_num_particles()
        self._update_field_names()
        self.passive_scalars = []
        #: per-ptype sampler diagnostics (moments from the kernel: trials/particle ...)
        self.diagnostics = diagnostics or {}

    def wait(self):
        """Block until the host arrays of a ``generate_*_particles(host=..., wait=False)`` call are
        complete.  Any access to the fields (``particles[key]``, ``particles.fields``, every
        method that reads them) does this by itself."""
        for hs in self._pending:
            hs.synchronize()
        self._pending = []
        return self

    @property
    def fields(self):
        if self._pending:
            self.wait()
        return self._fields

    @fields.setter
    def fields(self, value):
        self._fields = value

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

    @classmethod
    def from_fields(cls, fields):
        """particles.py:227-233: particle types in order of first appearance in the keys."""
        particle_types = []
        for key in fields:
            if key[0] not in particle_types:
                particle_types.append(key[0])
        return cls(particle_types, fields)

    def set_field(self, ptype, name, value, units=None, add=False, passive_scalar=False):
        """Add or update a particle field (particles.py:368-421).  ``value``: tensor / array with
        one entry per particle of ``ptype``; ``units``: the units the values are in (recorded in
        ``field_units``; the reference converts through unyt)."""
        from .utils import mylog

        ref = self.fields[ptype, "particle_mass"]
        value = torch.as_tensor(value, device=ref.device)
        num_particles = self.num_particles[ptype]
        if value.shape[0] != num_particles:
            raise ValueError(f"The length of the array needs to be {num_particles} particles!")
        exists = (ptype, name) in self.fields
        if exists:
            if add:
                self.fields[ptype, name] += value
            else:
                mylog.warning("Overwriting field (%s, %s).", ptype, name)
                self.fields[ptype, name] = value
        else:
            if add:
                raise RuntimeError(f"Field ({ptype}, {name}) does not exist and add=True!")
            self.fields[ptype, name] = value
            if passive_scalar and ptype == "gas":
                self.passive_scalars.append(name)
        if units is not None or not exists:
            self.field_units[ptype, name] = units if units is not None else FIELD_UNITS.get(name, "")
        self._update_field_names()

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
    # Same on-disk layouts as the reference (HDF5 via h5py there, via the native cgb_h5_*
    # writer/reader of libcgb200 here; see include/cgb200.h and h5io.py).  Device fields
    # stream to the file through pinned staging slots; no whole-field host copy is made.
    def _clip_to_box(self, ptype, box_size):
        """particles.py:105-107: mask of particles inside [0, box_size]^3."""
        pos = self.fields[ptype, "particle_position"]
        return ~((pos < 0.0).any(dim=1) | (pos > box_size).any(dim=1))

    def write_particles(self, output_filename, overwrite=False):
        """Write the particles to an HDF5 file: one group per particle type, one dataset per
        field, ``units`` string attribute on each (particles.py:337-363 + unyt ``write_hdf5``).
        A file name ending in ``.npz`` selects a NumPy archive with the same keys instead."""
        import os

        output_filename = str(output_filename)
        if os.path.exists(output_filename) and not overwrite:
            raise IOError(f"Cannot create {output_filename}. It exists and overwrite=False.")
        if output_filename.endswith(".npz"):
            return self._write_npz(output_filename)
        from .h5io import H5Writer

        with H5Writer(output_filename, overwrite=True) as f:
            for ptype in self.particle_types:
                f.create_group(f"/{ptype}")
            for (ptype, name), t in self.fields.items():
                path = f"/{ptype}/{name}"
                if t.is_cuda and t.dtype in (torch.float32, torch.float64):
                    f.write_device(path, t, out_dtype=str(t.dtype).split(".")[1])
                else:
                    f.write(path, t.detach().cpu().numpy())
                if name != "particle_index":
                    f.set_attr(path, "units", self.field_units.get((ptype, name), ""))
        return output_filename

    def write_particles_to_h5(self, output_filename, overwrite=False):
        return self.write_particles(output_filename, overwrite=overwrite)

    def _write_npz(self, output_filename):
        import numpy as np

        data = {}
        for (ptype, name), t in self.fields.items():
            data[f"{ptype}/{name}"] = t.detach().cpu().numpy()
            data[f"{ptype}/{name}.units"] = np.array(self.field_units.get((ptype, name), ""))
        np.savez(output_filename, **data)
        return output_filename

    @classmethod
    def from_file(cls, filename, ptypes=None, device=None):
        """Read particles written by :meth:`write_particles` or by the reference's
        ``write_particles`` (particles.py:235-274).  Floating-point fields come back as
        float64 in galactic units (kpc, Msun, Myr), as there."""
        import numpy as np

        device = device or ("cuda" if torch.cuda.is_available() else "cpu")
        filename = str(filename)
        with open(filename, "rb") as fh:
            magic = fh.read(4)
        if magic[:2] == b"PK":
            return cls._from_npz(filename, ptypes, device)
        from .h5io import H5Reader

        fields = OrderedDict()
        units = {}
        with H5Reader(filename) as f:
            all_ptypes = [k for k in f.keys("/") if f.is_group("/", k)]
            ptypes = all_ptypes if ptypes is None else ensure_list(ptypes)
            for ptype in ptypes:
                for name in f.keys(f"/{ptype}"):
                    path = f"/{ptype}/{name}"
                    if name == "particle_index":
                        fields[ptype, name] = _to_device_planes(f.read(path), device)
                        continue
                    unit = f.attrs(path).get("units", FIELD_UNITS.get(name, ""))
                    scale = _to_galactic(unit)
                    units[ptype, name] = _galactic_units(unit)
                    dt, shape, _ = f.info(path)
                    if torch.device(device).type == "cuda" and dt in (np.float32, np.float64) and len(shape) in (1, 2) \
                            and (len(shape) == 1 or shape[1] <= 4):
                        fields[ptype, name] = f.read_device(path, device, torch.float64, scale)
                        continue
                    a = f.read(path)
                    if a.dtype != np.float64 or scale != 1.0:  # (no extra host pass otherwise)
                        a = a.astype(np.float64) * scale
                    fields[ptype, name] = _to_device_planes(a, device)
        out = cls(ptypes, fields)
        out.field_units.update(units)
        return out

    @classmethod
    def from_h5_file(cls, filename, ptypes=None, device=None):
        return cls.from_file(filename, ptypes=ptypes, device=device)

    @classmethod
    def _from_npz(cls, filename, ptypes, device):
        import numpy as np

        z = np.load(filename)
        keys = [k for k in z.files if not k.endswith(".units")]
        all_ptypes = list(dict.fromkeys(k.split("/")[0] for k in keys))
        ptypes = all_ptypes if ptypes is None else ensure_list(ptypes)
        fields = OrderedDict()
        units = {}
        for k in keys:
            ptype, name = k.split("/", 1)
            if ptype not in ptypes:
                continue
            fields[ptype, name] = _to_device_planes(z[k], device)
            units[ptype, name] = str(z[k + ".units"]) if k + ".units" in z.files else ""
        out = cls(ptypes, fields)
        out.field_units.update(units)
        return out

    def write_to_gadget_file(self, ic_filename, box_size, dtype="float32", overwrite=False, code=None, *,
                             shard=None, shard_counts=None):
        """Write Gadget-HDF5 initial conditions (particles.py:453-552): ``PartType{0,1,2,4,5}``
        groups with Coordinates [kpc], Velocities [km/s], Masses [1e10 Msun], Density,
        InternalEnergy / Potential [km**2/s**2], sequential uint32 ParticleIDs, particles
        outside ``[0, box_size]^3`` dropped, and the ``Header`` attributes Gadget-family
        codes read.  ``code="arepo"`` adds the ``Config`` group.

        ``shard=(rank, world)`` (one process per GPU): this object holds rank's index shard and
        the ranks write one file each, ``<stem>.<rank><suffix>`` -- Gadget's multi-file layout
        (``NumFilesPerSnapshot = world``, ``NumPart_Total`` over all files, particle ids continuing
        from file to file in the order a single-file write would give them).  The per-rank counts
        after clipping come from one ``all_gather`` over the default process group, or from
        ``shard_counts`` (list over ranks of ``{ptype: count}``)."""
        import os

        import numpy as np

        from .h5io import H5Writer

        ic_filename = str(ic_filename)
        rank, world = (0, 1) if shard is None else shard
        if shard is not None:
            if not 0 <= rank < world:
                raise ValueError(f"bad shard {shard}")
            stem, ext = os.path.splitext(ic_filename)
            ic_filename = f"{stem}.{rank}{ext}"
        if os.path.exists(ic_filename) and not overwrite:
            raise IOError(f"Cannot create {ic_filename}. It exists and overwrite=False.")
        if np.dtype(dtype) not in (np.dtype("float32"), np.dtype("float64")):
            raise ValueError("dtype must be 'float32' or 'float64'")
        masks = {ptype: self._clip_to_box(ptype, box_size) for ptype in self.particle_types}
        num_particles = {ptype: int(m.sum().item()) for ptype, m in masks.items()}
        # counts[r][ptype] for every rank r
        if shard is None:
            counts = [num_particles]
        elif shard_counts is not None:
            counts = [dict(c) for c in shard_counts]
            if len(counts) != world or any(counts[rank].get(t, 0) != n for t, n in num_particles.items()):
                raise ValueError("shard_counts must list every rank's counts, this rank's included")
        else:
            import torch.distributed as dist

            mine = [num_particles.get(ptype_map[g], 0) for g in ptype_map]
            gathered = [None] * world
            dist.all_gather_object(gathered, mine)
            counts = [dict(zip(ptype_map.values(), g)) for g in gathered]
        totals = {t: sum(c.get(t, 0) for c in counts) for t in ptype_map.values()}
        mass_table = np.zeros(6)
        npart = 0  # ids already taken by the particle types written before
        with H5Writer(ic_filename, overwrite=True) as f:
            for ptype in self.particle_types:
                gptype = rptype_map[ptype]
                idxs = masks[ptype]
                n_in = num_particles[ptype]
                keep_all = n_in == idxs.numel()
                f.create_group(f"/{gptype}")
                for field in gadget_fields[ptype] + code_fields.get(code, {}).get(ptype, []):
                    if field == "ParticleIDs":
                        continue
                    if field == "PassiveScalars":
                        if self.num_passive_scalars > 0:
                            data = torch.stack([self.fields[ptype, s] for s in self.passive_scalars], dim=-1)
                            data = data if keep_all else data[idxs]
                            f.write(f"/{gptype}/PassiveScalars", data.detach().cpu().numpy())
                        continue
                    my_field = gadget_field_map[field]
                    if (ptype, my_field) not in self.fields:
                        continue
                    scale = gadget_unit_scale(field, self.field_units.get((ptype, my_field), ""))
                    fd = self.fields[ptype, my_field]
                    fd = fd if keep_all else fd[idxs]
                    if fd.is_cuda:
                        f.write_device(f"/{gptype}/{field}", fd, out_dtype=dtype, scale=scale)
                    else:
                        f.write(f"/{gptype}/{field}", (fd.double().numpy() * scale).astype(dtype))
                    if field == "Masses" and ptype in ("star", "dm", "black_hole") and n_in:
                        m0 = np.asarray(float(fd[0].item()) * scale).astype(dtype)
                        mass_table[int(gptype[-1])] = m0
                before = sum(c.get(ptype, 0) for c in counts[:rank])
                ids = np.arange(n_in, dtype=np.int64) + 1 + npart + before
                f.write(f"/{gptype}/ParticleIDs", ids.astype("uint32"))
                npart += totals[ptype]
            f.create_group("/Header")
            order = ("gas", "dm", "tracer", None, "star", "black_hole")
            nums = np.array([num_particles.get(t, 0) if t else 0 for t in order], dtype="uint32")
            tot = np.array([totals.get(t, 0) if t else 0 for t in order], dtype=np.int64)
            header = [("Time", 0.0), ("Redshift", 0.0), ("BoxSize", float(box_size)), ("Omega0", 0.0),
                      ("OmegaLambda", 0.0), ("HubbleParam", 1.0), ("NumPart_ThisFile", nums),
                      ("NumPart_Total", (tot & 0xFFFFFFFF).astype("uint32")),
                      ("NumPart_Total_HighWord", (tot >> 32).astype("uint32")),
                      ("NumFilesPerSnapshot", world), ("MassTable", mass_table), ("Flag_Sfr", 0),
                      ("Flag_Cooling", 0), ("Flag_StellarAge", 0), ("Flag_Metals", 0), ("Flag_Feedback", 0),
                      ("Flag_DoublePrecision", 0), ("Flag_IC_Info", 0)]
            for k, v in header:
                f.set_attr("/Header", k, v)
            if code == "arepo":
                f.create_group("/Config")
                f.set_attr("/Config", "VORONOI", 1)
        return ic_filename

    @property
    def num_passive_scalars(self):
        return len(self.passive_scalars)

    @classmethod
    def from_gadget_file(cls, filename, ptypes=None, device=None, all_files=False):
        """Read a Gadget-HDF5 snapshot / IC file into galactic units (particles.py:280-335);
        a missing ``Masses`` dataset is filled from ``Header.MassTable``.  ``all_files=True``:
        ``filename`` is one file ``<stem>.<i><suffix>`` of a multi-file snapshot; all
        ``NumFilesPerSnapshot`` files are read and concatenated in file order."""
        import numpy as np

        from .h5io import H5Reader

        device = device or ("cuda" if torch.cuda.is_available() else "cpu")
        if all_files:
            import re

            with H5Reader(str(filename)) as f:
                nfiles = int(f.attrs("/Header").get("NumFilesPerSnapshot", 1))
            m = re.match(r"^(.*)\.(\d+)(\.[^./]+)$", str(filename))
            if nfiles > 1 and m is None:
                raise ValueError(f"{filename} belongs to a {nfiles}-file snapshot but is not named <stem>.<i><suffix>")
            parts = [cls.from_gadget_file(f"{m.group(1)}.{i}{m.group(3)}" if nfiles > 1 else filename,
                                          ptypes=ptypes, device=device) for i in range(nfiles)]
            out = parts[0]
            for p in parts[1:]:
                out = out + p
            return out
        fields = OrderedDict()
        particle_types = []
        with H5Reader(str(filename)) as f:
            if ptypes is None:
                gptypes = [k for k in f.keys("/") if k.startswith("PartType")]
            else:
                gptypes = [rptype_map[k] for k in ensure_list(ptypes)]
            for gptype in gptypes:
                my_ptype = ptype_map[gptype]
                particle_types.append(my_ptype)
                present = f.keys(f"/{gptype}")
                for field in gadget_fields[my_ptype]:
                    if field not in present:
                        continue
                    a = f.read(f"/{gptype}/{field}")
                    if field == "ParticleIDs":
                        fields[my_ptype, "particle_index"] = torch.from_numpy(a.astype(np.int64)).to(device)
                    else:
                        a = a.astype(np.float64) * _to_galactic(gadget_field_units[field])
                        fields[my_ptype, gadget_field_map[field]] = _to_device_planes(a, device)
                if "Masses" not in present:
                    n_ptype = f.info(f"/{gptype}/ParticleIDs")[1][0]
                    m = float(f.attrs("/Header")["MassTable"][int(gptype[-1])])
                    fields[my_ptype, "particle_mass"] = torch.full(
                        (n_ptype,), m * _to_galactic(gadget_field_units["Masses"]), dtype=torch.float64, device=device)
        return cls(particle_types, fields)

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

    passive_scalars = list(passive_scalars) if passive_scalars else []
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
    if passive_scalars:
        # particles.py:606-637: mass-weighted scalar profiles, into the particles' own scalar fields
        ns = len(passive_scalars)
        tabs = (N.ctypes.c_void_p * (num_halos * ns))()
        outs = (N.ctypes.c_void_p * ns)()
        for i, hse in enumerate(hses):
            if "density" not in hse:
                continue
            rr = np.asarray(hse["radius"], dtype=np.float64)
            for j, name in enumerate(passive_scalars):
                sp = DeviceSpline.fit(rr, np.asarray(hse[name], dtype=np.float64))
                nb = L.cgb_ppform_workspace_bytes(sp.n)
                ws = N.workspace(nb)
                N.check(L.cgb_ppform_prepare(N.ptr(sp.t), N.ptr(sp.c), sp.n, N.ptr(ws), nb, N.stream_ptr()))
                keep.append((sp, ws))
                tabs[i * ns + j] = ws.data_ptr()
        for j, name in enumerate(passive_scalars):
            t = particles["gas", name]
            if t.dtype != dtype or not t.is_contiguous() or t.numel() != n:
                raise ValueError(f"passive scalar {name!r} must be a contiguous (n,) field of the gas dtype")
            outs[j] = t.data_ptr()
        N.check(L.cgb_resample_scalars(halos, num_halos, tabs, ns, N.ptr(pos), n, n, code, outs,
                                       1 if radii is not None else 0, N.stream_ptr()))
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


def resample_two_clusters(particles, hse1, hse2, center1, center2, velocity1, velocity2, radii,
                          passive_scalars=None):
    """particles.py:747-768"""
    return _sample_clusters(particles, [hse1, hse2], [center1, center2], [velocity1, velocity2],
                            radii=radii, resample=True, passive_scalars=passive_scalars)


def resample_three_clusters(particles, hse1, hse2, hse3, center1, center2, center3, velocity1,
                            velocity2, velocity3, radii, passive_scalars=None):
    """particles.py:771-795"""
    return _sample_clusters(particles, [hse1, hse2, hse3], [center1, center2, center3],
                            [velocity1, velocity2, velocity3], radii=radii, resample=True,
                            passive_scalars=passive_scalars)
