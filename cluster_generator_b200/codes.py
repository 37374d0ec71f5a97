"""
Particle files for the AMR codes (GAMER / FLASH / RAMSES): ``write_amr_particles``
(codes.py:16-73).  Only this writer -- the function on the particle path that turns a
``ClusterParticles`` into the flat file those codes read; the code-specific parameter-file
generators of ``codes.py`` stay out of scope.

File layout, as the reference writes it:

* ``format="hdf5"``: datasets ``particle_position`` (N, 3), ``particle_velocity`` (N, 3),
  ``particle_mass`` (N,) -- the listed particle types concatenated in the given order -- and
  ``particle_type`` (N,) float64 holding ``ptype_num[ptype]``.
* ``format="fortran"``: two unformatted sequential records (``scipy.io.FortranFile``): the
  particle count, then an (N, 7) float64 array of position, velocity and mass.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .utils import _M_PER_KPC, _MSUN_KG, _SEC_PER_MYR

# galactic (kpc, Msun, Myr) -> cgs, what unyt's ``convert_to_cgs`` multiplies by
_CGS = {
    "particle_position": _M_PER_KPC * 1.0e2,
    "particle_velocity": _M_PER_KPC * 1.0e2 / _SEC_PER_MYR,
    "particle_mass": _MSUN_KG * 1.0e3,
}


def write_amr_particles(particles, output_filename, ptypes, ptype_num, overwrite=True, in_cgs=False,
                        format="hdf5"):
    """Write ``ptypes`` of ``particles`` for GAMER, FLASH or RAMSES (codes.py:16-73)."""
    output_filename = str(output_filename)
    if os.path.exists(output_filename) and not overwrite:
        raise IOError(f"Cannot create {output_filename}. It exists and overwrite=False.")
    if format not in ("hdf5", "fortran"):
        raise ValueError("format must be 'hdf5' or 'fortran'")
    ptypes = list(ptypes)
    nparts = [particles.num_particles[ptype] for ptype in ptypes]
    fields = {}
    for field in ("particle_position", "particle_velocity", "particle_mass"):
        fd = torch.cat([particles[ptype, field] for ptype in ptypes], dim=0)
        fields[field] = (fd, _CGS[field] if in_cgs else 1.0)
    if format == "hdf5":
        from .h5io import H5Writer

        with H5Writer(output_filename, overwrite=True) as f:
            for field, (fd, scale) in fields.items():
                if fd.is_cuda and fd.dtype in (torch.float32, torch.float64):
                    f.write_device("/" + field, fd, out_dtype=str(fd.dtype).split(".")[1], scale=scale)
                else:
                    f.write("/" + field, fd.detach().cpu().numpy() * scale)
            ptype_field = np.concatenate([ptype_num[ptype] * np.ones(nparts[i]) for i, ptype in enumerate(ptypes)])
            f.write("/particle_type", ptype_field)
        return output_filename
    from scipy.io import FortranFile

    pdata = [(fd.detach().cpu().numpy().astype("float64") * scale).T for fd, scale in fields.values()]
    with FortranFile(output_filename, "w") as f:
        f.write_record(np.int32(sum(nparts)))
        f.write_record(np.vstack(pdata).T)
    return output_filename
