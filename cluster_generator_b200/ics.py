"""
Multi-halo initial conditions driver ("next" row 1 of SURVEY.md section 8(f)).

Mirror of ``cluster_generator.ics.ClusterICs`` (ics.py:60-375) and
``compute_centers_for_binary`` (ics.py:22-57) for the particle path: split the particle
budget over halos by enclosed mass, generate DM / star / gas / tracer particles of every
halo from ONE shared ``prng`` (stream order matters, ics.py:173-192), then offset and
combine them, re-sampling the gas fields where halos overlap.

``profiles`` are HDF5 model files (``ClusterModel.write_model_to_h5``) as in the reference, or
:class:`ClusterModel` objects directly.  With an ``output_directory`` the per-halo particles
are written to ``{output_directory}/{basename}_{i}_particles.h5`` and reused on the next call
unless ``regenerate_particles`` (ics.py:173-199); without one nothing touches the disk (the
reference would write to the file-system root).
"""
from __future__ import annotations

from collections import defaultdict
from numbers import Number

import numpy as np

from .particles import (ClusterParticles, combine_three_clusters, combine_two_clusters, resample_one_cluster,
                        resample_three_clusters, resample_two_clusters)
from .utils import ensure_list, parse_prng


def compute_centers_for_binary(center, d, b, a=0.0):
    """Centres of two clusters a distance d apart with impact parameters b (y) and a (z)
    around a common centre (ics.py:22-57)."""
    d = np.sqrt(d * d - b * b - a * a)
    diff = np.array([d, b, a])
    center1 = center - 0.5 * diff
    center2 = center + 0.5 * diff
    return center1, center2


class ClusterICs:
    def __init__(self, basename, num_halos, profiles, center, velocity, num_particles=None,
                 mag_file=None, particle_files=None, r_max=20000.0, r_max_tracer=None):
        self.basename = basename
        self.num_halos = int(num_halos)
        if not 1 <= self.num_halos <= 3:
            raise ValueError("num_halos must be 1, 2 or 3")
        self.profiles = ensure_list(profiles)
        if len(self.profiles) != self.num_halos:
            raise ValueError("one profile (ClusterModel) per halo is required")
        self.particle_files = [None] * self.num_halos
        if particle_files is not None:
            self.particle_files = ensure_list(particle_files)[:]
        self._models = None
        self.center = np.asarray(center, dtype=np.float64).reshape(self.num_halos, 3)
        self.velocity = np.asarray(velocity, dtype=np.float64).reshape(self.num_halos, 3)
        self.mag_file = mag_file
        if isinstance(r_max, Number):
            r_max = [r_max] * self.num_halos
        self.r_max = np.array(r_max)
        if r_max_tracer is None:
            r_max_tracer = r_max
        if isinstance(r_max_tracer, Number):
            r_max_tracer = [r_max_tracer] * self.num_halos
        self.r_max_tracer = np.array(r_max_tracer)
        self.tot_np = num_particles if num_particles is not None else {"dm": 0, "gas": 0, "star": 0, "tracer": 0}
        self._determine_num_particles()

    @property
    def models(self):
        """The halos' :class:`ClusterModel` objects (profile files are read once)."""
        if self._models is None:
            from .model import ClusterModel

            self._models = [p if isinstance(p, ClusterModel) else ClusterModel.from_h5_file(p)
                            for p in self.profiles]
        return self._models

    def _determine_num_particles(self):
        """n_i = rint(N_tot * M_i(<r_max) / sum M) per species (ics.py:100-160)."""
        dm_masses, gas_masses, star_masses, tracer_masses = [], [], [], []
        for i, p in enumerate(self.models):
            idxs = p["radius"].d < self.r_max[i]
            dm_masses.append(p["dark_matter_mass"].d[idxs][-1])
            gas_masses.append(p["gas_mass"].d[idxs][-1] if "gas_mass" in p else 0.0)
            star_masses.append(p["stellar_mass"].d[idxs][-1] if "stellar_mass" in p else 0.0)
            if self.tot_np.get("tracer", 0) > 0:
                idxst = p["radius"].d < self.r_max_tracer[i]
                tracer_masses.append(p["gas_mass"].d[idxst][-1])
            else:
                tracer_masses.append(0.0)
        totals = {"dm": np.sum(dm_masses), "gas": np.sum(gas_masses), "star": np.sum(star_masses),
                  "tracer": np.sum(tracer_masses)}
        masses = {"dm": dm_masses, "gas": gas_masses, "star": star_masses, "tracer": tracer_masses}
        self.num_particles = defaultdict(list)
        for i in range(self.num_halos):
            for key in ("dm", "gas", "star", "tracer"):
                if self.tot_np.get(key, 0) > 0:
                    n = int(np.rint(self.tot_np[key] * masses[key][i] / totals[key]))
                else:
                    n = 0
                self.num_particles[key].append(n)

    def _generate_particles(self, output_directory=None, regenerate_particles=False, prng=None, **kw):
        """ics.py:162-200: dm -> star -> gas -> tracer per halo, one shared prng."""
        prng = parse_prng(prng)
        parts = []
        for i, m in enumerate(self.models):
            if not (regenerate_particles or self.particle_files[i] is None):
                parts.append(ClusterParticles.from_file(self.particle_files[i]))
                continue
            p = m.generate_dm_particles(self.num_particles["dm"][i], r_max=self.r_max[i], prng=prng, **kw)
            if self.num_particles["star"][i] > 0:
                p = p + m.generate_star_particles(self.num_particles["star"][i], r_max=self.r_max[i],
                                                  prng=prng, **kw)
            if self.num_particles["gas"][i] > 0:
                p = p + m.generate_gas_particles(self.num_particles["gas"][i], r_max=self.r_max[i],
                                                 prng=prng, **kw)
            if self.num_particles["tracer"][i] > 0:
                p = p + m.generate_tracer_particles(self.num_particles["tracer"][i],
                                                    r_max=self.r_max_tracer[i], prng=prng, **kw)
            parts.append(p)
            if output_directory is not None:
                outfile = f"{output_directory}/{self.basename}_{i}_particles.h5"
                p.write_particles(outfile, overwrite=True)
                self.particle_files[i] = outfile
        return parts

    def setup_particle_ics(self, output_directory=None, regenerate_particles=False, prng=None, **kw):
        """Generate and combine the particles of all halos (ics.py:323-375)."""
        profiles = self.models
        parts = self._generate_particles(output_directory=output_directory,
                                         regenerate_particles=regenerate_particles, prng=prng, **kw)
        if self.num_halos == 1:
            all_parts = parts[0]
            all_parts.add_offsets(self.center[0], self.velocity[0])
        elif self.num_halos == 2:
            all_parts = combine_two_clusters(parts[0], parts[1], profiles[0], profiles[1],
                                             self.center[0], self.center[1], self.velocity[0],
                                             self.velocity[1])
        else:
            all_parts = combine_three_clusters(parts[0], parts[1], parts[2], profiles[0], profiles[1],
                                               profiles[2], self.center[0], self.center[1],
                                               self.center[2], self.velocity[0], self.velocity[1],
                                               self.velocity[2])
        return all_parts

    def resample_particle_ics(self, parts, passive_scalars=None):
        """Re-sample density, thermal energy and velocity of relaxed gas particles / cells from
        the initial hydrostatic profiles (ics.py:377-422)."""
        m, c, v = self.models, self.center, self.velocity
        if self.num_halos == 1:
            return resample_one_cluster(parts, m[0], c[0], v[0])
        if self.num_halos == 2:
            return resample_two_clusters(parts, m[0], m[1], c[0], c[1], v[0], v[1], self.r_max,
                                         passive_scalars=passive_scalars)
        return resample_three_clusters(parts, m[0], m[1], m[2], c[0], c[1], c[2], v[0], v[1], v[2], self.r_max,
                                       passive_scalars=passive_scalars)

    # -- parameter files (ics.py:202-320; same keys, written with PyYAML instead of ruamel) -----
    def to_file(self, filename, overwrite=False):
        """Write the initial-conditions description to a YAML file.  Profiles must be model
        files (``ClusterModel.write_model_to_h5``), as in the reference."""
        import os

        import yaml

        if os.path.exists(filename) and not overwrite:
            raise RuntimeError(f"{filename} exists and overwrite=False!")
        if not all(isinstance(p, (str, os.PathLike)) for p in self.profiles):
            raise ValueError("to_file needs the profiles as model file names (write_model_to_h5)")
        out = {"basename": self.basename, "num_halos": self.num_halos}
        for i in range(self.num_halos):
            out[f"profile{i + 1}"] = os.fspath(self.profiles[i])
            out[f"center{i + 1}"] = [float(x) for x in self.center[i]]
            out[f"velocity{i + 1}"] = [float(x) for x in self.velocity[i]]
            if self.particle_files[i] is not None:
                out[f"particle_file{i + 1}"] = os.fspath(self.particle_files[i])
        for key in ("dm", "gas", "star", "tracer"):
            if self.tot_np.get(key, 0) > 0:
                out[f"num_{key}_particles"] = int(self.tot_np[key])
        if self.mag_file is not None:
            out["mag_file"] = self.mag_file
        out["r_max"] = [float(x) for x in self.r_max]
        if self.tot_np.get("tracer", 0) > 0:
            out["r_max_tracer"] = [float(x) for x in self.r_max_tracer]
        with open(filename, "w") as f:
            yaml.safe_dump(out, f, sort_keys=False)
        return filename

    @classmethod
    def from_file(cls, filename):
        """Read a YAML file written by :meth:`to_file` (or by the reference)."""
        import yaml

        with open(filename, "r") as f:
            params = yaml.safe_load(f)
        num_halos = params["num_halos"]
        rng = range(1, num_halos + 1)
        num_particles = {k: params.get(f"num_{k}_particles", 0) for k in ("gas", "dm", "star", "tracer")}
        r_max = params.get("r_max", 20000.0)
        return cls(params["basename"], num_halos, [params[f"profile{i}"] for i in rng],
                   [np.array(params[f"center{i}"]) for i in rng], [np.array(params[f"velocity{i}"]) for i in rng],
                   num_particles=num_particles, mag_file=params.get("mag_file", None),
                   particle_files=[params.get(f"particle_file{i}", None) for i in rng],
                   r_max=r_max, r_max_tracer=params.get("r_max_tracer", r_max))
