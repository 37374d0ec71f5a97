"""
Cluster model: radial tables + the particle-generation API.

Drop-in mirror of the particle-path part of ``cluster_generator.model.ClusterModel``:
constructors ``from_arrays`` (model.py:129-146), ``no_gas`` (:620-656),
``from_dens_and_temp`` (:474-523), ``from_dens_and_tden`` (:565-618), the shared
``_from_scratch`` (:213-269), the lazy ``dm_virial / star_virial`` (:111-127),
``check_hse / check_dm_virial / check_star_virial`` (:667-716) and
``generate_{dm,star,gas,tracer}_particles`` (:766-1039) with the reference
signatures.  Table construction (N ~ 1000 ``quad`` calls) stays on
the host with the reference's own scipy calls; everything per-particle and the
f(E) build run in CUDA.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
from scipy.integrate import cumulative_trapezoid as cumtrapz
from scipy.integrate import quad
from scipy.interpolate import InterpolatedUnivariateSpline

from . import _native as N
from .particles import ClusterParticles
from .utils import _MH_G, _MSUN_G, G, kpc_to_cm, integrate, integrate_device, integrate_mass, integrate_mass_device, keV, mp, mu, mue, mylog, parse_prng, seed_from_prng, unit_array
from .virial import (PRIVATE_TRIAL_SEED, DeviceSpline, SamplerDiagnostics, SamplerTables,  # noqa: F401
                     VirialEquilibrium, _out_code, radius_cdf_table, shard_bounds)

#: default of the constructors' ``integrator`` argument: "host" = the reference's own scipy
#: ``quad`` loops (tables bit-equal to the reference, 1-3 s), "device" = the same QUADPACK
#: decisions on the GPU (tables equal to ~1e-13, 4-11 ms)
DEFAULT_INTEGRATOR = "host"

_UNITS = {
    "radius": "kpc", "density": "Msun/kpc**3", "total_density": "Msun/kpc**3",
    "dark_matter_density": "Msun/kpc**3", "stellar_density": "Msun/kpc**3",
    "total_mass": "Msun", "gas_mass": "Msun", "dark_matter_mass": "Msun", "stellar_mass": "Msun",
    "temperature": "keV", "pressure": "Msun/(Myr**2*kpc)", "gravitational_field": "kpc/Myr**2",
    "gravitational_potential": "kpc**2/Myr**2", "gas_fraction": "",
    "electron_number_density": "cm**-3", "entropy": "keV*cm**2",
}


def _tag(fields):
    for k in list(fields.keys()):
        fields[k] = unit_array(fields[k], _UNITS.get(k, getattr(fields[k], "units", "")))
    return fields


class ClusterModel:
    default_fields = [
        "density", "temperature", "pressure", "total_density", "gravitational_potential",
        "gravitational_field", "total_mass", "gas_mass", "dark_matter_mass", "dark_matter_density",
        "stellar_density", "stellar_mass",
    ]

    def __init__(self, num_elements, fields):
        self.num_elements = int(num_elements)
        self.fields = _tag(fields)
        self._dm_virial = None
        self._star_virial = None
        self._gas_tables = {}

    # -- virial models (model.py:111-127) ------------------------------------
    @property
    def dm_virial(self) -> VirialEquilibrium:
        if self._dm_virial is None:
            self._dm_virial = VirialEquilibrium(self, "dark_matter")
        return self._dm_virial

    @property
    def star_virial(self) -> VirialEquilibrium:
        if self._star_virial is None and "stellar_density" in self:
            self._star_virial = VirialEquilibrium(self, "stellar")
        return self._star_virial

    # -- model files (model.py:149-211, 342-397) -----------------------------
    _keep_units = ["entropy", "electron_number_density", "magnetic_field_strength"]

    def write_model_to_h5(self, output_filename, in_cgs=False, r_min=None, r_max=None, overwrite=False):
        """Write the model tables to HDF5 in the reference's layout: scalar dataset
        ``num_elements``, root attribute ``unit_system``, ``fields/<name>`` datasets with a
        ``units`` attribute, and ``dm_df`` / ``star_df`` when the distribution functions exist."""
        import os

        from .h5io import H5Writer

        if in_cgs:
            raise NotImplementedError("only galactic units (kpc, Msun, Myr) are written; cgs output needs unyt")
        if os.path.exists(output_filename) and not overwrite:
            raise IOError(f"Cannot create {output_filename}. It exists and overwrite=False.")
        r_min = 0.0 if r_min is None else r_min
        r_max = self.fields["radius"].d[-1] * 2 if r_max is None else r_max
        mask = np.logical_and(self.fields["radius"].d >= r_min, self.fields["radius"].d <= r_max)
        with H5Writer(output_filename, overwrite=True) as f:
            f.write("/num_elements", np.asarray(self.num_elements, dtype=np.int64))
            f.set_attr("/", "unit_system", "galactic")
            f.create_group("/fields")
            for k, v in self.fields.items():
                f.write(f"/fields/{k}", v.d[mask])
                f.set_attr(f"/fields/{k}", "units", v.units or "dimensionless")
            for name, vir in (("dm_df", self._dm_virial), ("star_df", self._star_virial)):
                if vir is not None:
                    f.write(f"/{name}", vir.df.d)
                    f.set_attr(f"/{name}", "units", vir.df.units)
        return output_filename

    @classmethod
    def from_h5_file(cls, filename, r_min=None, r_max=None):
        """Read a model written by :meth:`write_model_to_h5` (or by the reference in galactic
        units).  A stored ``dm_df`` / ``star_df`` is reused instead of rebuilding f(E)."""
        from .h5io import H5Reader

        fields = OrderedDict()
        df = {}
        with H5Reader(filename) as f:
            for name in f.keys("/fields"):
                units = f.attrs(f"/fields/{name}").get("units", _UNITS.get(name, ""))
                units = "" if units == "dimensionless" else units
                if name in _UNITS and units != _UNITS[name] and name not in cls._keep_units:
                    raise ValueError(f"field {name!r} is stored in {units!r}; only galactic units "
                                     f"({_UNITS[name]!r}) can be read without unyt")
                fields[name] = unit_array(f.read(f"/fields/{name}").astype(np.float64), units)
            for name in ("dm_df", "star_df"):
                if name in f.keys("/"):
                    df[name] = f.read(f"/{name}").astype(np.float64)
        r_min = 0.0 if r_min is None else r_min
        r_max = fields["radius"].d[-1] * 2 if r_max is None else r_max
        mask = np.logical_and(fields["radius"].d >= r_min, fields["radius"].d <= r_max)
        for name in fields:
            fields[name] = unit_array(fields[name].d[mask], fields[name].units)
        model = cls(int(mask.sum()), fields)
        if "dm_df" in df:
            model._dm_virial = VirialEquilibrium(model, ptype="dark_matter", df=df["dm_df"][mask])
        if "star_df" in df:
            model._star_virial = VirialEquilibrium(model, ptype="stellar", df=df["star_df"][mask])
        return model

    def write_model_to_binary(self, output_filename, fields_to_write=None, in_cgs=False, r_min=None,
                              r_max=None, overwrite=False):
        """Fortran unformatted file: one record with the number of radii, one with the
        (num_radii, num_fields) table (model.py:399-456)."""
        import os

        from scipy.io import FortranFile

        if in_cgs:
            raise NotImplementedError("only galactic units (kpc, Msun, Myr) are written; cgs output needs unyt")
        if fields_to_write is None:
            fields_to_write = list(self.fields.keys())
        if os.path.exists(output_filename) and not overwrite:
            raise IOError(f"Cannot create {output_filename}. It exists and overwrite=False.")
        r_min = 0.0 if r_min is None else r_min
        r_max = self.fields["radius"].d[-1] * 2 if r_max is None else r_max
        mask = np.logical_and(self.fields["radius"].d >= r_min, self.fields["radius"].d <= r_max)
        with FortranFile(str(output_filename), "w") as f:
            f.write_record(self.fields["radius"].d[mask].size)
            f.write_record(np.array([self.fields[k].d[mask] for k in fields_to_write]).T)
        return output_filename

    def write_model_to_ascii(self, output_filename, in_cgs=False, overwrite=False):
        """Plain-text table, one column per field, units in the header (model.py:307-340 writes an
        astropy QTable; astropy is not available, the column layout is the same)."""
        import os

        if in_cgs:
            raise NotImplementedError("only galactic units (kpc, Msun, Myr) are written; cgs output needs unyt")
        if os.path.exists(output_filename) and not overwrite:
            raise IOError(f"Cannot create {output_filename}. It exists and overwrite=False.")
        names = list(self.fields.keys())
        header = "unit_system=galactic\n" + " ".join(names) + "\n" + \
                 " ".join(self.fields[k].units or "dimensionless" for k in names)
        np.savetxt(str(output_filename), np.array([self.fields[k].d for k in names]).T, header=header)
        return output_filename

    def mass_in_radius(self, radius):
        """Mass of each matter type inside ``radius`` (model.py:1088-1107)."""
        r = self.fields["radius"].d
        return {mtype: unit_array(self.fields[f"{mtype}_mass"].d[r < radius][-1], "Msun")
                for mtype in ("total", "gas", "dark_matter", "stellar") if f"{mtype}_mass" in self.fields}

    def find_radius_for_density(self, density):
        """Radius at which the gas density falls to ``density`` [Msun/kpc**3] (model.py:1109-1113)."""
        r = self.fields["radius"].d[::-1]
        d = self.fields["density"].d[::-1]
        return unit_array(np.interp(float(density), d, r), "kpc")

    # -- construction --------------------------------------------------------
    @classmethod
    def from_arrays(cls, fields):
        return cls(np.asarray(fields["radius"]).size, OrderedDict(fields))

    @classmethod
    def _from_scratch(cls, fields, stellar_density=None, integrator=None):
        """model.py:213-269.  ``integrator="device"`` runs the N potential integrals
        (utils.py:137-142) and the mass integrals of the constructors (utils.py:129-134) -- ~1-3 s
        of host ``quad`` callbacks -- through the device QUADPACK restatement instead
        (utils.integrate_spline_device / quad_callable_device)."""
        integrator = DEFAULT_INTEGRATOR if integrator is None else integrator
        if integrator not in ("host", "device"):
            raise ValueError("integrator must be 'host' or 'device'")
        rr = np.asarray(fields["radius"])
        mylog.info("Integrating gravitational potential profile.")
        tdens_func = InterpolatedUnivariateSpline(rr, fields["total_density"])
        gpot_profile = lambda r: tdens_func(r) * r
        gpot1 = fields["total_mass"] / fields["radius"]
        if integrator == "device":
            from .utils import integrate_spline_device

            gpot2 = 4.0 * np.pi * integrate_spline_device(tdens_func, rr, np.full_like(rr, rr[-1]), power=1)
        else:
            gpot2 = 4.0 * np.pi * integrate(gpot_profile, rr)
        fields["gravitational_potential"] = -G * (gpot1 + gpot2)

        if "density" in fields and "gas_mass" not in fields:
            mylog.info("Integrating gas mass profile.")
            m0 = fields["density"][0] * rr[0] ** 3 / 3.0
            fields["gas_mass"] = (
                4.0 * np.pi * cumtrapz(fields["density"] * rr * rr, x=rr, initial=0.0) + m0
            )

        if stellar_density is not None:
            fields["stellar_density"] = stellar_density(rr)
            mylog.info("Integrating stellar mass profile.")
            fields["stellar_mass"] = (integrate_mass_device if integrator == "device" else integrate_mass)(stellar_density, rr)

        mdm = np.array(fields["total_mass"], dtype=np.float64)
        ddm = np.array(fields["total_density"], dtype=np.float64)
        if "density" in fields:
            mdm -= fields["gas_mass"]
            ddm -= fields["density"]
        if "stellar_mass" in fields:
            mdm -= fields["stellar_mass"]
            ddm -= fields["stellar_density"]
        mdm[ddm < 0.0] = mdm.max()
        ddm[ddm < 0.0] = 0.0
        if ddm.sum() < 0.0 or mdm.sum() < 0.0:
            mylog.warning("The total dark matter mass is either zero or negative!!")
        fields["dark_matter_density"] = ddm
        fields["dark_matter_mass"] = mdm

        if "density" in fields:
            fields["gas_fraction"] = fields["gas_mass"] / fields["total_mass"]
            # model.py:262-267: n_e = rho / (mue * m_H) in cm**-3 (unyt's number-density
            # equivalence), entropy = kT * n_e**(-2/3) in keV*cm**2
            fields["electron_number_density"] = (
                np.asarray(fields["density"]) * (_MSUN_G / kpc_to_cm**3) / (mue * _MH_G))
            fields["entropy"] = np.asarray(fields["temperature"]) * fields["electron_number_density"] ** (-2.0 / 3.0)
        return cls(rr.size, fields)

    @classmethod
    def no_gas(cls, rmin, rmax, total_density, stellar_density=None, num_points=1000, integrator=None):
        """Model without a gas component (model.py:620-656)."""
        rr = np.logspace(np.log10(rmin), np.log10(rmax), num_points, endpoint=True)
        fields = OrderedDict()
        fields["radius"] = rr
        fields["total_density"] = total_density(rr)
        mylog.info("Integrating total mass profile.")
        integrator = DEFAULT_INTEGRATOR if integrator is None else integrator
        fields["total_mass"] = (integrate_mass_device if integrator == "device" else integrate_mass)(total_density, rr)
        fields["gravitational_field"] = -G * fields["total_mass"] / (fields["radius"] ** 2)
        return cls._from_scratch(fields, stellar_density=stellar_density, integrator=integrator)

    @classmethod
    def from_dens_and_temp(cls, rmin, rmax, density, temperature, stellar_density=None, num_points=1000,
                           integrator=None):
        """Hydrostatic model from gas density and temperature (model.py:474-523)."""
        mylog.info("Computing the profiles from density and temperature.")
        rr = np.logspace(np.log10(rmin), np.log10(rmax), num_points, endpoint=True)
        fields = OrderedDict()
        fields["radius"] = rr
        fields["density"] = density(rr)
        fields["temperature"] = temperature(rr)
        fields["pressure"] = fields["density"] * fields["temperature"]
        fields["pressure"] = fields["pressure"] / (mu * mp)
        fields["pressure"] = fields["pressure"] * keV
        pressure_spline = InterpolatedUnivariateSpline(rr, fields["pressure"])
        dPdr = pressure_spline(rr, 1)
        fields["gravitational_field"] = dPdr / fields["density"]
        integrator = DEFAULT_INTEGRATOR if integrator is None else integrator
        fields["gas_mass"] = (integrate_mass_device if integrator == "device" else integrate_mass)(density, rr)
        fields["total_mass"] = -fields["radius"] ** 2 * fields["gravitational_field"] / G
        total_mass_spline = InterpolatedUnivariateSpline(rr, fields["total_mass"])
        dMdr = total_mass_spline(rr, nu=1)
        fields["total_density"] = dMdr / (4.0 * np.pi * fields["radius"] ** 2)
        return cls._from_scratch(fields, stellar_density=stellar_density, integrator=integrator)

    @classmethod
    def from_dens_and_entr(cls, rmin, rmax, density, entropy, stellar_density=None, num_points=1000,
                           integrator=None):
        """Hydrostatic model from gas density and entropy profiles (model.py:525-563):
        kT = S n_e^(2/3) with n_e = rho / (mue mp) in cm^-3 (entropy in keV cm^2)."""
        from .radial_profiles import RadialProfile
        from .utils import kpc_to_cm

        n_e = RadialProfile(lambda r: density(r) / (mue * mp * kpc_to_cm**3))
        temperature = RadialProfile(lambda r: entropy(r) * n_e(r) ** (2.0 / 3.0))
        return cls.from_dens_and_temp(rmin, rmax, density, temperature, stellar_density=stellar_density,
                                      num_points=num_points, integrator=integrator)

    @classmethod
    def from_dens_and_tden(cls, rmin, rmax, density, total_density, stellar_density=None, num_points=1000,
                           integrator=None):
        """Hydrostatic model from gas density and total density (model.py:565-618)."""
        integrator = DEFAULT_INTEGRATOR if integrator is None else integrator
        if integrator not in ("host", "device"):
            raise ValueError("integrator must be 'host' or 'device'")
        imass = integrate_mass_device if integrator == "device" else integrate_mass
        iint = integrate_device if integrator == "device" else integrate
        mylog.info("Computing the profiles from density and total density.")
        rr = np.logspace(np.log10(rmin), np.log10(rmax), num_points, endpoint=True)
        fields = OrderedDict()
        fields["radius"] = rr
        fields["density"] = density(rr)
        fields["total_density"] = total_density(rr)
        fields["total_mass"] = imass(total_density, rr)
        fields["gas_mass"] = imass(density, rr)
        fields["gravitational_field"] = -G * fields["total_mass"] / (fields["radius"] ** 2)
        g = fields["gravitational_field"]
        g_r = InterpolatedUnivariateSpline(rr, g)
        dPdr_int = lambda r: density(r) * g_r(r)
        P = -iint(dPdr_int, rr)
        dPdr_int2 = lambda r: density(r) * g[-1] * (rr[-1] / r) ** 2
        P -= quad(dPdr_int2, rr[-1], np.inf, limit=100)[0]
        fields["pressure"] = P
        fields["temperature"] = fields["pressure"] * mu * mp / fields["density"] / keV
        return cls._from_scratch(fields, stellar_density=stellar_density, integrator=integrator)

    def set_rmax(self, r_max):
        """Truncate the model at ``r_max`` (model.py:271-293; the reference passes kwargs its
        constructor does not accept -- here the truncated model simply rebuilds f(E) lazily)."""
        mask = self.fields["radius"].d <= r_max
        return ClusterModel(int(mask.sum()), OrderedDict((k, v.d[mask]) for k, v in self.fields.items()))

    # -- dict protocol (model.py:295-305) ------------------------------------
    def __getitem__(self, key):
        return self.fields[key]

    def __contains__(self, key):
        return key in self.fields

    def keys(self):
        return self.fields.keys()

    def set_field(self, name, value):
        """model.py:458-472"""
        value = np.asarray(value)
        if value.size != self.num_elements:
            raise ValueError(f"The length of the array needs to be {self.num_elements} elements!")
        if name in self.fields:
            mylog.warning("Overwriting field %s." % name)
        self.fields[name] = unit_array(value, _UNITS.get(name, ""))
        # The reference rebuilds its CDF and splines from self.fields on every generate_* call
        # (virial.py:148-166, model.py:879-930) while a VirialEquilibrium keeps the f(E) it was
        # built with: drop the cached device tables, keep the virial objects.
        self._gas_tables.clear()
        for vir in (self._dm_virial, self._star_virial):
            if vir is not None:
                vir._tables.clear()

    def find_field_at_radius(self, field, r):
        """model.py:658-665"""
        return unit_array(np.interp(r, self["radius"].d, self[field].d), self[field].units)

    # -- checks (model.py:667-716) -------------------------------------------
    def check_hse(self):
        if "pressure" not in self.fields:
            raise RuntimeError("This ClusterModel contains no gas!")
        rr = self.fields["radius"].d
        pressure_spline = InterpolatedUnivariateSpline(rr, self.fields["pressure"].d)
        dPdx = pressure_spline(rr, 1)
        rhog = self.fields["density"].d * self.fields["gravitational_field"].d
        chk = (dPdx - rhog) / rhog
        mylog.info("The maximum relative deviation of this profile from hydrostatic equilibrium is %g",
                   np.abs(chk).max())
        return chk

    def check_dm_virial(self):
        return self.dm_virial.check_virial()

    def check_star_virial(self):
        return self.star_virial.check_virial()

    # -- particle generation ---------------------------------------------------
    def generate_dm_particles(self, num_particles, r_max=None, sub_sample=1, compute_potential=False,
                              prng=None, **kw) -> ClusterParticles:
        """Dark matter particles in virial equilibrium (model.py:949-993)."""
        return self.dm_virial.generate_particles(
            num_particles, r_max=r_max, sub_sample=sub_sample, compute_potential=compute_potential,
            prng=prng, **kw)

    def generate_star_particles(self, num_particles, r_max=None, sub_sample=1, compute_potential=False,
                                prng=None, **kw) -> ClusterParticles:
        """Star particles in virial equilibrium (model.py:995-1039)."""
        return self.star_virial.generate_particles(
            num_particles, r_max=r_max, sub_sample=sub_sample, compute_potential=compute_potential,
            prng=prng, **kw)

    def _gas_sampler_tables(self, r_max, tracer) -> SamplerTables:
        key = (None if r_max is None else float(r_max), bool(tracer))
        if key not in self._gas_tables:
            rr = self["radius"].d
            # gas / tracer radii use ALL rows (model.py:803-809, 879-885): no rho > 0 mask
            P_r, r_cdf, mtot = radius_cdf_table(rr, self["gas_mass"].d, r_max)
            psi_spline = DeviceSpline.fit(rr, -self["gravitational_potential"].d)  # model.py:939-941
            aux = []
            if not tracer:
                e_arr = 1.5 * self["pressure"].d / self["density"].d                # model.py:908
                aux = [DeviceSpline.fit(rr, e_arr), DeviceSpline.fit(rr, self["density"].d)]
            self._gas_tables[key] = SamplerTables(P_r, r_cdf, mtot, psi_spline, None, aux)
        return self._gas_tables[key]

    def _generate_radial(self, ptype, num_particles, r_max, sub_sample, compute_potential, prng,
                         dtype, rng, shard, host=False, chunk=None, wait=True):
        if "gas_mass" not in self.fields:
            raise RuntimeError("This ClusterModel contains no gas!")
        num_particles, sub_sample = int(num_particles), int(sub_sample)
        if sub_sample < 1 or num_particles % sub_sample != 0:
            raise ValueError("num_particles must be a multiple of sub_sample")
        if rng not in ("device", "numpy"):
            raise ValueError("rng must be 'device' or 'numpy'")
        tracer = ptype == "tracer"
        prng = parse_prng(prng)
        dev = N.require_cuda()
        tabs = self._gas_sampler_tables(r_max, tracer)
        lo, hi = shard_bounds(num_particles, shard)
        n_loc = hi - lo
        code = _out_code(dtype)
        keep = []
        a = N.SampleArgs()
        n_sub = num_particles // sub_sample
        if sub_sample > 1:
            a.radius_period = n_sub            # np.tile(radius_sub, sub_sample) (model.py:887-890)
        if rng == "numpy":
            u_r = prng.uniform(size=n_sub)           # model.py:879-885 -> utils.py:161
            u_ct = prng.uniform(size=num_particles)  # model.py:892
            u_ph = prng.uniform(size=num_particles)  # model.py:893
            for name, arr in (("d_u_r", u_r if sub_sample > 1 else u_r[lo:hi]), ("d_u_ct", u_ct[lo:hi]),
                              ("d_u_ph", u_ph[lo:hi])):
                t = N.dev_f64(arr, dev)
                keep.append(t)
                setattr(a.hu, name, t.data_ptr())
            seed = PRIVATE_TRIAL_SEED
        else:
            seed = seed_from_prng(prng, 4 if tracer else 3)
        a.d_workspace = tabs.workspace.data_ptr()
        # tracers carry zero mass (model.py:834); gas mtot / n (model.py:917-919)
        a.particle_mass = 0.0 if tracer else (tabs.mtot / num_particles if num_particles else 0.0)
        a.seed = seed & 0xFFFFFFFFFFFFFFFF
        a.out_dtype = code
        a.flags = tabs.gas_flags
        if host and rng == "device":
            # host-resident output: chunks staged on the device, copies overlapped with the kernels
            from .hoststream import HostStreamer

            layout = OrderedDict([("particle_position", 3)])
            if not tracer:
                layout["thermal_energy"] = 1
            layout["particle_mass"] = 1
            if not tracer:
                layout["density"] = 1
            if compute_potential:
                layout["particle_potential"] = 1
            prev = getattr(host, "_host_buffers", {}).get(ptype) if isinstance(host, ClusterParticles) else None
            hs = HostStreamer(layout, n_loc, dtype, dev, chunk, out=prev)

            def launch(stage, b0, b1):
                a.id_begin, a.id_end, a.ld = lo + b0, lo + b1, hs.chunk
                a.d_pos, a.d_vel, a.d_mass = stage["particle_position"].data_ptr(), None, stage["particle_mass"].data_ptr()
                a.d_pot = stage["particle_potential"].data_ptr() if compute_potential else None
                N.check(N.lib().cgb_sample_gas(N.ctypes.byref(a), N.ptr(stage.get("thermal_energy")),
                                               N.ptr(stage.get("density")), N.stream_ptr()))

            hb = hs.run(launch, wait=wait) if n_loc > 0 else hs.host
            # the velocities are identically zero (model.py:829-831, 933-935): host-side zeros
            velh = getattr(host, "_host_vel", {}).get(ptype) if isinstance(host, ClusterParticles) else None
            if velh is None or velh.shape != (3, n_loc) or velh.dtype != dtype:
                velh = torch.zeros((3, n_loc), dtype=dtype, pin_memory=True)
            if wait:
                hs.copy_stream.synchronize()
            fields = OrderedDict()
            fields[ptype, "particle_position"] = hb.arrays["particle_position"].T
            if not tracer:
                fields[ptype, "thermal_energy"] = hb.arrays["thermal_energy"]
            fields[ptype, "particle_mass"] = hb.arrays["particle_mass"]
            if not tracer:
                fields[ptype, "density"] = hb.arrays["density"]
            fields[ptype, "particle_velocity"] = velh.T
            if compute_potential:
                fields[ptype, "particle_potential"] = hb.arrays["particle_potential"]
            out = ClusterParticles(ptype, fields)
            out._host_buffers, out._host_vel, out.d2h_bytes = {ptype: hb}, {ptype: velh}, hs.bytes
            if not wait:
                out._pending.append(hs)  # keeps the staging slots alive until the copies have left them
            return out
        pos = torch.empty((3, n_loc), dtype=dtype, device=dev)
        vel = torch.empty((3, n_loc), dtype=dtype, device=dev)
        mass = torch.empty(n_loc, dtype=dtype, device=dev)
        pot = torch.empty(n_loc, dtype=dtype, device=dev) if compute_potential else None
        therm = dens = None
        if not tracer:
            therm = torch.empty(n_loc, dtype=dtype, device=dev)
            dens = torch.empty(n_loc, dtype=dtype, device=dev)
        a.id_begin, a.id_end, a.ld = lo, hi, n_loc
        a.d_pos, a.d_vel, a.d_mass = pos.data_ptr(), vel.data_ptr(), mass.data_ptr()
        a.d_pot = pot.data_ptr() if pot is not None else None
        if n_loc > 0:
            N.check(N.lib().cgb_sample_gas(N.ctypes.byref(a), N.ptr(therm), N.ptr(dens), N.stream_ptr()))
        fields = OrderedDict()
        fields[ptype, "particle_position"] = pos.T
        if not tracer:
            fields[ptype, "thermal_energy"] = therm
        fields[ptype, "particle_mass"] = mass
        if not tracer:
            fields[ptype, "density"] = dens
        fields[ptype, "particle_velocity"] = vel.T
        if compute_potential:
            fields[ptype, "particle_potential"] = pot
        del keep
        if host:
            from .virial import _pinned_copy

            fields = OrderedDict((k, _pinned_copy(v)) for k, v in fields.items())
            torch.cuda.current_stream().synchronize()
        return ClusterParticles(ptype, fields)

    def generate_gas_particles(self, num_particles, r_max=None, sub_sample=1, compute_potential=False,
                               prng=None, *, dtype=torch.float64, rng="device", shard=None, host=False,
                               chunk=None, wait=True):
        """Gas particles in hydrostatic equilibrium (model.py:838-947).  ``host``, ``wait``: see
        :meth:`VirialEquilibrium.generate_particles`."""
        return self._generate_radial("gas", num_particles, r_max, sub_sample, compute_potential, prng,
                                     dtype, rng, shard, host, chunk, wait)

    def generate_tracer_particles(self, num_particles, r_max=None, sub_sample=1, prng=None, *,
                                  dtype=torch.float64, rng="device", shard=None, host=False, chunk=None,
                                  wait=True):
        """Tracer particles following the gas distribution (model.py:766-836)."""
        return self._generate_radial("tracer", num_particles, r_max, sub_sample, False, prng, dtype,
                                     rng, shard, host, chunk, wait)
