"""
TESTS ONLY -- a minimal stand-in for the ``unyt`` package (absent from this image).

It exists for one purpose: to let the reference's own modules
(``/root/reference/cluster_generator/{utils,virial,particles,model,ics}.py``) be
imported and executed UNMODIFIED, so that ``oracle/`` can be pinned to the numbers
the reference's code itself produces.  Nothing in the product
imports this.

What is restated from unyt's documented behaviour (the reference's call sites are the
specification; no unyt source is available here):

* ``unyt_array`` / ``unyt_quantity`` are ndarray subclasses carrying ``units``;
  ``.d`` is a plain-ndarray view, ``.v`` / ``.value`` a plain copy;
* units form a symbolic product of named base symbols with rational exponents, so
  that e.g. ``kpc**3/Msun/Myr**2 * Msun/kpc`` *is* ``kpc**2/Myr**2`` and converting
  between them multiplies by exactly 1.0 (as unyt's sympy expressions do);
* a conversion multiplies the values by ``old.base_value / new.base_value`` (unyt's
  ``get_conversion_factor``), the base value of a unit being the product over its
  symbols (sorted by name) of ``mks_value ** exponent``;
* multiplication / division combine units without touching the values;
  addition / subtraction / comparison convert the second operand to the units of
  the first.

The constants are unyt's CODATA-2014 mks values.  The arithmetic order inside a base
value is this shim's choice (unyt's depends on sympy's argument ordering, which
cannot be observed here); product, oracle and this shim therefore share ONE set of
converted constants -- ``tests/test_reference_cpu.py`` asserts that they are bitwise the same.
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np

__all__ = ["unyt_array", "unyt_quantity", "uconcatenate", "Unit", "physical_constants"]

# mks value and dimensions (mass, length, time, temperature) of every named symbol
_SYMBOLS: dict[str, tuple[float, tuple]] = {}


def _def(name, value, dims):
    _SYMBOLS[name] = (float(value), tuple(Fraction(d) for d in dims))


_M, _L, _T, _K = (1, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1)
_ENERGY = (1, 2, -2, 0)
_def("kg", 1.0, _M)
_def("g", 1.0e-3, _M)
_def("Msun", 1.98841586e30, _M)
_def("amu", 1.660539040e-27, _M)
_def("m", 1.0, _L)
_def("cm", 1.0e-2, _L)
_def("km", 1.0e3, _L)
_def("pc", 3.0856775809623245e16, _L)
_def("kpc", 3.0856775809623245e19, _L)
_def("Mpc", 3.0856775809623245e22, _L)
_def("s", 1.0, _T)
_def("yr", 31557600.0, _T)
_def("Myr", 31557600.0 * 1.0e6, _T)
_def("Gyr", 31557600.0 * 1.0e9, _T)
_def("K", 1.0, _K)
_def("J", 1.0, _ENERGY)
_def("erg", 1.0e-7, _ENERGY)
_def("eV", 1.6021766208e-19, _ENERGY)
_def("keV", 1.6021766208e-16, _ENERGY)
_def("N", 1.0, (1, 1, -2, 0))
_def("dyne", 1.0e-5, (1, 1, -2, 0))
_def("dimensionless", 1.0, (0, 0, 0, 0))

_GALACTIC = ("Msun", "kpc", "Myr", "K")
_CGS = ("g", "cm", "s", "K")


class Unit:
    """factor * prod(symbol ** exponent); immutable."""

    __slots__ = ("factor", "powers")
    __array_priority__ = 1000

    def __init__(self, spec="", factor=1.0, powers=None):
        if powers is not None:
            self.factor = float(factor)
            self.powers = {k: Fraction(v) for k, v in powers.items() if v != 0 and k != "dimensionless"}
            return
        if isinstance(spec, Unit):
            self.factor, self.powers = spec.factor, dict(spec.powers)
            return
        u = _parse(spec)
        self.factor, self.powers = u.factor, u.powers

    # -- algebra ------------------------------------------------------------ #
    def __mul__(self, other):
        if isinstance(other, Unit):
            p = dict(self.powers)
            for k, v in other.powers.items():
                p[k] = p.get(k, 0) + v
            return Unit(factor=self.factor * other.factor, powers=p)
        if isinstance(other, (int, float)):
            return Unit(factor=self.factor * other, powers=self.powers)
        if isinstance(other, np.ndarray):
            return unyt_array(np.asarray(other), self)
        return NotImplemented

    def __rmul__(self, other):
        if isinstance(other, (int, float)):
            if self.factor == 1.0 and not isinstance(other, bool):
                # number * unit inside a unit string ("1e10*Msun") or 1.0 * kpc
                return _NumberTimesUnit(other, self)
            return Unit(factor=self.factor * other, powers=self.powers)
        if isinstance(other, np.ndarray):
            return unyt_array(np.asarray(other), self)
        return NotImplemented

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return self * other ** -1
        if isinstance(other, (int, float)):
            return Unit(factor=self.factor / other, powers=self.powers)
        return NotImplemented

    def __rtruediv__(self, other):
        if isinstance(other, (int, float)):
            return (self ** -1) * float(other)
        return NotImplemented

    def __pow__(self, e):
        e = Fraction(e).limit_denominator(1000)
        return Unit(factor=self.factor ** float(e), powers={k: v * e for k, v in self.powers.items()})

    def __eq__(self, other):
        if isinstance(other, str):
            other = Unit(other)
        return isinstance(other, Unit) and self.factor == other.factor and self.powers == other.powers

    def __ne__(self, other):
        return not self == other

    def __hash__(self):
        return hash((self.factor, tuple(sorted(self.powers.items()))))

    # -- values ------------------------------------------------------------- #
    @property
    def base_value(self):
        v = self.factor
        for k in sorted(self.powers):
            v = v * _SYMBOLS[k][0] ** float(self.powers[k])
        return v

    @property
    def dimensions(self):
        d = [Fraction(0)] * 4
        for k, e in self.powers.items():
            for i, x in enumerate(_SYMBOLS[k][1]):
                d[i] += x * e
        return tuple(d)

    @property
    def is_dimensionless(self):
        return all(x == 0 for x in self.dimensions)

    def get_conversion_factor(self, other):
        other = Unit(other)
        if self.dimensions != other.dimensions:
            raise UnitConversionError(f"cannot convert {self} to {other}")
        if self == other:
            return 1.0
        return self.base_value / other.base_value

    def get_base_equivalent(self, system="mks"):
        names = {"galactic": _GALACTIC, "cgs": _CGS, "mks": ("kg", "m", "s", "K")}[system]
        return Unit(powers={n: e for n, e in zip(names, self.dimensions)})

    def __str__(self):
        if not self.powers:
            return "dimensionless" if self.factor == 1.0 else repr(self.factor)
        parts = []
        for k in sorted(self.powers):
            e = self.powers[k]
            if e == 1:
                parts.append(k)
            elif e.denominator == 1:
                parts.append(f"{k}**{int(e)}" if e > 0 else f"{k}**({int(e)})")
            else:
                parts.append(f"{k}**({float(e)!r})")
        s = "*".join(parts)
        return s if self.factor == 1.0 else f"{self.factor!r}*{s}"

    __repr__ = __str__


class _NumberTimesUnit(Unit):
    """`1e10*Msun` inside a unit string is a Unit; `1.0*kpc` in code is a quantity.
    unyt returns a unyt_quantity for number*Unit; when such an object is used as a
    unit specification its value becomes the unit's factor."""

    __slots__ = ("number",)

    def __init__(self, number, unit):
        Unit.__init__(self, factor=unit.factor * number, powers=unit.powers)
        self.number = number

    def _as_quantity(self):
        return unyt_quantity(self.number, Unit(factor=self.factor / self.number, powers=self.powers))

    def to_value(self, units):
        return self._as_quantity().to_value(units)

    def to(self, units, *a, **k):
        return self._as_quantity().to(units, *a, **k)

    in_units = to


class UnitConversionError(Exception):
    pass


_NAMESPACE = None


def _parse(spec):
    global _NAMESPACE
    if spec is None or (isinstance(spec, str) and spec.strip() in ("", "dimensionless", "1")):
        return Unit(powers={})
    if not isinstance(spec, str):
        raise TypeError(f"bad unit specification {spec!r}")
    if _NAMESPACE is None:
        _NAMESPACE = {k: Unit(powers={k: 1}) for k in _SYMBOLS}
        _NAMESPACE["sqrt"] = lambda u: u ** Fraction(1, 2)
        _NAMESPACE["__builtins__"] = {}
    u = eval(spec, _NAMESPACE)  # unit strings come from the reference's source files
    if isinstance(u, (int, float)):
        return Unit(factor=float(u), powers={})
    return Unit(factor=u.factor, powers=u.powers)


def _units_of(x):
    return x.units if isinstance(x, unyt_array) else None


_DIMLESS = Unit(powers={})


class unyt_array(np.ndarray):
    __array_priority__ = 2.0

    def __new__(cls, input_array, units=None, registry=None, dtype=None, bypass_validation=False, name=None):
        if isinstance(input_array, unyt_array):
            ret = input_array.view(cls)
            ret.units = input_array.units if units is None else Unit(units)
            return ret
        if isinstance(input_array, (list, tuple)) and len(input_array) and any(
            isinstance(x, unyt_array) for x in input_array
        ):
            first = next(x for x in input_array if isinstance(x, unyt_array))
            vals = [x.to(first.units).d if isinstance(x, unyt_array) else x for x in input_array]
            ret = np.asarray(vals, dtype=dtype).view(cls)
            ret.units = first.units if units is None else Unit(units)
            return ret
        arr = np.asarray(input_array, dtype=dtype)
        if arr.dtype.kind in "iub" and dtype is None:
            arr = arr.astype("float64")
        ret = arr.view(cls)
        ret.units = _DIMLESS if units is None else Unit(units)
        return ret

    def __array_finalize__(self, obj):
        self.units = getattr(obj, "units", _DIMLESS)

    # -- accessors ----------------------------------------------------------- #
    @property
    def d(self):
        return self.view(np.ndarray)

    ndview = ndarray_view = d

    @property
    def v(self):
        return np.array(self.view(np.ndarray))

    value = v

    def to_ndarray(self):
        return self.v

    @property
    def uq(self):
        return unyt_quantity(1.0, self.units)

    unit_quantity = uq

    def __getitem__(self, item):
        ret = super().__getitem__(item)
        if not isinstance(ret, np.ndarray):
            ret = unyt_quantity(ret, self.units)
        elif ret.shape == () and type(ret) is not unyt_quantity:
            ret = unyt_quantity(ret.view(np.ndarray), self.units)
        return ret

    def __reduce__(self):
        return (unyt_array, (self.v, str(self.units)))

    def __repr__(self):
        return f"unyt_array({self.d!r}, '{self.units}')"

    def __str__(self):
        return f"{self.d} {self.units}"

    def __format__(self, spec):
        return format(self.d if self.ndim else float(self.d), spec) + (" " + str(self.units) if not spec else "")

    def __deepcopy__(self, memo):
        return type(self)(self.v, self.units)

    def copy(self, order="C"):
        return type(self)(np.copy(self.d, order=order), self.units)

    # -- conversions --------------------------------------------------------- #
    def _equivalent(self, units, equivalence, **kw):
        new = Unit(units)
        if equivalence == "number_density":
            # unyt NumberDensityEquivalence: n = rho / (mu * m_H)
            from .physical_constants import mh

            mu = kw.get("mu", 0.6)
            rho = self.to("g/cm**3")
            n = unyt_array(rho.d / (mu * mh.to("g").d), "cm**-3")
            return n.to(new)
        if equivalence == "thermal":
            from .physical_constants import kboltz

            if new.dimensions == Unit("K").dimensions:
                return unyt_array(self.to("J").d / kboltz.to("J/K").d, "K").to(new)
            return unyt_array(self.to("K").d * kboltz.to("J/K").d, "J").to(new)
        raise NotImplementedError(f"equivalence {equivalence!r} is not in the test shim")

    def to(self, units, equivalence=None, **kw):
        if equivalence is not None:
            return self._equivalent(units, equivalence, **kw)
        new = Unit(units)
        f = self.units.get_conversion_factor(new)
        return type(self)(self.d * f, new)

    in_units = to

    def to_equivalent(self, units, equivalence, **kw):
        return self._equivalent(units, equivalence, **kw)

    def to_value(self, units=None, equivalence=None, **kw):
        if units is None:
            return self.v
        return self.to(units, equivalence, **kw).v

    def convert_to_units(self, units, equivalence=None, **kw):
        if equivalence is not None:
            out = self._equivalent(units, equivalence, **kw)
            self.d[...] = out.d
            self.units = out.units
            return
        new = Unit(units)
        f = self.units.get_conversion_factor(new)
        np.multiply(self.d, f, out=self.d)
        self.units = new

    def in_base(self, unit_system="mks"):
        return self.to(self.units.get_base_equivalent(unit_system))

    def convert_to_base(self, unit_system="mks"):
        self.convert_to_units(self.units.get_base_equivalent(unit_system))

    def in_cgs(self):
        return self.in_base("cgs")

    def in_mks(self):
        return self.in_base("mks")

    # -- I/O through the (in-memory) h5py stand-in -------------------------------- #
    def write_hdf5(self, filename, dataset_name=None, info=None, group_name=None):
        import h5py

        f = h5py.File(filename, "a")
        g = f
        if group_name is not None:
            g = f[group_name] if group_name in f else f.create_group(group_name)
        if dataset_name in g:
            del g[dataset_name]
        ds = g.create_dataset(dataset_name or "array_data", data=np.array(self.d))
        ds.attrs["units"] = str(self.units)
        f.close()

    @classmethod
    def from_hdf5(cls, filename, dataset_name=None, group_name=None):
        import h5py

        f = h5py.File(filename, "r")
        g = f if group_name is None else f[group_name]
        ds = g[dataset_name or "array_data"]
        out = cls(np.array(ds[:] if ds.shape else ds[()]), ds.attrs.get("units", ""))
        f.close()
        return out

    # -- arithmetic ------------------------------------------------------------ #
    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        plain = [x.view(np.ndarray) if isinstance(x, unyt_array) else x for x in inputs]
        units = [_units_of(x) for x in inputs]
        name = ufunc.__name__
        out_units = None
        if method == "__call__":
            if name in ("multiply", "matmul"):
                out_units = (units[0] or _DIMLESS) * (units[1] or _DIMLESS)
            elif name in ("divide", "true_divide", "floor_divide"):
                out_units = (units[0] or _DIMLESS) / (units[1] or _DIMLESS)
            elif name in ("add", "subtract", "maximum", "minimum", "fmax", "fmin", "hypot",
                          "less", "less_equal", "greater", "greater_equal", "equal", "not_equal",
                          "remainder", "fmod", "arctan2", "copysign", "nextafter"):
                u0, u1 = units
                if u0 is not None and u1 is not None and u0 != u1:
                    plain[1] = plain[1] * u1.get_conversion_factor(u0)
                base = u0 if u0 is not None else u1
                if u0 is None or u1 is None:
                    # unyt lets a dimensionless / zero plain operand through
                    pass
                if name in ("add", "subtract", "maximum", "minimum", "fmax", "fmin", "hypot",
                            "remainder", "fmod", "copysign", "nextafter"):
                    out_units = base
            elif name == "power":
                if units[1] is not None:
                    raise UnitConversionError("exponent must be a plain number")
                out_units = (units[0] or _DIMLESS) ** float(np.asarray(plain[1]))
            elif name == "sqrt":
                out_units = units[0] ** Fraction(1, 2)
            elif name == "cbrt":
                out_units = units[0] ** Fraction(1, 3)
            elif name == "square":
                out_units = units[0] ** 2
            elif name == "reciprocal":
                out_units = units[0] ** -1
            elif name in ("negative", "positive", "absolute", "fabs", "rint", "floor", "ceil",
                          "trunc", "conjugate"):
                out_units = units[0]
            # everything else (exp, log, sin, isfinite, sign, ...) returns plain numbers
        elif method in ("reduce", "accumulate", "reduceat"):
            if name in ("add", "maximum", "minimum", "fmax", "fmin"):
                out_units = units[0]
        if out is not None:
            kwargs["out"] = tuple(o.view(np.ndarray) if isinstance(o, unyt_array) else o for o in out)
        res = getattr(ufunc, method)(*plain, **kwargs)
        if out is not None:
            o = out[0]
            if isinstance(o, unyt_array) and out_units is not None:
                o.units = out_units
            return o
        if out_units is None or res is NotImplemented:
            return res
        if isinstance(res, np.ndarray) and res.ndim > 0:
            r = res.view(unyt_array)
            r.units = out_units
            return r
        return unyt_quantity(res, out_units)

    # reductions that numpy implements as methods
    def dot(self, other):
        res = np.dot(self.d, np.asarray(other))
        return unyt_array(res, self.units * (_units_of(other) or _DIMLESS))

    def __float__(self):
        return float(self.d)


class unyt_quantity(unyt_array):
    def __new__(cls, input_scalar, units=None, registry=None, dtype=None, bypass_validation=False, name=None):
        if isinstance(input_scalar, unyt_array):
            u = input_scalar.units if units is None else Unit(units)
            input_scalar = input_scalar.d
        else:
            u = _DIMLESS if units is None else Unit(units)
        arr = np.asarray(input_scalar, dtype=dtype)
        if arr.shape != ():
            raise RuntimeError("unyt_quantity values must be numeric")
        if arr.dtype.kind in "iub" and dtype is None:
            arr = arr.astype("float64")
        ret = arr.view(cls)
        ret.units = u
        return ret

    def __repr__(self):
        return f"unyt_quantity({float(self.d)!r}, '{self.units}')"

    def __reduce__(self):
        return (unyt_quantity, (float(self.d), str(self.units)))


def uconcatenate(arrs, axis=0):
    u = arrs[0].units
    if not all(a.units == u for a in arrs):
        raise RuntimeError("Your arrays must have identical units.")
    return unyt_array(np.concatenate([a.d for a in arrs], axis=axis), u)


def _install_unit_names():
    g = globals()
    for k in _SYMBOLS:
        if k != "dimensionless":
            g[k] = Unit(powers={k: 1})


_install_unit_names()

from . import physical_constants  # noqa: E402
