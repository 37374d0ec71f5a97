"""
TESTS ONLY -- executes the reference's own Python modules, unmodified, in this container.

``load_reference()`` imports ``/root/reference/cluster_generator/{utils, particles,
virial, model, radial_profiles, ics}.py`` straight from the read-only reference tree
(no copy, no edit) behind

* stand-ins for the packages this image lacks (``stubs/``: a minimal unit-aware
  ``unyt``, an in-memory ``h5py``, ``more_itertools.always_iterable``, empty
  ``ruamel.yaml`` / ``yt``), and
* an injected ``cluster_generator.opt.cython_utils.generate_velocities`` -- the one
  function whose source is absent from the snapshot -- bound to
  ``oracle.cg_oracle.generate_velocities_{sequential,counter}``.

The reference's package ``__init__`` (which eagerly imports yt-dependent modules) is
bypassed by registering an empty package object whose ``__path__`` is the reference
directory, so ``import cluster_generator.virial`` loads the reference's file itself.

It does not exist on the GPU box (``/root/reference`` is absent there): callers must
check ``available()`` and skip.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("CGB_REFERENCE_ROOT", "/root/reference")
_PKG = os.path.join(REFERENCE_ROOT, "cluster_generator")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")
_STUB_NAMES = ("unyt", "h5py", "more_itertools", "ruamel", "yt")
_REF = None


def available() -> bool:
    return os.path.isfile(os.path.join(_PKG, "virial.py"))


class _Velocities:
    """Stand-in for the absent Cython function (call site virial.py:198-206).

    ``mode = "sequential"``: one private uniform stream, particle after particle
    (oracle mode A); ``"counter"``: (u1,u2) = Philox(seed; particle id, trial)
    (mode B, what the CUDA path implements).  Records the per-particle trial counts
    of the last call."""

    def __init__(self):
        self.mode, self.seed, self.first_id = "sequential", None, 0
        self.last_trials = None
        self.calls = 0

    def __call__(self, psi, vesc, fv2esc, knots, coeffs, k, quiet):
        from oracle import cg_oracle as O

        self.calls += 1
        if self.mode == "sequential":
            kw = {} if self.seed is None else {"seed": self.seed}
            v, self.last_trials = O.generate_velocities_sequential(psi, vesc, fv2esc, knots, coeffs, k, **kw)
        else:
            v, self.last_trials = O.generate_velocities_counter(
                psi, vesc, fv2esc, knots, coeffs, k, seed=self.seed or 0, first_id=self.first_id)
        return v


class Reference(types.SimpleNamespace):
    """Namespace of the reference's modules: ``.utils .particles .virial .model
    .radial_profiles .ics`` plus ``.velocities`` (the injected seam) and ``.h5py``
    (the in-memory file store)."""


def load_reference() -> Reference:
    global _REF
    if _REF is not None:
        return _REF
    if not available():
        raise RuntimeError(f"the reference tree is not present at {REFERENCE_ROOT}")
    saved = {n: sys.modules.pop(n) for n in list(sys.modules)
             if n.split(".")[0] in _STUB_NAMES}
    sys.path.insert(0, _STUBS)
    try:
        pkg = types.ModuleType("cluster_generator")
        pkg.__path__ = [_PKG]
        pkg.__file__ = os.path.join(_PKG, "__init__.py")  # not executed
        sys.modules["cluster_generator"] = pkg
        opt = types.ModuleType("cluster_generator.opt")
        opt.__path__ = []
        cy = types.ModuleType("cluster_generator.opt.cython_utils")
        vel = _Velocities()
        cy.generate_velocities = vel
        sys.modules["cluster_generator.opt"] = opt
        sys.modules["cluster_generator.opt.cython_utils"] = cy
        pkg.opt, opt.cython_utils = opt, cy
        mods = {}
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for name in ("utils", "particles", "virial", "model", "radial_profiles", "ics"):
                mods[name] = importlib.import_module(f"cluster_generator.{name}")
        mods["utils"].mylog.setLevel("ERROR")
        stubs = {n: sys.modules[n] for n in _STUB_NAMES if n in sys.modules}
        _REF = Reference(velocities=vel, h5py=stubs["h5py"], unyt=stubs["unyt"], **mods)
    finally:
        sys.path.remove(_STUBS)
    # The stand-ins stay importable under their package names only for the
    # reference's own lazy imports (``unyt_array.write_hdf5`` -> ``import h5py``).
    # No other test may see them as "the real h5py": mark them.
    for n in _STUB_NAMES:
        top = sys.modules.get(n)
        if top is not None:
            top.__cgb_test_stub__ = True
    for n, m in saved.items():
        sys.modules.setdefault(n, m)
    return _REF
