"""
TEST INFRASTRUCTURE ONLY -- CPU oracle for the particle-sampling hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and there only as the checker (or the
thing timed *beside* the CUDA path), never as the thing shipped.

Parity status
-------------
* D / C / R / E / T / P / Gs rows of SURVEY.md section 8(a): the oracle calls the
  same scipy / numpy entry points as the reference (``scipy.integrate.quad``,
  ``InterpolatedUnivariateSpline``, ``np.interp``, ``RandomState``) in the same
  order, so it *is* the reference arithmetic by construction.  Pinned by the
  one numeric known-answer test the reference ships
  (``tests/test_radial_profiles.py:177-194``) and by the MT19937 stream KAT.
* V row (``opt/cython_utils.pyx::generate_velocities``): the source is ABSENT
  from the reference snapshot and the reference ships no golden vectors for it
  => **parity unpinned** for speed magnitudes; restated from the call-site
  contract (``virial.py:191-206``, ``docs/source/particles.rst:102-105``).
"""
