"""
TEST INFRASTRUCTURE ONLY -- CPU oracle for the particle-sampling hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and there only as the checker (or the
thing timed *beside* the CUDA path), never as the thing shipped.

Parity status
-------------
* D / C / R / E / T / P / M / Gs / K rows of SURVEY.md section 8(a) and the multi-halo rows
  of 8(f)1: **pinned to the reference's own code executed in the build container**.
  ``tests/_refshim`` imports ``/root/reference/cluster_generator/{utils,virial,model,
  particles,radial_profiles,ics}.py`` unmodified (stand-ins only for the absent
  unyt / h5py / more_itertools and for the absent Cython function), and
  ``tests/test_reference_cpu.py`` asserts BITWISE equality of every table, f(E),
  ``check_virial`` array and particle field the reference computes with what this
  oracle computes, for all three BASELINE models, ``sub_sample``, ``r_max``,
  ``compute_potential``, gas, tracers and 1/2/3-halo ``ClusterICs``.
  ``tests/golden/*.npz`` are written from that reference run
  (``scripts/make_golden.py``), so the GPU tests compare with the reference's numbers.
  Also pinned by the one numeric known-answer test the reference ships
  (``tests/test_radial_profiles.py:177-194``) and by the MT19937 stream KAT.
* V row (``opt/cython_utils.pyx::generate_velocities``): the source is ABSENT
  from the reference snapshot and the reference ships no golden vectors for it
  => **parity unpinned** for the private RNG of the speed draw (draw order, ``<`` vs
  ``<=`` at equality are assumptions); restated from the call-site contract
  (``virial.py:191-206``, ``docs/source/particles.rst:102-105``).  What the seam
  receives (psi, vesc, fv2esc, knots, coeffs, k, quiet) IS pinned: the tests capture
  the arguments the reference's ``generate_particles`` passes.
"""
